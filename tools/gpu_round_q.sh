#!/bin/bash
# usage: bash tools/gpu_round_q.sh <tag> <ngpus> -- gradient exchange: CTA-shape sweep alone (check tool) and beside the FK
# kernels (bench weak companion: ms_per_step with / without the exchange)
TAG=${1:-r2q}
N=${2:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
  tools/peer_exchange_check.py --time > gpurun_out/${TAG}_peer_exchange_n$N.json 2> gpurun_out/${TAG}_peer_exchange_n$N.err; echo check rc=$?
python - <<PY
import json
t=open('gpurun_out/${TAG}_peer_exchange_n$N.json').read(); d=json.loads(t[t.index('{'):])
print('cases', len(d['cases']), 'failures', d['failures'])
for k,v in d['timing_6p4MB'].items(): print(k, v)
PY
for V in "peer 16 512 -1" "peer 16 512 0" "peer 64 128 -1" "peer 64 128 0" "peer 32 256 -1" "nccl 16 512 0" "nccl 16 512 -1"; do
set -- $V
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29527 \
  bench.py --gpus $N --steps 20 --warmup 5 --exchange $1 --exchange-ctas $2 --exchange-threads $3 --exchange-priority $4 \
  --no-cpu-baseline --no-e2e --no-extras > gpurun_out/${TAG}_sweep.json 2> gpurun_out/${TAG}_sweep.err || { echo "FAILED $V"; grep -v "NCCL INFO" gpurun_out/${TAG}_sweep.err | tail -c 600; }
python - "$V" <<PY | tee -a gpurun_out/${TAG}_exchange_sweep_n$N.txt
import json,sys
d=json.load(open('gpurun_out/${TAG}_sweep.json')); w=d['weak']; a=d['grad_allreduce']
print('%-18s alone %.1f us (nccl %.1f) | weak step %.4f ms, without %.4f -> exposed %.1f us | strong %.4f ms, without %.4f -> exposed %.1f us' % (
  sys.argv[1], a['ms_alone']*1e3, a['ms_alone_nccl']*1e3, w['ms_per_step'], w['ms_per_step_without_allreduce'],
  (w['ms_per_step']-w['ms_per_step_without_allreduce'])*1e3, d['ms_per_step'], a['ms_per_step_without_allreduce'], a['ms_exposed_per_step']*1e3))
PY
done
