#!/bin/bash
# usage: bash tools/gpu_round_y.sh <tag> <ngpus> -- exchange beside the FK kernels: 8 vs 16 CTAs (8 loads in flight per thread)
TAG=${1:-r2y}
N=${2:-2}
for V in "peer 8 512 -1" "peer 16 512 -1" "peer 4 512 -1"; do
set -- $V
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29527 \
  bench.py --gpus $N --steps 20 --warmup 5 --exchange $1 --exchange-ctas $2 --exchange-threads $3 --exchange-priority $4 \
  --no-cpu-baseline --no-e2e --no-extras > gpurun_out/${TAG}_sweep.json 2> gpurun_out/${TAG}_sweep.err || { echo "FAILED $V"; grep -v "NCCL INFO" gpurun_out/${TAG}_sweep.err | tail -c 600; }
python - "$V" <<PY | tee -a gpurun_out/${TAG}_exchange_sweep_n$N.txt
import json,sys
d=json.load(open('gpurun_out/${TAG}_sweep.json')); w=d['weak']; a=d['grad_allreduce']
print('%-18s alone %.1f us (nccl %.1f) | weak step %.4f ms, without %.4f -> exposed %.1f us | strong %.4f ms, without %.4f -> exposed %.1f us' % (
  sys.argv[1], a['ms_alone']*1e3, a['ms_alone_nccl']*1e3, w['ms_per_step'], w['ms_per_step_without_allreduce'],
  w['ms_exposed_per_step']*1e3, d['ms_per_step'], a['ms_per_step_without_allreduce'], a['ms_exposed_per_step']*1e3))
PY
done
