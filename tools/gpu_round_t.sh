#!/bin/bash
# usage: bash tools/gpu_round_t.sh <tag> <ngpus> [steps] -- the hand-written gradient exchange at N GPUs: check against NCCL + timing,
# then the driver's configs[4] command with it
TAG=${1:-r2t}
N=${2:-8}
STEPS=${3:-20}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
  tools/peer_exchange_check.py --time > gpurun_out/${TAG}_peer_exchange_n$N.json 2> gpurun_out/${TAG}_peer_exchange_n$N.err; echo check rc=$?
python - <<PY
import json
t=open('gpurun_out/${TAG}_peer_exchange_n$N.json').read(); d=json.loads(t[t.index('{'):])
print('world', d['world'], 'multicast', d['multicast'], 'cases', len(d['cases']), 'failures', d['failures'])
for k,v in d['timing_6p4MB'].items(): print(k, v)
for k,v in d['timing_by_size'].items(): print(k, v)
PY
grep -v "NCCL INFO" gpurun_out/${TAG}_peer_exchange_n$N.err | tail -c 600
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29527 \
  bench.py --gpus $N --steps $STEPS --warmup 5 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo bench rc=$?
grep -v "NCCL INFO" gpurun_out/${TAG}_bench_n$N.err | tail -c 600
python - <<PY
import json; d=json.load(open('gpurun_out/${TAG}_bench_n$N.json'))
print('value %.4e ms/step %.4f n_gpus %d scaling %s' % (d['value'], d['ms_per_step'], d['n_gpus'], d['scaling']))
print('fwd %.3f bwd %.3f step %.3f' % (d['roofline_fwd']['frac'], d['roofline']['frac'], d['roofline_step']['frac']), d['clocks'])
print('weak', d.get('weak')); print('allreduce', d.get('grad_allreduce'))
e=d['e2e']; print('e2e', e['value'], e['ms_per_step'], e['copy_ceiling_ms'], e['frac_of_ceiling'], e['per_rank_gbs'])
PY
