#!/bin/bash
# usage: bash tools/gpu_round_g.sh <tag> <ngpus> [steps] -- configs[4] at N GPUs (native + reference arm), short
TAG=${1:-r2g}
N=${2:-8}
STEPS=${3:-100}
set -x
NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29527 \
  bench.py --gpus $N --steps $STEPS --warmup 5 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo bench rc=$?
grep -E "NCCL INFO" gpurun_out/${TAG}_bench_n$N.err | grep -iE "nranks|NVLS|Init COMPLETE" | head -6
grep -v "NCCL INFO" gpurun_out/${TAG}_bench_n$N.err | tail -c 1200
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_n$N.json'))
print('value %.4e ms/step %.4f n_gpus %d scaling %s' % (d['value'], d['ms_per_step'], d['n_gpus'], d['scaling']))
print('fwd %.3f bwd %.3f step %.3f' % (d['roofline_fwd']['frac'], d['roofline']['frac'], d['roofline_step']['frac']), d['clocks'])
print('weak', d.get('weak')); print('allreduce', d.get('grad_allreduce'))
e=d['e2e']; print('e2e', e['value'], e['ms_per_step'], e['copy_ceiling_ms'], e['frac_of_ceiling'], e['per_rank_gbs'])
"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29528 \
  bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref_n$N.json 2>/dev/null; echo ref rc=$?
cat gpurun_out/${TAG}_bench_ref_n$N.json | cut -c1-300
