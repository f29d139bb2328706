#!/bin/bash
# usage: bash tools/gpu_round_k.sh <tag> <ngpus> -- configs[4]: sweep of the all-reduce communicator's CTA limit, then of the e2e chunking
TAG=${1:-r2k}
N=${2:-4}
set -x
: > gpurun_out/${TAG}_nccl_sweep.txt
for C in 0 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + C)) \
    bench.py --gpus $N --steps 100 --warmup 5 --no-e2e --no-cpu-baseline --nccl-max-ctas $C 2>/dev/null | python -c "
import sys, json; d = json.loads(sys.stdin.read()); a = d['grad_allreduce']; w = d['weak']
print('max_ctas %d  strong %.4e poses/s %.4f ms/step (no allreduce %.4f, exposed %+.4f)  fwd %.3f bwd %.3f | weak %.4e (%.4f vs %.4f ms) | alone %.4f ms  [%s]' % (
  $C, d['value'], d['ms_per_step'], a['ms_per_step_without_allreduce'], a['ms_exposed_per_step'], d['roofline_fwd']['frac'], d['roofline']['frac'],
  w['value'], w['ms_per_step'], w['ms_per_step_without_allreduce'], a['ms_alone'], a['communicator']))" >> gpurun_out/${TAG}_nccl_sweep.txt
done
cat gpurun_out/${TAG}_nccl_sweep.txt
: > gpurun_out/${TAG}_e2e_sweep.txt
for spec in "131072 3" "65536 4" "32768 6" "262144 3"; do
  set -- $spec
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + $2)) \
    bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --e2e-chunk $1 --e2e-slots $2 2>/dev/null | python -c "
import sys, json; d = json.loads(sys.stdin.read()); e = d['e2e']
print('chunk %7d slots %d  e2e %.4e poses/s %.3f ms (ceiling %.3f ms, frac %.3f)  fwd-only %.3f ms  per-rank %s' % (
  $1, $2, e['value'], e['ms_per_step'], e['copy_ceiling_ms'], e['frac_of_ceiling'], e['forward_only']['ms_per_step'], {k: round(v, 1) for k, v in e['per_rank_gbs'].items()}))" >> gpurun_out/${TAG}_e2e_sweep.txt
done
cat gpurun_out/${TAG}_e2e_sweep.txt
