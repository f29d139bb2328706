#!/bin/bash
# usage: bash tools/gpu_round_d.sh <tag> <ngpus>  -- multi-GPU bench (configs[4]) native + reference arm, and an A/B of the
# wide-row switch on GPU 0
TAG=${1:-r2d}
N=${2:-2}
set -x
NW=$(ls -d dh-aug*/build_nw 2>/dev/null)/libdhfk.so
if [ -f "$NW" ]; then
  python tools/ab_bench.py default= nowide=$NW default= nowide=$NW > gpurun_out/${TAG}_ab_wide.txt 2>&1; cat gpurun_out/${TAG}_ab_wide.txt
fi
NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo bench rc=$?
grep -E "NCCL INFO (comm|Channel 00/|Connected|NVLS|ncclCommInitRank)" gpurun_out/${TAG}_bench_n$N.err | head -12
tail -c 1500 gpurun_out/${TAG}_bench_n$N.err
cat gpurun_out/${TAG}_bench_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
  bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref_n$N.json 2>/dev/null; echo ref rc=$?
cat gpurun_out/${TAG}_bench_ref_n$N.json
python -m pytest tests/test_parity_gpu.py -m gpu -q -k "nan_and_zero or strided" 2>&1 | tail -3
