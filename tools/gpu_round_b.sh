#!/bin/bash
# GPU round for the side kernels: all GPU tests, the unmodified-caller path, side-kernel timings (two bank-gather
# unrolls), and ncu --set full captures of the generator-mode, video-critic, bank and wide-row kernels.
# usage: bash tools/gpu_round_b.sh <tag>
TAG=${1:-r2b}
U8=$(ls -d dh-aug*/build_u8)/libdhfk.so
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/${TAG}_pytest_gpu.log
cat gpurun_out/${TAG}_pytest_gpu.log
python tools/dropin_path_bench.py --profile > gpurun_out/${TAG}_dropin_path.json 2> gpurun_out/${TAG}_dropin_path.err; echo dropin rc=$?
python tools/small_batch_latency.py > gpurun_out/${TAG}_small_batch.txt 2>&1; cat gpurun_out/${TAG}_small_batch.txt
python tools/aux_bench.py > gpurun_out/${TAG}_aux_bench.json 2> gpurun_out/${TAG}_aux_bench.err; echo aux rc=$?
if [ -f "$U8" ]; then DHFK_LIB_PATH=$PWD/$U8 python tools/aux_bench.py > gpurun_out/${TAG}_aux_bench_u8.json 2>/dev/null; fi
for T in gen video bank wide; do
  python tools/ncu_target.py $T > gpurun_out/ncu_plain_$T.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:dhfk_ -s 4 -c 3 -f -o gpurun_out/${TAG}_prof_$T python tools/ncu_target.py $T > gpurun_out/ncu_$T.log 2>&1
  tail -2 gpurun_out/ncu_$T.log
done
ls -la gpurun_out | tail -12
