#!/bin/bash
# usage: bash tools/gpu_round_e.sh <tag> -- GPU tests; burst-vs-sustained sweep of the headline step (launch size x run length);
# ncu of the video-critic and bank kernels
TAG=${1:-r2e}
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/${TAG}_pytest_gpu.log; cat gpurun_out/${TAG}_pytest_gpu.log
: > gpurun_out/${TAG}_sweep.txt
for spec in "1048576 200" "1048576 2000" "1048576 8000" "2097152 100" "2097152 1000" "8388608 25" "8388608 250" "16777216 12" "16777216 100"; do
  set -- $spec
  python bench.py --poses $1 --steps $2 --warmup 5 --no-e2e --no-cpu-baseline --no-extras --buffers $([ $1 -ge 8388608 ] && echo 1 || echo 4) 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('poses %9d steps %5d  %.4e poses/s  %.4f ms/step  fwd %.3f bwd %.3f  clk %s %s' % ($1, $2, d['value'], d['ms_per_step'], d['roofline_fwd']['frac'], d['roofline']['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons']))" >> gpurun_out/${TAG}_sweep.txt
done
cat gpurun_out/${TAG}_sweep.txt
python tools/ncu_target.py video > gpurun_out/ncu_plain_video.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:dhfk_ -s 3 -c 3 -f -o gpurun_out/${TAG}_prof_video python tools/ncu_target.py video > gpurun_out/ncu_video.log 2>&1
python tools/ncu_target.py bank > gpurun_out/ncu_plain_bank.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:dhfk_ -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_bank python tools/ncu_target.py bank > gpurun_out/ncu_bank.log 2>&1
ls -la gpurun_out | tail -5
