#!/bin/bash
# One full GPU round (run under gpurun): GPU tests, bench line, smoke, ncu launch list, ncu --set full of one forward +
# one backward launch.  usage: bash tools/gpu_round.sh [tag]     artefacts: gpurun_out/<tag>_*
# then, back in the build container: python tools/summarize_profile.py <tag>
TAG=${1:-r1}
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${TAG}_pytest_gpu.log
cat gpurun_out/${TAG}_pytest_gpu.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo bench rc=$?
cat gpurun_out/${TAG}_bench.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python tools/aux_bench.py > gpurun_out/${TAG}_aux_bench.json 2> gpurun_out/${TAG}_aux_bench.err; echo aux rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dhfk_ -s 6 -c 2 -f -o gpurun_out/${TAG}_prof python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out | tail -5
