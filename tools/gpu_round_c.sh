#!/bin/bash
# usage: bash tools/gpu_round_c.sh <tag>   -- all GPU tests, the N=1 bench line, smoke, the unmodified-caller path, side kernels
TAG=${1:-r2c}
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/${TAG}_pytest_gpu.log
cat gpurun_out/${TAG}_pytest_gpu.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo bench rc=$?
tail -c 3000 gpurun_out/${TAG}_bench.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python tools/dropin_path_bench.py > gpurun_out/${TAG}_dropin_path.json 2> gpurun_out/${TAG}_dropin_path.err; echo dropin rc=$?
python tools/aux_bench.py > gpurun_out/${TAG}_aux_bench.json 2> gpurun_out/${TAG}_aux_bench.err; echo aux rc=$?
