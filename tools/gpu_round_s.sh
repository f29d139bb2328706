#!/bin/bash
# usage: bash tools/gpu_round_s.sh <tag> -- one GPU: GPU tests, smoke, the driver's bench command
TAG=${1:-r2s}
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${TAG}_pytest_gpu.log; cat gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo bench rc=$?
python - <<PY
import json; d=json.load(open("gpurun_out/${TAG}_bench.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline_fwd"]["frac"], d["clocks"])
print(json.dumps(d["dropin_path"]["at_1024"])); print(json.dumps(d["dropin_path"]["at_4608"]))
PY
