#!/bin/bash
# usage: bash tools/gpu_round_j.sh <tag>  -- GPU tests + smoke with the packed-V3 build, A/B against the scalar build, side kernels
TAG=${1:-r2j}
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/${TAG}_pytest_gpu.log; cat gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
NP=$(ls -d dh-aug*/build_np 2>/dev/null)/libdhfk.so
python tools/ab_bench.py packed= scalar=$NP packed= scalar=$NP > gpurun_out/${TAG}_ab_packed_v3.txt 2>&1; cat gpurun_out/${TAG}_ab_packed_v3.txt
python tools/aux_bench.py > gpurun_out/${TAG}_aux_bench.json 2> gpurun_out/${TAG}_aux_bench.err; echo aux rc=$?
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_aux_bench.json'))
for k,v in d['kernels'].items():
    if 'video' in k or 'bank' in k or 'FK' in k: print('%-72s %.4f ms %.3f' % (k[:72], v['ms'], v['frac']))
"
