#!/bin/bash
TAG=${1:-r2f}
set -x
python -m pytest tests/test_video_critic.py -m gpu -x -q --tb=short 2>&1 | tail -40 > gpurun_out/${TAG}_video_tests.log; cat gpurun_out/${TAG}_video_tests.log
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/${TAG}_pytest_gpu.log; cat gpurun_out/${TAG}_pytest_gpu.log
python tools/aux_bench.py > gpurun_out/${TAG}_aux_bench.json 2> gpurun_out/${TAG}_aux_bench.err; echo aux rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/%s_aux_bench.json' % __import__('sys').argv[1] if False else 'gpurun_out/TAG_aux_bench.json'.replace('TAG', __import__('os').environ.get('TAGX','r2f'))))
for k,v in d['kernels'].items():
    if 'video' in k or 'bank' in k: print("%-70s %.4f ms %.3f" % (k[:70], v['ms'], v['frac']))
PY
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo bench rc=$?
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline_fwd']['frac'], d['clocks'])
print('sustained', d.get('sustained')); print('generator', d.get('generator_mode'))
print('dropin', {k: v for k, v in d.get('dropin_path', {}).items() if k.startswith('at_')})
"
