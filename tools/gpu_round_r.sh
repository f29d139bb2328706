#!/bin/bash
# usage: bash tools/gpu_round_r.sh <tag> <ngpus> -- gradient exchange: latency vs bandwidth (size sweep), strong vs weak memory ops
TAG=${1:-r2r}
N=${2:-2}
PKG=dh-aug-dh-forward-kinematics-model-driven-augmentation-for-3d-human-pose-estimation_b200
for V in default weakio; do
[ $V = weakio ] && export DHFK_LIB_PATH=$PWD/$PKG/build_w/libdhfk.so
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
  tools/peer_exchange_check.py --time > gpurun_out/${TAG}_peer_exchange_n${N}_$V.json 2> gpurun_out/${TAG}_peer_exchange_n${N}_$V.err; echo check $V rc=$?
python - <<PY
import json
t=open('gpurun_out/${TAG}_peer_exchange_n${N}_$V.json').read(); d=json.loads(t[t.index('{'):])
print('$V cases', len(d['cases']), 'failures', d['failures'])
for k,v in d['timing_6p4MB'].items(): print(k, v)
for k,v in d['timing_by_size'].items(): print(k, v)
PY
done
