#!/bin/bash
# usage: bash tools/gpu_round_u.sh <tag> <ngpus> -- exchange kernel: loads in flight per thread (DHFK_AR_UNROLL 4 / 8 / 16 builds);
# then the CUDA-graph capture of the caller's path through the bench extra's own code
TAG=${1:-r2u}
N=${2:-2}
PKG=dh-aug-dh-forward-kinematics-model-driven-augmentation-for-3d-human-pose-estimation_b200
for U in 4 8 16; do
[ $U = 4 ] && unset DHFK_LIB_PATH || export DHFK_LIB_PATH=$PWD/$PKG/build_u$U/libdhfk.so
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
  tools/peer_exchange_check.py --time > gpurun_out/${TAG}_u$U.json 2> gpurun_out/${TAG}_u$U.err; echo check unroll $U rc=$?
python - <<PY | tee -a gpurun_out/${TAG}_unroll_n$N.txt
import json
t=open('gpurun_out/${TAG}_u$U.json').read(); d=json.loads(t[t.index('{'):])
print('unroll $U: cases', len(d['cases']), 'failures', d['failures'])
for k,v in d['timing_6p4MB'].items():
    if k.startswith('multimem') or k=='nccl': print('  unroll $U', k, v)
PY
done
unset DHFK_LIB_PATH
timeout 300 python tools/graph_capture_probe.py --measure > gpurun_out/${TAG}_graph_measure.txt 2>&1; echo probe rc=$?
tail -25 gpurun_out/${TAG}_graph_measure.txt
