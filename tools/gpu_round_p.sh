#!/bin/bash
# usage: bash tools/gpu_round_p.sh <tag> <ngpus> -- the hand-written gradient exchange: check against NCCL + timing, then
# configs[4] with it and with NCCL; last, the CUDA-graph capture probe on one GPU
TAG=${1:-r2p}
N=${2:-2}
set -x
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
  tools/peer_exchange_check.py --time > gpurun_out/${TAG}_peer_exchange_n$N.json 2> gpurun_out/${TAG}_peer_exchange_n$N.err; echo check rc=$?
cat gpurun_out/${TAG}_peer_exchange_n$N.json | head -150
grep -v "NCCL INFO" gpurun_out/${TAG}_peer_exchange_n$N.err | tail -c 1500
for X in peer nccl; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29527 \
  bench.py --gpus $N --steps 20 --warmup 5 --exchange $X --no-cpu-baseline > gpurun_out/${TAG}_bench_n${N}_$X.json 2> gpurun_out/${TAG}_bench_n${N}_$X.err; echo bench $X rc=$?
grep -v "NCCL INFO" gpurun_out/${TAG}_bench_n${N}_$X.err | tail -c 800
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_n${N}_$X.json'))
print('value %.4e ms/step %.4f n_gpus %d scaling %s' % (d['value'], d['ms_per_step'], d['n_gpus'], d['scaling']))
print('clocks', d['clocks'])
print('weak', d.get('weak')); print('allreduce', d.get('grad_allreduce'))
"
done
timeout 300 python tools/graph_capture_probe.py > gpurun_out/${TAG}_graph_probe.txt 2>&1; echo probe rc=$?
tail -60 gpurun_out/${TAG}_graph_probe.txt
