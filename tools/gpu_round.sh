set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r1k_pytest_gpu.log
cat gpurun_out/r1k_pytest_gpu.log
python bench.py > gpurun_out/r1k_bench.json 2> gpurun_out/r1k_bench.err; echo bench rc=$?
cat gpurun_out/r1k_bench.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1k_launches.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dhfk_ -s 6 -c 2 -f -o gpurun_out/r1k_prof python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out | tail -5
