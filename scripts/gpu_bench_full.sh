#!/bin/bash
# full-size bench (both arms) + ncu launch list of a short run
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench exit $?" >> gpurun_out/bench_full.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?" >> gpurun_out/bench_ref.err
SHORT="python bench.py --cells 50000 --steps 2 --warmup 1 --no-e2e --no-cpu"
$SHORT > gpurun_out/plain_short.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $SHORT > gpurun_out/ncu_launches.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_full.json; tail -3 gpurun_out/bench_full.err; cat gpurun_out/bench_ref.json; tail -4 gpurun_out/bench_ref.err; tail -5 gpurun_out/launches.csv
