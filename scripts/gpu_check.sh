#!/bin/bash
# parity tests + full bench (our arm) on one GPU
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
AUCELL_B200_DEBUG=1 timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench exit $?" >> gpurun_out/bench_full.err
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_full.json; tail -8 gpurun_out/bench_full.err
