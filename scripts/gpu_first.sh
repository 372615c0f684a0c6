#!/bin/bash
# first GPU pass: parity tests, smoke, a short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
free -g >> gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --cells 200000 --steps 5 --warmup 3 > gpurun_out/bench_200k.json 2> gpurun_out/bench_200k.err; echo "bench exit $?" >> gpurun_out/bench_200k.err
tail -5 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; cat gpurun_out/bench_200k.json; tail -5 gpurun_out/bench_200k.err
