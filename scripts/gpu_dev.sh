#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python scripts/bench_shapes.py > gpurun_out/bench_shapes.jsonl 2> gpurun_out/bench_shapes.err; cut -c1-170 gpurun_out/bench_shapes.jsonl
timeout 400 python scripts/fuzz_gpu.py ${FUZZ:-240} 7 > gpurun_out/fuzz.log 2>&1; tail -3 gpurun_out/fuzz.log
AUCELL_B200_TIE_KERNEL=2 timeout 400 python scripts/fuzz_gpu.py 120 8 > gpurun_out/fuzz2.log 2>&1; tail -3 gpurun_out/fuzz2.log
