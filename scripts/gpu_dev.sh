#!/bin/bash
mkdir -p gpurun_out
for sk in 0 8 16 3; do
echo "== skip $sk"
AUCELL_B200_TIE_SKIP=$sk AUCELL_B200_TIE_KERNEL=1 timeout 300 python scripts/dev_tie.py 200000 2>&1 | grep -E '"weighted": true, "fast": 1' | cut -c1-100
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python scripts/fuzz_gpu.py 150 31 > gpurun_out/fuzz3.log 2>&1; tail -2 gpurun_out/fuzz3.log
