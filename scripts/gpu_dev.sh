#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -4 gpurun_out/pytest_gpu.log
AUCELL_B200_DEBUG=1 timeout 600 python scripts/profile_api.py > gpurun_out/profile_api.jsonl 2> gpurun_out/profile_api.err; cut -c1-700 gpurun_out/profile_api.jsonl; grep "aucell_b200\] csr host call: 300000" gpurun_out/profile_api.err | cut -c1-420; tail -3 gpurun_out/profile_api.err | cut -c1-300
