#!/bin/bash
mkdir -p gpurun_out
for v in 1 0; do
export AUCELL_B200_DENSE_CSR=$v
echo "== DENSE_CSR=$v"
timeout 600 python bench.py --workload rank --steps 5 --warmup 3 2>/dev/null | cut -c1-200
timeout 600 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | cut -c1-200
timeout 600 python bench.py --workload cfg2 --cells 100000 --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | cut -c1-200
done
