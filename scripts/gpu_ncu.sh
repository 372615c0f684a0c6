#!/bin/bash
# ncu capture of the fused kernel on a short bench run (plain run first, as the profiling recipe requires)
mkdir -p gpurun_out
CMD="python bench.py --workload ${WORKLOAD:-cfg4} --cells ${CELLS:-50000} --steps 1 --warmup 1 --no-e2e --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:aucell_fused_kernel -s 1 -c 1 -f -o gpurun_out/prof $CMD > gpurun_out/ncu.log 2>&1
echo "ncu exit $?" >> gpurun_out/ncu.log
tail -3 gpurun_out/plain.log; tail -5 gpurun_out/ncu.log
ls -la gpurun_out/
