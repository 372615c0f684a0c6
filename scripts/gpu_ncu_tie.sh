#!/bin/bash
# ncu capture of the tie-class kernel (plain run first): NAME=<report name> KREGEX=<kernel regex>
mkdir -p gpurun_out
CMD="python scripts/dev_tie_ncu.py 50000 ${MODE:-weighted}"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${KREGEX:-aucell_tie2_kernel} -s 1 -c 1 -f -o gpurun_out/${NAME:-tie2} $CMD > gpurun_out/ncu.log 2>&1
echo "ncu exit $?" >> gpurun_out/ncu.log
tail -2 gpurun_out/plain.log; tail -3 gpurun_out/ncu.log
