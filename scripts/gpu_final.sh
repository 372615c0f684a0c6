#!/bin/bash
# round-end record: full bench (both arms), launch list of a short run, ncu --set full of the tie-class kernel
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench exit $?" >> gpurun_out/bench_full.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?" >> gpurun_out/bench_ref.err
SHORT="python bench.py --cells 50000 --steps 2 --warmup 1 --no-e2e --no-cpu --no-api"
$SHORT > gpurun_out/plain_short.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $SHORT > gpurun_out/ncu_launches.log 2>&1
CMD="python scripts/dev_tie_ncu.py 50000 weighted"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:aucell_tie_kernel -s 1 -c 1 -f -o gpurun_out/tie_v11 $CMD > gpurun_out/ncu.log 2>&1
echo "ncu exit $?" >> gpurun_out/ncu.log
cat gpurun_out/bench_full.json | cut -c1-2500; tail -2 gpurun_out/bench_full.err; cat gpurun_out/bench_ref.json | cut -c1-600; tail -3 gpurun_out/launches.csv | cut -c1-300; tail -2 gpurun_out/ncu.log
