#!/bin/bash
# round-end record: GPU suite, full bench (both arms), secondary workloads and sweeps, launch list, ncu of the tie-class kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench exit $?" >> gpurun_out/bench_full.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?" >> gpurun_out/bench_ref.err
for wl in cfg3 cfg2 rank; do timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; done
for s in sweep_cfg5 bench_shapes bench_continuous bench_rank; do timeout 600 python scripts/$s.py > gpurun_out/$s.jsonl 2> gpurun_out/$s.err; done
SHORT="python bench.py --cells 50000 --steps 2 --warmup 1 --no-e2e --no-cpu --no-api"
$SHORT > gpurun_out/plain_short.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $SHORT > gpurun_out/ncu_launches.log 2>&1
CMD="python scripts/dev_tie_ncu.py 50000 weighted"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:aucell_tie_kernel -s 1 -c 1 -f -o gpurun_out/tie_v12 $CMD > gpurun_out/ncu.log 2>&1
echo "ncu exit $?" >> gpurun_out/ncu.log
tail -2 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cut -c1-400 gpurun_out/bench_full.json; tail -1 gpurun_out/bench_full.err; for wl in cfg3 cfg2 rank; do cut -c1-220 gpurun_out/bench_$wl.json; done; cut -c1-150 gpurun_out/sweep_cfg5.jsonl; cut -c1-150 gpurun_out/bench_shapes.jsonl; cut -c1-120 gpurun_out/bench_continuous.jsonl; cat gpurun_out/bench_rank.jsonl; tail -1 gpurun_out/ncu.log
