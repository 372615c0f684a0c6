#!/bin/bash
mkdir -p gpurun_out
for v in 1 2; do
  export AUCELL_B200_TIE_KERNEL=$v
  timeout 300 python scripts/dev_tie.py 200000 2>&1 | grep -E '"fast": '$v'|kernel '$v > gpurun_out/ab_dev_$v.log
  timeout 600 python scripts/sweep_cfg5.py > gpurun_out/ab_sweep_$v.jsonl 2> gpurun_out/ab_sweep_$v.err
  timeout 600 python scripts/bench_shapes.py > gpurun_out/ab_shapes_$v.jsonl 2> gpurun_out/ab_shapes_$v.err
done
for v in 1 2; do echo "== kernel $v"; cut -c1-200 gpurun_out/ab_dev_$v.log; cut -c1-140 gpurun_out/ab_sweep_$v.jsonl; cut -c1-160 gpurun_out/ab_shapes_$v.jsonl; done
