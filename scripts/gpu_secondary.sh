#!/bin/bash
mkdir -p gpurun_out
for s in profile_api sweep_cfg5 bench_shapes bench_continuous bench_rank; do
  AUCELL_B200_DEBUG=1 timeout 600 python scripts/$s.py > gpurun_out/$s.jsonl 2> gpurun_out/$s.err; echo "$s exit $?"
done
tail -n 20 gpurun_out/profile_api.jsonl gpurun_out/sweep_cfg5.jsonl gpurun_out/bench_shapes.jsonl gpurun_out/bench_continuous.jsonl gpurun_out/bench_rank.jsonl | cut -c1-700
tail -5 gpurun_out/profile_api.err
