#!/bin/bash
# N GPUs: PCIe ceiling, multi-GPU tests, the strong-scaling bench
N=${N:-8}
mkdir -p gpurun_out
if [ "$DEVTIE" == "1" ]; then timeout 300 python scripts/dev_tie.py 200000 2>&1 | grep -E '"fast": 1|kernel 1' | cut -c1-120; fi
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/pcie_probe_multi.py > gpurun_out/pcie_probe_${N}gpu.json 2> gpurun_out/pcie_probe_${N}gpu.err; cat gpurun_out/pcie_probe_${N}gpu.json
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q > gpurun_out/pytest_multigpu_${N}.log 2>&1; tail -3 gpurun_out/pytest_multigpu_${N}.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench exit $?"; grep "^{" gpurun_out/bench_${N}gpu.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(json.dumps({k: d[k] for k in ('value', 'e2e', 'sharded')})[:3000])"; grep -v "Warning\|warn\|OMP_NUM\|\*\*\*" gpurun_out/bench_${N}gpu.err | tail -4 | cut -c1-300
