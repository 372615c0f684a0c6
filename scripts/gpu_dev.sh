#!/bin/bash
mkdir -p gpurun_out
for sl in 3 4; do
echo "== slots $sl"
AUCELL_B200_TIE_SLOTS=$sl timeout 300 python scripts/dev_tie.py 200000 2>&1 | grep -E '"weighted": true, "fast": 1|True kernel 1' | cut -c1-130
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tie_class or byte_coded or config4 or config5" 2>&1 | tail -3
AUCELL_B200_TIE_SLOTS=3 timeout 300 python scripts/sweep_cfg5.py 2>/dev/null | cut -c1-140 | tail -5
