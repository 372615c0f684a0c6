#!/bin/bash
timeout 300 python scripts/dev_tie.py 200000 2>&1 | grep -E '"fast": 1|kernel 1' | cut -c1-120
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tie_class or byte_coded or config4 or config5" 2>&1 | tail -2
