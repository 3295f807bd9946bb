#!/bin/bash
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tma" 2>&1 | tail -15
AKB_MARCH_TMA=1 timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_hydro.py tests/test_gpu_smr_mhd.py -m gpu -q -x 2>&1 | tail -4
echo "== tma 64 blocks"; AKB_MARCH_TMA=1 python profiles/bench_variants.py --only "64 blocks" --steps 10 2>&1 | grep -v "^$" | tail -1
