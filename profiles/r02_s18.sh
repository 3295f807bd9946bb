#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_zamr.py -m gpu -q -x 2>&1 | tail -12
python profiles/bench_variants.py --steps 6 2>&1 | grep -v "^$" | tail -16
