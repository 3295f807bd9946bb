#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r02_s17_pytest.log 2>&1
tail -4 gpurun_out/r02_s17_pytest.log
python profiles/bench_variants.py --steps 6 2>&1 | grep -v "^$" | tail -16
