#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_gpu_smr.py tests/test_gpu_smr_mhd.py tests/test_gpu_zamr.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02_s16_pytest.log 2>&1
tail -4 gpurun_out/r02_s16_pytest.log
python profiles/bench_variants.py --only "SMR" --steps 6 2>&1 | tail -4
python profiles/bench_variants.py --only "config 5" --steps 4 2>&1 | tail -3
