#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > gpurun_out/r02_s21_pytest.log
cat gpurun_out/r02_s21_pytest.log
python bench.py > gpurun_out/r02_s21_bench.json 2> gpurun_out/r02_s21_bench.err
cat gpurun_out/r02_s21_bench.json
