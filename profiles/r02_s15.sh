#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02_s15_pytest.log 2>&1
tail -6 gpurun_out/r02_s15_pytest.log
