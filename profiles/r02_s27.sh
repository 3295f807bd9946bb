#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_hydro.py -m gpu -q -x 2>&1 | tail -3
AKB_PIPE_TRACE=1 python profiles/e2e_pipe.py --slots 4 --jobs 12 2>&1 | tail -6
python profiles/e2e_pipe.py --slots 3 --jobs 12 2>&1 | tail -1
python profiles/e2e_pipe.py --slots 5 --jobs 20 2>&1 | tail -1
