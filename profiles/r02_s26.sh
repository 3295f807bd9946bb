#!/bin/bash
CUDA_DEVICE_MAX_CONNECTIONS=32 AKB_PIPE_TRACE=1 python profiles/e2e_pipe.py --slots 4 --jobs 12 2>&1 | tail -6
CUDA_DEVICE_MAX_CONNECTIONS=32 python profiles/e2e_pipe.py --slots 3 --jobs 12 2>&1 | tail -1
CUDA_DEVICE_MAX_CONNECTIONS=32 python profiles/e2e_pipe.py --slots 5 --jobs 20 2>&1 | tail -1
CUDA_DEVICE_MAX_CONNECTIONS=32 AKB_SWEEP_STREAMS=0 python profiles/e2e_pipe.py --slots 4 --jobs 12 2>&1 | tail -1
AKB_SWEEP_STREAMS=0 AKB_OVERLAP_U=0 python profiles/e2e_pipe.py --slots 4 --jobs 12 2>&1 | tail -1
