#!/bin/bash
AKB_PIPE_TRACE=1 python profiles/e2e_pipe.py --slots 4 --jobs 12 2>&1 | tail -10
python profiles/e2e_pipe.py --slots 3 --jobs 12 2>&1 | tail -1
python profiles/e2e_pipe.py --slots 5 --jobs 20 2>&1 | tail -1
python profiles/e2e_pipe.py --slots 6 --jobs 24 2>&1 | tail -1
