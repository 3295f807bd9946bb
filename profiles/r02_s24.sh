#!/bin/bash
AKB_PIPE_TRACE=1 python profiles/e2e_pipe.py --slots 4 --jobs 12 2>&1 | tail -18
python profiles/e2e_pipe.py --slots 3 --jobs 12 2>&1 | tail -1
python profiles/e2e_pipe.py --slots 6 --jobs 18 2>&1 | tail -1
