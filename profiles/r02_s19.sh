#!/bin/bash
AKB_TRACE=1 python profiles/bench_variants.py --only headline --steps 10 2>&1 | grep -v "^$" | tail -4
AKB_TRACE=1 AKB_SWEEP_STREAMS=0 python profiles/bench_variants.py --only headline --steps 10 2>&1 | grep -v "^$" | tail -3
AKB_TRACE=1 AKB_OVERLAP_U=0 python profiles/bench_variants.py --only headline --steps 10 2>&1 | grep -v "^$" | tail -3
