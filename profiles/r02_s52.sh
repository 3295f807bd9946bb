#!/bin/bash
# where the 2-level MHD SMR variant's step time goes: device time between consecutive task bodies (AKB_TRACE=1), then the variant with the sweeps on one stream
mkdir -p gpurun_out
AKB_TRACE=1 python profiles/bench_variants.py --only "mhd SMR ppm4" --steps 3 > gpurun_out/r02_s52_trace.txt 2> gpurun_out/r02_s52_trace.err
grep AKB_TRACE gpurun_out/r02_s52_trace.err | tail -2 | cut -c1-1500
cat gpurun_out/r02_s52_trace.txt
AKB_SWEEP_STREAMS=0 python profiles/bench_variants.py --only "mhd SMR ppm4" --steps 10
python profiles/bench_variants.py --only "mhd SMR ppm4" --steps 10
python profiles/bench_variants.py --only "mhd SMR ppm4" --steps 40
