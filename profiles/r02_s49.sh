#!/bin/bash
# side measurement with the final library: the other solver families and mesh layouts that share the sweep kernels (not the headline)
mkdir -p gpurun_out
python profiles/bench_variants.py --steps 10 > gpurun_out/r02_variants.txt 2> gpurun_out/r02_variants.err
python profiles/bench_variants.py --steps 6 --only "config 5" >> gpurun_out/r02_variants.txt 2>> gpurun_out/r02_variants.err
cat gpurun_out/r02_variants.txt; tail -2 gpurun_out/r02_variants.err
