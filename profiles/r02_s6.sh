#!/bin/bash
mkdir -p gpurun_out
AKB_TAIL=0 ncu --set full --import-source on --clock-control none -k regex:flux_ -s 12 -c 3 -o gpurun_out/r02_s6_y -f \
   python profiles/bench_variants.py --only headline --steps 2 > gpurun_out/r02_s6_n.log 2>&1
ls -la gpurun_out/
