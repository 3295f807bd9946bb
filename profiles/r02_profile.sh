#!/bin/bash
# Round-2 evidence, same command as the bench: (1) launch list with time + DRAM bytes, (2) FP64 / instruction counts of the sweeps
# (cp.async and TMA form), (3) one --set full capture of the three sweeps of a stage.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 200 -c 80 --csv \
    --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_launches_bench.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread
ncu --metrics $M --clock-control none -k regex:flux_ -s 12 -c 3 --csv --log-file gpurun_out/r02_flux_counts.csv \
    python profiles/bench_variants.py --only headline --steps 2 > /dev/null 2>&1
AKB_MARCH_TMA=1 ncu --metrics $M --clock-control none -k regex:flux_ -s 12 -c 3 --csv --log-file gpurun_out/r02_flux_counts_tma.csv \
    python profiles/bench_variants.py --only headline --steps 2 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:flux_ -s 12 -c 3 -o gpurun_out/r02_flux_full -f \
    python profiles/bench_variants.py --only headline --steps 2 > gpurun_out/r02_flux_full.log 2>&1
ls -la gpurun_out | tail -8
