#!/bin/bash
# shared-interface PPM4 march (AKB_MARCH_SHARE) + marked-face repeat + AKB_COLD operators: full GPU suite, headline at two windows,
# launch list and FP64 / DRAM / local-memory counts of the sweeps
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r02_s45_pytest.log
cat gpurun_out/r02_s45_pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_s45_bench_k20.json 2> gpurun_out/r02_s45_bench_k20.err
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_s45_bench_k30.json 2> gpurun_out/r02_s45_bench_k30.err
python - <<'PY'
import json
for f in ('k20','k30'):
    d=json.load(open(f'gpurun_out/r02_s45_bench_{f}.json'))
    print(f, d['value'], d['ms_per_step'], d['e2e']['value'], {a:round(b['ms'],3) for a,b in d['roofline']['kernels'].items()}, d.get('config3_ot',{}).get('value'))
PY
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 200 -c 80 --csv \
    --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_launches_bench.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,sass__inst_executed_local_loads,sass__inst_executed_local_stores,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,sm__issue_active.avg.pct_of_peak_sustained_elapsed
ncu --metrics $M --clock-control none -k regex:flux_ -s 12 -c 3 --csv --log-file gpurun_out/r02_flux_counts.csv \
    python profiles/bench_variants.py --only headline --steps 2 > /dev/null 2>&1
grep -c flux_ gpurun_out/r02_flux_counts.csv
