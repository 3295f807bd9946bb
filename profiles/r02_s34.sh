#!/bin/bash
AKB_MARCH_TMA=1 timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bcs.py tests/test_gpu_smr_mhd.py -m gpu -q -x 2>&1 | tail -4
echo "== tma"; AKB_MARCH_TMA=1 AKB_TRACE=1 python profiles/bench_variants.py --only headline --steps 10 2>&1 | grep -v "^$" | tail -2
echo "== cp.async"; AKB_MARCH_TMA=0 AKB_TRACE=1 python profiles/bench_variants.py --only headline --steps 10 2>&1 | grep -v "^$" | tail -2
echo "== tma 64 blocks"; AKB_MARCH_TMA=1 python profiles/bench_variants.py --only "64 blocks" --steps 10 2>&1 | grep -v "^$" | tail -1
echo "== cp.async 64 blocks"; AKB_MARCH_TMA=0 python profiles/bench_variants.py --only "64 blocks" --steps 10 2>&1 | grep -v "^$" | tail -1
AKB_MARCH_TMA=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 12 --csv --log-file gpurun_out/r02_s34_launches.csv \
   python profiles/bench_variants.py --only headline --steps 3 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_s34_launches.csv')) if len(r)>5]
h=rows[0]; ik=h.index('Kernel Name'); iv=h.index('Metric Value'); ig=h.index('Grid Size')
for r in rows[1:]: print(r[ik][:60], r[ig], r[iv])
PY
