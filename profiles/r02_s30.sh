#!/bin/bash
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -4
for so in athenak_b200/libathenak_b200.so; do
  echo "== $so"
  AKB_TRACE=1 python profiles/bench_variants.py --only headline --steps 10 2>&1 | grep -v "^$" | tail -2
  AKB_TRACE=1 python profiles/bench_variants.py --only headline --steps 35 2>&1 | grep -v "^$" | tail -2
done
ncu --metrics gpu__time_duration.sum --clock-control none -s 1380 -c 14 --csv --log-file gpurun_out/r02_s30_launches.csv \
   python profiles/bench_variants.py --only headline --steps 35 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_s30_launches.csv')) if len(r)>5]
h=rows[0]; ik=h.index('Kernel Name'); iv=h.index('Metric Value'); ig=h.index('Grid Size')
for r in rows[1:]: print(r[ik][:60], r[ig], r[iv])
PY
