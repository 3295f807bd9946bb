#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 1380 -c 40 --csv --log-file gpurun_out/r02_s29_launches.csv \
   python profiles/bench_variants.py --only headline --steps 35 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_s29_launches.csv')) if len(r)>5]
h=rows[0]; ik=h.index('Kernel Name'); iv=h.index('Metric Value'); ig=h.index('Grid Size')
for r in rows[1:]: print(r[ik][:60], r[ig], r[iv])
PY
