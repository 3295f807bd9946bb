#!/bin/bash
mkdir -p gpurun_out
ATHENAK_B200_LIB=$PWD/athenak_b200/var/lib_g0.so ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/r02_s5_launches.csv \
   python profiles/bench_variants.py --only headline --steps 3 > gpurun_out/r02_s5_l.log 2>&1
ATHENAK_B200_LIB=$PWD/athenak_b200/var/lib_g0.so ncu --set full --import-source on --clock-control none -k regex:stage_tail -s 4 -c 1 -o gpurun_out/r02_s5_tail -f \
   python profiles/bench_variants.py --only headline --steps 2 > gpurun_out/r02_s5_n.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_s5_launches.csv')) if len(r)>5]
h=rows[0]; ik=h.index('Kernel Name'); iv=h.index('Metric Value'); 
for r in rows[1:]:
    print(r[ik][:70], r[iv])
PY
