#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 80 -c 40 --csv --log-file gpurun_out/r02_s14_c3.csv \
   python profiles/bench_variants.py --only "64 blocks" --steps 3 > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r02_s14_c3.csv')) if len(r)>5]
h=rows[0]; ik=h.index('Kernel Name'); iv=h.index('Metric Value'); im=h.index('Metric Name'); iid=h.index('ID'); ig=h.index('Grid Size')
d=collections.OrderedDict()
for r in rows[1:]:
    d.setdefault(r[iid],{'k':r[ik][:48],'g':r[ig]})[r[im]]=r[iv]
for k,v in list(d.items())[:40]: print(v['k'], v['g'], v.get('gpu__time_duration.sum'), v.get('dram__bytes_read.sum'), v.get('dram__bytes_write.sum'))
PY
python profiles/bench_variants.py --only "64 blocks" --steps 10
