#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_hydro.py tests/test_gpu_tail.py -m gpu -q -x 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 60 -c 33 --csv --log-file gpurun_out/r02_s20_launches.csv \
   python profiles/bench_variants.py --only headline --steps 3 > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r02_s20_launches.csv')) if len(r)>5]
h=rows[0]; ik=h.index('Kernel Name'); iv=h.index('Metric Value'); im=h.index('Metric Name'); iid=h.index('ID'); ig=h.index('Grid Size')
d=collections.OrderedDict()
for r in rows[1:]:
    d.setdefault(r[iid],{'k':r[ik][:48],'g':r[ig]})[r[im]]=r[iv]
for k,v in list(d.items())[:33]: print(v['k'], v['g'], v.get('gpu__time_duration.sum'), v.get('dram__bytes_read.sum'), v.get('dram__bytes_write.sum'))
PY
AKB_TRACE=1 python profiles/bench_variants.py --only headline --steps 10 2>&1 | grep -v "^$" | tail -2
