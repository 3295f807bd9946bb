#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tail.py -m gpu -x -q 2>&1 | tail -15
echo "== fluxupd / tasks"
python profiles/bench_variants.py --only headline --steps 10
AKB_FLUXUPD=0 python profiles/bench_variants.py --only headline --steps 10
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 60 -c 22 --csv --log-file gpurun_out/r02_s8_fu.csv \
   python profiles/bench_variants.py --only headline --steps 3 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_s8_fu.csv')) if len(r)>5]
h=rows[0]; ik=h.index('Kernel Name'); iv=h.index('Metric Value'); im=h.index('Metric Name'); iid=h.index('ID')
d={}
for r in rows[1:]:
    d.setdefault(r[iid],{'k':r[ik][:44]})[r[im]]=r[iv]
for k,v in list(d.items())[:22]: print(v['k'], v.get('gpu__time_duration.sum'), v.get('dram__bytes_read.sum'), v.get('dram__bytes_write.sum'))
PY
