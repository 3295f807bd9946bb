#!/bin/bash
mkdir -p gpurun_out
AKB_TAIL_SPLIT=1 timeout 900 python -m pytest tests/test_gpu_tail.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_tail.py -m gpu -x -q 2>&1 | tail -3
for so in athenak_b200/libathenak_b200.so athenak_b200/var/lib_f2.so athenak_b200/var/lib_g0.so; do
  echo "== $so  mono / split / off"
  ATHENAK_B200_LIB=$PWD/$so python profiles/bench_variants.py --only headline --steps 10
  AKB_TAIL_SPLIT=1 ATHENAK_B200_LIB=$PWD/$so python profiles/bench_variants.py --only headline --steps 10
  AKB_TAIL=0 ATHENAK_B200_LIB=$PWD/$so python profiles/bench_variants.py --only headline --steps 10
done
AKB_TAIL_SPLIT=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 60 -c 26 --csv --log-file gpurun_out/r02_s7_split.csv \
   python profiles/bench_variants.py --only headline --steps 3 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 60 -c 24 --csv --log-file gpurun_out/r02_s7_mono.csv \
   python profiles/bench_variants.py --only headline --steps 3 > /dev/null 2>&1
python - <<'PY'
import csv
for f in ('split','mono'):
    rows=[r for r in csv.reader(open(f'gpurun_out/r02_s7_{f}.csv')) if len(r)>5]
    h=rows[0]; ik=h.index('Kernel Name'); iv=h.index('Metric Value'); im=h.index('Metric Name'); iid=h.index('ID')
    d={}
    for r in rows[1:]:
        d.setdefault(r[iid],{'k':r[ik][:40]})[r[im]]=r[iv]
    print('==',f)
    for k,v in list(d.items())[:13]: print(v['k'], v.get('gpu__time_duration.sum'), v.get('dram__bytes_read.sum'), v.get('dram__bytes_write.sum'))
PY
