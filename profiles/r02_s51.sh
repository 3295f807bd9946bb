#!/bin/bash
# launch list of the 2-level MHD SMR variant (15 blocks of 64^3) with the final library, to compare with profiles/r01_launches_mhd_smr_session4b.txt
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 100 --csv --log-file gpurun_out/r02_launches_mhd_smr.csv \
   python profiles/bench_variants.py --only "mhd SMR ppm4" --steps 4 > /dev/null 2>&1
python - <<'PY'
import csv,collections,re
rows=[r for r in csv.reader(open('gpurun_out/r02_launches_mhd_smr.csv')) if len(r)>5]
h=rows[0]; ik=h.index('Kernel Name'); iv=h.index('Metric Value')
c=collections.OrderedDict()
for r in rows[1:]:
    k=re.sub(r'\(.*','',r[ik])[:60]; a=c.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=float(r[iv].replace(',',''))/1e3
tot=sum(v[1] for v in c.values())
for k,v in sorted(c.items(), key=lambda kv:-kv[1][1]): print(f"{k:62s} n={v[0]:3d} total={v[1]:9.1f} us  avg={v[1]/v[0]:8.1f} us  {100*v[1]/tot:5.1f}%")
print(f"total {tot:.1f} us over {sum(v[0] for v in c.values())} launches")
PY
