#!/bin/bash
mkdir -p gpurun_out
for so in athenak_b200/libathenak_b200.so athenak_b200/var/lib_m3.so athenak_b200/var/lib_x3.so athenak_b200/var/lib_m3x3.so athenak_b200/var/lib_g0.so; do
  echo "== $so"
  AKB_LIB_LAX=1 ATHENAK_B200_LIB=$PWD/$so python profiles/bench_variants.py --only headline --steps 10
done
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_s10_bench.json 2> gpurun_out/r02_s10_bench.err
tail -3 gpurun_out/r02_s10_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_s10_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e'],'c3',d.get('config3_ot'))
print('fp64',d['roofline']['fp64'])
print({k: round(v['ms'],3) for k,v in d['roofline']['kernels'].items()}, d.get('cpu_baseline'))
PY
