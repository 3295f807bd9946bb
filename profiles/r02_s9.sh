#!/bin/bash
for so in athenak_b200/libathenak_b200.so athenak_b200/var/lib_ppmng.so athenak_b200/var/lib_g0.so; do
  echo "== $so"
  ATHENAK_B200_LIB=$PWD/$so python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value'],1), round(d['ms_per_step'],3), {k: round(v['ms'],3) for k,v in d['roofline']['kernels'].items()})"
done
