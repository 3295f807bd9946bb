#!/bin/bash
# profiles/sweep.sh -- times tuning variants of the library (built by hand into
# athenak_b200/variants/) with the bench's per-kernel CUDA-event timings.
for so in athenak_b200/variants/lib_*.so; do
  echo "== $so"
  ATHENAK_B200_LIB=$PWD/$so python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); k = d['roofline']['kernels']
        print('value', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 2), {n: round(v['ms'], 3) for n, v in k.items()})
    elif 'Error' in l or 'error' in l: print(l.strip())
"
done
