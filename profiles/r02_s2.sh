#!/bin/bash
# round 2, GPU session 2: A/B of the division-guard forms on the headline bench; full GPU suite with the branchy guard
mkdir -p gpurun_out
rm -f gpurun_out/r02_s2_ab.txt
for so in athenak_b200/var/lib_g0.so athenak_b200/libathenak_b200.so athenak_b200/var/lib_g2.so; do
  echo "== $so" >> gpurun_out/r02_s2_ab.txt
  ATHENAK_B200_LIB=$PWD/$so python bench.py --steps 10 --warmup 3 --no-cpu-baseline >> gpurun_out/r02_s2_ab.txt 2>gpurun_out/r02_s2_ab.err
done
python - <<'PY'
import json
for l in open('gpurun_out/r02_s2_ab.txt'):
    if l.startswith('=='): print(l.strip())
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value'],1), round(d['ms_per_step'],3), {k: round(v['ms'],3) for k,v in d['roofline']['kernels'].items()})
PY
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02_s2_pytest.log 2>&1
tail -15 gpurun_out/r02_s2_pytest.log
