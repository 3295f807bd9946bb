#!/bin/bash
# round 2, GPU session 1: full GPU test suite with the division guard + tightened AMR tolerances,
# FP64 issue-peak sweep, A/B of the guard on the headline bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_s1_smi.txt
python profiles/fp64_peak.py > gpurun_out/r02_s1_fp64.txt 2>&1
for so in athenak_b200/libathenak_b200.so athenak_b200/var/lib_g0.so; do
  echo "== $so" >> gpurun_out/r02_s1_ab.txt
  ATHENAK_B200_LIB=$PWD/$so python bench.py --steps 10 --warmup 3 --no-cpu-baseline >> gpurun_out/r02_s1_ab.txt 2>gpurun_out/r02_s1_ab.err
done
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02_s1_pytest.log 2>&1
tail -5 gpurun_out/r02_s1_pytest.log
python - <<'PY'
import json
for l in open('gpurun_out/r02_s1_ab.txt'):
    if l.startswith('=='): print(l.strip())
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value'],1), round(d['ms_per_step'],3), {k: round(v['ms'],3) for k,v in d['roofline']['kernels'].items()})
PY
tail -3 gpurun_out/r02_s1_fp64.txt
