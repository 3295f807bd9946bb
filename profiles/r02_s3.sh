#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r02_s3_ab.txt
for so in athenak_b200/var/lib_g0.so athenak_b200/libathenak_b200.so; do
  echo "== $so" >> gpurun_out/r02_s3_ab.txt
  ATHENAK_B200_LIB=$PWD/$so python bench.py --steps 10 --warmup 3 --no-cpu-baseline >> gpurun_out/r02_s3_ab.txt 2>gpurun_out/r02_s3_ab.err
done
python - <<'PY'
import json
for l in open('gpurun_out/r02_s3_ab.txt'):
    if l.startswith('=='): print(l.strip())
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value'],1), round(d['ms_per_step'],3), {k: round(v['ms'],3) for k,v in d['roofline']['kernels'].items()})
PY
timeout 2400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zamr.py tests/test_gpu_lwave.py -m gpu -q > gpurun_out/r02_s3_pytest.log 2>&1
tail -8 gpurun_out/r02_s3_pytest.log
