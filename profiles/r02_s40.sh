#!/bin/bash
# marked-face repeat after the march (march_fix_faces) + AKB_COLD operators: full GPU suite, then the headline at two windows
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r02_s40_pytest.log
cat gpurun_out/r02_s40_pytest.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_s40_bench_k20.json 2> gpurun_out/r02_s40_bench_k20.err
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_s40_bench_k30.json 2> gpurun_out/r02_s40_bench_k30.err
python - <<'PY'
import json
for f in ('k20','k30'):
    d=json.load(open(f'gpurun_out/r02_s40_bench_{f}.json'))
    print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernels'], d.get('config3_ot',{}).get('value'))
PY
