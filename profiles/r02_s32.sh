#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r02_s32_pytest.log
cat gpurun_out/r02_s32_pytest.log
python bench.py > gpurun_out/r02_s32_bench.json 2> gpurun_out/r02_s32_bench.err
tail -3 gpurun_out/r02_s32_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_s32_bench.json'))
print(d['value'], d['ms_per_step'], d['clocks']); print(d['e2e']); print(d['roofline']['kernels']); print(d.get('config3_ot')); print(d['cpu_baseline'])
PY
