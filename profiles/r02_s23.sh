#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "pipeline or cxx_driver" 2>&1 | tail -5
python bench.py --no-cpu-baseline > gpurun_out/r02_s23_bench.json 2> gpurun_out/r02_s23_bench.err
tail -5 gpurun_out/r02_s23_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_s23_bench.json'))
print(d['value'], d['ms_per_step'], d['clocks']); print(d['e2e'])
PY
