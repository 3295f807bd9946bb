#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_s37_bench.json 2> gpurun_out/r02_s37_bench.err
tail -3 gpurun_out/r02_s37_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_s37_bench.json'))
print(d['value'], d['ms_per_step'], d['clocks']); print(d['e2e']['value'], d['e2e']['ms_per_step']); print(d.get('tma_march')); print(d.get('reference_gpu')); print(d['cpu_baseline']); print(d['roofline']['frac'], d['roofline']['fp64']['frac'], d['roofline']['traffic'])
PY
python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-400
