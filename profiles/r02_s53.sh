#!/bin/bash
# far-field-shell prediction (at_floor): full GPU suite, headline at the two windows
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r02_s53_pytest.log
cat gpurun_out/r02_s53_pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_s53_bench_k20.json 2> gpurun_out/r02_s53_bench_k20.err
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_s53_bench_k30.json 2> gpurun_out/r02_s53_bench_k30.err
python - <<'PY'
import json
for f in ('k20','k30'):
    d=json.load(open(f'gpurun_out/r02_s53_bench_{f}.json'))
    print(f, d['value'], d['ms_per_step'], d['e2e']['value'], {a:round(b['ms'],3) for a,b in d['roofline']['kernels'].items()}, d.get('config3_ot',{}).get('value'))
PY
