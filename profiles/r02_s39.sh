#!/bin/bash
# AKB_COLD (scaled exact division on the repeat side of the range guard): selftest + parity subset, then the headline at two windows
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zamr.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r02_s39_pytest.log
cat gpurun_out/r02_s39_pytest.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_s39_bench_k20.json 2> gpurun_out/r02_s39_bench_k20.err
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_s39_bench_k30.json 2> gpurun_out/r02_s39_bench_k30.err
python - <<'PY'
import json
for f in ('k20','k30'):
    d=json.load(open(f'gpurun_out/r02_s39_bench_{f}.json'))
    print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernels']['fluxes(x1+x2+x3)'], d.get('config3_ot',{}).get('value'))
PY
