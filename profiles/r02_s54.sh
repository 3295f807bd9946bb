#!/bin/bash
# final library (prediction skip reverted): the parity file with the new slab pattern, smoke(), one headline line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_s54_bench_k20.json 2> gpurun_out/r02_s54_bench_k20.err
python -c "
import json; d=json.load(open('gpurun_out/r02_s54_bench_k20.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
