#!/bin/bash
# round-end rehearsal: smoke(), the reference arm, the default bench command
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 | cut -c1-300
python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/r02_s47_bench.json 2> gpurun_out/r02_s47_bench.err; echo rc=$?
python -c "
import json; d=json.load(open('gpurun_out/r02_s47_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['cpu_baseline']['value'], d['roofline']['frac'], d['roofline']['fp64']['frac'])"
