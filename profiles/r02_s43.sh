#!/bin/bash
# N=2 sanity of the final library: parity leg (exit 2 if a bit differs) + headline + config-3 layout + pipelined e2e
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_s43_bench2.json 2> gpurun_out/r02_s43_bench2.err
echo rc=$?
grep -h '^{"metric"' gpurun_out/r02_s43_bench2.json gpurun_out/r02_s43_bench2.err > gpurun_out/r02_s43_line.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_s43_line.json').readline())
print(d['value'], d['ms_per_step'], d.get('parity')); print(d['e2e']['value'], d['e2e']['ms_per_step']); print(d.get('config3_ot'))
PY
tail -3 gpurun_out/r02_s43_bench2.err | cut -c1-300
