#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_s33_bench2.json 2> gpurun_out/r02_s33_bench2.err
echo rc=$?
tail -5 gpurun_out/r02_s33_bench2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_s33_bench2.json'))
print(d['value'], d['ms_per_step'], d.get('parity')); print(d['e2e']); print(d.get('config3_ot'))
PY
