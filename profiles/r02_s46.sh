#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02_s46_bench8.json 2> gpurun_out/r02_s46_bench8.err
echo rc=$?
grep -h '^{"metric"' gpurun_out/r02_s46_bench8.json gpurun_out/r02_s46_bench8.err > gpurun_out/r02_s46_line.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_s46_line.json').readline())
print(d['value'], d['ms_per_step'], d.get('parity')); print(d['e2e']); print(d.get('config3_ot')); print(d.get('tma_march'))
PY
tail -3 gpurun_out/r02_s46_bench8.err | cut -c1-300
