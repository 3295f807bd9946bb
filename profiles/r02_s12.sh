#!/bin/bash
mkdir -p gpurun_out
for ov in 1 0; do
  AKB_OVERLAP_PUSH=$ov timeout 900 python -X faulthandler -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 15 --warmup 4 > gpurun_out/r02_s12_bench2_ov$ov.json 2> gpurun_out/r02_s12_bench2_ov$ov.err
  grep -v "^$" gpurun_out/r02_s12_bench2_ov$ov.err | grep -A12 "Fatal Python\|FATAL" | head -30
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02_s12_bench2_ov$ov.json') if l.startswith('{')][-1])
    print('ov$ov value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'parity',d.get('parity',{}).get('max_abs_diff'),d.get('parity',{}).get('peer_push'),'e2e',round(d['e2e']['value'],1),d['e2e']['mode'][:40],'serial',round(d['e2e']['serial_value'],1),'c3',d.get('config3_ot',{}).get('value'), d.get('config3_ot',{}).get('error'))
except Exception as ex: print('parse fail',ex)
PY
done
