#!/bin/bash
# A/B of the tile mapping of the stencil kernels at 257^3 positions (corner EMF, CT): 3-D tiles (87.8 % of the launched threads in range,
# default) against the flattened (j,i) mapping (AKB_FLAT_THRESHOLD=0.95), per-task times of bench.py; then the N=1 launch list of the winner
mkdir -p gpurun_out
for thr in 0.85 0.95; do
  AKB_FLAT_THRESHOLD=$thr python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_s44_thr$thr.json 2> gpurun_out/r02_s44_thr$thr.err
done
python - <<'PY'
import json
for t in ('0.85','0.95'):
    d=json.load(open(f'gpurun_out/r02_s44_thr{t}.json'))
    print(t, round(d['value'],1), round(d['ms_per_step'],3), {a:round(b['ms'],3) for a,b in d['roofline']['kernels'].items()}, round(d.get('config3_ot',{}).get('value',0),1))
PY
