#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-control --no-sustained"
for rep in 1 2; do
  for a in 2 3 4; do
    UWM_SPX_ASTAGES=$a $B --config 2 > gpurun_out/spx_a${a}_c2_$rep.json 2>>gpurun_out/l2_err.log
  done
done
for a in 2 3 4; do UWM_SPX_ASTAGES=$a $B --config 3 > gpurun_out/spx_a${a}_c3_1.json 2>>gpurun_out/l2_err.log; done
for a in 2 3; do UWM_SPX_ASTAGES=$a $B --config 4 > gpurun_out/spx_a${a}_c4_1.json 2>>gpurun_out/l2_err.log; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/spx_a*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['roofline']['frac'])
    except Exception as e:
        print(f, 'ERR', e)
PY
