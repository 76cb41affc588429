#!/bin/bash
# A/B of the serpentine tile order and the N=64 CTA-pair switch (bench.py value, config 2 / 3 / 4)
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-control --no-sustained"
for cfg in 2; do
  for rep in 1 2; do
    UWM_SERP=0 $B --config $cfg > gpurun_out/ab_serp0_c${cfg}_$rep.json 2>gpurun_out/ab_err.log
    UWM_SERP=1 $B --config $cfg > gpurun_out/ab_serp1_c${cfg}_$rep.json 2>>gpurun_out/ab_err.log
    UWM_SERP=1 UWM_CG2_N64=1 $B --config $cfg > gpurun_out/ab_serp1_pair64_c${cfg}_$rep.json 2>>gpurun_out/ab_err.log
    UWM_SERP=0 UWM_CG2_N64=1 $B --config $cfg > gpurun_out/ab_serp0_pair64_c${cfg}_$rep.json 2>>gpurun_out/ab_err.log
  done
done
for cfg in 3 4; do
  UWM_SERP=0 $B --config $cfg > gpurun_out/ab_serp0_c${cfg}_1.json 2>>gpurun_out/ab_err.log
  UWM_SERP=1 $B --config $cfg > gpurun_out/ab_serp1_c${cfg}_1.json 2>>gpurun_out/ab_err.log
  UWM_SERP=1 UWM_CG2_N64=1 $B --config $cfg > gpurun_out/ab_serp1_pair64_c${cfg}_1.json 2>>gpurun_out/ab_err.log
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/ab_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['roofline']['frac'])
    except Exception as e:
        print(f, 'ERR', e)
PY
UWM_CG2_N64=1 timeout 900 python -m pytest tests/test_model_gpu.py -x -q 2>&1 | tail -3
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
