#!/bin/bash
# generic A/B: ENVS="A=1|B=2 C=3|..." (| separated environment sets, first one may be empty = baseline), CFGS="2 3 4"
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-control --no-sustained"
IFS='|' read -ra SETS <<< "${ENVS}"
for cfg in ${CFGS:-2}; do
  for rep in $(seq 1 ${REPS:-2}); do
    i=0
    for e in "${SETS[@]}"; do
      env $e $B --config $cfg > gpurun_out/ab_${TAG:-x}_${i}_c${cfg}_$rep.json 2>>gpurun_out/ab_err.log
      i=$((i+1))
    done
  done
done
python - <<'PY'
import json,glob,os
tag=os.environ.get('TAG','x')
for f in sorted(glob.glob(f'gpurun_out/ab_{tag}_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['value'], round(d['ms_per_step'],4), round(d['roofline']['frac'],4))
    except Exception as e:
        print(f, 'ERR', e)
PY
