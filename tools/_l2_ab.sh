#!/bin/bash
# A/B of the L2 management switches (UWM_L2 bit mask) + parity under them + warm DRAM traffic per kernel
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-control --no-sustained"
for m in ${MODES:-0 1 2 3}; do
  UWM_L2=$m timeout 600 python -m pytest tests/test_model_gpu.py -x -q 2>&1 | tail -2
done
for rep in 1 2; do
  for m in ${MODES:-0 1 2 3}; do
    UWM_L2=$m $B --config 2 > gpurun_out/l2_${m}_c2_$rep.json 2>>gpurun_out/l2_err.log
  done
done
for cfg in 3 4; do
  for m in ${MODES:-0 1 2 3}; do
    UWM_L2=$m $B --config $cfg > gpurun_out/l2_${m}_c${cfg}_1.json 2>>gpurun_out/l2_err.log
  done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/l2_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['roofline']['frac'])
    except Exception as e:
        print(f, 'ERR', e)
PY
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for m in ${NCU_MODES:-3}; do
  UWM_L2=$m ncu --metrics $M --cache-control none --clock-control none --graph-profiling node -c 420 --csv \
    --log-file gpurun_out/warm_dram_l2_$m.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-control --no-sustained \
    > gpurun_out/warm_dram_l2_$m.log 2>&1
done
