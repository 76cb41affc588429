#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -k "s2d" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_model_gpu.py -x -q 2>&1 | tail -2
UWM_VERBOSE=1 python tools/gpu_trace.py --filter "dec4 s2d" > gpurun_out/trace_s2d_b.txt 2>&1
grep "halo conv" gpurun_out/trace_s2d_b.txt | head -3
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-control --no-sustained"
for rep in 1 2; do
  for m in 0 1 3; do
    UWM_L2=$m $B --config 2 > gpurun_out/s2db_l2_${m}_c2_$rep.json 2>>gpurun_out/l2_err.log
  done
done
for m in 0 1; do UWM_L2=$m $B --config 3 > gpurun_out/s2db_l2_${m}_c3_1.json 2>>gpurun_out/l2_err.log; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s2db_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['roofline']['frac'])
    except Exception as e:
        print(f, 'ERR', e)
PY
python tools/gpu_layer_times.py 2>&1 | tail -12
