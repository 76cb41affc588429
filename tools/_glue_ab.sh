cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -k "prep or maxpool or stem" 2>&1 | tail -3
for r in 8 4 2; do UWM_POOL_ROWS=$r timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-gpu-control --no-sustained > gpurun_out/r02_glue_bench.json 2> gpurun_out/r02_glue_bench.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_glue_bench.json').read().strip().splitlines()[-1])
    print('POOL_ROWS=$r', d['value'], d['ms_per_step'], d['frac_of_bf16_peak'], d['roofline']['frac'], d['roofline']['eager_events'])
except Exception as e: print('ERR', e)
PY
done
UWM_POOL_ROWS=4 python tools/gpu_layer_times.py 2>/dev/null | head -8
