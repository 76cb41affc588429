cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py -x -q 2>&1 | tail -4
timeout 600 python tools/gpu_chain_ab.py 2>&1 | tail -9
for c in "0 0" "1 200"; do set -- $c; UWM_CHAIN=$1 UWM_CHAIN_MIN_TILES=$2 timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-gpu-control --no-sustained > gpurun_out/r02_chain_bench.json 2> gpurun_out/r02_chain_bench.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_chain_bench.json').read().strip().splitlines()[-1])
    print('CHAIN=$1 MIN_TILES=$2', d['value'], d['ms_per_step'], d['frac_of_bf16_peak'], d['roofline']['frac'], d['gpu_launches'])
except Exception as e: print('ERR', e); print(open('gpurun_out/r02_chain_bench.err').read()[-800:])
PY
done
