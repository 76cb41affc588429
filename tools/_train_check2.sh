cd $GRAFT_REPO_ROOT
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_training_gpu.py -q > $O/r02_train_tests2.log 2>&1
tail -12 $O/r02_train_tests2.log
timeout 200 python tools/gpu_train_profile.py --top 16 > $O/r02_train_profile_native2.txt 2>&1
sed -n 9,26p $O/r02_train_profile_native2.txt | cut -c1-120
F="--steps 20 --warmup 5 --no-cpu-baseline --no-gpu-control --no-sustained"
for v in "1 1" "1 0" "0 1" "0 0"; do
  set -- $v
  UWM_NATIVE_POOL=$1 UWM_NATIVE_PACK=$2 timeout 200 python bench.py --config 5 $F > $O/r02_c5_pool$1_pack$2.json 2> $O/r02_c5_pool$1_pack$2.err
  python - $O/r02_c5_pool$1_pack$2.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"], 1), "img/s", round(d["ms_per_step"], 3), "ms  e2e", round(d["e2e"]["value"], 1), "loss", d.get("final_loss"), "launches", d.get("gpu_launches"))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
  tail -2 $O/r02_c5_pool$1_pack$2.err
done
