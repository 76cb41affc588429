cd $GRAFT_REPO_ROOT
O=gpurun_out
mkdir -p $O
timeout 400 python -m pytest tests/test_training_gpu.py -q -x > $O/r02_train_tests.log 2>&1
tail -25 $O/r02_train_tests.log
timeout 300 python tools/gpu_train_profile.py --top 24 > $O/r02_train_profile_native.txt 2>&1
head -36 $O/r02_train_profile_native.txt
F="--steps 20 --warmup 5 --no-cpu-baseline --no-gpu-control --no-sustained"
for v in "1 1" "1 0" "0 1" "0 0"; do
  set -- $v
  UWM_TRAIN_GRAPH=$1 UWM_NATIVE_DGRAD=$2 timeout 200 python bench.py --config 5 $F > $O/r02_c5_g$1_d$2.json 2> $O/r02_c5_g$1_d$2.err
  python - $O/r02_c5_g$1_d$2.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d["value"], 1), "img/s", round(d["ms_per_step"], 3), "ms  e2e", round(d["e2e"]["value"], 1), "loss", d.get("final_loss"), "launches", d.get("gpu_launches"))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
  tail -2 $O/r02_c5_g$1_d$2.err
done
