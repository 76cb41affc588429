# round-2 follow-up (one gpurun call): sub-batch sweep, native-dgrad A/B on config 5, resnet50 per-kernel table, the
# new parity tests, bench lines of configs 3 / 4 with the sweep's best sub-batch, then the whole GPU suite
cd $GRAFT_REPO_ROOT
O=gpurun_out
mkdir -p $O
timeout 400 python tools/gpu_subbatch_sweep.py --json $O/r02_subbatch_sweep.json > $O/r02_subbatch_sweep.txt 2>&1
tail -14 $O/r02_subbatch_sweep.txt
timeout 300 python -m pytest tests/test_training_gpu.py tests/test_model_gpu.py -q -k "dgrad or sub_batched or train_mode or train_steps" > $O/r02_new_tests.log 2>&1
tail -5 $O/r02_new_tests.log
timeout 200 python tools/gpu_layer_times.py --encoder resnet50 --size 768 --batch 32 > $O/r02_layer_times_r50.txt 2>&1
timeout 200 python tools/gpu_layer_times.py --encoder resnet50 --size 768 --batch 4 > $O/r02_layer_times_r50_b4.txt 2>&1
F="--steps 20 --warmup 5 --no-cpu-baseline --no-gpu-control --no-sustained"
UWM_NATIVE_DGRAD=0 timeout 200 python bench.py --config 5 $F > $O/r02_c5_dgrad0.json 2> $O/r02_c5_dgrad0.err
UWM_NATIVE_DGRAD=1 timeout 200 python bench.py --config 5 $F > $O/r02_c5_dgrad1.json 2> $O/r02_c5_dgrad1.err
python - <<'PY' > $O/r02_subbatch_best.env
import json
rows = json.load(open("gpurun_out/r02_subbatch_sweep.json"))
for enc, cfg in (("resnet50", 4), ("resnet34", 3)):
    rs = [r for r in rows if r["encoder"] == enc and r["size"] != 512]
    best = max(rs, key=lambda r: r["images_per_s"])
    print(f"SB{cfg}={best['sub_batch']}")
PY
. $O/r02_subbatch_best.env; cat $O/r02_subbatch_best.env
G="--steps 20 --warmup 5 --no-cpu-baseline"
UWM_SUBBATCH=$SB4 timeout 300 python bench.py --config 4 $G > $O/r02_bench_c4_sub.json 2> $O/r02_bench_c4_sub.err
UWM_SUBBATCH=$SB3 timeout 300 python bench.py --config 3 $G > $O/r02_bench_c3_sub.json 2> $O/r02_bench_c3_sub.err
timeout 200 python bench.py --config 2 $G > $O/r02_bench_c2_again.json 2> $O/r02_bench_c2_again.err
for f in c5_dgrad0 c5_dgrad1 bench_c4_sub bench_c3_sub bench_c2_again; do python - $O/r02_$f.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["value"], d["ms_per_step"], d.get("frac_of_bf16_peak"), d.get("e2e", {}).get("value"))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
timeout 900 python -m pytest tests -m gpu -q > $O/r02_followup_pytest.log 2>&1; echo "pytest exit $?" >> $O/r02_followup_pytest.log
tail -4 $O/r02_followup_pytest.log
