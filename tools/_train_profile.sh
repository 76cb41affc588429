cd $GRAFT_REPO_ROOT
O=gpurun_out
mkdir -p $O
timeout 300 python tools/gpu_train_profile.py > $O/r02_train_profile.txt 2>&1
head -60 $O/r02_train_profile.txt
G="--steps 20 --warmup 5 --no-cpu-baseline --no-gpu-control --no-sustained"
timeout 200 python bench.py --config 4 $G > $O/r02_bench_c4_bb.json 2> $O/r02_bench_c4_bb.err; tail -c 1500 $O/r02_bench_c4_bb.json; tail -3 $O/r02_bench_c4_bb.err
timeout 200 python bench.py --config 2 $G > $O/r02_bench_c2_bb.json 2> $O/r02_bench_c2_bb.err; tail -c 1500 $O/r02_bench_c2_bb.json; tail -3 $O/r02_bench_c2_bb.err
