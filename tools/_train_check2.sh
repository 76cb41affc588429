cd $GRAFT_REPO_ROOT
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_training_gpu.py -q -x > $O/r02_train_tests2.log 2>&1
tail -25 $O/r02_train_tests2.log
timeout 200 python tools/gpu_train_profile.py --top 24 > $O/r02_train_profile_native2.txt 2>&1
head -34 $O/r02_train_profile_native2.txt
timeout 250 python bench.py --config 5 --steps 20 --warmup 5 > $O/r02_c5_final.json 2> $O/r02_c5_final.err
tail -c 2500 $O/r02_c5_final.json; tail -3 $O/r02_c5_final.err
