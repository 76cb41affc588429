# end-of-round check (one gpurun call): the whole GPU suite, smoke, the training profile, bench lines of configs 5 and 2
cd $GRAFT_REPO_ROOT
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/r02_final2_pytest.log 2>&1; echo "pytest exit $?" >> $O/r02_final2_pytest.log
tail -6 $O/r02_final2_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_final2_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $O/r02_final2_smoke.log
timeout 200 python tools/gpu_train_profile.py --top 40 > $O/r02_train_profile_final.txt 2>&1
sed -n 1,5p $O/r02_train_profile_final.txt; sed -n 9,34p $O/r02_train_profile_final.txt | cut -c1-120
timeout 250 python bench.py --config 5 --steps 20 --warmup 5 > $O/r02_final2_bench_c5.json 2> $O/r02_final2_bench_c5.err
tail -c 1800 $O/r02_final2_bench_c5.json; tail -3 $O/r02_final2_bench_c5.err
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r02_final2_bench_c2.json 2> $O/r02_final2_bench_c2.err
head -c 700 $O/r02_final2_bench_c2.json; tail -3 $O/r02_final2_bench_c2.err
