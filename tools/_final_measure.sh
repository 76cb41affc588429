# round-2 measurement suite (one gpurun call): tests, benches of every config, ncu launch list, per-conv ncu table
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > $O/r02_final_pytest.log 2>&1; echo "pytest exit $?" >> $O/r02_final_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_final_smoke.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r02_final_bench_c2.json 2> $O/r02_final_bench_c2.err
for c in 3 4 5; do timeout 300 python bench.py --config $c --steps 20 --warmup 5 > $O/r02_final_bench_c$c.json 2> $O/r02_final_bench_c$c.err; done
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_final_bench_ref.json 2> $O/r02_final_bench_ref.err
timeout 200 python tools/gpu_layer_times.py > $O/r02_final_layer_times.txt 2>&1
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-control --no-sustained"
$CMD > $O/r02_plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/r02_launches.csv $CMD > $O/r02_ncu_launches.log 2>&1
CMD2="python tools/profile_run.py --passes 3"
$CMD2 > $O/r02_plain_profile.log 2>&1 && ncu --set full --clock-control none -k regex:conv_halo_kernel -s 98 -c 49 -o $O/r02_convs -f $CMD2 > $O/r02_ncu_convs.log 2>&1
ncu -i $O/r02_convs.ncu-rep --page raw --csv > $O/r02_convs_raw.csv 2>/dev/null; rm -f $O/r02_convs.ncu-rep
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
ncu --metrics $M --cache-control none --clock-control none --graph-profiling node -c 420 --csv --log-file $O/r02_warm_dram_final.csv $CMD > $O/r02_warm_dram_final.log 2>&1
tail -3 $O/r02_final_pytest.log; ls -la $O | tail -15
