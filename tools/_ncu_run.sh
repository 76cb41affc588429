cd $GRAFT_REPO_ROOT
CMD="python tools/profile_run.py --passes 3"
$CMD > gpurun_out/r02_plain.log 2>&1 || { echo plain run failed; tail -5 gpurun_out/r02_plain.log; exit 1; }
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:prep_s2d_kernel -s 2 -c 1 -o gpurun_out/r02_prof_prep -f $CMD > gpurun_out/r02_ncu_prep.log 2>&1
$NCU -k regex:maxpool3x3s2_kernel -s 2 -c 1 -o gpurun_out/r02_prof_pool -f $CMD > gpurun_out/r02_ncu_pool.log 2>&1
for n in prep pool; do
  ncu -i gpurun_out/r02_prof_$n.ncu-rep --page raw --csv > gpurun_out/r02_prof_${n}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r02_prof_$n.ncu-rep --page source --csv --print-source sass > gpurun_out/r02_prof_${n}_src.csv 2>/dev/null
  rm -f gpurun_out/r02_prof_$n.ncu-rep
done
ls -la gpurun_out/ | tail -8
