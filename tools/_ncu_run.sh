cd $GRAFT_REPO_ROOT
CMD="python tools/profile_run.py --passes 3"
$CMD > gpurun_out/r02_plain.log 2>&1 || { echo plain run failed; tail -5 gpurun_out/r02_plain.log; exit 1; }
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
$NCU -k 'regex:conv_halo_kernel<\(int\)64, \(int\)3, \(int\)3, \(int\)2, \(bool\)1, \(bool\)1, \(int\)0, \(bool\)0, \(bool\)0' -s 14 -c 2 -o gpurun_out/r02_prof_l1 -f $CMD > gpurun_out/r02_ncu_l1.log 2>&1
$NCU -k 'regex:conv_halo_kernel<\(int\)64, \(int\)3, \(int\)3, \(int\)2, \(bool\)0, \(bool\)1, \(int\)0, \(bool\)0, \(bool\)0' -s 44 -c 2 -o gpurun_out/r02_prof_l2 -f $CMD > gpurun_out/r02_ncu_l2.log 2>&1
$NCU -k 'regex:conv_halo_kernel<\(int\)64, \(int\)3, \(int\)3, \(int\)2, \(bool\)1, \(bool\)1, \(int\)0, \(bool\)1, \(bool\)0' -s 4 -c 2 -o gpurun_out/r02_prof_tail -f $CMD > gpurun_out/r02_ncu_tail.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -3 gpurun_out/r02_ncu_l1.log
for n in l1 l2 tail; do
  ncu -i gpurun_out/r02_prof_$n.ncu-rep --page raw --csv > gpurun_out/r02_prof_${n}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r02_prof_$n.ncu-rep --page source --csv --print-source sass > gpurun_out/r02_prof_${n}_src.csv 2>/dev/null
  rm -f gpurun_out/r02_prof_$n.ncu-rep
done
ls -la gpurun_out/ | tail -12
