cd $GRAFT_REPO_ROOT
O=gpurun_out
for d in 1 0; do echo "== UWM_NATIVE_DGRAD=$d"; UWM_NATIVE_DGRAD=$d timeout 200 python tools/gpu_train_e2e_diag.py 2>&1 | tee $O/r02_train_e2e_diag_d$d.txt | tail -13; done
