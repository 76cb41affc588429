cd $GRAFT_REPO_ROOT
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 150 $TR --master-port 29513 bench.py --gpus 8 --config 5 --steps 20 --warmup 5 > $O/r02_scale8_c5_final.json 2> $O/r02_scale8_c5_final.err
tail -c 1500 $O/r02_scale8_c5_final.json; tail -c 600 $O/r02_scale8_c5_final.err
