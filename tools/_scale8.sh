cd $GRAFT_REPO_ROOT
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --no-gpu-control > $O/r02_scale8_c2.json 2> $O/r02_scale8_c2.err
timeout 400 $TR --master-port 29512 bench.py --gpus 8 --config 3 --steps 20 --warmup 5 --no-gpu-control --no-sustained > $O/r02_scale8_c3.json 2> $O/r02_scale8_c3.err
timeout 400 $TR --master-port 29513 bench.py --gpus 8 --config 5 --steps 10 --warmup 3 > $O/r02_scale8_c5.json 2> $O/r02_scale8_c5.err
for c in 2 3 5; do tail -c 400 $O/r02_scale8_c$c.err; echo; done
nvidia-smi topo -m 2>/dev/null | head -12
