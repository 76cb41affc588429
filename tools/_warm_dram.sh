#!/bin/bash
# per-kernel DRAM traffic of a WARM forward (no cache flush between kernels): serpentine order off / on
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for s in 0 1; do
  UWM_SERP=$s ncu --metrics $M --cache-control none --clock-control none --graph-profiling node -c 420 --csv \
    --log-file gpurun_out/warm_dram_serp$s.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-gpu-control --no-sustained \
    > gpurun_out/warm_dram_serp$s.log 2>&1
done
tail -3 gpurun_out/warm_dram_serp1.csv | cut -c1-300
