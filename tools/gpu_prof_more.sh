#!/bin/bash
# ncu --set full of the R = 64 search kernel (config 4) and of one wavefront step (search + sub-pel kernels of the median policy)
mkdir -p gpurun_out
python bench.py --workload 1080p_r64_41blk_int_4ref --steps 1 --warmup 3 --no-cpu --no-parity --no-graph --no-extras > gpurun_out/plain_r64.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'me_int_kernel' -s 2 -c 1 -o gpurun_out/prof_r02_r64 -f \
    python bench.py --workload 1080p_r64_41blk_int_4ref --steps 1 --warmup 3 --no-cpu --no-parity --no-graph --no-extras > gpurun_out/ncu_r64.log 2>&1
tail -2 gpurun_out/ncu_r64.log
python tools/time_median.py --slices 1 --iters 1 --graph 0 > gpurun_out/plain_median.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'me_int_tb_kernel|me_subpel_kernel' -s 300 -c 2 -o gpurun_out/prof_r02_wave -f \
    python tools/time_median.py --slices 1 --iters 1 --graph 0 > gpurun_out/ncu_wave.log 2>&1
tail -2 gpurun_out/ncu_wave.log
ls -la gpurun_out/*.ncu-rep
