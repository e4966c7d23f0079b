#!/bin/bash
# microbenchmark + ncu launch list + ncu full capture (with source) of the hot kernels of the bench step
mkdir -p gpurun_out
h264-jm-commentary_b200/csrc/microbench 150 > gpurun_out/INT_PEAKS.json 2> gpurun_out/microbench.err; tail -2 gpurun_out/microbench.err
python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --no-graph --no-extras > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --no-graph --no-extras > gpurun_out/ncu_launches.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --no-graph --no-extras > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'me_int|me_subpel|interp_kernel' -s 6 -c 3 \
    -o gpurun_out/prof_r02 -f python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --no-graph --no-extras > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out | head -30
