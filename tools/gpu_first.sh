#!/bin/bash
# first GPU call: microbenchmarks, parity tests, K sweep
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
./h264-jm-commentary_b200/csrc/microbench > gpurun_out/int_peaks.json 2> gpurun_out/microbench.err
cat gpurun_out/int_peaks.json
python -m pytest tests -q -m gpu -x --timeout=900 > gpurun_out/pytest_gpu.log 2>&1
tail -30 gpurun_out/pytest_gpu.log
python tools/time_search.py > gpurun_out/time_1080p.log 2>&1
cat gpurun_out/time_1080p.log
