#!/bin/bash
# ncu --set full of the integer kernel for the variants in $1 (comma list)
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x --timeout=900 2>&1 | tail -3
for v in ${1//,/ }; do
python tools/time_search.py --subpel 0 --iters 3 --Ks $v > gpurun_out/plain_$v.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:me_int -s 3 -c 1 -o gpurun_out/int_$v -f \
    python tools/time_search.py --subpel 0 --iters 3 --Ks $v > gpurun_out/ncu_$v.log 2>&1
tail -2 gpurun_out/ncu_$v.log
done
