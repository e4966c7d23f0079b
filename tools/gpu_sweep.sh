#!/bin/bash
# launch-shape sweep of the zero-predictor integer search: whole 1080p frame and the stripes of 2/4/8-GPU ranks
mkdir -p gpurun_out
for rows in 0 34 17 9; do
  for t in "no_split=1" "" "no_split_pdl=1" "group=4" "group=4,no_split_pdl=1" "group=4,no_split=1"; do
    python tools/time_search.py --Ks 0 --rows $rows --iters 20 --tuning "$t" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()); continue
    print('rows', d['rows'], 'subpel', d['subpel'], 'tuning', '$t', 'ms', d['ms_search'], 'setref', d['ms_set_reference'], d['kernel'][-60:])
"
  done
done | tee gpurun_out/sweep.txt
