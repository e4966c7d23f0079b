#!/bin/bash
# full GPU check of a revision: tests, smoke, microbenchmark, bench N=1 and the CPU arm (no ncu: see gpu_prof.sh)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader | head -1; nproc
python -m pytest tests -q -m gpu --timeout=900 -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
h264-jm-commentary_b200/csrc/microbench 150 > gpurun_out/INT_PEAKS.json 2> gpurun_out/microbench.err; tail -2 gpurun_out/microbench.err
( time python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -3 gpurun_out/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python - <<'PY'
import json
for f in ("bench_n1", "bench_ref"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "e2e", "kernel_ms", "parity", "clocks")}, d.get("roofline", {}).get("frac"), d.get("roofline", {}).get("peak"))
    except Exception as e:
        print(f, "unreadable", e)
PY
