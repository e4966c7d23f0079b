"""profiles/r02_ncu.json from an `ncu --set full` capture of the bench step: per hot kernel the counters bench.py
quotes (DRAM bytes per launch, ALU-pipe and issue utilisation) — `python tools/ncu_json.py x.ncu-rep > out.json`."""
import csv
import json
import subprocess
import sys

KEYS = {"gpu__time_duration.sum": "duration_ns", "dram__bytes_read.sum": "dram_bytes_read", "dram__bytes_write.sum": "dram_bytes_write",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed": "fmaheavy_pipe_pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
        "smsp__inst_executed.sum": "warp_instructions", "launch__registers_per_thread": "registers",
        "launch__grid_size": "grid", "launch__block_size": "block",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts", "lts__t_bytes.sum": "l2_bytes"}
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e3, "ms": 1e6, "ns": 1.0, "s": 1e9}
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
res = {}
for r in rows[2:]:
    d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
    name = d["Kernel Name"]
    key = "interp" if "interp" in name else "me_subpel" if "subpel" in name else "me_int" if "me_int" in name else None
    if key is None:
        continue
    e = {"kernel": name[:160]}
    for k, short in KEYS.items():
        if k in d:
            v = float(d[k].replace(",", ""))
            e[short] = v * UNIT_SCALE.get(u[k], 1.0) if short in ("duration_ns", "dram_bytes_read", "dram_bytes_write", "l2_bytes") else v
    if key == "me_int" and key in res:            # split launch: main + tail, the main launch is the longer one
        if e.get("duration_ns", 0) < res[key].get("duration_ns", 0):
            res["me_int_tail"] = e
            continue
        res["me_int_tail"] = res[key]
    res[key] = e
json.dump(res, sys.stdout, indent=1)
