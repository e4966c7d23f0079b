"""Summarise an `ncu --page source --csv` export: executed warp-instructions and stall samples by opcode."""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ex, smp, shw, shi = Counter(), Counter(), Counter(), Counter()
tot = 0
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    src = r[ix["Source"]].strip()
    toks = src.split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0] if not op.startswith(("LDS", "STS", "LDG", "IMAD", "ATOMS")) else ".".join(op.split(".")[:2])
    n = int(r[ix["Instructions Executed"]])
    ex[op] += n
    tot += n
    smp[op] += int(r[ix["# Samples"]])
    shw[op] += int(r[ix["L1 Wavefronts Shared"]])
    shi[op] += int(r[ix["L1 Wavefronts Shared Ideal"]])
print(f"total warp-instructions executed: {tot}")
ts = sum(smp.values())
for op, n in ex.most_common(28):
    extra = f"  smem wavefronts {shw[op]} (ideal {shi[op]})" if shw[op] else ""
    print(f"{op:16s} {n:12d} {100.0 * n / tot:6.2f}%   samples {100.0 * smp[op] / max(ts, 1):5.1f}%{extra}")
