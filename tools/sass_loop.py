"""Static opcode histogram of the innermost loop that contains VABSDIFF4 in a kernel's SASS."""
import re
import subprocess
import sys
from collections import Counter

obj, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
on, ins = False, []
for line in txt.splitlines():
    if "Function :" in line:
        on = pat in line
    elif on:
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
addr = {a: i for i, (a, _) in enumerate(ins)}
# backward branches = loops; pick the smallest loop containing >= 32 VABSDIFF4
best = None
for i, (a, s) in enumerate(ins):
    m = re.search(r"BRA(?:\.\w+)* .*?(0x[0-9a-f]+)", s)
    if m:
        t = int(m.group(1), 16)
        if t in addr and addr[t] < i:
            body = ins[addr[t]:i + 1]
            n = sum("VABSDIFF4" in x for _, x in body)
            if n >= 32 and (best is None or len(body) < len(best)):
                best = body
if best is None:
    sys.exit("no loop found")
c = Counter()
for _, s in best:
    t = s.split()
    op = t[1] if t[0].startswith("@") else t[0]
    if op.startswith("IMAD.U32") and "RZ, RZ, UR" in s:
        op = "IMAD.U32(UR->R)"
    c[op] += 1
print(f"{pat}: loop of {len(best)} instr:", "; ".join(f"{o} {n}" for o, n in c.most_common(24)))
