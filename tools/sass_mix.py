"""Static SASS opcode mix of the kernels of an object file whose demangled name matches a regex:
python tools/sass_mix.py h264-jm-commentary_b200/csrc/me_int_tb.o 'me_int_tb_kernel<6, 4, 3, false, 126' [top]"""
import re
import subprocess
import sys
from collections import Counter

obj, pat = sys.argv[1], re.compile(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 16
elf = subprocess.run(["cuobjdump", "-elf", obj], capture_output=True, text=True).stdout
syms = sorted(set(re.findall(r"\.text\.(_Z\w+)", elf)))
for s in syms:
    name = subprocess.run(["c++filt", s], capture_output=True, text=True).stdout.strip().replace("(anonymous namespace)::", "")
    if not pat.search(name):
        continue
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", s, obj], capture_output=True, text=True).stdout
    ops = Counter()
    n = 0
    for line in sass.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
        if m:
            ops[m.group(2)] += 1
            n += 1
    print(f"== {name}: {n} instructions")
    print("   " + ", ".join(f"{k} {v}" for k, v in ops.most_common(top)))
