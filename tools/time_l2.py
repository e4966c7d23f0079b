"""Per-kernel device times of set_reference + search with and without the 256 MB L2 flush bench.py does between steps
(is the sub-pel kernel reading its planes from L2 or from HBM?), whole 1080p frame and a 9-row stripe."""
import json, pathlib, sys
import torch
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "h264-jm-commentary_b200"))
import jmme
from jmme import synth
from jmme.torch_api import DeviceSearch
W, H, R = 1920, 1080, 32
lib = jmme.load()
cur, refs = synth.frame_pair(W, H, seed=1, search_range=R)
dcur, dref = torch.from_numpy(cur).cuda(), torch.from_numpy(refs[0]).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
prop = torch.cuda.get_device_properties(0)
print(json.dumps({"l2_bytes": prop.L2_cache_size, "name": prop.name}))
for rows in (0, 9):
    for mode in ("noflush", "flush", "flush-read"):
        s = DeviceSearch(lib, width=W, height=H, search_range=R, subpel=1, qp=28, mb_row_end=rows)
        s.ctx.set_profiling(True)
        acc = []
        for it in range(25):
            if mode == "flush":
                flush.fill_(it & 255)
            elif mode == "flush-read":
                flush.sum()                      # evicts with clean lines instead of dirty ones
            s.set_reference(0, dref)
            s.search(dcur)
            torch.cuda.synchronize()
            if it >= 5:
                acc.append(s.ctx.kernel_times())
        avg = {k: round(sum(a[k] for a in acc) / len(acc), 4) for k in acc[0]}
        print(json.dumps({"rows": rows or 68, "mode": mode, **avg}), flush=True)
        s.close()
