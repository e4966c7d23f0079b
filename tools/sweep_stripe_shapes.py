"""Short stripes (the ranks of an 8- or 4-GPU run): integer-search CTA shapes and item sizes, search and search + sub-pel."""
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
spin = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
for _ in range(200):
    spin.fill_(1)
torch.cuda.synchronize()
for rows in (9, 17):
    for tn in [dict(), dict(variant=65), dict(variant=65, group=1), dict(variant=65, group=4), dict(variant=64), dict(variant=64, group=1),
               dict(variant=64, group=4), dict(group=1), dict(variant=48), dict(variant=66)]:
        r = {}
        for subpel in (0, 1):
            s = DeviceSearch(lib, width=W, height=H, search_range=R, subpel=subpel, qp=28, mb_row_end=rows, tuning=tn)
            s.set_reference(0, dref)
            for _ in range(30):
                s.search(dcur)
            torch.cuda.synchronize()
            best = 1e9
            for rep in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(40):
                    s.search(dcur)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / 40)
            r[subpel] = round(best, 4); k = s.ctx.last_kernel()[16:60]
            s.close()
        print(json.dumps(dict(rows=rows, tuning=tn, int_ms=r[0], total_ms=r[1], kernel=k)), flush=True)
