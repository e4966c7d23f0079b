"""Integer-search kernel shapes on the whole 1080p frame (zero predictors, balanced ranges): device time of the search
alone, one process, steady clocks."""
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
base = None
for tn in [dict(), dict(variant=65), dict(variant=64), dict(variant=88), dict(variant=48), dict(variant=66), dict(variant=65, group=2),
           dict(variant=66, group=2), dict(variant=88, group=2), dict(variant=65, balance=2), dict(variant=66, balance=2)]:
    try:
        s = DeviceSearch(lib, width=W, height=H, search_range=R, subpel=0, qp=28, tuning=tn)
        s.set_reference(0, dref)
        for _ in range(30):
            out = s.search(dcur)
        torch.cuda.synchronize()
        res = s.to_numpy(out)["cost"].copy()
        if base is None:
            base = res
        best = 1e9
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(40):
                s.search(dcur)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 40)
        print(json.dumps(dict(tuning=tn, ms=round(best, 4), same=bool((res == base).all()), kernel=s.ctx.last_kernel()[16:])), flush=True)
        s.close()
    except Exception as e:
        print(json.dumps(dict(tuning=tn, error=str(e)[:100])), flush=True)
