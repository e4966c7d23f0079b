"""Launch-shape sweep of the zero-predictor search in ONE process (steady clocks): whole 1080p frame and the
stripes of 2/4/8-GPU ranks, per jmme_tuning setting; device time of search only and of search + sub-pel."""
import json
import pathlib
import sys

import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "h264-jm-commentary_b200"))
import jmme  # noqa: E402
from jmme import synth  # noqa: E402
from jmme.torch_api import DeviceSearch  # noqa: E402

W, H, R = 1920, 1080, 32
TUNINGS = [dict(), dict(no_pair_tail=1), dict(group=2), dict(early_subpel=2), dict(balance=1)]
lib = jmme.load()
cur, refs = synth.frame_pair(W, H, seed=1, search_range=R)
dcur, dref = torch.from_numpy(cur).cuda(), torch.from_numpy(refs[0]).cuda()
spin = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
for _ in range(200):                                   # clocks up before anything is timed
    spin.fill_(1)
torch.cuda.synchronize()
for rows in (0, 34, 17, 9):
    for tn in TUNINGS:
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
            print(json.dumps(dict(rows=rows or 68, subpel=subpel, tuning=tn, ms=round(best, 4), kernel=s.ctx.last_kernel()[-70:])), flush=True)
            s.close()
