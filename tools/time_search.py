"""Quick device-resident timing of the search at a BASELINE config for several K (tuning aid)."""
import argparse
import json
import pathlib
import sys

import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "h264-jm-commentary_b200"))
import jmme  # noqa: E402
from jmme import synth  # noqa: E402
from jmme.torch_api import DeviceSearch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--w", type=int, default=1920)
ap.add_argument("--h", type=int, default=1080)
ap.add_argument("--R", type=int, default=32)
ap.add_argument("--refs", type=int, default=1)
ap.add_argument("--subpel", type=int, default=1)
ap.add_argument("--mask", type=lambda s: int(s, 0), default=0xFE)
ap.add_argument("--pred", type=int, default=0)
ap.add_argument("--Ks", default="0,68,48,88,32,51")
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--group", type=int, default=0, help="jmme_tuning.group")
ap.add_argument("--tuning", default="", help="more jmme_tuning fields: key=val,key=val")
ap.add_argument("--rows", type=int, default=0, help="search only the first N MB rows (stripe)")
a = ap.parse_args()

lib = jmme.load()
cur, refs = synth.frame_pair(a.w, a.h, seed=1, search_range=a.R, num_refs=a.refs)
dcur = torch.from_numpy(cur).cuda()
drefs = [torch.from_numpy(r).cuda() for r in refs]
for K in [int(k) for k in a.Ks.split(",")]:
    for subpel in sorted({0, a.subpel}):
        s = DeviceSearch(lib, width=a.w, height=a.h, search_range=a.R, num_refs=a.refs, subpel=subpel,
                         blocktype_mask=a.mask, pred_policy=a.pred, qp=28, mb_row_end=a.rows,
                         tuning={**dict(variant=K, group=a.group), **{k: int(v) for k, v in (kv.split('=') for kv in a.tuning.split(',') if kv)}})
        pred = None
        if a.pred:
            nb = 1 if a.pred == 1 else 41
            pred = torch.from_numpy(synth.random_pred(a.refs, s.n_mb, nb, 3, 4 * a.R)).cuda()
        for i, r in enumerate(drefs):
            s.set_reference(i, r)
        for _ in range(3):
            s.search(dcur, pred)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            s.search(dcur, pred)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        e0.record()
        for _ in range(a.iters):
            s.set_reference(0, drefs[0])
        e1.record()
        torch.cuda.synchronize()
        ms_ref = e0.elapsed_time(e1) / a.iters
        print(json.dumps(dict(K=K, subpel=subpel, w=a.w, h=a.h, R=a.R, refs=a.refs, mask=a.mask, pred=a.pred,
                              ms_search=round(ms, 4), kernel=s.ctx.last_kernel(), ms_set_reference=round(ms_ref, 4),
                              rows=a.rows, mb_per_s=round((s.n_mb if not a.rows else a.rows * s.ctx.mb_w) / (ms * 1e-3)))), flush=True)
        s.close()
