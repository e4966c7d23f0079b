"""Closed per-frame loop at 1080p: pass 1 (zero predictors) -> commit -> median predictors -> pass 2 (PER_BLOCK)."""
import pathlib
import sys

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "h264-jm-commentary_b200"))
import jmme  # noqa: E402
from jmme import abi, synth  # noqa: E402
from jmme.torch_api import DeviceSearch  # noqa: E402

lib = jmme.load()
w, h, R = 1920, 1080, 32
cur, refs = synth.frame_pair(w, h, 1, R)
with lib.context(width=w, height=h, search_range=R, subpel=1, qp=28) as c:
    c.set_reference(0, refs[0])
    res1 = c.search_frame(cur)
    mv4, ref4, mode = c.commit_field(res1)
    pred = c.predict_frame(mv4, ref4)
uniform = np.all(pred[0] == pred[0][:, :1], axis=(1, 2)).mean()
print(f"MBs whose 41 median predictors are all equal: {100 * uniform:.1f} %; modes 1/2/3/8: "
      f"{[int((mode[:, 0] == m).sum()) for m in (1, 2, 3, 8)]}")
dcur, dref = torch.from_numpy(cur).cuda(), torch.from_numpy(refs[0]).cuda()
dpred = torch.from_numpy(pred).cuda()
for name, kw, p in (("pass 1 zero predictors", dict(), None), ("pass 2 median predictors", dict(pred_policy=abi.PRED_PER_BLOCK), dpred),
                    ("pass 2, rdopt lambda (rate from the table)", dict(pred_policy=abi.PRED_PER_BLOCK, rdopt=1), dpred),
                    ("pass 2, integer only", dict(pred_policy=abi.PRED_PER_BLOCK, subpel=0), dpred),
                    ("FullPelBlockMotionSearch: a window per block (me_full.cu), integer only",
                     dict(pred_policy=abi.PRED_PER_BLOCK, search_mode=abi.SEARCH_FULL, subpel=0), dpred)):
    s = DeviceSearch(lib, width=w, height=h, search_range=R, qp=28, **dict(dict(subpel=1), **kw))
    s.set_reference(0, dref)
    for _ in range(3):
        s.search(dcur, p)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        s.search(dcur, p)
    e1.record()
    torch.cuda.synchronize()
    out = s.to_numpy(s.out)
    print(f"{name}: {e0.elapsed_time(e1) / 20:.4f} ms, total 16x16 cost {int(out['cost'][:, 0].astype(np.int64).sum())}")
    s.close()
