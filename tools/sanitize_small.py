"""Small end-to-end run for compute-sanitizer: every kernel once on tiny inputs."""
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "h264-jm-commentary_b200"))
import jmme  # noqa: E402
from jmme import abi, synth  # noqa: E402

lib = jmme.load()
cur, refs = synth.frame_pair(80, 64, seed=1, search_range=8, num_refs=2)
for kw in (dict(subpel=1), dict(subpel=1, pred_policy=abi.PRED_PER_BLOCK), dict(blocktype_mask=abi.MASK_16x16),
           dict(search_mode=abi.SEARCH_FULL, pred_policy=abi.PRED_PER_BLOCK), dict(search_range=40),
           dict(search_range=32, subpel=1), dict(search_range=32, subpel=1, tuning=dict(balance=1)),     # R = 32 zero-predictor forms
           dict(search_range=32, subpel=0, tuning=dict(balance=1, group=2)),
           dict(search_mode=abi.SEARCH_FULL, pred_policy=abi.PRED_PER_BLOCK, cost_domain=1, subpel=1),
           dict(subpel=1, pred_policy=abi.PRED_MEDIAN, slice_rows=0),
           dict(subpel=1, rdopt=1, jm_center=1, max_pred_qpel=160, pred_policy=abi.PRED_PER_BLOCK, chroma_me=1)):
    R = kw.pop("search_range", 8)
    with lib.context(width=80, height=64, search_range=R, num_refs=2, **kw) as ctx:
        n = ctx.mb_w * ctx.mb_h
        pred = synth.random_pred(2, n, 41, 3, 40) if kw.get("pred_policy") in (abi.PRED_PER_BLOCK,) else None
        for i, r in enumerate(refs):
            ctx.set_reference(i, r)
            if kw.get("chroma_me"):
                ctx.set_reference_chroma(i, refs[i][::2, ::2].copy(), refs[i][1::2, ::2].copy())
        if kw.get("chroma_me"):
            ctx.set_current_chroma(cur[::2, ::2].copy(), cur[::2, 1::2].copy())
        out = ctx.search_frame(cur, pred)
        print(kw, int(out["cost"][0, 0]))
print(lib.satd(np.ones((4, 16), np.int16)))
print("sanitize run done")
