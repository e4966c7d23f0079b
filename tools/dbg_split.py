"""debug: 1080p zero-predictor search against the oracle; argv[1] = repo root of the revision to test"""
import pathlib, sys
import numpy as np
ROOT = pathlib.Path(sys.argv[1] if len(sys.argv) > 1 else pathlib.Path(__file__).resolve().parent.parent).resolve()
sys.path.insert(0, str(ROOT / "h264-jm-commentary_b200")); sys.path.insert(0, str(ROOT / "oracle"))
import jmme, oracle as om
from jmme import synth
lib, orc = jmme.load(), om.load()
orc.dll.jmme_oracle_set_threads(0)
R = 32
def run(L, cur, refs, w, h, tuning=None, **kw):
    with L.context(width=w, height=h, search_range=R, qp=28, **kw) as c:
        if tuning: c.set_tuning(**tuning)
        c.set_reference(0, refs[0])
        out = c.search_frame(cur)
        return out, (c.last_kernel() if (L is lib and hasattr(c, "last_kernel")) else "")
for (w, h, seed) in ((1920, 1080, 1), (1280, 720, 1), (1280, 720, 6)):
    cur, refs = synth.frame_pair(w, h, seed=seed, search_range=R)
    for subpel in (0,):
        o, _ = run(orc, cur, refs, w, h, subpel=subpel)
        tns = [dict()] + ([dict(balance=2), dict(group=1)] if len(sys.argv) <= 1 else [])
        for tn in tns:
            g, k = run(lib, cur, refs, w, h, tn, subpel=subpel)
            bad = np.nonzero(np.any(g["cost"] != o["cost"], axis=1))[0]
            print(ROOT.name, (w, h, seed), subpel, tn, "bad MBs", len(bad), bad[:6], bad[-3:] if len(bad) else "", k[-60:], flush=True)
            if len(bad) and tn == {}:
                m = bad[0]
                for b in (0, 1, 5, 9, 25, 40):
                    print("   mb", m, "blk", b, "gpu", g["mv"][m, b], g["cost"][m, b], "oracle", o["mv"][m, b], o["cost"][m, b])
                nb = np.array([np.count_nonzero(g["cost"][bad, b] != o["cost"][bad, b]) for b in range(41)])
                print("   bad per block", nb)
