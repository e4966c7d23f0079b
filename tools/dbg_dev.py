"""debug: device-resident path (one launch over the frame, nothing else in the stream) vs host path vs oracle"""
import pathlib, sys
import numpy as np, torch
ROOT = pathlib.Path(sys.argv[1]).resolve()
sys.path.insert(0, str(ROOT / "h264-jm-commentary_b200")); sys.path.insert(0, str(ROOT / "oracle"))
import jmme, oracle as om
from jmme import synth
from jmme.torch_api import DeviceSearch
lib, orc = jmme.load(), om.load()
orc.dll.jmme_oracle_set_threads(0)
w, h, R = 1920, 1080, 32
cur, refs = synth.frame_pair(w, h, seed=1, search_range=R)
with orc.context(width=w, height=h, search_range=R, qp=28, subpel=0) as c:
    c.set_reference(0, refs[0]); o = c.search_frame(cur)
dcur, dref = torch.from_numpy(cur).cuda(), torch.from_numpy(refs[0]).cuda()
for tn in (dict(), dict(balance=2), dict(group=1, balance=2)):
    ds = DeviceSearch(lib, width=w, height=h, search_range=R, qp=28, subpel=0, tuning=tn)
    ds.set_reference(0, dref); torch.cuda.synchronize()
    for rep in range(2):
        out = ds.search(dcur); torch.cuda.synchronize()
        g = ds.to_numpy(out)
        bad = np.nonzero(np.any(g["cost"] != o["cost"], axis=1))[0]
        print(ROOT.name, "dev", tn, rep, "bad", len(bad), bad[:4], bad[-3:] if len(bad) else "", ds.ctx.last_kernel()[-50:], flush=True)
    ds.close()
