"""In-frame median prediction (JMME_PRED_MEDIAN) at 1080p: device time per frame for several slice heights."""
import argparse
import pathlib
import sys

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "h264-jm-commentary_b200"))
import jmme  # noqa: E402
from jmme import abi, synth  # noqa: E402
from jmme.torch_api import DeviceSearch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--slices", default="1,2,4,0")
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--refs", type=int, default=1)
ap.add_argument("--graph", type=int, default=1)
ap.add_argument("--subpel", type=int, default=1)
a = ap.parse_args()
lib = jmme.load()
w, h, R = 1920, 1080, 32
cur, refs = synth.frame_pair(w, h, 1, R, num_refs=a.refs)
dcur = torch.from_numpy(cur).cuda()
drefs = [torch.from_numpy(r).cuda() for r in refs]
for k in [int(x) for x in a.slices.split(",")]:
    s = DeviceSearch(lib, width=w, height=h, search_range=R, subpel=a.subpel, qp=28, num_refs=a.refs,
                     pred_policy=abi.PRED_MEDIAN, slice_rows=k)
    for i, d in enumerate(drefs):
        s.set_reference(i, d)
    n0 = s.launch_count()
    s.search(dcur)
    launches = s.launch_count() - n0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        e0.record()
        for _ in range(a.iters):
            s.search(dcur)
        e1.record()
    torch.cuda.synchronize()
    eager = e0.elapsed_time(e1) / a.iters
    graph_ms = float("nan")
    if a.graph:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            s.search(dcur)
        g.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(a.iters):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        graph_ms = e0.elapsed_time(e1) / a.iters
    out = s.to_numpy(s.out)
    print(f"slice_rows={k}: eager {eager:.3f} ms, graph {graph_ms:.3f} ms per frame ({launches} launches), "
          f"{8160 / (min(eager, graph_ms) if a.graph else eager) / 1e3:.2f} M MB/s, total 16x16 cost "
          f"{int(out['cost'][:, 0].astype(np.int64).sum())}", flush=True)
    s.close()
