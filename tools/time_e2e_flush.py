"""The end-to-end host step with and without the L2 flush bench.py does between timed steps (per-step wall clock,
like bench.py's e2e): where do the extra microseconds of the flushed step come from?"""
import ctypes as C, pathlib, sys, time
import numpy as np, torch
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "h264-jm-commentary_b200"))
import jmme
from jmme import synth
lib = jmme.load()
w, h, R = 1920, 1080, 32
cur, refs = synth.frame_pair(w, h, 1, R)
hc, hr = torch.from_numpy(cur).pin_memory(), torch.from_numpy(refs[0]).pin_memory()
ho = torch.zeros(120 * 68 * 372, dtype=torch.uint8).pin_memory()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
small = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
pu8 = C.POINTER(C.c_uint8)
ctx = lib.context(width=w, height=h, search_range=R, subpel=1, qp=28, async_reference=1)
def step():
    lib.dll.jmme_set_reference(ctx.handle, 0, C.cast(hr.data_ptr(), pu8), w)
    lib.dll.jmme_search_frame(ctx.handle, C.cast(hc.data_ptr(), pu8), w, None, C.c_void_p(ho.data_ptr()), None)
for _ in range(5):
    step()
for mode in ("none", "fill256", "sum256", "fill1", "sleep"):
    ts = []
    for s in range(30):
        if mode == "fill256": flush.fill_(s & 255)
        elif mode == "sum256": flush.sum()
        elif mode == "fill1": small.fill_(s & 255)
        torch.cuda.synchronize()
        if mode == "sleep": time.sleep(0.002)
        t0 = time.perf_counter(); step(); ts.append(time.perf_counter() - t0)
    print(mode, "mean %.1f us  median %.1f us  min %.1f us" % (1e6 * np.mean(ts[5:]), 1e6 * np.median(ts[5:]), 1e6 * np.min(ts[5:])), flush=True)
# three frame pairs in turn (bench.py cycles seeds 1-3), separate pinned buffers
pairs = [synth.frame_pair(w, h, sd, R) for sd in (1, 2, 3)]
hp = [(torch.from_numpy(c).pin_memory(), torch.from_numpy(r[0]).pin_memory()) for c, r in pairs]
def step_i(i):
    c_, r_ = hp[i]
    lib.dll.jmme_set_reference(ctx.handle, 0, C.cast(r_.data_ptr(), pu8), w)
    lib.dll.jmme_search_frame(ctx.handle, C.cast(c_.data_ptr(), pu8), w, None, C.c_void_p(ho.data_ptr()), None)
for mode in ("same-pair", "three-pairs"):
    ts = []
    for s in range(30):
        flush.fill_(s & 255); torch.cuda.synchronize()
        t0 = time.perf_counter(); step_i(0 if mode == "same-pair" else s % 3); ts.append(time.perf_counter() - t0)
    print(mode, "mean %.1f us  median %.1f us  min %.1f us" % (1e6 * np.mean(ts[5:]), 1e6 * np.median(ts[5:]), 1e6 * np.min(ts[5:])), flush=True)
ctx.close()
