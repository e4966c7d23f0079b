"""Where the end-to-end host-buffer step spends its time (copies vs kernels vs call overhead)."""
import ctypes as C
import pathlib
import sys
import time

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "h264-jm-commentary_b200"))
import jmme  # noqa: E402
from jmme import abi, synth  # noqa: E402

lib = jmme.load()
w, h, R = 1920, 1080, 32
cur, refs = synth.frame_pair(w, h, 1, R)
hc = torch.from_numpy(cur).pin_memory()
hr = torch.from_numpy(refs[0]).pin_memory()
n_mb = 120 * 68
ho = torch.zeros(n_mb * 372, dtype=torch.uint8).pin_memory()
dbuf = torch.empty(8 << 20, dtype=torch.uint8, device="cuda")


def t(f, n=50):
    for _ in range(5):
        f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


print("H2D 2.07 MB pinned  %.1f us" % t(lambda: (dbuf[: w * h].copy_(hc.view(-1), non_blocking=True), torch.cuda.synchronize())))
print("D2H 3.04 MB pinned  %.1f us" % t(lambda: (ho.copy_(dbuf[: n_mb * 372], non_blocking=True), torch.cuda.synchronize())))
print("empty sync          %.1f us" % t(lambda: torch.cuda.synchronize()))
pu8 = C.POINTER(C.c_uint8)
for asyncref, ng, tn in ((0, 1, {}), (1, 1, {}), (1, 1, dict(early_subpel=2)), (1, 1, dict(even_parts=1)), (1, 1, dict(balance=1)),
                         (1, 1, dict(pipe_parts=4)), (1, 1, dict(pipe_parts=2)), (1, 1, dict(pipe_parts=1)), (1, 1, {})):
    ctx = lib.context(width=w, height=h, search_range=R, subpel=1, qp=28, async_reference=asyncref, n_gpus=ng,
                      device_ids=[0] * ng, tuning=tn)

    def setref():
        lib.dll.jmme_set_reference(ctx.handle, 0, C.cast(hr.data_ptr(), pu8), w)

    def search():
        lib.dll.jmme_search_frame(ctx.handle, C.cast(hc.data_ptr(), pu8), w, None, C.c_void_p(ho.data_ptr()), None)

    def step():
        setref()
        search()
    setref()
    print(f"async_reference={asyncref} stripes={ng} {tn}: set_reference %.1f us, search_frame %.1f us, step %.1f us" % (t(setref), t(search), t(step)))
    ctx.close()
