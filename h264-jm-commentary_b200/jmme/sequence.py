"""ME-only pass over a YUV 4:2:0 sequence, the way lencod walks its input (IPPP).

Frame 0 is the I picture (no search); every later frame is searched against the previous `num_refs` frames,
reference 0 being the nearest.  There is no transform / reconstruction path in this library (out of scope),
so the *original* earlier frames stand in for the reconstructed ones.
"""
from __future__ import annotations

import time

import numpy as np

from . import abi, synth


def search_sequence(lib: abi.Lib, frames, pred_policy=abi.PRED_ZERO, **params):
    """frames: iterable of 2-D uint8 luma arrays, or of (luma, cb, cr) triples (needed with chroma_me = 1).
    Yields (frame_no, records, seconds) for frames 1..n-1; records = jmme_mbresult array of the whole frame.
    With fewer decoded frames than `num_refs` the oldest available frame is repeated (JM shortens the list
    instead; costs of the duplicated references never win because the lowest reference index wins ties)."""
    chroma = bool(params.get("chroma_me"))

    def split(fr):
        if isinstance(fr, (tuple, list)):
            return np.ascontiguousarray(fr[0], np.uint8), (fr[1], fr[2])
        if chroma:
            raise ValueError("chroma_me = 1 needs (luma, cb, cr) frames")
        return np.ascontiguousarray(fr, np.uint8), None

    frames = iter(frames)
    first, first_c = split(next(frames))
    h, w = first.shape
    params = dict(params, width=w, height=h, pred_policy=pred_policy)
    n_refs = params.setdefault("num_refs", 1)
    if pred_policy not in (abi.PRED_ZERO, abi.PRED_MEDIAN):
        raise ValueError("a sequence is searched with zero or in-frame median predictors")
    history = [(first, first_c)]
    with lib.context(**params) as ctx:
        for n, fr in enumerate(frames, start=1):
            cur, cur_c = split(fr)
            if cur.shape != (h, w):
                raise ValueError(f"frame {n} is {cur.shape}, expected {(h, w)}")
            t0 = time.perf_counter()
            for r in range(n_refs):
                ref, ref_c = history[min(r, len(history) - 1)]
                ctx.set_reference(r, ref)
                if chroma:
                    ctx.set_reference_chroma(r, *ref_c)
            if chroma:
                ctx.set_current_chroma(*cur_c)
            rec = ctx.search_frame(cur)
            yield n, rec, time.perf_counter() - t0
            history.insert(0, (cur, cur_c))
            del history[n_refs:]


def yuv_frames(path, w, h, count=None, start=0, chroma=False):
    """Luma planes — or (luma, cb, cr) triples with chroma=True — of a planar 8-bit YUV 4:2:0 file."""
    n = start
    while count is None or n < start + count:
        try:
            yield synth.read_yuv420(path, w, h, n) if chroma else synth.read_yuv420_luma(path, w, h, n)
        except ValueError:
            if count is None:
                return
            raise
        n += 1
