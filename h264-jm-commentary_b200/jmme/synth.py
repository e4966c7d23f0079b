"""Deterministic, integer-only synthetic YUV420 luma content (SURVEY.md §8(d) "Synthetic inputs").

Everything is derived from splitmix64 of (seed, coordinates), so the same bytes come out on every
machine.  Only luma matters to motion estimation; `yuv420_frame` adds flat chroma planes so that a
frame can be written in the YUV420 container the reference encoder reads.
"""
from __future__ import annotations

import numpy as np

KINDS = ("texture", "const", "noise", "checker", "gradient")
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x):
    """Vectorised splitmix64 finaliser on uint64 arrays."""
    with np.errstate(over="ignore"):
        x = (np.asarray(x, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15)) & _M64
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return x ^ (x >> np.uint64(31))


def _hash2d(w, h, seed, salt=0):
    y, x = np.mgrid[0:h, 0:w].astype(np.uint64)
    with np.errstate(over="ignore"):
        k = (y * np.uint64(0x100000001B3) + x) ^ (np.uint64(seed) * np.uint64(0xD6E8FEB86659FD93)
                                                  + np.uint64(salt))
    return splitmix64(k)


def gen_luma(w, h, seed=1, kind="texture"):
    """One 8-bit luma plane."""
    if kind == "const":
        return np.full((h, w), 128, np.uint8)
    if kind == "noise":
        return (_hash2d(w, h, seed) & np.uint64(255)).astype(np.uint8)
    if kind == "checker":
        y, x = np.mgrid[0:h, 0:w]
        return (((x + y) & 1) * 255).astype(np.uint8)
    if kind == "gradient":
        y, x = np.mgrid[0:h, 0:w]
        return ((x * 2 + y * 3) & 255).astype(np.uint8)
    if kind != "texture":
        raise ValueError(kind)
    y, x = np.mgrid[0:h, 0:w].astype(np.int64)
    r = splitmix64(np.arange(8, dtype=np.uint64) + np.uint64(seed) * np.uint64(1000003))
    p = [int(v % 23) + 5 for v in r[:6]]
    # three integer plaid / triangle-wave patterns plus +-8 hash noise
    tri = lambda v, per: np.abs((v % (2 * per)) - per) * 96 // per          # noqa: E731
    img = 40 + tri(x + 2 * y, p[0] * 3) + tri(3 * x - y + 7 * p[1], p[2] * 2) // 2
    img += ((x // (p[3] + 3) + y // (p[4] + 3)) & 1) * 24 + tri(x * y // 64 + p[5], 37) // 4
    img += (_hash2d(w, h, seed, 1) % np.uint64(17)).astype(np.int64) - 8
    return np.clip(img, 0, 255).astype(np.uint8)


def warp_qpel(ref, mvx, mvy):
    """cur(x,y) = bilinear sample of ref at (x + mvx/4, y + mvy/4), edges replicated.
    mvx/mvy: integer arrays (quarter-pel) broadcastable to ref.shape.  Integer arithmetic only."""
    h, w = ref.shape
    y, x = np.mgrid[0:h, 0:w]
    qx, qy = 4 * x + mvx, 4 * y + mvy
    ix, iy, fx, fy = qx >> 2, qy >> 2, qx & 3, qy & 3
    r = ref.astype(np.int64)
    g = lambda yy, xx: r[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)]      # noqa: E731
    v = ((4 - fx) * (4 - fy) * g(iy, ix) + fx * (4 - fy) * g(iy, ix + 1)
         + (4 - fx) * fy * g(iy + 1, ix) + fx * fy * g(iy + 1, ix + 1) + 8) >> 4
    return v.astype(np.uint8)


def motion_field(w, h, seed, max_qpel, tile=64):
    """Piecewise-constant motion: one MV per tile x tile region, uniform in +-max_qpel."""
    ty, tx = (h + tile - 1) // tile, (w + tile - 1) // tile
    k = splitmix64(np.arange(ty * tx * 2, dtype=np.uint64) + np.uint64(seed) * np.uint64(7919))
    mv = (k % np.uint64(2 * max_qpel + 1)).astype(np.int64) - max_qpel
    mv = mv.reshape(ty, tx, 2)
    up = lambda a: np.repeat(np.repeat(a, tile, 0), tile, 1)[:h, :w]          # noqa: E731
    return up(mv[..., 0]), up(mv[..., 1])


def frame_pair(w, h, seed=1, search_range=32, kind="texture", num_refs=1, noise=2):
    """(cur, [ref_0..ref_{n-1}]): cur is ref_0 displaced by a tiled quarter-pel motion field
    (|mv| <= 3/4 of the search range) plus +-noise; ref_k are further displaced copies."""
    base = gen_luma(w, h, seed, kind)
    refs = [base]
    for k in range(1, num_refs):
        mx, my = motion_field(w, h, seed * 131 + k, 3 * search_range)
        refs.append(warp_qpel(base, mx, my))
    mx, my = motion_field(w, h, seed * 17 + 3, 3 * search_range)
    cur = warp_qpel(base, mx, my).astype(np.int64)
    if noise and kind not in ("const", "checker"):
        cur += (_hash2d(w, h, seed, 99) % np.uint64(2 * noise + 1)).astype(np.int64) - noise
    return np.clip(cur, 0, 255).astype(np.uint8), refs


def yuv420_frame(luma):
    """Planar YUV420 bytes for one frame: luma + flat 128 chroma (unused by ME)."""
    h, w = luma.shape
    c = np.full(((h + 1) // 2) * ((w + 1) // 2), 128, np.uint8)
    return np.concatenate([luma.reshape(-1), c, c])


def random_pred(num_refs, n_mb, nb, seed, max_qpel):
    """Synthetic MV predictors, int16 [num_refs][n_mb][nb][2], uniform in +-max_qpel."""
    k = splitmix64(np.arange(num_refs * n_mb * nb * 2, dtype=np.uint64) + np.uint64(seed) * np.uint64(104729))
    return ((k % np.uint64(2 * max_qpel + 1)).astype(np.int64) - max_qpel).astype(np.int16).reshape(
        num_refs, n_mb, nb, 2)


# ---- planar YUV 4:2:0 files (the container JM's lencod reads with InputFile / SourceWidth / SourceHeight) ----
def yuv420_frame_bytes(w, h):
    return w * h + 2 * (((w + 1) // 2) * ((h + 1) // 2))


def read_yuv420_luma(path, w, h, frame=0):
    """Luma plane of frame `frame` of a planar 8-bit YUV 4:2:0 file."""
    with open(path, "rb") as f:
        f.seek(frame * yuv420_frame_bytes(w, h))
        buf = f.read(w * h)
    if len(buf) != w * h:
        raise ValueError(f"{path}: frame {frame} of {w}x{h} is not in the file")
    return np.frombuffer(buf, np.uint8).reshape(h, w).copy()


def read_yuv420(path, w, h, frame=0):
    """(luma, cb, cr) of frame `frame` of a planar 8-bit YUV 4:2:0 file."""
    cw, ch = (w + 1) // 2, (h + 1) // 2
    with open(path, "rb") as f:
        f.seek(frame * yuv420_frame_bytes(w, h))
        buf = f.read(yuv420_frame_bytes(w, h))
    if len(buf) != yuv420_frame_bytes(w, h):
        raise ValueError(f"{path}: frame {frame} of {w}x{h} is not in the file")
    a = np.frombuffer(buf, np.uint8)
    return (a[:w * h].reshape(h, w).copy(), a[w * h:w * h + cw * ch].reshape(ch, cw).copy(),
            a[w * h + cw * ch:].reshape(ch, cw).copy())


def write_yuv420(path, frames):
    """Write a planar YUV 4:2:0 sequence; a frame is a luma plane (flat chroma is written) or (luma, cb, cr)."""
    with open(path, "wb") as f:
        for fr in frames:
            if isinstance(fr, (tuple, list)):
                for pl in fr:
                    f.write(np.ascontiguousarray(pl, np.uint8).tobytes())
            else:
                f.write(yuv420_frame(np.ascontiguousarray(fr, np.uint8)).tobytes())
