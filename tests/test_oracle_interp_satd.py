"""Decoder-equivalence of the oracle's quarter-pel planes (a12) against a per-sample restatement
of H.264 8.4.2.2.1, and algebraic properties of SATD (a11)."""
import numpy as np
import pytest

import refimpl
from jmme import synth


@pytest.mark.parametrize("kind,seed", [("texture", 1), ("noise", 2), ("checker", 0)])
def test_planes_match_per_sample_formulas(oracle, kind, seed):
    w, h, pad = 32, 16, 6
    img = synth.gen_luma(w, h, seed, kind)
    planes = oracle.get_sub_images_luma(img, pad)
    assert planes.shape == (4, 4, h + 2 * pad, w + 2 * pad)
    it = refimpl.Interp(img)
    rng = np.random.default_rng(seed)
    # every phase on a band crossing all four borders + random interior points
    pts = [(x, y) for y in range(-pad, h + pad) for x in (-pad, -3, -1, 0, 1, w - 2, w - 1, w, w + pad - 1)]
    pts += [(x, y) for x in range(-pad, w + pad) for y in (-pad, -2, 0, h - 1, h + pad - 1)]
    pts += [(int(rng.integers(-pad, w + pad)), int(rng.integers(-pad, h + pad))) for _ in range(300)]
    for (x, y) in pts:
        for fy in range(4):
            for fx in range(4):
                assert planes[fy, fx, y + pad, x + pad] == it.sample(4 * x + fx, 4 * y + fy), (x, y, fx, fy)


def test_integer_plane_is_edge_replication(oracle):
    img = synth.gen_luma(48, 32, 3, "noise")
    pad = 20
    planes = oracle.get_sub_images_luma(img, pad)
    assert np.array_equal(planes[0, 0], np.pad(img, pad, mode="edge"))


def test_constant_image_gives_constant_planes(oracle):
    img = np.full((16, 16), 77, np.uint8)
    assert np.all(oracle.get_sub_images_luma(img, 4) == 77)


def test_context_planes_equal_leaf_and_pad_to_16(oracle):
    # 20x18 picture -> padded to 32x32 by replication, then pad border
    img = synth.gen_luma(20, 18, 5, "texture")
    with oracle.context(width=20, height=18, search_range=4, subpel=1) as ctx:
        ctx.set_reference(0, img)
        assert (ctx.mb_w, ctx.mb_h) == (2, 2)
        ext = np.pad(img, ((0, 14), (0, 12)), mode="edge")
        leaf = oracle.get_sub_images_luma(ext, ctx.pad)
        for fy in range(4):
            for fx in range(4):
                assert np.array_equal(ctx.get_subimage(0, fx, fy), leaf[fy, fx])


def test_satd_properties(oracle):
    rng = np.random.default_rng(7)
    assert np.all(refimpl.H4 @ refimpl.H4.T == 4 * np.eye(4, dtype=np.int64))
    d = rng.integers(-255, 256, size=(500, 16)).astype(np.int16)
    d[0] = 0
    d[1] = 255
    d[2] = -255
    d[3] = np.tile([255, -255], 8)
    d[4] = np.array([[255, -255, 255, -255], [-255, 255, -255, 255]] * 2).reshape(-1)
    for rnd in (0, 1):
        got = oracle.satd(d, rnd)
        exp = [refimpl.satd4x4(x, rnd) for x in d]
        assert got.tolist() == exp
    assert oracle.satd(d[:1])[0] == 0
    # DC-only difference v: SATD = (16|v|)>>1
    for v in (1, -3, 100, 255):
        assert oracle.satd(np.full(16, v, np.int16))[0] == (16 * abs(v)) >> 1
    assert oracle.satd(d[1:2])[0] == 2040 and oracle.satd(d[4:5])[0] == 2040
