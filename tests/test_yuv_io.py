"""Planar YUV 4:2:0 container round trip (the file format either side of the path)."""
import numpy as np

from jmme import synth


def test_yuv420_round_trip(tmp_path):
    w, h = 52, 38                                            # odd chroma sizes
    frames = [synth.gen_luma(w, h, s, "texture") for s in (1, 2, 3)]
    p = tmp_path / "seq.yuv"
    synth.write_yuv420(p, frames)
    assert p.stat().st_size == 3 * synth.yuv420_frame_bytes(w, h) == 3 * (w * h + 2 * 26 * 19)
    for i, f in enumerate(frames):
        assert np.array_equal(synth.read_yuv420_luma(p, w, h, i), f)
    try:
        synth.read_yuv420_luma(p, w, h, 3)
        raise AssertionError("reading past the end must fail")
    except ValueError:
        pass
