"""JM encoder.cfg -> jmme_params, and the IPPP pass over a YUV file (host logic; the search runs on the oracle
here and on the GPU in test_gpu_parity.py)."""
import numpy as np
import pytest

from jmme import abi, synth
from jmme.cfg import EncoderCfg, parse_encoder_cfg
from jmme.sequence import search_sequence, yuv_frames

CFG = """
# New Input File Format is as follows
# <ParameterName> = <ParameterValue> # Comment
InputFile             = "clip.yuv"       # Input sequence
FramesToBeEncoded     = 3   # Number of frames to be coded
SourceWidth           = 64  # Frame width
SourceHeight          = 48  # Frame height
QPPSlice              = 30
SearchRange           = 6       # Max search range
NumberReferenceFrames = 2
InterSearch16x16      = 1
InterSearch8x4        = 0   # off
InterSearch4x8        = 0
UseHadamard           = 1
RDOptimization        = 1
SearchMode            = -1      # full search
SliceMode             = 1
SliceArgument         = 8       # two MB rows of 4 MBs
free text without an equals sign is ignored
"""


def write_cfg(tmp_path, text=CFG):
    p = tmp_path / "encoder.cfg"
    p.write_text(text)
    return p


def test_parse_and_map(tmp_path):
    raw = parse_encoder_cfg(CFG)
    assert raw["InputFile"] == "clip.yuv" and raw["SearchRange"] == "6" and "free" not in raw
    cfg = EncoderCfg.load(write_cfg(tmp_path), ["SearchRange=9", "InterSearch4x4 = 0"])
    kw = cfg.params()
    assert kw == dict(width=64, height=48, search_range=9, num_refs=2, blocktype_mask=0x1E, qp=30, rdopt=1,
                      use_hadamard=1, subpel=1, search_mode=abi.SEARCH_FULL, slice_rows=2)
    assert cfg.frames == 3 and cfg.input_file == "clip.yuv"


@pytest.mark.parametrize("override,msg", [("SearchMode=1", "UMHex"), ("UseFME=1", "UMHexagonS"),
                                          ("MEDistortionQPel=3", "MEDistortion"), ("SourceWidth=0", "SourceWidth"),
                                          ("QPPSlice=abc", "integer"), ("MEDistortionFPel=2", "integer-pel"),
                                          ("ChromaMEEnable=2", "ChromaMEEnable")])
def test_unsupported_keys_are_refused(tmp_path, override, msg):
    with pytest.raises(ValueError, match=msg):
        EncoderCfg.load(write_cfg(tmp_path), [override]).params()


def test_distortion_keys_of_newer_jm(tmp_path):
    kw = EncoderCfg.load(write_cfg(tmp_path), ["MEDistortionHPel=0", "MEDistortionQPel=0", "DisableSubpelME=0",
                                               "SearchMode=0"]).params()
    assert kw["use_hadamard"] == 0 and kw["subpel"] == 1 and kw["search_mode"] == abi.SEARCH_FASTFULL


def test_jm12_distortion_keys(tmp_path):
    """MEDistortionFPel/HPel/QPel (0 SAD, 1 SSE, 2 Hadamard), Transform8x8Mode, ChromaMEEnable; the cost domain is a
    compile-time switch of JM and therefore an argument."""
    kw = EncoderCfg.load(write_cfg(tmp_path), ["MEDistortionFPel=1", "MEDistortionQPel=1", "Transform8x8Mode=1",
                                               "ChromaMEEnable=1"]).params(cost_domain=1)
    assert (kw["me_distortion"], kw["me_distortion_fpel"], kw["me_distortion_hpel"], kw["me_distortion_qpel"]) == (1, 1, 2, 1)
    assert kw["transform8x8"] == 1 and kw["chroma_me"] == 1 and kw["cost_domain"] == 1 and kw["use_hadamard"] == 1
    kw = EncoderCfg.load(write_cfg(tmp_path), []).params()
    assert "me_distortion" not in kw and "cost_domain" not in kw and "chroma_me" not in kw
    with pytest.raises(ValueError, match="sub-pel"):
        EncoderCfg.load(write_cfg(tmp_path), ["ChromaMEEnable=1", "DisableSubpelME=1"]).params()


def test_sequence_pass_with_chroma_me(oracle, tmp_path):
    """A YUV file with real chroma planes through the IPPP pass with ChromaMEEnable = 1: equals frame-by-frame calls."""
    w, h, R = 64, 48, 4
    frames = []
    for k in range(3):
        frames.append((synth.gen_luma(w, h, 3 + k), synth.gen_luma(w // 2, h // 2, 30 + k), synth.gen_luma(w // 2, h // 2, 40 + k)))
    synth.write_yuv420(tmp_path / "clip.yuv", frames)
    back = synth.read_yuv420(tmp_path / "clip.yuv", w, h, 2)
    assert all(np.array_equal(a, b) for a, b in zip(back, frames[2]))
    assert np.array_equal(synth.read_yuv420_luma(tmp_path / "clip.yuv", w, h, 1), frames[1][0])
    kw = dict(search_range=R, subpel=1, chroma_me=1, qp=30)
    got = list(search_sequence(oracle, yuv_frames(tmp_path / "clip.yuv", w, h, 3, chroma=True), **kw))
    assert [n for n, _, _ in got] == [1, 2]
    for n, rec, _ in got:
        with oracle.context(width=w, height=h, **kw) as ctx:
            ctx.set_reference(0, frames[n - 1][0])
            ctx.set_reference_chroma(0, *frames[n - 1][1:])
            ctx.set_current_chroma(*frames[n][1:])
            assert ctx.search_frame(frames[n][0]).tobytes() == rec.tobytes()
    with pytest.raises(ValueError, match="chroma_me"):
        list(search_sequence(oracle, yuv_frames(tmp_path / "clip.yuv", w, h, 3), **kw))


def test_sequence_pass_matches_frame_by_frame_calls(oracle, tmp_path):
    w, h, R = 64, 48, 6
    lumas = [synth.gen_luma(w, h, 3)]
    for k in range(3):
        lumas.append(synth.frame_pair(w, h, seed=3, search_range=R + k)[0])
    synth.write_yuv420(tmp_path / "clip.yuv", lumas)
    cfg = EncoderCfg.load(write_cfg(tmp_path), ["FramesToBeEncoded=4"])
    kw = cfg.params()
    kw.pop("slice_rows")
    assert kw.pop("width") == w and kw.pop("height") == h
    got = list(search_sequence(oracle, yuv_frames(tmp_path / cfg.input_file, w, h, cfg.frames), **kw))
    assert [n for n, _, _ in got] == [1, 2, 3]
    for n, rec, _ in got:
        with oracle.context(width=w, height=h, **kw) as ctx:
            ctx.set_reference(0, lumas[n - 1])
            ctx.set_reference(1, lumas[max(n - 2, 0)])
            assert ctx.search_frame(lumas[n]).tobytes() == rec.tobytes()
    # in-frame median over the same file, slices from the configuration
    med = list(search_sequence(oracle, yuv_frames(tmp_path / "clip.yuv", w, h, 2), abi.PRED_MEDIAN,
                               **dict(kw, slice_rows=cfg.params()["slice_rows"])))
    assert len(med) == 1 and med[0][1].shape == (12,)
    with pytest.raises(ValueError):
        list(yuv_frames(tmp_path / "clip.yuv", w, h, 6))           # the file holds 4 frames
