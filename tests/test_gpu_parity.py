"""GPU parity: libjmme_cuda.so (through the C ABI) against the CPU oracle on the same seeded inputs.
Bit-exact: MV, ref_idx and cost of all 41 blocks of every macroblock; every byte of the 16 planes."""
import ctypes

import numpy as np
import pytest

import refimpl
from jmme import abi, synth

pytestmark = pytest.mark.gpu
BLOCKS = abi.block_table()


LAST = {}          # what the last CUDA run() launched: kernel name (jmme_last_kernel) and tuning in effect


def run(lib, cur, refs, pred=None, per_ref=False, tuning=None, **kw):
    """tuning: jmme_tuning fields for the product library (explicit per-context state; the oracle has none)."""
    h, w = cur.shape
    is_cuda = lib.backend().startswith("cuda")
    with lib.context(width=w, height=h, num_refs=len(refs), tuning=tuning if is_cuda else None, **kw) as ctx:
        for i, r in enumerate(refs):
            ctx.set_reference(i, r)
        out = ctx.search_frame(cur, pred, per_ref)
        if is_cuda:
            LAST["kernel"], LAST["tuning"] = ctx.last_kernel(), ctx.get_tuning()
        return out


def oracle_threads(oracle, n):
    """All host threads for the big oracle runs (0 = every core), 1 = back to the JM-like single thread."""
    oracle.dll.jmme_oracle_set_threads.restype = ctypes.c_int
    return oracle.dll.jmme_oracle_set_threads(n)


# regular expressions of the integer-search instantiation each BASELINE config launches by default (whole frame)
DEFAULT_KERNELS = {
    "config1": r"me_int_kernel<K=3,NW=4,MINB=3,PER_BLOCK=0,ONLY16=1,RS_CT=0,MODE=0>$",
    # zero predictors, R = 32: whole items of MB pairs while a launch has less than two rounds of them (720p through
    # the host path: parts of 9 / 20 / 16 MB rows), else of 4 MBs
    "config2": r"me_int_tb_kernel<K=6,NW=4,MINB=3,PER_BLOCK=0,RS_CT=94,KEYG=0,KRTAB=1,NMB=2,CL=1,WP=0,LIN=0,BAL=0>$",
    "config3": r"me_int_tb_kernel<K=6,NW=4,MINB=3,PER_BLOCK=0,RS_CT=126,KEYG=0,KRTAB=1,NMB=4,CL=1,WP=0,LIN=0,BAL=0>$",
    "config4": r"me_int_kernel<K=5,NW=8,MINB=1,PER_BLOCK=0,ONLY16=0,RS_CT=144,MODE=0>$",
}


def assert_default_kernel(prefix):
    """The run used the library's default launch (no forced variant) and the kernel family the bench times."""
    assert LAST["tuning"]["variant"] == 0, LAST
    assert LAST["kernel"].startswith(prefix), LAST


def assert_same(got, exp, what=""):
    if got.tobytes() == exp.tobytes():
        return
    for f in ("mv", "cost", "ref_idx"):
        bad = np.argwhere(got[f] != exp[f])
        if len(bad):
            i = tuple(bad[0])
            idx = i[:-1] if f == "mv" else i
            raise AssertionError(
                f"{what}: field {f}: {len(bad)} mismatches; first at index {i}: "
                f"gpu mv={got['mv'][idx]} cost={got['cost'][idx]} ref={got['ref_idx'][idx]} | "
                f"oracle mv={exp['mv'][idx]} cost={exp['cost'][idx]} ref={exp['ref_idx'][idx]}")
    raise AssertionError(f"{what}: reserved bytes differ")


def test_library_is_the_cuda_backend(cuda):
    assert cuda.backend() == "cuda-sm_100a"
    assert cuda.dll.jmme_abi_version() == abi.ABI_VERSION


def test_tables_leaf(cuda, oracle):
    for R in (1, 7, 32, 64):
        a, b = cuda.init_motion_search_module(R, 700, 16), oracle.init_motion_search_module(R, 700, 16)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    for qp in range(52):
        for rd in (0, 1):
            assert cuda.lambda_factor(qp, rd) == oracle.lambda_factor(qp, rd)


@pytest.mark.parametrize("w,h,pad,kind,seed", [(16, 16, 8, "noise", 1), (64, 48, 24, "texture", 2),
                                               (176, 144, 48, "noise", 3), (48, 32, 12, "checker", 0),
                                               (32, 32, 16, "const", 0)])
def test_get_sub_images_luma_bit_exact(cuda, oracle, w, h, pad, kind, seed):
    img = synth.gen_luma(w, h, seed, kind)
    a, b = cuda.get_sub_images_luma(img, pad), oracle.get_sub_images_luma(img, pad)
    for fy in range(4):
        for fx in range(4):
            bad = np.argwhere(a[fy, fx] != b[fy, fx])
            assert len(bad) == 0, (f"plane ({fx},{fy}): {len(bad)} bytes differ, first at {bad[0]} "
                                   f"gpu {a[fy, fx][tuple(bad[0])]} oracle {b[fy, fx][tuple(bad[0])]}")


def test_context_planes_non_multiple_of_16(cuda, oracle):
    img = synth.gen_luma(52, 38, 4, "texture")
    kw = dict(width=52, height=38, search_range=6, subpel=1)
    with cuda.context(**kw) as g, oracle.context(**kw) as o:
        g.set_reference(0, img)
        o.set_reference(0, img)
        assert (g.mb_w, g.mb_h, g.pad) == (o.mb_w, o.mb_h, o.pad)
        for fy in range(4):
            for fx in range(4):
                assert np.array_equal(g.get_subimage(0, fx, fy), o.get_subimage(0, fx, fy)), (fx, fy)


def test_satd_leaf(cuda, oracle):
    rng = np.random.default_rng(3)
    d = rng.integers(-255, 256, size=(4096, 16)).astype(np.int16)
    d[0] = 255
    d[1] = np.tile([255, -255], 8)
    for rnd in (0, 1):
        assert np.array_equal(cuda.satd(d, rnd), oracle.satd(d, rnd))


def test_setup_fastfull_and_block_search_leaves(cuda, oracle):
    R = 7
    cur, refs = synth.frame_pair(64, 48, seed=5, search_range=R)
    pad = 2 * R + 16
    refp = refimpl.padded(refs[0], pad)
    f = oracle.lambda_factor(30, 1)
    for (mbx, mby, cx, cy, bonus) in [(0, 0, 0, 0, 0), (3, 2, -3, 5, 96), (1, 1, 7, -7, 0)]:
        c = cur[16 * mby:16 * mby + 16, 16 * mbx:16 * mbx + 16]
        a = cuda.setup_fast_full_pel_search(c, refp, pad, mbx, mby, cx, cy, R, bonus)
        b = oracle.setup_fast_full_pel_search(c, refp, pad, mbx, mby, cx, cy, R, bonus)
        assert np.array_equal(a, b)
        for blk in (0, 2, 7, 20, 40):
            for (px, py, pre) in [(0, 0, 1), (4 * cx + 1, 4 * cy - 2, 0), (-9, 30, 1)]:
                ga = cuda.fast_full_pel_block_motion_search(b[blk], R, cx, cy, px, py, f, pre)
                gb = oracle.fast_full_pel_block_motion_search(b[blk], R, cx, cy, px, py, f, pre)
                assert ga == gb, (mbx, mby, blk, px, py, pre)
    for (bx, by, bw, bh, px, py) in [(16, 16, 16, 16, 0, 0), (36, 20, 4, 8, -11, 6), (8, 40, 8, 4, 40, -40)]:
        ga = cuda.full_pel_block_motion_search(cur, refp, pad, bx, by, bw, bh, px, py, R, f, 37)
        gb = oracle.full_pel_block_motion_search(cur, refp, pad, bx, by, bw, bh, px, py, R, f, 37)
        assert ga == gb


def test_subpel_leaf(cuda, oracle):
    cur, refs = synth.frame_pair(64, 48, seed=6, search_range=4)
    pad = 16
    planes = oracle.get_sub_images_luma(refs[0], pad)
    f = oracle.lambda_factor(26, 0)
    for (bx, by, bw, bh) in [(16, 16, 16, 16), (32, 8, 8, 16), (4, 4, 4, 4), (40, 28, 8, 4)]:
        for (mv, had, rnd, bonus) in [((0, 0), 1, 0, 64), ((8, -4), 1, 1, 0), ((-12, 4), 0, 0, 0)]:
            ga = cuda.sub_pel_block_motion_search(cur, planes, pad, bx, by, bw, bh, 3, -5, f, mv, 5000, had, rnd, bonus)
            gb = oracle.sub_pel_block_motion_search(cur, planes, pad, bx, by, bw, bh, 3, -5, f, mv, 5000, had, rnd,
                                                    bonus)
            assert ga == gb, (bx, by, bw, bh, mv, had)


CASES = [
    # w, h, R, kwargs
    (64, 48, 4, dict()),
    (64, 48, 4, dict(rdopt=1, qp=33)),
    (80, 64, 9, dict(qp=20)),
    (96, 64, 16, dict(rdopt=1, qp=40)),
    (64, 64, 32, dict(qp=28)),
    (80, 48, 32, dict(qp=30)),                                 # 5 MBs per row: item groups of 4 + 1
    (112, 32, 32, dict(rdopt=1, qp=26, subpel=1)),            # 7 MBs per row
    (48, 48, 5, dict(subpel=1)),
    (64, 48, 8, dict(subpel=1, rdopt=1, qp=24, satd_round=1)),
    (64, 48, 6, dict(subpel=1, use_hadamard=0)),
    (52, 38, 6, dict(subpel=1)),                               # not a multiple of 16
    (64, 48, 7, dict(blocktype_mask=abi.MASK_16x16, search_mode=abi.SEARCH_FULL)),
    (64, 48, 7, dict(blocktype_mask=abi.MASK_16x16, subpel=1)),
    (64, 48, 7, dict(blocktype_mask=0x92, subpel=1)),            # 16x16, 8x8, 4x4 only
    (64, 48, 6, dict(search_mode=abi.SEARCH_FULL, rdopt=0)),
    (48, 32, 1, dict(subpel=1)),
    (48, 32, 2, dict()),
    (48, 32, 3, dict(rdopt=1)),
]


@pytest.mark.parametrize("variant", [0, 32, 47, 69, 51])
@pytest.mark.parametrize("w,h,R,kw", CASES)
def test_search_frame_matches_oracle(cuda, oracle, w, h, R, kw, variant):
    for kind, seed in (("texture", 1), ("noise", 2)):
        cur, refs = synth.frame_pair(w, h, seed=seed, search_range=R, kind=kind)
        got = run(cuda, cur, refs, search_range=R, tuning=dict(variant=variant), **kw)
        exp = run(oracle, cur, refs, search_range=R, **kw)
        assert_same(got, exp, f"{w}x{h} R={R} {kw} {kind}")


@pytest.mark.parametrize("variant", [0, 30, 31, 51, 22, 32, 47, 48, 68, 88, 69, 66])
@pytest.mark.parametrize("policy,nb", [(abi.PRED_PER_MB, 1), (abi.PRED_PER_BLOCK, 41)])
@pytest.mark.parametrize("rdopt", [0, 1])
def test_predictor_policies(cuda, oracle, policy, nb, rdopt, variant):
    w, h, R = 64, 48, 8
    cur, refs = synth.frame_pair(w, h, seed=4, search_range=R, num_refs=2)
    pred = synth.random_pred(2, 12, nb, seed=7 + rdopt, max_qpel=4 * R + 30)   # some centres get clamped
    kw = dict(search_range=R, qp=31, rdopt=rdopt, pred_policy=policy, subpel=1)
    g, gp = run(cuda, cur, refs, pred, True, tuning=dict(variant=variant), **kw)
    o, op = run(oracle, cur, refs, pred, True, **kw)
    assert_same(gp, op, "per-ref")
    assert_same(g, o, "best-ref")


def test_constant_and_checker_frames_all_ties(cuda, oracle):
    for kind in ("const", "checker"):
        cur = synth.gen_luma(48, 48, 0, kind)
        for rdopt in (0, 1):
            for lf in (0, 1):
                kw = dict(search_range=6, rdopt=rdopt, lambda_factor=lf, qp=28, subpel=1)
                assert_same(run(cuda, cur, [cur], **kw), run(oracle, cur, [cur], **kw),
                            f"{kind} rdopt={rdopt} lf={lf}")
    pred = np.array([8, -4], np.int16) * np.ones((1, 9, 1, 2), np.int16)
    for rdopt in (0, 1):
        kw = dict(search_range=6, rdopt=rdopt, lambda_factor=1, pred_policy=abi.PRED_PER_MB)
        cur = synth.gen_luma(48, 48, 0, "const")
        assert_same(run(cuda, cur, [cur], pred, **kw), run(oracle, cur, [cur], pred, **kw))


def test_max_sad_no_overflow_in_packed_cost(cuda, oracle):
    """0/255 opposition gives the largest possible 16x16 SAD (65280) with the largest lambda."""
    cur = np.zeros((32, 32), np.uint8)
    ref = np.full((32, 32), 255, np.uint8)
    pred = synth.random_pred(1, 4, 41, seed=1, max_qpel=2048)
    for rdopt in (0, 1):
        kw = dict(search_range=4, qp=51, rdopt=rdopt, pred_policy=abi.PRED_PER_BLOCK)
        assert_same(run(cuda, cur, [ref], pred, **kw), run(oracle, cur, [ref], pred, **kw))


def test_multi_ref_and_stripes(cuda, oracle):
    w, h, R = 96, 80, 8
    cur, refs = synth.frame_pair(w, h, seed=12, search_range=R, num_refs=4)
    kw = dict(search_range=R, qp=30, rdopt=1, subpel=1)
    g, gp = run(cuda, cur, refs, None, True, **kw)
    o, op = run(oracle, cur, refs, None, True, **kw)
    assert_same(gp, op, "per-ref")
    assert_same(g, o, "best")
    parts = np.zeros_like(g)
    for (a, b) in [(0, 2), (2, 3), (3, 5)]:
        s = run(cuda, cur, refs, mb_row_begin=a, mb_row_end=b, **kw)
        parts[a * 6:b * 6] = s[a * 6:b * 6]
        assert not s[:a * 6].tobytes().strip(b"\0") and not s[b * 6:].tobytes().strip(b"\0")
    assert parts.tobytes() == g.tobytes()


def test_errors_mirror_the_oracle(cuda, oracle):
    for lib in (cuda, oracle):
        for kw in (dict(width=0, height=16), dict(width=16, height=16, search_range=65),
                   dict(width=16, height=16, num_refs=5), dict(width=16, height=16, blocktype_mask=0),
                   dict(width=16, height=16, mb_row_begin=1, mb_row_end=1)):
            with pytest.raises(abi.JmmeError) as e:
                lib.context(**kw)
            assert e.value.code == abi.ERR_PARAM
        with pytest.raises(abi.JmmeError) as e:
            lib.context(width=16, height=16, me_distortion=1, me_distortion_fpel=abi.DIST_HADAMARD)
        assert e.value.code == abi.ERR_UNSUPPORTED
        with lib.context(width=32, height=32, search_range=4) as ctx:
            with pytest.raises(abi.JmmeError) as e:
                ctx.search_frame(np.zeros((32, 32), np.uint8))
            assert e.value.code == abi.ERR_STATE
        with lib.context(width=32, height=32, search_range=4, pred_policy=abi.PRED_PER_MB) as ctx:
            ctx.set_reference(0, np.zeros((32, 32), np.uint8))
            with pytest.raises(abi.JmmeError) as e:
                ctx.search_frame(np.zeros((32, 32), np.uint8), np.full((1, 4, 1, 2), 3000, np.int16))
            assert e.value.code == abi.ERR_PARAM


def test_config1_cif_16x16_full_search(cuda, oracle):
    """BASELINE config 1: CIF 352x288, 16x16 only, +-16, 1 ref, integer (FullPelBlockMotionSearch)."""
    cur, refs = synth.frame_pair(352, 288, seed=1, search_range=16)
    kw = dict(search_range=16, blocktype_mask=abi.MASK_16x16, search_mode=abi.SEARCH_FULL, qp=28)
    assert_same(run(cuda, cur, refs, **kw), run(oracle, cur, refs, **kw), "config 1")
    assert_default_kernel("me_int_kernel<")


def test_config2_720p_whole_frame_against_oracle_and_planted_motion(cuda, oracle):
    """BASELINE config 2 at full size: every MB of the frame against the oracle (all host threads), under the
    default kernel; planted motion everywhere."""
    w, h, R = 1280, 720, 32
    cur, refs = synth.frame_pair(w, h, seed=1, search_range=R)
    kw = dict(search_range=R, qp=28)
    g = run(cuda, cur, refs, **kw)
    assert_default_kernel("me_int_tb_kernel<")
    oracle_threads(oracle, 0)
    try:
        o = run(oracle, cur, refs, **kw)
    finally:
        oracle_threads(oracle, 1)
    assert len(g) == 80 * 45
    assert_same(g, o, "720p whole frame")
    ref = synth.gen_luma(w, h, 5, "noise")
    dx, dy = -19, 27
    cur2 = np.roll(ref, (-dy, -dx), axis=(0, 1))
    g2 = run(cuda, cur2, [ref], search_range=R, qp=28, rdopt=1).reshape(45, 80)
    inner = g2[3:-3, 3:-3]
    assert np.all(inner["mv"] == [4 * dx, 4 * dy])
    lam = oracle.lambda_factor(28, 1)
    exp = refimpl.weighted_cost(lam, refimpl.se_bits(4 * dx) + refimpl.se_bits(4 * dy)) + refimpl.weighted_cost(lam, 1)
    assert np.all(inner["cost"] == exp)


@pytest.mark.parametrize("seed", [1, 2])
def test_config3_1080p_whole_frame_against_oracle(cuda, oracle, seed):
    """BASELINE config 3 (the headline) at full size: all 8160 MBs x 41 blocks against the oracle (all host
    threads; includes the rows replicated from 1080 to 1088), under the default kernels; seed 1 is the frame
    pair bench.py times.  Stripes reproduce the whole-frame result."""
    w, h, R = 1920, 1080, 32
    cur, refs = synth.frame_pair(w, h, seed=seed, search_range=R)
    kw = dict(search_range=R, qp=28, subpel=1)
    g = run(cuda, cur, refs, **kw)
    assert_default_kernel("me_int_tb_kernel<")
    assert len(g) == 120 * 68
    oracle_threads(oracle, 0)
    try:
        o = run(oracle, cur, refs, **kw)
    finally:
        oracle_threads(oracle, 1)
    assert_same(g, o, f"1080p whole frame, seed {seed}")
    s = run(cuda, cur, refs, mb_row_begin=30, mb_row_end=41, **kw)
    assert s[30 * 120:41 * 120].tobytes() == g[30 * 120:41 * 120].tobytes()


@pytest.mark.parametrize("w,h,subpel", [(1920, 1080, 0), (3840, 2160, 1)])
def test_config4_and_config5_r64_four_refs(cuda, oracle, w, h, subpel):
    """BASELINE configs 4 (1080p, integer) and 5 (4K, quarter-pel): R = 64, 4 references, at full size on the
    GPU under the default R = 64 kernel; the oracle (all host threads) on ten MB rows spread over the frame
    — top, bottom, the seams of the 2/4/8-rank stripe partitions and rows in between; two stripes reproduce the
    whole-frame field."""
    R, nref = 64, 4
    cur, refs = synth.frame_pair(w, h, seed=4, search_range=R, num_refs=nref)
    kw = dict(search_range=R, qp=28, subpel=subpel)
    mb_w, mb_h = (w + 15) // 16, (h + 15) // 16
    g, gp = run(cuda, cur, refs, None, True, **kw)
    assert LAST["tuning"]["variant"] == 0 and "R64" not in LAST["kernel"], LAST
    default_kernel = LAST["kernel"]
    per8 = -(-mb_h // 8)
    rows = sorted({0, 1, per8 - 1, per8, 2 * per8, 4 * per8 - 1, 4 * per8, 6 * per8, mb_h - 2, mb_h - 1})
    oracle_threads(oracle, 0)
    try:
        for row in rows:
            o, op = run(oracle, cur, refs, None, True, mb_row_begin=row, mb_row_end=row + 1, **kw)
            sl = slice(row * mb_w, (row + 1) * mb_w)
            assert_same(gp[:, sl], op[:, sl], f"per-ref row {row}")
            assert_same(g[sl], o[sl], f"best row {row}")
    finally:
        oracle_threads(oracle, 1)
    assert len(np.unique(g["ref_idx"])) > 1                           # more than one reference wins somewhere
    half = mb_h // 2
    s0 = run(cuda, cur, refs, mb_row_begin=0, mb_row_end=half, **kw)
    s1 = run(cuda, cur, refs, mb_row_begin=half, mb_row_end=mb_h, **kw)
    assert LAST["kernel"] == default_kernel
    assert s0[:half * mb_w].tobytes() == g[:half * mb_w].tobytes()
    assert s1[half * mb_w:].tobytes() == g[half * mb_w:].tobytes()


def test_default_kernels_of_the_baseline_configs(cuda):
    """Which integer-search instantiation each BASELINE config launches by default (jmme_last_kernel) — the
    names bench.py prints in its JSON line (`roofline.kernel_instance`).  Updated whenever a default changes."""
    import re
    expect = {
        (352, 288, 16, 1, abi.MASK_16x16, abi.SEARCH_FULL): DEFAULT_KERNELS["config1"],
        (1280, 720, 32, 1, abi.MASK_ALL, abi.SEARCH_FASTFULL): DEFAULT_KERNELS["config2"],
        (1920, 1080, 32, 1, abi.MASK_ALL, abi.SEARCH_FASTFULL): DEFAULT_KERNELS["config3"],
        (1920, 1080, 64, 4, abi.MASK_ALL, abi.SEARCH_FASTFULL): DEFAULT_KERNELS["config4"],
    }
    for (w, h, R, nref, mask, mode), pat in expect.items():
        cur, refs = synth.frame_pair(w, h, seed=1, search_range=R, num_refs=nref)
        run(cuda, cur, refs, search_range=R, blocktype_mask=mask, search_mode=mode)
        assert re.match(pat, LAST["kernel"]), (w, h, R, LAST)


@pytest.mark.parametrize("tuning", [dict(balance=1), dict(balance=1, group=2), dict(balance=1, group=1), dict(balance=1, variant=65), dict(balance=1, variant=48)])
@pytest.mark.parametrize("w,h,rows,nref,subpel", [(1920, 1080, None, 1, 1), (1920, 1080, (0, 9), 1, 1), (1920, 1080, (59, 68), 1, 0),
                                                  (1280, 720, None, 2, 0), (1288, 728, None, 1, 1), (336, 64, None, 2, 1),
                                                  (80, 48, None, 1, 1), (16, 16, None, 1, 0)])
def test_balanced_task_ranges_reproduce_whole_items(cuda, oracle, tuning, w, h, rows, nref, subpel):
    """The zero-predictor search with balanced task ranges (me_int_tb.cu BAL: every CTA takes an equal range of the
    stripe's tasks, MBs split over CTAs meet in global memory, the consuming kernel decodes) in groups of 1, 2 and 4
    MBs and other CTA shapes, with one and two references, with the sub-pel kernel or the reference selection as
    the consumer, on whole frames (odd MB counts too), on the stripes an 8-GPU run gives a rank and on frames much
    smaller than the grid: every shape gives the field of the whole-item launch (balance = 2), twice in a row (the
    consumer resets the packed words), and small frames are checked against the oracle."""
    R = 32
    cur, refs = synth.frame_pair(w, h, seed=6, search_range=R, num_refs=nref)
    kw = dict(search_range=R, qp=28, subpel=subpel)
    if rows:
        kw.update(mb_row_begin=rows[0], mb_row_end=rows[1])
    ref = run(cuda, cur, refs, tuning=dict(balance=2, group=2, early_subpel=2), **kw)
    assert "BAL=0" in LAST["kernel"], LAST
    dflt = run(cuda, cur, refs, **kw)                            # the default: whole items + early sub-pel start
    assert "BAL=0" in LAST["kernel"] and dflt.tobytes() == ref.tobytes(), LAST
    with cuda.context(width=w, height=h, num_refs=nref, tuning=tuning, **kw) as ctx:
        for i, r in enumerate(refs):
            ctx.set_reference(i, r)
        for rep in range(2):
            got = ctx.search_frame(cur)
            assert "BAL=1" in ctx.last_kernel(), ctx.last_kernel()
            assert got.tobytes() == ref.tobytes(), (tuning, rep, ctx.last_kernel())
    if w * h <= 336 * 64:
        assert_same(got, run(oracle, cur, refs, **kw), f"{w}x{h} {tuning}")


@pytest.mark.parametrize("tuning", [dict(), dict(early_subpel=2), dict(no_pair_tail=1), dict(group=2), dict(group=4, early_subpel=2),
                                    dict(even_parts=1, pipe_parts=4)])
@pytest.mark.parametrize("nref", [1, 2])
def test_early_subpel_start_and_item_shapes(cuda, tuning, nref):
    """The whole-item search with the early sub-pel start (per-MB ready flags, programmatic dependent launch), items of
    4 MBs with a tail of pairs, through the host path and the device path, twice in a row in one context (the sub-pel
    kernel lowers the flags): every launch shape gives the same field as the plain serial form."""
    import torch
    from jmme.torch_api import DeviceSearch
    w, h, R = 1920, 400, 32                                   # 25 MB rows: 4-MB items in whole rounds + pair rows
    cur, refs = synth.frame_pair(w, h, seed=17, search_range=R, num_refs=nref)
    kw = dict(search_range=R, qp=28, subpel=1)
    ref = run(cuda, cur, refs, tuning=dict(early_subpel=2, no_pair_tail=1, group=2, balance=2, pipe_parts=1), **kw)
    with cuda.context(width=w, height=h, num_refs=nref, tuning=tuning, **kw) as ctx:
        for i, r in enumerate(refs):
            ctx.set_reference(i, r)
        for rep in range(2):
            assert ctx.search_frame(cur).tobytes() == ref.tobytes(), (tuning, rep, ctx.last_kernel())
    ds = DeviceSearch(cuda, width=w, height=h, num_refs=nref, tuning=tuning, **kw)
    dcur = torch.from_numpy(cur).cuda()
    for i, r in enumerate(refs):
        ds.set_reference(i, torch.from_numpy(r).cuda())
    for rep in range(3):
        out = ds.search(dcur)
        torch.cuda.synchronize()
        assert ds.to_numpy(out).tobytes() == ref.tobytes(), (tuning, rep)
    ds.close()


def test_launch_counter_counts_kernels(cuda):
    cur, refs = synth.frame_pair(64, 48, seed=1, search_range=4)
    with cuda.context(width=64, height=48, search_range=4, subpel=1) as ctx:
        ctx.set_reference(0, refs[0])
        n0 = ctx.launch_count()
        ctx.search_frame(cur)
        assert n0 == 1 and ctx.launch_count() - n0 == 2          # me_int, me_subpel (writes the records itself)


@pytest.mark.parametrize("n", [2, 3, 8])
def test_in_library_multi_gpu_mode_on_virtual_devices(cuda, oracle, n):
    """jmme_params.n_gpus splits the frame into MB-row stripes, one sub-context each, and gathers the MV
    field to the first device.  With fewer GPUs than stripes the same device is listed several times
    (N virtual GPUs on one device must reproduce the one-stripe field byte for byte)."""
    import torch
    ndev = torch.cuda.device_count()
    w, h, R = 96, 144, 8                                     # 9 MB rows
    cur, refs = synth.frame_pair(w, h, seed=21, search_range=R, num_refs=2)
    kw = dict(search_range=R, qp=29, subpel=1)
    ids = [i % ndev for i in range(n)]
    g, gp = run(cuda, cur, refs, None, True, n_gpus=n, device_ids=ids, **kw)
    o, op = run(oracle, cur, refs, None, True, **kw)
    assert_same(gp, op, f"per-ref n_gpus={n}")
    assert_same(g, o, f"best n_gpus={n}")


@pytest.mark.parametrize("rdopt,subpel", [(0, 0), (1, 1)])
def test_full_search_with_per_block_windows(cuda, oracle, rdopt, subpel):
    """search_mode FULL + 41 predictors per MB: every block searches a window centred on its own predictor
    (JM FullPelBlockMotionSearch, a8) — the dedicated me_full kernel."""
    w, h, R = 64, 48, 7
    cur, refs = synth.frame_pair(w, h, seed=15, search_range=R, num_refs=2)
    pred = synth.random_pred(2, 12, 41, seed=3, max_qpel=4 * R + 9)
    kw = dict(search_range=R, qp=27, rdopt=rdopt, subpel=subpel, search_mode=abi.SEARCH_FULL,
              pred_policy=abi.PRED_PER_BLOCK)
    g, gp = run(cuda, cur, refs, pred, True, **kw)
    o, op = run(oracle, cur, refs, pred, True, **kw)
    assert_same(gp, op, "per-ref")
    assert_same(g, o, "best")
    kw["blocktype_mask"] = 0x8A                                  # 16x16, 8x16, 4x4
    assert_same(run(cuda, cur, refs, pred, **kw), run(oracle, cur, refs, pred, **kw), "masked")


def test_predictor_commit_and_closed_loop(cuda, oracle):
    """a3: SetMotionVectorPredictor, the ME-only field commit and the frame-wide predictor kernel, then the
    closed loop search -> commit -> predict -> search with per-block predictors, all against the oracle."""
    rng = np.random.default_rng(9)
    for _ in range(500):
        nb = [(int(rng.integers(-64, 65)), int(rng.integers(-64, 65)), int(rng.integers(-1, 3)), int(rng.integers(0, 2)))
              for _ in range(3)]
        t, part, ref = int(rng.integers(1, 8)), int(rng.integers(0, 2)), int(rng.integers(0, 3))
        assert cuda.set_motion_vector_predictor(t, part, ref, *nb) == oracle.set_motion_vector_predictor(t, part, ref, *nb)
    w, h, R = 96, 80, 8
    cur, refs = synth.frame_pair(w, h, seed=17, search_range=R, num_refs=2)
    for mask in (abi.MASK_ALL, 0x92):
        kw = dict(width=w, height=h, search_range=R, num_refs=2, qp=31, rdopt=1, subpel=1, blocktype_mask=mask)
        with cuda.context(**kw) as g, oracle.context(**kw) as o:
            for i, r in enumerate(refs):
                g.set_reference(i, r)
                o.set_reference(i, r)
            rg, ro = g.search_frame(cur), o.search_frame(cur)
            assert_same(rg, ro, "pass 1")
            fg, fo = g.commit_field(rg), o.commit_field(ro)
            for a, b in zip(fg, fo):
                assert np.array_equal(a, b)
            fo[1][3, 5] = -1
            pg, po = g.predict_frame(fo[0], fo[1]), o.predict_frame(fo[0], fo[1])
            assert np.array_equal(pg, po)
        kw["pred_policy"] = abi.PRED_PER_BLOCK
        assert_same(run(cuda, cur, refs, po, **{k: v for k, v in kw.items() if k not in ("width", "height", "num_refs")}),
                    run(oracle, cur, refs, po, **{k: v for k, v in kw.items() if k not in ("width", "height", "num_refs")}),
                    "pass 2")


def test_randomised_configurations(cuda, oracle):
    """Seeded random sweep over sizes, search ranges, lambda, masks, policies, references and sub-pel options."""
    rng = np.random.default_rng(2026)
    for case in range(24):
        w, h = int(rng.integers(17, 130)), int(rng.integers(17, 100))
        R = int(rng.choice([1, 2, 3, 5, 8, 13, 16, 21, 32, 40]))
        refs_n = int(rng.integers(1, 4))
        mask = int(rng.choice([0xFE, 0x02, 0x92, 0x0E, 0xF0, 0x80, 0xFE]))
        policy = int(rng.integers(0, 3))
        mode = int(rng.integers(0, 2))
        kw = dict(search_range=R, qp=int(rng.integers(0, 52)), rdopt=int(rng.integers(0, 2)), subpel=int(rng.integers(0, 2)),
                  use_hadamard=int(rng.integers(0, 2)), satd_round=int(rng.integers(0, 2)), blocktype_mask=mask,
                  pred_policy=policy, search_mode=mode)
        kind = str(rng.choice(["texture", "noise", "gradient"]))
        cur, refs = synth.frame_pair(w, h, seed=case + 1, search_range=R, kind=kind, num_refs=refs_n)
        n_mb = ((w + 15) // 16) * ((h + 15) // 16)
        pred = None if policy == 0 else synth.random_pred(refs_n, n_mb, 1 if policy == 1 else 41, case, 4 * R + 40)
        g, gp = run(cuda, cur, refs, pred, True, **kw)
        o, op = run(oracle, cur, refs, pred, True, **kw)
        assert_same(gp, op, f"case {case} per-ref {w}x{h} {kw}")
        assert_same(g, o, f"case {case} best {w}x{h} {kw}")


# ---- JMME_PRED_MEDIAN: predictor loop closed inside the frame (wavefront on the GPU, raster order in the oracle) ----
def run_median(lib, cur, refs, tuning=None, **kw):
    h, w = cur.shape
    is_cuda = lib.backend().startswith("cuda")
    with lib.context(width=w, height=h, num_refs=len(refs), pred_policy=abi.PRED_MEDIAN,
                     tuning=tuning if is_cuda else None, **kw) as ctx:
        for i, r in enumerate(refs):
            ctx.set_reference(i, r)
        out, opr = ctx.search_frame(cur, None, True)
        if is_cuda:
            LAST["kernel"], LAST["tuning"] = ctx.last_kernel(), ctx.get_tuning()
        return out, opr, ctx.get_predictors()


MEDIAN_CASES = [
    (80, 64, 6, dict(slice_rows=0, qp=30, subpel=1), 2),
    (80, 64, 6, dict(slice_rows=1, qp=30, subpel=1), 1),
    (96, 80, 8, dict(slice_rows=2, qp=26, subpel=1, rdopt=1), 2),
    (72, 56, 5, dict(slice_rows=0, qp=36, subpel=0), 1),                       # sizes not multiples of 16
    (64, 48, 5, dict(slice_rows=0, qp=34, rdopt=1, search_mode=abi.SEARCH_FULL, subpel=1), 1),
    (64, 48, 5, dict(slice_rows=2, qp=28, blocktype_mask=0x92, subpel=1), 2),
    (64, 48, 5, dict(slice_rows=0, blocktype_mask=abi.MASK_16x16, subpel=1), 1),
    (176, 144, 16, dict(slice_rows=3, qp=28, subpel=1), 1),
    (128, 96, 32, dict(slice_rows=0, qp=32, subpel=1, use_hadamard=0), 1),
    (64, 64, 40, dict(slice_rows=0, qp=30, subpel=0), 1),                      # R > 32: the me_int.cu kernel
]


@pytest.mark.parametrize("w,h,R,kw,nref", MEDIAN_CASES)
def test_in_frame_median_matches_oracle(cuda, oracle, w, h, R, kw, nref):
    cur, refs = synth.frame_pair(w, h, seed=w + R, search_range=R, num_refs=nref)
    g, gp, gpred = run_median(cuda, cur, refs, search_range=R, **kw)
    o, op, opred = run_median(oracle, cur, refs, search_range=R, **kw)
    assert np.array_equal(gpred, opred), "predictors"
    assert_same(gp, op, "per-ref")
    assert_same(g, o, "best")


@pytest.mark.parametrize("variant", [68, 64, 65, 47, 32])
def test_in_frame_median_kernel_variants(cuda, oracle, variant):
    w, h, R = 96, 64, 8
    cur, refs = synth.frame_pair(w, h, seed=3, search_range=R)
    kw = dict(search_range=R, slice_rows=2, qp=30, subpel=1)
    g, _, gpred = run_median(cuda, cur, refs, tuning=dict(variant=variant), **kw)
    assert LAST["tuning"]["variant"] == variant
    o, _, opred = run_median(oracle, cur, refs, **kw)
    assert np.array_equal(gpred, opred)
    assert_same(g, o, f"variant {variant}")


@pytest.mark.parametrize("n", [2, 3])
def test_in_frame_median_stripes_and_virtual_devices(cuda, oracle, n):
    """Whole slices per device: the result does not depend on the device count; stripes that cut a slice
    are refused like the oracle refuses them."""
    import torch
    ndev = torch.cuda.device_count()
    w, h, R = 64, 112, 6                                     # 7 MB rows, slices of 2 rows -> 4 slices
    cur, refs = synth.frame_pair(w, h, seed=9, search_range=R, num_refs=2)
    kw = dict(search_range=R, slice_rows=2, qp=30, subpel=1)
    o, op, opred = run_median(oracle, cur, refs, **kw)
    g, gp, gpred = run_median(cuda, cur, refs, n_gpus=n, device_ids=[i % ndev for i in range(n)], **kw)
    assert np.array_equal(gpred, opred)
    assert_same(gp, op, f"per-ref n_gpus={n}")
    assert_same(g, o, f"best n_gpus={n}")
    for lib in (cuda, oracle):
        with pytest.raises(abi.JmmeError) as e:
            lib.context(width=w, height=h, pred_policy=abi.PRED_MEDIAN, mb_row_begin=1, mb_row_end=4, **kw)
        assert e.value.code == abi.ERR_PARAM


def test_in_frame_median_device_api_and_repeat(cuda, oracle):
    """Device-resident entry point, two frames in a row through one context (the field of the first frame
    must not leak into the second)."""
    import torch
    from jmme.torch_api import DeviceSearch
    w, h, R = 96, 64, 8
    kw = dict(width=w, height=h, search_range=R, pred_policy=abi.PRED_MEDIAN, slice_rows=0, qp=30, subpel=1)
    ds = DeviceSearch(cuda, **kw)
    for seed in (5, 6):
        cur, refs = synth.frame_pair(w, h, seed=seed, search_range=R)
        ds.set_reference(0, torch.from_numpy(refs[0]).cuda())
        got = ds.to_numpy(ds.search(torch.from_numpy(cur).cuda()))
        torch.cuda.synchronize()
        o, _, _ = run_median(oracle, cur, refs, **{k: v for k, v in kw.items() if k not in ("width", "height", "pred_policy")})
        assert_same(got, o, f"seed {seed}")
    ds.close()


def test_yuv_sequence_pass_matches_oracle(cuda, oracle, tmp_path):
    """encoder.cfg + planar YUV file -> IPPP ME pass (jmme/cfg.py, jmme/sequence.py), GPU against oracle."""
    from jmme.cfg import EncoderCfg
    from jmme.sequence import search_sequence, yuv_frames
    w, h, R = 96, 80, 8
    lumas = [synth.gen_luma(w, h, 5)] + [synth.frame_pair(w, h, seed=5, search_range=R + k)[0] for k in range(3)]
    synth.write_yuv420(tmp_path / "seq.yuv", lumas)
    (tmp_path / "encoder.cfg").write_text(
        f"InputFile = seq.yuv\nFramesToBeEncoded = 4\nSourceWidth = {w}\nSourceHeight = {h}\nSearchRange = {R}\n"
        "NumberReferenceFrames = 2\nQPPSlice = 30\nSliceMode = 1\nSliceArgument = 12\n")
    cfg = EncoderCfg.load(tmp_path / "encoder.cfg")
    kw = cfg.params()
    kw.pop("width"), kw.pop("height")
    for policy in (abi.PRED_ZERO, abi.PRED_MEDIAN):
        k2 = dict(kw) if policy == abi.PRED_MEDIAN else {k: v for k, v in kw.items() if k != "slice_rows"}
        a = list(search_sequence(cuda, yuv_frames(tmp_path / "seq.yuv", w, h, cfg.frames), policy, **k2))
        b = list(search_sequence(oracle, yuv_frames(tmp_path / "seq.yuv", w, h, cfg.frames), policy, **k2))
        assert len(a) == len(b) == 3
        for (n, ga, _), (_, gb, _) in zip(a, b):
            assert_same(ga, gb, f"policy {policy} frame {n}")


def test_get_predictors_states_mirror_the_oracle(cuda, oracle):
    w, h, R = 32, 32, 4
    cur, refs = synth.frame_pair(w, h, seed=1, search_range=R)
    for lib in (cuda, oracle):
        with lib.context(width=w, height=h, search_range=R) as ctx:
            with pytest.raises(abi.JmmeError) as e:
                ctx.get_predictors()
            assert e.value.code == abi.ERR_STATE
        with lib.context(width=w, height=h, search_range=R, pred_policy=abi.PRED_MEDIAN) as ctx:
            with pytest.raises(abi.JmmeError) as e:
                ctx.get_predictors()
            assert e.value.code == abi.ERR_STATE
            ctx.set_reference(0, refs[0])
            ctx.search_frame(cur)
            assert ctx.get_predictors().shape == (1, 4, 41, 2)
        for bad in (dict(pred_policy=abi.PRED_MEDIAN, slice_rows=-1), dict(pred_policy=4)):
            with pytest.raises(abi.JmmeError) as e:
                lib.context(width=w, height=h, search_range=R, **bad)
            assert e.value.code == abi.ERR_PARAM


def test_randomised_in_frame_median(cuda, oracle):
    """Seeded random sweep of the in-frame median policy: sizes, ranges, lambda, masks, slices, references,
    search modes and sub-pel options; predictors and records against the oracle."""
    rng = np.random.default_rng(77)
    for case in range(20):
        w, h = int(rng.integers(17, 150)), int(rng.integers(17, 120))
        R = int(rng.choice([1, 2, 3, 5, 8, 13, 16, 21, 32, 40]))
        refs_n = int(rng.integers(1, 4))
        kw = dict(search_range=R, qp=int(rng.integers(0, 52)), rdopt=int(rng.integers(0, 2)), subpel=int(rng.integers(0, 2)),
                  use_hadamard=int(rng.integers(0, 2)), satd_round=int(rng.integers(0, 2)),
                  blocktype_mask=int(rng.choice([0xFE, 0x02, 0x92, 0x0E, 0xF0, 0x80, 0xFE])),
                  search_mode=int(rng.integers(0, 2)), slice_rows=int(rng.choice([0, 0, 1, 2, 3])))
        kind = str(rng.choice(["texture", "noise", "gradient"]))
        cur, refs = synth.frame_pair(w, h, seed=100 + case, search_range=R, kind=kind, num_refs=refs_n)
        g, gp, gpred = run_median(cuda, cur, refs, **kw)
        o, op, opred = run_median(oracle, cur, refs, **kw)
        assert np.array_equal(gpred, opred), f"case {case} predictors {w}x{h} {kw}"
        assert_same(gp, op, f"case {case} per-ref {w}x{h} {kw}")
        assert_same(g, o, f"case {case} best {w}x{h} {kw}")


@pytest.mark.parametrize("force_wave_step", [0, 1])
def test_in_frame_median_tall_frame(cuda, oracle, force_wave_step):
    """135 MB rows (4K height): more MBs per wavefront step than one chunk of wave_step_kernel (128) and more
    (MB, ref) items than SMs; both predictor paths (search-kernel prologue / wave_step_kernel)."""
    w, h, R = 256, 2160, 6
    cur, refs = synth.frame_pair(w, h, seed=8, search_range=R, num_refs=2)
    kw = dict(search_range=R, slice_rows=1, qp=30, subpel=1)
    g, gp, gpred = run_median(cuda, cur, refs, tuning=dict(wave_step=force_wave_step), **kw)
    assert LAST["kernel"].startswith("me_int_tb_kernel<") and ("WP=0" if force_wave_step else "WP=1") in LAST["kernel"], LAST
    oracle_threads(oracle, 0)
    try:
        o, op, opred = run_median(oracle, cur, refs, **kw)
    finally:
        oracle_threads(oracle, 1)
    assert np.array_equal(gpred, opred)
    assert_same(gp, op, "per-ref")
    assert_same(g, o, "best")


@pytest.mark.parametrize("h,stripe", [(208, None), (400, (4, 21)), (1080, None)])
def test_async_reference_with_the_pipelined_host_path(cuda, oracle, h, stripe):
    """async_reference = 1 (jmme_set_reference returns before its copy and plane kernel have run) together with
    the host path that searches the stripe in parts on separate streams."""
    w, R = (1920, 32) if h == 1080 else (128, 8)
    cur, refs = synth.frame_pair(w, h, seed=h, search_range=R, num_refs=1 if h == 1080 else 2)
    kw = dict(search_range=R, qp=28, subpel=1)
    if stripe:
        kw.update(mb_row_begin=stripe[0], mb_row_end=stripe[1])
    g, gp = run(cuda, cur, refs, None, True, async_reference=1, **kw)
    if h == 1080:                                            # oracle on a few rows around the part boundaries
        for rb in (21, 44, 66):
            o, op = run(oracle, cur, refs, None, True, **dict(kw, mb_row_begin=rb, mb_row_end=rb + 2))
            sl = slice(rb * 120, (rb + 2) * 120)
            assert_same(g[sl], o[sl], f"rows {rb}..")
        g0 = run(cuda, cur, refs, **kw)
        assert g0.tobytes() == g.tobytes()
        return
    o, op = run(oracle, cur, refs, None, True, **kw)
    mb_w = w // 16
    sl = slice((stripe[0] if stripe else 0) * mb_w, (stripe[1] if stripe else (h + 15) // 16) * mb_w)
    assert_same(gp[:, sl], op[:, sl], "per-ref")
    assert_same(g[sl], o[sl], "best")
    # planes of an asynchronous build are the planes of a synchronous one
    with cuda.context(width=w, height=h, num_refs=2, async_reference=1, **kw) as a, \
            cuda.context(width=w, height=h, num_refs=2, **kw) as b:
        a.set_reference(1, refs[1]); b.set_reference(1, refs[1])
        for fx, fy in ((0, 0), (2, 1), (3, 3)):
            pa, pb = a.get_subimage(1, fx, fy), b.get_subimage(1, fx, fy)
            if stripe is None:
                assert np.array_equal(pa, pb)


@pytest.mark.parametrize("parts", [2, 3, 4])
@pytest.mark.parametrize("extra", [dict(), dict(mb_row_begin=3, mb_row_end=19), dict(rdopt=1, jm_center=1, max_pred_qpel=160, pred_policy=abi.PRED_PER_MB)])
def test_reference_chunks_of_the_pipelined_host_path(cuda, parts, extra):
    """async_reference + the pipelined host path: jmme_set_reference queues the first chunk of the reference only, the
    others go up with the parts of the search.  Three frames through one context (each frame a new reference and a
    new current picture), two references, stripes, wider borders: the same fields as a context that uploads
    synchronously; a tuning change, a plane read-back and a device-pointer search in between flush the pending
    chunks."""
    import torch
    w, h, R = 160, 336, 8
    frames = [synth.frame_pair(w, h, seed=40 + i, search_range=R, num_refs=2) for i in range(3)]
    kw = dict(width=w, height=h, search_range=R, num_refs=2, qp=28, subpel=1, **extra)
    n_mb = (w // 16) * (h // 16)
    pred = synth.random_pred(2, n_mb, 1, seed=3, max_qpel=150) if "pred_policy" in extra else None
    with cuda.context(async_reference=1, tuning=dict(pipe_parts=parts), **kw) as a, cuda.context(**kw) as b:
        for i, (cur, refs) in enumerate(frames):
            for r in range(2):
                a.set_reference(r, refs[r]); b.set_reference(r, refs[r])
            if i == 1:                                           # read a plane back while chunks are pending
                assert np.array_equal(a.get_subimage(1, 2, 3), b.get_subimage(1, 2, 3)) or "mb_row_begin" in extra
                for r in range(2):
                    a.set_reference(r, refs[r])
            if i == 2:                                           # a tuning change while chunks are pending
                a.set_tuning(pipe_parts=parts % 3 + 2)
            ga, gb = a.search_frame(cur, pred), b.search_frame(cur, pred)
            assert ga.tobytes() == gb.tobytes(), (i, parts, extra)
        # pending chunks and a device-pointer search
        cur, refs = frames[0]
        for r in range(2):
            a.set_reference(r, refs[r]); b.set_reference(r, refs[r])
        dcur = torch.from_numpy(cur).cuda()
        outs = []
        for ctx in (a, b):
            out = torch.zeros(n_mb * abi.MBRESULT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
            dp = torch.from_numpy(pred).cuda() if pred is not None else None
            cuda.check(cuda.dll.jmme_search_frame_dev(ctx.handle, ctypes.c_void_p(dcur.data_ptr()), w,
                                                      ctypes.c_void_p(dp.data_ptr()) if dp is not None else None,
                                                      ctypes.c_void_p(out.data_ptr()), None, None), ctx.handle)
            torch.cuda.synchronize()
            outs.append(out.cpu().numpy().tobytes())
        assert outs[0] == outs[1]


@pytest.mark.parametrize("nref,subpel,median", [(1, 1, False), (2, 1, False), (1, 0, False), (1, 1, True), (2, 1, True)])
def test_fused_peer_stores(cuda, nref, subpel, median):
    """jmme_set_peer_fields_dev: the kernel that writes a record also stores it into the peer buffers (here two more
    buffers on the same device stand in for the peers; NULL and the output buffer itself are skipped)."""
    import torch
    from jmme.torch_api import DeviceSearch
    w, h, R = 96, 112, 8
    cur, refs = synth.frame_pair(w, h, seed=2, search_range=R, num_refs=nref)
    kw = dict(width=w, height=h, search_range=R, num_refs=nref, subpel=subpel, qp=30, mb_row_begin=2, mb_row_end=6)
    if median:
        kw.update(pred_policy=abi.PRED_MEDIAN, slice_rows=2)
    ds = DeviceSearch(cuda, **kw)
    for i, r in enumerate(refs):
        ds.set_reference(i, torch.from_numpy(r).cuda())
    peers = [torch.full_like(ds.out, 0xAB) for _ in range(2)]
    ds.set_peer_fields([peers[0].data_ptr(), 0, ds.out.data_ptr(), peers[1].data_ptr()])
    out = ds.search(torch.from_numpy(cur).cuda())
    torch.cuda.synchronize()
    mb_w = w // 16
    lo, hi = 2 * mb_w, 6 * mb_w
    for p in peers:
        assert torch.equal(p[lo:hi], out[lo:hi])                       # the stripe landed in every peer
        assert bool((p[:lo] == 0xAB).all()) and bool((p[hi:] == 0xAB).all())   # and nothing else was touched
    n0 = ds.launch_count()
    ds.set_peer_fields([])                                             # off again: peers keep their contents
    for p in peers:
        p.fill_(0xCD)
    ds.search(torch.from_numpy(cur).cuda())
    torch.cuda.synchronize()
    assert all(bool((p == 0xCD).all()) for p in peers) and ds.launch_count() > n0
    ds.close()
