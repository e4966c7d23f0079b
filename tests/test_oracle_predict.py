"""MV prediction (a3) and the ME-only field commit: oracle against independent restatements, hand cases
of H.264 8.4.1.3, and the closed predictor -> search -> commit loop over two frames."""
import itertools

import numpy as np
import pytest

import refimpl
from jmme import abi, synth


def test_median_rules_by_hand(oracle):
    P = oracle.set_motion_vector_predictor
    A, B, Cc = (4, -8, 0, 1), (12, 0, 0, 1), (-4, 20, 0, 1)
    assert P(1, 0, 0, A, B, Cc) == (4, 0)                                     # plain median
    assert P(1, 0, 0, A, (12, 0, 1, 1), (-4, 20, 2, 1)) == (4, -8)           # only A has the same reference
    assert P(1, 0, 0, A, (0, 0, -1, 0), (0, 0, -1, 0)) == (4, -8)             # B, C unavailable -> A
    assert P(1, 0, 0, (0, 0, -1, 0), B, Cc) == (0, 0)                          # median(0, 12, -4), median(0,0,20)
    assert P(1, 0, 1, A, B, Cc) == (4, 0)                                     # nobody matches ref 1 -> median
    assert P(2, 0, 0, A, B, Cc) == (12, 0) and P(2, 1, 0, A, B, Cc) == (4, -8)    # 16x8 upper: B, lower: A
    assert P(3, 0, 0, A, B, Cc) == (4, -8) and P(3, 1, 0, A, B, Cc) == (-4, 20)   # 8x16 left: A, right: C
    assert P(2, 0, 0, A, (12, 0, 1, 1), Cc) == (4, 0)                          # directional ref mismatch -> median
    assert P(1, 0, 0, (9, 9, -1, 1), (0, 0, -1, 0), (0, 0, -1, 0)) == (0, 0)    # intra A is available but has no MV


def test_predictor_function_exhaustive_against_restatement(oracle):
    rng = np.random.default_rng(5)
    for _ in range(4000):
        nb = [(int(rng.integers(-64, 65)), int(rng.integers(-64, 65)), int(rng.integers(-1, 3)), int(rng.integers(0, 2)))
              for _ in range(3)]
        t, part, ref = int(rng.integers(1, 8)), int(rng.integers(0, 2)), int(rng.integers(0, 3))
        assert oracle.set_motion_vector_predictor(t, part, ref, *nb) == refimpl.mv_predict(t, part, ref, *nb)
    for av in itertools.product((0, 1), repeat=3):                              # every availability pattern
        nb = [(5 * (i + 1), -3 * (i + 1), 0, a) for i, a in enumerate(av)]
        for t in (1, 2, 3, 4):
            for part in (0, 1):
                assert oracle.set_motion_vector_predictor(t, part, 0, *nb) == refimpl.mv_predict(t, part, 0, *nb)


@pytest.mark.parametrize("mask", [abi.MASK_ALL, 0x92, 0x0E])
def test_commit_and_predict_frame_against_restatement(oracle, mask):
    w, h, R = 64, 48, 6
    cur, refs = synth.frame_pair(w, h, seed=31, search_range=R, num_refs=2)
    with oracle.context(width=w, height=h, search_range=R, num_refs=2, qp=30, blocktype_mask=mask) as ctx:
        for i, r in enumerate(refs):
            ctx.set_reference(i, r)
        res = ctx.search_frame(cur)
        mv4, ref4, mode = ctx.commit_field(res)
        e_mv, e_ref, e_mode = refimpl.commit_field(res, ctx.mb_w, ctx.mb_h, mask)
        assert np.array_equal(mode, e_mode) and np.array_equal(mv4, e_mv) and np.array_equal(ref4, e_ref)
        assert set(np.unique(mode[:, 0])) <= {1, 2, 3, 8}
        ref4[1, 2] = -1                                                         # an intra cell
        pred = ctx.predict_frame(mv4, ref4)
        assert np.array_equal(pred, refimpl.predict_frame(mv4, ref4, ctx.mb_w, ctx.mb_h, 2))
        assert np.all(pred[:, 0, 0] == 0)                                       # first MB: no neighbours at all


def test_closed_loop_two_passes(oracle):
    """pass 1 with zero predictors -> commit -> median predictors -> pass 2 (PER_BLOCK): runs end to end and
    lowers the total rate+distortion cost on smooth synthetic motion."""
    w, h, R = 96, 64, 8
    cur, refs = synth.frame_pair(w, h, seed=8, search_range=R)
    kw = dict(width=w, height=h, search_range=R, qp=32, rdopt=1)
    with oracle.context(**kw) as c1:
        c1.set_reference(0, refs[0])
        res1 = c1.search_frame(cur)
        mv4, ref4, _ = c1.commit_field(res1)
        pred = c1.predict_frame(mv4, ref4)
    with oracle.context(pred_policy=abi.PRED_PER_BLOCK, **kw) as c2:
        c2.set_reference(0, refs[0])
        res2 = c2.search_frame(cur, pred)
    assert res2["cost"][:, 0].astype(np.int64).sum() <= res1["cost"][:, 0].astype(np.int64).sum()


# ---- JMME_PRED_MEDIAN: the predictor loop closed inside the frame (DESIGN.md §2 "in-frame median") ------
def _median_fixed_point(lib, cur, refs, slice_rows, **kw):
    """The policy is pinned by two facts that determine the result MB by MB in raster order:
    (1) the predictors it used are the in-frame 8.4.1.3 predictors (independent restatement) of the field
        committed from its own result; (2) a PER_BLOCK search with those predictors reproduces the result."""
    n = len(refs)
    with lib.context(pred_policy=abi.PRED_MEDIAN, slice_rows=slice_rows, num_refs=n, **kw) as ctx:
        for i, r in enumerate(refs):
            ctx.set_reference(i, r)
        out, opr = ctx.search_frame(cur, per_ref=True)
        used = ctx.get_predictors()
        mv4, ref4, _ = ctx.commit_field(out)
        mb_w, mb_h = ctx.mb_w, ctx.mb_h
    e_mv, e_ref, _ = refimpl.commit_field(out, mb_w, mb_h, kw.get("blocktype_mask", abi.MASK_ALL))
    assert np.array_equal(mv4, e_mv) and np.array_equal(ref4, e_ref)
    assert np.array_equal(used, refimpl.predict_frame(mv4, ref4, mb_w, mb_h, n, slice_rows=slice_rows))
    with lib.context(pred_policy=abi.PRED_PER_BLOCK, num_refs=n, **kw) as ctx:
        for i, r in enumerate(refs):
            ctx.set_reference(i, r)
        out2, opr2 = ctx.search_frame(cur, used, per_ref=True)
    assert out.tobytes() == out2.tobytes() and opr.tobytes() == opr2.tobytes()
    return out, used


@pytest.mark.parametrize("slice_rows", [0, 1, 2])
def test_in_frame_median_is_the_fixed_point(oracle, slice_rows):
    w, h, R = 80, 64, 6
    cur, refs = synth.frame_pair(w, h, seed=12, search_range=R, num_refs=2)
    out, used = _median_fixed_point(oracle, cur, refs, slice_rows, width=w, height=h, search_range=R, qp=30, subpel=1)
    mb_w = w // 16
    assert np.all(used[:, 0] == 0)                                  # first MB of the frame: nothing to predict from
    if slice_rows == 1:                                             # every row is its own slice: only A exists
        assert np.all(used[:, ::mb_w] == 0)
        # "only A available" (8.4.1.3): the 16x16 predictor is the vector committed for the left MB's
        # cell (3, 0) whenever that cell uses the same reference
        mv4, ref4, _ = refimpl.commit_field(out, mb_w, h // 16)
        for mb in range(1, mb_w):
            if ref4[0, 4 * mb - 1] == 0:
                assert tuple(used[0, mb, 0]) == tuple(mv4[0, 4 * mb - 1])
    assert np.any(used != 0)


def test_in_frame_median_full_search_and_masks(oracle):
    w, h, R = 64, 48, 5
    cur, refs = synth.frame_pair(w, h, seed=4, search_range=R)
    _median_fixed_point(oracle, cur, refs, 0, width=w, height=h, search_range=R, qp=34, rdopt=1,
                        search_mode=abi.SEARCH_FULL)
    _median_fixed_point(oracle, cur, refs, 2, width=w, height=h, search_range=R, qp=26, blocktype_mask=0x92)
    _median_fixed_point(oracle, cur, refs, 0, width=w, height=h, search_range=R, blocktype_mask=abi.MASK_16x16)


def test_in_frame_median_stripes_are_slice_aligned(oracle):
    """Stripes that start and end on slice boundaries reproduce the whole-frame result; others are refused."""
    w, h, R = 64, 96, 4
    cur, refs = synth.frame_pair(w, h, seed=6, search_range=R)
    kw = dict(width=w, height=h, search_range=R, pred_policy=abi.PRED_MEDIAN, slice_rows=2)
    with oracle.context(**kw) as ctx:
        ctx.set_reference(0, refs[0])
        whole = ctx.search_frame(cur)
    got = np.zeros_like(whole)
    for rb, re in ((0, 2), (2, 6)):
        with oracle.context(mb_row_begin=rb, mb_row_end=re, **kw) as ctx:
            ctx.set_reference(0, refs[0])
            part = ctx.search_frame(cur)
            got[rb * 4:re * 4] = part[rb * 4:re * 4]
    assert got.tobytes() == whole.tobytes()
    with pytest.raises(abi.JmmeError) as e:
        oracle.context(mb_row_begin=1, mb_row_end=4, **kw)
    assert e.value.code == abi.ERR_PARAM
    with pytest.raises(abi.JmmeError):
        oracle.context(width=w, height=h, search_range=R, pred_policy=abi.PRED_MEDIAN, slice_rows=0, mb_row_end=3)


def test_in_frame_median_lowers_cost_against_zero_predictors(oracle):
    w, h, R = 96, 64, 8
    cur, refs = synth.frame_pair(w, h, seed=8, search_range=R)
    kw = dict(width=w, height=h, search_range=R, qp=32, rdopt=1)
    with oracle.context(**kw) as c0:
        c0.set_reference(0, refs[0])
        res0 = c0.search_frame(cur)
    with oracle.context(pred_policy=abi.PRED_MEDIAN, **kw) as c1:
        c1.set_reference(0, refs[0])
        res1 = c1.search_frame(cur)
    assert res1["cost"][:, 0].astype(np.int64).sum() < res0["cost"][:, 0].astype(np.int64).sum()


def test_get_predictors_states(oracle):
    w, h, R = 32, 32, 4
    cur, refs = synth.frame_pair(w, h, seed=1, search_range=R)
    with oracle.context(width=w, height=h, search_range=R) as ctx:          # not a median context
        with pytest.raises(abi.JmmeError) as e:
            ctx.get_predictors()
        assert e.value.code == abi.ERR_STATE
    with oracle.context(width=w, height=h, search_range=R, pred_policy=abi.PRED_MEDIAN) as ctx:
        with pytest.raises(abi.JmmeError) as e:                              # nothing searched yet
            ctx.get_predictors()
        assert e.value.code == abi.ERR_STATE
        ctx.set_reference(0, refs[0])
        ctx.search_frame(cur, pred=np.zeros(3, np.int16))                    # `pred` is ignored by this policy
        assert ctx.get_predictors().shape == (1, 4, 41, 2)
    with pytest.raises(abi.JmmeError) as e:
        oracle.context(width=w, height=h, search_range=R, pred_policy=abi.PRED_MEDIAN, slice_rows=-1)
    assert e.value.code == abi.ERR_PARAM
    with pytest.raises(abi.JmmeError):
        oracle.context(width=w, height=h, search_range=R, pred_policy=4)
