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
