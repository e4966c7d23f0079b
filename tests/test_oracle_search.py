"""Oracle search behaviour: 41-block FastFull == brute-force per-block search, tie-break,
(0,0) pre-test, 16x16 bonus, planted motion, sub-pel refinement, multi-ref, stripes."""
import numpy as np
import pytest

import refimpl
from jmme import abi, synth

BLOCKS = abi.block_table()


def run(oracle, cur, refs, pred=None, per_ref=False, **kw):
    h, w = cur.shape
    with oracle.context(width=w, height=h, num_refs=len(refs), **kw) as ctx:
        for i, r in enumerate(refs):
            ctx.set_reference(i, r)
        return ctx.search_frame(cur, pred, per_ref), ctx.lambda_factor, ctx.pad


@pytest.mark.parametrize("rdopt,qp", [(0, 28), (1, 30), (0, 40)])
def test_fastfull_equals_bruteforce_per_block(oracle, rdopt, qp):
    w, h, R = 48, 32, 6
    cur, refs = synth.frame_pair(w, h, seed=3, search_range=R)
    res, f, pad = run(oracle, cur, refs, search_range=R, qp=qp, rdopt=rdopt)
    refp = refimpl.padded(refs[0], pad)
    bonus = 0 if rdopt else refimpl.weighted_cost(f, 16)
    for mb in range(6):
        mbx, mby = mb % 3, mb // 3
        for b, (t, x0, y0, bw, bh) in enumerate(BLOCKS):
            mx, my, c = refimpl.brute_search(cur, refp, pad, 16 * mbx + x0, 16 * mby + y0, bw, bh, 0, 0, 0, 0, R, f,
                                             bonus if t == 1 else 0, pretest=not rdopt)
            assert tuple(res[mb]["mv"][b]) == (4 * mx, 4 * my), (mb, b)
            assert res[mb]["cost"][b] == c + (refimpl.weighted_cost(f, 1) if rdopt else 0)
            assert res[mb]["ref_idx"][b] == 0


def test_per_block_predictors_and_centre_clamp(oracle):
    w, h, R = 32, 32, 5
    cur, refs = synth.frame_pair(w, h, seed=9, search_range=R)
    pred = synth.random_pred(1, 4, 41, seed=4, max_qpel=60)      # centre pred/4 up to 15 > R: clamped
    res, f, pad = run(oracle, cur, refs, pred=pred, search_range=R, qp=33, rdopt=1, pred_policy=abi.PRED_PER_BLOCK)
    refp = refimpl.padded(refs[0], pad)
    for mb in range(4):
        mbx, mby = mb % 2, mb // 2
        p16 = pred[0, mb, 0]
        cx = int(np.clip(int(p16[0] / 4), -R, R))           # C division truncates toward zero
        cy = int(np.clip(int(p16[1] / 4), -R, R))
        for b, (t, x0, y0, bw, bh) in enumerate(BLOCKS):
            px, py = int(pred[0, mb, b, 0]), int(pred[0, mb, b, 1])
            mx, my, c = refimpl.brute_search(cur, refp, pad, 16 * mbx + x0, 16 * mby + y0, bw, bh, cx, cy, px, py,
                                             R, f)
            assert tuple(res[mb]["mv"][b]) == (4 * mx, 4 * my)
            assert res[mb]["cost"][b] == c + refimpl.weighted_cost(f, 1)


def test_full_mode_centres_each_block_on_its_predictor(oracle):
    w, h, R = 32, 16, 4
    cur, refs = synth.frame_pair(w, h, seed=11, search_range=R)
    pred = synth.random_pred(1, 2, 41, seed=5, max_qpel=24)
    res, f, pad = run(oracle, cur, refs, pred=pred, search_range=R, qp=28, rdopt=0, pred_policy=abi.PRED_PER_BLOCK,
                      search_mode=abi.SEARCH_FULL)
    refp = refimpl.padded(refs[0], pad)
    bonus = refimpl.weighted_cost(f, 16)
    for mb in range(2):
        for b, (t, x0, y0, bw, bh) in enumerate(BLOCKS):
            px, py = int(pred[0, mb, b, 0]), int(pred[0, mb, b, 1])
            cx, cy = int(np.clip(int(px / 4), -R, R)), int(np.clip(int(py / 4), -R, R))
            mx, my, c = refimpl.brute_search(cur, refp, pad, 16 * mb + x0, y0, bw, bh, cx, cy, px, py, R, f,
                                             bonus if t == 1 else 0)
            assert tuple(res[mb]["mv"][b]) == (4 * mx, 4 * my)
            assert res[mb]["cost"][b] == c


def test_constant_frames_all_ties(oracle):
    """Every candidate has SAD 0: winner is the cheapest rate; with a zero predictor that is MV (0,0);
    with predictor (4*2, 4*-1) FastFull/!rdopt still returns (0,0) only if it is not beaten on rate."""
    cur = np.full((16, 16), 90, np.uint8)
    res, f, _ = run(oracle, cur, [cur], search_range=4, qp=28, rdopt=0)
    assert np.all(res["mv"] == 0)
    assert res[0]["cost"][0] == refimpl.weighted_cost(f, 2) - refimpl.weighted_cost(f, 16)   # 16x16 bonus
    assert np.all(res[0]["cost"][1:] == refimpl.weighted_cost(f, 2))
    pred = np.array([8, -4], np.int16).reshape(1, 1, 1, 2)
    res, f, _ = run(oracle, cur, [cur], pred=pred, search_range=4, qp=28, rdopt=1, pred_policy=abi.PRED_PER_MB)
    assert np.all(res["mv"] == [8, -4])                      # rate-optimal candidate = predictor
    # spiral tie-break: rdopt, zero SAD, lambda tiny so every rate rounds to 0 -> first spiral position (centre)
    res, _, _ = run(oracle, cur, [cur], pred=pred, search_range=4, lambda_factor=1, rdopt=1,
                    pred_policy=abi.PRED_PER_MB)
    assert np.all(res["mv"] == [8, -4])                      # centre = pred/4 = (2,-1) is spiral pos 0
    res, _, _ = run(oracle, cur, [cur], pred=pred, search_range=4, lambda_factor=1, rdopt=0,
                    pred_policy=abi.PRED_PER_MB)
    assert np.all(res["mv"] == 0)                            # !rdopt: (0,0) is tested first and keeps ties


def test_planted_integer_motion_is_found(oracle):
    R = 8
    ref = synth.gen_luma(96, 64, 21, "noise")
    for (dx, dy) in [(3, -2), (-8, 8), (0, 5)]:
        cur = np.roll(ref, (-dy, -dx), axis=(0, 1))         # cur(x,y) = ref(x+dx, y+dy)
        res, f, _ = run(oracle, cur, [ref], search_range=R, qp=20, rdopt=1)
        inner = [mby * 6 + mbx for mby in range(1, 3) for mbx in range(1, 5)]
        for mb in inner:
            assert np.all(res[mb]["mv"] == [4 * dx, 4 * dy])
            exp = refimpl.weighted_cost(f, refimpl.se_bits(4 * dx) + refimpl.se_bits(4 * dy)) + refimpl.weighted_cost(f, 1)
            assert np.all(res[mb]["cost"] == exp)


@pytest.mark.parametrize("hadamard,rnd", [(1, 0), (1, 1), (0, 0)])
def test_subpel_refinement_against_python_restatement(oracle, hadamard, rnd):
    w, h, R = 32, 32, 4
    cur, refs = synth.frame_pair(w, h, seed=6, search_range=R)
    kw = dict(search_range=R, qp=30, rdopt=0, use_hadamard=hadamard, satd_round=rnd)
    (res_i), f, pad = run(oracle, cur, refs, subpel=0, **kw)
    (res_q), _, _ = run(oracle, cur, refs, subpel=1, **kw)
    it = refimpl.Interp(refs[0])
    sp = refimpl.spiral(1)
    bonus16 = refimpl.weighted_cost(f, 16)

    def dist(bx, by, bw, bh, qx, qy):
        blk = cur[by:by + bh, bx:bx + bw].astype(np.int64)
        r = np.array([[it.sample(4 * (bx + i) + qx, 4 * (by + j) + qy) for i in range(bw)] for j in range(bh)])
        d = blk - r
        if not hadamard:
            return int(np.abs(d).sum())
        return sum(refimpl.satd4x4(d[j:j + 4, i:i + 4], rnd) for j in range(0, bh, 4) for i in range(0, bw, 4))

    for mb in (0, 3):
        mbx, mby = mb % 2, mb // 2
        for b, (t, x0, y0, bw, bh) in enumerate(BLOCKS):
            if b % 3 and t > 4:
                continue                                      # subsample small blocks for speed
            mvx, mvy = (int(v) for v in res_i[mb]["mv"][b])
            mn = None if hadamard else int(res_i[mb]["cost"][b])
            for step in (2, 1):
                ox, oy, best = mvx, mvy, 0
                for pos in range(0 if (step == 2 and hadamard) else 1, 9):
                    qx, qy = ox + step * sp[pos][0], oy + step * sp[pos][1]
                    c = refimpl.weighted_cost(f, refimpl.se_bits(qx) + refimpl.se_bits(qy))
                    c += dist(16 * mbx + x0, 16 * mby + y0, bw, bh, qx, qy)
                    if t == 1 and qx == 0 and qy == 0:
                        c -= bonus16
                    if mn is None or c < mn:
                        mn, best = c, pos
                mvx, mvy = ox + step * sp[best][0], oy + step * sp[best][1]
            assert tuple(res_q[mb]["mv"][b]) == (mvx, mvy), (mb, b)
            assert res_q[mb]["cost"][b] == mn


def test_planted_quarter_pel_motion(oracle):
    """cur sampled from the oracle's own plane (1,3) shifted by (2,-1): SATD 0 at mv (4*2+1, 4*-1+3)."""
    R = 4
    ref = synth.gen_luma(64, 48, 13, "texture")
    planes = oracle.get_sub_images_luma(ref, 8)
    cur = planes[3, 1, 8 - 1:8 - 1 + 48, 8 + 2:8 + 2 + 64]
    res, f, _ = run(oracle, np.ascontiguousarray(cur), [ref], search_range=R, qp=12, rdopt=1, subpel=1)
    mb = 1 * 4 + 1
    assert np.all(res[mb]["mv"] == [9, -1])
    assert np.all(res[mb]["cost"] == refimpl.weighted_cost(f, refimpl.se_bits(9) + refimpl.se_bits(-1)) +
                  refimpl.weighted_cost(f, 1))


def test_multi_ref_best_ref_and_ref_cost(oracle):
    w, h, R = 48, 32, 4
    cur, refs = synth.frame_pair(w, h, seed=2, search_range=R, num_refs=3)
    refs[2] = cur.copy()                                     # ref 2 is a perfect match
    for rdopt in (0, 1):
        (res, per), f, _ = run(oracle, cur, refs, per_ref=True, search_range=R, qp=26, rdopt=rdopt)
        lam = f >> 16
        rc = [refimpl.weighted_cost(f, refimpl.ue_bits(r)) if rdopt else (2 * lam if r else 0) for r in range(3)]
        tot = np.stack([per[r]["cost"].astype(np.int64) + rc[r] for r in range(3)])     # [ref][mb][41]
        best = tot.argmin(0)                                 # first minimum = lowest ref on ties
        assert np.array_equal(res["ref_idx"], best)
        assert np.array_equal(res["cost"], tot.min(0))
        for r in range(3):
            sel = best == r
            assert np.array_equal(res["mv"][sel], per[r]["mv"][sel])
        assert np.all(per[2]["mv"] == 0) and (res["ref_idx"] == 2).mean() > 0.9
        # each per-ref result equals a single-ref search on that reference (minus its ref cost, bonus only on ref 0)
        single, _, _ = run(oracle, cur, [refs[0]], search_range=R, qp=26, rdopt=rdopt)
        assert np.array_equal(single["mv"], per[0]["mv"])
        assert np.array_equal(single["cost"], per[0]["cost"] + rc[0])


def test_blocktype_mask_and_stripes(oracle):
    w, h, R = 64, 64, 4
    cur, refs = synth.frame_pair(w, h, seed=8, search_range=R)
    full, _, _ = run(oracle, cur, refs, search_range=R)
    only16, _, _ = run(oracle, cur, refs, search_range=R, blocktype_mask=abi.MASK_16x16)
    assert np.array_equal(only16["mv"][:, 0], full["mv"][:, 0])
    assert np.array_equal(only16["cost"][:, 0], full["cost"][:, 0])
    assert np.all(only16["cost"][:, 1:] == abi.INT32_MAX) and np.all(only16["ref_idx"][:, 1:] == -1)
    top, _, _ = run(oracle, cur, refs, search_range=R, mb_row_begin=0, mb_row_end=1)
    bot, _, _ = run(oracle, cur, refs, search_range=R, mb_row_begin=1, mb_row_end=4)
    assert top[:4].tobytes() == full[:4].tobytes() and bot[4:].tobytes() == full[4:].tobytes()
    assert not top[4:].tobytes().strip(b"\0")                # rows outside the stripe untouched


def test_non_multiple_of_16_frames_are_padded_by_replication(oracle):
    cur, refs = synth.frame_pair(40, 24, seed=5, search_range=4)
    a, _, _ = run(oracle, cur, refs, search_range=4, subpel=1)
    ext = lambda im: np.pad(im, ((0, 8), (0, 8)), mode="edge")          # noqa: E731
    b, _, _ = run(oracle, ext(cur), [ext(refs[0])], search_range=4, subpel=1)
    assert a.tobytes() == b.tobytes() and len(a) == 6


def test_leaf_entry_points_agree_with_frame_path(oracle):
    w, h, R = 32, 32, 5
    cur, refs = synth.frame_pair(w, h, seed=14, search_range=R)
    res, f, pad = run(oracle, cur, refs, search_range=R, qp=28, rdopt=0, subpel=1)
    resi, _, _ = run(oracle, cur, refs, search_range=R, qp=28, rdopt=0, subpel=0)
    refp = refimpl.padded(refs[0], pad)
    planes = oracle.get_sub_images_luma(refs[0], pad)
    bonus = refimpl.weighted_cost(f, 16)
    mb, mbx, mby = 3, 1, 1
    sad = oracle.setup_fast_full_pel_search(cur[16:32, 16:32], refp, pad, mbx, mby, 0, 0, R, bonus)
    assert sad.shape == (41, 121)
    for b, (t, x0, y0, bw, bh) in enumerate(BLOCKS):
        mx, my, c = oracle.fast_full_pel_block_motion_search(sad[b], R, 0, 0, 0, 0, f, 1)
        assert (4 * mx, 4 * my, c) == (*resi[mb]["mv"][b], resi[mb]["cost"][b])
        fx, fy, fc = oracle.full_pel_block_motion_search(cur, refp, pad, 16 + x0, 16 + y0, bw, bh, 0, 0, R, f,
                                                         bonus if t == 1 else 0)
        assert fc == c                                          # same minimum; tie winners may differ (pre-test)
        qx, qy, qc = oracle.sub_pel_block_motion_search(cur, planes, pad, 16 + x0, 16 + y0, bw, bh, 0, 0, f,
                                                        (4 * mx, 4 * my), c, 1, 0, bonus if t == 1 else 0)
        assert (qx, qy, qc) == (*res[mb]["mv"][b], res[mb]["cost"][b])
