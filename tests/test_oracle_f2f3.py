"""Round-2 oracle features against independent restatements (tests/refimpl.py): the JM >= 12 scaled-up cost
domain (a2), SSE and 8x8-Hadamard distortion with per-stage metrics (f2), chroma ME with eighth-pel bilinear
chroma samples (f3)."""
import numpy as np
import pytest

import refimpl
from jmme import abi, synth

BLOCKS = abi.block_table()


def test_hadamard_sad8x8_leaf(oracle):
    rng = np.random.default_rng(5)
    d = rng.integers(-255, 256, size=(300, 64)).astype(np.int16)
    d[0] = 255
    d[1] = np.tile([255, -255], 32)
    d[2] = 0
    d[3] = 1                                                   # DC only: 64 -> (64 + 2) >> 2 = 16
    for rnd in (0, 1):
        got = oracle.hadamard_sad8x8(d, rnd)
        assert got[3] == 16 and got[2] == 0
        assert list(got) == [refimpl.satd8x8(x, rnd) for x in d]
    # an 8x8 transform is not the sum of its four 4x4 transforms
    x = d[7].reshape(8, 8)
    four = sum(refimpl.satd4x4(x[i:i + 4, j:j + 4]) for i in (0, 4) for j in (0, 4))
    assert refimpl.satd8x8(x) != four


def test_get_sub_images_chroma_per_sample(oracle):
    img = synth.gen_luma(24, 16, 3, "noise")
    pad = 5
    planes = oracle.get_sub_images_chroma(img, pad)
    assert planes.shape == (8, 8, 16 + 2 * pad, 24 + 2 * pad)
    rng = np.random.default_rng(1)
    for _ in range(600):
        x, y = int(rng.integers(0, 24 + 2 * pad)), int(rng.integers(0, 16 + 2 * pad))
        xf, yf = int(rng.integers(0, 8)), int(rng.integers(0, 8))
        assert planes[yf, xf, y, x] == refimpl.chroma_sample(img, 8 * (x - pad) + xf, 8 * (y - pad) + yf)
    assert np.array_equal(planes[0, 0], np.pad(img, pad, mode="edge"))


def lam_of(qp, rdopt):
    q = min(max(qp - 12, 0), 39)
    tab = [1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 4, 4, 4, 5, 6, 6, 7, 8, 9, 10, 11, 13, 14, 16, 18, 20, 23, 25, 29, 32, 36, 40, 45, 51,
           57, 64, 72, 81, 91]
    return (0.85 * 2.0 ** (q / 3.0)) ** 0.5 if rdopt else float(tab[q])


CASES = [
    # (params, StageSpec metrics, t8, chroma)
    (dict(cost_domain=1), (0, 0, 0), False, False),
    (dict(cost_domain=1, subpel=1, rdopt=1, qp=31), (0, 2, 2), False, False),
    (dict(cost_domain=1, subpel=1, qp=24, use_hadamard=0), (0, 0, 0), False, False),
    (dict(me_distortion=1, me_distortion_fpel=1, me_distortion_hpel=1, me_distortion_qpel=1, subpel=1, qp=26), (1, 1, 1), False, False),
    (dict(me_distortion=1, me_distortion_fpel=0, me_distortion_hpel=2, me_distortion_qpel=2, subpel=1, transform8x8=1, satd_round=1,
          rdopt=1, qp=30), (0, 2, 2), True, False),
    (dict(me_distortion=1, me_distortion_fpel=1, me_distortion_hpel=0, me_distortion_qpel=2, subpel=1, transform8x8=1,
          cost_domain=1, qp=29), (1, 0, 2), True, False),
    (dict(subpel=1, chroma_me=1, qp=27), (0, 2, 2), False, True),
    (dict(subpel=1, chroma_me=1, use_hadamard=0, cost_domain=1, rdopt=1, qp=33), (0, 0, 0), False, True),
    (dict(me_distortion=1, me_distortion_fpel=0, me_distortion_hpel=1, me_distortion_qpel=2, subpel=1, chroma_me=1, qp=25,
          satd_round=1), (0, 1, 2), False, True),
]


@pytest.mark.parametrize("kw,metrics,t8,chroma", CASES)
@pytest.mark.parametrize("mode", [abi.SEARCH_FASTFULL, abi.SEARCH_FULL])
def test_stage_metrics_cost_domains_and_chroma_against_the_restatement(oracle, kw, metrics, t8, chroma, mode):
    w, h, R = 32, 16, 2
    cur, refs = synth.frame_pair(w, h, seed=7, search_range=R)
    cur_c = [synth.gen_luma(w // 2, h // 2, 11 + k, "texture") for k in range(2)]
    ref_c = [synth.gen_luma(w // 2, h // 2, 21 + k, "texture") for k in range(2)]
    pred = synth.random_pred(1, 2, 41, seed=2, max_qpel=9)
    rdopt, qp = kw.get("rdopt", 0), kw.get("qp", 28)
    with oracle.context(width=w, height=h, search_range=R, pred_policy=abi.PRED_PER_BLOCK, search_mode=mode, **kw) as ctx:
        ctx.set_reference(0, refs[0])
        if chroma:
            ctx.set_reference_chroma(0, *ref_c)
            ctx.set_current_chroma(*cur_c)
        res = ctx.search_frame(cur, pred)
    spec = refimpl.StageSpec(lam_of(qp, rdopt), kw.get("cost_domain", 0), metrics, t8, kw.get("satd_round", 0), chroma)
    last = 2 if kw.get("subpel") else 0
    refc = spec.rate(last, 1) if rdopt else 0                     # reference 0: ue(0) = 1 bit with rdopt, free without
    for mb in range(2):
        p16 = pred[0, mb, 0]
        for b, (t, x0, y0, bw, bh) in enumerate(BLOCKS):
            if b not in (0, 1, 4, 6, 12, 20, 27, 40):           # one or two blocks of every blocktype
                continue
            px, py = int(pred[0, mb, b, 0]), int(pred[0, mb, b, 1])
            own = (px, py) if mode == abi.SEARCH_FULL else (int(p16[0]), int(p16[1]))
            cx, cy = int(np.clip(int(own[0] / 4), -R, R)), int(np.clip(int(own[1] / 4), -R, R))
            mvx, mvy, c = refimpl.block_search(spec, cur, refs[0], 16 * mb + x0, y0, bw, bh, cx, cy, px, py, R,
                                               bonus16=(t == 1 and not rdopt),
                                               pretest=(not rdopt and mode == abi.SEARCH_FASTFULL),
                                               subpel=kw.get("subpel", 0), cur_c=cur_c, ref_c=ref_c)
            assert tuple(res[mb]["mv"][b]) == (mvx, mvy), (mb, b, kw)
            assert res[mb]["cost"][b] == c + refc, (mb, b, kw)


def test_legacy_parameters_are_the_zero_settings(oracle):
    """me_distortion = 0 / cost_domain = 0 / transform8x8 = 0 / chroma_me = 0 reproduce the round-1 results: the
    explicit form of the legacy metrics gives the same field."""
    w, h, R = 48, 32, 4
    cur, refs = synth.frame_pair(w, h, seed=5, search_range=R, num_refs=2)
    outs = []
    for kw in (dict(use_hadamard=1), dict(me_distortion=1, me_distortion_fpel=0, me_distortion_hpel=2, me_distortion_qpel=2)):
        with oracle.context(width=w, height=h, search_range=R, num_refs=2, subpel=1, qp=30, **kw) as ctx:
            for i, r in enumerate(refs):
                ctx.set_reference(i, r)
            outs.append(ctx.search_frame(cur))
    assert outs[0].tobytes() == outs[1].tobytes()


def test_new_parameter_errors(oracle):
    for kw, code in ((dict(cost_domain=2), abi.ERR_PARAM), (dict(me_distortion=1, me_distortion_hpel=3), abi.ERR_PARAM),
                     (dict(me_distortion=1, me_distortion_fpel=2), abi.ERR_UNSUPPORTED), (dict(chroma_me=1), abi.ERR_PARAM),
                     (dict(transform8x8=2), abi.ERR_PARAM)):
        with pytest.raises(abi.JmmeError) as e:
            oracle.context(width=32, height=32, **kw)
        assert e.value.code == code, kw
    with oracle.context(width=32, height=32, search_range=2, subpel=1, chroma_me=1) as ctx:
        ctx.set_reference(0, np.zeros((32, 32), np.uint8))
        with pytest.raises(abi.JmmeError) as e:
            ctx.search_frame(np.zeros((32, 32), np.uint8))
        assert e.value.code == abi.ERR_STATE
    with oracle.context(width=32, height=32, search_range=2) as ctx:
        with pytest.raises(abi.JmmeError) as e:
            ctx.set_current_chroma(np.zeros((16, 16), np.uint8), np.zeros((16, 16), np.uint8))
        assert e.value.code == abi.ERR_STATE


@pytest.mark.parametrize("kw,metric0", [(dict(subpel=1, qp=30), 0), (dict(subpel=0, rdopt=1, qp=26, cost_domain=1), 0),
                                        (dict(subpel=1, me_distortion=1, me_distortion_fpel=1, me_distortion_hpel=2, me_distortion_qpel=2), 1)])
def test_bipred_refinement_against_the_restatement(oracle, kw, metric0):
    """jmme_search_frame_bipred: two uni-directional searches give the start pair, the refinement is compared block by
    block with the per-sample restatement (quarter-pel fixed vectors, both rates, spiral order, strict <)."""
    w, h, R, rng, iters = 32, 32, 3, 2, 3
    cur, refs = synth.frame_pair(w, h, seed=9, search_range=R, num_refs=2)
    ref1 = synth.frame_pair(w, h, seed=10, search_range=R)[1][0]
    rdopt, qp = kw.get("rdopt", 0), kw.get("qp", 28)
    pred0 = synth.random_pred(2, 4, 41, seed=3, max_qpel=7)
    pred1 = synth.random_pred(1, 4, 41, seed=4, max_qpel=7)
    with oracle.context(width=w, height=h, search_range=R, num_refs=2, pred_policy=abi.PRED_PER_BLOCK, **kw) as ctx, \
            oracle.context(width=w, height=h, search_range=R, num_refs=1, pred_policy=abi.PRED_PER_BLOCK, **kw) as ctx1:
        for i, r in enumerate(refs):
            ctx.set_reference(i, r)
        ctx1.set_reference(0, ref1)
        l0, l1 = ctx.search_frame(cur, pred0), ctx1.search_frame(cur, pred1)
        ctx.set_reference_l1(ref1)
        got = ctx.search_frame_bipred(cur, l0, l1, pred0, pred1, rng, iters)
        pad = ctx.pad
    spec = refimpl.StageSpec(lam_of(qp, rdopt), kw.get("cost_domain", 0), (metric0, 2, 2))
    assert len(np.unique(l0["ref_idx"])) > 1
    moved = 0
    for mb in range(4):
        for b, (t, x0, y0, bw, bh) in enumerate(BLOCKS):
            if b not in (0, 2, 3, 7, 10, 19, 33):
                continue
            r0 = int(l0[mb]["ref_idx"][b])
            e0, e1, c = refimpl.bipred_block(spec, cur, refs[r0], ref1, 16 * (mb % 2) + x0, 16 * (mb // 2) + y0, bw, bh,
                                             l0[mb]["mv"][b], l1[mb]["mv"][b], [int(v) for v in pred0[r0, mb, b]],
                                             [int(v) for v in pred1[0, mb, b]], rng, iters, pad)
            assert (tuple(got[mb]["mv0"][b]), tuple(got[mb]["mv1"][b]), int(got[mb]["cost"][b]), int(got[mb]["ref0"][b])) == (e0, e1, c, r0), (mb, b)
            moved += e0 != tuple(l0[mb]["mv"][b]) or e1 != tuple(l1[mb]["mv"][b])
    assert moved > 0                                                  # the refinement changes some pairs


def test_bipred_errors_and_masks(oracle):
    w, h, R = 32, 32, 3
    cur, refs = synth.frame_pair(w, h, seed=9, search_range=R)
    with oracle.context(width=w, height=h, search_range=R, blocktype_mask=0x92) as ctx:
        ctx.set_reference(0, refs[0])
        l0 = ctx.search_frame(cur)
        with pytest.raises(abi.JmmeError) as e:
            ctx.search_frame_bipred(cur, l0, l0)
        assert e.value.code == abi.ERR_STATE                          # no list-1 picture yet
        ctx.set_reference_l1(refs[0])
        for bad in (dict(search_range=0), dict(search_range=16), dict(iterations=0), dict(iterations=9)):
            with pytest.raises(abi.JmmeError) as e:
                ctx.search_frame_bipred(cur, l0, l0, **bad)
            assert e.value.code == abi.ERR_PARAM
        out = ctx.search_frame_bipred(cur, l0, l0, search_range=2, iterations=2)
        off = [b for b, blk in enumerate(BLOCKS) if blk[0] not in (1, 4, 7)]
        assert np.all(out["ref0"][:, off] == -1) and np.all(out["cost"][:, off] == abi.INT32_MAX)
        on = [b for b, blk in enumerate(BLOCKS) if blk[0] in (1, 4, 7)]
        assert np.all(out["ref0"][:, on] == 0)
        # identical pictures on both lists: averaging two copies of the best uni-directional block cannot cost more
        # distortion than it, so the pair cost is at most the uni cost plus the second vector's rate
        lf = ctx.lambda_factor
        for mb in range(4):
            for b in on:
                bits = refimpl.se_bits(int(l0[mb]["mv"][b][0])) + refimpl.se_bits(int(l0[mb]["mv"][b][1]))
                assert out[mb]["cost"][b] <= l0[mb]["cost"][b] + refimpl.weighted_cost(lf, bits) + 1


@pytest.mark.parametrize("mode", [abi.SEARCH_FASTFULL, abi.SEARCH_FULL])
@pytest.mark.parametrize("rdopt", [1, 0])
def test_jm_center_rule_against_the_restatement(oracle, mode, rdopt):
    """jm_center = 1 (JM's BlockMotionSearch): with rdopt the window centre is pred/4
    truncated, NOT clamped to +-R — predictors of up to 7 R here, windows far outside the picture (unbounded edge
    replication); without rdopt the clamp stays.  Checked block by block against the per-sample restatement, and
    the replication border grows with max_pred_qpel."""
    w, h, R = 32, 16, 2
    cur, refs = synth.frame_pair(w, h, seed=9, search_range=R)
    pred = synth.random_pred(1, 2, 41, seed=4, max_qpel=56)
    assert np.abs(pred).max() > 4 * 3 * R
    kw = dict(width=w, height=h, search_range=R, pred_policy=abi.PRED_PER_BLOCK, search_mode=mode, subpel=1, qp=30, rdopt=rdopt)
    with oracle.context(jm_center=1, max_pred_qpel=64, **kw) as ctx:
        assert ctx.pad == ((64 // 4 + R + 16 + 15) & ~15 if rdopt else (2 * R + 16 + 15) & ~15)
        ctx.set_reference(0, refs[0])
        res = ctx.search_frame(cur, pred)
        with pytest.raises(abi.JmmeError):                       # beyond max_pred_qpel
            ctx.search_frame(cur, np.full_like(pred, 65))
    with oracle.context(**kw) as ctx:                             # the default rule: clamped always
        ctx.set_reference(0, refs[0])
        clamped = ctx.search_frame(cur, pred)
    assert (res.tobytes() == clamped.tobytes()) == (not rdopt)
    spec = refimpl.StageSpec(lam_of(30, rdopt))
    refc = spec.rate(2, 1) if rdopt else 0
    for mb in range(2):
        p16 = pred[0, mb, 0]
        for b, (t, x0, y0, bw, bh) in enumerate(BLOCKS):
            if b not in (0, 2, 3, 7, 11, 22, 31):
                continue
            px, py = int(pred[0, mb, b, 0]), int(pred[0, mb, b, 1])
            own = (px, py) if mode == abi.SEARCH_FULL else (int(p16[0]), int(p16[1]))
            cx, cy = int(own[0] / 4), int(own[1] / 4)            # C truncation toward zero
            if not rdopt:
                cx, cy = int(np.clip(cx, -R, R)), int(np.clip(cy, -R, R))
            mvx, mvy, c = refimpl.block_search(spec, cur, refs[0], 16 * mb + x0, y0, bw, bh, cx, cy, px, py, R,
                                               bonus16=(t == 1 and not rdopt),
                                               pretest=(not rdopt and mode == abi.SEARCH_FASTFULL), subpel=1)
            assert tuple(res[mb]["mv"][b]) == (mvx, mvy), (mb, b)
            assert res[mb]["cost"][b] == c + refc, (mb, b)
