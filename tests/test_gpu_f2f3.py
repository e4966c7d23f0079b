"""GPU parity of the round-2 features against the oracle, through the C ABI: the JM >= 12 scaled-up cost domain
(a2), SSE / 8x8-Hadamard distortion with per-stage metrics (f2), chroma ME (f3).  Bit-exact records."""
import numpy as np
import pytest

from jmme import abi, synth
from test_gpu_parity import assert_same

pytestmark = pytest.mark.gpu


def chroma_pair(w, h, seed):
    """Two chroma pictures correlated with a luma-like texture so that chroma changes the decisions."""
    return [synth.gen_luma((w + 1) // 2, (h + 1) // 2, seed + k, "texture") for k in range(2)]


def run(lib, cur, refs, pred=None, per_ref=False, chroma=None, tuning=None, **kw):
    h, w = cur.shape
    is_cuda = lib.backend().startswith("cuda")
    with lib.context(width=w, height=h, num_refs=len(refs), tuning=tuning if is_cuda else None, **kw) as ctx:
        for i, r in enumerate(refs):
            ctx.set_reference(i, r)
        if chroma is not None:
            cur_c, ref_c = chroma
            for i in range(len(refs)):
                ctx.set_reference_chroma(i, *ref_c[i])
            ctx.set_current_chroma(*cur_c)
        out = ctx.search_frame(cur, pred, per_ref)
        return out, (ctx.last_kernel() if is_cuda else None)


def test_new_leaves(cuda, oracle):
    rng = np.random.default_rng(11)
    d = rng.integers(-255, 256, size=(2048, 64)).astype(np.int16)
    d[0] = 255
    d[1] = np.tile([255, -255], 32)
    for rnd in (0, 1):
        assert np.array_equal(cuda.hadamard_sad8x8(d, rnd), oracle.hadamard_sad8x8(d, rnd))
    for (w, h, pad, kind) in ((24, 16, 5, "noise"), (40, 22, 9, "texture"), (8, 8, 3, "checker")):
        img = synth.gen_luma(w, h, 3, kind)
        assert np.array_equal(cuda.get_sub_images_chroma(img, pad), oracle.get_sub_images_chroma(img, pad))


SSE3 = dict(me_distortion=1, me_distortion_fpel=1, me_distortion_hpel=1, me_distortion_qpel=1)
CASES = [
    # cost domain 1 (a2): integer only, with sub-pel (SATD / SAD), rdopt and not, per-block predictors come below
    dict(cost_domain=1),
    dict(cost_domain=1, rdopt=1, qp=33),
    dict(cost_domain=1, subpel=1, qp=24),
    dict(cost_domain=1, subpel=1, use_hadamard=0, rdopt=1, qp=30),
    dict(cost_domain=1, subpel=1, blocktype_mask=abi.MASK_16x16),
    dict(cost_domain=1, subpel=1, blocktype_mask=0x92, satd_round=1),
    dict(cost_domain=1, search_mode=abi.SEARCH_FULL, qp=36),
    # SSE (f2)
    dict(subpel=1, qp=26, **SSE3),
    dict(subpel=0, rdopt=1, qp=35, **SSE3),
    dict(subpel=1, cost_domain=1, qp=22, **SSE3),
    dict(subpel=1, blocktype_mask=abi.MASK_16x16, **SSE3),
    dict(subpel=1, search_mode=abi.SEARCH_FULL, **SSE3),
    # mixed metrics: every stage restarts
    dict(me_distortion=1, me_distortion_fpel=1, me_distortion_hpel=0, me_distortion_qpel=2, subpel=1, qp=29),
    dict(me_distortion=1, me_distortion_fpel=0, me_distortion_hpel=1, me_distortion_qpel=0, subpel=1, rdopt=1, qp=27),
    dict(me_distortion=1, me_distortion_fpel=0, me_distortion_hpel=2, me_distortion_qpel=1, subpel=1, cost_domain=1),
    # 8x8 Hadamard (f2)
    dict(subpel=1, transform8x8=1),
    dict(subpel=1, transform8x8=1, satd_round=1, rdopt=1, qp=31),
    dict(subpel=1, transform8x8=1, cost_domain=1, satd_round=1),
    dict(subpel=1, transform8x8=1, blocktype_mask=0x1E),                 # only the blocktypes that use it
    dict(subpel=1, transform8x8=1, use_hadamard=0),                      # no Hadamard stage: the flag is inert
]


@pytest.mark.parametrize("kw", CASES)
@pytest.mark.parametrize("policy", [abi.PRED_ZERO, abi.PRED_PER_BLOCK])
def test_cost_domain_sse_hadamard8_match_oracle(cuda, oracle, kw, policy):
    for (w, h, R, kind, seed) in ((64, 48, 5, "texture", 1), (52, 38, 9, "noise", 2)):
        cur, refs = synth.frame_pair(w, h, seed=seed, search_range=R, kind=kind, num_refs=2)
        n_mb = ((w + 15) // 16) * ((h + 15) // 16)
        pred = None if policy == abi.PRED_ZERO else synth.random_pred(2, n_mb, 41, seed, 4 * R + 20)
        k2 = dict(kw, search_range=R, pred_policy=policy)
        k2.setdefault("qp", 28)
        (g, gp), kern = run(cuda, cur, refs, pred, True, **k2)
        (o, op), _ = run(oracle, cur, refs, pred, True, **k2)
        assert_same(gp, op, f"per-ref {kw} {kind} ({kern})")
        assert_same(g, o, f"best {kw} {kind} ({kern})")
        wide = kw.get("cost_domain") or kw.get("me_distortion_fpel") == 1
        if wide and kw.get("search_mode", 0) == abi.SEARCH_FASTFULL:
            assert kern.startswith("me_int_kernel<") and ("MODE=2" if kw.get("me_distortion_fpel") == 1 else "MODE=1") in kern, kern


def test_extreme_costs_in_the_wide_kernels(cuda, oracle):
    """0/255 opposition with the largest lambda: the largest SSE (16.6 M for 16x16, 2^29 in domain 1) and the
    all-ties frames, where only the 64-bit (cost, key) order decides."""
    cur = np.zeros((32, 32), np.uint8)
    ref = np.full((32, 32), 255, np.uint8)
    pred = synth.random_pred(1, 4, 41, seed=1, max_qpel=2048)
    for kw in (dict(cost_domain=1), dict(**SSE3), dict(cost_domain=1, **SSE3)):
        for rdopt in (0, 1):
            k2 = dict(kw, search_range=4, qp=51, rdopt=rdopt, pred_policy=abi.PRED_PER_BLOCK, subpel=1)
            assert_same(run(cuda, cur, [ref], pred, **k2)[0], run(oracle, cur, [ref], pred, **k2)[0], str(k2))
    for kind in ("const", "checker"):
        img = synth.gen_luma(48, 48, 0, kind)
        for kw in (dict(cost_domain=1), dict(**SSE3), dict(cost_domain=1, transform8x8=1)):
            for rdopt in (0, 1):
                for lf in (0, 1):
                    k2 = dict(kw, search_range=6, rdopt=rdopt, lambda_factor=lf, subpel=1)
                    assert_same(run(cuda, img, [img], **k2)[0], run(oracle, img, [img], **k2)[0], f"{kind} {k2}")


CHROMA_CASES = [
    dict(subpel=1, chroma_me=1, qp=27),
    dict(subpel=1, chroma_me=1, use_hadamard=0, rdopt=1, qp=33),
    dict(subpel=1, chroma_me=1, cost_domain=1, satd_round=1),
    dict(subpel=1, chroma_me=1, transform8x8=1, qp=30),
    dict(subpel=1, chroma_me=1, **SSE3),
    dict(subpel=1, chroma_me=1, me_distortion=1, me_distortion_fpel=0, me_distortion_hpel=1, me_distortion_qpel=2, cost_domain=1),
    dict(subpel=1, chroma_me=1, blocktype_mask=0x92),
    dict(subpel=1, chroma_me=1, search_mode=abi.SEARCH_FULL),
]


@pytest.mark.parametrize("kw", CHROMA_CASES)
def test_chroma_me_matches_oracle(cuda, oracle, kw):
    for (w, h, R, nref, policy) in ((64, 48, 6, 2, abi.PRED_ZERO), (52, 38, 9, 1, abi.PRED_PER_BLOCK), (96, 32, 32, 1, abi.PRED_PER_MB)):
        cur, refs = synth.frame_pair(w, h, seed=3, search_range=R, num_refs=nref)
        chroma = (chroma_pair(w, h, 50), [chroma_pair(w, h, 60 + 7 * i) for i in range(nref)])
        n_mb = ((w + 15) // 16) * ((h + 15) // 16)
        nb = {abi.PRED_ZERO: 0, abi.PRED_PER_MB: 1, abi.PRED_PER_BLOCK: 41}[policy]
        pred = synth.random_pred(nref, n_mb, nb, 5, 4 * R + 20) if nb else None
        k2 = dict(kw, search_range=R, pred_policy=policy)
        (g, gp), kern = run(cuda, cur, refs, pred, True, chroma=chroma, **k2)
        (o, op), _ = run(oracle, cur, refs, pred, True, chroma=chroma, **k2)
        assert_same(gp, op, f"per-ref {kw} {w}x{h}")
        assert_same(g, o, f"best {kw} {w}x{h}")
        # chroma really takes part: without it some decision differs
        (g0, _), _ = run(cuda, cur, refs, pred, True, **{k: v for k, v in k2.items() if k != "chroma_me"})
        assert g0.tobytes() != g.tobytes()


def test_in_frame_median_with_the_new_cost_functions(cuda, oracle):
    """The wavefront (JMME_PRED_MEDIAN) with cost domain 1 / SSE / 8x8 Hadamard / chroma: the wide integer kernels
    take their predictors from the wave-step kernel, the general sub-pel kernel commits."""
    w, h, R = 96, 64, 6
    cur, refs = synth.frame_pair(w, h, seed=4, search_range=R, num_refs=2)
    chroma = (chroma_pair(w, h, 70), [chroma_pair(w, h, 80 + i) for i in range(2)])
    for kw, ch in ((dict(cost_domain=1, subpel=1), None), (dict(subpel=1, **SSE3), None), (dict(subpel=1, transform8x8=1, slice_rows=2), None),
                   (dict(subpel=1, chroma_me=1, slice_rows=1), chroma), (dict(cost_domain=1, subpel=0), None)):
        k2 = dict(kw, search_range=R, pred_policy=abi.PRED_MEDIAN, qp=30)
        (g, gp), kern = run(cuda, cur, refs, None, True, chroma=ch, **k2)
        (o, op), _ = run(oracle, cur, refs, None, True, chroma=ch, **k2)
        assert_same(gp, op, f"per-ref {kw} ({kern})")
        assert_same(g, o, f"best {kw} ({kern})")


def test_errors_of_the_new_parameters_mirror_the_oracle(cuda, oracle):
    for lib in (cuda, oracle):
        for kw, code in ((dict(cost_domain=2), abi.ERR_PARAM), (dict(me_distortion=1, me_distortion_hpel=3), abi.ERR_PARAM),
                         (dict(me_distortion=1, me_distortion_fpel=2), abi.ERR_UNSUPPORTED), (dict(chroma_me=1), abi.ERR_PARAM),
                         (dict(transform8x8=2), abi.ERR_PARAM)):
            with pytest.raises(abi.JmmeError) as e:
                lib.context(width=32, height=32, **kw)
            assert e.value.code == code, kw
        with lib.context(width=32, height=32, search_range=2, subpel=1, chroma_me=1) as ctx:
            ctx.set_reference(0, np.zeros((32, 32), np.uint8))
            with pytest.raises(abi.JmmeError) as e:
                ctx.search_frame(np.zeros((32, 32), np.uint8))
            assert e.value.code == abi.ERR_STATE
        with lib.context(width=32, height=32, search_range=2) as ctx:
            with pytest.raises(abi.JmmeError) as e:
                ctx.set_current_chroma(np.zeros((16, 16), np.uint8), np.zeros((16, 16), np.uint8))
            assert e.value.code == abi.ERR_STATE


def test_chroma_device_api(cuda, oracle):
    import torch
    from jmme.torch_api import DeviceSearch
    w, h, R = 64, 48, 6
    cur, refs = synth.frame_pair(w, h, seed=3, search_range=R)
    cur_c, ref_c = chroma_pair(w, h, 50), chroma_pair(w, h, 60)
    kw = dict(search_range=R, subpel=1, chroma_me=1, qp=29)
    ds = DeviceSearch(cuda, width=w, height=h, **kw)
    ds.set_reference(0, torch.from_numpy(refs[0]).cuda())
    ds.set_reference_chroma(0, *[torch.from_numpy(x).cuda() for x in ref_c])
    ds.set_current_chroma(*[torch.from_numpy(x).cuda() for x in cur_c])
    got = ds.to_numpy(ds.search(torch.from_numpy(cur).cuda()))
    torch.cuda.synchronize()
    (o, _), _ = run(oracle, cur, refs, None, True, chroma=(cur_c, [ref_c]), **kw)
    assert_same(got, o, "device chroma API")
    ds.close()


@pytest.mark.parametrize("kw", [dict(subpel=1, qp=30), dict(subpel=0, rdopt=1, qp=26, cost_domain=1), dict(subpel=1, **SSE3),
                                dict(subpel=1, blocktype_mask=0x92, qp=34)])
@pytest.mark.parametrize("policy", [abi.PRED_ZERO, abi.PRED_PER_BLOCK])
def test_bipred_refinement_matches_oracle(cuda, oracle, kw, policy):
    """jmme_search_frame_bipred (f2): list 0 = two references, list 1 = one more picture; start pairs from the
    uni-directional searches of each library (already bit-exact), refinement records compared bit for bit."""
    for (w, h, R, rng, iters) in ((64, 48, 6, 3, 2), (52, 38, 9, 8, 3), (96, 64, 16, 15, 4)):
        cur, refs = synth.frame_pair(w, h, seed=9, search_range=R, num_refs=2)
        ref1 = synth.frame_pair(w, h, seed=10, search_range=R)[1][0]
        n_mb = ((w + 15) // 16) * ((h + 15) // 16)
        nb = 41 if policy == abi.PRED_PER_BLOCK else 0
        pred0 = synth.random_pred(2, n_mb, nb, 3, 4 * R + 10) if nb else None
        pred1 = synth.random_pred(1, n_mb, nb, 4, 4 * R + 10) if nb else None
        outs = []
        for lib in (cuda, oracle):
            with lib.context(width=w, height=h, search_range=R, num_refs=2, pred_policy=policy, **kw) as ctx, \
                    lib.context(width=w, height=h, search_range=R, num_refs=1, pred_policy=policy, **kw) as ctx1:
                for i, r in enumerate(refs):
                    ctx.set_reference(i, r)
                ctx1.set_reference(0, ref1)
                l0, l1 = ctx.search_frame(cur, pred0), ctx1.search_frame(cur, pred1)
                ctx.set_reference_l1(ref1)
                outs.append((l0, l1, ctx.search_frame_bipred(cur, l0, l1, pred0, pred1, rng, iters)))
        assert outs[0][0].tobytes() == outs[1][0].tobytes() and outs[0][1].tobytes() == outs[1][1].tobytes()
        g, o = outs[0][2], outs[1][2]
        for f in ("mv0", "mv1", "cost", "ref0"):
            bad = np.argwhere(g[f] != o[f])
            assert len(bad) == 0, f"{w}x{h} {kw} {f}: {len(bad)} mismatches, first {bad[0]}: gpu {g[f][tuple(bad[0][:2])]} oracle {o[f][tuple(bad[0][:2])]}"
        assert g.tobytes() == o.tobytes()
        assert np.any(g["mv0"] != outs[0][0]["mv"]) or np.any(g["mv1"] != outs[0][1]["mv"])       # something was refined


def test_bipred_stripes_virtual_devices_and_errors(cuda, oracle):
    import torch
    w, h, R = 64, 112, 6
    cur, refs = synth.frame_pair(w, h, seed=5, search_range=R)
    ref1 = synth.frame_pair(w, h, seed=6, search_range=R)[1][0]
    kw = dict(width=w, height=h, search_range=R, subpel=1, qp=29)
    with oracle.context(**kw) as o:
        o.set_reference(0, refs[0])
        l0 = o.search_frame(cur)
        o.set_reference(0, ref1)
        l1 = o.search_frame(cur)
        o.set_reference(0, refs[0])
        o.set_reference_l1(ref1)
        exp = o.search_frame_bipred(cur, l0, l1, search_range=4, iterations=2)
    ndev = torch.cuda.device_count()
    for extra in (dict(), dict(n_gpus=3, device_ids=[i % ndev for i in range(3)]), dict(mb_row_begin=2, mb_row_end=5)):
        with cuda.context(**kw, **extra) as g:
            with pytest.raises(abi.JmmeError) as e:
                g.set_reference(0, refs[0])
                g.search_frame_bipred(cur, l0, l1)
            assert e.value.code == abi.ERR_STATE
            g.set_reference_l1(ref1)
            got = g.search_frame_bipred(cur, l0, l1, search_range=4, iterations=2)
            rb, re = extra.get("mb_row_begin", 0), extra.get("mb_row_end", 7)
            assert got[rb * 4:re * 4].tobytes() == exp[rb * 4:re * 4].tobytes(), extra
            assert not got[:rb * 4].tobytes().strip(b"\0") and not got[re * 4:].tobytes().strip(b"\0")
    for lib in (cuda, oracle):
        with lib.context(width=w, height=h, search_range=R, subpel=0) as c:          # integer plane only: qpel vectors refused
            c.set_reference(0, refs[0])
            c.set_reference_l1(ref1)
            with pytest.raises(abi.JmmeError) as e:
                c.search_frame_bipred(cur, l0, l1, search_range=2, iterations=1)
            assert e.value.code == abi.ERR_PARAM
            bad = l0.copy()
            bad["mv"] = 0
            bad["ref_idx"][3, 7] = 2                                                     # not a reference of the context
            with pytest.raises(abi.JmmeError) as e:
                c.search_frame_bipred(cur, bad, bad, search_range=2, iterations=1)
            assert e.value.code == abi.ERR_PARAM


@pytest.mark.parametrize("policy,mode,R,extra", [
    (abi.PRED_PER_BLOCK, abi.SEARCH_FASTFULL, 8, dict(subpel=1)),
    (abi.PRED_PER_BLOCK, abi.SEARCH_FULL, 6, dict(subpel=1, num_refs=2)),
    (abi.PRED_PER_MB, abi.SEARCH_FASTFULL, 32, dict(subpel=1)),
    (abi.PRED_PER_BLOCK, abi.SEARCH_FASTFULL, 40, dict(subpel=0)),                       # R > 32: me_int.cu
    (abi.PRED_PER_BLOCK, abi.SEARCH_FASTFULL, 8, dict(subpel=1, cost_domain=1, chroma_me=1)),
    (abi.PRED_MEDIAN, abi.SEARCH_FASTFULL, 8, dict(subpel=1, slice_rows=0)),
])
def test_jm_center_rule_matches_oracle(cuda, oracle, policy, mode, R, extra):
    """jm_center = 1 with rdopt = 1: window centres that follow predictors of several R (not clamped to +-R), planes
    with the wider replication border — every kernel family against the oracle, and the default rule differs."""
    w, h = 112, 80
    nref = extra.get("num_refs", 1)
    cur, refs = synth.frame_pair(w, h, seed=12, search_range=R, num_refs=nref)
    n_mb = 7 * 5
    nb = {abi.PRED_PER_MB: 1, abi.PRED_PER_BLOCK: 41}.get(policy)
    pred = synth.random_pred(nref, n_mb, nb, seed=8, max_qpel=200) if nb else None
    kw = dict(width=w, height=h, search_range=R, pred_policy=policy, search_mode=mode, qp=29, rdopt=1, jm_center=1,
              max_pred_qpel=200, **extra)
    cc = [synth.gen_luma(w // 2, h // 2, 31 + k, "texture") for k in range(2)]
    outs = []
    for lib in (cuda, oracle):
        with lib.context(**kw) as ctx:
            assert ctx.pad == (200 // 4 + R + 16 + 15) & ~15
            for i, r in enumerate(refs):
                ctx.set_reference(i, r)
                if extra.get("chroma_me"):
                    ctx.set_reference_chroma(i, *cc)
            if extra.get("chroma_me"):
                ctx.set_current_chroma(*cc)
            outs.append(ctx.search_frame(cur, pred))
    assert outs[0].tobytes() == outs[1].tobytes()
    if pred is not None:
        with cuda.context(**dict(kw, jm_center=0)) as ctx:
            for i, r in enumerate(refs):
                ctx.set_reference(i, r)
                if extra.get("chroma_me"):
                    ctx.set_reference_chroma(i, *cc)
            if extra.get("chroma_me"):
                ctx.set_current_chroma(*cc)
            assert ctx.search_frame(cur, pred).tobytes() != outs[0].tobytes()


def test_randomised_round2_parameter_space(cuda, oracle):
    """Seeded random sweep over everything round 2 added, combined freely with the older parameters: cost domain,
    per-stage metrics, 8x8 Hadamard, chroma ME, JM's centre rule with a predictor limit, predictor policies incl. the
    in-frame median, stripes, and for the product library the launch shape (balanced task ranges, item groups, parts
    of the host path).  Even frame sizes (4:2:0 chroma)."""
    rng = np.random.default_rng(20262)
    kernels = set()
    for case in range(40):
        w, h = 2 * int(rng.integers(9, 72)), 2 * int(rng.integers(9, 56))
        R = int(rng.choice([2, 4, 7, 12, 16, 32, 32, 32, 36]))
        nref = int(rng.integers(1, 3))
        policy = int(rng.choice([0, 0, 1, 2, 3]))
        rdopt = int(rng.integers(0, 2))
        subpel = int(rng.integers(0, 2))
        kw = dict(search_range=R, qp=int(rng.integers(8, 48)), rdopt=rdopt, subpel=subpel, pred_policy=policy,
                  search_mode=int(rng.integers(0, 2)), satd_round=int(rng.integers(0, 2)), cost_domain=int(rng.integers(0, 2)),
                  blocktype_mask=int(rng.choice([0xFE, 0xFE, 0x92, 0x1E])))
        if rng.integers(0, 2):
            kw.update(me_distortion=1, me_distortion_fpel=int(rng.integers(0, 2)), me_distortion_hpel=int(rng.integers(0, 3)),
                      me_distortion_qpel=int(rng.integers(0, 3)), transform8x8=int(rng.integers(0, 2)))
        else:
            kw.update(use_hadamard=int(rng.integers(0, 2)))
        chroma = None
        if subpel and rng.integers(0, 2):
            kw["chroma_me"] = 1
            chroma = (chroma_pair(w, h, 100 + case), [chroma_pair(w, h, 200 + case + 7 * r) for r in range(nref)])
        maxp = 4 * R + 40
        if rdopt and rng.integers(0, 2):
            maxp = int(rng.choice([4 * R + 40, 240]))
            kw.update(jm_center=1, max_pred_qpel=maxp)
        mb_w, mb_h = (w + 15) // 16, (h + 15) // 16
        if policy == 3:
            kw["slice_rows"] = int(rng.choice([0, 1, 2]))
            if kw["slice_rows"] == 1 and mb_h >= 3 and rng.integers(0, 2):
                kw.update(mb_row_begin=1, mb_row_end=mb_h - 1)
        elif mb_h >= 3 and rng.integers(0, 3) == 0:
            kw.update(mb_row_begin=int(rng.integers(0, 2)), mb_row_end=mb_h - int(rng.integers(0, 2)))
        tuning = dict(balance=int(rng.integers(0, 3)), group=int(rng.choice([0, 1, 2, 4])), pipe_parts=int(rng.integers(1, 5)),
                      even_parts=int(rng.integers(0, 2)), early_subpel=int(rng.choice([0, 0, 2])), no_pair_tail=int(rng.integers(0, 2)))
        cur, refs = synth.frame_pair(w, h, seed=300 + case, search_range=R, kind=str(rng.choice(["texture", "noise"])), num_refs=nref)
        pred = None if policy in (0, 3) else synth.random_pred(nref, mb_w * mb_h, 1 if policy == 1 else 41, case, maxp)
        g, k = run(cuda, cur, refs, pred, chroma=chroma, tuning=tuning, **kw)
        o, _ = run(oracle, cur, refs, pred, chroma=chroma, **kw)
        rb, re = kw.get("mb_row_begin", 0), kw.get("mb_row_end", mb_h)
        sl = slice(rb * mb_w, re * mb_w)
        assert_same(g[sl], o[sl], f"case {case}: {w}x{h} refs {nref} {kw} {tuning} {k}")
        kernels.add(k.split("<")[0] + ("BAL=1" if "BAL=1" in k else ""))
    assert len(kernels) >= 3, kernels
