"""Frozen vectors (tests/golden/*.npz, made by tools/make_golden.py): the oracle must keep reproducing
them (CPU test) and the CUDA library must reproduce them through the C ABI (GPU test)."""
import pathlib

import numpy as np
import pytest

GOLD = sorted((pathlib.Path(__file__).parent / "golden").glob("*.npz"))


def replay(lib, g):
    w, h, R, refs, pol = (int(v) for v in g["params"])
    kw = {str(k): int(v) for k, v in zip(g["kw_keys"], g["kw_vals"])}
    pred = g["pred"] if pol in (1, 2) else None            # 0: zero predictors, 3: in-frame median
    with lib.context(width=w, height=h, search_range=R, num_refs=refs, pred_policy=pol, **kw) as ctx:
        for i in range(refs):
            ctx.set_reference(i, g["refs"][i])
        if kw.get("chroma_me"):
            for i in range(refs):
                ctx.set_reference_chroma(i, g["ref_c"][i, 0], g["ref_c"][i, 1])
            ctx.set_current_chroma(g["cur_c"][0], g["cur_c"][1])
        res, per = ctx.search_frame(g["cur"], pred, per_ref=True)
        planes = None
        if kw.get("subpel"):
            c = ctx.pad
            planes = np.stack([ctx.get_subimage(0, fx, fy)[c - 8:c + 16, c - 8:c + 16] for fy in range(4) for fx in range(4)])
    return res, per, planes


def check(lib, path):
    g = np.load(path)
    res, per, planes = replay(lib, g)
    assert np.array_equal(res["mv"], g["mv"]) and np.array_equal(res["cost"], g["cost"])
    assert np.array_equal(res["ref_idx"], g["ref_idx"])
    assert np.array_equal(per["mv"], g["per_mv"]) and np.array_equal(per["cost"], g["per_cost"])
    if planes is not None:
        assert np.array_equal(planes, g["planes_crop"])


def test_golden_files_exist():
    assert len(GOLD) >= 12


@pytest.mark.parametrize("path", GOLD, ids=lambda p: p.stem)
def test_oracle_reproduces_golden(oracle, path):
    check(oracle, path)


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=lambda p: p.stem)
def test_cuda_reproduces_golden(cuda, path):
    check(cuda, path)
