"""Known-answer tests of the oracle's tables and scalar cost arithmetic."""
import numpy as np
import pytest

import refimpl
from jmme import abi


def test_backend_and_abi(oracle):
    assert oracle.backend() == "cpu-oracle"
    assert oracle.dll.jmme_abi_version() == abi.ABI_VERSION


def test_mvbits_known_answers(oracle):
    mvbits, refbits, _, _ = oracle.init_motion_search_module(2, max_mvd=600)
    c = 600
    assert mvbits[c] == 1
    # |v| in [2^(k-1), 2^k - 1] -> 2k+1
    expect = {1: 3, 2: 5, 3: 5, 4: 7, 7: 7, 8: 9, 15: 9, 16: 11, 31: 11, 32: 13, 255: 17, 256: 19, 511: 19, 512: 21}
    for v, b in expect.items():
        assert mvbits[c + v] == b and mvbits[c - v] == b, v
    for v in range(-600, 601):
        assert mvbits[c + v] == refimpl.se_bits(v)
    assert list(refbits[:8]) == [1, 3, 3, 5, 5, 5, 5, 7]
    for r in range(16):
        assert refbits[r] == refimpl.ue_bits(r)


def test_spiral_hand_written(oracle):
    _, _, sx, sy = oracle.init_motion_search_module(1)
    assert list(zip(sx, sy)) == [(0, 0), (0, -1), (0, 1), (-1, -1), (1, -1), (-1, 0), (1, 0), (-1, 1), (1, 1)]
    _, _, sx, sy = oracle.init_motion_search_module(2)
    assert list(zip(sx, sy))[9:15] == [(-1, -2), (-1, 2), (0, -2), (0, 2), (1, -2), (1, 2)]
    assert list(zip(sx, sy))[15:19] == [(-2, -2), (2, -2), (-2, -1), (2, -1)]
    assert len(sx) == 25


@pytest.mark.parametrize("R", [1, 2, 5, 16, 32, 64])
def test_spiral_is_permutation_and_closed_form(oracle, R):
    _, _, sx, sy = oracle.init_motion_search_module(R)
    pts = list(zip(sx.tolist(), sy.tolist()))
    assert pts == refimpl.spiral(R)
    assert len(set(pts)) == (2 * R + 1) ** 2
    assert all(abs(x) <= R and abs(y) <= R for x, y in pts)
    for i, (x, y) in enumerate(pts):
        assert refimpl.spiral_index_closed_form(x, y) == i


def test_lambda_factor(oracle):
    # !rdopt: QP2QUANT[max(0,qp-12)] exactly representable
    q2q = [1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 4, 4, 4, 5, 6, 6, 7, 8, 9, 10, 11, 13, 14, 16, 18, 20, 23, 25, 29, 32,
           36, 40, 45, 51, 57, 64, 72, 81, 91]
    for qp in range(0, 52):
        assert oracle.lambda_factor(qp, 0) == q2q[max(0, qp - 12)] << 16
    # rdopt: sqrt(0.85 * 2^((qp-12)/3)); qp=12 -> sqrt(0.85)
    assert oracle.lambda_factor(12, 1) == int(65536 * np.sqrt(0.85) + 0.5)
    assert oracle.lambda_factor(28, 1) == int(65536 * np.sqrt(0.85 * 2 ** (16 / 3)) + 0.5)
    assert oracle.lambda_factor(28, 0) == 6 << 16


def test_create_rejects_bad_params(oracle):
    from jmme import abi
    for kw in (dict(width=0, height=16), dict(width=16, height=16, search_range=65),
               dict(width=16, height=16, num_refs=5), dict(width=16, height=16, blocktype_mask=0),
               dict(width=16, height=16, blocktype_mask=1), dict(width=16, height=16, qp=52),
               dict(width=16, height=16, mb_row_begin=1, mb_row_end=1)):
        with pytest.raises(abi.JmmeError) as e:
            oracle.context(**kw)
        assert e.value.code == abi.ERR_PARAM
    with pytest.raises(abi.JmmeError) as e:
        oracle.context(width=16, height=16, cost_domain=2)
    assert e.value.code == abi.ERR_PARAM
    with pytest.raises(abi.JmmeError) as e:                       # a Hadamard integer stage is the one refused metric
        oracle.context(width=16, height=16, me_distortion=1, me_distortion_fpel=abi.DIST_HADAMARD)
    assert e.value.code == abi.ERR_UNSUPPORTED


def test_search_before_reference_is_state_error(oracle):
    from jmme import abi
    with oracle.context(width=32, height=32, search_range=4) as ctx:
        with pytest.raises(abi.JmmeError) as e:
            ctx.search_frame(np.zeros((32, 32), np.uint8))
        assert e.value.code == abi.ERR_STATE


def test_spiral_inverse_closed_form():
    """The arithmetic inverse of the spiral order that the CUDA kernels use at their result write (jmme_dev.cuh
    d_spiral_xy: ring from an integer square root, top/bottom rows interleaved, then left/right columns), restated
    here and checked against the generated order for every position of a +-64 window."""
    import math

    def inverse(k):
        if k == 0:
            return 0, 0
        s = math.isqrt(k)
        l = (s + 1) >> 1
        w = 2 * l - 1
        off = k - w * w
        if off < 2 * w:
            return (off >> 1) - l + 1, (l if off & 1 else -l)
        o2 = off - 2 * w
        return (l if o2 & 1 else -l), (o2 >> 1) - l

    order = refimpl.spiral(64)
    assert len(order) == 129 * 129
    for k, xy in enumerate(order):
        assert inverse(k) == tuple(xy), (k, xy)
