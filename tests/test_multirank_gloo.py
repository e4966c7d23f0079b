"""N > 1 host logic on CPU: world_size 2 and 3 over gloo.  Each rank searches its MB-row stripe (with
the oracle standing in for the GPU) and the all-gathered MV field must equal the single-rank field
byte for byte — the partition is invisible in the result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from jmme import abi, synth
from jmme.dist import StripeGather, stripe_of


def test_stripes_cover_rows_exactly():
    for mb_h in (1, 5, 18, 45, 68, 135):
        for world in (1, 2, 3, 4, 8):
            if world > mb_h:
                continue
            parts = [stripe_of(r, world, mb_h) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == mb_h
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in parts]
            assert max(sizes) == -(-mb_h // world)           # no stripe larger than an even split's largest
            assert all(b == min(r * max(sizes), mb_h) for r, (b, e) in enumerate(parts))


def test_slice_aligned_stripes():
    """unit > 1 (the slices of the in-frame median policy): every stripe starts and ends on a slice boundary."""
    for mb_h in (7, 18, 68, 135):
        for unit in (2, 3, 4):
            for world in (1, 2, 3, 8):
                parts = [stripe_of(r, world, mb_h, unit) for r in range(world)]
                assert parts[0][0] == 0 and max(e for _, e in parts) == mb_h
                assert all(b % unit == 0 and (e % unit == 0 or e == mb_h) for b, e in parts if e > b)
                assert all(parts[i][1] == parts[i + 1][0] or parts[i + 1][0] == parts[i + 1][1] == mb_h
                           for i in range(world - 1))


def _worker(rank, world, port, tmp, median=False):
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
    import oracle as oracle_mod
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = oracle_mod.load()
    w, h, R = 80, 112, 6                                     # 7 MB rows: uneven stripes
    cur, refs = synth.frame_pair(w, h, seed=3, search_range=R, num_refs=2)
    unit = 2 if median else 1                                # median policy: slices of 2 MB rows, whole slices per rank
    kw = dict(pred_policy=abi.PRED_MEDIAN, slice_rows=2) if median else {}
    rb, re = stripe_of(rank, world, 7, unit)
    g = StripeGather(5, 7, "cpu", unit=unit)
    if re > rb:                                              # a rank past the end of the frame has no rows
        with orc.context(width=w, height=h, search_range=R, num_refs=2, subpel=1, mb_row_begin=rb, mb_row_end=re, **kw) as c:
            for i, r in enumerate(refs):
                c.set_reference(i, r)
            mine = c.search_frame(cur)                       # whole-frame indexing, stripe rows valid
        rec = torch.from_numpy(mine.view(np.uint8).reshape(-1, abi.MBRESULT_DTYPE.itemsize))
        g.field[rb * 5:re * 5].copy_(rec[rb * 5:re * 5])
    full = g.gather()
    np.save(os.path.join(tmp, f"full_{world}_{rank}_{int(median)}.npy"), full.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3, 4])
def test_gathered_field_equals_single_rank_field(oracle, tmp_path, world):
    w, h, R = 80, 112, 6
    cur, refs = synth.frame_pair(w, h, seed=3, search_range=R, num_refs=2)
    with oracle.context(width=w, height=h, search_range=R, num_refs=2, subpel=1) as c:
        for i, r in enumerate(refs):
            c.set_reference(i, r)
        ref_field = c.search_frame(cur)
    port = 29600 + world + (os.getpid() % 200)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        got = np.load(tmp_path / f"full_{world}_{r}_0.npy")
        assert got.tobytes() == ref_field.tobytes(), f"rank {r} of {world}"


def test_gathered_in_frame_median_field_equals_single_rank_field(oracle, tmp_path):
    """The in-frame median policy shards by whole slices: two ranks reproduce the one-rank field."""
    w, h, R, world = 80, 112, 6, 2
    cur, refs = synth.frame_pair(w, h, seed=3, search_range=R, num_refs=2)
    with oracle.context(width=w, height=h, search_range=R, num_refs=2, subpel=1, pred_policy=abi.PRED_MEDIAN,
                        slice_rows=2) as c:
        for i, r in enumerate(refs):
            c.set_reference(i, r)
        ref_field = c.search_frame(cur)
    port = 29850 + (os.getpid() % 100)
    mp.spawn(_worker, args=(world, port, str(tmp_path), True), nprocs=world, join=True)
    for r in range(world):
        got = np.load(tmp_path / f"full_{world}_{r}_1.npy")
        assert got.tobytes() == ref_field.tobytes(), f"rank {r} of {world}"
