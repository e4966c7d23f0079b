"""Multi-GPU plumbing of the search: MB-row stripes (one per rank) and the single collective of the
path, an all-gather of the MV field (torch.distributed: NCCL over NVLink on GPUs, gloo in CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import abi

REC = abi.MBRESULT_DTYPE.itemsize


def stripe_of(rank: int, world: int, mb_h: int, unit: int = 1):
    """Contiguous MB rows [begin, end) of `rank`.  Every stripe has ceil(mb_h/world) rows except the
    last non-empty one: the largest stripe is as small as an even split would make
    it, and rank r's rows start at r*ceil(mb_h/world), so that the stripes, padded to that size, tile
    the all-gather buffer in frame order and the gather needs no re-packing.  A rank past the end of
    the frame gets an empty stripe (begin == end).
    unit > 1: stripes are whole groups of `unit` rows (the slices of JMME_PRED_MEDIAN, whose stripes must
    start and end on slice boundaries)."""
    rows = -(-(-(-mb_h // unit)) // world) * unit
    b = min(rank * rows, mb_h)
    return b, min(b + rows, mb_h)


class StripeGather:
    """All-gather of the per-rank stripes of jmme_mbresult records into the whole-frame MV field.

    `field` is a uint8 tensor [world * rows_per_rank * mb_w, 372] whose first mb_w*mb_h records are the
    frame; rank r's stripe is chunk r.  A rank lets its search write straight into `field` (whole-frame
    indexing puts its stripe into its own chunk), then `gather()` runs ONE in-place all-gather — the
    only collective of the path."""

    def __init__(self, mb_w: int, mb_h: int, device, group=None, unit: int = 1):
        """unit: stripes are whole groups of `unit` MB rows (see stripe_of)."""
        self.mb_w, self.mb_h, self.group = mb_w, mb_h, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.rows = -(-(-(-mb_h // unit)) // self.world) * unit
        self.chunk = self.rows * mb_w
        self.field = torch.zeros((self.world * self.chunk, REC), dtype=torch.uint8, device=device)

    def frame(self) -> torch.Tensor:
        return self.field[: self.mb_w * self.mb_h]

    def gather(self) -> torch.Tensor:
        if self.world > 1:
            mine = self.field[self.rank * self.chunk:(self.rank + 1) * self.chunk]
            dist.all_gather_into_tensor(self.field, mine, group=self.group)
        return self.frame()


class PeerPushGather:
    """The same gather without a collective library: the MV field lives in symmetric memory
    (torch.distributed._symmetric_memory: every rank's buffer is mapped into every process), each rank
    pushes its stripe into all peers' fields with one kernel of NVLink peer stores
    (jmme_push_stripe_dev) and a symmetric-memory barrier orders the reads."""

    def __init__(self, mb_w: int, mb_h: int, device, group=None, unit: int = 1):
        import torch.distributed._symmetric_memory as symm_mem
        self.mb_w, self.mb_h = mb_w, mb_h
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.rows = -(-(-(-mb_h // unit)) // self.world) * unit
        n = self.world * self.rows * mb_w
        self.field = symm_mem.empty((n, REC), dtype=torch.uint8, device=device)
        self.field.zero_()
        self.hdl = symm_mem.rendezvous(self.field, group if group is not None else dist.group.WORLD)
        self.peer_ptrs = [int(p) for p in self.hdl.buffer_ptrs]

    def frame(self) -> torch.Tensor:
        return self.field[: self.mb_w * self.mb_h]

    def gather(self, search) -> torch.Tensor:
        """search: the DeviceSearch whose last result was written into self.field."""
        self.hdl.barrier(channel=0)            # peers are done reading the previous field
        search.push_stripe(self.field, self.peer_ptrs)
        self.hdl.barrier(channel=1)            # every stripe has landed everywhere
        return self.frame()

    # ---- fused variant: the search kernels store their records into the peers' fields themselves ----------
    def attach(self, search, multicast=True):
        """Every later search of `search` (with out=self.field) also writes its records into all peers' fields
        (jmme_set_peer_fields_dev): no push kernel; bracket the search with pre() and post()."""
        mc = int(getattr(self.hdl, "multicast_ptr", 0) or 0)
        if mc and multicast:
            search.set_multicast_field(mc)     # one multimem.st per word, replicated by the NVSwitch
            self.mode = "multicast"
        else:
            search.set_peer_fields(self.peer_ptrs)
            self.mode = "peer stores"

    def pre(self):
        self.hdl.barrier(channel=0)            # peers are done reading the previous field

    def post(self) -> torch.Tensor:
        self.hdl.barrier(channel=1)            # every stripe has landed everywhere
        return self.frame()
