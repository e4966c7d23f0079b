"""Device-resident use of the C ABI from PyTorch: torch provides device memory, streams and
torch.distributed; all compute is libjmme_cuda.so (jmme_set_reference_dev / jmme_search_frame_dev)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import abi


class DeviceSearch:
    """One jmme_ctx bound to the current CUDA device, fed with torch uint8 tensors."""

    def __init__(self, lib: abi.Lib, **params):
        self.dev = torch.cuda.current_device()
        params.setdefault("device_ids", [self.dev])
        self.ctx = lib.context(**params)
        self.lib = lib
        n = self.ctx.mb_w * self.ctx.mb_h
        self.n_mb = n
        self.out = torch.zeros((n, abi.MBRESULT_DTYPE.itemsize), dtype=torch.uint8, device="cuda")
        self.out_per_ref = None

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def set_reference(self, ref_idx: int, luma: torch.Tensor):
        assert luma.is_cuda and luma.dtype == torch.uint8 and luma.dim() == 2 and luma.stride(1) == 1
        self.lib.check(self.lib.dll.jmme_set_reference_dev(self.ctx.handle, ref_idx, C.c_void_p(luma.data_ptr()),
                                                           luma.stride(0), self._stream()), self.ctx.handle)

    def set_reference_chroma(self, ref_idx: int, cb: torch.Tensor, cr: torch.Tensor):
        """chroma_me: the 4:2:0 chroma planes of reference `ref_idx` (uint8 cuda tensors of equal stride)."""
        assert cb.is_cuda and cr.is_cuda and cb.dtype == cr.dtype == torch.uint8 and cb.stride() == cr.stride() and cb.stride(1) == 1
        self.lib.check(self.lib.dll.jmme_set_reference_chroma_dev(self.ctx.handle, ref_idx, C.c_void_p(cb.data_ptr()),
                                                                  C.c_void_p(cr.data_ptr()), cb.stride(0), self._stream()),
                       self.ctx.handle)

    def set_current_chroma(self, cb: torch.Tensor, cr: torch.Tensor):
        assert cb.is_cuda and cr.is_cuda and cb.dtype == cr.dtype == torch.uint8 and cb.stride() == cr.stride() and cb.stride(1) == 1
        self.lib.check(self.lib.dll.jmme_set_current_chroma_dev(self.ctx.handle, C.c_void_p(cb.data_ptr()),
                                                                C.c_void_p(cr.data_ptr()), cb.stride(0), self._stream()),
                       self.ctx.handle)

    def search(self, cur: torch.Tensor, pred: torch.Tensor | None = None, per_ref: bool = False,
               out: torch.Tensor | None = None) -> torch.Tensor:
        """out: optional uint8 cuda tensor with room for mb_w*mb_h records (whole-frame indexing)."""
        assert cur.is_cuda and cur.dtype == torch.uint8 and cur.dim() == 2 and cur.stride(1) == 1
        if out is not None:
            assert out.is_cuda and out.dtype == torch.uint8 and out.is_contiguous()
            assert out.numel() >= self.n_mb * abi.MBRESULT_DTYPE.itemsize
            self.out = out
        if per_ref and self.out_per_ref is None:
            self.out_per_ref = torch.zeros((self.ctx.num_refs, self.n_mb, abi.MBRESULT_DTYPE.itemsize),
                                           dtype=torch.uint8, device="cuda")
        pp = C.c_void_p(pred.data_ptr()) if pred is not None else None
        po = C.c_void_p(self.out_per_ref.data_ptr()) if per_ref else None
        self.lib.check(self.lib.dll.jmme_search_frame_dev(self.ctx.handle, C.c_void_p(cur.data_ptr()), cur.stride(0),
                                                          pp, C.c_void_p(self.out.data_ptr()), po, self._stream()),
                       self.ctx.handle)
        return self.out

    def push_stripe(self, field: torch.Tensor, peer_ptrs):
        """Copy this rank's stripe of `field` into the same offsets of the peer buffers (NVLink stores)."""
        arr = (C.c_void_p * len(peer_ptrs))(*[C.c_void_p(int(p)) for p in peer_ptrs])
        self.lib.check(self.lib.dll.jmme_push_stripe_dev(self.ctx.handle, C.c_void_p(field.data_ptr()), arr, len(peer_ptrs),
                                                         self._stream()), self.ctx.handle)

    def set_peer_fields(self, peer_ptrs):
        """Every later search also stores its records into the peer buffers (fused gather); [] turns it off."""
        arr = (C.c_void_p * max(len(peer_ptrs), 1))(*[C.c_void_p(int(p)) for p in peer_ptrs])
        self.lib.check(self.lib.dll.jmme_set_peer_fields_dev(self.ctx.handle, arr, len(peer_ptrs)), self.ctx.handle)

    def set_multicast_field(self, mc_ptr):
        """Every later search stores its records once through the NVLS multicast mapping `mc_ptr` of the symmetric
        field (they land in every rank's field); 0 / None turns it off."""
        self.lib.check(self.lib.dll.jmme_set_multicast_field_dev(self.ctx.handle, C.c_void_p(int(mc_ptr or 0))), self.ctx.handle)

    def stripe_rows(self):
        return self.ctx.params.mb_row_begin, (self.ctx.params.mb_row_end or self.ctx.mb_h)

    @staticmethod
    def to_numpy(out: torch.Tensor) -> np.ndarray:
        return out.cpu().numpy().view(abi.MBRESULT_DTYPE).reshape(out.shape[:-1])

    def launch_count(self):
        return self.ctx.launch_count()

    def close(self):
        self.ctx.close()
