"""ctypes view of include/jmme.h, shared by the product binding and by the test-side oracle loader.

Nothing here computes anything: it declares the C structs / prototypes of the C ABI and wraps the
calls with numpy buffers.  Which shared library is bound is the caller's choice
(`h264-jm-commentary_b200/jmme/__init__.py` binds libjmme_cuda.so; `oracle/oracle.py` binds the oracle).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

ABI_VERSION = 6          # JMME_ABI_VERSION of include/jmme.h
BLOCKS_PER_MB = 41
MAX_REFS = 4
MAX_GPUS = 8
INT32_MAX = 2**31 - 1

OK, ERR_PARAM, ERR_CUDA, ERR_NOMEM, ERR_UNSUPPORTED, ERR_STATE, ERR_NODEVICE = 0, -1, -2, -3, -4, -5, -6
SEARCH_FASTFULL, SEARCH_FULL = 0, 1
PRED_ZERO, PRED_PER_MB, PRED_PER_BLOCK, PRED_MEDIAN = 0, 1, 2, 3
DIST_SAD, DIST_SSE, DIST_HADAMARD = 0, 1, 2
MASK_16x16, MASK_ALL = 0x02, 0xFE

# block geometry (JM blc_size): blocktype -> (w, h); result index bases
BLC = {1: (16, 16), 2: (16, 8), 3: (8, 16), 4: (8, 8), 5: (8, 4), 6: (4, 8), 7: (4, 4)}
BLK_BASE = {1: 0, 2: 1, 3: 3, 4: 5, 5: 9, 6: 17, 7: 25}


def block_table():
    """[(blocktype, x0, y0, w, h)] for the 41 blocks in result order."""
    out = []
    for t in range(1, 8):
        w, h = BLC[t]
        for j in range(16 // h):
            for i in range(16 // w):
                out.append((t, i * w, j * h, w, h))
    return out


class Params(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "width", "height", "search_range", "num_refs", "blocktype_mask", "lambda_factor", "qp", "rdopt",
        "use_hadamard", "subpel", "search_mode", "pred_policy", "satd_round", "cost_domain",
        "mb_row_begin", "mb_row_end", "n_gpus")] + [("device_ids", C.c_int32 * MAX_GPUS),
                                                     ("async_reference", C.c_int32), ("slice_rows", C.c_int32)] + \
        [(n, C.c_int32) for n in ("me_distortion", "me_distortion_fpel", "me_distortion_hpel", "me_distortion_qpel",
                                  "transform8x8", "chroma_me", "jm_center", "max_pred_qpel")]


class Tuning(C.Structure):
    """jmme_tuning: launch knobs of the product library, 0 = default."""
    _fields_ = [(n, C.c_int32) for n in ("variant", "group", "cluster", "table_rate", "wave_step", "no_pdl",
                                         "pipe_parts", "balance", "even_parts", "early_subpel", "no_pair_tail")] + \
        [("reserved", C.c_int32 * 5)]


MBRESULT_DTYPE = np.dtype([("mv", np.int16, (BLOCKS_PER_MB, 2)), ("cost", np.int32, (BLOCKS_PER_MB,)),
                           ("ref_idx", np.int8, (BLOCKS_PER_MB,)), ("reserved", np.int8, (3,))], align=True)
assert MBRESULT_DTYPE.itemsize == 372, MBRESULT_DTYPE.itemsize
BIPRED_DTYPE = np.dtype([("mv0", np.int16, (BLOCKS_PER_MB, 2)), ("mv1", np.int16, (BLOCKS_PER_MB, 2)),
                         ("cost", np.int32, (BLOCKS_PER_MB,)), ("ref0", np.int8, (BLOCKS_PER_MB,)), ("reserved", np.int8, (3,))],
                        align=True)
assert BIPRED_DTYPE.itemsize == 536, BIPRED_DTYPE.itemsize

EXPORTS = [
    "jmme_default_params", "jmme_create", "jmme_destroy", "jmme_strerror", "jmme_last_error", "jmme_backend",
    "jmme_abi_version", "jmme_mb_width", "jmme_mb_height", "jmme_pad", "jmme_lambda_factor_of",
    "jmme_lambda_factor", "jmme_set_reference_l1", "jmme_search_frame_bipred", "jmme_set_reference", "jmme_set_reference_chroma", "jmme_set_current_chroma", "jmme_search_frame", "jmme_get_predictors", "jmme_get_subimage",
    "jmme_set_reference_dev", "jmme_search_frame_dev", "jmme_set_reference_chroma_dev", "jmme_set_current_chroma_dev", "jmme_push_stripe_dev", "jmme_set_peer_fields_dev", "jmme_set_multicast_field_dev",
    "jmme_launch_count", "jmme_set_tuning", "jmme_get_tuning", "jmme_last_kernel",
    "jmme_set_profiling",
    "jmme_get_kernel_times", "jmme_InitMotionSearchModule", "jmme_SetMotionVectorPredictor",
    "jmme_commit_field", "jmme_predict_frame",
    "jmme_getSubImagesLuma", "jmme_getSubImagesChroma", "jmme_SATD", "jmme_HadamardSAD8x8", "jmme_SetupFastFullPelSearch", "jmme_FastFullPelBlockMotionSearch",
    "jmme_FullPelBlockMotionSearch", "jmme_SubPelBlockMotionSearch",
]


class JmmeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"jmme error {code}: {msg}")
        self.code = code


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(C.POINTER(C.c_uint8))


class Lib:
    """One loaded implementation of include/jmme.h."""

    def __init__(self, path):
        self.path = str(path)
        self.dll = C.CDLL(self.path)          # raises OSError if the library is missing
        d = self.dll
        vp, i32, pu8 = C.c_void_p, C.c_int, C.POINTER(C.c_uint8)
        pi16, pi32 = C.POINTER(C.c_int16), C.POINTER(C.c_int32)
        protos = {
            "jmme_default_params": (None, [C.POINTER(Params)]),
            "jmme_create": (i32, [C.POINTER(vp), C.POINTER(Params)]),
            "jmme_destroy": (i32, [vp]),
            "jmme_strerror": (C.c_char_p, [i32]),
            "jmme_last_error": (C.c_char_p, [vp]),
            "jmme_backend": (C.c_char_p, []),
            "jmme_abi_version": (i32, []),
            "jmme_mb_width": (i32, [vp]), "jmme_mb_height": (i32, [vp]), "jmme_pad": (i32, [vp]),
            "jmme_lambda_factor_of": (i32, [vp]),
            "jmme_lambda_factor": (i32, [i32, i32]),
            "jmme_set_reference": (i32, [vp, i32, pu8, i32]),
            "jmme_set_reference_l1": (i32, [vp, pu8, i32]),
            "jmme_search_frame_bipred": (i32, [vp, pu8, i32, vp, vp, pi16, pi16, i32, i32, vp]),
            "jmme_set_reference_chroma": (i32, [vp, i32, pu8, pu8, i32]),
            "jmme_set_current_chroma": (i32, [vp, pu8, pu8, i32]),
            "jmme_search_frame": (i32, [vp, pu8, i32, pi16, vp, vp]),
            "jmme_get_predictors": (i32, [vp, pi16]),
            "jmme_get_subimage": (i32, [vp, i32, i32, i32, pu8, i32]),
            "jmme_set_reference_dev": (i32, [vp, i32, vp, i32, vp]),
            "jmme_search_frame_dev": (i32, [vp, vp, i32, vp, vp, vp, vp]),
            "jmme_set_reference_chroma_dev": (i32, [vp, i32, vp, vp, i32, vp]),
            "jmme_set_current_chroma_dev": (i32, [vp, vp, vp, i32, vp]),
            "jmme_push_stripe_dev": (i32, [vp, vp, C.POINTER(vp), i32, vp]),
            "jmme_set_peer_fields_dev": (i32, [vp, C.POINTER(vp), i32]),
            "jmme_set_multicast_field_dev": (i32, [vp, vp]),
            "jmme_launch_count": (C.c_longlong, [vp]),
            "jmme_set_tuning": (i32, [vp, C.POINTER(Tuning)]),
            "jmme_get_tuning": (i32, [vp, C.POINTER(Tuning)]),
            "jmme_last_kernel": (C.c_char_p, [vp]),
            "jmme_set_profiling": (i32, [vp, i32]),
            "jmme_get_kernel_times": (i32, [vp, C.POINTER(C.c_float)]),
            "jmme_InitMotionSearchModule": (i32, [i32, i32, pi32, i32, pi32, pi16, pi16]),
            "jmme_SetMotionVectorPredictor": (i32, [i32, i32, i32, pi16, i32, i32, pi16, i32, i32, pi16, i32, i32, pi16]),
            "jmme_commit_field": (i32, [vp, vp, pi16, C.POINTER(C.c_int8), pu8]),
            "jmme_predict_frame": (i32, [vp, pi16, C.POINTER(C.c_int8), pi16]),
            "jmme_getSubImagesLuma": (i32, [pu8, i32, i32, i32, i32, pu8]),
            "jmme_SATD": (i32, [pi16, i32, i32, pi32]),
            "jmme_HadamardSAD8x8": (i32, [pi16, i32, i32, pi32]),
            "jmme_getSubImagesChroma": (i32, [pu8, i32, i32, i32, i32, pu8]),
            "jmme_SetupFastFullPelSearch": (i32, [pu8, i32, pu8, i32, i32, i32, i32, i32, i32, i32, pi32]),
            "jmme_FastFullPelBlockMotionSearch": (i32, [pi32, i32, i32, i32, i32, i32, i32, i32, pi16, pi16, pi32]),
            "jmme_FullPelBlockMotionSearch": (i32, [pu8, i32, pu8, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32,
                                                    pi16, pi16, pi32]),
            "jmme_SubPelBlockMotionSearch": (i32, [pu8, i32, pu8, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32,
                                                   i32, i32, i32, pi16, pi16, pi32]),
        }
        assert set(protos) == set(EXPORTS)
        for name, (res, args) in protos.items():
            fn = getattr(d, name)             # AttributeError if a declared symbol is not exported
            fn.restype, fn.argtypes = res, args

    # ---- small helpers -------------------------------------------------------------------
    def backend(self):
        return self.dll.jmme_backend().decode()

    def check(self, rc, ctx=None):
        if rc != OK:
            msg = self.dll.jmme_strerror(rc).decode()
            if ctx:
                det = self.dll.jmme_last_error(ctx)
                if det:
                    msg += ": " + det.decode()
            raise JmmeError(rc, msg)

    def default_params(self, **kw):
        p = Params()
        self.dll.jmme_default_params(C.byref(p))
        for k, v in kw.items():
            if k == "device_ids":
                for i, d in enumerate(v):
                    p.device_ids[i] = d
            else:
                if not hasattr(p, k):
                    raise AttributeError(k)
                setattr(p, k, v)
        return p

    def lambda_factor(self, qp, rdopt):
        return self.dll.jmme_lambda_factor(qp, rdopt)

    def context(self, tuning=None, **kw):
        """tuning: dict of jmme_tuning fields applied right after jmme_create (product library only)."""
        ctx = Context(self, self.default_params(**kw))
        if tuning:
            ctx.set_tuning(**tuning)
        return ctx

    # ---- leaf entry points ---------------------------------------------------------------
    def init_motion_search_module(self, R, max_mvd=64, n_refbits=16):
        mvbits = np.zeros(2 * max_mvd + 1, np.int32)
        refbits = np.zeros(n_refbits, np.int32)
        n = (2 * R + 1) ** 2
        sx, sy = np.zeros(n, np.int16), np.zeros(n, np.int16)
        self.check(self.dll.jmme_InitMotionSearchModule(
            R, max_mvd, mvbits.ctypes.data_as(C.POINTER(C.c_int32)), n_refbits,
            refbits.ctypes.data_as(C.POINTER(C.c_int32)), sx.ctypes.data_as(C.POINTER(C.c_int16)),
            sy.ctypes.data_as(C.POINTER(C.c_int16))))
        return mvbits, refbits, sx, sy

    def set_motion_vector_predictor(self, blocktype, part, ref_idx, A, B, Cn):
        """A, B, Cn: (mvx, mvy, ref, avail) of the left, up and up-right(-or-up-left) neighbours."""
        arr = [np.array(n[:2], np.int16) for n in (A, B, Cn)]
        out = np.zeros(2, np.int16)
        p16 = lambda a: a.ctypes.data_as(C.POINTER(C.c_int16))                       # noqa: E731
        self.check(self.dll.jmme_SetMotionVectorPredictor(blocktype, part, ref_idx, p16(arr[0]), A[2], A[3], p16(arr[1]),
                                                          B[2], B[3], p16(arr[2]), Cn[2], Cn[3], p16(out)))
        return int(out[0]), int(out[1])

    def get_sub_images_luma(self, luma, pad):
        luma, p = _u8(luma)
        h, w = luma.shape
        out = np.zeros((4, 4, h + 2 * pad, w + 2 * pad), np.uint8)     # [yfrac][xfrac]
        self.check(self.dll.jmme_getSubImagesLuma(p, w, h, luma.strides[0], pad,
                                                  out.ctypes.data_as(C.POINTER(C.c_uint8))))
        return out

    def satd(self, diffs, satd_round=0):
        d = np.ascontiguousarray(diffs, dtype=np.int16).reshape(-1, 16)
        out = np.zeros(len(d), np.int32)
        self.check(self.dll.jmme_SATD(d.ctypes.data_as(C.POINTER(C.c_int16)), len(d), satd_round,
                                      out.ctypes.data_as(C.POINTER(C.c_int32))))
        return out

    def hadamard_sad8x8(self, diffs, satd_round=1):
        d = np.ascontiguousarray(diffs, dtype=np.int16).reshape(-1, 64)
        out = np.zeros(len(d), np.int32)
        self.check(self.dll.jmme_HadamardSAD8x8(d.ctypes.data_as(C.POINTER(C.c_int16)), len(d), satd_round,
                                                out.ctypes.data_as(C.POINTER(C.c_int32))))
        return out

    def get_sub_images_chroma(self, chroma, pad):
        """The 64 eighth-pel planes [yF][xF] of one chroma component, padded by `pad`."""
        chroma, p = _u8(chroma)
        h, w = chroma.shape
        out = np.zeros((8, 8, h + 2 * pad, w + 2 * pad), np.uint8)
        self.check(self.dll.jmme_getSubImagesChroma(p, w, h, chroma.strides[0], pad,
                                                    out.ctypes.data_as(C.POINTER(C.c_uint8))))
        return out

    def setup_fast_full_pel_search(self, cur_mb, ref_padded, pad, mb_x, mb_y, cx, cy, R, bonus=0):
        cur_mb, pc = _u8(cur_mb)
        ref_padded, _ = _u8(ref_padded)
        rs = ref_padded.strides[0]
        p00 = C.cast(ref_padded.ctypes.data + pad * rs + pad, C.POINTER(C.c_uint8))
        out = np.zeros((BLOCKS_PER_MB, (2 * R + 1) ** 2), np.int32)
        self.check(self.dll.jmme_SetupFastFullPelSearch(pc, cur_mb.strides[0], p00, rs, mb_x, mb_y, cx, cy, R,
                                                        bonus, out.ctypes.data_as(C.POINTER(C.c_int32))))
        return out

    def fast_full_pel_block_motion_search(self, blocksad, R, cx, cy, px, py, lambda_factor, pretest00):
        s = np.ascontiguousarray(blocksad, dtype=np.int32)
        mx, my, mc = C.c_int16(), C.c_int16(), C.c_int32()
        self.check(self.dll.jmme_FastFullPelBlockMotionSearch(
            s.ctypes.data_as(C.POINTER(C.c_int32)), R, cx, cy, px, py, lambda_factor, pretest00,
            C.byref(mx), C.byref(my), C.byref(mc)))
        return mx.value, my.value, mc.value

    def full_pel_block_motion_search(self, cur, ref_padded, pad, bx, by, bw, bh, px, py, R, lambda_factor,
                                     bonus=0):
        cur, pc = _u8(cur)
        ref_padded, _ = _u8(ref_padded)
        rs = ref_padded.strides[0]
        p00 = C.cast(ref_padded.ctypes.data + pad * rs + pad, C.POINTER(C.c_uint8))
        mx, my, mc = C.c_int16(), C.c_int16(), C.c_int32()
        self.check(self.dll.jmme_FullPelBlockMotionSearch(pc, cur.strides[0], p00, rs, bx, by, bw, bh, px, py, R,
                                                          lambda_factor, bonus, C.byref(mx), C.byref(my),
                                                          C.byref(mc)))
        return mx.value, my.value, mc.value

    def sub_pel_block_motion_search(self, cur, planes, pad, bx, by, bw, bh, px, py, lambda_factor, mv, cost,
                                    use_hadamard=1, satd_round=0, bonus=0):
        cur, pc = _u8(cur)
        planes, pp = _u8(planes)
        h, w = planes.shape[-2] - 2 * pad, planes.shape[-1] - 2 * pad
        mx, my, mc = C.c_int16(mv[0]), C.c_int16(mv[1]), C.c_int32(cost)
        self.check(self.dll.jmme_SubPelBlockMotionSearch(pc, cur.strides[0], pp, w, h, pad, bx, by, bw, bh, px, py,
                                                         lambda_factor, use_hadamard, satd_round, bonus,
                                                         C.byref(mx), C.byref(my), C.byref(mc)))
        return mx.value, my.value, mc.value


@dataclass
class Context:
    """jmme_ctx with numpy in/out.  Mirrors the C calls one to one."""
    lib: Lib
    params: Params
    handle: C.c_void_p = field(default=None, repr=False)

    def __post_init__(self):
        h = C.c_void_p()
        self.lib.check(self.lib.dll.jmme_create(C.byref(h), C.byref(self.params)))
        self.handle = h
        d = self.lib.dll
        self.mb_w, self.mb_h, self.pad = d.jmme_mb_width(h), d.jmme_mb_height(h), d.jmme_pad(h)
        self.lambda_factor = d.jmme_lambda_factor_of(h)
        self.num_refs = self.params.num_refs

    def close(self):
        if self.handle:
            self.lib.dll.jmme_destroy(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_reference(self, ref_idx, luma):
        luma, p = _u8(luma)
        self.lib.check(self.lib.dll.jmme_set_reference(self.handle, ref_idx, p, luma.strides[0]), self.handle)

    def set_reference_chroma(self, ref_idx, cb, cr):
        cb, pb = _u8(cb)
        cr, pr = _u8(cr)
        assert cb.shape == cr.shape and cb.strides == cr.strides
        self.lib.check(self.lib.dll.jmme_set_reference_chroma(self.handle, ref_idx, pb, pr, cb.strides[0]), self.handle)

    def set_current_chroma(self, cb, cr):
        cb, pb = _u8(cb)
        cr, pr = _u8(cr)
        assert cb.shape == cr.shape and cb.strides == cr.strides
        self.lib.check(self.lib.dll.jmme_set_current_chroma(self.handle, pb, pr, cb.strides[0]), self.handle)

    def search_frame(self, cur, pred=None, per_ref=False):
        cur, p = _u8(cur)
        n = self.mb_w * self.mb_h
        out = np.zeros(n, MBRESULT_DTYPE)
        opr = np.zeros((self.num_refs, n), MBRESULT_DTYPE) if per_ref else None
        pp = None
        if pred is not None:
            pred = np.ascontiguousarray(pred, dtype=np.int16)
            nb = {PRED_PER_MB: 1, PRED_PER_BLOCK: BLOCKS_PER_MB}.get(self.params.pred_policy)
            if nb is not None and pred.size != self.num_refs * n * nb * 2:
                raise ValueError(f"pred has {pred.size} elements, expected {self.num_refs * n * nb * 2}")
            pp = pred.ctypes.data_as(C.POINTER(C.c_int16))
        self.lib.check(self.lib.dll.jmme_search_frame(self.handle, p, cur.strides[0], pp, out.ctypes.data,
                                                      opr.ctypes.data if per_ref else None), self.handle)
        return (out, opr) if per_ref else out

    def set_reference_l1(self, luma):
        luma, p = _u8(luma)
        self.lib.check(self.lib.dll.jmme_set_reference_l1(self.handle, p, luma.strides[0]), self.handle)

    def search_frame_bipred(self, cur, l0, l1, pred0=None, pred1=None, search_range=8, iterations=2):
        """Bi-predictive refinement of the vector pairs of two uni-directional searches -> BIPRED_DTYPE records."""
        cur, p = _u8(cur)
        l0, l1 = np.ascontiguousarray(l0), np.ascontiguousarray(l1)
        n = self.mb_w * self.mb_h
        assert l0.dtype == MBRESULT_DTYPE and l1.dtype == MBRESULT_DTYPE and len(l0) == len(l1) == n
        out = np.zeros(n, BIPRED_DTYPE)
        pp = []
        for pr in (pred0, pred1):
            if pr is None:
                pp.append(None)
            else:
                pr = np.ascontiguousarray(pr, dtype=np.int16)
                pp.append((pr, pr.ctypes.data_as(C.POINTER(C.c_int16))))
        self.lib.check(self.lib.dll.jmme_search_frame_bipred(self.handle, p, cur.strides[0], l0.ctypes.data, l1.ctypes.data,
                                                             pp[0][1] if pp[0] else None, pp[1][1] if pp[1] else None,
                                                             search_range, iterations, out.ctypes.data), self.handle)
        return out

    def commit_field(self, res):
        """ME-only mode decision -> (mv4 [4mb_h,4mb_w,2] int16, ref4 [4mb_h,4mb_w] int8, mode [n_mb,5] uint8)."""
        res = np.ascontiguousarray(res)
        mv4 = np.zeros((4 * self.mb_h, 4 * self.mb_w, 2), np.int16)
        ref4 = np.zeros((4 * self.mb_h, 4 * self.mb_w), np.int8)
        mode = np.zeros((self.mb_w * self.mb_h, 5), np.uint8)
        self.lib.check(self.lib.dll.jmme_commit_field(self.handle, res.ctypes.data, mv4.ctypes.data_as(C.POINTER(C.c_int16)),
                                                      ref4.ctypes.data_as(C.POINTER(C.c_int8)),
                                                      mode.ctypes.data_as(C.POINTER(C.c_uint8))), self.handle)
        return mv4, ref4, mode

    def predict_frame(self, mv4, ref4):
        """PER_BLOCK predictors int16 [num_refs, n_mb, 41, 2] from a committed field."""
        mv4 = np.ascontiguousarray(mv4, np.int16)
        ref4 = np.ascontiguousarray(ref4, np.int8)
        pred = np.zeros((self.num_refs, self.mb_w * self.mb_h, BLOCKS_PER_MB, 2), np.int16)
        self.lib.check(self.lib.dll.jmme_predict_frame(self.handle, mv4.ctypes.data_as(C.POINTER(C.c_int16)),
                                                       ref4.ctypes.data_as(C.POINTER(C.c_int8)),
                                                       pred.ctypes.data_as(C.POINTER(C.c_int16))), self.handle)
        return pred

    def get_predictors(self):
        """Predictors used by the last PRED_MEDIAN search: int16 [num_refs, n_mb, 41, 2]."""
        pred = np.zeros((self.num_refs, self.mb_w * self.mb_h, BLOCKS_PER_MB, 2), np.int16)
        self.lib.check(self.lib.dll.jmme_get_predictors(self.handle, pred.ctypes.data_as(C.POINTER(C.c_int16))),
                       self.handle)
        return pred

    def get_subimage(self, ref_idx, xfrac, yfrac):
        ph, pw = self.mb_h * 16 + 2 * self.pad, self.mb_w * 16 + 2 * self.pad
        dst = np.zeros((ph, pw), np.uint8)
        self.lib.check(self.lib.dll.jmme_get_subimage(self.handle, ref_idx, xfrac, yfrac,
                                                      dst.ctypes.data_as(C.POINTER(C.c_uint8)), pw), self.handle)
        return dst

    def launch_count(self):
        return int(self.lib.dll.jmme_launch_count(self.handle))

    def set_tuning(self, **kw):
        t = Tuning()
        for k, v in kw.items():
            if not hasattr(t, k):
                raise AttributeError(k)
            setattr(t, k, int(v))
        self.lib.check(self.lib.dll.jmme_set_tuning(self.handle, C.byref(t)), self.handle)

    def get_tuning(self):
        t = Tuning()
        self.lib.check(self.lib.dll.jmme_get_tuning(self.handle, C.byref(t)), self.handle)
        return {n: getattr(t, n) for n, _ in Tuning._fields_ if n != "reserved"}

    def last_kernel(self):
        """Integer-search kernel instantiation the last search launched (jmme_last_kernel)."""
        return self.lib.dll.jmme_last_kernel(self.handle).decode()

    def set_profiling(self, enable=True):
        self.lib.check(self.lib.dll.jmme_set_profiling(self.handle, int(enable)), self.handle)

    def kernel_times(self):
        """ms of the last (interp, me_int, me_subpel, select) kernels; needs set_profiling(True)."""
        ms = (C.c_float * 4)()
        self.lib.check(self.lib.dll.jmme_get_kernel_times(self.handle, ms), self.handle)
        return dict(interp=ms[0], me_int=ms[1], me_subpel=ms[2], select=ms[3])
