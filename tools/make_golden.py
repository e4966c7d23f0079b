"""Regenerates tests/golden/*.npz: small frozen input/output vectors of the CPU oracle.

The mounted reference has no sources, tests or vectors, so these are NOT reference
outputs: they pin the oracle's behaviour (the frozen spec of DESIGN.md §2) against accidental change,
and give the GPU tests a fixture that does not depend on the oracle build of the day.
Run from the repo root:  python tools/make_golden.py
"""
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "h264-jm-commentary_b200"))
sys.path.insert(0, str(ROOT / "oracle"))
import oracle as oracle_mod  # noqa: E402
from jmme import abi, synth  # noqa: E402

CASES = {
    # name: (w, h, R, refs, kwargs, pred policy)
    "cif_quarter_16x16_full_r16": (176, 144, 16, 1, dict(blocktype_mask=abi.MASK_16x16, search_mode=abi.SEARCH_FULL, qp=28), 0),
    "all41_fastfull_r8_qpel": (96, 64, 8, 1, dict(qp=28, subpel=1), 0),
    "all41_rdopt_r12_2refs_qpel_perblock": (64, 64, 12, 2, dict(qp=33, rdopt=1, subpel=1, satd_round=1), 2),
    "all41_r32_int_permb": (64, 48, 32, 1, dict(qp=24), 1),
    "odd_size_sad_subpel": (52, 38, 6, 1, dict(qp=30, subpel=1, use_hadamard=0), 0),
    # in-frame median (JMME_PRED_MEDIAN): slices of two MB rows with two references; one slice = the frame
    "median_slices2_2refs_qpel": (80, 96, 8, 2, dict(qp=30, subpel=1, slice_rows=2), 3),
    "median_wholeframe_rdopt_r6": (96, 64, 6, 1, dict(qp=34, rdopt=1, subpel=1, slice_rows=0), 3),
    # round 2: JM >= 12 scaled-up cost domain, SSE / mixed per-stage metrics, 8x8 Hadamard, chroma ME
    "costdomain1_r8_2refs_qpel_perblock": (64, 48, 8, 2, dict(qp=29, subpel=1, cost_domain=1), 2),
    "sse_all_stages_r6": (64, 48, 6, 1, dict(qp=26, subpel=1, me_distortion=1, me_distortion_fpel=1, me_distortion_hpel=1,
                                             me_distortion_qpel=1), 0),
    "mixed_metrics_hadamard8_r6_domain1": (64, 64, 6, 1, dict(qp=31, rdopt=1, subpel=1, me_distortion=1, me_distortion_fpel=1,
                                                              me_distortion_hpel=0, me_distortion_qpel=2, transform8x8=1,
                                                              satd_round=1, cost_domain=1), 1),
    "chroma_me_r6_2refs": (64, 48, 6, 2, dict(qp=28, subpel=1, chroma_me=1), 0),
    "chroma_me_hadamard8_median": (80, 64, 5, 1, dict(qp=30, subpel=1, chroma_me=1, transform8x8=1, slice_rows=2), 3),
}


def main():
    orc = oracle_mod.load()
    out = ROOT / "tests" / "golden"
    out.mkdir(parents=True, exist_ok=True)
    for name, (w, h, R, refs, kw, pol) in CASES.items():
        if (out / f"{name}.npz").exists() and "--all" not in sys.argv:
            continue                                      # frozen: existing fixtures are only rewritten with --all
        cur, ref_l = synth.frame_pair(w, h, seed=11, search_range=R, num_refs=refs)
        with orc.context(width=w, height=h, search_range=R, num_refs=refs, pred_policy=pol, **kw) as ctx:
            n_mb = ctx.mb_w * ctx.mb_h
            pred = None
            if pol in (1, 2):
                pred = synth.random_pred(refs, n_mb, 1 if pol == 1 else 41, seed=5, max_qpel=4 * R + 20)
            for i, r in enumerate(ref_l):
                ctx.set_reference(i, r)
            cur_c = ref_c = np.zeros(0, np.uint8)
            if kw.get("chroma_me"):
                cur_c = np.stack([synth.gen_luma(w // 2, h // 2, 40 + k, "texture") for k in range(2)])
                ref_c = np.stack([np.stack([synth.gen_luma(w // 2, h // 2, 50 + 3 * i + k, "texture") for k in range(2)])
                                  for i in range(refs)])
                for i in range(refs):
                    ctx.set_reference_chroma(i, ref_c[i, 0], ref_c[i, 1])
                ctx.set_current_chroma(cur_c[0], cur_c[1])
            res, per = ctx.search_frame(cur, pred, per_ref=True)
            planes = None
            if kw.get("subpel"):
                planes = np.stack([ctx.get_subimage(0, fx, fy) for fy in range(4) for fx in range(4)])
                c = ctx.pad                                   # keep the fixture small: a 24x24 crop at the corner
                planes = planes[:, c - 8:c + 16, c - 8:c + 16].copy()
        np.savez_compressed(out / f"{name}.npz", cur=cur, refs=np.stack(ref_l), pred=pred if pred is not None else np.zeros(0, np.int16),
                            params=np.array([w, h, R, refs, pol], np.int32), kw_keys=np.array(list(kw.keys())),
                            kw_vals=np.array(list(kw.values()), np.int32), mv=res["mv"], cost=res["cost"], ref_idx=res["ref_idx"],
                            per_mv=per["mv"], per_cost=per["cost"], cur_c=cur_c, ref_c=ref_c,
                            planes_crop=planes if planes is not None else np.zeros(0, np.uint8))
        print(name, (out / f"{name}.npz").stat().st_size, "bytes")


if __name__ == "__main__":
    main()
