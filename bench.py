#!/usr/bin/env python
"""bench.py — ME macroblocks/sec on BASELINE config 3 (1080p, all 41 blocks, +-32 full search,
quarter-pel SATD refinement, 1 reference; synthetic YUV420 luma).

A step = one pass of the hot path over one frame pair:
    set_reference (border replication + 16 quarter-pel planes, a12)
  + search_frame  (integer 41-block search a6/a7 + sub-pel SATD refinement a10/a11 + reference choice)
  + (N > 1) all-gather of the MV field over NCCL/NVLink.
N ranks split the frame's 68 MB rows into contiguous stripes (strong scaling: total work fixed).

  value   MB/s with the frame pair already resident in HBM, device-timed with CUDA events
  e2e     MB/s through the C-ABI host-buffer calls (jmme_set_reference + jmme_search_frame) with
          pinned host buffers: H2D of both pictures and D2H of the MV field inside the timed region
  roofline / roofline_interp / cpu_baseline: see DESIGN.md §5

`--impl reference` times the CPU restatement of the reference algorithm (oracle/, kind "port":
/root/reference holds no sources) on all host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import pathlib
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / "h264-jm-commentary_b200"))

WORKLOADS = {
    # name: (width, height, R, refs, subpel, blocktype_mask)
    "1080p_r32_41blk_qpel_1ref": (1920, 1080, 32, 1, 1, 0xFE),          # BASELINE config 3 (headline)
    "720p_r32_41blk_int_1ref": (1280, 720, 32, 1, 0, 0xFE),             # config 2
    "1080p_r64_41blk_int_4ref": (1920, 1080, 64, 4, 0, 0xFE),           # config 4
    "cif_r16_16x16_int_1ref": (352, 288, 16, 1, 0, 0x02),               # config 1
    "2160p_r64_41blk_qpel_4ref": (3840, 2160, 64, 4, 1, 0xFE),          # config 5
}
OPS_PER_CAND_41 = 171      # 64 VABSDIFF4.ACC + 25 partition adds + 41 x (pack + min), SURVEY §8(d)
OPS_PER_CAND_16 = 66
QP = 28
SEEDS = (1, 2, 3)          # synthetic frame pairs; timed step s uses seed SEEDS[s % 3]


def config_of(args, rows=None):
    """The `config` object of the JSON line — identical in both arms (`--impl ours` / `--impl reference`); how
    an arm ran it (partition, launch mode, gather) goes into its own `run` object."""
    w, h, R, refs, subpel, mask = WORKLOADS[args.workload]
    mb_h, mb_w = (h + 15) // 16, (w + 15) // 16
    wave = args.pred_policy == "median"
    k = args.slice_rows or mb_h
    return {"workload": args.workload, "frame": f"{w}x{h}", "mbs": mb_h * mb_w, "search_range": R, "refs": refs,
            "blocks": 41 if mask != 0x02 else 1, "subpel": "half+quarter SATD" if subpel else "none", "qp": QP,
            "seeds": SEEDS,
            "pred_policy": ("zero" if not wave else f"in-frame median (JMME_PRED_MEDIAN), slice_rows={args.slice_rows}: "
                            f"slices of {k} MB rows, each a 2:1 wavefront")}


def policy_kw(args):
    """Context parameters of --pred-policy (3 = JMME_PRED_MEDIAN)."""
    return dict(pred_policy=3, slice_rows=args.slice_rows) if args.pred_policy == "median" else {}


def slice_unit(args, mb_h):
    """Stripes and CPU samples are whole slices under the median policy."""
    return (args.slice_rows or mb_h) if args.pred_policy == "median" else 1


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        busy = [s for s in sm if s > 0.5 * (max(mx) if mx else 1)] or sm
        pw = []
        for r in self.rows:
            try:
                pw.append(float(r[3]))
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(pw) if pw else None}


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _ncu_table():
    """Per-kernel counters of the committed `ncu --set full` capture of this command (profiles/r02_ncu.json, made
    by tools/ncu_key.py from the .ncu-rep; keys me_int / me_subpel / interp), else the round-1 file."""
    for name in ("r02_ncu.json", "r01_traffic.json"):
        f = ROOT / "profiles" / name
        if f.exists():
            d = json.loads(f.read_text())
            if name.startswith("r01"):
                d = {"me_int": d.get("me_int_tb_kernel"), "interp": d.get("interp_kernel"), "me_subpel": d.get("me_subpel_kernel")}
            return d
    return {}


def ncu_traffic(kernel):
    """DRAM bytes per launch (read + write) of `kernel` from the committed ncu capture, or None."""
    try:
        d = _ncu_table()[kernel]
        return d["dram_bytes_read"] + d["dram_bytes_write"]
    except Exception:  # noqa: BLE001
        return None


def ncu_metric(kernel, key):
    try:
        return _ncu_table()[kernel][key]
    except Exception:  # noqa: BLE001
        return None


def int_peaks(dev):
    """The two ceilings of the integer search:
    mix_peak       the search kernel's own per-candidate instruction mix (64 VABSDIFF4 + 66 IMAD + 41 min) issued
                   from registers by every SM, measured on this GPU right now by csrc/microbench as a SUSTAINED rate
                   (back-to-back launches until the clock has settled under the power cap, like bf16_tflops_sustained)
    issue_ceiling  SMs x 4 schedulers x 32 lanes x clocks.max.sm: what no kernel can exceed"""
    import torch
    prop = torch.cuda.get_device_properties(dev)
    max_mhz = 1965.0
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.max.sm", "--format=csv,noheader,nounits", "-i", str(dev)],
                             capture_output=True, text=True, timeout=20).stdout.strip()
        max_mhz = float(out.splitlines()[0])
    except Exception:  # noqa: BLE001
        pass
    res = {"issue_ceiling": prop.multi_processor_count * 128 * max_mhz * 1e6 * 1e-12}
    exe = ROOT / "h264-jm-commentary_b200" / "csrc" / "microbench"
    try:
        env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",")[dev]
                   if os.environ.get("CUDA_VISIBLE_DEVICES") else str(dev))
        d = json.loads(subprocess.run([str(exe), "120", "mix-only"], capture_output=True, text=True, timeout=120,
                                      check=True, env=env).stdout)
        m = d["mix_64sad_66imad_41min"]
        res.update(mix_peak=m["tera_lane_ops_per_s"], sm_mhz=m.get("mhz_nvml") or m.get("mhz_clock64"),
                   source="measured live, sustained (csrc/microbench: kernel-mix issue rate from registers, "
                          f"{m.get('launches')} back-to-back launches, SM clock {m.get('mhz_nvml')} MHz by NVML / "
                          f"{m.get('mhz_clock64')} MHz by clock64, {m.get('power_w_max')} W)")
    except Exception as e:  # noqa: BLE001
        for name in ("INT_PEAKS_r02.json", "INT_PEAKS_r01.json"):
            f = ROOT / "profiles" / name
            if f.exists():
                m = json.loads(f.read_text())["mix_64sad_66imad_41min"]
                res.update(mix_peak=m["tera_lane_ops_per_s"], sm_mhz=m.get("mhz_nvml"), source=f"profiles/{name} ({type(e).__name__})")
                break
        else:
            res.update(mix_peak=25.9, sm_mhz=None, source="fallback constant (round-1 measurement)")
    return res


# ------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """CPU arm: the oracle port on all host cores, bounded sample per step (rank 0 only)."""
    if rank != 0:
        return
    sys.path.insert(0, str(ROOT / "oracle"))
    import oracle as oracle_mod
    from jmme import synth
    orc = oracle_mod.load()
    w, h, R, refs, subpel, mask = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    orc.dll.jmme_oracle_set_threads.restype = C.c_int
    cores = orc.dll.jmme_oracle_set_threads(cores)
    cur, ref_l = synth.frame_pair(w, h, seed=1, search_range=R, num_refs=refs)
    mb_h = (h + 15) // 16
    mb_w = (w + 15) // 16
    # calibrate: one MB row, then size the per-step sample to ~args.ref_seconds of work
    rows = 1
    times = []
    with orc.context(width=w, height=h, search_range=R, num_refs=refs, subpel=subpel, blocktype_mask=mask, qp=QP,
                     mb_row_begin=mb_h // 2, mb_row_end=mb_h // 2 + 1) as c:      # calibration: zero predictors
        t0 = time.perf_counter()
        for i, r in enumerate(ref_l):
            c.set_reference(i, r)
        c.search_frame(cur)
        t_row = time.perf_counter() - t0
    unit = slice_unit(args, mb_h)
    rows = int(max(1, min(mb_h, args.ref_seconds / max(t_row, 1e-6))))
    rows = min(mb_h, -(-rows // unit) * unit)
    b = max(0, (mb_h - rows) // 2) // unit * unit
    with orc.context(width=w, height=h, search_range=R, num_refs=refs, subpel=subpel, blocktype_mask=mask, qp=QP,
                     mb_row_begin=b, mb_row_end=min(mb_h, b + rows), **policy_kw(args)) as c:
        for s in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            for i, r in enumerate(ref_l):
                c.set_reference(i, r)
            t1 = time.perf_counter()
            c.search_frame(cur)
            t2 = time.perf_counter()
            # the oracle always interpolates the whole reference: a sample of `rows` MB rows is charged its share
            dt = (t2 - t1) + (t1 - t0) * rows / mb_h
            if s >= args.warmup:
                times.append(dt)
    ms = 1e3 * float(np.mean(times))
    n_mb = rows * mb_w
    v = n_mb / (ms * 1e-3)
    sample = (f"{rows} of {mb_h} MB rows ({n_mb} MBs) of {args.workload} per step: search of those rows + {rows}/{mb_h} of the "
              f"time of interpolating the whole reference")
    print(json.dumps({
        "impl": "reference", "metric": "ME macroblocks/sec", "value": v, "unit": "MB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": config_of(args),
        "note": "CPU restatement (oracle/) of the JM path on all host cores; the mounted reference holds no sources, "
                "so this is a port, not JM itself; each step is a bounded sample of the workload (cpu_baseline.sample)",
        "cpu_baseline": {"value": v, "unit": "MB/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def cpu_baseline(args):
    """Single-thread oracle on a bounded sample (rank 0, N=1)."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import oracle as oracle_mod
    from jmme import synth
    orc = oracle_mod.load()
    orc.dll.jmme_oracle_set_threads.restype = C.c_int
    orc.dll.jmme_oracle_set_threads(1)
    w, h, R, refs, subpel, mask = WORKLOADS[args.workload]
    cur, ref_l = synth.frame_pair(w, h, seed=1, search_range=R, num_refs=refs)
    mb_h, mb_w = (h + 15) // 16, (w + 15) // 16

    unit = slice_unit(args, mb_h)

    def go(rows):
        rows = min(mb_h, -(-rows // unit) * unit)
        b = max(0, (mb_h - rows) // 2) // unit * unit
        with orc.context(width=w, height=h, search_range=R, num_refs=refs, subpel=subpel, blocktype_mask=mask, qp=QP,
                         mb_row_begin=b, mb_row_end=min(mb_h, b + rows), **policy_kw(args)) as c:
            t0 = time.perf_counter()
            for i, r in enumerate(ref_l):
                c.set_reference(i, r)
            t1 = time.perf_counter()
            c.search_frame(cur)
            # (the oracle interpolates the whole reference: the sample is charged its share of that)
            return (time.perf_counter() - t1) + (t1 - t0) * rows / mb_h
    t1 = go(1) / min(mb_h, unit)
    rows = int(max(1, min(mb_h, args.cpu_seconds / max(t1, 1e-6))))
    rows = min(mb_h, -(-rows // unit) * unit)
    t = go(rows)
    n = rows * mb_w
    return {"value": n / t, "unit": "MB/s", "cores": 1, "kind": "port",
            "sample": f"{rows} of {mb_h} MB rows ({n} MBs) of {args.workload}, {t:.1f} s, gcc -O2 single thread, "
                      f"host has {os.cpu_count()} cores"}


# ------------------------------------------------------------------------------------------------
def _watchdog(seconds):
    """Hard exit if the run has not finished in time: a hung collective must not stall the caller."""
    def fire():
        sys.stderr.write(f"bench.py watchdog: no exit after {seconds} s, leaving\n")
        sys.stderr.flush()
        os._exit(3)
    t = threading.Timer(seconds, fire)
    t.daemon = True
    t.start()


def oracle_parity(args, fields, search_mode=0):
    """Parity guard of the bench line: the field the GPU produced for every seed against the CPU oracle (all host
    threads, not timed).  Whole frame when the oracle needs about `--parity-seconds` or less for it, else MB rows
    spread over the frame (first, last and evenly in between, whole slices under the median policy).
    fields: {seed: numpy MBRESULT array of the whole frame}.  Returns the `parity` object; raises on a mismatch."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import oracle as oracle_mod
    from jmme import synth
    orc = oracle_mod.load()
    orc.dll.jmme_oracle_set_threads.restype = C.c_int
    cores = orc.dll.jmme_oracle_set_threads(0)
    w, h, R, refs, subpel, mask = WORKLOADS[args.workload]
    mb_h, mb_w = (h + 15) // 16, (w + 15) // 16
    unit = slice_unit(args, mb_h)
    kw = dict(width=w, height=h, search_range=R, num_refs=refs, subpel=subpel, blocktype_mask=mask, qp=QP, search_mode=search_mode,
              **policy_kw(args))

    def rows_of(seed, spans, timing=None):
        """oracle records of the MB-row spans [(rb, re)] of the frame pair of `seed` (one context per span: the stripe
        of an oracle context is fixed at creation, and every context interpolates the whole reference)"""
        cur, ref_l = synth.frame_pair(w, h, seed=seed, search_range=R, num_refs=refs)
        out = {}
        for rb, re in spans:
            with orc.context(mb_row_begin=rb, mb_row_end=re, **kw) as c:
                t0 = time.perf_counter()
                for i, r in enumerate(ref_l):
                    c.set_reference(i, r)
                t1 = time.perf_counter()
                out[(rb, re)] = c.search_frame(cur)[rb * mb_w:re * mb_w]
                if timing is not None:
                    timing += [t1 - t0, time.perf_counter() - t1]
        return out

    try:
        n_groups = -(-mb_h // unit)
        g0 = n_groups // 2
        tm = []
        first = rows_of(SEEDS[0], [(g0 * unit, min(mb_h, (g0 + 1) * unit))], tm)       # calibration: one group of rows
        t_ref, t_group = tm
        budget = args.parity_seconds / len(fields)                                        # per seed
        if t_ref + n_groups * t_group <= budget:
            spans, first = [(0, mb_h)], {}                                                # the whole frame, one context
        else:
            n = max(1, int(budget / (t_ref + t_group)))
            gs = sorted({round(i * (n_groups - 1) / max(n - 1, 1)) for i in range(n)})
            spans = [(g * unit, min(mb_h, (g + 1) * unit)) for g in gs]
        checked = mism = 0
        for seed, field in fields.items():
            exp = rows_of(seed, spans)
            if seed == SEEDS[0]:
                exp.update(first)
            for (rb, re), o in exp.items():
                g = field[rb * mb_w:re * mb_w]
                checked += len(o)
                if g.tobytes() != o.tobytes():
                    mism += int(np.count_nonzero(np.any(g["mv"] != o["mv"], axis=(1, 2)) | np.any(g["cost"] != o["cost"], axis=1) |
                                                 np.any(g["ref_idx"] != o["ref_idx"], axis=1))) or 1
    finally:
        orc.dll.jmme_oracle_set_threads(1)
    whole = spans == [(0, mb_h)]
    par = {"mbs": checked, "mismatches": mism, "blocks_per_mb": 41, "seeds": list(fields),
           "rows": "whole frame" if whole else f"{len(spans)} of {n_groups} row groups of {unit} MB rows, spread over the frame",
           "checker": f"oracle/libjmme_oracle.so, {cores} threads, every MV / ref_idx / cost byte of the records"}
    if mism:
        raise SystemExit(f"bench.py: GPU field differs from the oracle: {json.dumps(par)}")
    return par


EXTRAS = [
    # (workload, --pred-policy, --slice-rows, timed steps): the other BASELINE configs and the in-frame median policy,
    # measured in the default N = 1 run after the headline so that the driver's record holds them too
    ("cif_r16_16x16_int_1ref", "zero", 1, 10),
    ("720p_r32_41blk_int_1ref", "zero", 1, 10),
    ("1080p_r64_41blk_int_4ref", "zero", 1, 5),
    ("2160p_r64_41blk_qpel_4ref", "zero", 1, 3),
    ("1080p_r32_41blk_qpel_1ref", "median", 1, 5),
]


def measure_extra(args, lib, name, policy, slice_rows, steps, mix_peak, flush):
    """One short device-resident measurement of another workload (whole frame on this GPU): `steps` timed steps of
    set_reference + search with the L2 flushed in between, CUDA events per step; the field of the last step against
    the oracle on MB rows spread over the frame (whole frame when the oracle does it in a few seconds)."""
    import torch
    from jmme import synth
    from jmme.torch_api import DeviceSearch
    a = argparse.Namespace(**vars(args))
    a.workload, a.pred_policy, a.slice_rows, a.parity_seconds = name, policy, slice_rows, args.extra_parity_seconds
    w, h, R, refs, subpel, mask = WORKLOADS[name]
    mb_h, mb_w = (h + 15) // 16, (w + 15) // 16
    cur, ref_l = synth.frame_pair(w, h, seed=SEEDS[0], search_range=R, num_refs=refs)
    d_cur, d_refs = torch.from_numpy(cur).cuda(), [torch.from_numpy(r).cuda() for r in ref_l]
    mode = 1 if mask == 0x02 else 0                                   # config 1: FullPelBlockMotionSearch (search_mode FULL)
    ds = DeviceSearch(lib, width=w, height=h, search_range=R, num_refs=refs, subpel=subpel, blocktype_mask=mask, qp=QP,
                      search_mode=mode, **policy_kw(a))
    try:
        def step():
            for i, r in enumerate(d_refs):
                ds.set_reference(i, r)
            return ds.search(d_cur)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        ms = []
        for s in range(steps):
            flush.fill_(s & 255)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = step()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        ms_step = float(np.mean(ms))
        ds.ctx.set_profiling(True)
        step()
        torch.cuda.synchronize()
        kt = ds.ctx.kernel_times()
        field = ds.to_numpy(out).copy()
        n_mb = mb_h * mb_w
        ncand = (2 * R + 1) ** 2
        ops = n_mb * refs * ncand * (OPS_PER_CAND_16 if mask == 0x02 else OPS_PER_CAND_41)
        wave = policy == "median"
        k_int = (ms_step - kt["interp"] * refs) if wave else kt["me_int"]
        res = {"config": config_of(a), "value": n_mb / (ms_step * 1e-3), "unit": "MB/s", "ms_per_step": ms_step, "steps": steps,
               "kernel_ms": {"interp": kt["interp"], "me_int": k_int, "me_subpel": kt["me_subpel"], "select_ref": kt["select"]},
               "kernel_instance": ds.ctx.last_kernel(),
               "roofline_frac": ops / (max(k_int, 1e-9) * 1e-3) * 1e-12 / mix_peak}
        if not args.no_parity:
            a.search_mode = mode
            res["parity"] = oracle_parity(a, {SEEDS[0]: field}, search_mode=mode)
        return res
    finally:
        ds.close()


def run_ours(args, rank, world, local_rank):
    _watchdog(args.watchdog)
    import torch
    import torch.distributed as dist
    import jmme
    from jmme import abi, synth
    from jmme.dist import PeerPushGather, StripeGather, stripe_of
    from jmme.torch_api import DeviceSearch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = jmme.load()
    w, h, R, refs, subpel, mask = WORKLOADS[args.workload]
    mb_h, mb_w = (h + 15) // 16, (w + 15) // 16
    n_mb = mb_h * mb_w
    rb, re = stripe_of(rank, world, mb_h, slice_unit(args, mb_h))
    pairs = [synth.frame_pair(w, h, seed=sd, search_range=R, num_refs=refs) for sd in SEEDS]
    # the step's inputs live in these device buffers; the frame pair of step s (seed SEEDS[s % 3]) is copied
    # into them before the timed region of the step starts
    d_pairs = [(torch.from_numpy(c).cuda(), [torch.from_numpy(r).cuda() for r in rl]) for c, rl in pairs]
    d_cur = d_pairs[0][0].clone()
    d_refs = [r.clone() for r in d_pairs[0][1]]

    def load_pair(i):
        d_cur.copy_(d_pairs[i][0])
        for a, b_ in zip(d_refs, d_pairs[i][1]):
            a.copy_(b_)
    if re <= rb:
        raise SystemExit(f"rank {rank}: empty stripe — {mb_h} MB rows cannot feed {world} ranks of {-(-mb_h // world)} rows")
    ds = DeviceSearch(lib, width=w, height=h, search_range=R, num_refs=refs, subpel=subpel, blocktype_mask=mask, qp=QP,
                      mb_row_begin=rb, mb_row_end=re, **policy_kw(args))
    ds.ctx.set_profiling(True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")          # > 126 MB L2
    rec = abi.MBRESULT_DTYPE.itemsize
    # the search writes its stripe straight into the field; N > 1: gathered either by peer stores over
    # NVLink into symmetric memory (default) or by one in-place NCCL all-gather (--gather nccl)
    gather, gather_mode = None, "none (1 rank)"
    fused = False
    if world > 1 and args.gather in ("p2p", "p2p-peer", "p2p-push"):
        try:
            gather = PeerPushGather(mb_w, mb_h, "cuda", unit=slice_unit(args, mb_h))
            fused = args.gather in ("p2p", "p2p-peer")
            if fused:
                gather.attach(ds, multicast=args.gather == "p2p")
                gather_mode = ("one multimem.st per record word through the NVLS multicast mapping of the symmetric field, from "
                               "the kernels that write the records (jmme_set_multicast_field_dev) + symm-mem barriers"
                               if gather.mode == "multicast" else
                               "peer stores into symmetric memory from the kernels that write the records "
                               "(jmme_set_peer_fields_dev) + symm-mem barriers")
            else:
                gather_mode = "peer stores into symmetric memory (jmme_push_stripe_dev) + symm-mem barrier"
        except Exception as e:  # noqa: BLE001
            sys.stderr.write(f"rank {rank}: symmetric memory unavailable ({type(e).__name__}: {e}); using NCCL\n")
            gather, fused = None, False
    if gather is None:
        gather = StripeGather(mb_w, mb_h, "cuda", unit=slice_unit(args, mb_h))
        if world > 1:
            gather_mode = "in-place NCCL all_gather_into_tensor"
    p2p = isinstance(gather, PeerPushGather)

    def step_device():
        for i, r in enumerate(d_refs):
            ds.set_reference(i, r)
        if fused:
            gather.pre()
            ds.search(d_cur, out=gather.field)
            return gather.post()
        ds.search(d_cur, out=gather.field)
        return gather.gather(ds) if p2p else gather.gather()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- device-resident timing ----------------------------------------------------------------
    sampler = ClockSampler(local_rank)      # started early: nvidia-smi needs a moment before its first line
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    l_a = ds.launch_count()
    step_device()
    launches_per_step = ds.launch_count() - l_a          # kernels of one step (a graph replay launches the same ones)
    kernel_instance = ds.ctx.last_kernel()
    barrier()
    # per-kernel device times (CUDA events around each kernel, eager launches) for the roofline
    ktimes = []
    for s in range(min(args.steps, 10)):
        load_pair(s % len(SEEDS))
        flush.fill_(s & 255)
        step_device()
        torch.cuda.synchronize()
        ktimes.append(ds.ctx.kernel_times())
    ds.ctx.set_profiling(False)
    # the whole step (kernels + the all-gather) is captured once in a CUDA graph and replayed
    launch_mode, graph = "eager", None
    if not args.no_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step_device()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                step_device()
            launch_mode = "cuda-graph replay of the captured step"
        except Exception as e:  # noqa: BLE001
            graph, launch_mode = None, f"eager (graph capture failed: {type(e).__name__})"
            try:
                torch.cuda.synchronize()
            except Exception:  # noqa: BLE001
                pass
    run_step = graph.replay if graph is not None else step_device
    for _ in range(3):
        run_step()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    l0 = ds.launch_count()
    barrier()
    fields = {}
    for s in range(args.steps):
        load_pair(s % len(SEEDS))         # this step's frame pair, resident in HBM before the timed region
        flush.fill_(s & 255)              # L2 flush between timed iterations (outside the event pair)
        barrier()
        ev[s][0].record()
        run_step()
        ev[s][1].record()
        torch.cuda.synchronize()
        if rank == 0 and SEEDS[s % len(SEEDS)] not in fields and not args.no_parity:
            fields[SEEDS[s % len(SEEDS)]] = ds.to_numpy(gather.frame()).copy()      # the timed step's own output
    barrier()
    launches = ds.launch_count() - l0
    if graph is not None:                 # replays launch the captured kernels without passing the counter
        launches = args.steps * launches_per_step
    step_ms = [a.elapsed_time(b) for a, b in ev]
    ms_dev = float(np.mean(step_ms))

    # ---- end to end through the C-ABI host calls (pinned host buffers) -------------------------
    h_pairs = [(torch.from_numpy(c).pin_memory(), [torch.from_numpy(r).pin_memory() for r in rl]) for c, rl in pairs]
    h_out = torch.zeros(n_mb * rec, dtype=torch.uint8).pin_memory()
    hctx = lib.context(width=w, height=h, search_range=R, num_refs=refs, subpel=subpel, blocktype_mask=mask, qp=QP,
                       mb_row_begin=rb, mb_row_end=re, device_ids=[local_rank], async_reference=1, **policy_kw(args))
    pu8 = C.POINTER(C.c_uint8)

    def step_host(i):
        h_cur, h_refs = h_pairs[i]
        for k, r in enumerate(h_refs):
            lib.check(lib.dll.jmme_set_reference(hctx.handle, k, C.cast(r.data_ptr(), pu8), w), hctx.handle)
        lib.check(lib.dll.jmme_search_frame(hctx.handle, C.cast(h_cur.data_ptr(), pu8), w, None,
                                            C.c_void_p(h_out.data_ptr()), None), hctx.handle)

    # warm-up: every pinned frame pair several times — the first DMA transfers out of a pinned buffer are slow (0.9 ms
    # for the first step from a fresh buffer, 0.6, 0.58, 0.56, ... for the next ones; measured), whichever buffer it is
    e2e_warm = max(args.warmup, 8)           # (a buffer keeps getting faster over its first ~8 transfers)
    for _ in range(e2e_warm):
        for i in range(len(SEEDS)):
            step_host(i)
    e2e_t = []
    barrier()
    last_seed_i = 0
    for s in range(args.steps):
        flush.fill_(s & 255)
        barrier()
        last_seed_i = s % len(SEEDS)
        t0 = time.perf_counter()
        step_host(last_seed_i)
        e2e_t.append(time.perf_counter() - t0)
    barrier()
    ms_e2e = 1e3 * float(np.mean(e2e_t))
    launches_e2e = hctx.launch_count()
    clocks = sampler.stop() if rank == 0 else None

    # the host path and the device path must agree byte for byte on this rank's stripe (same frame pair)
    load_pair(last_seed_i)
    step_device()
    torch.cuda.synchronize()
    got = ds.to_numpy(gather.frame())[rb * mb_w:re * mb_w]
    exp = h_out.numpy().view(abi.MBRESULT_DTYPE)[rb * mb_w:re * mb_w]
    assert got.tobytes() == exp.tobytes(), "device-resident and host-buffer paths disagree"

    # the gathered field must be the same on every rank and equal to an independent NCCL gather
    if world > 1:
        chk = StripeGather(mb_w, mb_h, "cuda", unit=slice_unit(args, mb_h))
        chk.field[rb * mb_w:re * mb_w].copy_(gather.field[rb * mb_w:re * mb_w])
        ref_field = chk.gather()
        torch.cuda.synchronize()
        assert torch.equal(ref_field, gather.frame()), f"rank {rank}: gathered MV field differs from the NCCL gather"
        chk = None

    # ---- max over ranks -------------------------------------------------------------------------
    t = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = t.tolist()

    if rank == 0:
        ncand = (2 * R + 1) ** 2
        ops_cand = OPS_PER_CAND_16 if mask == 0x02 else OPS_PER_CAND_41
        k_int = float(np.mean([k["me_int"] for k in ktimes]))
        k_sub = float(np.mean([k["me_subpel"] for k in ktimes]))
        k_itp = float(np.mean([k["interp"] for k in ktimes]))
        k_sel = float(np.mean([k["select"] for k in ktimes]))
        peaks = int_peaks(local_rank)
        hbm, hbm_src = measured_peaks()
        alg_ops = (re - rb) * mb_w * refs * ncand * ops_cand          # this rank's launch
        wave = args.pred_policy == "median"
        if wave:            # no per-kernel brackets inside the wavefront: the whole search (all steps) is the "launch"
            k_int = max(ms_dev - k_itp * refs, 1e-9)
        achieved = alg_ops / (k_int * 1e-3) * 1e-12
        pad = hctx.pad
        rows_itp = ((16 * mb_h + 2 * pad) if re == mb_h else min(16 * mb_h + 2 * pad, pad + 16 * re + 2 * R + 4)) - \
            (0 if rb == 0 else max(0, pad + 16 * rb - 2 * R - 4))
        itp_bytes = 17 * rows_itp * (16 * mb_w + 2 * pad)
        itp_gbs = itp_bytes / (k_itp * 1e-3) * 1e-9 if k_itp > 0 else None
        # rows the host path really uploads for this rank's stripe (jmme_api.cu: set_reference / search_frame)
        yb = 0 if rb == 0 else max(0, pad + 16 * rb - 2 * R - 4)
        ye = (16 * mb_h + 2 * pad) if re == mb_h else min(16 * mb_h + 2 * pad, pad + 16 * re + 2 * R + 4)
        ref_rows = min(max(ye - pad + 3, 1), h) - min(max(yb - pad - 3, 0), h - 1)
        cur_rows = max(min(16 * re, h) - min(16 * rb, h - 1), 1)
        h2d_bytes = (refs * ref_rows + cur_rows) * w
        pad_ = (2 * R + 16 + 15) & ~15
        plane_mb = (16 if subpel else 1) * (((w + 15) & ~15) + 2 * pad_) * (((h + 15) & ~15) + 2 * pad_) / 1e6
        headline = world == 1 and not wave and args.workload.startswith("1080p_r32")
        line = {
            "metric": "ME macroblocks/sec", "value": n_mb / (ms_dev * 1e-3), "unit": "MB/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev, "ms_steps_rank0": [round(x, 4) for x in step_ms],
            "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": config_of(args, re - rb),
            "run": {"partition": f"{world} MB-row stripes",
                    "l2": "256 MB buffer written between timed steps (outside the event pair)",
                    "timing": "CUDA events per step on the launching stream, mean over steps, max over ranks",
                    "launch": launch_mode, "gather": gather_mode},
            "e2e": {"value": n_mb / (ms_e2e * 1e-3), "unit": "MB/s", "ms_per_step": ms_e2e,
                    "ms_steps_rank0": [round(1e3 * x, 4) for x in e2e_t],
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": (re - rb) * mb_w * rec,
                    "api": "jmme_set_reference + jmme_search_frame (C ABI, pinned host buffers)",
                    "warmup": f"{e2e_warm} untimed steps from each of the {len(SEEDS)} pinned frame pairs"},
            "gpu_launches": int(launches + launches_e2e),
            "kernel_ms": {"interp": k_itp, "me_int": k_int, "me_subpel": k_sub, "select_ref": k_sel,
                          "share_me_int": k_int / max(k_itp + k_int + k_sub + k_sel, 1e-9)},
            "roofline": {"bound": "int_alu", "achieved": achieved, "peak": peaks["mix_peak"], "unit": "Tlane-op/s",
                         "frac": achieved / peaks["mix_peak"],
                         # the same achieved rate against the two other ceilings the review asked for
                         "frac_of_issue_ceiling": achieved / peaks["issue_ceiling"],
                         "issue_ceiling": peaks["issue_ceiling"],
                         "alu_pipe_pct_ncu": ncu_metric("me_int", "alu_pipe_pct") if headline else None,
                         "traffic": ncu_traffic("me_int") if headline else None,
                         "kernel": ("wavefront step chain (me_int_tb_kernel clusters + me_subpel_kernel), whole search" if wave
                                    else kernel_instance.split("<")[0]),
                         "kernel_instance": kernel_instance,
                         "algorithmic_ops_per_candidate": ops_cand, "peak_source": peaks["source"],
                         "sm_mhz_microbench": peaks.get("sm_mhz")},
            "roofline_interp": {"bound": "alu (the planes stay in the 126 MB L2; HBM figure for reference)", "achieved": itp_gbs,
                                "peak": hbm, "unit": "GB/s", "frac": (itp_gbs / hbm) if itp_gbs else None,
                                "traffic": ncu_traffic("interp") if (world == 1 and args.workload.startswith("1080p_r32")) else None,
                                "note": (f"{plane_mb:.1f} MB of planes per reference are written; they stay in the 126 MB L2 for the "
                                         "sub-pel kernel when they fit" if subpel else "integer plane only"),
                                "kernel": "interp_kernel", "algorithmic_bytes_per_pixel": 17, "peak_source": hbm_src},
            "clocks": clocks,
        }
        if fields:
            line["parity"] = oracle_parity(args, fields)
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(args)
        if world == 1 and headline and not args.no_extras:
            graph = run_step = None                       # the headline's contexts and graph make room first
            hctx.close()
            ds.close()
            extras = []
            for name, policy, srows, steps in EXTRAS:
                try:
                    extras.append(measure_extra(args, lib, name, policy, srows, steps, peaks["mix_peak"], flush))
                except SystemExit:
                    raise
                except Exception as e:  # noqa: BLE001
                    extras.append({"config": {"workload": name, "pred_policy": policy}, "error": f"{type(e).__name__}: {e}"})
            line["other_workloads"] = extras
        print(json.dumps(line), flush=True)
    # Orderly exit through the interpreter (the driver records the loaded .so files at exit).  A captured graph
    # holds the step's kernels (and NCCL / symmetric-memory work): it goes first, then the contexts, every rank
    # meets at a barrier and the process group is destroyed; the watchdog bounds all of it.
    sys.stdout.flush()
    graph = run_step = None
    torch.cuda.synchronize()
    hctx.close()
    ds.close()
    if world > 1:
        dist.barrier()
        gather = None
        torch.cuda.synchronize()
        dist.destroy_process_group()
    sys.stdout.flush()
    sys.stderr.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="1080p_r32_41blk_qpel_1ref", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work of the cpu_baseline sample")
    ap.add_argument("--ref-seconds", type=float, default=4.0, help="CPU work per step of --impl reference")
    ap.add_argument("--pred-policy", default="zero", choices=["zero", "median"],
                    help="median: JMME_PRED_MEDIAN, the predictor loop closed inside the frame (a wavefront)")
    ap.add_argument("--slice-rows", type=int, default=1, help="MB rows per slice of --pred-policy median (0 = whole frame)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle comparison of the timed steps' output")
    ap.add_argument("--parity-seconds", type=float, default=20.0,
                    help="CPU time budget of the oracle comparison (whole frame if it fits, else spread rows)")
    ap.add_argument("--no-extras", action="store_true", help="N = 1: skip the short measurements of the other BASELINE configs")
    ap.add_argument("--extra-parity-seconds", type=float, default=6.0, help="oracle time budget per extra workload")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "p2p-peer", "p2p-push", "nccl"],
                    help="N > 1: how the MV field is gathered: stores from the search kernels through the NVLS multicast "
                         "mapping (p2p; peer stores when there is none), peer stores from the search kernels (p2p-peer), "
                         "from a separate push kernel (p2p-push), or one NCCL all-gather")
    ap.add_argument("--watchdog", type=float, default=300.0, help="hard exit after this many seconds")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if world == 1 and args.gpus > 1 and args.impl == "ours":
        # launched without torchrun: re-exec under it
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29511", __file__] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
