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
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/r01_traffic.json), or None."""
    try:
        d = json.loads((ROOT / "profiles" / "r01_traffic.json").read_text())[kernel]
        return d["dram_bytes_read"] + d["dram_bytes_write"]
    except Exception:  # noqa: BLE001
        return None


def int_peak_live():
    """Integer issue peak measured on this GPU right now by csrc/microbench (same instruction mix as the
    search kernel: 64 VABSDIFF4 + 66 IMAD + 41 VIMNMX per candidate)."""
    exe = ROOT / "h264-jm-commentary_b200" / "csrc" / "microbench"
    try:
        d = json.loads(subprocess.run([str(exe)], capture_output=True, text=True, timeout=120, check=True).stdout)
        return d["mix_64sad_66imad_41min"]["tera_lane_ops_per_s"], "measured live (csrc/microbench, kernel-mix issue rate)", d
    except Exception as e:  # noqa: BLE001
        p = ROOT / "profiles" / "INT_PEAKS_r01.json"
        if p.exists():
            d = json.loads(p.read_text())
            return d["mix_64sad_66imad_41min"]["tera_lane_ops_per_s"], f"profiles/INT_PEAKS_r01.json ({type(e).__name__})", d
        return 25.9, "fallback constant", {}


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
            c.search_frame(cur)
            dt = time.perf_counter() - t0
            if s >= args.warmup:
                times.append(dt)
    ms = 1e3 * float(np.mean(times))
    n_mb = rows * mb_w
    v = n_mb / (ms * 1e-3)
    sample = f"{rows} of {mb_h} MB rows ({n_mb} MBs) of {args.workload} per step, interpolation of the whole reference included"
    print(json.dumps({
        "impl": "reference", "metric": "ME macroblocks/sec", "value": v, "unit": "MB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": args.workload, "qp": QP, "pred_policy": args.pred_policy, "note": "CPU restatement (oracle/) of the JM path; the mounted "
                   "reference holds no sources, so this is a port, not JM itself"},
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
            c.search_frame(cur)
            return time.perf_counter() - t0
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
    cur, ref_l = synth.frame_pair(w, h, seed=1, search_range=R, num_refs=refs)
    d_cur = torch.from_numpy(cur).cuda()
    d_refs = [torch.from_numpy(r).cuda() for r in ref_l]
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
    if world > 1 and args.gather in ("p2p", "p2p-push"):
        try:
            gather = PeerPushGather(mb_w, mb_h, "cuda", unit=slice_unit(args, mb_h))
            fused = args.gather == "p2p"
            if fused:
                gather.attach(ds)
                gather_mode = ("peer stores into symmetric memory from the kernels that write the records "
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
    barrier()
    # per-kernel device times (CUDA events around each kernel, eager launches) for the roofline
    ktimes = []
    for s in range(min(args.steps, 10)):
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
    for s in range(args.steps):
        flush.fill_(s & 255)              # L2 flush between timed iterations (outside the event pair)
        barrier()
        ev[s][0].record()
        run_step()
        ev[s][1].record()
        torch.cuda.synchronize()
    barrier()
    launches = ds.launch_count() - l0
    if graph is not None:                 # replays launch the captured kernels without passing the counter
        launches = args.steps * launches_per_step
    step_ms = [a.elapsed_time(b) for a, b in ev]
    ms_dev = float(np.mean(step_ms))

    # ---- end to end through the C-ABI host calls (pinned host buffers) -------------------------
    h_cur = torch.from_numpy(cur).pin_memory()
    h_refs = [torch.from_numpy(r).pin_memory() for r in ref_l]
    h_out = torch.zeros(n_mb * rec, dtype=torch.uint8).pin_memory()
    hctx = lib.context(width=w, height=h, search_range=R, num_refs=refs, subpel=subpel, blocktype_mask=mask, qp=QP,
                       mb_row_begin=rb, mb_row_end=re, device_ids=[local_rank], async_reference=1, **policy_kw(args))
    pu8 = C.POINTER(C.c_uint8)

    def step_host():
        for i, r in enumerate(h_refs):
            lib.check(lib.dll.jmme_set_reference(hctx.handle, i, C.cast(r.data_ptr(), pu8), w), hctx.handle)
        lib.check(lib.dll.jmme_search_frame(hctx.handle, C.cast(h_cur.data_ptr(), pu8), w, None,
                                            C.c_void_p(h_out.data_ptr()), None), hctx.handle)

    for _ in range(max(args.warmup, 3)):
        step_host()
    e2e_t = []
    barrier()
    for s in range(args.steps):
        flush.fill_(s & 255)
        barrier()
        t0 = time.perf_counter()
        step_host()
        e2e_t.append(time.perf_counter() - t0)
    barrier()
    ms_e2e = 1e3 * float(np.mean(e2e_t))
    launches_e2e = hctx.launch_count()
    clocks = sampler.stop() if rank == 0 else None

    # parity guard: the host path and the device path must agree byte for byte on this rank's stripe
    got = ds.to_numpy(gather.frame())[rb * mb_w:re * mb_w]
    exp = h_out.numpy().view(abi.MBRESULT_DTYPE)[rb * mb_w:re * mb_w]
    assert got.tobytes() == exp.tobytes(), "device-resident and host-buffer paths disagree"

    # the gathered field must be the same on every rank and equal to an independent NCCL gather
    if world > 1:
        step_device()
        torch.cuda.synchronize()
        chk = StripeGather(mb_w, mb_h, "cuda", unit=slice_unit(args, mb_h))
        chk.field[rb * mb_w:re * mb_w].copy_(gather.field[rb * mb_w:re * mb_w])
        ref_field = chk.gather()
        torch.cuda.synchronize()
        assert torch.equal(ref_field, gather.frame()), f"rank {rank}: gathered MV field differs from the NCCL gather"

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
        peak, peak_src, _ = int_peak_live()
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
        line = {
            "metric": "ME macroblocks/sec", "value": n_mb / (ms_dev * 1e-3), "unit": "MB/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": args.workload, "frame": f"{w}x{h}", "mbs": n_mb, "search_range": R, "refs": refs,
                       "blocks": 41 if mask != 0x02 else 1, "subpel": "half+quarter SATD" if subpel else "none",
                       "qp": QP, "pred_policy": ("zero" if not wave else
                                                 f"in-frame median (JMME_PRED_MEDIAN), slice_rows={args.slice_rows}: "
                                                 f"a 2:1 wavefront of {mb_w + 2 * (min(args.slice_rows or mb_h, re - rb) - 1)} steps"),
                       "partition": f"{world} MB-row stripes",
                       "l2": "256 MB buffer written between timed steps (outside the event pair)",
                       "timing": "CUDA events per step on the launching stream, mean over steps, max over ranks",
                       "launch": launch_mode, "gather": gather_mode},
            "e2e": {"value": n_mb / (ms_e2e * 1e-3), "unit": "MB/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": (re - rb) * mb_w * rec,
                    "api": "jmme_set_reference + jmme_search_frame (C ABI, pinned host buffers)"},
            "gpu_launches": int(launches + launches_e2e),
            "kernel_ms": {"interp": k_itp, "me_int": k_int, "me_subpel": k_sub, "select_ref": k_sel,
                          "share_me_int": k_int / max(k_itp + k_int + k_sub + k_sel, 1e-9)},
            "roofline": {"bound": "int_alu", "achieved": achieved, "peak": peak, "unit": "Tlane-op/s",
                         "frac": achieved / peak,
                         "traffic": ncu_traffic("me_int_tb_kernel") if (world == 1 and not wave and args.workload.startswith("1080p_r32")) else None,
                         "kernel": ("wavefront step chain (me_int_tb_kernel clusters + me_subpel_kernel), whole search" if wave
                                    else "me_int_tb_kernel" if (R <= 32 and mask != 0x02) else "me_int_kernel"),
                         "algorithmic_ops_per_candidate": ops_cand, "peak_source": peak_src},
            "roofline_interp": {"bound": "hbm", "achieved": itp_gbs, "peak": hbm, "unit": "GB/s",
                                "frac": (itp_gbs / hbm) if itp_gbs else None,
                                "traffic": ncu_traffic("interp_kernel") if (world == 1 and args.workload.startswith("1080p_r32")) else None,
                                "note": (f"{plane_mb:.1f} MB of planes per reference are written; they stay in the 126 MB L2 for the "
                                         "sub-pel kernel when they fit" if subpel else "integer plane only"),
                                "kernel": "interp_kernel", "algorithmic_bytes_per_pixel": 17, "peak_source": hbm_src},
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(args)
        print(json.dumps(line), flush=True)
    # Orderly exit.  A captured graph holds NCCL work; destroying the process group with it alive was seen
    # to hang, so the graph goes first, every rank meets at a barrier, and the process leaves without
    # running the communicator's destructor (a watchdog bounds everything above in case a rank is stuck).
    sys.stdout.flush()
    graph = None
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


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
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "p2p-push", "nccl"],
                    help="N > 1: how the MV field is gathered: peer stores from the search kernels (p2p), from a "
                         "separate push kernel (p2p-push), or one NCCL all-gather")
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
