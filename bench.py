#!/usr/bin/env python
"""bench.py -- throughput of the ASW hot path (raw cost + 4 support tables + r x (V,H) + WTA).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3|cfg4|cfg5|cfg2] [--impl reference]

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over this rank's
batch of synthetic stereo pairs.  Metric: Mpix*disp/s = W*H*D*pairs / time (the reference
thesis' own "disparities per second" metric).
  value     : device-resident inputs, CUDA-event timed on the library's stream, max over ranks
  e2e       : the same through asw_disparity() with pinned HOST buffers (H2D + D2H inside)
  roofline  : dominant kernel (the slower of the V / H aggregation passes) vs the FP32 peak
  cpu_baseline : the CPU oracle (port of the reference kernels) on a bounded sample, rank 0, N=1
Multi-GPU: pair sharding (each rank owns its pairs, weak scaling, no data-path collective;
one all-gather of the uint8 disparity maps at the end of a step) -- or, with --workload cfg4,
row-band sharding of one 4K frame with a shrinking halo (strong scaling).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (synth config, pairs per rank, description)
    "cfg2": ("cfg2_teddy_shape", 1, "synthetic 450x375, 61 disparities (Middlebury teddy/cones shape)"),
    "cfg3": ("cfg3_1800x1500_d256", 1, "synthetic 1800x1500 pair, 256 disparities, 33-tap window, r=7"),
    "cfg4": ("cfg4_3840x2160_d256", 1, "synthetic 3840x2160 pair, 256 disparities, row-band sharded"),
    "cfg5": ("cfg5_1280x720_d128", 4, "synthetic 1280x720 pairs, 128 disparities, pair sharded"),
}
T_TAPS = 33
FP32_PEAK_FALLBACK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12   # SMs x lanes x 2 x sm_max_mhz (MEASURED_PEAKS.json)


def alg_flops(W, H, D, r):          # 8*T*r*W*H*D
    return 8.0 * T_TAPS * r * W * H * D


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def oracle_sample(L, R, D, r, rows):
    """Times the CPU oracle on the top `rows` rows of the workload (bounded sample)."""
    from oracle import asw_oracle as O
    O.lib()
    # all host threads this process may use: torchrun exports OMP_NUM_THREADS=1 to its workers, which would
    # otherwise turn the 16-core baseline into a single-threaded one
    try:
        ncpu = len(os.sched_getaffinity(0))
    except AttributeError:
        ncpu = os.cpu_count() or 1
    O.set_num_threads(ncpu)
    rows = min(rows, L.shape[0])
    Ls, Rs = np.ascontiguousarray(L[:rows]), np.ascontiguousarray(R[:rows])
    p = O.OracleParams(ndisp=D, iterations=r)
    t0 = time.perf_counter()
    res = O.asw_hot_path(Ls, Rs, p, use_fma=True)
    dt = time.perf_counter() - t0
    W = L.shape[1]
    return {"value": W * rows * D / dt / 1e6, "unit": "Mpix*disp/s", "cores": O.num_threads(), "kind": "port",
            "sample": f"top {rows} of {L.shape[0]} rows of the same pair ({W}x{rows}x{D}, r={r}), {dt:.1f} s, OpenMP oracle/asw_oracle.c",
            "seconds": dt}, res


def run_reference_arm(args, rank):
    """--impl reference: the reference's CPU implementation of the path (OpenMP port of its kernels;
    the OpenCL original cannot run here) on all host threads, bounded sample per step."""
    if rank != 0:
        return
    from stereo_matchin_b200 import synth
    cfg, _, desc = WORKLOADS[args.workload]
    L, R, _, D = synth.make_config(cfg)
    H, W, _ = L.shape
    rows = max(33, min(H, int(args.ref_rows)))
    times = []
    info = None
    for i in range(args.warmup_ref + args.steps):
        info, _ = oracle_sample(L, R, D, args.iterations, rows)
        if i >= args.warmup_ref:
            times.append(info["seconds"])
    ms = float(np.mean(times)) * 1e3
    v = W * rows * D / (ms * 1e-3) / 1e6
    line = {"impl": "reference", "metric": "Mpix*disp/s (ASW agg+WTA)", "value": v, "unit": "Mpix*disp/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup_ref, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}", "W": W, "H": H, "ndisp": D, "iterations": args.iterations,
                       "sample_rows": rows},
            "cpu_baseline": {"value": v, "unit": "Mpix*disp/s", "cores": info["cores"], "kind": "port", "sample": info["sample"]},
            "e2e": {"value": v, "unit": "Mpix*disp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--iterations", type=int, default=7)
    ap.add_argument("--grid", default="", help="cfg4: force the (row bands)x(disparity shards) grid, e.g. 4x2")
    ap.add_argument("--bands-only", action="store_true", help="cfg4: shard by row bands only (default: row bands x disparity shards)")
    ap.add_argument("--family", type=int, default=0, help="0 = TMA-fed kernels (default), 1 = basic kernels, 2 = tiled kernels without TMA")
    ap.add_argument("--cpu-rows", type=int, default=400, help="rows of the pair the CPU baseline sample covers")
    ap.add_argument("--ref-rows", type=int, default=400)
    ap.add_argument("--warmup-ref", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    from stereo_matchin_b200 import api, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback exists)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    cfg, pairs_per_rank, desc = WORKLOADS[args.workload]
    W, H, D, _ = synth.CONFIGS[cfg]
    r = args.iterations
    band_mode = args.workload == "cfg4"
    params = api.AswParams(ndisp=D, iterations=r)
    ctx = api.AswContext(local_rank)
    ctx.set_kernel_family(args.family)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local_rank))

    # ---- inputs: pinned host copies and device-resident copies --------------------------------
    dshard, grid = None, (world, 1)
    if band_mode:
        from stereo_matchin_b200 import sharding
        pairs = [synth.make_config(cfg, 0)[:2]]                      # every rank holds the full frame
        grid = sharding.shard_grid(world, D) if args.family == 0 and not args.bands_only else (world, 1)
        if args.grid:
            grid = tuple(int(v) for v in args.grid.lower().split("x"))
            assert grid[0] * grid[1] == world, "--grid must multiply to the number of ranks"
        if H % grid[0]:
            grid = (world, 1)
        if grid[1] > 1:     # 2-D grid: row bands x disparity shards (no halo work along d), ranks of a band are consecutive
            band, dshard, _, _ = sharding.rank_shard(rank, world, H, D, grid)
        else:
            band = ((H * rank) // world, (H * (rank + 1)) // world)
        out_rows = band[1] - band[0]
    else:
        pairs = [synth.make_config(cfg, rank * pairs_per_rank + i)[:2] for i in range(pairs_per_rank)]
        band, out_rows = None, H
    host_in = [(torch.from_numpy(L).pin_memory(), torch.from_numpy(R).pin_memory()) for L, R in pairs]
    host_d = [torch.empty((out_rows, W), dtype=torch.uint8).pin_memory() for _ in pairs]
    host_full_d = [torch.empty((H, W), dtype=torch.uint8).pin_memory() for _ in pairs]
    dev_in = [(a.cuda(non_blocking=False), b.cuda(non_blocking=False)) for a, b in host_in]
    dev_d = [torch.empty((out_rows, W), dtype=torch.uint8, device="cuda") for _ in pairs]
    gather = torch.empty((world, len(pairs), out_rows, W), dtype=torch.uint8, device="cuda") if world > 1 and not band_mode else None
    gather_band = torch.empty((H, W), dtype=torch.uint8, device="cuda") if world > 1 and band_mode else None
    nd = grid[1]
    if dshard is not None:  # partial WTA triples of this rank, the gathered triples of all ranks, this band's shards re-laid out
        parts = torch.empty((3, out_rows, W), dtype=torch.float32, device="cuda")
        allparts = torch.empty((world, 3, out_rows, W), dtype=torch.float32, device="cuda")
        gather_band = torch.empty((world, out_rows, W), dtype=torch.uint8, device="cuda")
    units_per_step_rank = W * out_rows * ((dshard[1] - dshard[0]) if dshard else D) * len(pairs)   # pix*disp this rank produces per step
    torch.cuda.synchronize()

    stage_log = []   # per-stage CUDA-event times of every hot-path call made inside the timed region

    def shard_step(timing):
        """cfg4 on a 2-D grid: aggregate this rank's (band, disparity shard), all-gather the partial WTA triples (the one
        real exchange step of this sharding: 12 B per pixel and rank), merge this band's shards, all-gather the bands."""
        l, rr = dev_in[0]
        tmr = ctx.disparity_shard_raw(l.data_ptr(), rr.data_ptr(), W, H, params, band, dshard, parts[0].data_ptr(), parts[1].data_ptr(),
                                      parts[2].data_ptr(), timing=timing)
        with torch.cuda.stream(stream):
            dist.all_gather_into_tensor(allparts, parts)
            b0 = (rank // nd) * nd
            mine = allparts[b0:b0 + nd].permute(1, 0, 2, 3).contiguous()           # [3][shard][rows][W]
            ctx.merge_shards(W, out_rows, D, nd, mine[0].data_ptr(), mine[1].data_ptr(), mine[2].data_ptr(), None, dev_d[0].data_ptr(), None)
            dist.all_gather_into_tensor(gather_band, dev_d[0])                      # rank b*nd holds band b
        return tmr

    def step_device():
        if dshard is not None:
            stage_log.append(shard_step(True))
            return
        for (l, rr), o in zip(dev_in, dev_d):
            # timing=True: the library brackets every kernel group with events on its own stream (and waits for
            # them at the end of the call, a ~20 us host gap per call that stays inside the timed region)
            stage_log.append(ctx.disparity_raw(l.data_ptr(), rr.data_ptr(), W, H, params, None, o.data_ptr(), None, timing=True, band=band))
        if world > 1:   # the single collective: all-gather of the uint8 disparity maps / bands
            with torch.cuda.stream(stream):
                if band_mode and H % world == 0:
                    dist.all_gather_into_tensor(gather_band, dev_d[0])
                elif not band_mode:
                    dist.all_gather_into_tensor(gather, torch.stack(dev_d) if len(dev_d) > 1 else dev_d[0].unsqueeze(0))

    def step_host():
        if dshard is not None:
            with torch.cuda.stream(stream):
                dev_in[0][0].copy_(host_in[0][0], non_blocking=True)
                dev_in[0][1].copy_(host_in[0][1], non_blocking=True)
            shard_step(False)
            with torch.cuda.stream(stream):
                host_d[0].copy_(dev_d[0], non_blocking=True)
            stream.synchronize()
            return
        for (l, rr), o, of in zip(host_in, host_d, host_full_d):
            if band_mode:   # upload, run the band on device pointers, download the band
                with torch.cuda.stream(stream):
                    dev_in[0][0].copy_(l, non_blocking=True)
                    dev_in[0][1].copy_(rr, non_blocking=True)
                    ctx.disparity_raw(dev_in[0][0].data_ptr(), dev_in[0][1].data_ptr(), W, H, params, None, dev_d[0].data_ptr(), None, band=band)
                    o.copy_(dev_d[0], non_blocking=True)
                stream.synchronize()
            else:           # the user-facing call: host buffers in, host buffers out, synchronous
                ctx.disparity_raw(l.data_ptr(), rr.data_ptr(), W, H, params, None, of.data_ptr(), None, host=True)

    def timed(fn, steps, warm):
        for _ in range(warm):
            fn()
        stream.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        stream.synchronize()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        ms = e0.elapsed_time(e1)
        if fn is step_host and not band_mode:
            ms = wall       # host entry point is synchronous per call: the wall clock is the honest number
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.barrier()
        return float(t.item())

    # ---- timed regions ---------------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev = timed(step_device, args.steps, max(args.warmup, 3))
    clocks = sampler.stop() if rank == 0 else None
    timed_calls = stage_log[-args.steps * len(pairs):]           # the calls of the timed region (warm-up calls dropped)
    ms_host = timed(step_host, args.steps, 1)

    total_units = units_per_step_rank
    if world > 1:
        tu = torch.tensor([float(units_per_step_rank)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tu)
        total_units = float(tu.item())
    value = total_units * args.steps / (ms_dev * 1e-3) / 1e6
    e2e_value = total_units * args.steps / (ms_host * 1e-3) / 1e6

    # ---- per-stage timing of one instrumented step (outside the timed region) ----------------------
    tm = {k: float(np.mean([c[k] for c in timed_calls])) for k in timed_calls[0]}
    launches_per_pair = int(timed_calls[0]["kernel_launches"])

    if rank == 0:
        # roofline of the dominant kernel: F_alg of one pass = 4*T*W*rows*D, over its mean duration
        rows_mean = out_rows
        if band_mode:   # halo rows computed per pass, averaged over the r iterations
            rows_mean = float(np.mean([min(H, band[1] + (r - 1 - it) * 16) - max(0, band[0] - (r - 1 - it) * 16) for it in range(r)]))
        D_rank = (dshard[1] - dshard[0]) if dshard else D             # disparities this rank aggregates
        pass_flops = 4.0 * T_TAPS * W * rows_mean * D_rank
        # V pass = main kernel + two small launches (diagonal fix-up, edge padding): the dominant KERNEL is the main one
        v_ms, h_ms = tm["vagg_mean_ms"] - tm["vfix_mean_ms"], tm["hagg_mean_ms"]
        tma = args.family == 0                                   # radius 16: the TMA-fed kernels (D padded to 128 internally)
        v_name = "k_vagg_v2 (asw_vCostAggregation; mean of its %d launches per frame)" % r if tma else "k_vagg_t (asw_vCostAggregation)"
        h_name = "k_hagg_split (asw_hCostAggregation)" if tma else "k_hagg_t (asw_hCostAggregation)"
        dom, dom_ms = (v_name, v_ms) if v_ms >= h_ms else (h_name, h_ms)
        if v_ms >= h_ms and tma:
            pass_flops *= (D - 1.5) / D                         # the fix-up launch produces 1.5 of the D outputs per pixel
        peak, peak_src = FP32_PEAK_FALLBACK_TFLOPS, "computed: 148 SMs x 128 lanes x 2 x 1.965 GHz (MEASURED_PEAKS.json has no FP32 figure)"
        ffma = None
        try:
            ub = C.CDLL(os.path.join(ROOT, "stereo_matchin_b200", "libasw_ubench.so"))
            ub.asw_ubench_ffma_tflops.restype = C.c_double
            ffma = {"ffma_tflops": ub.asw_ubench_ffma_tflops(0), "ffma2_tflops": ub.asw_ubench_ffma_tflops(1)}
        except Exception as e:  # measurement helper only
            ffma = {"error": str(e)}
        achieved = pass_flops / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
        peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm_peak = json.load(open(peaks_file)).get("hbm_gbs") if os.path.exists(peaks_file) else 6650.0
        Dp = (D_rank + 63) // 64 * 64 if args.family == 0 else (D + 31) // 32 * 32
        pass_bytes = 3.0 * 4 * W * rows_mean * Dp      # tiled kernels: read cost + read denominator + write cost
        traffic, traffic_src = None, None                       # DRAM bytes per launch of that kernel, from the committed ncu --set full capture
        tfiles = sorted(f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.endswith("_traffic.json")) if os.path.isdir(os.path.join(ROOT, "profiles")) else []
        if tfiles and args.workload == "cfg3" and world == 1:
            tj = json.load(open(os.path.join(ROOT, "profiles", tfiles[-1])))
            key = "k_vagg_v2" if v_ms >= h_ms else "k_hagg_split"
            if key in tj:
                traffic, traffic_src = tj[key]["dram_bytes_per_launch"], "profiles/" + tfiles[-1]
        roofline = {"bound": "fp32", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes": pass_bytes,
                    "timing": "CUDA events on the library's stream around every launch group, mean over the timed region",
                    "peak_source": peak_src, "measured_ffma_microbench": ffma,
                    "whole_path_frac": alg_flops(W, rows_mean, D_rank, r) / (tm["total_ms"] * 1e-3) / 1e12 / peak,
                    "v_pass_ms": v_ms, "h_pass_ms": h_ms,
                    "hbm": {"algorithmic_gbs": pass_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0, "peak_gbs": hbm_peak,
                            "bytes_per_pass": pass_bytes}}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu, _ = oracle_sample(pairs[0][0], pairs[0][1], D, r, args.cpu_rows)
            cpu.pop("seconds", None)
        also = None
        if world == 1 and args.workload == "cfg3" and not args.no_cpu_baseline:
            # BASELINE.json configs[1] (teddy / cones shape, 61 disparities) beside the headline: sub-millisecond kernels,
            # launch- and tail-dominated, so it is reported, not used for the roofline
            W2, H2, D2, _ = synth.CONFIGS["cfg2_teddy_shape"]
            L2, R2 = synth.make_config("cfg2_teddy_shape", 0)[:2]
            p2 = api.AswParams(ndisp=D2, iterations=r)
            a2, b2 = torch.from_numpy(L2).cuda(), torch.from_numpy(R2).cuda()
            o2 = torch.empty((H2, W2), dtype=torch.uint8, device="cuda")
            ho2 = np.empty((H2, W2), np.uint8)
            for _ in range(3):
                ctx.disparity_raw(a2.data_ptr(), b2.data_ptr(), W2, H2, p2, None, o2.data_ptr(), None)
            ctx.sync()
            t0 = time.perf_counter()
            for _ in range(20):
                ctx.disparity_raw(a2.data_ptr(), b2.data_ptr(), W2, H2, p2, None, o2.data_ptr(), None)
            ctx.sync()
            dev_ms = (time.perf_counter() - t0) * 1e3 / 20
            t0 = time.perf_counter()
            for _ in range(10):
                ctx.disparity_raw(L2.ctypes.data, R2.ctypes.data, W2, H2, p2, None, ho2.ctypes.data, None, host=True)
            host_ms = (time.perf_counter() - t0) * 1e3 / 10
            also = {"cfg2": {"workload": "synthetic 450x375 pair, 61 disparities (BASELINE.json configs[1] shape), r=%d" % r,
                             "ms_per_frame_device": dev_ms, "Mpix_disp_per_s": W2 * H2 * D2 / dev_ms / 1e3,
                             "ms_per_frame_e2e_host_buffers": host_ms, "e2e_Mpix_disp_per_s": W2 * H2 * D2 / host_ms / 1e3,
                             "timing": "wall clock around 20 back-to-back calls + stream sync (frames this small are launch-bound)"}}
            del a2, b2, o2
        npx_in = sum(a.numel() + b.numel() for a, b in host_in)
        line = {
            "metric": "Mpix*disp/s (ASW agg+WTA, device-timed)", "value": value, "unit": "Mpix*disp/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "strong" if band_mode else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}", "W": W, "H": H, "ndisp": D, "radius": 16, "iterations": r,
                       "pairs_per_rank": len(pairs), "sharding": (("%d row bands x %d disparity shards, all-gather of partial WTA triples + bands" % grid) if dshard is not None
                                                                  else "row bands + shrinking halo, all-gather of bands" if band_mode
                                                                  else "pairs (independent), all-gather of disparity maps"),
                       "l2": "inputs larger than L2: every pass streams a %.2f GB cost volume" % (4.0 * W * H * Dp / 1e9),
                       "kernel_family": {0: "tma", 1: "basic", 2: "tiled"}[args.family]},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mpix*disp/s", "h2d_bytes_per_step": int(npx_in), "d2h_bytes_per_step": int(W * out_rows * len(pairs)),
                    "ms_per_step": ms_host / args.steps},
            "gpu_launches": int(launches_per_pair * len(pairs) * args.steps),
            "stage_ms": {k: tm[k] for k in ("raw_ms", "supp_ms", "vagg_mean_ms", "vfix_mean_ms", "hagg_mean_ms", "agg_total_ms", "wta_ms", "total_ms")},
            "roofline": roofline,
            "cpu_baseline": cpu,
            "also": also,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    # tensors allocated under the library's stream must be released before that stream is destroyed
    del dev_in, dev_d, gather, gather_band, host_in, host_d, host_full_d
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    ctx.close()


if __name__ == "__main__":
    main()
