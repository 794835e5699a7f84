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
one all-gather of the uint8 disparity maps at the end of a step).  The strong-scaling configuration
of BASELINE.json (one 3840x2160x256 frame split into row bands, `radius` halo rows exchanged between
iterations over NCCL) is measured in every run as `also.cfg4_strong`, with an in-run check that the
gathered map equals the one-GPU map; `--workload cfg4` makes it the headline instead.
Both arms (`--impl ours` / `--impl reference`) print the same `metric`, `unit` and `config`.
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
METRIC = "Mpix*disp/s (ASW agg+WTA)"          # the same string in both arms: the driver pairs the arms by it
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


def make_config_dict(args, W, H, D, pairs_per_rank, sharding_desc, Dp):
    """`config` of the JSON line: identical in both arms."""
    _, _, desc = WORKLOADS[args.workload]
    return {"workload": f"{args.workload}: {desc}", "W": W, "H": H, "ndisp": D, "radius": 16, "iterations": args.iterations,
            "pairs_per_rank": pairs_per_rank, "sharding": sharding_desc,
            "l2": "inputs larger than L2: every pass streams a %.2f GB cost volume" % (4.0 * W * H * Dp / 1e9),
            "kernel_family": {0: "tma", 1: "basic"}[args.family],
            "reference_sample": "the reference arm (--impl reference) and cpu_baseline time the CPU port on the top %d of %d rows "
                                "of the same pair, all host threads" % (min(H, max(33, int(args.ref_rows))), H)}


def default_sharding_desc(args, world):
    if args.workload == "cfg4":
        return "row bands, %d halo rows exchanged per iteration (NCCL send/recv), all-gather of bands" % 16
    return "pairs (independent), all-gather of disparity maps"


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
    warm = max(args.warmup, 3)                                   # the same warm-up count as our arm
    times = []
    info = None
    for i in range(warm + args.steps):
        info, _ = oracle_sample(L, R, D, args.iterations, rows)
        if i >= warm:
            times.append(info["seconds"])
    ms = float(np.mean(times)) * 1e3
    v = W * rows * D / (ms * 1e-3) / 1e6
    Dp = (D + 63) // 64 * 64
    ppr = WORKLOADS[args.workload][1]
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "Mpix*disp/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong" if args.workload == "cfg4" else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "timing": "host wall clock around the CPU port",
            "config": make_config_dict(args, W, H, D, ppr, default_sharding_desc(args, args.gpus), Dp),
            "cpu_baseline": {"value": v, "unit": "Mpix*disp/s", "cores": info["cores"], "kind": "port", "sample": info["sample"]},
            "e2e": {"value": v, "unit": "Mpix*disp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def parity_block(ctx, api, r):
    """north_star: aggregated costs within 1e-5 relative of the reference, disparity flips counted and reported.  The CUDA
    path (FMA-contracted tap accumulation, bit-identical to oracle(use_fma=1)) against the OTHER arithmetic an OpenCL build
    of the reference may legally use (separately rounded multiply and add, oracle(use_fma=0)), on BASELINE.json configs[1]'s
    shape (450x375, 61 disparities, r iterations)."""
    from oracle import asw_oracle as O
    from stereo_matchin_b200 import synth
    L, R, _, D = synth.make_config("cfg2_teddy_shape")
    H, W, _ = L.shape
    p = api.AswParams(ndisp=D, iterations=r)
    ctx.set_keep_volume(True)
    dl, dr = ctx.to_device(L), ctx.to_device(R)
    od = ctx.alloc(W * H)
    ctx.disparity_raw(dl.ptr, dr.ptr, W, H, p, None, od.ptr, None)
    ctx.sync()
    cost = np.empty((D, H, W), np.float32)
    ctx._check(ctx.lib.asw_memcpy_d2h(ctx.h, cost.ctypes.data, ctx.final_volume_ptr(), cost.nbytes))
    d_gpu = od.download((H, W), np.uint8)
    ctx.set_keep_volume(False)
    for b in (dl, dr, od):
        b.free()
    ref = O.asw_hot_path(L, R, O.OracleParams(ndisp=D, iterations=r), use_fma=False, want_cost=True)
    fma = O.asw_hot_path(L, R, O.OracleParams(ndisp=D, iterations=r), use_fma=True, want_cost=True)
    rel = np.abs(cost - ref["cost"]) / np.maximum(np.abs(ref["cost"]), 1e-30)
    d_ref = ref["d_ref"].astype(np.int32)
    flips = d_gpu.astype(np.int32) != d_ref
    # a flip is a tie flip when the reference's own two best costs are within the tolerance of each other
    srt = np.sort(ref["cost"], axis=0)
    near_tie = (srt[1] - srt[0]) <= 1e-5 * np.abs(srt[1])
    return {"workload": "cfg2 shape: synthetic 450x375, 61 disparities, r=%d" % r,
            "against": "CPU oracle with separately rounded multiply/add (use_fma=0), the other arithmetic the reference's OpenCL build may use",
            "cost_rel_err_max": float(rel.max()), "cost_rel_err_tolerance": 1e-5,
            "tie_flips": int((flips & near_tie).sum()), "other_flips": int((flips & ~near_tie).sum()),
            "tie_flip_pct": float(100.0 * flips.sum() / flips.size), "tie_flip_pct_limit": 0.1,
            "bit_identical_to_fma_oracle": bool(np.array_equal(cost.view(np.uint32), fma["cost"].view(np.uint32))
                                                and np.array_equal(d_gpu, fma["d_ref"].astype(np.uint8)))}


def cfg4_strong_block(ctx, api, torch, dist, rank, world, r, steps=3):
    """BASELINE.json configs[3]: ONE 3840x2160x256 frame on all ranks, row bands, `radius` halo rows exchanged over NCCL
    between iterations, one all-gather of the uint8 bands.  Device-timed (CUDA events on the library's stream, max over
    ranks); the gathered map is compared byte for byte with rank 0's own unsharded result in the same run."""
    from stereo_matchin_b200 import sharding, synth
    L, R, _, D = synth.make_config("cfg4_3840x2160_d256")
    H, W, _ = L.shape
    p = api.AswParams(ndisp=D, iterations=r)
    dl, dr = torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()
    y0, y1 = sharding.row_bands(H, world)[rank]
    band = torch.empty((y1 - y0, W), dtype=torch.uint8, device="cuda")
    stream = torch.cuda.ExternalStream(ctx.stream())
    full = None

    def step():
        nonlocal full
        sharding.disparity_row_exchange_cuda(ctx, dl.data_ptr(), dr.data_ptr(), W, H, p, rank, world, band)
        if world > 1:
            ctx.sync()
            full = sharding.gather_bands(band, H, W, rank, world)

    step()                                                    # warm-up: allocations, NCCL channels
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    times = []
    for _ in range(steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step()
        e1.record(stream)
        stream.synchronize()
        torch.cuda.synchronize()
        # the exchange callbacks synchronise with the host, so the frame time is the wall clock of the slowest rank;
        # the event span on the library's stream is reported beside it
        t = torch.tensor([(time.perf_counter() - t0) * 1e3, e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(t.tolist())
    ms_wall = float(np.median([t[0] for t in times]))
    ms_dev = float(np.median([t[1] for t in times]))
    equals = True
    if world > 1 and rank == 0:
        one = torch.empty((H, W), dtype=torch.uint8, device="cuda")
        ctx.disparity_raw(dl.data_ptr(), dr.data_ptr(), W, H, p, None, one.data_ptr(), None)
        ctx.sync()
        equals = bool(torch.equal(full, one))
    one_process = None
    if world > 1:
        # the same frame through the C ABI's own multi-GPU entry (asw_multi_*): ONE process (rank 0) drives all `world` GPUs
        # with a thread per band, copy-engine peer copies ordered by events, exchange hidden under the interior rows.  The
        # other ranks wait at a HOST-side barrier (gloo): an NCCL barrier would park a spinning kernel on every GPU rank 0 is
        # about to use (measured: 37.0 instead of 17.0 ms per frame at 8 GPUs).
        torch.cuda.synchronize()
        host_group = dist.new_group(backend="gloo")
        dist.barrier(group=host_group)
        if rank == 0:
            d_one = one.cpu().numpy()
            with api.AswMulti(list(range(world))) as m:
                m.disparity(L, R, p, want_conf=False)
                runs = [m.disparity(L, R, p, want_conf=False) for _ in range(steps)]
            ms1 = float(np.median([x["timing"]["compute_ms"] for x in runs]))
            one_process = {"how": "asw_multi_disparity from one process: %d band threads, cudaMemcpyPeerAsync of the halo rows on a communication "
                                  "stream per band, ordered by CUDA events, under the interior rows of the iteration" % world,
                           "ms_per_frame": ms1, "Mpix_disp_per_s": W * H * D / ms1 / 1e3,
                           "upload_ms": float(np.median([x["timing"]["upload_ms"] for x in runs])),
                           "download_ms": float(np.median([x["timing"]["download_ms"] for x in runs])),
                           "equals_1gpu": bool(all(np.array_equal(x["disp_d"], d_one) for x in runs)),
                           "timing": "host wall clock from the first launch to the slowest band's last kernel, median of %d frames" % steps}
            del one
        dist.barrier(group=host_group)
        dist.destroy_process_group(host_group)
    del dl, dr, band, full
    torch.cuda.synchronize()
    return {"workload": "cfg4: ONE synthetic 3840x2160 pair, 256 disparities, r=%d, strong scaling" % r, "n_gpus": world,
            "one_process": one_process,
            "sharding": "%d row bands; %d halo rows exchanged with each neighbour per iteration (NCCL send/recv, %.0f MB each way); "
                        "no rows recomputed; one all-gather of uint8 bands" % (world, 16, 16 * (((W + 63) // 64) * 64 + 32) * 256 * 4 / 1e6),
            "ms_per_frame": ms_wall, "ms_per_frame_device_events": ms_dev, "Mpix_disp_per_s": W * H * D / ms_wall / 1e3,
            "halo_rows_recomputed_fraction": 0.0, "equals_1gpu": equals,
            "equals_1gpu_how": "rank 0 runs the unsharded frame after the timed steps and compares all %d bytes" % (W * H) if world > 1
                               else "N=1 is the unsharded frame",
            "timing": "median of %d frames; wall clock from a barrier to the last rank's completed all-gather, max over ranks" % steps}


def cfg5_batch_block(api, torch, dist, rank, local_rank, world, r, pairs_per_rank=8, rounds=2):
    """BASELINE.json configs[4]: a batch of 1280x720x128 pairs, pair sharded.  Host buffers (pinned) in and out through
    asw_disparity_async on two contexts per GPU that alternate between pairs, so the upload / download of one pair runs
    under the kernels of the other.  A bounded sample of the 1024-pair batch; the full batch time is extrapolated."""
    from stereo_matchin_b200 import synth
    W, H, D, _ = synth.CONFIGS["cfg5_1280x720_d128"]
    p = api.AswParams(ndisp=D, iterations=r)
    pairs = [synth.make_config("cfg5_1280x720_d128", rank * pairs_per_rank + i)[:2] for i in range(pairs_per_rank)]
    hin = [(torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory()) for a, b in pairs]
    hout = [torch.empty((H, W), dtype=torch.uint8).pin_memory() for _ in pairs]
    ctxs = [api.AswContext(local_rank), api.AswContext(local_rank)]

    def run_batch():
        for i, ((a, b), o) in enumerate(zip(hin, hout)):
            ctxs[i & 1].disparity_async(a.data_ptr(), b.data_ptr(), W, H, p, None, o.data_ptr(), None)
        for c in ctxs:
            c.sync()

    run_batch()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(rounds):
        run_batch()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    for c in ctxs:
        c.close()
    npairs = pairs_per_rank * rounds * world
    return {"workload": "cfg5: batch of synthetic 1280x720 pairs, 128 disparities, r=%d, pair sharded" % r, "n_gpus": world,
            "pairs_timed": npairs, "pairs_per_s": npairs / dt, "Mpix_disp_per_s": npairs * W * H * D / dt / 1e6,
            "seconds_for_1024_pairs_extrapolated": 1024.0 / (npairs / dt),
            "h2d_bytes_per_pair": 2 * W * H * 4, "d2h_bytes_per_pair": W * H,
            "timing": "host wall clock over %d pairs per rank (%d distinct x %d rounds), pinned host buffers in and out, "
                      "two asw contexts per GPU alternating (asw_disparity_async), max over ranks" % (pairs_per_rank * rounds, pairs_per_rank, rounds)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--iterations", type=int, default=7)
    ap.add_argument("--family", type=int, default=0, choices=[0, 1], help="0 = TMA-fed kernels (default), 1 = generic kernels")
    ap.add_argument("--cpu-rows", type=int, default=400, help="rows of the pair the CPU baseline sample covers")
    ap.add_argument("--ref-rows", type=int, default=400)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the also.* blocks (cfg2, parity, cfg4_strong, cfg5_batch)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    from stereo_matchin_b200 import api, sharding, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback exists)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    cfg, pairs_per_rank, desc = WORKLOADS[args.workload]
    W, H, D, _ = synth.CONFIGS[cfg]
    r = args.iterations
    band_mode = args.workload == "cfg4"                          # one frame on all ranks: row bands + halo exchange
    params = api.AswParams(ndisp=D, iterations=r)
    ctx = api.AswContext(local_rank)
    ctx.set_kernel_family(args.family)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local_rank))

    # ---- inputs: pinned host copies and device-resident copies --------------------------------
    if band_mode:
        pairs = [synth.make_config(cfg, 0)[:2]]                      # every rank holds the full frame
        band = sharding.row_bands(H, world)[rank]
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
    units_per_step_rank = W * out_rows * D * len(pairs)           # pix*disp this rank produces per step
    torch.cuda.synchronize()

    stage_log = []   # per-stage CUDA-event times of every hot-path call made inside the timed region

    def step_device():
        if band_mode:
            stage_log.append(sharding.disparity_row_exchange_cuda(ctx, dev_in[0][0].data_ptr(), dev_in[0][1].data_ptr(), W, H, params, rank,
                                                                  world, dev_d[0], timing=True))
            if world > 1:   # uneven bands are padded inside gather_bands
                sharding.gather_bands(dev_d[0], H, W, rank, world)
            return
        for (l, rr), o in zip(dev_in, dev_d):
            # timing=True: the library brackets every kernel group with events on its own stream (and waits for
            # them at the end of the call, a ~20 us host gap per call that stays inside the timed region)
            stage_log.append(ctx.disparity_raw(l.data_ptr(), rr.data_ptr(), W, H, params, None, o.data_ptr(), None, timing=True))
        if world > 1:   # the single collective: all-gather of the uint8 disparity maps
            with torch.cuda.stream(stream):
                dist.all_gather_into_tensor(gather, torch.stack(dev_d) if len(dev_d) > 1 else dev_d[0].unsqueeze(0))

    def step_host():
        for (l, rr), o, of in zip(host_in, host_d, host_full_d):
            if band_mode:   # upload, run the band on device pointers, download the band
                with torch.cuda.stream(stream):
                    dev_in[0][0].copy_(l, non_blocking=True)
                    dev_in[0][1].copy_(rr, non_blocking=True)
                stream.synchronize()
                sharding.disparity_row_exchange_cuda(ctx, dev_in[0][0].data_ptr(), dev_in[0][1].data_ptr(), W, H, params, rank, world, dev_d[0])
                with torch.cuda.stream(stream):
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
        if fn is step_host or band_mode:
            ms = wall       # synchronous host entry point / host-synchronised halo exchange: the wall clock is the honest number
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.barrier()
        return float(t.item())

    # ---- timed regions ---------------------------------------------------------------------------
    warm = max(args.warmup, 3)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev = timed(step_device, args.steps, warm)
    clocks = sampler.stop() if rank == 0 else None
    timed_calls = stage_log[-args.steps * len(pairs):]           # the calls of the timed region (warm-up calls dropped)
    ms_host = timed(step_host, args.steps, 1)

    total_units = float(W * H * D) if band_mode else float(units_per_step_rank)
    if world > 1 and not band_mode:
        tu = torch.tensor([float(units_per_step_rank)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tu)
        total_units = float(tu.item())
    value = total_units * args.steps / (ms_dev * 1e-3) / 1e6
    e2e_value = total_units * args.steps / (ms_host * 1e-3) / 1e6

    tm = {k: float(np.mean([c[k] for c in timed_calls])) for k in timed_calls[0]}
    launches_per_pair = int(timed_calls[0]["kernel_launches"])

    # ---- the other configurations of BASELINE.json, measured beside the headline (all ranks take part) ----------
    also = {}
    if not args.no_also and args.workload == "cfg3":
        del dev_in, dev_d, gather                               # make room: the 4K frame needs 38 GB of scratch on one GPU
        dev_in = dev_d = gather = None
        torch.cuda.empty_cache()
        also["cfg4_strong"] = cfg4_strong_block(ctx, api, torch, dist, rank, world, r)
        also["cfg5_batch"] = cfg5_batch_block(api, torch, dist, rank, local_rank, world, r)

    if rank == 0:
        # roofline of the dominant kernel: F_alg of one pass = 4*T*W*rows*D, over its mean duration
        rows_mean = out_rows
        pass_flops = 4.0 * T_TAPS * W * rows_mean * D
        # V pass = main kernel + the edge padding launch: the dominant KERNEL is the main one
        v_ms, h_ms = tm["vagg_mean_ms"] - tm["vfix_mean_ms"], tm["hagg_mean_ms"]
        tma = args.family == 0                                   # radius 16: the TMA-fed kernels (D padded to 64 internally)
        v_name = "k_vagg_v2 (asw_vCostAggregation; mean of its %d launches per frame)" % r if tma else "k_vagg_t (asw_vCostAggregation)"
        h_name = "k_hagg_split (asw_hCostAggregation)" if tma else "k_hagg_t (asw_hCostAggregation)"
        dom, dom_ms = (v_name, v_ms) if v_ms >= h_ms else (h_name, h_ms)
        peak = FP32_PEAK_FALLBACK_TFLOPS
        peak_src = "computed: 148 SMs x 128 lanes x 2 x 1.965 GHz (MEASURED_PEAKS.json holds HBM GB/s and bf16 TFLOP/s, no FP32 figure)"
        ffma = None
        try:
            ub = C.CDLL(os.path.join(ROOT, "stereo_matchin_b200", "libasw_ubench.so"))
            ub.asw_ubench_ffma_tflops.restype = C.c_double
            ffma = {"ffma_tflops": ub.asw_ubench_ffma_tflops(0), "ffma2_tflops": ub.asw_ubench_ffma_tflops(1)}
        except Exception as e:  # measurement helper only
            ffma = {"error": str(e)}
        peak_measured = max(ffma.get("ffma_tflops", 0.0), ffma.get("ffma2_tflops", 0.0)) if "error" not in ffma else None
        achieved = pass_flops / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
        peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm_peak = json.load(open(peaks_file)).get("hbm_gbs") if os.path.exists(peaks_file) else 6650.0
        Dp = (D + 63) // 64 * 64 if args.family == 0 else (D + 31) // 32 * 32
        pass_bytes = 3.0 * 4 * W * rows_mean * Dp      # read cost + read denominator + write cost
        traffic, traffic_src = None, None              # DRAM bytes per launch of that kernel, from the committed ncu --set full capture
        pdir = os.path.join(ROOT, "profiles")
        tfiles = sorted(f for f in os.listdir(pdir) if f.endswith("_traffic.json")) if os.path.isdir(pdir) else []
        if tfiles and args.workload == "cfg3":         # every rank runs the same cfg3 launches: the per-launch traffic holds at any N
            tj = json.load(open(os.path.join(pdir, tfiles[-1])))
            key = "k_vagg_v2" if v_ms >= h_ms else "k_hagg_split"
            if key in tj:
                traffic, traffic_src = tj[key]["dram_bytes_per_launch"], "profiles/" + tfiles[-1]
        whole = alg_flops(W, rows_mean, D, r) / (tm["total_ms"] * 1e-3) / 1e12
        roofline = {"bound": "fp32", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes": pass_bytes,
                    "timing": "CUDA events on the library's stream around every launch group, mean over the timed region",
                    "peak_source": peak_src, "peak_measured": peak_measured,
                    "peak_measured_source": "FP32 FMA microbenchmark run by this process on this GPU (libasw_ubench.so: dependent FFMA / FFMA2 chains, ILP 8, best of 5)",
                    "frac_of_peak_measured": achieved / peak_measured if peak_measured else None,
                    "measured_ffma_microbench": ffma,
                    "whole_path_frac": whole / peak, "whole_path_frac_of_peak_measured": whole / peak_measured if peak_measured else None,
                    "v_pass_ms": v_ms, "h_pass_ms": h_ms,
                    "other_pass": {"kernel": h_name if v_ms >= h_ms else v_name,
                                   "frac": pass_flops / (min(v_ms, h_ms) * 1e-3) / 1e12 / peak if min(v_ms, h_ms) > 0 else None},
                    "hbm": {"algorithmic_gbs": pass_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0, "peak_gbs": hbm_peak,
                            "bytes_per_pass": pass_bytes}}
        # the boxes of this pool run this load at their power cap (sw_power_cap): the FMA rate available at the SM clock
        # sampled DURING the timed region, beside the fraction against the peak at the maximum clock
        if clocks and clocks.get("sm_mhz"):
            peak_clk = 148 * 128 * 2 * clocks["sm_mhz"] * 1e6 / 1e12
            roofline["frac_at_sampled_sm_clock"] = {"sm_mhz": clocks["sm_mhz"], "peak_tflops_at_that_clock": peak_clk, "frac": achieved / peak_clk,
                                                    "whole_path_frac": whole / peak_clk}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu, _ = oracle_sample(pairs[0][0], pairs[0][1], D, r, args.cpu_rows)
            cpu.pop("seconds", None)
        if world == 1 and args.workload == "cfg3" and not args.no_also:
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
            hl2, hr2 = torch.from_numpy(L2).pin_memory(), torch.from_numpy(R2).pin_memory()   # pinned, like the headline's e2e
            hq2 = torch.empty((H2, W2), dtype=torch.uint8).pin_memory()
            for _ in range(3):
                ctx.disparity_raw(hl2.data_ptr(), hr2.data_ptr(), W2, H2, p2, None, hq2.data_ptr(), None, host=True)
            t0 = time.perf_counter()
            for _ in range(10):
                ctx.disparity_raw(hl2.data_ptr(), hr2.data_ptr(), W2, H2, p2, None, hq2.data_ptr(), None, host=True)
            host_ms = (time.perf_counter() - t0) * 1e3 / 10
            ho2 = hq2.numpy()
            also["cfg2"] = {"workload": "synthetic 450x375 pair, 61 disparities (BASELINE.json configs[1] shape), r=%d" % r,
                            "ms_per_frame_device": dev_ms, "Mpix_disp_per_s": W2 * H2 * D2 / dev_ms / 1e3,
                            "ms_per_frame_e2e_host_buffers": host_ms, "e2e_Mpix_disp_per_s": W2 * H2 * D2 / host_ms / 1e3,
                            "timing": "wall clock around 20 back-to-back calls + stream sync; repeated calls replay a CUDA graph of the ~30 launches; e2e: pinned host buffers in and out, one call at a time"}
            del a2, b2, o2
            also["parity"] = parity_block(ctx, api, r)
        npx_in = sum(a.numel() + b.numel() for a, b in host_in)
        sharding_desc = default_sharding_desc(args, world)
        line = {
            "metric": METRIC, "value": value, "unit": "Mpix*disp/s", "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "strong" if band_mode else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "timing": "device: CUDA events on the library's stream, max over ranks" if not band_mode
                      else "wall clock of the slowest rank (the halo exchange synchronises with the host)",
            "config": make_config_dict(args, W, H, D, len(pairs), sharding_desc, Dp),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mpix*disp/s", "h2d_bytes_per_step": int(npx_in), "d2h_bytes_per_step": int(W * out_rows * len(pairs)),
                    "ms_per_step": ms_host / args.steps},
            "gpu_launches": int(launches_per_pair * len(pairs) * args.steps),
            "stage_ms": {k: tm[k] for k in ("raw_ms", "supp_ms", "vagg_mean_ms", "vfix_mean_ms", "hagg_mean_ms", "agg_total_ms", "wta_ms", "total_ms")},
            "roofline": roofline,
            "cpu_baseline": cpu,
            "also": also or None,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    # tensors allocated under the library's stream must be released before that stream is destroyed
    del dev_in, dev_d, gather, host_in, host_d, host_full_d
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    ctx.close()


if __name__ == "__main__":
    main()
