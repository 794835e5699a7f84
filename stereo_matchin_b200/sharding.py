"""Multi-GPU sharding of the ASW hot path: one process per GPU, torch.distributed for the plumbing.

The path shards without any mid-computation exchange:
  * pair sharding  -- a batch of independent stereo pairs is split across ranks;
  * row-band sharding -- one large frame is split into row bands.  A vertical pass reaches R rows
    per iteration, so a band needs r*R real halo rows on each side (clipped at the frame border,
    where the reference's clamp-to-edge applies); the library shrinks the halo by R per iteration
    (asw_disparity_band_device), so the band is bit-identical to the same rows of a 1-GPU run.
  * disparity sharding -- planes of the cost volume never interact during aggregation, so a rank can own
    the disparities [d0, d1) of (a band of) a frame with NO halo work; it returns per pixel the partial
    winner-take-all triple (min1, min2, argmin) and the triples are merged after one all-gather
    (asw_disparity_shard_device / asw_merge_shards).  A 2-D grid (row bands x disparity shards) keeps the
    halo overhead of an 8-GPU run of a 4K frame at +4 % instead of +31 %.
The only collectives are all-gathers of results: uint8 disparity bands / maps, or the 12-byte partial
triples of disparity shards (NCCL on GPUs; the CPU tests run the same code over gloo).  The compute callback is injected so that the host logic can be
tested without a GPU; the product passes the CUDA library's band entry point.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import numpy as np


def row_bands(H: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced row bands [y0, y1) covering 0..H; ranks beyond H get empty bands."""
    if H <= 0 or world <= 0:
        raise ValueError("H and world must be positive")
    return [((H * r) // world, (H * (r + 1)) // world) for r in range(world)]


def band_input_rows(y0: int, y1: int, H: int, radius: int, iterations: int) -> Tuple[int, int]:
    """Rows of the frame that can influence output rows [y0, y1): r*R halo rows, clipped to the frame."""
    halo = radius * iterations
    return max(0, y0 - halo), min(H, y1 + halo)


def iteration_rows(y0: int, y1: int, H: int, radius: int, iterations: int) -> List[Tuple[int, int]]:
    """Rows each aggregation iteration has to produce (the halo shrinks by R per iteration)."""
    ya, yb = band_input_rows(y0, y1, H, radius, iterations)
    return [(max(ya, y0 - (iterations - 1 - it) * radius), min(yb, y1 + (iterations - 1 - it) * radius))
            for it in range(iterations)]


def band_work_fraction(H: int, world: int, radius: int = 16, iterations: int = 7) -> float:
    """Rows processed by all ranks (sum over iterations) relative to a 1-GPU run: the halo overhead."""
    total = 0
    for y0, y1 in row_bands(H, world):
        if y1 > y0:
            total += sum(b - a for a, b in iteration_rows(y0, y1, H, radius, iterations))
    return total / float(H * iterations)


def pair_shards(n_pairs: int, world: int) -> List[range]:
    """Contiguous, balanced split of a batch of independent stereo pairs."""
    return [range((n_pairs * r) // world, (n_pairs * (r + 1)) // world) for r in range(world)]


def gather_bands(local_band, H: int, W: int, rank: int, world: int, group=None):
    """All-gather of the uint8 disparity bands into the full H x W map (every rank gets it).

    `local_band` is a torch uint8 tensor of shape (rows_of_this_rank, W) on the device of the
    process group's backend.  Bands may differ by one row, so the gather is padded to the tallest.
    """
    import torch
    import torch.distributed as dist

    bands = row_bands(H, world)
    tall = max(b - a for a, b in bands)
    pad = torch.zeros((tall, W), dtype=torch.uint8, device=local_band.device)
    y0, y1 = bands[rank]
    pad[: y1 - y0] = local_band
    out = torch.empty((world * tall * W,), dtype=torch.uint8, device=local_band.device)
    if world > 1:
        dist.all_gather_into_tensor(out, pad.reshape(-1), group=group)     # flat: accepted by NCCL and gloo alike
    else:
        out.copy_(pad.reshape(-1))
    out = out.reshape(world, tall, W)
    full = torch.empty((H, W), dtype=torch.uint8, device=local_band.device)
    for r, (a, b) in enumerate(bands):
        full[a:b] = out[r, : b - a]
    return full


def gather_maps(local_maps, group=None):
    """All-gather of per-rank stacks of disparity maps (pair sharding, equal pairs per rank)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local_maps
    out = torch.empty((world * local_maps.numel(),), dtype=local_maps.dtype, device=local_maps.device)
    dist.all_gather_into_tensor(out, local_maps.contiguous().reshape(-1), group=group)
    return out.reshape((-1,) + tuple(local_maps.shape[1:]))


BandFn = Callable[[np.ndarray, np.ndarray, int, int], np.ndarray]


def disparity_row_sharded(left: np.ndarray, right: np.ndarray, rank: int, world: int, compute_band: BandFn, device="cpu",
                          group=None):
    """Row-band sharded disparity of one frame: this rank computes its band with `compute_band(left,
    right, y0, y1) -> uint8 (y1-y0, W)`, then all ranks all-gather the bands.  Returns the full map."""
    import torch

    H, W, _ = left.shape
    y0, y1 = row_bands(H, world)[rank]
    band = compute_band(left, right, y0, y1) if y1 > y0 else np.zeros((0, W), np.uint8)
    t = torch.from_numpy(np.ascontiguousarray(band)).to(device)
    return gather_bands(t, H, W, rank, world, group)


def disparity_pair_sharded(pairs: Sequence[Tuple[np.ndarray, np.ndarray]], rank: int, world: int,
                           compute_pair: Callable[[np.ndarray, np.ndarray], np.ndarray], device="cpu", group=None):
    """Pair-sharded batch: this rank computes its share of the pairs, then one all-gather of the maps.
    The batch size must be divisible by the world size (the benchmark shapes are)."""
    import torch

    if len(pairs) % world:
        raise ValueError("batch size must be divisible by the world size")
    mine = pair_shards(len(pairs), world)[rank]
    maps = np.stack([compute_pair(*pairs[i]) for i in mine])
    return gather_maps(torch.from_numpy(maps).to(device), group)


def cuda_band_fn(ctx, params) -> BandFn:
    """The product's compute callback: asw_disparity_band_device on this rank's GPU."""
    import torch

    def fn(left, right, y0, y1):
        H, W, _ = left.shape
        dl = torch.from_numpy(np.ascontiguousarray(left)).cuda()
        dr = torch.from_numpy(np.ascontiguousarray(right)).cuda()
        out = torch.empty((y1 - y0, W), dtype=torch.uint8, device="cuda")
        ctx.disparity_raw(dl.data_ptr(), dr.data_ptr(), W, H, params, None, out.data_ptr(), None, band=(y0, y1))
        ctx.sync()
        return out.cpu().numpy()

    return fn


# ---- 2-D sharding of one frame: row bands x disparity shards -------------------------------------------

def shard_grid(world: int, ndisp: int, window: int = 128) -> Tuple[int, int]:
    """(row bands, disparity shards) for `world` ranks: as many disparity shards as there are 128-disparity
    windows (no halo work), the rest of the ranks as row bands.  (Measured on the 4K x 256 frame: shards of 64
    disparities lose to 128 because the per-tile fixed costs of the vertical pass and the replicated weight
    tables weigh more: 4 ranks 2x2 39.1 ms vs 1x4 43.7 ms; 8 ranks 4x2 23.0, 2x4 23.2, 8x1 24.9 ms.)"""
    nwin = max(1, (ndisp + window - 1) // window)
    nd = 1
    for cand in range(min(world, nwin), 0, -1):
        if world % cand == 0 and nwin % cand == 0:
            nd = cand
            break
    return world // nd, nd


def disparity_shards(ndisp: int, nd: int, window: int = 64) -> List[Tuple[int, int]]:
    """nd contiguous disparity ranges whose starts are multiples of `window` (the library's requirement)."""
    nwin = (ndisp + window - 1) // window
    if nd < 1 or nd > nwin:
        raise ValueError("between 1 and ceil(ndisp / window) disparity shards")
    out = []
    for i in range(nd):
        w0, w1 = (nwin * i) // nd, (nwin * (i + 1)) // nd
        out.append((w0 * window, min(w1 * window, ndisp)))
    return out


def rank_shard(rank: int, world: int, H: int, ndisp: int, grid: Tuple[int, int] | None = None) -> Tuple[Tuple[int, int], Tuple[int, int], int, int]:
    """((y0, y1), (d0, d1), band index, shard index) of `rank`; ranks of one band are consecutive."""
    nb, nd = grid if grid else shard_grid(world, ndisp)
    bi, di = rank // nd, rank % nd
    return row_bands(H, nb)[bi], disparity_shards(ndisp, nd)[di], bi, di


def merge_triples(min1: np.ndarray, min2: np.ndarray, arg: np.ndarray):
    """Numpy statement of Min2::merge over the leading (shard) axis, shards in ascending disparity order:
    the result equals one sequential two-minimum scan over all disparities (lowest index wins ties)."""
    cur, last, a = min1[0].copy(), min2[0].copy(), arg[0].copy()
    for s in range(1, min1.shape[0]):
        oc, ol, oa = min1[s], min2[s], arg[s]
        take = (oc < cur) | ((oc == cur) & (oa < a))
        lo = np.where(take, oc, cur)
        other = np.where(take, cur, oc)
        last = np.minimum(other, np.minimum(last, ol))
        a = np.where(take, oa, a)
        cur = lo
    return cur, last, a


ShardFn = Callable[[np.ndarray, np.ndarray, Tuple[int, int], Tuple[int, int]], Tuple[np.ndarray, np.ndarray, np.ndarray]]


def disparity_2d_sharded(left: np.ndarray, right: np.ndarray, ndisp: int, rank: int, world: int, compute_shard: ShardFn,
                         device="cpu", group=None, grid: Tuple[int, int] | None = None):
    """One frame on a (row bands x disparity shards) grid of ranks.  `compute_shard(left, right, (y0, y1), (d0, d1))`
    returns the partial (min1, min2, argmin) float32 / float32 / int32 arrays of shape (y1-y0, W).  One all-gather of
    the triples, then every rank merges each band's shards.  Returns (argmin map uint8/int32 (H, W), confidence (H, W)).
    H must be divisible by the number of bands."""
    import torch
    import torch.distributed as dist

    H, W, _ = left.shape
    nb, nd = grid if grid else shard_grid(world, ndisp)
    if H % nb:
        raise ValueError("H must be divisible by the number of row bands")
    (y0, y1), (d0, d1), _, _ = rank_shard(rank, world, H, ndisp, (nb, nd))
    m1, m2, a = compute_shard(left, right, (y0, y1), (d0, d1))
    rows = y1 - y0
    mine = torch.from_numpy(np.stack([m1.astype(np.float32).view(np.int32), m2.astype(np.float32).view(np.int32), a.astype(np.int32)])).to(device)
    if world > 1:
        allp = torch.empty((world,) + tuple(mine.shape), dtype=torch.int32, device=mine.device)
        dist.all_gather_into_tensor(allp.view(-1), mine.contiguous().view(-1), group=group)
    else:
        allp = mine.unsqueeze(0)
    allp = allp.cpu().numpy().reshape(nb, nd, 3, rows, W)
    cur, last, arg = merge_triples(allp[:, :, 0].view(np.float32).transpose(1, 0, 2, 3).reshape(nd, H, W),
                                   allp[:, :, 1].view(np.float32).transpose(1, 0, 2, 3).reshape(nd, H, W),
                                   allp[:, :, 2].transpose(1, 0, 2, 3).reshape(nd, H, W))
    with np.errstate(invalid="ignore", divide="ignore"):
        conf = ((last - cur) / last).astype(np.float32)
    return arg, conf


def cuda_shard_fn(ctx, params) -> ShardFn:
    """The product's compute callback: asw_disparity_shard_device on this rank's GPU."""
    import torch

    def fn(left, right, band, dshard):
        H, W, _ = left.shape
        rows = band[1] - band[0]
        dl = torch.from_numpy(np.ascontiguousarray(left)).cuda()
        dr = torch.from_numpy(np.ascontiguousarray(right)).cuda()
        m1 = torch.empty((rows, W), dtype=torch.float32, device="cuda")
        m2 = torch.empty_like(m1)
        a = torch.empty((rows, W), dtype=torch.int32, device="cuda")
        ctx.disparity_shard_raw(dl.data_ptr(), dr.data_ptr(), W, H, params, band, dshard, m1.data_ptr(), m2.data_ptr(), a.data_ptr())
        ctx.sync()
        return m1.cpu().numpy(), m2.cpu().numpy(), a.cpu().numpy()

    return fn


# ---- row bands with a per-iteration halo exchange (alternative (i)) -------------------------------------
# A vertical pass reaches R rows, so instead of carrying r*R recomputed halo rows a band can carry R rows and fetch its
# neighbours' boundary rows between two iterations: no row is aggregated twice, the price is one neighbour exchange per
# iteration (R * W * D * 4 bytes each way: 63 MB at 4K x 256, ~90 us over NVLink).  This is the one real exchange step of
# the row-band sharding; everything else stays an all-gather of results.

def halo_exchange(top_send, bottom_send, top_recv, bottom_recv, rank: int, world: int, group=None):
    """Neighbour exchange of boundary rows between row bands (rank r owns band r; band r-1 lies above band r).
    Arguments are torch tensors (or None towards a frame border): this rank's first / last rows to hand out, and
    where the rows of the band above / below belong.  Works on NCCL (device tensors) and gloo (CPU tensors)."""
    import torch.distributed as dist

    if world == 1:
        return
    ops = []
    if rank > 0:
        ops.append(dist.P2POp(dist.isend, top_send, rank - 1, group))
        ops.append(dist.P2POp(dist.irecv, top_recv, rank - 1, group))
    if rank + 1 < world:
        ops.append(dist.P2POp(dist.isend, bottom_send, rank + 1, group))
        ops.append(dist.P2POp(dist.irecv, bottom_recv, rank + 1, group))
    for req in dist.batch_isend_irecv(ops):
        req.wait()


class _DevView:
    """Zero-copy view of raw device memory for torch (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def cuda_exchange_fn(ctx, rank: int, world: int, group=None, device=None):
    """The product's exchange callback for AswContext.disparity_band_exchange: the library hands out device addresses into
    its cost volume, NCCL send/recv moves the rows.  The library's stream is drained first (the rows of the iteration are
    final), the NCCL work is drained before returning (the next vertical pass reads the received rows)."""
    import torch

    def fn(_it, top_send, bottom_send, top_recv, bottom_recv, nbytes):
        ctx.sync()
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        view = lambda p: None if not p else torch.as_tensor(_DevView(p, nbytes), device=dev)
        halo_exchange(view(top_send), view(bottom_send), view(top_recv), view(bottom_recv), rank, world, group)
        torch.cuda.current_stream(dev).synchronize()

    return fn


class CudaHaloOverlap:
    """The product's callbacks for AswContext.disparity_band_exchange_async: the halo rows travel (NCCL send/recv on a
    communication stream of this rank) while the library aggregates the interior rows; nothing blocks the host.
      begin: the communication stream waits for the library's boundary stream (the band's border rows of this iteration),
             then the batched isend/irecv is enqueued on it;
      end:   the library's main stream waits for the communication stream (rows received, send rows free)."""

    def __init__(self, rank: int, world: int, group=None, device=None):
        import torch
        self.rank, self.world, self.group = rank, world, group
        self.dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.comm = torch.cuda.Stream(device=self.dev)
        self.keep = []                                   # tensor views of the library's rows, alive until the transfer is over

    def begin(self, _it, top_send, bottom_send, top_recv, bottom_recv, nbytes, boundary_stream):
        import torch
        import torch.distributed as dist
        view = lambda p: None if not p else torch.as_tensor(_DevView(p, nbytes), device=self.dev)
        ts, bs, tr, br = view(top_send), view(bottom_send), view(top_recv), view(bottom_recv)
        self.keep = [ts, bs, tr, br]
        self.comm.wait_stream(torch.cuda.ExternalStream(boundary_stream, device=self.dev))
        ops = []
        if self.rank > 0:
            ops += [dist.P2POp(dist.isend, ts, self.rank - 1, self.group), dist.P2POp(dist.irecv, tr, self.rank - 1, self.group)]
        if self.rank + 1 < self.world:
            ops += [dist.P2POp(dist.isend, bs, self.rank + 1, self.group), dist.P2POp(dist.irecv, br, self.rank + 1, self.group)]
        with torch.cuda.stream(self.comm):
            for req in dist.batch_isend_irecv(ops):
                req.wait()                               # NCCL: orders the communication stream after the transfer, does not block the host

    def end(self, _it, main_stream):
        import torch
        torch.cuda.ExternalStream(main_stream, device=self.dev).wait_stream(self.comm)


def disparity_row_exchange_cuda(ctx, d_left: int, d_right: int, W: int, H: int, params, rank: int, world: int, out_band,
                                group=None, timing: bool = False, overlap: bool = False):
    """This rank's band of one frame with halo exchange (device pointers of the FULL images in, `out_band` = torch uint8
    tensor (rows, W) on this rank's GPU out).  Returns the library's timing dict if asked.  Bands: row_bands(H, world)."""
    # `overlap`: hide the exchange under the interior rows (CudaHaloOverlap).  Measured on 8 B200s with NCCL 2.27 it LOSES
    # to the host-synchronous exchange (25.6 vs 23.7 ms on the 4K frame): NCCL's send/recv kernels occupy SMs, and the
    # persistent vertical-pass grid (one CTA per SM, static tile list) then runs with late CTAs.  The in-process path
    # (asw_multi_*: copy-engine peer copies ordered by events) does profit from the overlap (20.7 ms).
    y0, y1 = row_bands(H, world)[rank]
    if world == 1:
        return ctx.disparity_raw(d_left, d_right, W, H, params, None, out_band.data_ptr(), None, timing=timing)
    if overlap:
        ops = CudaHaloOverlap(rank, world, group)
        return ctx.disparity_band_exchange_async(d_left, d_right, W, H, params, (y0, y1), None, out_band.data_ptr(), None, ops.begin, ops.end,
                                                 timing=timing)
    return ctx.disparity_band_exchange(d_left, d_right, W, H, params, (y0, y1), None, out_band.data_ptr(), None,
                                       cuda_exchange_fn(ctx, rank, world, group), timing=timing)
