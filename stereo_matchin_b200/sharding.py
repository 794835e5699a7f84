"""Multi-GPU sharding of the ASW hot path: one process per GPU, torch.distributed for the plumbing.

The path shards without any mid-computation exchange:
  * pair sharding  -- a batch of independent stereo pairs is split across ranks;
  * row-band sharding -- one large frame is split into row bands.  A vertical pass reaches R rows
    per iteration, so a band needs r*R real halo rows on each side (clipped at the frame border,
    where the reference's clamp-to-edge applies); the library shrinks the halo by R per iteration
    (asw_disparity_band_device), so the band is bit-identical to the same rows of a 1-GPU run.
The only collective is one all-gather of the uint8 disparity bands / maps (NCCL on GPUs; the CPU
tests run the same code over gloo).  The compute callback is injected so that the host logic can be
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
