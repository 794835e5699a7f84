"""Multi-process host logic of the sharded paths, on CPU over gloo (world_size 2 and 3).

The compute callback is the CPU oracle run on the rows that can influence the band (r*R halo rows,
clipped at the frame border) -- the same halo rule asw_disparity_band_device implements on the GPU --
so the test checks the band geometry, the padding of uneven bands and the all-gather, and that the
sharded result is bit-identical to the unsharded one."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import crop_pair


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_band(left, right, y0, y1, D=24, r=2):
    from oracle import asw_oracle as O
    from stereo_matchin_b200.sharding import band_input_rows
    H = left.shape[0]
    ya, yb = band_input_rows(y0, y1, H, 16, r)
    res = O.asw_hot_path(np.ascontiguousarray(left[ya:yb]), np.ascontiguousarray(right[ya:yb]),
                         O.OracleParams(ndisp=D, iterations=r), use_fma=True)
    return res["d_ref"][y0 - ya:y1 - ya].astype(np.uint8)


def _two_min_scan(cost, d0):
    """Sequential two-minimum scan of asw_wta.cl over the planes of `cost` (global index = d0 + plane)."""
    cur = np.full(cost.shape[1:], 100000.0, np.float32)
    last = cur.copy()
    arg = np.zeros(cost.shape[1:], np.int32)
    for i in range(cost.shape[0]):
        t = cost[i]
        last = np.where(t < last, t, last)
        arg = np.where(t < cur, d0 + i, arg)
        last = np.where(t < cur, cur, last)
        cur = np.where(t < cur, t, cur)
    return cur, last, arg


def _oracle_shard(left, right, band, dshard, D=128, r=1):
    """Partial WTA triple of disparities [d0, d1) on rows [y0, y1): planes d0..d1-1 of the oracle's aggregated volume."""
    from oracle import asw_oracle as O
    from stereo_matchin_b200.sharding import band_input_rows
    (y0, y1), (d0, d1) = band, dshard
    ya, yb = band_input_rows(y0, y1, left.shape[0], 16, r)
    res = O.asw_hot_path(np.ascontiguousarray(left[ya:yb]), np.ascontiguousarray(right[ya:yb]),
                         O.OracleParams(ndisp=D, iterations=r), use_fma=True, want_cost=True)
    return _two_min_scan(res["cost"][d0:d1, y0 - ya:y1 - ya], d0)


def _oracle_band_exchange(left, right, rank, world, D=24, r=3, R=16):
    """CPU statement of asw_disparity_band_exchange_device: this rank keeps R halo rows of the cost volume, aggregates
    exactly its own rows every iteration (oracle operators on the local rows: their clamping only touches halo rows, which
    are discarded) and swaps boundary rows with its neighbours through sharding.halo_exchange between iterations."""
    import torch
    from oracle import asw_oracle as O
    from stereo_matchin_b200.sharding import row_bands, halo_exchange
    H = left.shape[0]
    y0, y1 = row_bands(H, world)[rank]
    ya, yb = max(0, y0 - R), min(H, y1 + R)
    L, Rr = np.ascontiguousarray(left[ya:yb]), np.ascontiguousarray(right[ya:yb])
    a, b = y0 - ya, y1 - ya                                      # own rows inside the local arrays
    cost = np.ascontiguousarray(O.asw_aggr(L, Rr, D))             # (D, rows, W): raw cost needs no exchange
    vl, vr = O.asw_support(L, True), O.asw_support(Rr, True)
    hl, hr = O.asw_support(L, False), O.asw_support(Rr, False)
    for it in range(r):
        h = O.asw_hcost_aggregation(hl, hr, O.asw_vcost_aggregation(vl, vr, cost, use_fma=True)[0], use_fma=True)
        cost[:, a:b] = h[:, a:b]                                 # only the own rows are valid (and needed)
        if it + 1 < r:
            # volume rows are (D, W) slabs here; the exchange buffers are contiguous copies
            ts = torch.from_numpy(np.ascontiguousarray(cost[:, a:a + R])) if rank > 0 else None
            bs = torch.from_numpy(np.ascontiguousarray(cost[:, b - R:b])) if rank + 1 < world else None
            tr = torch.empty_like(ts) if ts is not None else None
            br = torch.empty_like(bs) if bs is not None else None
            halo_exchange(ts, bs, tr, br, rank, world)
            if tr is not None:
                cost[:, a - R:a] = tr.numpy()
            if br is not None:
                cost[:, b:b + R] = br.numpy()
    return O.asw_wta(np.ascontiguousarray(cost[:, a:b]), right_view=False)["d_ref"].astype(np.uint8), cost[:, a:b].copy()


def _worker(rank, world, port, mode, q):
    import torch.distributed as dist
    from oracle import asw_oracle as O
    from stereo_matchin_b200 import sharding
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    O.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        if mode == "grid":
            L, R = crop_pair("teddy", 150, 100, 150, 64)
            arg, conf = sharding.disparity_2d_sharded(L, R, 128, rank, world, _oracle_shard, grid=(2, 2))
            q.put((rank, np.stack([arg.astype(np.float32), conf])))
        elif mode == "exchange":
            L, R = crop_pair("teddy", 60, 10, 96, 118)      # 118 rows: uneven bands, every band >= 16 rows
            band, vol = _oracle_band_exchange(L, R, rank, world)
            import torch
            full = sharding.gather_bands(torch.from_numpy(band), L.shape[0], L.shape[1], rank, world)
            q.put((rank, (full.numpy(), vol)))
        elif mode == "bands":
            L, R = crop_pair("teddy", 40, 0, 120, 151)      # 151 rows: uneven bands
            full = sharding.disparity_row_sharded(L, R, rank, world, _oracle_band)
            q.put((rank, full.numpy()))
        else:
            pairs = [crop_pair("cones", 30 * i, 20 * i, 64, 40) for i in range(2 * world)]
            maps = sharding.disparity_pair_sharded(pairs, rank, world, lambda l, r: _oracle_band(l, r, 0, l.shape[0]))
            q.put((rank, maps.numpy()))
    finally:
        dist.destroy_process_group()


def _run(world, mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = {}
    try:
        for _ in range(world):
            rank, arr = q.get(timeout=90)
            out[rank] = arr
    finally:
        for p in procs:
            p.join(30)
            if p.is_alive():
                p.kill()
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    return out


@pytest.mark.parametrize("world", [2, 3])
def test_row_band_sharding_matches_single_process(world):
    out = _run(world, "bands")
    L, R = crop_pair("teddy", 40, 0, 120, 151)
    ref = _oracle_band(L, R, 0, L.shape[0])
    for r in range(world):
        assert np.array_equal(out[r], ref), f"rank {r}: gathered map differs from the unsharded result"


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_bands_match_single_process(world):
    """Row bands with R exchanged halo rows per iteration (no recomputed halo): disparity map and the final cost volume of
    every band are bit-identical to the unsharded oracle."""
    from oracle import asw_oracle as O
    from stereo_matchin_b200.sharding import row_bands
    out = _run(world, "exchange")
    L, R = crop_pair("teddy", 60, 10, 96, 118)
    ref = O.asw_hot_path(L, R, O.OracleParams(ndisp=24, iterations=3), use_fma=True, want_cost=True)
    for r in range(world):
        full, vol = out[r]
        y0, y1 = row_bands(L.shape[0], world)[r]
        assert np.array_equal(full, ref["d_ref"].astype(np.uint8)), f"rank {r}: gathered map differs from the unsharded result"
        assert np.array_equal(vol.view(np.uint32), ref["cost"][:, y0:y1].view(np.uint32)), f"rank {r}: cost volume of the band differs"


def test_pair_sharding_gathers_every_map():
    world = 2
    out = _run(world, "pairs")
    pairs = [crop_pair("cones", 30 * i, 20 * i, 64, 40) for i in range(2 * world)]
    ref = np.stack([_oracle_band(l, r, 0, l.shape[0]) for l, r in pairs])
    for r in range(world):
        assert np.array_equal(out[r], ref)


def test_band_geometry():
    from stereo_matchin_b200 import sharding as s
    assert s.row_bands(2160, 8) == [(270 * i, 270 * (i + 1)) for i in range(8)]
    assert s.row_bands(10, 4) == [(0, 2), (2, 5), (5, 7), (7, 10)]
    assert s.band_input_rows(270, 540, 2160, 16, 7) == (158, 652)          # 112 halo rows per side
    assert s.band_input_rows(0, 270, 2160, 16, 7) == (0, 382)              # clipped at the frame border
    rows = s.iteration_rows(270, 540, 2160, 16, 7)
    assert rows[0] == (174, 636) and rows[-1] == (270, 540)                # halo shrinks by R per iteration
    assert [b - a for a, b in rows] == [462 - 32 * i for i in range(7)]
    # halo overhead of the 4K frame: +36 % at 8 GPUs, +4 % at 2
    assert abs(s.band_work_fraction(2160, 8) - 1.311) < 0.01
    assert abs(s.band_work_fraction(2160, 2) - 1.044) < 0.01
    assert s.band_work_fraction(2160, 1) == 1.0
    assert [len(r) for r in s.pair_shards(1024, 8)] == [128] * 8


def test_2d_sharding_matches_single_process():
    """4 ranks = 2 row bands x 2 disparity shards: merged result == unsharded oracle (indices and confidence)."""
    from oracle import asw_oracle as O
    from stereo_matchin_b200 import sharding
    assert sharding.shard_grid(4, 256) == (2, 2) and sharding.shard_grid(8, 256) == (4, 2) and sharding.shard_grid(4, 130) == (2, 2)
    assert sharding.shard_grid(4, 128, window=64) == (2, 2) and sharding.shard_grid(3, 256) == (3, 1)
    assert sharding.disparity_shards(256, 4) == [(0, 64), (64, 128), (128, 192), (192, 256)]
    assert sharding.disparity_shards(130, 3) == [(0, 64), (64, 128), (128, 130)]
    L, R = crop_pair("teddy", 150, 100, 150, 64)
    ref = O.asw_hot_path(L, R, O.OracleParams(ndisp=128, iterations=1), use_fma=True)
    out = _run(4, "grid")
    for rank in range(4):
        assert np.array_equal(out[rank][0].astype(np.int32), ref["d_ref"].astype(np.int32)), f"rank {rank}: disparity indices differ"
        assert np.array_equal(out[rank][1].view(np.uint32), ref["conf_ref"].view(np.uint32)), f"rank {rank}: confidence differs"


def test_merge_triples_equals_sequential_scan():
    from stereo_matchin_b200.sharding import merge_triples
    rng = np.random.default_rng(11)
    cost = rng.integers(0, 6, (200, 9, 13)).astype(np.float32)        # many exact ties
    whole = _two_min_scan(cost, 0)
    parts = [_two_min_scan(cost[a:b], a) for a, b in ((0, 64), (64, 128), (128, 192), (192, 200))]
    merged = merge_triples(np.stack([p[0] for p in parts]), np.stack([p[1] for p in parts]), np.stack([p[2] for p in parts]))
    for w, m in zip(whole, merged):
        assert np.array_equal(w, m)
