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
