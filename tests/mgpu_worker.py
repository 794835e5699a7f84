"""torchrun worker of tests/test_gpu_multi.py: one frame split into row bands over the ranks' GPUs, halo rows exchanged
with NCCL send/recv every iteration, bands all-gathered; rank 0 compares the map with its own one-GPU result."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

from stereo_matchin_b200 import api, sharding, synth


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L, R = synth.make_pair(640, 360, 128, seed=11)[:2]
    H, W, _ = L.shape
    p = api.AswParams(ndisp=128)
    ctx = api.AswContext(local)
    dl, dr = torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()
    y0, y1 = sharding.row_bands(H, world)[rank]
    band = torch.empty((y1 - y0, W), dtype=torch.uint8, device="cuda")
    fulls = []
    for overlap in (True, False):       # exchange hidden under the interior rows / host-synchronous exchange
        band.zero_()
        sharding.disparity_row_exchange_cuda(ctx, dl.data_ptr(), dr.data_ptr(), W, H, p, rank, world, band, overlap=overlap)
        ctx.sync()
        fulls.append(sharding.gather_bands(band, H, W, rank, world))
    ok = True
    if rank == 0:
        one = torch.empty((H, W), dtype=torch.uint8, device="cuda")
        ctx.disparity_raw(dl.data_ptr(), dr.data_ptr(), W, H, p, None, one.data_ptr(), None)
        ctx.sync()
        ok = all(bool(torch.equal(f, one)) for f in fulls)
        print("equals_1gpu", ok, "world", world, flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
