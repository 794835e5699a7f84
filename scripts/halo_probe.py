"""torchrun probe: time of the halo exchange alone (cfg4 geometry: 16 rows x 3872 cols x 256 d x 4 B = 63.4 MB each way)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from stereo_matchin_b200 import sharding
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 16 * 3872 * 256 * 4
ts, bs, tr, br = (torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(4))
def ex():
    sharding.halo_exchange(ts if rank > 0 else None, bs if rank + 1 < world else None, tr if rank > 0 else None, br if rank + 1 < world else None, rank, world)
    torch.cuda.current_stream().synchronize()
for _ in range(3): ex()
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): ex()
dt = (time.perf_counter() - t0) / 20
t = torch.tensor([dt], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0: print("halo exchange: %.3f ms per exchange (max over ranks), %.0f GB/s per direction per link" % (t.item() * 1e3, n / t.item() / 1e9), "env", {k: v for k, v in os.environ.items() if k.startswith("NCCL")})
dist.destroy_process_group()
