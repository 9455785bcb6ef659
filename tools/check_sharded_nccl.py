"""Multi-GPU check (torchrun, NCCL): price_batch_sharded / ladder_batch_sharded with the closing all_gather equal the
single-GPU results bit for bit on every rank."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastamericanoptionpricing_b200 import batch, workloads as wl, price_batch_sharded, ladder_batch_sharded
from fastamericanoptionpricing_b200.batch import AloConfig
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = AloConfig(**wl.C2_CFG)
N = 100_003
b = wl.c2_batch(N)
a = [torch.as_tensor(b[k]).cuda() for k in ("r", "q", "sigma", "K", "T", "S", "is_put")]
full, (lo, hi) = price_batch_sharded(*a, cfg=cfg, gather=True)
ref = batch.price_batch(*a, cfg=cfg, want_boundary=False)["price"]
eq = lambda x, y: bool(((x == y) | (torch.isnan(x) & torch.isnan(y))).all())      # bit-equal, NaN for NaN (upstream-NaN options exist)
ok1 = eq(full, ref)
G, Kl = 1001, np.linspace(60, 140, 33)
g = [x[:G] for x in a]
lad, rows = ladder_batch_sharded(g[0], g[1], g[2], g[4], g[5], g[6], Kl, cfg=cfg, gather=True)
lref = batch.ladder_batch(g[0], g[1], g[2], g[4], g[5], g[6], Kl, cfg=cfg)
ok2 = eq(lad, lref)
flag = torch.tensor([int(ok1 and ok2)], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("world %d: price_batch_sharded gather equal %s, ladder_batch_sharded gather equal %s, all ranks ok %d; slice of rank 0: [%d, %d)"
          % (world, ok1, ok2, flag.item(), lo, hi))
dist.destroy_process_group()
