import os, sys, json
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import elmerfem_b200 as B
from elmerfem_b200 import synth
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
def allsum(v):
    t = torch.tensor([v], dtype=torch.float64, device="cuda"); dist.all_reduce(t); return float(t.item())
kind = sys.argv[1]
if kind == "elasticity":
    p = synth.elasticity_slab(137, 137, 140 * world - 1, rank, world, allreduce_sum=allsum)
else:
    p = synth.heat_slab(200, 200, 200 * world, rank, world, allreduce_sum=allsum)
M = B.Matrix()
ids = [B.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
M.comm_init(world, rank, ids[0])
M.set_partition(p["gn"], p["rows"], p["cols"], p["goffset"], 1, p["ndeg"])
M.set_values(p["vals"])
for rep in range(2):
    ms = M.time_matvec(20)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0: print(kind, "HALO_DEBUG", os.environ.get("B200_HALO_DEBUG", "0"), "matvec ms (max over ranks)", float(t.item()), flush=True)
M.close()
dist.barrier(); dist.destroy_process_group()
