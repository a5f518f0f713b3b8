import os, sys, time
sys.path.insert(0, '.')
import numpy as np, torch, torch.distributed as dist
from bella_b200 import distributed as bd, frontend as fe
import bench
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank = dist.get_rank()
if rank != 0: dist.barrier()
inp = bench.load_workload(dict(bench.WORKLOAD), need_seqs=False)
if rank == 0: dist.barrier()
sh = bd.ShardedOverlapSpGEMM(local, mode=os.environ.get("BELLA_MG_MODE","nvlink")); sh.load_shard(inp, pinned=True)
for _ in range(3): sh.step()
sh.prof = {}
for _ in range(5): sh.step()
if rank in (0, dist.get_world_size() - 1):
    print(rank, {k: round(v / 5, 3) for k, v in sh.prof.items()}, flush=True)
dist.barrier(); dist.destroy_process_group()
