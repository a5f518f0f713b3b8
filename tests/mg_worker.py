"""Worker of tests/test_gpu_multi.py (launched under torchrun, one rank per GPU): runs the sharded
path on a seeded input and checks this rank's columns against the CPU oracle, bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    from bella_b200 import distributed as bd, frontend as fe
    import oracle_lib as ol
    mode = sys.argv[1] if len(sys.argv) > 1 else "exchange"
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    cases = [dict(n_reads=3000, read_len=6000, seed=31), dict(n_reads=700, read_len=5000, coverage=40.0, err=0.01, seed=13, hi=80)]
    for kw in cases:
        inp = fe.synthetic(**kw)
        want = ol.oracle_spgemm(inp)
        sh = bd.ShardedOverlapSpGEMM(local, mode=mode)
        sh.load_shard(inp)
        for _ in range(2):
            Z, flops, (lo, hi), res = sh.step(fetch=True)
        colptrC, rows, cnt, pH, pV = res
        z0, z1 = int(want.colptrC[lo]), int(want.colptrC[hi])
        np.testing.assert_array_equal(colptrC.astype(np.int64), want.colptrC[lo:hi + 1].astype(np.int64) - z0)
        np.testing.assert_array_equal(rows, want.rowids[z0:z1])
        np.testing.assert_array_equal(cnt, want.count[z0:z1])
        np.testing.assert_array_equal(pH, want.posH[z0:z1])
        np.testing.assert_array_equal(pV, want.posV[z0:z1])
        tot = torch.tensor([Z, hi - lo], dtype=torch.int64, device="cuda")
        dist.all_reduce(tot)
        assert int(tot[0]) == want.nnz and int(tot[1]) == inp.n_reads, (tot, want.nnz)
        sh.close()
        if rank == 0:
            print(f"mg ok: mode={mode} world={world} n={inp.n_reads} Z={want.nnz}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
