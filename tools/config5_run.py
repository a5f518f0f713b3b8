#!/usr/bin/env python
"""BASELINE.json configs[4]: HiFi reads (e = 0.005), overlap SpGEMM + batched X-drop alignment chained on N GPUs.
    python -m torch.distributed.run --nproc-per-node N tools/config5_run.py [n_reads] [out.json]
Every rank: sharded SpGEMM (bella_b200.distributed, NVLink mode: it ends with the result of the rank's own columns on the device)
-> seeds taken from that device result -> X-drop extension + accept/reject of its own pairs (bella_b200.xdrop; all reads resident on
every GPU, no collective: SURVEY.md 8f / bella_b200/distributed_xdrop.py).  Rank 0 runs the same chain on one GPU and compares a
checksum of all alignment results.  Variant (i) of SURVEY.md 8d: coverage 6x with [l,u] = [2,8]."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import bench
from bella_b200 import distributed as bd, spgemm, xdrop

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
out_path = sys.argv[2] if len(sys.argv) > 2 else None
local = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
w = dict(bench.CONFIGS[5], n_reads=n_reads)
if rank != 0: dist.barrier()
inp = bench.load_workload(w, need_seqs=True)
if rank == 0: dist.barrier()
XD, RATIO, DELTA = 7, 0.7, 0.1

def csum(rows, cols, out8):
    with np.errstate(over="ignore"):
        x = (rows.astype(np.uint64) << np.uint64(32) | cols.astype(np.uint64)) * np.uint64(0x9E3779B97F4A7C15)
        for c in range(8):
            x = (x ^ out8[:, c].astype(np.int64).astype(np.uint64)) * np.uint64(0xBF58476D1CE4E5B9)
            x ^= x >> np.uint64(31)
        return int(x.sum(dtype=np.uint64))

sh = single = None
if world > 1:
    sh = bd.ShardedOverlapSpGEMM(local, mode="nvlink")
    sh.load_shard(inp, pinned=True)
else:
    single = spgemm.OverlapSpGEMM(local)
al = xdrop.XdropAligner(local)
al.set_reads(inp.seqs, inp.seq_off)
al.set_params(inp.kmer_size, XD, RATIO, DELTA, -1)
stream = torch.cuda.current_stream(dev)

def chain():
    """-> (Z, accepted, t_spgemm_ms, t_xdrop_ms, (rows, cols, out8) on the device)"""
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record(stream)
    if world > 1:
        Z, flops, (lo, hi) = sh.step()
        g = sh.g
    else:
        g = single
        g.set_inputs(inp)
        g.symbolic(); g.numeric_device()
        lo, hi = 0, inp.n_reads
    r = g.result_device()
    Z = int(r["nnz"])
    e1.record(stream)
    ncols = hi - lo
    def dev_view(ptr, n, dtype, esize):
        # wrap a raw device pointer (valid until the next pass on the handle) as a tensor through __cuda_array_interface__
        class A: pass
        a = A(); a.__cuda_array_interface__ = {"shape": (n,), "typestr": {4: "<i4", 2: "<i2"}[esize], "data": (ptr, False), "version": 2}
        return torch.as_tensor(a, device=dev)
    colptr = dev_view(r["colptrC"], ncols + 1, torch.int32, 4).to(torch.int64)
    cols = torch.repeat_interleave(torch.arange(lo, hi, device=dev, dtype=torch.int64), colptr[1:] - colptr[:-1]).to(torch.int32)
    out = torch.zeros((max(Z, 1), 8), dtype=torch.int32, device=dev)
    torch.cuda.synchronize(dev)
    if Z:
        al.align_device(Z, r["rowids"], cols, r["posH"], r["posV"], out)
        al.sync()
    e2.record(stream)
    torch.cuda.synchronize(dev)
    rows = dev_view(r["rowids"], max(Z, 1), torch.int32, 4)[:Z].clone()
    return Z, int(out[:Z, 7].sum()), e0.elapsed_time(e1), al.stats()["kernel_ms"], (rows, cols, out[:Z])

for _ in range(2):
    chain()
dist.barrier()
t0 = time.perf_counter()
steps = 3
acc = np.zeros(2)
for _ in range(steps):
    Z, ok, ts, tx, res = chain()
    acc += [ts, tx]
torch.cuda.synchronize(dev); dist.barrier()
wall = (time.perf_counter() - t0) / steps * 1e3
rows, cols, out8 = (t.cpu().numpy() for t in res)
mine = torch.tensor([Z, ok, csum(rows.view(np.uint32), cols.view(np.uint32), out8) >> 1], dtype=torch.int64, device=dev)   # >> 1: keep it in int64
allv = [torch.zeros_like(mine) for _ in range(world)]
dist.all_gather(allv, mine)
tmax = torch.tensor(list(acc / steps) + [wall], dtype=torch.float64, device=dev)
dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
if rank == 0:
    Zt, okt = sum(int(v[0]) for v in allv), sum(int(v[1]) for v in allv)
    line = {"config": bench.workload_name(w) + " + X-drop (x=7)", "n_gpus": world, "pairs": Zt, "accepted": okt, "ms_per_step_wall": float(tmax[2]),
            "spgemm_ms_max": float(tmax[0]), "xdrop_kernel_ms_max": float(tmax[1]), "pairs_per_s": Zt / (float(tmax[2]) * 1e-3), "nnz_A": inp.nnz}
    # the same chain on one GPU (rank 0), result by result
    one = spgemm.overlap_spgemm(inp, device=local)
    cols1 = np.repeat(np.arange(inp.n_reads, dtype=np.uint32), np.diff(one["colptrC"].astype(np.int64)))
    o1 = al.align(one["rowids"], cols1, one["posH"], one["posV"])
    line["single_gpu_pairs"], line["single_gpu_accepted"] = int(len(cols1)), int(o1[:, 7].sum())
    # per-rank checksums are over disjoint pair sets: compare the multiset through the per-rank slices of the single-GPU result
    bounds = np.searchsorted(cols1, np.array(sh.cuts if world > 1 else [0, inp.n_reads], dtype=np.uint32), side="left")
    ref = [csum(one["rowids"][a:b], cols1[a:b], o1[a:b]) >> 1 for a, b in zip(bounds[:-1], bounds[1:])]
    line["parity"] = "every rank's alignments == the single-GPU chain's on the same pairs" if ref == [int(v[2]) for v in allv] and Zt == len(cols1) else "MISMATCH"
    print(json.dumps(line), flush=True)
    if out_path:
        json.dump(line, open(out_path, "w"), indent=1)
    if line["parity"] == "MISMATCH":
        sys.exit(1)
dist.barrier()
dist.destroy_process_group()
