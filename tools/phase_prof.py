#!/usr/bin/env python
"""Per-phase SM cycles of k_group_fold / k_bucket on the bench workload (profiling build: -DBELLA_PHASE_CLOCKS).
   nvcc ... -DBELLA_PHASE_CLOCKS -o scratch/libbella_b200_phase.so bella_b200/csrc/bella_b200.cu
   python tools/phase_prof.py [n_reads] [out.json]"""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from bella_b200 import _build
os.environ["BELLA_B200_LIB"] = os.path.join(ROOT, "scratch", "libbella_b200_phase.so")
import bench
from bella_b200 import spgemm
import torch

w = dict(bench.WORKLOAD)
if len(sys.argv) > 1:
    w["n_reads"] = int(sys.argv[1])
inp = bench.load_workload(w, need_seqs=False)
g = spgemm.OverlapSpGEMM(0)
L = spgemm.lib()
dev = torch.device("cuda", 0)
def dev_t(a):
    return torch.from_numpy(a.view(np.int32) if a.dtype == np.uint32 else a.view(np.int16) if a.dtype == np.uint16 else a).to(dev)
d = {k: dev_t(getattr(inp, k)) for k in ("B_colptr", "B_rowids", "B_values", "B_strand", "read_len")}
def step():
    g.set_inputs_device(inp.n_reads, inp.n_kmers, inp.nnz, (d["B_colptr"], d["B_rowids"], d["B_values"]), d["read_len"], d["B_strand"], inp.kmer_size, inp.bin_size)
    return g.run_resident()
for _ in range(3):
    step()
torch.cuda.synchronize()
L.bella_b200_debug_phases(None, 1)
steps = 3
for _ in range(steps):
    Z, F = step()
torch.cuda.synchronize()
out = (ctypes.c_ulonglong * 32)()
L.bella_b200_debug_phases(out, 0)
v = np.array(list(out), dtype=np.float64) / steps
names = {15: "gf unit params + clear", 0: "gf wait for the bulk load", 1: "gf row bitmap (atomicOr)", 2: "gf popc scan", 3: "gf pair index + slot (atomic_add16)", 4: "gf count scan", 5: "gf placement + multiply",
         6: "gf rank in pair", 7: "gf fold + pair results", 8: "gf warp path + huge + unit end", 16: "bk TMA wait + count", 17: "bk scan", 18: "bk group by k-mer", 19: "bk rank by read + write"}
gf = sum(v[i] for i in range(9)) + v[15]; bk = sum(v[i] for i in range(16, 20))
res = {"Z": int(Z), "products": int(F), "timings": g.timings()}
for i, nme in names.items():
    tot = gf if i < 16 else bk
    res[nme] = {"cycles": v[i], "share": v[i] / tot if tot else 0}
    print(f"{nme:40s} {v[i]:16.0f} cyc  {100 * v[i] / tot if tot else 0:5.1f}%")
print(f"fold queue: long pairs: {v[13]:.0f} pairs, {v[11]:.0f} products, {v[9]:.0f} warp-cycles ({v[9] / max(v[11], 1):.1f} per product); "
      f"short: {v[14]:.0f} pairs, {v[12]:.0f} products, {v[10]:.0f} warp-cycles per tile-sum ({v[10] / max(v[12], 1):.1f} per product)")
res["long"] = {"pairs": v[13], "products": v[11], "warp_cycles": v[9]}
res["short"] = {"pairs": v[14], "products": v[12], "warp_cycles": v[10]}
print(json.dumps(res["timings"]))
if len(sys.argv) > 2:
    json.dump(res, open(sys.argv[2], "w"), indent=1)
