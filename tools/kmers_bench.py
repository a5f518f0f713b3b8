"""Row f3 on the GPU box: device k-mer selection (bella_kmers) against the oracle for parity and against the reference's own
SplitCount (oracle/_ref, all host threads it can use -- one, see oracle/ref_driver.cpp) for time.
  python tools/kmers_bench.py [n_reads] [read_len] [out.json]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402

import oracle_lib as ol  # noqa: E402
from bella_b200 import frontend as fe, kmers  # noqa: E402

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
read_len = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
G = max(int(n_reads * read_len / 30.0), 4 * read_len)
seqs, offs = fe.simulate_reads(G, n_reads, read_len, 0.15, (0.10, 0.60, 0.30), 2)
inp = fe.OverlapInputs(n_reads=n_reads, n_kmers=0, nnz=0, A_colptr=None, A_rowids=None, A_values=None, A_strand=None, B_colptr=None,
                       B_rowids=None, B_values=None, B_strand=None, read_len=None, kmer_size=17, seqs=seqs, seq_off=offs)
out = {"n_reads": n_reads, "read_len": read_len, "bases": int(len(seqs))}
c = kmers.KmerCounter(0)
c.count(seqs, offs, 17, 2, 8)                                   # warm-up (allocations)
t = time.time()
got = c.count(seqs, offs, 17, 2, 8)
out["device_e2e_s"] = time.time() - t
out["device_kernel_ms"] = c.stats()["kernel_ms"]
out["n_kmers"], out["n_tuples"] = int(got["n_kmers"]), int(len(got["t_read"]))
t = time.time()
r, p, nk = ol.oracle_reliable_occurrences(inp, 17, 2, 8)
out["oracle_s"] = time.time() - t
out["parity"] = bool(nk == got["n_kmers"] and np.array_equal(r, got["t_read"]) and np.array_equal(p, got["t_pos"]))
if ol.have_ref() and n_reads * read_len <= 400_000_000:
    t = time.time()
    rr, rp, rk = ol.ref_reliable_occurrences(inp, 17, 2, 8, "/tmp/bella_kmers_bench.fastq")
    out["reference_splitcount_s"] = time.time() - t
    out["reference_agrees"] = bool(rk == nk and np.array_equal(rr, r) and np.array_equal(rp, p))
print(json.dumps(out))
dst = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", "kmers_bench.json")
os.makedirs(os.path.dirname(dst), exist_ok=True)
json.dump(out, open(dst, "w"), indent=1)
