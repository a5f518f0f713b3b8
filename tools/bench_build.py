#!/usr/bin/env python
"""Measurement of the "next" row f2 (matrix construction): tuples -> B on the device
(bella_b200_set_inputs_tuples) next to the reference's own CSC constructor + MergeDuplicates + Transpose
(oracle/_ref, src/CSC.cpp:289-479) on the host cores, on the tuples of BASELINE.json configs[1].
Prints one JSON line.  Not part of bench.py's contract (that measures the SpGEMM)."""
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    from bella_b200 import frontend as fe, spgemm
    import oracle_lib as ol
    n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
    inp = fe.synthetic(n_reads, 10000, coverage=30.0, err=0.15, seed=2, keep_tuples=True)
    tk, tr, tp = inp.tuples
    order = np.lexsort((tp, tr))                      # read by read, in position order (src/main.cpp:393-416)
    tk, tr, tp = np.ascontiguousarray(tk[order]), np.ascontiguousarray(tr[order]), np.ascontiguousarray(tp[order])
    T = len(tk)
    pin = lambda a, dt: torch.from_numpy(a.view(dt)).pin_memory().numpy().view(a.dtype)
    tk, tr, tp = pin(tk, np.int32), pin(tr, np.int32), pin(tp, np.int16)
    strand = pin(np.zeros((T + 7) // 8 + 8, dtype=np.uint8), np.uint8)      # strand bits do not change the construction
    g = spgemm.OverlapSpGEMM(0)
    build, wall = [], []
    for it in range(5):
        t0 = time.perf_counter()
        g.set_inputs_tuples(inp.n_kmers, inp.n_reads, tk, tr, tp, strand, inp.read_len, inp.kmer_size, inp.bin_size)
        wall.append((time.perf_counter() - t0) * 1e3)
        build.append(g.get_B()[4] if it == 4 else None)
    colptr, rows, vals, _, ms = g.get_B()
    same = bool((colptr == inp.B_colptr).all() and (rows == inp.B_rowids).all() and (vals == inp.B_values).all())
    line = {"stage": "matrix construction (tuples -> B)", "tuples": int(T), "nnz": int(inp.nnz), "device_build_ms": float(ms),
            "device_wall_ms_with_upload": float(np.median(wall[1:])), "identical_to_host_front_end": same}
    if ol.have_ref():
        L = ol.ref()
        Bc = np.zeros(inp.n_reads + 1, np.uint32); Br = np.zeros(T, np.uint32); Bv = np.zeros(T, np.uint16)
        Ac = np.zeros(inp.n_kmers + 1, np.uint32); Ar = np.zeros(T, np.uint32); Av = np.zeros(T, np.uint16)
        t0 = time.perf_counter()
        L.bella_ref_build(ctypes.c_uint32(inp.n_kmers), ctypes.c_uint32(inp.n_reads), ctypes.c_uint64(T), ol._p(tk), ol._p(tr), ol._p(tp),
                          ctypes.c_int(0), ol._p(Bc), ol._p(Br), ol._p(Bv), ol._p(Ac), ol._p(Ar), ol._p(Av))
        line["reference_cpu_ms"] = (time.perf_counter() - t0) * 1e3
        line["reference_cores"] = L.bella_ref_max_threads()
        line["reference_note"] = "CSC(tuples) + MergeDuplicates + Transpose of the unmodified reference, incl. copying the tuples into its vector<tuple>"
    print(json.dumps(line), flush=True)
    g.close()


if __name__ == "__main__":
    main()
