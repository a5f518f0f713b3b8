#!/usr/bin/env python
"""f1 against its bar: the reference's CUDA aligner LOGAN (loganGPU/functions.cuh:223-689, compiled for sm_100a from
/root/reference into oracle/_ref/libbella_logan.so by oracle/Makefile) and bella_b200.xdrop on the SAME pairs on the same box.
Two batches: the GPU test's (400 reads x 3 kb, 6 k pairs, where LOGAN reproduces the oracle) and the bench batch (3 k reads x 8 kb,
60 k pairs).  LOGAN is called twice and the second call counts (the first pays CUDA context / allocation set-up).
    python tools/xdrop_vs_logan.py [out.json]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as ol
import test_xdrop_gpu as T
from bella_b200 import frontend as fe

out = []
for name, inp, npairs in (("400 reads x 3 kb", fe.synthetic(400, 3000, seed=101), 6000), ("3000 reads x 8 kb", fe.synthetic(3000, 8000, seed=13), 60000)):
    pairs = T.candidate_pairs(inp, npairs, seed=1 if npairs == 6000 else 2)
    want = ol.oracle_align(inp, *pairs, 7)
    a = T.aligner(inp, 7, (-1, -1))
    a.align(*pairs)
    ms = []
    for _ in range(3):
        got = a.align(*pairs)
        ms.append(a.stats()["kernel_ms"])
    st = a.stats()
    t0 = time.perf_counter(); a.align(*pairs); e2e_ms = (time.perf_counter() - t0) * 1e3
    a.close()
    row = {"batch": name, "pairs": len(pairs[0]), "b200_shape": [st["lanes"], st["cells_per_lane"]], "b200_kernel_ms": sorted(ms)[1], "b200_call_ms_host_buffers": e2e_ms,
           "b200_identical_to_oracle": float((got[:, :6] == want).all(axis=1).mean())}
    if ol.have_logan():
        ol.logan_align(inp, *pairs, 7)
        lg, sec = ol.logan_align(inp, *pairs, 7)
        row.update({"logan_extendSeedL_ms": sec * 1e3, "logan_identical_to_oracle": float((lg == want).all(axis=1).mean()),
                    "speedup_kernel": sec * 1e3 / sorted(ms)[1], "speedup_call": sec * 1e3 / e2e_ms})
    print(row, flush=True)
    out.append(row)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
