"""Timing / profiling driver of the X-drop row: one batch of candidate pairs through several register shapes.
  python tools/xdrop_prof.py               timing of every shape (kernel ms by CUDA events) + parity against the oracle
  python tools/xdrop_prof.py --quick       the same for three shapes only
  ... --logan                              also runs the reference's CUDA aligner (oracle/_ref/libbella_logan.so) on the batch
  python tools/xdrop_prof.py --prof G T    one un-warmed batch with shape (G,T), for ncu"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402

import oracle_lib as ol  # noqa: E402
import test_xdrop_gpu as T  # noqa: E402
from bella_b200 import frontend as fe  # noqa: E402

prof = "--prof" in sys.argv
big = fe.synthetic(3000, 8000, seed=13)
pairs = T.candidate_pairs(big, 20000 if prof else 60000, seed=2)
n = len(pairs[0])
if prof:
    i = sys.argv.index("--prof")
    a = T.aligner(big, 7, (int(sys.argv[i + 1]), int(sys.argv[i + 2])))
    a.align(*pairs)
    print(a.stats())
    sys.exit(0)
t = time.time()
want = ol.oracle_align(big, *pairs, 7)
cpu_s = time.time() - t
out = []
plan = ((7, [(1, 64), (1, 32), (32, 1), (16, 2), (8, 4), (32, 2), (16, 4), (8, 8), (2, 64), (3, 64), (4, 64), (5, 64), (5, 32), (3, 32)]), (25, [(32, 2), (16, 4), (8, 8), (32, 4), (5, 128), (5, 256)]))
if "--quick" in sys.argv:
    plan = ((7, [(1, 64), (1, 32), (32, 2)]),)
for x, shapes in plan:
    w = want if x == 7 else ol.oracle_align(big, *pairs, x)
    for shape in shapes:
        try:
            a = T.aligner(big, x, shape)
            a.align(*pairs)
            ms = []
            for _ in range(3):
                got = a.align(*pairs)
                ms.append(a.stats()["kernel_ms"])
            st = a.stats()
            out.append({"xdrop": x, "shape": shape, "pairs": n, "kernel_ms": sorted(ms)[1], "wide": st["wide_extensions"],
                        "parity": bool(np.array_equal(got[:, :6], w)), "cpu_oracle_s_x7": cpu_s})
            a.close()
        except Exception as e:  # noqa: BLE001 -- a faulting shape must not hide the numbers of the others
            out.append({"xdrop": x, "shape": shape, "error": repr(e)[:200]})
        print(out[-1], flush=True)
        if "error" in out[-1]:
            break                                                 # the CUDA context may be gone
name = "xdrop_shapes_quick.json" if "--quick" in sys.argv else "xdrop_shapes.json"
json.dump(out, open(os.path.join(ROOT, "gpurun_out", name), "w"), indent=1)
if "--logan" in sys.argv and ol.have_logan():                     # the bar of SURVEY.md 8f: the reference kernel on the same box
    got, sec = ol.logan_align(big, *pairs, 7)
    out.append({"xdrop": 7, "shape": "LOGAN (reference, recompiled sm_100a)", "pairs": n, "extendSeedL_s": sec,
                "identical_to_oracle": float((got == want).all(axis=1).mean())})
    print(out[-1], flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", name), "w"), indent=1)
