"""One-shot GPU check of the X-drop row (f1) that avoids pytest/torch start-up: runs the functions of
tests/test_xdrop_gpu.py directly, then times one larger batch per shape and the CPU oracle beside it.
Usage (GPU box):  python tools/xdrop_check.py [out.json]"""
import json
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402

import oracle_lib as ol  # noqa: E402
import test_xdrop_gpu as T  # noqa: E402
from bella_b200 import frontend as fe  # noqa: E402

out = {"tests": {}, "perf": []}
t0 = time.time()
inp = fe.synthetic(400, 3000, seed=101)
reads = (inp, T.candidate_pairs(inp, 6000, seed=1))


def run(name, fn, *a):
    t = time.time()
    try:
        fn(*a)
        out["tests"][name] = "ok"
    except Exception as e:  # noqa: BLE001
        out["tests"][name] = "FAIL: " + repr(e)[:300]
        traceback.print_exc()
    print(f"{name}: {out['tests'][name][:80]}  ({time.time() - t:.1f}s)", flush=True)


for lanes, cells, x in [(32, 1, 7), (-1, -1, 7), (32, 2, 15), (32, 4, 30), (16, 1, 3), (16, 2, 7), (0, 0, 7), (-1, -1, 25), (-1, -1, 120)]:
    run(f"matches_oracle[{lanes},{cells},x={x}]", T.test_xdrop_matches_oracle, reads, lanes, cells, x)
run("overflow_to_wide", T.test_window_overflow_goes_through_the_wide_kernel, reads)
run("low_error_and_edges", T.test_low_error_reads_and_seeds_at_the_read_ends)
run("golden_fixture", T.test_reference_golden_fixture)
run("bad_arguments", T.test_bad_arguments_are_refused, reads)

# timing: every output pair of a 3000-read x 8 kb batch (reads resident, pairs from the host)
try:
    big = fe.synthetic(3000, 8000, seed=13)
    pairs = T.candidate_pairs(big, 60000, seed=2)
    n = len(pairs[0])
    t = time.time()
    want = ol.oracle_align(big, *pairs, 7)
    cpu_s = time.time() - t
    span = (want[:, 3] - want[:, 2]).astype(np.int64) + (want[:, 5] - want[:, 4])       # ~ anti-diagonals per pair
    for shape in [(32, 1), (16, 2), (32, 2), (0, 0)]:
        a = T.aligner(big, 7, shape)
        a.align(*pairs)                                                                 # warm-up (allocations)
        t = time.time()
        got = a.align(*pairs)
        wall = time.time() - t
        st = a.stats()
        ok = bool(np.array_equal(got[:, :6], want))
        out["perf"].append({"shape": shape, "pairs": n, "kernel_ms": st["kernel_ms"], "e2e_ms": wall * 1e3, "wide": st["wide_extensions"],
                            "pairs_per_s_kernel": n / (st["kernel_ms"] * 1e-3), "antidiagonals": int(span.sum()), "parity": ok,
                            "cpu_oracle_s": cpu_s, "cpu_threads": os.cpu_count()})
        print(out["perf"][-1], flush=True)
        a.close()
except Exception as e:  # noqa: BLE001
    out["perf_error"] = repr(e)[:300]
    traceback.print_exc()
out["total_s"] = time.time() - t0
dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "xdrop_check.json")
os.makedirs(os.path.dirname(dst), exist_ok=True)
json.dump(out, open(dst, "w"), indent=1)
# last (needs torch, whose first import on a fresh box is slow): the SpGEMM result aligned on the device as it is
run("chained_behind_spgemm", T.test_chained_behind_the_overlap_spgemm_on_the_device, fe.synthetic(2000, 5000, seed=7))
out["total_s"] = time.time() - t0
json.dump(out, open(dst, "w"), indent=1)
print("ALL OK" if all(v == "ok" for v in out["tests"].values()) else "SOME FAILED", f"{out['total_s']:.1f}s")
