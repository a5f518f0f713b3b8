"""CPU: the result checks bench.py relies on ("tuples compared ... must be bit-identical")."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle_lib as ol
from bella_b200 import frontend as fe
from bella_b200.checks import tuple_checksum


@pytest.fixture(scope="module")
def result():
    inp = fe.synthetic(1200, 4000, seed=9)
    return inp, ol.oracle_spgemm(inp, want_aux=False)


def test_checksums_of_column_ranges_add_up_to_the_whole(result):
    """the multi-GPU bench sums the per-rank checksums (mod 2^64) and compares with the single-GPU result"""
    inp, r = result
    whole, z = tuple_checksum(0, r.colptrC, (r.rowids, r.count, r.posH, r.posV))
    assert z == r.nnz
    cuts = [0, 100, 101, 700, inp.n_reads]
    tot, zt = 0, 0
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        z0, z1 = int(r.colptrC[lo]), int(r.colptrC[hi])
        c, zz = tuple_checksum(lo, r.colptrC[lo:hi + 1].astype(np.int64) - z0, (r.rowids[z0:z1], r.count[z0:z1], r.posH[z0:z1], r.posV[z0:z1]))
        tot = (tot + c) % (1 << 64)
        zt += zz
    assert (tot, zt) == (whole, z)


def test_checksum_sees_a_single_changed_field_and_a_moved_tuple(result):
    inp, r = result
    whole, _ = tuple_checksum(0, r.colptrC, (r.rowids, r.count, r.posH, r.posV))
    for name in ("rowids", "count", "posH", "posV"):
        a = {k: getattr(r, k).copy() for k in ("rowids", "count", "posH", "posV")}
        a[name][r.nnz // 2] ^= 1
        assert tuple_checksum(0, r.colptrC, (a["rowids"], a["count"], a["posH"], a["posV"]))[0] != whole
    cp = r.colptrC.copy()
    j = int(np.flatnonzero(np.diff(cp.astype(np.int64)) > 0)[3])
    cp[j + 1] -= 1                                             # the last tuple of column j now belongs to column j + 1
    assert tuple_checksum(0, cp, (r.rowids, r.count, r.posH, r.posV))[0] != whole


def test_bench_parity_check_aborts_on_a_difference(result):
    import bench
    inp, r = result
    n = 900
    z = int(r.colptrC[n])
    good = (r.rowids.copy(), r.count.copy(), r.posH.copy(), r.posV.copy())
    assert bench.check_parity(r.colptrC, good, r, n) == z
    bad = tuple(a.copy() for a in good)
    bad[2][z - 1] ^= 1
    with pytest.raises(SystemExit, match="PARITY FAILURE: posH"):
        bench.check_parity(r.colptrC, bad, r, n)
    cp = r.colptrC.copy(); cp[5] += 1
    with pytest.raises(SystemExit, match="PARITY FAILURE: colptrC"):
        bench.check_parity(cp, good, r, n)


def test_both_arms_print_the_same_config():
    import bench
    for c, w in bench.CONFIGS.items():
        assert isinstance(bench.workload_name(w), str) and str(w["k"]) in bench.workload_name(w)
    assert bench.workload_name(bench.CONFIGS[2]) == "synthetic 50000 PacBio reads x 10000 bp, e=0.15, k=17, [l,u]=[2,8], 30x, seed 2"
