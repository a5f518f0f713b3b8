"""GPU parity: the CUDA path through the C-ABI (include/bella_b200.h) against the CPU oracle and the
committed reference goldens -- bit-exact (integer work)."""
import numpy as np
import pytest

import golden_util
import oracle_lib as ol

pytestmark = pytest.mark.gpu


def gpu_result(inp, with_A=True):
    from bella_b200 import spgemm
    r = spgemm.overlap_spgemm(inp, with_A=with_A, aux=True)
    assert r["flops"] == int(r["flopC"].astype(np.uint64).sum())
    return ol.Result(r["flopC"], r["colptrC"], r["rowids"], r["count"], r["posH"], r["posV"], r["aux"])


@pytest.mark.parametrize("name", golden_util.SPGEMM_FIXTURES)
@pytest.mark.parametrize("with_A", [True, False])
def test_gpu_reproduces_reference_golden(name, with_A):
    inp, ref = golden_util.load(name)
    ol.assert_same(gpu_result(inp, with_A), ref)


def test_gpu_vs_oracle_small(small_inputs):
    ol.assert_same(gpu_result(small_inputs), ol.oracle_spgemm(small_inputs))


def test_gpu_vs_oracle_medium(medium_inputs):
    ol.assert_same(gpu_result(medium_inputs, with_A=False), ol.oracle_spgemm(medium_inputs))


def test_gpu_unsorted_A_columns(small_inputs):
    # the reference's Transpose() leaves A's columns in schedule-dependent order: results must not depend on it
    import copy
    inp = copy.copy(small_inputs)
    rng = np.random.default_rng(0)
    Ar, Av = inp.A_rowids.copy(), inp.A_values.copy()
    sA = np.unpackbits(inp.A_strand, bitorder="little")[:inp.nnz].copy()
    for c in rng.choice(inp.n_kmers, 20000, replace=False):
        s, e = inp.A_colptr[c], inp.A_colptr[c + 1]
        p = rng.permutation(e - s)
        Ar[s:e], Av[s:e], sA[s:e] = Ar[s:e][p], Av[s:e][p], sA[s:e][p]
    inp.A_rowids, inp.A_values = Ar, Av
    inp.A_strand = np.concatenate([np.packbits(sA, bitorder="little"), np.zeros(8, np.uint8)])
    ol.assert_same(gpu_result(inp), ol.oracle_spgemm(small_inputs))


def test_gpu_staged_numeric_and_column_range(small_inputs):
    from bella_b200 import spgemm
    want = ol.oracle_spgemm(small_inputs)
    g = spgemm.OverlapSpGEMM(0)
    g.set_inputs(small_inputs)
    flops, flopC, colptrC = g.symbolic()
    np.testing.assert_array_equal(colptrC, want.colptrC)
    # HashSpGEMM's stage loop (overlap.hpp:712-719): numeric on consecutive column ranges
    bounds = [0, 1, 700, 701, 1500, small_inputs.n_reads]
    rows = np.concatenate([g.numeric(a, b)[0] for a, b in zip(bounds[:-1], bounds[1:])])
    np.testing.assert_array_equal(rows, want.rowids)
    # row sharding: a handle restricted to [lo, hi)
    lo, hi = 300, 1100
    g.set_column_range(lo, hi)
    f2, flopC2, colptrC2 = g.symbolic()
    np.testing.assert_array_equal(flopC2, want.flopC[lo:hi])
    np.testing.assert_array_equal(colptrC2, want.colptrC[lo:hi + 1] - want.colptrC[lo])
    r, c, h, v = g.numeric()
    z0, z1 = int(want.colptrC[lo]), int(want.colptrC[hi])
    np.testing.assert_array_equal(r, want.rowids[z0:z1])
    np.testing.assert_array_equal(c, want.count[z0:z1])
    np.testing.assert_array_equal(h, want.posH[z0:z1])
    np.testing.assert_array_equal(v, want.posV[z0:z1])
    g.close()


def test_gpu_empty_and_degenerate():
    from bella_b200 import frontend as fe, spgemm
    # reads that share nothing: every k-mer is unique -> no reliable k-mers -> empty matrices
    inp = fe.synthetic(8, 400, coverage=0.01, seed=3)
    assert inp.nnz == 0
    r = spgemm.overlap_spgemm(inp, with_A=False)
    assert r["flops"] == 0 and int(r["colptrC"][-1]) == 0 and len(r["rowids"]) == 0


def test_gpu_errors_are_reported(small_inputs):
    from bella_b200 import spgemm
    g = spgemm.OverlapSpGEMM(0)
    with pytest.raises(spgemm.BellaB200Error):
        g.symbolic()            # no inputs yet
    g.set_inputs(small_inputs)
    with pytest.raises(spgemm.BellaB200Error):
        g.set_column_range(5, small_inputs.n_reads + 1)
    g.close()
