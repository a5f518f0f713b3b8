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


def test_gpu_heavy_columns_row_range_units():
    # low-error reads with a permissive multiplicity filter: thousands of products per column and hundreds per
    # pair -> columns are split into row-range units, long pairs fold on warps / the whole CTA
    from bella_b200 import frontend as fe
    inp = fe.synthetic(700, 5000, coverage=40.0, err=0.01, seed=13, hi=80)
    want = ol.oracle_spgemm(inp)
    assert int(want.flopC.max()) > 3 * 8192
    ol.assert_same(gpu_result(inp), want)


def test_gpu_single_pair_larger_than_shared_memory():
    # a pair sharing more than 8192 k-mers: the unit is one row and is folded from global memory
    from bella_b200 import frontend as fe
    inp = fe.synthetic(24, 14000, coverage=8.0, err=0.002, seed=17, hi=40)
    want = ol.oracle_spgemm(inp)
    per_pair = want.flopC.astype(np.int64).sum() / max(want.nnz, 1)
    assert per_pair > 2000
    ol.assert_same(gpu_result(inp), want)


def test_gpu_positions_near_the_u16_limit(small_inputs):
    # positions above 65535-k cannot use the packed 16-bit far test: the kernel must switch to the exact one.
    # (Such positions are not valid k-mer starts; the oracle applies the same arithmetic to them.)
    import copy
    inp = copy.copy(small_inputs)
    rng = np.random.default_rng(1)
    Bv = inp.B_values.copy()
    idx = rng.choice(inp.nnz, inp.nnz // 50, replace=False)
    Bv[idx] = rng.integers(65400, 65536, idx.size).astype(np.uint16)
    inp.B_values = Bv
    # the device derives A from B, but the oracle reads A's values: keep A == B^T by matching (k-mer, read) keys
    reads_of = np.repeat(np.arange(inp.n_reads, dtype=np.int64), np.diff(inp.B_colptr.astype(np.int64)))
    kmers_of = np.repeat(np.arange(inp.n_kmers, dtype=np.int64), np.diff(inp.A_colptr.astype(np.int64)))
    keyA = kmers_of * inp.n_reads + inp.A_rowids.astype(np.int64)
    keyB = inp.B_rowids.astype(np.int64) * inp.n_reads + reads_of
    Av = np.empty_like(Bv)
    Av[np.argsort(keyA, kind="stable")] = Bv[np.argsort(keyB, kind="stable")]
    inp.A_values = Av
    ol.assert_same(gpu_result(inp), ol.oracle_spgemm(inp))


def test_gpu_panel_format_strand_in_rowids(small_inputs):
    # the multi-GPU exchange format: strand bit in bit 31 of B.rowids, strand_B = NULL, device-resident inputs
    import torch
    from bella_b200 import spgemm
    inp = small_inputs
    want = ol.oracle_spgemm(inp)
    strand = np.unpackbits(inp.B_strand, bitorder="little")[:inp.nnz].astype(np.uint32)
    dev = torch.device("cuda", 0)
    t = lambda a, dt: torch.from_numpy(a.view(dt)).to(dev)
    colptr, rows = t(inp.B_colptr, np.int32), t(inp.B_rowids | (strand << 31), np.int32)
    vals, lens = t(inp.B_values, np.int16), t(inp.read_len, np.int32)
    g = spgemm.OverlapSpGEMM(0)
    g.set_inputs_device(inp.n_reads, inp.n_kmers, inp.nnz, (colptr, rows, vals), lens, None, inp.kmer_size, inp.bin_size)
    flops, flopC, colptrC = g.symbolic()
    r = g.numeric(aux=True)
    g.close()
    ol.assert_same(ol.Result(flopC, colptrC, *r), want)


def test_gpu_ragged_edge_inputs():
    # one read; reads without any reliable k-mer between overlapping ones (empty columns of B, rows that never occur)
    from bella_b200 import frontend as fe
    rng = np.random.default_rng(5)
    bases = np.frombuffer(b"ACGT", dtype=np.uint8)
    genome = rng.integers(0, 4, 6000)
    rd = lambda a, b: bases[genome[a:b]].tobytes().decode()
    lone = bases[rng.integers(0, 4, 900)].tobytes().decode()
    cases = [[rd(0, 3000)],                                                     # n = 1: nothing to overlap with
             [rd(0, 3000), lone, rd(1000, 4000), lone[::-1], rd(2000, 5000), rd(100, 140)]]   # unrelated and very short reads in between
    for reads in cases:
        s, o = fe.reads_from_strings(reads)
        inp = fe.build_matrices(s, o)
        want = ol.oracle_spgemm(inp) if inp.nnz else None
        got = gpu_result(inp)
        if want is None:
            assert int(got.colptrC[-1]) == 0
        else:
            ol.assert_same(got, want)


def test_gpu_many_reads_two_level_row_bitmap():
    # more than 65 536 reads: a light column's rows span more 32-row words than the direct bitmap of the
    # smallest class holds, so the pair index goes through the two-level bitmap
    from bella_b200 import frontend as fe
    inp = fe.synthetic(70000, 1000, coverage=20.0, seed=41)
    assert inp.n_reads > 65536 + 2048
    ol.assert_same(gpu_result(inp), ol.oracle_spgemm(inp))


def test_gpu_csr_surface_and_unpinned_count(small_inputs):
    """a14: A handed over row-major (bella_csr_view; CSR.h:15-67) is B's CSC reinterpreted -- same results; and the count
    of pairs with more than 16 bins (choose()'s unpinned tie order, common.h:162-170) equals the oracle's."""
    from bella_b200 import spgemm
    want = ol.oracle_spgemm(small_inputs)
    g = spgemm.OverlapSpGEMM(0)
    try:
        g.set_inputs_csr(small_inputs)
        flops, flopC, colptrC = g.symbolic()
        rows, cnt, pH, pV, aux = g.numeric(aux=True)
        got = ol.Result(flopC, colptrC, rows, cnt, pH, pV, aux)
        ol.assert_same(got, want)
        assert g.n_unpinned() == want.unpinned == int((want.aux[:, 0] > 16).sum())
    finally:
        g.close()


def test_gpu_one_based_csr_is_refused(small_inputs):
    import ctypes
    from bella_b200 import spgemm
    inp = small_inputs
    g = spgemm.OverlapSpGEMM(0)
    try:
        v = spgemm.CsrView(inp.n_reads, inp.n_kmers, inp.nnz, inp.B_colptr.ctypes.data, inp.B_rowids.ctypes.data, inp.B_values.ctypes.data, 0)
        rc = spgemm.lib().bella_b200_set_inputs_csr(g._h, ctypes.byref(v), inp.read_len.ctypes.data, inp.B_strand.ctypes.data, 17, 500)
        assert rc == -1
    finally:
        g.close()


@pytest.mark.parametrize("nrange", ["2", "4"])
def test_gpu_scatter_group_pipeline_with_heavy_columns(monkeypatch, nrange):
    """the scatter | group + fold pipeline over column ranges (on by default only for large inputs) on inputs whose heavy
    columns are cut into row-range units: a column's units must all be folded after the scatter pass that covers the column"""
    from bella_b200 import frontend as fe
    monkeypatch.setenv("BELLA_B200_NRANGE", nrange)
    inp, ref = golden_util.load("heavy_units")
    ol.assert_same(gpu_result(inp, False), ref)
    hifi = fe.synthetic(n_reads=700, read_len=5000, coverage=40.0, err=0.01, seed=13, hi=80)
    ol.assert_same(gpu_result(hifi, False), ol.oracle_spgemm(hifi))
    clr = fe.synthetic(n_reads=2500, read_len=6000, seed=77)
    ol.assert_same(gpu_result(clr, False), ol.oracle_spgemm(clr))


def test_gpu_minimizer_sparsified_matrix():
    """BASELINE.json configs[3]: the read x k-mer matrix of BELLA's minimizer mode (-w 10; only the sparsity pattern changes,
    include/minimizer.hpp:49-77 through the host front end) -- the SpGEMM must not care"""
    from bella_b200 import frontend as fe
    inp = fe.synthetic(n_reads=6000, read_len=8000, err=0.20, split=(0.35, 0.25, 0.40), seed=4, window=10)
    assert inp.nnz > 100000
    ol.assert_same(gpu_result(inp, False), ol.oracle_spgemm(inp))


@pytest.mark.parametrize("nrange", ["1", "3"])
def test_gpu_streamed_output_buffers(monkeypatch, medium_inputs, nrange):
    """bella_b200_set_output_buffers: the results of a column range are copied into the caller's page-locked buffers while the
    later ranges still fold; numeric() with the same buffers has nothing left to do.  Same bytes as the plain sequence; a
    capacity that is too small falls back to the plain copy."""
    from bella_b200 import spgemm
    monkeypatch.setenv("BELLA_B200_NRANGE", nrange)
    want = ol.oracle_spgemm(medium_inputs)
    hifi, href = golden_util.load("heavy_units")
    for inp, ref, cap in ((medium_inputs, want, want.nnz + 100), (hifi, href, href.nnz + 7), (medium_inputs, want, 10)):
        g = spgemm.OverlapSpGEMM(0)
        try:
            g.set_output_buffers(cap)
            for _ in range(2):                              # twice: the buffers are reused
                g.set_inputs(inp)
                flops, flopC, colptrC = g.symbolic(pinned=True)
                rows, cnt, pH, pV = g.numeric(pinned=True)
                np.testing.assert_array_equal(colptrC, ref.colptrC)
                np.testing.assert_array_equal(rows, ref.rowids)
                np.testing.assert_array_equal(cnt, ref.count)
                np.testing.assert_array_equal(pH, ref.posH)
                np.testing.assert_array_equal(pV, ref.posV)
            a = g.numeric(aux=True)[4]                      # the device arrays are complete too
            np.testing.assert_array_equal(a, ref.aux)
        finally:
            g.close()
