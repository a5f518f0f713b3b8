"""CPU: the oracle of the "next" row f1 (gapped X-drop seed-and-extend, oracle/bella_oracle.c
oracle_xdrop_align) against the reference's own alignSeqAn -> seqan::extendSeed(GappedXDrop)
(oracle/_ref, built from /root/reference) and against the committed fixture generated from it.
There is no device implementation of this row yet: the oracle comes first."""
import os

import numpy as np
import pytest

import golden_util
import oracle_lib as ol

GOLD = os.path.join(golden_util.GOLDEN, "xdrop.npz")


def candidate_pairs(inp, limit=None, seed=0):
    """the seeds the overlap SpGEMM hands to the aligner: (row, col, posH, posV) of every output nonzero"""
    r = ol.oracle_spgemm(inp, want_aux=False)
    cols = np.repeat(np.arange(inp.n_reads, dtype=np.uint32), np.diff(r.colptrC.astype(np.int64)))
    idx = np.arange(r.nnz)
    if limit and r.nnz > limit:
        idx = np.sort(np.random.default_rng(seed).choice(r.nnz, limit, replace=False))
    return r.rowids[idx], cols[idx], r.posH[idx], r.posV[idx]


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref/libbella_ref.so not built (needs /root/reference)")
@pytest.mark.parametrize("xdrop", [7, 3, 25])
def test_xdrop_oracle_matches_reference(small_inputs, xdrop):
    rows, cols, pH, pV = candidate_pairs(small_inputs, limit=4000, seed=xdrop)
    want = ol.ref_align(small_inputs, rows, cols, pH, pV, xdrop)
    got = ol.oracle_align(small_inputs, rows, cols, pH, pV, xdrop)
    np.testing.assert_array_equal(got, want)
    assert (want[:, 1] == ord("c")).any() and (want[:, 1] == ord("n")).any()
    assert want[:, 0].max() > 500                       # real overlaps extend over hundreds of bases


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref/libbella_ref.so not built (needs /root/reference)")
def test_xdrop_oracle_matches_reference_low_error_and_edges():
    from bella_b200 import frontend as fe
    inp = fe.synthetic(150, 3000, coverage=20.0, err=0.02, seed=77, hi=40)       # long extensions that run into the read ends
    rows, cols, pH, pV = candidate_pairs(inp, limit=1500)
    np.testing.assert_array_equal(ol.oracle_align(inp, rows, cols, pH, pV, 7), ol.ref_align(inp, rows, cols, pH, pV, 7))
    # seeds at the very start / end of a read (empty prefix or suffix)
    k = inp.kmer_size
    n = 40
    r = np.arange(1, n + 1, dtype=np.uint32)
    c = np.zeros(n, dtype=np.uint32)
    lens = inp.read_len
    pH = np.where(np.arange(n) % 2 == 0, 0, lens[r] - k).astype(np.uint16)
    pV = np.where(np.arange(n) % 3 == 0, 0, lens[c] - k).astype(np.uint16)
    np.testing.assert_array_equal(ol.oracle_align(inp, r, c, pH, pV, 7), ol.ref_align(inp, r, c, pH, pV, 7))


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref/libbella_ref.so not built (needs /root/reference)")
def test_xdrop_oracle_follows_seqan_dna5_on_unusual_bases(small_inputs):
    """alignSeqAn converts the reads to seqan::Dna5String: lower case == upper case, U == T, everything else is N and N matches N"""
    import copy
    rows, cols, pH, pV = candidate_pairs(small_inputs, limit=3000, seed=11)
    dirty = copy.copy(small_inputs)
    s = small_inputs.seqs.copy()
    rng = np.random.default_rng(5)
    idx = rng.choice(len(s), len(s) // 50, replace=False)
    s[idx] = np.frombuffer(b"NnacgtRUuYx-", dtype=np.uint8)[rng.integers(0, 12, len(idx))]
    dirty.seqs = s
    want = ol.ref_align(dirty, rows, cols, pH, pV, 7)
    np.testing.assert_array_equal(ol.oracle_align(dirty, rows, cols, pH, pV, 7), want)
    assert (want != ol.ref_align(small_inputs, rows, cols, pH, pV, 7)).any()


def test_xdrop_oracle_reproduces_reference_golden():
    z = np.load(GOLD)
    from bella_b200.frontend import OverlapInputs
    inp = OverlapInputs(n_reads=int(z["n_reads"]), n_kmers=0, nnz=0, A_colptr=None, A_rowids=None, A_values=None, A_strand=None,
                        B_colptr=None, B_rowids=None, B_values=None, B_strand=None, read_len=None, kmer_size=int(z["k"]),
                        seqs=z["seqs"], seq_off=z["seq_off"])
    got = ol.oracle_align(inp, z["rows"], z["cols"], z["posH"], z["posV"], int(z["xdrop"]))
    np.testing.assert_array_equal(got, z["ref_out"])
