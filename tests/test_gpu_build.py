"""GPU: matrix construction on the device ("next" row f2): tuples -> B must reproduce, array for array, what the
reference's CSC constructor + MergeDuplicates produced (tests/golden/build_csc.npz, generated from the unmodified
reference), and the SpGEMM that follows must give the same result as the one fed with the host-built matrices."""
import os

import numpy as np
import pytest

import golden_util
import oracle_lib as ol

pytestmark = pytest.mark.gpu


def tuple_strands(inp, tk, tr, tp):
    """strand bit of every tuple, from the read sequence: window <= its reverse complement (any consistent
    canonical form gives the same oriented/not-oriented decisions)"""
    comp = np.zeros(256, dtype=np.uint8)
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    k = inp.kmer_size
    out = np.zeros(len(tk), dtype=np.uint8)
    for t in range(len(tk)):
        w = inp.seqs[int(inp.seq_off[tr[t]]) + int(tp[t]):int(inp.seq_off[tr[t]]) + int(tp[t]) + k]
        rc = comp[w[::-1]]
        out[t] = 1 if w.tobytes() <= rc.tobytes() else 0
    return out


def test_device_build_matches_reference_golden():
    from bella_b200 import spgemm
    z = np.load(os.path.join(golden_util.GOLDEN, "build_csc.npz"))
    m, n = int(z["n_kmers"]), int(z["n_reads"])
    tk, tr, tp = z["t_kmer"], z["t_read"], z["t_pos"]
    strand = np.packbits(np.zeros(len(tk), dtype=np.uint8), bitorder="little")
    g = spgemm.OverlapSpGEMM(0)
    g.set_inputs_tuples(m, n, tk, tr, tp, np.concatenate([strand, np.zeros(8, np.uint8)]), np.full(n, 2000, dtype=np.uint32))
    colptr, rows, vals, _, ms = g.get_B()
    g.close()
    np.testing.assert_array_equal(colptr, z["B_colptr"])
    np.testing.assert_array_equal(rows, z["B_rowids"])
    np.testing.assert_array_equal(vals, z["B_values"])


def test_device_build_then_spgemm_equals_host_built_path():
    from bella_b200 import frontend as fe, spgemm
    inp = fe.synthetic(1500, 4000, seed=23, keep_tuples=True)
    tk, tr, tp = inp.tuples
    # the front end emits tuples k-mer major; BELLA emits them read by read in position order (src/main.cpp:393-416)
    order = np.lexsort((tp, tr))
    tk, tr, tp = tk[order], tr[order], tp[order]
    st = tuple_strands(inp, tk, tr, tp)
    g = spgemm.OverlapSpGEMM(0)
    g.set_inputs_tuples(inp.n_kmers, inp.n_reads, tk, tr, tp, np.concatenate([np.packbits(st, bitorder="little"), np.zeros(8, np.uint8)]),
                        inp.read_len, inp.kmer_size, inp.bin_size)
    colptr, rows, vals, strand, ms = g.get_B()
    np.testing.assert_array_equal(colptr, inp.B_colptr)
    np.testing.assert_array_equal(rows, inp.B_rowids)
    np.testing.assert_array_equal(vals, inp.B_values)
    np.testing.assert_array_equal(np.unpackbits(strand, bitorder="little")[:inp.nnz], np.unpackbits(inp.B_strand, bitorder="little")[:inp.nnz])
    flops, flopC, colptrC = g.symbolic()
    r = g.numeric(aux=True)
    g.close()
    ol.assert_same(ol.Result(flopC, colptrC, *r), ol.oracle_spgemm(inp))


def test_device_build_refuses_scattered_tuples():
    from bella_b200 import spgemm
    g = spgemm.OverlapSpGEMM(0)
    tk = np.array([5, 6, 7, 8], dtype=np.uint32)
    tr = np.array([0, 1, 0, 1], dtype=np.uint32)          # read 0's tuples are not contiguous
    tp = np.array([1, 2, 3, 4], dtype=np.uint16)
    with pytest.raises(spgemm.BellaB200Error):
        g.set_inputs_tuples(10, 2, tk, tr, tp, np.zeros(9, np.uint8), np.full(2, 100, dtype=np.uint32))
    g.close()
