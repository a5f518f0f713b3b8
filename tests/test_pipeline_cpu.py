"""CPU: the oracle side of the whole chain (tests/pipeline_util.py) is consistent with the pieces that are pinned one by
one: tuples with rank ids rebuild the front end's matrices up to the renaming of k-mer ids, and the chain finds the same
overlapping pairs as the front-end path (the pattern of C does not depend on the ids; seeds and counts may, through the
fold order)."""
import numpy as np

import oracle_lib as ol
import pipeline_util as pu
from bella_b200 import frontend as fe


def test_matrices_from_front_end_tuples_are_the_front_end_matrices():
    seqs, offs = fe.simulate_reads(200000, 400, 4000, 0.15, (0.10, 0.60, 0.30), 9)
    inp = fe.build_matrices(seqs, offs, 17, 2, 8, keep_tuples=True)
    tk, tr, tp = inp.tuples
    order = np.lexsort((tp, tr))                                # read-major, position order (src/main.cpp:393-416)
    tk, tr, tp = tk[order], tr[order], tp[order]
    code = pu.kmer_codes()
    g = offs[tr].astype(np.int64) + tp.astype(np.int64)
    w = code[seqs[g[:, None] + np.arange(17)[None, :]]].astype(np.int64)
    st = np.array([tuple(a) <= tuple(3 - a[::-1]) for a in w], dtype=np.uint8)
    got = pu.inputs_from_tuples(inp.n_kmers, inp.n_reads, tk, tr, tp, st, seqs, offs)
    for name in ("B_colptr", "B_rowids", "B_values", "A_colptr", "A_rowids", "A_values", "read_len"):
        np.testing.assert_array_equal(getattr(got, name), getattr(inp, name), err_msg=name)
    for name in ("A_strand", "B_strand"):
        a, b = (np.unpackbits(getattr(x, name), bitorder="little")[:inp.nnz] for x in (got, inp))
        np.testing.assert_array_equal(a, b, err_msg=name)


def test_chain_finds_the_pairs_of_the_front_end_path():
    seqs, offs = fe.simulate_reads(150000, 300, 3000, 0.15, (0.10, 0.60, 0.30), 4)
    ch = pu.oracle_chain(seqs, offs)
    inp = fe.build_matrices(seqs, offs, 17, 2, 8)
    want = ol.oracle_spgemm(inp, want_aux=False)
    np.testing.assert_array_equal(ch["C"].colptrC, want.colptrC)
    np.testing.assert_array_equal(ch["C"].rowids, want.rowids)
    assert ch["inp"].nnz == inp.nnz and ch["inp"].n_kmers == inp.n_kmers
    assert 0 < len(ch["lines"]) <= want.nnz and (ch["out8"][:, 0] >= 17 - 2 * 7).all()     # each side ends within x of its best


def test_reads_with_n_or_lower_case_are_refused_before_any_device_call():
    """the k-mer code maps N -> G and lower case -> upper case, the reference compares raw substrings
    (chain.hpp:35-44): such reads must fail loudly, not diverge silently.  No GPU is needed to be refused."""
    import pytest
    from bella_b200 import pipeline
    seqs = np.frombuffer(b"ACGTACGTACGTACGTACGTNACGTACGTACGTACGTACGTACGTacgtACGT", dtype=np.uint8)
    off = np.array([0, 27, len(seqs)], dtype=np.uint64)
    with pytest.raises(ValueError, match="read 0 has the byte 'N'"):
        pipeline.overlap_reads(seqs, off)
    with pytest.raises(ValueError, match="read 1 has the byte 'a'"):
        pipeline.overlap_reads(seqs[27:], np.array([0, 0, len(seqs) - 27], dtype=np.uint64))     # read 0 is empty, read 1 holds the lower case
