"""GPU tests of the rows next to the hot path: the packed X-drop thread kernels (f1), reliable k-mer selection and tuple
emission on the device (f3), and the whole reads -> overlaps chain.  All of them ran green on a B200 at the start of round 2
(gpurun_out/first_unvalidated_tests.txt: 18 of 18), so they are ordinary tests now: a regression turns the suite red."""
import numpy as np
import pytest

import oracle_lib as ol
from bella_b200 import frontend as fe

pytestmark = [pytest.mark.gpu]


def candidate_pairs(inp, limit, seed=0):
    r = ol.oracle_spgemm(inp, want_aux=False)
    cols = np.repeat(np.arange(inp.n_reads, dtype=np.uint32), np.diff(r.colptrC.astype(np.int64)))
    idx = np.sort(np.random.default_rng(seed).choice(r.nnz, min(limit, r.nnz), replace=False))
    return r.rowids[idx], cols[idx], r.posH[idx], r.posV[idx]


# ---- f1: the packed-word thread kernels (2, W), (4, W) and the longest-first job order (3, W), (5, W) ------------------
@pytest.mark.parametrize("lanes,cells,xdrop", [(2, 64, 7), (3, 64, 7), (2, 32, 7), (3, 32, 3), (4, 64, 7), (5, 64, 7), (4, 32, 7), (5, 32, 3), (5, 128, 25), (4, 256, 60)])
def test_xdrop_unmeasured_shapes_match_oracle(lanes, cells, xdrop):
    from bella_b200 import xdrop as xd
    inp = fe.synthetic(400, 3000, seed=101)
    pairs = candidate_pairs(inp, 6000, seed=1)
    a = xd.XdropAligner(0)
    a.set_reads(inp.seqs, inp.seq_off)
    a.set_params(inp.kmer_size, xdrop, 0.55, 0.1, -1)
    a.set_shape(lanes, cells)
    got = a.align(*pairs)
    a.close()
    np.testing.assert_array_equal(got, ol.oracle_align_post(inp, *pairs, xdrop, 0.55, 0.1, -1))


# ---- f3: reliable k-mer selection + tuple emission on the device ------------------------------------------------------
def count_on_device(inp, k, lower, upper):
    from bella_b200 import kmers
    c = kmers.KmerCounter(0)
    out = c.count(inp.seqs, inp.seq_off, k, lower, upper)
    out["stats"] = c.stats()
    c.close()
    return out


@pytest.mark.parametrize("k,lower,upper", [(17, 2, 8), (15, 2, 4), (32, 2, 8)])
def test_kmers_device_selects_the_oracle_occurrences(k, lower, upper):
    inp = fe.synthetic(2000, 5000, seed=7)
    got = count_on_device(inp, k, lower, upper)
    r, p, n_kmers = ol.oracle_reliable_occurrences(inp, k, lower, upper)
    assert got["n_kmers"] == n_kmers
    np.testing.assert_array_equal(got["t_read"], r)
    np.testing.assert_array_equal(got["t_pos"], p)
    assert got["t_kmer"].max() == n_kmers - 1 and got["stats"]["launches"] == 4


def test_kmers_device_feeds_the_matrix_construction_and_the_spgemm():
    """reads -> tuples (f3) -> B on the device (f2) -> overlap SpGEMM: same C as from the host front end's matrices"""
    from bella_b200 import spgemm
    seqs, offs = fe.simulate_reads(400000, 800, 5000, 0.15, (0.10, 0.60, 0.30), 21)
    inp = fe.build_matrices(seqs, offs, 17, 2, 8)
    t = count_on_device(inp, 17, 2, 8)
    assert t["n_kmers"] == inp.n_kmers
    want = ol.oracle_spgemm(inp, want_aux=False)
    g = spgemm.OverlapSpGEMM(0)
    strand = np.concatenate([t["t_strand"], np.zeros(8, np.uint8)])
    g.set_inputs_tuples(inp.n_kmers, inp.n_reads, t["t_kmer"], t["t_read"], t["t_pos"], strand, inp.read_len, inp.kmer_size, inp.bin_size)
    flops, flopC, colptrC = g.symbolic()
    np.testing.assert_array_equal(colptrC, want.colptrC)        # the pattern of C does not depend on the k-mer ids
    g.close()


# ---- the whole chain: reads -> k-mers (f3) -> matrices (f2) -> SpGEMM -> alignment (f1) -> lines (f4) -------------------
def test_reads_to_overlaps_equals_the_oracle_chain():
    import pipeline_util as pu
    from bella_b200 import pipeline
    seqs, offs = fe.simulate_reads(400000, 800, 5000, 0.15, (0.10, 0.60, 0.30), 21)
    want = pu.oracle_chain(seqs, offs, ratiophi=0.55)
    got = pipeline.overlap_reads(seqs, offs, ratiophi=0.55)
    np.testing.assert_array_equal(got["colptrC"], want["C"].colptrC)
    for name, ref in (("rows", want["C"].rowids), ("count", want["C"].count), ("posH", want["C"].posH), ("posV", want["C"].posV)):
        np.testing.assert_array_equal(got[name], ref, err_msg=name)
    np.testing.assert_array_equal(got["out8"], want["out8"])
    assert got["lines"] == want["lines"] and len(got["lines"]) > 1000
