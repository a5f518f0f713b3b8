"""CPU: the "next" row f3, reliable k-mer selection.  The oracle (oracle_reliable_occurrences: exact sort-and-count over
canonical k-mers) against the reference's own SplitCount (HyperLogLog -> Bloom -> cuckoo counts -> [l,u] filter,
include/kmercount.hpp:466-677) run on a FASTQ file of the same reads, followed by the reference's tuple emission
(src/main.cpp:393-416).  K-mer ids are arbitrary in the reference, so the id-free content is compared: every reliable
occurrence (read, pos) and the number of distinct reliable k-mers.  The host front end that feeds the SpGEMM tests and
bench.py (bella_b200/csrc/frontend.cpp) is then checked against the oracle.  No device code for this row yet."""
import copy

import numpy as np
import pytest

import oracle_lib as ol
from bella_b200 import frontend as fe

needs_ref = pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref/libbella_ref.so not built (needs /root/reference)")


def compare(inp, k, lower, upper, fastq):
    want = ol.ref_reliable_occurrences(inp, k, lower, upper, fastq)
    got = ol.oracle_reliable_occurrences(inp, k, lower, upper)
    assert got[2] == want[2] and got[2] > 1000
    np.testing.assert_array_equal(got[0], want[0])
    np.testing.assert_array_equal(got[1], want[1])


# one k per process: the reference's Kmer::set_k may be called with a single value (kmercode/Kmer.cpp:574-584)
@needs_ref
@pytest.mark.parametrize("lower,upper", [(2, 8), (2, 4), (3, 50)])
def test_oracle_selects_what_splitcount_selects(tmp_path, lower, upper):
    compare(fe.synthetic(500, 4000, seed=3 + upper), 17, lower, upper, str(tmp_path / "reads.fastq"))


@needs_ref
@pytest.mark.parametrize("k", [15, 21])
def test_other_kmer_lengths_in_their_own_process(tmp_path, k):
    import subprocess
    import sys
    code = (f"import sys; sys.path[:0] = {[ol.ROOT, ol.ROOT + '/tests']!r}\n"
            "import test_oracle_kmers as t\nfrom bella_b200 import frontend as fe\n"
            f"t.compare(fe.synthetic(400, 3000, seed=5), {k}, 2, 8, {str(tmp_path / 'reads.fastq')!r})\nprint('same')")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "same" in out.stdout, out.stderr[-2000:]


@needs_ref
def test_unusual_bases_follow_kmer_set_kmer(tmp_path):
    """no N check on this path: Kmer::set_kmer packs by two bits of the byte (N counts as G, lower case as upper case)"""
    inp = fe.synthetic(300, 3000, seed=8)
    dirty = copy.copy(inp)
    s = inp.seqs.copy()
    rng = np.random.default_rng(2)
    idx = rng.choice(len(s), len(s) // 60, replace=False)
    s[idx] = np.frombuffer(b"NnacgtRYK", dtype=np.uint8)[rng.integers(0, 9, len(idx))]
    dirty.seqs = s
    want = ol.ref_reliable_occurrences(dirty, 17, 2, 8, str(tmp_path / "reads.fastq"))
    got = ol.oracle_reliable_occurrences(dirty, 17, 2, 8)
    assert got[2] == want[2]
    np.testing.assert_array_equal(got[0], want[0])
    np.testing.assert_array_equal(got[1], want[1])


def test_host_front_end_emits_the_oracle_occurrences():
    """the tuples bella_fe_build emits (before MergeDuplicates) are exactly the reliable occurrences"""
    seqs, offs = fe.simulate_reads(400000, 800, 5000, 0.15, (0.10, 0.60, 0.30), 21)
    inp = fe.build_matrices(seqs, offs, 17, 2, 8, keep_tuples=True)
    t_kmer, t_read, t_pos = inp.tuples
    r, p, n_kmers = ol.oracle_reliable_occurrences(inp, 17, 2, 8)
    assert n_kmers == inp.n_kmers and len(r) == len(t_read)
    a = np.sort(t_read.astype(np.uint64) << np.uint64(16) | t_pos.astype(np.uint64))
    b = np.sort(r.astype(np.uint64) << np.uint64(16) | p.astype(np.uint64))
    np.testing.assert_array_equal(a, b)
    # one id per canonical k-mer: occurrences with the same id carry the same canonical k-mer and vice versa
    assert len(np.unique(t_kmer)) == n_kmers


# ---- minimizers (BASELINE.json configs[3], -w): the oracle's restatement against the reference's own functions ----
@needs_ref
@pytest.mark.parametrize("window", [10, 5, 32])
def test_oracle_minimizers_are_the_references(window):
    """getMinimizers (include/minimizer.hpp:49-77) under Kmer::rep().hash() (MurmurHash3_x64_128 of the packed canonical k-mer,
    seed 313): positions identical read by read, including the reference's quirk that the first `window` k-mers of a read are
    never sampled (its range test subtracts a size_t from an int) and reads with repeats (equal orders in one window)."""
    inp = fe.synthetic(60, 3000, seed=41)
    rng = np.random.default_rng(7)
    reads = [bytes(inp.seqs[int(inp.seq_off[r]):int(inp.seq_off[r + 1])]) for r in range(inp.n_reads)]
    reads += [b"ACGT" * 200, b"A" * 500, bytes(rng.choice(np.frombuffer(b"AC", dtype=np.uint8), 700)), b"ACGTACGTACGTACGTAC", b"ACGTACGTACGTACGTA", b"ACGT"]
    nonempty = 0
    for s in reads:
        want = ol.ref_minimizers(s, 17, window)
        got = ol.oracle_minimizers(s, 17, window)
        np.testing.assert_array_equal(got, want)
        nonempty += len(want) > 0
        if len(want):
            assert want.min() >= window                      # the quirk, pinned
    assert nonempty >= 60


@needs_ref
@pytest.mark.parametrize("window,lower,upper", [(10, 2, 8), (10, 2, 4), (5, 2, 8)])
def test_oracle_minimizer_selection_is_minimizercounts(tmp_path, window, lower, upper):
    """MinimizerCount (include/kmercount.hpp:690-832) + the minimizer branch of the tuple emission (src/main.cpp:363-388)"""
    inp = fe.synthetic(500, 4000, seed=11 + window)
    want = ol.ref_minimizer_occurrences(inp, 17, window, lower, upper, str(tmp_path / "reads.fastq"))
    got = ol.oracle_minimizer_occurrences(inp, 17, window, lower, upper)
    assert got[2] == want[2] and got[2] > 300
    np.testing.assert_array_equal(got[0], want[0])
    np.testing.assert_array_equal(got[1], want[1])


@pytest.mark.parametrize("window", [10, 5])
def test_front_end_minimizer_mode_emits_the_oracles_occurrences(window):
    """the product's host front end (bella_b200/csrc/frontend.cpp, window > 0) against the oracle's minimizer selection (which
    is pinned against the reference above): same reliable k-mer count, and B holds exactly the selected (read, position)s"""
    inp = fe.synthetic(500, 4000, seed=23, window=window)
    reads, pos, nk = ol.oracle_minimizer_occurrences(inp, 17, window, 2, 8)
    assert inp.n_kmers == nk and nk > 300
    cols = np.repeat(np.arange(inp.n_reads, dtype=np.uint32), np.diff(inp.B_colptr.astype(np.int64)))
    got = np.unique(np.stack([cols.astype(np.int64), inp.B_values.astype(np.int64)], axis=1), axis=0)
    # a k-mer sampled twice in one read keeps its last position in B (MergeDuplicates): compare as sets of (read, k-mer) through positions
    want = np.unique(np.stack([reads.astype(np.int64), pos.astype(np.int64)], axis=1), axis=0)
    assert len(got) <= len(want) and len(want) - len(got) <= 0.01 * len(want)
    assert set(map(tuple, got)) <= set(map(tuple, want))
    full = fe.synthetic(500, 4000, seed=23)
    assert 0.1 < inp.nnz / full.nnz < 0.35                  # -w 10 keeps ~15 % of the nonzeros on E. coli-sim
