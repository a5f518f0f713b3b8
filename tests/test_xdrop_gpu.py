"""GPU parity of the "next" row f1 (batched gapped X-drop seed-and-extend) through the C-ABI of include/bella_xdrop.h:
bit-exact against the oracle (oracle_xdrop_align + oracle_xdrop_post, pinned against the reference's alignSeqAn and
PostAlignDecision in test_oracle_xdrop.py) and against the committed reference fixture."""
import os

import numpy as np
import pytest

import golden_util
import oracle_lib as ol
from bella_b200 import frontend as fe

pytestmark = pytest.mark.gpu


def candidate_pairs(inp, limit=None, seed=0):
    r = ol.oracle_spgemm(inp, want_aux=False)
    cols = np.repeat(np.arange(inp.n_reads, dtype=np.uint32), np.diff(r.colptrC.astype(np.int64)))
    idx = np.arange(r.nnz)
    if limit and r.nnz > limit:
        idx = np.sort(np.random.default_rng(seed).choice(r.nnz, limit, replace=False))
    return r.rowids[idx], cols[idx], r.posH[idx], r.posV[idx]


def aligner(inp, xdrop=7, shape=(-1, -1), ratiophi=0.55, delta=0.1, fixed_threshold=-1):
    from bella_b200 import xdrop as xd
    a = xd.XdropAligner(0)
    a.set_reads(inp.seqs, inp.seq_off)
    a.set_params(inp.kmer_size, xdrop, ratiophi, delta, fixed_threshold)
    a.set_shape(*shape)
    return a


@pytest.fixture(scope="module")
def reads():
    inp = fe.synthetic(400, 3000, seed=101)
    return inp, candidate_pairs(inp, 6000, seed=1)


# every instantiation the library ships; (0,0) = wide path only; (-1,-1) = chosen from xdrop
@pytest.mark.parametrize("lanes,cells,xdrop", [(-1, -1, 7), (1, 64, 7), (1, 32, 7), (32, 1, 7), (32, 2, 15), (32, 4, 30), (16, 1, 3), (16, 2, 7), (16, 4, 15), (8, 4, 7),
                                                (8, 8, 15), (0, 0, 7), (-1, -1, 25), (-1, -1, 120)])
def test_xdrop_matches_oracle(reads, lanes, cells, xdrop):
    inp, pairs = reads
    a = aligner(inp, xdrop, (lanes, cells))
    got = a.align(*pairs)
    want = ol.oracle_align_post(inp, *pairs, xdrop, 0.55, 0.1, -1)
    np.testing.assert_array_equal(got, want)
    st = a.stats()
    # main kernel + wide kernel + compose (the wide-only path has no main kernel); the shapes that start the longest extensions
    # first (3 and 5) add the estimate and the sort in front
    assert 3 - (st["lanes"] == 0) <= st["launches"] <= 8 and st["kernel_ms"] > 0
    if (lanes, cells) == (-1, -1):      # round 2: (3, 64) measured fastest at x = 7 (profiles/xdrop_r02.md)
        assert (st["lanes"], st["cells_per_lane"]) == ((3, 64) if xdrop <= 12 else (32, 2) if xdrop <= 40 else (0, 0))
    a.close()


def test_window_overflow_goes_through_the_wide_kernel(reads):
    inp, pairs = reads
    want = ol.oracle_align_post(inp, *pairs, 7, 0.55, 0.1, 200)
    for shape, least in (((16, 1), 100), ((1, 32), 1)):             # 16 slots: about a quarter of the extensions outgrow them
        a = aligner(inp, 7, shape, fixed_threshold=200)
        got = a.align(*pairs)
        assert a.stats()["wide_extensions"] >= least
        np.testing.assert_array_equal(got, want)
        a.close()


def test_low_error_reads_and_seeds_at_the_read_ends():
    inp = fe.synthetic(150, 3000, coverage=20.0, err=0.02, seed=77, hi=40)
    pairs = candidate_pairs(inp, 1500)
    for shape in ((-1, -1), (32, 1)):
        a = aligner(inp, 7, shape)
        np.testing.assert_array_equal(a.align(*pairs), ol.oracle_align_post(inp, *pairs, 7, 0.55, 0.1, -1))
        a.close()
    a = aligner(inp)
    k, n = inp.kmer_size, 40
    r = np.arange(1, n + 1, dtype=np.uint32)
    c = np.zeros(n, dtype=np.uint32)
    lens = inp.read_len
    pH = np.where(np.arange(n) % 2 == 0, 0, lens[r] - k).astype(np.uint16)
    pV = np.where(np.arange(n) % 3 == 0, 0, lens[c] - k).astype(np.uint16)
    np.testing.assert_array_equal(a.align(r, c, pH, pV), ol.oracle_align_post(inp, r, c, pH, pV, 7, 0.55, 0.1, -1))
    assert a.align(r[:0], c[:0], pH[:0], pV[:0]).shape == (0, 8)    # empty batch
    a.close()


def test_bases_are_dna5_as_seqan_sees_them(reads):
    """N, lower case, U and IUPAC letters: alignSeqAn converts to seqan::Dna5 first (N matches N, 'a' == 'A')"""
    import copy
    inp, pairs = reads
    dirty = copy.copy(inp)
    s = inp.seqs.copy()
    rng = np.random.default_rng(5)
    idx = rng.choice(len(s), len(s) // 40, replace=False)
    s[idx] = np.frombuffer(b"NnacgtRUuYx-", dtype=np.uint8)[rng.integers(0, 12, len(idx))]
    dirty.seqs = s
    want = ol.oracle_align_post(dirty, *pairs, 7, 0.55, 0.1, -1)
    for shape in ((-1, -1), (32, 1), (0, 0)):
        a = aligner(dirty, 7, shape)
        np.testing.assert_array_equal(a.align(*pairs), want)
        a.close()


def test_maximum_read_length_and_scores_beyond_16_bits():
    """two 65 535-base reads overlapping end to end at 3 % error: coordinates up to the u16 limit, scores near 60 000"""
    rng = np.random.default_rng(11)
    base = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 65535)
    other = base.copy()
    idx = rng.choice(len(other), len(other) * 3 // 100, replace=False)
    other[idx] = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), len(idx))
    for p in (0, 30000, 65535 - 17):
        other[p:p + 17] = base[p:p + 17]
    inp = fe.OverlapInputs(n_reads=2, n_kmers=0, nnz=0, A_colptr=None, A_rowids=None, A_values=None, A_strand=None, B_colptr=None,
                           B_rowids=None, B_values=None, B_strand=None, read_len=np.array([65535, 65535], dtype=np.uint32), kmer_size=17,
                           seqs=np.concatenate([base, other]), seq_off=np.array([0, 65535, 131070], dtype=np.uint64))
    rows = np.array([1, 1, 1], dtype=np.uint32); cols = np.array([0, 0, 0], dtype=np.uint32)
    pos = np.array([0, 30000, 65535 - 17], dtype=np.uint16)
    want = ol.oracle_align_post(inp, rows, cols, pos, pos, 50, 0.5, 0.1, -1)
    assert want[:, 0].min() > 40000
    for shape in ((-1, -1), (1, 64), (32, 4), (0, 0)):
        a = aligner(inp, 50, shape, ratiophi=0.5)
        np.testing.assert_array_equal(a.align(rows, cols, pos, pos), want)
        a.close()


def test_reference_golden_fixture():
    z = np.load(os.path.join(golden_util.GOLDEN, "xdrop.npz"))
    inp = fe.OverlapInputs(n_reads=int(z["n_reads"]), n_kmers=0, nnz=0, A_colptr=None, A_rowids=None, A_values=None, A_strand=None,
                           B_colptr=None, B_rowids=None, B_values=None, B_strand=None, read_len=None, kmer_size=int(z["k"]),
                           seqs=z["seqs"], seq_off=z["seq_off"])
    a = aligner(inp, int(z["xdrop"]))
    got = a.align(z["rows"], z["cols"], z["posH"], z["posV"])
    np.testing.assert_array_equal(got[:, :6], z["ref_out"])
    a.close()


def test_bad_arguments_are_refused(reads):
    from bella_b200 import xdrop as xd
    inp, pairs = reads
    a = aligner(inp)
    rows, cols, pH, pV = (x[:16].copy() for x in pairs)
    pH[5] = inp.read_len[rows[5]] - 3                               # seed sticks out of its read
    with pytest.raises(xd.BellaXdropError, match="-3"):
        a.align(rows, cols, pH, pV)
    rows[2] = inp.n_reads + 7
    with pytest.raises(xd.BellaXdropError, match="-1"):
        a.align(rows, cols, pH, pV)
    with pytest.raises(xd.BellaXdropError):
        a.set_shape(8, 1)
    b = xd.XdropAligner(0)
    with pytest.raises(xd.BellaXdropError, match="set_reads"):
        b.align(*(x[:4] for x in pairs))
    a.close(); b.close()


def test_chained_behind_the_overlap_spgemm_on_the_device(small_inputs):
    """SpGEMM result stays on the device and is aligned as it is (CSC form): the reference's RunPairWiseAlignments loop"""
    import torch
    from bella_b200 import spgemm
    inp = small_inputs
    g = spgemm.OverlapSpGEMM(0)
    g.set_inputs(inp)
    g.symbolic()
    g.numeric_device()
    res = g.result_device()
    want_c = ol.oracle_spgemm(inp, want_aux=False)
    assert res["nnz"] == want_c.nnz
    a = aligner(inp)
    out = torch.empty((res["nnz"], 8), dtype=torch.int32, device="cuda:0")
    torch.cuda.synchronize()
    a.align_csc_device(inp.n_reads, res["colptrC"], res["nnz"], res["rowids"], res["posH"], res["posV"], out)
    a.sync()
    cols = np.repeat(np.arange(inp.n_reads, dtype=np.uint32), np.diff(want_c.colptrC.astype(np.int64)))
    want = ol.oracle_align_post(inp, want_c.rowids, cols, want_c.posH, want_c.posV, 7, 0.55, 0.1, -1)
    np.testing.assert_array_equal(out.cpu().numpy(), want)
    a.close()


@pytest.mark.skipif(os.environ.get("BELLA_RUN_LOGAN") != "1" or not ol.have_logan(),
                    reason="opt-in (BELLA_RUN_LOGAN=1): the reference's CUDA aligner recompiled for sm_100a as a second live reference; "
                           "not yet run on a B200, so it does not gate the suite")
def test_logan_recompiled_agrees(reads):
    """LOGAN keeps its anti-diagonals in `short`, so it can only agree while scores stay below 32767 (true here)."""
    inp, pairs = reads
    got, seconds = ol.logan_align(inp, *pairs, 7)
    want = ol.oracle_align(inp, *pairs, 7)
    same = (got == want).all(axis=1).mean()
    print(f"LOGAN vs oracle: {same:.4f} of {len(want)} pairs identical, {seconds:.3f} s in extendSeedL")
    assert same > 0.99
