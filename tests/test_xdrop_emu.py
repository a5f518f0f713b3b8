"""CPU: the DEVICE source of the X-drop kernels (bella_b200/csrc/xdrop.cuh -- register path, overflow hand-over, wide
path, the fused threshold test) compiled for the host under the lane-fiber emulator of tests/emu/ and compared with the
oracle (oracle_xdrop_align + oracle_xdrop_post, both pinned against the reference in test_oracle_xdrop.py).  Same code
the GPU runs, so logic errors show up here without a B200; the GPU parity tests are tests/test_xdrop_gpu.py."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import golden_util
import oracle_lib as ol
from bella_b200 import frontend as fe

EMU_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")
_emu = None


def emu_lib():
    global _emu
    if _emu is None:
        subprocess.run(["make", "-s", "-C", EMU_DIR], check=True)
        _emu = ctypes.CDLL(os.path.join(EMU_DIR, "_build", "libxdrop_emu.so"))
    return _emu


def _p(a):
    return ctypes.c_void_p(a.ctypes.data)


def emu_align(inp, rows, cols, posH, posV, xdrop, lanes, cells, ratiophi=0.5, delta=0.1, fixed_threshold=-1, warps=8, colptr=None):
    rows = np.ascontiguousarray(rows, dtype=np.uint32)
    posH = np.ascontiguousarray(posH, dtype=np.uint16); posV = np.ascontiguousarray(posV, dtype=np.uint16)
    if colptr is None:
        cols = np.ascontiguousarray(cols, dtype=np.uint32)
        c_cols, c_colptr, n_cols = _p(cols), None, 0
    else:                                                           # CSC form: the column of a pair comes from colptr
        colptr = np.ascontiguousarray(colptr, dtype=np.uint32)
        c_cols, c_colptr, n_cols = None, _p(colptr), len(colptr) - 1
    out = np.zeros((len(rows), 8), dtype=np.int32)
    n_wide = ctypes.c_int(0)
    rc = emu_lib().xdrop_emu_align(lanes, cells, ctypes.c_uint64(len(rows)), _p(rows), c_cols, _p(posH), _p(posV), _p(inp.seqs),
                                   _p(inp.seq_off), ctypes.c_uint32(len(inp.seq_off) - 1), ctypes.c_int(inp.kmer_size), ctypes.c_int(xdrop),
                                   ctypes.c_double(ratiophi), ctypes.c_double(delta), ctypes.c_int(fixed_threshold), ctypes.c_int(warps),
                                   _p(out), ctypes.byref(n_wide), c_colptr, ctypes.c_int(n_cols))
    return rc, out, n_wide.value


def candidate_pairs(inp, limit, seed=0):
    r = ol.oracle_spgemm(inp, want_aux=False)
    cols = np.repeat(np.arange(inp.n_reads, dtype=np.uint32), np.diff(r.colptrC.astype(np.int64)))
    idx = np.arange(r.nnz)
    if r.nnz > limit:
        idx = np.sort(np.random.default_rng(seed).choice(r.nnz, limit, replace=False))
    return r.rowids[idx], cols[idx], r.posH[idx], r.posV[idx]


@pytest.fixture(scope="module")
def reads():
    inp = fe.synthetic(120, 1200, coverage=14.0, seed=31)
    return inp, candidate_pairs(inp, 400, seed=3)


# (lanes per extension, cells per lane, xdrop): every instantiation the library ships; (1, W) = one thread per extension
# with W window slots, (0,0) = wide path only
# (2, W) / (3, W) = the packed-word form of the thread path / the same with the longest-first job order;
# (4, W) / (5, W) = both anti-diagonals of a column in one word / the same with the longest-first job order
@pytest.mark.parametrize("lanes,cells,xdrop", [(1, 64, 7), (1, 32, 7), (2, 64, 7), (3, 64, 7), (2, 32, 3), (3, 32, 15), (4, 64, 7), (5, 64, 7),
                                                (4, 32, 3), (5, 32, 15), (4, 128, 30), (5, 256, 60), (32, 1, 7), (32, 2, 15), (32, 4, 30), (16, 1, 3), (16, 2, 7), (16, 4, 15),
                                                (8, 4, 7), (8, 8, 15), (0, 0, 7)])
def test_device_source_matches_oracle(reads, lanes, cells, xdrop):
    inp, pairs = reads
    want = ol.oracle_align_post(inp, *pairs, xdrop, 0.55, 0.1, -1)
    rc, got, _ = emu_align(inp, *pairs, xdrop, lanes, cells, 0.55, 0.1, -1)
    assert rc == 0
    np.testing.assert_array_equal(got, want)
    assert want[:, 0].max() > 300 and (want[:, 1] == ord("c")).any() and (want[:, 1] == ord("n")).any()
    assert 0 < want[:, 7].sum() < len(want) or xdrop != 7          # the threshold test separates the pairs


def test_window_overflow_hands_over_to_the_wide_path(reads):
    inp, pairs = reads
    want = ol.oracle_align_post(inp, *pairs, 7, 0.55, 0.1, 200)
    rc, got, n_wide = emu_align(inp, *pairs, 7, 16, 1, 0.55, 0.1, 200)     # 16 slots: a quarter of the extensions outgrow them
    assert rc == 0 and n_wide > 20
    np.testing.assert_array_equal(got, want)
    for lanes in (1, 3, 5):                                                # the thread-per-extension paths hand over the same way
        rc, got, n_wide = emu_align(inp, *pairs, 7, lanes, 16, 0.55, 0.1, 200)
        assert rc == 0 and n_wide > 20
        np.testing.assert_array_equal(got, want)
    rc, got, n_wide = emu_align(inp, *pairs, 25, 32, 1, 0.55, 0.1, 200)    # x = 25 needs ~34 columns: most go wide
    assert rc == 0 and n_wide > len(want)
    np.testing.assert_array_equal(got, ol.oracle_align_post(inp, *pairs, 25, 0.55, 0.1, 200))


def test_low_error_reads_and_seeds_at_the_read_ends():
    inp = fe.synthetic(60, 1500, coverage=15.0, err=0.02, seed=77, hi=40)       # extensions that run into the read ends
    pairs = candidate_pairs(inp, 200)
    for lanes, cells in ((32, 1), (1, 64), (2, 64), (4, 64)):
        rc, got, _ = emu_align(inp, *pairs, 7, lanes, cells)
        assert rc == 0
        np.testing.assert_array_equal(got, ol.oracle_align_post(inp, *pairs, 7))
    k, n = inp.kmer_size, 30
    r = np.arange(1, n + 1, dtype=np.uint32)
    c = np.zeros(n, dtype=np.uint32)
    lens = inp.read_len
    pH = np.where(np.arange(n) % 2 == 0, 0, lens[r] - k).astype(np.uint16)      # empty prefix / empty suffix
    pV = np.where(np.arange(n) % 3 == 0, 0, lens[c] - k).astype(np.uint16)
    for lanes, cells in ((1, 64), (3, 64), (5, 64), (32, 1), (16, 2), (0, 0)):
        rc, got, _ = emu_align(inp, r, c, pH, pV, 7, lanes, cells)
        assert rc == 0
        np.testing.assert_array_equal(got, ol.oracle_align_post(inp, r, c, pH, pV, 7))


def test_reference_golden_fixture():
    z = np.load(os.path.join(golden_util.GOLDEN, "xdrop.npz"))
    inp = fe.OverlapInputs(n_reads=int(z["n_reads"]), n_kmers=0, nnz=0, A_colptr=None, A_rowids=None, A_values=None, A_strand=None,
                           B_colptr=None, B_rowids=None, B_values=None, B_strand=None, read_len=None, kmer_size=int(z["k"]),
                           seqs=z["seqs"], seq_off=z["seq_off"])
    sel = slice(0, 3000, 6)
    for lanes, cells in ((32, 1), (1, 64)):
        rc, got, _ = emu_align(inp, z["rows"][sel], z["cols"][sel], z["posH"][sel], z["posV"][sel], int(z["xdrop"]), lanes, cells)
        assert rc == 0
        np.testing.assert_array_equal(got[:, :6], z["ref_out"][sel])


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
    want = ol.oracle_align_post(dirty, *pairs, 7)
    assert (want[:, :6] != ol.oracle_align(inp, *pairs, 7)).any()
    for lanes, cells in ((1, 64), (2, 64), (4, 64), (32, 1), (0, 0)):
        rc, got, _ = emu_align(dirty, *pairs, 7, lanes, cells)
        assert rc == 0
        np.testing.assert_array_equal(got, want)


def test_csc_form_takes_the_spgemm_result_as_it_is():
    inp = fe.synthetic(40, 900, coverage=10.0, seed=9)
    r = ol.oracle_spgemm(inp, want_aux=False)
    assert (np.diff(r.colptrC.astype(np.int64)) == 0).any()        # empty columns are part of the case
    cols = np.repeat(np.arange(inp.n_reads, dtype=np.uint32), np.diff(r.colptrC.astype(np.int64)))
    for lanes, cells in ((32, 1), (1, 64)):
        rc, got, _ = emu_align(inp, r.rowids, None, r.posH, r.posV, 7, lanes, cells, colptr=r.colptrC)
        assert rc == 0
        np.testing.assert_array_equal(got, ol.oracle_align_post(inp, r.rowids, cols, r.posH, r.posV, 7))


def test_seed_outside_its_read_is_reported(reads):
    inp, pairs = reads
    rows, cols, pH, pV = (a[:8].copy() for a in pairs)
    pH[3] = inp.read_len[rows[3]] - 3
    for lanes, cells in ((32, 1), (1, 64), (3, 64), (0, 0)):
        rc, _, _ = emu_align(inp, rows, cols, pH, pV, 7, lanes, cells)
        assert rc == -1


def _inputs_from_strings(reads, k=17):
    seqs = np.frombuffer("".join(reads).encode(), dtype=np.uint8).copy()
    off = np.zeros(len(reads) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(r) for r in reads])
    return fe.OverlapInputs(n_reads=len(reads), n_kmers=0, nnz=0, A_colptr=None, A_rowids=None, A_values=None, A_strand=None,
                            B_colptr=None, B_rowids=None, B_values=None, B_strand=None,
                            read_len=np.array([len(r) for r in reads], dtype=np.uint32), kmer_size=k, seqs=seqs, seq_off=off)


@pytest.mark.parametrize("xdrop", [0, 1, 7, 200])
def test_degenerate_reads_and_extreme_xdrop(xdrop):
    """homopolymers, tandem repeats, reads of exactly k bases, exact reverse complements, arbitrary (non-matching) seeds,
    x = 0 (the initial gap cells are undefined, seeds_extension.h:476) and x larger than any score: reference (when built),
    oracle and every execution shape of the device source agree"""
    rng = np.random.default_rng(1)
    rnd = lambda n: "".join(rng.choice(list("ACGT"), n))  # noqa: E731
    rc = lambda s: s[::-1].translate(str.maketrans("ACGT", "TGCA"))  # noqa: E731

    def mutate(s, e):
        out = []
        for c in s:
            u = rng.random()
            if u < e / 3:
                continue
            out.append(rng.choice(list("ACGT")) if u < 2 * e / 3 else c)
            if 2 * e / 3 <= u < e:
                out.append(rng.choice(list("ACGT")))
        return "".join(out)

    base = rnd(700)
    reads = [base, mutate(base, 0.1), rc(mutate(base, 0.1)), "A" * 300, "A" * 250 + "C" * 50, "AC" * 150, base[:17], base[:40], rnd(200),
             base[100:500], rc(base[100:500])]
    inp = _inputs_from_strings(reads)
    rows, cols, pH, pV = [], [], [], []
    for i in range(len(reads)):
        for j in range(len(reads)):
            if i == j:
                continue
            for _ in range(3):
                rows.append(i); cols.append(j)
                pH.append(int(rng.integers(0, len(reads[i]) - 16))); pV.append(int(rng.integers(0, len(reads[j]) - 16)))
            rows += [i, i]; cols += [j, j]
            pH += [0, len(reads[i]) - 17]; pV += [0, len(reads[j]) - 17]
    for a in (0, 1, 150, 382, 383):                                   # exact reverse complements: twin(seedH) == seedV
        rows.append(9); cols.append(10); pH.append(a); pV.append(400 - a - 17)
        rows.append(10); cols.append(9); pH.append(400 - a - 17); pV.append(a)
    pairs = (np.array(rows, dtype=np.uint32), np.array(cols, dtype=np.uint32), np.array(pH, dtype=np.uint16), np.array(pV, dtype=np.uint16))
    want = ol.oracle_align_post(inp, *pairs, xdrop, 0.4, 0.1, -1)
    assert (want[:, 1] == ord("c")).sum() >= 10 and want[:, 0].max() >= 400
    if ol.have_ref():
        np.testing.assert_array_equal(want[:, :6], ol.ref_align(inp, *pairs, xdrop))
    for lanes, cells in ((1, 64), (2, 64), (4, 64), (32, 1), (32, 4), (16, 2), (0, 0)):
        rc_, got, _ = emu_align(inp, *pairs, xdrop, lanes, cells, 0.4, 0.1, -1)
        assert rc_ == 0
        np.testing.assert_array_equal(got, want)


def test_device_source_is_clean_under_address_sanitizer(tmp_path):
    """every execution shape once more with the emulator built with -fsanitize=address: the per-thread windows, the base
    rings and the wide path's scratch are heap blocks there, so an out-of-bounds slot is reported instead of going unnoticed"""
    import shutil
    import sys
    cxx = shutil.which("g++") or "g++"
    asan = subprocess.run([cxx, "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("libasan not available")
    subprocess.run(["make", "-s", "-C", EMU_DIR, "asan"], check=True)
    code = f"""
import ctypes, sys
sys.path[:0] = {[os.path.dirname(EMU_DIR), os.path.dirname(os.path.dirname(EMU_DIR))]!r}
import numpy as np
import test_xdrop_emu as E, oracle_lib as ol
from bella_b200 import frontend as fe
E._emu = ctypes.CDLL({os.path.join(EMU_DIR, "_build", "libxdrop_emu_asan.so")!r})
inp = fe.synthetic(80, 1000, coverage=12.0, seed=31)
pairs = E.candidate_pairs(inp, 120, seed=3)
for G, T, x in [(1, 64, 7), (1, 16, 7), (2, 64, 7), (2, 16, 7), (4, 64, 7), (4, 16, 7), (5, 128, 30), (32, 1, 7), (32, 4, 30), (16, 2, 7), (8, 4, 7),
                (16, 1, 3), (0, 0, 7)]:
    rc, got, _ = E.emu_align(inp, *pairs, x, G, T, 0.55, 0.1, -1, warps=2)
    assert rc == 0 and np.array_equal(got, ol.oracle_align_post(inp, *pairs, x, 0.55, 0.1, -1)), (G, T, x)
print("clean")
"""
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0 and "clean" in out.stdout, (out.stdout[-500:], out.stderr[-3000:])


def test_maximum_read_length_and_scores_beyond_16_bits():
    """two 65 535-base reads (BELLA's position type is unsigned short) overlapping end to end at 3 % error: coordinates up to
    the u16 limit and a score near 60 000 -- beyond what the reference's CUDA port keeps in `short` -- through every
    thread-kernel encoding (plain ints, score << 6 | bases, scores relative to the drop-off limit)"""
    rng = np.random.default_rng(11)
    base = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 65535)
    other = base.copy()
    idx = rng.choice(len(other), len(other) * 3 // 100, replace=False)
    other[idx] = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), len(idx))         # substitutions only: same length
    # a clean seed in the middle and one at each end
    for p in (0, 30000, 65535 - 17):
        other[p:p + 17] = base[p:p + 17]
    seqs = np.concatenate([base, other])
    off = np.array([0, 65535, 131070], dtype=np.uint64)
    inp = fe.OverlapInputs(n_reads=2, n_kmers=0, nnz=0, A_colptr=None, A_rowids=None, A_values=None, A_strand=None, B_colptr=None,
                           B_rowids=None, B_values=None, B_strand=None, read_len=np.array([65535, 65535], dtype=np.uint32), kmer_size=17,
                           seqs=seqs, seq_off=off)
    rows = np.array([1, 1, 1], dtype=np.uint32); cols = np.array([0, 0, 0], dtype=np.uint32)
    pos = np.array([0, 30000, 65535 - 17], dtype=np.uint16)
    want = ol.oracle_align_post(inp, rows, cols, pos, pos, 50, 0.5, 0.1, -1)
    assert want[:, 0].min() > 40000 and (want[:, 3] == 65535).all() and (want[:, 2] == 0).all()
    if ol.have_ref():
        np.testing.assert_array_equal(want[:, :6], ol.ref_align(inp, rows, cols, pos, pos, 50))
    for lanes, cells in ((1, 64), (2, 64), (4, 64), (5, 128), (32, 4), (0, 0)):
        rc, got, _ = emu_align(inp, rows, cols, pos, pos, 50, lanes, cells, 0.5, 0.1, -1, warps=1)
        assert rc == 0
        np.testing.assert_array_equal(got, want)
