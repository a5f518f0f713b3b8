"""CPU: the per-element functions the f3 kernels are made of (bella_b200/csrc/kmers.cuh -- extract, classify, place, emit)
compiled for the host and run in plain loops (tests/emu/kmers_host.cpp; std::sort / prefix sums where the device uses cub)
against the oracle of the row (oracle_reliable_occurrences, pinned against the reference's SplitCount in
test_oracle_kmers.py) and against the host front end the SpGEMM tests use.  The device library itself
(bella_b200/libbella_kmers.so) has not run on a B200 yet; its GPU tests are tests/test_kmers_gpu.py."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
from bella_b200 import frontend as fe

EMU_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")


def host_count(inp, k, lower, upper):
    subprocess.run(["make", "-s", "-C", EMU_DIR, "all"], check=True)
    L = ctypes.CDLL(os.path.join(EMU_DIR, "_build", "libkmers_host.so"))
    cap = len(inp.seqs)
    t_kmer = np.zeros(cap, dtype=np.uint32); t_read = np.zeros(cap, dtype=np.uint32); t_pos = np.zeros(cap, dtype=np.uint16)
    bits = np.zeros((cap + 7) // 8, dtype=np.uint8)
    nk, nt = ctypes.c_uint64(0), ctypes.c_uint64(0)
    p = lambda a: ctypes.c_void_p(a.ctypes.data)  # noqa: E731
    rc = L.kmers_host_count(p(inp.seqs), p(inp.seq_off), ctypes.c_uint32(inp.n_reads), k, lower, upper, p(t_kmer), p(t_read), p(t_pos), p(bits),
                            ctypes.c_uint64(cap), ctypes.byref(nk), ctypes.byref(nt))
    assert rc == 0
    n = nt.value
    return t_kmer[:n], t_read[:n], t_pos[:n], np.unpackbits(bits, bitorder="little")[:n], nk.value


def check_against_oracle(inp, k, lower, upper):
    t_kmer, t_read, t_pos, strand, n_kmers = host_count(inp, k, lower, upper)
    r, p, want_kmers = ol.oracle_reliable_occurrences(inp, k, lower, upper)
    assert n_kmers == want_kmers
    np.testing.assert_array_equal(t_read, r)                      # read / position order, as src/main.cpp:393-416 emits
    np.testing.assert_array_equal(t_pos, p)
    assert n_kmers > 0 and t_kmer.max() == n_kmers - 1 and len(np.unique(t_kmer)) == n_kmers
    # strand bit: 1 iff the window equals its canonical representative (<= its reverse complement in Kmer's packing order)
    code = np.zeros(256, dtype=np.uint8)
    for c in range(256):
        x = (c & 4) >> 1
        code[c] = x + ((x ^ (c & 2)) >> 1)
    for t in np.random.default_rng(0).choice(len(t_read), min(300, len(t_read)), replace=False):
        g = int(inp.seq_off[t_read[t]]) + int(t_pos[t])
        w = code[inp.seqs[g:g + k]]
        assert strand[t] == (tuple(w) <= tuple(3 - w[::-1]))
    # same id <=> same canonical k-mer (spot check on the first ids)
    for i in range(min(20, n_kmers)):
        occ = np.nonzero(t_kmer == i)[0]
        canon = set()
        for t in occ:
            g = int(inp.seq_off[t_read[t]]) + int(t_pos[t])
            w = code[inp.seqs[g:g + k]]
            canon.add(min(tuple(w), tuple(3 - w[::-1])))
        assert len(canon) == 1 and lower <= len(occ) <= upper
    return t_kmer, t_read, t_pos, strand


@pytest.mark.parametrize("k,lower,upper", [(17, 2, 8), (15, 2, 4), (21, 3, 50), (32, 2, 8), (10, 1, 3)])
def test_kernel_functions_select_the_oracle_occurrences(k, lower, upper):
    check_against_oracle(fe.synthetic(200, 2500, seed=k), k, lower, upper)


def test_ragged_and_unusual_reads():
    import copy
    rng = np.random.default_rng(4)
    inp = fe.synthetic(150, 2000, seed=9)
    s = inp.seqs.copy()
    idx = rng.choice(len(s), len(s) // 60, replace=False)
    s[idx] = np.frombuffer(b"NnacgtRYK", dtype=np.uint8)[rng.integers(0, 9, len(idx))]
    # ragged: reads shorter than k, empty reads, a read of exactly k bases
    lens = np.diff(inp.seq_off.astype(np.int64))
    lens[3] = 0; lens[4] = 5; lens[5] = 17; lens[-1] = 16
    off = np.zeros(inp.n_reads + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    dirty = copy.copy(inp)
    dirty.seqs, dirty.seq_off = s[:int(off[-1])].copy(), off
    check_against_oracle(dirty, 17, 2, 8)


def test_ids_and_strand_bits_match_the_host_front_end():
    """same tuples as bella_fe_build (ids there are ranks in (hash bucket, k-mer) order, so compare up to a renaming)"""
    seqs, offs = fe.simulate_reads(300000, 600, 4000, 0.15, (0.10, 0.60, 0.30), 5)
    inp = fe.build_matrices(seqs, offs, 17, 2, 8, keep_tuples=True)
    t_kmer, t_read, t_pos, strand = check_against_oracle(inp, 17, 2, 8)
    f_kmer, f_read, f_pos = inp.tuples
    key = lambda r, p: r.astype(np.uint64) << np.uint64(16) | p.astype(np.uint64)  # noqa: E731
    o1, o2 = np.argsort(key(t_read, t_pos)), np.argsort(key(f_read, f_pos))
    a, b = t_kmer[o1], f_kmer[o2]
    # a renaming: equal ids here <=> equal ids there
    m = {}
    assert all(m.setdefault(int(x), int(y)) == int(y) for x, y in zip(a, b)) and len(set(m.values())) == len(m)
