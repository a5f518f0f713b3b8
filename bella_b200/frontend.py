"""ctypes wrapper over libbella_frontend.so (bella_b200/csrc/frontend.cpp): the host front end that
builds the read x k-mer matrices the overlap SpGEMM consumes (mirror of the reference's
src/main.cpp:339-489 + include/kmercount.hpp reliable-k-mer selection) and the seeded read simulator
used by tests and bench.py.  Host-only; never on the timed GPU path."""
import ctypes
import os
from dataclasses import dataclass

import numpy as np

from . import _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = _build.build_frontend()
        L = ctypes.CDLL(path)
        L.bella_fe_simulate.restype = ctypes.c_int
        L.bella_fe_simulate.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_double,
                                        ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_uint64,
                                        ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p)]
        L.bella_fe_free_buf.argtypes = [ctypes.c_void_p]
        L.bella_fe_build.restype = ctypes.c_void_p
        L.bella_fe_build.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
        L.bella_fe_build_w.restype = ctypes.c_void_p
        L.bella_fe_build_w.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
        for f, rt in (("bella_fe_n", ctypes.c_uint32), ("bella_fe_m", ctypes.c_uint32),
                      ("bella_fe_nnz", ctypes.c_uint64), ("bella_fe_ntuples", ctypes.c_uint64)):
            getattr(L, f).restype = rt
            getattr(L, f).argtypes = [ctypes.c_void_p]
        L.bella_fe_array.restype = ctypes.c_void_p
        L.bella_fe_array.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.bella_fe_free.argtypes = [ctypes.c_void_p]
        _lib = L
    return _lib


@dataclass
class OverlapInputs:
    """Everything the SpGEMM boundary takes (include/bella_b200.h): A = reads x k-mers, B = Aᵀ, both CSC
    with uint32 colptr/rowids and uint16 values (k-mer positions), bit-packed strand bits in each
    matrix's array order, read lengths, k and the bin size."""
    n_reads: int
    n_kmers: int
    nnz: int
    A_colptr: np.ndarray
    A_rowids: np.ndarray
    A_values: np.ndarray
    A_strand: np.ndarray
    B_colptr: np.ndarray
    B_rowids: np.ndarray
    B_values: np.ndarray
    B_strand: np.ndarray
    read_len: np.ndarray
    kmer_size: int = 17
    bin_size: int = 500
    seqs: np.ndarray = None      # concatenated read characters (uint8); only the CPU reference needs them
    seq_off: np.ndarray = None   # uint64 [n_reads+1]
    tuples: tuple = None         # (kmer, read, pos) before de-duplication, when requested

    def save(self, path):
        d = {k: v for k, v in self.__dict__.items() if isinstance(v, np.ndarray)}
        d["meta"] = np.array([self.n_reads, self.n_kmers, self.nnz, self.kmer_size, self.bin_size], dtype=np.int64)
        np.savez_compressed(path, **d)

    @staticmethod
    def load(path):
        z = np.load(path)
        meta = z["meta"]
        kw = {k: z[k] for k in z.files if k != "meta"}
        return OverlapInputs(n_reads=int(meta[0]), n_kmers=int(meta[1]), nnz=int(meta[2]),
                             kmer_size=int(meta[3]), bin_size=int(meta[4]), **kw)


def simulate_reads(genome_len, n_reads, read_len, err=0.15, split=(0.10, 0.60, 0.30), seed=1):
    """Uniform random genome, fixed-length reads, random start/strand, sub/ins/del errors
    (SURVEY.md section 8d).  Returns (seqs uint8[n*L], offs uint64[n+1])."""
    L = lib()
    ps, po = ctypes.c_void_p(), ctypes.c_void_p()
    rc = L.bella_fe_simulate(genome_len, n_reads, read_len, err, split[0], split[1], split[2], seed,
                             ctypes.byref(ps), ctypes.byref(po))
    if rc != 0:
        raise ValueError(f"bella_fe_simulate failed: {rc}")
    try:
        seqs = np.ctypeslib.as_array(ctypes.cast(ps, ctypes.POINTER(ctypes.c_uint8)), (n_reads * read_len,)).copy()
        offs = np.ctypeslib.as_array(ctypes.cast(po, ctypes.POINTER(ctypes.c_uint64)), (n_reads + 1,)).copy()
    finally:
        L.bella_fe_free_buf(ps)
        L.bella_fe_free_buf(po)
    return seqs, offs


def reads_from_strings(reads):
    seqs = np.frombuffer("".join(reads).encode(), dtype=np.uint8).copy()
    offs = np.zeros(len(reads) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(r) for r in reads])
    return seqs, offs


def read_fastq(path):
    names, reads = [], []
    with open(path) as f:
        lines = [l.rstrip("\n") for l in f]
    for i in range(0, len(lines) - 3, 4):
        names.append(lines[i][1:].split()[0])
        reads.append(lines[i + 1])
    return names, reads


def build_matrices(seqs, offs, k=17, lo=2, hi=8, bin_size=500, keep_tuples=False, nthreads=0, keep_seqs=True, window=0):
    """reads -> OverlapInputs (reliable k-mers in [lo,hi], B with the reference's MergeDuplicates
    order, A = Bᵀ, strand bits, read lengths).  window > 0: BELLA's minimizer mode (-w): only the positions the reference's
    getMinimizers samples (include/minimizer.hpp:49-77) are counted and kept."""
    L = lib()
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
    offs = np.ascontiguousarray(offs, dtype=np.uint64)
    n = len(offs) - 1
    err = ctypes.c_int(0)
    h = L.bella_fe_build_w(seqs.ctypes.data, offs.ctypes.data, n, k, lo, hi, int(keep_tuples), nthreads, int(window), ctypes.byref(err))
    if not h:
        raise ValueError({-1: "bad arguments", -2: "read contains a character other than upper-case ACGT "
                          "(strand-bit contract, SURVEY 8c)", -3: "read longer than 65535 (u16 positions)",
                          -4: "more than 2^32-1 k-mers or tuples"}.get(err.value, f"error {err.value}"))
    try:
        m, nnz, nt = L.bella_fe_m(h), L.bella_fe_nnz(h), L.bella_fe_ntuples(h)

        def arr(which, dtype, count):
            p = L.bella_fe_array(h, which)
            if count == 0:
                return np.zeros(0, dtype=dtype)
            ct = {np.uint32: ctypes.c_uint32, np.uint16: ctypes.c_uint16, np.uint8: ctypes.c_uint8}[dtype]
            return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ct)), (count,)).copy()

        nb = (nnz + 7) // 8 + 8
        out = OverlapInputs(
            n_reads=n, n_kmers=m, nnz=nnz,
            B_colptr=arr(0, np.uint32, n + 1), B_rowids=arr(1, np.uint32, nnz), B_values=arr(2, np.uint16, nnz),
            B_strand=arr(3, np.uint8, nb),
            A_colptr=arr(4, np.uint32, m + 1), A_rowids=arr(5, np.uint32, nnz), A_values=arr(6, np.uint16, nnz),
            A_strand=arr(7, np.uint8, nb), read_len=arr(8, np.uint32, n), kmer_size=k, bin_size=bin_size)
        if keep_tuples:
            out.tuples = (arr(9, np.uint32, nt), arr(10, np.uint32, nt), arr(11, np.uint16, nt))
        if keep_seqs:
            out.seqs, out.seq_off = seqs, offs
        return out
    finally:
        L.bella_fe_free(h)


def synthetic(n_reads, read_len, coverage=30.0, err=0.15, split=(0.10, 0.60, 0.30), seed=1, k=17, lo=2, hi=8,
              bin_size=500, keep_tuples=False, nthreads=0, window=0):
    """One call: simulate + build.  genome_len = n*L/coverage."""
    G = max(int(n_reads * read_len / coverage), 2 * read_len + 64 + int(read_len * 1.5) + 64)
    seqs, offs = simulate_reads(G, n_reads, read_len, err, split, seed)
    return build_matrices(seqs, offs, k, lo, hi, bin_size, keep_tuples, nthreads, window=window)
