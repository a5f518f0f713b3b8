"""Host-side Python binding of include/bella_kmers.h over libbella_kmers.so: the "next" row f3, reliable k-mer selection and
tuple emission on the device (the reference's SplitCount, include/kmercount.hpp:466-677, and the tuple loop of
src/main.cpp:339-423).  No CPU fallback.  Not yet run on a B200 (see the header)."""
import ctypes
import os

import numpy as np

from . import _build

_lib = None

EXPORTS = ["bella_kmers_create", "bella_kmers_destroy", "bella_kmers_last_error", "bella_kmers_count", "bella_kmers_get_tuples",
           "bella_kmers_get_stats"]


class BellaKmersError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        path = _build.LIB_KMERS
        path = _build.build_kmers()                        # returns at once when the .so matches its sources
        L = ctypes.CDLL(path)
        vp, H = ctypes.c_void_p, ctypes.c_void_p
        L.bella_kmers_create.argtypes = [ctypes.c_int]
        L.bella_kmers_create.restype = H
        L.bella_kmers_destroy.argtypes = [H]
        L.bella_kmers_destroy.restype = None
        L.bella_kmers_last_error.argtypes = [H]
        L.bella_kmers_last_error.restype = ctypes.c_char_p
        L.bella_kmers_count.argtypes = [H, vp, vp, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]
        L.bella_kmers_get_tuples.argtypes = [H, vp, vp, vp, vp]
        L.bella_kmers_get_stats.argtypes = [H, ctypes.POINTER(ctypes.c_double)]
        _lib = L
    return _lib


class KmerCounter:
    def __init__(self, device=0):
        self._h = lib().bella_kmers_create(device)
        if not self._h:
            raise BellaKmersError("no usable sm_100 device for the k-mer counter (there is no CPU fallback)")

    def close(self):
        if self._h:
            lib().bella_kmers_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise BellaKmersError(f"bella_kmers error {rc}: {lib().bella_kmers_last_error(self._h).decode()}")

    def count(self, seqs, seq_off, k=17, lower=2, upper=8):
        """-> dict(t_kmer, t_read, t_pos, t_strand (bits, LSB first), n_kmers)"""
        seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
        seq_off = np.ascontiguousarray(seq_off, dtype=np.uint64)
        nk, nt = ctypes.c_uint64(0), ctypes.c_uint64(0)
        p = lambda a: ctypes.c_void_p(a.ctypes.data)  # noqa: E731
        self._check(lib().bella_kmers_count(self._h, p(seqs), p(seq_off), len(seq_off) - 1, k, lower, upper, ctypes.byref(nk), ctypes.byref(nt)))
        n = nt.value
        t_kmer = np.zeros(n, dtype=np.uint32); t_read = np.zeros(n, dtype=np.uint32); t_pos = np.zeros(n, dtype=np.uint16)
        bits = np.zeros((n + 7) // 8, dtype=np.uint8)
        self._check(lib().bella_kmers_get_tuples(self._h, p(t_kmer), p(t_read), p(t_pos), p(bits)))
        return {"t_kmer": t_kmer, "t_read": t_read, "t_pos": t_pos, "t_strand": bits, "n_kmers": nk.value}

    def stats(self):
        s = (ctypes.c_double * 3)()
        self._check(lib().bella_kmers_get_stats(self._h, s))
        return {"kernel_ms": s[0], "positions": int(s[1]), "launches": int(s[2])}
