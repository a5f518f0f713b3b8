"""Host-side Python binding of include/bella_xdrop.h over libbella_xdrop.so: the "next" row f1, batched gapped
X-drop seed-and-extend (the reference's alignSeqAn / alignLogan + PostAlignDecision, include/align.hpp:93-255,
include/overlap.hpp:415-462).  No CPU fallback: creating a handle without a B200 raises."""
import ctypes
import os

import numpy as np

from . import _build

_lib = None

FIELDS = ("score", "strand", "begH", "endH", "begV", "endV", "ov", "passed")
EXPORTS = ["bella_xdrop_create", "bella_xdrop_destroy", "bella_xdrop_last_error", "bella_xdrop_set_reads",
           "bella_xdrop_set_params", "bella_xdrop_set_shape", "bella_xdrop_align", "bella_xdrop_align_device",
           "bella_xdrop_align_csc_device",
           "bella_xdrop_get_stats", "bella_xdrop_stream", "bella_xdrop_sync"]


class BellaXdropError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        path = _build.LIB_XDROP
        path = _build.build_xdrop()                        # returns at once when the .so matches its sources
        L = ctypes.CDLL(path)
        vp, H = ctypes.c_void_p, ctypes.c_void_p
        L.bella_xdrop_create.argtypes = [ctypes.c_int]
        L.bella_xdrop_create.restype = H
        L.bella_xdrop_destroy.argtypes = [H]
        L.bella_xdrop_destroy.restype = None
        L.bella_xdrop_last_error.argtypes = [H]
        L.bella_xdrop_last_error.restype = ctypes.c_char_p
        L.bella_xdrop_set_reads.argtypes = [H, vp, vp, ctypes.c_uint32]
        L.bella_xdrop_set_params.argtypes = [H, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_int]
        L.bella_xdrop_set_shape.argtypes = [H, ctypes.c_int, ctypes.c_int]
        L.bella_xdrop_align.argtypes = [H, ctypes.c_uint64, vp, vp, vp, vp, vp]
        L.bella_xdrop_align_device.argtypes = [H, ctypes.c_uint64, vp, vp, vp, vp, vp]
        L.bella_xdrop_align_csc_device.argtypes = [H, ctypes.c_uint32, vp, ctypes.c_uint64, vp, vp, vp, vp]
        L.bella_xdrop_get_stats.argtypes = [H, ctypes.POINTER(ctypes.c_double)]
        L.bella_xdrop_stream.argtypes = [H]
        L.bella_xdrop_stream.restype = vp
        L.bella_xdrop_sync.argtypes = [H]
        _lib = L
    return _lib


def _p(a):
    return ctypes.c_void_p(a.ctypes.data)


class XdropAligner:
    """One handle per GPU; the reads stay on the device between batches."""

    def __init__(self, device=0):
        self._h = lib().bella_xdrop_create(device)
        if not self._h:
            raise BellaXdropError("no usable sm_100 device for the X-drop aligner (there is no CPU fallback)")
        self._keep = None

    def close(self):
        if self._h:
            lib().bella_xdrop_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise BellaXdropError(f"bella_xdrop error {rc}: {lib().bella_xdrop_last_error(self._h).decode()}")

    def set_reads(self, seqs, seq_off):
        seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
        seq_off = np.ascontiguousarray(seq_off, dtype=np.uint64)
        self._check(lib().bella_xdrop_set_reads(self._h, _p(seqs), _p(seq_off), len(seq_off) - 1))

    def set_params(self, kmer_len=17, xdrop=7, ratiophi=0.0, delta_chernoff=0.1, fixed_threshold=-1):
        self._check(lib().bella_xdrop_set_params(self._h, kmer_len, xdrop, ratiophi, delta_chernoff, fixed_threshold))

    def set_shape(self, lanes=-1, cells_per_lane=-1):
        self._check(lib().bella_xdrop_set_shape(self._h, lanes, cells_per_lane))

    def align(self, rows, cols, posH, posV):
        """-> int32 [n][8]: score, strand ('n' = 110 / 'c' = 99), begH, endH, begV, endV, ov, passed"""
        rows = np.ascontiguousarray(rows, dtype=np.uint32); cols = np.ascontiguousarray(cols, dtype=np.uint32)
        posH = np.ascontiguousarray(posH, dtype=np.uint16); posV = np.ascontiguousarray(posV, dtype=np.uint16)
        out = np.zeros((len(rows), len(FIELDS)), dtype=np.int32)
        self._check(lib().bella_xdrop_align(self._h, len(rows), _p(rows), _p(cols), _p(posH), _p(posV), _p(out)))
        return out

    def align_device(self, n_pairs, d_rows, d_cols, d_posH, d_posV, d_out):
        """device pointers (ints or torch tensors); asynchronous on the handle's stream -- call sync()"""
        ptr = lambda a: ctypes.c_void_p(a.data_ptr() if hasattr(a, "data_ptr") else int(a))  # noqa: E731
        self._check(lib().bella_xdrop_align_device(self._h, n_pairs, ptr(d_rows), ptr(d_cols), ptr(d_posH), ptr(d_posV), ptr(d_out)))

    def align_csc_device(self, n_cols, d_colptrC, n_pairs, d_rowids, d_posH, d_posV, d_out):
        """the overlap SpGEMM's device result (OverlapSpGEMM.result_device()) in, int32 [n_pairs][8] on the device out"""
        ptr = lambda a: ctypes.c_void_p(a.data_ptr() if hasattr(a, "data_ptr") else int(a))  # noqa: E731
        self._check(lib().bella_xdrop_align_csc_device(self._h, n_cols, ptr(d_colptrC), n_pairs, ptr(d_rowids), ptr(d_posH), ptr(d_posV), ptr(d_out)))

    def sync(self):
        self._check(lib().bella_xdrop_sync(self._h))

    def stats(self):
        s = (ctypes.c_double * 5)()
        self._check(lib().bella_xdrop_get_stats(self._h, s))
        return {"kernel_ms": s[0], "wide_extensions": int(s[1]), "launches": int(s[2]), "lanes": int(s[3]), "cells_per_lane": int(s[4])}
