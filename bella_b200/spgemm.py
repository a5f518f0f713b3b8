"""Host-side Python binding of the C-ABI (include/bella_b200.h) over libbella_b200.so.

`OverlapSpGEMM` mirrors the reference's HashSpGEMM structure (include/overlap.hpp:650-789):
set_inputs(A, B, reads-as-lengths+strand-bits, k, binSize) -> symbolic() == estimateFLOP +
estimateNNZ_Hash + prefix sums -> numeric(col_begin, col_end) == LocalSpGEMM + choose().
There is no CPU fallback: importing works anywhere, but creating a handle without a B200 raises."""
import ctypes
import os

import numpy as np

from . import _build

_lib = None

ERRORS = {-6: "multi-GPU exchange buffer too small", -1: "bad argument", -2: "CUDA error / no usable sm_100 device (there is no CPU fallback)",
          -3: "device out of memory", -4: "count exceeds the reference's index types", -5: "internal device error"}


class CsrView(ctypes.Structure):
    """bella_csr_view: mirror of CSR<uint32_t, unsigned short> (reference include/common/CSR.h:15-67)"""
    _fields_ = [("rows", ctypes.c_uint32), ("cols", ctypes.c_uint32), ("nnz", ctypes.c_uint32),
                ("rowptr", ctypes.c_void_p), ("colids", ctypes.c_void_p), ("values", ctypes.c_void_p), ("zerobased", ctypes.c_int)]


class CscView(ctypes.Structure):
    _fields_ = [("rows", ctypes.c_uint32), ("cols", ctypes.c_uint32), ("nnz", ctypes.c_uint32),
                ("colptr", ctypes.c_void_p), ("rowids", ctypes.c_void_p), ("values", ctypes.c_void_p)]


def lib():
    global _lib
    if _lib is None:
        path = os.environ.get("BELLA_B200_LIB")          # A/B runs of two builds of the same library (profiling only)
        if not path:
            path = _build.build_cuda()                    # returns at once when the in-tree .so is newer than its sources
        L = ctypes.CDLL(path)
        vp, H = ctypes.c_void_p, ctypes.c_void_p
        L.bella_b200_create.argtypes = [ctypes.POINTER(H), ctypes.c_int]
        L.bella_b200_destroy.argtypes = [H]
        L.bella_b200_last_error.argtypes = [H]
        L.bella_b200_last_error.restype = ctypes.c_char_p
        for f in ("bella_b200_set_inputs", "bella_b200_set_inputs_device"):
            getattr(L, f).argtypes = [H, ctypes.POINTER(CscView), ctypes.POINTER(CscView), vp, vp, vp, ctypes.c_uint16, ctypes.c_uint16]
        L.bella_b200_set_inputs_csr.argtypes = [H, ctypes.POINTER(CsrView), vp, vp, ctypes.c_uint16, ctypes.c_uint16]
        L.bella_b200_set_column_range.argtypes = [H, ctypes.c_uint32, ctypes.c_uint32]
        L.bella_b200_symbolic.argtypes = [H, ctypes.POINTER(ctypes.c_uint64), vp, vp]
        L.bella_b200_numeric.argtypes = [H, ctypes.c_uint32, ctypes.c_uint32, vp, vp, vp, vp]
        L.bella_b200_set_output_buffers.argtypes = [H, vp, vp, vp, vp, ctypes.c_uint64]
        L.bella_b200_numeric_aux.argtypes = [H, ctypes.c_uint32, ctypes.c_uint32, vp, vp, vp]
        L.bella_b200_numeric_device.argtypes = [H]
        L.bella_b200_n_unpinned.argtypes = [H, ctypes.POINTER(ctypes.c_uint64)]
        L.bella_b200_get_flops.argtypes = [H, ctypes.POINTER(ctypes.c_uint64)]
        L.bella_b200_result_device.argtypes = [H] + [ctypes.POINTER(vp)] * 5 + [ctypes.POINTER(ctypes.c_uint64)]
        L.bella_b200_run_resident.argtypes = [H, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]
        L.bella_b200_get_timings.argtypes = [H, ctypes.POINTER(ctypes.c_float)]
        L.bella_b200_stream.argtypes = [H]
        L.bella_b200_stream.restype = vp
        L.bella_b200_set_stream.argtypes = [H, vp]
        L.bella_b200_mg_transpose.argtypes = [H, ctypes.c_uint32, ctypes.c_uint32, vp]
        L.bella_b200_mg_scatter.argtypes = [H, vp, vp]
        L.bella_b200_mg_finish.argtypes = [H, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, vp, vp, vp, vp]
        L.bella_b200_get_colptr.argtypes = [H, vp]
        pp = ctypes.POINTER(vp)
        L.bella_b200_mg_geometry.argtypes = [ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int, vp]
        L.bella_b200_mg_route_push.argtypes = [H, ctypes.c_uint32, ctypes.c_uint32, vp, vp, vp, ctypes.c_uint32, vp, ctypes.c_int, ctypes.c_int, pp, pp, pp]
        L.bella_b200_mg_transpose_coarse.argtypes = [H, ctypes.c_uint32, ctypes.c_uint32, vp, ctypes.c_int, vp, vp, vp, ctypes.c_uint64, vp]
        L.bella_b200_mg_post.argtypes = [H, ctypes.c_uint64, vp, ctypes.c_int, ctypes.c_uint64, pp]
        L.bella_b200_mg_exchange.argtypes = [H, ctypes.c_int, ctypes.c_int, vp, vp, pp, ctypes.c_int, vp, vp, ctypes.c_uint64, ctypes.c_uint64, vp, vp, vp, vp, vp, pp]
        L.bella_b200_set_inputs_tuples.argtypes = [H, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint64, vp, vp, vp, vp, vp, ctypes.c_uint16, ctypes.c_uint16]
        L.bella_b200_get_B.argtypes = [H, ctypes.POINTER(ctypes.c_uint32), vp, vp, vp, vp, ctypes.POINTER(ctypes.c_float)]
        _lib = L
    return _lib


EXPORTS = ["bella_b200_set_output_buffers", "bella_b200_get_flops", "bella_b200_mg_geometry", "bella_b200_mg_route_push", "bella_b200_mg_transpose_coarse", "bella_b200_mg_post", "bella_b200_mg_exchange",
           "bella_b200_set_inputs_csr", "bella_b200_n_unpinned", "bella_b200_create", "bella_b200_destroy", "bella_b200_last_error", "bella_b200_set_inputs",
           "bella_b200_set_inputs_device", "bella_b200_set_column_range", "bella_b200_symbolic",
           "bella_b200_numeric", "bella_b200_numeric_aux", "bella_b200_numeric_device",
           "bella_b200_result_device", "bella_b200_run_resident", "bella_b200_get_timings", "bella_b200_stream",
           "bella_b200_set_stream", "bella_b200_mg_transpose", "bella_b200_mg_scatter", "bella_b200_mg_finish",
           "bella_b200_get_colptr", "bella_b200_set_inputs_tuples", "bella_b200_get_B"]


class BellaB200Error(RuntimeError):
    pass


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return ctypes.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):          # torch tensor (device-resident inputs)
        return ctypes.c_void_p(a.data_ptr())
    return ctypes.c_void_p(int(a))


class OverlapSpGEMM:
    """One handle = one GPU (one process per GPU in the multi-GPU layout)."""

    def __init__(self, device=0):
        self._L = lib()
        self._h = ctypes.c_void_p()
        rc = self._L.bella_b200_create(ctypes.byref(self._h), device)
        if rc != 0:
            self._h = None
            raise BellaB200Error(f"bella_b200_create(device={device}) failed: {ERRORS.get(rc, rc)}")
        self.n = self.m = 0
        self.lo = self.hi = 0
        self.flops = 0
        self.colptrC = None
        self._keep = None
        self._pinned = {}

    def _host(self, name, dtype, count, pinned):
        """Output buffer: a fresh numpy array, or (pinned=True) a reused page-locked one -- device->host
        copies into pageable memory run at a fraction of the PCIe rate."""
        count = max(int(count), 1)
        if not pinned:
            return np.zeros(count, dtype=dtype)
        import torch
        tdt = {np.uint32: torch.int32, np.uint16: torch.int16}[dtype]
        t = self._pinned.get(name)
        if t is None or t.numel() < count or t.dtype != tdt:
            t = torch.empty(int(count * 1.25) + 16, dtype=tdt).pin_memory()
            self._pinned[name] = t
        return t.numpy()[:count].view(dtype)

    def _check(self, rc, what):
        if rc != 0:
            msg = self._L.bella_b200_last_error(self._h)
            e = BellaB200Error(f"{what}: {ERRORS.get(rc, rc)}: {msg.decode() if msg else ''}")
            e.code = rc
            raise e

    def close(self):
        if self._h:
            self._L.bella_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _views(self, n, m, nnz, A, B):
        vB = CscView(m, n, nnz, _ptr(B[0]), _ptr(B[1]), _ptr(B[2]))
        vA = CscView(n, m, nnz, _ptr(A[0]), _ptr(A[1]), _ptr(A[2])) if A is not None else None
        return vA, vB

    def set_inputs(self, inp, with_A=False):
        """inp: frontend.OverlapInputs (host numpy arrays).  A (== Bᵀ, what BELLA always passes) is only
        checked for shape: the device derives it from B, so its arrays are never copied."""
        A = (inp.A_colptr, inp.A_rowids, inp.A_values) if with_A else None
        vA, vB = self._views(inp.n_reads, inp.n_kmers, inp.nnz, A, (inp.B_colptr, inp.B_rowids, inp.B_values))
        self._keep = inp
        rc = self._L.bella_b200_set_inputs(self._h, ctypes.byref(vA) if vA is not None else None, ctypes.byref(vB),
                                           _ptr(inp.read_len), _ptr(inp.A_strand) if with_A else None, _ptr(inp.B_strand),
                                           inp.kmer_size, inp.bin_size)
        self._check(rc, "bella_b200_set_inputs")
        self.n, self.m, self.lo, self.hi = inp.n_reads, inp.n_kmers, 0, inp.n_reads

    def set_inputs_device(self, n, m, nnz, B, read_len, strand_B, kmer_size, bin_size, A=None, strand_A=None):
        """Device-resident inputs (torch CUDA tensors or raw device pointers); nothing is copied."""
        vA, vB = self._views(n, m, nnz, A, B)
        self._keep = (B, read_len, strand_B, A, strand_A)
        rc = self._L.bella_b200_set_inputs_device(self._h, ctypes.byref(vA) if vA is not None else None, ctypes.byref(vB),
                                                  _ptr(read_len), _ptr(strand_A), _ptr(strand_B), kmer_size, bin_size)
        self._check(rc, "bella_b200_set_inputs_device")
        self.n, self.m, self.lo, self.hi = n, m, 0, n

    def set_stream(self, cuda_stream_ptr):
        """Run on the caller's CUDA stream (e.g. torch.cuda.current_stream().cuda_stream)."""
        self._check(self._L.bella_b200_set_stream(self._h, ctypes.c_void_p(cuda_stream_ptr)), "bella_b200_set_stream")

    def set_column_range(self, lo, hi):
        self._check(self._L.bella_b200_set_column_range(self._h, lo, hi), "bella_b200_set_column_range")
        self.lo, self.hi = lo, hi

    def symbolic(self, want_flopC=True, pinned=False):
        """-> (flops, flopC[ncols] or None, colptrC[ncols+1])"""
        nc = self.hi - self.lo
        flops = ctypes.c_uint64(0)
        flopC = self._host("flopC", np.uint32, nc, pinned)[:nc] if want_flopC else None
        colptrC = self._host("colptrC", np.uint32, nc + 1, pinned)
        self._check(self._L.bella_b200_symbolic(self._h, ctypes.byref(flops), _ptr(flopC), _ptr(colptrC)), "bella_b200_symbolic")
        self.flops, self.colptrC = flops.value, colptrC
        return flops.value, flopC, colptrC

    def set_output_buffers(self, capacity):
        """Register page-locked host buffers of `capacity` entries for the whole result BEFORE symbolic(): the results of a column
        range are then copied out while the later ranges still fold, and numeric(pinned=True) finds them delivered
        (include/bella_b200.h bella_b200_set_output_buffers).  capacity = 0 switches it off."""
        if not capacity:
            self._check(self._L.bella_b200_set_output_buffers(self._h, None, None, None, None, 0), "bella_b200_set_output_buffers")
            return
        for name, dt in (("rows", np.uint32), ("cnt", np.uint16), ("pH", np.uint16), ("pV", np.uint16)):      # the buffers numeric(pinned=True) hands out
            self._host(name, dt, capacity, True)
        self._registered = {k: self._pinned[k] for k in ("rows", "cnt", "pH", "pV")}      # kept alive while the library may write to them
        p = {k: ctypes.c_void_p(t.data_ptr()) for k, t in self._registered.items()}
        cap = min(t.numel() for t in self._registered.values())
        self._check(self._L.bella_b200_set_output_buffers(self._h, p["rows"], p["cnt"], p["pH"], p["pV"], cap), "bella_b200_set_output_buffers")

    def numeric(self, col_begin=None, col_end=None, aux=False, pinned=False):
        """-> (rowids, count, posH, posV[, aux(nnz,3)]) for global columns [col_begin, col_end).
        pinned=True returns views of page-locked buffers that the next call overwrites."""
        c0 = self.lo if col_begin is None else col_begin
        c1 = self.hi if col_end is None else col_end
        z = int(self.colptrC[c1 - self.lo]) - int(self.colptrC[c0 - self.lo])
        rows = self._host("rows", np.uint32, z, pinned)
        cnt = self._host("cnt", np.uint16, z, pinned)
        pH = self._host("pH", np.uint16, z, pinned)
        pV = self._host("pV", np.uint16, z, pinned)
        self._check(self._L.bella_b200_numeric(self._h, c0, c1, _ptr(rows), _ptr(cnt), _ptr(pH), _ptr(pV)), "bella_b200_numeric")
        out = (rows[:z], cnt[:z], pH[:z], pV[:z])
        if aux:
            a = np.zeros((3, max(z, 1)), dtype=np.uint16)
            self._check(self._L.bella_b200_numeric_aux(self._h, c0, c1, _ptr(a[0]), _ptr(a[1]), _ptr(a[2])), "bella_b200_numeric_aux")
            out = out + (np.ascontiguousarray(a[:, :z].T),)
        return out

    # ---- matrix construction on the device ("next" row f2) ----
    def set_inputs_tuples(self, n_kmers, n_reads, t_kmer, t_read, t_pos, t_strand, read_len, kmer_size=17, bin_size=500):
        """Host tuples (k-mer id, read id, position; one strand bit per tuple) -> B on the device, in the reference's order."""
        self._keep = (t_kmer, t_read, t_pos, t_strand, read_len)
        self._check(self._L.bella_b200_set_inputs_tuples(self._h, n_kmers, n_reads, len(t_kmer), _ptr(t_kmer), _ptr(t_read), _ptr(t_pos),
                                                         _ptr(t_strand), _ptr(read_len), kmer_size, bin_size), "bella_b200_set_inputs_tuples")
        self.n, self.m, self.lo, self.hi = n_reads, n_kmers, 0, n_reads

    def get_B(self):
        """-> (colptr, rowids, values, strand bits, build_ms) of the handle's B, on the host"""
        nnz, ms = ctypes.c_uint32(0), ctypes.c_float(0)
        self._check(self._L.bella_b200_get_B(self._h, ctypes.byref(nnz), None, None, None, None, ctypes.byref(ms)), "bella_b200_get_B")
        z = nnz.value
        colptr = np.zeros(self.n + 1, dtype=np.uint32)
        rows, vals = np.zeros(max(z, 1), dtype=np.uint32), np.zeros(max(z, 1), dtype=np.uint16)
        strand = np.zeros((z + 7) // 8 + 8, dtype=np.uint8)
        self._check(self._L.bella_b200_get_B(self._h, None, _ptr(colptr), _ptr(rows), _ptr(vals), _ptr(strand), None), "bella_b200_get_B")
        return colptr, rows[:z], vals[:z], strand, ms.value

    # ---- multi-GPU stages (device tensors / pointers; bella_b200/distributed.py runs the collectives between them) ----
    def mg_transpose(self, kmer_lo, kmer_hi, cnt_local):
        self._check(self._L.bella_b200_mg_transpose(self._h, kmer_lo, kmer_hi, _ptr(cnt_local)), "bella_b200_mg_transpose")

    @staticmethod
    def _pp(ptrs):
        return (ctypes.c_void_p * len(ptrs))(*[ctypes.c_void_p(int(p)) for p in ptrs])

    def mg_route_push(self, read_lo, read_hi, colptr_global_ptr, rowids, values, n_kmers, geom, world, me, peer_E, peer_K, peer_cnt):
        self._check(self._L.bella_b200_mg_route_push(self._h, read_lo, read_hi, _ptr(colptr_global_ptr), _ptr(rowids), _ptr(values), n_kmers, _ptr(geom),
                                                     world, me, self._pp(peer_E), self._pp(peer_K), self._pp(peer_cnt)), "bella_b200_mg_route_push")

    def mg_transpose_coarse(self, kmer_lo, kmer_hi, geom, world, E, K, cnt, nnz_cap, cnt_local):
        self._check(self._L.bella_b200_mg_transpose_coarse(self._h, kmer_lo, kmer_hi, _ptr(geom), world, _ptr(E), _ptr(K), _ptr(cnt), nnz_cap, _ptr(cnt_local)),
                    "bella_b200_mg_transpose_coarse")

    def mg_post(self, count, src, world, at, peer_dst):
        self._check(self._L.bella_b200_mg_post(self._h, count, _ptr(src), world, at, self._pp(peer_dst)), "bella_b200_mg_post")

    def mg_exchange(self, world, me, cuts, cnt_local, peer_counts_all, phase, counts_all=None, scan=None, cap_recv=0, cap_send=0, sendoff=None, segoff=None,
                    recvbase=None, push=None, sendbuf=None, peer_recv=None):
        self._check(self._L.bella_b200_mg_exchange(self._h, world, me, _ptr(cuts), _ptr(cnt_local), self._pp(peer_counts_all), phase, _ptr(counts_all),
                                                   _ptr(scan), cap_recv, cap_send, _ptr(sendoff), _ptr(segoff), _ptr(recvbase), _ptr(push), _ptr(sendbuf),
                                                   self._pp(peer_recv if peer_recv is not None else peer_counts_all)), "bella_b200_mg_exchange")

    def mg_scatter(self, sendoff, sendbuf):
        self._check(self._L.bella_b200_mg_scatter(self._h, _ptr(sendoff), _ptr(sendbuf)), "bella_b200_mg_scatter")

    def mg_finish(self, col_lo, col_hi, world, counts_all, segoff, recvbase, recv):
        self._check(self._L.bella_b200_mg_finish(self._h, col_lo, col_hi, world, _ptr(counts_all), _ptr(segoff), _ptr(recvbase), _ptr(recv)),
                    "bella_b200_mg_finish")
        self.lo, self.hi = col_lo, col_hi

    def get_colptr(self, pinned=False):
        colptrC = self._host("colptrC", np.uint32, self.hi - self.lo + 1, pinned)
        self._check(self._L.bella_b200_get_colptr(self._h, _ptr(colptrC)), "bella_b200_get_colptr")
        self.colptrC = colptrC
        return colptrC

    def result_flops(self):
        v = ctypes.c_uint64(0)
        self._check(self._L.bella_b200_get_flops(self._h, ctypes.byref(v)), "bella_b200_get_flops")
        return int(v.value)

    def result_nnz(self):
        z = ctypes.c_uint64(0)
        self._check(self._L.bella_b200_result_device(self._h, None, None, None, None, None, ctypes.byref(z)), "bella_b200_result_device")
        return z.value

    def result_device(self):
        """device pointers of the whole result after numeric_device() / run_resident():
        -> dict(colptrC, rowids, count, posH, posV: int addresses; nnz).  Valid until the next pass on this handle."""
        ptrs = [ctypes.c_void_p(0) for _ in range(5)]
        z = ctypes.c_uint64(0)
        self._check(self._L.bella_b200_result_device(self._h, *[ctypes.byref(p) for p in ptrs], ctypes.byref(z)), "bella_b200_result_device")
        out = {k: (p.value or 0) for k, p in zip(("colptrC", "rowids", "count", "posH", "posV"), ptrs)}
        out["nnz"] = z.value
        return out

    def n_unpinned(self):
        """pairs with more than 16 bins (choose()'s tie order is unpinned there, common.h:162-170)"""
        v = ctypes.c_uint64(0)
        self._check(self._L.bella_b200_n_unpinned(self._h, ctypes.byref(v)), "bella_b200_n_unpinned")
        return int(v.value)

    def set_inputs_csr(self, inp):
        """The CSR surface: A row-major == B's CSC arrays (include/bella_b200.h bella_csr_view)."""
        v = CsrView(inp.n_reads, inp.n_kmers, inp.nnz, _ptr(inp.B_colptr), _ptr(inp.B_rowids), _ptr(inp.B_values), 1)
        self._keep = inp
        self._check(self._L.bella_b200_set_inputs_csr(self._h, ctypes.byref(v), _ptr(inp.read_len), _ptr(inp.B_strand), inp.kmer_size, inp.bin_size),
                    "bella_b200_set_inputs_csr")
        self.n, self.m, self.lo, self.hi = inp.n_reads, inp.n_kmers, 0, inp.n_reads

    def numeric_device(self):
        self._check(self._L.bella_b200_numeric_device(self._h), "bella_b200_numeric_device")

    def run_resident(self):
        """layout + symbolic + numeric on device-resident inputs, no result copy. -> (nnzC, flops)"""
        z, f = ctypes.c_uint64(0), ctypes.c_uint64(0)
        self._check(self._L.bella_b200_run_resident(self._h, ctypes.byref(z), ctypes.byref(f)), "bella_b200_run_resident")
        return z.value, f.value

    def timings(self):
        t = (ctypes.c_float * 8)()
        self._L.bella_b200_get_timings(self._h, t)
        return {"partition_ms": t[0], "bucket_plan_ms": t[6], "transpose_ms": t[0] + t[6], "scatter_ms": t[7], "group_fold_ms": t[1],
                "output_ms": t[2], "h2d_ms": t[3], "d2h_ms": t[4], "launches": int(t[5])}


def overlap_spgemm(inp, device=0, with_A=False, aux=False):
    """Convenience: whole pass through the C-ABI with host buffers. -> dict of numpy arrays"""
    g = OverlapSpGEMM(device)
    try:
        g.set_inputs(inp, with_A=with_A)
        flops, flopC, colptrC = g.symbolic()
        res = g.numeric(aux=aux)
        return {"flops": flops, "flopC": flopC, "colptrC": colptrC, "rowids": res[0], "count": res[1],
                "posH": res[2], "posV": res[3], "aux": res[4] if aux else None, "timings": g.timings()}
    finally:
        g.close()
