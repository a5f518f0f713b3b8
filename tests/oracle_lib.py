"""TEST INFRASTRUCTURE: ctypes loaders for the CPU oracle (oracle/bella_oracle.c) and, when it has
been built, the unmodified reference (oracle/_ref/libbella_ref.so, see oracle/ref_driver.cpp)."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "libbella_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libbella_ref.so")

_oracle = None
_ref = None
u32p, u16p, u8p = (ctypes.c_void_p,) * 3


def oracle():
    global _oracle
    if _oracle is None:
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True)
        L = ctypes.CDLL(ORACLE_SO)
        L.oracle_local_spgemm.restype = ctypes.c_int64
        L.oracle_build_csc.restype = ctypes.c_int64
        _oracle = L
    return _oracle


def have_ref():
    return os.path.exists(REF_SO)


def ref():
    global _ref
    if _ref is None:
        L = ctypes.CDLL(REF_SO)
        L.bella_ref_create.restype = ctypes.c_void_p
        L.bella_ref_flops.restype = ctypes.c_uint64
        L.bella_ref_flopC.restype = ctypes.c_void_p
        L.bella_ref_colptrC.restype = ctypes.c_void_p
        L.bella_ref_cols.restype = ctypes.c_uint32
        L.bella_ref_build.restype = ctypes.c_int64
        _ref = L
    return _ref


def _p(a):
    return ctypes.c_void_p(a.ctypes.data)


def sort_tuples(cols, rows, *vals):
    order = np.lexsort((rows, cols))
    return (cols[order], rows[order]) + tuple(v[order] for v in vals)


class Result:
    """Canonical SpGEMM result: CSC of C with rows ascending in each column."""

    def __init__(self, flopC, colptrC, rowids, count, posH, posV, aux=None, times=None, unpinned=0):
        self.flopC, self.colptrC, self.rowids = flopC, colptrC, rowids
        self.count, self.posH, self.posV, self.aux = count, posH, posV, aux
        self.times, self.unpinned = times, unpinned

    @property
    def nnz(self):
        return int(self.colptrC[-1])


def oracle_spgemm(inp, ncols=None, nthreads=0, want_aux=True):
    """oracle/bella_oracle.c on OverlapInputs; ncols restricts to output columns [0, ncols)."""
    L = oracle()
    if nthreads:
        L.oracle_set_threads(nthreads)
    n = inp.n_reads if ncols is None else min(ncols, inp.n_reads)
    flopC = np.zeros(n, dtype=np.uint32)
    nnzC = np.zeros(n, dtype=np.uint32)
    colptrC = np.zeros(n + 1, dtype=np.uint32)
    import time
    t0 = time.perf_counter()
    L.oracle_estimate_flop(ctypes.c_uint32(n), _p(inp.A_colptr), _p(inp.A_rowids), _p(inp.B_colptr), _p(inp.B_rowids), _p(flopC))
    t1 = time.perf_counter()
    L.oracle_estimate_nnz(ctypes.c_uint32(n), _p(inp.A_colptr), _p(inp.A_rowids), _p(inp.B_colptr), _p(inp.B_rowids), _p(flopC), _p(nnzC))
    L.oracle_prefixsum(_p(nnzC), ctypes.c_uint32(n), _p(colptrC))
    t2 = time.perf_counter()
    Z = int(colptrC[n])
    rows = np.zeros(max(Z, 1), dtype=np.uint32)
    cnt = np.zeros(max(Z, 1), dtype=np.uint16)
    pH = np.zeros(max(Z, 1), dtype=np.uint16)
    pV = np.zeros(max(Z, 1), dtype=np.uint16)
    aux = np.zeros((max(Z, 1), 3), dtype=np.uint16) if want_aux else None
    rc = L.oracle_local_spgemm(ctypes.c_uint32(0), ctypes.c_uint32(n),
                               _p(inp.A_colptr), _p(inp.A_rowids), _p(inp.A_values), _p(inp.A_strand),
                               _p(inp.B_colptr), _p(inp.B_rowids), _p(inp.B_values), _p(inp.B_strand),
                               _p(inp.read_len), ctypes.c_uint16(inp.kmer_size), ctypes.c_uint16(inp.bin_size),
                               _p(colptrC), _p(rows), _p(cnt), _p(pH), _p(pV), _p(aux) if want_aux else None)
    t3 = time.perf_counter()
    if rc < 0:
        raise RuntimeError("oracle_local_spgemm failed")
    return Result(flopC, colptrC, rows[:Z], cnt[:Z], pH[:Z], pV[:Z], aux[:Z] if want_aux else None,
                  times=(t1 - t0, t2 - t1, t3 - t2), unpinned=int(rc))


def ref_spgemm(inp, ncols=None, nthreads=0, want_aux=True):
    """The UNMODIFIED reference (oracle/_ref) on the same arrays; needs inp.seqs."""
    L = ref()
    assert inp.seqs is not None, "the reference multiply compares read substrings: sequences required"
    h = L.bella_ref_create(ctypes.c_uint32(inp.n_reads), ctypes.c_uint32(inp.n_kmers),
                           ctypes.c_uint32(inp.nnz), _p(inp.A_colptr), _p(inp.A_rowids), _p(inp.A_values),
                           ctypes.c_uint32(inp.nnz), _p(inp.B_colptr), _p(inp.B_rowids), _p(inp.B_values),
                           _p(inp.seqs), _p(inp.seq_off), ctypes.c_uint16(inp.kmer_size), ctypes.c_uint16(inp.bin_size),
                           ctypes.c_uint32(0 if ncols is None else ncols), ctypes.c_int(nthreads))
    if not h:
        raise RuntimeError("bella_ref_create: empty matrix")
    h = ctypes.c_void_p(h)
    try:
        n = L.bella_ref_cols(h)
        flopC = np.ctypeslib.as_array(ctypes.cast(L.bella_ref_flopC(h), ctypes.POINTER(ctypes.c_uint32)), (n,)).copy()
        colptrC = np.ctypeslib.as_array(ctypes.cast(L.bella_ref_colptrC(h), ctypes.POINTER(ctypes.c_uint32)), (n + 1,)).copy()
        Z = int(colptrC[n])
        rows = np.zeros(max(Z, 1), dtype=np.uint32)
        cnt = np.zeros(max(Z, 1), dtype=np.uint16)
        pH = np.zeros(max(Z, 1), dtype=np.uint16)
        pV = np.zeros(max(Z, 1), dtype=np.uint16)
        aux = np.zeros((max(Z, 1), 3), dtype=np.uint16) if want_aux else None
        rc = L.bella_ref_numeric(h, ctypes.c_uint32(0), ctypes.c_uint32(n), _p(rows), _p(cnt), _p(pH), _p(pV),
                                 _p(aux) if want_aux else None)
        if rc != 0:
            raise RuntimeError(f"bella_ref_numeric failed: {rc}")
        t = (ctypes.c_double * 3)()
        L.bella_ref_times(h, t)
        return Result(flopC, colptrC, rows[:Z], cnt[:Z], pH[:Z], pV[:Z], aux[:Z] if want_aux else None, times=tuple(t))
    finally:
        L.bella_ref_destroy(h)


def assert_same(a, b, aux=True):
    np.testing.assert_array_equal(a.flopC, b.flopC)
    np.testing.assert_array_equal(a.colptrC, b.colptrC)
    np.testing.assert_array_equal(a.rowids, b.rowids)
    np.testing.assert_array_equal(a.count, b.count)
    np.testing.assert_array_equal(a.posH, b.posH)
    np.testing.assert_array_equal(a.posV, b.posV)
    if aux and a.aux is not None and b.aux is not None:
        np.testing.assert_array_equal(a.aux, b.aux)


# ---- "next" row f1: gapped X-drop seed-and-extend ------------------------------------------------

def _align(fn, inp, rows, cols, posH, posV, xdrop):
    rows = np.ascontiguousarray(rows, dtype=np.uint32); cols = np.ascontiguousarray(cols, dtype=np.uint32)
    posH = np.ascontiguousarray(posH, dtype=np.uint16); posV = np.ascontiguousarray(posV, dtype=np.uint16)
    out = np.zeros((len(rows), 6), dtype=np.int32)
    rc = fn(ctypes.c_uint64(len(rows)), _p(rows), _p(cols), _p(posH), _p(posV), _p(inp.seqs), _p(inp.seq_off),
            ctypes.c_int(inp.kmer_size), ctypes.c_int(xdrop), _p(out))
    if rc != 0:
        raise RuntimeError(f"alignment failed: {rc}")
    return out


def oracle_align(inp, rows, cols, posH, posV, xdrop=7):
    """oracle/bella_oracle.c oracle_xdrop_align -> int32 [n][6] = score, strand char, begH, endH, begV, endV"""
    return _align(oracle().oracle_xdrop_align, inp, rows, cols, posH, posV, xdrop)


def ref_align(inp, rows, cols, posH, posV, xdrop=7):
    """the reference's alignSeqAn (oracle/_ref) on the same pairs"""
    return _align(ref().bella_ref_align, inp, rows, cols, posH, posV, xdrop)


def ref_align_xavier(inp, rows, cols, posH, posV, xdrop=7):
    """the aligner the reference's default CPU build calls (xavierAlign, a different fixed-band algorithm); informational"""
    return _align(ref().bella_ref_align_xavier, inp, rows, cols, posH, posV, xdrop)


def oracle_align_post(inp, rows, cols, posH, posV, xdrop=7, ratiophi=0.5, delta=0.1, fixed_threshold=-1):
    """oracle_xdrop_align + oracle_xdrop_post -> int32 [n][8] = the six alignment fields, ov, passed"""
    a = oracle_align(inp, rows, cols, posH, posV, xdrop)
    rows = np.ascontiguousarray(rows, dtype=np.uint32); cols = np.ascontiguousarray(cols, dtype=np.uint32)
    post = np.zeros((len(rows), 2), dtype=np.int32)
    rc = oracle().oracle_xdrop_post(ctypes.c_uint64(len(rows)), _p(rows), _p(cols), _p(inp.seq_off), _p(a),
                                    ctypes.c_double(ratiophi), ctypes.c_double(delta), ctypes.c_int(fixed_threshold), _p(post))
    if rc != 0:
        raise RuntimeError(f"post-alignment decision failed: {rc}")
    return np.concatenate([a, post], axis=1)


def ref_align_post(inp, rows, cols, posH, posV, xdrop=7, ratiophi=0.5, delta=0.1, fixed_threshold=-1):
    """the reference's alignSeqAn + PostAlignDecision -> int32 [n][8]; column 6 (ov) is -1 where the pair is rejected,
    because the reference only prints it for accepted pairs"""
    rows = np.ascontiguousarray(rows, dtype=np.uint32); cols = np.ascontiguousarray(cols, dtype=np.uint32)
    posH = np.ascontiguousarray(posH, dtype=np.uint16); posV = np.ascontiguousarray(posV, dtype=np.uint16)
    out = np.zeros((len(rows), 8), dtype=np.int32)
    rc = ref().bella_ref_align_post(ctypes.c_uint64(len(rows)), _p(rows), _p(cols), _p(posH), _p(posV), _p(inp.seqs), _p(inp.seq_off),
                                    ctypes.c_int(inp.kmer_size), ctypes.c_int(xdrop), ctypes.c_double(ratiophi), ctypes.c_double(delta),
                                    ctypes.c_int(fixed_threshold), _p(out))
    if rc != 0:
        raise RuntimeError(f"reference post-alignment failed: {rc}")
    return out


LOGAN = os.path.join(ROOT, "oracle", "_ref", "libbella_logan.so")


def have_logan():
    return os.path.exists(LOGAN)


def logan_align(inp, rows, cols, posH, posV, xdrop=7):
    """the reference's CUDA aligner (LOGAN, recompiled for sm_100a from /root/reference; needs a GPU)
    -> (int32 [n][6] like ref_align, seconds spent in extendSeedL)"""
    L = ctypes.CDLL(LOGAN)
    rows = np.ascontiguousarray(rows, dtype=np.uint32); cols = np.ascontiguousarray(cols, dtype=np.uint32)
    posH = np.ascontiguousarray(posH, dtype=np.uint16); posV = np.ascontiguousarray(posV, dtype=np.uint16)
    out = np.zeros((len(rows), 6), dtype=np.int32)
    sec = ctypes.c_double(0.0)
    rc = L.bella_logan_align(ctypes.c_uint64(len(rows)), _p(rows), _p(cols), _p(posH), _p(posV), _p(inp.seqs), _p(inp.seq_off),
                             ctypes.c_int(inp.kmer_size), ctypes.c_int(xdrop), _p(out), ctypes.byref(sec))
    if rc != 0:
        raise RuntimeError(f"LOGAN failed: {rc}")
    return out, sec.value


# ---- "next" row f3: reliable k-mer selection -------------------------------------------------------

def _occurrences(fn, head, inp, k, lower, upper):
    total = int(len(inp.seqs))
    out_read = np.zeros(total, dtype=np.uint32); out_pos = np.zeros(total, dtype=np.uint16)
    n_out, n_kmers = ctypes.c_uint64(0), ctypes.c_uint64(0)
    kw = [ctypes.c_int(x) for x in k] if isinstance(k, tuple) else [ctypes.c_int(k)]          # (k, window) for the minimizer variants
    rc = fn(*head, ctypes.c_uint32(inp.n_reads), _p(inp.seqs), _p(inp.seq_off), *kw, ctypes.c_int(lower), ctypes.c_int(upper),
            _p(out_read), _p(out_pos), ctypes.c_uint64(total), ctypes.byref(n_out), ctypes.byref(n_kmers))
    if rc != 0:
        raise RuntimeError(f"reliable k-mer selection failed: {rc}")
    return out_read[:n_out.value].copy(), out_pos[:n_out.value].copy(), n_kmers.value


def oracle_reliable_occurrences(inp, k, lower, upper):
    """oracle_reliable_occurrences -> (read u32[], pos u16[], number of distinct reliable k-mers)"""
    return _occurrences(oracle().oracle_reliable_occurrences, (), inp, k, lower, upper)


def write_fastq(inp, path):
    with open(path, "wb") as f:
        for r in range(inp.n_reads):
            s = inp.seqs[int(inp.seq_off[r]):int(inp.seq_off[r + 1])].tobytes()
            f.write(b"@read%d\n" % r + s + b"\n+\n" + b"I" * len(s) + b"\n")
    return os.path.getsize(path)


def oracle_minimizers(seq_bytes, k, window):
    """oracle_minimizers: sampled k-mer positions of one read (numpy int32)"""
    out = np.zeros(max(len(seq_bytes), 1), dtype=np.int32)
    n = oracle().oracle_minimizers(ctypes.c_char_p(bytes(seq_bytes)), ctypes.c_int(len(seq_bytes)), ctypes.c_int(k), ctypes.c_int(window), _p(out),
                                   ctypes.c_int(len(out)))
    assert n >= 0
    return out[:n].copy()


def ref_minimizers(seq_bytes, k, window):
    """the reference's getMinimizers (include/minimizer.hpp:49-77) on the read's Kmer objects"""
    out = np.zeros(max(len(seq_bytes), 1), dtype=np.int32)
    n = ref().bella_ref_minimizers(ctypes.c_char_p(bytes(seq_bytes)), ctypes.c_int(len(seq_bytes)), ctypes.c_int(k), ctypes.c_int(window), _p(out),
                                   ctypes.c_int(len(out)))
    assert n >= 0
    return out[:n].copy()


def oracle_minimizer_occurrences(inp, k, window, lower, upper):
    return _occurrences(oracle().oracle_minimizer_occurrences, (), inp, (k, window), lower, upper)


def ref_minimizer_occurrences(inp, k, window, lower, upper, fastq_path):
    """the reference's MinimizerCount on a FASTQ of the reads + the minimizer branch of its tuple emission"""
    size = write_fastq(inp, fastq_path)
    return _occurrences(ref().bella_ref_minimizer_occurrences, (fastq_path.encode(), ctypes.c_uint64(size)), inp, (k, window), lower, upper)


def ref_reliable_occurrences(inp, k, lower, upper, fastq_path):
    """the reference's SplitCount on a FASTQ of the reads + its tuple emission loop -> same triple"""
    size = write_fastq(inp, fastq_path)
    return _occurrences(ref().bella_ref_reliable_occurrences, (fastq_path.encode(), ctypes.c_uint64(size)), inp, k, lower, upper)
