"""TEST INFRASTRUCTURE: the oracle side of the whole chain reads -> reliable k-mers (f3) -> matrices (f2) -> overlap SpGEMM
(a-e) -> X-drop alignment + decision (f1) -> output lines (f4), every stage by the CPU oracle, with the SAME k-mer ids the
device assigns (rank of the canonical k-mer), so that the device chain can be compared value for value."""
import ctypes

import numpy as np

import oracle_lib as ol
from bella_b200 import frontend as fe


def kmer_codes():
    code = np.zeros(256, dtype=np.uint64)                      # Kmer::set_kmer, kmercode/Kmer.cpp:215-216
    for c in range(256):
        x = (c & 4) >> 1
        code[c] = x + ((x ^ (c & 2)) >> 1)
    return code


def rank_tuples(inp, k, lower, upper):
    """(t_kmer, t_read, t_pos, t_strand) with id = rank of the canonical k-mer's packed value -- what bella_kmers emits"""
    r, p, n_kmers = ol.oracle_reliable_occurrences(inp, k, lower, upper)
    code = kmer_codes()
    g = inp.seq_off[r].astype(np.int64) + p.astype(np.int64)
    w = code[inp.seqs[g[:, None] + np.arange(k)[None, :]]]     # [n][k] base codes
    sh = np.uint64(2) * np.arange(k - 1, -1, -1, dtype=np.uint64)
    fw = (w << sh[None, :]).sum(axis=1, dtype=np.uint64)
    rv = ((np.uint64(3) - w) << sh[::-1][None, :]).sum(axis=1, dtype=np.uint64)
    canon = np.minimum(fw, rv)
    uniq, ids = np.unique(canon, return_inverse=True)
    assert len(uniq) == n_kmers
    return ids.astype(np.uint32), r, p, (fw <= rv).astype(np.uint8), n_kmers


def inputs_from_tuples(n_kmers, n_reads, t_kmer, t_read, t_pos, t_strand, seqs, seq_off, k=17, bin_size=500):
    """tuples (read-major, position order) -> OverlapInputs through the oracle's CSC constructor + MergeDuplicates + transpose"""
    nt = len(t_kmer)
    tk = np.ascontiguousarray(t_kmer, dtype=np.uint32); tr = np.ascontiguousarray(t_read, dtype=np.uint32)
    tp = np.ascontiguousarray(t_pos, dtype=np.uint16)
    Bc = np.zeros(n_reads + 1, np.uint32); Br = np.zeros(nt, np.uint32); Bv = np.zeros(nt, np.uint16)
    Ac = np.zeros(n_kmers + 1, np.uint32); Ar = np.zeros(nt, np.uint32); Av = np.zeros(nt, np.uint16)
    nnz = ol.oracle().oracle_build_csc(ctypes.c_uint32(n_kmers), ctypes.c_uint32(n_reads), ctypes.c_uint64(nt), ol._p(tk), ol._p(tr), ol._p(tp),
                                       ol._p(Bc), ol._p(Br), ol._p(Bv), ol._p(Ac), ol._p(Ar), ol._p(Av))
    assert nnz >= 0
    Br, Bv, Ar, Av = Br[:nnz], Bv[:nnz], Ar[:nnz], Av[:nnz]
    # strand of a nonzero = strand of the occurrence (read, pos) it kept
    key = tr.astype(np.uint64) << np.uint64(16) | tp.astype(np.uint64)
    order = np.argsort(key)
    skey, sst = key[order], np.asarray(t_strand, dtype=np.uint8)[order]

    def strand_of(reads, pos):
        q = reads.astype(np.uint64) << np.uint64(16) | pos.astype(np.uint64)
        i = np.searchsorted(skey, q)
        assert (skey[i] == q).all()
        return np.concatenate([np.packbits(sst[i], bitorder="little"), np.zeros(8, np.uint8)])      # padded like the front end's

    b_reads = np.repeat(np.arange(n_reads, dtype=np.uint32), np.diff(Bc.astype(np.int64)))
    lens = np.diff(np.asarray(seq_off).astype(np.int64)).astype(np.uint32)
    return fe.OverlapInputs(n_reads=n_reads, n_kmers=n_kmers, nnz=int(nnz), A_colptr=Ac, A_rowids=Ar, A_values=Av, A_strand=strand_of(Ar, Av),
                            B_colptr=Bc, B_rowids=Br, B_values=Bv, B_strand=strand_of(b_reads, Bv), read_len=lens, kmer_size=k,
                            bin_size=bin_size, seqs=seqs, seq_off=seq_off)


def output_lines(inp, rows, cols, count, out8, paf=False):
    """BELLA's output lines (include/overlap.hpp:470-488), reads named read<i>"""
    lens = np.diff(inp.seq_off.astype(np.int64))
    out = []
    for p in np.nonzero(out8[:, 7])[0]:
        r, v = int(rows[p]), int(cols[p])
        score, strand, bH, eH, bV, eV, ov = (int(x) for x in out8[p, :7])
        if not paf:
            out.append(f"read{v}\tread{r}\t{int(count[p])}\t{score}\t{ov}\t{chr(strand)}\t{bV}\t{eV}\t{lens[v]}\t{bH}\t{eH}\t{lens[r]}")
        else:
            if chr(strand) == "c":
                bH, eH = lens[r] - eH, lens[r] - bH
            out.append(f"read{v}\t{lens[v]}\t{bV}\t{eV}\t{'+' if chr(strand) == 'n' else '-'}\tread{r}\t{lens[r]}\t{bH}\t{eH}\t{score}\t{ov}\t255")
    return out


def oracle_chain(seqs, seq_off, k=17, lower=2, upper=8, bin_size=500, xdrop=7, ratiophi=0.55, delta=0.1, fixed_threshold=-1):
    """-> dict(inp, C (oracle Result), cols, out8, lines)"""
    n_reads = len(seq_off) - 1
    raw = fe.OverlapInputs(n_reads=n_reads, n_kmers=0, nnz=0, A_colptr=None, A_rowids=None, A_values=None, A_strand=None, B_colptr=None,
                           B_rowids=None, B_values=None, B_strand=None, read_len=None, kmer_size=k, seqs=seqs, seq_off=seq_off)
    tk, tr, tp, ts, n_kmers = rank_tuples(raw, k, lower, upper)
    inp = inputs_from_tuples(n_kmers, n_reads, tk, tr, tp, ts, seqs, seq_off, k, bin_size)
    C = ol.oracle_spgemm(inp, want_aux=False)
    cols = np.repeat(np.arange(n_reads, dtype=np.uint32), np.diff(C.colptrC.astype(np.int64)))
    out8 = ol.oracle_align_post(inp, C.rowids, cols, C.posH, C.posV, xdrop, ratiophi, delta, fixed_threshold)
    return {"inp": inp, "tuples": (tk, tr, tp, ts), "C": C, "cols": cols, "out8": out8, "lines": output_lines(inp, C.rowids, cols, C.count, out8)}
