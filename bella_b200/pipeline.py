"""reads -> overlaps on one B200, every stage behind its C-ABI: reliable k-mers and tuples (f3, include/bella_kmers.h) ->
matrix construction (f2) + overlap SpGEMM (include/bella_b200.h) -> gapped X-drop alignment + accept/reject (f1,
include/bella_xdrop.h) -> BELLA's output lines (f4, include/overlap.hpp:470-488).  This is src/main.cpp from the k-mer
counting to the output file (:282-525) with the FASTQ parser left to the caller.  No CPU fallback anywhere.
The whole chain runs on the B200 (tests/test_gpu_rows_f1_f3.py::test_reads_to_overlaps_equals_the_oracle_chain)."""
import numpy as np

from . import kmers, spgemm, xdrop


def overlap_reads(seqs, seq_off, names=None, k=17, lower=2, upper=8, bin_size=500, xdrop_value=7, ratiophi=0.0, delta_chernoff=0.1,
                  fixed_threshold=-1, paf=False, device=0):
    """-> dict(lines, rows, cols, count, posH, posV, out8, n_kmers, stage_ms)"""
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
    seq_off = np.ascontiguousarray(seq_off, dtype=np.uint64)
    # The strand bit of a nonzero ("the window equals its canonical k-mer") replaces the reference's raw substring comparison
    # (chain.hpp:35-44) only for upper-case ACGT windows: the k-mer code maps N to G and lower case to upper case
    # (Kmer.cpp:215-216), so two windows that differ as strings can get the same code and the same bit, and the overlap
    # estimate, the bins and the count of such a pair would silently differ from BELLA's.  Refuse such reads, as the C++ shim does
    # (bella_b200/csrc/overlap_b200.hpp); INTEGRATION.md "Reads with N or lower-case bases".
    ok = np.zeros(256, dtype=bool)
    ok[list(b"ACGT")] = True
    if seqs.size and not ok[seqs].all():
        bad = int(np.flatnonzero(~ok[seqs])[0])
        raise ValueError(f"read {int(np.searchsorted(seq_off, bad, side='right')) - 1} has the byte {chr(int(seqs[bad]))!r} at offset {bad}: only "
                         "upper-case ACGT reads are accepted (the strand-bit form of checkstrand is exact for those only)")
    n_reads = len(seq_off) - 1
    lens = np.diff(seq_off.astype(np.int64)).astype(np.uint32)
    kc = kmers.KmerCounter(device)
    g = spgemm.OverlapSpGEMM(device)
    al = xdrop.XdropAligner(device)
    try:
        t = kc.count(seqs, seq_off, k, lower, upper)
        strand = np.concatenate([t["t_strand"], np.zeros(8, np.uint8)])
        g.set_inputs_tuples(t["n_kmers"], n_reads, t["t_kmer"], t["t_read"], t["t_pos"], strand, lens, k, bin_size)
        flops, flopC, colptrC = g.symbolic()
        rows, count, posH, posV = (a.copy() for a in g.numeric()[:4])
        cols = np.repeat(np.arange(n_reads, dtype=np.uint32), np.diff(colptrC.astype(np.int64)))
        al.set_reads(seqs, seq_off)
        al.set_params(k, xdrop_value, ratiophi, delta_chernoff, fixed_threshold)
        out8 = al.align(rows, cols, posH, posV)
        stage_ms = {"kmers": kc.stats()["kernel_ms"], "spgemm": g.timings(), "xdrop": al.stats()["kernel_ms"]}
    finally:
        kc.close(); g.close(); al.close()
    name = (lambda i: names[i]) if names is not None else (lambda i: f"read{i}")
    lines = []
    for p in np.nonzero(out8[:, 7])[0]:
        r, v = int(rows[p]), int(cols[p])
        score, strand_c, bH, eH, bV, eV, ov = (int(x) for x in out8[p, :7])
        if not paf:
            lines.append(f"{name(v)}\t{name(r)}\t{int(count[p])}\t{score}\t{ov}\t{chr(strand_c)}\t{bV}\t{eV}\t{lens[v]}\t{bH}\t{eH}\t{lens[r]}")
        else:
            if chr(strand_c) == "c":
                bH, eH = int(lens[r]) - eH, int(lens[r]) - bH
            lines.append(f"{name(v)}\t{lens[v]}\t{bV}\t{eV}\t{'+' if chr(strand_c) == 'n' else '-'}\t{name(r)}\t{lens[r]}\t{bH}\t{eH}\t{score}\t{ov}\t255")
    return {"lines": lines, "rows": rows, "cols": cols, "count": count, "posH": posH, "posV": posV, "out8": out8, "colptrC": colptrC,
            "n_kmers": t["n_kmers"], "stage_ms": stage_ms}
