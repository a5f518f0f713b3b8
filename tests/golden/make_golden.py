"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libbella_ref.so, built
by oracle/Makefile from /root/reference).  Run in the build container only:

    python tests/golden/make_golden.py

Each fixture stores the exact inputs handed to the reference (CSC arrays, read lengths, strand bits,
k, bin size -- and the read sequences the reference's multiply compares) and the reference's outputs
(flopC, colptrC, rowids, count, posH, posV, aux = {nbins, support, overlap of the chosen bin}).
Fixtures:
  sanity        the reference's own sanitytests/reversecomptest.fastq (3 reads, fwd / revcomp / perturbed)
  tiny_clr      300 reads x 3 kb, e=0.15, 30x
  tiny_hifi     200 reads x 3 kb, e=0.01, u=40 (hundreds of products per pair, multi-bin folds)
  tiny_bin50    tiny_clr with binSize=50 (many bins per pair)
  repeats       reads from a genome with diverged tandem repeats, u=60, binSize 200: several bins per
                pair, bins merging and re-splitting (exercises the orphan / near-bin paths of chainop)
  build_csc     tuples -> B (CSC ctor + MergeDuplicates) -> A (Transpose, 1 thread) of the reference
  heavy_units   250 reads x 5 kb, e=0.01, 40x, u=80: up to 188 k products per column, hundreds per pair (the GPU splits
                such columns into row-range units)
  huge_pair     24 reads x 14 kb, e=0.002, 8x, u=40: read pairs sharing more than 8192 k-mers (more than fits shared memory)
  xdrop         "next" row f1: (score, strand, begH, endH, begV, endV) of the reference's alignSeqAn (gapped X-drop) on
                3000 candidate pairs of the tiny_clr reads
  ragged        160 reads of 300..30 000 bp (log-uniform), 10 % substitutions, both strands: ragged columns, short reads
                with a handful of k-mers next to very long ones
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib as ol  # noqa: E402
from bella_b200 import frontend as fe  # noqa: E402


def save(name, inp, res):
    # sequences are only needed by the reference run above, not by the checks: leave them out
    d = {k: v for k, v in inp.__dict__.items() if isinstance(v, np.ndarray) and k not in ("seqs", "seq_off")}
    d["meta"] = np.array([inp.n_reads, inp.n_kmers, inp.nnz, inp.kmer_size, inp.bin_size], dtype=np.int64)
    d.update(ref_flopC=res.flopC, ref_colptrC=res.colptrC, ref_rowids=res.rowids, ref_count=res.count,
             ref_posH=res.posH, ref_posV=res.posV, ref_aux=res.aux)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
    print(f"{name}: n={inp.n_reads} m={inp.n_kmers} nnz={inp.nnz} F={int(res.flopC.sum())} Z={res.nnz} "
          f"max_nbins={int(res.aux[:, 0].max()) if res.nnz else 0}")


def repeat_reads(seed=5, unit=1500, copies=6, div=0.03, n_reads=260, L=4000, err=0.05):
    rng = np.random.default_rng(seed)
    B = np.frombuffer(b"ACGT", dtype=np.uint8)
    base = rng.integers(0, 4, unit)
    parts = []
    for c in range(copies):
        u = base.copy()
        mut = rng.random(unit) < div
        u[mut] = (u[mut] + rng.integers(1, 4, mut.sum())) % 4
        parts.append(u)
        parts.append(rng.integers(0, 4, int(rng.integers(50, 600))))
    genome = np.concatenate(parts + [rng.integers(0, 4, 3000)])
    comp = np.array([3, 2, 1, 0])
    reads = []
    for r in range(n_reads):
        s0 = int(rng.integers(0, len(genome) - L))
        t = genome[s0:s0 + L].copy()
        if rng.random() < 0.5:
            t = comp[t[::-1]]
        out = []
        for b in t:
            x = rng.random()
            if x < err * 0.3:
                out.append((b + rng.integers(1, 4)) % 4)
            elif x < err * 0.6:
                out.append(rng.integers(0, 4)); out.append(b)
            elif x < err:
                continue
            else:
                out.append(b)
        reads.append(B[np.array(out[:L], dtype=np.int64)].tobytes().decode())
    return reads


def save_repeats():
    s, o = fe.reads_from_strings(repeat_reads())
    inp = fe.build_matrices(s, o, lo=2, hi=60, bin_size=200)
    save("repeats", inp, ol.ref_spgemm(inp, nthreads=1))


def ragged_reads(seed=9, n_reads=160, genome_len=90000, err=0.10):
    rng = np.random.default_rng(seed)
    B = np.frombuffer(b"ACGT", dtype=np.uint8)
    genome = rng.integers(0, 4, genome_len)
    comp = np.array([3, 2, 1, 0])
    reads = []
    for r in range(n_reads):
        L = int(np.exp(rng.uniform(np.log(300), np.log(30000))))
        s0 = int(rng.integers(0, genome_len - L))
        t = genome[s0:s0 + L].copy()
        mut = rng.random(L) < err
        t[mut] = (t[mut] + rng.integers(1, 4, int(mut.sum()))) % 4
        if rng.random() < 0.5:
            t = comp[t[::-1]]
        reads.append(B[t].tobytes().decode())
    return reads


def main_round1_additions():
    """Fixtures added later in round 1 (the earlier files stay byte-identical)."""
    assert ol.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    inp = fe.synthetic(250, 5000, coverage=40.0, err=0.01, seed=13, hi=80)
    save("heavy_units", inp, ol.ref_spgemm(inp, nthreads=1))
    inp = fe.synthetic(24, 14000, coverage=8.0, err=0.002, seed=17, hi=40)
    save("huge_pair", inp, ol.ref_spgemm(inp, nthreads=1))
    s, o = fe.reads_from_strings(ragged_reads())
    inp = fe.build_matrices(s, o)
    save("ragged", inp, ol.ref_spgemm(inp, nthreads=1))
    save_xdrop()


def save_xdrop():
    """"Next" row f1: the reference's alignSeqAn (gapped X-drop, x = 7) on 3000 candidate pairs of the tiny_clr reads."""
    inp = fe.synthetic(300, 3000, seed=101)
    r = ol.oracle_spgemm(inp, want_aux=False)
    cols = np.repeat(np.arange(inp.n_reads, dtype=np.uint32), np.diff(r.colptrC.astype(np.int64)))
    idx = np.sort(np.random.default_rng(1).choice(r.nnz, 3000, replace=False))
    rows, cols, pH, pV = r.rowids[idx], cols[idx], r.posH[idx], r.posV[idx]
    ref = ol.ref_align(inp, rows, cols, pH, pV, 7)
    np.savez_compressed(os.path.join(HERE, "xdrop.npz"), n_reads=inp.n_reads, k=inp.kmer_size, xdrop=7, seqs=inp.seqs, seq_off=inp.seq_off,
                        rows=rows, cols=cols, posH=pH, posV=pV, ref_out=ref)
    print(f"xdrop: {len(rows)} pairs, scores {ref[:, 0].min()}..{ref[:, 0].max()}, {(ref[:, 1] == ord('c')).sum()} on the reverse strand")


def main():
    assert ol.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    names, reads = fe.read_fastq("/root/reference/sanitytests/reversecomptest.fastq")
    s, o = fe.reads_from_strings(reads)
    inp = fe.build_matrices(s, o)
    save("sanity", inp, ol.ref_spgemm(inp, nthreads=1))

    inp = fe.synthetic(300, 3000, seed=101)
    save("tiny_clr", inp, ol.ref_spgemm(inp, nthreads=1))

    inp = fe.synthetic(200, 3000, err=0.01, seed=102, hi=40)
    save("tiny_hifi", inp, ol.ref_spgemm(inp, nthreads=1))

    inp = fe.synthetic(300, 3000, seed=101, bin_size=50)
    save("tiny_bin50", inp, ol.ref_spgemm(inp, nthreads=1))

    save_repeats()

    # matrix construction
    inp = fe.synthetic(120, 2000, seed=103, keep_tuples=True)
    tk, tr, tp = inp.tuples
    import ctypes
    L = ol.ref()
    nt = len(tk)
    Bc = np.zeros(inp.n_reads + 1, np.uint32); Br = np.zeros(nt, np.uint32); Bv = np.zeros(nt, np.uint16)
    Ac = np.zeros(inp.n_kmers + 1, np.uint32); Ar = np.zeros(nt, np.uint32); Av = np.zeros(nt, np.uint16)
    nnz = L.bella_ref_build(ctypes.c_uint32(inp.n_kmers), ctypes.c_uint32(inp.n_reads), ctypes.c_uint64(nt),
                            ol._p(tk), ol._p(tr), ol._p(tp), ctypes.c_int(1),
                            ol._p(Bc), ol._p(Br), ol._p(Bv), ol._p(Ac), ol._p(Ar), ol._p(Av))
    np.savez_compressed(os.path.join(HERE, "build_csc.npz"), n_kmers=inp.n_kmers, n_reads=inp.n_reads,
                        t_kmer=tk, t_read=tr, t_pos=tp, B_colptr=Bc, B_rowids=Br[:nnz], B_values=Bv[:nnz],
                        A_colptr=Ac, A_rowids=Ar[:nnz], A_values=Av[:nnz])
    print(f"build_csc: tuples={nt} nnz={nnz} (duplicates merged: {nt - nnz})")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "additions":
        main_round1_additions()
    else:
        main()
        main_round1_additions()
