#!/usr/bin/env python
"""BASELINE.json configs[0] ("dataset/selfSampleData E. coli, k=17, --skip-alignment") restated as SURVEY.md 8d does: the FASTQ is
not in the reference repository, so reads are SIMULATED from its own dataset files -- the genome of
/root/reference/dataset/selfSampleData/reference.fasta at the intervals (start, end) of /root/reference/dataset/ecsample-truth.txt
(>= 500 bp), strand uniform, error 0.15 split 10/60/30 % substitution / insertion / deletion, numpy default_rng(1).

    python tools/make_config1.py          (in the build container: /root/reference exists only here)

writes scratch/config1_reads.npz (2-bit packed reads + offsets, ~33 MB): scratch/ is git-ignored but travels to the GPU box with
the gpurun snapshot, where `bench.py --config 1` and tests/test_gpu_fullsize.py::test_config1_ecoli_sim read it (both skip / refuse
cleanly when it is absent).  Nothing of the reference is copied into the repository."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/dataset"
out = os.path.join(ROOT, "scratch", "config1_reads.npz")

genome = []
with open(os.path.join(REF, "selfSampleData", "reference.fasta")) as f:
    for line in f:
        if not line.startswith(">"):
            genome.append(line.strip().upper())
g = np.frombuffer("".join(genome).encode(), dtype=np.uint8)
code = np.full(256, 255, dtype=np.uint8)
for i, c in enumerate(b"ACGT"):
    code[c] = i
g = code[g]
assert (g < 4).all(), "the reference genome is expected to be ACGT only (SURVEY.md 8d)"
iv = []
with open(os.path.join(REF, "ecsample-truth.txt")) as f:
    for line in f:
        p = line.split()
        if len(p) >= 4:
            s, e = int(p[-2]), int(p[-1])
            if e - s >= 500 and 0 <= s < e <= len(g):
                iv.append((s, e))
rng = np.random.default_rng(1)
E, SUB, INS = 0.15, 0.10, 0.60
reads, offs = [], [0]
for s, e in iv:
    t = g[s:e]
    if rng.integers(2):
        t = (3 - t)[::-1]
    u = rng.random(len(t))
    sub = u < E * SUB
    ins = (u >= E * SUB) & (u < E * (SUB + INS))
    dele = (u >= E * (SUB + INS)) & (u < E)
    base = np.where(sub, (t + 1 + rng.integers(0, 3, len(t))) & 3, t).astype(np.uint8)
    emit = np.where(dele, 0, np.where(ins, 2, 1))
    r = np.repeat(base, emit)
    first_of_ins = np.cumsum(emit)[ins] - 2                      # an insertion emits a random base in front of the template base
    r[first_of_ins] = rng.integers(0, 4, len(first_of_ins)).astype(np.uint8)
    r = r[:65000]
    reads.append(r); offs.append(offs[-1] + len(r))
allr = np.concatenate(reads)
pad = (-len(allr)) % 4
a = np.concatenate([allr, np.zeros(pad, np.uint8)]).reshape(-1, 4)
packed = (a[:, 0] | (a[:, 1] << 2) | (a[:, 2] << 4) | (a[:, 3] << 6)).astype(np.uint8)
os.makedirs(os.path.dirname(out), exist_ok=True)
np.savez(out, packed=packed, offs=np.array(offs, dtype=np.uint64), n_bases=np.array([len(allr)], dtype=np.int64))
print(f"{len(iv)} reads, {len(allr)} bases -> {out} ({os.path.getsize(out) / 1e6:.1f} MB)")
