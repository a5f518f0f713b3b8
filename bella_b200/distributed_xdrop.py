"""Multi-GPU for the "next" row f1 (one process per GPU): the alignment step shards by PAIRS and needs no collective.

The pairs of a batch are the nonzeros of C in CSC order, so a contiguous slice of pairs is a column range of C -- the same
ownership the row-sharded overlap SpGEMM ends with (bella_b200/distributed.py: every GPU holds the result of the columns it
owns).  Every GPU keeps the whole read set resident (config 2: 0.5 GB) and aligns its own slice; outputs are disjoint and the
host concatenates them, as the reference concatenates its per-thread output (include/overlap.hpp:603-639).  `scaling` is
"weak" when every rank brings its own batch, "strong" when one batch is cut with shard_pairs().

Not yet run on more than one GPU; tests/test_distributed_cpu.py checks the cutting and the union on CPU (gloo)."""
import numpy as np


def pair_costs(seq_off, rows, cols, posH, posV, kmer_len):
    """expected anti-diagonals of a pair: both directions extend at most to the nearer read end (the strand is not known
    before the seed comparison, so the H side uses the position as given -- a balance heuristic, not a result)"""
    seq_off = np.asarray(seq_off, dtype=np.int64)
    lens = np.diff(seq_off)
    rows = np.asarray(rows, dtype=np.int64); cols = np.asarray(cols, dtype=np.int64)
    pH = np.asarray(posH, dtype=np.int64); pV = np.asarray(posV, dtype=np.int64)
    left = np.minimum(pH, pV)
    right = np.minimum(lens[rows] - pH - kmer_len, lens[cols] - pV - kmer_len)
    return 2 * (np.maximum(left, 0) + np.maximum(right, 0)) + 64      # + a constant per pair (set-up, result)


def shard_pairs(costs, world):
    """contiguous cut points [world + 1] with (nearly) equal cumulative cost"""
    costs = np.asarray(costs, dtype=np.int64)
    n = len(costs)
    if n == 0:
        return np.zeros(world + 1, dtype=np.int64)
    pre = np.cumsum(costs)
    targets = pre[-1] * np.arange(1, world, dtype=np.float64) / world
    cuts = np.searchsorted(pre, targets, side="left") + 1
    bounds = np.concatenate([[0], np.minimum(cuts, n), [n]]).astype(np.int64)
    return np.maximum.accumulate(bounds)


class ShardedXdropAligner:
    """aligner_factory() -> object with set_reads / set_params / align (bella_b200.xdrop.XdropAligner on a GPU box)"""

    def __init__(self, rank, world, aligner_factory=None):
        self.rank, self.world = rank, world
        if aligner_factory is None:
            from . import xdrop
            aligner_factory = lambda: xdrop.XdropAligner(rank)  # noqa: E731
        self.aligner = aligner_factory()
        self.kmer_len = 17

    def set_reads(self, seqs, seq_off):
        self.seq_off = np.asarray(seq_off)
        self.aligner.set_reads(seqs, seq_off)

    def set_params(self, kmer_len=17, xdrop=7, ratiophi=0.0, delta_chernoff=0.1, fixed_threshold=-1):
        self.kmer_len = kmer_len
        self.aligner.set_params(kmer_len, xdrop, ratiophi, delta_chernoff, fixed_threshold)

    def my_slice(self, rows, cols, posH, posV):
        b = shard_pairs(pair_costs(self.seq_off, rows, cols, posH, posV, self.kmer_len), self.world)
        return int(b[self.rank]), int(b[self.rank + 1])

    def align_my_share(self, rows, cols, posH, posV):
        """one batch known to every rank (strong scaling): -> (lo, hi, out[hi - lo][8])"""
        lo, hi = self.my_slice(rows, cols, posH, posV)
        return lo, hi, self.aligner.align(rows[lo:hi], cols[lo:hi], posH[lo:hi], posV[lo:hi])

    def close(self):
        if hasattr(self.aligner, "close"):
            self.aligner.close()
