"""Result checks shared by bench.py and the multi-GPU bench (host side, numpy)."""
import numpy as np


def tuple_checksum(col_lo, colptrC, res):
    """Order-independent 64-bit checksum of the tuples (col,row,count,posH,posV) of a column range: the sum modulo 2^64 of a
    mixed 64-bit word per tuple, so that the per-rank checksums of a sharded run add up to the single-GPU one."""
    rows, cnt, pH, pV = (np.asarray(a) for a in res)
    cp = np.asarray(colptrC, dtype=np.int64)
    z = int(cp[-1] - cp[0])
    cols = np.repeat(np.arange(col_lo, col_lo + len(cp) - 1, dtype=np.uint64), np.diff(cp))
    with np.errstate(over="ignore"):
        w = (cols << np.uint64(32)) | rows[:z].astype(np.uint64)
        v = (cnt[:z].astype(np.uint64) << np.uint64(32)) | (pH[:z].astype(np.uint64) << np.uint64(16)) | pV[:z].astype(np.uint64)
        x = (w * np.uint64(0x9E3779B97F4A7C15)) ^ (v * np.uint64(0xC2B2AE3D27D4EB4F))
        x ^= x >> np.uint64(29)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(32)
        return int(x.sum(dtype=np.uint64)), z
