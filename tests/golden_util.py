"""TEST INFRASTRUCTURE: load tests/golden/*.npz (made by tests/golden/make_golden.py from the unmodified reference)."""
import os

import numpy as np

from bella_b200.frontend import OverlapInputs
from oracle_lib import Result

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SPGEMM_FIXTURES = ["sanity", "tiny_clr", "tiny_hifi", "tiny_bin50", "repeats", "heavy_units", "huge_pair", "ragged"]


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = z["meta"]
    kw = {k: z[k] for k in z.files if k != "meta" and not k.startswith("ref_")}
    inp = OverlapInputs(n_reads=int(meta[0]), n_kmers=int(meta[1]), nnz=int(meta[2]), kmer_size=int(meta[3]),
                        bin_size=int(meta[4]), **kw)
    ref = Result(z["ref_flopC"], z["ref_colptrC"], z["ref_rowids"], z["ref_count"], z["ref_posH"], z["ref_posV"], z["ref_aux"])
    return inp, ref
