"""GPU, BASELINE.json's full single-GPU size (configs[1]: 50 k reads x 10 kb, e=0.15, k=17, [2,8], 30x):
size-independent properties of the result plus exact agreement with the oracle on a column prefix
(the oracle needs ~4 s of 16 cores for the whole matrix; the prefix keeps the test in seconds)."""
import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def config2():
    from bella_b200 import frontend as fe
    return fe.synthetic(50000, 10000, coverage=30.0, err=0.15, seed=2)


def test_full_size_properties_and_oracle_prefix(config2):
    from bella_b200 import spgemm
    inp = config2
    g = spgemm.OverlapSpGEMM(0)
    g.set_inputs(inp)
    flops, flopC, colptrC = g.symbolic()
    rows, cnt, pH, pV, aux = g.numeric(aux=True)
    n = inp.n_reads
    Z = int(colptrC[-1])
    # structure: checksum of the per-column counts, monotone colptr, strictly lower triangle, rows ascending per column
    assert flops == int(flopC.astype(np.uint64).sum()) and Z == rows.size
    assert (np.diff(colptrC.astype(np.int64)) >= 0).all()
    cols = np.repeat(np.arange(n, dtype=np.int64), np.diff(colptrC.astype(np.int64)))
    assert (rows.astype(np.int64) > cols).all()
    same_col = cols[1:] == cols[:-1]
    assert (rows[1:][same_col] > rows[:-1][same_col]).all()
    # semiring invariants: 1 <= nbins, 1 <= support <= products of the pair; seeds are valid k-mer starts
    assert (aux[:, 0] >= 1).all() and (aux[:, 1] >= 1).all()
    assert (pH.astype(np.int64) + inp.kmer_size <= inp.read_len[rows].astype(np.int64)).all()
    assert (pV.astype(np.int64) + inp.kmer_size <= inp.read_len[cols].astype(np.int64)).all()
    # every pair has at least one product and the products sum to flops: count >= 1 unless it wrapped (u16)
    assert Z <= flops
    # exact agreement with the oracle on the first 1200 output columns
    want = ol.oracle_spgemm(inp, ncols=1200)
    z = want.nnz
    np.testing.assert_array_equal(flopC[:1200], want.flopC)
    np.testing.assert_array_equal(colptrC[:1201], want.colptrC)
    np.testing.assert_array_equal(rows[:z], want.rowids)
    np.testing.assert_array_equal(cnt[:z], want.count)
    np.testing.assert_array_equal(pH[:z], want.posH)
    np.testing.assert_array_equal(pV[:z], want.posV)
    np.testing.assert_array_equal(aux[:z], want.aux)
    # idempotence: a second pass over the same handle gives the same bytes
    flops2, flopC2, colptrC2 = g.symbolic()
    rows2, cnt2, pH2, pV2 = g.numeric()
    assert flops2 == flops and (colptrC2 == colptrC).all() and (rows2 == rows).all() and (cnt2 == cnt).all() and (pH2 == pH).all() and (pV2 == pV).all()
    # a column range of the same matrix reproduces the slice of the whole result (row sharding, staged numeric)
    lo, hi = 20000, 23000
    g.set_column_range(lo, hi)
    _, flopC3, colptrC3 = g.symbolic()
    rows3, cnt3, pH3, pV3 = g.numeric()
    z0, z1 = int(colptrC[lo]), int(colptrC[hi])
    np.testing.assert_array_equal(colptrC3.astype(np.int64), colptrC[lo:hi + 1].astype(np.int64) - z0)
    np.testing.assert_array_equal(rows3, rows[z0:z1])
    np.testing.assert_array_equal(cnt3, cnt[z0:z1])
    np.testing.assert_array_equal(pH3, pH[z0:z1])
    g.close()


def test_config1_ecoli_sim_whole_matrix_against_the_oracle():
    """BASELINE.json configs[0] as SURVEY.md 8d restates it: E. coli-sim (tools/make_config1.py: reads simulated from the
    reference's own genome at its own truth intervals).  The packed reads live in scratch/ (git-ignored, travels with the
    gpurun snapshot); without them the test has nothing to run on.  The WHOLE result is compared with the oracle."""
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench
    if not os.path.exists(bench.CONFIG1_READS):
        pytest.skip("scratch/config1_reads.npz is absent (tools/make_config1.py writes it in the build container)")
    from bella_b200 import spgemm
    inp = bench.load_workload(dict(bench.CONFIGS[1]), need_seqs=False)
    assert inp.n_reads == 14939 and 5.9e6 < inp.n_kmers < 6.1e6 and 15.0e6 < inp.nnz < 15.5e6        # 14,947 / 6.00 M / 15.24 M
    r = spgemm.overlap_spgemm(inp, aux=True)
    got = ol.Result(r["flopC"], r["colptrC"], r["rowids"], r["count"], r["posH"], r["posV"], r["aux"])
    want = ol.oracle_spgemm(inp)
    ol.assert_same(got, want)
    assert 1.70e6 < want.nnz < 1.73e6 and 14.4e6 < r["flops"] < 14.8e6                                 # Z = 1,713,290, F = 14,606,754
