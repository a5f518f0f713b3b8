"""CPU, build container only: the C oracle against the UNMODIFIED reference (oracle/_ref) on fresh
seeded inputs.  Skipped where oracle/_ref has not been built (it needs /root/reference)."""
import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref/libbella_ref.so not built (needs /root/reference)")


def test_small(small_inputs):
    ol.assert_same(ol.oracle_spgemm(small_inputs), ol.ref_spgemm(small_inputs))


def test_reference_thread_invariance(small_inputs):
    a = ol.ref_spgemm(small_inputs, nthreads=1)
    b = ol.ref_spgemm(small_inputs, nthreads=4)
    ol.assert_same(a, b)


def test_column_prefix_sample(small_inputs):
    # the bounded cpu_baseline sample of bench.py: output columns [0, c) of the same workload
    a = ol.ref_spgemm(small_inputs, ncols=500)
    b = ol.oracle_spgemm(small_inputs, ncols=500)
    ol.assert_same(a, b)
    full = ol.oracle_spgemm(small_inputs)
    np.testing.assert_array_equal(a.colptrC, full.colptrC[:501])


@pytest.mark.parametrize("k,bin_size,hi", [(15, 500, 8), (17, 100, 12), (21, 1000, 6)])
def test_parameters(k, bin_size, hi):
    from bella_b200 import frontend as fe
    inp = fe.synthetic(400, 3000, seed=k, k=k, hi=hi, bin_size=bin_size)
    ol.assert_same(ol.oracle_spgemm(inp), ol.ref_spgemm(inp))
