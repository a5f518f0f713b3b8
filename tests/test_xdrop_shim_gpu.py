"""Drop-in check of the reference-side binding of the "next" row f1 (bella_b200/csrc/align_b200.hpp):
oracle/align_shim_driver.cpp includes the UNMODIFIED reference headers, writes BELLA's output file once with the reference's
own alignSeqAn + PostAlignDecision per nonzero and once through RunPairWiseAlignments_b200 (B200 behind the C-ABI of
include/bella_xdrop.h); the files must hold the same lines and the returned statistics must agree.
The CPU test pins the reference arm's file against the oracle (same fields, BELLA and PAF format)."""
import ctypes
import os

import numpy as np
import pytest

import oracle_lib as ol
from bella_b200 import frontend as fe

LIB = os.path.join(ol.ROOT, "oracle", "_ref", "libbella_align_shim_test.so")
needs_lib = pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libbella_align_shim_test.so not built (needs /root/reference)")


def _p(a):
    return ctypes.c_void_p(a.ctypes.data)


def run(inp, c, tmp_path, which, xdrop=7, ratiophi=0.55, delta=0.1, fixed=-1, paf=0):
    L = ctypes.CDLL(LIB)
    out_ref, out_b200 = str(tmp_path / "ref.out"), str(tmp_path / "b200.out")
    s_ref, s_b200 = np.zeros(6, dtype=np.uint64), np.zeros(6, dtype=np.uint64)
    rc = L.align_shim_compare(ctypes.c_uint32(inp.n_reads), _p(inp.seqs), _p(inp.seq_off), _p(c.colptrC), _p(c.rowids), _p(c.count),
                              _p(c.posH), _p(c.posV), ctypes.c_ushort(inp.kmer_size), ctypes.c_ushort(xdrop), ctypes.c_double(ratiophi),
                              ctypes.c_double(delta), ctypes.c_int(fixed), ctypes.c_int(paf), out_ref.encode(), out_b200.encode(),
                              ctypes.c_int(which), _p(s_ref), _p(s_b200))
    assert rc == 0
    return out_ref, out_b200, s_ref, s_b200


def lines(path):
    with open(path) as f:
        return sorted(f.read().splitlines())


def oracle_lines(inp, c, xdrop, ratiophi, delta, fixed, paf):
    """BELLA's output lines (overlap.hpp:470-488) from the oracle's alignment + decision"""
    cols = np.repeat(np.arange(inp.n_reads, dtype=np.uint32), np.diff(c.colptrC.astype(np.int64)))
    o = ol.oracle_align_post(inp, c.rowids, cols, c.posH, c.posV, xdrop, ratiophi, delta, fixed)
    lens = np.diff(inp.seq_off.astype(np.int64))
    out = []
    for p in np.nonzero(o[:, 7])[0]:
        r, v = int(c.rowids[p]), int(cols[p])
        score, strand, bH, eH, bV, eV, ov = (int(x) for x in o[p, :7])
        if not paf:
            out.append(f"read{v}\tread{r}\t{int(c.count[p])}\t{score}\t{ov}\t{chr(strand)}\t{bV}\t{eV}\t{lens[v]}\t{bH}\t{eH}\t{lens[r]}")
        else:
            if chr(strand) == "c":
                bH, eH = lens[r] - eH, lens[r] - bH
            out.append(f"read{v}\t{lens[v]}\t{bV}\t{eV}\t{'+' if chr(strand) == 'n' else '-'}\tread{r}\t{lens[r]}\t{bH}\t{eH}\t{score}\t{ov}\t255")
    return sorted(out), o


@pytest.fixture(scope="module")
def case():
    inp = fe.synthetic(300, 3000, seed=101)
    return inp, ol.oracle_spgemm(inp, want_aux=False)


@needs_lib
@pytest.mark.parametrize("paf", [0, 1])
def test_reference_arm_writes_what_the_oracle_computes(case, tmp_path, paf):
    inp, c = case
    out_ref, _, s_ref, _ = run(inp, c, tmp_path, 1, ratiophi=0.7, delta=0.2, paf=paf)
    want, o = oracle_lines(inp, c, 7, 0.7, 0.2, -1, paf)
    assert lines(out_ref) == want and 0 < len(want) < c.nnz
    assert int(s_ref[0]) == c.nnz and int(s_ref[3]) == len(want) and int(s_ref[1]) == int((o[:, 5] - o[:, 4]).sum())


@needs_lib
@pytest.mark.gpu
@pytest.mark.parametrize("paf,fixed", [(0, -1), (1, -1), (0, 400)])
def test_b200_binding_writes_the_reference_output_file(case, tmp_path, paf, fixed):
    inp, c = case
    out_ref, out_b200, s_ref, s_b200 = run(inp, c, tmp_path, 3, ratiophi=0.7, delta=0.2, fixed=fixed, paf=paf)
    a, b = lines(out_ref), lines(out_b200)
    assert len(a) > 100 and a == b
    np.testing.assert_array_equal(s_ref, s_b200)
