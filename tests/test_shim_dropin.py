"""Drop-in check at the reference's own call boundary: the reference's HashSpGEMM (src/main.cpp:499,
--skip-alignment) and HashSpGEMM_b200 from bella_b200/csrc/overlap_b200.hpp, both compiled in one
translation unit with the UNMODIFIED reference headers (oracle/shim_driver.cpp), must write the same
BELLA output file (compared as sorted lines; the reference's line order is schedule dependent).
The library is built in the container (it needs /root/reference) and travels prebuilt to the GPU box."""
import ctypes
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM_SO = os.path.join(ROOT, "oracle", "_ref", "libbella_shim_test.so")


def _run(inp, tmp_path, memory_mb, which):
    L = ctypes.CDLL(SHIM_SO)
    p = lambda a: ctypes.c_void_p(a.ctypes.data)
    out_ref, out_b200 = str(tmp_path / "ref.out"), str(tmp_path / "b200.out")
    rc = L.shim_compare(ctypes.c_uint32(inp.n_reads), ctypes.c_uint32(inp.n_kmers), ctypes.c_uint32(inp.nnz),
                        p(inp.A_colptr), p(inp.A_rowids), p(inp.A_values), p(inp.B_colptr), p(inp.B_rowids), p(inp.B_values),
                        p(inp.seqs), p(inp.seq_off), ctypes.c_uint16(inp.kmer_size), ctypes.c_uint16(inp.bin_size),
                        ctypes.c_double(memory_mb), out_ref.encode(), out_b200.encode(), ctypes.c_int(which))
    assert rc == 0
    return out_ref, out_b200


def _lines(path):
    with open(path) as f:
        return sorted(f.read().splitlines())


@pytest.mark.skipif(not os.path.exists(SHIM_SO), reason="oracle/_ref/libbella_shim_test.so not built (needs the reference tree)")
def test_reference_side_of_the_shim_library_runs(small_inputs, tmp_path):
    # CPU: the library loads (it links libbella_b200.so) and the reference arm writes one line per output nonzero
    import oracle_lib as ol
    out_ref, _ = _run(small_inputs, tmp_path, 8000.0, 1)
    lines = _lines(out_ref)
    assert len(lines) == ol.oracle_spgemm(small_inputs, want_aux=False).nnz
    assert all(len(l.split("\t")) == 6 for l in lines[:50])


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(SHIM_SO), reason="oracle/_ref/libbella_shim_test.so not built (needs the reference tree)")
def test_shim_writes_the_reference_output_file(small_inputs, tmp_path):
    # One stage only: with several stages the reference's own writer (overlap.hpp:603-639) reopens the file and
    # fseek()s every thread's slice from offset 0 again, so later stages overwrite earlier ones -- the file is
    # schedule-dependent garbage for the reference and the shim alike (both call that same function).  The staged
    # numeric calls themselves are covered by test_gpu_staged_numeric_and_column_range.
    out_ref, out_b200 = _run(small_inputs, tmp_path, 8000.0, 3)
    ref, got = _lines(out_ref), _lines(out_b200)
    assert len(ref) > 1000
    assert got == ref


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(SHIM_SO), reason="oracle/_ref/libbella_shim_test.so not built (needs the reference tree)")
def test_tuples_shim_writes_the_reference_output_file(small_inputs, tmp_path):
    # "next" row f2 at the reference's own boundary: from the tuples, the reference runs its CSC constructor
    # (MergeDuplicates) + Transpose + HashSpGEMM; the shim's OverlapFromTuples_b200 builds B on the device
    out_ref, out_b200 = _run(small_inputs, tmp_path, 8000.0, 1 | 4)
    ref, got = _lines(out_ref), _lines(out_b200)
    assert len(ref) > 1000
    assert got == ref
