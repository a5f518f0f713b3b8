"""CPU: the C-ABI library builds, loads and exports every symbol include/bella_b200.h declares; the
product path fails loudly (no CPU fallback) when no B200 is present.  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header="bella_b200.h", prefix="bella_b200_"):
    hdr = open(os.path.join(ROOT, "include", header)).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"\w+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from bella_b200 import _build, spgemm
    path = _build.build_cuda()
    L = ctypes.CDLL(path)
    syms = declared_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/bella_b200.h but not exported"
    assert sorted(spgemm.EXPORTS) == syms


def test_xdrop_library_exports_every_declared_symbol():
    from bella_b200 import _build, xdrop
    L = ctypes.CDLL(_build.build_xdrop())
    syms = declared_symbols("bella_xdrop.h", "bella_xdrop_")
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/bella_xdrop.h but not exported"
    assert sorted(xdrop.EXPORTS) == syms


def test_kmers_library_exports_every_declared_symbol():
    from bella_b200 import _build, kmers
    L = ctypes.CDLL(_build.build_kmers())
    syms = declared_symbols("bella_kmers.h", "bella_kmers_")
    assert len(syms) >= 6
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/bella_kmers.h but not exported"
    assert sorted(kmers.EXPORTS) == syms


def test_product_path_does_not_touch_the_oracle():
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "bella_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".hpp", ".cuh")):
                src = open(os.path.join(d, f)).read()
                if re.search(r"oracle_lib|bella_oracle|libbella_ref|oracle/", src):
                    bad.append(f)
    assert not bad, f"product files reference the oracle: {bad}"


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from bella_b200 import spgemm
    with pytest.raises(spgemm.BellaB200Error):
        spgemm.OverlapSpGEMM(0)
    from bella_b200 import xdrop
    with pytest.raises(xdrop.BellaXdropError):
        xdrop.XdropAligner(0)
    from bella_b200 import kmers
    with pytest.raises(kmers.BellaKmersError):
        kmers.KmerCounter(0)
