"""In-tree builds of the native libraries (no JIT cache: the built .so files are git-ignored but travel to the GPU box with the
snapshot of the tree, each with a `.src` digest of the sources it was built from).

  libbella_b200.so      bella_b200/csrc/bella_b200.cu  -- CUDA kernels (sm_100a) + the C-ABI (include/bella_b200.h)
  libbella_xdrop.so     bella_b200/csrc/bella_xdrop.cu -- "next" row f1: X-drop seed-and-extend kernels + C-ABI (include/bella_xdrop.h)
  libbella_kmers.so     bella_b200/csrc/bella_kmers.cu -- "next" row f3: reliable k-mer selection + tuple emission (include/bella_kmers.h)
  libbella_frontend.so  bella_b200/csrc/frontend.cpp   -- host front end (matrix construction, read simulator)
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_CUDA = os.path.join(HERE, "libbella_b200.so")
LIB_XDROP = os.path.join(HERE, "libbella_xdrop.so")
LIB_KMERS = os.path.join(HERE, "libbella_kmers.so")
LIB_FE = os.path.join(HERE, "libbella_frontend.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-fopenmp", "-shared", "--use_fast_math",
]


def _digest(sources, flags=()):
    import hashlib
    h = hashlib.sha256(" ".join(flags).encode())
    for s in sources:
        with open(s, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _stale(target, sources, flags=()):
    """A library is current when the digest of its sources (written beside it when it was built) matches: file times do
    not survive the copy to the GPU box, a digest does.  The bindings call build_*() on every load, so a parity test can
    never run against a binary built from older kernels."""
    side = target + ".src"
    if not (os.path.exists(target) and os.path.exists(side)):
        return True
    with open(side) as f:
        return f.read().strip() != _digest(sources, flags)


def _built(target, sources, flags=()):
    with open(target + ".src", "w") as f:
        f.write(_digest(sources, flags))


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: cannot build libbella_b200.so (there is no CPU fallback)")


def build_cuda(force=False, verbose=False):
    srcs = [os.path.join(CSRC, "bella_b200.cu"), os.path.join(CSRC, "kernels.cuh"), os.path.join(ROOT, "include", "bella_b200.h")]
    if not force and not _stale(LIB_CUDA, srcs):
        return LIB_CUDA
    cmd = [_nvcc()] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), "-o", LIB_CUDA,
                                    os.path.join(CSRC, "bella_b200.cu"), "-lgomp"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.run(cmd, check=True)
    _built(LIB_CUDA, srcs)
    return LIB_CUDA


def build_xdrop(force=False, verbose=False):
    srcs = [os.path.join(CSRC, "bella_xdrop.cu"), os.path.join(CSRC, "xdrop.cuh"), os.path.join(ROOT, "include", "bella_xdrop.h")]
    if not force and not _stale(LIB_XDROP, srcs):
        return LIB_XDROP
    flags = [f for f in NVCC_FLAGS if f != "--use_fast_math"]
    cmd = [_nvcc()] + flags + ["-I", os.path.join(ROOT, "include"), "-o", LIB_XDROP, srcs[0]]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.run(cmd, check=True)
    _built(LIB_XDROP, srcs)
    return LIB_XDROP


def build_kmers(force=False, verbose=False):
    srcs = [os.path.join(CSRC, "bella_kmers.cu"), os.path.join(CSRC, "kmers.cuh"), os.path.join(ROOT, "include", "bella_kmers.h")]
    if not force and not _stale(LIB_KMERS, srcs):
        return LIB_KMERS
    flags = [f for f in NVCC_FLAGS if f != "--use_fast_math"]
    cmd = [_nvcc()] + flags + ["-I", os.path.join(ROOT, "include"), "-o", LIB_KMERS, srcs[0]]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.run(cmd, check=True)
    _built(LIB_KMERS, srcs)
    return LIB_KMERS


def build_frontend(force=False):
    src = os.path.join(CSRC, "frontend.cpp")
    if not force and not _stale(LIB_FE, [src]):
        return LIB_FE
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-O3", "-std=c++17", "-fopenmp", "-fPIC", "-shared", "-Wall", "-o", LIB_FE, src], check=True)
    _built(LIB_FE, [src])
    return LIB_FE


if __name__ == "__main__":
    build_frontend()
    build_cuda(verbose=True)
    build_xdrop(verbose=True)
    build_kmers(verbose=True)
