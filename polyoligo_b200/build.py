"""Builds libpolyoligo_b200.so in-tree with nvcc for sm_100a.

The library is plain CUDA C++ behind a C ABI (include/polyoligo_b200.h); it does
not link against torch.  -fmad=false keeps every double operation individually
rounded: the thal kernels promise results bit-identical to the scalar algorithm.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpolyoligo_b200.so")
SOURCES = ["po_lib.cu", "po_tables.cpp"]
PUBLIC_HEADER = os.path.join(HERE, "..", "include", "polyoligo_b200.h")
SOURCE_SUFFIXES = (".cu", ".cuh", ".h", ".inc", ".cpp")


def nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def source_files():
    """Every file the library is compiled from: all of csrc/ plus the public header."""
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(SOURCE_SUFFIXES)]
    return files + [PUBLIC_HEADER]


def source_hash():
    """sha1 over the library's sources (names + contents); stamps profiles/current.json."""
    import hashlib
    h = hashlib.sha1()
    for f in source_files():
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in source_files())


def build(force=False, verbose=False, out=None, defines=()):
    """``out`` / ``defines`` build an alternate library (A/B experiments, selected at run
    time with POLYOLIGO_B200_LIB); the default builds the in-tree product library."""
    if out is None and not force and not stale():
        return LIB
    out = out or LIB
    cmd = [nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-fmad=false", "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++", "-shared", "-cudart", "static",
           "-o", out] + ["-D" + d for d in defines] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("-D") and not a.startswith("--out=")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in args, verbose="-v" in args, out=outs[0] if outs else None,
                defines=[a[2:] for a in sys.argv[1:] if a.startswith("-D")]))
