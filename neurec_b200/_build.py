"""Builds neurec_b200/libneurec_b200.so (the C-ABI CUDA library) with nvcc for sm_100a.

In-tree build so the library travels to the GPU box with the repository snapshot.  Every .cu is
compiled to its own object (in parallel, only when it or a header changed) and the objects are
linked into one shared library.
"""
from __future__ import annotations

import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libneurec_b200.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "neurec_b200.h")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-diag-suppress", "549",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + [HEADER, __file__]


def _obj_of(src):
    return os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")


def _stale(target, deps):
    if not os.path.isfile(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build() -> bool:
    return _stale(LIB, sources() + _headers())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _headers()
    todo = [s for s in sources() if force or _stale(_obj_of(s), [s] + hdrs)]

    def compile_one(src):
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", _obj_of(src), src]
        r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
        return src, r

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(todo)))) as pool:
        for src, r in pool.map(compile_one, todo):
            if verbose or r.returncode:
                print(r.stdout + r.stderr)
            if r.returncode:
                raise subprocess.CalledProcessError(r.returncode, "nvcc -c %s" % src)
    objs = [_obj_of(s) for s in sources()]
    subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs,
                   check=True, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
