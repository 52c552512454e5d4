"""Builds libnagp.so (hand-written sm_100a kernels + the C ABI) in-tree with nvcc.

`python -m nowcastautogp_b200.build [--force] [--verbose]`. nvcc cross-compiles without a GPU.
Each translation unit is compiled to an object under `csrc/_obj/` (git-ignored `*.o`), in parallel, and
only when it or a header changed; the objects are linked into `libnagp.so`.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libnagp.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ARCH + ["-lineinfo", "-O3", "-std=c++17", "-fmad=false", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _headers():
    return glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "nagp.h")]


def _newer(deps, target) -> bool:
    if not os.path.exists(target):
        return True
    mt = os.path.getmtime(target)
    return any(os.path.getmtime(d) > mt for d in deps)


def _compile(src: str, verbose: bool):
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    cmd = [NVCC] + CFLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
    res = subprocess.run(cmd, capture_output=True, text=True)
    return src, obj, res


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _headers()
    todo, objs = [], []
    for src in sources():
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or _newer([src] + hdrs, obj):
            todo.append(src)
    if not todo and not _newer(objs, LIB):
        return LIB
    failed = False
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(todo)))) as ex:
        for src, obj, res in ex.map(lambda s: _compile(s, verbose), todo):
            if verbose or res.returncode:
                sys.stderr.write(f"--- {os.path.basename(src)}\n" + res.stdout + res.stderr)
            failed = failed or res.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libnagp.so")
    # drop objects of sources that no longer exist
    for stale in set(glob.glob(os.path.join(OBJ, "*.o"))) - set(objs):
        os.remove(stale)
    res = subprocess.run([NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-ldl"], capture_output=True, text=True)
    if res.returncode:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libnagp.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
