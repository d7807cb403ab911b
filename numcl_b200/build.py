"""Builds numcl_b200/libnumcl_cuda.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

  python -m numcl_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
SO = os.path.join(HERE, "libnumcl_cuda.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

# No --use_fast_math, no -ftz: float elementwise ops must be IEEE round-to-nearest with denormals
# (SURVEY.md A.2).  -fmad=false keeps the compiler from contracting a*b+c in the bit-exact paths;
# the GEMM kernels ask for FMA explicitly (tensor cores / __fma_rn).
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC", "-fmad=false", "--expt-relaxed-constexpr", "-Xptxas", "-v",
         "-diag-suppress", "177,550",
         # the library is several hundred kernel instantiations with -lineinfo: 91 MB of fatbin with the default compression,
         # 15 MB with this one (the driver decompresses a kernel's image when it is first used)
         "--compress-mode=size"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


_INC = re.compile(r'^\s*#\s*include\s+"([^"]+)"', re.M)


def _deps_mtime(src, _seen=None):
    """newest mtime among `src` and the local headers it includes, transitively (a header change rebuilds only the
    translation units that see it: the kernel instantiation files take minutes)"""
    seen = _seen if _seen is not None else set()
    path = os.path.normpath(os.path.join(CSRC, src))
    if path in seen or not os.path.exists(path):
        return 0.0
    seen.add(path)
    m = os.path.getmtime(path)
    for inc in _INC.findall(open(path, errors="replace").read()):
        m = max(m, _deps_mtime(os.path.join(os.path.dirname(src), inc), seen))
    return m


def _compile(src, verbose):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    cmd = [NVCC, *FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(OBJ, src[:-3] + ".log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stderr[-6000:]}")
    if verbose:
        print(f"[build] {src} ok")
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    todo, objs = [], []
    for src in _sources():
        obj = os.path.join(OBJ, src[:-3] + ".o")
        objs.append(obj)
        sm = _deps_mtime(src)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < sm:
            todo.append(src)
    if todo:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            list(ex.map(lambda s: _compile(s, verbose), todo))
    if todo or not os.path.exists(SO) or any(os.path.getmtime(o) > os.path.getmtime(SO) for o in objs):
        cmd = [NVCC, "-shared", "-o", SO, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
               "-cudart", "static", "-ldl", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stderr[-4000:])
        if verbose:
            print("[build] linked", SO)
    return SO


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
