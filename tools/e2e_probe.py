"""End-to-end time of one call on PINNED HOST arrays (what a numcl user's call costs: copy in, compute, copy out), with the
panel pipeline (default) or without it (NUMCL_NO_PIPELINE=1), panel size NUMCL_PANEL_MB.
usage: python tools/e2e_probe.py [case ...]     cases: c1 c2a c2a_resident c3d c3s c4 c5a"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numcl_b200 as nc
from numcl_b200 import _lib

L = nc.lib()


def H(shape, dt, seed, lo=-1.0, hi=1.0):
    d = nc.random_array(shape, dt, seed, 0, lo, hi)
    h = nc.NArray(shape, dt, where="host")
    h.set_raw(d.raw())
    return h


def timed(f, n=5):
    f(); f(); L.ncl_sync()
    st0 = _lib.transfer_stats() + _lib.pipeline_stats()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); L.ncl_sync(); ts.append(time.perf_counter() - t0)
    st1 = _lib.transfer_stats() + _lib.pipeline_stats()
    return sorted(ts)[len(ts) // 2] * 1e3, [(b - a) / n for a, b in zip(st0, st1)]


cases = sys.argv[1:] or ["c1", "c2a", "c2a_resident", "c5a", "c4", "c3s", "c3d"]
out = {"pipeline": os.environ.get("NUMCL_NO_PIPELINE", "0") != "1", "panel_mb": os.environ.get("NUMCL_PANEL_MB", "32")}
for c in cases:
    if c == "c1":
        a, b = H((4096, 4096), "f32", 1), H((4096,), "f32", 2)
        f = lambda: nc.add(a, b)
    elif c == "c2a":
        a = H((16384, 16384), "f32", 3); f = lambda: nc.sum(a, axes=1)
    elif c == "c2a_resident":
        a = H((16384, 16384), "f32", 3).make_resident()
        def f():
            a.dirty(); nc.sum(a, axes=1); nc.sum(a)
    elif c == "c5a":
        x, y = H((64, 32, 32, 256), "ub8", 4, 0, 255), H((256,), "fixnum", 5, -2 ** 40, 2 ** 40); f = lambda: nc.mul(x, y)
    elif c == "c4":
        A, B = H((256, 512, 512), "f32", 6), H((256, 512, 512), "f32", 7); f = lambda: nc.einsum("bij bjk -> bik", A, B)
    elif c == "c3s":
        A, B = H((8192, 8192), "f32", 8), H((8192, 8192), "f32", 9); f = lambda: nc.matmul(A, B)
    elif c == "c3d":
        A, B = H((8192, 8192), "f64", 8), H((8192, 8192), "f64", 9); f = lambda: nc.matmul(A, B)
    else:
        raise SystemExit(c)
    ms, (h2d, d2h, pc, pp) = timed(f)
    out[c] = {"ms": round(ms, 3), "h2d_MiB": round(h2d / 2 ** 20, 2), "d2h_MiB": round(d2h / 2 ** 20, 2), "serial_pcie_ms_at_55GBs": round((h2d + d2h) / 55e6, 2),
              "duplex_bound_ms": round(max(h2d, d2h) / 55e6, 2), "pipelined": pc, "panels": pp, "kernel": L.ncl_last_kernel().decode()}
    del f
print(json.dumps(out))
