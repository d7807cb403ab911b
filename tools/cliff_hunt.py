"""Throughput of assorted everyday shapes (scratch tool): looks for fast-path cliffs.

Each case is issued `iters` times back to back on the library stream and synchronised once, so the figure is
max(kernel time, host time per call); a single call timed alone ("lat") shows the host overhead."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import numcl_b200 as nc
from numcl_b200 import numeric as nn

PEAK = 6450.3
torch.cuda.init()
L = nc.lib()
only = sys.argv[1:] or None


def run(name, fn, nbytes, iters=6):
    if only and not any(o in name for o in only):
        return
    try:
        fn(); L.ncl_sync()
    except Exception as e:                                     # noqa: BLE001
        print(f"{name:52s} ERROR {type(e).__name__}: {e}", flush=True)
        return
    kern = L.ncl_last_kernel().decode()
    L.ncl_sync(); t0 = time.perf_counter(); fn(); L.ncl_sync(); lat = time.perf_counter() - t0
    best = 1e9
    for _ in range(3):
        L.ncl_sync(); t0 = time.perf_counter()
        for _ in range(iters):
            fn()
        L.ncl_sync(); best = min(best, (time.perf_counter() - t0) / iters)
    t = best
    frac = nbytes / t / 1e9 / PEAK
    flag = "  <-- slow" if frac < 0.5 else ""
    print(f"{name:52s} {t*1e6:9.1f} us (lat {lat*1e6:7.1f}) {nbytes/t/1e9:7.0f} GB/s {frac:6.3f}  {kern}{flag}", flush=True)


def R(shape, dt="f32", seed=1):
    return nc.random_array(shape, dt, seed, 0, -1.0, 1.0) if dt in ("f32", "f64") else nc.random_array(shape, dt, seed, 0, 0, 100)


a = R((1 << 26, 3)); v = R((3,), seed=2); w = R((1 << 26, 1), seed=3)
run("(64M,3) + (3,)", lambda: nc.add(a, v), 2 * a.nbytes)
run("(64M,3) < (3,) -> bit", lambda: nn.lt(a, v), a.nbytes + a.nbytes // 32)
run("(64M,3) * (64M,1)", lambda: nc.mul(a, w), 2 * a.nbytes + w.nbytes)
run("sum (64M,3) axis 1", lambda: nc.sum(a, axes=1), a.nbytes * 4 // 3)
run("sum (64M,3) axis 0", lambda: nc.sum(a, axes=0), a.nbytes)
run("transpose (64M,3)", lambda: nc.transpose(a), 2 * a.nbytes)
del a, w
a = R((3, 1 << 26)); c = R((3, 1), seed=3)
run("(3,64M) + (3,1)", lambda: nc.add(a, c), 2 * a.nbytes)
run("sum (3,64M) axis 0", lambda: nc.sum(a, axes=0), a.nbytes * 4 // 3)
run("sum (3,64M) axis 1", lambda: nc.sum(a, axes=1), a.nbytes)
run("transpose (3,64M)", lambda: nc.transpose(a), 2 * a.nbytes)
del a
a = R((32767, 8193)); v = R((8193,), seed=2)
run("(32767,8193) + (8193,)", lambda: nc.add(a, v), 2 * a.nbytes)
run("sum (32767,8193) axis 1", lambda: nc.sum(a, axes=1), a.nbytes)
run("sum (32767,8193) axis 0", lambda: nc.sum(a, axes=0), a.nbytes)
run("transpose (32767,8193)", lambda: nc.transpose(a), 2 * a.nbytes)
del a
a = R((1 << 22, 48)); a2 = R((1 << 21, 100), seed=5)
run("sum (4M,48) axis 1", lambda: nc.sum(a, axes=1), a.nbytes)
run("sum (4M,48) axis 0", lambda: nc.sum(a, axes=0), a.nbytes)
run("sum (2M,100) axis 1", lambda: nc.sum(a2, axes=1), a2.nbytes)
run("sum (2M,100) axis 0", lambda: nc.sum(a2, axes=0), a2.nbytes)
run("transpose (4M,48)", lambda: nc.transpose(a), 2 * a.nbytes)
del a, a2
a = R((1024, 512, 512)); b1 = R((1024, 1, 512), seed=4); b2 = R((1, 512, 1), seed=5)
run("(1024,512,512) + (1024,1,512)", lambda: nc.add(a, b1), 2 * a.nbytes)
run("(1024,512,512) * (1,512,1)", lambda: nc.mul(a, b2), 2 * a.nbytes)
run("sum (1024,512,512) axis 1", lambda: nc.sum(a, axes=1), a.nbytes)
run("sum (1024,512,512) axis 0", lambda: nc.sum(a, axes=0), a.nbytes)
run("sum (1024,512,512) axes (0,2)", lambda: nc.sum(a, axes=(0, 2)), a.nbytes)
run("sum (1024,512,512) axes (1,2)", lambda: nc.sum(a, axes=(1, 2)), a.nbytes)
run("amax (1024,512,512) axis 2", lambda: nc.amax(a, axes=2), a.nbytes)
run("mean (1024,512,512) axis 2", lambda: nc.mean(a, axes=2), a.nbytes)
run("einsum abc -> cab", lambda: nc.einsum("abc -> cab", a), 2 * a.nbytes)
run("einsum abc -> acb", lambda: nc.einsum("abc -> acb", a), 2 * a.nbytes)
run("einsum abc -> bac", lambda: nc.einsum("abc -> bac", a), 2 * a.nbytes)
run("einsum abc -> cba", lambda: nc.einsum("abc -> cba", a), 2 * a.nbytes)
run("sqrt (map)", lambda: nn.sqrt(a), 2 * a.nbytes)
run("exp (map)", lambda: nn.exp(a), 2 * a.nbytes)
run("a * 2.0 (scalar)", lambda: nc.mul(a, 2.0), 2 * a.nbytes)
run("a < b1 -> bit", lambda: nn.lt(a, b1), a.nbytes + a.nbytes // 32)
a2 = R((1024, 512, 512), seed=77)
run("a < a2 same shape -> bit", lambda: nn.lt(a, a2), 2 * a.nbytes + a.nbytes // 32)
del a2
run("a < 0.5 (scalar) -> bit", lambda: nn.lt(a, 0.5), a.nbytes + a.nbytes // 32)
d = R((512, 512, 512), "f64", 6); ah = R((512, 512, 512), seed=30)
run("f32 + f64 same shape -> f64", lambda: nc.add(ah, d), ah.nbytes + 2 * d.nbytes)
d2 = R((512, 512, 512), "f64", 61)
run("f64 + f64", lambda: nc.add(d, d2), 3 * d.nbytes)
del d2
del d, ah
u = R((1024, 512, 512), "ub8", 7); u2 = R((1024, 512, 512), "ub8", 8)
run("ub8 + ub8 -> ub15", lambda: nc.add(u, u2), u.nbytes * 4)
run("ub8 * 3 (scalar)", lambda: nc.mul(u, 3), u.nbytes * 3)
run("ub8 max ub8 -> ub8", lambda: nc.maximum(u, u2), u.nbytes * 3)
run("ub8 < ub8 -> bit", lambda: nn.lt(u, u2), u.nbytes * 2 + u.nbytes // 8)
run("astype ub8 -> f32", lambda: nc.astype(u, "f32"), u.nbytes * 5)
run("ub8 + f32 -> f32", lambda: nc.add(u, a), u.nbytes + 2 * a.nbytes)
run("sum ub8 (1024,512,512) axis 2 -> fixnum", lambda: nc.sum(u, axes=2), u.nbytes)
run("sum ub8 all", lambda: nc.sum(u), u.nbytes)
del u, u2, a
fx = R((8192, 8192), "fixnum", 9); fy = R((8192, 8192), "fixnum", 10)
run("fixnum + fixnum", lambda: nc.add(fx, fy), 3 * fx.nbytes)
run("sum fixnum axis 0", lambda: nc.sum(fx, axes=0), fx.nbytes)
del fx, fy
x = R((32768, 8192), seed=9); y = R((32768, 8192), seed=10)
run("einsum ij ij -> i (rowwise dot)", lambda: nc.einsum("ij ij -> i", x, y), 2 * x.nbytes)
run("einsum ij ij -> (full dot)", lambda: nc.einsum("ij ij -> ", x, y), 2 * x.nbytes)
run("einsum ij ij -> j (colwise dot)", lambda: nc.einsum("ij ij -> j", x, y), 2 * x.nbytes)
del y
yv = R((8192,), seed=14); xv = R((32768,), seed=15)
run("matvec ij j -> i 32768x8192", lambda: nc.einsum("ij j -> i", x, yv), x.nbytes)
run("vecmat i ij -> j 32768x8192", lambda: nc.einsum("i ij -> j", xv, x), x.nbytes)
del x
x1 = R((16384,), seed=11); y1 = R((16384,), seed=12)
run("outer (i j -> ij) 16384x16384", lambda: nc.outer(x1, y1), 16384 * 16384 * 4)
v1 = R((1 << 27,), seed=16); v2 = R((1 << 27,), seed=17)
run("vdot 128M", lambda: nc.vdot(v1, v2), 2 * v1.nbytes)
run("inner 128M", lambda: nc.inner(v1, v2), 2 * v1.nbytes)
del v1, v2
s1 = R((1000,), seed=20); s2 = R((1000,), seed=21)
run("tiny add (1000)", lambda: nc.add(s1, s2), 12000, iters=200)
run("tiny sum (1000)", lambda: nc.sum(s1), 4000, iters=200)
