"""Throughput of assorted everyday shapes (scratch tool): looks for fast-path cliffs."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import numcl_b200 as nc

PEAK = 6450.3
torch.cuda.init()
L = nc.lib()


def run(name, fn, nbytes, iters=5):
    fn(); L.ncl_sync()
    ts = []
    for _ in range(iters):
        L.ncl_sync(); t0 = time.perf_counter(); fn(); L.ncl_sync(); ts.append(time.perf_counter() - t0)
    t = sorted(ts)[len(ts) // 2]
    flag = "  <-- slow" if nbytes / t / 1e9 / PEAK < 0.3 else ""
    print(f"{name:52s} {t*1e6:10.1f} us {nbytes/t/1e9:8.0f} GB/s {nbytes/t/1e9/PEAK:6.3f}  {L.ncl_last_kernel().decode()}{flag}", flush=True)


R = lambda shape, dt="f32", seed=1: nc.random_array(shape, dt, seed, 0, -1.0, 1.0) if dt in ("f32", "f64") else nc.random_array(shape, dt, seed, 0, 0, 100)
a = R((1 << 24, 3)); v = R((3,), seed=2)
run("(16M,3) + (3,)", lambda: nc.add(a, v), 2 * a.nbytes)
run("sum (16M,3) axis 1", lambda: nc.sum(a, axes=1), a.nbytes * 4 // 3)
run("sum (16M,3) axis 0", lambda: nc.sum(a, axes=0), a.nbytes)
run("transpose (16M,3)", lambda: nc.transpose(a), 2 * a.nbytes)
del a
a = R((3, 1 << 24)); c = R((3, 1), seed=3)
run("(3,16M) + (3,1)", lambda: nc.add(a, c), 2 * a.nbytes)
run("sum (3,16M) axis 0", lambda: nc.sum(a, axes=0), a.nbytes * 4 // 3)
del a
a = R((256, 512, 512)); b1 = R((256, 1, 512), seed=4); b2 = R((1, 512, 1), seed=5)
run("(256,512,512) + (256,1,512)", lambda: nc.add(a, b1), 2 * a.nbytes)
run("(256,512,512) * (1,512,1)", lambda: nc.mul(a, b2), 2 * a.nbytes)
run("sum (256,512,512) axis 1", lambda: nc.sum(a, axes=1), a.nbytes)
run("sum (256,512,512) axes (0,2)", lambda: nc.sum(a, axes=(0, 2)), a.nbytes)
run("amax (256,512,512) axis 2", lambda: nc.amax(a, axes=2), a.nbytes)
run("einsum abc -> cab", lambda: nc.einsum("abc -> cab", a), 2 * a.nbytes)
run("einsum abc -> acb", lambda: nc.einsum("abc -> acb", a), 2 * a.nbytes)
run("einsum abc -> bac", lambda: nc.einsum("abc -> bac", a), 2 * a.nbytes)
run("sqrt (map)", lambda: nc.numeric.sqrt(a), 2 * a.nbytes)
d = R((256, 512, 512), "f64", 6)
run("f32 + f64 same shape -> f64", lambda: nc.add(a, d), a.nbytes + 2 * d.nbytes)
del d
u = R((256, 512, 512), "ub8", 7); u2 = R((256, 512, 512), "ub8", 8)
run("ub8 + ub8 -> ub15", lambda: nc.add(u, u2), u.nbytes * 4)
run("ub8 * 3 (scalar) -> ub15?", lambda: nc.mul(u, 3), u.nbytes * 3)
run("ub8 < ub8 -> bit", lambda: nc.lt(u, u2), u.nbytes * 2 + u.nbytes // 8)
run("astype ub8 -> f32", lambda: nc.astype(u, "f32"), u.nbytes * 5)
run("einsum ij ij -> i (rowwise dot) f32 16384x4096", (lambda x=R((16384, 4096), seed=9), y=R((16384, 4096), seed=10): (lambda: nc.einsum("ij ij -> i", x, y)))(), 2 * 16384 * 4096 * 4)
run("outer (i j -> ij) 16384x16384", (lambda x=R((16384,), seed=11), y=R((16384,), seed=12): (lambda: nc.outer(x, y)))(), 16384 * 16384 * 4)
run("matvec ij j -> i 16384x16384", (lambda x=R((16384, 16384), seed=13), y=R((16384,), seed=14): (lambda: nc.einsum("ij j -> i", x, y)))(), 16384 * 16384 * 4)
