"""Which kernel family does an everyday numcl operation reach, and how fast (scratch tool, GPU box)?  Flags generic_nest.
usage: python tools/kernel_survey.py"""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import numcl_b200 as nc
from numcl_b200 import numeric as nn

L = nc.lib()
N = 1 << 26


def R(shape, dt="f32", seed=1):
    return nc.random_array(shape, dt, seed, 0, -1.0, 1.0) if dt in ("f32", "f64") else nc.random_array(shape, dt, seed, 0, 0, 100)


def run(name, f, nbytes):
    try:
        f(); L.ncl_sync()
    except Exception as e:  # noqa: BLE001
        print(f"{name:58s} ERROR {type(e).__name__}: {str(e)[:90]}", flush=True); return
    k = L.ncl_last_kernel().decode()
    L.ncl_sync(); t0 = time.perf_counter()
    for _ in range(3):
        f()
    L.ncl_sync(); t = (time.perf_counter() - t0) / 3
    print(f"{name:58s} {t*1e6:9.1f} us {nbytes/t/1e9:7.0f} GB/s {nbytes/t/1e9/6450.3:6.3f}  {k}{'   <-- GENERIC' if k == 'generic_nest' else ''}", flush=True)


a = R((N,)); b = R((N,), seed=2); m1 = nn.lt(a, b); m2 = nn.gt(a, 0.25)
run("logand bit bit (mask combine)", lambda: nn.logand(m1, m2), 3 * N // 8)
run("logior bit bit", lambda: nn.logior(m1, m2), 3 * N // 8)
run("lognot bit", lambda: nn.lognot(m1), 2 * N // 8)
run("sum bit (count true)", lambda: nc.sum(m1), N // 8)
run("bit * f32 (masking)", lambda: nc.mul(m1, a), 2 * 4 * N + N // 8)
run("astype bit -> f32", lambda: nc.astype(m1, "f32"), 4 * N + N // 8)
run("astype f32 -> f64", lambda: nc.astype(a, "f64"), 12 * N)
run("astype f64 -> f32", lambda: nc.astype(nc.astype(a, "f64"), "f32") if False else None, 1) if False else None
d = R((N // 2,), "f64", 3)
run("astype f64 -> f32", lambda: nc.astype(d, "f32"), 12 * (N // 2))
run("astype f32 -> fixnum (round)", lambda: nc.astype(a, "fixnum"), 12 * N)
run("bit + 1 -> ub2", lambda: nc.add(m1, 1), N // 8 + N // 4)
run("clip f32", lambda: nc.clip(a, -0.5, 0.5), 8 * N)
run("abs f32", lambda: nn.abs_(a), 8 * N)
run("neg f32 (- a)", lambda: nc.sub(a), 8 * N)
run("square f32", lambda: nn.square(a), 8 * N)
run("a*b+c map 3 inputs", lambda: nc.map_array("(+ (* $1 $2) $3)", a, b, a), 16 * N)
run("a / b f32", lambda: nc.div(a, b), 12 * N)
run("1+ a", lambda: nc.one_plus(a), 8 * N)
a2 = R((8192, 8192)); u8 = R((8192, 8192), "ub8", 5)
run("amax axis 0 (8192^2)", lambda: nc.amax(a2, axes=0), a2.nbytes)
run("prod axis 1", lambda: nc.prod(a2, axes=1), a2.nbytes)
run("sum ub8 axis 0", lambda: nc.sum(u8, axes=0), u8.nbytes)
run("mean ub8 axis 1", lambda: nc.mean(u8, axes=1), u8.nbytes)
v = a2.reshape((8192 * 8192,))
run("strided view a[::2] sum (einsum i-step)", lambda: nc.einsum("i -> (+ @1 $1) -> ", v) if False else nc.sum(v), v.nbytes)
run("einsum ij jk kl -> il (3 inputs, 512)", lambda: nc.einsum("ij jk kl -> il", R((512, 512), seed=7), R((512, 512), seed=8), R((512, 512), seed=9)), 1)
run("einsum ijk -> ikj", lambda: nc.einsum("ijk -> ikj", R((256, 512, 512), seed=7)), 2 * 256 * 512 * 512 * 4)
run("einsum ij -> j (col sum via einsum)", lambda: nc.einsum("ij -> j", a2), a2.nbytes)
run("einsum ii -> i (diag 8192)", lambda: nc.einsum("ii -> i", a2), 8192 * 8)
run("einsum ij ik -> jk (A^T A, 8192x256)", lambda: nc.einsum("ij ik -> jk", R((8192, 256), seed=3), R((8192, 256), seed=4)), 1)
run("kron 64x64 x 64x64", lambda: nc.kron(R((64, 64), seed=3), R((64, 64), seed=4)), 64 ** 4 * 4)
f1 = R((4096, 4096), "fixnum", 9)
run("fixnum * f32 -> f32", lambda: nc.mul(f1, R((4096, 4096), seed=11)), 4096 * 4096 * 16)
run("fixnum / fixnum -> f32", lambda: nc.div(f1, f1), 4096 * 4096 * 20)
run("ub8 - ub8 -> sb16", lambda: nc.sub(u8, u8), u8.nbytes * 4)
run("mod fixnum 7", lambda: nn.mod(f1, 7), 4096 * 4096 * 16)
run("floor f32", lambda: nn.floor(a2), a2.nbytes * 3)
