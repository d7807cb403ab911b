"""Device time of assorted elementwise shapes on 1 GiB-class operands (scratch tool; bench.py is the contract)."""
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
import numcl_b200 as nc
from numcl_b200 import _lib
from tools.quick_perf import timed

PEAK = 6450.3
torch.cuda.init()
L = nc.lib()


def run(name, fn, nbytes):
    med, best = timed(fn, iters=10, warm=2)
    print(f"{name:44s} {med*1e6:9.1f} us  {nbytes/med/1e9:7.0f} GB/s  {nbytes/med/1e9/PEAK:.3f}  {L.ncl_last_kernel().decode()}", flush=True)


n = 16384
for dt, sz in (("f32", 4), ("f64", 8), ("fixnum", 8)):
    m = n if sz == 4 else n // 2
    lo, hi = (-1.0, 1.0) if dt != "fixnum" else (-1000, 1000)
    a = nc.random_array((m, n), dt, 1, 0, lo, hi); b = nc.random_array((m, n), dt, 2, 0, lo, hi); v = nc.random_array((n,), dt, 3, 0, lo, hi)
    r = nc.empty((m, n), dt)
    ta, tb, tv, tr = a.tensor(), b.tensor(), v.tensor(), r.tensor()
    ab = (_lib.Tensor * 2)(ta, tb); av = (_lib.Tensor * 2)(ta, tv)
    one = m * n * sz
    run(f"{dt} a+b same shape (numcl:+, fold)", lambda: L.ncl_fold(0, 1, ab, 2, C.byref(tr)), 3 * one)
    run(f"{dt} a*b same shape (broadcast)", lambda: L.ncl_broadcast(_lib.OPS["MUL"], C.byref(ta), C.byref(tb), C.byref(tr)), 3 * one)
    run(f"{dt} a+rowvec (numcl:+, fold)", lambda: L.ncl_fold(0, 1, av, 2, C.byref(tr)), 2 * one)
    run(f"{dt} copy/astype same type (ncl_convert)", lambda: L.ncl_convert(C.byref(ta), C.byref(tr)), 2 * one)
    if dt != "fixnum":
        from numcl_b200.einsum import _memoized
        _, d = _memoized("i -> (sqrt $1) -> i")
        fa = a.flatten(); fr = r.flatten()
        tin = (_lib.Tensor * 1)(fa.tensor()); tout = (_lib.Tensor * 1)(fr.tensor())
        run(f"{dt} sqrt (einsum i -> (sqrt $1) -> i)", lambda: L.ncl_einsum(C.byref(d), tin, 1, tout, 1, 1), 2 * one)
    del a, b, r
x = nc.random_array((n, n), "ub8", 4, 0, 0, 255); y = nc.random_array((n,), "fixnum", 5, 0, -1000, 1000); r = nc.empty((n, n), "fixnum")
xy = (_lib.Tensor * 2)(x.tensor(), y.tensor()); tr = r.tensor()
run("ub8 * fixnum rowvec -> fixnum (widen)", lambda: L.ncl_fold(2, 1, xy, 2, C.byref(tr)), n * n * 9)
r32 = nc.empty((n, n), "f32"); f = nc.random_array((n, n), "f32", 6, 0, -1.0, 1.0)
xf = (_lib.Tensor * 2)(x.tensor(), f.tensor()); t32 = r32.tensor()
run("ub8 * f32 same shape -> f32", lambda: L.ncl_fold(2, 1, xf, 2, C.byref(t32)), n * n * 9)
