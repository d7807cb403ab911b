"""Device time on odd (unaligned) shapes: no fast-path cliff should be catastrophic (scratch tool)."""
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
    med, best = timed(fn, iters=5, warm=1)
    print(f"{name:40s} {med*1e6:10.1f} us {nbytes/med/1e9:8.0f} GB/s {nbytes/med/1e9/PEAK:6.3f}  {L.ncl_last_kernel().decode()}", flush=True)


n = 16383
a = nc.random_array((n, n), "f32", 1, 0, -1.0, 1.0); b = nc.random_array((n, n), "f32", 2, 0, -1.0, 1.0); v = nc.random_array((n,), "f32", 3, 0, -1.0, 1.0)
r = nc.empty((n, n), "f32"); rows = nc.empty((n,), "f32"); tot = nc.empty((), "f32")
ta, tb, tv, tr, trows, ttot = a.tensor(), b.tensor(), v.tensor(), r.tensor(), rows.tensor(), tot.tensor()
one = n * n * 4
ab = (_lib.Tensor * 2)(ta, tb); av = (_lib.Tensor * 2)(ta, tv)
ax1 = (C.c_int * 1)(1); ax0 = (C.c_int * 1)(0)
run("f32 16383^2 a+b", lambda: L.ncl_fold(0, 1, ab, 2, C.byref(tr)), 3 * one)
run("f32 16383^2 a+rowvec", lambda: L.ncl_fold(0, 1, av, 2, C.byref(tr)), 2 * one)
run("f32 16383^2 sum axis 1", lambda: L.ncl_reduce(0, C.byref(ta), ax1, 1, C.byref(trows), 1), one)
run("f32 16383^2 sum axis 0", lambda: L.ncl_reduce(0, C.byref(ta), ax0, 1, C.byref(trows), 1), one)
run("f32 16383^2 sum all", lambda: L.ncl_reduce(0, C.byref(ta), ax1, 0, C.byref(ttot), 1), one)
from numcl_b200.einsum import _memoized
_, d = _memoized("ij -> ji")
tin = (_lib.Tensor * 1)(ta); tout = (_lib.Tensor * 1)(tr)
run("f32 16383^2 transpose", lambda: L.ncl_einsum(C.byref(d), tin, 1, tout, 1, 1), 2 * one)
del a, b, r
m = 4099
A = nc.random_array((m, m), "f32", 4, 0, -1.0, 1.0); B = nc.random_array((m, m), "f32", 5, 0, -1.0, 1.0); Cc = nc.empty((m, m), "f32")
_, d = _memoized("ij jk -> ik")
tin = (_lib.Tensor * 2)(A.tensor(), B.tensor()); tout = (_lib.Tensor * 1)(Cc.tensor())
med, best = timed(lambda: L.ncl_einsum(C.byref(d), tin, 2, tout, 1, 1), iters=3, warm=1)
print(f"f32 4099^3 matmul {med*1e3:.2f} ms {2*m**3/med/1e12:.1f} TFLOP/s {L.ncl_last_kernel().decode()}")
A = nc.random_array((m, m), "f64", 4, 0, -1.0, 1.0); B = nc.random_array((m, m), "f64", 5, 0, -1.0, 1.0); Cc = nc.empty((m, m), "f64")
tin = (_lib.Tensor * 2)(A.tensor(), B.tensor()); tout = (_lib.Tensor * 1)(Cc.tensor())
med, best = timed(lambda: L.ncl_einsum(C.byref(d), tin, 2, tout, 1, 1), iters=3, warm=1)
print(f"f64 4099^3 matmul {med*1e3:.2f} ms {2*m**3/med/1e12:.1f} TFLOP/s {L.ncl_last_kernel().decode()}")
