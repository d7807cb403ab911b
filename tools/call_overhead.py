"""Host cost of one C-ABI call on tiny device-resident arrays (what a numcl program that works on small arrays pays per
operation): N calls enqueued back to back, one sync at the end.  usage: python tools/call_overhead.py"""
import ctypes as C
import sys
import time

sys.path.insert(0, ".")
import numcl_b200 as nc
from numcl_b200 import _lib

L = nc.lib()
a = nc.random_array((1000,), "f32", 1, 0, -1.0, 1.0)
b = nc.random_array((1000,), "f32", 2, 0, -1.0, 1.0)
r = nc.empty((1000,), "f32")
s = nc.empty((), "f32")
m = nc.random_array((32, 32), "f32", 3, 0, -1.0, 1.0)
mr = nc.empty((32, 32), "f32")
args = (_lib.Tensor * 2)(a.tensor(), b.tensor())
tr, ts, ta = r.tensor(), s.tensor(), a.tensor()
ax = (C.c_int * 1)(0)
margs = (_lib.Tensor * 2)(m.tensor(), m.tensor())
tmr = mr.tensor()


def bench(name, f, n=2000):
    for _ in range(20):
        f()
    L.ncl_sync()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    t1 = time.perf_counter()
    L.ncl_sync()
    t2 = time.perf_counter()
    print(f"{name:44s} enqueue {(t1 - t0) / n * 1e6:6.2f} us/call   with drain {(t2 - t0) / n * 1e6:6.2f} us/call   {L.ncl_last_kernel().decode()}", flush=True)


bench("ncl_fold + (1000) f32, identity", lambda: L.ncl_fold(_lib.OPS["ADD"], 1, args, 2, C.byref(tr)))
bench("ncl_broadcast * (1000) f32", lambda: L.ncl_broadcast(_lib.OPS["MUL"], C.byref(args[0]), C.byref(args[1]), C.byref(tr)))
bench("ncl_reduce sum (1000) f32 -> scalar", lambda: L.ncl_reduce(_lib.OPS["ADD"], C.byref(ta), ax, 0, C.byref(ts), _lib.OUT_FRESH))
from numcl_b200.einsum import _memoized
specs, desc = _memoized("ij jk -> ik")
bench("ncl_einsum (ij jk -> ik) 32x32 f32", lambda: L.ncl_einsum(C.byref(desc), margs, 2, C.byref(tmr), 1, _lib.OUT_FRESH))
bench("python nc.add (1000)", lambda: nc.add(a, b), 500)
bench("python nc.sum (1000) -> number", lambda: nc.sum(a), 500)
