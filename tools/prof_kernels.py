"""Launches each hot-path kernel a few times (device resident) so ncu can capture them.
usage: python tools/prof_kernels.py [c1 c5a c5b c2a c2b c4 c3s c3d]"""
import ctypes as C
import sys
sys.path.insert(0, ".")
import numcl_b200 as nc
from numcl_b200 import _lib
from numcl_b200.einsum import _memoized

which = sys.argv[1:] or ["c1", "c5a", "c5b", "c2a", "c2b"]
L = nc.lib()
import os
REP = int(os.environ.get("REP", "3"))


def fold(op, arrs, r):
    ts = (_lib.Tensor * len(arrs))(*[a.tensor() for a in arrs])
    tr = r.tensor()
    for _ in range(REP):
        _lib.check(L.ncl_fold(op, 1, ts, len(arrs), C.byref(tr)))


def einsum(spec, ins, out):
    _, d = _memoized(spec)
    tin = (_lib.Tensor * len(ins))(*[a.tensor() for a in ins])
    tout = (_lib.Tensor * 1)(out.tensor())
    for _ in range(REP):
        _lib.check(L.ncl_einsum(C.byref(d), tin, len(ins), tout, 1, 1))


if "c1" in which:
    a = nc.random_array((4096, 4096), "f32", 1, 2, -1.0, 1.0); b = nc.random_array((4096,), "f32", 2, 2, -1.0, 1.0)
    fold(0, [a, b], nc.empty((4096, 4096), "f32"))
if "c5a" in which or "c5b" in which:
    x = nc.random_array((64, 32, 32, 256), "ub8", 4, 0, 0, 255); y = nc.random_array((256,), "fixnum", 5, 0, -(2 ** 40), 2 ** 40)
    p = nc.empty((64, 32, 32, 256), "fixnum")
    fold(2, [x, y], p)
    if "c5b" in which:
        einsum("abcd -> dbca", [p], nc.empty((256, 32, 32, 64), "fixnum"))
if "c2a" in which or "c2b" in which:
    big = nc.random_array((16384, 16384), "f32", 3, 0, -1.0, 1.0)
    tb = big.tensor(); ax = (C.c_int * 1)(1)
    rows = nc.empty((16384,), "f32"); tot = nc.empty((), "f32"); tr, tt = rows.tensor(), tot.tensor()
    for _ in range(REP):
        if "c2a" in which:
            _lib.check(L.ncl_reduce(0, C.byref(tb), ax, 1, C.byref(tr), 1))
        if "c2b" in which:
            _lib.check(L.ncl_reduce(0, C.byref(tb), ax, 0, C.byref(tt), 1))
    del big
if "c4" in which:
    A = nc.random_array((256, 512, 512), "f32", 7, 0, -1.0, 1.0); B = nc.random_array((256, 512, 512), "f32", 8, 0, -1.0, 1.0)
    einsum("bij bjk -> bik", [A, B], nc.empty((256, 512, 512), "f32"))
if "c3s" in which:
    A = nc.random_array((4096, 4096), "f32", 7, 0, -1.0, 1.0); B = nc.random_array((4096, 4096), "f32", 8, 0, -1.0, 1.0)
    einsum("ij jk -> ik", [A, B], nc.empty((4096, 4096), "f32"))
if "c3d" in which:
    A = nc.random_array((4096, 4096), "f64", 7, 0, -1.0, 1.0); B = nc.random_array((4096, 4096), "f64", 8, 0, -1.0, 1.0)
    einsum("ij jk -> ik", [A, B], nc.empty((4096, 4096), "f64"))
L.ncl_sync()
print("ok", L.ncl_launch_count())
