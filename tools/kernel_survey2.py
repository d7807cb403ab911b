"""Second sweep of tools/kernel_survey.py: reductions, einsum forms, map forms, conversions (scratch tool, GPU box)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import numcl_b200 as nc
from numcl_b200 import numeric as nn

L = nc.lib()


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


a = R((4096, 4096)); b = R((4096, 4096), seed=2); fx = R((4096, 4096), "fixnum", 3); u8 = R((4096, 4096), "ub8", 4); d = R((4096, 4096), "f64", 5)
nb = a.nbytes
run("sum fixnum all", lambda: nc.sum(fx), fx.nbytes)
run("amin ub8 all", lambda: nc.amin(u8), u8.nbytes)
run("prod f64 axis 0", lambda: nc.prod(d, axes=0), d.nbytes)
run("mean f32 axes (0,1)", lambda: nc.mean(a), nb)
run("variance f32 axis 1", lambda: nc.variance(a, axes=1), nb)
run("stdev f32 axis 0", lambda: nc.stdev(a, axes=0), nb)
c3 = R((256, 256, 256)); 
run("einsum ijk -> k", lambda: nc.einsum("ijk -> k", c3), c3.nbytes)
run("einsum ijk -> j", lambda: nc.einsum("ijk -> j", c3), c3.nbytes)
run("einsum ijk -> ik", lambda: nc.einsum("ijk -> ik", c3), c3.nbytes)
v = R((4096,), seed=7)
run("einsum ij j -> ij (row scale)", lambda: nc.einsum("ij j -> ij", a, v), 2 * nb)
run("einsum ij i -> ij (col scale)", lambda: nc.einsum("ij i -> ij", a, v), 2 * nb)
run("einsum ij ij -> ij (hadamard)", lambda: nc.einsum("ij ij -> ij", a, b), 3 * nb)
run("einsum ij jk -> ik fixnum 1024", lambda: nc.einsum("ij jk -> ik", R((1024, 1024), "fixnum", 8), R((1024, 1024), "fixnum", 9)), 1)
run("einsum ij jk -> ik ub8 -> fixnum 1024", lambda: nc.einsum("ij jk -> ik", R((1024, 1024), "ub8", 8), R((1024, 1024), "ub8", 9)), 1)
run("einsum bij bjk -> bik f64 (64,256,256)", lambda: nc.einsum("bij bjk -> bik", R((64, 256, 256), "f64", 8), R((64, 256, 256), "f64", 9)), 1)
run("einsum ij jk -> ik f32 x f64 (mixed) 512", lambda: nc.einsum("ij jk -> ik", R((512, 512), seed=8), R((512, 512), "f64", 9)), 1)
run("einsum (i :step 2) stride sum", lambda: nc.einsum("i -> (+ @1 $1) -> ", v), v.nbytes)
run("map (* $1 $1)", lambda: nc.map_array("(* $1 $1)", a), 2 * nb)
run("map (+ $1 1)", lambda: nc.map_array("(+ $1 1)", a), 2 * nb)
run("map (max $1 $2)", lambda: nc.map_array("max", a, b), 3 * nb)
run("map (sqrt (+ (* $1 $1) (* $2 $2))) composite", lambda: nc.map_array("(sqrt (+ (* $1 $1) (* $2 $2)))", a, b), 3 * nb)
run("map (* 2.0 (+ $1 $2)) composite", lambda: nc.map_array("(* 2.0 (+ $1 $2))", a, b), 3 * nb)
run("ub8 < 50 (scalar)", lambda: nn.lt(u8, 50), u8.nbytes)
run("fixnum < fixnum", lambda: nn.lt(fx, fx), 2 * fx.nbytes)
run("f32 == f64 (mixed)", lambda: nn.eq(a, d), a.nbytes + d.nbytes)
run("astype ub8 -> fixnum", lambda: nc.astype(u8, "fixnum"), 9 * u8.nbytes)
run("astype fixnum -> ub8 (wrap)", lambda: nc.astype(fx, "ub8"), fx.nbytes * 9 // 8)
run("astype f64 -> ub8 (round)", lambda: nc.astype(d, "ub8"), d.nbytes * 9 // 8)
run("copy f32", lambda: nc.copy(a), 2 * nb)
run("a + b + a (3-ary fold)", lambda: nc.add(a, b, a), 4 * nb)
run("maximum(a, b, a)", lambda: nc.maximum(a, b, a), 4 * nb)
run("transpose 4096^2 f32", lambda: nc.transpose(a), 2 * nb)
run("transpose 4096^2 ub8", lambda: nc.transpose(u8), u8.nbytes * 9)
run("diag(4096)", lambda: nc.diag(a), 4096 * 8)
run("outer fixnum 4096", lambda: nc.outer(R((4096,), "fixnum", 3), R((4096,), "fixnum", 4)), 4096 * 4096 * 8)
