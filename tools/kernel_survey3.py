"""Third sweep: nests where ONE generic-kernel thread could end up walking a long reduced loop (scratch tool, GPU box)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import numcl_b200 as nc
from numcl_b200 import numeric as nn

L = nc.lib()


def R(shape, dt="f32", seed=1):
    return nc.random_array(shape, dt, seed, 0, -1.0, 1.0) if dt in ("f32", "f64") else nc.random_array(shape, dt, seed, 0, 0, 100)


def run(name, f):
    try:
        f(); L.ncl_sync()
    except Exception as e:  # noqa: BLE001
        print(f"{name:62s} ERROR {type(e).__name__}: {str(e)[:80]}", flush=True); return
    k = L.ncl_last_kernel().decode()
    L.ncl_sync(); t0 = time.perf_counter(); f(); L.ncl_sync(); t = time.perf_counter() - t0
    print(f"{name:62s} {t*1e3:10.3f} ms  {k}{'   <-- GENERIC' if k == 'generic_nest' else ''}", flush=True)


a = R((4096, 4096)); fx = R((4096, 4096), "fixnum", 3); u8 = R((4096, 4096), "ub8", 4); d = R((4096, 4096), "f64", 5)
out64 = nc.zeros((), type="f64"); outf = nc.zeros((), type="f32"); rows64 = nc.zeros((4096,), type="f64")
run("sum f32 -> supplied f64 scalar (einsum ij -> )", lambda: nc.einsum("ij -> ", a, out64))
run("sum ub8 -> supplied f32 scalar", lambda: nc.einsum("ij -> ", u8, outf))
run("row sums f32 -> supplied f64 vector", lambda: nc.einsum("ij -> i", a, rows64))
run("trace ii -> (4096)", lambda: nc.einsum("ii -> ", a))
run("trace of a 16M-diagonal? ii -> i then sum", lambda: nc.sum(nc.einsum("ii -> i", a)))
run("einsum (i :step 2) j -> j  strided rows", lambda: nc.einsum("ij -> (+ @1 $1) -> j", a))
run("einsum ij -> (max @1 $1) -> i  (row max via einsum)", lambda: nc.einsum("ij -> (max @1 $1) -> i", a))
run("einsum ij -> (+ @1 (abs $1)) ->  (L1 norm)", lambda: nc.einsum("ij -> (+ @1 (abs $1)) -> ", a))
run("einsum ij -> (max @1 (abs $1)) ->  (inf norm)", lambda: nc.einsum("ij -> (max @1 (abs $1)) -> ", a))
run("einsum ij ij -> (+ @1 (abs (- $1 $2))) ->  (L1 distance)", lambda: nc.einsum("ij ij -> (+ @1 (abs (- $1 $2))) -> ", a, a))
run("einsum i i i ->  (triple product)", lambda: nc.einsum("i i i -> ", a.reshape((4096 * 4096,)), a.reshape((4096 * 4096,)), a.reshape((4096 * 4096,))))
run("einsum ij j ->  (sum of row-scaled)", lambda: nc.einsum("ij j -> ", a, R((4096,), seed=9)))
run("einsum ij i j ->  (bilinear form)", lambda: nc.einsum("ij i j -> ", a, R((4096,), seed=8), R((4096,), seed=9)))
run("prod f32 all", lambda: nc.prod(a))
run("amax fixnum all", lambda: nc.amax(fx))
run("mean ub8 all", lambda: nc.mean(u8))
run("variance f64 all", lambda: nc.variance(d))
run("sum of (a < 0.5) * a  (masked sum)", lambda: nc.sum(nc.mul(nn.lt(a, 0.5), a)))
