"""Quick device-side timings of the hot-path kernels (CUDA events on the library's stream).
Scratch tool for bring-up; bench.py is the contract."""
import ctypes as C
import json
import sys
import time

import torch

sys.path.insert(0, ".")
import numcl_b200 as nc  # noqa: E402

PEAK = 6450.3


def timed(fn, iters=20, warm=3, flush=None):
    L = nc.lib()
    s = torch.cuda.Stream()
    L.ncl_set_stream(C.c_void_p(s.cuda_stream))
    with torch.cuda.stream(s):
        for _ in range(warm):
            fn()
        ts = []
        for _ in range(iters):
            if flush is not None:
                flush.add_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            fn()
            e1.record(s)
            e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3)
    L.ncl_set_stream(None)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    torch.cuda.init()
    L = nc.lib()
    flush = torch.empty(256 << 20, dtype=torch.float32, device="cuda")   # 1 GiB > L2
    out = {}
    # C1
    a = nc.random_array((4096, 4096), "f32", 1, 2, -1.0, 1.0)
    b = nc.random_array((4096,), "f32", 2, 2, -1.0, 1.0)
    r = nc.empty((4096, 4096), "f32")
    ta, tb, tr = a.tensor(), b.tensor(), r.tensor()
    args = (nc._lib.Tensor * 2)(ta, tb)
    f = lambda: L.ncl_fold(0, 1, args, 2, C.byref(tr))
    med, best = timed(f, flush=flush)
    out["C1_add_f32"] = dict(us=med * 1e6, best_us=best * 1e6, gbs=134234112 / med / 1e9, frac=134234112 / med / 1e9 / PEAK, kernel=L.ncl_last_kernel().decode())
    # C2
    big = nc.random_array((16384, 16384), "f32", 3, 0, -1.0, 1.0)
    rows = nc.empty((16384,), "f32")
    tot = nc.empty((), "f32")
    tbig, trows, ttot = big.tensor(), rows.tensor(), tot.tensor()
    ax1 = (C.c_int * 1)(1)
    med, best = timed(lambda: L.ncl_reduce(0, C.byref(tbig), ax1, 1, C.byref(trows), 1))
    out["C2a_sum_axis1"] = dict(us=med * 1e6, best_us=best * 1e6, gbs=1073807360 / med / 1e9, frac=1073807360 / med / 1e9 / PEAK, kernel=L.ncl_last_kernel().decode())
    med, best = timed(lambda: L.ncl_reduce(0, C.byref(tbig), ax1, 0, C.byref(ttot), 1))
    out["C2b_sum_all"] = dict(us=med * 1e6, best_us=best * 1e6, gbs=1073741828 / med / 1e9, frac=1073741828 / med / 1e9 / PEAK, kernel=L.ncl_last_kernel().decode())
    ax0 = (C.c_int * 1)(0)
    cols = nc.empty((16384,), "f32")
    tcols = cols.tensor()
    med, best = timed(lambda: L.ncl_reduce(0, C.byref(tbig), ax0, 1, C.byref(tcols), 1))
    out["C2c_sum_axis0"] = dict(us=med * 1e6, best_us=best * 1e6, gbs=1073807360 / med / 1e9, frac=1073807360 / med / 1e9 / PEAK, kernel=L.ncl_last_kernel().decode())
    del big
    # C5a / C5b
    x = nc.random_array((64, 32, 32, 256), "ub8", 4, 0, 0, 255)
    y = nc.random_array((256,), "fixnum", 5, 0, -(2 ** 40), 2 ** 40)
    p = nc.empty((64, 32, 32, 256), "fixnum")
    tx, ty, tp = x.tensor(), y.tensor(), p.tensor()
    args2 = (nc._lib.Tensor * 2)(tx, ty)
    med, best = timed(lambda: L.ncl_fold(2, 1, args2, 2, C.byref(tp)), flush=flush)
    out["C5a_mul_u8_fix"] = dict(us=med * 1e6, best_us=best * 1e6, gbs=150996992 / med / 1e9, frac=150996992 / med / 1e9 / PEAK, kernel=L.ncl_last_kernel().decode())
    t0 = time.time()
    q = nc.einsum("abcd -> dbca", p)
    specs, desc = nc.einsum.__globals__["_memoized"]("abcd -> dbca")
    tq = q.tensor()
    tin = (nc._lib.Tensor * 1)(tp)
    tout = (nc._lib.Tensor * 1)(tq)
    med, best = timed(lambda: L.ncl_einsum(C.byref(desc), tin, 1, tout, 1, 1), flush=flush)
    out["C5b_transpose"] = dict(us=med * 1e6, best_us=best * 1e6, gbs=268435456 / med / 1e9, frac=268435456 / med / 1e9 / PEAK, kernel=L.ncl_last_kernel().decode())
    # contractions
    from numcl_b200.einsum import _memoized
    for name, dt, spec, sa, sb, sc, flops in [("C3_f64", "f64", "ij jk -> ik", (8192, 8192), (8192, 8192), (8192, 8192), 2 * 8192 ** 3),
                                               ("C3_f32", "f32", "ij jk -> ik", (4096, 4096), (4096, 4096), (4096, 4096), 2 * 4096 ** 3),
                                               ("C4_f32", "f32", "bij bjk -> bik", (256, 512, 512), (256, 512, 512), (256, 512, 512), 2 * 256 * 512 ** 3)]:
        _, d = _memoized(spec)
        A = nc.random_array(sa, dt, 7, 0, -1.0, 1.0); B = nc.random_array(sb, dt, 8, 0, -1.0, 1.0); Cc = nc.empty(sc, dt)
        tin = (nc._lib.Tensor * 2)(A.tensor(), B.tensor()); tout = (nc._lib.Tensor * 1)(Cc.tensor())
        med, best = timed(lambda: L.ncl_einsum(C.byref(d), tin, 2, tout, 1, 1), iters=3, warm=1)
        out[name] = dict(ms=med * 1e3, tflops=flops / med / 1e12, kernel=L.ncl_last_kernel().decode())
        del A, B, Cc
    # denominators for the contractions: cuBLAS f64 and tf32 GEMM peaks
    for name, dt, tf32 in [("f64", torch.float64, False), ("tf32", torch.float32, True), ("f32", torch.float32, False)]:
        torch.backends.cuda.matmul.allow_tf32 = tf32
        n = 8192
        A = torch.randn(n, n, device="cuda", dtype=dt)
        B = torch.randn(n, n, device="cuda", dtype=dt)
        for _ in range(2):
            A @ B
        torch.cuda.synchronize()
        best_t = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); A @ B; e1.record(); e1.synchronize()
            best_t = min(best_t, e0.elapsed_time(e1) * 1e-3)
        out["cublas_" + name] = dict(ms=best_t * 1e3, tflops=2 * n ** 3 / best_t / 1e12)
        del A, B
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
