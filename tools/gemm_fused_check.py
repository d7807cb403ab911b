"""3xTF32 GEMM: fused (in-kernel split) vs packed (pre-pass) -- accuracy vs float64 and device time.
Scratch tool for bring-up; bench.py is the contract.   python tools/gemm_fused_check.py [quick]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import numcl_b200 as nc  # noqa: E402

MODES = os.environ.get("GEMM_MODES", "packed,fused").split(",")


def run(mode, a, b, spec):
    os.environ["NUMCL_TF32_MODE"] = mode
    r = nc.einsum(spec, a, b)
    return r, nc.lib().ncl_last_kernel().decode()


def accuracy():
    rng = np.random.default_rng(0)
    ok = True
    for (bt, M, N, K) in [(1, 128, 256, 32), (1, 128, 256, 64), (1, 128, 256, 256), (1, 128, 256, 512), (1, 256, 512, 512),
                          (3, 512, 512, 512), (1, 384, 768, 1024), (2, 128, 512, 96), (1, 300, 700, 2500), (2, 640, 256, 4096)]:
        A = rng.uniform(-1, 1, (bt, M, K)).astype(np.float32)
        B = rng.uniform(-1, 1, (bt, K, N)).astype(np.float32)
        ref = A.astype(np.float64) @ B.astype(np.float64)
        a, b = nc.asarray(A, type="f32"), nc.asarray(B, type="f32")
        for mode in MODES:
            r, k = run(mode, a, b, "bij bjk -> bik")
            got = r.numpy().astype(np.float64)
            rel = np.linalg.norm(got - ref) / np.linalg.norm(ref)
            bad = not (rel < 1e-5)
            ok &= not bad
            print(f"b={bt} M={M} N={N} K={K} mode={mode:6s} kernel={k:28s} rel_frob={rel:.3e} {'BAD' if bad else ''}", flush=True)
            if bad and mode != "packed":
                e = np.abs(got - ref)[0]
                print("   max err", e.max(), "at", np.unravel_index(e.argmax(), e.shape), " err by 32-col block:",
                      np.array2string(e.reshape(M, N // 32, 32).max(axis=(0, 2)), precision=2),
                      " by 8-row block:", np.array2string(e.reshape(M // 8, 8, N).max(axis=(1, 2))[:8], precision=2))
    # accumulate into a supplied output
    A = rng.uniform(-1, 1, (256, 512)).astype(np.float32); B = rng.uniform(-1, 1, (512, 256)).astype(np.float32)
    C0 = rng.uniform(-1, 1, (256, 256)).astype(np.float32)
    os.environ["NUMCL_TF32_MODE"] = MODES[-1]
    c = nc.asarray(C0, type="f32")
    nc.einsum("ij jk -> ik", nc.asarray(A, type="f32"), nc.asarray(B, type="f32"), c)
    ref = C0.astype(np.float64) + A.astype(np.float64) @ B.astype(np.float64)
    rel = np.linalg.norm(c.numpy() - ref) / np.linalg.norm(ref)
    print("accumulate:", nc.lib().ncl_last_kernel().decode(), f"rel={rel:.3e}")
    return ok and rel < 1e-5


def timed(fn, iters=20, warm=3):
    L = nc.lib()
    s = torch.cuda.Stream()
    L.ncl_set_stream(C.c_void_p(s.cuda_stream))
    ts = []
    with torch.cuda.stream(s):
        for _ in range(warm):
            fn()
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s); fn(); e1.record(s); e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3)
    L.ncl_set_stream(None)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def timing():
    L = nc.lib()
    shapes = [(256, 512, 512, 512), (64, 1024, 1024, 1024), (1024, 256, 256, 256), (4, 2048, 2048, 1024), (1, 4096, 4096, 1024)]
    if "big" in sys.argv:
        shapes = [(1, 4096, 4096, 4096), (1, 8192, 8192, 8192), (1, 4096, 4096, 1024), (3, 1000, 1500, 700)]
    for (bt, M, N, K) in shapes:
        a = nc.random_array((bt, M, K), "f32", 1, 0, -1.0, 1.0)
        b = nc.random_array((bt, K, N), "f32", 2, 0, -1.0, 1.0)
        r = nc.empty((bt, M, N), "f32")
        from numcl_b200.einsum import _memoized
        specs, desc = _memoized("bij bjk -> bik")
        tin = (nc._lib.Tensor * 2)(a.tensor(), b.tensor()); tout = (nc._lib.Tensor * 1)(r.tensor())
        for mode in MODES:
            os.environ["NUMCL_TF32_MODE"] = mode
            med, best = timed(lambda: L.ncl_einsum(C.byref(desc), tin, 2, tout, 1, 1))
            fl = 2.0 * bt * M * N * K
            print(f"b={bt} M={M} N={N} K={K} mode={mode:6s} {L.ncl_last_kernel().decode():28s} median {med*1e3:.3f} ms best {best*1e3:.3f} ms"
                  f" = {fl/med/1e12:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    torch.cuda.init()
    good = accuracy()
    print("ACCURACY", "OK" if good else "FAILED", flush=True)
    if good or "force" in sys.argv:
        timing()
