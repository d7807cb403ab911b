"""Throughput of (N, C) <-> (C, N) transposes over the short side C (scratch tool, GPU box): ~768 MiB moved per call."""
import sys, time
sys.path.insert(0, ".")
import numcl_b200 as nc

PEAK = 6450.3
L = nc.lib()
for dt, wb in (("f32", 4), ("f64", 8)):
    for C in (9, 12, 16, 24, 33, 48, 64, 100, 128, 200):
        n = (384 << 20) // (C * wb)
        for shape in ((n, C), (C, n)):
            a = nc.random_array(shape, dt, 1, 0, -1.0, 1.0)
            f = lambda: nc.transpose(a)
            f(); L.ncl_sync()
            best = 1e9
            for _ in range(3):
                L.ncl_sync(); t0 = time.perf_counter()
                for _ in range(5):
                    f()
                L.ncl_sync(); best = min(best, (time.perf_counter() - t0) / 5)
            print(f"{dt} {str(shape):22s} {best*1e6:8.1f} us {2*a.nbytes/best/1e9:7.0f} GB/s {2*a.nbytes/best/1e9/PEAK:5.2f}  {L.ncl_last_kernel().decode()}", flush=True)
            del a
