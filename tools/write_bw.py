"""Write-only / read-only HBM bandwidth ceilings next to the C5a kernel (scratch measurement tool)."""
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
import numcl_b200 as nc
from tools.quick_perf import timed

torch.cuda.init()
L = nc.lib()
for mib in (128, 1024):
    n = mib << 17                                        # int64 elements
    bufs = [nc.empty((n,), "sb64") for _ in range(3 if mib == 128 else 1)]
    ts = [b.tensor() for b in bufs]
    k = [0]
    def fill():
        k[0] += 1
        L.ncl_fill(C.byref(ts[k[0] % len(ts)]), 0, 7, 0.0)
    med, best = timed(fill)
    print(f"ncl_fill {mib} MiB: {mib * 1.048576e6 / med / 1e9:.0f} GB/s median, {mib * 1.048576e6 / best / 1e9:.0f} best")
    t = [torch.empty(n, dtype=torch.int64, device="cuda") for _ in range(3 if mib == 128 else 1)]
    def z():
        k[0] += 1
        t[k[0] % len(t)].zero_()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): z()
        tt = []
        for _ in range(20):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s); z(); e1.record(s); e1.synchronize(); tt.append(e0.elapsed_time(e1) * 1e-3)
    tt.sort()
    print(f"torch zero_ {mib} MiB: {mib * 1.048576e6 / tt[10] / 1e9:.0f} GB/s median, {mib * 1.048576e6 / tt[0] / 1e9:.0f} best")
    def ssum():
        k[0] += 1
        t[k[0] % len(t)].sum()
    with torch.cuda.stream(s):
        for _ in range(3): ssum()
        tt = []
        for _ in range(20):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s); ssum(); e1.record(s); e1.synchronize(); tt.append(e0.elapsed_time(e1) * 1e-3)
    tt.sort()
    print(f"torch sum (read) {mib} MiB: {mib * 1.048576e6 / tt[10] / 1e9:.0f} GB/s median")
    del bufs, t
