"""Strong-scaling probe on ONE GPU: how long do the two reductions of BASELINE configs[1] take on the
row shard a rank owns at N = 1, 2, 4, 8 (16384/N rows of 16384 single-floats)?  The step is captured
into a CUDA graph over rotating slabs (a 128 MiB shard would otherwise sit in the 126 MB L2).
Prints one JSON object; scratch tool, bench.py is the contract."""
import ctypes as C
import json
import sys

import torch

sys.path.insert(0, ".")
import numcl_b200 as nc  # noqa: E402
from numcl_b200 import _lib  # noqa: E402

COLS = 16384


def main():
    torch.cuda.init()
    L = nc.lib()
    stream = torch.cuda.Stream()
    _lib.check(L.ncl_set_stream(C.c_void_p(stream.cuda_stream)))
    ax1 = (C.c_int * 1)(1)
    out = {}
    for n in (1, 2, 4, 8):
        rows = 16384 // n
        nslab = max(3, min(8, (768 << 20) // (rows * COLS * 4) + 1))
        slabs = [nc.random_array((rows, COLS), "f32", 7 + i, 0, -1.0, 1.0) for i in range(nslab)]
        r = nc.empty((rows,), "f32")
        t = nc.empty((), "f32")
        ts = [s.tensor() for s in slabs]
        tr, tt = r.tensor(), t.tensor()

        def k_all(i):
            _lib.check(L.ncl_reduce(0, C.byref(ts[i]), ax1, 0, C.byref(tt), 1))

        def k_row(i):
            _lib.check(L.ncl_reduce(0, C.byref(ts[i]), ax1, 1, C.byref(tr), 1))

        def both(i):
            k_all(i)
            k_row(i)

        res = {}
        for name, fn in (("all", k_all), ("row", k_row), ("step", both)):
            for i in range(nslab):
                fn(i)
            stream.synchronize()
            batch = nslab * 4
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream, capture_error_mode="thread_local"):
                for j in range(batch):
                    fn(j % nslab)
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(9)]
            with torch.cuda.stream(stream):
                g.replay()
                for e0, e1 in evs:
                    e0.record(stream)
                    g.replay()
                    e1.record(stream)
            stream.synchronize()
            us = sorted(e0.elapsed_time(e1) * 1e3 / batch for e0, e1 in evs)
            res[name + "_us"] = round(us[len(us) // 2], 2)
        byt = rows * COLS * 4
        res["ideal_us_at_6800"] = round(byt / 6.8e3, 2) * 1e-3 * 1e3 / 1e3
        res["step_GBs"] = round(2 * byt / res["step_us"] / 1e3, 1)
        res["slabs"] = nslab
        out[f"N={n} rows={rows}"] = res
        del slabs
    base = out["N=1 rows=16384"]["step_us"]
    for n in (2, 4, 8):
        k = f"N={n} rows={16384 // n}"
        out[k]["predicted_efficiency_no_exchange"] = round(base / n / out[k]["step_us"], 3)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
