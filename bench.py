#!/usr/bin/env python
"""bench.py -- the hot path's headline benchmark (contract: see the task statement, section (4)).

Workload (BASELINE.json configs[1]): numcl:sum over axis 1 AND over the full array of ONE
16384 x 16384 single-float array.  One "step" = both reductions.  With N GPUs the array is sharded by
rows, 16384/N rows per rank (STRONG scaling: the global shape is fixed); the axis-1 sums need no
exchange, the full sum combines one float64 partial per GPU over NVLink (peer-mapped mailboxes written
by the reduction kernel itself, ncl_reduce_all_sharded_async; ncclAllReduce if peer access is missing).
Steps rotate over 3 arrays (3 seeds) so that a 128 MiB shard is not served from the 126 MB L2.

  metric/value : GB/s of ALGORITHMIC bytes of the whole job, inputs resident in HBM; the
                 timed region replays CUDA graphs of the step's C-ABI calls (ncl_capture_begin/end)
  e2e          : the same through the C ABI with the shard in pinned HOST memory, copies inside
  roofline     : the dominant kernel (reduce_row_block) vs the measured HBM peak, CUDA events
  cpu_baseline : the oracle's typed C loops (numcl's sequential scans), 1 thread, on the host
  others       : the remaining BASELINE configs timed the same way (device resident); at N > 1 the
                 shardable ones (C1 and C5a by rows, C4 along b) on each rank's shard, max over ranks,
                 plus the weak-scaling variant of the headline (one whole 16384^2 array per GPU)

`--impl reference` times the reference's own algorithm on the host CPU (the oracle port: the
reference is Common Lisp and no Lisp exists in this image) on the same array.
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ROWS = N_COLS = 16384
BYTES_ROW = N_ROWS * N_COLS * 4 + N_ROWS * 4          # read a, write the row sums      (1 073 807 360)
BYTES_ALL = N_ROWS * N_COLS * 4 + 4                    # read a, write one element       (1 073 741 828)
SEED = 20261019
N_ARRAYS = 3                                           # rotating arrays (seeds SEED, SEED+1, SEED+2)
BIG_REPS = 4                                           # rotations per replay of the long graph (12 steps)
METRIC = "einsum GB/s (elementwise/reduce) & TFLOP/s (matmul) vs B200 roofline"
WORKLOAD = "numcl:sum axis 1 + numcl:sum full of a 16384x16384 single-float array (BASELINE configs[1])"


def workload_config(world):
    return {"workload": WORKLOAD, "global_shape": [N_ROWS, N_COLS],
            "sharding": f"rows (outer axis), {N_ROWS // max(world, 1)} per GPU; axis-1 sums need no exchange, the full sum combines one float64 partial per GPU",
            "l2": f"steps rotate over {N_ARRAYS} arrays: consecutive steps never read the same bytes (a shard is {N_ROWS * N_COLS * 4 // max(world, 1) >> 20} MiB, L2 is 126 MB)",
            "algorithmic_bytes_per_step": BYTES_ROW + BYTES_ALL}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)", d
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)", {}


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  The timed region of this
    workload lasts milliseconds, far below `nvidia-smi -lms` granularity, so NVML is polled directly
    from a thread (same counters as the recipe's nvidia-smi query)."""

    def __init__(self, dev):
        self.dev, self.sm, self.reasons, self.stop_flag, self.max = dev, [], set(), False, None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(dev)
            self.max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                break
            time.sleep(0.0005)        # yield the GIL: the launch loop of this rank must not be starved

    def start(self):
        if self.nv:
            self.th = threading.Thread(target=self._loop, daemon=True)
            self.th.start()

    def stop(self):
        if not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        self.stop_flag = True
        self.th.join(timeout=2)
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max, "reasons": sorted(self.reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
def cpu_reference_step(O, a_host):
    """numcl's own loops on the host: sequential row-major scans (src/5reduce.lisp:78-87), 1 thread."""
    t0 = time.perf_counter()
    rows = O.fast_sum_axis1_f32(a_host)
    total = O.fast_sum_all_f32(a_host)
    return time.perf_counter() - t0, rows, total


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path, timed on the host cores, on the
    same 16384 x 16384 array (all of it) as the GPU arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    a = O.fill_random((N_ROWS, N_COLS), "f32", SEED, 0, -1.0, 1.0).values()
    for _ in range(min(args.warmup, 1)):
        cpu_reference_step(O, a)
    ts = [cpu_reference_step(O, a)[0] for _ in range(args.steps)]
    t = sum(ts) / len(ts)
    value = (BYTES_ROW + BYTES_ALL) / t / 1e9
    cfg = workload_config(args.gpus)
    cfg["note"] = "reference = C transcription of numcl's reduce-array nest (oracle port), 1 thread because numcl is single-threaded; SBCL is not installed in this image"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": 1, "kind": "port", "host_cores": os.cpu_count(),
                         "sample": f"the whole {N_ROWS}x{N_COLS} array (1 GiB) per step, {args.steps} steps, mean"},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
def time_other_configs(nc, L, torch, stream, peak, rank, world, dist, sampler_dev, contractions=True):
    """C1, C3, C4, C5 of BASELINE.json, device resident, rotating buffer sets so that no operand of
    the next launch is still in L2 (126 MB).  world > 1: only what shards without an exchange runs
     -- C1 and C5a by rows of the outer axis, C4 along b -- every rank on its own shard,
    the time is the max over ranks, the figure is the whole job's."""
    from numcl_b200 import _lib
    out = {}
    sampler = ClockSampler(sampler_dev)
    if rank == 0:
        sampler.start()

    def timed(make_call, nsets, iters, warm=3, batch=None):
        """median over `iters` event pairs of the average duration of a back-to-back batch of
        launches (one per rotating buffer set).  The batch is recorded once into a CUDA graph
        (ncl_capture_begin/end) and replayed, so host-side launch cost (ctypes + planning, ~20 us per
        call) cannot starve the stream: the figure is kernel time."""
        calls = [make_call(i) for i in range(nsets)]
        batch = batch or nsets
        for i in range(max(warm, nsets)):
            calls[i % nsets]()
        stream.synchronize()
        with _lib.Graph() as graph:
            for j in range(batch):
                calls[j % nsets]()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
        graph.launch()
        for i in range(iters):
            ev[i][0].record(stream)
            graph.launch()
            ev[i][1].record(stream)
        stream.synchronize()
        ts = sorted(e0.elapsed_time(e1) * 1e-3 / batch for e0, e1 in ev)
        t = ts[len(ts) // 2]
        if world > 1:
            tt = torch.tensor([t], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t = tt.item()
        return t, L.ncl_last_kernel().decode()

    def fold_call(op, arrs, r):
        ts = (_lib.Tensor * len(arrs))(*[a.tensor() for a in arrs])
        tr = r.tensor()
        keep = (ts, tr, arrs, r)
        return lambda: _lib.check(L.ncl_fold(op, 1, keep[0], len(arrs), C.byref(keep[1])))

    out["timing"] = "CUDA-graph replay (ncl_capture_begin/end) of back-to-back launches over rotating buffer sets; max over ranks"
    shard = f", sharded by rows: {4096 // world} per GPU" if world > 1 else ""
    # C1: (numcl:+ a b), 4096x4096 f32 + 4096 row vector; rank r owns rows [r*4096/world, (r+1)*4096/world)
    r1 = 4096 // world
    b = nc.random_array((4096,), "f32", SEED + 1, 2, -1.0, 1.0)
    nset = 4 if world == 1 else 8
    sets = [(nc.random_array((r1, 4096), "f32", SEED + 10 + i, 2, -1.0, 1.0, first=rank * r1 * 4096), nc.empty((r1, 4096), "f32")) for i in range(nset)]
    t, k = timed(lambda i: fold_call(_lib.OPS["ADD"], [sets[i][0], b], sets[i][1]), nset, 12, batch=2 * nset)
    byt = 2 * 4096 * 4096 * 4 + 4096 * 4 * world
    out["C1 (numcl:+ a b) 4096x4096 f32 + 4096" + shard] = {"us": t * 1e6, "GB/s": byt / t / 1e9, "frac_hbm": byt / t / 1e9 / (peak * world), "kernel": k}
    del sets
    # C5a: (numcl:* u8 fixnum) -> fixnum, (64,32,32,256), outer axis sharded; C5b: (abcd -> dbca) of the result (1 GPU only:
    # the permutation moves the sharded axis)
    a5 = 64 // world
    y = nc.random_array((256,), "fixnum", SEED + 2, 0, -(2 ** 40), 2 ** 40)
    nset = 3 if world == 1 else 8
    sets = [(nc.random_array((a5, 32, 32, 256), "ub8", SEED + 20 + i, 0, 0, 255, first=rank * a5 * 32 * 32 * 256), nc.empty((a5, 32, 32, 256), "fixnum")) for i in range(nset)]
    t, k = timed(lambda i: fold_call(_lib.OPS["MUL"], [sets[i][0], y], sets[i][1]), nset, 12, batch=2 * nset)
    byt = 16777216 + 134217728 + 2048 * world
    out["C5a (numcl:* ub8 fixnum) (64,32,32,256)" + (f", outer axis sharded: {a5} per GPU" if world > 1 else "")] = {"us": t * 1e6, "GB/s": byt / t / 1e9, "frac_hbm": byt / t / 1e9 / (peak * world), "kernel": k}
    from numcl_b200.einsum import _memoized
    if world == 1:
        _, desc = _memoized("abcd -> dbca")
        outs = [nc.empty((256, 32, 32, 64), "fixnum") for _ in range(3)]

        def tr_call(i):
            tin = (_lib.Tensor * 1)(sets[i][1].tensor())
            tout = (_lib.Tensor * 1)(outs[i].tensor())
            keep = (tin, tout)
            return lambda: _lib.check(L.ncl_einsum(C.byref(desc), keep[0], 1, keep[1], 1, _lib.OUT_FRESH))
        t, k = timed(tr_call, 3, 12, batch=6)
        out["C5b (einsum '(abcd -> dbca)) fixnum"] = {"us": t * 1e6, "GB/s": 268435456 / t / 1e9, "frac_hbm": 268435456 / t / 1e9 / peak, "kernel": k}
        del outs
    del sets
    out["clocks_elementwise"] = sampler.stop() if rank == 0 else None

    # C3 / C4: contractions (TFLOP/s; denominators measured with cuBLAS in the same run, clocks sampled around them)
    def cublas_peak(dt, tf32, n=8192):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        A = torch.randn(n, n, device="cuda", dtype=dt)
        B = torch.randn(n, n, device="cuda", dtype=dt)
        for _ in range(2):
            A @ B
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); A @ B; e1.record(); e1.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e-3)
        torch.backends.cuda.matmul.allow_tf32 = False
        return 2 * n ** 3 / best / 1e12
    if not contractions:
        return out
    try:
        _, mm = _memoized("ij jk -> ik")
        _, bmm = _memoized("bij bjk -> bik")
        sampler = ClockSampler(sampler_dev)
        if rank == 0:
            sampler.start()
        peaks = {"tf32": cublas_peak(torch.float32, True)}
        if world == 1:
            peaks["f64"] = cublas_peak(torch.float64, False)
        out["clocks_cublas_denominators"] = sampler.stop() if rank == 0 else None
        sampler = ClockSampler(sampler_dev)
        if rank == 0:
            sampler.start()
        bl = 256 // world
        cases = []
        if world == 1:
            cases += [("C3 matmul 8192^3 f64", "f64", mm, (8192, 8192), (8192, 8192), (8192, 8192), 2 * 8192 ** 3, peaks["f64"], 3),
                      ("C3 matmul 8192^3 f32", "f32", mm, (8192, 8192), (8192, 8192), (8192, 8192), 2 * 8192 ** 3, peaks["tf32"] / 3, 5)]
        cases += [("C4 (bij bjk -> bik) b=256 512^3 f32" + (f", sharded along b: {bl} per GPU" if world > 1 else ""), "f32", bmm,
                   (bl, 512, 512), (bl, 512, 512), (bl, 512, 512), 2 * 256 * 512 ** 3, peaks["tf32"] / 3 * world, 10)]
        for name, dt, spec_desc, shp_a, shp_b, shp_c, flops, denom, iters in cases:
            first = rank * shp_a[0] * shp_a[1] * shp_a[2] if len(shp_a) == 3 else 0
            nset = 1 if world == 1 else 4                  # a b/8 shard (96 MiB) would otherwise be served from L2
            ABC = [(nc.random_array(shp_a, dt, SEED + 30 + 2 * i, 0, -1.0, 1.0, first=first), nc.random_array(shp_b, dt, SEED + 31 + 2 * i, 0, -1.0, 1.0, first=first),
                    nc.empty(shp_c, dt)) for i in range(nset)]

            def mk(i):
                tin = (_lib.Tensor * 2)(ABC[i][0].tensor(), ABC[i][1].tensor())
                tout = (_lib.Tensor * 1)(ABC[i][2].tensor())
                keep = (tin, tout)
                return lambda: _lib.check(L.ncl_einsum(C.byref(spec_desc), keep[0], 2, keep[1], 1, _lib.OUT_FRESH))
            t, k = timed(mk, nset, iters, warm=1)
            out[name] = {"ms": t * 1e3, "TFLOP/s": flops / t / 1e12, "denominator_TFLOP/s": denom, "frac_tensor": flops / t / 1e12 / denom, "kernel": k,
                         "denominator": ("cuBLAS f64 GEMM measured in this run" if dt == "f64" else "cuBLAS TF32 GEMM measured in this run / 3") + (f" x {world} GPUs" if world > 1 else "")}
            del ABC
        out["clocks_contractions"] = sampler.stop() if rank == 0 else None
    except Exception as e:                                   # never lose the headline line to an auxiliary config
        out["contractions"] = {"error": repr(e)[:300]}
    return out


class _CudaView:
    """A library buffer seen through __cuda_array_interface__, so that torch can read it (the in-bench fp64 check)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}


def time_e2e_other_configs(nc, L, contractions=True):
    """The other BASELINE configs END TO END on pinned HOST arrays through the operator interface (numcl:+, numcl:*,
    einsum, matmul): copy in, compute, copy out inside the timed call, which returns when the result is in host memory.
    The library cuts such a call into panels of its outermost kept loop and overlaps upload, kernels and write-back
    (include/numcl_cuda.h: ncl_pipeline_stats); result arrays come from the pinned-block cache."""
    from numcl_b200 import _lib

    def H(shape, dt, seed, lo=-1.0, hi=1.0):
        d = nc.random_array(shape, dt, seed, 0, lo, hi)
        h = nc.NArray(shape, dt, where="host")
        h.set_raw(d.raw())
        return h

    def timed(f, n=5):
        f(); f(); L.ncl_sync()
        st0 = _lib.transfer_stats() + _lib.pipeline_stats()
        ts = []
        for _ in range(n):
            t0 = time.perf_counter(); f(); L.ncl_sync(); ts.append(time.perf_counter() - t0)
        st1 = _lib.transfer_stats() + _lib.pipeline_stats()
        return sorted(ts)[len(ts) // 2], [(b - a) / n for a, b in zip(st0, st1)]

    out = {"how": "median of 5 calls on pinned host arrays, host clock around the call; the call returns with the result in host memory"}
    cases = [("C1 (numcl:+ a b) 4096x4096 f32 + 4096", 134234112, None, lambda: (H((4096, 4096), "f32", 1), H((4096,), "f32", 2)), lambda a, b: nc.add(a, b)),
             ("C5a (numcl:* ub8 fixnum) (64,32,32,256)", 150996992, None, lambda: (H((64, 32, 32, 256), "ub8", 4, 0, 255), H((256,), "fixnum", 5, -2 ** 40, 2 ** 40)), lambda x, y: nc.mul(x, y))]
    if contractions:
        cases += [("C4 (bij bjk -> bik) b=256 512^3 f32", 805306368, 2 * 256 * 512 ** 3, lambda: (H((256, 512, 512), "f32", 6), H((256, 512, 512), "f32", 7)), lambda A, B: nc.einsum("bij bjk -> bik", A, B)),
                  ("C3 matmul 8192^3 f32", 805306368, 2 * 8192 ** 3, lambda: (H((8192, 8192), "f32", 8), H((8192, 8192), "f32", 9)), lambda A, B: nc.matmul(A, B)),
                  ("C3 matmul 8192^3 f64", 1610612736, 2 * 8192 ** 3, lambda: (H((8192, 8192), "f64", 8), H((8192, 8192), "f64", 9)), lambda A, B: nc.matmul(A, B))]
    for name, nbytes, flop, make, call in cases:
        ops = make()
        t, (h2d, d2h, pc, pp) = timed(lambda: call(*ops))
        r = {"ms": t * 1e3, "GB/s": nbytes / t / 1e9, "h2d_bytes": int(h2d), "d2h_bytes": int(d2h), "panels": pp if pc else 0,
             "pcie_one_way_ms_at_55GBs": max(h2d, d2h) / 55e6, "kernel": L.ncl_last_kernel().decode()}
        if flop:
            r["TFLOP/s"] = flop / t / 1e12
        out[name] = r
        del ops
    L.ncl_trim()
    return out


def source_stamp():
    """hash of the reduction kernels' sources: roofline.traffic is an ncu figure taken from a build with this stamp"""
    h = hashlib.sha256()
    for f in ("k_reduce.cuh", "k_reduce.cu", "ncl_device.cuh"):
        h.update(open(os.path.join(ROOT, "numcl_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ["NUMCL_CUDA_DEVICE"] = str(local)
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import numcl_b200 as nc
    from numcl_b200 import _lib
    L = nc.lib()
    if world > 1:                                            # share one NCCL unique id through torch.distributed
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_uint8 * 128)()
            _lib.check(L.ncl_comm_unique_id(buf))
            uid = torch.tensor(list(buf), dtype=torch.uint8)
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        buf = (C.c_uint8 * 128)(*uid.cpu().tolist())
        _lib.check(L.ncl_comm_init(buf, rank, world))
    peak, peak_src, _ = measured_peaks()

    stream = torch.cuda.Stream()
    _lib.check(L.ncl_set_stream(C.c_void_p(stream.cuda_stream)))
    ax1 = (C.c_int * 1)(1)
    from numcl_b200.dist import row_partition

    class Job:
        """this rank's row shards of `narr` synthetic (rows x 16384) arrays + the step over them"""

        def __init__(self, rows_global, narr):
            self.start, self.rows = row_partition(rows_global, world)[rank]
            self.narr = narr
            self.a = [nc.random_array((self.rows, N_COLS), "f32", SEED + s, 0, -1.0, 1.0, first=self.start * N_COLS) for s in range(narr)]
            self.rowsum = [nc.empty((self.rows,), "f32") for _ in range(narr)]
            self.total = [nc.empty((), "f32") for _ in range(narr)]
            self.ta = [x.tensor() for x in self.a]
            self.tr = [x.tensor() for x in self.rowsum]
            self.tt = [x.tensor() for x in self.total]
            self.graph = self.big = None

        def step(self, i, ev=None, join=True):
            # N > 1: the full sum goes first: its kernel pushes the partial into every rank's mailbox and the
            # one-CTA collect runs on the communication stream BESIDE the axis-1 kernel, which needs no
            # exchange.  ncl_comm_join orders the compute stream after the collects still in flight; inside the
            # timed region it is called once per graph / at the end of run(), not per step: the library keeps two
            # exchanges in flight (a push waits only for the collect TWO steps back), so the next step's kernels
            # start while this step's collect still waits for the slowest rank.  Every result is complete when the
            # timed region closes (join + barrier + synchronize).
            s = i % self.narr
            if ev:
                ev[1].record(stream)
            if world > 1:
                _lib.check(L.ncl_reduce_all_sharded_async(_lib.OPS["ADD"], C.byref(self.ta[s]), C.byref(self.tt[s])))     # (sum a)
            else:
                _lib.check(L.ncl_reduce(_lib.OPS["ADD"], C.byref(self.ta[s]), ax1, 0, C.byref(self.tt[s]), _lib.OUT_FRESH))  # (sum a)
            if ev:
                ev[2].record(stream)
                ev[0].record(stream)
            _lib.check(L.ncl_reduce(_lib.OPS["ADD"], C.byref(self.ta[s]), ax1, 1, C.byref(self.tr[s]), _lib.OUT_FRESH))      # (sum a :axes 1)
            if ev:
                ev[3].record(stream)
            if world > 1 and join:
                _lib.check(L.ncl_comm_join())

        def capture(self):
            """two graphs: BIG_REPS rotations of the arrays (12 steps) and one rotation (3 steps).  A graph is one launch and
            ends with the join of the communication stream, so longer graphs amortise both; the steps inside are the same
            C-ABI calls in the same order."""
            graphs = []
            for reps in (BIG_REPS, 1):
                with _lib.Graph() as g:                  # ncl_capture_end joins the communication stream
                    for s in range(self.narr * reps):
                        self.step(s, join=False)
                graphs.append(g)
            self.big, self.graph = graphs

        def run(self, steps):
            """exactly `steps` steps: replays of the 12-step graph, then of the 3-step graph, the remainder eagerly"""
            big = self.narr * BIG_REPS
            for _ in range(steps // big):
                self.big.launch()
            steps %= big
            for _ in range(steps // self.narr):
                self.graph.launch()
            for i in range(steps % self.narr):
                self.step(i, join=False)
            if world > 1:
                _lib.check(L.ncl_comm_join())

    def barrier():
        stream.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sync_a = nc.zeros((8,), type="f32")
    sync_r = nc.empty((), type="f32")
    tsync_a, tsync_r = sync_a.tensor(), sync_r.tensor()

    def timed_steps(job, steps):
        """device time of `steps` steps, max over ranks"""
        t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            # device-side alignment: ranks leave the host barrier up to milliseconds apart; a one-element
            # exchange enqueued BEFORE the start event makes every GPU begin its timed region together,
            # so the max-over-ranks time does not charge the first rank for the last rank's late start
            _lib.check(L.ncl_reduce_all_sharded(_lib.OPS["ADD"], C.byref(tsync_a), C.byref(tsync_r)))
        t_start.record(stream)
        h0 = time.perf_counter()
        job.run(steps)
        host = time.perf_counter() - h0
        t_end.record(stream)
        barrier()
        el = t_start.elapsed_time(t_end) * 1e-3
        if world > 1:
            tt = torch.tensor([el], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            el = tt.item()
        return el, host

    job = Job(args.rows, N_ARRAYS)
    for i in range(max(args.warmup, N_ARRAYS)):
        job.step(i)
    barrier()
    job.capture()
    job.run(N_ARRAYS * (BIG_REPS + 1))                      # one replay of each graph as warm-up of the graphs themselves
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = L.ncl_launch_count()
    elapsed, host_enqueue = timed_steps(job, args.steps)
    launches = L.ncl_launch_count() - l0
    # per-kernel durations: the same steps issued one call at a time with CUDA events around each kernel
    ev_n = max(6, min(30, args.steps))
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(ev_n)]
    for i in range(ev_n):
        job.step(i, evs[i])
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t_row = sorted(e[0].elapsed_time(e[3]) for e in evs)[ev_n // 2] * 1e-3
    t_all = sorted(e[1].elapsed_time(e[2]) for e in evs)[ev_n // 2] * 1e-3
    per_rank = None
    parity = None
    if world > 1:
        mine = torch.tensor([t_row, t_all, host_enqueue / args.steps], device="cuda", dtype=torch.float64)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"row_us": [round(x[0].item() * 1e6, 1) for x in allr], "full_local_us": [round(x[1].item() * 1e6, 1) for x in allr],
                    "host_enqueue_us_per_step": [round(x[2].item() * 1e6, 1) for x in allr]}
        tt = torch.tensor([t_row, t_all], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_row, t_all = tt.tolist()
        # correctness of the NVLink exchange, checked in the run: every rank's result of (sum a) against the float64
        # sum of float64 per-rank partials computed by torch from the same device buffers
        errs, same = [], True
        for s in range(N_ARRAYS):
            view = torch.as_tensor(_CudaView(job.a[s].device_ptr(), job.rows * N_COLS), device="cuda")
            part = view.double().sum().reshape(1)
            dist.all_reduce(part)
            got = torch.tensor([float(job.total[s].numpy().reshape(-1)[0])], device="cuda", dtype=torch.float64)
            gl = [torch.zeros_like(got) for _ in range(world)]
            dist.all_gather(gl, got)
            same = same and all(g.item() == gl[0].item() for g in gl)
            errs.append(max(abs(g.item() - part.item()) for g in gl) / abs(part.item()))        # every rank's copy of the result
        parity = {"parity_full_sum_rel_err": max(errs), "same_bits_on_every_rank": same, "parity_ok": bool(max(errs) <= 1e-6 and same),
                  "how": "float32 result of the sharded (sum a) vs torch float64 sum of the same shards, all-reduced; per array, max"}
    scale_rows = args.rows / N_ROWS
    value = (BYTES_ROW + BYTES_ALL) * scale_rows * args.steps / elapsed / 1e9

    # ---- e2e: this rank's shard in pinned host memory, copies inside the timed region, through the C ABI ----------
    host = nc.NArray((job.rows, N_COLS), "f32", where="host")
    host.set_raw(job.a[0].raw())
    rows_h = np.empty(job.rows, dtype=np.float32)
    tot_h = np.empty(1, dtype=np.float32)
    e2e_steps = max(1, min(args.steps, 5))
    shard_bytes = job.rows * N_COLS * 4

    def e2e_step():
        _lib.check(L.ncl_h2d(job.a[0]._buf, 0, host._host.ctypes.data, shard_bytes))
        job.step(0)
        _lib.check(L.ncl_d2h(rows_h.ctypes.data, job.rowsum[0]._buf, 0, job.rows * 4))
        _lib.check(L.ncl_d2h(tot_h.ctypes.data, job.total[0]._buf, 0, 4))
    e2e_step()
    barrier()
    w0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_t = time.perf_counter() - w0
    if world > 1:
        tt = torch.tensor([e2e_t], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_t = tt.item()
    e2e_value = (BYTES_ROW + BYTES_ALL) * e2e_steps / e2e_t / 1e9

    # ---- e2e, warm: the same HOST arrays carry a resident device mirror (ncl_host_register: what the Lisp glue attaches in
    # %make-array); the calls take the host tensors themselves, the first one uploads, the timed ones move no input bytes ----
    host.make_resident()
    rows_host = nc.NArray((job.rows,), "f32", where="host").make_resident(fresh=True)       # written back into the host vector every call
    th, trh = host.tensor(), rows_host.tensor()

    def warm_step():
        if world > 1:
            _lib.check(L.ncl_reduce_all_sharded_async(_lib.OPS["ADD"], C.byref(th), C.byref(job.tt[0])))
        else:
            _lib.check(L.ncl_reduce(_lib.OPS["ADD"], C.byref(th), ax1, 0, C.byref(job.tt[0]), _lib.OUT_FRESH))
        _lib.check(L.ncl_reduce(_lib.OPS["ADD"], C.byref(th), ax1, 1, C.byref(trh), _lib.OUT_FRESH))
        if world > 1:
            _lib.check(L.ncl_comm_join())
        _lib.check(L.ncl_d2h(tot_h.ctypes.data, job.total[0]._buf, 0, 4))
    warm_step()                                              # cold call: uploads the shard once
    barrier()
    hb0, db0 = _lib.transfer_stats()
    warm_steps = max(3, min(args.steps, 20))
    w0 = time.perf_counter()
    for _ in range(warm_steps):
        warm_step()
    barrier()
    warm_t = time.perf_counter() - w0
    hb1, db1 = _lib.transfer_stats()
    if world > 1:
        tt = torch.tensor([warm_t], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        warm_t = tt.item()
    e2e_warm = {"value": (BYTES_ROW + BYTES_ALL) * warm_steps / warm_t / 1e9, "unit": "GB/s", "steps": warm_steps,
                "h2d_bytes_per_step": (hb1 - hb0) // warm_steps * world, "d2h_bytes_per_step": (db1 - db0) // warm_steps * world,
                "api": "ncl_reduce x2 on HOST tensors whose base vectors are registered (ncl_host_register); results written back to host memory every step"}
    del host, rows_host

    # ---- the other configs (every rank takes part: sharded), the weak-scaling variant, the CPU baseline ------------
    others = {}
    if not args.skip_others:
        others = time_other_configs(nc, L, torch, stream, peak, rank, world, dist, local, contractions=not args.no_contractions)
        if world == 1:
            others["e2e_host_arrays"] = time_e2e_other_configs(nc, L, contractions=not args.no_contractions)
    if world > 1 and not args.skip_others:
        del job
        wjob = Job(N_ROWS * world, N_ARRAYS)               # weak scaling: one whole 16384^2 array per GPU
        for i in range(N_ARRAYS):
            wjob.step(i)
        barrier()
        wjob.capture()
        wjob.run(N_ARRAYS * (BIG_REPS + 1))
        barrier()
        wsteps = max(N_ARRAYS, min(args.steps, 30))
        wel, _ = timed_steps(wjob, wsteps)
        others["weak scaling: one 16384x16384 array per GPU"] = {"ms_per_step": wel / wsteps * 1e3, "GB/s": (BYTES_ROW + BYTES_ALL) * world * wsteps / wel / 1e9,
                                                                  "global_shape": [N_ROWS * world, N_COLS]}
        del wjob
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1:
        from oracle import oracle as O
        ah = O.fill_random((N_ROWS, N_COLS), "f32", SEED, 0, -1.0, 1.0).values()
        reps = 12
        ct = min(cpu_reference_step(O, ah)[0] for _ in range(reps))
        cpu = {"value": (BYTES_ROW + BYTES_ALL) / ct / 1e9, "unit": "GB/s", "cores": 1, "kind": "port",
               "sample": f"the whole {N_ROWS}x{N_COLS} array (1 GiB), best of {reps}; C transcription of numcl's sequential reduce nest, not SBCL",
               "host_cores": os.cpu_count()}
        # parity spot check of the timed arrays against the oracle on the same synthetic stream
        got = job.rowsum[0].numpy().astype(np.float64)
        want = O.fast_sum_axis1_f32(ah).astype(np.float64)
        cpu["parity_rel_err_axis1_vs_port"] = float(np.linalg.norm(got - want) / np.linalg.norm(want))
        truth = float(ah.astype(np.float64).sum())
        cpu["parity_full_sum_rel_err_vs_fp64"] = abs(float(job.total[0].numpy().reshape(-1)[0]) - truth) / abs(truth)
        cpu["port_full_sum_rel_err_vs_fp64"] = abs(float(O.fast_sum_all_f32(ah)) - truth) / abs(truth)
    traffic, traffic_note = None, None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        traffic = tj.get("reduce_row_block")
        traffic_note = {"source": tj.get("source"), "taken_at_source_stamp": tj.get("source_stamp"), "current_source_stamp": source_stamp(),
                        "stale": tj.get("source_stamp") != source_stamp()}
    bytes_row_local = BYTES_ROW / world
    achieved = bytes_row_local / t_row / 1e9
    cfg = workload_config(world)
    line = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "roofline": {"bound": "hbm", "kernel": "reduce_row_kernel (sum :axes 1)" + (f" on a {N_ROWS // world}-row shard" if world > 1 else ""),
                     "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic if world == 1 else None, "traffic_note": traffic_note, "peak_source": peak_src,
                     "how": f"median of {ev_n} CUDA-event pairs around the kernel in eager steps issued right after the timed region",
                     "second_kernel": {"kernel": "reduce_all_kernel (sum full)" + (" local phase + push into the peers' mailboxes; the collect runs beside the axis-1 kernel" if world > 1 else ""),
                                       "achieved": BYTES_ALL / world / t_all / 1e9, "frac": BYTES_ALL / world / t_all / 1e9 / peak}},
        "e2e": {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": shard_bytes * world, "d2h_bytes_per_step": (job.rows * 4 + 4) * world if world == 1 else (N_ROWS * 4 + 4 * world),
                "steps": e2e_steps, "api": "ncl_h2d + ncl_reduce x2 + ncl_d2h (pinned host shard per rank)"},
        "e2e_warm": e2e_warm,
        "gpu_launches": int(launches),
        "timed_region": f"{args.steps // (N_ARRAYS * BIG_REPS)} replays of a CUDA graph of {N_ARRAYS * BIG_REPS} steps + {args.steps % (N_ARRAYS * BIG_REPS) // N_ARRAYS} replays of a graph of {N_ARRAYS} steps + {args.steps % N_ARRAYS} eager steps; at N > 1 the communication stream is joined once per graph",
        "host_enqueue_us_per_step": host_enqueue / args.steps * 1e6,
        "clocks": clocks,
    }
    if per_rank:
        line["per_rank"] = per_rank
    if parity:
        line.update(parity)
    if cpu:
        line["cpu_baseline"] = cpu
    if others:
        line["others"] = others
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--skip-others", action="store_true")
    ap.add_argument("--no-contractions", action="store_true", help="skip C3/C4 in `others` (bring-up)")
    ap.add_argument("--rows", type=int, default=N_ROWS, help="rows of the global array (experiments only: the contract is 16384)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
