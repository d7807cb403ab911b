#!/usr/bin/env python
"""bench.py -- the hot path's headline benchmark (contract: see the task statement, section (4)).

Workload (BASELINE.json configs[1]): numcl:sum over axis 1 AND over the full array of a
16384 x 16384 single-float array.  One "step" = both reductions over one per-GPU slab.  With
N GPUs the array is the row-concatenation of N such slabs (weak scaling: (16384*N) x 16384), the
axis-1 sums need no exchange and the full sum combines one float64 partial per GPU with an NCCL
all-reduce (ncl_reduce_all_sharded).

  metric/value : GB/s of ALGORITHMIC bytes over all ranks, inputs resident in HBM
  e2e          : the same through the C ABI with the slab in pinned HOST memory, copies inside
  roofline     : the dominant kernel (reduce_row_block) vs the measured HBM peak
  cpu_baseline : the oracle's typed C loops (numcl's sequential scans), 1 thread, on the host
  others       : the remaining BASELINE configs timed the same way (device resident)

`--impl reference` times the reference's own algorithm on the host CPU (the oracle port: the
reference is Common Lisp and no Lisp exists in this image).
"""
import argparse
import ctypes as C
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
METRIC = "einsum GB/s (elementwise/reduce) & TFLOP/s (matmul) vs B200 roofline"


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
    """--impl reference: the reference's CPU implementation of the path, timed on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from oracle import oracle as O
    sample_rows = 4096                                   # bounded sample: a quarter of the slab per step
    a = O.fill_random((sample_rows, N_COLS), "f32", SEED, 0, -1.0, 1.0).values()
    for _ in range(min(args.warmup, 1)):
        cpu_reference_step(O, a)
    ts = [cpu_reference_step(O, a)[0] for _ in range(args.steps)]
    frac = sample_rows / N_ROWS
    t = sum(ts) / len(ts)
    value = (BYTES_ROW + BYTES_ALL) * frac / t / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t / frac * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "numcl:sum axis 1 + full, 16384x16384 single-float (BASELINE configs[1])",
                   "note": "reference = C transcription of numcl's reduce-array nest (oracle port); SBCL is not installed in this image"},
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": 1, "kind": "port",
                         "sample": f"{sample_rows} of {N_ROWS} rows per step, {args.steps} steps, ms_per_step scaled to the full slab"},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
def time_other_configs(nc, L, torch, stream, peak, contractions=True):
    """C1, C3, C4, C5 of BASELINE.json, device resident, rotating buffer sets so that no operand of
    the next launch is still in L2 (126 MB)."""
    from numcl_b200 import _lib
    out = {}

    def timed(make_call, nsets, iters, warm=3, batch=None):
        """median over `iters` event pairs of the average duration of a back-to-back batch of
        launches (one per rotating buffer set).  The batch is captured once into a CUDA graph and
        replayed, so host-side launch cost (ctypes + planning, ~20 us per call) cannot starve the
        stream: the figure is kernel time.  Falls back to direct launches if capture is refused."""
        calls = [make_call(i) for i in range(nsets)]
        batch = batch or nsets
        for i in range(max(warm, nsets)):
            calls[i % nsets]()
        stream.synchronize()
        graph = None
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream, capture_error_mode="thread_local"):
                for j in range(batch):
                    calls[j % nsets]()
        except Exception:
            graph = None
            torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
        with torch.cuda.stream(stream):
            if graph is not None:
                graph.replay()
            for i in range(iters):
                ev[i][0].record(stream)
                if graph is not None:
                    graph.replay()
                else:
                    for j in range(batch):
                        calls[j % nsets]()
                ev[i][1].record(stream)
        stream.synchronize()
        ts = sorted(e0.elapsed_time(e1) * 1e-3 / batch for e0, e1 in ev)
        timed.graph_used = graph is not None
        return ts[len(ts) // 2], L.ncl_last_kernel().decode()

    def fold_call(op, arrs, r):
        ts = (_lib.Tensor * len(arrs))(*[a.tensor() for a in arrs])
        tr = r.tensor()
        keep = (ts, tr, arrs, r)
        return lambda: (L.ncl_fold(op, 1, keep[0], len(arrs), C.byref(keep[1])))

    # C1: (numcl:+ a b), 4096x4096 f32 + 4096 row vector
    b = nc.random_array((4096,), "f32", SEED + 1, 2, -1.0, 1.0)
    sets = [(nc.random_array((4096, 4096), "f32", SEED + 10 + i, 2, -1.0, 1.0), nc.empty((4096, 4096), "f32")) for i in range(4)]
    t, k = timed(lambda i: fold_call(_lib.OPS["ADD"], [sets[i][0], b], sets[i][1]), 4, 12, batch=8)
    out["timing"] = "CUDA-graph replay of back-to-back launches over rotating buffer sets" if getattr(timed, "graph_used", False) else "direct back-to-back launches over rotating buffer sets"
    out["C1 (numcl:+ a b) 4096x4096 f32 + 4096"] = {"us": t * 1e6, "GB/s": 134234112 / t / 1e9, "frac_hbm": 134234112 / t / 1e9 / peak, "kernel": k}
    del sets
    # C5a: (numcl:* u8 fixnum) -> fixnum, (64,32,32,256); C5b: (abcd -> dbca) of the result
    y = nc.random_array((256,), "fixnum", SEED + 2, 0, -(2 ** 40), 2 ** 40)
    sets = [(nc.random_array((64, 32, 32, 256), "ub8", SEED + 20 + i, 0, 0, 255), nc.empty((64, 32, 32, 256), "fixnum")) for i in range(3)]
    t, k = timed(lambda i: fold_call(_lib.OPS["MUL"], [sets[i][0], y], sets[i][1]), 3, 12, batch=6)
    out["C5a (numcl:* ub8 fixnum) (64,32,32,256)"] = {"us": t * 1e6, "GB/s": 150996992 / t / 1e9, "frac_hbm": 150996992 / t / 1e9 / peak, "kernel": k}
    from numcl_b200.einsum import _memoized
    _, desc = _memoized("abcd -> dbca")
    outs = [nc.empty((256, 32, 32, 64), "fixnum") for _ in range(3)]

    def tr_call(i):
        tin = (_lib.Tensor * 1)(sets[i][1].tensor())
        tout = (_lib.Tensor * 1)(outs[i].tensor())
        keep = (tin, tout)
        return lambda: L.ncl_einsum(C.byref(desc), keep[0], 1, keep[1], 1, _lib.OUT_FRESH)
    t, k = timed(tr_call, 3, 12, batch=6)
    out["C5b (einsum '(abcd -> dbca)) fixnum"] = {"us": t * 1e6, "GB/s": 268435456 / t / 1e9, "frac_hbm": 268435456 / t / 1e9 / peak, "kernel": k}
    del sets, outs
    # C3 / C4: contractions (TFLOP/s; denominators measured with cuBLAS in the same run)
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
        peaks = {"f64": cublas_peak(torch.float64, False), "tf32": cublas_peak(torch.float32, True)}
        for name, dt, spec_desc, shp_a, shp_b, shp_c, flops, denom, iters in [
            ("C3 matmul 8192^3 f64", "f64", mm, (8192, 8192), (8192, 8192), (8192, 8192), 2 * 8192 ** 3, peaks["f64"], 3),
            ("C3 matmul 8192^3 f32", "f32", mm, (8192, 8192), (8192, 8192), (8192, 8192), 2 * 8192 ** 3, peaks["tf32"] / 3, 5),
            ("C4 (bij bjk -> bik) b=256 512^3 f32", "f32", bmm, (256, 512, 512), (256, 512, 512), (256, 512, 512), 2 * 256 * 512 ** 3, peaks["tf32"] / 3, 10),
        ]:
            A = nc.random_array(shp_a, dt, SEED + 30, 0, -1.0, 1.0)
            B = nc.random_array(shp_b, dt, SEED + 31, 0, -1.0, 1.0)
            Cc = nc.empty(shp_c, dt)
            tin = (_lib.Tensor * 2)(A.tensor(), B.tensor())
            tout = (_lib.Tensor * 1)(Cc.tensor())
            t, k = timed(lambda i: (lambda: L.ncl_einsum(C.byref(spec_desc), tin, 2, tout, 1, _lib.OUT_FRESH)), 1, iters, warm=1)
            out[name] = {"ms": t * 1e3, "TFLOP/s": flops / t / 1e12, "denominator_TFLOP/s": denom, "frac_tensor": flops / t / 1e12 / denom, "kernel": k,
                         "denominator": "cuBLAS f64 GEMM measured in this run" if dt == "f64" else "cuBLAS TF32 GEMM measured in this run / 3"}
            del A, B, Cc
    except Exception as e:                                   # never lose the headline line to an auxiliary config
        out["contractions"] = {"error": repr(e)[:300]}
    return out


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
    # this rank's slab of the (16384*world, 16384) array: rows [rank*16384, (rank+1)*16384)
    a = nc.random_array((N_ROWS, N_COLS), "f32", SEED + 1000 * rank, 0, -1.0, 1.0)
    rows = nc.empty((N_ROWS,), "f32")
    total = nc.empty((), "f32")
    ta, trows, ttot = a.tensor(), rows.tensor(), total.tensor()
    ax1 = (C.c_int * 1)(1)

    sync_a = nc.zeros((8,), type="f32")
    sync_r = nc.empty((), type="f32")
    tsync_a, tsync_r = sync_a.tensor(), sync_r.tensor()

    def step(ev=None):
        # N > 1: the full sum goes first so that its one-element all-reduce (communication stream)
        # hides behind the axis-1 kernel, which needs no exchange; ncl_comm_join closes the step.
        if ev:
            ev[1].record(stream)
        if world > 1:
            _lib.check(L.ncl_reduce_all_sharded_async(_lib.OPS["ADD"], C.byref(ta), C.byref(ttot)))          # (sum a): local partial, async all-reduce
        else:
            _lib.check(L.ncl_reduce(_lib.OPS["ADD"], C.byref(ta), ax1, 0, C.byref(ttot), _lib.OUT_FRESH))    # (sum a)
        if ev:
            ev[2].record(stream)
            ev[0].record(stream)
        _lib.check(L.ncl_reduce(_lib.OPS["ADD"], C.byref(ta), ax1, 1, C.byref(trows), _lib.OUT_FRESH))       # (sum a :axes 1)
        if ev:
            ev[3].record(stream)
        if world > 1:
            _lib.check(L.ncl_comm_join())

    def barrier():
        stream.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # per-kernel events on every 4th step only: with N ranks in lock-step the host side of a step must
    # stay far below the 0.34 ms of device work, or the slowest rank's launch loop sets the pace
    ev_steps = [i for i in range(args.steps) if i % 4 == 0]
    evmap = {i: [torch.cuda.Event(enable_timing=True) for _ in range(4)] for i in ev_steps}
    evs = list(evmap.values())
    l0 = L.ncl_launch_count()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        # device-side alignment: ranks leave the host barrier up to milliseconds apart; a one-element
        # all-reduce enqueued BEFORE the start event makes every GPU begin its timed region together,
        # so the max-over-ranks time does not charge the first rank for the last rank's late start
        _lib.check(L.ncl_reduce_all_sharded(_lib.OPS["ADD"], C.byref(tsync_a), C.byref(tsync_r)))
    t_start.record(stream)
    h0 = time.perf_counter()
    for i in range(args.steps):
        step(evmap.get(i))
    host_enqueue = time.perf_counter() - h0
    t_end.record(stream)
    barrier()
    launches = L.ncl_launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    elapsed = t_start.elapsed_time(t_end) * 1e-3
    t_row = sum(e[0].elapsed_time(e[3]) for e in evs) * 1e-3 / len(evs)
    t_all = sum(e[1].elapsed_time(e[2]) for e in evs) * 1e-3 / len(evs)
    per_rank = None
    if world > 1:
        mine = torch.tensor([elapsed, t_row, t_all, host_enqueue / args.steps], device="cuda", dtype=torch.float64)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"ms_per_step": [round(x[0].item() / args.steps * 1e3, 4) for x in allr], "row_us": [round(x[1].item() * 1e6, 1) for x in allr],
                    "full_us": [round(x[2].item() * 1e6, 1) for x in allr], "host_enqueue_us": [round(x[3].item() * 1e6, 1) for x in allr]}
        tt = torch.tensor([elapsed, t_row, t_all], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        elapsed, t_row, t_all = tt.tolist()
    value = (BYTES_ROW + BYTES_ALL) * world * args.steps / elapsed / 1e9

    # ---- e2e: slab in pinned host memory, copies inside the timed region, through the C ABI ----------
    host = nc.NArray((N_ROWS, N_COLS), "f32", where="host")
    host.set_raw(a.raw())
    rows_h = np.empty(N_ROWS, dtype=np.float32)
    tot_h = np.empty(1, dtype=np.float32)
    e2e_steps = max(1, min(args.steps, 5))

    def e2e_step():
        _lib.check(L.ncl_h2d(a._buf, 0, host._host.ctypes.data, N_ROWS * N_COLS * 4))
        step()
        _lib.check(L.ncl_d2h(rows_h.ctypes.data, rows._buf, 0, N_ROWS * 4))
        _lib.check(L.ncl_d2h(tot_h.ctypes.data, total._buf, 0, 4))
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
    e2e_value = (BYTES_ROW + BYTES_ALL) * world * e2e_steps / e2e_t / 1e9

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    # ---- rank 0 only: CPU baseline, the other configs, the JSON line -----------------------------------
    cpu = None
    others = {}
    if world == 1:
        from oracle import oracle as O
        sample_rows = N_ROWS                  # the whole slab: ~0.35 s per pass of both scans, ~10 s in total with the host fill
        ah = O.fill_random((sample_rows, N_COLS), "f32", SEED, 0, -1.0, 1.0).values()
        reps = 12
        ct = min(cpu_reference_step(O, ah)[0] for _ in range(reps))
        cpu = {"value": (BYTES_ROW + BYTES_ALL) * (sample_rows / N_ROWS) / ct / 1e9, "unit": "GB/s", "cores": 1, "kind": "port",
               "sample": f"{sample_rows} of {N_ROWS} rows ({sample_rows * N_COLS * 4 >> 20} MiB), best of {reps}; C transcription of numcl's sequential reduce nest, not SBCL",
               "host_cores": os.cpu_count()}
        # parity spot check of the timed arrays against the oracle on the same synthetic stream
        got = rows.numpy()[:sample_rows].astype(np.float64)
        want = O.fast_sum_axis1_f32(ah).astype(np.float64)
        cpu["parity_rel_err_axis1_vs_port"] = float(np.linalg.norm(got - want) / np.linalg.norm(want))
        if not args.skip_others:
            others = time_other_configs(nc, L, torch, stream, peak, contractions=not args.no_contractions)
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("reduce_row_block")
    achieved = BYTES_ROW / t_row / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "numcl:sum axis 1 + numcl:sum full, 16384x16384 single-float slab per GPU (BASELINE configs[1])",
                   "global_shape": [N_ROWS * world, N_COLS], "sharding": "rows (outer axis); full sum = NCCL all-reduce of one float64 partial per GPU",
                   "l2": "input slab (1 GiB) is 8x the 126 MB L2, no flush needed", "algorithmic_bytes_per_step_per_gpu": BYTES_ROW + BYTES_ALL},
        "roofline": {"bound": "hbm", "kernel": "reduce_row_kernel (sum :axes 1)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "second_kernel": {"kernel": "reduce_all_kernel (sum full)" + (" local phase; the all-reduce overlaps the axis-1 kernel" if world > 1 else ""),
                                       "achieved": BYTES_ALL / t_all / 1e9, "frac": BYTES_ALL / t_all / 1e9 / peak}},
        "e2e": {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": N_ROWS * N_COLS * 4 * 1, "d2h_bytes_per_step": N_ROWS * 4 + 4,
                "steps": e2e_steps, "api": "ncl_h2d + ncl_reduce x2 + ncl_d2h (pinned host slab)"},
        "gpu_launches": int(launches),
        "host_enqueue_us_per_step": host_enqueue / args.steps * 1e6,
        "clocks": clocks,
    }
    if per_rank:
        line["per_rank"] = per_rank
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
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
