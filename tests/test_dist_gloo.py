"""world_size-2 gloo runs on the CPU: the host-side logic of the sharded path (row partition,
rendezvous, combining per-rank reduction partials).  The kernels themselves are GPU-only; here the
per-rank partials come from the oracle so that the combination logic is what is under test."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_row_partition_properties():
    from numcl_b200.dist import row_partition
    for n, w in [(16384, 8), (16384, 1), (10, 4), (3, 8), (256, 2), (0, 2)]:
        parts = row_partition(n, w)
        assert len(parts) == w and sum(c for _, c in parts) == n
        assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(w - 1))
        assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from numcl_b200 import dist as ndist
    from oracle import oracle as O
    ndist.init_process_group("gloo")
    rows, cols = 64, 96
    start, cnt = ndist.row_partition(rows, world)[rank]
    # every rank regenerates the full synthetic array (counter-based) and keeps its slab
    full = O.fill_random((rows, cols), "f32", 20261019, 1, -8, 8).values()
    slab = full[start:start + cnt]
    out = {}
    # full reduce: one partial per rank, combined in float64, one rounding
    out["total"] = float(ndist.allreduce_host(np.float32(O.fast_sum_all_f32(np.ascontiguousarray(slab))), "+"))
    # reduce over the sharded axis: kept-axis vector of partials
    out["cols"] = ndist.allreduce_host(slab.sum(axis=0, dtype=np.float32), "+").tolist()
    out["colmax"] = ndist.allreduce_host(slab.max(axis=0), "max").tolist()
    # reduce over the non-sharded axis needs no exchange: gather only to check
    out["rows"] = (start, O.fast_sum_axis1_f32(np.ascontiguousarray(slab)).tolist())
    ints = O.fill_random((rows, cols), "fixnum", 7, 0, -1000, 1000).values()[start:start + cnt]
    out["int_total"] = int(ndist.allreduce_host(np.int64(ints.sum()), "+"))
    dist.barrier()
    q.put((rank, out))
    dist.destroy_process_group()


def test_sharded_reduction_partials_combine_over_gloo():
    import torch.multiprocessing as mp
    from oracle import oracle as O
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = O.fill_random((64, 96), "f32", 20261019, 1, -8, 8).values()
    ints = O.fill_random((64, 96), "fixnum", 7, 0, -1000, 1000).values()
    for r in range(world):
        assert res[r]["total"] == float(O.fast_sum_all_f32(np.ascontiguousarray(full)))          # exact inputs: any order agrees
        assert res[r]["cols"] == full.sum(axis=0, dtype=np.float32).tolist()
        assert res[r]["colmax"] == full.max(axis=0).tolist()
        assert res[r]["int_total"] == int(ints.sum())
    rows = np.concatenate([np.array(res[r]["rows"][1], dtype=np.float32) for r in range(world)])
    assert np.array_equal(rows, O.fast_sum_axis1_f32(np.ascontiguousarray(full)))
