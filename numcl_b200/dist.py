"""Multi-GPU layer: one process per GPU, the outer axis sharded, NCCL only where partials meet.

The reference has no collectives.  The path shards naturally:
  * broadcast elementwise, batched einsum, reductions over non-sharded axes: independent row slabs,
    no data-path collective at all;
  * full reduce: each GPU reduces its slab to ONE float64/int64 partial on the device, the partials
    are all-reduced over NVLink (ncl_reduce_all_sharded -> ncclAllReduce of one element) and rounded
    once into the result element type;
  * reduce over the sharded (outer) axis: each GPU yields a full-length partial vector, all-reduced
    in place (ncl_allreduce).
torch.distributed is plumbing only: rendezvous, the unique-id broadcast, barriers.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from . import typeinfer as ti
from .array import NArray, empty

_OPS = {"+": "ADD", "*": "MUL", "max": "MAX", "min": "MIN"}
_comm_ready = False


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def row_partition(n_rows: int, world: int):
    """Contiguous, balanced split of the outer axis: [(start, count)] per rank; the first
    n_rows % world ranks get one extra row."""
    base, extra = divmod(n_rows, world)
    out, start = [], 0
    for r in range(world):
        cnt = base + (1 if r < extra else 0)
        out.append((start, cnt))
        start += cnt
    return out


def init_process_group(backend: str | None = None):
    """torch.distributed rendezvous on 127.0.0.1 (the container hostname may not resolve)."""
    import torch.distributed as dist
    if dist.is_initialized():
        return dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    rank, world = rank_world()
    if backend is None:
        import torch
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    dist.init_process_group(backend, rank=rank, world_size=world)
    return dist


def init_comm():
    """Create the library's NCCL communicator: rank 0 makes the unique id, torch.distributed carries it."""
    global _comm_ready
    rank, world = rank_world()
    if _comm_ready or world == 1:
        return
    import torch
    dist = init_process_group()
    L = _lib.lib()
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = (C.c_uint8 * 128)()
        _lib.check(L.ncl_comm_unique_id(buf))
        uid = torch.tensor(list(buf), dtype=torch.uint8)
    if dist.get_backend() == "nccl":
        uid = uid.cuda()
    dist.broadcast(uid, 0)
    buf = (C.c_uint8 * 128)(*uid.cpu().tolist())
    _lib.check(L.ncl_comm_init(buf, rank, world))
    _comm_ready = True


def allreduce_host(values: np.ndarray, op: str = "+") -> np.ndarray:
    """Combine per-rank partials that live on the host (small kept-axis vectors, scalars) through
    torch.distributed -- gloo on CPU-only hosts, NCCL otherwise.  float32 partials travel as float64
    and are rounded once, like the device path."""
    import torch
    import torch.distributed as dist
    rank, world = rank_world()
    v = np.asarray(values)
    if world == 1:
        return v
    wide = v.astype(np.float64) if v.dtype.kind == "f" else v.astype(np.int64)
    t = torch.from_numpy(np.ascontiguousarray(wide))
    if dist.get_backend() == "nccl":
        t = t.cuda()
    red = {"+": dist.ReduceOp.SUM, "*": dist.ReduceOp.PRODUCT, "max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN}[op]
    dist.all_reduce(t, op=red)
    return t.cpu().numpy().astype(v.dtype)


def reduce_sharded(fn: str, local: NArray, axes=None, type=None):
    """numcl:sum / prod / amax / amin of an array whose OUTER axis is sharded across ranks.
    `local` is this rank's slab.  Returns the slab of the result (axes without 0), or the full
    replicated result (axes containing 0 / a full reduce)."""
    from .reduce import _fresh
    f = fn.lower().split(":")[-1]
    op = _OPS[f]
    rank, world = rank_world()
    ax = list(range(local.rank)) if axes is None else sorted([axes] if isinstance(axes, int) else list(axes))
    if 0 not in ax or world == 1:
        return _fresh(op, local, axes, type)                       # no exchange: rows are independent
    init_comm()
    L = _lib.lib()
    rtype = type or ti.upgrade(ti.reduce_result_type(f, ti.type_of_dtype(local.dtype)))
    if len(ax) == local.rank:                                       # full reduce: one element crosses NVLink
        r = empty((), type=rtype)
        ta, tr = local.tensor(), r.tensor()
        _lib.check(L.ncl_reduce_all_sharded(_lib.OPS[op], C.byref(ta), C.byref(tr)))
        return r.item()
    part = _fresh(op, local, axes, type)                            # kept-axis vector of partials
    tp = part.tensor()
    _lib.check(L.ncl_allreduce(_lib.OPS[op], C.byref(tp)))
    return part


def sum_sharded(local: NArray, axes=None, type=None):
    return reduce_sharded("+", local, axes, type)
