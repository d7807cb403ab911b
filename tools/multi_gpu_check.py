"""Run under torchrun (one rank per GPU): sharded reductions through the library's NCCL path,
checked against the oracle on the full (concatenated) synthetic array."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.environ["NUMCL_CUDA_DEVICE"] = str(local)
import torch  # noqa: E402

torch.cuda.set_device(local)
import numcl_b200 as nc  # noqa: E402
from numcl_b200 import dist as ndist  # noqa: E402
from oracle import oracle as O  # noqa: E402

ndist.init_process_group("nccl")
ndist.init_comm()
rows, cols = 512 * world, 2048
start, cnt = ndist.row_partition(rows, world)[rank]
for dt, dist_kind, lo, hi in [("f32", 1, -8, 8), ("fixnum", 0, -10 ** 6, 10 ** 6), ("f64", 1, -8, 8)]:
    full = O.fill_random((rows, cols), dt, 20261019, dist_kind, lo, hi)
    fv = full.values()
    slab = nc.asarray(np.ascontiguousarray(fv[start:start + cnt]), type=dt)
    total = ndist.reduce_sharded("+", slab, None)
    want_total = fv.astype(np.float64).sum() if dt != "fixnum" else int(fv.sum())
    assert total == want_total, (dt, total, want_total)
    col = ndist.reduce_sharded("+", slab, (0,)).numpy()
    assert np.array_equal(col.astype(np.float64), fv.astype(np.float64).sum(axis=0)), dt
    mx = ndist.reduce_sharded("max", slab, (0,)).numpy()
    assert np.array_equal(mx, fv.max(axis=0)), dt
    row = ndist.reduce_sharded("+", slab, (1,)).numpy()                    # no collective: this rank's rows
    assert np.array_equal(row.astype(np.float64), fv[start:start + cnt].astype(np.float64).sum(axis=1)), dt
# ---- the mailbox exchange under stress: many back-to-back calls (two-parity ring), max / min, an EMPTY slab on the last
# rank, and the whole thing recorded into a CUDA graph and replayed with fresh data -------------------------------------
import ctypes as C  # noqa: E402
from numcl_b200 import _lib  # noqa: E402
L = nc.lib()
rows2, cols2 = 64 * world, 1024
start2, cnt2 = ndist.row_partition(rows2, world)[rank]
fulls = [O.fill_random((rows2, cols2), "f32", 99 + k, 1, -8, 8).values() for k in range(5)]
slabs = [nc.asarray(np.ascontiguousarray(f[start2:start2 + cnt2]), type="f32") for f in fulls]
outs = [nc.empty((), type="f32") for _ in fulls]
tsl = [x.tensor() for x in slabs]
tou = [x.tensor() for x in outs]
for rep in range(4):                                               # 20 exchanges without a host sync in between
    for k in range(5):
        _lib.check(L.ncl_reduce_all_sharded_async(_lib.OPS["ADD"], C.byref(tsl[k]), C.byref(tou[k])))
_lib.check(L.ncl_comm_join())
for k in range(5):
    assert outs[k].item() == float(fulls[k].astype(np.float64).sum()), (k, outs[k].item())
for opname, fn in (("MAX", np.max), ("MIN", np.min)):
    for k in range(5):
        _lib.check(L.ncl_reduce_all_sharded(_lib.OPS[opname], C.byref(tsl[k]), C.byref(tou[k])))
        assert outs[k].item() == float(fn(fulls[k])), (opname, k)
# empty slab: rank world-1 owns no rows (n_rows < world); its max / min contribution must be the identity
tiny = O.fill_random((world - 1, 256), "f32", 5, 1, -8, 8).values()
st, ct = ndist.row_partition(world - 1, world)[rank]
tslab = nc.asarray(np.ascontiguousarray(tiny[st:st + ct]).reshape(ct, 256), type="f32") if ct else nc.empty((0, 256), type="f32")
for opname, want in (("MAX", float(tiny.max())), ("MIN", float(tiny.min())), ("ADD", float(tiny.astype(np.float64).sum()))):
    got = ndist.reduce_sharded({"MAX": "max", "MIN": "min", "ADD": "+"}[opname], tslab, None)
    assert got == want, (opname, got, want, rank)
# recorded into a graph: sharded full sum + a row sum per array; replayed three times, then with new contents
rowout = [nc.empty((cnt2,), type="f32") for _ in fulls]
trow = [x.tensor() for x in rowout]
ax1 = (C.c_int * 1)(1)


def step(k):
    _lib.check(L.ncl_reduce_all_sharded_async(_lib.OPS["ADD"], C.byref(tsl[k]), C.byref(tou[k])))
    _lib.check(L.ncl_reduce(_lib.OPS["ADD"], C.byref(tsl[k]), ax1, 1, C.byref(trow[k]), _lib.OUT_FRESH))
    _lib.check(L.ncl_comm_join())


for k in range(5):
    step(k)
_lib.check(L.ncl_sync())
with _lib.Graph() as g:
    for k in range(5):
        step(k)
for rep in range(3):
    for o in outs:
        o.set_raw(np.zeros(o.nbytes, dtype=np.uint8))
    g.launch()
    for k in range(5):
        assert outs[k].item() == float(fulls[k].astype(np.float64).sum()), ("graph", rep, k)
        assert np.array_equal(rowout[k].numpy().astype(np.float64), fulls[k][start2:start2 + cnt2].astype(np.float64).sum(axis=1))
fulls2 = [O.fill_random((rows2, cols2), "f32", 199 + k, 1, -8, 8).values() for k in range(5)]
for k in range(5):
    slabs[k].set_raw(np.ascontiguousarray(fulls2[k][start2:start2 + cnt2]))
g.launch()
for k in range(5):
    assert outs[k].item() == float(fulls2[k].astype(np.float64).sum()), ("graph, new data", k)
torch.distributed.barrier()
if rank == 0:
    print(f"multi-gpu check ok: world={world}, kernels launched on rank 0 = {nc.lib().ncl_launch_count()}")
torch.distributed.destroy_process_group()
