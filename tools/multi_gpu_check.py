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
torch.distributed.barrier()
if rank == 0:
    print(f"multi-gpu check ok: world={world}, kernels launched on rank 0 = {nc.lib().ncl_launch_count()}")
torch.distributed.destroy_process_group()
