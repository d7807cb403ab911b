"""Which kernels the GEMM fuzz cases reach (scratch tool)."""
import collections, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import numcl_b200 as nc
from oracle import oracle as O
import test_gpu_fuzz as F
from test_gpu_parity import SEED, to_gpu
hist = collections.Counter()
for case in range(48):
    rng = np.random.default_rng(SEED + 3000 + case)
    dt = "f32" if case % 2 == 0 else "f64"
    b = int(rng.choice([1, 1, 2, 3])); m = int(rng.choice([128, 256, 384, 512, 130, 200, 64])); n = int(rng.choice([256, 512, 300, 64, 768]))
    k = int(rng.choice([32, 64, 96, 128, 512, 1024, 70, 2048]))
    A = F._mk(O, rng, (b, m, k), dt, -3, 3); B = F._mk(O, rng, (b, k, n), dt, -3, 3)
    nc.einsum("bij bjk -> bik", to_gpu(nc, A), to_gpu(nc, B))
    hist[nc.lib().ncl_last_kernel().decode()] += 1
print(hist)
