"""3xTF32 accuracy probe: error of the tcgen05 GEMM vs float64 truth as K grows."""
import sys
import numpy as np
sys.path.insert(0, ".")
import numcl_b200 as nc

rng = np.random.default_rng(0)
for K in (256, 512, 1024, 2048, 4096, 8192):
    A = rng.uniform(-1, 1, (256, K)).astype(np.float32)
    B = rng.uniform(-1, 1, (K, 256)).astype(np.float32)
    got = nc.matmul(nc.asarray(A, type="f32"), nc.asarray(B, type="f32")).numpy().astype(np.float64)
    k = nc.lib().ncl_last_kernel().decode()
    ref = A.astype(np.float64) @ B.astype(np.float64)
    err = got - ref
    rel = np.linalg.norm(err) / np.linalg.norm(ref)
    shrink = np.mean(err * np.sign(ref)) / np.mean(np.abs(ref))       # < 0: magnitudes systematically too small
    seq = np.zeros((256, 256), dtype=np.float32)
    print(f"K={K:5d} kernel={k} rel_frob={rel:.3e} signed_bias/mean|ref|={shrink:+.3e}")
