"""Seeded differential fuzzing of the CUDA path against the oracle: random broadcast shapes, reduction axes and
einsum specs over random element types.  Float data are small integers stored as floats (fill_random dist 1), so
every accumulation order is exact and EVERYTHING here is compared bit for bit -- whichever kernel family
(elementwise / widening / reduce / transpose / GEMM / generic nest) the dispatcher picks."""
import numpy as np
import pytest

import refsem
from test_gpu_parity import SEED, assert_bit_exact, rand, to_gpu

pytestmark = pytest.mark.gpu

DTYPES = ["f32", "f64", "fixnum", "ub8", "sb8", "ub15", "sb16", "ub31", "sb32", "sb64", "bit", "ub4"]
DIMS = [1, 2, 3, 4, 5, 7, 8, 16, 31, 64, 130]


def _mk(oracle, rng, shape, dt, lo=-6, hi=6):
    from numcl_b200.dtypes import dtype
    d = dtype(dt)
    if dt in ("f32", "f64"):
        return rand(oracle, shape, dt, int(rng.integers(1 << 30)), 1, lo, hi)
    return rand(oracle, shape, dt, int(rng.integers(1 << 30)), 0, max(d.lo, lo), min(d.hi, hi))


def _bshapes(rng):
    """two shapes that numpy-broadcast against each other (right-aligned; dims equal or 1), total <= ~70k elements"""
    rank = int(rng.integers(1, 5))
    full = []
    for _ in range(rank):
        full.append(int(rng.choice(DIMS)))
    while np.prod(full) > 70000:
        full[int(np.argmax(full))] = int(rng.choice(DIMS[:6]))
    def cut(sh):
        sh = [d if rng.random() > 0.3 else 1 for d in sh]
        drop = int(rng.integers(0, len(sh)))               # leading dims may be missing altogether
        return tuple(sh[drop:])
    a, b = cut(full), cut(full)
    if rng.random() < 0.5:
        a = tuple(full)
    return a, b


@pytest.mark.parametrize("case", range(120))
def test_fuzz_broadcast_bit_exact(nc, oracle, case):
    rng = np.random.default_rng(SEED + case)
    op = str(rng.choice(["+", "-", "*", "max", "min"]))
    tx, ty = str(rng.choice(DTYPES)), str(rng.choice(DTYPES))
    sa, sb = _bshapes(rng)
    x, y = _mk(oracle, rng, sa, tx), _mk(oracle, rng, sb, ty)
    f = {"+": nc.add, "-": nc.sub, "*": nc.mul, "max": nc.maximum, "min": nc.minimum}[op]
    want = refsem.arith(oracle, op, [x, y])
    assert_bit_exact(f(to_gpu(nc, x), to_gpu(nc, y)), want)


@pytest.mark.parametrize("case", range(80))
def test_fuzz_reduce_bit_exact(nc, oracle, case):
    rng = np.random.default_rng(SEED + 1000 + case)
    dt = str(rng.choice(["f32", "f64", "fixnum", "ub8", "sb16", "sb32"]))
    rank = int(rng.integers(1, 5))
    shape = [int(rng.choice(DIMS)) for _ in range(rank)]
    while np.prod(shape) > 60000:
        shape[int(np.argmax(shape))] = int(rng.choice(DIMS[:6]))
    axes = sorted(int(a) for a in rng.choice(rank, size=int(rng.integers(1, rank + 1)), replace=False))
    fn = str(rng.choice(["+", "max", "min"]))
    a = _mk(oracle, rng, tuple(shape), dt, -5, 5)
    want = refsem.reduce_array(oracle, fn, a, tuple(axes))
    got = {"+": nc.sum, "max": nc.amax, "min": nc.amin}[fn](to_gpu(nc, a), axes=tuple(axes))
    if want.size == 1 and not hasattr(got, "raw"):
        assert got == want.values().reshape(-1)[0]
    else:
        assert_bit_exact(got, want)


def _spec(rng):
    """random einsum spec with 1 or 2 inputs over indices a..e, default transform, output = any subset in any order"""
    letters = "abcde"
    n_in = int(rng.integers(1, 3))
    ins = []
    for _ in range(n_in):
        r = int(rng.integers(1, 4))
        ins.append("".join(rng.choice(list(letters), size=r, replace=False)))
    used = sorted(set("".join(ins)))
    k = int(rng.integers(0, len(used) + 1))
    out = "".join(rng.permutation(used)[:k])
    return ins, out


@pytest.mark.parametrize("case", range(120))
def test_fuzz_einsum_bit_exact(nc, oracle, case):
    rng = np.random.default_rng(SEED + 2000 + case)
    ins, out = _spec(rng)
    ext = {c: int(rng.choice([1, 2, 3, 4, 5, 8, 16, 33, 64])) for c in "abcde"}
    big = rng.random() < 0.25                       # sometimes GEMM-sized, tile-aligned dims
    if big:
        ext = {c: int(rng.choice([32, 64, 128, 256])) for c in "abcde"}
    total = np.prod([ext[c] for c in set("".join(ins))])
    while total > (1 << 22):
        c = max(set("".join(ins)), key=lambda z: ext[z]); ext[c] //= 2
        total = np.prod([ext[c] for c in set("".join(ins))])
    family = str(rng.choice(["f32", "f64", "int"]))
    arrs = []
    for s in ins:
        dt = family if family != "int" else str(rng.choice(["fixnum", "ub8", "sb16", "sb32"]))
        arrs.append(_mk(oracle, rng, tuple(ext[c] for c in s), dt, -3, 3))
    spec = " ".join(ins) + " -> " + out
    want = refsem.einsum(oracle, spec, arrs)
    got = nc.einsum(spec, *[to_gpu(nc, a) for a in arrs])
    if out == "":
        w = want.values().reshape(-1)[0] if hasattr(want, "values") else want
        g = got.numpy().reshape(-1)[0] if hasattr(got, "numpy") else got
        assert g == w, (spec, g, w)
    else:
        assert_bit_exact(got, want)


@pytest.mark.parametrize("case", range(48))
def test_fuzz_gemm_shapes_bit_exact(nc, oracle, case):
    """GEMM-shaped specs over random tile-aligned and ragged sizes; integer-valued float data make the tensor-core
    paths (DMMA, 3xTF32 on tcgen05 -- packed, fused, fused on CTA pairs) exact, so they too are compared bit for bit
    with numcl's nest"""
    rng = np.random.default_rng(SEED + 3000 + case)
    dt = "f32" if case % 2 == 0 else "f64"
    b = int(rng.choice([1, 1, 2, 3]))
    m = int(rng.choice([128, 256, 384, 512, 130, 200, 64]))
    n = int(rng.choice([256, 512, 300, 64, 768]))
    k = int(rng.choice([32, 64, 96, 128, 512, 1024, 70, 2048]))
    A = _mk(oracle, rng, (b, m, k), dt, -3, 3)
    B = _mk(oracle, rng, (b, k, n), dt, -3, 3)
    if dt == "f32":
        ref = oracle.fast_bmm_f32(A.values(), B.values())
    else:
        ref = np.stack([oracle.fast_matmul(A.values()[i], B.values()[i]) for i in range(b)])
    if b == 1 and rng.random() < 0.5:
        got = nc.matmul(to_gpu(nc, oracle.OArr.from_values(A.values()[0], dt)), to_gpu(nc, oracle.OArr.from_values(B.values()[0], dt)))
        ref = ref[0]
    else:
        got = nc.einsum("bij bjk -> bik", to_gpu(nc, A), to_gpu(nc, B))
    assert got.dtype.name == dt
    assert np.array_equal(got.numpy(), ref), (nc.lib().ncl_last_kernel(), b, m, n, k)
