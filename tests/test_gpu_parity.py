"""Parity of the CUDA path (through the C ABI) with the CPU oracle on seeded inputs.

Bar (north_star): bit-exact for integer / bit / fixnum element types and for single- and
double-float ELEMENTWISE ops; for float reductions bit-exact on exactly-summable inputs and
<= 1e-5 (single) / 1e-12 (double) relative otherwise.  Inputs travel as raw base-vector bytes, so
"bit-exact" is literally a comparison of the stored words.
"""
import numpy as np
import pytest

import refsem
from numcl_b200 import typeinfer as ti

pytestmark = pytest.mark.gpu
SEED = 20261019


def to_gpu(nc, o):
    a = nc.NArray(o.shape, o.dtype)
    a.set_raw(o.storage)
    return a


def live_bytes(dtype_name, size):
    from numcl_b200.dtypes import dtype
    return (size * dtype(dtype_name).slot_bits + 7) // 8


def assert_bit_exact(g, o):
    """same element type, same shape, same stored words"""
    assert g.dtype.name == o.dtype, (g.dtype.name, o.dtype)
    assert tuple(g.shape) == tuple(o.shape)
    n = live_bytes(o.dtype, o.size)
    gb, ob = g.raw()[:n], o.storage[:n]
    if np.array_equal(gb, ob):
        return
    gv, ov = g.numpy().reshape(-1), o.values().reshape(-1)
    if o.dtype in ("f32", "f64"):
        # inf + -inf etc.: SBCL traps (:invalid) where IEEE hardware returns a NaN, and x86 and the GPU
        # pick different default-NaN payloads.  NaN positions must coincide; every other word must match.
        wt = np.uint32 if o.dtype == "f32" else np.uint64
        gw, ow = gb[: gv.size * gv.itemsize].view(wt), ob[: ov.size * ov.itemsize].view(wt)
        both_nan = (gv != gv) & (ov != ov)
        bad = np.nonzero((gw != ow) & ~both_nan)[0]
    else:
        bad = np.nonzero(gv != ov)[0]
    if bad.size:
        raise AssertionError(f"{bad.size} of {gv.size} elements differ; first at {bad[:5]}: gpu {gv[bad[:5]]} oracle {ov[bad[:5]]}")


def rand(O, shape, dtype, seed, dist=0, lo=0.0, hi=1.0):
    return O.fill_random(shape, dtype, seed, dist, lo, hi)


# ---- K1: broadcast elementwise -------------------------------------------------------------------------
@pytest.mark.parametrize("shape_a,shape_b", [((256, 512), (512,)), ((64, 1, 32), (5, 32)), ((7, 13), (7, 1)), ((1000,), ()),
                                              ((33, 5, 250), (250,)), ((4096, 4096), (4096,))])
@pytest.mark.parametrize("dist", [2, 1])
def test_add_f32_bit_exact(nc, oracle, shape_a, shape_b, dist):
    """config C1: (numcl:+ a b), sets R (specials: signed zeros, inf, denormals) and E (small ints)."""
    lo, hi = (-1.0, 1.0) if dist == 2 else (-8, 8)
    a, b = rand(oracle, shape_a, "f32", SEED, dist, lo, hi), rand(oracle, shape_b, "f32", SEED + 1, dist, lo, hi)
    want = refsem.arith(oracle, "+", [a, b])
    got = nc.add(to_gpu(nc, a), to_gpu(nc, b))
    assert_bit_exact(got, want)
    if a.size >= 1024 and shape_a[-1] % 4 == 0:
        assert nc.lib().ncl_last_kernel() == b"bcast_vec"


def test_add_f32_c1_matches_fast_oracle(nc, oracle):
    """the typed two-pass loop the CPU baseline times == the interpreter == the GPU, at C1 size"""
    a, b = rand(oracle, (4096, 4096), "f32", SEED, 2, -1.0, 1.0), rand(oracle, (4096,), "f32", SEED + 1, 2, -1.0, 1.0)
    fast = oracle.fast_add2_f32(a.values(), b.values())
    got = nc.add(to_gpu(nc, a), to_gpu(nc, b)).raw()[: 4096 * 4096 * 4].view(np.uint32)
    want = fast.view(np.uint32).reshape(-1)
    nan = np.isnan(fast.reshape(-1)) & np.isnan(got.view(np.float32))     # inf + -inf: payloads differ by platform
    assert np.array_equal(got[~nan], want[~nan])


def test_device_generator_matches_oracle(nc, oracle):
    for dt, dist, lo, hi in [("f32", 2, -1.0, 1.0), ("f64", 2, -1.0, 1.0), ("f32", 1, -8, 8), ("ub8", 0, 0, 255),
                             ("fixnum", 0, -(2 ** 40), 2 ** 40), ("bit", 0, 0, 1), ("sb16", 0, -300, 300)]:
        o = rand(oracle, (1000, 37), dt, SEED, dist, lo, hi)
        g = nc.random_array((1000, 37), dt, SEED, dist, lo, hi)
        assert_bit_exact(g, o)


OPS2 = ["+", "-", "*", "/", "max", "min"]


@pytest.mark.parametrize("op", OPS2)
@pytest.mark.parametrize("tx,ty", [("f32", "f32"), ("f64", "f64"), ("f32", "f64"), ("ub8", "fixnum"), ("fixnum", "f32"), ("bit", "ub8"),
                                   ("sb16", "ub31"), ("ub4", "sb8"), ("f64", "sb32"), ("ub8", "ub8")])
def test_binary_dtype_matrix_bit_exact(nc, oracle, op, tx, ty):
    def mk(t, seed, shape):
        if t in ("f32", "f64"):
            return rand(oracle, shape, t, seed, 2 if op != "/" else 0, -4.0, 4.0)
        from numcl_b200.dtypes import dtype
        d = dtype(t)
        lo, hi = max(d.lo, -1000), min(d.hi, 1000)
        if op == "/":
            lo = max(lo, 1)
        return rand(oracle, shape, t, seed, 0, lo, hi)
    x, y = mk(tx, SEED, (48, 40)), mk(ty, SEED + 7, (40,))
    want = refsem.arith(oracle, op, [x, y])
    f = {"+": nc.add, "-": nc.sub, "*": nc.mul, "/": nc.div, "max": nc.maximum, "min": nc.minimum}[op]
    assert_bit_exact(f(to_gpu(nc, x), to_gpu(nc, y)), want)


def test_u8_times_fixnum_wraps_at_63_bits(nc, oracle):
    """config C5a, set W: products overflow 63 bits and must wrap exactly like %coerce into fixnum"""
    x = rand(oracle, (8, 4, 4, 256), "ub8", SEED, 0, 0, 255)
    y = rand(oracle, (256,), "fixnum", SEED + 1, 0, 2 ** 61 - 1000, 2 ** 61 + 1000)
    want = refsem.arith(oracle, "*", [x, y])
    got = nc.mul(to_gpu(nc, x), to_gpu(nc, y))
    assert want.dtype == "fixnum"
    assert_bit_exact(got, want)
    assert nc.lib().ncl_last_kernel() == b"bcast_widen"
    fast = oracle.fast_mul2_u8_fix(x.values().astype(np.uint8), y.storage[: 256 * 8].view(np.int64))
    assert np.array_equal(got.raw()[: x.size * 8].view(np.int64), fast.reshape(-1))


@pytest.mark.parametrize("op", ["+", "-", "*", "max", "min"])
@pytest.mark.parametrize("narrow", ["ub8", "sb8", "ub15", "sb16", "ub31", "sb32"])
@pytest.mark.parametrize("wide", ["fixnum", "sb64"])
def test_widening_integer_family_bit_exact(nc, oracle, op, narrow, wide):
    """k_ew_widen.cu: narrow integer array (op) 64-bit operand -> 64-bit result, in both operand orders and in all
    four layouts of the 64-bit operand (row vector shorter / longer than a warp item, same shape, scalar), with a
    ragged last item; bit-exact against numcl's broadcast, including the 63/64-bit wrap"""
    from numcl_b200.dtypes import dtype
    d = dtype(narrow)
    f = {"+": nc.add, "-": nc.sub, "*": nc.mul, "max": nc.maximum, "min": nc.minimum}[op]
    big = 2 ** 61 if op == "*" else 2 ** 40
    for shape_x, shape_y in [((37, 64), (64,)), ((5, 512), (512,)), ((3, 7, 130), (3, 7, 130)), ((1000,), ())]:
        x = rand(oracle, shape_x, narrow, SEED, 0, max(d.lo, -30000), min(d.hi, 30000))
        if shape_y == ():
            for s in (7, -3):
                assert_bit_exact(f(to_gpu(nc, x), s), refsem.arith(oracle, op, [x, s]))
                assert_bit_exact(f(s, to_gpu(nc, x)), refsem.arith(oracle, op, [s, x]))
            continue
        y = rand(oracle, shape_y, wide, SEED + 1, 0, -big, big)
        for args in ([x, y], [y, x]):
            want = refsem.arith(oracle, op, args)
            got = f(*[to_gpu(nc, a) for a in args])
            assert_bit_exact(got, want)
            if want.dtype in ("fixnum", "sb64"):
                assert nc.lib().ncl_last_kernel() == b"bcast_widen"


@pytest.mark.parametrize("op", ["=/bit", "</bit", ">=/bit"])
def test_comparisons_to_bit_arrays(nc, oracle, op):
    x, y = rand(oracle, (64, 128), "f32", SEED, 1, -3, 3), rand(oracle, (128,), "f32", SEED + 3, 1, -3, 3)
    want = refsem.broadcast(oracle, op, x, y)
    assert want.dtype == "bit"
    assert_bit_exact(nc.broadcast(op, to_gpu(nc, x), to_gpu(nc, y)), want)


def test_scalar_operands(nc, oracle):
    a = rand(oracle, (100, 30), "f32", SEED, 2, -1.0, 1.0)
    for op, s in [("+", 2.5), ("*", 3), ("-", 1), ("/", 2), ("max", 0.25)]:
        f = {"+": nc.add, "-": nc.sub, "*": nc.mul, "/": nc.div, "max": nc.maximum}[op]
        assert_bit_exact(f(to_gpu(nc, a), s), refsem.arith(oracle, op, [a, s]))
    assert_bit_exact(nc.sub(to_gpu(nc, a)), refsem.arith(oracle, "-", [a]))      # (- a) = 0 - a
    b = rand(oracle, (100, 30), "f32", SEED, 0, 1.0, 2.0)
    assert_bit_exact(nc.div(to_gpu(nc, b)), refsem.arith(oracle, "/", [b]))      # (/ a) = 1 / a


def test_three_way_fold(nc, oracle):
    a, b, c = (rand(oracle, s, "f32", SEED + i, 2, -1.0, 1.0) for i, s in enumerate([(50, 20), (20,), (50, 1)]))
    assert_bit_exact(nc.add(to_gpu(nc, a), to_gpu(nc, b), to_gpu(nc, c)), refsem.arith(oracle, "+", [a, b, c]))
    i1, i2, i3 = (rand(oracle, (9, 11), t, SEED + k, 0, 0, 200) for k, t in enumerate(["ub8", "ub8", "sb16"]))
    assert_bit_exact(nc.mul(to_gpu(nc, i1), to_gpu(nc, i2), to_gpu(nc, i3)), refsem.arith(oracle, "*", [i1, i2, i3]))


def test_broadcast_shape_errors(nc):
    with pytest.raises(AssertionError):
        nc.add(nc.zeros((3, 4), type="f32"), nc.zeros((5,), type="f32"))


# ---- K3: reductions -------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape,axes", [((5, 5, 5), None), ((5, 5, 5), (0,)), ((5, 5, 5), (0, 1)), ((5, 5, 5), (1,)), ((5, 5, 5), (2,)),
                                        ((5, 5, 5), (0, 2)), ((64, 1024), (1,)), ((64, 1024), (0,)), ((300, 77), (1,)), ((300, 77), (0,)),
                                        ((8, 16, 4096), (2,)), ((16, 2048, 8), (1,)), ((1 << 16,), None)])
@pytest.mark.parametrize("fn", ["+", "max", "min"])
@pytest.mark.parametrize("dt", ["ub8", "fixnum", "f32", "f64"])
def test_reduce_exact_inputs_bit_exact(nc, oracle, shape, axes, fn, dt):
    """integers, and floats holding small integers (every partial sum exact): all orders agree"""
    a = rand(oracle, shape, dt, SEED, 1, 0 if dt == "ub8" else -8, 8)
    want = refsem.reduce_array(oracle, fn, a, axes)
    got = {"+": nc.sum, "max": nc.amax, "min": nc.amin}[fn](to_gpu(nc, a), axes=axes)
    if want.shape == ():
        assert got == want.values().reshape(-1)[0]
    else:
        assert_bit_exact(got, want)


def test_prod_int(nc, oracle):
    a = rand(oracle, (6, 40), "fixnum", SEED, 0, -3, 3)
    for axes in (None, (0,), (1,)):
        want = refsem.reduce_array(oracle, "*", a, axes)
        got = nc.prod(to_gpu(nc, a), axes=axes)
        if want.shape == ():
            assert got == want.values().reshape(-1)[0]
        else:
            assert_bit_exact(got, want)


@pytest.mark.parametrize("dt,tol", [("f32", 1e-5), ("f64", 1e-12)])
@pytest.mark.parametrize("shape,axes", [((512, 4096), (1,)), ((512, 4096), (0,)), ((512, 4096), None), ((33, 1000), (1,))])
def test_reduce_random_floats_within_tolerance(nc, oracle, dt, tol, shape, axes):
    a = rand(oracle, shape, dt, SEED, 0, -1.0, 1.0)
    want = refsem.reduce_array(oracle, "+", a, axes)
    got = nc.sum(to_gpu(nc, a), axes=axes)
    truth = a.values().astype(np.float64).sum(axis=axes)
    g = np.asarray(got.numpy() if hasattr(got, "numpy") else got, dtype=np.float64)
    w = want.values().astype(np.float64).reshape(g.shape)
    assert np.linalg.norm(g - w) <= tol * max(np.linalg.norm(w), 1e-30) + (tol * np.abs(truth).max() if g.shape == () else 0)
    # and the tree order is at least as close to the float64 truth as the reference's sequential scan
    assert np.linalg.norm(g - truth) <= max(np.linalg.norm(w - truth) * 1.5, tol * np.linalg.norm(truth) + 1e-30)


def test_sum_c2_row_sums(nc, oracle):
    """config C2a at full width (16384 columns), 256 rows: GPU vs the sequential scan numcl does"""
    a = rand(oracle, (256, 16384), "f32", SEED, 0, -1.0, 1.0)
    want = oracle.fast_sum_axis1_f32(a.values())
    got = nc.sum(to_gpu(nc, a), axes=1)
    assert nc.lib().ncl_last_kernel() == b"reduce_row_block"
    g = got.numpy().astype(np.float64)
    assert np.linalg.norm(g - want) <= 1e-5 * np.linalg.norm(want)


def test_sum_c2_full_size_properties(nc, oracle):
    """config C2 at BASELINE size, set E (integers in [-1, 1]: every order gives the same bits)"""
    n = 16384
    a = nc.random_array((n, n), "f32", SEED, 1, -1, 1)
    rows = nc.sum(a, axes=1)
    total = nc.sum(a)
    assert nc.lib().ncl_last_kernel() == b"reduce_all"
    cols = nc.sum(a, axes=0)
    # checksum of checksums: the three agree exactly because every partial sum is a small integer
    assert float(rows.numpy().astype(np.float64).sum()) == float(total) == float(cols.numpy().astype(np.float64).sum())
    # first rows against the oracle's sequential scan of the identical synthetic stream
    o = rand(oracle, (64, n), "f32", SEED, 1, -1, 1)
    assert np.array_equal(rows.numpy()[:64], oracle.fast_sum_axis1_f32(o.values()))


def test_mean_and_reduce_initial_element(nc, oracle):
    a = rand(oracle, (20, 30), "f32", SEED, 1, -8, 8)
    s = refsem.reduce_array(oracle, "+", a, (1,))
    want = refsem.broadcast(oracle, "/", s, 30)
    assert_bit_exact(nc.mean(to_gpu(nc, a), axes=1), want)
    r = nc.reduce_array("max", to_gpu(nc, a), axes=(0,), initial_element=3.0)
    w = oracle.reduce_array("max", a, [0], "f32", 3.0)
    assert_bit_exact(r, w)
    # amax starts from most-negative-single-float, not -infinity (5reduce.lisp:113-119)
    ninf = nc.full((4, 8), float("-inf"), type="f32")
    assert np.all(nc.amax(ninf, axes=1).numpy() == np.float32(-3.4028234663852886e38))


# ---- K5: permutations ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(61, 7, 13, 250), (16, 8, 8, 64), (3, 1, 2, 5)])
@pytest.mark.parametrize("dt", ["fixnum", "f32", "ub8", "f64"])
def test_transpose_abcd_dbca_bit_exact(nc, oracle, shape, dt):
    """config C5b (+ the odd shape): `(abcd -> dbca)`; integer inputs come out as fixnum arrays"""
    a = rand(oracle, shape, dt, SEED, 2 if dt in ("f32", "f64") else 0, -100 if dt != "ub8" else 0, 100)
    want = refsem.einsum(oracle, "abcd -> dbca", [a])
    got = nc.einsum("abcd -> dbca", to_gpu(nc, a))
    assert_bit_exact(got, want)
    if a.size > 1000:      # ub8 -> fixnum changes the element type on the way: the converting variant of the tiled kernel
        assert nc.lib().ncl_last_kernel() in ((b"transpose_convert",) if dt == "ub8" else (b"transpose_tile", b"transpose_thin"))


def test_transpose_c5_chain(nc, oracle):
    """C5a output (fixnum words) fed to C5b, against the typed fast loops, at 1/8 of BASELINE size"""
    x = rand(oracle, (8, 32, 32, 256), "ub8", SEED, 0, 0, 255)
    y = rand(oracle, (256,), "fixnum", SEED + 1, 0, -(2 ** 40), 2 ** 40)
    prod = nc.mul(to_gpu(nc, x), to_gpu(nc, y))
    out = nc.einsum("abcd -> dbca", prod)
    pw = oracle.fast_mul2_u8_fix(x.values().astype(np.uint8), y.storage[: 256 * 8].view(np.int64))
    tw = oracle.fast_transpose_abcd_dbca(pw)
    assert np.array_equal(out.raw()[: tw.size * 8].view(np.int64), tw.reshape(-1))
    # involution: transposing back restores the words
    back = nc.einsum("dbca -> abcd", out)
    assert np.array_equal(back.raw()[: pw.size * 8].view(np.int64), pw.reshape(-1))


def test_matrix_transpose_and_accumulate_into_supplied_output(nc, oracle):
    a = rand(oracle, (130, 70), "f32", SEED, 2, -1.0, 1.0)
    assert_bit_exact(nc.transpose(to_gpu(nc, a)), refsem.einsum(oracle, "ab -> ba", [a]))
    # a caller-supplied output is accumulated into, not cleared (3einsum.lisp:468)
    r0 = rand(oracle, (70, 130), "f32", SEED + 5, 0, -1.0, 1.0)
    want = oracle.OArr((70, 130), "f32", storage=r0.storage.copy())
    refsem.einsum(oracle, "ab -> ba", [a], [want])
    got = to_gpu(nc, r0)
    nc.einsum("ab -> ba", to_gpu(nc, a), got)
    assert_bit_exact(got, want)


# ---- generic reference-ordered kernel: bit-exact even for float contractions ----------------------------------------
SPECS = [("ij jk -> ik", [(7, 9), (9, 5)]), ("bij bjk -> bik", [(3, 4, 6), (3, 6, 5)]), ("ii -> i", [(6, 6)]), ("ii ->", [(6, 6)]),
         ("ij kl -> ikjl", [(2, 3), (4, 5)]), ("i i ->", [(17,), (17,)]), ("i j -> ij", [(5,), (7,)]), ("ijk -> ij ik jk", [(3, 4, 5)]),
         ("-i -i -> -i", [(3, 4), (2, 1, 4)]), ("- - -> (cl:* $1 $2) -> -", [(5,), (4, 1)]), ("ij -> (max @1 $1) -> i", [(6, 7)]),
         ("ij ij -> (+ @1 (* $1 $2) 1.5) -> i", [(4, 5), (4, 5)]), ("i -> (cl:+ @1 $1) -> -> ((i :start 2 :end 9 :step 3))", [(11,)]),
         ("kji -> ijk", [(3, 4, 5)]), ("ab bc ca ->", [(3, 4), (4, 5), (5, 3)])]


@pytest.mark.parametrize("spec,shapes", SPECS)
@pytest.mark.parametrize("dt", ["f32", "f64", "fixnum", "ub8"])
def test_generic_nest_bit_exact(nc, oracle, spec, shapes, dt):
    if dt in ("fixnum", "ub8") and "1.5" in spec:
        pytest.skip("float literal with integer arrays rounds into a fixnum output: covered by f32/f64")
    ins = [rand(oracle, s, dt, SEED + k, 0, *((-1.0, 1.0) if dt in ("f32", "f64") else (0, 9))) for k, s in enumerate(shapes)]
    want = refsem.einsum(oracle, spec, ins)
    nc.lib().ncl_force_generic(1)
    try:
        got = nc.einsum(spec, *[to_gpu(nc, a) for a in ins])
    finally:
        nc.lib().ncl_force_generic(0)
    wants = want if isinstance(want, list) else [want]
    gots = got if isinstance(got, tuple) else [got]
    for g, w in zip(gots, wants):
        if w.shape == ():
            wv = w.values().reshape(-1)[0]
            assert g == wv or (g != g and wv != wv)
        else:
            assert_bit_exact(g, w)


@pytest.mark.parametrize("spec,shapes", SPECS)
def test_dispatched_kernels_match_oracle(nc, oracle, spec, shapes):
    """default dispatch (fast kernel families where they apply) on exact inputs: bit-exact"""
    ins = [rand(oracle, s, "f32", SEED + k, 1, -4, 4) for k, s in enumerate(shapes)]
    want = refsem.einsum(oracle, spec, ins)
    got = nc.einsum(spec, *[to_gpu(nc, a) for a in ins])
    wants = want if isinstance(want, list) else [want]
    gots = got if isinstance(got, tuple) else [got]
    for g, w in zip(gots, wants):
        if w.shape == ():
            assert g == w.values().reshape(-1)[0]
        else:
            assert_bit_exact(g, w)


def test_unary_mappers(nc, oracle):
    a = rand(oracle, (37, 20), "f32", SEED, 0, 0.5, 4.0)
    g = to_gpu(nc, a)
    from numcl_b200 import numeric
    assert_bit_exact(numeric.sqrt(g), oracle.map_array("(sqrt $1)", [a], "f32"))           # IEEE sqrt: exact
    assert_bit_exact(numeric.square(g), oracle.map_array("(%square $1)", [a], "f32"))
    for name in ("exp", "log", "sin", "cos", "tanh"):                                       # libm: tolerance only
        got = getattr(numeric, name)(g).numpy()
        want = oracle.map_array(f"({name} $1)", [a], "f32").values()
        assert np.allclose(got, want, rtol=2e-6, atol=1e-7), name


def test_unsupported_is_loud(nc):
    with pytest.raises(NotImplementedError):
        nc.einsum("i -> (gamma $1) -> i", nc.zeros(4, type="f32"))
    with pytest.raises(nc.NumclUnsupportedError):
        nc.einsum("i- -> i-", nc.zeros((4, 3), type="f32"))


# ---- K4: dense contractions ---------------------------------------------------------------------------------------
def _frob(got, ref):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    return np.linalg.norm(got - ref) / np.linalg.norm(ref)


def _tf32_kernel(m, n, k):
    """which single-float GEMM kernel k_gemm_tf32.cu::gemm_f32_tf32x3 picks"""
    eligible = m % 128 == 0 and n % 256 == 0 and k % 32 == 0 and k <= 1024
    if eligible and m % 256 == 0 and m * n <= 2048 * 2048:
        return b"gemm_tf32x3_fused_2cta_tcgen05"
    if eligible and m * n <= 1024 * 1024:
        return b"gemm_tf32x3_fused_tcgen05"
    return b"gemm_tf32x3_tcgen05"


@pytest.mark.parametrize("dt,tol", [("f64", 1e-12), ("f32", 1e-5)])
@pytest.mark.parametrize("m,k,n", [(256, 192, 320), (130, 70, 50), (1024, 1024, 1024), (128, 4096, 128), (33, 2, 65), (129, 67, 131)])
def test_matmul_float_within_tolerance(nc, oracle, dt, tol, m, k, n):
    """config C3 at oracle-affordable sizes: relative Frobenius error vs numcl's i,j,k nest"""
    A, B = rand(oracle, (m, k), dt, SEED, 0, -1.0, 1.0), rand(oracle, (k, n), dt, SEED + 1, 0, -1.0, 1.0)
    ref = oracle.fast_matmul(A.values(), B.values())
    got = nc.matmul(to_gpu(nc, A), to_gpu(nc, B))
    assert got.dtype.name == dt and got.shape == (m, n)
    assert _frob(got.numpy(), ref) <= tol
    if dt == "f64" and m >= 32 and n >= 32:
        assert nc.lib().ncl_last_kernel() == b"gemm_f64_dmma"         # odd leading dimensions: 8-byte cp.async variant
    if dt == "f32" and m * n >= 128 * 128 and k >= 32:
        assert nc.lib().ncl_last_kernel() == _tf32_kernel(m, n, k)


def test_matmul_f64_interpreter_agrees_with_fast_loop(oracle):
    A, B = rand(oracle, (9, 7), "f64", SEED, 0, -1.0, 1.0), rand(oracle, (7, 5), "f64", SEED + 1, 0, -1.0, 1.0)
    assert np.array_equal(refsem.einsum(oracle, "ij jk -> ik", [A, B]).values(), oracle.fast_matmul(A.values(), B.values()))


@pytest.mark.parametrize("dt,tol", [("f64", 1e-12), ("f32", 1e-5)])
def test_batched_einsum_within_tolerance(nc, oracle, dt, tol):
    """config C4 shape family: (bij bjk -> bik)"""
    A, B = rand(oracle, (6, 160, 96), dt, SEED, 0, -1.0, 1.0), rand(oracle, (6, 96, 224), dt, SEED + 1, 0, -1.0, 1.0)
    if dt == "f32":
        ref = oracle.fast_bmm_f32(A.values(), B.values())
    else:
        ref = np.stack([oracle.fast_matmul(A.values()[i], B.values()[i]) for i in range(6)])
    got = nc.einsum("bij bjk -> bik", to_gpu(nc, A), to_gpu(nc, B))
    assert _frob(got.numpy(), ref) <= tol


@pytest.mark.parametrize("b,m,k,n", [(1, 128, 32, 256), (3, 256, 96, 512), (2, 384, 544, 256), (5, 128, 1024, 768),
                                     (1, 256, 32, 256), (7, 512, 288, 512), (150, 256, 64, 256), (2, 1024, 1024, 1024)])
def test_batched_einsum_fused_split_kernel(nc, oracle, b, m, k, n):
    """tile-aligned single-float shapes take the kernels that split hi/lo in shared memory (one launch, no
    pack pre-pass; on CTA pairs when M % 256 == 0); same tolerance vs numcl's nest as the packed path"""
    A, B = rand(oracle, (b, m, k), "f32", SEED, 0, -1.0, 1.0), rand(oracle, (b, k, n), "f32", SEED + 1, 0, -1.0, 1.0)
    ref = oracle.fast_bmm_f32(A.values(), B.values())
    got = nc.einsum("bij bjk -> bik", to_gpu(nc, A), to_gpu(nc, B))
    assert nc.lib().ncl_last_kernel() == _tf32_kernel(m, n, k) != b"gemm_tf32x3_tcgen05"
    assert _frob(got.numpy(), ref) <= 1e-5
    truth = A.values().astype(np.float64) @ B.values().astype(np.float64)
    assert _frob(got.numpy(), truth) <= 4e-6          # 3xTF32 + chunked accumulation: ~2e-6


@pytest.mark.parametrize("m", [128, 256])
def test_fused_split_kernel_specials_and_accumulate(nc, oracle, m):
    """inf / NaN operands stay non-finite (the split keeps them in the hi part; inf * lo-part may turn an
    IEEE inf into NaN, as in every split-precision GEMM -- SBCL would trap there, SURVEY 8b "Errors"), NaN
    stays NaN, and a caller-supplied result is accumulated into (src/3einsum.lisp:468)"""
    rng = np.random.default_rng(5)
    A = rng.uniform(-1, 1, (m, 64)).astype(np.float32); B = rng.uniform(-1, 1, (64, 256)).astype(np.float32)
    A[3, 5] = np.inf; A[7, 9] = np.nan; B[11, 17] = -np.inf
    A[9, 0] = np.float32(np.frombuffer(np.uint32(0x7fffffff).tobytes(), dtype=np.float32)[0])    # CUDA's canonical NaN
    got = nc.matmul(nc.asarray(A, type="f32"), nc.asarray(B, type="f32")).numpy()
    assert nc.lib().ncl_last_kernel() == _tf32_kernel(m, 256, 64)
    with np.errstate(invalid="ignore", over="ignore"):
        want = A.astype(np.float64) @ B.astype(np.float64)
    assert np.array_equal(np.isfinite(got), np.isfinite(want))
    assert np.all(np.isnan(got[np.isnan(want)]))
    fin = np.isfinite(want)
    assert np.allclose(got[fin], want[fin], rtol=0, atol=1e-4)
    C0 = rng.uniform(-1, 1, (m, 256)).astype(np.float32)
    A2 = rng.uniform(-1, 1, (m, 512)).astype(np.float32); B2 = rng.uniform(-1, 1, (512, 256)).astype(np.float32)
    c = nc.asarray(C0, type="f32")
    nc.matmul(nc.asarray(A2, type="f32"), nc.asarray(B2, type="f32"), c)
    assert nc.lib().ncl_last_kernel() == _tf32_kernel(m, 256, 512)
    assert _frob(c.numpy(), C0.astype(np.float64) + A2.astype(np.float64) @ B2.astype(np.float64)) <= 4e-6


def test_batched_einsum_c4_full_size(nc, oracle):
    """config C4 at full size (256 x 512^3): three batches checked against numcl's nest, all against linearity"""
    bsz = 256
    a = nc.random_array((bsz, 512, 512), "f32", SEED, 0, -1.0, 1.0)
    b = nc.random_array((bsz, 512, 512), "f32", SEED + 1, 0, -1.0, 1.0)
    r = nc.einsum("bij bjk -> bik", a, b)
    assert nc.lib().ncl_last_kernel() == b"gemm_tf32x3_fused_2cta_tcgen05"
    got = r.numpy()
    A, B = a.numpy(), b.numpy()
    for i in (0, 131, 255):
        ref = oracle.fast_bmm_f32(A[i:i + 1], B[i:i + 1])
        assert _frob(got[i:i + 1], ref) <= 1e-5
    # row sums of every product: sum_k (sum_j A[b,i,j] B[b,j,k]) = A[b,i,:] . rowsum(B[b])
    want = np.einsum("bij,bj->bi", A.astype(np.float64), B.astype(np.float64).sum(axis=2))
    assert _frob(got.astype(np.float64).sum(axis=2), want) <= 1e-5


@pytest.mark.parametrize("ta,tb", [("fixnum", "fixnum"), ("ub8", "ub8"), ("ub8", "fixnum"), ("sb16", "ub4")])
def test_integer_matmul_bit_exact(nc, oracle, ta, tb):
    A, B = rand(oracle, (37, 29), ta, SEED, 0, 0, 9), rand(oracle, (29, 41), tb, SEED + 1, 0, 0, 7)
    want = refsem.einsum(oracle, "ij jk -> ik", [A, B])
    got = nc.matmul(to_gpu(nc, A), to_gpu(nc, B))
    assert want.dtype == "fixnum"
    assert_bit_exact(got, want)
    # operands of different integer types (a packed sub-byte one included) are brought to fixnum first (ncl_plan.cu:
    # harmonize_contraction), then the integer GEMM runs
    assert nc.lib().ncl_last_kernel() == b"gemm_int_simt"


def test_integer_matmul_wraps_like_coerce(nc, oracle):
    A = rand(oracle, (20, 16), "fixnum", SEED, 0, 2 ** 60, 2 ** 61)
    B = rand(oracle, (16, 24), "fixnum", SEED + 1, 0, -5, 5)
    assert_bit_exact(nc.matmul(to_gpu(nc, A), to_gpu(nc, B)), refsem.einsum(oracle, "ij jk -> ik", [A, B]))


def test_matmul_accumulates_into_supplied_result(nc, oracle):
    A, B = rand(oracle, (96, 64), "f64", SEED, 1, -4, 4), rand(oracle, (64, 160), "f64", SEED + 1, 1, -4, 4)
    R = rand(oracle, (96, 160), "f64", SEED + 2, 1, -4, 4)
    want = oracle.OArr((96, 160), "f64", storage=R.storage.copy())
    refsem.einsum(oracle, "ij jk -> ik", [A, B], [want])
    got = to_gpu(nc, R)
    nc.matmul(to_gpu(nc, A), to_gpu(nc, B), got)
    assert_bit_exact(got, want)                     # small integers: every order is exact


def test_matmul_c3_sampled_rows(nc, oracle):
    """config C3 at N=2048 (f64): 64 sampled rows in the reference's own accumulation order"""
    n = 2048
    A, B = rand(oracle, (n, n), "f64", SEED, 0, -1.0, 1.0), rand(oracle, (n, n), "f64", SEED + 1, 0, -1.0, 1.0)
    got = nc.matmul(to_gpu(nc, A), to_gpu(nc, B)).numpy()
    for r0 in (0, 1000, 2040):
        ref = oracle.fast_matmul(A.values(), B.values(), rows=(r0, min(r0 + 22, n)))
        assert _frob(got[r0:r0 + ref.shape[0]], ref) <= 1e-12


def test_matmul_c3_f32_full_size_sampled_rows(nc, oracle):
    """config C3 single-float at full size (8192^3, pack + 3xTF32 tcgen05 GEMM): 3 x 8 sampled rows in the
    reference's own accumulation order, and the whole product against a linear property"""
    n = 8192
    a = nc.random_array((n, n), "f32", SEED, 0, -1.0, 1.0)
    b = nc.random_array((n, n), "f32", SEED + 1, 0, -1.0, 1.0)
    r = nc.matmul(a, b)
    assert nc.lib().ncl_last_kernel() == b"gemm_tf32x3_tcgen05"
    A, B, got = a.numpy(), b.numpy(), r.numpy()
    for r0 in (0, 4099, 8184):
        ref = oracle.fast_matmul(A, B, rows=(r0, r0 + 8))
        assert _frob(got[r0:r0 + 8], ref) <= 1e-5
    # C @ 1 = A @ (B @ 1): every row of the product is covered
    want = A.astype(np.float64) @ B.astype(np.float64).sum(axis=1)
    assert _frob(got.astype(np.float64).sum(axis=1), want) <= 1e-5


def test_matmul_c3_f64_full_size_sampled_rows(nc, oracle):
    """config C3 double-float at full size (8192^3, DMMA): sampled rows in the reference order (1e-12) + linearity"""
    n = 8192
    a = nc.random_array((n, n), "f64", SEED, 0, -1.0, 1.0)
    b = nc.random_array((n, n), "f64", SEED + 1, 0, -1.0, 1.0)
    r = nc.matmul(a, b)
    assert nc.lib().ncl_last_kernel() == b"gemm_f64_dmma"
    A, B, got = a.numpy(), b.numpy(), r.numpy()
    for r0 in (0, 4099, 8188):
        ref = oracle.fast_matmul(A, B, rows=(r0, r0 + 4))
        assert _frob(got[r0:r0 + 4], ref) <= 1e-12
    want = A @ B.sum(axis=1)
    assert _frob(got.sum(axis=1), want) <= 1e-12
