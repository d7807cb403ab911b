"""Edge cases the reference's tests touch or imply: empty and rank-0 arrays, ragged / odd shapes that
leave the vector paths, displaced views, multi-axis reductions, aliasing, host-resident operands."""
import numpy as np
import pytest

import refsem
from test_gpu_parity import SEED, assert_bit_exact, rand, to_gpu

pytestmark = pytest.mark.gpu


def test_empty_arrays(nc, oracle):
    # a size-0 axis against a size-1 axis: the reference's result shape is (mapcar #'max ...) = 1 and its
    # loop writes nothing (4numeric.lisp:187-199,263-273) -- an array of uninitialised memory.  Fenced.
    with pytest.raises(AssertionError):
        nc.add(nc.zeros((0, 5), type="f32"), nc.ones((5,), type="f32"))
    e = nc.add(nc.zeros((0, 5), type="f32"), nc.zeros((0, 5), type="f32"))
    assert e.shape == (0, 5) and e.numpy().size == 0
    z = nc.asarray(np.zeros((3, 0), dtype=np.float32), type="f32")
    assert nc.sum(z, axes=1).numpy().tolist() == [0.0, 0.0, 0.0]          # (full kept-shape 0.0), nothing to add
    assert nc.sum(z) == 0.0
    assert nc.amax(z, axes=1).numpy().tolist() == [np.float32(-3.4028234663852886e38)] * 3
    A = nc.asarray(np.zeros((4, 0), dtype=np.float64), type="f64")
    B = nc.asarray(np.zeros((0, 6), dtype=np.float64), type="f64")
    assert np.array_equal(nc.matmul(A, B).numpy(), np.zeros((4, 6)))        # zeros output, zero contraction steps
    t = nc.einsum("ij -> ji", nc.asarray(np.zeros((0, 7), dtype=np.int64), type="fixnum"))
    assert t.shape == (7, 0)


def test_rank0_arrays_and_scalars(nc, oracle):
    x = nc.full((), 3.5, type="f32")
    y = nc.full((), 2, type="ub2")
    r = nc.mul(x, y)
    assert r.shape == () and r.item() == 7.0
    assert nc.add(2, 3) == 5                                  # scalars only: plain CL arithmetic (4numeric.lisp:205-206)
    assert nc.einsum("i i ->", nc.asarray(np.arange(5), type="fixnum"), nc.asarray(np.arange(5), type="fixnum")) == 30


@pytest.mark.parametrize("shape_a,shape_b", [((3, 5, 7), (7,)), ((3, 5, 7), (5, 1)), ((1, 9), (9, 1)), ((2, 3, 1, 5), (4, 1)), ((6, 1, 11), (1, 13, 1)),
                                              ((1023,), (1023,)), ((17, 1001), (1001,)), ((5, 3, 2, 2, 3, 2), (3, 2))])
@pytest.mark.parametrize("dt", ["f32", "f64", "fixnum", "ub8", "sb16"])
def test_ragged_broadcast_shapes_bit_exact(nc, oracle, shape_a, shape_b, dt):
    lo, hi = (-3.0, 3.0) if dt in ("f32", "f64") else (0, 100)
    a = rand(oracle, shape_a, dt, SEED, 2 if dt in ("f32", "f64") else 0, lo, hi)
    b = rand(oracle, shape_b, dt, SEED + 1, 0, lo, hi)
    for op, f in (("+", nc.add), ("*", nc.mul), ("max", nc.maximum), ("-", nc.sub)):
        assert_bit_exact(f(to_gpu(nc, a), to_gpu(nc, b)), refsem.arith(oracle, op, [a, b]))


def test_displaced_views_in_broadcast_and_reduce(nc, oracle):
    """broadcast and reduce-array honour the displacement offset (array-displacement's second value,
    4numeric.lisp:225-227); einsum does not (3einsum.lisp:485-487), so the host mirror refuses it."""
    base = rand(oracle, (40,), "f32", SEED, 1, -8, 8)
    v = base.values()
    g = to_gpu(nc, base)
    view = nc.NArray((5, 4), "f32", _storage=g._buf, offset=7)
    view._base = g
    oview = oracle.OArr((5, 4), "f32", storage=base.storage, offset=7)
    assert np.array_equal(view.numpy(), v[7:27].reshape(5, 4))
    b = rand(oracle, (4,), "f32", SEED + 1, 1, -8, 8)
    assert_bit_exact(nc.add(view, to_gpu(nc, b)), refsem.arith(oracle, "+", [oview, b]))
    assert_bit_exact(nc.sum(view, axes=1), refsem.reduce_array(oracle, "+", oview, (1,)))
    assert nc.sum(view) == float(v[7:27].sum())
    with pytest.raises(AssertionError):
        nc.einsum("ij -> ji", view)


@pytest.mark.parametrize("shape,axes", [((4, 5, 6, 7), (1, 3)), ((4, 5, 6, 7), (0, 2)), ((4, 5, 6, 7), (0, 1, 2)), ((4, 5, 6, 7), (1, 2, 3)),
                                        ((3, 1, 5), (1,)), ((2, 3, 4, 5, 6), (0, 2, 4)), ((129, 33), (0,)), ((1, 4096), (1,)), ((4099,), None)])
@pytest.mark.parametrize("dt", ["f64", "sb32", "f32"])
def test_multi_axis_reductions_bit_exact(nc, oracle, shape, axes, dt):
    a = rand(oracle, shape, dt, SEED, 1, -8, 8)
    for fn, f in (("+", nc.sum), ("max", nc.amax), ("min", nc.amin), ("*", nc.prod)):
        src = a if fn != "*" else rand(oracle, shape, dt, SEED, 1, 1, 2)          # products stay exact
        want = refsem.reduce_array(oracle, fn, src, axes)
        got = f(to_gpu(nc, src), axes=axes)
        if want.shape == ():
            assert got == want.values().reshape(-1)[0]
        else:
            assert_bit_exact(got, want)


def test_map_array_forms_and_aliasing(nc, oracle):
    a = rand(oracle, (33, 7), "f32", SEED, 2, -2.0, 2.0)
    b = rand(oracle, (33, 7), "f32", SEED + 1, 0, -2.0, 2.0)
    c = rand(oracle, (33, 7), "ub8", SEED + 2, 0, 0, 9)
    form = "(max (* $1 $2) (+ $3 0.5))"
    assert_bit_exact(nc.map_array(form, to_gpu(nc, a), to_gpu(nc, b), to_gpu(nc, c)), oracle.map_array(form, [a, b, c], "f32"))
    # in place: (map-array-into a 'cl:+ a b)
    ga = to_gpu(nc, a)
    nc.map_array_into(ga, "cl:+", ga, to_gpu(nc, b))
    assert_bit_exact(ga, oracle.map_array("(+ $1 $2)", [a, b], "f32"))
    with pytest.raises(AssertionError):
        nc.map_array("cl:+", to_gpu(nc, a), nc.zeros((7, 33), type="f32"))


def test_float_into_integer_output_rounds_half_even(nc, oracle):
    """%coerce: (round x) then wrap (1type.lisp:673-694)"""
    v = np.array([0.5, 1.5, 2.5, -0.5, -1.5, 127.5, 128.5, 300.7, -129.5], dtype=np.float32)
    a = oracle.OArr.from_values(v, "f32")
    for dt in ("sb8", "ub8", "fixnum", "ub4"):
        r = nc.NArray(v.shape, dt)
        nc.map_array_into(r, "identity", to_gpu(nc, a))
        assert_bit_exact(r, oracle.map_array("(identity $1)", [a], dt))


def test_host_resident_operands_match_device(nc, oracle):
    A, B = rand(oracle, (200, 96), "f32", SEED, 1, -4, 4), rand(oracle, (96, 136), "f32", SEED + 1, 1, -4, 4)
    dev = nc.matmul(to_gpu(nc, A), to_gpu(nc, B))
    host = nc.matmul(to_gpu(nc, A).to("host"), to_gpu(nc, B).to("host"))
    assert host.where == "host" and np.array_equal(host.numpy(), dev.numpy())
    x = rand(oracle, (300, 77), "f64", SEED, 1, -8, 8)
    assert np.array_equal(nc.sum(to_gpu(nc, x).to("host"), axes=0).numpy(), nc.sum(to_gpu(nc, x), axes=0).numpy())


def test_int_division_correctly_rounded(nc, oracle):
    """(/ int int) is an exact ratio in CL, rounded ONCE into single-float"""
    n = rand(oracle, (4096,), "fixnum", SEED, 0, -(2 ** 40), 2 ** 40)
    d = rand(oracle, (4096,), "fixnum", SEED + 1, 0, 1, 2 ** 30)
    assert_bit_exact(nc.div(to_gpu(nc, n), to_gpu(nc, d)), refsem.arith(oracle, "/", [n, d]))
    big_n = rand(oracle, (512,), "fixnum", SEED + 2, 0, 2 ** 58, 2 ** 61)       # beyond 2^53: the long-division path
    big_d = rand(oracle, (512,), "fixnum", SEED + 3, 0, 3, 2 ** 57)
    assert_bit_exact(nc.div(to_gpu(nc, big_n), to_gpu(nc, big_d)), refsem.arith(oracle, "/", [big_n, big_d]))


def test_errors_mirror_reference_assertions(nc):
    a = nc.zeros((3, 4), type="f32")
    with pytest.raises(AssertionError):
        nc.einsum("ij jk -> ik", a, nc.zeros((5, 6), type="f32"))          # (assert (= ...)) in shape-resolver
    with pytest.raises(nc.EinsumSpecError):
        nc.einsum("i0 -> i", a)
    with pytest.raises(TypeError):
        nc.einsum("ij jk -> ik", a)
    with pytest.raises(nc.NumclUnsupportedError):
        nc.reduce_array("logior", nc.asarray(np.arange(4), type="fixnum"))


def test_matmul_star_copy_astype(nc, oracle):
    mats = [rand(oracle, s, "fixnum", SEED + k, 0, 0, 5) for k, s in enumerate([(10, 30), (30, 5), (5, 60), (60, 4)])]
    want = mats[0].values()
    for m in mats[1:]:
        want = want @ m.values()
    got = nc.matmul_star(*[to_gpu(nc, m) for m in mats])
    assert np.array_equal(got.numpy(), want)
    a = rand(oracle, (50, 7), "f32", SEED, 0, -300.0, 300.0)
    g = to_gpu(nc, a)
    assert_bit_exact(nc.copy(g), a)
    for dt in ("fixnum", "sb8", "f64", "ub8"):
        assert_bit_exact(nc.astype(g, dt), oracle.map_array("(identity $1)", [a], dt))


@pytest.mark.gpu
@pytest.mark.parametrize("dt,value", [("ub8", 200), ("sb16", -1234), ("ub31", 123456789), ("fixnum", -(2 ** 40) - 3), ("sb64", 2 ** 62 + 5),
                                      ("f32", 1.5), ("f64", -2.25)])
@pytest.mark.parametrize("n,offset", [(4099, 0), (4099, 1), (100003, 3), (65536, 7), (513, 5)])
def test_full_vectorised_fill_with_displacement(nc, dt, value, n, offset):
    """(full shape value), src/2zeros.lisp:46-50, on a displaced array: the 128-bit body, the scalar head and
    tail give the same slots as the scalar kernel, and nothing outside [offset, offset + n) is written"""
    import ctypes as C
    from numcl_b200 import _lib
    from numcl_b200.array import NArray
    a = NArray((n,), dt, offset=offset)
    a.set_raw(np.full(a.nbytes, 0xAB, dtype=np.uint8))
    t = a.tensor()
    isf = isinstance(value, float)
    _lib.check(_lib.lib().ncl_fill(C.byref(t), 1 if isf else 0, 0 if isf else int(value), float(value)))
    got = a.numpy()
    assert np.all(got == (np.float64(value) if isf else np.int64(value)))
    raw = a.raw()
    sb = a.dtype.slot_bits // 8
    assert np.all(raw[: offset * sb] == 0xAB) and np.all(raw[(offset + n) * sb:] == 0xAB)
    if n * sb >= 4096:
        assert nc.lib().ncl_last_kernel() == b"fill_vec"


@pytest.mark.gpu
@pytest.mark.parametrize("m,k,n,kernel", [(256, 64, 256, b"gemm_tf32x3_fused_2cta_tcgen05"), (128, 96, 512, b"gemm_tf32x3_fused_tcgen05"),
                                          (200, 96, 300, b"gemm_tf32x3_tcgen05"), (40, 24, 56, b"gemm_fma_simt")])
def test_gemm_writes_only_its_output_region(nc, m, k, n, kernel):
    """C-ABI level: the result tensor is a displaced window (elem_offset 16) into a larger base vector filled
    with a canary; every single-float GEMM kernel must leave the bytes before and after the window alone"""
    import ctypes as C
    from numcl_b200 import _lib
    from numcl_b200.array import NArray
    from numcl_b200.einsum import _memoized
    rng = np.random.default_rng(11)
    A = rng.uniform(-1, 1, (m, k)).astype(np.float32); B = rng.uniform(-1, 1, (k, n)).astype(np.float32)
    a, b = nc.asarray(A, type="f32"), nc.asarray(B, type="f32")
    off, tail = 16, 40
    store = NArray((off + m * n + tail,), "f32")
    store.set_raw(np.full(store.nbytes, 0xAB, dtype=np.uint8))
    t = store.tensor()
    t.rank = 2; t.dims[0] = m; t.dims[1] = n; t.elem_offset = off
    _, desc = _memoized("ij jk -> ik")
    tin = (_lib.Tensor * 2)(a.tensor(), b.tensor()); tout = (_lib.Tensor * 1)(t)
    _lib.check(_lib.lib().ncl_einsum(C.byref(desc), tin, 2, tout, 1, _lib.OUT_FRESH))
    assert nc.lib().ncl_last_kernel() == kernel
    raw = store.raw()
    assert np.all(raw[: off * 4] == 0xAB) and np.all(raw[(off + m * n) * 4: (off + m * n + tail) * 4] == 0xAB)
    got = raw[off * 4: (off + m * n) * 4].view(np.float32).reshape(m, n)
    ref = A.astype(np.float64) @ B.astype(np.float64)
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) <= 1e-5


@pytest.mark.gpu
def test_widening_kernel_writes_only_its_output_region(nc, oracle):
    """same canary check for the widening integer family and a ragged element count"""
    import ctypes as C
    from numcl_b200 import _lib
    from numcl_b200.array import NArray
    n = 128 * 2 * 37 + 6                       # not a multiple of the 256-element warp item
    x = rand(oracle, (n,), "ub8", SEED, 0, 0, 255); y = rand(oracle, (n,), "fixnum", SEED + 1, 0, -(2 ** 40), 2 ** 40)
    want = refsem.arith(oracle, "*", [x, y])
    off, tail = 2, 10
    store = NArray((off + n + tail,), "fixnum")
    store.set_raw(np.full(store.nbytes, 0xAB, dtype=np.uint8))
    t = store.tensor(); t.rank = 1; t.dims[0] = n; t.elem_offset = off
    gx, gy = to_gpu(nc, x), to_gpu(nc, y)
    tx, ty = gx.tensor(), gy.tensor()
    _lib.check(_lib.lib().ncl_broadcast(_lib.OPS["MUL"], C.byref(tx), C.byref(ty), C.byref(t)))
    assert nc.lib().ncl_last_kernel() == b"bcast_widen"
    raw = store.raw()
    assert np.all(raw[: off * 8] == 0xAB) and np.all(raw[(off + n) * 8: (off + n + tail) * 8] == 0xAB)
    assert np.array_equal(raw[off * 8: (off + n) * 8].view(np.int64), want.storage[: n * 8].view(np.int64))


@pytest.mark.gpu
@pytest.mark.parametrize("dt", ["f32", "f64", "fixnum"])
def test_matmul_with_empty_contraction(nc, oracle, dt):
    """(ij jk -> ik) with an empty j: the reference's loops never run and the fresh `zeros` output stays zero;
    tile-aligned shapes must not reach the GEMM kernels with K = 0"""
    a = nc.empty((256, 0), dt); b = nc.empty((0, 256), dt)
    r = nc.matmul(a, b)
    assert r.shape == (256, 256) and np.all(r.numpy() == 0)
    c = nc.full((256, 256), 3, type=dt)
    nc.matmul(a, b, c)
    assert np.all(c.numpy() == 3)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(1500, 48), (48, 1500), (1027, 127), (127, 1027), (4096, 9), (9, 4096)])
@pytest.mark.parametrize("dt,sb", [("f32", 4), ("f64", 8)])
def test_strip_transpose_writes_only_its_output_region(nc, shape, dt, sb):
    """the strip transpose (cp.async into a padded tile, partial last strip) behind a canary on both sides, from an input
    that is itself a displaced window: nothing outside the output window may change"""
    import ctypes as C
    from numcl_b200 import _lib
    from numcl_b200.array import NArray
    from numcl_b200.einsum import _memoized
    rng = np.random.default_rng(12)
    A = rng.integers(-8, 9, shape).astype(np.float32 if dt == "f32" else np.float64)
    n = A.size
    src = NArray((5 + n + 3,), dt)
    buf = np.full(src.nbytes, 0xCD, dtype=np.uint8); buf[5 * sb:(5 + n) * sb] = A.view(np.uint8).reshape(-1)
    src.set_raw(buf)
    ts = src.tensor(); ts.rank = 2; ts.dims[0], ts.dims[1] = shape; ts.elem_offset = 5
    off, tail = 7, 29
    store = NArray((off + n + tail,), dt)
    store.set_raw(np.full(store.nbytes, 0xAB, dtype=np.uint8))
    t = store.tensor(); t.rank = 2; t.dims[0], t.dims[1] = shape[1], shape[0]; t.elem_offset = off
    _, desc = _memoized("ij -> ji")
    tin = (_lib.Tensor * 1)(ts); tout = (_lib.Tensor * 1)(t)
    _lib.check(_lib.lib().ncl_einsum(C.byref(desc), tin, 1, tout, 1, _lib.OUT_FRESH))
    assert nc.lib().ncl_last_kernel() == b"transpose_strip"
    raw = store.raw()
    assert np.all(raw[: off * sb] == 0xAB) and np.all(raw[(off + n) * sb: (off + n + tail) * sb] == 0xAB)
    got = raw[off * sb: (off + n) * sb].view(A.dtype).reshape(shape[1], shape[0])
    assert np.array_equal(got, A.T)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(700, 515), (64, 259), (2050, 1001)])
@pytest.mark.parametrize("dt,sb", [("f32", 4), ("f64", 8)])
def test_column_reduce_of_ragged_rows_writes_only_its_output_region(nc, shape, dt, sb):
    """k_reduce_rot.cu reads whole aligned chunks (a few elements of the neighbouring rows travel along) but must store
    nothing outside the result's window"""
    import ctypes as C
    from numcl_b200 import _lib
    from numcl_b200.array import NArray
    rng = np.random.default_rng(13)
    A = rng.integers(-8, 9, shape).astype(np.float32 if dt == "f32" else np.float64)
    a = nc.asarray(A, type=dt)
    off, tail, ncols = 3, 21, shape[1]
    store = NArray((off + ncols + tail,), dt)
    store.set_raw(np.full(store.nbytes, 0xAB, dtype=np.uint8))
    t = store.tensor(); t.rank = 1; t.dims[0] = ncols; t.elem_offset = off
    ta = a.tensor(); ax = (C.c_int * 1)(0)
    _lib.check(_lib.lib().ncl_reduce(_lib.OPS["ADD"], C.byref(ta), ax, 1, C.byref(t), _lib.OUT_FRESH))
    assert nc.lib().ncl_last_kernel() == b"reduce_col_rot"
    raw = store.raw()
    assert np.all(raw[: off * sb] == 0xAB) and np.all(raw[(off + ncols) * sb: (off + ncols + tail) * sb] == 0xAB)
    assert np.array_equal(raw[off * sb: (off + ncols) * sb].view(A.dtype), A.sum(axis=0))
    # accumulate into the same window: r <- r + column sums
    _lib.check(_lib.lib().ncl_reduce(_lib.OPS["ADD"], C.byref(ta), ax, 1, C.byref(t), 0))
    raw = store.raw()
    assert np.all(raw[: off * sb] == 0xAB) and np.array_equal(raw[off * sb: (off + ncols) * sb].view(A.dtype), 2 * A.sum(axis=0))
