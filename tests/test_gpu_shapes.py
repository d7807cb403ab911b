"""Everyday shapes off the BASELINE configs: a short trailing axis ((N,3) point arrays), odd row lengths, one value per
row, reductions over short / ragged axes, tall-and-thin transposes.  The planner has dedicated paths for them
(element-granular flat mode, small-row / few-column reductions, thin transposes); whatever it picks is compared with
the oracle bit for bit (integers, and floats holding small integers so that every summation order is exact), and
random floats within the north-star tolerance for the reductions."""
import numpy as np
import pytest

import refsem
from test_gpu_parity import SEED, assert_bit_exact, rand, to_gpu

pytestmark = pytest.mark.gpu


def _last(nc):
    return nc.lib().ncl_last_kernel().decode()


# ---- broadcast elementwise -------------------------------------------------------------------------------------------
@pytest.mark.parametrize("sa,sb", [((40000, 3), (3,)), ((40000, 3), (40000, 1)), ((1001, 3), (3,)), ((333, 7), (7,)), ((333, 7), (333, 1)),
                                   ((50, 20, 5), (5,)), ((50, 20, 5), (20, 1)), ((129, 1001), (1001,)), ((129, 1001), (129, 1)),
                                   ((3,), (40000, 3)), ((4097, 6), (1, 6)), ((4097, 6), ())])
@pytest.mark.parametrize("ta,tb,op", [("f32", "f32", "+"), ("f32", "f32", "/"), ("f64", "f32", "*"), ("ub8", "fixnum", "*"), ("ub8", "ub8", "+"),
                                      ("sb16", "ub8", "-"), ("fixnum", "fixnum", "max"), ("f32", "ub8", "min")])
def test_short_and_odd_rows_bit_exact(nc, oracle, sa, sb, ta, tb, op):
    x = rand(oracle, sa, ta, SEED + 1, 2 if ta == "f32" and op != "/" else 0, *((-1.0, 1.0) if ta in ("f32", "f64") else (0, 100)))
    y = rand(oracle, sb, tb, SEED + 2, 0, *((0.5, 2.0) if tb in ("f32", "f64") else (1, 100)))
    f = {"+": nc.add, "-": nc.sub, "*": nc.mul, "/": nc.div, "max": nc.maximum, "min": nc.minimum}[op]
    assert_bit_exact(f(to_gpu(nc, x), to_gpu(nc, y)), refsem.arith(oracle, op, [x, y]))


def test_short_rows_take_the_flat_vector_path(nc, oracle):
    x = rand(oracle, (40000, 3), "f32", SEED, 0, -1.0, 1.0)
    y = rand(oracle, (3,), "f32", SEED + 1, 0, -1.0, 1.0)
    nc.add(to_gpu(nc, x), to_gpu(nc, y))
    assert _last(nc) == "bcast_vec_flat"
    # total not a multiple of the chunk: vector body + the tail nests, and nothing past the end is written
    x = rand(oracle, (1001, 3), "f32", SEED, 0, -1.0, 1.0)
    g = nc.add(to_gpu(nc, x), to_gpu(nc, y))
    assert_bit_exact(g, refsem.arith(oracle, "+", [x, y]))
    assert not g.raw()[3003 * 4:].any()


# ---- reductions over short and ragged axes --------------------------------------------------------------------------
RSHAPES = [((50000, 3), (1,)), ((50000, 3), (0,)), ((3, 50000), (0,)), ((3, 50000), (1,)), ((20000, 48), (1,)), ((20000, 48), (0,)),
           ((9000, 100), (1,)), ((9000, 100), (0,)), ((1000, 33), (1,)), ((1000, 33), (0,)), ((257, 1001), (1,)), ((257, 1001), (0,)),
           ((300, 7, 5), (2,)), ((300, 7, 5), (0,)), ((300, 7, 5), (0, 2)), ((300, 7, 5), (0, 1)), ((64, 100, 36), (0, 2)),
           ((70001, 2), (1,)), ((70001, 2), (0,)), ((12345, 17), (0,)), ((12345, 17), (1,))]


@pytest.mark.parametrize("shape,axes", RSHAPES)
@pytest.mark.parametrize("fn", ["+", "max", "min"])
@pytest.mark.parametrize("dt", ["ub8", "fixnum", "f32", "f64", "sb16"])
def test_reduce_short_axes_bit_exact(nc, oracle, shape, axes, fn, dt):
    a = rand(oracle, shape, dt, SEED + 3, 1, 0 if dt == "ub8" else -8, 8)
    want = refsem.reduce_array(oracle, fn, a, axes)
    got = {"+": nc.sum, "max": nc.amax, "min": nc.amin}[fn](to_gpu(nc, a), axes=axes)
    assert_bit_exact(got, want)


# ---- column reduce of a matrix whose rows are not whole 16-byte chunks (k_reduce_rot.cu) -------------------------------------
@pytest.mark.parametrize("shape", [(1000, 257), (4099, 301), (333, 1001), (64, 259), (2050, 8193), (129, 4097), (97, 515)])
@pytest.mark.parametrize("fn", ["+", "max", "min", "*"])
@pytest.mark.parametrize("dt", ["f32", "f64", "fixnum", "sb32", "ub31", "sb64"])
def test_column_reduce_with_ragged_row_pitch_bit_exact(nc, oracle, shape, fn, dt):
    if fn == "*":
        a = rand(oracle, shape, dt, SEED + 31, 1, 1, 1) if dt in ("f32", "f64") else rand(oracle, shape, dt, SEED + 31, 0, 1, 1)
        v = a.values().copy(); v[::7, ::5] = 2 if shape[0] < 400 else 1; v[3::11, 1::3] = -1 if dt not in ("ub31",) else 1     # products stay exact
        a = oracle.OArr.from_values(v, dt)
    else:
        a = rand(oracle, shape, dt, SEED + 31, 1 if dt in ("f32", "f64") else 0, 0 if dt == "ub31" else -8, 8)
    want = refsem.reduce_array(oracle, fn, a, (0,))
    got = {"+": nc.sum, "max": nc.amax, "min": nc.amin, "*": nc.prod}[fn](to_gpu(nc, a), axes=0)
    assert _last(nc) == "reduce_col_rot", _last(nc)
    assert_bit_exact(got, want)


@pytest.mark.parametrize("dt", ["f32", "f64", "sb32"])
@pytest.mark.parametrize("offset", [1, 2, 3, 5])
def test_column_reduce_of_a_misaligned_view(nc, oracle, dt, offset):
    """a displaced view that starts in the middle of a chunk: the classes' misalignment follows the element offset"""
    rows, cols = 700, 515
    base = rand(oracle, (rows * cols + 8,), dt, SEED + 32, 1 if dt != "sb32" else 0, -8, 8)
    g = to_gpu(nc, base)
    view = nc.NArray((rows, cols), dt, _storage=g._buf, offset=offset)
    got = nc.sum(view, axes=0)
    assert _last(nc) == "reduce_col_rot"
    v = base.values()[offset:offset + rows * cols].reshape(rows, cols)
    assert np.array_equal(got.numpy(), v.sum(axis=0).astype(got.numpy().dtype))
    m = nc.mean(view, axes=0) if dt != "sb32" else None
    if m is not None:
        assert np.allclose(m.numpy(), v.mean(axis=0), rtol=1e-6)


@pytest.mark.parametrize("dt,tol", [("f32", 1e-5), ("f64", 1e-12)])
@pytest.mark.parametrize("shape,axes", [((50000, 3), (1,)), ((50000, 3), (0,)), ((20000, 48), (1,)), ((20000, 48), (0,)), ((257, 1001), (1,)),
                                        ((257, 1001), (0,)), ((64, 100, 36), (0, 2))])
def test_reduce_short_axes_random_floats(nc, oracle, dt, tol, shape, axes):
    a = rand(oracle, shape, dt, SEED + 4, 0, -1.0, 1.0)
    want = refsem.reduce_array(oracle, "+", a, axes)
    got = nc.sum(to_gpu(nc, a), axes=axes)
    truth = a.values().astype(np.float64).sum(axis=axes)
    g = got.numpy().astype(np.float64)
    w = want.values().astype(np.float64).reshape(g.shape)
    assert np.linalg.norm(g - w) <= tol * np.linalg.norm(w)
    assert np.linalg.norm(g - truth) <= max(np.linalg.norm(w - truth) * 1.5, tol * np.linalg.norm(truth))


def test_reduce_into_supplied_output_accumulates(nc, oracle):
    """reduce-array folds into what the result already holds when it is not fresh (einsum with a caller-supplied output)"""
    a = rand(oracle, (30000, 3), "f32", SEED + 5, 1, -8, 8)
    for spec, oshape in (("ij -> i", (30000,)), ("ij -> j", (3,))):
        o = rand(oracle, oshape, "f32", SEED + 6, 1, -8, 8)
        go = to_gpu(nc, o)                                   # before the oracle accumulates into o in place
        want = refsem.einsum(oracle, spec, [a], [o])
        nc.einsum(spec, to_gpu(nc, a), go)
        assert_bit_exact(go, want if not isinstance(want, (list, tuple)) else want[0])


# ---- tall-and-thin transposes ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(50000, 3), (3, 50000), (20000, 48), (48, 20000), (4099, 7), (7, 4099), (1000, 33), (70001, 2), (2, 70001),
                                   (40000, 5), (5, 40000), (40000, 8), (8, 40000), (2, 40000), (40000, 2), (40004, 6), (6, 40004), (7, 40000), (4, 4096), (4096, 4)])
@pytest.mark.parametrize("dt", ["f32", "f64", "fixnum", "ub8", "sb16"])
def test_thin_transpose_bit_exact(nc, oracle, shape, dt):
    a = rand(oracle, shape, dt, SEED + 7, 1 if dt in ("f32", "f64") else 0, *((-8, 8) if dt != "ub8" else (0, 255)))
    assert_bit_exact(nc.transpose(to_gpu(nc, a)), refsem.einsum(oracle, "ij -> ji", [a]))


# ---- strips: a short side of up to 128 whose rows lie back to back (transpose_strip_kernel) ------------------------------------
@pytest.mark.parametrize("shape", [(4096, 9), (9, 4096), (5000, 12), (12, 5000), (3001, 16), (16, 3001), (2049, 33), (33, 2049), (20000, 48), (48, 20000),
                                   (1500, 64), (64, 1500), (1024, 100), (100, 1024), (1025, 128), (128, 1025), (1027, 127), (127, 1027)])
@pytest.mark.parametrize("dt", ["f32", "f64", "fixnum"])
def test_strip_transpose_bit_exact(nc, oracle, shape, dt):
    a = rand(oracle, shape, dt, SEED + 8, 1 if dt in ("f32", "f64") else 0, -8, 8)
    got = nc.transpose(to_gpu(nc, a))
    assert _last(nc) == "transpose_strip", _last(nc)
    assert_bit_exact(got, refsem.einsum(oracle, "ij -> ji", [a]))


@pytest.mark.parametrize("spec,shape", [("bnc -> bcn", (3, 2000, 24)), ("bcn -> bnc", (3, 24, 2000)), ("anbc -> abcn", (2, 1500, 3, 10)), ("nc -> cn", (1100, 40))])
@pytest.mark.parametrize("dt", ["f32", "f64"])
def test_strip_transpose_batched_and_signed_zero(nc, oracle, spec, shape, dt):
    """batch loops in front of the plane; the default transform into a fresh output is 0 + x: -0.0 is stored as +0.0"""
    a = rand(oracle, shape, dt, SEED + 9, 2, -1.0, 1.0)                       # set R: signed zeros, infinities, denormals
    assert_bit_exact(nc.einsum(spec, to_gpu(nc, a)), refsem.einsum(oracle, spec, [a]))


# ---- contractions that are really reductions (vdot, row-wise dot, matrix x vector) ---------------------------------------
DOTS = [("i i -> ", [(70001,), (70001,)]), ("ij ij -> i", [(300, 1000), (300, 1000)]), ("ij ij -> i", [(20000, 3), (20000, 3)]),
        ("ij ij -> ", [(300, 1001), (300, 1001)]), ("ij ij -> j", [(1000, 300), (1000, 300)]), ("ij j -> i", [(300, 1000), (1000,)]),
        ("ij j -> i", [(301, 999), (999,)]), ("i ij -> j", [(1000,), (1000, 300)]), ("i ij -> j", [(999,), (999, 301)]),
        ("ij j -> i", [(67, 4096), (4096,)]), ("bij bj -> bi", [(40, 30, 52), (40, 52)]), ("ij i -> i", [(300, 1000), (300,)]), ("j ij -> i", [(1000,), (300, 1000)])]


@pytest.mark.parametrize("spec,shapes", DOTS)
@pytest.mark.parametrize("dt", ["f32", "f64", "fixnum", "ub8"])
def test_dot_like_contractions_bit_exact(nc, oracle, spec, shapes, dt):
    ins = [rand(oracle, s, dt, SEED + 10 + i, 1 if dt in ("f32", "f64") else 0, *((0, 9) if dt == "ub8" else (-4, 4))) for i, s in enumerate(shapes)]
    want = refsem.einsum(oracle, spec, ins)
    got = nc.einsum(spec, *[to_gpu(nc, a) for a in ins])
    assert _last(nc).startswith("dot_"), _last(nc)
    if want.shape == ():
        assert got == want.values().reshape(-1)[0]
    else:
        assert_bit_exact(got, want)


@pytest.mark.parametrize("spec,shapes", [DOTS[0], DOTS[1], DOTS[4], DOTS[5], DOTS[7], DOTS[9]])
@pytest.mark.parametrize("dt,tol", [("f32", 1e-5), ("f64", 1e-12)])
def test_dot_like_contractions_random_floats(nc, oracle, spec, shapes, dt, tol):
    ins = [rand(oracle, s, dt, SEED + 20 + i, 0, -1.0, 1.0) for i, s in enumerate(shapes)]
    want = refsem.einsum(oracle, spec, ins)
    got = nc.einsum(spec, *[to_gpu(nc, a) for a in ins])
    g = np.asarray(got.numpy() if hasattr(got, "numpy") else got, dtype=np.float64).reshape(-1)
    w = want.values().astype(np.float64).reshape(-1)
    scale = np.sqrt(sum(float(np.prod(s)) for s in shapes[:1])) if w.size == 1 else 0.0     # a single sum of mean-zero terms
    assert np.linalg.norm(g - w) <= tol * max(np.linalg.norm(w), scale)


def test_dot_accumulates_into_supplied_output(nc, oracle):
    a = rand(oracle, (300, 1000), "f32", SEED + 30, 1, -4, 4)
    x = rand(oracle, (1000,), "f32", SEED + 31, 1, -4, 4)
    o = rand(oracle, (300,), "f32", SEED + 32, 1, -4, 4)
    go = to_gpu(nc, o)                                       # before the oracle accumulates into o in place
    want = refsem.einsum(oracle, "ij j -> i", [a, x], [o])
    nc.einsum("ij j -> i", to_gpu(nc, a), to_gpu(nc, x), go)
    assert_bit_exact(go, want if not isinstance(want, (list, tuple)) else want[0])


# ---- comparisons into packed bit arrays (words assembled across lanes) ------------------------------------------------------
@pytest.mark.parametrize("sa,sb", [((64, 128), (128,)), ((33, 5, 64), (5, 64)), ((1000, 96), (96,)), ((1000, 96), (1000, 1)), ((7, 1056), (7, 1056)),
                                   ((129, 2048), ()), ((3, 40000), (3, 1)), ((50, 33), (33,)), ((2049, 32), (32,))])
@pytest.mark.parametrize("ta,tb", [("f32", "f32"), ("f64", "f64"), ("ub8", "ub8"), ("fixnum", "fixnum"), ("sb16", "ub8"), ("f32", "ub8"), ("ub31", "sb32")])
@pytest.mark.parametrize("op", ["</bit", ">=/bit", "=/bit"])
def test_comparisons_broadcast_shapes(nc, oracle, sa, sb, ta, tb, op):
    lo_hi = lambda t: (-3, 3) if t in ("f32", "f64", "fixnum", "sb16", "sb32") else (0, 6)
    x = rand(oracle, sa, ta, SEED + 40, 1, *lo_hi(ta))
    y = rand(oracle, sb, tb, SEED + 41, 1, *lo_hi(tb))
    want = refsem.broadcast(oracle, op, x, y)
    assert want.dtype == "bit"
    assert_bit_exact(nc.broadcast(op, to_gpu(nc, x), to_gpu(nc, y)), want)


@pytest.mark.parametrize("sa,sb", [((50000, 3), (3,)), ((50000, 3), (50000, 1)), ((4099, 33), (33,)), ((700, 100), ()), ((10001, 7), (10001, 7)), ((3000, 5), (1, 5))])
@pytest.mark.parametrize("ta,tb", [("f32", "f32"), ("f64", "f32"), ("ub8", "ub8"), ("fixnum", "f32"), ("sb32", "sb32")])
@pytest.mark.parametrize("op", ["</bit", "=/bit"])
def test_comparisons_whose_rows_are_not_whole_result_words(nc, oracle, sa, sb, ta, tb, op):
    """`(N,3) < (3)`: the result's rows are 3 bits long.  The element-granular flat mode walks the result as one flat run of
    32-bit words (a lane owns a word = 32 consecutive elements, operands of another shape are gathered per element); the
    last partial word is a tiny nest of its own"""
    lo_hi = lambda t: (-3, 3) if t in ("f32", "f64", "fixnum", "sb32") else (0, 6)
    x = rand(oracle, sa, ta, SEED + 42, 1, *lo_hi(ta))
    y = rand(oracle, sb, tb, SEED + 43, 1, *lo_hi(tb))
    want = refsem.broadcast(oracle, op, x, y)
    got = nc.broadcast(op, to_gpu(nc, x), to_gpu(nc, y))
    assert_bit_exact(got, want)
    if sa[0] * sa[1] >= 64 * 32 and not (ta == "fixnum" and tb == "f32"):
        assert _last(nc) in ("bcast_vec_flat", "bcast_vec", "cmp_pack"), _last(nc)


@pytest.mark.parametrize("n", [33, 1000, 70001, 1 << 20 | 5])
@pytest.mark.parametrize("ta", ["f32", "f64", "ub8", "fixnum"])
def test_comparison_of_an_odd_sized_array_with_a_scalar(nc, oracle, n, ta):
    """(numcl:< a 2): whole result words on the vector kernels, the last partial word on its own"""
    x = rand(oracle, (n,), ta, SEED + 44, 1, 0 if ta == "ub8" else -3, 6 if ta == "ub8" else 3)
    y = rand(oracle, (), ta, SEED + 45, 1, 1, 1)
    for op in ("</bit", ">=/bit"):
        got = nc.broadcast(op, to_gpu(nc, x), to_gpu(nc, y))
        assert_bit_exact(got, refsem.broadcast(oracle, op, x, y))
    y2 = rand(oracle, (n,), ta, SEED + 46, 1, 0 if ta == "ub8" else -3, 6 if ta == "ub8" else 3)
    assert_bit_exact(nc.broadcast("=/bit", to_gpu(nc, x), to_gpu(nc, y2)), refsem.broadcast(oracle, "=/bit", x, y2))
