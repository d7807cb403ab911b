"""Bit arrays at speed: every numcl comparison returns one (src/4numeric.lisp:609-655), masks are combined with the bitwise
functions (:584-602), counted with numcl:sum and applied with numcl:* -- none of which may fall to the one-element-per-thread
paths.  Also: map-array over three arrays with a multiply-add form."""
import numpy as np
import pytest

import refsem
from test_gpu_parity import SEED, assert_bit_exact, rand, to_gpu

pytestmark = pytest.mark.gpu


def _last(nc):
    return nc.lib().ncl_last_kernel().decode()


def bits(oracle, shape, seed):
    return rand(oracle, shape, "bit", seed, 0, 0, 1)


@pytest.mark.parametrize("shape", [(64,), (1000,), (4099, 33), (1 << 20 | 5,), (256, 512), (7, 32, 33)])
@pytest.mark.parametrize("fn", ["logand", "logior", "logxor", "logeqv", "lognand", "lognor", "logandc1", "logandc2", "logorc1", "logorc2"])
def test_bitwise_functions_of_bit_arrays_run_on_words(nc, oracle, shape, fn):
    x, y = bits(oracle, shape, SEED + 50), bits(oracle, shape, SEED + 51)
    want = refsem.broadcast(oracle, fn, x, y)
    got = nc.broadcast(fn, to_gpu(nc, x), to_gpu(nc, y))
    assert_bit_exact(got, want)
    if int(np.prod(shape)) >= 4096:
        assert _last(nc) != "generic_nest" or int(np.prod(shape)) % 32 != 0, _last(nc)


@pytest.mark.parametrize("shape", [(96,), (1000,), (513, 65), (1 << 18,)])
def test_lognot_of_a_bit_array(nc, oracle, shape):
    """numcl infers (integer -2 -1) for (lognot bit): a (signed-byte 8) array of -1 / -2, not a bit array"""
    x = bits(oracle, shape, SEED + 52)
    got = nc.lognot(to_gpu(nc, x))
    assert got.dtype.name == "sb8"
    assert np.array_equal(got.numpy(), ~x.values())


def test_bitwise_on_a_window_that_does_not_start_on_a_word(nc, oracle):
    """offsets that are not multiples of 32 keep the element-wise path: same answer, neighbours untouched"""
    base = bits(oracle, (5000,), SEED + 53)
    g = to_gpu(nc, base)
    a = nc.NArray((4000,), "bit", _storage=g._buf, offset=7)
    b = to_gpu(nc, bits(oracle, (4000,), SEED + 54))
    got = nc.broadcast("logand", a, b)
    assert np.array_equal(got.numpy(), base.values()[7:4007] & b.numpy())
    assert np.array_equal(g.numpy(), base.values())


@pytest.mark.parametrize("shape,axes", [((1 << 20,), None), ((1000, 1000), None), ((1 << 16 | 3,), None), ((512, 4096), (1,)), ((512, 4096), (0,)),
                                        ((300, 70), (1,)), ((64, 64, 64), (0, 2))])
@pytest.mark.parametrize("where", ["device", "host"])
def test_counting_the_ones_of_a_bit_array(nc, oracle, shape, axes, where):
    """(numcl:sum (numcl:< a b)): word popcounts for the whole array, an (unsigned-byte 8) copy for everything else"""
    x = bits(oracle, shape, SEED + 55)
    g = to_gpu(nc, x) if where == "device" else nc.asarray(x.values(), type="bit", where="host")
    got = nc.sum(g, axes=axes)
    want = x.values().sum(axis=axes)
    if axes is None:
        assert got == want
    else:
        assert got.dtype.name == "fixnum" and np.array_equal(got.numpy(), want)
    assert _last(nc) != "generic_nest"


def test_amax_and_mean_of_packed_arrays(nc, oracle):
    x = bits(oracle, (300, 1000), SEED + 56)
    assert np.array_equal(nc.amax(to_gpu(nc, x), axes=1).numpy(), x.values().max(axis=1))
    u4 = rand(oracle, (100, 5000), "ub4", SEED + 57, 0, 0, 15)
    assert np.array_equal(nc.sum(to_gpu(nc, u4), axes=1).numpy(), u4.values().sum(axis=1))
    assert nc.sum(to_gpu(nc, u4)) == u4.values().sum()


@pytest.mark.parametrize("dt", ["f32", "f64", "fixnum", "ub8"])
@pytest.mark.parametrize("shape", [(1 << 16,), (1000, 37), (4099,)])
def test_masking_and_conversion_of_bit_arrays(nc, oracle, dt, shape):
    """mask * array (numcl's `where`) and astype of a bit array read E consecutive bits per chunk"""
    m = bits(oracle, shape, SEED + 58)
    a = rand(oracle, shape, dt, SEED + 59, 1 if dt in ("f32", "f64") else 0, 0 if dt == "ub8" else -8, 8)
    assert_bit_exact(nc.mul(to_gpu(nc, m), to_gpu(nc, a)), refsem.arith(oracle, "*", [m, a]))
    assert_bit_exact(nc.mul(to_gpu(nc, a), to_gpu(nc, m)), refsem.arith(oracle, "*", [a, m]))
    c = nc.astype(to_gpu(nc, m), dt)
    assert c.dtype.name == dt and np.array_equal(c.numpy(), m.values().astype(c.numpy().dtype))
    if int(np.prod(shape)) >= 4096 and int(np.prod(shape)) % 32 == 0:
        assert _last(nc) == "bcast_vec", _last(nc)


@pytest.mark.parametrize("dt", ["f32", "f64", "fixnum"])
@pytest.mark.parametrize("form", ["(+ (* $1 $2) $3)", "(+ $3 (* $1 $2))"])
def test_map_array_multiply_add_over_three_arrays(nc, oracle, dt, form):
    shape = (513, 257)
    xs = [rand(oracle, shape, dt, SEED + 60 + i, 0 if dt != "fixnum" else 0, -1.0 if dt != "fixnum" else -9, 1.0 if dt != "fixnum" else 9) for i in range(3)]
    want = oracle.map_array(form, xs, dt)
    got = nc.map_array(form, *[to_gpu(nc, x) for x in xs])
    assert_bit_exact(got, want)
    assert _last(nc) != "generic_nest", _last(nc)


# ---- contractions with a transposed operand: A^T B, A B^T (k_gemm.cu: the operand is re-laid out, then the GEMM kernels run) -----
@pytest.mark.parametrize("spec,sa,sb", [("ij ik -> jk", (300, 70), (300, 90)), ("ij kj -> ik", (130, 257), (64, 257)), ("ij ik -> jk", (1024, 256), (1024, 256)),
                                        ("ji ki -> jk", (128, 512), (256, 512)), ("bij bik -> bjk", (6, 200, 48), (6, 200, 80)), ("bij bkj -> bik", (5, 128, 96), (5, 256, 96)),
                                        ("ij ik -> jk", (4099, 33), (4099, 17))])
@pytest.mark.parametrize("dt", ["f32", "f64", "fixnum"])
def test_contractions_with_transposed_operands(nc, oracle, spec, sa, sb, dt):
    exact = dt == "fixnum"
    A = rand(oracle, sa, dt, SEED + 70, 1, -4, 4)                    # integer-valued data: every order of summation is exact
    B = rand(oracle, sb, dt, SEED + 71, 1, -4, 4)
    want = refsem.einsum(oracle, spec, [A, B])
    got = nc.einsum(spec, to_gpu(nc, A), to_gpu(nc, B))
    assert_bit_exact(got, want)
    assert _last(nc).startswith("gemm") or _last(nc).startswith("dot"), _last(nc)
    if not exact:                                                    # random floats: within the contraction tolerance
        A = rand(oracle, sa, dt, SEED + 72, 0, -1.0, 1.0); B = rand(oracle, sb, dt, SEED + 73, 0, -1.0, 1.0)
        w = refsem.einsum(oracle, spec, [A, B]).values().astype(np.float64)
        g = nc.einsum(spec, to_gpu(nc, A), to_gpu(nc, B)).numpy().astype(np.float64)
        assert np.linalg.norm(g - w) / np.linalg.norm(w) <= (1e-5 if dt == "f32" else 1e-12)


def test_transposed_contraction_accumulates_into_a_supplied_output(nc, oracle):
    A = rand(oracle, (500, 64), "f64", SEED + 74, 1, -4, 4); B = rand(oracle, (500, 96), "f64", SEED + 75, 1, -4, 4)
    C0 = rand(oracle, (64, 96), "f64", SEED + 76, 1, -4, 4)
    g = to_gpu(nc, C0)
    nc.einsum("ij ik -> jk", to_gpu(nc, A), to_gpu(nc, B), g)
    assert np.array_equal(g.numpy(), C0.values() + A.values().T @ B.values())


# ---- (unsigned-byte 2) / (unsigned-byte 4) results and operands: one 32-bit word of packed fields per chunk ---------------------
@pytest.mark.parametrize("n", [64, 1000, 4099, 1 << 18, (1 << 18) + 13])
def test_small_integer_results_are_packed_words(nc, oracle, n):
    """numcl infers the narrowest type: mask + mask and mask + 1 are (unsigned-byte 2), (mod a 7) is (unsigned-byte 4)"""
    m1, m2 = bits(oracle, (n,), SEED + 80), bits(oracle, (n,), SEED + 81)
    got = nc.add(to_gpu(nc, m1), to_gpu(nc, m2))
    assert got.dtype.name == "ub2"
    assert_bit_exact(got, refsem.arith(oracle, "+", [m1, m2]))
    assert_bit_exact(nc.add(to_gpu(nc, m1), 1), refsem.arith(oracle, "+", [m1, 1]))
    f = rand(oracle, (n,), "fixnum", SEED + 82, 0, -1000, 1000)
    r = nc.mod(to_gpu(nc, f), 7)
    assert r.dtype.name == "ub4" and np.array_equal(r.numpy(), np.mod(f.values(), 7))
    if n >= 4096:
        assert _last(nc) != "generic_nest", _last(nc)
    u4 = rand(oracle, (n,), "ub4", SEED + 83, 0, 0, 7)
    s = nc.add(to_gpu(nc, u4), r)                             # ub4 + ub4: (integer 0 13) -> ub4
    assert_bit_exact(s, refsem.arith(oracle, "+", [u4, oracle.OArr.from_values(np.mod(f.values(), 7), "ub4")]))
    assert_bit_exact(nc.mul(to_gpu(nc, u4), to_gpu(nc, m1)), refsem.arith(oracle, "*", [u4, m1]))
    w = nc.astype(to_gpu(nc, u4), "f32")
    assert np.array_equal(w.numpy(), u4.values().astype(np.float32))


def test_packed_results_in_two_dimensions_and_views(nc, oracle):
    a = rand(oracle, (300, 64), "ub4", SEED + 84, 0, 0, 7); b = rand(oracle, (64,), "ub4", SEED + 85, 0, 0, 7)
    assert_bit_exact(nc.add(to_gpu(nc, a), to_gpu(nc, b)), refsem.arith(oracle, "+", [a, b]))
    c = rand(oracle, (301, 33), "ub2", SEED + 86, 0, 0, 1)
    assert_bit_exact(nc.add(to_gpu(nc, c), to_gpu(nc, c)), refsem.arith(oracle, "+", [c, c]))
    # a result window that does not start on a word boundary keeps the element-wise path: neighbours untouched
    base = nc.NArray((5000,), "ub4"); base.set_raw(np.full(base.nbytes, 0xAB, dtype=np.uint8))
    view = nc.NArray((4000,), "ub4", _storage=base._buf, offset=3)
    x = rand(oracle, (4000,), "ub4", SEED + 87, 0, 0, 15)
    nc.map_array_into(view, "identity", to_gpu(nc, x))
    got = base.numpy()
    sentinel = np.tile(np.array([0xB, 0xA]), 2500)                          # the nibbles of 0xAB, low one first
    assert np.array_equal(got[3:4003], x.values()) and np.array_equal(got[:3], sentinel[:3]) and np.array_equal(got[4003:], sentinel[4003:])


# ---- n-ary folds and reductions over non-adjacent axes through einsum -----------------------------------------------------------
@pytest.mark.parametrize("fn,op", [("add", "+"), ("mul", "*"), ("sub", "-"), ("div", "/"), ("maximum", "max"), ("minimum", "min")])
@pytest.mark.parametrize("dts", [("f32", "f32", "f32"), ("f32", "f64", "f32"), ("ub8", "ub8", "ub8"), ("fixnum", "f32", "ub8"), ("f64", "f64", "f64", "f64")])
@pytest.mark.parametrize("shapes", [((300, 70), (300, 70), (300, 70), (300, 70)), ((300, 1), (1, 70), (70,), (300, 70)), ((4099,), (), (4099,), ())])
def test_folds_over_three_and_four_arrays(nc, oracle, fn, op, dts, shapes):
    """(numcl:+ a b c): the reference's chain of binary passes with typed intermediates"""
    if op in ("max", "min") and len(set(dts)) > 1 and "fixnum" in dts:
        pytest.skip("max / min of mixed integer / float arrays return the original object in CL: single-pass path")
    xs = []
    for i, dt in enumerate(dts):
        lo, hi = ((1, 9) if dt in ("ub8", "fixnum") else (1.0, 9.0)) if op == "/" else ((0, 9) if dt == "ub8" else (-9, 9))
        xs.append(rand(oracle, shapes[i], dt, SEED + 90 + i, 1, lo, hi))
    want = refsem.arith(oracle, op, xs)
    got = getattr(nc, fn)(*[to_gpu(nc, x) for x in xs])
    assert_bit_exact(got, want)
    if int(np.prod(want.shape)) >= 4096:
        assert _last(nc) != "generic_nest", _last(nc)


@pytest.mark.parametrize("spec,shape", [("ijk -> j", (40, 50, 60)), ("ijkl -> jl", (12, 30, 14, 40)), ("ijkl -> k", (12, 30, 14, 40)), ("ijk -> j", (256, 64, 256))])
@pytest.mark.parametrize("dt", ["f32", "f64", "fixnum", "ub8"])
def test_einsum_reductions_over_non_adjacent_axes(nc, oracle, spec, shape, dt):
    a = rand(oracle, shape, dt, SEED + 95, 1, 0 if dt == "ub8" else -8, 8)
    assert_bit_exact(nc.einsum(spec, to_gpu(nc, a)), refsem.einsum(oracle, spec, [a]))
    assert _last(nc) != "generic_nest", _last(nc)
    b = rand(oracle, shape, dt, SEED + 96, 1, 0 if dt == "ub8" else -8, 8)
    ax = tuple(i for i, ch in enumerate(spec.split(" -> ")[0]) if ch not in spec.split(" -> ")[1])
    assert_bit_exact(nc.amax(to_gpu(nc, b), axes=ax), refsem.reduce_array(oracle, "max", b, ax))


@pytest.mark.parametrize("spec,sa,sb", [("ij jk -> ik", (200, 96), (96, 300)), ("ij ik -> jk", (256, 128), (256, 256)), ("bij bjk -> bik", (4, 128, 64), (4, 64, 256))])
@pytest.mark.parametrize("da,db", [("f32", "f64"), ("f64", "f32"), ("fixnum", "f32"), ("f64", "ub8"), ("ub8", "f32")])
def test_contractions_of_mixed_element_classes(nc, oracle, spec, sa, sb, da, db):
    """single-float x double-float, integer x float: contagion converts the lower operand before the product, which is what
    converting the array first does; the result has the higher class"""
    A = rand(oracle, sa, da, SEED + 97, 1, 0 if da == "ub8" else -4, 4); B = rand(oracle, sb, db, SEED + 98, 1, 0 if db == "ub8" else -4, 4)
    want = refsem.einsum(oracle, spec, [A, B])
    got = nc.einsum(spec, to_gpu(nc, A), to_gpu(nc, B))
    assert_bit_exact(got, want)                                      # integer-valued data: exact in any order
    assert _last(nc).startswith("gemm"), _last(nc)
    if "fixnum" not in (da, db) and "ub8" not in (da, db):
        A = rand(oracle, sa, da, SEED + 99, 0, -1.0, 1.0); B = rand(oracle, sb, db, SEED + 100, 0, -1.0, 1.0)
        w = refsem.einsum(oracle, spec, [A, B]).values().astype(np.float64)
        g = nc.einsum(spec, to_gpu(nc, A), to_gpu(nc, B)).numpy().astype(np.float64)
        assert np.linalg.norm(g - w) / np.linalg.norm(w) <= 1e-12


# ---- composite transforms: node-by-node passes of the engine instead of the per-element interpreter -------------------------------
@pytest.mark.parametrize("form", ["(sqrt (+ (* $1 $1) (* $2 $2)))", "(* 2.0 (+ $1 $2))", "(+ (* $1 $2) (* $1 $1) $2)", "(max (abs $1) (abs $2))",
                                  "(- (* $1 $1) (/ $2 3.0d0))", "(* (+ $1 1) (- $2 1))", "(+ (square $1) (abs (- $2)))"])
@pytest.mark.parametrize("dts", [("f32", "f32"), ("f64", "f64"), ("f32", "f64"), ("fixnum", "fixnum"), ("ub8", "f32")])
def test_composite_transforms_match_the_interpreter(nc, oracle, form, dts):
    if "sqrt" in form and "fixnum" in dts:
        pytest.skip("sqrt of integers goes through single-float substitution: tolerance only")
    shape = (300, 257)
    xs = [rand(oracle, shape, dt, SEED + 110 + i, 1, 0 if dt == "ub8" else -6, 6) for i, dt in enumerate(dts)]
    gs = [to_gpu(nc, x) for x in xs]
    got = nc.map_array(form, *gs)
    assert _last(nc) != "generic_nest", (_last(nc), form)
    L = nc.lib(); L.ncl_force_generic(1)
    try:
        ref = nc.map_array(form, *gs)                            # the interpreter: one kernel, every node rounded in its own class
        assert _last(nc) == "generic_nest"
    finally:
        L.ncl_force_generic(0)
    assert got.dtype.name == ref.dtype.name and np.array_equal(got.raw(), ref.raw())
    if "sqrt" not in form and "3.0d0" not in form:
        want = oracle.map_array(form, xs, got.dtype.name)
        assert_bit_exact(got, want)


def test_composite_transform_in_an_einsum_with_broadcasting(nc, oracle):
    a = rand(oracle, (400, 300), "f32", SEED + 120, 1, -6, 6); v = rand(oracle, (300,), "f32", SEED + 121, 1, -6, 6)
    spec = "ij j -> (* $2 (+ $1 1.0)) -> ij"
    got = nc.einsum(spec, to_gpu(nc, a), to_gpu(nc, v))
    assert _last(nc) != "generic_nest"
    assert np.array_equal(got.numpy(), (v.values() * (a.values() + np.float32(1.0))).astype(np.float32))


@pytest.mark.parametrize("dt", ["f32", "f64", "fixnum"])
def test_sum_of_squares_of_one_array(nc, oracle, dt):
    """(einsum '(i -> (+ @1 (* $1 $1)) -> )): the squared norm; the same array is both operands of the dot-like reduction"""
    x = rand(oracle, (1 << 18,), dt, SEED + 122, 1, -4, 4)
    got = nc.einsum("i -> (+ @1 (* $1 $1)) -> ", to_gpu(nc, x))
    assert _last(nc).startswith("dot_"), _last(nc)
    assert got == (x.values().astype(np.float64) ** 2).sum()
    m = rand(oracle, (300, 1000), dt, SEED + 123, 1, -4, 4)
    rows = nc.einsum("ij -> (+ @1 (* $1 $1)) -> i", to_gpu(nc, m))
    assert np.array_equal(rows.numpy().astype(np.float64), (m.values().astype(np.float64) ** 2).sum(axis=1))


@pytest.mark.parametrize("spec,sa,sb", [("i i -> ", (70001,), (70001,)), ("ij ij -> i", (300, 1000), (300, 1000)), ("ij j -> i", (300, 1000), (1000,)),
                                        ("i ij -> j", (999,), (999, 301)), ("ij ij -> ", (257, 513), (257, 513))])
@pytest.mark.parametrize("da,db", [("f32", "f64"), ("f64", "f32"), ("ub8", "fixnum"), ("fixnum", "f32"), ("sb16", "ub8"), ("f64", "ub8")])
def test_dot_like_contractions_of_mixed_element_types(nc, oracle, spec, sa, sb, da, db):
    """a full dot product of a single-float and a double-float vector used to be ONE thread of the generic kernel"""
    lo = lambda t: 0 if t == "ub8" else -4
    A = rand(oracle, sa, da, SEED + 130, 1, lo(da), 4); B = rand(oracle, sb, db, SEED + 131, 1, lo(db), 4)
    want = refsem.einsum(oracle, spec, [A, B])
    got = nc.einsum(spec, to_gpu(nc, A), to_gpu(nc, B))
    assert _last(nc).startswith("dot_") or _last(nc).startswith("gemm"), _last(nc)
    if want.shape == ():
        assert got == want.values().reshape(-1)[0]
    else:
        assert_bit_exact(got, want)


@pytest.mark.parametrize("spec,sa,sb", [("ij jk -> ki", (300, 70), (70, 90)), ("ij jk -> ki", (256, 128), (128, 256)), ("bij bjk -> bki", (4, 128, 64), (4, 64, 256)),
                                        ("ij ik -> kj", (256, 128), (256, 64))])
@pytest.mark.parametrize("dt", ["f32", "f64", "fixnum"])
def test_contraction_with_a_transposed_result(nc, oracle, spec, sa, sb, dt):
    """C^T = B^T A^T: the operands swap roles when the result is contiguous along the loops of the first operand"""
    A = rand(oracle, sa, dt, SEED + 140, 1, -4, 4); B = rand(oracle, sb, dt, SEED + 141, 1, -4, 4)
    got = nc.einsum(spec, to_gpu(nc, A), to_gpu(nc, B))
    assert_bit_exact(got, refsem.einsum(oracle, spec, [A, B]))
    assert _last(nc).startswith("gemm"), _last(nc)


# ---- reductions of a function of the arrays, and sums into a result of another class (ncl_plan.cu: run_map_reduce) ---------------
@pytest.mark.parametrize("spec,nin", [("ij -> (+ @1 (abs $1)) -> ", 1), ("ij -> (max @1 (abs $1)) -> ", 1), ("ij ij -> (+ @1 (abs (- $1 $2))) -> ", 2),
                                      ("ij -> (+ @1 (abs $1)) -> i", 1), ("ij -> (max @1 (* $1 $1)) -> j", 1), ("ij ij -> (+ @1 (* (+ $1 $2) $2)) -> i", 2),
                                      ("ij -> (min @1 (- $1)) -> ", 1)])
@pytest.mark.parametrize("dt", ["f32", "f64", "fixnum"])
def test_reductions_of_a_function_of_the_arrays(nc, oracle, spec, nin, dt):
    """norms and distances written as einsum transforms: one thread of the generic kernel used to walk the whole array"""
    xs = [rand(oracle, (300, 1000), dt, SEED + 150 + i, 1, -6, 6) for i in range(nin)]
    want = refsem.einsum(oracle, spec, xs)
    got = nc.einsum(spec, *[to_gpu(nc, x) for x in xs])
    assert _last(nc) == "map_reduce", _last(nc)
    if want.shape == ():
        assert got == want.values().reshape(-1)[0]
    else:
        assert_bit_exact(got, want)


def test_sum_into_a_result_of_another_class(nc, oracle):
    """a single-float array summed into a supplied double-float scalar: every element is converted before it is added, so the
    sum runs in double-float; an (unsigned-byte 8) array into a single-float vector; accumulation into what the result held"""
    a = rand(oracle, (300, 1000), "f32", SEED + 160, 0, -1.0, 1.0)
    out = nc.asarray(np.array(2.5), type="f64")
    nc.einsum("ij -> ", to_gpu(nc, a), out)
    truth = 2.5 + a.values().astype(np.float64).sum()
    assert abs(out.item() - truth) <= 1e-12 * max(1.0, abs(truth)) * 300
    assert _last(nc) == "map_reduce"
    u = rand(oracle, (300, 1000), "ub8", SEED + 161, 0, 0, 9)
    rows = nc.zeros((300,), type="f32")
    nc.einsum("ij -> i", to_gpu(nc, u), rows)
    assert np.array_equal(rows.numpy(), u.values().sum(axis=1).astype(np.float32))


@pytest.mark.parametrize("dt", ["f32", "f64", "fixnum"])
def test_sum_of_an_outer_product(nc, oracle, dt):
    """(einsum '(i j -> ) a b): no input spans both loops; the (small) iteration space is materialised, then folded"""
    a = rand(oracle, (700,), dt, SEED + 170, 1, -4, 4); b = rand(oracle, (900,), dt, SEED + 171, 1, -4, 4)
    got = nc.einsum("i j -> ", to_gpu(nc, a), to_gpu(nc, b))
    assert _last(nc) == "map_reduce", _last(nc)
    assert got == a.values().astype(np.float64).sum() * b.values().astype(np.float64).sum()
