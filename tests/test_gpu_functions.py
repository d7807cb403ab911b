"""The rest of numcl's scalar-function surface on the GPU (src/4numeric.lisp:466-519,533,568-602), against the oracle:
bitwise family, floor / ceiling / round / truncate / mod / rem, expt, isqrt / logcount / integer-length, inverse and
hyperbolic functions, exact integer-vs-float comparisons, the (unsigned-byte 64) fence and the DOMAIN flag."""
import numpy as np
import pytest

import refsem
from numcl_b200 import _lib
from numcl_b200 import typeinfer as ti
from test_gpu_parity import SEED, assert_bit_exact, rand, to_gpu

pytestmark = pytest.mark.gpu

BITWISE = ["logandc1", "logandc2", "logeqv", "lognand", "lognor", "logorc1", "logorc2"]
DIVLIKE = ["floor", "ceiling", "round", "truncate", "mod", "rem"]


def _run_broadcast(nc, oracle, fn, a, b, out_dt):
    """the op through ncl_broadcast into an output of a given element type, against the oracle's broadcast-core"""
    ga, gb = to_gpu(nc, a), to_gpu(nc, b)
    got = nc.broadcast(fn, ga, gb, type=out_dt)
    want = oracle.broadcast(fn, a, b, out_dt)
    return got, want


@pytest.mark.parametrize("fn", BITWISE)
@pytest.mark.parametrize("dta,dtb,out", [("ub8", "ub8", "sb16"), ("fixnum", "fixnum", "fixnum"), ("sb32", "ub16", "sb64"), ("ub8", "fixnum", "fixnum"),
                                          ("sb8", "sb8", "sb8"), ("ub32", "ub32", "ub32")])
def test_bitwise_family_bit_exact(nc, oracle, fn, dta, dtb, out):
    lo = {"ub8": (0, 255), "fixnum": (-2 ** 61, 2 ** 61), "sb32": (-2 ** 31, 2 ** 31 - 1), "ub16": (0, 65535), "sb8": (-128, 127), "ub32": (0, 2 ** 32 - 1)}
    for shape_a, shape_b in (((33, 40), (40,)), ((4096,), (4096,)), ((7, 5, 3), (5, 1))):
        a = rand(oracle, shape_a, dta, SEED, 0, *lo[dta])
        b = rand(oracle, shape_b, dtb, SEED + 1, 0, *lo[dtb])
        got, want = _run_broadcast(nc, oracle, fn, a, b, out)
        assert_bit_exact(got, want)


def test_bitwise_result_types_follow_the_reference_inferers(nc, oracle):
    a = rand(oracle, (50,), "ub8", SEED, 0, 0, 255)
    b = rand(oracle, (50,), "ub8", SEED + 1, 0, 0, 255)
    for fn in BITWISE:
        got = getattr(nc, fn)(to_gpu(nc, a), to_gpu(nc, b))
        want = refsem.broadcast(oracle, fn, a, b)
        assert_bit_exact(got, want)


@pytest.mark.parametrize("fn", DIVLIKE)
@pytest.mark.parametrize("dta,dtb", [("fixnum", "fixnum"), ("sb16", "ub8"), ("ub8", "sb8"), ("sb64", "sb32")])
def test_rounding_family_on_integers_bit_exact(nc, oracle, fn, dta, dtb):
    rng = {"fixnum": (-10 ** 12, 10 ** 12), "sb16": (-32768, 32767), "ub8": (1, 255), "sb8": (-128, -1), "sb64": (-2 ** 62, 2 ** 62), "sb32": (3, 2 ** 31 - 1)}
    for shape_a, shape_b in (((65, 48), (48,)), ((3000,), ()), ((9, 7), (9, 1))):
        a = rand(oracle, shape_a, dta, SEED, 0, *rng[dta])
        b = rand(oracle, shape_b, dtb, SEED + 1, 0, *rng[dtb])
        if dtb in ("fixnum", "sb32", "sb64"):                       # no zero divisors
            bv = b.values(); bv[bv == 0] = 7
            b = oracle.OArr.from_values(bv, dtb)
        got, want = _run_broadcast(nc, oracle, fn, a, b, "fixnum")
        assert_bit_exact(got, want)
    # ties and signs, exhaustively on a small range
    n = np.arange(-12, 13, dtype=np.int64)
    for d in (1, 2, 3, 4, -1, -2, -3, -4, 5, -5):
        a = oracle.OArr.from_values(n, "fixnum")
        b = oracle.OArr.from_values(np.array(d, dtype=np.int64), "fixnum")
        got, want = _run_broadcast(nc, oracle, fn, a, b, "fixnum")
        assert_bit_exact(got, want)
        ref = {"floor": lambda x: x // d, "mod": lambda x: x % d, "truncate": lambda x: int(x / d), "rem": lambda x: x - d * int(x / d),
               "ceiling": lambda x: -((-x) // d), "round": lambda x: int(round(x / d))}[fn]
        assert got.numpy().tolist() == [ref(int(x)) for x in n], (fn, d)


def test_rounding_family_api_and_result_types(nc, oracle):
    """(numcl:floor x 2) on an (unsigned-byte 8) array is (integer 0 127) -> (unsigned-byte 7); one argument means divisor 1"""
    a = rand(oracle, (300,), "ub8", SEED, 0, 0, 255)
    g = to_gpu(nc, a)
    r = nc.floor(g, 2)
    assert r.dtype.name == "ub7" and np.array_equal(r.numpy(), a.values() // 2)
    assert np.array_equal(nc.round_(g, 2).numpy(), np.array([int(round(int(v) / 2)) for v in a.values()]))
    assert np.array_equal(nc.mod(g, 10).numpy(), a.values() % 10)
    assert np.array_equal(nc.truncate(g).numpy(), a.values())
    with pytest.raises(ti.TypeInferenceError):
        nc.floor(g, to_gpu(nc, a))                                      # a divisor type that contains 0: (integer * *) -> T array


@pytest.mark.parametrize("fn", DIVLIKE)
@pytest.mark.parametrize("dt", ["f32", "f64"])
def test_rounding_family_on_floats_bit_exact(nc, oracle, fn, dt):
    """floats: SBCL's definitions (truncate of the rounded quotient, adjusted by the remainder's sign), each op rounded"""
    a = rand(oracle, (41, 64), dt, SEED, 0, -50.0, 50.0)
    b = rand(oracle, (64,), dt, SEED + 1, 0, 0.25, 7.0)
    nb = oracle.OArr.from_values(-b.values(), dt)
    out = dt if fn in ("mod", "rem") else "fixnum"
    for bb in (b, nb):
        got, want = _run_broadcast(nc, oracle, fn, a, bb, out)
        assert_bit_exact(got, want)
    # divisor 1 (the default): exact rounding, ties to even for round; halves included
    h = oracle.OArr.from_values((np.arange(-40, 41) / 2.0).astype(np.float32 if dt == "f32" else np.float64), dt)
    one = oracle.OArr.from_values(np.array(1, dtype=np.int64), "bit")
    got, want = _run_broadcast(nc, oracle, fn, h, one, out)
    assert_bit_exact(got, want)
    if fn == "round":
        assert got.numpy().tolist() == [int(np.rint(v)) for v in h.values()]


def test_float_to_integer_stores_wrap_like_coerce_beyond_2_63(nc, oracle):
    """(round x) of a float >= 2^63 is a bignum that %coerce wraps; conversion intrinsics would saturate"""
    vals = np.array([2.0 ** 63, -2.0 ** 63, 2.0 ** 64, 3.0 * 2.0 ** 62, 2.0 ** 70 + 2.0 ** 20, -(2.0 ** 66 + 2.0 ** 14), 1e19, 123456.5, -0.5, 2.5], dtype=np.float64)
    a = oracle.OArr.from_values(vals, "f64")
    for out in ("fixnum", "sb64", "ub8", "ub63"):
        got = nc.map_array_into(nc.empty(vals.shape, out), "identity", to_gpu(nc, a))
        want = oracle.map_array("(identity $1)", [a], out)
        assert_bit_exact(got, want)
    f = oracle.OArr.from_values(np.array([2.0 ** 63, 2.0 ** 90, -2.0 ** 64, 1e10], dtype=np.float32), "f32")
    assert_bit_exact(nc.map_array_into(nc.empty((4,), "fixnum"), "identity", to_gpu(nc, f)), oracle.map_array("(identity $1)", [f], "fixnum"))


def test_integer_functions(nc, oracle):
    a = rand(oracle, (500,), "fixnum", SEED, 0, -2 ** 61, 2 ** 61)
    p = rand(oracle, (500,), "fixnum", SEED + 1, 0, 0, 2 ** 62 - 1)
    small = rand(oracle, (777,), "ub8", SEED + 2, 0, 0, 255)
    for fn, x, out in (("logcount", a, "ub8"), ("integer-length", a, "ub8"), ("isqrt", p, "fixnum"), ("isqrt", small, "ub4"),
                       ("logcount", small, "ub4"), ("integer-length", small, "fixnum")):
        got = nc.map_array_into(nc.empty(x.shape, out), fn, to_gpu(nc, x))
        want = oracle.map_array(f"({fn} $1)", [x], out)
        assert_bit_exact(got, want)
    sq = np.array([0, 1, 2, 3, 4, 15, 16, 17, 2 ** 62 - 1, (2 ** 31 - 1) ** 2, (2 ** 31 - 1) ** 2 - 1], dtype=np.int64)
    r = nc.map_array_into(nc.empty(sq.shape, "fixnum"), "isqrt", nc.asarray(sq, type="fixnum")).numpy()
    import math
    assert r.tolist() == [math.isqrt(int(v)) for v in sq]
    assert nc.isqrt(to_gpu(nc, small)).dtype.name == "ub4"               # (floor (sqrt (integer 0 255))) = (integer 0 15)


@pytest.mark.parametrize("fn,lo,hi", [("asin", -1.0, 1.0), ("acos", -1.0, 1.0), ("atan", -30.0, 30.0), ("sinh", -8.0, 8.0), ("cosh", -8.0, 8.0),
                                       ("asinh", -1e4, 1e4), ("acosh", 1.0, 1e4), ("atanh", -0.99, 0.99), ("%log2", 1e-3, 1e6), ("sqrt", 0.0, 1e6),
                                       ("%logr", 1e-3, 1e6)])
@pytest.mark.parametrize("dt,rtol", [("f32", 4e-6), ("f64", 1e-13)])
def test_inverse_and_hyperbolic_functions_in_tolerance(nc, oracle, fn, lo, hi, dt, rtol):
    a = rand(oracle, (64, 33), dt, SEED, 0, lo, hi)
    got = nc.map_array_into(nc.empty(a.shape, dt), fn, to_gpu(nc, a)).numpy()
    want = oracle.map_array(f"({fn} $1)", [a], dt).values()
    assert np.allclose(got, want, rtol=rtol, atol=rtol), fn
    # integer arguments are converted to single-float first (float substitution)
    if fn in ("atan", "sinh", "asinh", "sqrt", "%log2"):
        u = rand(oracle, (200,), "ub8", SEED, 0, 1, 9)
        got = nc.map_array_into(nc.empty(u.shape, "f32"), fn, to_gpu(nc, u)).numpy()
        assert np.allclose(got, oracle.map_array(f"({fn} $1)", [u], "f32").values(), rtol=4e-6), fn


def test_unary_mapper_types(nc, oracle):
    u = rand(oracle, (64,), "ub8", SEED, 0, 0, 255)
    g = to_gpu(nc, u)
    assert nc.sqrt(g).dtype.name == "f32" and nc.asinh(g).dtype.name == "f32" and nc.log2(nc.add(g, 1)).dtype.name == "f32"
    assert not ti.reference_infers_complex("sqrt", ti.type_of_dtype("ub8"))
    assert ti.reference_infers_complex("sqrt", ti.type_of_dtype("f32"))          # (single-float * *): not provably >= 0
    assert ti.reference_infers_complex("asin", ti.type_of_dtype("bit"))


def test_domain_errors_are_reported_not_silently_nan(nc, oracle):
    L = nc.lib()
    ok = nc.asarray(np.array([0.0, 1.0, 4.0], dtype=np.float32), type="f32")
    assert nc.sqrt(ok).numpy().tolist() == [0.0, 1.0, 2.0]
    bad = nc.asarray(np.array([4.0, -1.0, 9.0], dtype=np.float32), type="f32")
    with pytest.raises(nc.NumclDomainError):
        nc.sqrt(bad).numpy()                                         # reported at the d2h that follows
    _lib.check(L.ncl_sync())                                         # the flag is cleared once reported
    for fn, v in (("%logr", -2.0), ("asin", 1.5), ("acos", -1.5), ("acosh", 0.5), ("atanh", 2.0), ("%log2", -1.0)):
        x = nc.asarray(np.array([0.5, v], dtype=np.float64), type="f64")
        r = nc.map_array_into(nc.empty((2,), "f64"), fn, x)
        with pytest.raises(nc.NumclDomainError):
            _lib.check(L.ncl_sync())
        _lib.check(L.ncl_sync())
    # integer division by zero, isqrt of a negative, an integer power beyond 64 bits
    a = nc.asarray(np.array([7, 8, 9], dtype=np.int64), type="fixnum")
    z = nc.asarray(np.array([1, 0, 3], dtype=np.int64), type="fixnum")
    for fn in ("floor", "mod", "rem", "truncate"):
        nc.broadcast(fn, a, z, type="fixnum")
        with pytest.raises(nc.NumclDomainError):
            _lib.check(L.ncl_sync())
    nc.map_array_into(nc.empty((3,), "fixnum"), "isqrt", nc.asarray(np.array([4, -4, 9], dtype=np.int64), type="fixnum"))
    with pytest.raises(nc.NumclDomainError):
        _lib.check(L.ncl_sync())
    nc.broadcast("expt", a, nc.asarray(np.array([2, 30, 3], dtype=np.int64), type="fixnum"), type="f32")
    with pytest.raises(nc.NumclDomainError):
        _lib.check(L.ncl_sync())
    # host-resident operands: reported by the call itself
    hbad = nc.asarray(np.array([1.0, -1.0], dtype=np.float32), type="f32", where="host")
    with pytest.raises(nc.NumclDomainError):
        nc.sqrt(hbad)
    _lib.check(L.ncl_sync())


def test_expt(nc, oracle):
    # integer ^ non-negative integer: exact, stored as the reference's inferred single-float (exp (* (log b) p))
    b = rand(oracle, (40,), "ub8", SEED, 0, 0, 15)
    p = rand(oracle, (40,), "ub4", SEED + 1, 0, 0, 9)
    got = nc.expt(to_gpu(nc, b), to_gpu(nc, p))
    want = refsem.broadcast(oracle, "expt", b, p)
    assert got.dtype.name == "f32"
    assert_bit_exact(got, want)
    # into an integer array: exact; negative powers are exact ratios rounded once into a float array
    bi = rand(oracle, (64,), "fixnum", SEED, 0, -9, 9)
    pi = rand(oracle, (64,), "fixnum", SEED + 1, 0, 0, 19)
    assert_bit_exact(*_run_broadcast(nc, oracle, "expt", bi, pi, "fixnum"))
    bnz = oracle.OArr.from_values(np.where(bi.values() == 0, 3, bi.values()), "fixnum")
    pn = oracle.OArr.from_values(-(pi.values() % 17), "fixnum")        # |b|^p < 2^53: the generic kernel's ratio -> double-float is one exact division
    assert_bit_exact(*_run_broadcast(nc, oracle, "expt", bnz, pn, "f32"))
    assert_bit_exact(*_run_broadcast(nc, oracle, "expt", bnz, pn, "f64"))
    # float ^ integer: SBCL's intexp (repeated squaring), every product rounded -- bit-exact
    for dt in ("f32", "f64"):
        x = rand(oracle, (50, 8), dt, SEED, 0, -3.0, 3.0)
        k = rand(oracle, (8,), "sb8", SEED + 1, 0, -12, 12)
        assert_bit_exact(*_run_broadcast(nc, oracle, "expt", x, k, dt))
    # float ^ float: pow, tolerance
    x = rand(oracle, (300,), "f32", SEED, 0, 0.1, 5.0)
    y = rand(oracle, (300,), "f32", SEED + 1, 0, -3.0, 3.0)
    got, want = _run_broadcast(nc, oracle, "expt", x, y, "f32")
    assert np.allclose(got.numpy(), want.values(), rtol=4e-6)


@pytest.mark.parametrize("op", ["=/bit", "/=/bit", "<=/bit", ">=/bit", "</bit", ">/bit"])
def test_integer_vs_float_comparisons_are_exact(nc, oracle, op):
    """CLHS 12.1.4.1: the float is converted with `rational` and the comparison is exact -- 16777217 is not = 1.6777216e7"""
    ints = np.array([2 ** 24, 2 ** 24 + 1, 2 ** 24 + 2, -(2 ** 24) - 1, 2 ** 53 + 1, 2 ** 53, -(2 ** 53) - 1, 2 ** 62 - 1, 7, 0, -3, 2 ** 31 - 1, 2 ** 31 + 1], dtype=np.int64)
    for fdt in ("f32", "f64"):
        ft = np.float32 if fdt == "f32" else np.float64
        floats = ints.astype(ft)                                    # rounded neighbours of the integers
        for idt in ("fixnum", "sb64"):
            a = oracle.OArr.from_values(ints, idt)
            b = oracle.OArr.from_values(floats, fdt)
            for x, y in ((a, b), (b, a)):
                got = nc.broadcast(op, to_gpu(nc, x), to_gpu(nc, y))
                want = refsem.broadcast(oracle, op, x, y)
                assert_bit_exact(got, want)
    # 32-bit integers against single-floats take the vector path in double-float lanes
    i32 = rand(oracle, (64, 128), "sb32", SEED, 0, 2 ** 24, 2 ** 31 - 1)
    f32 = oracle.OArr.from_values(i32.values().astype(np.float32), "f32")
    got = nc.broadcast(op, to_gpu(nc, i32), to_gpu(nc, f32))
    assert_bit_exact(got, refsem.broadcast(oracle, op, i32, f32))
    assert nc.lib().ncl_last_kernel() == b"cmp_pack"                 # the vector path (double-float lanes), not the generic kernel
    frac = oracle.OArr.from_values(np.array([0.5, -0.5, 1e30, -1e30, np.inf, -np.inf, 2.0 ** 63, -2.0 ** 63]), "f64")
    k = oracle.OArr.from_values(np.array([0, 0, 2 ** 62 - 1, -2 ** 62, 5, 5, 2 ** 62 - 1, -2 ** 62], dtype=np.int64), "fixnum")
    assert_bit_exact(nc.broadcast(op, to_gpu(nc, k), to_gpu(nc, frac)), refsem.broadcast(oracle, op, k, frac))


def test_mixed_max_min_with_large_integers_match_exact_comparison(nc, oracle):
    """max / min of an integer and a float: comparing after conversion gives the same VALUE as CL's exact comparison followed
    by the conversion of the winner (rounding is monotonic); checked against the oracle, which compares exactly"""
    ints = np.array([2 ** 24 + 1, 2 ** 24 + 3, -(2 ** 24) - 1, 2 ** 53 + 1, 2 ** 61 + 12345, -(2 ** 61) - 7, 5, -5], dtype=np.int64)
    for fdt in ("f32", "f64"):
        ft = np.float32 if fdt == "f32" else np.float64
        floats = np.nextafter(ints.astype(ft), np.array([np.inf, -np.inf] * 4, dtype=ft))
        a = oracle.OArr.from_values(ints, "fixnum")
        for fl in (ints.astype(ft), floats):
            b = oracle.OArr.from_values(fl, fdt)
            for fn, f in (("max", nc.maximum), ("min", nc.minimum)):
                assert_bit_exact(f(to_gpu(nc, a), to_gpu(nc, b)), refsem.arith(oracle, fn, [a, b]))
                assert_bit_exact(f(to_gpu(nc, b), to_gpu(nc, a)), refsem.arith(oracle, fn, [b, a]))


def test_ub64_is_exact_where_it_can_be_and_refused_elsewhere(nc, oracle):
    big = np.array([2 ** 64 - 1, 2 ** 63, 2 ** 63 + 5, 17, 0], dtype=np.uint64)
    a = nc.NArray((5,), "ub64"); a.set_raw(big)
    b = nc.NArray((5,), "ub64"); b.set_raw(np.array([1, 2 ** 63, 3, 2 ** 64 - 1, 9], dtype=np.uint64))
    # + - * and the bitwise family only need the value modulo 2^64: exact (the fixnum result is the low 63 bits, like %coerce)
    for fn, ref in (("+", lambda x, y: x + y), ("*", lambda x, y: x * y), ("-", lambda x, y: x - y)):
        r = nc.broadcast(fn, a, b, type="fixnum").numpy()
        xs = [int(v) for v in big]; ys = [1, 2 ** 63, 3, 2 ** 64 - 1, 9]
        want = [((ref(x, y) + 2 ** 62) % 2 ** 63) - 2 ** 62 for x, y in zip(xs, ys)]
        assert r.tolist() == want, fn
    assert nc.broadcast("logxor", a, b, type="ub64").raw()[:40].view(np.uint64).tolist() == (big ^ np.array([1, 2 ** 63, 3, 2 ** 64 - 1, 9], dtype=np.uint64)).tolist()
    assert nc.sum(a) == ((sum(int(v) for v in big) + 2 ** 62) % 2 ** 63) - 2 ** 62
    # anything that looks at order or magnitude is refused, not silently wrong
    for f in (lambda: nc.maximum(a, b), lambda: nc.lt(a, b), lambda: nc.amax(a), lambda: nc.broadcast("+", a, nc.ones((5,), type="f32")),
              lambda: nc.astype(a, "f64"), lambda: nc.broadcast("floor", a, b, type="fixnum"), lambda: nc.div(a, b)):
        with pytest.raises(nc.NumclUnsupportedError):
            f()
