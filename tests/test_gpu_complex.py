"""Complex element types (complex single-float) / (complex double-float) through the C ABI: SURVEY 8(b)'s dtype enum,
src/1type.lisp:341-344,446-451, src/4numeric.lisp:486-489, vdot's conjugate (src/4linear-algebra2.lisp:54-59).
Checker: oracle/complex_oracle.py (restatement of SBCL's complex number dispatch; parity unpinned, see its header)."""
import numpy as np
import pytest

from numcl_b200 import _lib
from oracle import complex_oracle as CO

pytestmark = pytest.mark.gpu
RNG = np.random.default_rng(20261019)


def cints(shape, ct=np.complex64, lo=-8, hi=8):
    """complex numbers with small integer parts: sums and products are exact in any order"""
    v = RNG.integers(lo, hi + 1, size=shape) + 1j * RNG.integers(lo, hi + 1, size=shape)
    return v.astype(ct)


def cfloats(shape, ct=np.complex64):
    v = RNG.uniform(-1, 1, size=shape) + 1j * RNG.uniform(-1, 1, size=shape)
    return v.astype(ct)


def bits(x):
    x = np.ascontiguousarray(x)
    return x.view(np.uint32 if x.dtype in (np.complex64, np.float32) else np.uint64)


def same_bits(got, want):
    assert got.shape == want.shape and got.dtype == want.dtype, (got.shape, want.shape, got.dtype, want.dtype)
    assert np.array_equal(bits(got), bits(want))


@pytest.mark.parametrize("ct", [np.complex64, np.complex128])
@pytest.mark.parametrize("op,fn", [("+", "add"), ("-", "sub"), ("*", "mul")])
@pytest.mark.parametrize("shapes", [((64, 33), (64, 33)), ((64, 33), (33,)), ((5, 1, 7), (6, 7)), ((1000,), ())])
def test_complex_arithmetic_bit_exact(nc, ct, op, fn, shapes):
    """numcl:+ - * on complex arrays, random floats: every float operation rounded on its own, componentwise"""
    a, b = cfloats(shapes[0], ct), cfloats(shapes[1], ct)
    got = getattr(nc, fn)(nc.asarray(a), nc.asarray(b))
    assert got.dtype.name == ("c64" if ct == np.complex64 else "c128")
    # numcl:+ / * fold from the identity: (0 + a) + b, (1 * a) * b; - folds from the first argument
    first = CO.binary(op, np.int64(0 if op == "+" else 1), a) if op in "+*" else a
    same_bits(got.numpy(), CO.binary(op, first, b))


@pytest.mark.parametrize("ct,tol", [(np.complex64, 2e-6), (np.complex128, 1e-14)])
def test_complex_division(nc, ct, tol):
    a, b = cfloats((128, 65), ct), cfloats((128, 65), ct) + ct(0.5)
    got = nc.div(nc.asarray(a), nc.asarray(b)).numpy()
    want = CO.binary("/", a, b)
    assert np.max(np.abs(got - want) / np.abs(want)) <= tol
    assert np.max(np.abs(got - a.astype(np.complex128) / b.astype(np.complex128)) / np.abs(want)) <= 8 * tol
    e = cints((16, 16), ct)
    two = nc.div(nc.asarray(e), 2)                                            # complex / real: both parts divided
    same_bits(two.numpy(), (e / 2).astype(ct))


@pytest.mark.parametrize("other", ["f32", "f64", "ub8", "fixnum"])
@pytest.mark.parametrize("op,fn", [("+", "add"), ("*", "mul"), ("-", "sub")])
def test_complex_with_real_operands_contagion(nc, other, op, fn):
    """(complex single-float) o double-float -> (complex double-float); integers become floats of the part type"""
    a = cfloats((40, 24))
    r = RNG.integers(0, 100, size=(24,)) if other in ("ub8", "fixnum") else RNG.uniform(-2, 2, size=(24,)).astype(np.float32 if other == "f32" else np.float64)
    gr = nc.asarray(r, type=other)
    for x, y, gx, gy in ((a, r, nc.asarray(a), gr), (r, a, gr, nc.asarray(a))):
        got = getattr(nc, fn)(gx, gy)
        assert got.dtype.name == ("c128" if other == "f64" else "c64")
        # numcl:+ / * fold from the identity; a real first operand meets it as a real: (0 + r), (1 * r) stay r
        first = CO.binary(op, np.int64(0 if op == "+" else 1), x) if op in "+*" and np.iscomplexobj(x) else x
        same_bits(got.numpy(), CO.binary(op, first, y))


def test_unary_functions_of_complex_arrays(nc):
    a = cfloats((31, 17))
    g = nc.asarray(a)
    same_bits(nc.conjugate(g).numpy(), np.conj(a))
    rp, ip = nc.realpart(g), nc.imagpart(g)
    assert rp.dtype.name == "f32" and ip.dtype.name == "f32"
    same_bits(rp.numpy(), a.real.copy()); same_bits(ip.numpy(), a.imag.copy())
    same_bits(nc.sub(g).numpy(), -a)                                          # (- x): both parts negated
    same_bits(nc.square(g).numpy(), CO.binary("*", a, a))
    ab = nc.map_array("abs", g)
    assert ab.dtype.name == "f32" and np.allclose(ab.numpy(), np.abs(a), rtol=2e-6)
    # on real arrays conjugate / realpart are the identity and imagpart is zero
    r = RNG.uniform(-1, 1, size=(50,)).astype(np.float32)
    gr = nc.asarray(r)
    same_bits(nc.conjugate(gr).numpy(), r); same_bits(nc.realpart(gr).numpy(), r)
    assert np.array_equal(nc.imagpart(gr).numpy(), np.zeros(50, np.float32))
    z = cints((9, 9), np.complex128)
    same_bits(nc.conjugate(nc.asarray(z)).numpy(), np.conj(z))


@pytest.mark.parametrize("ct", [np.complex64, np.complex128])
@pytest.mark.parametrize("axes", [None, (0,), (1,), (0, 2), (1, 2)])
def test_complex_sums_exact(nc, ct, axes):
    a = cints((48, 20, 36), ct)
    got = nc.sum(nc.asarray(a), axes=axes)
    want = a.sum(axis=axes)
    if axes is None:
        assert got == want
    else:
        same_bits(got.numpy(), want.astype(ct))


def test_sums_and_copies_of_complex_arrays_run_on_the_real_kernels(nc):
    """real and imaginary parts never meet in + - sum transpose: the nest is lowered to a float nest with one more
    innermost loop of extent 2 and keeps the fast kernel families; the generic kernel gives the same bits"""
    L = nc.lib()
    a, b = cints((512, 256)), cints((512, 256))
    ga, gb = nc.asarray(a), nc.asarray(b)
    r = nc.broadcast("+", ga, gb)
    assert L.ncl_last_kernel() != b"generic_nest"
    same_bits(r.numpy(), a + b)
    s = nc.sum(ga)
    assert L.ncl_last_kernel() != b"generic_nest" and s == a.sum()
    t = nc.transpose(ga)
    assert L.ncl_last_kernel() != b"generic_nest"
    same_bits(t.numpy(), (a.T + 0).astype(np.complex64))
    L.ncl_force_generic(1)
    try:
        same_bits(nc.broadcast("+", ga, gb).numpy(), a + b)
        assert L.ncl_last_kernel() == b"generic_nest"
        same_bits(nc.transpose(ga).numpy(), (a.T + 0).astype(np.complex64))
    finally:
        L.ncl_force_generic(0)


@pytest.mark.parametrize("ct", [np.complex64, np.complex128])
def test_complex_matmul_and_vdot(nc, ct):
    A, B = cints((24, 40), ct, -4, 4), cints((40, 16), ct, -4, 4)
    got = nc.matmul(nc.asarray(A), nc.asarray(B)).numpy()
    same_bits(got, CO.matmul_reference_order(A, B))
    same_bits(got, (A @ B).astype(ct))
    Af, Bf = cfloats((24, 40), ct), cfloats((40, 16), ct)
    same_bits(nc.matmul(nc.asarray(Af), nc.asarray(Bf)).numpy(), CO.matmul_reference_order(Af, Bf))     # the reference's own order
    x, y = cints((300,), ct), cints((300,), ct)
    assert nc.vdot(nc.asarray(x), nc.asarray(y)) == np.vdot(x, y)                                        # conjugates its FIRST argument
    assert nc.inner(nc.asarray(x), nc.asarray(y)) == np.dot(x, y)
    o = nc.outer(nc.asarray(x[:20]), nc.asarray(y[:30])).numpy()
    # the default transform accumulates into zeros: @1 <- (+ @1 (* $1 $2)), so a product of -0.0 is stored as +0.0
    same_bits(o, CO.binary("+", np.zeros((20, 30), ct), CO.binary("*", x[:20, None], y[None, :30])))


def test_constructors_conversion_and_host_arrays(nc):
    z = nc.zeros((5, 3), type="c64")
    assert z.numpy().dtype == np.complex64 and not z.numpy().any()
    f = nc.full((4, 6), 2.5, type="c128")
    assert np.array_equal(f.numpy(), np.full((4, 6), 2.5 + 0j))
    r = RNG.uniform(-1, 1, size=(33,)).astype(np.float32)
    c = nc.astype(nc.asarray(r), "c64")
    same_bits(c.numpy(), r.astype(np.complex64))
    wide = nc.astype(nc.asarray(cfloats((17,))), "c128")
    assert wide.dtype.name == "c128"
    a, b = cfloats((64, 64)), cfloats((64,))
    h = nc.mul(nc.asarray(a, where="host"), nc.asarray(b, where="host"))
    assert h.where == "host"
    same_bits(h.numpy(), CO.binary("*", CO.binary("*", np.int64(1), a), b))
    rnd = nc.random_array((1000,), "c64", 7, 0, -1.0, 1.0).numpy()
    assert rnd.dtype == np.complex64 and np.all(np.abs(rnd.real) <= 1) and np.all(np.abs(rnd.imag) <= 1) and rnd.imag.std() > 0.3


def test_functions_outside_the_complex_whitelist_are_refused(nc):
    g = nc.asarray(cints((8,)))
    with pytest.raises(Exception):
        nc.maximum(g, g)                                                      # complex numbers are not ordered
    with pytest.raises((_lib.NumclUnsupportedError, NotImplementedError, TypeError)):
        nc.sqrt(g)
    out = nc.empty((8,), type="f32")
    with pytest.raises(_lib.NumclCudaError):
        nc.map_array_into(out, "identity", g)                                 # a complex value does not fit a real array
