"""The complex checker itself (oracle/complex_oracle.py) against numpy's complex arithmetic on the CPU: exactly representable
data must agree bit for bit (+ - * are then order-independent and exact), division within rounding, and the real-meets-complex
rules must leave the untouched part untouched (signed zeros included)."""
import numpy as np
import pytest

from oracle import complex_oracle as CO

RNG = np.random.default_rng(7)


def cints(shape, ct):
    return (RNG.integers(-8, 9, size=shape) + 1j * RNG.integers(-8, 9, size=shape)).astype(ct)


@pytest.mark.parametrize("ct", [np.complex64, np.complex128])
@pytest.mark.parametrize("op,f", [("+", np.add), ("-", np.subtract), ("*", np.multiply)])
def test_exact_data_agrees_with_numpy(ct, op, f):
    a, b = cints((50, 7), ct), cints((7,), ct)
    got = CO.binary(op, a, b)
    assert got.dtype == ct and np.array_equal(got, f(a, b).astype(ct))


@pytest.mark.parametrize("ct,tol", [(np.complex64, 2e-6), (np.complex128, 1e-14)])
def test_division_is_smiths_method_within_rounding(ct, tol):
    a = (RNG.uniform(-1, 1, 500) + 1j * RNG.uniform(-1, 1, 500)).astype(ct)
    b = (RNG.uniform(0.5, 2, 500) * np.exp(1j * RNG.uniform(0, 6.28, 500))).astype(ct)
    got = CO.binary("/", a, b)
    want = a.astype(np.complex128) / b.astype(np.complex128)
    assert np.max(np.abs(got - want) / np.abs(want)) <= tol
    assert np.array_equal(CO.binary("/", cints((9,), ct) * 2, np.float64(2.0) if ct == np.complex128 else np.float32(2.0)).real % 1, np.zeros(9))


def test_a_real_operand_meets_only_the_real_part():
    z = np.empty(3, dtype=np.complex64)
    z.real = [1.0, -0.0, 2.5]; z.imag = [-0.0, -0.0, 3.0]
    s = CO.binary("+", np.int64(0), z)                      # (+ 0 #C(a b)) = #C((+ 0 a) b): -0.0 real becomes +0.0, the imaginary part is untouched
    assert np.signbit(s.imag[0]) and np.signbit(s.imag[1]) and not np.signbit(s.real[1])
    m = CO.binary("*", np.float32(2.0), z)
    assert np.array_equal(m.real, 2 * z.real) and np.array_equal(m.imag, 2 * z.imag)
    d = CO.binary("-", np.float32(1.0), z)
    assert np.array_equal(d.imag, -z.imag)


def test_contagion_of_the_parts():
    a = cints((4,), np.complex64)
    assert CO.binary("+", a, np.float64(1.0)).dtype == np.complex128
    assert CO.binary("*", a, np.int64(3)).dtype == np.complex64
    assert CO.binary("+", a, a.astype(np.complex128)).dtype == np.complex128


@pytest.mark.parametrize("ct", [np.complex64, np.complex128])
def test_matmul_in_the_reference_order(ct):
    A, B = cints((6, 9), ct), cints((9, 5), ct)
    assert np.array_equal(CO.matmul_reference_order(A, B), (A @ B).astype(ct))
