"""CPU restatement of complex arithmetic for the parity tests (TEST INFRASTRUCTURE: only tests/ may import this).

numcl has no complex code of its own: a (complex single-float) array goes through the same loop generators as every other
array (src/4einsum-backends/common-lisp.lisp:226-240, src/4numeric.lisp:228-293, src/5reduce.lisp:67-87) and the leaf calls
CL's generic + - * / on two numbers, i.e. SBCL's number dispatch (src/code/numbers.lisp of SBCL, a third-party dependency
that is not under /root/reference and not version-pinned).  What is restated here is that dispatch as published:

  complex o complex:  + - componentwise;  (a+bi)(c+di) = (ac - bd) + (ad + bc)i, four products and two sums each rounded;
                      /  by Smith's method: scale by the larger part of the divisor
  real o complex:     the real meets only the parts it takes part in:  x + (a+bi) = (x+a) + bi,  x * (a+bi) = xa + xb i,
                      (a+bi) / x = a/x + (b/x)i;  x / (a+bi) is the general quotient with a zero imaginary part
  contagion:          the part types follow float contagion (src/1type.lisp:446-451); integers become floats of that type

PARITY UNPINNED: the reference's tests hold no complex known-answer vector for this path (t/package.lisp:742 only checks
an element TYPE), and no Lisp exists in this image; + - * on exactly representable data are order-independent and exact,
everything else is checked to tolerance."""
import numpy as np


def _parts(x, ft):
    x = np.asarray(x)
    if np.iscomplexobj(x):
        return True, x.real.astype(ft), x.imag.astype(ft)
    return False, x.astype(ft), None


def result_float(x, y):
    """part type of the result: double if either side is double precision"""
    def prec(v):
        v = np.asarray(v)
        return 64 if v.dtype in (np.float64, np.complex128) else 32
    return np.float64 if max(prec(x), prec(y)) == 64 else np.float32


def binary(op, x, y):
    ft = result_float(x, y)
    xc, xr, xi = _parts(x, ft)
    yc, yr, yi = _parts(y, ft)
    xr, yr = np.broadcast_arrays(xr, yr)
    shape = xr.shape
    xi = np.broadcast_to(xi, shape) if xc else None
    yi = np.broadcast_to(yi, shape) if yc else None
    with np.errstate(all="ignore"):
        if op == "+":
            rr = xr + yr
            ri = (xi + yi) if xc and yc else (xi if xc else yi)
        elif op == "-":
            rr = xr - yr
            ri = (xi - yi) if xc and yc else (xi if xc else -yi)
        elif op == "*":
            if xc and yc:
                rr = (xr * yr) - (xi * yi)
                ri = (xr * yi) + (xi * yr)
            elif xc:
                rr, ri = xr * yr, xi * yr
            else:
                rr, ri = xr * yr, xr * yi
        elif op == "/":
            if not yc:
                rr, ri = xr / yr, xi / yr
            else:
                if not xc:
                    xi = np.zeros(shape, ft)
                big = np.abs(yr) > np.abs(yi)
                r1 = yi / yr
                d1 = yr + r1 * yi
                r2 = yr / yi
                d2 = yi + r2 * yr
                rr = np.where(big, (xr + xi * r1) / d1, (xr * r2 + xi) / d2)
                ri = np.where(big, (xi - xr * r1) / d1, (xi * r2 - xr) / d2)
        else:
            raise ValueError(op)
    out = np.empty(shape, np.complex64 if ft == np.float32 else np.complex128)
    out.real, out.imag = rr, ri
    return out


def matmul_reference_order(a, b):
    """(einsum '(ij jk -> ik) a b) on complex arrays: for every output the contraction index ascends, the leaf is
    @1 <- (+ @1 (* $1 $2)) starting from #C(0 0)"""
    acc = np.zeros((a.shape[0], b.shape[1]), np.result_type(a, b))
    for j in range(a.shape[1]):
        acc = binary("+", acc, binary("*", a[:, j:j + 1], b[j:j + 1, :]))
    return acc


def sum_reference_order(a, axes):
    """(numcl:sum a :axes axes): sequential left-to-right componentwise sums (exact data: any order)"""
    return a.sum(axis=axes)
