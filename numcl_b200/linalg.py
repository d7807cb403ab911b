"""matmul, transpose, inner, vdot, outer, kron, diag: the one-line einsum wrappers of
src/4linear-algebra2.lisp:38-95, with the same specs."""
from __future__ import annotations

from .array import NArray
from .einsum import einsum


def transpose(matrix: NArray, result: NArray | None = None):
    """Reverses the axes: (einsum `(,indices -> ,(reverse indices)) matrix), :38-45."""
    idx = "abcdefgh"[: matrix.rank]
    spec = f"{idx} -> {idx[::-1]}"
    return einsum(spec, matrix, result) if result is not None else einsum(spec, matrix)


def matmul(a: NArray, b: NArray, result: NArray | None = None):
    """(einsum '(ij jk -> ik) a b), :47-52."""
    return einsum("ij jk -> ik", a, b, result) if result is not None else einsum("ij jk -> ik", a, b)


def inner(a, b, result=None):
    """(einsum '(i i -> ) a b), :61-66."""
    return einsum("i i -> ", a, b, result) if result is not None else einsum("i i -> ", a, b)


def vdot(a, b, result=None):
    """(einsum '(i i -> ) (conjugate a) b), :54-59; reals only, so conjugate is the identity."""
    return inner(a, b, result)


def outer(a, b, result=None):
    """(einsum '(i j -> ij) a b), :68-73."""
    return einsum("i j -> ij", a, b, result) if result is not None else einsum("i j -> ij", a, b)


def kron(a, b, result=None):
    """(reshape (einsum '(ij kl -> ikjl) a b) (mapcar #'* (shape a) (shape b))), :75-81."""
    r = einsum("ij kl -> ikjl", a, b, result) if result is not None else einsum("ij kl -> ikjl", a, b)
    return r.reshape([x * y for x, y in zip(a.shape, b.shape)])


def diag(a, result=None):
    """(einsum '(ii -> i) a), :90-95."""
    return einsum("ii -> i", a, result) if result is not None else einsum("ii -> i", a)


def eye(n, type="bit", where="device"):
    """eye (src/4linear-algebra2.lisp:143-150), a host-side constructor."""
    import numpy as np
    from .array import asarray
    return asarray(np.eye(n, dtype=np.int64), type=type, where=where)
