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
    """(einsum '(i i -> ) (conjugate a) b), :54-59.  On reals conjugate is the identity; for complex vectors the
    conjugation is folded into the transform (one pass instead of the reference's mapper pass + einsum, same bits:
    conjugation is exact)."""
    if getattr(a, "dtype", None) is not None and a.dtype.kind in ("c64", "c128"):
        spec = "i i -> (+ @1 (* (conjugate $1) $2)) -> "
        return einsum(spec, a, b, result) if result is not None else einsum(spec, a, b)
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


def matrix_chain_order(dims):
    """matrix-chain-order, src/4linear-algebra3.lisp:179-215: the classic O(n^3) DP over
    single-float costs; index[i][j] = split point of the product of matrices i..j."""
    import numpy as np
    n = len(dims) - 1
    d = [np.float32(x) for x in dims]
    cost = [[np.float32(0.0)] * n for _ in range(n)]
    index = [[0] * n for _ in range(n)]
    for length in range(2, n + 1):
        for i in range(0, n - length + 1):
            j = i + length - 1
            cost[i][j] = np.float32(3.4028235e38)
            for k in range(i, j):
                c = np.float32(cost[i][k] + cost[k + 1][j] + d[i] * d[k + 1] * d[j + 1])
                if c < cost[i][j]:
                    cost[i][j] = c
                    index[i][j] = k
    return index


def matmul_star(*args):
    """matmul*, src/4linear-algebra3.lisp:152-177: multiply a chain of matrices in the order that
    minimises the intermediate sizes; every product is an einsum '(ij jk -> ik) on the backend."""
    n = len(args)
    if n == 1:
        return args[0]
    dims = [args[0].shape[0]] + [a.shape[1] for a in args]
    order = matrix_chain_order(dims)

    def rec(i, j):
        if i == j:
            return args[i]
        k = order[i][j]
        return matmul(rec(i, k), rec(k + 1, j))
    return rec(0, n - 1)
