"""Reference semantics composed from oracle primitives, for the tests only.

The oracle (oracle/) restates numcl's three loop generators; the API functions above them
(numcl:+ as an identity-seeded fold of `broadcast`, numcl:sum as reduce-array with the zero value,
kron as einsum + reshape) are composed here exactly as the reference composes them, with element
types decided by numcl_b200.typeinfer (the host mirror of 2typeinfer.lisp, pinned by the golden
type assertions)."""
import numpy as np

from numcl_b200 import typeinfer as ti
from numcl_b200.einsum import transform_to_typeform
from numcl_b200.einsum_spec import einsum_normalize_subscripts


def _is_scalar(x):
    return isinstance(x, (int, float, np.integer, np.floating))


def oarr(O, x):
    if _is_scalar(x):
        return O.OArr.from_values(np.array(x), ti.upgrade(ti.strict_type_of(x)))     # (full nil x)
    return x


def typ(x):
    return ti.strict_type_of(x) if _is_scalar(x) else ti.type_of_dtype(x.dtype)


def broadcast(O, fn, x, y):
    """(broadcast fn x y), src/4numeric.lisp:201-294."""
    rt = ti.upgrade(ti.infer_type(fn, typ(x), typ(y)))
    return O.broadcast(fn, oarr(O, x), oarr(O, y), rt)


def fold(O, fn, args, use_identity):
    """numcl:+ * - / max min, src/4numeric.lisp:528-542: a chain of binary passes."""
    args = list(args)
    if use_identity:
        cur = 1 if fn in ("*", "/") else 0
    else:
        cur = args.pop(0)
    for a in args:
        cur = broadcast(O, fn, cur, a)
    return cur


def arith(O, op, args):
    if op in ("+", "*"):
        return fold(O, op, args, True)
    if op in ("-", "/"):
        return fold(O, op, args, len(args) == 1)
    return fold(O, op, args, False)


def einsum(O, spec, inputs, outputs=None):
    specs = einsum_normalize_subscripts(spec)
    n_out = len(specs.o_specs)
    if outputs:
        return O.einsum(spec, *inputs, out=list(outputs))
    o_types = ti.einsum_output_types([transform_to_typeform(t) for t in specs.transforms],
                                     [ti.type_of_dtype(a.dtype) for a in inputs], n_out)
    return O.einsum(spec, *inputs, out_dtypes=[ti.upgrade(t) for t in o_types])


_IDENT = {"+": ti.zero_value, "*": ti.one_value, "max": ti.most_negative_value, "min": ti.most_positive_value}


def reduce_array(O, fn, a, axes=None, initial="numcl"):
    """reduce-array with the identities numcl:sum/prod/amax/amin pass, src/5reduce.lisp:42-126.
    initial="zero" gives bare reduce-array's default (zero-value type)."""
    f = fn.lower().split(":")[-1]
    rt = ti.reduce_result_type(f, ti.type_of_dtype(a.dtype))
    init = ti.zero_value(rt) if initial == "zero" else _IDENT[f](rt)
    return O.reduce_array(f, a, axes, ti.upgrade(rt), init)
