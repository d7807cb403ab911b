"""numcl:+ - * / max min, broadcast, map-array on the CUDA backend (src/4numeric.lisp).

In the reference these never reach `einsum-body`: `broadcast` (:201-399) and `map-array-into`
(:37-113) carry their own loop generators.  The `:cuda` glue re-routes them (lisp/cuda.lisp) to
`ncl_broadcast` / `ncl_fold` / `ncl_map`; this module is the Python mirror of that glue with the
reference's names, argument meaning and result element types.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from . import typeinfer as ti
from .array import NArray, empty, full
from .einsum import compile_transform, ensure_singleton
from .einsum_spec import read_subscripts

_OP = {"+": "ADD", "-": "SUB", "*": "MUL", "/": "DIV", "max": "MAX", "min": "MIN",
       "=/bit": "EQ", "/=/bit": "NE", "<=/bit": "LE", ">=/bit": "GE", "</bit": "LT", ">/bit": "GT",
       "logand": "LOGAND", "logior": "LOGIOR", "logxor": "LOGXOR",
       "logandc1": "LOGANDC1", "logandc2": "LOGANDC2", "logeqv": "LOGEQV", "lognand": "LOGNAND", "lognor": "LOGNOR",
       "logorc1": "LOGORC1", "logorc2": "LOGORC2",
       "floor": "FLOOR", "ceiling": "CEILING", "round": "ROUND", "truncate": "TRUNCATE", "mod": "MOD", "rem": "REM",
       "expt": "EXPT"}


def _is_scalar(x):
    return isinstance(x, (int, float, complex, np.integer, np.floating, np.complexfloating)) and not isinstance(x, bool)


def _scalar_op(fn, x, y):
    import math
    if fn in ("floor", "ceiling", "truncate", "round") and isinstance(x, int) and isinstance(y, int):
        from fractions import Fraction
        q = Fraction(x, y)
        return {"floor": math.floor, "ceiling": math.ceil, "truncate": math.trunc, "round": round}[fn](q)
    if fn == "mod":
        return x % y                          # sign of the divisor, like CL's mod
    if fn == "rem":
        return math.fmod(x, y) if isinstance(x, float) or isinstance(y, float) else int(math.fmod(x, y))
    if fn == "expt":
        return x ** y
    if fn == "+":
        return x + y
    if fn == "-":
        return x - y
    if fn == "*":
        return x * y
    if fn == "/":
        return x / y
    if fn == "max":
        return max(x, y)
    if fn == "min":
        return min(x, y)
    raise NotImplementedError(fn)


def _as_array(x, like=None) -> NArray:
    """Scalars are wrapped as rank-0 arrays typed by strict-type-of (src/4numeric.lisp:211-214)."""
    if isinstance(x, NArray):
        return x
    where = like.where if like is not None else "device"
    if isinstance(x, (complex, np.complexfloating)):
        from .array import asarray
        v = np.asarray(x)
        return asarray(v.astype(np.complex64) if isinstance(x, complex) else v, where=where)
    return full((), x, type=ti.upgrade(ti.strict_type_of(x)), where=where)


def _type_of(x):
    return ti.type_of_dtype(x.dtype) if isinstance(x, NArray) else ti.strict_type_of(x)


def broadcast_result_shape(xs, ys):
    """broadcast-p + broadcast-result-shape, src/4numeric.lisp:174-199."""
    rr = max(len(xs), len(ys))
    xs = (1,) * (rr - len(xs)) + tuple(xs)
    ys = (1,) * (rr - len(ys)) + tuple(ys)
    for dx, dy in zip(xs, ys):
        if not (dx == dy or dx == 1 or dy == 1):
            raise AssertionError(f"Given arrays have incompatible shape for broadcasting: {xs} and {ys}")
    return tuple(max(a, b) for a, b in zip(xs, ys))


def broadcast(function: str, x, y, type=None):
    """(broadcast function x y &key type), src/4numeric.lisp:201-294: one binary op, numpy-style
    right-aligned broadcasting, result element type from infer-type, stored through %coerce."""
    fn = function.lower().split(":")[-1]
    if _is_scalar(x) and _is_scalar(y):
        return _scalar_op(fn, x, y)
    like = x if isinstance(x, NArray) else y
    rtype = type or ti.upgrade(ti.infer_type(fn, _type_of(x), _type_of(y)))
    xa, ya = _as_array(x, like), _as_array(y, like)
    r = empty(broadcast_result_shape(xa.shape, ya.shape), type=rtype, where=like.where)
    tx, ty, tr = xa.tensor(), ya.tensor(), r.tensor()
    _lib.check(_lib.lib().ncl_broadcast(_lib.OPS[_OP[fn]], C.byref(tx), C.byref(ty), C.byref(tr)))
    return r


def _fold(fn: str, args, use_identity: bool):
    """The n-ary folds of src/4numeric.lisp:528-542 as ONE fused pass (ncl_fold).  Element types
    follow the reference's chain of intermediate arrays: each step's inferred type is upgraded to an
    array element type before the next step sees it."""
    if all(_is_scalar(a) for a in args):
        import functools
        init = [0 if fn in "+-" else 1] if use_identity else []
        return functools.reduce(lambda p, q: _scalar_op(fn, p, q), init + list(args))
    like = next(a for a in args if isinstance(a, NArray))
    ident = ti.strict_type_of(1 if fn in ("*", "/") else 0)
    chain = ([ident] if use_identity else []) + [_type_of(a) for a in args]
    acc = chain[0]
    shape = ()
    for k, t in enumerate(chain[1:]):
        acc = ti.type_of_dtype(ti.upgrade(ti.infer_type(fn, acc, t)))
    shapes = [a.shape if isinstance(a, NArray) else () for a in args]
    shape = shapes[0]
    for sh in shapes[1:]:
        shape = broadcast_result_shape(shape, sh)
    arrs = [_as_array(a, like) for a in args]
    r = empty(shape, type=ti.upgrade(acc), where=like.where)
    ts = (_lib.Tensor * len(arrs))(*[a.tensor() for a in arrs])
    tr = r.tensor()
    try:
        _lib.check(_lib.lib().ncl_fold(_lib.OPS[_OP[fn]], 1 if use_identity else 0, ts, len(arrs), C.byref(tr)))
    except _lib.NumclUnsupportedError:
        # e.g. a chain of integer divisions: run the reference's own sequence of binary passes
        it = iter(args)
        cur = broadcast(fn, 1 if fn in ("*", "/") else 0, next(it)) if use_identity else next(it)
        for a in it:
            cur = broadcast(fn, cur, a)
        return cur
    return r


def add(*args):
    """numcl:+  = (reduce (lambda (x y) (broadcast '+ x y)) args :initial-value 0), :528."""
    return _fold("+", args, True) if args else 0


def mul(*args):
    """numcl:*  = (reduce ... :initial-value 1), :529."""
    return _fold("*", args, True) if args else 1


def maximum(*args):
    """numcl:max = (reduce (lambda (x y) (broadcast 'max x y)) args), :530."""
    return args[0] if len(args) == 1 else _fold("max", args, False)


def minimum(*args):
    return args[0] if len(args) == 1 else _fold("min", args, False)


def sub(first, *args):
    """numcl:- : fold from the first argument; unary is (broadcast '- 0 first), :535-538."""
    if args:
        return _fold("-", (first,) + args, False)
    return broadcast("-", 0, first)


def div(first, *args):
    """numcl:/ : fold from the first argument; unary is (broadcast '/ 1 first), :539-542."""
    if args:
        return _fold("/", (first,) + args, False)
    return broadcast("/", 1, first)


def clip(array, lo, hi):
    """(broadcast 'max min (broadcast 'min array max)), :544."""
    return broadcast("max", lo, broadcast("min", array, hi))


def one_plus(array):
    return add(array, 1)


def one_minus(array):
    return sub(array, 1)


def _cmp(name):
    def f(x, y):
        if _is_scalar(x) and _is_scalar(y):
            return {"=/bit": x == y, "/=/bit": x != y, "<=/bit": x <= y, ">=/bit": x >= y, "</bit": x < y, ">/bit": x > y}[name]
        return broadcast(name, x, y)
    return f


eq, ne, le, ge, lt, gt = (_cmp(n) for n in ("=/bit", "/=/bit", "<=/bit", ">=/bit", "</bit", ">/bit"))


def logand(x, y):
    return broadcast("logand", x, y)


def logior(x, y):
    return broadcast("logior", x, y)


def logxor(x, y):
    return broadcast("logxor", x, y)


# ---- map-array -------------------------------------------------------------------------------------------
def _fn_form(function, n):
    """'cl:+ or "+" -> (+ $1 ... $n); a full form with $k variables is passed through."""
    if isinstance(function, str) and "(" in function:
        forms = read_subscripts(function)
        return forms[0]
    name = function.split(":")[-1]
    from .einsum_spec import Sym
    return [Sym(name.upper())] + [Sym(f"${i + 1}") for i in range(n)]


def map_array_into(result: NArray, function, *arrays: NArray) -> NArray:
    """(map-array-into result function arrays...), src/4numeric.lisp:37-51: same-shape n-ary map,
    each value stored through %coerce into the result's element type."""
    for a in arrays:
        if tuple(a.shape) != tuple(result.shape):
            raise AssertionError("map-array-into: (assert (equal (shape sequence) result-shape)) failed")
    e = compile_transform(_fn_form(function, len(arrays)))
    ts = (_lib.Tensor * len(arrays))(*[a.tensor() for a in arrays])
    tr = result.tensor()
    _lib.check(_lib.lib().ncl_map(C.byref(e), ts, len(arrays), C.byref(tr)))
    return result


def map_array(function, *arrays: NArray) -> NArray:
    """(map-array function arrays...), :117-122: result type = infer-type over the element types."""
    from .einsum import transform_to_typeform
    form = _fn_form(function, len(arrays))
    env = {f"${i + 1}": ti.type_of_dtype(a.dtype) for i, a in enumerate(arrays)}
    rtype = ti.upgrade(ti.interpret_type(transform_to_typeform(form), env))
    return map_array_into(empty(arrays[0].shape, type=rtype, where=arrays[0].where), function, *arrays)


def _simple_mapper(cl_fn):
    """define-simple-mapper, src/4numeric.lisp:455-465: (einsum '(i -> (f $1) -> i) (flatten x) (flatten y))."""
    def mapper(x):
        if _is_scalar(x):
            import math
            return {"sqrt": math.sqrt, "exp": math.exp, "log": math.log, "%logr": math.log, "%log2": math.log2, "sin": math.sin,
                    "cos": math.cos, "tan": math.tan, "tanh": math.tanh, "abs": abs, "square": lambda v: v * v,
                    "asin": math.asin, "acos": math.acos, "atan": math.atan, "sinh": math.sinh, "cosh": math.cosh,
                    "asinh": math.asinh, "acosh": math.acosh, "atanh": math.atanh, "isqrt": math.isqrt,
                    "lognot": lambda v: ~v, "signum": lambda v: (v > 0) - (v < 0),
                    "conjugate": lambda v: v.conjugate(), "realpart": lambda v: v.real, "imagpart": lambda v: v.imag}[cl_fn](x)
        from .einsum import einsum
        t = ti.infer_type("%square" if cl_fn == "square" else cl_fn, ti.type_of_dtype(x.dtype))
        y = empty(x.shape, type=ti.upgrade(t), where=x.where)
        f = "%square" if cl_fn == "square" else cl_fn
        einsum(f"i -> ({f} $1) -> i", x.flatten(), y.flatten())
        return y
    mapper.__name__ = cl_fn
    return mapper


sqrt, exp, log, sin, cos, tan, tanh, abs_, square = (_simple_mapper(n) for n in
                                                     ("sqrt", "exp", "log", "sin", "cos", "tan", "tanh", "abs", "square"))
# the rest of src/4numeric.lisp:466-519.  numcl:log is %logr (positive reals assumed), numcl:logc is CL's log, numcl:log2 is
# %log2; logcount / integer-length have no type inferer in the reference (2typeinfer.lisp:484-497 is commented out), so the
# mappers there produce T arrays: use map_array_into with a typed result for those two.
asin, acos, atan, sinh, cosh, asinh, acosh, atanh, isqrt, lognot = (
    _simple_mapper(n) for n in ("asin", "acos", "atan", "sinh", "cosh", "asinh", "acosh", "atanh", "isqrt", "lognot"))
# complex numbers, src/4numeric.lisp:486-489
conjugate, realpart, imagpart = (_simple_mapper(n) for n in ("conjugate", "realpart", "imagpart"))
logc = _simple_mapper("log")
log = _simple_mapper("%logr")
log2 = _simple_mapper("%log2")


def _binary(name):
    def f(x, y):
        return broadcast(name, x, y)
    f.__name__ = name
    return f


def _divlike(name):
    def f(number, divisor=1):
        """(numcl:NAME number &optional (divisor 1)) = (broadcast 'NAME number divisor), src/4numeric.lisp:572-577"""
        return broadcast(name, number, divisor)
    f.__name__ = name
    return f


floor, ceiling, round_, truncate, mod, rem = (_divlike(n) for n in ("floor", "ceiling", "round", "truncate", "mod", "rem"))
expt = _binary("expt")
logandc1, logandc2, logeqv, lognand, lognor, logorc1, logorc2 = (
    _binary(n) for n in ("logandc1", "logandc2", "logeqv", "lognand", "lognor", "logorc1", "logorc2"))
