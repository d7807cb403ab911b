"""numcl:einsum on the CUDA backend -- host side of src/3einsum.lisp:24-37,470-518.

What `einsum-lambda` does around the backend hook is restated here for the Python host mirror:
resolve the loop limits from the array shapes (shape-resolver, :287-367), decide output element
types (einsum-output-types, :495-518), create missing outputs (%output-generator, :439-468), then
hand the serialised `einsum-specs` to the backend -- `ncl_einsum`, the call the `:cuda` method of
`einsum-body` emits -- and return (values (ensure-singleton O0) ...).
"""
from __future__ import annotations

import ctypes as C
import functools

from . import _lib
from . import typeinfer as ti
from .array import NArray, empty
from .einsum_spec import Double, EinsumSpecError, EinsumSpecs, Sym, einsum_normalize_subscripts

_BINARY = {"+": "ADD", "-": "SUB", "*": "MUL", "/": "DIV", "MAX": "MAX", "MIN": "MIN",
           "=/BIT": "EQ", "/=/BIT": "NE", "<=/BIT": "LE", ">=/BIT": "GE", "</BIT": "LT", ">/BIT": "GT",
           "LOGAND": "LOGAND", "LOGIOR": "LOGIOR", "LOGXOR": "LOGXOR",
           "LOGANDC1": "LOGANDC1", "LOGANDC2": "LOGANDC2", "LOGEQV": "LOGEQV", "LOGNAND": "LOGNAND", "LOGNOR": "LOGNOR",
           "LOGORC1": "LOGORC1", "LOGORC2": "LOGORC2", "EXPT": "EXPT"}
# (floor number &optional (divisor 1)) ...: one argument means divisor 1 (src/4numeric.lisp:572-577)
_DIVLIKE = {"FLOOR": "FLOOR", "CEILING": "CEILING", "ROUND": "ROUND", "TRUNCATE": "TRUNCATE", "MOD": "MOD", "REM": "REM"}
_UNARY = {"ABS": "ABS", "%SQUARE": "SQUARE", "SQUARE": "SQUARE", "SQRT": "SQRT", "LOGNOT": "LOGNOT",
          "IDENTITY": "IDENTITY", "EXP": "EXP", "LOG": "LOG", "%LOGR": "LOG", "SIN": "SIN", "COS": "COS",
          "TAN": "TAN", "TANH": "TANH", "SIGNUM": "SIGNUM",
          "ASIN": "ASIN", "ACOS": "ACOS", "ATAN": "ATAN", "SINH": "SINH", "COSH": "COSH", "ASINH": "ASINH", "ACOSH": "ACOSH",
          "ATANH": "ATANH", "%LOG2": "LOG2", "LOG2": "LOG2", "ISQRT": "ISQRT", "LOGCOUNT": "LOGCOUNT",
          "INTEGER-LENGTH": "INTEGER_LENGTH", "CONJUGATE": "CONJUGATE", "REALPART": "REALPART", "IMAGPART": "IMAGPART"}


class TransformError(NotImplementedError):
    """The TRANSFORM uses a function outside the CUDA whitelist (no CPU fallback)."""


def compile_transform(form) -> _lib.Expr:
    """Flatten a TRANSFORM form into the post-order node list of include/numcl_cuda.h's ncl_expr.

    N-ary Lisp forms become left-associated binary nodes, except max/min which SBCL nests to the
    right ((max a b c) = (max a (max b c)))."""
    e = _lib.Expr()
    e.n = 0

    def emit(op, a=0, b=0, ival=0, fval=0.0):
        if e.n >= _lib.MAX_EXPR:
            raise TransformError("transform too large")
        nd = e.node[e.n]
        nd.op, nd.a, nd.b, nd.ival, nd.fval = _lib.OPS[op], a, b, ival, fval
        e.n += 1
        return e.n - 1

    def rec(f):
        if isinstance(f, bool):
            raise TransformError("booleans are not numbers")
        if isinstance(f, int):
            return emit("X_CONST_INT", ival=f)
        if isinstance(f, Double):
            return emit("X_CONST_F64", fval=float(f))
        if isinstance(f, float):
            return emit("X_CONST_F32", fval=f)
        if isinstance(f, str):
            name = str(f).upper()
            if name[0] in "$@" and name[1:].isdigit():
                return emit("X_INPUT" if name[0] == "$" else "X_OUTPUT", a=int(name[1:]))
            raise TransformError(f"free variable {f} in transform")
        if not isinstance(f, list) or not f or not isinstance(f[0], str):
            raise TransformError(f"bad transform form {f!r}")
        head = str(f[0]).upper().split(":")[-1]
        args = f[1:]
        if head in ("+", "*"):
            if not args:
                return emit("X_CONST_INT", ival=0 if head == "+" else 1)
            acc = rec(args[0])
            for a in args[1:]:
                acc = emit(_BINARY[head], acc, rec(a))
            return acc
        if head in ("-", "/"):
            if len(args) == 1:
                if head == "-":
                    return emit("NEG", rec(args[0]))
                one = emit("X_CONST_INT", ival=1)
                return emit("DIV", one, rec(args[0]))
            acc = rec(args[0])
            for a in args[1:]:
                acc = emit(_BINARY[head], acc, rec(a))
            return acc
        if head in ("MAX", "MIN"):
            vals = [rec(a) for a in args]          # arguments are evaluated left to right
            acc = vals[-1]
            for v in reversed(vals[:-1]):
                acc = emit(_BINARY[head], v, acc)
            return acc
        if head == "1+":
            return emit("ADD", rec(args[0]), emit("X_CONST_INT", ival=1))
        if head == "1-":
            return emit("SUB", rec(args[0]), emit("X_CONST_INT", ival=1))
        if head in _DIVLIKE:
            if len(args) not in (1, 2):
                raise TransformError(f"{head} takes one or two arguments")
            return emit(_DIVLIKE[head], rec(args[0]), rec(args[1]) if len(args) == 2 else emit("X_CONST_INT", ival=1))
        if head in _BINARY:
            if head.startswith("LOG") and len(args) != 2:
                acc = rec(args[0])
                for a in args[1:]:
                    acc = emit(_BINARY[head], acc, rec(a))
                return acc
            if len(args) != 2:
                raise TransformError(f"{head} takes two arguments here")
            return emit(_BINARY[head], rec(args[0]), rec(args[1]))
        if head in _UNARY:
            if len(args) != 1:
                raise TransformError(f"{head} takes one argument")
            return emit(_UNARY[head], rec(args[0]))
        raise TransformError(f"function {head} is not on the CUDA whitelist (no CPU fallback)")

    rec(form)
    return e


def transform_to_typeform(form):
    """TRANSFORM form -> the nested-list form typeinfer.interpret_type reads."""
    if isinstance(form, list):
        return [str(form[0]).lower().split(":")[-1]] + [transform_to_typeform(a) for a in form[1:]]
    if isinstance(form, Double):
        return ("double-float", float(form), float(form))
    if isinstance(form, str):
        return str(form).lower() if str(form)[0] in "$@" else str(form)
    return form


# ---- shape-resolver (src/3einsum.lisp:274-367) on the host: needed to create the outputs -------------
def _normalize_index(index, dim):
    if index < 0:
        return index % dim if dim else 0
    return min(index, dim)


def _range_width(dim, opt):
    start = opt.start if (opt and opt.start is not None) else 0
    end = opt.end if (opt and opt.end is not None) else dim
    step = opt.step if opt else 1
    return max(0, -((-(_normalize_index(end, dim) - _normalize_index(start, dim))) // step))


def _find_opt(opts, key):
    for o in opts or []:
        if o.key == key:
            return o
    return None


def resolve_shapes(specs: EinsumSpecs, in_shapes, out_shapes=None):
    """Returns ({index: ?n}, plan_out_dims, [output shapes]).  Raises AssertionError like the
    reference's (assert (= ...)) when the dims do not unify."""
    q: dict = {}
    blocks = []

    def unify(idx, w):
        if idx in q and q[idx] != w:
            raise AssertionError(f"einsum: index {idx} is used with sizes {q[idx]} and {w}")
        q[idx] = w

    def one(spec, shape, opts, what):
        if -1 not in spec:
            if len(shape) != len(spec):
                raise AssertionError(f"einsum: {what} of rank {len(shape)} for a spec of {len(spec)} indices")
            for ax, idx in enumerate(spec):
                unify(idx, _range_width(shape[ax], _find_opt(opts, idx)))
            return None
        b = spec.index(-1)
        for ax in range(b):
            unify(spec[ax], _range_width(shape[ax], _find_opt(opts, spec[ax])))
        after = len(spec) - 1 - b
        for i in range(after):
            ax = len(shape) - 1 - i
            idx = spec[len(spec) - 1 - i]
            unify(idx, _range_width(shape[ax], _find_opt(opts, idx)))
        return tuple(shape[b: len(shape) - after])

    for spec, shape, opts in zip(specs.i_specs, in_shapes, specs.i_options):
        blk = one(spec, shape, opts, "input")
        if blk is not None:
            blocks.append(blk)
    plan_out = ()
    if blocks:                                      # plan-broadcast, :378-425
        m = max(len(b) for b in blocks)
        padded = [(1,) * (m - len(b)) + tuple(b) for b in blocks]
        plan_out = tuple(max(p[j] for p in padded) for j in range(m))
        for p in padded:
            for j in range(m):
                if not (p[j] == 1 or p[j] == plan_out[j]):
                    raise AssertionError("einsum: broadcast blocks are incompatible")
    shapes = []
    for k, spec in enumerate(specs.o_specs):
        shp = []
        for idx in spec:
            if idx == -1:
                shp.extend(plan_out)
            else:
                shp.append(q[idx])
        shapes.append(tuple(shp))
        if out_shapes is not None and out_shapes[k] is not None:
            blk = one(spec, out_shapes[k], specs.o_options[k], "output")
            if blk is not None and tuple(blk) != tuple(plan_out):
                raise AssertionError("The output shapes do not match the broadcast plan!")
    return q, plan_out, shapes


# ---- descriptor ------------------------------------------------------------------------------------------
def build_desc(specs: EinsumSpecs) -> _lib.EinsumDesc:
    d = _lib.EinsumDesc()
    d.n_in, d.n_out, d.n_iter = len(specs.i_specs), len(specs.o_specs), len(specs.iter_specs)
    if d.n_in > _lib.MAX_OPERANDS or d.n_out > _lib.MAX_OPERANDS or d.n_iter > _lib.MAX_INDEX:
        raise EinsumSpecError("einsum spec exceeds the backend's limits")
    for i, v in enumerate(specs.iter_specs):
        d.iter[i] = v

    def fill(specs_, ranks, arr, opts_in, nopt, optarr):
        for a, spec in enumerate(specs_):
            if len(spec) > _lib.MAX_RANK:
                raise EinsumSpecError("rank exceeds the backend's limit")
            ranks[a] = len(spec)
            for k, v in enumerate(spec):
                arr[a][k] = v
            n = 0
            for o in opts_in[a] if a < len(opts_in) else []:
                if not isinstance(o.key, int):
                    continue                        # an option keyed by a symbol that is no index: never found
                e = optarr[a][n]
                e.key = o.key
                e.has_start, e.start = (1, o.start) if o.start is not None else (0, 0)
                e.has_end, e.end = (1, o.end) if o.end is not None else (0, 0)
                e.step = o.step
                n += 1
            nopt[a] = n

    fill(specs.i_specs, d.in_rank, d.in_spec, specs.i_options, d.in_nopt, d.in_opt)
    fill(specs.o_specs, d.out_rank, d.out_spec, specs.o_options, d.out_nopt, d.out_opt)
    for k, tr in enumerate(specs.transforms):
        d.transform[k] = compile_transform(tr)
    return d


@functools.lru_cache(maxsize=512)
def _memoized(subscripts_key: str):
    """memoized-einsum-function (src/3einsum.lisp:30-31): one normalisation + descriptor per spec."""
    specs = einsum_normalize_subscripts(subscripts_key)
    return specs, build_desc(specs)


def einsum(subscripts, *args):
    """(numcl:einsum subscripts inputs... [outputs...]), src/3einsum.lisp:24-28.

    Missing outputs are created with the element type einsum-output-types infers and are never
    zero-filled on the host: the backend is told they are fresh.  Supplied outputs are
    accumulated into, as in the reference (:468)."""
    if not isinstance(subscripts, str):
        raise TypeError("subscripts: pass the text of the Lisp list, e.g. 'ij jk -> ik'")
    specs, desc = _memoized(subscripts)
    n_in, n_out = len(specs.i_specs), len(specs.o_specs)
    if len(args) < n_in or len(args) > n_in + n_out:
        raise TypeError(f"einsum: {len(args)} arrays for {n_in} inputs and up to {n_out} outputs")
    ins = list(args[:n_in])
    outs = list(args[n_in:]) + [None] * (n_in + n_out - len(args))
    for a in list(ins) + [o for o in outs if o is not None]:
        if not isinstance(a, NArray):
            raise TypeError("einsum inputs must be NArray")
        # einsum-lambda keeps only the base vector (src/3einsum.lisp:485-487): a displaced view would
        # silently be read from element 0.  Fenced rather than reproduced.
        assert a.offset == 0, "einsum on a displaced view: the reference drops the displacement offset"
    _, _, shapes = resolve_shapes(specs, [a.shape for a in ins], [o.shape if o is not None else None for o in outs])
    fresh = all(o is None for o in outs)
    if not fresh and any(o is None for o in outs):
        # (or O (zeros ...)): a mix of supplied and missing outputs -- materialise the missing as zeros
        from .array import zeros
        missing_as_zeros = zeros
    else:
        missing_as_zeros = None
    if any(o is None for o in outs):
        in_types = [ti.type_of_dtype(a.dtype) for a in ins]
        o_types = ti.einsum_output_types([transform_to_typeform(t) for t in specs.transforms], in_types, n_out)
        where = ins[0].where if ins else "device"
        for k in range(n_out):
            if outs[k] is None:
                dt = ti.upgrade(o_types[k])
                outs[k] = missing_as_zeros(shapes[k], type=dt, where=where) if missing_as_zeros else empty(shapes[k], type=dt, where=where)
    tin = (_lib.Tensor * max(1, n_in))(*[a.tensor() for a in ins])
    tout = (_lib.Tensor * n_out)(*[o.tensor() for o in outs])
    L = _lib.lib()
    _lib.check(L.ncl_einsum(C.byref(desc), tin, n_in, tout, n_out, _lib.OUT_FRESH if fresh else _lib.OUT_ACCUMULATE))
    res = tuple(ensure_singleton(o) for o in outs)
    return res[0] if n_out == 1 else res


def ensure_singleton(a: NArray):
    """src/2aref.lisp:177-182: a rank-0 result is returned as a bare number."""
    return a.item() if a.shape == () else a
