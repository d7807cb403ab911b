"""Host mirror of numcl's element-type inference for the hot path.

In a real deployment this stays in Lisp (src/2typeinfer.lisp, src/1type.lisp; BASELINE north_star:
"2typeinfer.lisp still decides result element types") and the C ABI just receives dtype enums.  The
Python host mirror needs the same decisions to create result arrays, so the parts the hot path
uses are restated here: interval arithmetic on (head lo hi) type specifiers
(src/1type.lisp:193-291), float contagion / substitution (:358-483),
infer-rational-arithmetic-result with the fixnum clip (:575-621), the per-op inferers
(src/2typeinfer.lisp:105-151, 501-660), einsum-output-types (src/3einsum.lisp:495-518),
%reduce-array-result-type (src/5reduce.lisp:34-40) and SBCL's upgraded-array-element-type.

A type is a tuple (head, lo, hi); head in {"integer", "single-float", "double-float"};
lo/hi are numbers or None for `*`.
"""
from __future__ import annotations

import math

from .dtypes import DTYPES, dtype

MOST_POSITIVE_FIXNUM = (1 << 62) - 1     # SBCL x86-64 [unverified here]
MOST_NEGATIVE_FIXNUM = -(1 << 62)
DEFAULT_FLOAT = "single-float"           # +numcl-default-float-format+, src/1constants.lisp:23-25

_RANK = {"integer": 0, "single-float": 1, "double-float": 2}
# complex element types: the head is "complex-<part>", bounds are always None.  Only (complex *-float) arrays exist
# (src/1type.lisp:341-344); contagion works on the part types (src/1type.lisp:446-451).
COMPLEX_PREFIX = "complex-"


def is_complex(head) -> bool:
    return isinstance(head, str) and head.startswith(COMPLEX_PREFIX)


def complex_part(head) -> str:
    return head[len(COMPLEX_PREFIX):] if is_complex(head) else head


class TypeInferenceError(TypeError):
    pass


def type_of_dtype(dt) -> tuple:
    """array-element-type of an array, as the type specifier numcl's inferers see."""
    dt = dtype(dt)
    if dt.kind == "f32":
        return ("single-float", None, None)
    if dt.kind == "f64":
        return ("double-float", None, None)
    if dt.kind == "c64":
        return (COMPLEX_PREFIX + "single-float", None, None)
    if dt.kind == "c128":
        return (COMPLEX_PREFIX + "double-float", None, None)
    return ("integer", dt.lo, dt.hi)


def strict_type_of(x) -> tuple:
    """strict-type-of, src/1type.lisp:725-744: (integer x x) / (single-float x x) ..."""
    if isinstance(x, bool):
        raise TypeInferenceError("booleans are not numbers")
    if isinstance(x, int):
        return ("integer", x, x)
    try:
        import numpy as np
        if isinstance(x, np.floating):
            return ("double-float" if x.dtype == np.float64 else "single-float", float(x), float(x))
        if isinstance(x, np.integer):
            return ("integer", int(x), int(x))
    except ImportError:  # pragma: no cover
        pass
    if isinstance(x, float):
        return ("single-float", x, x)    # the reader's default float format
    if isinstance(x, complex):
        return (COMPLEX_PREFIX + "single-float", None, None)
    try:
        import numpy as np
        if isinstance(x, np.complexfloating):
            return (COMPLEX_PREFIX + ("double-float" if x.dtype == np.complex128 else "single-float"), None, None)
    except ImportError:  # pragma: no cover
        pass
    raise TypeInferenceError(f"unsupported scalar {x!r}")


# SBCL's specialized array element types in upgrade order [SBCL x86-64, unverified here]
_UPGRADE_ORDER = ["bit", "ub2", "ub4", "ub7", "ub8", "ub15", "ub16", "ub31", "ub32", "ub62", "ub63", "ub64",
                  "sb8", "sb16", "sb32", "fixnum", "sb64"]


def upgrade(t) -> str:
    """upgraded-array-element-type: the first specialized type that contains t."""
    head, lo, hi = t
    if head == "single-float":
        return "f32"
    if head == "double-float":
        return "f64"
    if is_complex(head):
        return "c128" if complex_part(head) == "double-float" else "c64"
    if head != "integer":
        raise TypeInferenceError(f"no specialized array for {t}")
    if lo is None or hi is None:
        raise TypeInferenceError("unbounded integer type upgrades to T (general array): unsupported")
    for name in _UPGRADE_ORDER:
        d = DTYPES[name]
        if d.lo <= lo and hi <= d.hi:
            return name
    raise TypeInferenceError(f"integer range {lo}..{hi} exceeds 64 bits")


# ---- interval helpers (None = unbounded) ----------------------------------------------------
def _lo(x):
    return -math.inf if x is None else x


def _hi(x):
    return math.inf if x is None else x


def _post(lo, hi):
    def p(v, neg):
        if isinstance(v, float) and (math.isinf(v) or math.isnan(v)):
            return None
        return v
    return p(lo, True), p(hi, False)


def _mul(a, b):
    if (a == 0 and isinstance(b, float) and math.isinf(b)) or (b == 0 and isinstance(a, float) and math.isinf(a)):
        return math.nan
    return a * b


def interval_add(l1, h1, l2, h2):
    return _post(_lo(l1) + _lo(l2), _hi(h1) + _hi(h2))


def interval_sub(l1, h1, l2, h2):
    return _post(_lo(l1) - _hi(h2), _hi(h1) - _lo(l2))


def interval_mul(l1, h1, l2, h2):
    c = [_mul(_lo(l1), _lo(l2)), _mul(_lo(l1), _hi(h2)), _mul(_hi(h1), _lo(l2)), _mul(_hi(h1), _hi(h2))]
    # a NaN corner (0 * inf) means the bound is unbounded on the side it could grow (1type.lisp:238-275)
    los = [-math.inf if isinstance(v, float) and math.isnan(v) else v for v in c]
    his = [math.inf if isinstance(v, float) and math.isnan(v) else v for v in c]
    return _post(min(los), max(his))


def interval_max(l1, h1, l2, h2):
    return _post(max(_lo(l1), _lo(l2)), max(_hi(h1), _hi(h2)))


def interval_min(l1, h1, l2, h2):
    return _post(min(_lo(l1), _lo(l2)), min(_hi(h1), _hi(h2)))


def interval_div(l1, h1, l2, h2):
    # only the head type matters to upgrade(); keep the bounds open (a superset of the reference's)
    return None, None


def float_substitution(t, int_result="integer"):
    head = t[0]
    if head == "integer":
        return int_result
    return head


def float_contagion(t1, t2, int_int_result="integer"):
    a, b = float_substitution(t1), float_substitution(t2)
    if is_complex(a) or is_complex(b):                  # (complex <contagion of the parts>); an integer part becomes a float
        pa, pb = complex_part(a), complex_part(b)
        part = pa if _RANK[pa] >= _RANK[pb] else pb
        return COMPLEX_PREFIX + (DEFAULT_FLOAT if part == "integer" else part)
    if a == "integer" and b == "integer":
        return float_substitution((int_int_result, None, None))
    return a if _RANK[a] >= _RANK[b] else b


def _coerce_bound(v, head):
    if v is None:
        return None
    if head == "integer":
        return v
    return float(v)


def infer_rational_arithmetic_result(interval_op, types, int_int_result="integer"):
    """src/1type.lisp:575-621: fold the types pairwise; integer bounds are clipped to fixnums."""
    prev = types[0]
    for now in types[1:]:
        head = float_contagion(prev, now, int_int_result)
        if is_complex(head):
            if interval_op in (interval_max, interval_min):
                raise TypeInferenceError("max / min of complex numbers")
            prev = (head, None, None)
            continue
        lo, hi = interval_op(prev[1], prev[2], now[1], now[2])
        if head == "integer":
            lo = MOST_NEGATIVE_FIXNUM if lo is None else max(lo, MOST_NEGATIVE_FIXNUM)
            hi = MOST_POSITIVE_FIXNUM if hi is None else min(hi, MOST_POSITIVE_FIXNUM)
            prev = ("integer", int(lo), int(hi))
        else:
            prev = (head, _coerce_bound(lo, head), _coerce_bound(hi, head))
    return prev


def _bitwise(kind):
    def infer(x, y=None):
        def ilen(v):
            return None if v is None else int(v).bit_length() if v >= 0 else int(~v).bit_length()
        def width(t):
            a, b = ilen(t[1]), ilen(t[2])
            return None if a is None or b is None else max(a, b)
        if x[0] != "integer" or (y is not None and y[0] != "integer"):
            raise TypeInferenceError("bitwise functions need integer arrays")
        nx = _lo(x[1]) < 0
        if kind == "lognot":
            w = width(x)
            signed = (~x[1] < 0 if x[1] is not None else True) or (~x[2] < 0 if x[2] is not None else True)
        else:
            ny = _lo(y[1]) < 0
            wx, wy = width(x), width(y)
            if kind == "logior":
                signed, w = nx or ny, (None if wx is None or wy is None else max(wx, wy))
            elif kind == "logand":
                signed = nx and ny
                w = wy if wx is None else wx if wy is None else min(wx, wy)
            else:  # logxor / logeqv
                signed, w = True, (None if wx is None or wy is None else max(wx, wy))
        if w is None:
            return ("integer", MOST_NEGATIVE_FIXNUM, MOST_POSITIVE_FIXNUM)
        if signed:                                   # (signed-byte w): the reference does not add the sign bit
            w = max(w, 1)
            return ("integer", -(1 << (w - 1)), (1 << (w - 1)) - 1)
        return ("integer", 0, (1 << w) - 1) if w > 0 else ("integer", 0, 0)
    return infer


def _irrational(t):
    head = float_substitution(t, int_result=DEFAULT_FLOAT)
    return (head, None, None)


_BIT = ("integer", 0, 1)
_INFERERS = {
    "+": lambda *ts: infer_rational_arithmetic_result(interval_add, list(ts)),
    "*": lambda *ts: infer_rational_arithmetic_result(interval_mul, list(ts)),
    "max": lambda *ts: infer_rational_arithmetic_result(interval_max, list(ts)),
    "min": lambda *ts: infer_rational_arithmetic_result(interval_min, list(ts)),
    "=/bit": lambda *ts: _BIT, "/=/bit": lambda *ts: _BIT, "<=/bit": lambda *ts: _BIT,
    ">=/bit": lambda *ts: _BIT, "</bit": lambda *ts: _BIT, ">/bit": lambda *ts: _BIT,
    "logior": _bitwise("logior"), "logand": _bitwise("logand"), "logxor": _bitwise("logxor"),
    "logeqv": _bitwise("logxor"), "lognot": _bitwise("lognot"),
    "sqrt": _irrational, "exp": _irrational, "log": _irrational, "%logr": _irrational,
    "sin": _irrational, "cos": _irrational, "tan": _irrational, "tanh": _irrational,
    "identity": lambda t: t,
}


def _infer_sub(first, *rest):            # src/2typeinfer.lisp:114-120
    if rest:
        return infer_rational_arithmetic_result(interval_sub, [first, *rest])
    return infer_rational_arithmetic_result(interval_mul, [first, ("integer", -1, -1)])


def _infer_div(first, *rest):            # src/2typeinfer.lisp:131-139
    if rest:
        return infer_rational_arithmetic_result(interval_div, [first, *rest], DEFAULT_FLOAT)
    return infer_rational_arithmetic_result(interval_div, [("integer", 1, 1), first], DEFAULT_FLOAT)


def _infer_abs(t):
    head, lo, hi = t
    if is_complex(head):
        return (complex_part(head), 0.0, None)
    l, h = _lo(lo), _hi(hi)
    nl = 0 if l <= 0 <= h else min(abs(l), abs(h))
    nh = max(abs(l), abs(h))
    nl, nh = _post(nl, nh)
    if head == "integer":
        return infer_rational_arithmetic_result(interval_max, [("integer", nl, nh), ("integer", nl, nh)])
    return (head, nl, nh)


def _infer_part(t):                       # complex-part-inferer, src/2typeinfer.lisp:283-296
    return (complex_part(t[0]), None, None) if is_complex(t[0]) else t


_INFERERS.update({
    "conjugate": lambda t: t, "realpart": _infer_part, "imagpart": _infer_part,
    "-": _infer_sub, "/": _infer_div, "abs": _infer_abs,
    "%square": lambda t: _INFERERS["*"](t, t), "square": lambda t: _INFERERS["*"](t, t),
    "1+": lambda t: _INFERERS["+"](("integer", 1, 1), t),
    "1-": lambda t: _infer_sub(t, ("integer", 1, 1)),
})


# ---- the rest of the scalar-function surface (src/2typeinfer.lisp:255-454,606-660) -----------------------------------
def _f32(v):
    import numpy as np
    return float(np.float32(v))


def _divlike(kind):
    """floor / ceiling / round / truncate inferers (:396-454): (integer ,@(interval-KIND l1 h1 l2 h2)).  interval-divlike
    (src/1type.lisp:238-281) evaluates the function on the four corners with the divisor converted by (float y), i.e. in
    single-float; a divisor type that contains or touches zero, or any open bound, gives an unbounded integer -> a T array."""
    fn = {"floor": math.floor, "ceiling": math.ceil, "truncate": math.trunc,
          "round": lambda q: int(round(q))}[kind]                      # Python's round is half-even like CL's

    def infer(x, y=("integer", 1, 1)):
        if None in (x[1], x[2], y[1], y[2]) or y[1] <= 0 <= y[2]:
            raise TypeInferenceError(f"({kind} x y): the quotient's bounds are open: (integer * *) upgrades to T (general array)")
        corners = []
        for a in (x[1], x[2]):
            for b in (y[1], y[2]):
                q = _f32(_f32(a) / _f32(b)) if (x[0] == "integer") else a / b
                corners.append(fn(q))
        lo, hi = min(corners), max(corners)
        return ("integer", max(lo, MOST_NEGATIVE_FIXNUM), min(hi, MOST_POSITIVE_FIXNUM))
    return infer


def _interval_intersection(l1, h1, l2, h2):
    lo = l2 if l1 is None else l1 if l2 is None else max(l1, l2)
    hi = h2 if h1 is None else h1 if h2 is None else min(h1, h2)
    if lo is not None and hi is not None and lo > hi:
        raise TypeInferenceError("empty interval: the reference infers the type NIL")
    return lo, hi


def _infer_mod(x, y):                     # (set-type-inferer 'mod 'intersection-to-float-type), :456 -- replicated, not fixed
    return infer_rational_arithmetic_result(_interval_intersection, [x, y])


def _infer_rem(x, y):                     # :459-481: x ^ (y v -y); the two pieces of the union are taken as their hull
    m = None if (y[1] is None or y[2] is None) else max(abs(y[1]), abs(y[2]))
    hull = (y[0], None if m is None else -m, m)
    return infer_rational_arithmetic_result(_interval_intersection, [x, hull])


COMPLEX_WHEN_UNPROVEN = {}


def _real_when(name, pred, why):
    """Functions that leave the reals outside a domain.  When the argument's TYPE is not provably inside the domain the
    reference infers (complex <float>) and stores #C(x 0.0) for every in-domain element.  The CUDA backend has no complex
    element types: it keeps the real float type (the real parts are the same numbers) and the kernels raise NCL_ERR_DOMAIN
    if any ELEMENT is outside the domain -- exactly the elements whose imaginary part would be non-zero.
    `reference_infers_complex(name, type)` tells the two cases apart."""
    COMPLEX_WHEN_UNPROVEN[name] = (pred, why)

    def infer(t):
        return (float_substitution(t, int_result=DEFAULT_FLOAT), None, None)
    return infer


def reference_infers_complex(name: str, t) -> bool:
    name = name.lower().split(":")[-1]
    return name in COMPLEX_WHEN_UNPROVEN and not COMPLEX_WHEN_UNPROVEN[name][0](t)


def _infer_isqrt(t):                      # (floor (sqrt x)), :644
    if t[1] is None or t[1] < 0 or t[2] is None:
        raise TypeInferenceError("(isqrt x): open upper bound: (integer * *) upgrades to T")
    return ("integer", math.isqrt(int(t[1])), math.isqrt(int(t[2])))


def _infer_expt(base, power):             # (exp (* (log base) power)), :255
    head = float_contagion((float_substitution(base, DEFAULT_FLOAT), None, None), power)
    return (float_substitution((head, None, None), DEFAULT_FLOAT) if head == "integer" else head, None, None)


_nonneg = lambda t: t[1] is not None and t[1] >= 0                                  # noqa: E731
_INFERERS.update({
    "floor": _divlike("floor"), "ceiling": _divlike("ceiling"), "round": _divlike("round"), "truncate": _divlike("truncate"),
    "mod": _infer_mod, "rem": _infer_rem, "expt": _infer_expt, "isqrt": _infer_isqrt,
    # sqrt = (exp (/ (log x) 2)) and log (numcl:logc) leave the reals unless the type is provably non-negative (:196-214, 642);
    # numcl:log (%logr) and %log2 ignore that on purpose (:216-240, 257-283)
    "sqrt": _real_when("sqrt", _nonneg, "(sqrt x) of a type that may be negative"),
    "log": _real_when("log", _nonneg, "(log x) of a type that may be negative"),
    "%log2": _irrational, "log2": _irrational,
    "asin": _real_when("asin", lambda t: t[1] is not None and t[2] is not None and -1 < t[1] and t[2] < 1, "(asin x) unless -1 < x < 1"),
    "acos": _real_when("acos", lambda t: t[1] is not None and t[2] is not None and -1 < t[1] and t[2] < 1, "(acos x) unless -1 < x < 1"),
    "atan": _irrational, "sinh": _irrational, "cosh": _irrational,
    "asinh": _real_when("asinh", _nonneg, "(asinh x) = (log (+ x (sqrt (+ 1 (* x x))))) of a type that may be negative"),
    "acosh": _real_when("acosh", lambda t: t[1] is not None and t[1] >= 1, "(acosh x) unless x >= 1"),
    "atanh": _real_when("atanh", lambda t: t[1] is not None and t[2] is not None and t[1] >= -1 and t[2] <= 1, "(atanh x) unless -1 <= x <= 1"),
    "logandc1": lambda x, y: _bitwise("logand")(_bitwise("lognot")(x), y),
    "logandc2": lambda x, y: _bitwise("logand")(x, _bitwise("lognot")(y)),
    "logorc1": lambda x, y: _bitwise("logior")(_bitwise("lognot")(x), y),
    "logorc2": lambda x, y: _bitwise("logior")(x, _bitwise("lognot")(y)),
    "lognand": lambda x, y: _bitwise("lognot")(_bitwise("logand")(x, y)),
    "lognor": lambda x, y: _bitwise("lognot")(_bitwise("logior")(x, y)),
})


def infer_type(name: str, *types) -> tuple:
    """infer-type, src/2typeinfer.lisp:40-45."""
    name = name.lower().split(":")[-1]
    if name not in _INFERERS:
        raise TypeInferenceError(f"Missing type inference function for {name}")
    return _INFERERS[name](*types)


def interpret_type(form, env=None) -> tuple:
    """interpret-type, src/2typeinfer.lisp:47-93, over a TRANSFORM form.

    form: nested lists [fname, arg...], variable names (looked up in env), type tuples, numbers."""
    env = env or {}
    if isinstance(form, tuple):
        return form
    if isinstance(form, str):
        if form in env:
            return env[form]
        raise TypeInferenceError(f"unbound variable {form} in transform")
    if isinstance(form, (int, float)):
        return strict_type_of(form)
    if isinstance(form, list):
        return infer_type(form[0], *[interpret_type(a, env) for a in form[1:]])
    return strict_type_of(form)


def einsum_output_types(transforms, in_types, n_out: int):
    """einsum-output-types, src/3einsum.lisp:495-518: simulate 10 rounds from (integer 0 0)."""
    o_types = [("integer", 0, 0)] * n_out
    prev = None
    for _ in range(10):
        if prev is not None:
            env = {f"${i + 1}": t for i, t in enumerate(in_types)}
            env.update({f"@{i + 1}": t for i, t in enumerate(prev)})
            o_types = [interpret_type(tr, env) for tr in transforms]
        if prev is not None and o_types == prev:
            return o_types
        prev = o_types
    return [(float_substitution(t, int_result="fixnum"), None, None) if t[0] != "integer"
            else ("integer", MOST_NEGATIVE_FIXNUM, MOST_POSITIVE_FIXNUM) for t in o_types]


def reduce_result_type(fn: str, elt_type) -> tuple:
    """%reduce-array-result-type, src/5reduce.lisp:34-40."""
    t = infer_type(fn, elt_type, elt_type)
    head = float_substitution(t, int_result="fixnum")
    if head == "fixnum":
        return ("integer", MOST_NEGATIVE_FIXNUM, MOST_POSITIVE_FIXNUM)
    return (head, None, None)


# identities and extremes, src/1constants.lisp:30-89
def zero_value(t):
    return 0.0 if t[0] != "integer" else 0


def one_value(t):
    return 1.0 if t[0] != "integer" else 1


_F32_MAX = 3.4028234663852886e38
_F64_MAX = 1.7976931348623157e308


def most_positive_value(t):
    head, lo, hi = t
    if head == "integer":
        if hi is not None and hi <= MOST_POSITIVE_FIXNUM and lo is not None and lo >= MOST_NEGATIVE_FIXNUM:
            return hi                                    # (fixnum-subtype _ high)
        return _F32_MAX
    return _F64_MAX if head == "double-float" else _F32_MAX


def most_negative_value(t):
    head, lo, hi = t
    if head == "integer":
        if hi is not None and hi <= MOST_POSITIVE_FIXNUM and lo is not None and lo >= MOST_NEGATIVE_FIXNUM:
            return lo
        return -_F32_MAX
    return -_F64_MAX if head == "double-float" else -_F32_MAX
