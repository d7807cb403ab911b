"""reduce-array, sum, prod, amax, amin, mean on the CUDA backend (src/5reduce.lisp).

`reduce-array` builds and compiles its own loop nest in the reference (:42-87); the `:cuda` glue
routes it to `ncl_reduce`.  Results are identical for integer element types and for float inputs
whose sums are exact; float reductions otherwise use a tree order and are held to 1e-5 (single) /
1e-12 (double) relative tolerance (north_star).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from . import typeinfer as ti
from .array import NArray, empty, full
from .einsum import ensure_singleton
from .numeric import div, map_array, sqrt, square

_OP = {"+": "ADD", "*": "MUL", "max": "MAX", "min": "MIN"}


def _axes(array: NArray, axes):
    if axes is None:
        return list(range(array.rank))                      # nil -> (iota (rank array)), :48-50
    if isinstance(axes, (int, np.integer)):
        return [int(axes)]
    return sorted(int(a) for a in axes)


def reduce_array(fn: str, array: NArray, axes=None, type=None, initial_element=None):
    """(reduce-array fn array &key axes type initial-element), src/5reduce.lisp:42-65.

    Without `initial_element` the accumulator starts from the zero value of the result type,
    exactly as the keyword default (zero-value type) does."""
    f = fn.lower().split(":")[-1]
    if f not in _OP:
        raise _lib.NumclUnsupportedError(_lib.ERR_UNSUPPORTED, f"reduce-array: {fn} is not on the CUDA whitelist (+ * max min)")
    rtype = type or ti.upgrade(ti.reduce_result_type(f, ti.type_of_dtype(array.dtype)))
    ax = _axes(array, axes)
    kept = [array.shape[d] for d in range(array.rank) if d not in ax]
    L = _lib.lib()
    ta = array.tensor()
    axc = (C.c_int * max(1, len(ax)))(*ax)
    if not ax:                                              # :axes '() "does nothing": copy into the result type
        from .numeric import map_array_into
        return map_array_into(empty(array.shape, type=rtype, where=array.where), "identity", array)
    zero = ti.zero_value(ti.type_of_dtype(rtype))
    if initial_element is None or (f == "+" and initial_element == zero):
        r = empty(kept, type=rtype, where=array.where)
        tr = r.tensor()
        if f == "+":
            _lib.check(L.ncl_reduce(_lib.OPS["ADD"], C.byref(ta), axc, len(ax), C.byref(tr), _lib.OUT_FRESH))
            return ensure_singleton(r)
        initial_element = zero
    r = full(kept, initial_element, type=rtype, where=array.where)   # (full result-shape initial-element :type type), :56-59
    tr = r.tensor()
    _lib.check(L.ncl_reduce(_lib.OPS[_OP[f]], C.byref(ta), axc, len(ax), C.byref(tr), _lib.OUT_ACCUMULATE))
    return ensure_singleton(r)


def _fresh(op, array, axes, type):
    rtype = type or ti.upgrade(ti.reduce_result_type({"ADD": "+", "MUL": "*", "MAX": "max", "MIN": "min"}[op], ti.type_of_dtype(array.dtype)))
    ax = _axes(array, axes)
    if not ax:
        return reduce_array({"ADD": "+", "MUL": "*", "MAX": "max", "MIN": "min"}[op], array, axes=axes, type=type)
    kept = [array.shape[d] for d in range(array.rank) if d not in ax]
    r = empty(kept, type=rtype, where=array.where)
    ta, tr = array.tensor(), r.tensor()
    axc = (C.c_int * len(ax))(*ax)
    _lib.check(_lib.lib().ncl_reduce(_lib.OPS[op], C.byref(ta), axc, len(ax), C.byref(tr), _lib.OUT_FRESH))
    return ensure_singleton(r)


def sum(array, axes=None, type=None):        # noqa: A001  (numcl:sum)
    """numcl:sum, :103-105."""
    return _fresh("ADD", array, axes, type)


def prod(array, axes=None, type=None):
    """numcl:prod, :106-112: initial element (one-value type)."""
    return _fresh("MUL", array, axes, type)


def amax(array, axes=None, type=None):
    """numcl:amax, :113-119: starts from most-negative-value (NOT -infinity)."""
    return _fresh("MAX", array, axes, type)


def amin(array, axes=None, type=None):
    """numcl:amin, :120-126: starts from most-positive-value."""
    return _fresh("MIN", array, axes, type)


def _count(array, axes):
    n = 1
    for a in _axes(array, axes):
        n *= array.shape[a]
    return n


def _stat(stat: int, array: NArray, axes):
    """One fused launch (ncl_reduce_stat) when the array is device resident; None when the composition has to run
    (host arrays, packed / ub64 element types, and full reductions of integer arrays, whose result in the reference is
    an exact Lisp rational, not a float)."""
    if not isinstance(array, NArray) or array.where != "device" or array.dtype.slot_bits < 8 or array.dtype.name == "ub64":
        return None
    ax = _axes(array, axes)
    is_int = array.dtype.name not in ("f32", "f64")
    if not ax or (is_int and len(ax) == array.rank) or array.size == 0:
        return None
    kept = [array.shape[d] for d in range(array.rank) if d not in ax]
    r = empty(kept, type="f64" if array.dtype.name == "f64" else "f32")
    ta, tr = array.tensor(), r.tensor()
    axc = (C.c_int * len(ax))(*ax)
    _lib.check(_lib.lib().ncl_reduce_stat(stat, C.byref(ta), axc, len(ax), C.byref(tr)))
    return ensure_singleton(r)


def mean_composed(array, axes=None):
    """numcl:mean = (numcl:/ (numcl:sum array :axes axes) (array-dimension* array axes)), :128-133, pass by pass."""
    return div(sum(array, axes=axes), _count(array, axes))


def variance_composed(array, axes=None):
    return div(sum(square(array), axes=axes), _count(array, axes))


def standard_deviation_composed(array, axes=None):
    v = variance_composed(array, axes=axes)
    if isinstance(v, NArray):
        return sqrt(v)
    import math
    return math.sqrt(v)


def mean(array, axes=None):
    """numcl:mean, :128-133; fused into the sum kernel's final store (bit-identical to the composition)."""
    r = _stat(_lib.STAT_MEAN, array, axes)
    return mean_composed(array, axes) if r is None else r


avg = mean


def variance(array, axes=None):
    """numcl:variance as the reference defines it (:134-139): the mean of SQUARES, no mean
    subtraction (replicate, do not fix)."""
    r = _stat(_lib.STAT_VARIANCE, array, axes)
    return variance_composed(array, axes) if r is None else r


var = variance


def standard_deviation(array, axes=None):
    """numcl:standard-deviation, :140-146 = (numcl:sqrt variance)."""
    r = _stat(_lib.STAT_STDEV, array, axes)
    return standard_deviation_composed(array, axes) if r is None else r


stdev = standard_deviation
