"""Host-side logic of the product (no GPU): the Python mirror of numcl's einsum front end and type
inference, checked against the independently written C oracle and the reference's type assertions."""
import ctypes as C
import re

import numpy as np
import pytest

from numcl_b200 import _lib
from numcl_b200 import typeinfer as ti
from numcl_b200.dtypes import DTYPES, pack, storage_bytes, unpack
from numcl_b200.einsum import build_desc, compile_transform, resolve_shapes
from numcl_b200.einsum_spec import EinsumSpecError, einsum_normalize_subscripts

SPECS = [
    "ij jk -> ik", "(i k) (k j) -> (i j)", "ij -> ji", "ji", "ii", "ii -> i", "ii ->", "ij ->", "ij -> i", "i", "i -> i",
    "i ->", "bij bjk -> bik", "abcd -> dbca", "ij kl -> ikjl", "ijk -> ij ik jk", "i i -> ", "i j -> ij",
    "-i -i -> -i", "- - -> (cl:+ $1 $2) -> -", "- - -> (cl:* $1 $2) -> -", "i -> (sin $1) -> i",
    "i -> (cl:+ @1 $1) -> -> ((i :start 0 :end 10 :step 2))", "i -> (cl:+ @1 $1) -> -> ((i :start 0 :end -1))",
    "ij jk kl -> il", "abc cd -> abd", "ab ab -> ", "ba ab ->", "ijk jk -> i", "kji -> ijk", "ab bc ca ->",
    "ij -> (max @1 $1) -> i", "ij ij -> (+ @1 (* $1 $2) 1.5) -> i", "zy yx -> zx", "i -> (* 2.0d0 $1) -> i",
]


@pytest.mark.parametrize("spec", SPECS)
def test_normaliser_matches_oracle(oracle, spec):
    assert einsum_normalize_subscripts(spec).describe() == oracle.normalize(spec)


@pytest.mark.parametrize("spec", ["i00", "i1 -> i", "ij -> ik", "i -> (+ $1 1) (+ $1 2) -> i"])
def test_normaliser_rejects(oracle, spec):
    with pytest.raises((EinsumSpecError,)):
        einsum_normalize_subscripts(spec)
    with pytest.raises(oracle.OracleError):
        oracle.normalize(spec)


def test_sort_locality_gemm_is_ijk():
    # "ikj-loop is the fastest" -- src/1docstrings.lisp:166-171: i, then j (contracted), then k
    assert einsum_normalize_subscripts("ij jk -> ik").iter_specs == [0, 1, 2]
    assert einsum_normalize_subscripts("bij bjk -> bik").iter_specs == [0, 1, 2, 3]


@pytest.mark.parametrize("spec,shapes", [
    ("ij jk -> ik", [(3, 4), (4, 5)]), ("bij bjk -> bik", [(2, 3, 4), (2, 4, 5)]), ("abcd -> dbca", [(2, 3, 4, 5)]),
    ("-i -i -> -i", [(2, 2), (2, 1, 2)]), ("- - -> (cl:+ $1 $2) -> -", [(5,), (5, 1)]), ("ii -> i", [(4, 4)]),
    ("i -> (cl:+ @1 $1) -> -> ((i :start 5 :end 9 :step 2))", [(10,)]), ("ijk -> ij ik jk", [(2, 3, 4)]),
    ("-ij -jk -> -ik", [(7, 2, 3), (1, 3, 4)]),
])
def test_shape_resolver_matches_oracle(oracle, spec, shapes):
    specs = einsum_normalize_subscripts(spec)
    _, _, mine = resolve_shapes(specs, shapes)
    ins = [oracle.OArr(s, "fixnum") for s in shapes]
    assert [tuple(s) for s in mine] == oracle.einsum_shapes(spec, ins, len(specs.o_specs))


def test_shape_resolver_asserts():
    with pytest.raises(AssertionError):
        resolve_shapes(einsum_normalize_subscripts("ij jk -> ik"), [(3, 4), (5, 6)])
    with pytest.raises(AssertionError):
        resolve_shapes(einsum_normalize_subscripts("-i -i -> -i"), [(2, 3), (4, 3)])


# ---- element types: the reference's own assertions (t/package.lisp) --------------------------------
def T(name):
    return ti.type_of_dtype(name)


def test_type_assertions_from_reference():
    assert ti.upgrade(ti.infer_type("/", T("bit"), ti.strict_type_of(2))) == "f32"                 # :372
    for op in ("*", "+", "-", "max", "min"):                                                          # :373-377
        d = DTYPES[ti.upgrade(ti.infer_type(op, T("bit"), ti.strict_type_of(2)))]
        assert d.kind == "int" and DTYPES["fixnum"].lo <= d.lo and d.hi <= DTYPES["fixnum"].hi
    # (1+ (zeros 10)) is (array (unsigned-byte 2)), :348-350: 0 + bit -> bit, then + 1 -> (integer 1 2)
    step1 = T(ti.upgrade(ti.infer_type("+", ti.strict_type_of(0), T("bit"))))
    assert ti.upgrade(ti.infer_type("+", step1, ti.strict_type_of(1))) == "ub2"
    assert ti.infer_type("*", ("single-float", None, None), T("bit"))[0] == "single-float"          # issue-44, :735-738
    assert ti.upgrade(ti.infer_type("-", ti.strict_type_of(1), T("bit"))) == "bit"                  # issue-19, :672-676
    neg = ti.infer_type("-", ("integer", 0, 10))                                                      # issue-18, :665-667
    assert neg == ("integer", -10, 0)


def test_survey_dtype_table():
    # SURVEY.md Appendix A.1
    assert ti.upgrade(ti.infer_type("*", T("ub8"), T("fixnum"))) == "fixnum"
    assert ti.upgrade(ti.infer_type("+", T("ub8"), T("ub8"))) == "ub15"
    assert ti.upgrade(ti.infer_type("+", T("f32"), T("fixnum"))) == "f32"
    assert ti.upgrade(ti.infer_type("+", T("f64"), T("f32"))) == "f64"
    default = ["+", "@1", ["*", "$1", "$2"]]
    assert ti.upgrade(ti.einsum_output_types([default], [T("f32"), T("f32")], 1)[0]) == "f32"
    assert ti.upgrade(ti.einsum_output_types([default], [T("f64"), T("f64")], 1)[0]) == "f64"
    assert ti.upgrade(ti.einsum_output_types([default], [T("bit"), T("ub8")], 1)[0]) == "fixnum"
    assert ti.upgrade(ti.einsum_output_types([["+", "@1", ["*", "$1"]]], [T("fixnum")], 1)[0]) == "fixnum"
    assert ti.upgrade(ti.einsum_output_types([["+", "$1", "$2"]], [T("ub4"), T("ub4")], 1)[0]) == "ub7"   # fixpoint
    assert ti.upgrade(ti.reduce_result_type("+", T("f32"))) == "f32"
    assert ti.upgrade(ti.reduce_result_type("max", T("ub8"))) == "fixnum"
    assert ti.most_negative_value(("single-float", None, None)) == -3.4028234663852886e38


# ---- storage formats ----------------------------------------------------------------------------------
@pytest.mark.parametrize("name", [n for n in DTYPES if DTYPES[n].kind not in ("c64", "c128")])      # the C oracle has no complex types
def test_pack_matches_oracle_storage(oracle, name):
    rng = np.random.default_rng(7)
    d = DTYPES[name]
    if d.kind == "int":
        lo, hi = max(d.lo, -(2 ** 62)), min(d.hi, 2 ** 62)
        v = rng.integers(lo, hi, size=37, dtype=np.int64, endpoint=True)
    else:
        v = rng.standard_normal(37)
    mine = pack(v, name)
    theirs = oracle.OArr.from_values(v, name).storage
    assert mine.size == theirs.size == storage_bytes(name, 37)
    assert np.array_equal(mine, theirs)
    back = unpack(mine, 37, name)
    assert np.array_equal(back, v.astype(back.dtype))


def test_pack_of_complex_types_is_interleaved_parts():
    """(complex single-float) / (complex double-float): pairs (real, imaginary), 8 / 16 bytes per element"""
    v = np.array([1 + 2j, -0.5 + 0.25j, 3j])
    for name, ft, ct in (("c64", np.float32, np.complex64), ("c128", np.float64, np.complex128)):
        raw = pack(v, name)
        assert raw.size == storage_bytes(name, 3) == 8 * DTYPES[name].slot_bits // 8
        assert np.array_equal(raw[: 3 * 2 * np.dtype(ft).itemsize].view(ft), np.array([1, 2, -0.5, 0.25, 0, 3], ft))
        assert np.array_equal(unpack(raw, 3, name), v.astype(ct)) and np.array_equal(unpack(raw, 2, name, offset=1), v[1:].astype(ct))


def test_complex_type_inference_follows_float_contagion_of_the_parts():
    """src/1type.lisp:446-451"""
    c, z = ti.type_of_dtype("c64"), ti.type_of_dtype("c128")
    T = ti.type_of_dtype
    assert ti.upgrade(ti.infer_type("+", c, T("f32"))) == "c64" and ti.upgrade(ti.infer_type("*", c, T("f64"))) == "c128"
    assert ti.upgrade(ti.infer_type("+", c, T("ub8"))) == "c64" and ti.upgrade(ti.infer_type("/", T("fixnum"), z)) == "c128"
    assert ti.upgrade(ti.infer_type("-", c)) == "c64" and ti.upgrade(ti.infer_type("+", c, z)) == "c128"
    assert ti.upgrade(ti.infer_type("realpart", z)) == "f64" and ti.upgrade(ti.infer_type("imagpart", c)) == "f32"
    assert ti.upgrade(ti.infer_type("conjugate", c)) == "c64" and ti.infer_type("realpart", T("ub8")) == T("ub8")
    assert ti.upgrade(ti.einsum_output_types([["+", "@1", ["*", "$1", "$2"]]], [c, c], 1)[0]) == "c64"
    assert ti.upgrade(ti.reduce_result_type("+", z)) == "c128"
    with pytest.raises(ti.TypeInferenceError):
        ti.infer_type("max", c, c)


def test_pack_wraps_like_coerce(oracle):
    v = np.array([-5, 300, 2 ** 62 + 3, -(2 ** 62) - 1, 127, 128], dtype=np.int64)
    for name in ("ub7", "ub8", "sb8", "fixnum", "ub4", "bit", "sb16", "ub31"):
        assert np.array_equal(pack(v, name), oracle.OArr.from_values(v, name).storage), name


# ---- descriptor / expression serialisation ------------------------------------------------------------
def test_compile_transform_shapes():
    from numcl_b200.einsum_spec import read_subscripts
    e = compile_transform(read_subscripts("(+ @1 (* $1 $2))")[0])
    ops = [e.node[i].op for i in range(e.n)]
    assert ops == [_lib.OPS["X_OUTPUT"], _lib.OPS["X_INPUT"], _lib.OPS["X_INPUT"], _lib.OPS["MUL"], _lib.OPS["ADD"]]
    e = compile_transform(read_subscripts("(max $1 $2 $3)")[0])        # right-nested on SBCL
    assert (e.node[3].a, e.node[3].b) == (1, 2) and (e.node[4].a, e.node[4].b) == (0, 3)
    e = compile_transform(read_subscripts("(- $1)")[0])
    assert e.node[e.n - 1].op == _lib.OPS["NEG"]
    with pytest.raises(NotImplementedError):
        compile_transform(read_subscripts("(gamma $1)")[0])


def test_descriptor_layout_matches_header():
    """sizeof / offsetof of the ctypes mirrors == the C structs (compiled from include/numcl_cuda.h)."""
    import os, subprocess, tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "numcl_cuda.h"
int main(void){
 printf("%zu %zu %zu %zu %zu\n", sizeof(ncl_expr_node), sizeof(ncl_expr), sizeof(ncl_iter_opt), sizeof(ncl_einsum_desc), sizeof(ncl_tensor));
 printf("%zu %zu %zu %zu\n", offsetof(ncl_einsum_desc,in_spec), offsetof(ncl_einsum_desc,in_opt), offsetof(ncl_einsum_desc,transform), offsetof(ncl_tensor,elem_offset));
 return 0; }'''
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "a.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(root, "include"), os.path.join(td, "a.c"), "-o", os.path.join(td, "a")])
        out = subprocess.check_output([os.path.join(td, "a")]).decode().split()
    sizes = [int(x) for x in out]
    assert sizes[:5] == [C.sizeof(_lib.ExprNode), C.sizeof(_lib.Expr), C.sizeof(_lib.IterOpt), C.sizeof(_lib.EinsumDesc), C.sizeof(_lib.Tensor)]
    assert sizes[5:] == [_lib.EinsumDesc.in_spec.offset, _lib.EinsumDesc.in_opt.offset, _lib.EinsumDesc.transform.offset, _lib.Tensor.elem_offset.offset]


def test_build_desc_roundtrip():
    d = build_desc(einsum_normalize_subscripts("i -> (cl:+ @1 $1) -> -> ((i :start 5 :end 9 :step 2))"))
    assert (d.n_in, d.n_out, d.n_iter, d.iter[0]) == (1, 1, 1, 0)
    assert d.in_nopt[0] == 1 and (d.in_opt[0][0].key, d.in_opt[0][0].start, d.in_opt[0][0].end, d.in_opt[0][0].step) == (0, 5, 9, 2)


def test_matrix_chain_order_is_optimal():
    """matmul*'s DP (src/4linear-algebra3.lisp:179-215) against brute force on small chains"""
    import itertools
    from numcl_b200.linalg import matrix_chain_order

    def cost_of(order, dims, i, j):
        if i == j:
            return 0
        k = order[i][j]
        return cost_of(order, dims, i, k) + cost_of(order, dims, k + 1, j) + dims[i] * dims[k + 1] * dims[j + 1]

    def brute(dims, i, j):
        if i == j:
            return 0
        return min(brute(dims, i, k) + brute(dims, k + 1, j) + dims[i] * dims[k + 1] * dims[j + 1] for k in range(i, j))
    rng = np.random.default_rng(3)
    for n in (2, 3, 4, 5, 6):
        for _ in range(5):
            dims = [int(x) for x in rng.integers(1, 40, size=n + 1)]
            order = matrix_chain_order(dims)
            assert cost_of(order, dims, 0, n - 1) == brute(dims, 0, n - 1)
