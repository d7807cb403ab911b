"""Writes tests/golden/numcl_tests.json: the known-answer vectors of numcl's own test-suite for the
einsum / broadcast / reduce hot path, transcribed by hand from /root/reference/t/package.lisp and
example.lisp (file:line in each case).  The reference is Common Lisp and cannot run in this image,
so these literals -- not generated outputs -- are the only reference-produced values there are.

Array literals: {"v": nested list, "t": element type}.  Types follow what the Lisp expression
creates: (ones 5) is a bit array, (arange 5) an (unsigned-byte 4) array, :type 'fixnum a fixnum one.
"""
import json
import os

F, B = "fixnum", "bit"


def A(v, t):
    return {"v": v, "t": t}


ones5 = A([1] * 5, B)
m0123 = A([[0, 1], [2, 3]], F)
m5678 = A([[5, 6], [7, 8]], F)
ones55 = A([[1] * 5] * 5, F)
ones10 = A([1] * 10, B)

cases = []


def einsum(name, ref, spec, inputs, expect, outputs=None):
    cases.append(dict(kind="einsum", name=name, ref=ref, spec=spec, inputs=inputs, outputs=outputs, expect=expect))


# ---- (test einsum ...) t/package.lisp:492-540 ------------------------------------------------
einsum("identity-implicit", "t/package.lisp:497-498", "i", [ones5], [1] * 5)
einsum("identity", "t/package.lisp:499-500", "i -> i", [ones5], [1] * 5)
einsum("sum-vector", "t/package.lisp:501-502", "i ->", [A([1] * 5, F)], 5)
einsum("sum-matrix", "t/package.lisp:504", "ij ->", [ones55], 25)
einsum("trace", "t/package.lisp:505", "ii ->", [ones55], 5)
einsum("diagonal", "t/package.lisp:506-507", "ii -> i", [ones55], [1] * 5)
einsum("identity-matrix-implicit", "t/package.lisp:509-510", "ij", [ones55], [[1] * 5] * 5)
einsum("row-sum", "t/package.lisp:511-512", "ij -> i", [ones55], [5] * 5)
einsum("matmul", "t/package.lisp:514-519", "ij jk -> ik", [m0123, m5678], [[7, 8], [31, 36]])
einsum("matmul-alt-notation", "t/package.lisp:521-526", "(i k) (k j) -> (i j)", [m0123, m5678], [[7, 8], [31, 36]])
einsum("transpose", "t/package.lisp:528-530", "ij -> ji", [m0123], [[0, 2], [1, 3]])
einsum("transpose-implicit", "t/package.lisp:531-532", "ji", [m0123], [[0, 2], [1, 3]])
einsum("diagonal-implicit", "t/package.lisp:533-534", "ii", [m0123], [0, 3])
einsum("sum-0123", "t/package.lisp:535-536", "ij ->", [m0123], 6)
einsum("transpose-into-f32", "t/package.lisp:538-540", "ij -> ji", [m0123], [[0.0, 2.0], [1.0, 3.0]],
       outputs=[A([[0.0, 0.0], [0.0, 0.0]], "f32")])

# ---- (test einsum-stride ...) t/package.lisp:542-566 -------------------------------------------
einsum("stride-default", "t/package.lisp:544", "i -> ", [ones10], 10)
einsum("stride-explicit-transform", "t/package.lisp:545", "i -> (cl:+ @1 $1) -> ", [ones10], 10)
for line, opts, want in [
    (546, ":start 0 :end 10 :step 1", 10), (547, ":start 0 :end 10 :step 2", 5), (548, ":start 0 :end 10 :step 3", 4),
    (550, ":start 0 :end -1", 9),
    (552, ":start 0 :end 3  :step 1", 3), (553, ":start 0 :end 5  :step 1", 5), (554, ":start 0 :end 7  :step 1", 7),
    (556, ":start 3 :end 10 :step 1", 7), (557, ":start 5 :end 10 :step 1", 5), (558, ":start 7 :end 10 :step 1", 3),
    (560, ":start 0 :end 5  :step 2", 3), (561, ":start 0 :end 5  :step 3", 2),
    (563, ":start 5 :end 10 :step 2", 3), (564, ":start 5 :end 10 :step 3", 2),
    (566, ":start 5 :end 9  :step 2", 2),
]:
    einsum(f"stride-{line}", f"t/package.lisp:{line}", f"i -> (cl:+ @1 $1) -> -> ((i {opts}))", [ones10], want)

# ---- (test einsum-broadcast ...) t/package.lisp:568-597 ----------------------------------------
einsum("broadcast-contract", "t/package.lisp:570-575", "-i -i -> -i",
       [A([[0, 1], [2, 3]], "ub2"), A([[[0, 1]], [[2, 3]]], "ub2")],
       [[[0, 1], [0, 3]], [[0, 3], [4, 9]]])
ar5 = A([0, 1, 2, 3, 4], "ub4")
ar51 = A([[0], [1], [2], [3], [4]], "ub4")
einsum("broadcast-outer-add", "t/package.lisp:577-586", "- - -> (cl:+ $1 $2) -> -", [ar5, ar51],
       [[i + j for j in range(5)] for i in range(5)])
einsum("broadcast-outer-mul", "t/package.lisp:588-597", "- - -> (cl:* $1 $2) -> -", [ar5, ar51],
       [[i * j for j in range(5)] for i in range(5)])

# ---- example.lisp:169 multiple outputs (values not printed in the reference; shape/semantics only)
einsum("multi-output", "example.lisp:168-170", "ijk -> ij ik jk", [A([[[0, 1], [2, 3]], [[4, 5], [6, 7]]], "ub4")],
       [[[1, 5], [9, 13]], [[2, 4], [10, 12]], [[4, 6], [8, 10]]])

# ---- (test linarg ...) t/package.lisp:599-606 ---------------------------------------------------
cases.append(dict(kind="kron", name="kron-eye-ones", ref="t/package.lisp:599-606",
                  inputs=[A([[1, 0], [0, 1]], B), A([[1, 1], [1, 1]], B)],
                  expect=[[1, 1, 0, 0], [1, 1, 0, 0], [0, 0, 1, 1], [0, 0, 1, 1]]))

# ---- (test map-array ...) t/package.lisp:352-365 ------------------------------------------------
ar9 = A([[0, 1, 2], [3, 4, 5], [6, 7, 8]], "ub4")
ones33 = A([[1] * 3] * 3, B)
cases.append(dict(kind="map-array-into", name="map-array-into-f32", ref="t/package.lisp:353-359", fn="cl:+",
                  result=A([[0.0] * 3] * 3, "f32"), inputs=[ar9, ones33],
                  expect=[[1.0, 2.0, 3.0], [4.0, 5.0, 6.0], [7.0, 8.0, 9.0]]))
cases.append(dict(kind="map-array", name="map-array", ref="t/package.lisp:361-365", fn="cl:+", inputs=[ar9, ones33],
                  expect=[[1, 2, 3], [4, 5, 6], [7, 8, 9]], expect_type_subtype_of=F))

# ---- (test arithmetic ...) t/package.lisp:369-378, (test 1+/1- ...) :348-350 ---------------------
for op, want, ty in [("/", [0.5] * 5, "f32"), ("*", [2] * 5, None), ("+", [3] * 5, None), ("-", [-1] * 5, None),
                     ("max", [2] * 5, None), ("min", [1] * 5, None)]:
    cases.append(dict(kind="arith", name=f"arith-{op}", ref="t/package.lisp:373-378", op=op, args=[ones5, 2],
                      expect=want, expect_type=ty, expect_type_subtype_of=None if ty else F))
cases.append(dict(kind="arith", name="1+zeros", ref="t/package.lisp:348-350", op="+", args=[A([0] * 10, B), 1],
                  expect=[1] * 10, expect_type="ub2", expect_type_subtype_of=None))

# ---- (test reduce ...) t/package.lisp:447-471 ---------------------------------------------------
cases.append(dict(kind="reduce", name="reduce-array-sum", ref="t/package.lisp:448", fn="cl:+", input=ar5, axes=None, expect=10))
cases.append(dict(kind="sum", name="sum-arange5", ref="t/package.lisp:449", input=ar5, axes=None, expect=10))

# ---- (test issue-42 ...) t/package.lisp:712-733 --------------------------------------------------
cases.append(dict(kind="arith", name="issue-42-float", ref="t/package.lisp:713-716", op="/",
                  args=[A([1.0, 2.0, 3.0], "f32"), A([2.0, 4.0, 6.0], "f32")], expect=[0.5] * 3, expect_type="f32", expect_type_subtype_of=None))
for i, (x, y, want) in enumerate([([1, 2, 3], [2, 4, 6], [0.5, 0.5, 0.5]), ([1, 2, 3], [2, -4, 6], [0.5, -0.5, 0.5]),
                                  ([1, 2, -3], [2, 4, 6], [0.5, 0.5, -0.5]), ([1, 2, -3], [2, -4, 6], [0.5, -0.5, -0.5])]):
    tx = "ub2" if min(x) >= 0 else "sb8"
    ty = "ub4" if min(y) >= 0 else "sb8"
    cases.append(dict(kind="arith", name=f"issue-42-int-{i}", ref="t/package.lisp:718-733", op="/",
                      args=[A(x, tx), A(y, ty)], expect=want, expect_type="f32", expect_type_subtype_of=None))

# ---- (test issue-19 ...) t/package.lisp:672-676: (- 1 bit-vector) stays a bit-vector ---------------
cases.append(dict(kind="arith", name="issue-19", ref="t/package.lisp:672-676", op="-",
                  args=[1, A([0, 1, 1, 0, 1, 0, 0, 0, 1, 1], B)], expect=[1, 0, 0, 1, 0, 1, 1, 1, 0, 0],
                  expect_type="bit", expect_type_subtype_of=None))
# ---- (test issue-18 ...) t/package.lisp:665-670: (- a) of (integer 0 10) values holds (integer -10 0)
cases.append(dict(kind="arith", name="issue-18-neg", ref="t/package.lisp:665-667", op="-",
                  args=[A([0, 3, 10, 7, 1, 9, 2, 5, 8, 4], "ub4")], expect=[0, -3, -10, -7, -1, -9, -2, -5, -8, -4],
                  expect_type=None, expect_type_superset_of=[-10, 0]))

# ---- example.lisp:49-50 broadcasting tour ---------------------------------------------------------
ar3c = A([[0], [1], [2]], "ub2")
cases.append(dict(kind="arith", name="example-broadcast-add", ref="example.lisp:49", op="+", args=[ar5, ar3c],
                  expect=[[i + j for j in range(5)] for i in range(3)], expect_type=None, expect_type_subtype_of=F))
cases.append(dict(kind="arith", name="example-broadcast-mul", ref="example.lisp:50", op="*", args=[ar5, ar3c],
                  expect=[[i * j for j in range(5)] for i in range(3)], expect_type=None, expect_type_subtype_of=F))

# ---- plan-broadcast literal, src/3einsum.lisp:427-429 ----------------------------------------------
cases.append(dict(kind="plan-broadcast", name="plan-broadcast", ref="src/3einsum.lisp:427-429",
                  shapes=[[2, 4, 1, 3], [4, 5, 3], [1, 5, 1]],
                  expect=[[[2, 4, 5, 3], [2, 4, 1, 3], [1, 4, 5, 3], [1, 1, 5, 1]],
                          [[60, 15, 3, 1], [12, 3, 3, 1], [60, 15, 3, 1], [5, 5, 1, 1]]]))

# ---- normaliser behaviour -------------------------------------------------------------------------
cases.append(dict(kind="normalize-error", name="bad-spec-char", ref="t/package.lisp:494-495", spec="i00"))
cases.append(dict(kind="normalize", name="docstring-normal-form", ref="src/3einsum.lisp:113-117", spec="ik kj -> ij",
                  expect="iter=(0 1 2) i=((0 1) (1 2)) o=((0 2)) t=((+ @1 (* $1 $2))) io=(nil nil) oo=(nil)"))

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "numcl_tests.json"), "w") as f:
    json.dump(cases, f, indent=1)
print(len(cases), "cases")
