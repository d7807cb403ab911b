"""Pins the CPU oracle against every known-answer vector numcl's own tests hold for the hot path
(tests/golden/numcl_tests.json, transcribed from t/package.lisp / example.lisp / 3einsum.lisp)."""
import numpy as np
import pytest

import refsem
from golden_runner import CASE_IDS, CASES, run_case


class OracleBackend:
    def __init__(self, O):
        self.O = O

    def array(self, v, t):
        return self.O.OArr.from_values(np.array(v), t)

    def values(self, r):
        if isinstance(r, self.O.OArr):
            return r.values() if r.shape != () else r.values().reshape(())
        return r

    def dtype_name(self, r):
        return r.dtype

    def einsum(self, spec, ins, outs):
        return refsem.einsum(self.O, spec, ins, outs)

    def kron(self, a, b):
        r = refsem.einsum(self.O, "ij kl -> ikjl", [a, b])
        v = r.values()
        return v.reshape([x * y for x, y in zip(a.shape, b.shape)])

    def map_array_into(self, r, fn, ins):
        form = "(" + fn + " " + " ".join(f"${i + 1}" for i in range(len(ins))) + ")"
        got = self.O.map_array(form, ins, r.dtype)
        r.storage[:] = got.storage

    def map_array(self, fn, ins):
        from numcl_b200 import typeinfer as ti
        rt = ti.upgrade(ti.infer_type(fn, *[ti.type_of_dtype(a.dtype) for a in ins]))
        form = "(" + fn + " " + " ".join(f"${i + 1}" for i in range(len(ins))) + ")"
        return self.O.map_array(form, ins, rt)

    def arith(self, op, args):
        return refsem.arith(self.O, op, args)

    def reduce_array(self, fn, a, axes):
        return refsem.reduce_array(self.O, fn, a, axes, initial="zero")

    def sum(self, a, axes):
        return refsem.reduce_array(self.O, "+", a, axes)

    def plan_broadcast(self, shapes):
        return self.O.plan_broadcast([tuple(s) for s in shapes])

    def expect_normalize_error(self, spec):
        with pytest.raises(self.O.OracleError, match="non-alpha"):
            self.O.normalize(spec)

    def normalize(self, spec):
        return self.O.normalize(spec)


@pytest.mark.parametrize("case", CASES, ids=CASE_IDS)
def test_oracle_matches_reference_vector(oracle, case):
    run_case(case, OracleBackend(oracle))
