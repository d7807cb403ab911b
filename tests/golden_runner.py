"""Runs the cases of tests/golden/numcl_tests.json against a backend adapter (oracle or GPU)."""
import json
import os

import numpy as np

from numcl_b200 import typeinfer as ti
from numcl_b200.dtypes import dtype

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "numcl_tests.json")) as f:
    CASES = json.load(f)
CASE_IDS = [c["name"] for c in CASES]


def _subtype(dt_name, of):
    d, o = dtype(dt_name), dtype(of)
    if d.kind != o.kind:
        return False
    return d.kind != "int" or (o.lo <= d.lo and d.hi <= o.hi)


def check_types(case, dt_name):
    if case.get("expect_type"):
        assert dt_name == case["expect_type"], f"{case['name']}: element type {dt_name}, reference says {case['expect_type']}"
    if case.get("expect_type_subtype_of"):
        assert _subtype(dt_name, case["expect_type_subtype_of"]), f"{case['name']}: {dt_name} is not a subtype of {case['expect_type_subtype_of']}"
    if case.get("expect_type_superset_of"):
        lo, hi = case["expect_type_superset_of"]
        d = dtype(dt_name)
        assert d.lo <= lo and hi <= d.hi


def equalp(got, want):
    """cl:equalp on arrays/numbers: same dimensions, elements compared with `=`."""
    g, w = np.asarray(got), np.asarray(want)
    assert g.shape == w.shape, f"shape {g.shape} vs {w.shape}"
    assert np.array_equal(g.astype(np.float64), w.astype(np.float64)), f"{g} vs {w}"


def run_case(case, be):
    """be: adapter with array / scalar-aware ops; see tests/test_oracle_golden.py and test_gpu_golden.py."""
    kind = case["kind"]
    mk = lambda a: be.array(a["v"], a["t"]) if isinstance(a, dict) else a
    if kind == "einsum":
        ins = [mk(a) for a in case["inputs"]]
        outs = [mk(a) for a in case["outputs"]] if case.get("outputs") else None
        res = be.einsum(case["spec"], ins, outs)
        if isinstance(res, (list, tuple)):
            for r, w in zip(res, case["expect"]):
                equalp(be.values(r), w)
        else:
            equalp(be.values(res), case["expect"])
    elif kind == "kron":
        a, b = (mk(x) for x in case["inputs"])
        equalp(be.values(be.kron(a, b)), case["expect"])
    elif kind == "map-array-into":
        r = mk(case["result"])
        be.map_array_into(r, case["fn"], [mk(a) for a in case["inputs"]])
        equalp(be.values(r), case["expect"])
    elif kind == "map-array":
        r = be.map_array(case["fn"], [mk(a) for a in case["inputs"]])
        equalp(be.values(r), case["expect"])
        check_types(case, be.dtype_name(r))
    elif kind == "arith":
        r = be.arith(case["op"], [mk(a) for a in case["args"]])
        equalp(be.values(r), case["expect"])
        check_types(case, be.dtype_name(r))
    elif kind == "reduce":
        equalp(be.values(be.reduce_array(case["fn"], mk(case["input"]), case["axes"])), case["expect"])
    elif kind == "sum":
        equalp(be.values(be.sum(mk(case["input"]), case["axes"])), case["expect"])
    elif kind == "plan-broadcast":
        equalp(be.plan_broadcast(case["shapes"]), case["expect"])
    elif kind == "normalize-error":
        be.expect_normalize_error(case["spec"])
    elif kind == "normalize":
        assert be.normalize(case["spec"]) == case["expect"]
    else:
        raise AssertionError(kind)
