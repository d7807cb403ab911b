"""The oracle on the everyday shapes of tests/test_gpu_shapes.py, against numpy on exact data (CPU only).

The GPU tests compare the CUDA path with the oracle; this file is the oracle's own check on those shapes: reductions,
permutations and dot-like contractions of small integers (every order of summation exact) must equal numpy's result."""
import numpy as np
import pytest

import refsem

SEED = 20261019


def _vals(o):
    return o.values().astype(np.float64)


@pytest.mark.parametrize("shape,axes", [((5000, 3), (1,)), ((5000, 3), (0,)), ((3, 5000), (0,)), ((3, 5000), (1,)), ((2000, 48), (1,)),
                                        ((2000, 48), (0,)), ((300, 7, 5), (0, 2)), ((300, 7, 5), (0, 1)), ((7001, 2), (0,)), ((64, 100, 36), (1,))])
@pytest.mark.parametrize("dt", ["ub8", "fixnum", "f32", "f64", "sb16"])
def test_oracle_reduce_matches_numpy(oracle, shape, axes, dt):
    a = oracle.fill_random(shape, dt, SEED, 1, 0 if dt == "ub8" else -8, 8)
    for fn, npf in (("+", np.sum), ("max", np.max), ("min", np.min)):
        want = npf(_vals(a), axis=axes)
        got = _vals(refsem.reduce_array(oracle, fn, a, axes)).reshape(want.shape)
        assert np.array_equal(got, want), (fn, shape, axes, dt)


@pytest.mark.parametrize("shape", [(5000, 3), (3, 5000), (409, 7), (2, 7001), (1000, 33)])
@pytest.mark.parametrize("dt", ["f32", "fixnum", "ub8"])
def test_oracle_transpose_matches_numpy(oracle, shape, dt):
    a = oracle.fill_random(shape, dt, SEED + 1, 1 if dt == "f32" else 0, 0 if dt == "ub8" else -100, 100)
    got = refsem.einsum(oracle, "ij -> ji", [a])
    assert got.dtype == ("f32" if dt == "f32" else "fixnum")           # integer einsum outputs are fixnum (3einsum.lisp:517-518)
    assert np.array_equal(_vals(got), _vals(a).T)


@pytest.mark.parametrize("spec,shapes,np_spec", [("i i -> ", [(7001,), (7001,)], "i,i->"), ("ij ij -> i", [(300, 100), (300, 100)], "ij,ij->i"),
                                                 ("ij ij -> j", [(100, 300), (100, 300)], "ij,ij->j"), ("ij j -> i", [(301, 99), (99,)], "ij,j->i"),
                                                 ("i ij -> j", [(99,), (99, 301)], "i,ij->j"), ("bij bj -> bi", [(4, 30, 52), (4, 52)], "bij,bj->bi")])
@pytest.mark.parametrize("dt", ["f32", "f64", "fixnum", "ub8"])
def test_oracle_dot_like_matches_numpy(oracle, spec, shapes, np_spec, dt):
    ins = [oracle.fill_random(s, dt, SEED + 10 + i, 1 if dt in ("f32", "f64") else 0, 0 if dt == "ub8" else -4, 9 if dt == "ub8" else 4)
           for i, s in enumerate(shapes)]
    want = np.einsum(np_spec, *[_vals(a) for a in ins])
    got = _vals(refsem.einsum(oracle, spec, ins)).reshape(np.shape(want))
    assert np.array_equal(got, want)


@pytest.mark.parametrize("sa,sb", [((4000, 3), (3,)), ((4000, 3), (4000, 1)), ((1001, 3), (3,)), ((129, 101), (101,)), ((3,), (4000, 3))])
@pytest.mark.parametrize("op,npf", [("+", np.add), ("-", np.subtract), ("*", np.multiply), ("max", np.maximum), ("min", np.minimum)])
def test_oracle_broadcast_matches_numpy(oracle, sa, sb, op, npf):
    for ta, tb in (("f32", "f32"), ("fixnum", "ub8"), ("sb16", "sb16"), ("f64", "fixnum")):
        x = oracle.fill_random(sa, ta, SEED + 20, 1 if ta in ("f32", "f64") else 0, -9, 9)
        y = oracle.fill_random(sb, tb, SEED + 21, 1 if tb in ("f32", "f64") else 0, 0, 9)
        got = refsem.arith(oracle, op, [x, y])
        assert np.array_equal(_vals(got), npf(_vals(x), _vals(y))), (op, ta, tb, sa, sb)
