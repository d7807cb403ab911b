"""Arrays beyond 2^31 elements stay on the fast paths (a 2 GiB (unsigned-byte 8) array is ordinary on a 180 GB part):
elementwise (cut on the host into < 2^31-element launches), row / full reductions and the tiled transpose (64-bit indexing).
Checked against the oracle on sampled windows of the synthetic stream, plus size-independent properties."""
import ctypes as C

import numpy as np
import pytest

import refsem
from numcl_b200 import _lib
from test_gpu_parity import SEED

pytestmark = pytest.mark.gpu

ROWS, COLS = 65537, 32768                      # 2^31 + 32768 elements
N = ROWS * COLS


def _window(nc, arr, first, count):
    """elements [first, first + count) of a device array's base vector, as logical values"""
    from numcl_b200.dtypes import unpack
    sb = arr.dtype.slot_bits // 8
    out = np.empty(count * sb, dtype=np.uint8)
    _lib.check(nc.lib().ncl_d2h(out.ctypes.data, arr._buf, first * sb, count * sb))
    return unpack(out, count, arr.dtype, 0)


def _free_gb():
    import torch
    return torch.cuda.mem_get_info()[0] / 2 ** 30


def test_elementwise_reduce_transpose_beyond_2_31_elements(nc, oracle):
    if _free_gb() < 56:
        pytest.skip("needs ~45 GiB of free device memory (the transposes of an integer array are fixnum arrays, 16 GiB each)")
    L = nc.lib()
    a = nc.random_array((ROWS, COLS), "ub8", SEED, 0, 0, 255)
    b = nc.random_array((ROWS, COLS), "ub8", SEED + 1, 0, 0, 255)
    windows = [0, 2 ** 31 - 2048, 2 ** 31 - 8, N - 4096]

    def owin(seed, first, count):
        return oracle.fill_random((count,), "ub8", seed, 0, 0, 255, first=first)

    # ---- a + b -> (unsigned-byte 15): fast path on every piece, bit-exact on the windows -------------------------
    r = nc.add(a, b)
    assert r.dtype.name == "ub15" and L.ncl_last_kernel() != b"generic_nest"
    for w0 in windows:
        want = refsem.arith(oracle, "+", [owin(SEED, w0, 4096), owin(SEED + 1, w0, 4096)]).values()
        assert np.array_equal(_window(nc, r, w0, 4096), want), w0
    # a comparison (packed bits) and a scalar operand over the same range
    lt = nc.lt(a, b)
    assert L.ncl_last_kernel() != b"generic_nest"
    for w0 in (0, 2 ** 31 - 2048, N - 4096):
        aw, bw = owin(SEED, w0, 4096).values(), owin(SEED + 1, w0, 4096).values()
        bits = np.unpackbits(_window_bytes(nc, lt, w0 // 8, 512), bitorder="little")
        assert np.array_equal(bits, (aw < bw).astype(np.uint8)), w0
    del r, lt
    # ---- row sums and the full sum: exact integers; checksum of checksums + sampled rows against the oracle -------
    rows = nc.sum(a, axes=1)
    assert L.ncl_last_kernel() != b"generic_nest"
    total = nc.sum(a)
    assert L.ncl_last_kernel() != b"generic_nest"
    rv = rows.numpy()
    assert int(rv.sum()) == int(total)
    for i in (0, 1, 32768, 65535, 65536):
        assert int(rv[i]) == int(owin(SEED, i * COLS, COLS).values().sum()), i
    # ---- transpose: sampled rows of the result = columns of the source; involution on windows ------------------------
    del b, rows
    t = nc.transpose(a)                                            # integer einsum results are fixnum arrays (3einsum.lisp:517-518)
    assert t.shape == (COLS, ROWS) and t.dtype.name == "fixnum" and L.ncl_last_kernel() != b"generic_nest"
    for j in (0, 777, COLS - 1):
        got = _window(nc, t, j * ROWS, ROWS)                      # row j of the transpose
        idx = np.array([0, 1, 4097, 32768, 65535, 65536])
        want = np.array([int(owin(SEED, int(i) * COLS + j, 1).values()[0]) for i in idx])
        assert np.array_equal(got[idx], want), j
    tt = nc.transpose(t)
    for w0 in windows:
        assert np.array_equal(_window(nc, tt, w0, 4096), owin(SEED, w0, 4096).values()), w0


def _window_bytes(nc, arr, first_byte, nbytes):
    out = np.empty(nbytes, dtype=np.uint8)
    _lib.check(nc.lib().ncl_d2h(out.ctypes.data, arr._buf, first_byte, nbytes))
    return out
