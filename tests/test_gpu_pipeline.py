"""Panel pipeline of host-resident arrays (ncl_api.cu: Stager::try_pipeline): a call on pinned host arrays is cut into
panels of its outermost kept loop, upload / compute / write-back overlapped on three streams.  The results must be the
same bits as the unpipelined path and as the device-resident path, only the tensor's own window of a host output may be
written, and a registered base vector must end up completely in sync."""
import numpy as np
import pytest

import refsem
from numcl_b200 import _lib
from test_gpu_parity import SEED, assert_bit_exact, rand, to_gpu

pytestmark = pytest.mark.gpu


def host(nc, o, resident=False):
    a = nc.NArray(o.shape, o.dtype, where="host")
    a.set_raw(o.storage)
    return a.make_resident() if resident else a


def pipelined(f):
    """run f, return (result, number of calls that took the panel pipeline, panels)"""
    c0, p0 = _lib.pipeline_stats()
    r = f()
    c1, p1 = _lib.pipeline_stats()
    return r, c1 - c0, p1 - p0


def same_bytes(x, y):
    n = min(x.nbytes, y.nbytes)
    assert x.dtype.name == y.dtype.name and tuple(x.shape) == tuple(y.shape)
    assert np.array_equal(x.raw()[:n], y.raw()[:n])


def test_broadcast_add_on_host_arrays_runs_in_panels(nc, oracle):
    """C1's shape in host memory: 64 MiB up, 64 MiB back, the row vector uploaded once"""
    a = rand(oracle, (4096, 4096), "f32", SEED, 2, -1.0, 1.0)
    b = rand(oracle, (4096,), "f32", SEED + 1, 0, -1.0, 1.0)
    got, calls, panels = pipelined(lambda: nc.add(host(nc, a), host(nc, b)))
    assert calls == 1 and panels >= 2 and got.where == "host"
    assert_bit_exact(got, refsem.arith(oracle, "+", [a, b]))
    same_bytes(got, nc.add(to_gpu(nc, a), to_gpu(nc, b)))


def test_widening_multiply_on_host_arrays(nc, oracle):
    """C5a: 16 MiB of (unsigned-byte 8) up, 128 MiB of fixnums back"""
    x = rand(oracle, (64, 32, 32, 256), "ub8", SEED + 3, 0, 0, 255)
    y = rand(oracle, (256,), "fixnum", SEED + 4, 0, 2 ** 61 - 100, 2 ** 61 + 100)
    got, calls, panels = pipelined(lambda: nc.mul(host(nc, x), host(nc, y)))
    assert calls == 1 and panels >= 2
    assert_bit_exact(got, refsem.arith(oracle, "*", [x, y]))


@pytest.mark.parametrize("axes", [(1,), (1, 2)])
def test_row_reductions_on_host_arrays(nc, oracle, axes):
    """the kept axis is the outermost one: panels of rows; the small result comes back in one piece"""
    e = rand(oracle, (2048, 64, 32), "f32", SEED + 2, 1, -8, 8)
    got, calls, panels = pipelined(lambda: nc.sum(host(nc, e), axes=axes))
    assert calls == 1 and panels >= 2
    assert_bit_exact(got, refsem.reduce_array(oracle, "+", e, axes))


def test_reduction_over_the_outer_axis_is_not_cut(nc, oracle):
    """a reduced loop cannot be cut into panels (they would all write the same elements): the plain path, same answer"""
    e = rand(oracle, (4096, 1024), "f32", SEED + 2, 1, -8, 8)
    got, calls, _ = pipelined(lambda: nc.sum(host(nc, e), axes=0))
    assert calls == 0
    assert_bit_exact(got, refsem.reduce_array(oracle, "+", e, (0,)))
    tot, calls, _ = pipelined(lambda: nc.sum(host(nc, e)))
    assert calls == 0 and tot == refsem.reduce_array(oracle, "+", e, None).values().reshape(-1)[0]


@pytest.mark.parametrize("dt", ["f64", "f32"])
def test_matmul_on_host_arrays_row_panels(nc, oracle, dt):
    """A and C move in row panels, B goes up once; a panel product accumulates over K exactly like the whole product"""
    A = rand(oracle, (2048, 1024), dt, SEED + 5, 0, -1.0, 1.0)
    B = rand(oracle, (1024, 1024), dt, SEED + 6, 0, -1.0, 1.0)
    got, calls, panels = pipelined(lambda: nc.matmul(host(nc, A), host(nc, B)))
    assert calls == 1 and 2 <= panels <= 4
    same_bytes(got, nc.matmul(to_gpu(nc, A), to_gpu(nc, B)))
    ref = oracle.fast_matmul(A.values(), B.values()).astype(np.float64)
    err = np.linalg.norm(got.numpy().astype(np.float64) - ref) / np.linalg.norm(ref)
    assert err <= (1e-12 if dt == "f64" else 1e-5)


def test_batched_einsum_on_host_arrays(nc, oracle):
    A = rand(oracle, (32, 256, 128), "f32", SEED + 7, 0, -1.0, 1.0)
    B = rand(oracle, (32, 128, 256), "f32", SEED + 8, 0, -1.0, 1.0)
    got, calls, panels = pipelined(lambda: nc.einsum("bij bjk -> bik", host(nc, A), host(nc, B)))
    assert calls == 1 and panels >= 2
    same_bytes(got, nc.einsum("bij bjk -> bik", to_gpu(nc, A), to_gpu(nc, B)))


def test_accumulating_into_a_supplied_host_output(nc, oracle):
    """a caller-supplied output is accumulated into, not cleared (src/3einsum.lisp:468): its panels go up too"""
    A = rand(oracle, (1024, 512), "f64", SEED + 9, 1, -8, 8)
    B = rand(oracle, (512, 1024), "f64", SEED + 10, 1, -8, 8)
    C0 = rand(oracle, (1024, 1024), "f64", SEED + 11, 1, -8, 8)
    hc = host(nc, C0)
    _, calls, _ = pipelined(lambda: nc.einsum("ij jk -> ik", host(nc, A), host(nc, B), hc))
    assert calls == 1
    dc = to_gpu(nc, C0)
    nc.einsum("ij jk -> ik", to_gpu(nc, A), to_gpu(nc, B), dc)
    same_bytes(hc, dc)
    want = C0.values() + A.values() @ B.values()
    assert np.array_equal(hc.numpy(), want)


def test_host_output_view_in_panels_only_its_window_is_written(nc, oracle):
    """map-array-into rows 3 .. 4098 of a 4100-row host matrix: the rows in front of and behind the view survive"""
    rows, cols = 4100, 1024
    base = nc.NArray((rows, cols), "f32", where="host")
    sentinel = (np.arange(rows * cols, dtype=np.float32) % 8191) + 7
    base.set_raw(sentinel)
    view = nc.NArray((rows - 5, cols), "f32", where="host", _storage=base._host, offset=3 * cols)
    x = rand(oracle, (rows - 5, cols), "f32", SEED + 12, 1, -8, 8)
    y = rand(oracle, (rows - 5, cols), "f32", SEED + 13, 1, -8, 8)
    _, calls, panels = pipelined(lambda: nc.map_array_into(view, "*", host(nc, x), host(nc, y)))
    assert calls == 1 and panels >= 2
    out = base.raw().view(np.float32)[: rows * cols]
    want = sentinel.copy()
    want[3 * cols:(rows - 2) * cols] = (x.values() * y.values()).reshape(-1)
    assert np.array_equal(out, want)


def test_resident_vector_is_uploaded_in_panels_and_ends_in_sync(nc, oracle):
    """a registered base vector whose host copy is newer goes up panel by panel under the first call -- all of it, also the
    rows in front of the view the call works on -- and later calls move nothing"""
    e = rand(oracle, (4096, 1024), "f32", SEED + 14, 1, -8, 8)
    he = host(nc, e, resident=True)
    L = nc.lib()
    assert L.ncl_host_state(he._host.ctypes.data) == 1
    tail = nc.NArray((4096 - 100, 1024), "f32", where="host", _storage=he._host, offset=100 * 1024)
    h0, _ = _lib.transfer_stats()
    got, calls, panels = pipelined(lambda: nc.sum(tail, axes=1))
    assert calls == 1 and panels >= 2
    assert L.ncl_host_state(he._host.ctypes.data) == 0
    h1, _ = _lib.transfer_stats()
    assert h1 - h0 == he.nbytes                                    # the whole vector, exactly once
    want = refsem.reduce_array(oracle, "+", e, (1,))
    assert np.array_equal(got.numpy(), want.values()[100:])
    whole, calls, _ = pipelined(lambda: nc.sum(he, axes=1))         # the rows in front of the view are there as well
    h2, _ = _lib.transfer_stats()
    assert calls == 0 and h2 == h1
    assert_bit_exact(whole, want)
    he.set_raw(rand(oracle, (4096, 1024), "f32", SEED + 15, 1, -8, 8).storage)     # the host writes: the next use uploads again
    got, calls, _ = pipelined(lambda: nc.sum(he))
    assert L.ncl_host_state(he._host.ctypes.data) == 0


def test_lazy_resident_output_stays_on_the_device(nc, oracle):
    a = rand(oracle, (4096, 1024), "f32", SEED + 16, 1, -8, 8)
    b = rand(oracle, (1024,), "f32", SEED + 17, 1, -8, 8)
    out = nc.NArray((4096, 1024), "f32", where="host")
    out._host[:] = 0xEE
    out.make_resident(lazy=True, fresh=True)
    _, d0 = _lib.transfer_stats()
    b2 = oracle.OArr.from_values(np.tile(b.values(), (4096, 1)).astype(np.float32), "f32")       # map-array wants equal shapes
    _, calls, _ = pipelined(lambda: nc.map_array_into(out, "+", host(nc, a), host(nc, b2)))
    _, d1 = _lib.transfer_stats()
    assert calls == 1 and d1 == d0 and nc.lib().ncl_host_state(out._host.ctypes.data) == 2
    assert bytes(out._host[:4]) == b"\xee" * 4
    assert np.array_equal(out.numpy(), a.values() + b.values())      # fetch() brings it back (exact data)
    assert nc.lib().ncl_host_state(out._host.ctypes.data) == 0


def test_pageable_host_memory_takes_the_plain_path(nc, oracle):
    """asynchronous copies of pageable memory block the host: no panels, same result"""
    a = rand(oracle, (4096, 1024), "f32", SEED + 18, 1, -8, 8)
    store = np.empty(a.storage.size + 64, dtype=np.uint8)
    store[: a.storage.size] = a.storage
    ha = nc.NArray(a.shape, "f32", where="host", _storage=store)
    got, calls, _ = pipelined(lambda: nc.sum(ha, axes=1))
    assert calls == 0
    assert_bit_exact(got, refsem.reduce_array(oracle, "+", a, (1,)))


def test_transposed_result_has_no_common_outer_loop(nc, oracle):
    a = rand(oracle, (2048, 2048), "f32", SEED + 19, 1, -8, 8)
    got, calls, _ = pipelined(lambda: nc.transpose(host(nc, a)))
    assert calls == 0
    assert np.array_equal(got.numpy(), a.values().T)
