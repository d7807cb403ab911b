"""Round-2 additions, checked through the C ABI: sparse fresh einsum outputs, window-precise staging of host outputs,
CUDA-graph capture of call sequences, sharded synthetic streams, transfer accounting."""
import ctypes as C

import numpy as np
import pytest

import refsem
from numcl_b200 import _lib
from test_gpu_parity import SEED, assert_bit_exact, rand, to_gpu

pytestmark = pytest.mark.gpu


def _poison_pool(nc, nbytes):
    """leave a recycled block full of 0xA5 in the library's buffer cache: the next allocation of that size gets it"""
    junk = nc.NArray((nbytes,), "ub8")
    junk.set_raw(np.full(nbytes, 0xA5, dtype=np.uint8))
    del junk


@pytest.mark.parametrize("spec,shape,dt", [("i -> ii", (5,), "fixnum"), ("i -> ii", (37,), "f32"), ("ij -> iij", (4, 6), "f64"),
                                           ("i -> (+ @1 $1) -> (i :step 2)", (6,), "fixnum")])
@pytest.mark.parametrize("where", ["device", "host"])
def test_einsum_fresh_output_not_densely_covered_reads_zero(nc, oracle, spec, shape, dt, where):
    """`i -> ii` writes the diagonal only; the reference builds the missing output with `zeros` (3einsum.lisp:468), so
    everything else must be 0 -- also when the fresh output is a recycled pool block full of junk."""
    a = rand(oracle, shape, dt, SEED, 1, -8, 8)
    try:
        want = refsem.einsum(oracle, spec, [a])
    except Exception as e:                                   # the o-option case needs a caller-sized output in the oracle too
        pytest.skip(f"oracle cannot size this output: {e}")
    for _ in range(2):
        _poison_pool(nc, 8 * max(64, want.size + 8))
        _poison_pool(nc, 4 * max(64, want.size + 8))
        g = nc.asarray(a.values(), type=dt, where=where)
        got = nc.einsum(spec, g)
        assert_bit_exact(got, want)


def test_host_output_view_only_its_window_is_written(nc, oracle):
    """map-array-into a displaced host view (a row of a matrix): the bytes of the base vector in front of and behind
    the view must survive the call (the staging arena there was never initialised)."""
    base = nc.NArray((6, 8), "f32", where="host")
    sentinel = np.arange(48, dtype=np.float32) + 1000
    base.set_raw(sentinel)
    row = nc.NArray((8,), "f32", where="host", _storage=base._host, offset=16)       # row 2
    x = nc.asarray(np.arange(8, dtype=np.float32), type="f32", where="host")
    y = nc.asarray(np.full(8, 2, dtype=np.float32), type="f32", where="host")
    nc.map_array_into(row, "*", x, y)
    out = base.raw().view(np.float32)[:48]
    want = sentinel.copy()
    want[16:24] = np.arange(8, dtype=np.float32) * 2
    assert np.array_equal(out, want)
    # astype into a view (ncl_convert) and a fold into a view
    for f in (lambda: nc.map_array_into(row, "+", x, y),):
        f()
    want[16:24] = np.arange(8, dtype=np.float32) + 2
    assert np.array_equal(base.raw().view(np.float32)[:48], want)


@pytest.mark.parametrize("dt,bits", [("bit", 1), ("ub2", 2), ("ub4", 4)])
@pytest.mark.parametrize("offset,n", [(3, 7), (5, 40), (1, 3), (8, 16), (13, 100)])
def test_packed_host_output_view_keeps_neighbouring_fields(nc, dt, bits, offset, n):
    """bit / ub2 / ub4 outputs whose window does not start or end on a byte boundary share bytes with their neighbours"""
    total = offset + n + 11
    rng = np.random.default_rng(offset * 131 + n)
    vals = rng.integers(0, 1 << bits, size=total)
    base = nc.asarray(vals, type=dt, where="host")
    before = base.numpy().copy()
    view = nc.NArray((n,), dt, where="host", _storage=base._host, offset=offset)
    src = nc.asarray(rng.integers(0, 1 << bits, size=n), type=dt, where="host")
    nc.map_array_into(view, "identity", src)
    after = base.numpy()
    want = before.copy()
    want[offset:offset + n] = src.numpy()
    assert np.array_equal(after, want)


def test_capture_and_replay_of_a_call_sequence(nc, oracle):
    """ncl_capture_begin/end: the calls are recorded, not run; every replay recomputes from the CURRENT contents of the
    arrays named at capture time."""
    L = nc.lib()
    a = rand(oracle, (64, 256), "f32", SEED, 1, -8, 8)
    b = rand(oracle, (256,), "f32", SEED + 1, 1, -8, 8)
    ga, gb = to_gpu(nc, a), to_gpu(nc, b)
    r = nc.empty((64, 256), "f32")
    rows = nc.empty((64,), "f32")
    tot = nc.empty((), "f32")
    args = (_lib.Tensor * 2)(ga.tensor(), gb.tensor())
    tr, trows, ttot = r.tensor(), rows.tensor(), tot.tensor()
    ax1 = (C.c_int * 1)(1)

    def seq():
        _lib.check(L.ncl_fold(_lib.OPS["ADD"], 1, args, 2, C.byref(tr)))
        _lib.check(L.ncl_reduce(_lib.OPS["ADD"], C.byref(tr), ax1, 1, C.byref(trows), _lib.OUT_FRESH))
        _lib.check(L.ncl_reduce(_lib.OPS["ADD"], C.byref(tr), ax1, 0, C.byref(ttot), _lib.OUT_FRESH))
    seq()                                                       # eager once: scratch arenas get their size
    _lib.check(L.ncl_sync())
    n0 = L.ncl_launch_count()
    with _lib.Graph() as g:
        seq()
        # nothing that must synchronise is allowed while recording
        assert L.ncl_d2h(np.empty(4, dtype=np.uint8).ctypes.data, tot._buf, 0, 4) == _lib.ERR_STATE
        assert L.ncl_sync() == _lib.ERR_STATE
    assert g.kernels == 3 and L.ncl_launch_count() == n0        # recorded, not run
    want_r = refsem.arith(oracle, "+", [a, b])
    for k in range(3):
        rows.set_raw(np.zeros(rows.nbytes, dtype=np.uint8))
        g.launch()
        assert_bit_exact(r, want_r)
        assert_bit_exact(rows, refsem.reduce_array(oracle, "+", want_r, (1,)))
        assert tot.item() == refsem.reduce_array(oracle, "+", want_r, None).values().reshape(-1)[0]
    assert L.ncl_launch_count() == n0 + 9
    # new input contents, same graph
    a2 = rand(oracle, (64, 256), "f32", SEED + 5, 1, -8, 8)
    ga.set_raw(a2.storage)
    g.launch()
    assert_bit_exact(r, refsem.arith(oracle, "+", [a2, b]))
    # host-resident tensors cannot be recorded
    h = nc.asarray(np.ones(8, dtype=np.float32), type="f32", where="host")
    hr = nc.empty((8,), "f32", where="host")
    th = (_lib.Tensor * 2)(h.tensor(), h.tensor())
    thr = hr.tensor()
    _lib.check(L.ncl_capture_begin())
    rc = L.ncl_fold(_lib.OPS["ADD"], 1, th, 2, C.byref(thr))
    gp = C.c_void_p()
    _lib.check(L.ncl_capture_end(C.byref(gp)))
    L.ncl_graph_free(gp)
    assert rc == _lib.ERR_STATE


def test_fill_random_at_is_a_window_of_the_stream(nc, oracle):
    """a rank's row shard of a larger synthetic array: element k = stream element first + k"""
    for dt, dist, lo, hi in (("f32", 2, -1.0, 1.0), ("fixnum", 0, -2 ** 40, 2 ** 40), ("ub8", 0, 0, 255), ("f64", 0, -1.0, 1.0)):
        full = oracle.fill_random((8, 96), dt, SEED, dist, lo, hi)
        fv = full.values()
        for r0, nr in ((0, 8), (3, 2), (7, 1)):
            g = nc.random_array((nr, 96), dt, SEED, dist, lo, hi, first=r0 * 96)
            assert np.array_equal(g.numpy().view(np.uint8), np.ascontiguousarray(fv[r0:r0 + nr]).view(np.uint8)), (dt, r0)
            o = oracle.fill_random((nr, 96), dt, SEED, dist, lo, hi, first=r0 * 96)
            assert np.array_equal(o.values().view(np.uint8), np.ascontiguousarray(fv[r0:r0 + nr]).view(np.uint8))


def test_transfer_stats_count_staging_and_explicit_copies(nc):
    h0, d0 = _lib.transfer_stats()
    a = nc.asarray(np.arange(1024, dtype=np.float32), type="f32", where="host")
    b = nc.asarray(np.arange(1024, dtype=np.float32), type="f32", where="host")
    r = nc.add(a, b)
    h1, d1 = _lib.transfer_stats()
    assert h1 - h0 == 2 * 4096 and d1 - d0 == 4096
    ga = a.to("device")
    gb = b.to("device")
    h2, d2 = _lib.transfer_stats()
    gr = nc.add(ga, gb)
    nc.lib().ncl_sync()
    assert _lib.transfer_stats() == (h2, d2)                   # device-resident operands: nothing crosses PCIe
    assert np.array_equal(gr.numpy(), r.numpy())


def test_trim_releases_the_cache_and_the_library_keeps_working(nc):
    x = nc.ones((1 << 20,), type="f32")
    del x
    _lib.check(nc.lib().ncl_trim())
    y = nc.add(nc.ones((1000,), type="f32"), nc.ones((1000,), type="f32"))
    assert float(nc.sum(y)) == 2000.0
