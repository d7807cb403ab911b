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


@pytest.mark.parametrize("dt,dist,lo,hi", [("f32", 0, -3.0, 3.0), ("f64", 0, -3.0, 3.0), ("fixnum", 0, -10 ** 6, 10 ** 6), ("ub8", 0, 0, 255), ("sb16", 0, -3000, 3000),
                                           ("f32", 1, -8, 8)])
@pytest.mark.parametrize("shape,axes", [((64, 4096), 1), ((64, 4096), 0), ((64, 4096), None), ((33, 257), 1), ((33, 257), 0), ((7, 12, 40), (1, 2)), ((7, 12, 40), 1),
                                        ((5, 3, 4, 6), (0, 3)), ((4000, 3), 0), ((4000, 3), 1), ((3, 70000), 1)])
def test_fused_mean_variance_stdev_equal_the_composition(nc, oracle, dt, dist, lo, hi, shape, axes):
    """ncl_reduce_stat (one launch: the sum kernel's final store divides by the count and takes the root) against
    (numcl:sqrt (numcl:/ (numcl:sum (square a)) count)) run pass by pass through the same library.  Integers and exactly
    summable floats: the same bits (the epilogue reproduces the roundings of the separate passes).  Random floats: the fused
    and the composed sums may take different (equally valid) orders -- the usual reduction tolerance."""
    from numcl_b200 import reduce as R
    a = rand(oracle, shape, dt, SEED, dist, lo, hi)
    g = to_gpu(nc, a)
    exact = dt not in ("f32", "f64") or dist == 1
    tol = 1e-12 if dt == "f64" else 1e-5
    n = R._count(g, axes)
    for fused, composed in ((R.mean, R.mean_composed), (R.variance, R.variance_composed), (R.standard_deviation, R.standard_deviation_composed)):
        x, y = fused(g, axes=axes), composed(g, axes=axes)
        if isinstance(y, nc.NArray):
            assert isinstance(x, nc.NArray) and x.dtype.name == y.dtype.name and x.shape == y.shape
            if exact:
                assert np.array_equal(x.raw()[: x.size * x.dtype.slot_bits // 8], y.raw()[: y.size * y.dtype.slot_bits // 8]), (fused.__name__, x.numpy().reshape(-1)[:4], y.numpy().reshape(-1)[:4])
            else:
                xv, yv = x.numpy().astype(np.float64), y.numpy().astype(np.float64)
                assert np.linalg.norm(xv - yv) <= tol * np.linalg.norm(yv), fused.__name__
        else:
            # a full reduce yields a bare number; the reference divides single-float by integer in single-float, the mirror's
            # scalar composition in Python doubles -- compare with the single-float arithmetic
            ft = np.float64 if dt == "f64" else np.float32
            s = ft(R.sum(g if fused is R.mean else R.square(g)))
            want = ft(s / ft(n))
            if fused is R.standard_deviation:
                want = np.sqrt(want)
            if exact:
                assert ft(x) == want, (fused.__name__, x, want)
            else:
                assert abs(x - want) <= tol * abs(want), (fused.__name__, x, want)


def test_fused_statistics_are_one_launch_and_match_numpy(nc, oracle):
    a = rand(oracle, (512, 2048), "f32", SEED, 1, -8, 8)           # exact data: sums are order independent
    g = to_gpu(nc, a)
    v = a.values().astype(np.float64)
    L = nc.lib()
    n0 = L.ncl_launch_count()
    m = nc.mean(g, axes=1)
    assert L.ncl_launch_count() - n0 == 1
    assert np.array_equal(m.numpy(), (v.sum(axis=1).astype(np.float32) / np.float32(2048)))
    n0 = L.ncl_launch_count()
    s = nc.stdev(g, axes=1)
    assert L.ncl_launch_count() - n0 == 1
    want = np.sqrt(((v * v).sum(axis=1).astype(np.float32) / np.float32(2048)).astype(np.float32))
    assert np.array_equal(s.numpy(), want)
    i = rand(oracle, (300, 77), "ub8", SEED, 0, 0, 255)
    mi = nc.mean(to_gpu(nc, i), axes=0)
    assert mi.dtype.name == "f32" and np.array_equal(mi.numpy(), (i.values().sum(axis=0) / 300.0).astype(np.float32))


def test_resident_host_arrays_upload_once(nc, oracle):
    """ncl_host_register: the device mirror outlives the call; a warm call moves ZERO input bytes over PCIe"""
    a = rand(oracle, (256, 1024), "f32", SEED, 1, -8, 8)
    b = rand(oracle, (1024,), "f32", SEED + 1, 1, -8, 8)
    ha = nc.asarray(a.values(), type="f32", where="host").make_resident()
    hb = nc.asarray(b.values(), type="f32", where="host").make_resident()
    want = refsem.arith(oracle, "+", [a, b])
    h0, d0 = _lib.transfer_stats()
    r1 = nc.add(ha, hb)                                             # cold: both inputs cross once
    h1, d1 = _lib.transfer_stats()
    assert h1 - h0 == ha.nbytes + hb.nbytes and d1 - d0 == r1.nbytes
    assert_bit_exact(r1, want)
    r2 = nc.add(ha, hb)                                             # warm: nothing goes up
    h2, d2 = _lib.transfer_stats()
    assert h2 - h1 == 0 and d2 - d1 == r2.nbytes
    assert_bit_exact(r2, want)
    s = nc.sum(ha, axes=1)                                          # another op on the same resident array: still no upload
    assert _lib.transfer_stats()[0] == h2
    assert_bit_exact(s, refsem.reduce_array(oracle, "+", a, (1,)))
    # the host writes into the array: the next use uploads it again (and sees the new contents)
    a2 = rand(oracle, (256, 1024), "f32", SEED + 9, 1, -8, 8)
    ha.set_raw(a2.storage)                                          # marks it dirty
    h3 = _lib.transfer_stats()[0]
    r3 = nc.add(ha, hb)
    assert _lib.transfer_stats()[0] - h3 == ha.nbytes
    assert_bit_exact(r3, refsem.arith(oracle, "+", [a2, b]))


def test_lazy_resident_results_stay_on_the_device_until_fetched(nc, oracle):
    a = rand(oracle, (128, 512), "f32", SEED, 1, -8, 8)
    ha = nc.asarray(a.values(), type="f32", where="host").make_resident()
    out = nc.NArray((128, 512), "f32", where="host")
    out._host[:] = 0xEE
    out.make_resident(lazy=True, fresh=True)
    tmp = nc.NArray((128,), "f32", where="host").make_resident(lazy=True, fresh=True)
    h0, d0 = _lib.transfer_stats()
    nc.map_array_into(out, "+", ha, ha)                             # out = a + a, left on the device
    L = nc.lib()
    assert L.ncl_host_state(out._host.ctypes.data) == 2 and bytes(out._host[:4]) == b"\xee" * 4
    t_out, t_tmp = out.tensor(), tmp.tensor()
    ax = (C.c_int * 1)(1)
    _lib.check(L.ncl_reduce(_lib.OPS["ADD"], C.byref(t_out), ax, 1, C.byref(t_tmp), _lib.OUT_FRESH))     # consumes the device-side result
    h1, d1 = _lib.transfer_stats()
    assert h1 - h0 == ha.nbytes and d1 - d0 == 0                     # one upload of a, nothing came back yet
    two = refsem.arith(oracle, "+", [a, a])
    assert_bit_exact(tmp, refsem.reduce_array(oracle, "+", two, (1,)))   # raw() fetches
    assert_bit_exact(out, two)
    assert _lib.transfer_stats()[1] - d1 == out.nbytes + tmp.nbytes
    assert L.ncl_host_state(out._host.ctypes.data) == 0
    # a view into a registered base vector (displaced array) is served from the same mirror
    row = nc.NArray((512,), "f32", where="host", _storage=ha._host, offset=512)
    h2 = _lib.transfer_stats()[0]
    assert nc.sum(row) == float(a.values()[1].astype(np.float64).sum())
    assert _lib.transfer_stats()[0] == h2


@pytest.mark.parametrize("n", [4096, 65536 + 16, 1 << 20])
def test_eight_bit_simd_in_register_family_bit_exact(nc, oracle, n):
    """k_ew_narrow.cu: ub8 (op) ub8 / scalar with the result types numcl infers (ub15, sb16, ub8), against the oracle"""
    L = nc.lib()
    a = rand(oracle, (n,), "ub8", SEED, 0, 0, 255)
    b = rand(oracle, (n,), "ub8", SEED + 1, 0, 0, 255)
    sa = rand(oracle, (n,), "sb8", SEED + 2, 0, -128, 127)
    sb = rand(oracle, (n,), "sb8", SEED + 3, 0, -128, 127)
    ga, gb, gsa, gsb = (to_gpu(nc, x) for x in (a, b, sa, sb))
    hits = 0
    for fn, f in (("+", nc.add), ("-", nc.sub), ("max", nc.maximum), ("min", nc.minimum)):
        assert_bit_exact(f(ga, gb), refsem.arith(oracle, fn, [a, b])); hits += L.ncl_last_kernel() == b"bcast_narrow"
    for fn in ("logand", "logior", "logxor"):
        assert_bit_exact(getattr(nc, fn)(ga, gb), refsem.broadcast(oracle, fn, a, b)); hits += L.ncl_last_kernel() == b"bcast_narrow"
    for fn, f in (("max", nc.maximum), ("min", nc.minimum)):
        assert_bit_exact(f(gsa, gsb), refsem.arith(oracle, fn, [sa, sb])); hits += L.ncl_last_kernel() == b"bcast_narrow"
    for k in (3, 1, 255, 200):
        assert_bit_exact(nc.mul(ga, k), refsem.arith(oracle, "*", [a, k])); hits += L.ncl_last_kernel() == b"bcast_narrow"
        assert_bit_exact(nc.mul(k, ga), refsem.arith(oracle, "*", [k, a]))
        assert_bit_exact(nc.add(ga, k), refsem.arith(oracle, "+", [a, k]))
        assert_bit_exact(nc.sub(ga, k), refsem.arith(oracle, "-", [a, k]))
        assert_bit_exact(nc.sub(k, ga), refsem.arith(oracle, "-", [k, a]))
    assert hits >= 8                                         # the family is what actually ran
    # caller-supplied narrower results wrap like %coerce: (unsigned-byte 15) keeps 15 bits, ub8 + ub8 into ub8 stays on the engine
    for out in ("ub15", "ub16", "sb16"):
        assert_bit_exact(nc.broadcast("+", ga, gb, type=out), oracle.broadcast("+", a, b, out))
        assert_bit_exact(nc.broadcast("-", ga, gb, type=out), oracle.broadcast("-", a, b, out))
