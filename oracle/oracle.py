"""ctypes harness over oracle/libnumcl_oracle.so -- the CPU restatement of numcl's loop generators.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under numcl_b200/ may import this module.

Parity pinning: against the reference's own known-answer vectors (tests/test_oracle_golden.py,
from t/package.lisp); "parity unpinned" beyond them because no Common Lisp exists in this image.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libnumcl_oracle.so")

MAX_RANK = 8

DTYPES = ["bit", "ub2", "ub4", "ub7", "ub8", "ub15", "ub16", "ub31", "ub32", "ub62", "ub63", "ub64",
          "sb8", "sb16", "sb32", "fixnum", "sb64", "f32", "f64"]
DT = {n: i for i, n in enumerate(DTYPES)}
SLOT_BITS = dict(bit=1, ub2=2, ub4=4, ub7=8, ub8=8, ub15=16, ub16=16, ub31=32, ub32=32, ub62=64,
                 ub63=64, ub64=64, sb8=8, sb16=16, sb32=32, fixnum=64, sb64=64, f32=32, f64=64)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "numcl_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return _SO


class _Tensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("dtype", C.c_int32), ("rank", C.c_int32),
                ("dims", C.c_int64 * MAX_RANK), ("offset", C.c_int64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.nco_last_error.restype = C.c_char_p
        L.nco_storage_bytes.restype = C.c_size_t
        L.nco_storage_bytes.argtypes = [C.c_int, C.c_int64]
        L.nco_fast_sum_all_f32.restype = C.c_float
        L.nco_fast_sum_all_f32.argtypes = [C.c_void_p, C.c_int64]
        L.nco_fill_random.argtypes = [C.POINTER(_Tensor), C.c_uint64, C.c_int, C.c_double, C.c_double]
        L.nco_fill_random_at.argtypes = [C.POINTER(_Tensor), C.c_uint64, C.c_uint64, C.c_int, C.c_double, C.c_double]
        _lib = L
    return _lib


class OracleError(RuntimeError):
    pass


def _check(rc):
    if rc != 0:
        raise OracleError(lib().nco_last_error().decode())


class OArr:
    """A numcl array on the host: base vector in SBCL storage format + shape + element type."""

    def __init__(self, shape, dtype: str, storage: np.ndarray | None = None, offset: int = 0):
        self.shape = tuple(int(d) for d in shape)
        self.dtype = dtype
        n = int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1
        self.size = n
        nbytes = lib().nco_storage_bytes(DT[dtype], n + offset)
        if storage is None:
            storage = np.zeros(nbytes, dtype=np.uint8)
        self.storage = storage
        self.offset = offset

    def c(self, with_data=True) -> _Tensor:
        t = _Tensor()
        t.data = self.storage.ctypes.data if with_data else None
        t.dtype = DT[self.dtype]
        t.rank = len(self.shape)
        for i, d in enumerate(self.shape):
            t.dims[i] = d
        t.offset = self.offset
        return t

    # logical values <-> storage -----------------------------------------------------------
    @staticmethod
    def from_values(values, dtype: str) -> "OArr":
        v = np.asarray(values)
        a = OArr(v.shape, dtype)
        flat = np.ascontiguousarray(v).reshape(-1)
        if dtype == "f32":
            a.storage[: flat.size * 4] = flat.astype(np.float32).view(np.uint8)
        elif dtype == "f64":
            a.storage[: flat.size * 8] = flat.astype(np.float64).view(np.uint8)
        else:
            iv = np.ascontiguousarray(flat.astype(np.int64))
            _check(lib().nco_pack_i64(iv.ctypes.data_as(C.c_void_p), C.c_int64(iv.size), DT[dtype],
                                      a.storage.ctypes.data_as(C.c_void_p)))
        return a

    def values(self) -> np.ndarray:
        n = self.size
        if self.dtype == "f32":
            out = self.storage[: (n + self.offset) * 4].view(np.float32)[self.offset:].copy()
        elif self.dtype == "f64":
            out = self.storage[: (n + self.offset) * 8].view(np.float64)[self.offset:].copy()
        else:
            tmp = np.zeros(n + self.offset, dtype=np.int64)
            _check(lib().nco_unpack_i64(self.storage.ctypes.data_as(C.c_void_p), C.c_int64(n + self.offset),
                                        DT[self.dtype], tmp.ctypes.data_as(C.c_void_p)))
            out = tmp[self.offset:]
        return out.reshape(self.shape)

    def item(self):
        """ensure-singleton (src/2aref.lisp:177-182): rank-0 arrays read as a bare number."""
        v = self.values()
        return v.reshape(-1)[0].item() if self.shape == () else v


def normalize(spec: str) -> str:
    buf = C.create_string_buffer(4096)
    _check(lib().nco_normalize(spec.encode(), buf, 4096))
    return buf.value.decode()


def einsum_shapes(spec: str, inputs, n_out: int):
    ins = (_Tensor * len(inputs))(*[a.c() for a in inputs])
    outs = (_Tensor * n_out)()
    _check(lib().nco_einsum_shapes(spec.encode(), len(inputs), ins, n_out, outs))
    return [tuple(outs[k].dims[i] for i in range(outs[k].rank)) for k in range(n_out)]


def n_outputs(spec: str) -> int:
    ni, no = C.c_int(), C.c_int()
    _check(lib().nco_spec_counts(spec.encode(), C.byref(ni), C.byref(no)))
    return no.value


def einsum(spec: str, *inputs: OArr, out=None, out_dtypes=None):
    """(numcl:einsum spec inputs... [outputs...]).  Fresh outputs are `zeros` of out_dtypes."""
    n_out = len(out) if out is not None else len(out_dtypes)
    if out is None:
        shapes = einsum_shapes(spec, inputs, n_out)
        out = [OArr(s, dt) for s, dt in zip(shapes, out_dtypes)]   # zero storage == (zeros ...)
    ins = (_Tensor * len(inputs))(*[a.c() for a in inputs])
    outs = (_Tensor * n_out)(*[a.c() for a in out])
    _check(lib().nco_einsum(spec.encode(), len(inputs), ins, n_out, outs))
    return out[0] if n_out == 1 else out


def broadcast_shape(x: OArr, y: OArr):
    r = _Tensor()
    _check(lib().nco_broadcast_shape(C.byref(x.c()), C.byref(y.c()), C.byref(r)))
    return tuple(r.dims[i] for i in range(r.rank))


def broadcast(op: str, x: OArr, y: OArr, out_dtype: str) -> OArr:
    r = OArr(broadcast_shape(x, y), out_dtype)
    rt = r.c()
    _check(lib().nco_broadcast(op.encode(), C.byref(x.c()), C.byref(y.c()), C.byref(rt)))
    return r


def reduce_array(op: str, a: OArr, axes, out_dtype: str, init) -> OArr:
    """reduce-array (src/5reduce.lisp:42-65): result = (full kept-shape init), then the nest."""
    rank = len(a.shape)
    axes = list(range(rank)) if axes is None else sorted([axes] if isinstance(axes, int) else list(axes))
    kept = [a.shape[d] for d in range(rank) if d not in axes]
    r = OArr.from_values(np.full(kept, init), out_dtype)
    ax = (C.c_int * max(1, len(axes)))(*axes)
    rt = r.c()
    _check(lib().nco_reduce(op.encode(), C.byref(a.c()), ax, len(axes), C.byref(rt)))
    return r


def map_array(form: str, args, out_dtype: str) -> OArr:
    r = OArr(args[0].shape, out_dtype)
    ts = (_Tensor * len(args))(*[a.c() for a in args])
    rt = r.c()
    _check(lib().nco_map(form.encode(), len(args), ts, C.byref(rt)))
    return r


def plan_broadcast(shapes):
    n = len(shapes)
    ranks = (C.c_int * n)(*[len(s) for s in shapes])
    dims = (C.c_int64 * (n * MAX_RANK))()
    for i, s in enumerate(shapes):
        for k, d in enumerate(s):
            dims[i * MAX_RANK + k] = d
    M = max(len(s) for s in shapes)
    plan = (C.c_int64 * (2 * (n + 1) * max(M, 1)))()
    m = lib().nco_plan_broadcast(n, ranks, dims, plan)
    if m < 0:
        raise OracleError(lib().nco_last_error().decode())
    return np.array(list(plan), dtype=np.int64).reshape(2, n + 1, M)


def fill_random(shape, dtype: str, seed: int, dist: int, lo: float, hi: float, first: int = 0) -> OArr:
    """Synthetic stream element first + k for element k (first > 0: a row shard of a larger array)."""
    a = OArr(shape, dtype)
    t = a.c()
    _check(lib().nco_fill_random_at(C.byref(t), C.c_uint64(seed), C.c_uint64(first), dist, lo, hi))
    return a


# ---- typed fast loops ------------------------------------------------------------------------
def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def fast_add2_f32(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    rows, cols = a.shape
    tmp = np.empty_like(a)
    r = np.empty_like(a)
    lib().nco_fast_add2_f32(_p(a), _p(b), _p(tmp), _p(r), C.c_int64(rows), C.c_int64(cols))
    return r


def fast_mul2_u8_fix(x: np.ndarray, y_tagged: np.ndarray) -> np.ndarray:
    inner = y_tagged.size
    outer = x.size // inner
    tmp = np.empty(x.size, dtype=np.uint8)
    r = np.empty(x.size, dtype=np.int64)
    lib().nco_fast_mul2_u8_fix(_p(x), _p(y_tagged), _p(tmp), _p(r), C.c_int64(outer), C.c_int64(inner))
    return r.reshape(x.shape)


def fast_sum_axis1_f32(a: np.ndarray) -> np.ndarray:
    r = np.empty(a.shape[0], dtype=np.float32)
    lib().nco_fast_sum_axis1_f32(_p(a), _p(r), C.c_int64(a.shape[0]), C.c_int64(a.shape[1]))
    return r


def fast_sum_all_f32(a: np.ndarray) -> np.float32:
    return np.float32(lib().nco_fast_sum_all_f32(_p(a), C.c_int64(a.size)))


def fast_matmul(a: np.ndarray, b: np.ndarray, rows=None) -> np.ndarray:
    m, k = a.shape
    n = b.shape[1]
    sfx = "f64" if a.dtype == np.float64 else "f32"
    if rows is None:
        c = np.zeros((m, n), dtype=a.dtype)
        getattr(lib(), "nco_fast_matmul_" + sfx)(_p(a), _p(b), _p(c), C.c_int64(m), C.c_int64(k), C.c_int64(n))
        return c
    r0, r1 = rows
    c = np.zeros((r1 - r0, n), dtype=a.dtype)
    getattr(lib(), "nco_fast_matmul_rows_" + sfx)(_p(a), _p(b), _p(c), C.c_int64(r0), C.c_int64(r1), C.c_int64(k), C.c_int64(n))
    return c


def fast_bmm_f32(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    nb, m, k = a.shape
    n = b.shape[2]
    c = np.zeros((nb, m, n), dtype=np.float32)
    lib().nco_fast_bmm_f32(_p(a), _p(b), _p(c), C.c_int64(nb), C.c_int64(m), C.c_int64(k), C.c_int64(n))
    return c


def fast_transpose_abcd_dbca(in_tagged: np.ndarray) -> np.ndarray:
    A, B, Cc, D = in_tagged.shape
    out = np.zeros((D, B, Cc, A), dtype=np.int64)
    lib().nco_fast_transpose_abcd_dbca_w64(_p(in_tagged), _p(out), C.c_int64(A), C.c_int64(B), C.c_int64(Cc), C.c_int64(D))
    return out
