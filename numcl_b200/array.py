"""numcl arrays behind the CUDA backend: a base vector in device memory (or pinned host memory)
plus the displaced N-D header -- what src/1instantiate.lisp:81-96 (%make-array) gains when the
`:cuda` backend is active ("3array.lisp gains pinned-host and device buffers", north_star).

Constructors mirror src/2zeros.lisp:41-76 and src/3array.lisp (asarray), src/3arange.lisp.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from . import typeinfer as ti
from .dtypes import DType, dtype, pack, storage_bytes, unpack


class NArray:
    """A numcl array: `shape`, element type `dtype`, storage = device buffer or pinned host bytes."""

    def __init__(self, shape, dt, where: str = "device", _storage=None, offset: int = 0):
        self.shape = tuple(int(d) for d in (shape if shape is not None else ()))
        self.dtype: DType = dtype(dt)
        self.offset = int(offset)
        self.size = int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1
        self.where = where
        self.nbytes = storage_bytes(self.dtype, self.size + self.offset)
        self._buf = None
        self._host = None
        self._pinned = None
        if _storage is not None:
            if where == "device":
                self._buf = _storage
            else:
                self._host = _storage
            self._owner = False
        elif where == "device":
            L = _lib.lib()
            self._buf = L.ncl_alloc(self.nbytes, 0)
            if not self._buf:
                raise _lib.NumclCudaError(_lib.ERR_NOMEM, L.ncl_last_error().decode())
            self._owner = True
        elif where == "host":
            L = _lib.lib()
            p = L.ncl_host_alloc_pinned(self.nbytes)
            if not p:
                raise _lib.NumclCudaError(_lib.ERR_NOMEM, L.ncl_last_error().decode())
            self._pinned = p
            self._host = np.ctypeslib.as_array((C.c_uint8 * self.nbytes).from_address(p))
            self._owner = True
        else:
            raise ValueError(where)

    def __del__(self):
        try:
            if getattr(self, "_resident", False) and self._host is not None:
                _lib.load().ncl_host_unregister(self._host.ctypes.data)
            if getattr(self, "_owner", False):
                L = _lib.load()
                if self._buf:
                    L.ncl_free(self._buf)
                if self._pinned:
                    L.ncl_host_free(self._pinned)
        except Exception:
            pass

    # ---- ABI view ---------------------------------------------------------------------------
    def tensor(self) -> _lib.Tensor:
        t = _lib.Tensor()
        t.buf = self._buf
        t.host = self._host.ctypes.data if self._host is not None else None
        t.dtype = self.dtype.code
        t.rank = len(self.shape)
        for i, d in enumerate(self.shape):
            t.dims[i] = d
        t.elem_offset = self.offset
        return t

    @property
    def rank(self):
        return len(self.shape)

    def device_ptr(self) -> int:
        return _lib.lib().ncl_buf_ptr(self._buf)

    # ---- data movement ----------------------------------------------------------------------
    # ---- residency (host arrays): what the `:cuda` glue does in %make-array, src/1instantiate.lisp:81-96 ---------
    def make_resident(self, lazy: bool = False, fresh: bool = False) -> "NArray":
        """Attach a device mirror that outlives the calls (ncl_host_register): the array is uploaded once, later calls
        on it move nothing over PCIe.  lazy: results written into this array stay on the device until fetch()."""
        assert self.where == "host" and self.offset == 0 and self._host is not None
        flags = (_lib.RESIDENT_LAZY if lazy else 0) | (_lib.RESIDENT_FRESH if fresh else 0)
        _lib.check(_lib.lib().ncl_host_register(self._host.ctypes.data, self.nbytes, flags))
        self._resident = True
        return self

    def dirty(self):
        """The host bytes were modified behind the library's back ((setf aref) in Lisp): upload again on next use."""
        _lib.check(_lib.lib().ncl_host_dirty(self._host.ctypes.data))

    def fetch(self):
        _lib.check(_lib.lib().ncl_host_fetch(self._host.ctypes.data))

    def raw(self) -> np.ndarray:
        """The base vector's bytes (SBCL storage format)."""
        if self.where == "host":
            if getattr(self, "_resident", False):
                self.fetch()                                   # a lazily kept result comes back now
            return self._host.copy()
        out = np.empty(self.nbytes, dtype=np.uint8)
        _lib.check(_lib.lib().ncl_d2h(out.ctypes.data, self._buf, 0, self.nbytes))
        return out

    def set_raw(self, storage: np.ndarray):
        storage = np.ascontiguousarray(storage).view(np.uint8).reshape(-1)
        n = min(storage.size, self.nbytes)
        if self.where == "host":
            self._host[:n] = storage[:n]
            if getattr(self, "_resident", False):
                self.dirty()
        else:
            L = _lib.lib()
            _lib.check(L.ncl_h2d(self._buf, 0, storage.ctypes.data, n))
            _lib.check(L.ncl_sync())

    def numpy(self) -> np.ndarray:
        """Logical element values (float32 / float64 / int64)."""
        return unpack(self.raw(), self.size, self.dtype, self.offset).reshape(self.shape)

    def item(self):
        v = self.numpy().reshape(-1)[0]
        return v.item()

    def to(self, where: str) -> "NArray":
        if where == self.where:
            return self
        out = NArray(self.shape, self.dtype, where=where, offset=self.offset)
        out.set_raw(self.raw())
        return out

    def reshape(self, shape) -> "NArray":
        """A view sharing the base vector (src/2alias.lisp:50-104); -1 is inferred."""
        shape = list(shape) if not isinstance(shape, int) else [shape]
        if -1 in shape:
            known = int(np.prod([d for d in shape if d != -1], dtype=np.int64))
            shape[shape.index(-1)] = self.size // max(known, 1)
        if int(np.prod(shape, dtype=np.int64)) != self.size:
            raise ValueError("reshape: size mismatch")
        v = NArray(shape, self.dtype, where=self.where, _storage=self._buf if self.where == "device" else self._host,
                   offset=self.offset)
        v._base = self   # keep the owner alive
        return v

    def flatten(self):
        return self.reshape([self.size])

    def __repr__(self):
        return f"NArray(shape={self.shape}, dtype={self.dtype.name}, where={self.where})"


# ---- constructors ---------------------------------------------------------------------------------
def empty(shape, type="bit", where="device") -> NArray:
    return NArray(_shape(shape), type, where=where)


def full(shape, value, type=None, where="device") -> NArray:
    """(full shape value :type type): src/2zeros.lisp:46-50; type defaults to (type-of value)."""
    if type is None:
        type = ti.upgrade(ti.strict_type_of(value))
    a = NArray(_shape(shape), type, where=where)
    if where == "host":
        a.set_raw(pack(np.full(a.shape, value), a.dtype))
    else:
        t = a.tensor()
        isf = isinstance(value, (float, np.floating))
        _lib.check(_lib.lib().ncl_fill(C.byref(t), 1 if isf else 0, int(value) if not isf else 0, float(value)))
    return a


def zeros(shape, type="bit", where="device") -> NArray:
    return full(shape, 0, type=type, where=where)


def ones(shape, type="bit", where="device") -> NArray:
    return full(shape, 1, type=type, where=where)


def _shape(shape):
    if shape is None:
        return ()
    if isinstance(shape, (int, np.integer)):
        return (int(shape),)
    return tuple(shape)


def infer_asarray_type(values: np.ndarray) -> str:
    """asarray's element type: the interval union of the contents (src/3array.lisp:62-118)."""
    if values.dtype == np.float32:
        return "f32"
    if values.dtype == np.float64:
        return "f64"
    if values.dtype == np.complex64:
        return "c64"
    if values.dtype == np.complex128:
        return "c128"
    if values.dtype.kind in "iub":
        if values.size == 0:
            return "bit"
        return ti.upgrade(("integer", int(values.min()), int(values.max())))
    raise TypeError(f"unsupported numpy dtype {values.dtype}")


def asarray(values, type=None, where="device") -> NArray:
    if isinstance(values, NArray):
        return values
    v = np.asarray(values)
    if v.dtype == np.float64 and not isinstance(values, (np.ndarray, np.generic)) and type is None:
        v = v.astype(np.float32)           # Lisp reads 1.0 as a single-float
    if v.dtype == np.complex128 and not isinstance(values, (np.ndarray, np.generic)) and type is None:
        v = v.astype(np.complex64)         # ... and #C(1.0 2.0) as a (complex single-float)
    if type is None:
        type = infer_asarray_type(v)
    a = NArray(v.shape, type, where=where)
    a.set_raw(pack(v, a.dtype))
    return a


def arange(*args, type=None, where="device") -> NArray:
    """(arange stop) / (arange start stop [step]); integer results are typed by their range
    (src/3arange.lisp), e.g. (arange 5) is an (unsigned-byte 4) array."""
    v = np.arange(*args)
    if v.dtype.kind == "f" and type is None:
        v = v.astype(np.float32)
    return asarray(v, type=type, where=where)


def random_array(shape, type, seed, dist=0, lo=0.0, hi=1.0, first=0) -> NArray:
    """Synthetic input generated on the device (splitmix64(seed ^ index)).  `first`: index of this array's
    first element in the stream, so that a rank can generate its row shard of a larger array."""
    a = NArray(_shape(shape), type)
    t = a.tensor()
    _lib.check(_lib.lib().ncl_fill_random_at(C.byref(t), C.c_uint64(seed), C.c_uint64(first), dist, float(lo), float(hi)))
    return a


def astype(array: NArray, type) -> NArray:
    """(astype array type), src/3copy.lisp:38-43: every element through %coerce into `type`
    (ncl_convert: a conversion kernel instead of the host map-into)."""
    r = NArray(array.shape, type, where=array.where)
    ta, tr = array.tensor(), r.tensor()
    _lib.check(_lib.lib().ncl_convert(C.byref(ta), C.byref(tr)))
    return r


def copy(array: NArray) -> NArray:
    """(copy array), src/3copy.lisp:27-36: a fresh zero-offset array with the same contents."""
    return astype(array, array.dtype)
