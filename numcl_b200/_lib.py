"""ctypes binding of libnumcl_cuda.so (include/numcl_cuda.h).

This is the Python twin of the CFFI stubs in lisp/numcl-cuda-ffi.lisp.  There is no CPU fallback:
if the shared library is missing or no B200 is visible, every compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libnumcl_cuda.so")

MAX_RANK, MAX_INDEX, MAX_OPERANDS, MAX_EXPR, MAX_OPTS = 8, 16, 8, 48, 8

OK, ERR_INVALID, ERR_SHAPE, ERR_UNSUPPORTED, ERR_CUDA, ERR_NOMEM, ERR_STATE, ERR_DOMAIN = 0, -1, -2, -3, -4, -5, -6, -7
OUT_ACCUMULATE, OUT_FRESH = 0, 1
STAT_MEAN, STAT_VARIANCE, STAT_STDEV = 1, 2, 3
RESIDENT_FRESH, RESIDENT_LAZY = 1, 2

OPS = dict(ADD=0, SUB=1, MUL=2, DIV=3, MAX=4, MIN=5, EQ=6, NE=7, LE=8, GE=9, LT=10, GT=11,
           LOGAND=12, LOGIOR=13, LOGXOR=14,
           LOGANDC1=15, LOGANDC2=16, LOGEQV=17, LOGNAND=18, LOGNOR=19, LOGORC1=20, LOGORC2=21,
           FLOOR=22, CEILING=23, ROUND=24, TRUNCATE=25, MOD=26, REM=27, EXPT=28,
           NEG=32, ABS=33, SQUARE=34, SQRT=35, LOGNOT=36, IDENTITY=37, EXP=38, LOG=39, SIN=40, COS=41,
           TAN=42, TANH=43, SIGNUM=44,
           ASIN=45, ACOS=46, ATAN=47, SINH=48, COSH=49, ASINH=50, ACOSH=51, ATANH=52, LOG2=53, ISQRT=54, LOGCOUNT=55,
           INTEGER_LENGTH=56, CONJUGATE=57, REALPART=58, IMAGPART=59,
           X_INPUT=64, X_OUTPUT=65, X_CONST_INT=66, X_CONST_F32=67, X_CONST_F64=68)

# every symbol include/numcl_cuda.h declares (tests/test_abi.py checks the .so exports them all)
EXPORTS = [
    "ncl_init", "ncl_shutdown", "ncl_last_error", "ncl_version", "ncl_device_count", "ncl_set_stream",
    "ncl_get_stream", "ncl_sync", "ncl_launch_count", "ncl_last_kernel", "ncl_force_generic",
    "ncl_alloc", "ncl_free", "ncl_buf_ptr", "ncl_buf_bytes", "ncl_wrap", "ncl_h2d", "ncl_d2h",
    "ncl_host_alloc_pinned", "ncl_host_free", "ncl_storage_bytes", "ncl_fill", "ncl_fill_random",
    "ncl_einsum", "ncl_broadcast", "ncl_fold", "ncl_map", "ncl_reduce", "ncl_convert",
    "ncl_comm_unique_id", "ncl_comm_init", "ncl_comm_destroy", "ncl_allreduce", "ncl_reduce_all_sharded",
    "ncl_reduce_all_sharded_async", "ncl_comm_join",
    "ncl_trim", "ncl_transfer_stats", "ncl_pipeline_stats", "ncl_capture_begin", "ncl_capture_end", "ncl_graph_launch", "ncl_graph_kernels",
    "ncl_graph_free", "ncl_fill_random_at", "ncl_reduce_stat",
    "ncl_host_register", "ncl_host_unregister", "ncl_host_dirty", "ncl_host_fetch", "ncl_host_state",
]


class ExprNode(C.Structure):
    _fields_ = [("op", C.c_int32), ("a", C.c_int32), ("b", C.c_int32), ("pad", C.c_int32),
                ("ival", C.c_int64), ("fval", C.c_double)]


class Expr(C.Structure):
    _fields_ = [("n", C.c_int32), ("pad", C.c_int32), ("node", ExprNode * MAX_EXPR)]


class IterOpt(C.Structure):
    _fields_ = [("key", C.c_int32), ("has_start", C.c_int32), ("has_end", C.c_int32), ("pad", C.c_int32),
                ("start", C.c_int64), ("end", C.c_int64), ("step", C.c_int64)]


class EinsumDesc(C.Structure):
    _fields_ = [("n_in", C.c_int32), ("n_out", C.c_int32), ("n_iter", C.c_int32),
                ("iter", C.c_int32 * MAX_INDEX),
                ("in_rank", C.c_int32 * MAX_OPERANDS),
                ("in_spec", (C.c_int32 * MAX_RANK) * MAX_OPERANDS),
                ("out_rank", C.c_int32 * MAX_OPERANDS),
                ("out_spec", (C.c_int32 * MAX_RANK) * MAX_OPERANDS),
                ("in_nopt", C.c_int32 * MAX_OPERANDS),
                ("out_nopt", C.c_int32 * MAX_OPERANDS),
                ("in_opt", (IterOpt * MAX_OPTS) * MAX_OPERANDS),
                ("out_opt", (IterOpt * MAX_OPTS) * MAX_OPERANDS),
                ("transform", Expr * MAX_OPERANDS)]


class Tensor(C.Structure):
    _fields_ = [("buf", C.c_void_p), ("host", C.c_void_p), ("dtype", C.c_int32), ("rank", C.c_int32),
                ("dims", C.c_int64 * MAX_RANK), ("elem_offset", C.c_int64)]


class NumclCudaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"numcl-cuda error {code}: {msg}")
        self.code = code


class NumclShapeError(NumclCudaError, AssertionError):
    """The reference signals these as failed `assert`s before the loop nest runs."""


class NumclUnsupportedError(NumclCudaError, NotImplementedError):
    pass


class NumclDomainError(NumclCudaError, ArithmeticError):
    """An element left the real domain of its function (sqrt of a negative, division by zero ...): the reference signals a
    Lisp error or produces a complex number there.  Reported at the first synchronisation point after the call."""


_lib = None
_inited = False


def load():
    """dlopen the library (no GPU needed); raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(f"{SO_PATH} is missing: run `python -m numcl_b200.build` (there is no CPU fallback)")
    L = C.CDLL(SO_PATH, mode=C.RTLD_GLOBAL)
    L.ncl_last_error.restype = C.c_char_p
    L.ncl_version.restype = C.c_char_p
    L.ncl_last_kernel.restype = C.c_char_p
    L.ncl_launch_count.restype = C.c_int64
    L.ncl_get_stream.restype = C.c_void_p
    L.ncl_set_stream.argtypes = [C.c_void_p]
    L.ncl_alloc.restype = C.c_void_p
    L.ncl_alloc.argtypes = [C.c_size_t, C.c_int]
    L.ncl_free.argtypes = [C.c_void_p]
    L.ncl_buf_ptr.restype = C.c_void_p
    L.ncl_buf_ptr.argtypes = [C.c_void_p]
    L.ncl_buf_bytes.restype = C.c_size_t
    L.ncl_buf_bytes.argtypes = [C.c_void_p]
    L.ncl_wrap.restype = C.c_void_p
    L.ncl_wrap.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
    L.ncl_h2d.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    L.ncl_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]
    L.ncl_host_alloc_pinned.restype = C.c_void_p
    L.ncl_host_alloc_pinned.argtypes = [C.c_size_t]
    L.ncl_host_free.argtypes = [C.c_void_p]
    L.ncl_storage_bytes.restype = C.c_size_t
    L.ncl_storage_bytes.argtypes = [C.c_int, C.c_int64]
    L.ncl_fill.argtypes = [C.POINTER(Tensor), C.c_int, C.c_int64, C.c_double]
    L.ncl_fill_random.argtypes = [C.POINTER(Tensor), C.c_uint64, C.c_int, C.c_double, C.c_double]
    L.ncl_fill_random_at.argtypes = [C.POINTER(Tensor), C.c_uint64, C.c_uint64, C.c_int, C.c_double, C.c_double]
    L.ncl_transfer_stats.restype = None
    L.ncl_transfer_stats.argtypes = [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.ncl_pipeline_stats.restype = None
    L.ncl_pipeline_stats.argtypes = [C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.ncl_capture_end.argtypes = [C.POINTER(C.c_void_p)]
    L.ncl_graph_launch.argtypes = [C.c_void_p]
    L.ncl_graph_kernels.restype = C.c_int64
    L.ncl_graph_kernels.argtypes = [C.c_void_p]
    L.ncl_graph_free.argtypes = [C.c_void_p]
    L.ncl_einsum.argtypes = [C.POINTER(EinsumDesc), C.POINTER(Tensor), C.c_int, C.POINTER(Tensor), C.c_int, C.c_uint]
    L.ncl_broadcast.argtypes = [C.c_int, C.POINTER(Tensor), C.POINTER(Tensor), C.POINTER(Tensor)]
    L.ncl_fold.argtypes = [C.c_int, C.c_int, C.POINTER(Tensor), C.c_int, C.POINTER(Tensor)]
    L.ncl_map.argtypes = [C.POINTER(Expr), C.POINTER(Tensor), C.c_int, C.POINTER(Tensor)]
    L.ncl_reduce.argtypes = [C.c_int, C.POINTER(Tensor), C.POINTER(C.c_int), C.c_int, C.POINTER(Tensor), C.c_uint]
    L.ncl_convert.argtypes = [C.POINTER(Tensor), C.POINTER(Tensor)]
    L.ncl_host_register.argtypes = [C.c_void_p, C.c_size_t, C.c_uint]
    L.ncl_host_unregister.argtypes = [C.c_void_p]
    L.ncl_host_dirty.argtypes = [C.c_void_p]
    L.ncl_host_fetch.argtypes = [C.c_void_p]
    L.ncl_host_state.argtypes = [C.c_void_p]
    L.ncl_reduce_stat.argtypes = [C.c_int, C.POINTER(Tensor), C.POINTER(C.c_int), C.c_int, C.POINTER(Tensor)]
    L.ncl_comm_unique_id.argtypes = [C.c_void_p]
    L.ncl_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.ncl_allreduce.argtypes = [C.c_int, C.POINTER(Tensor)]
    L.ncl_reduce_all_sharded.argtypes = [C.c_int, C.POINTER(Tensor), C.POINTER(Tensor)]
    L.ncl_reduce_all_sharded_async.argtypes = [C.c_int, C.POINTER(Tensor), C.POINTER(Tensor)]
    _lib = L
    return L


def check(rc):
    if rc == OK:
        return
    msg = load().ncl_last_error().decode(errors="replace")
    if rc == ERR_SHAPE:
        raise NumclShapeError(rc, msg)
    if rc == ERR_UNSUPPORTED:
        raise NumclUnsupportedError(rc, msg)
    if rc == ERR_DOMAIN:
        raise NumclDomainError(rc, msg)
    raise NumclCudaError(rc, msg)


def lib():
    """The initialised library bound to this process's GPU (LOCAL_RANK, default 0)."""
    global _inited
    L = load()
    if not _inited:
        dev = int(os.environ.get("NUMCL_CUDA_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        devs = (C.c_int * 1)(dev)
        check(L.ncl_init(1, devs))
        _inited = True
    return L


def shutdown():
    global _inited
    if _lib is not None and _inited:
        _lib.ncl_shutdown()
        _inited = False


class Graph:
    """A recorded call sequence (ncl_capture_begin .. ncl_capture_end): `with Graph() as g: <calls>`; then g.launch()."""

    def __init__(self):
        self._g = C.c_void_p()

    def __enter__(self):
        check(lib().ncl_capture_begin())
        return self

    def __exit__(self, et, ev, tb):
        rc = lib().ncl_capture_end(C.byref(self._g))
        if et is None:
            check(rc)
        return False

    def launch(self):
        check(lib().ncl_graph_launch(self._g))

    @property
    def kernels(self):
        return lib().ncl_graph_kernels(self._g)

    def __del__(self):
        try:
            if self._g:
                load().ncl_graph_free(self._g)
        except Exception:
            pass


def transfer_stats():
    h, d = C.c_uint64(), C.c_uint64()
    lib().ncl_transfer_stats(C.byref(h), C.byref(d))
    return h.value, d.value


def pipeline_stats():
    c, p = C.c_int64(), C.c_int64()
    lib().ncl_pipeline_stats(C.byref(c), C.byref(p))
    return c.value, p.value
