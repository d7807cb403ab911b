"""numcl_b200 -- Python host mirror of numcl's einsum / broadcast / reduce API over the B200 backend.

The product is numcl_b200/libnumcl_cuda.so (C ABI in include/numcl_cuda.h, kernels in csrc/).
These modules mirror the reference's operator interface for the hot path -- same names, argument
meaning and error behaviour -- so the parity tests read like numcl's own tests.  No CPU fallback.
"""
from ._lib import (Graph, NumclCudaError, NumclDomainError, NumclShapeError, NumclUnsupportedError, lib, load, shutdown,  # noqa: F401
                   transfer_stats)
from .array import (NArray, arange, asarray, astype, copy, empty, full, ones, random_array, zeros)  # noqa: F401
from .einsum import einsum  # noqa: F401
from .einsum_spec import EinsumSpecError, einsum_normalize_subscripts  # noqa: F401
from .linalg import diag, eye, inner, kron, matmul, matmul_star, matrix_chain_order, outer, transpose, vdot  # noqa: F401
from .numeric import (add, broadcast, clip, div, map_array, map_array_into, maximum, minimum, mul,  # noqa: F401
                      one_minus, one_plus, sub)
from .numeric import (acos, acosh, asin, asinh, atan, atanh, ceiling, conjugate, cos, cosh, eq, exp, expt, floor, ge, gt, isqrt, le,  # noqa: F401
                      log, log2, logand, logandc1, logandc2, logc, logeqv, logior, lognand, lognor, lognot, logorc1,
                      logorc2, logxor, lt, mod, ne, imagpart, realpart, rem, round_, sin, sinh, sqrt, square, tan, tanh, truncate)
from .reduce import amax, amin, avg, mean, prod, reduce_array, stdev, var, variance  # noqa: F401
from .reduce import sum  # noqa: F401,A004
