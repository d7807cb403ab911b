"""Adapter that runs golden cases through the product (numcl_b200 -> C ABI -> CUDA kernels)."""
import numpy as np
import pytest


class GpuBackend:
    def __init__(self, nc, where="device"):
        self.nc = nc
        self.where = where

    def array(self, v, t):
        return self.nc.asarray(np.array(v), type=t, where=self.where)

    def values(self, r):
        return r.numpy() if isinstance(r, self.nc.NArray) else r

    def dtype_name(self, r):
        return r.dtype.name

    def einsum(self, spec, ins, outs):
        res = self.nc.einsum(spec, *ins, *(outs or []))
        return res

    def kron(self, a, b):
        return self.nc.kron(a, b)

    def map_array_into(self, r, fn, ins):
        self.nc.map_array_into(r, fn, *ins)

    def map_array(self, fn, ins):
        return self.nc.map_array(fn, *ins)

    def arith(self, op, args):
        f = {"+": self.nc.add, "-": self.nc.sub, "*": self.nc.mul, "/": self.nc.div,
             "max": self.nc.maximum, "min": self.nc.minimum}[op]
        return f(*args)

    def reduce_array(self, fn, a, axes):
        return self.nc.reduce_array(fn, a, axes=axes)

    def sum(self, a, axes):
        return self.nc.sum(a, axes=axes)

    def plan_broadcast(self, shapes):
        # the plan is internal to ncl_einsum; its observable part is the broadcast result shape
        pytest.skip("plan-broadcast is internal to the backend (covered by the -i -i -> -i case)")

    def expect_normalize_error(self, spec):
        with pytest.raises(self.nc.EinsumSpecError):
            self.nc.einsum_normalize_subscripts(spec)

    def normalize(self, spec):
        return self.nc.einsum_normalize_subscripts(spec).describe()
