"""One slow-path case per invocation, launched twice (ncu: -s 1 -c 1 captures the warm one).
usage: python tools/prof_r2.py <case>"""
import sys
sys.path.insert(0, ".")
import numcl_b200 as nc
from numcl_b200 import numeric as nn

case = sys.argv[1]
L = nc.lib()


def R(shape, dt="f32", seed=1):
    return nc.random_array(shape, dt, seed, 0, -1.0, 1.0) if dt in ("f32", "f64") else nc.random_array(shape, dt, seed, 0, 0, 100)


if case == "cmp_f32":
    a = R((1024, 512, 512)); b = R((1024, 1, 512), seed=4); f = lambda: nn.lt(a, b)
elif case == "cmp_u8":
    a = R((1024, 512, 512), "ub8", 7); b = R((1024, 512, 512), "ub8", 8); f = lambda: nn.lt(a, b)
elif case == "max_u8":
    a = R((1024, 512, 512), "ub8", 7); b = R((1024, 512, 512), "ub8", 8); f = lambda: nc.maximum(a, b)
elif case == "add_u8":
    a = R((1024, 512, 512), "ub8", 7); b = R((1024, 512, 512), "ub8", 8); f = lambda: nc.add(a, b)
elif case == "col_ragged":
    a = R((32767, 8193)); f = lambda: nc.sum(a, axes=0)
elif case == "dot_rows":
    a = R((32768, 8192), seed=9); b = R((32768, 8192), seed=10); f = lambda: nc.einsum("ij ij -> i", a, b)
elif case == "dot_cols":
    a = R((32768, 8192), seed=9); b = R((32768, 8192), seed=10); f = lambda: nc.einsum("ij ij -> j", a, b)
elif case == "vecmat":
    a = R((32768, 8192), seed=9); b = R((32768,), seed=15); f = lambda: nc.einsum("i ij -> j", b, a)
elif case == "col_fixnum":
    a = R((8192, 8192), "fixnum", 9); f = lambda: nc.sum(a, axes=0)
elif case == "sum_u8_all":
    a = R((1024, 512, 512), "ub8", 7); f = lambda: nc.sum(a)
elif case == "transpose_thin":
    a = R((1 << 22, 48)); f = lambda: nc.transpose(a)
elif case == "c3d":
    a = R((8192, 8192), "f64", 8); b = R((8192, 8192), "f64", 9); f = lambda: nc.matmul(a, b)
elif case == "c3s":
    a = R((8192, 8192), seed=8); b = R((8192, 8192), seed=9); f = lambda: nc.matmul(a, b)
elif case == "c4":
    a = R((256, 512, 512), seed=6); b = R((256, 512, 512), seed=7); f = lambda: nc.einsum("bij bjk -> bik", a, b)
elif case == "bitmul":
    a = R((1 << 26,)); b = R((1 << 26,), seed=2); m = nn.lt(a, b); f = lambda: nc.mul(m, a)
elif case == "fill":
    f = lambda: nc.full((16384, 16384), 1.5, type="f32")
elif case == "scale":
    a = R((16384, 16384), seed=5); f = lambda: nc.mul(a, 2.0)
elif case == "outer":
    a = R((16384,), seed=5); b = R((16384,), seed=6); f = lambda: nc.outer(a, b)
elif case == "sub_N3_N1":
    a = R((1 << 26, 3)); w = R((1 << 26, 1), seed=3); f = lambda: nc.mul(a, w)
elif case == "rows_reg":
    a = R((1 << 26, 3)); f = lambda: nc.sum(a, axes=1)
elif case == "sqrt":
    a = nc.random_array((1024, 512, 512), "f32", 3, 0, 0.0, 4.0); f = lambda: nn.sqrt(a)
else:
    raise SystemExit("unknown case")
r = f(); L.ncl_sync()
r = f(); L.ncl_sync()
print(case, L.ncl_last_kernel().decode())
