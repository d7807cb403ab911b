"""Differential fuzz of the everyday-shape paths against the oracle with fresh seeds (scratch tool, GPU box).

usage: python tools/fuzz_shapes.py [seed] [cases]      (--dry: generate the cases and run only the oracle)

Data are small integers (stored in the array's own element type), so every summation order is exact and every case is
compared bit for bit.  Sizes are chosen so that the dedicated paths trigger (element-granular flat mode, few-column /
short-row / few-long-row reductions, register and thin transposes, dot-like contractions, 32-bit integer lanes)."""
import collections
import sys
import time

sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np

from oracle import oracle as O
import refsem

args = [a for a in sys.argv[1:] if not a.startswith("--")]
DRY = "--dry" in sys.argv
seed0 = int(args[0]) if args else int(time.time()) % 100000
ncases = int(args[1]) if len(args) > 1 else 160
O.build()
if not DRY:
    import numcl_b200 as nc
    from test_gpu_parity import assert_bit_exact, to_gpu

INT = ["fixnum", "ub8", "sb8", "ub15", "sb16", "ub31", "sb32", "sb64"]
FLT = ["f32", "f64"]


def mk(rng, shape, dt, lo=-5, hi=5):
    from numcl_b200.dtypes import dtype
    d = dtype(dt)
    if dt in FLT:
        return O.fill_random(shape, dt, int(rng.integers(1 << 30)), 1, lo, hi)
    return O.fill_random(shape, dt, int(rng.integers(1 << 30)), 0, max(d.lo, lo), min(d.hi, hi))


def case(rng):
    kind = str(rng.choice(["bcast", "bcast", "reduce", "reduce", "transpose", "dot", "cmp"]))
    if kind in ("bcast", "cmp"):
        P = int(rng.choice([2, 3, 5, 6, 7, 12, 31, 33, 96, 100, 257] if kind == "bcast" else [32, 64, 96, 128, 1056, 3, 33]))
        R = int(rng.integers(64, max(65, 300000 // P)))
        sa = (R, P)
        sb = [(P,), (R, 1), (), (1, P), (R, P)][int(rng.integers(5))]
        if rng.random() < 0.3:
            sa, sb = sb, sa
        ta, tb = str(rng.choice(INT + FLT)), str(rng.choice(INT + FLT))
        op = str(rng.choice(["+", "-", "*", "max", "min"])) if kind == "bcast" else str(rng.choice(["</bit", ">=/bit", "=/bit"]))
        return kind, (sa, sb, ta, tb, op)
    if kind == "reduce":
        if rng.random() < 0.7:
            C = int(rng.choice([2, 3, 4, 5, 7, 8, 9, 16, 17, 48, 100, 129, 513])); R = int(rng.integers(200, 400000 // C))
            shape = (R, C) if rng.random() < 0.7 else (C, R)
            axes = (int(rng.integers(2)),)
        else:
            shape = (int(rng.integers(2, 200)), int(rng.integers(2, 60)), int(rng.choice([2, 3, 5, 8, 33])))
            axes = [(0,), (1,), (2,), (0, 1), (0, 2), (1, 2)][int(rng.integers(6))]
        return kind, (shape, axes, str(rng.choice(["+", "max", "min"])), str(rng.choice(INT[:6] + FLT)))
    if kind == "transpose":
        C = int(rng.choice([2, 3, 4, 5, 6, 7, 8, 9, 12, 16, 17])); R = int(rng.integers(128, 200000 // C))
        if rng.random() < 0.6:
            R = R // 4 * 4
        shape = (R, C) if rng.random() < 0.5 else (C, R)
        return kind, (shape, str(rng.choice(["f32", "f64", "fixnum", "ub8", "sb32"])))
    m, k = int(rng.integers(2, 600)), int(rng.integers(2, 3000))
    spec, shapes = [("i i -> ", [(m * k,), (m * k,)]), ("ij ij -> i", [(m, k), (m, k)]), ("ij ij -> j", [(m, k), (m, k)]),
                    ("ij j -> i", [(m, k), (k,)]), ("i ij -> j", [(m,), (m, k)]), ("ij ij -> ", [(m, k), (m, k)]),
                    ("ij j -> i", [(8, 4096 + 4 * int(rng.integers(0, 512))), None])][int(rng.integers(7))]
    if shapes[1] is None:
        shapes = [shapes[0], (shapes[0][1],)]
    return "dot", (spec, shapes, str(rng.choice(["f32", "f64", "fixnum", "ub8"])))


hist, bad, t0 = collections.Counter(), [], time.time()
for c in range(ncases):
    rng = np.random.default_rng(seed0 * 1000 + c)
    kind, a = case(rng)
    try:
        if kind in ("bcast", "cmp"):
            sa, sb, ta, tb, op = a
            x, y = mk(rng, sa, ta, 0 if op == "*" else -5, 5), mk(rng, sb, tb, 0 if op == "*" else -5, 5)
            want = refsem.arith(O, op, [x, y]) if kind == "bcast" else refsem.broadcast(O, op, x, y)
            if not DRY:
                f = {"+": nc.add, "-": nc.sub, "*": nc.mul, "max": nc.maximum, "min": nc.minimum}.get(op)
                got = f(to_gpu(nc, x), to_gpu(nc, y)) if kind == "bcast" else nc.broadcast(op, to_gpu(nc, x), to_gpu(nc, y))
        elif kind == "reduce":
            shape, axes, fn, dt = a
            x = mk(rng, shape, dt)
            want = refsem.reduce_array(O, fn, x, axes)
            if not DRY:
                got = {"+": nc.sum, "max": nc.amax, "min": nc.amin}[fn](to_gpu(nc, x), axes=axes)
        elif kind == "transpose":
            shape, dt = a
            x = mk(rng, shape, dt, 0 if dt == "ub8" else -100, 100)
            want = refsem.einsum(O, "ij -> ji", [x])
            if not DRY:
                got = nc.transpose(to_gpu(nc, x))
        else:
            spec, shapes, dt = a
            ins = [mk(rng, s, dt, 0 if dt == "ub8" else -3, 3) for s in shapes]
            want = refsem.einsum(O, spec, ins)
            if not DRY:
                got = nc.einsum(spec, *[to_gpu(nc, i) for i in ins])
        if DRY:
            hist[kind] += 1
            continue
        hist[kind + ":" + nc.lib().ncl_last_kernel().decode()] += 1
        if want.shape == () and not hasattr(got, "raw"):
            assert got == want.values().reshape(-1)[0], (got, want.values())
        else:
            assert_bit_exact(got, want)
    except Exception as e:                                      # noqa: BLE001
        bad.append((c, kind, a, f"{type(e).__name__}: {str(e)[:200]}"))
print(f"seed {seed0}: {ncases} cases in {time.time() - t0:.1f} s, {len(bad)} failures")
for k, v in sorted(hist.items()):
    print(f"  {v:4d}  {k}")
for b in bad:
    print("FAIL", b)
sys.exit(1 if bad else 0)
