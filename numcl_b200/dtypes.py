"""Element types of numcl arrays as SBCL stores them, and host-side packing.

numcl arrays are specialized CL arrays (doc/DETAILS.org:74-97); the backend reads and writes the
base vector in SBCL's own storage format so that pinned Lisp vectors can be handed over without a
conversion pass:
  * bit / ub2 / ub4 are packed LSB-first,
  * ub7 / ub15 / ub31 live in 8 / 16 / 32-bit slots,
  * fixnum and ub62 hold tagged words (value << 1), sb64 / ub63 / ub64 raw words.
The numbering is include/numcl_cuda.h's ncl_dtype.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class DType:
    name: str
    code: int
    slot_bits: int      # storage bits per element
    value_bits: int     # %coerce wrap width (src/1type.lisp:673-694); 0 for floats
    signed: bool
    tagged: bool
    kind: str           # "int" | "f32" | "f64" | "c64" | "c128"

    @property
    def lo(self):
        if self.kind != "int":
            return None
        return -(1 << (self.value_bits - 1)) if self.signed else 0

    @property
    def hi(self):
        if self.kind != "int":
            return None
        return (1 << (self.value_bits - 1)) - 1 if self.signed else (1 << self.value_bits) - 1


_TABLE = [
    ("bit", 1, 1, False, False), ("ub2", 2, 2, False, False), ("ub4", 4, 4, False, False),
    ("ub7", 8, 7, False, False), ("ub8", 8, 8, False, False), ("ub15", 16, 15, False, False),
    ("ub16", 16, 16, False, False), ("ub31", 32, 31, False, False), ("ub32", 32, 32, False, False),
    ("ub62", 64, 62, False, True), ("ub63", 64, 63, False, False), ("ub64", 64, 64, False, False),
    ("sb8", 8, 8, True, False), ("sb16", 16, 16, True, False), ("sb32", 32, 32, True, False),
    ("fixnum", 64, 63, True, True), ("sb64", 64, 64, True, False),
]
DTYPES: dict[str, DType] = {}
for _i, (_n, _s, _v, _sg, _tg) in enumerate(_TABLE):
    DTYPES[_n] = DType(_n, _i, _s, _v, _sg, _tg, "int")
DTYPES["f32"] = DType("f32", 17, 32, 0, True, False, "f32")
DTYPES["f64"] = DType("f64", 18, 64, 0, True, False, "f64")
# (complex single-float) / (complex double-float): pairs (real, imaginary); only float complexes exist as array element
# types (src/1type.lisp:341-344)
DTYPES["c64"] = DType("c64", 19, 64, 0, True, False, "c64")
DTYPES["c128"] = DType("c128", 20, 128, 0, True, False, "c128")
BY_CODE = {d.code: d for d in DTYPES.values()}

# Lisp spellings accepted by the host mirror
ALIASES = {
    "single-float": "f32", "double-float": "f64", "float": "f32", "(complex single-float)": "c64", "(complex double-float)": "c128", "(unsigned-byte 8)": "ub8",
    "(unsigned-byte 1)": "bit", "(unsigned-byte 2)": "ub2", "(unsigned-byte 4)": "ub4",
    "(unsigned-byte 7)": "ub7", "(unsigned-byte 15)": "ub15", "(unsigned-byte 16)": "ub16",
    "(unsigned-byte 31)": "ub31", "(unsigned-byte 32)": "ub32", "(unsigned-byte 62)": "ub62",
    "(unsigned-byte 63)": "ub63", "(unsigned-byte 64)": "ub64", "(signed-byte 8)": "sb8",
    "(signed-byte 16)": "sb16", "(signed-byte 32)": "sb32", "(signed-byte 64)": "sb64",
}


def dtype(name) -> DType:
    if isinstance(name, DType):
        return name
    name = ALIASES.get(name, name)
    return DTYPES[name]


def storage_elements(n: int) -> int:
    """Base vectors are padded to a multiple of 8 elements (src/1instantiate.lisp:81-84)."""
    return max(8, 8 * ((n + 7) // 8))


def storage_bytes(dt, n: int) -> int:
    dt = dtype(dt)
    return (storage_elements(n) * dt.slot_bits + 7) // 8


_SLOT_NP = {8: (np.uint8, np.int8), 16: (np.uint16, np.int16), 32: (np.uint32, np.int32), 64: (np.uint64, np.int64)}


def wrap(values: np.ndarray, dt) -> np.ndarray:
    """%coerce of integers into dt: wrap to value_bits (ub -> logand, sb -> two's complement)."""
    dt = dtype(dt)
    v = np.asarray(values).astype(np.int64)
    w = dt.value_bits
    if w >= 64:
        return v
    u = v.view(np.uint64) if v.ndim else np.uint64(int(v) & 0xFFFFFFFFFFFFFFFF)
    mask = np.uint64((1 << w) - 1)
    if dt.signed:                                   # two's complement wrap in unsigned arithmetic (no overflow warnings)
        half = np.uint64(1 << (w - 1))
        return (((u + half) & mask) - half).astype(np.int64) if isinstance(u, np.ndarray) else np.int64(int(((u + half) & mask)) - int(half))
    return (u & mask).astype(np.int64) if isinstance(u, np.ndarray) else np.int64(int(u & mask))


def pack(values: np.ndarray, dt) -> np.ndarray:
    """Logical values -> base vector bytes (uint8 array of storage_bytes length)."""
    dt = dtype(dt)
    flat = np.ascontiguousarray(np.asarray(values)).reshape(-1)
    n = flat.size
    out = np.zeros(storage_bytes(dt, n), dtype=np.uint8)
    if dt.kind == "f32":
        out[: n * 4] = flat.astype(np.float32).view(np.uint8)
        return out
    if dt.kind == "f64":
        out[: n * 8] = flat.astype(np.float64).view(np.uint8)
        return out
    if dt.kind in ("c64", "c128"):
        out[: n * dt.slot_bits // 8] = flat.astype(np.complex64 if dt.kind == "c64" else np.complex128).view(np.uint8)
        return out
    if flat.dtype.kind == "f":
        flat = np.rint(flat)                    # (round x): half to even
    v = wrap(flat, dt)
    if dt.slot_bits < 8:
        per = 8 // dt.slot_bits
        padded = np.zeros(((n + per - 1) // per) * per, dtype=np.uint8)
        padded[:n] = v.astype(np.uint8)
        padded = padded.reshape(-1, per)
        byte = np.zeros(padded.shape[0], dtype=np.uint8)
        for k in range(per):
            byte |= (padded[:, k] << np.uint8(k * dt.slot_bits)).astype(np.uint8)
        out[: byte.size] = byte
        return out
    if dt.tagged:
        v = v << np.int64(1)
    npdt = _SLOT_NP[dt.slot_bits][0]
    out[: n * dt.slot_bits // 8] = v.astype(np.int64).view(np.uint64).astype(npdt).view(np.uint8)
    return out


def unpack(storage: np.ndarray, n: int, dt, offset: int = 0) -> np.ndarray:
    """Base vector bytes -> logical values (float32 / float64 / int64; ub64 above 2^63 wraps)."""
    dt = dtype(dt)
    storage = np.ascontiguousarray(storage).view(np.uint8)
    if dt.kind == "f32":
        return storage[: (n + offset) * 4].view(np.float32)[offset:].copy()
    if dt.kind == "f64":
        return storage[: (n + offset) * 8].view(np.float64)[offset:].copy()
    if dt.kind in ("c64", "c128"):
        return storage[: (n + offset) * dt.slot_bits // 8].view(np.complex64 if dt.kind == "c64" else np.complex128)[offset:].copy()
    if dt.slot_bits < 8:
        per = 8 // dt.slot_bits
        idx = np.arange(offset, offset + n)
        byte = storage[idx // per]
        return ((byte >> ((idx % per) * dt.slot_bits).astype(np.uint8)) & np.uint8((1 << dt.slot_bits) - 1)).astype(np.int64)
    u, s = _SLOT_NP[dt.slot_bits]
    raw = storage[: (n + offset) * dt.slot_bits // 8].view(s if dt.signed or dt.tagged else u)[offset:]
    v = raw.astype(np.int64) if dt.slot_bits < 64 else raw.view(np.int64).copy()
    if dt.tagged:
        v = v >> np.int64(1)
    return v
