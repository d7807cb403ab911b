// ncl_device.cuh -- device-side building blocks shared by all kernel families.
//
// Semantics are Common Lisp's as SBCL x86-64 compiles them:
//   * float + - * / are IEEE round-to-nearest-even, each op rounded separately: the __f*_rn /
//     __d*_rn intrinsics are never contracted into FMA, and the library is built without
//     --use_fast_math / -ftz so denormals survive;
//   * integer ops are exact in 64-bit two's complement and wrapped only by the final %coerce
//     (src/1type.lisp:673-694), which commutes with + - * (mod 2^w);
//   * (max a b) = (if (>= a b) a b), (min a b) = (if (<= a b) a b).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "ncl_internal.h"

// ---- load / store kinds: what the host canonicalises an ncl_dtype to ---------------------------
enum LoadKind { LK_U8 = 0, LK_S8, LK_U16, LK_S16, LK_U32, LK_S32, LK_W64, LK_T64, LK_F32, LK_F64,
                LK_P1, LK_P2, LK_P4 };
enum StoreKind { SK_8 = 0, SK_16, SK_32, SK_64, SK_F32, SK_F64, SK_P1, SK_P2, SK_P4 };

struct StoreSpec { int kind; int shift; unsigned long long mask; };   // int: slot = (v & mask) << shift

__host__ __device__ inline int lk_bits(int lk) {
  switch (lk) { case LK_U8: case LK_S8: return 8; case LK_U16: case LK_S16: return 16;
    case LK_U32: case LK_S32: case LK_F32: return 32; case LK_W64: case LK_T64: case LK_F64: return 64;
    case LK_P1: return 1; case LK_P2: return 2; default: return 4; }
}
__host__ __device__ inline int sk_bits(int sk) {
  switch (sk) { case SK_8: return 8; case SK_16: return 16; case SK_32: case SK_F32: return 32;
    case SK_64: case SK_F64: return 64; case SK_P1: return 1; case SK_P2: return 2; default: return 4; }
}
inline int dtype_load_kind(int dt) {
  switch (dt) {
    case NCL_BIT: return LK_P1; case NCL_UB2: return LK_P2; case NCL_UB4: return LK_P4;
    case NCL_UB7: case NCL_UB8: return LK_U8; case NCL_UB15: case NCL_UB16: return LK_U16;
    case NCL_UB31: case NCL_UB32: return LK_U32; case NCL_UB62: case NCL_FIXNUM: return LK_T64;
    case NCL_UB63: case NCL_UB64: case NCL_SB64: return LK_W64;   // ub64 >= 2^63 reads as negative: documented limit
    case NCL_SB8: return LK_S8; case NCL_SB16: return LK_S16; case NCL_SB32: return LK_S32;
    case NCL_F32: return LK_F32; default: return LK_F64;
  }
}
inline StoreSpec dtype_store_spec(int dt) {
  const DtInfo& d = g_dt[dt];
  StoreSpec s; s.shift = d.tagged ? 1 : 0;
  s.mask = (d.is_signed || d.value_bits >= 64) ? ~0ull : ((1ull << d.value_bits) - 1);
  if (d.cls == CLS_F) s.kind = SK_F32; else if (d.cls == CLS_D) s.kind = SK_F64;
  else switch (d.slot_bits) { case 1: s.kind = SK_P1; break; case 2: s.kind = SK_P2; break; case 4: s.kind = SK_P4; break;
    case 8: s.kind = SK_8; break; case 16: s.kind = SK_16; break; case 32: s.kind = SK_32; break; default: s.kind = SK_64; }
  return s;
}

// ---- class value types -----------------------------------------------------------------------------
template <int C> struct ClsT;
template <> struct ClsT<CLS_I> { typedef long long T; };
template <> struct ClsT<CLS_F> { typedef float T; };
template <> struct ClsT<CLS_D> { typedef double T; };

template <class T> __device__ __forceinline__ T from_i64(long long v);
template <> __device__ __forceinline__ long long from_i64<long long>(long long v) { return v; }
template <> __device__ __forceinline__ int from_i64<int>(long long v) { return (int)v; }     // 32-bit lane class: values mod 2^32
template <> __device__ __forceinline__ float from_i64<float>(long long v) { return __ll2float_rn(v); }
template <> __device__ __forceinline__ double from_i64<double>(long long v) { return __ll2double_rn(v); }
template <class T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ long long from_f32<long long>(float v) { return (long long)v; }   // unreachable by typing
template <> __device__ __forceinline__ int from_f32<int>(float v) { return (int)v; }                        // unreachable by typing
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ double from_f32<double>(float v) { return (double)v; }
template <class T> __device__ __forceinline__ T from_f64(double v);
template <> __device__ __forceinline__ long long from_f64<long long>(double v) { return (long long)v; }   // unreachable
template <> __device__ __forceinline__ int from_f64<int>(double v) { return (int)v; }                      // unreachable
template <> __device__ __forceinline__ float from_f64<float>(double v) { return __double2float_rn(v); }   // unreachable
template <> __device__ __forceinline__ double from_f64<double>(double v) { return v; }

// scalar element load of any kind, converted to class type T (the int->float conversion is the
// contagion CL performs before a mixed op: round-to-nearest)
template <class T>
__device__ __forceinline__ T load_scalar(const char* base, long long idx, int lk) {
  switch (lk) {
    case LK_U8:  return from_i64<T>(((const unsigned char*)base)[idx]);
    case LK_S8:  return from_i64<T>(((const signed char*)base)[idx]);
    case LK_U16: return from_i64<T>(((const unsigned short*)base)[idx]);
    case LK_S16: return from_i64<T>(((const short*)base)[idx]);
    case LK_U32: return from_i64<T>(((const unsigned int*)base)[idx]);
    case LK_S32: return from_i64<T>(((const int*)base)[idx]);
    case LK_W64: return from_i64<T>(((const long long*)base)[idx]);
    case LK_T64: return from_i64<T>(((const long long*)base)[idx] >> 1);
    case LK_F32: return from_f32<T>(((const float*)base)[idx]);
    case LK_F64: return from_f64<T>(((const double*)base)[idx]);
    case LK_P1:  return from_i64<T>((((const unsigned char*)base)[idx >> 3] >> (idx & 7)) & 1);
    case LK_P2:  return from_i64<T>((((const unsigned char*)base)[idx >> 2] >> ((idx & 3) * 2)) & 3);
    default:     return from_i64<T>((((const unsigned char*)base)[idx >> 1] >> ((idx & 1) * 4)) & 15);
  }
}

// the same with the element kind known at compile time (callers hoist the kind switch out of their element loops)
template <class T, int LK>
__device__ __forceinline__ T load_scalar_k(const char* base, long long idx) {
  if constexpr (LK == LK_U8)  return from_i64<T>(((const unsigned char*)base)[idx]);
  else if constexpr (LK == LK_S8)  return from_i64<T>(((const signed char*)base)[idx]);
  else if constexpr (LK == LK_U16) return from_i64<T>(((const unsigned short*)base)[idx]);
  else if constexpr (LK == LK_S16) return from_i64<T>(((const short*)base)[idx]);
  else if constexpr (LK == LK_U32) return from_i64<T>(((const unsigned int*)base)[idx]);
  else if constexpr (LK == LK_S32) return from_i64<T>(((const int*)base)[idx]);
  else if constexpr (LK == LK_W64) return from_i64<T>(((const long long*)base)[idx]);
  else if constexpr (LK == LK_T64) return from_i64<T>(((const long long*)base)[idx] >> 1);
  else if constexpr (LK == LK_F32) return from_f32<T>(((const float*)base)[idx]);
  else if constexpr (LK == LK_F64) return from_f64<T>(((const double*)base)[idx]);
  else if constexpr (LK == LK_P1)  return from_i64<T>((((const unsigned char*)base)[idx >> 3] >> (idx & 7)) & 1);
  else if constexpr (LK == LK_P2)  return from_i64<T>((((const unsigned char*)base)[idx >> 2] >> ((idx & 3) * 2)) & 3);
  else return from_i64<T>((((const unsigned char*)base)[idx >> 1] >> ((idx & 1) * 4)) & 15);
}
template <int K> struct LkTag { static constexpr int value = K; };
// calls f(LkTag<kind>()) for the runtime kind lk: one warp-uniform switch, the body compiled per kind
template <class F> __device__ __forceinline__ void dispatch_lk(int lk, F f) {
  switch (lk) {
    case LK_U8: f(LkTag<LK_U8>()); break;   case LK_S8: f(LkTag<LK_S8>()); break;
    case LK_U16: f(LkTag<LK_U16>()); break; case LK_S16: f(LkTag<LK_S16>()); break;
    case LK_U32: f(LkTag<LK_U32>()); break; case LK_S32: f(LkTag<LK_S32>()); break;
    case LK_W64: f(LkTag<LK_W64>()); break; case LK_T64: f(LkTag<LK_T64>()); break;
    case LK_F32: f(LkTag<LK_F32>()); break; case LK_F64: f(LkTag<LK_F64>()); break;
    case LK_P1: f(LkTag<LK_P1>()); break;   case LK_P2: f(LkTag<LK_P2>()); break;
    default: f(LkTag<LK_P4>()); break;
  }
}

// %coerce of a class value into an integer slot (src/1type.lisp:673-694): (round x) half-even, then wrap.
__device__ __forceinline__ long long round_to_i64(long long v) { return v; }
__device__ __forceinline__ long long round_to_i64(int v) { return v; }
// an integer-valued double of any magnitude -> its value modulo 2^64 (what %coerce keeps after wrapping)
__device__ __forceinline__ long long f64_integer_to_i64_wrap(double v) {
  if (fabs(v) < 9223372036854775808.0) return (long long)v;
  if (!(fabs(v) <= 1.7976931348623157e308)) return 0;                       // inf / nan: the reference signals
  const long long bits = __double_as_longlong(v);
  const int e = (int)((bits >> 52) & 0x7ff) - 1075;                          // v = m * 2^e, e >= 11 here
  const unsigned long long m = ((unsigned long long)bits & 0xfffffffffffffull) | 0x10000000000000ull;
  const unsigned long long mag = e >= 64 ? 0ull : (m << e);
  return (long long)(bits < 0 ? 0ull - mag : mag);
}
// (round x) of a float of ANY magnitude, then modulo 2^64 -- the conversion intrinsics saturate at +-2^63, while the reference
// rounds to the exact (big) integer and %coerce wraps it
__device__ __forceinline__ long long round_to_i64(double v) { return fabs(v) < 9223372036854775808.0 ? __double2ll_rn(v) : f64_integer_to_i64_wrap(v); }
__device__ __forceinline__ long long round_to_i64(float v) { return round_to_i64((double)v); }
__device__ __forceinline__ float to_f32(long long v) { return __ll2float_rn(v); }
__device__ __forceinline__ float to_f32(int v) { return __int2float_rn(v); }
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(double v) { return __double2float_rn(v); }
__device__ __forceinline__ double to_f64(long long v) { return __ll2double_rn(v); }
__device__ __forceinline__ double to_f64(int v) { return (double)v; }
__device__ __forceinline__ double to_f64(float v) { return (double)v; }
__device__ __forceinline__ double to_f64(double v) { return v; }

// scalar store; sub-byte slots are updated with word atomics so neighbouring threads may own
// neighbouring fields of the same word.
template <class T>
__device__ __forceinline__ void store_scalar(char* base, long long idx, StoreSpec s, T v) {
  if (s.kind == SK_F32) { ((float*)base)[idx] = to_f32(v); return; }
  if (s.kind == SK_F64) { ((double*)base)[idx] = to_f64(v); return; }
  unsigned long long w = (((unsigned long long)round_to_i64(v)) & s.mask) << s.shift;
  switch (s.kind) {
    case SK_8:  ((unsigned char*)base)[idx] = (unsigned char)w; break;
    case SK_16: ((unsigned short*)base)[idx] = (unsigned short)w; break;
    case SK_32: ((unsigned int*)base)[idx] = (unsigned int)w; break;
    case SK_64: ((unsigned long long*)base)[idx] = w; break;
    default: {
      int bits = (s.kind == SK_P1) ? 1 : (s.kind == SK_P2) ? 2 : 4;
      long long bitpos = idx * bits;
      unsigned int* word = (unsigned int*)base + (bitpos >> 5);
      int sh = (int)(bitpos & 31);
      unsigned int m = ((1u << bits) - 1u) << sh;
      atomicAnd(word, ~m);
      atomicOr(word, ((unsigned int)w << sh) & m);
    }
  }
}

// ---- correctly rounded (coerce (/ n d) 'single-float) for integers n, d ----------------------------
// CL produces the exact ratio and rounds once.  Fast path: both exactly representable in double;
// quotient rounded toward zero, made odd when inexact (sticky), then one RN to 24 bits.
__device__ inline float ratio_to_f32(long long n, long long d) {
  if (d == 0) return (n > 0) ? __int_as_float(0x7f800000) : (n < 0 ? __int_as_float(0xff800000) : __int_as_float(0x7fc00000));
  const long long lim = 1ll << 53;
  if (n > -lim && n < lim && d > -lim && d < lim) {
    double dn = (double)n, dd = (double)d;
    double q = __ddiv_rz(dn, dd);
    double r = __fma_rn(q, dd, -dn);                 // exact residual: q*d - n
    if (r != 0.0) q = __longlong_as_double(__double_as_longlong(q) | 1ll);
    return __double2float_rn(q);
  }
  // slow path: binary long division for 26 quotient bits + sticky
  bool neg = (n < 0) != (d < 0);
  unsigned long long un = n < 0 ? (unsigned long long)(-(n + 1)) + 1ull : (unsigned long long)n;
  unsigned long long ud = d < 0 ? (unsigned long long)(-(d + 1)) + 1ull : (unsigned long long)d;
  if (un == 0) return 0.0f;
  int bn = 64 - __clzll(un), bd = 64 - __clzll(ud);
  int s = 26 + bd - bn;                              // quotient of (un << s) / ud has 26 or 27 bits
  // compute floor(un * 2^s / ud) by restoring division on the bits of un followed by s zero bits
  unsigned long long rem = 0, q = 0;
  int total = bn + (s > 0 ? s : 0);
  for (int i = 0; i < total; i++) {
    int bit = (i < bn) ? (int)((un >> (bn - 1 - i)) & 1ull) : 0;
    bool carry = (rem >> 63) != 0;                   // rem < ud <= 2^64-1; doubling may overflow
    rem = (rem << 1) | (unsigned long long)bit;
    q <<= 1;
    if (carry || rem >= ud) { rem -= ud; q |= 1ull; }
  }
  int exp2 = -(s > 0 ? s : 0);
  if (s < 0) {                                        // drop -s low quotient bits into sticky
    unsigned long long low = q & ((1ull << (-s)) - 1ull);
    q >>= (-s); if (low) rem = 1; exp2 = -s;
  }
  if (rem) q |= 1ull;
  float f = __ull2float_rn(q);
  f = ldexpf(f, exp2);
  return neg ? -f : f;
}

// ---- scalar functions of the whitelist, per class ------------------------------------------------------
template <class T> struct Ops;
template <> struct Ops<long long> {
  typedef long long T; typedef unsigned long long U;
  static __device__ __forceinline__ T add(T a, T b) { return (T)((U)a + (U)b); }
  static __device__ __forceinline__ T sub(T a, T b) { return (T)((U)a - (U)b); }
  static __device__ __forceinline__ T mul(T a, T b) { return (T)((U)a * (U)b); }
  static __device__ __forceinline__ T mx(T a, T b) { return a >= b ? a : b; }
  static __device__ __forceinline__ T mn(T a, T b) { return a <= b ? a : b; }
  static __device__ __forceinline__ T neg(T a) { return (T)(0ull - (U)a); }
  static __device__ __forceinline__ T abs_(T a) { return a < 0 ? neg(a) : a; }
  static __device__ __forceinline__ T zero() { return 0; }
};
// 32-bit lanes for integer arrays whose slots (inputs and result) are at most 32 bits wide: + - * and the bitwise functions
// are exact modulo 2^32, which is all a store into a <= 32-bit slot keeps (%coerce masks / wraps to the slot's width);
// max / min / abs / comparisons are only routed here when every operand fits a signed 32-bit value.
template <> struct Ops<int> {
  typedef int T; typedef unsigned int U;
  static __device__ __forceinline__ T add(T a, T b) { return (T)((U)a + (U)b); }
  static __device__ __forceinline__ T sub(T a, T b) { return (T)((U)a - (U)b); }
  static __device__ __forceinline__ T mul(T a, T b) { return (T)((U)a * (U)b); }
  static __device__ __forceinline__ T mx(T a, T b) { return a >= b ? a : b; }
  static __device__ __forceinline__ T mn(T a, T b) { return a <= b ? a : b; }
  static __device__ __forceinline__ T neg(T a) { return (T)(0u - (U)a); }
  static __device__ __forceinline__ T abs_(T a) { return a < 0 ? neg(a) : a; }
  static __device__ __forceinline__ T zero() { return 0; }
};
template <> struct Ops<float> {
  typedef float T;
  static __device__ __forceinline__ T add(T a, T b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ T sub(T a, T b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ T mul(T a, T b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ T div(T a, T b) { return __fdiv_rn(a, b); }
  static __device__ __forceinline__ T mx(T a, T b) { return a >= b ? a : b; }
  static __device__ __forceinline__ T mn(T a, T b) { return a <= b ? a : b; }
  static __device__ __forceinline__ T neg(T a) { return -a; }
  static __device__ __forceinline__ T abs_(T a) { return fabsf(a); }
  static __device__ __forceinline__ T zero() { return 0.0f; }
};
template <> struct Ops<double> {
  typedef double T;
  static __device__ __forceinline__ T add(T a, T b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ T sub(T a, T b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ T mul(T a, T b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ T div(T a, T b) { return __ddiv_rn(a, b); }
  static __device__ __forceinline__ T mx(T a, T b) { return a >= b ? a : b; }
  static __device__ __forceinline__ T mn(T a, T b) { return a <= b ? a : b; }
  static __device__ __forceinline__ T neg(T a) { return -a; }
  static __device__ __forceinline__ T abs_(T a) { return fabs(a); }
  static __device__ __forceinline__ T zero() { return 0.0; }
};

#include <type_traits>
// ---- vector chunk I/O ----------------------------------------------------------------------------------
template <int NBYTES> struct Chunk { unsigned int w[(NBYTES + 3) / 4]; };

template <int NBYTES> __device__ __forceinline__ void ld_chunk(Chunk<NBYTES>& c, const char* p) {
  if constexpr (NBYTES >= 16) {
#pragma unroll
    for (int i = 0; i < NBYTES / 16; i++) {
      uint4 v = __ldg((const uint4*)p + i);
      c.w[4 * i] = v.x; c.w[4 * i + 1] = v.y; c.w[4 * i + 2] = v.z; c.w[4 * i + 3] = v.w;
    }
  } else if constexpr (NBYTES == 8) { uint2 v = __ldg((const uint2*)p); c.w[0] = v.x; c.w[1] = v.y; }
  else if constexpr (NBYTES == 4) c.w[0] = __ldg((const unsigned int*)p);
  else if constexpr (NBYTES == 2) c.w[0] = __ldg((const unsigned short*)p);
  else c.w[0] = __ldg((const unsigned char*)p);
}
template <int NBYTES> __device__ __forceinline__ void st_chunk(const Chunk<NBYTES>& c, char* p) {
  if constexpr (NBYTES >= 16) {
#pragma unroll
    for (int i = 0; i < NBYTES / 16; i++) ((uint4*)p)[i] = make_uint4(c.w[4 * i], c.w[4 * i + 1], c.w[4 * i + 2], c.w[4 * i + 3]);
  } else if constexpr (NBYTES == 8) *(uint2*)p = make_uint2(c.w[0], c.w[1]);
  else if constexpr (NBYTES == 4) *(unsigned int*)p = c.w[0];
  else if constexpr (NBYTES == 2) *(unsigned short*)p = (unsigned short)c.w[0];
  else *(unsigned char*)p = (unsigned char)c.w[0];
}
// element i of a chunk of BITS-wide slots
template <int BITS, int NBYTES> __device__ __forceinline__ unsigned long long chunk_get(const Chunk<NBYTES>& c, int i) {
  if constexpr (BITS == 64) return ((unsigned long long)c.w[2 * i + 1] << 32) | c.w[2 * i];
  else if constexpr (BITS == 32) return c.w[i];
  else if constexpr (BITS == 16) return (c.w[i >> 1] >> ((i & 1) * 16)) & 0xffffu;
  else return (c.w[i >> 2] >> ((i & 3) * 8)) & 0xffu;
}
template <int BITS, int NBYTES> __device__ __forceinline__ void chunk_put(Chunk<NBYTES>& c, int i, unsigned long long v) {
  if constexpr (BITS == 64) { c.w[2 * i] = (unsigned int)v; c.w[2 * i + 1] = (unsigned int)(v >> 32); }
  else if constexpr (BITS == 32) c.w[i] = (unsigned int)v;
  else if constexpr (BITS == 16) { if ((i & 1) == 0) c.w[i >> 1] = (unsigned int)(v & 0xffffu); else c.w[i >> 1] |= (unsigned int)(v & 0xffffu) << 16; }
  else { if ((i & 3) == 0) c.w[i >> 2] = (unsigned int)(v & 0xffu); else c.w[i >> 2] |= (unsigned int)(v & 0xffu) << ((i & 3) * 8); }
}

template <class T, int E, int BITS, bool SIGNED, bool TAGGED>
__device__ __forceinline__ void load_int_chunk(const char* base, long long idx, T (&v)[E]) {
  Chunk<E * BITS / 8> c; ld_chunk(c, base + idx * (BITS / 8));
#pragma unroll
  for (int i = 0; i < E; i++) {
    unsigned long long u = chunk_get<BITS>(c, i);
    long long s;
    if constexpr (BITS == 64) s = TAGGED ? ((long long)u >> 1) : (long long)u;
    else if constexpr (SIGNED) s = (long long)((long long)(u << (64 - BITS)) >> (64 - BITS));
    else s = (long long)u;
    v[i] = from_i64<T>(s);
  }
}

// E <= 32 consecutive elements of a packed BIT array starting at element idx: one 32-bit word, or two when the run straddles a
// word boundary (the second word then holds elements of this very run, so it lies inside the array).  Bit arrays are what
// every numcl comparison returns; they used to be read one element per thread (mask * array: 0.03 of the HBM peak).
template <class T, int E>
__device__ __forceinline__ void load_bits(const char* base, long long idx, T (&v)[E]) {
  static_assert(E <= 32, "at most one result word per chunk");
  const unsigned* w = reinterpret_cast<const unsigned*>(base) + (idx >> 5);
  const unsigned sh = (unsigned)idx & 31u;
  // two loads, no branch (a branch around the second load kept the loads of an unrolled step one behind the other); when the
  // run fits one word the second address is the first one again
  const unsigned lo = __ldg(w), hi = __ldg(w + ((sh + E > 32) ? 1 : 0));
  const unsigned bits = __funnelshift_r(lo, hi, sh);
#pragma unroll
  for (int i = 0; i < E; i++) v[i] = from_i64<T>((long long)((bits >> i) & 1u));
}
// the same in two steps, so that a caller can issue the loads of several chunks before the first conversion; B = bits per
// element (1, 2, 4: bit, ub2, ub4 arrays), E * B <= 32
template <int E, int B = 1> __device__ __forceinline__ unsigned load_bits_raw(const char* base, long long idx) {
  static_assert(E * B <= 32, "one word's worth of packed elements per chunk");
  const long long bit = idx * B;
  const unsigned* w = reinterpret_cast<const unsigned*>(base) + (bit >> 5);
  const unsigned sh = (unsigned)bit & 31u;
  return __funnelshift_r(__ldg(w), __ldg(w + ((sh + E * B > 32) ? 1 : 0)), sh);
}
template <class T, int E, int B = 1> __device__ __forceinline__ void bits_to_elems(unsigned bits, T (&v)[E]) {
#pragma unroll
  for (int i = 0; i < E; i++) v[i] = from_i64<T>((long long)((bits >> (i * B)) & ((1u << B) - 1u)));
}
// E consecutive elements starting at element idx (innermost stride 1), or one element splat (stride 0)
template <class T, int E>
__device__ __forceinline__ void load_elems(const char* base, long long idx, int lk, bool contiguous, T (&v)[E]) {
  if (!contiguous || E == 1) {
    T s = load_scalar<T>(base, idx, lk);
#pragma unroll
    for (int i = 0; i < E; i++) v[i] = s;
    return;
  }
  if constexpr (E > 1) {
    switch (lk) {
      case LK_U8:  load_int_chunk<T, E, 8, false, false>(base, idx, v); break;
      case LK_S8:  load_int_chunk<T, E, 8, true, false>(base, idx, v); break;
      case LK_U16: load_int_chunk<T, E, 16, false, false>(base, idx, v); break;
      case LK_S16: load_int_chunk<T, E, 16, true, false>(base, idx, v); break;
      case LK_U32: load_int_chunk<T, E, 32, false, false>(base, idx, v); break;
      case LK_S32: load_int_chunk<T, E, 32, true, false>(base, idx, v); break;
      case LK_W64: load_int_chunk<T, E, 64, true, false>(base, idx, v); break;
      case LK_T64: load_int_chunk<T, E, 64, true, true>(base, idx, v); break;
      case LK_F32: { Chunk<E * 4> c; ld_chunk(c, base + idx * 4);
#pragma unroll
        for (int i = 0; i < E; i++) v[i] = from_f32<T>(__uint_as_float(c.w[i])); break; }
      case LK_P1: if constexpr (E <= 32) load_bits<T, E>(base, idx, v); break;
      case LK_P2: if constexpr (E <= 16) bits_to_elems<T, E, 2>(load_bits_raw<E, 2>(base, idx), v); break;
      case LK_P4: if constexpr (E <= 8) bits_to_elems<T, E, 4>(load_bits_raw<E, 4>(base, idx), v); break;
      default: { Chunk<E * 8> c; ld_chunk(c, base + idx * 8);       // LK_F64 (ub2 / ub4 never reach the fast path)
#pragma unroll
        for (int i = 0; i < E; i++) v[i] = from_f64<T>(__longlong_as_double((long long)chunk_get<64>(c, i))); break; }
    }
  }
}

// The same 16-byte chunk with the element kind reduced to two warp-uniform flags (bit 0: signed slots, bit 1: tagged
// words) and NO branch: within one (value class, E) instance the slot width is fixed, so a chunk is one LDG.128 followed by
// selects / shifts.  The kernels that stream (reductions) use this instead of load_elems: with the kind switch of
// load_elems around every load, consecutive loads of an unrolled step were ~30 instructions apart and, for integer
// kinds, each was followed directly by its conversion -- ONE load in flight per thread (profiles/r01_sass_reduce_row_kind_switch.txt).
__host__ __device__ inline unsigned lk_flags(int lk) {
  return ((lk == LK_S8 || lk == LK_S16 || lk == LK_S32 || lk == LK_W64 || lk == LK_T64) ? 1u : 0u) | (lk == LK_T64 ? 2u : 0u);
}
__device__ __forceinline__ uint4 load_raw16(const char* p) { return __ldg((const uint4*)p); }
template <class T, int E>
__device__ __forceinline__ void convert_raw16(const uint4 w, unsigned kf, T (&v)[E]) {
  static_assert(E == 2 || E == 4 || E == 8 || E == 16, "a 16-byte chunk holds 2, 4, 8 or 16 elements");
  const unsigned x[4] = {w.x, w.y, w.z, w.w};
  if constexpr (std::is_same<T, float>::value) {
    static_assert(E == 4, "float lanes read f32 slots");
#pragma unroll
    for (int i = 0; i < 4; i++) v[i] = __uint_as_float(x[i]);
  } else if constexpr (std::is_same<T, double>::value) {
    static_assert(E == 2, "double lanes read f64 slots");
    v[0] = __hiloint2double((int)x[1], (int)x[0]); v[1] = __hiloint2double((int)x[3], (int)x[2]);
  } else {
    const bool sg = (kf & 1u) != 0; const int tg = (int)((kf >> 1) & 1u);
    if constexpr (E == 2) {
#pragma unroll
      for (int i = 0; i < 2; i++) v[i] = (T)((long long)(((unsigned long long)x[2 * i + 1] << 32) | x[2 * i]) >> tg);
    } else if constexpr (E == 4) {
#pragma unroll
      for (int i = 0; i < 4; i++) v[i] = sg ? (T)(long long)(int)x[i] : (T)(long long)x[i];
    } else if constexpr (E == 8) {
#pragma unroll
      for (int i = 0; i < 8; i++) { const unsigned h = (x[i >> 1] >> ((i & 1) * 16)) & 0xffffu; v[i] = sg ? (T)(long long)(short)h : (T)(long long)h; }
    } else {
#pragma unroll
      for (int i = 0; i < 16; i++) { const unsigned b = (x[i >> 2] >> ((i & 3) * 8)) & 0xffu; v[i] = sg ? (T)(long long)(signed char)b : (T)(long long)b; }
    }
  }
}

// programmatic dependent launch (ncl_internal.h: PdlTracker): no-ops unless the launch carries the attribute
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <class TO, int E>
__device__ __forceinline__ void store_elems(char* base, long long idx, const StoreSpec& sk, const TO (&v)[E]) {
  if constexpr (E == 1) { store_scalar<TO>(base, idx, sk, v[0]); return; }
  else if constexpr (sizeof(TO) == 4 && !std::is_integral<TO>::value) {               // float -> f32 slots
    Chunk<E * 4> c;
#pragma unroll
    for (int i = 0; i < E; i++) c.w[i] = __float_as_uint(v[i]);
    st_chunk(c, base + idx * 4);
  } else if constexpr (!std::is_integral<TO>::value) {                                  // double -> f64 slots
    Chunk<E * 8> c;
#pragma unroll
    for (int i = 0; i < E; i++) chunk_put<64>(c, i, (unsigned long long)__double_as_longlong(v[i]));
    st_chunk(c, base + idx * 8);
  } else {                                                                              // integers: %coerce wrap, then slot
    if constexpr (E == 32) {                                                            // packed bits: one word
      unsigned int w = 0;
#pragma unroll
      for (int i = 0; i < 32; i++) w |= ((unsigned int)v[i] & 1u) << i;
      *(unsigned int*)(base + (idx >> 3)) = w;
    } else if ((E == 16 && sk.kind == SK_P2) || (E == 8 && sk.kind == SK_P4)) {         // ub2 / ub4: E elements = one 32-bit word
      constexpr int B = 32 / E;
      unsigned int w = 0;
#pragma unroll
      for (int i = 0; i < E; i++) w |= ((unsigned int)v[i] & ((1u << B) - 1u)) << (i * B);
      *(unsigned int*)(base + ((idx * B) >> 3)) = w;
    } else {
      if constexpr (E == 16) { Chunk<16> c;
#pragma unroll
        for (int i = 0; i < E; i++) chunk_put<8>(c, i, (((unsigned long long)v[i]) & sk.mask) << sk.shift);
        st_chunk(c, base + idx); }
      else if constexpr (E == 8) { Chunk<16> c;
#pragma unroll
        for (int i = 0; i < E; i++) chunk_put<16>(c, i, (((unsigned long long)v[i]) & sk.mask) << sk.shift);
        st_chunk(c, base + idx * 2); }
      else if constexpr (E == 4) { Chunk<16> c;
#pragma unroll
        for (int i = 0; i < E; i++) chunk_put<32>(c, i, (((unsigned long long)v[i]) & sk.mask) << sk.shift);
        st_chunk(c, base + idx * 4); }
      else { Chunk<16> c;
#pragma unroll
        for (int i = 0; i < E; i++) chunk_put<64>(c, i, (((unsigned long long)v[i]) & sk.mask) << sk.shift);
        st_chunk(c, base + idx * 8); }
    }
  }
}


// counter-based generator shared bit for bit with oracle/numcl_oracle.c (nco_fill_random)
__host__ __device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// ---- peer mailboxes of the sharded full reduce (ncl_comm.cu maps every rank's mailbox into every process) --------------
// One mailbox = [slot 4][source rank W][2 words] + [push_seq][fin_seq].  A message of sequence number n lives in slot n & 3; each of its
// two words is {n : 32 | half of the 64-bit value : 32}, so a single 8-byte store publishes it.  Slot n & 3 is overwritten
// by message n + 4 only, and the library orders a rank's push n after its own finalize n - 2: a rank that pushes n + 4 has
// finished its finalize n + 2, which needed every peer's push n + 2, which that peer issued after ITS finalize n -- so
// every reader of n is done before n + 4 arrives, while two exchanges can be in flight per rank.
__host__ __device__ inline size_t mbox_words(int world) { return (size_t)4 * world * 2 + 8; }
__device__ __forceinline__ unsigned long long mbox_push_seq(const unsigned long long* mb, int world) { return ((const volatile unsigned long long*)mb)[(size_t)8 * world]; }
__device__ __forceinline__ void mbox_set_push_seq(unsigned long long* mb, int world, unsigned long long seq) { ((volatile unsigned long long*)mb)[(size_t)8 * world] = seq; }
__device__ __forceinline__ void mbox_send(unsigned long long* peer_mb, int world, unsigned long long seq, int from, unsigned long long bits) {
  volatile unsigned long long* slot = peer_mb + (((size_t)(seq & 3ull) * world) + from) * 2;
  const unsigned long long tag = (seq & 0xffffffffull) << 32;
  slot[0] = tag | (bits & 0xffffffffull);
  slot[1] = tag | (bits >> 32);
}
// spins until message `seq` from rank `from` has arrived in my mailbox; false after ~10 s (the caller traps)
__device__ __forceinline__ bool mbox_recv(const unsigned long long* my_mb, int world, unsigned long long seq, int from, unsigned long long* bits) {
  const unsigned long long* slot = my_mb + (((size_t)(seq & 3ull) * world) + from) * 2;
  const unsigned long long tag = seq & 0xffffffffull;
  const long long t0 = clock64();
  unsigned long long lo, hi;
  for (;;) {
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(lo) : "l"(slot) : "memory");
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(hi) : "l"(slot + 1) : "memory");
    if ((lo >> 32) == tag && (hi >> 32) == tag) break;
    if (clock64() - t0 > 20000000000LL) return false;
  }
  *bits = (hi << 32) | (lo & 0xffffffffull);
  return true;
}

// warp / block reductions
template <class T, class F> __device__ __forceinline__ T warp_reduce(T v, F f) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = f(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
