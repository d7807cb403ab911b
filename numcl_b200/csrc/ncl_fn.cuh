// ncl_fn.cuh -- the scalar functions of the whitelist beyond + - * / max min: exact mixed comparisons, the rounding family,
// expt, isqrt / logcount / integer-length, the inverse and hyperbolic functions, and the DOMAIN flag.
//
// Reference: the functions numcl hands to `broadcast` and to the unary mappers (src/4numeric.lisp:466-519,533,568-602) are
// plain Common Lisp functions; their definitions are SBCL's (src/code/numbers.lisp, src/code/irrat.lisp -- not part of the
// numcl tree, restated here from their published source):
//   (truncate x y), floats : q = (%unary-truncate (/ x y)), remainder (- x (* q y)), each operation rounded on its own;
//   (floor x y)            : truncate, then q - 1 / rem + y when rem /= 0 and (if (minusp y) (plusp x) (minusp x));
//   (ceiling x y)          : truncate, then q + 1 / rem - y when rem /= 0 and (if (minusp y) (minusp x) (plusp x));
//   (round x y)            : truncate, then away by one when |rem| > |y|/2, or = |y|/2 and q odd   ((round x 1) = half-even);
//   (mod x y) / (rem x y)  : the remainders of floor / truncate;
//   (expt b p), integer p  : intexp -- total = (if (oddp p) b 1), then repeatedly b <- b*b, total <- b*total on odd bits;
//                            negative p = (/ (intexp b (- p))); (expt b 0) = 1 of b's type;
//   comparisons of a rational with a float are EXACT (CLHS 12.1.4.1: the float is converted with `rational`).
// Where Common Lisp signals an error or leaves the reals (sqrt of a negative is a complex number) the kernels OR a code into
// a sticky device word instead; the host reports it as NCL_ERR_DOMAIN at the next synchronisation point.
#pragma once
#include "ncl_device.cuh"

enum { FAULT_DOMAIN = 1, FAULT_DIV0 = 2, FAULT_OVERFLOW = 4 };
// `f` points to a THREAD-LOCAL accumulator; the kernel flushes it once, warp-aggregated, when it is done (an atomic per faulting
// element on one global word serialised the whole kernel: sqrt over data that is half negative ran at 0.05 of the HBM peak)
__device__ __forceinline__ void ncl_fault(unsigned int* f, unsigned int code) { *f |= code; }
__device__ __forceinline__ void ncl_fault_flush(unsigned int* global_flag, unsigned int local) {
  const unsigned int am = __activemask();
  const unsigned int any = __reduce_or_sync(am, local);
  if (any && global_flag && (threadIdx.x & 31) == (unsigned)(__ffs(am) - 1)) atomicOr(global_flag, any);
}

// sign of (a - b) for an integer a and a float b, compared exactly: -1, 0, 1; 2 when b is a NaN
__device__ __forceinline__ int cmp_i64_f64(long long a, double b) {
  if (b != b) return 2;
  if (b >= 9223372036854775808.0) return -1;
  if (b < -9223372036854775808.0) return 1;
  const long long t = (long long)b;                       // toward zero; |b| < 2^63 so t is exact
  if (a != t) return a < t ? -1 : 1;
  const double frac = b - (double)t;                      // exact
  return frac > 0.0 ? -1 : (frac < 0.0 ? 1 : 0);
}
__device__ __forceinline__ int cmp_result(int op, int c) {   // c: -1 / 0 / 1 / 2 (unordered)
  switch (op) {
    case NCL_OP_EQ: return c == 0; case NCL_OP_NE: return c != 0;                 // (/= x nan) is true
    case NCL_OP_LE: return c == -1 || c == 0; case NCL_OP_GE: return c == 1 || c == 0;
    case NCL_OP_LT: return c == -1; default: return c == 1;
  }
}

// ---- integers -----------------------------------------------------------------------------------------------------
// quotient of the rounding family, exact; *rem receives the matching remainder
__device__ __forceinline__ long long int_divide(int op, long long a, long long b, long long* rem, unsigned int* flt) {
  if (b == 0) { ncl_fault(flt, FAULT_DIV0); *rem = 0; return 0; }
  if (b == -1) { *rem = 0; return (long long)(0ull - (unsigned long long)a); }      // most-negative / -1 wraps like %coerce
  long long q = a / b, r = a % b;                                                    // truncate
  const bool opposite = (r < 0) != (b < 0);
  switch (op) {
    case NCL_OP_FLOOR: case NCL_OP_MOD: if (r != 0 && opposite) { q -= 1; r += b; } break;
    case NCL_OP_CEILING: if (r != 0 && !opposite) { q += 1; r -= b; } break;
    case NCL_OP_ROUND: {
      const unsigned long long ar = r < 0 ? 0ull - (unsigned long long)r : (unsigned long long)r;
      const unsigned long long ab = b < 0 ? 0ull - (unsigned long long)b : (unsigned long long)b;
      if (ar > ab - ar || (ar == ab - ar && (q & 1))) { if (opposite) { q -= 1; r += b; } else { q += 1; r -= b; } }
      break; }
    default: break;                                                                  // truncate / rem
  }
  *rem = r; return q;
}
__device__ __forceinline__ long long int_isqrt(long long n, unsigned int* flt) {
  if (n < 0) { ncl_fault(flt, FAULT_DOMAIN); return 0; }
  unsigned long long u = (unsigned long long)n, r = (unsigned long long)__dsqrt_rn((double)u);
  while (r * r > u) r--;
  while ((r + 1) * (r + 1) <= u) r++;
  return (long long)r;
}
__device__ __forceinline__ long long int_logcount(long long n) { return __popcll((unsigned long long)(n < 0 ? ~n : n)); }
__device__ __forceinline__ long long int_length(long long n) { return 64 - __clzll(n < 0 ? ~n : n); }
// base^power for power >= 0, exact or FAULT_OVERFLOW
__device__ inline long long int_expt(long long base, long long power, unsigned int* flt) {
  if (base == 2) { if (power >= 63) { ncl_fault(flt, FAULT_OVERFLOW); return 0; } return 1ll << power; }
  bool over = false;
  auto mul = [&](long long x, long long y) { const __int128 p = (__int128)x * (__int128)y; if (p != (__int128)(long long)p) over = true; return (long long)p; };
  long long total = (power & 1) ? base : 1;
  for (long long n = power >> 1; n != 0; n >>= 1) {
    base = mul(base, base);
    if (n & 1) total = mul(base, total);
    if (over) { ncl_fault(flt, FAULT_OVERFLOW); return 0; }
  }
  return total;
}

// ---- floats ---------------------------------------------------------------------------------------------------------
template <class T> __device__ __forceinline__ long long float_trunc_i64(T q) { return f64_integer_to_i64_wrap(trunc((double)q)); }

// quotient and remainder of the rounding family on floats, SBCL's definitions
template <class T>
__device__ __forceinline__ long long float_divide(int op, T x, T y, T* rem_out) {
  const T q = Ops<T>::div(x, y);
  long long tru = float_trunc_i64<T>(q);
  T rem = Ops<T>::sub(x, Ops<T>::mul(from_i64<T>(tru), y));
  const T zero = Ops<T>::zero();
  switch (op) {
    case NCL_OP_FLOOR: case NCL_OP_MOD:
      if (rem != zero && (y < zero ? x > zero : x < zero)) { tru -= 1; rem = Ops<T>::add(rem, y); } break;
    case NCL_OP_CEILING:
      if (rem != zero && (y < zero ? x < zero : x > zero)) { tru += 1; rem = Ops<T>::sub(rem, y); } break;
    case NCL_OP_ROUND:
      if (rem != zero) {
        const T thresh = Ops<T>::div(y < zero ? -y : y, (T)2);
        if (rem > thresh || (rem == thresh && (tru & 1))) { if (y < zero) { tru -= 1; rem = Ops<T>::add(rem, y); } else { tru += 1; rem = Ops<T>::sub(rem, y); } }
        else if (rem < -thresh || (rem == -thresh && (tru & 1))) { if (y < zero) { tru += 1; rem = Ops<T>::sub(rem, y); } else { tru -= 1; rem = Ops<T>::add(rem, y); } }
      }
      break;
    default: break;
  }
  *rem_out = rem; return tru;
}
// (round x 1) = (round x): to nearest, ties to even, of any magnitude
template <class T> __device__ __forceinline__ long long float_round_unary(T x) { return round_to_i64(x); }

template <class T> __device__ inline T float_intexp(T base, long long power) {
  if (power == 0) return (T)1;
  const bool neg = power < 0;
  unsigned long long p = neg ? 0ull - (unsigned long long)power : (unsigned long long)power;
  T total = (p & 1) ? base : (T)1;
  for (unsigned long long n = p >> 1; n != 0; n >>= 1) {
    base = Ops<T>::mul(base, base);
    if (n & 1) total = Ops<T>::mul(base, total);
  }
  return neg ? Ops<T>::div((T)1, total) : total;
}

// unary functions that leave the integers / the reals; op is an ncl_op.  libm vs SBCL's libm: tolerance only (except sqrt).
template <class T> __device__ __forceinline__ T float_unary(int op, T x, unsigned int* flt);
template <> __device__ __forceinline__ float float_unary<float>(int op, float x, unsigned int* flt) {
  switch (op) {
    case NCL_OP_SQRT: if (x < 0.0f) ncl_fault(flt, FAULT_DOMAIN); return __fsqrt_rn(x);
    case NCL_OP_LOG: if (x < 0.0f) ncl_fault(flt, FAULT_DOMAIN); return logf(x);
    case NCL_OP_LOG2: if (x < 0.0f) ncl_fault(flt, FAULT_DOMAIN); return __fdiv_rn(logf(x), 0.6931472f);     // (/ (log x) (log 2))
    case NCL_OP_EXP: return expf(x); case NCL_OP_SIN: return sinf(x); case NCL_OP_COS: return cosf(x); case NCL_OP_TAN: return tanf(x);
    case NCL_OP_TANH: return tanhf(x); case NCL_OP_SINH: return sinhf(x); case NCL_OP_COSH: return coshf(x);
    case NCL_OP_ASIN: if (fabsf(x) > 1.0f) ncl_fault(flt, FAULT_DOMAIN); return asinf(x);
    case NCL_OP_ACOS: if (fabsf(x) > 1.0f) ncl_fault(flt, FAULT_DOMAIN); return acosf(x);
    case NCL_OP_ATAN: return atanf(x); case NCL_OP_ASINH: return asinhf(x);
    case NCL_OP_ACOSH: if (x < 1.0f) ncl_fault(flt, FAULT_DOMAIN); return acoshf(x);
    default: if (fabsf(x) > 1.0f) ncl_fault(flt, FAULT_DOMAIN); return atanhf(x);
  }
}
template <> __device__ __forceinline__ double float_unary<double>(int op, double x, unsigned int* flt) {
  switch (op) {
    case NCL_OP_SQRT: if (x < 0.0) ncl_fault(flt, FAULT_DOMAIN); return __dsqrt_rn(x);
    case NCL_OP_LOG: if (x < 0.0) ncl_fault(flt, FAULT_DOMAIN); return log(x);
    case NCL_OP_LOG2: if (x < 0.0) ncl_fault(flt, FAULT_DOMAIN); return __ddiv_rn(log(x), (double)0.6931472f);  // (log 2) is a single-float
    case NCL_OP_EXP: return exp(x); case NCL_OP_SIN: return sin(x); case NCL_OP_COS: return cos(x); case NCL_OP_TAN: return tan(x);
    case NCL_OP_TANH: return tanh(x); case NCL_OP_SINH: return sinh(x); case NCL_OP_COSH: return cosh(x);
    case NCL_OP_ASIN: if (fabs(x) > 1.0) ncl_fault(flt, FAULT_DOMAIN); return asin(x);
    case NCL_OP_ACOS: if (fabs(x) > 1.0) ncl_fault(flt, FAULT_DOMAIN); return acos(x);
    case NCL_OP_ATAN: return atan(x); case NCL_OP_ASINH: return asinh(x);
    case NCL_OP_ACOSH: if (x < 1.0) ncl_fault(flt, FAULT_DOMAIN); return acosh(x);
    default: if (fabs(x) > 1.0) ncl_fault(flt, FAULT_DOMAIN); return atanh(x);
  }
}
__device__ __forceinline__ long long int_bitwise(int op, long long a, long long b) {
  switch (op) {
    case NCL_OP_LOGAND: return a & b; case NCL_OP_LOGIOR: return a | b; case NCL_OP_LOGXOR: return a ^ b;
    case NCL_OP_LOGANDC1: return ~a & b; case NCL_OP_LOGANDC2: return a & ~b; case NCL_OP_LOGEQV: return ~(a ^ b);
    case NCL_OP_LOGNAND: return ~(a & b); case NCL_OP_LOGNOR: return ~(a | b); case NCL_OP_LOGORC1: return ~a | b;
    default: return a | ~b;
  }
}
