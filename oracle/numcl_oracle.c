/*
 * numcl_oracle.c -- CPU restatement of numcl's einsum / broadcast / reduce loop generators.
 * TEST INFRASTRUCTURE ONLY (see numcl_oracle.h).  Parity pinning: the reference's own
 * known-answer vectors (t/package.lisp); "parity unpinned" beyond them (no Lisp here).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fno-tree-vectorize (oracle/Makefile), so that
 * every float op is rounded separately like SBCL's scalar SSE code and the typed loops stay the
 * scalar nests the reference runs.
 *
 * Citations are file:line in the numcl tree.
 */
#define _GNU_SOURCE
#include "numcl_oracle.h"
#include <ctype.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef __int128 i128;
typedef unsigned __int128 u128;

static __thread char g_err[512];
const char* nco_last_error(void) { return g_err; }
static int fail(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
  return -1;
}

/* ------------------------------------------------------------------------------------------
 * dtype table: SBCL x86-64 specialized array element types [SBCL, unverified here]
 * ---------------------------------------------------------------------------------------- */
typedef struct { int slot_bits, value_bits, is_signed, tagged, is_float; const char* name; } dtinfo;
static const dtinfo DT[NCO_DTYPE_COUNT] = {
  /* BIT    */ {1, 1, 0, 0, 0, "bit"},
  /* UB2    */ {2, 2, 0, 0, 0, "ub2"},
  /* UB4    */ {4, 4, 0, 0, 0, "ub4"},
  /* UB7    */ {8, 7, 0, 0, 0, "ub7"},
  /* UB8    */ {8, 8, 0, 0, 0, "ub8"},
  /* UB15   */ {16, 15, 0, 0, 0, "ub15"},
  /* UB16   */ {16, 16, 0, 0, 0, "ub16"},
  /* UB31   */ {32, 31, 0, 0, 0, "ub31"},
  /* UB32   */ {32, 32, 0, 0, 0, "ub32"},
  /* UB62   */ {64, 62, 0, 1, 0, "ub62"},
  /* UB63   */ {64, 63, 0, 0, 0, "ub63"},
  /* UB64   */ {64, 64, 0, 0, 0, "ub64"},
  /* SB8    */ {8, 8, 1, 0, 0, "sb8"},
  /* SB16   */ {16, 16, 1, 0, 0, "sb16"},
  /* SB32   */ {32, 32, 1, 0, 0, "sb32"},
  /* FIXNUM */ {64, 63, 1, 1, 0, "fixnum"},
  /* SB64   */ {64, 64, 1, 0, 0, "sb64"},
  /* F32    */ {32, 0, 1, 0, 1, "f32"},
  /* F64    */ {64, 0, 1, 0, 2, "f64"},
};

size_t nco_storage_bytes(int dtype, int64_t n) {
  /* 8 * ceil(n/8) elements: src/1instantiate.lisp:81-84 */
  int64_t padded = 8 * ((n + 7) / 8);
  if (padded == 0) padded = 8;
  return (size_t)((padded * DT[dtype].slot_bits + 7) / 8);
}

/* ------------------------------------------------------------------------------------------
 * Lisp numbers: integer (exact, 128-bit), ratio, single-float, double-float
 * ---------------------------------------------------------------------------------------- */
enum { V_INT = 0, V_RATIO, V_F32, V_F64 };
typedef struct { int k; i128 i; i128 den; float f; double d; } val;

static val v_int(i128 i) { val v; memset(&v, 0, sizeof v); v.k = V_INT; v.i = i; v.den = 1; return v; }
static val v_f32(float f) { val v; memset(&v, 0, sizeof v); v.k = V_F32; v.f = f; return v; }
static val v_f64(double d) { val v; memset(&v, 0, sizeof v); v.k = V_F64; v.d = d; return v; }
static i128 iabs128(i128 x) { return x < 0 ? -x : x; }
static i128 gcd128(i128 a, i128 b) { a = iabs128(a); b = iabs128(b); while (b) { i128 t = a % b; a = b; b = t; } return a; }
static val v_ratio(i128 n, i128 d) {           /* canonical rational, CLHS 12.1.3.2 */
  if (d < 0) { n = -n; d = -d; }
  i128 g = gcd128(n, d); if (g > 1) { n /= g; d /= g; }
  if (d == 1) return v_int(n);
  val v; memset(&v, 0, sizeof v); v.k = V_RATIO; v.i = n; v.den = d; return v;
}
static int bits128(u128 x) { int b = 0; while (x) { b++; x >>= 1; } return b; }

/* correctly rounded n/d -> float with p = 24 or 53 bits, via round-to-odd on p+2.. bits */
static double ratio_to_fp(i128 n, i128 d, int p) {
  if (n == 0) return 0.0;
  int neg = (n < 0) != (d < 0);
  u128 un = (u128)iabs128(n), ud = (u128)iabs128(d);
  int s = p + 2 + bits128(ud) - bits128(un);
  u128 N = un, D = ud;
  if (s >= 0) N <<= s; else D <<= (-s);
  u128 q = N / D, r = N % D;
  if (r) q |= 1;                                      /* sticky */
  double res;
  if (p == 24) res = ldexp((double)(float)(uint64_t)q, -s);   /* one RN to 24 bits */
  else         res = ldexp((double)(uint64_t)q, -s);          /* one RN to 53 bits */
  return neg ? -res : res;
}
static float int_to_f32(i128 i) { return (float)i; }   /* libgcc __floattisf: round-to-nearest-even */
static double int_to_f64(i128 i) { return (double)i; }

static float as_f32(val v) {
  switch (v.k) {
    case V_INT: return int_to_f32(v.i);
    case V_RATIO: return (float)ratio_to_fp(v.i, v.den, 24);
    case V_F32: return v.f;
    default: return (float)v.d;
  }
}
static double as_f64(val v) {
  switch (v.k) {
    case V_INT: return int_to_f64(v.i);
    case V_RATIO: return ratio_to_fp(v.i, v.den, 53);
    case V_F32: return (double)v.f;
    default: return v.d;
  }
}
/* float contagion for a binary op (CLHS 12.1.4): rational < single < double */
static int contagion(val a, val b) {
  if (a.k == V_F64 || b.k == V_F64) return V_F64;
  if (a.k == V_F32 || b.k == V_F32) return V_F32;
  return V_INT; /* rational */
}
/* exact comparison of two reals: -1, 0, 1 (NaN -> 2) */
static int cmp_val(val a, val b) {
  if (a.k <= V_RATIO && b.k <= V_RATIO) {
    i128 l = a.i * b.den, r = b.i * a.den;
    return l < r ? -1 : (l > r ? 1 : 0);
  }
  /* SBCL compares rational with float exactly; long double holds any 64-bit integer exactly */
  long double x = (a.k == V_INT) ? (long double)a.i : (a.k == V_RATIO) ? (long double)a.i / (long double)a.den
                  : (a.k == V_F32) ? (long double)a.f : (long double)a.d;
  long double y = (b.k == V_INT) ? (long double)b.i : (b.k == V_RATIO) ? (long double)b.i / (long double)b.den
                  : (b.k == V_F32) ? (long double)b.f : (long double)b.d;
  if (x != x || y != y) return 2;
  return x < y ? -1 : (x > y ? 1 : 0);
}

enum {
  F_ADD, F_SUB, F_MUL, F_DIV, F_MAX, F_MIN, F_EQ, F_NE, F_LE, F_GE, F_LT, F_GT,
  F_LOGAND, F_LOGIOR, F_LOGXOR,
  F_LOGANDC1, F_LOGANDC2, F_LOGEQV, F_LOGNAND, F_LOGNOR, F_LOGORC1, F_LOGORC2,
  F_FLOOR, F_CEILING, F_ROUND, F_TRUNCATE, F_MOD, F_REM, F_EXPT,
  F_ASIN, F_ACOS, F_ATAN, F_SINH, F_COSH, F_ASINH, F_ACOSH, F_ATANH, F_LOG2, F_ISQRT, F_LOGCOUNT, F_INTEGER_LENGTH,
  F_NEG, F_ABS, F_SQUARE, F_SQRT, F_LOGNOT, F_IDENTITY, F_EXP, F_LOG, F_SIN, F_COS, F_TAN, F_TANH,
  F_SIGNUM, F_ONEPLUS, F_ONEMINUS, F_COUNT
};
static const struct { const char* name; int f; int arity; } FN[] = {
  {"+", F_ADD, -1}, {"-", F_SUB, -1}, {"*", F_MUL, -1}, {"/", F_DIV, -1},
  {"MAX", F_MAX, -1}, {"MIN", F_MIN, -1},
  {"=/BIT", F_EQ, 2}, {"/=/BIT", F_NE, 2}, {"<=/BIT", F_LE, 2}, {">=/BIT", F_GE, 2},
  {"</BIT", F_LT, 2}, {">/BIT", F_GT, 2},
  {"LOGAND", F_LOGAND, -1}, {"LOGIOR", F_LOGIOR, -1}, {"LOGXOR", F_LOGXOR, -1},
  {"ABS", F_ABS, 1}, {"%SQUARE", F_SQUARE, 1}, {"SQUARE", F_SQUARE, 1}, {"SQRT", F_SQRT, 1},
  {"LOGNOT", F_LOGNOT, 1}, {"IDENTITY", F_IDENTITY, 1}, {"EXP", F_EXP, 1}, {"LOG", F_LOG, 1},
  {"%LOGR", F_LOG, 1}, {"SIN", F_SIN, 1}, {"COS", F_COS, 1}, {"TAN", F_TAN, 1}, {"TANH", F_TANH, 1},
  {"SIGNUM", F_SIGNUM, 1}, {"1+", F_ONEPLUS, 1}, {"1-", F_ONEMINUS, 1},
  /* src/4numeric.lisp:592-599 */
  {"LOGANDC1", F_LOGANDC1, 2}, {"LOGANDC2", F_LOGANDC2, 2}, {"LOGEQV", F_LOGEQV, 2}, {"LOGNAND", F_LOGNAND, 2},
  {"LOGNOR", F_LOGNOR, 2}, {"LOGORC1", F_LOGORC1, 2}, {"LOGORC2", F_LOGORC2, 2},
  /* :568-581 (first value: the quotient; mod / rem: the remainder) and :533 */
  {"FLOOR", F_FLOOR, 2}, {"CEILING", F_CEILING, 2}, {"ROUND", F_ROUND, 2}, {"TRUNCATE", F_TRUNCATE, 2},
  {"MOD", F_MOD, 2}, {"REM", F_REM, 2}, {"EXPT", F_EXPT, 2},
  /* :466-519 */
  {"ASIN", F_ASIN, 1}, {"ACOS", F_ACOS, 1}, {"ATAN", F_ATAN, 1}, {"SINH", F_SINH, 1}, {"COSH", F_COSH, 1},
  {"ASINH", F_ASINH, 1}, {"ACOSH", F_ACOSH, 1}, {"ATANH", F_ATANH, 1}, {"%LOG2", F_LOG2, 1}, {"LOG2", F_LOG2, 1},
  {"ISQRT", F_ISQRT, 1}, {"LOGCOUNT", F_LOGCOUNT, 1}, {"INTEGER-LENGTH", F_INTEGER_LENGTH, 1},
};
static int fn_lookup(const char* name) {
  for (size_t i = 0; i < sizeof FN / sizeof FN[0]; i++) if (!strcmp(FN[i].name, name)) return FN[i].f;
  return -1;
}

static int g_trap;  /* set when SBCL would have signalled (division-by-zero etc.) */

/* ---- the rounding family (floor ceiling round truncate mod rem), CLHS 12.2 / SBCL src/code/numbers.lisp -------------
 * rationals: exact.  floats: SBCL's definitions -- (truncate x y) = (%unary-truncate (/ x y)) with remainder
 * (- x (* q y)), each op rounded on its own; floor / ceiling / round adjust the truncated quotient by the sign of the
 * remainder; (round x 1) is the one-argument round (ties to even). */
static i128 floor_div128(i128 n, i128 d) { i128 q = n / d, r = n % d; if (r != 0 && ((r < 0) != (d < 0))) q -= 1; return q; }
static val rational_divide(int f, val a, val b, int want_rem) {
  if (b.i == 0) { g_trap = 1; return v_int(0); }
  i128 n = a.i * b.den, d = a.den * b.i;                    /* a / b = n / d */
  if (d < 0) { n = -n; d = -d; }
  i128 q;
  switch (f) {
    case F_FLOOR: case F_MOD: q = floor_div128(n, d); break;
    case F_CEILING: q = -floor_div128(-n, d); break;
    case F_TRUNCATE: case F_REM: q = n / d; break;
    default: {                                              /* round: nearest, ties to even */
      q = floor_div128(n, d); i128 r = n - q * d, twice = 2 * r;
      if (twice > d || (twice == d && (q & 1))) q += 1;
    }
  }
  if (!want_rem) return v_int(q);
  return v_ratio(a.i * b.den - q * b.i * a.den, a.den * b.den);   /* a - q*b */
}
#define FLOAT_DIVIDE(T, NAME, TRUNCF, FABSF)                                                              \
  static i128 NAME(int f, T x, T y, T* rem_out) {                                                         \
    T q = x / y; T tq = TRUNCF(q);                                                                        \
    if (!isfinite((double)tq)) { g_trap = 1; *rem_out = 0; return 0; }                                    \
    i128 tru = (i128)tq; T prod = (T)tru * y; T rem = x - prod;                                           \
    switch (f) {                                                                                          \
      case F_FLOOR: case F_MOD: if (rem != 0 && (y < 0 ? x > 0 : x < 0)) { tru -= 1; rem = rem + y; } break; \
      case F_CEILING: if (rem != 0 && (y < 0 ? x < 0 : x > 0)) { tru += 1; rem = rem - y; } break;        \
      case F_ROUND:                                                                                       \
        if (rem != 0) { T thresh = FABSF(y) / (T)2;                                                       \
          if (rem > thresh || (rem == thresh && (tru & 1))) { if (y < 0) { tru -= 1; rem = rem + y; } else { tru += 1; rem = rem - y; } } \
          else if (rem < -thresh || (rem == -thresh && (tru & 1))) { if (y < 0) { tru += 1; rem = rem - y; } else { tru -= 1; rem = rem + y; } } } \
        break;                                                                                            \
      default: break; }                                                                                   \
    *rem_out = rem; return tru; }
FLOAT_DIVIDE(float, float_divide_f32, truncf, fabsf)
FLOAT_DIVIDE(double, float_divide_f64, trunc, fabs)

/* SBCL's intexp (src/code/irrat.lisp): repeated squaring; every product is a separate CL multiplication */
static val apply2(int f, val a, val b);
static i128 cl_round(val v);
static val intexp(val base, i128 power) {
  val total = (power & 1) ? base : v_int(1);
  for (i128 n = power >> 1; n != 0; n >>= 1) {
    base = apply2(F_MUL, base, base);
    if (n & 1) total = apply2(F_MUL, base, total);
  }
  return total;
}

/* two-argument CL arithmetic on SBCL x86-64: each float op separately rounded */
static val apply2(int f, val a, val b) {
  int c = contagion(a, b);
  switch (f) {
    case F_LOGANDC1: return v_int(~a.i & b.i);
    case F_LOGANDC2: return v_int(a.i & ~b.i);
    case F_LOGEQV: return v_int(~(a.i ^ b.i));
    case F_LOGNAND: return v_int(~(a.i & b.i));
    case F_LOGNOR: return v_int(~(a.i | b.i));
    case F_LOGORC1: return v_int(~a.i | b.i);
    case F_LOGORC2: return v_int(a.i | ~b.i);
    case F_FLOOR: case F_CEILING: case F_ROUND: case F_TRUNCATE: case F_MOD: case F_REM: {
      const int want_rem = f == F_MOD || f == F_REM;
      const int df = f == F_REM ? F_TRUNCATE : f;
      if (c == V_INT) return rational_divide(df, a, b, want_rem);
      if (f == F_ROUND && b.k == V_INT && b.i == 1) return v_int(cl_round(a));          /* (round x 1) = (round x) */
      if (c == V_F32) { float r; i128 q = float_divide_f32(df, as_f32(a), as_f32(b), &r); return want_rem ? v_f32(r) : v_int(q); }
      double r; i128 q = float_divide_f64(df, as_f64(a), as_f64(b), &r); return want_rem ? v_f64(r) : v_int(q);
    }
    case F_EXPT:
      if (b.k == V_INT) {                                                                 /* integer power: intexp */
        if (b.i == 0) return a.k == V_F32 ? v_f32(1.0f) : a.k == V_F64 ? v_f64(1.0) : v_int(1);
        if (b.i > 0) { if (a.k <= V_RATIO && bits128((u128)iabs128(a.i)) * b.i > 120) { g_trap = 1; return v_int(0); } return intexp(a, b.i); }
        if (a.k <= V_RATIO && a.i == 0) { g_trap = 1; return v_int(0); }
        if (a.k <= V_RATIO && bits128((u128)iabs128(a.i)) * (-b.i) > 120) { g_trap = 1; return v_int(0); }
        return apply2(F_DIV, v_int(1), intexp(a, -b.i));
      } else {                                                                            /* real-expt: pow in double-float */
        double x = as_f64(a), y = as_f64(b);
        if (x < 0 && y != rint(y)) { g_trap = 1; return v_int(0); }                       /* complex result */
        double r = pow(x, y);
        return (a.k == V_F64 || b.k == V_F64) ? v_f64(r) : v_f32((float)r);
      }
    case F_ADD: case F_SUB: case F_MUL:
      if (c == V_INT) {
        i128 n, d = a.den * b.den;
        if (f == F_ADD) n = a.i * b.den + b.i * a.den;
        else if (f == F_SUB) n = a.i * b.den - b.i * a.den;
        else n = a.i * b.i;
        return v_ratio(n, d);
      } else if (c == V_F32) {
        float x = as_f32(a), y = as_f32(b);
        return v_f32(f == F_ADD ? x + y : f == F_SUB ? x - y : x * y);
      } else {
        double x = as_f64(a), y = as_f64(b);
        return v_f64(f == F_ADD ? x + y : f == F_SUB ? x - y : x * y);
      }
    case F_DIV:
      if (c == V_INT) {
        if (b.i == 0) { g_trap = 1; return v_int(0); }       /* division-by-zero error in CL */
        return v_ratio(a.i * b.den, a.den * b.i);
      } else if (c == V_F32) return v_f32(as_f32(a) / as_f32(b));
      else return v_f64(as_f64(a) / as_f64(b));
    case F_MAX: { int r = cmp_val(a, b); return (r == 2) ? b : (r >= 0 ? a : b); }  /* (if (>= a b) a b) */
    case F_MIN: { int r = cmp_val(a, b); return (r == 2) ? b : (r <= 0 ? a : b); }  /* (if (<= a b) a b) */
    case F_EQ: return v_int(cmp_val(a, b) == 0);
    case F_NE: { int r = cmp_val(a, b); return v_int(r != 0); }
    case F_LE: { int r = cmp_val(a, b); return v_int(r == -1 || r == 0); }
    case F_GE: { int r = cmp_val(a, b); return v_int(r == 1 || r == 0); }
    case F_LT: return v_int(cmp_val(a, b) == -1);
    case F_GT: return v_int(cmp_val(a, b) == 1);
    case F_LOGAND: return v_int(a.i & b.i);
    case F_LOGIOR: return v_int(a.i | b.i);
    case F_LOGXOR: return v_int(a.i ^ b.i);
  }
  g_trap = 1; return v_int(0);
}
static val apply1(int f, val a) {
  switch (f) {
    case F_NEG: return a.k <= V_RATIO ? v_ratio(-a.i, a.den) : (a.k == V_F32 ? v_f32(-a.f) : v_f64(-a.d));
    case F_ABS: return a.k <= V_RATIO ? v_ratio(iabs128(a.i), a.den) : a.k == V_F32 ? v_f32(fabsf(a.f)) : v_f64(fabs(a.d));
    case F_SQUARE: return apply2(F_MUL, a, a);
    case F_IDENTITY: return a;
    case F_LOGNOT: return v_int(~a.i);
    case F_ONEPLUS: return apply2(F_ADD, a, v_int(1));
    case F_ONEMINUS: return apply2(F_SUB, a, v_int(1));
    case F_SIGNUM:
      if (a.k <= V_RATIO) return v_int(a.i > 0 ? 1 : a.i < 0 ? -1 : 0);
      if (a.k == V_F32) return v_f32(a.f > 0 ? 1.0f : a.f < 0 ? -1.0f : a.f);
      return v_f64(a.d > 0 ? 1.0 : a.d < 0 ? -1.0 : a.d);
    case F_ISQRT: {
      if (a.k != V_INT || a.i < 0) { g_trap = 1; return v_int(0); }
      u128 u = (u128)a.i, r = (u128)sqrtl((long double)a.i);
      while (r * r > u) r--;
      while ((r + 1) * (r + 1) <= u) r++;
      return v_int((i128)r);
    }
    case F_LOGCOUNT: { if (a.k != V_INT) { g_trap = 1; return v_int(0); } u128 u = (u128)(a.i < 0 ? ~a.i : a.i); int c = 0; while (u) { c += (int)(u & 1); u >>= 1; } return v_int(c); }
    case F_INTEGER_LENGTH: { if (a.k != V_INT) { g_trap = 1; return v_int(0); } return v_int(bits128((u128)(a.i < 0 ? ~a.i : a.i))); }
    case F_ASIN: case F_ACOS: case F_ATAN: case F_SINH: case F_COSH: case F_ASINH: case F_ACOSH: case F_ATANH: case F_LOG2: {
      /* outside the real domain CL returns a complex number: trapped */
      double xd = as_f64(a);
      if (((f == F_ASIN || f == F_ACOS || f == F_ATANH) && fabs(xd) > 1.0) || (f == F_ACOSH && xd < 1.0) || (f == F_LOG2 && xd < 0.0)) { g_trap = 1; return v_int(0); }
      if (a.k == V_F64) {
        double x = a.d;
        return v_f64(f == F_ASIN ? asin(x) : f == F_ACOS ? acos(x) : f == F_ATAN ? atan(x) : f == F_SINH ? sinh(x) : f == F_COSH ? cosh(x)
                     : f == F_ASINH ? asinh(x) : f == F_ACOSH ? acosh(x) : f == F_ATANH ? atanh(x) : log(x) / (double)logf(2.0f));   /* (/ (log x) (log 2)) */
      }
      float x = as_f32(a);
      return v_f32(f == F_ASIN ? asinf(x) : f == F_ACOS ? acosf(x) : f == F_ATAN ? atanf(x) : f == F_SINH ? sinhf(x) : f == F_COSH ? coshf(x)
                   : f == F_ASINH ? asinhf(x) : f == F_ACOSH ? acoshf(x) : f == F_ATANH ? atanhf(x) : logf(x) / logf(2.0f));
    }
    /* irrational functions: rational -> single-float (float substitution, 1type.lisp:358-416) */
    case F_SQRT: case F_EXP: case F_LOG: case F_SIN: case F_COS: case F_TAN: case F_TANH: {
      if ((f == F_SQRT || f == F_LOG) && as_f64(a) < 0.0) { g_trap = 1; return v_int(0); }      /* complex result */
      if (a.k == V_F64) {
        double x = a.d;
        return v_f64(f == F_SQRT ? sqrt(x) : f == F_EXP ? exp(x) : f == F_LOG ? log(x) : f == F_SIN ? sin(x)
                     : f == F_COS ? cos(x) : f == F_TAN ? tan(x) : tanh(x));
      }
      float x = as_f32(a);
      return v_f32(f == F_SQRT ? sqrtf(x) : f == F_EXP ? expf(x) : f == F_LOG ? logf(x) : f == F_SIN ? sinf(x)
                   : f == F_COS ? cosf(x) : f == F_TAN ? tanf(x) : tanhf(x));
    }
  }
  g_trap = 1; return v_int(0);
}

/* (round x): to nearest, ties to even; on integers identity */
static i128 cl_round(val v) {
  switch (v.k) {
    case V_INT: return v.i;
    case V_RATIO: {
      i128 q = v.i / v.den, r = v.i % v.den;          /* trunc */
      if (r < 0) { q -= 1; r += v.den; }               /* floor, 0 <= r < den */
      i128 twice = 2 * r;
      if (twice > v.den || (twice == v.den && (q & 1))) q += 1;
      return q;
    }
    case V_F32: { double d = (double)v.f; if (!isfinite(d)) { g_trap = 1; return 0; } return (i128)nearbyint(d); }
    default: { if (!isfinite(v.d)) { g_trap = 1; return 0; }
      double r = nearbyint(v.d);                       /* FE_TONEAREST: ties to even */
      if (fabs(r) < 1.7e38) return (i128)r;
      g_trap = 1; return 0; }
  }
}

/* ------------------------------------------------------------------------------------------
 * %coerce + element access (src/1type.lisp:642-694): the first width w in 1..64 with
 * type <= (unsigned-byte w) -> (logand (round x) (2^w - 1)); type <= (signed-byte w) -> sb w
 * ---------------------------------------------------------------------------------------- */
static i128 wrap_int(i128 x, int dtype) {
  int w = DT[dtype].value_bits;
  u128 mask = (w == 128) ? ~(u128)0 : (((u128)1 << w) - 1);
  if (!DT[dtype].is_signed) return (i128)((u128)x & mask);                         /* ub, :644-647 */
  i128 half = (i128)1 << (w - 1), mod = (i128)1 << w;
  i128 m = (x + half) % mod; if (m < 0) m += mod;                                  /* sb, :648-653 */
  return m - half;
}
static val coerce_to(val v, int dtype) {
  if (DT[dtype].is_float == 1) return v_f32(as_f32(v));                            /* (coerce x 'single-float) */
  if (DT[dtype].is_float == 2) return v_f64(as_f64(v));
  return v_int(wrap_int(cl_round(v), dtype));
}

static val load_el(const nco_tensor* t, int64_t idx) {
  const dtinfo* d = &DT[t->dtype];
  const uint8_t* p = (const uint8_t*)t->data;
  if (d->is_float == 1) return v_f32(((const float*)p)[idx]);
  if (d->is_float == 2) return v_f64(((const double*)p)[idx]);
  if (d->slot_bits < 8) {
    int per = 8 / d->slot_bits; int64_t byte = idx / per; int sh = (int)(idx % per) * d->slot_bits;
    return v_int((p[byte] >> sh) & ((1 << d->slot_bits) - 1));
  }
  i128 raw;
  switch (d->slot_bits) {
    case 8:  raw = d->is_signed ? (i128)((const int8_t*)p)[idx]  : (i128)((const uint8_t*)p)[idx]; break;
    case 16: raw = d->is_signed ? (i128)((const int16_t*)p)[idx] : (i128)((const uint16_t*)p)[idx]; break;
    case 32: raw = d->is_signed ? (i128)((const int32_t*)p)[idx] : (i128)((const uint32_t*)p)[idx]; break;
    default:
      if (d->tagged) raw = (i128)(((const int64_t*)p)[idx] >> 1);                 /* fixnum tag bit */
      else raw = d->is_signed ? (i128)((const int64_t*)p)[idx] : (i128)((const uint64_t*)p)[idx];
  }
  return v_int(raw);
}
static void store_el(nco_tensor* t, int64_t idx, val v) {
  const dtinfo* d = &DT[t->dtype];
  uint8_t* p = (uint8_t*)t->data;
  val c = coerce_to(v, t->dtype);
  if (d->is_float == 1) { ((float*)p)[idx] = c.f; return; }
  if (d->is_float == 2) { ((double*)p)[idx] = c.d; return; }
  if (d->slot_bits < 8) {
    int per = 8 / d->slot_bits; int64_t byte = idx / per; int sh = (int)(idx % per) * d->slot_bits;
    uint8_t m = (uint8_t)(((1 << d->slot_bits) - 1) << sh);
    p[byte] = (uint8_t)((p[byte] & ~m) | (((uint8_t)c.i << sh) & m));
    return;
  }
  switch (d->slot_bits) {
    case 8:  ((uint8_t*)p)[idx] = (uint8_t)c.i; break;
    case 16: ((uint16_t*)p)[idx] = (uint16_t)c.i; break;
    case 32: ((uint32_t*)p)[idx] = (uint32_t)c.i; break;
    default: ((uint64_t*)p)[idx] = d->tagged ? ((uint64_t)c.i << 1) : (uint64_t)c.i;
  }
}

int nco_coerce_f64(double v, int dtype, void* slot) {
  nco_tensor t; memset(&t, 0, sizeof t); t.data = slot; t.dtype = dtype; memset(slot, 0, 8);
  g_trap = 0; store_el(&t, 0, v_f64(v)); return g_trap ? fail("trap in %%coerce") : 0;
}
int nco_coerce_i64(int64_t v, int dtype, void* slot) {
  nco_tensor t; memset(&t, 0, sizeof t); t.data = slot; t.dtype = dtype; memset(slot, 0, 8);
  store_el(&t, 0, v_int(v)); return 0;
}
int nco_pack_i64(const int64_t* v, int64_t n, int dtype, void* storage) {
  nco_tensor t; memset(&t, 0, sizeof t); t.data = storage; t.dtype = dtype;
  for (int64_t i = 0; i < n; i++) store_el(&t, i, DT[dtype].is_float ? v_int(v[i]) : v_int(v[i]));
  return 0;
}
int nco_unpack_i64(const void* storage, int64_t n, int dtype, int64_t* v) {
  nco_tensor t; memset(&t, 0, sizeof t); t.data = (void*)storage; t.dtype = dtype;
  for (int64_t i = 0; i < n; i++) { val x = load_el(&t, i); v[i] = (x.k == V_INT) ? (int64_t)x.i : (int64_t)as_f64(x); }
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * S-expression reader (the Lisp reader, as far as einsum subscripts need it)
 * ---------------------------------------------------------------------------------------- */
enum { SX_SYM, SX_INT, SX_F32, SX_F64, SX_LIST };
typedef struct sx { int kind; char* name; long long ival; double fval; struct sx** it; int n; int is_key; } sx;

typedef struct { sx** items; int n, cap; } arena_t;
static __thread arena_t g_arena;
static void* arena_track(void* p) {
  if (g_arena.n == g_arena.cap) { g_arena.cap = g_arena.cap ? g_arena.cap * 2 : 256;
    g_arena.items = (sx**)realloc(g_arena.items, sizeof(sx*) * g_arena.cap); }
  g_arena.items[g_arena.n++] = (sx*)p; return p;
}
static void arena_free(void) { for (int i = 0; i < g_arena.n; i++) free(g_arena.items[i]); g_arena.n = 0; }
static sx* sx_new(int kind) { sx* s = (sx*)arena_track(calloc(1, sizeof(sx))); s->kind = kind; return s; }
static char* sdup_up(const char* b, int n) {
  char* s = (char*)arena_track(malloc(n + 1));
  for (int i = 0; i < n; i++) s[i] = (char)toupper((unsigned char)b[i]);   /* readtable-case :upcase */
  s[n] = 0; return s;
}
static void sx_push(sx* l, sx* c) {
  sx** n = (sx**)arena_track(malloc(sizeof(sx*) * (l->n + 1)));
  if (l->n) memcpy(n, l->it, sizeof(sx*) * l->n);
  n[l->n] = c; l->it = n; l->n++;
}
static sx* parse_atom(const char* b, int n) {
  /* number? */
  char tmp[128]; if (n >= (int)sizeof tmp) n = sizeof tmp - 1;
  memcpy(tmp, b, n); tmp[n] = 0;
  char* end; long long iv = strtoll(tmp, &end, 10);
  if (*end == 0 && n > 0 && (isdigit((unsigned char)tmp[n - 1]))) { sx* s = sx_new(SX_INT); s->ival = iv; return s; }
  /* float: digits with '.' and/or exponent marker e/f (single) d (double) */
  int has_digit = 0, ok = 1, is_double = 0, seen_exp = 0; char conv[128]; int ci = 0;
  for (int i = 0; i < n && ok; i++) {
    char c = tmp[i];
    if (isdigit((unsigned char)c)) { has_digit = 1; conv[ci++] = c; }
    else if (c == '.' && !seen_exp) conv[ci++] = c;
    else if ((c == '+' || c == '-') && (i == 0 || strchr("eEdDfF", tmp[i - 1]))) conv[ci++] = c;
    else if (strchr("eEfFdD", c) && has_digit && !seen_exp && i + 1 < n) { seen_exp = 1; is_double = (c == 'd' || c == 'D'); conv[ci++] = 'e'; }
    else ok = 0;
  }
  conv[ci] = 0;
  if (ok && has_digit && (strchr(tmp, '.') || seen_exp)) {
    sx* s = sx_new(is_double ? SX_F64 : SX_F32);    /* *read-default-float-format* = single-float */
    s->fval = is_double ? strtod(conv, NULL) : (double)strtof(conv, NULL); return s;
  }
  sx* s = sx_new(SX_SYM);
  /* strip package prefix (cl:+ -> +); a leading colon marks a keyword */
  const char* colon = NULL; for (int i = 0; i < n; i++) if (b[i] == ':') colon = b + i;
  if (b[0] == ':') { s->is_key = 1; s->name = sdup_up(b + 1, n - 1); }
  else if (colon && colon + 1 < b + n) { s->name = sdup_up(colon + 1, (int)(b + n - colon - 1)); }
  else s->name = sdup_up(b, n);
  return s;
}
static sx* parse_list(const char** pp, int toplevel) {
  sx* l = sx_new(SX_LIST); const char* p = *pp;
  for (;;) {
    while (*p && isspace((unsigned char)*p)) p++;
    if (!*p) { if (!toplevel) { fail("unbalanced parens"); return NULL; } break; }
    if (*p == ')') { if (toplevel) { fail("unbalanced )"); return NULL; } p++; break; }
    if (*p == '(') { p++; sx* c = parse_list(&p, 0); if (!c) return NULL; sx_push(l, c); continue; }
    if (*p == '\'') { p++; continue; }
    const char* b = p; while (*p && !isspace((unsigned char)*p) && *p != '(' && *p != ')') p++;
    sx_push(l, parse_atom(b, (int)(p - b)));
  }
  *pp = p; return l;
}
static int sx_is_sym(const sx* s, const char* name) { return s->kind == SX_SYM && !s->is_key && !strcmp(s->name, name); }
static int sx_is_nil(const sx* s) { return (s->kind == SX_LIST && s->n == 0) || sx_is_sym(s, "NIL"); }

/* ------------------------------------------------------------------------------------------
 * einsum-specs (src/3einsum.lisp:44-62)
 * ---------------------------------------------------------------------------------------- */
typedef struct { int key_is_num; int key; char keyname[32]; int has_start, has_end; int64_t start, end, step; } opt_t;
typedef struct {
  int n_in, n_out;
  int n_iter, iter[NCO_MAX_IDX];
  int irank[NCO_MAX_T], ispec[NCO_MAX_T][NCO_MAX_RANK];
  int orank[NCO_MAX_T], ospec[NCO_MAX_T][NCO_MAX_RANK];
  int n_iopt[NCO_MAX_T], n_oopt[NCO_MAX_T];
  opt_t iopt[NCO_MAX_T][NCO_MAX_RANK], oopt[NCO_MAX_T][NCO_MAX_RANK];
  sx* transforms[NCO_MAX_T];
} specs_t;

/* explode, src/3einsum.lisp:124-138.  names[] receives index names */
static int explode(const sx* s, char names[][32], int* n) {
  *n = 0;
  if (s->kind == SX_LIST) {
    for (int i = 0; i < s->n; i++) {
      if (s->it[i]->kind != SX_SYM) return fail("(assert (symbolp c)) failed in explode");   /* :127 */
      snprintf(names[(*n)++], 32, "%s", s->it[i]->name);
    }
    return 0;
  }
  if (s->kind != SX_SYM) return fail("spec is neither a symbol nor a list");
  for (const char* c = s->name; *c; c++) {
    if (!(*c == '-' || isalpha((unsigned char)*c)))                                            /* :131-136 */
      return fail("type-error: Tried to make a spec from a non-alpha char %c", *c);
    names[*n][0] = *c; names[*n][1] = 0; (*n)++;
    if (*n >= NCO_MAX_RANK + 1) break;
  }
  return 0;
}

/* sort-locality, src/3einsum.lisp:523-565 */
static void sort_locality(int* indices, int n_idx, int subs[][NCO_MAX_RANK], const int* subrank, int n_subs, int* out) {
  /* hash table keyed by the list of subscripts an index occurs in -> encode as a bitmask */
  unsigned keys[64]; int vals[64]; int nk = 0;
  int remaining[NCO_MAX_IDX]; int nr = n_idx; memcpy(remaining, indices, sizeof(int) * n_idx);
  int no = 0;
  while (nr > 0) {
    int best = -1; long best_score = 0;
    for (int t = 0; t < nr; t++) {
      int index = remaining[t]; long score = 0; unsigned key = 0;
      for (int s = 0; s < n_subs; s++) {
        int pos = -1; for (int a = 0; a < subrank[s]; a++) if (subs[s][a] == index) { pos = a; break; }  /* position */
        if (pos >= 0) { score += 1 + pos; key |= 1u << s; }
      }
      int sticky = 0; for (int k = 0; k < nk; k++) if (keys[k] == key) sticky = vals[k];
      long total = score + (long)nr * sticky;                  /* (* (length indices) (sticky-score index)) */
      if (best < 0 || total < best_score) { best = t; best_score = total; }   /* finding ... minimizing: first minimum */
    }
    int next = remaining[best]; unsigned key = 0;
    for (int s = 0; s < n_subs; s++) for (int a = 0; a < subrank[s]; a++) if (subs[s][a] == next) { key |= 1u << s; break; }
    int found = 0; for (int k = 0; k < nk; k++) if (keys[k] == key) { vals[k]++; found = 1; }
    if (!found) { keys[nk] = key; vals[nk] = 1; nk++; }
    out[no++] = next;
    for (int t = best; t + 1 < nr; t++) remaining[t] = remaining[t + 1];
    nr--;
  }
}

static int parse_opts(const sx* alist, opt_t* opts, int* n_opts, char idx_names[][32], const int* idx_num, int n_names) {
  *n_opts = 0;
  if (!alist || sx_is_nil(alist)) return 0;
  if (alist->kind != SX_LIST) return fail("iteration options must be an alist");
  for (int e = 0; e < alist->n; e++) {
    const sx* ent = alist->it[e];
    if (ent->kind != SX_LIST || ent->n < 1) return fail("bad iteration option entry");
    opt_t* o = &opts[(*n_opts)++]; memset(o, 0, sizeof *o); o->step = 1;
    const sx* key = ent->it[0];
    if (key->kind == SX_INT) { o->key_is_num = 1; o->key = (int)key->ival; }
    else if (key->kind == SX_SYM) {
      /* (sublis alist i-options-full): index symbols become their numbers, :181-182 */
      snprintf(o->keyname, 32, "%s", key->name);
      for (int k = 0; k < n_names; k++) if (!strcmp(idx_names[k], key->name)) { o->key_is_num = 1; o->key = idx_num[k]; }
    }
    for (int k = 1; k + 1 < ent->n; k += 2) {
      const sx* kw = ent->it[k]; const sx* v = ent->it[k + 1];
      if (kw->kind != SX_SYM) return fail("bad iteration option keyword");
      int is_t = sx_is_sym(v, "T");
      if (!is_t && v->kind != SX_INT) return fail("iteration specifier option should be constant");  /* :311-313 */
      if (!strcmp(kw->name, "START")) { if (!is_t) { o->has_start = 1; o->start = v->ival; } }
      else if (!strcmp(kw->name, "END")) { if (!is_t) { o->has_end = 1; o->end = v->ival; } }
      else if (!strcmp(kw->name, "STEP")) o->step = v->ival;
    }
  }
  return 0;
}

/* einsum-normalize-subscripts + %einsum-normalize-subscripts, src/3einsum.lisp:64-200 */
static int normalize(const char* text, specs_t* sp) {
  memset(sp, 0, sizeof *sp);
  const char* p = text; sx* top = parse_list(&p, 1);
  if (!top) return -1;
  /* split on `->` (:66-96) */
  int arrows[8], na = 0;
  for (int i = 0; i < top->n; i++) if (sx_is_sym(top->it[i], "->")) { if (na < 8) arrows[na] = i; na++; }
  if (na > 4) return fail("ecase: %d arrows", na);
  int seg_b[5], seg_e[5];
  for (int s = 0; s <= na; s++) { seg_b[s] = s == 0 ? 0 : arrows[s - 1] + 1; seg_e[s] = s == na ? top->n : arrows[s]; }
  int i_b = seg_b[0], i_e = seg_e[0];
  int have_t = na >= 2, have_o = na >= 1, have_io = na >= 3, have_oo = na >= 4;
  int t_seg = 1, o_seg = na >= 2 ? 2 : 1, io_seg = 3, oo_seg = 4;

  /* explode input specs, collect indices in order of first appearance (:149-154) */
  char names[NCO_MAX_IDX][32]; int n_names = 0; int name_num[NCO_MAX_IDX];
  char tmp[NCO_MAX_RANK + 1][32]; int tn;
  char ispec_names[NCO_MAX_T][NCO_MAX_RANK][32], ospec_names[NCO_MAX_T][NCO_MAX_RANK][32];
  sp->n_in = i_e - i_b;
  if (sp->n_in > NCO_MAX_T) return fail("too many inputs");
  for (int a = 0; a < sp->n_in; a++) {
    if (explode(top->it[i_b + a], tmp, &tn)) return -1;
    if (tn > NCO_MAX_RANK) return fail("rank too large");
    sp->irank[a] = tn;
    for (int k = 0; k < tn; k++) {
      strcpy(ispec_names[a][k], tmp[k]);
      int seen = 0; for (int q = 0; q < n_names; q++) if (!strcmp(names[q], tmp[k])) seen = 1;
      if (!seen) { if (n_names >= NCO_MAX_IDX) return fail("too many indices"); strcpy(names[n_names++], tmp[k]); }
    }
  }
  /* make-map (:139-148): position in the index list, `-` -> -1 (the counter still advances) */
  for (int q = 0; q < n_names; q++) name_num[q] = !strcmp(names[q], "-") ? -1 : q;

  /* output specs (:156-160): given, or '(()) after a bare arrow, or all indices sorted by string< */
  if (have_o) {
    sp->n_out = seg_e[o_seg] - seg_b[o_seg];
    if (sp->n_out == 0) { sp->n_out = 1; sp->orank[0] = 0; }
    else for (int a = 0; a < sp->n_out; a++) {
      if (explode(top->it[seg_b[o_seg] + a], tmp, &tn)) return -1;
      if (tn > NCO_MAX_RANK) return fail("rank too large");
      sp->orank[a] = tn; for (int k = 0; k < tn; k++) strcpy(ospec_names[a][k], tmp[k]);
    }
  } else {
    sp->n_out = 1; sp->orank[0] = n_names;
    if (n_names > NCO_MAX_RANK) return fail("rank too large");
    int ord[NCO_MAX_IDX]; for (int q = 0; q < n_names; q++) ord[q] = q;
    for (int i = 0; i < n_names; i++) for (int j = i + 1; j < n_names; j++)
      if (strcmp(names[ord[j]], names[ord[i]]) < 0) { int t = ord[i]; ord[i] = ord[j]; ord[j] = t; }
    for (int k = 0; k < n_names; k++) strcpy(ospec_names[0][k], names[ord[k]]);
  }
  /* sublis: names -> numbers */
  for (int a = 0; a < sp->n_in; a++) for (int k = 0; k < sp->irank[a]; k++)
    for (int q = 0; q < n_names; q++) if (!strcmp(names[q], ispec_names[a][k])) sp->ispec[a][k] = name_num[q];
  for (int a = 0; a < sp->n_out; a++) for (int k = 0; k < sp->orank[a]; k++) {
    int found = 0;
    for (int q = 0; q < n_names; q++) if (!strcmp(names[q], ospec_names[a][k])) { sp->ospec[a][k] = name_num[q]; found = 1; }
    if (!found) return fail("The output spec contains %s which are not used in the input specs", ospec_names[a][k]); /* :191-194 */
  }
  /* transforms (:165-178) */
  if (have_t && seg_e[t_seg] > seg_b[t_seg]) {
    int nt = seg_e[t_seg] - seg_b[t_seg];
    if (nt != sp->n_out) return fail("(assert (= o-len (length transforms))) failed");             /* :190 */
    for (int a = 0; a < nt; a++) sp->transforms[a] = top->it[seg_b[t_seg] + a];
  } else {
    for (int o = 0; o < sp->n_out; o++) {              /* `(+ @o (* $1 ... $N)) */
      sx* mul = sx_new(SX_LIST); sx* star = sx_new(SX_SYM); star->name = sdup_up("*", 1); sx_push(mul, star);
      for (int i = 1; i <= sp->n_in; i++) { char b[16]; snprintf(b, 16, "$%d", i); sx* s = sx_new(SX_SYM); s->name = sdup_up(b, (int)strlen(b)); sx_push(mul, s); }
      sx* add = sx_new(SX_LIST); sx* plus = sx_new(SX_SYM); plus->name = sdup_up("+", 1); sx_push(add, plus);
      char b[16]; snprintf(b, 16, "@%d", o + 1); sx* at = sx_new(SX_SYM); at->name = sdup_up(b, (int)strlen(b)); sx_push(add, at);
      sx_push(add, mul); sp->transforms[o] = add;
    }
  }
  /* options (:179-182): (replace full given) stops at the shorter sequence */
  if (have_io) for (int a = 0; a < sp->n_in && a < seg_e[io_seg] - seg_b[io_seg]; a++)
    if (parse_opts(top->it[seg_b[io_seg] + a], sp->iopt[a], &sp->n_iopt[a], names, name_num, n_names)) return -1;
  if (have_oo) for (int a = 0; a < sp->n_out && a < seg_e[oo_seg] - seg_b[oo_seg]; a++)
    if (parse_opts(top->it[seg_b[oo_seg] + a], sp->oopt[a], &sp->n_oopt[a], names, name_num, n_names)) return -1;

  /* i-flat / o-flat = (remove-duplicates (flatten ...)) keeps the LAST occurrence (:183-184) */
  int iflat[NCO_MAX_IDX * 2], nif = 0, oflat[NCO_MAX_IDX * 2], nof = 0;
  {
    int all[NCO_MAX_T * NCO_MAX_RANK], n = 0;
    for (int a = 0; a < sp->n_in; a++) for (int k = 0; k < sp->irank[a]; k++) all[n++] = sp->ispec[a][k];
    for (int i = 0; i < n; i++) { int later = 0; for (int j = i + 1; j < n; j++) if (all[j] == all[i]) later = 1; if (!later) iflat[nif++] = all[i]; }
    n = 0;
    for (int a = 0; a < sp->n_out; a++) for (int k = 0; k < sp->orank[a]; k++) all[n++] = sp->ospec[a][k];
    for (int i = 0; i < n; i++) { int later = 0; for (int j = i + 1; j < n; j++) if (all[j] == all[i]) later = 1; if (!later) oflat[nof++] = all[i]; }
  }
  /* (union i-flat o-flat): SBCL pushes the members of list1 missing from list2 onto list2
   * [SBCL list-based union for short lists, unverified here] */
  int uni[NCO_MAX_IDX * 2], nu = 0;
  {
    int pushed[NCO_MAX_IDX * 2], np = 0;
    for (int i = 0; i < nif; i++) { int in2 = 0; for (int j = 0; j < nof; j++) if (oflat[j] == iflat[i]) in2 = 1; if (!in2) pushed[np++] = iflat[i]; }
    for (int i = np - 1; i >= 0; i--) uni[nu++] = pushed[i];
    for (int j = 0; j < nof; j++) uni[nu++] = oflat[j];
  }
  if (nu > NCO_MAX_IDX) return fail("too many indices");
  int subs[2 * NCO_MAX_T][NCO_MAX_RANK], subrank[2 * NCO_MAX_T], ns = 0;
  for (int a = 0; a < sp->n_in; a++) { memcpy(subs[ns], sp->ispec[a], sizeof subs[ns]); subrank[ns++] = sp->irank[a]; }
  for (int a = 0; a < sp->n_out; a++) { memcpy(subs[ns], sp->ospec[a], sizeof subs[ns]); subrank[ns++] = sp->orank[a]; }
  sort_locality(uni, nu, subs, subrank, ns, sp->iter);
  sp->n_iter = nu;
  /* (assert (subsetp o-flat i-flat)) :191 -- checked above while mapping names */
  return 0;
}

/* ---- printing for the harness ------------------------------------------------------------ */
typedef struct { char* b; int cap, n; } sbuf;
static void sb_put(sbuf* s, const char* fmt, ...) {
  va_list ap; va_start(ap, fmt);
  int r = vsnprintf(s->b + s->n, s->cap > s->n ? s->cap - s->n : 0, fmt, ap); va_end(ap);
  if (r > 0) s->n += r; if (s->n >= s->cap) s->n = s->cap - 1;
}
static void sb_sx(sbuf* s, const sx* x) {
  switch (x->kind) {
    case SX_SYM: { sb_put(s, "%s", x->is_key ? ":" : ""); for (const char* c = x->name; *c; c++) sb_put(s, "%c", tolower((unsigned char)*c)); break; }
    case SX_INT: sb_put(s, "%lld", x->ival); break;
    case SX_F32: sb_put(s, "%.9g", x->fval); break;
    case SX_F64: sb_put(s, "%.17gd0", x->fval); break;
    default: sb_put(s, "("); for (int i = 0; i < x->n; i++) { if (i) sb_put(s, " "); sb_sx(s, x->it[i]); } sb_put(s, ")");
  }
}
static void sb_intlist(sbuf* s, const int* v, int n) { sb_put(s, "("); for (int i = 0; i < n; i++) sb_put(s, i ? " %d" : "%d", v[i]); sb_put(s, ")"); }
static void sb_opts(sbuf* s, const opt_t* o, int n) {
  if (!n) { sb_put(s, "nil"); return; }
  sb_put(s, "(");
  for (int i = 0; i < n; i++) {
    if (i) sb_put(s, " ");
    if (o[i].key_is_num) sb_put(s, "(%d", o[i].key); else sb_put(s, "(%s", o[i].keyname);
    if (o[i].has_start) sb_put(s, " :start %lld", (long long)o[i].start);
    if (o[i].has_end) sb_put(s, " :end %lld", (long long)o[i].end);
    if (o[i].step != 1) sb_put(s, " :step %lld", (long long)o[i].step);
    sb_put(s, ")");
  }
  sb_put(s, ")");
}
int nco_spec_counts(const char* spec, int* n_in, int* n_out) {
  specs_t sp; int r = normalize(spec, &sp);
  if (!r) { *n_in = sp.n_in; *n_out = sp.n_out; }
  arena_free(); return r;
}
int nco_normalize(const char* spec, char* out, int cap) {
  specs_t sp; int r = normalize(spec, &sp);
  if (r) { arena_free(); return r; }
  sbuf s = {out, cap, 0};
  sb_put(&s, "iter="); sb_intlist(&s, sp.iter, sp.n_iter);
  sb_put(&s, " i=("); for (int a = 0; a < sp.n_in; a++) { if (a) sb_put(&s, " "); sb_intlist(&s, sp.ispec[a], sp.irank[a]); } sb_put(&s, ")");
  sb_put(&s, " o=("); for (int a = 0; a < sp.n_out; a++) { if (a) sb_put(&s, " "); sb_intlist(&s, sp.ospec[a], sp.orank[a]); } sb_put(&s, ")");
  sb_put(&s, " t=("); for (int a = 0; a < sp.n_out; a++) { if (a) sb_put(&s, " "); sb_sx(&s, sp.transforms[a]); } sb_put(&s, ")");
  sb_put(&s, " io=("); for (int a = 0; a < sp.n_in; a++) { if (a) sb_put(&s, " "); sb_opts(&s, sp.iopt[a], sp.n_iopt[a]); } sb_put(&s, ")");
  sb_put(&s, " oo=("); for (int a = 0; a < sp.n_out; a++) { if (a) sb_put(&s, " "); sb_opts(&s, sp.oopt[a], sp.n_oopt[a]); } sb_put(&s, ")");
  arena_free(); return 0;
}

/* ------------------------------------------------------------------------------------------
 * plan-broadcast, src/3einsum.lisp:378-425
 * ---------------------------------------------------------------------------------------- */
#define PLAN(p, M, n1, a, b, c) (p)[((a) * (n1) + (b)) * (M) + (c)]
static int plan_broadcast(int n, const int* ranks, const int64_t (*dims)[NCO_MAX_RANK], int64_t* plan, int* M_out) {
  int M = 0; for (int i = 0; i < n; i++) if (ranks[i] > M) M = ranks[i];        /* rrank */
  int n1 = n + 1;
  for (int i = 0; i < 2 * n1 * (M ? M : 1); i++) plan[i] = 1;                   /* :initial-element 1 */
  for (int i = 0; i < n; i++)                                                    /* populate, right aligned */
    for (int k = 0; k < ranks[i]; k++) PLAN(plan, M, n1, 0, i + 1, M - ranks[i] + k) = dims[i][k];
  for (int i = 0; i < n; i++) for (int j = 0; j < M; j++)                        /* result shape = max */
    if (PLAN(plan, M, n1, 0, i + 1, j) > PLAN(plan, M, n1, 0, 0, j)) PLAN(plan, M, n1, 0, 0, j) = PLAN(plan, M, n1, 0, i + 1, j);
  for (int i = 0; i < n; i++) for (int j = 0; j < M; j++)                        /* verify :413-416 */
    if (!(PLAN(plan, M, n1, 0, i + 1, j) == 1 || PLAN(plan, M, n1, 0, i + 1, j) == PLAN(plan, M, n1, 0, 0, j)))
      return fail("plan-broadcast: incompatible broadcast dims");
  for (int i = 0; i <= n; i++) for (int j = M - 2; j >= 0; j--)                  /* step sizes */
    PLAN(plan, M, n1, 1, i, j) = PLAN(plan, M, n1, 0, i, j + 1) * PLAN(plan, M, n1, 1, i, j + 1);
  *M_out = M; return 0;
}
int nco_plan_broadcast(int n, const int* ranks, const int64_t* dims, int64_t* plan_out) {
  int M; if (plan_broadcast(n, ranks, (const int64_t (*)[NCO_MAX_RANK])dims, plan_out, &M)) return -1;
  return M;
}

/* ------------------------------------------------------------------------------------------
 * shape-resolver (src/3einsum.lisp:287-367), range-width / normalize-index (:274-285)
 * ---------------------------------------------------------------------------------------- */
static int64_t normalize_index(int64_t index, int64_t dimension) {
  if (index < 0) { if (dimension == 0) return 0; int64_t m = index % dimension; if (m < 0) m += dimension; return m; }
  return index < dimension ? index : dimension;
}
static int64_t ceil_div(int64_t a, int64_t b) { int64_t q = a / b, r = a % b; if (r != 0 && ((r > 0) == (b > 0))) q++; return q; }

typedef struct {
  specs_t sp;
  int64_t q[NCO_MAX_IDX]; int q_set[NCO_MAX_IDX];    /* ?n loop limits */
  int bc_used; int M;                                  /* broadcast block rank */
  int64_t plan[2 * (NCO_MAX_T + 1) * NCO_MAX_RANK];
  nco_tensor arr[2 * NCO_MAX_T];                       /* outputs first, then inputs (as the backend iterates) */
  int n_arr;
} ctx_t;

static const opt_t* find_opt(const opt_t* o, int n, int key) { for (int i = 0; i < n; i++) if (o[i].key_is_num && o[i].key == key) return &o[i]; return NULL; }

/* one pass of shape-resolver over `count` arrays; collects broadcast source shapes */
static int resolve(ctx_t* c, const nco_tensor* arrs, int count, int (*spec)[NCO_MAX_RANK], const int* rank,
                   opt_t (*opts)[NCO_MAX_RANK], const int* nopt, int out_p,
                   int* bshape_rank, int64_t (*bshape)[NCO_MAX_RANK], int* n_bshapes) {
  *n_bshapes = 0;
  for (int a = 0; a < count; a++) {
    const nco_tensor* t = &arrs[a];
    int used = 0, bstart = 0, bend = 0;
    /* rank check: without `-` the spec length must equal the array rank (array-dimension would fail) */
    int has_b = 0; for (int k = 0; k < rank[a]; k++) if (spec[a][k] == -1) has_b = 1;
    if (!has_b && t->rank != rank[a]) return fail("array %d has rank %d but its spec has %d indices", a, t->rank, rank[a]);
    if (has_b && t->rank < rank[a] - 1) return fail("array %d has too few axes for its spec", a);
    for (int i = 0; i < rank[a]; i++) {                                        /* forward until `-` */
      int dim = spec[a][i];
      if (dim == -1) { used = 1; bstart = i; break; }
      const opt_t* o = find_opt(opts[a], nopt[a], dim);                        /* (assoc dim option): by index number */
      int64_t D = t->dims[i];
      int64_t start = (o && o->has_start) ? o->start : 0, end = (o && o->has_end) ? o->end : D, step = o ? o->step : 1;
      int64_t w = ceil_div(normalize_index(end, D) - normalize_index(start, D), step);
      if (c->q_set[dim] && c->q[dim] != w) return fail("(assert (= ...)) failed: index %d has widths %lld and %lld", dim, (long long)c->q[dim], (long long)w);
      c->q[dim] = w; c->q_set[dim] = 1;
    }
    if (used) {                                                                /* backward until `-` */
      for (int i = 0; i < rank[a]; i++) {
        int dim = spec[a][rank[a] - 1 - i];
        if (dim == -1) { bend = i; break; }
        const opt_t* o = find_opt(opts[a], nopt[a], dim);
        int axis = t->rank - (i + 1); int64_t D = t->dims[axis];
        int64_t start = (o && o->has_start) ? o->start : 0, end = (o && o->has_end) ? o->end : D, step = o ? o->step : 1;
        int64_t w = ceil_div(normalize_index(end, D) - normalize_index(start, D), step);
        if (c->q_set[dim] && c->q[dim] != w) return fail("(assert (= ...)) failed: index %d", dim);
        c->q[dim] = w; c->q_set[dim] = 1;
      }
      int r = t->rank - bend - bstart; if (r < 0) return fail("broadcast block has negative rank");
      bshape_rank[*n_bshapes] = r;
      for (int k = 0; k < r; k++) bshape[*n_bshapes][k] = t->dims[bstart + k];
      (*n_bshapes)++;
      (void)out_p;
    }
  }
  return 0;
}

static int64_t total_size(const nco_tensor* t) { int64_t n = 1; for (int i = 0; i < t->rank; i++) n *= t->dims[i]; return n; }

/* shape-resolver for inputs, %output-generator, shape-resolver for outputs (einsum-lambda :476-483) */
static int setup(ctx_t* c, const char* text, int n_in, const nco_tensor* in, int n_out, nco_tensor* out) {
  memset(c->q_set, 0, sizeof c->q_set); c->bc_used = 0; c->M = 0;
  if (normalize(text, &c->sp)) return -1;
  specs_t* sp = &c->sp;
  if (n_in != sp->n_in) return fail("einsum: %d input arrays for %d input specs", n_in, sp->n_in);
  if (n_out != sp->n_out) return fail("einsum: %d output arrays for %d output specs", n_out, sp->n_out);
  int brank[NCO_MAX_T]; int64_t bshape[NCO_MAX_T][NCO_MAX_RANK]; int nb;
  if (resolve(c, in, n_in, sp->ispec, sp->irank, sp->iopt, sp->n_iopt, 0, brank, bshape, &nb)) return -1;
  if (nb) {
    c->bc_used = 1;
    if (plan_broadcast(nb, brank, (const int64_t (*)[NCO_MAX_RANK])bshape, c->plan, &c->M)) return -1;
    /* the generated code indexes plan rows by array position: it needs every input to carry the block */
    if (nb != n_in) return fail("oracle: every input must use `-` when one does (plan rows are per input)");
  }
  /* %output-generator: shapes of fresh outputs (:439-468) */
  for (int o = 0; o < n_out; o++) {
    if (out[o].data != NULL) continue;
    int r = 0, seen_b = 0;
    for (int k = 0; k < sp->orank[o]; k++) {
      int dim = sp->ospec[o][k];
      if (dim == -1) { seen_b = 1; for (int j = 0; j < c->M; j++) out[o].dims[r++] = PLAN(c->plan, c->M, nb + 1, 0, 0, j); }
      else { if (!c->q_set[dim]) return fail("output index %d unbound", dim); out[o].dims[r++] = c->q[dim]; }
      if (r > NCO_MAX_RANK) return fail("output rank too large");
    }
    (void)seen_b; out[o].rank = r;
  }
  /* outputs may be supplied externally: resolve them too (out-p) */
  int obrank[NCO_MAX_T]; int64_t obshape[NCO_MAX_T][NCO_MAX_RANK]; int nob;
  if (resolve(c, out, n_out, sp->ospec, sp->orank, sp->oopt, sp->n_oopt, 1, obrank, obshape, &nob)) return -1;
  for (int b = 0; b < nob; b++) {                                               /* %output-result-shapes-match-p :431-437 */
    for (int k = 0; k < obrank[b]; k++)
      if (k >= c->M || obshape[b][k] != PLAN(c->plan, c->M, nb + 1, 0, 0, k)) return fail("The output shapes do not match the broadcast plan!");
  }
  c->n_arr = n_out + n_in;
  for (int o = 0; o < n_out; o++) c->arr[o] = out[o];
  for (int i = 0; i < n_in; i++) c->arr[n_out + i] = in[i];
  return 0;
}

int nco_einsum_shapes(const char* spec, int n_in, const nco_tensor* in, int n_out, nco_tensor* out) {
  ctx_t* c = (ctx_t*)calloc(1, sizeof(ctx_t));
  int r = setup(c, spec, n_in, in, n_out, out);
  free(c); arena_free(); return r;
}

/* ------------------------------------------------------------------------------------------
 * TRANSFORM evaluation
 * ---------------------------------------------------------------------------------------- */
typedef struct { val ev[2 * NCO_MAX_T]; int n_out; } env_t;   /* evars: @1.. then $1.. */

static val eval_form(const sx* f, env_t* e) {
  switch (f->kind) {
    case SX_INT: return v_int(f->ival);
    case SX_F32: return v_f32((float)f->fval);
    case SX_F64: return v_f64(f->fval);
    case SX_SYM: {
      if ((f->name[0] == '$' || f->name[0] == '@') && isdigit((unsigned char)f->name[1])) {
        int n = atoi(f->name + 1);
        return f->name[0] == '@' ? e->ev[n - 1] : e->ev[e->n_out + n - 1];
      }
      g_trap = 1; fail("unbound variable %s in transform", f->name); return v_int(0);
    }
    default: break;
  }
  if (f->n == 0 || f->it[0]->kind != SX_SYM) { g_trap = 1; fail("bad transform form"); return v_int(0); }
  int fn = fn_lookup(f->it[0]->name);
  if (fn < 0) { g_trap = 1; fail("oracle: function %s not in the whitelist", f->it[0]->name); return v_int(0); }
  int na = f->n - 1;
  val a[16]; if (na > 16) { g_trap = 1; return v_int(0); }
  for (int i = 0; i < na; i++) a[i] = eval_form(f->it[i + 1], e);          /* left-to-right evaluation */
  switch (fn) {
    case F_ADD: case F_MUL: {                                                /* (+) = 0, (*) = 1, left fold */
      if (na == 0) return v_int(fn == F_ADD ? 0 : 1);
      val r = a[0]; for (int i = 1; i < na; i++) r = apply2(fn, r, a[i]); return r;
    }
    case F_SUB: case F_DIV: {
      if (na == 1) return fn == F_SUB ? apply1(F_NEG, a[0]) : apply2(F_DIV, v_int(1), a[0]);
      val r = a[0]; for (int i = 1; i < na; i++) r = apply2(fn, r, a[i]); return r;
    }
    case F_MAX: case F_MIN: case F_LOGAND: case F_LOGIOR: case F_LOGXOR: {
      if (na == 0) { if (fn == F_LOGAND) return v_int(-1); return v_int(0); }
      /* (max a b c) = (let ((m (max b c))) (if (>= a m) a m)): right-nested on SBCL */
      if (fn == F_MAX || fn == F_MIN) { val r = a[na - 1]; for (int i = na - 2; i >= 0; i--) r = apply2(fn, a[i], r); return r; }
      val r = a[0]; for (int i = 1; i < na; i++) r = apply2(fn, r, a[i]); return r;
    }
    default:
      if (na == 2) return apply2(fn, a[0], a[1]);
      if (na == 1) return apply1(fn, a[0]);
  }
  g_trap = 1; return v_int(0);
}

/* ------------------------------------------------------------------------------------------
 * The loop nest (src/4einsum-backends/common-lisp.lisp:44-240), interpreted
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  ctx_t* c;
  int n_out, n_in, n_arr;
  int (*aspec)[NCO_MAX_RANK]; int arank[2 * NCO_MAX_T];      /* (append o-specs i-specs) */
  int aspec_store[2 * NCO_MAX_T][NCO_MAX_RANK];
  const opt_t* aopt[2 * NCO_MAX_T]; int anopt[2 * NCO_MAX_T]; /* (append o-options i-options) */
  env_t env;
  int n_plan_rows;
} nest_t;

typedef struct { int64_t idx[2 * NCO_MAX_T]; unsigned initialized, bound; } frame_t;

static int64_t dim_step(nest_t* n, int spec, int array_index, int out_p) {     /* %dimension-step-size :60-66 */
  ctx_t* c = n->c;
  if (spec == -1) return out_p ? PLAN(c->plan, c->M, n->n_plan_rows, 1, 0, 0) : PLAN(c->plan, c->M, n->n_plan_rows, 1, array_index, 0);
  return c->q[spec];
}
static int depends_on(const int* spec, int rank, const int* rest, int nrest) {  /* spec-depends-on */
  for (int i = 0; i < rank; i++) for (int j = 0; j < nrest; j++) if (spec[i] == rest[j]) return 1;
  return 0;
}
static void run_loop(nest_t* n, const int* iter, int niter, frame_t fr);

/* %einsum-early-binding :68-106 */
static void early_binding(nest_t* n, const int* rest, int nrest, frame_t fr) {
  unsigned newly = 0;
  for (int i = 0; i < n->n_arr; i++) {
    if (!depends_on(n->aspec[i], n->arank[i], rest, nrest) && !(fr.bound & (1u << i))) {
      fr.bound |= 1u << i; newly |= 1u << i;
      n->env.ev[i] = load_el(&n->c->arr[i], fr.idx[i]);                     /* (evar (aref var index)) */
    }
  }
  run_loop(n, rest, nrest, fr);
  for (int i = 0; i < n->n_out; i++)                                          /* ,@store : outputs only */
    if (newly & (1u << i)) store_el(&n->c->arr[i], fr.idx[i], n->env.ev[i]);
}

/* %einsum-leaf :226-240 */
static void leaf(nest_t* n) {
  for (int o = 0; o < n->n_out; o++) {
    val v = eval_form(n->c->sp.transforms[o], &n->env);
    n->env.ev[o] = coerce_to(v, n->c->arr[o].dtype);                         /* (setf @k (%coerce transform type)) */
  }
}

/* %einsum-static :108-174 */
static void run_static(nest_t* n, const int* iter, int niter, frame_t fr) {
  int spec = iter[0]; ctx_t* c = n->c;
  int64_t inc[2 * NCO_MAX_T]; int touched[2 * NCO_MAX_T];
  for (int i = 0; i < n->n_arr; i++) {
    int out_p = i < n->n_out; inc[i] = 0; touched[i] = 0;
    for (int ax = 0; ax < n->arank[i]; ax++) {
      if (n->aspec[i][ax] != spec) continue;
      int64_t start = 0, step = 1;                                           /* (assoc a-axis a-opts): by AXIS POSITION :129 */
      const opt_t* o = find_opt(n->aopt[i], n->anopt[i], ax);
      if (o) { if (o->has_start) start = o->start; step = o->step; }
      int64_t ss = 1; for (int r2 = ax + 1; r2 < n->arank[i]; r2++) ss *= dim_step(n, n->aspec[i][r2], i, out_p);
      if (touched[i]) inc[i] += step * ss;                                   /* repeated index (ii): strides add up :137-143 */
      else {
        int64_t off = (fr.initialized & (1u << i)) ? fr.idx[i] : 0;
        fr.idx[i] = off + start * ss; inc[i] = step * ss; touched[i] = 1;
        fr.initialized |= 1u << i;
      }
    }
    if (n->arank[i] == 0 && !(fr.initialized & (1u << i))) { fr.initialized |= 1u << i; fr.idx[i] = 0; }   /* singleton :156-163 */
  }
  for (int64_t k = 0; k < c->q[spec]; k++) {                                 /* (do* ((&n 0 (1+ &n)) ...) ((<= ?n &n))) */
    early_binding(n, iter + 1, niter - 1, fr);
    for (int i = 0; i < n->n_arr; i++) if (touched[i]) fr.idx[i] += inc[i];
  }
}

/* %einsum-broadcast :176-224 */
typedef struct { int arr; int plan_row; int out_p; int64_t step; } bidx_t;
static void bc_rec(nest_t* n, const int* rest, int nrest, frame_t fr, const bidx_t* b, int nb, int axis) {
  ctx_t* c = n->c;
  if (c->M <= axis) { early_binding(n, rest, nrest, fr); return; }
  int64_t ext = PLAN(c->plan, c->M, n->n_plan_rows, 0, 0, axis);
  for (int64_t i = 0; i < ext; i++) {
    bc_rec(n, rest, nrest, fr, b, nb, axis + 1);
    for (int k = 0; k < nb; k++) {
      if (b[k].out_p) fr.idx[b[k].arr] += PLAN(c->plan, c->M, n->n_plan_rows, 1, 0, axis) * b[k].step;
      else fr.idx[b[k].arr] += (PLAN(c->plan, c->M, n->n_plan_rows, 0, b[k].plan_row, axis) == 1) ? 0
                               : PLAN(c->plan, c->M, n->n_plan_rows, 1, b[k].plan_row, axis) * b[k].step;
    }
  }
}
static void run_broadcast(nest_t* n, const int* iter, int niter, frame_t fr) {
  bidx_t b[2 * NCO_MAX_T * 2]; int nb = 0;
  for (int i = 0; i < n->n_arr; i++) {
    int out_p = i < n->n_out;
    for (int ax = 0; ax < n->arank[i]; ax++) {
      if (n->aspec[i][ax] != -1) continue;
      int64_t ss = 1; for (int r2 = ax + 1; r2 < n->arank[i]; r2++) ss *= dim_step(n, n->aspec[i][r2], i, out_p);
      if (!(fr.initialized & (1u << i))) { fr.idx[i] = 0; fr.initialized |= 1u << i; }
      b[nb].arr = i; b[nb].plan_row = i; /* (aref plan 0 ,i axis): position in (append o-specs i-specs) */
      b[nb].out_p = out_p; b[nb].step = ss; nb++;
    }
  }
  bc_rec(n, iter + 1, niter - 1, fr, b, nb, 0);
}

static void run_loop(nest_t* n, const int* iter, int niter, frame_t fr) {       /* %einsum-loop :49-58 */
  if (niter == 0) { leaf(n); return; }
  if (iter[0] == -1) run_broadcast(n, iter, niter, fr);
  else run_static(n, iter, niter, fr);
}

int nco_einsum(const char* spec, int n_in, const nco_tensor* in, int n_out, nco_tensor* out) {
  ctx_t* c = (ctx_t*)calloc(1, sizeof(ctx_t));
  for (int o = 0; o < n_out; o++) if (!out[o].data) { free(c); return fail("nco_einsum: outputs must be allocated"); }
  if (setup(c, spec, n_in, in, n_out, out)) { free(c); arena_free(); return -1; }
  nest_t* n = (nest_t*)calloc(1, sizeof(nest_t));
  n->c = c; n->n_out = n_out; n->n_in = n_in; n->n_arr = n_out + n_in; n->aspec = n->aspec_store;
  n->n_plan_rows = n_in + 1;
  if (c->bc_used && n_out != 1) { free(n); free(c); arena_free(); return fail("oracle: `-` with %d outputs indexes the plan out of step (reference quirk)", n_out); }
  for (int o = 0; o < n_out; o++) { memcpy(n->aspec[o], c->sp.ospec[o], sizeof n->aspec[o]); n->arank[o] = c->sp.orank[o]; n->aopt[o] = c->sp.oopt[o]; n->anopt[o] = c->sp.n_oopt[o]; }
  for (int i = 0; i < n_in; i++) { memcpy(n->aspec[n_out + i], c->sp.ispec[i], sizeof n->aspec[0]); n->arank[n_out + i] = c->sp.irank[i]; n->aopt[n_out + i] = c->sp.iopt[i]; n->anopt[n_out + i] = c->sp.n_iopt[i]; }
  n->env.n_out = n_out;
  /* einsum-lambda keeps only the base vector: the displacement offset is dropped (3einsum.lisp:485-487, A.7) */
  frame_t fr; memset(&fr, 0, sizeof fr);
  g_trap = 0; g_err[0] = 0;
  if (c->sp.n_iter == 0) {   /* no loops: bind everything, run the leaf, store */
    for (int i = 0; i < n->n_arr; i++) n->env.ev[i] = load_el(&c->arr[i], 0);
    leaf(n);
    for (int o = 0; o < n_out; o++) store_el(&c->arr[o], 0, n->env.ev[o]);
  } else run_loop(n, c->sp.iter, c->sp.n_iter, fr);
  int trapped = g_trap;
  free(n); free(c); arena_free();
  if (trapped) { if (!g_err[0]) fail("arithmetic trap (SBCL would signal)"); return -1; }
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * broadcast, src/4numeric.lisp:174-294
 * ---------------------------------------------------------------------------------------- */
int nco_broadcast_shape(const nco_tensor* x, const nco_tensor* y, nco_tensor* r) {
  int rr = x->rank > y->rank ? x->rank : y->rank;
  for (int k = 0; k < rr; k++) {                                             /* broadcast-p, right aligned :174-185 */
    int xi = x->rank - rr + k, yi = y->rank - rr + k;
    int64_t dx = xi >= 0 ? x->dims[xi] : 1, dy = yi >= 0 ? y->dims[yi] : 1;
    if (!(dx == dy || dx == 1 || dy == 1)) return fail("Given arrays have incompatible shape for broadcasting");
    r->dims[k] = dx > dy ? dx : dy;                                          /* (mapcar #'max xshape yshape) */
  }
  r->rank = rr; return 0;
}
typedef struct { int fn; const nco_tensor *x, *y; nco_tensor* r; } bc_t;
static int64_t prod(const int64_t* d, int n) { int64_t p = 1; for (int i = 0; i < n; i++) p *= d[i]; return p; }
static void broadcast_core(bc_t* b, int64_t xo, const int64_t* xs, int64_t yo, const int64_t* ys, int64_t ro, const int64_t* rs, int n) {
  if (n == 0) {   /* rank-0 x rank-0: not reachable through the match clauses with lists, handled as one element */
    store_el(b->r, ro, apply2(b->fn, load_el(b->x, xo), load_el(b->y, yo))); return;
  }
  if (n == 1) {
    if (xs[0] == 1) {                                                        /* ((list 1) (list ydim)) :230-235 */
      val xv = load_el(b->x, xo);
      for (int64_t i = 0; i < ys[0]; i++) store_el(b->r, ro + i, apply2(b->fn, xv, load_el(b->y, yo + i)));
    } else if (ys[0] == 1) {                                                 /* ((list xdim) (list 1)) :236-241 */
      val yv = load_el(b->y, yo);
      for (int64_t i = 0; i < xs[0]; i++) store_el(b->r, ro + i, apply2(b->fn, load_el(b->x, xo + i), yv));
    } else {                                                                 /* ((list xdim) _) :242-249 */
      for (int64_t i = 0; i < xs[0]; i++) store_el(b->r, ro + i, apply2(b->fn, load_el(b->x, xo + i), load_el(b->y, yo + i)));
    }
    return;
  }
  int64_t xrest = prod(xs + 1, n - 1), yrest = prod(ys + 1, n - 1), rrest = prod(rs + 1, n - 1);
  if (xs[0] == 1) {                                                          /* :251-261 */
    for (int64_t i = 0; i < ys[0]; i++) { broadcast_core(b, xo, xs + 1, yo, ys + 1, ro, rs + 1, n - 1); yo += yrest; ro += rrest; }
  } else if (ys[0] == 1) {                                                   /* :263-273 */
    for (int64_t i = 0; i < xs[0]; i++) { broadcast_core(b, xo, xs + 1, yo, ys + 1, ro, rs + 1, n - 1); xo += xrest; ro += rrest; }
  } else {                                                                   /* :275-287 */
    for (int64_t i = 0; i < xs[0]; i++) { broadcast_core(b, xo, xs + 1, yo, ys + 1, ro, rs + 1, n - 1); xo += xrest; yo += yrest; ro += rrest; }
  }
}
int nco_broadcast(const char* op, const nco_tensor* x, const nco_tensor* y, nco_tensor* r) {
  char up[32]; int i = 0; for (; op[i] && i < 31; i++) up[i] = (char)toupper((unsigned char)op[i]); up[i] = 0;
  const char* colon = strrchr(up, ':'); int fn = fn_lookup(colon ? colon + 1 : up);
  if (fn < 0) return fail("oracle: broadcast function %s not in the whitelist", op);
  nco_tensor chk = *r; if (nco_broadcast_shape(x, y, &chk)) return -1;
  if (chk.rank != r->rank) return fail("result rank mismatch");
  for (int k = 0; k < chk.rank; k++) if (chk.dims[k] != r->dims[k]) return fail("result shape mismatch");
  int rr = r->rank; int64_t xs[NCO_MAX_RANK], ys[NCO_MAX_RANK];
  for (int k = 0; k < rr; k++) {                                             /* left-pad with 1s :191-194 */
    int xi = x->rank - rr + k, yi = y->rank - rr + k;
    xs[k] = xi >= 0 ? x->dims[xi] : 1; ys[k] = yi >= 0 ? y->dims[yi] : 1;
  }
  bc_t b = {fn, x, y, r}; g_trap = 0;
  broadcast_core(&b, x->offset, xs, y->offset, ys, 0, r->dims, rr);
  return g_trap ? fail("arithmetic trap in broadcast") : 0;
}

/* ------------------------------------------------------------------------------------------
 * reduce-array nest, src/5reduce.lisp:67-87: dotimes over all axes in order 0..rank-1,
 * leaf (setf (aref r kept...) (funcall fn (aref r kept...) (aref a all...)))
 * ---------------------------------------------------------------------------------------- */
int nco_reduce(const char* op, const nco_tensor* a, const int* axes, int n_axes, nco_tensor* r) {
  char up[32]; int i = 0; for (; op[i] && i < 31; i++) up[i] = (char)toupper((unsigned char)op[i]); up[i] = 0;
  const char* colon = strrchr(up, ':'); int fn = fn_lookup(colon ? colon + 1 : up);
  if (fn < 0) return fail("oracle: reduce function %s not in the whitelist", op);
  int summed[NCO_MAX_RANK] = {0};
  for (int k = 0; k < n_axes; k++) { if (axes[k] < 0 || axes[k] >= a->rank) return fail("bad axis"); summed[axes[k]] = 1; }
  /* result shape = remaining axes in order */
  int rk = 0; for (int ax = 0; ax < a->rank; ax++) if (!summed[ax]) { if (rk >= r->rank || r->dims[rk] != a->dims[ax]) return fail("reduce: result shape mismatch"); rk++; }
  if (rk != r->rank) return fail("reduce: result rank mismatch");
  int64_t astride[NCO_MAX_RANK], rstride[NCO_MAX_RANK]; int64_t s = 1;
  for (int ax = a->rank - 1; ax >= 0; ax--) { astride[ax] = s; s *= a->dims[ax]; }
  s = 1; { int k = r->rank - 1; for (int ax = a->rank - 1; ax >= 0; ax--) { if (summed[ax]) rstride[ax] = 0; else { rstride[ax] = s; s *= r->dims[k--]; } } }
  int64_t cnt[NCO_MAX_RANK] = {0}; int64_t n = total_size(a);
  g_trap = 0;
  for (int64_t it = 0; it < n; it++) {                                       /* row-major odometer == the dotimes nest */
    int64_t ai = a->offset, ri = r->offset;
    for (int ax = 0; ax < a->rank; ax++) { ai += cnt[ax] * astride[ax]; ri += cnt[ax] * rstride[ax]; }
    /* reduce-array full-calls the function symbol: (max r a) = (if (> a r) a r); same tie behaviour */
    store_el(r, ri, apply2(fn, load_el(r, ri), load_el(a, ai)));
    for (int ax = a->rank - 1; ax >= 0; ax--) { if (++cnt[ax] < a->dims[ax]) break; cnt[ax] = 0; }
  }
  return g_trap ? fail("arithmetic trap in reduce") : 0;
}

/* map-array-into :37-51 */
int nco_map(const char* form, int n, const nco_tensor* args, nco_tensor* r) {
  const char* p = form; sx* top = parse_list(&p, 1);
  if (!top || top->n != 1) { arena_free(); return fail("nco_map: expected one form"); }
  for (int k = 0; k < n; k++) {
    if (args[k].rank != r->rank) { arena_free(); return fail("map-array: shapes differ"); }
    for (int d = 0; d < r->rank; d++) if (args[k].dims[d] != r->dims[d]) { arena_free(); return fail("map-array: shapes differ"); }
  }
  env_t e; memset(&e, 0, sizeof e); e.n_out = 0; g_trap = 0;
  int64_t total = total_size(r);
  for (int64_t i = 0; i < total; i++) {
    for (int k = 0; k < n; k++) e.ev[k] = load_el(&args[k], args[k].offset + i);
    store_el(r, r->offset + i, eval_form(top->it[0], &e));
  }
  arena_free();
  return g_trap ? fail("arithmetic trap in map") : 0;
}

/* ------------------------------------------------------------------------------------------
 * typed fast loops (see header)
 * ---------------------------------------------------------------------------------------- */
void nco_fast_add2_f32(const float* a, const float* b, float* tmp, float* r, int64_t rows, int64_t cols) {
  /* pass 1: (broadcast '+ 0 a): x is the rank-0 bit array holding 0, loaded once and converted, :230-235 */
  const float zero = (float)0;
  for (int64_t i = 0; i < rows; i++)
    for (int64_t j = 0; j < cols; j++) tmp[i * cols + j] = zero + a[i * cols + j];
  /* pass 2: (broadcast '+ tmp b): rows advance only tmp (y has leading dim 1), inner loop elementwise */
  for (int64_t i = 0; i < rows; i++)
    for (int64_t j = 0; j < cols; j++) r[i * cols + j] = tmp[i * cols + j] + b[j];
}
void nco_fast_mul2_u8_fix(const uint8_t* x, const int64_t* yt, uint8_t* tmp, int64_t* rt, int64_t outer, int64_t inner) {
  for (int64_t i = 0; i < outer * inner; i++) tmp[i] = (uint8_t)(1 * x[i]);                 /* (integer 0 255) stays ub8 */
  for (int64_t i = 0; i < outer; i++)
    for (int64_t j = 0; j < inner; j++) {
      i128 p = (i128)tmp[i * inner + j] * (i128)(yt[j] >> 1);
      rt[i * inner + j] = (int64_t)((uint64_t)wrap_int(p, NCO_FIXNUM) << 1);
    }
}
void nco_fast_sum_axis1_f32(const float* a, float* r, int64_t rows, int64_t cols) {
  for (int64_t i = 0; i < rows; i++) {
    float acc = 0.0f;                                                         /* (zero-value 'single-float) */
    for (int64_t j = 0; j < cols; j++) acc = acc + a[i * cols + j];
    r[i] = acc;
  }
}
float nco_fast_sum_all_f32(const float* a, int64_t n) {
  float acc = 0.0f; for (int64_t i = 0; i < n; i++) acc = acc + a[i]; return acc;
}
#define MATMUL_BODY(T)                                                                      \
  for (int64_t i = r0; i < r1; i++)                                                         \
    for (int64_t j = 0; j < k; j++) {                                                       \
      T av = a[i * k + j];                      /* $1 hoisted out of the k loop */          \
      T* crow = c + (i - c_row0) * n; const T* brow = b + j * n;                            \
      for (int64_t l = 0; l < n; l++) { T t = av * brow[l]; crow[l] = crow[l] + t; }        \
    }
void nco_fast_matmul_f64(const double* a, const double* b, double* c, int64_t m, int64_t k, int64_t n) {
  int64_t r0 = 0, r1 = m, c_row0 = 0; MATMUL_BODY(double)
}
void nco_fast_matmul_f32(const float* a, const float* b, float* c, int64_t m, int64_t k, int64_t n) {
  int64_t r0 = 0, r1 = m, c_row0 = 0; MATMUL_BODY(float)
}
void nco_fast_matmul_rows_f64(const double* a, const double* b, double* c, int64_t r0, int64_t r1, int64_t k, int64_t n) {
  int64_t c_row0 = r0; MATMUL_BODY(double)
}
void nco_fast_matmul_rows_f32(const float* a, const float* b, float* c, int64_t r0, int64_t r1, int64_t k, int64_t n) {
  int64_t c_row0 = r0; MATMUL_BODY(float)
}
void nco_fast_bmm_f32(const float* a, const float* b, float* c, int64_t nb, int64_t m, int64_t k, int64_t n) {
  for (int64_t q = 0; q < nb; q++) nco_fast_matmul_f32(a + q * m * k, b + q * k * n, c + q * m * n, m, k, n);
}
void nco_fast_transpose_abcd_dbca_w64(const int64_t* in, int64_t* out, int64_t A, int64_t B, int64_t C, int64_t D) {
  /* (+ @1 (* $1)) on fixnums into a zeroed output; loops a,b,c,d */
  for (int64_t a = 0; a < A; a++) for (int64_t b = 0; b < B; b++) for (int64_t c = 0; c < C; c++) for (int64_t d = 0; d < D; d++) {
    int64_t o = ((d * B + b) * C + c) * A + a;
    i128 v = (i128)(out[o] >> 1) + (i128)(in[((a * B + b) * C + c) * D + d] >> 1);
    out[o] = (int64_t)((uint64_t)wrap_int(v, NCO_FIXNUM) << 1);
  }
}

/* ------------------------------------------------------------------------------------------
 * synthetic inputs: splitmix64(seed ^ index), mirrored bit for bit by the CUDA fill kernel
 * ---------------------------------------------------------------------------------------- */
static uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
int nco_fill_random(nco_tensor* t, uint64_t seed, int dist, double lo, double hi) { return nco_fill_random_at(t, seed, 0, dist, lo, hi); }
/* element k of the tensor is element first + k of the synthetic stream (a row shard of a larger array) */
int nco_fill_random_at(nco_tensor* t, uint64_t seed, uint64_t first, int dist, double lo, double hi) {
  int64_t n = total_size(t); const dtinfo* d = &DT[t->dtype];
  for (int64_t i = 0; i < n; i++) {
    uint64_t h = splitmix64(seed ^ (first + (uint64_t)i));
    val v;
    if (d->is_float && dist != 1) {
      if (d->is_float == 1) { float u = (float)(h >> 40) * 0x1p-24f; float w = (float)hi - (float)lo; float m = w * u; v = v_f32((float)lo + m); }
      else { double u = (double)(h >> 11) * 0x1p-53; double w = hi - lo; double m = w * u; v = v_f64(lo + m); }
      if (dist == 2) {
        uint64_t h2 = splitmix64(h);
        if ((h2 & 1023) == 0) {
          int sel = (int)((h2 >> 10) % 6);
          if (d->is_float == 1) { const float sp[6] = {0.0f, -0.0f, INFINITY, -INFINITY, 0x1p-149f, 0x1.fffffep+126f}; v = v_f32(sp[sel]); }
          else { const double sp[6] = {0.0, -0.0, INFINITY, -INFINITY, 0x1p-1074, 0x1.fffffffffffffp+1022}; v = v_f64(sp[sel]); }
        }
      }
    } else {
      int64_t l = (int64_t)lo, hh = (int64_t)hi; uint64_t span = (uint64_t)(hh - l) + 1;
      int64_t iv = l + (int64_t)(span ? h % span : h);
      v = v_int(iv);
    }
    store_el(t, t->offset + i, v);
  }
  return 0;
}
