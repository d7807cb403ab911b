// k_generic.cu -- the reference-ordered generic nest kernel.
//
// One thread owns one output element and walks the contracted loops serially in the nest's own
// order, so each output accumulates exactly as the reference's loop nest does
// (src/4einsum-backends/common-lisp.lisp:108-240: for every output element the contraction
// indices ascend in iter-specs order and the leaf is  @k <- (%coerce transform type)).
// That makes this kernel bit-identical to the reference even for float contractions.  It is the
// completeness path -- every spec/transform on the whitelist that no specialised kernel family
// claims (diagonals `ii`, kron `ij kl -> ikjl`, multi-output specs, iteration options, odd
// transforms) runs here -- not a fast path: operands are read with whatever strides the spec
// gives and the transform is interpreted.
#include "ncl_fn.cuh"

struct GNode { short op, a, b, cls; long long c; };

struct GenericParams {
  int n_kept, n_con, n_arr, n_nodes;
  long long ext[NCL_MAX_LOOPS];                       // kept loops first (fastest last), then contracted (nest order)
  long long str[NCL_MAX_ARR][NCL_MAX_LOOPS];          // array 0 = the output, 1.. = inputs
  char* base[NCL_MAX_ARR]; long long off[NCL_MAX_ARR]; int lk[NCL_MAX_ARR]; int cls[NCL_MAX_ARR];
  StoreSpec st; int out_cls, out_bits, out_signed;
  int fresh; int has_identity; long long ident_i; double ident_f;
  long long total;
  unsigned int* fault;                                // sticky DOMAIN flag (ncl_fn.cuh)
  GNode node[NCL_MAX_EXPR];
};

struct GV { long long i, j; };   // i: int / float bits / double bits / ratio numerator; j: ratio denominator

__device__ __forceinline__ GV gv_i(long long v) { GV g; g.i = v; g.j = 1; return g; }
__device__ __forceinline__ GV gv_f(float v) { GV g; g.i = __float_as_int(v); g.j = 0; return g; }
__device__ __forceinline__ GV gv_d(double v) { GV g; g.i = __double_as_longlong(v); g.j = 0; return g; }
__device__ __forceinline__ float gf(GV g) { return __int_as_float((int)g.i); }
__device__ __forceinline__ double gd(GV g) { return __longlong_as_double(g.i); }
// complex numbers: i = real part, j = imaginary part (float or double bits by class)
__device__ __forceinline__ GV gv_cf(float re, float im) { GV g; g.i = __float_as_int(re); g.j = __float_as_int(im); return g; }
__device__ __forceinline__ GV gv_cd(double re, double im) { GV g; g.i = __double_as_longlong(re); g.j = __double_as_longlong(im); return g; }
__device__ __forceinline__ float gfi(GV g) { return __int_as_float((int)g.j); }
__device__ __forceinline__ double gdi(GV g) { return __longlong_as_double(g.j); }
__device__ __forceinline__ bool is_cx(int cls) { return cls == CLS_CF || cls == CLS_CD; }

__device__ inline double ratio_to_f64(long long n, long long d) {
  // exact operands -> one correctly rounded division when both fit 53 bits; beyond that the
  // operands are rounded first (documented limit of the generic path)
  return __ddiv_rn((double)n, (double)d);
}
__device__ inline float gv_to_f32(GV g, int cls) {
  switch (cls) { case CLS_I: return __ll2float_rn(g.i); case CLS_F: return gf(g); case CLS_D: return __double2float_rn(gd(g));
    default: return ratio_to_f32(g.i, g.j); }
}
__device__ inline double gv_to_f64(GV g, int cls) {
  switch (cls) { case CLS_I: return __ll2double_rn(g.i); case CLS_F: return (double)gf(g); case CLS_D: return gd(g);
    default: return ratio_to_f64(g.i, g.j); }
}
// (round x): half to even
__device__ inline long long gv_round(GV g, int cls) {
  switch (cls) {
    case CLS_I: return g.i;
    case CLS_F: return round_to_i64(gf(g));
    case CLS_D: return round_to_i64(gd(g));
    default: {
      long long n = g.i, d = g.j; if (d < 0) { n = -n; d = -d; }
      if (d == 0) return 0;
      long long q = n / d, r = n % d; if (r < 0) { q -= 1; r += d; }
      long long twice = 2 * r;
      if (twice > d || (twice == d && (q & 1))) q += 1;
      return q;
    }
  }
}
__device__ __forceinline__ GV gv_convert(GV g, int from, int to) {
  if (from == to) return g;
  if (to == CLS_F) return gv_f(gv_to_f32(g, from));
  if (to == CLS_D) return gv_d(gv_to_f64(g, from));
  return g;
}
// %coerce into the output's logical value (what a later @k read sees)
__device__ inline GV coerce_out(GV v, int cls, const GenericParams& p) {
  if (p.out_cls == CLS_CF) {                                           // (coerce x '(complex single-float))
    if (cls == CLS_CF) return v;
    if (cls == CLS_CD) return gv_cf(__double2float_rn(gd(v)), __double2float_rn(gdi(v)));
    return gv_cf(gv_to_f32(v, cls), 0.0f);
  }
  if (p.out_cls == CLS_CD) {
    if (cls == CLS_CD) return v;
    if (cls == CLS_CF) return gv_cd((double)gf(v), (double)gfi(v));
    return gv_cd(gv_to_f64(v, cls), 0.0);
  }
  if (p.out_cls == CLS_F) return gv_f(gv_to_f32(v, cls));
  if (p.out_cls == CLS_D) return gv_d(gv_to_f64(v, cls));
  long long x = gv_round(v, cls);
  if (p.out_bits < 64) {
    if (p.out_signed) { int sh = 64 - p.out_bits; x = (long long)((unsigned long long)x << sh) >> sh; }
    else x &= (long long)((1ull << p.out_bits) - 1ull);
  }
  return gv_i(x);
}


// ---- complex arithmetic as SBCL's number dispatch does it (src/code/numbers.lisp [SBCL, unverified]): every float operation
// rounded on its own; complex o complex: + - componentwise, * = (ac - bd, ad + bc), / = Smith's scaling by the larger part of
// the divisor; real o complex touches only the parts the real takes part in ((+ x #C(a b)) = #C((+ x a) b), (* x #C(a b)) =
// #C((* x a) (* x b)), (/ #C(a b) x) = #C((/ a x) (/ b x))); real / complex goes through the general quotient.
template <class T> struct CxOps;
template <> struct CxOps<float> {
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); } static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); } static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
};
template <> struct CxOps<double> {
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); } static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); } static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
};
template <class T> __device__ inline void cx_binary(int op, bool acx, T ar, T ai, bool bcx, T br, T bi, T* rr, T* ri) {
  typedef CxOps<T> O;
  switch (op) {
    case NCL_OP_ADD: *rr = O::add(ar, br); *ri = acx && bcx ? O::add(ai, bi) : (acx ? ai : bi); return;
    case NCL_OP_SUB: *rr = O::sub(ar, br); *ri = acx && bcx ? O::sub(ai, bi) : (acx ? ai : -bi); return;
    case NCL_OP_MUL:
      if (acx && bcx) { *rr = O::sub(O::mul(ar, br), O::mul(ai, bi)); *ri = O::add(O::mul(ar, bi), O::mul(ai, br)); }
      else if (acx) { *rr = O::mul(ar, br); *ri = O::mul(ai, br); }
      else { *rr = O::mul(ar, br); *ri = O::mul(ar, bi); }
      return;
    default:                                                          // NCL_OP_DIV
      if (!bcx) { *rr = O::div(ar, br); *ri = O::div(ai, br); return; }
      if (!acx) ai = (T)0;
      if (fabs((double)br) > fabs((double)bi)) {
        const T r = O::div(bi, br), dn = O::add(br, O::mul(r, bi));
        *rr = O::div(O::add(ar, O::mul(ai, r)), dn); *ri = O::div(O::sub(ai, O::mul(ar, r)), dn);
      } else {
        const T r = O::div(br, bi), dn = O::add(bi, O::mul(r, br));
        *rr = O::div(O::add(O::mul(ar, r), ai), dn); *ri = O::div(O::sub(O::mul(ai, r), ar), dn);
      }
  }
}

__device__ inline GV eval_node(const GNode& nd, const GV* val, const GNode* nodes, unsigned int* flt) {
  int op = nd.op;
  if (op < NCL_OP_BINARY_END) {
    int ca = nodes[nd.a].cls, cb = nodes[nd.b].cls;
    GV a = val[nd.a], b = val[nd.b];
    if (is_cx(ca) || is_cx(cb)) {                                        // the host typer admits + - * / only
      const bool acx = is_cx(ca), bcx = is_cx(cb);
      if (nd.cls == CLS_CD) {
        const double ar = acx ? (ca == CLS_CD ? gd(a) : (double)gf(a)) : gv_to_f64(a, ca), ai = acx ? (ca == CLS_CD ? gdi(a) : (double)gfi(a)) : 0.0;
        const double br = bcx ? (cb == CLS_CD ? gd(b) : (double)gf(b)) : gv_to_f64(b, cb), bi = bcx ? (cb == CLS_CD ? gdi(b) : (double)gfi(b)) : 0.0;
        double rr, ri; cx_binary<double>(op, acx, ar, ai, bcx, br, bi, &rr, &ri); return gv_cd(rr, ri);
      }
      const float ar = acx ? gf(a) : gv_to_f32(a, ca), ai = acx ? gfi(a) : 0.0f, br = bcx ? gf(b) : gv_to_f32(b, cb), bi = bcx ? gfi(b) : 0.0f;
      float rr, ri; cx_binary<float>(op, acx, ar, ai, bcx, br, bi, &rr, &ri); return gv_cf(rr, ri);
    }
    // comparisons between an integer and a float are exact in Common Lisp (CLHS 12.1.4.1), not after contagion
    if (op >= NCL_OP_EQ && op <= NCL_OP_GT && ((ca == CLS_I) != (cb == CLS_I)) && ca != CLS_R && cb != CLS_R) {
      int c;
      if (ca == CLS_I) c = cmp_i64_f64(a.i, cb == CLS_F ? (double)gf(b) : gd(b));
      else { c = cmp_i64_f64(b.i, ca == CLS_F ? (double)gf(a) : gd(a)); if (c != 2) c = -c; }
      return gv_i(cmp_result(op, c));
    }
    if (op == NCL_OP_EXPT) {
      if (cb == CLS_I) {                                                  // integer power: intexp
        const long long pw = b.i;
        if (ca == CLS_I) {
          if (pw >= 0) return gv_i(int_expt(a.i, pw, flt));
          if (a.i == 0) { ncl_fault(flt, FAULT_DIV0); return gv_i(0); }
          GV g; g.i = 1; g.j = int_expt(a.i, -pw, flt); if (g.j < 0) { g.i = -1; g.j = -g.j; } return g;      // exact ratio until %coerce
        }
        if (ca == CLS_F) return gv_f(float_intexp<float>(gf(a), pw));
        return gv_d(float_intexp<double>(gd(a), pw));
      }
      // float power: real-expt (pow in double-float, coerced back): tolerance only
      const double x = gv_to_f64(a, ca), y = gv_to_f64(b, cb);
      if (x < 0.0 && y != rint(y)) ncl_fault(flt, FAULT_DOMAIN);
      const double r = pow(x, y);
      return (ca == CLS_D || cb == CLS_D) ? gv_d(r) : gv_f((float)r);
    }
    if (op >= NCL_OP_FLOOR && op <= NCL_OP_REM) {
      const bool want_rem = op == NCL_OP_MOD || op == NCL_OP_REM;
      const int dop = op == NCL_OP_REM ? NCL_OP_TRUNCATE : op;
      if (ca == CLS_I && cb == CLS_I) { long long r; const long long q = int_divide(dop, a.i, b.i, &r, flt); return gv_i(want_rem ? r : q); }
      if (ca == CLS_D || cb == CLS_D) {
        const double x = gv_to_f64(a, ca), y = gv_to_f64(b, cb); double r;
        if (op == NCL_OP_ROUND && cb == CLS_I && b.i == 1) return gv_i(float_round_unary<double>(x));
        const long long q = float_divide<double>(dop, x, y, &r); return want_rem ? gv_d(r) : gv_i(q);
      }
      const float x = gv_to_f32(a, ca), y = gv_to_f32(b, cb); float r;
      if (op == NCL_OP_ROUND && cb == CLS_I && b.i == 1) return gv_i(float_round_unary<float>(x));
      const long long q = float_divide<float>(dop, x, y, &r); return want_rem ? gv_f(r) : gv_i(q);
    }
    int cc = (ca == CLS_D || cb == CLS_D) ? CLS_D : (ca == CLS_F || cb == CLS_F) ? CLS_F : CLS_I;   // contagion
    if (cc == CLS_I && (ca == CLS_R || cb == CLS_R)) cc = CLS_F;       // rejected by the host typer; defensive
    if (cc == CLS_I) {
      long long x = a.i, y = b.i;
      switch (op) {
        case NCL_OP_ADD: return gv_i(Ops<long long>::add(x, y));
        case NCL_OP_SUB: return gv_i(Ops<long long>::sub(x, y));
        case NCL_OP_MUL: return gv_i(Ops<long long>::mul(x, y));
        case NCL_OP_DIV: { GV g; g.i = x; g.j = y; return g; }          // exact ratio until %coerce
        case NCL_OP_MAX: return gv_i(x >= y ? x : y);
        case NCL_OP_MIN: return gv_i(x <= y ? x : y);
        case NCL_OP_EQ: return gv_i(x == y); case NCL_OP_NE: return gv_i(x != y);
        case NCL_OP_LE: return gv_i(x <= y); case NCL_OP_GE: return gv_i(x >= y);
        case NCL_OP_LT: return gv_i(x < y);  case NCL_OP_GT: return gv_i(x > y);
        default: return gv_i(int_bitwise(op, x, y));
      }
    } else if (cc == CLS_F) {
      float x = gv_to_f32(a, ca), y = gv_to_f32(b, cb);
      switch (op) {
        case NCL_OP_ADD: return gv_f(__fadd_rn(x, y)); case NCL_OP_SUB: return gv_f(__fsub_rn(x, y));
        case NCL_OP_MUL: return gv_f(__fmul_rn(x, y)); case NCL_OP_DIV: return gv_f(__fdiv_rn(x, y));
        case NCL_OP_MAX: return gv_f(x >= y ? x : y);  case NCL_OP_MIN: return gv_f(x <= y ? x : y);
        case NCL_OP_EQ: return gv_i(x == y); case NCL_OP_NE: return gv_i(x != y);
        case NCL_OP_LE: return gv_i(x <= y); case NCL_OP_GE: return gv_i(x >= y);
        case NCL_OP_LT: return gv_i(x < y);  default: return gv_i(x > y);
      }
    } else {
      double x = gv_to_f64(a, ca), y = gv_to_f64(b, cb);
      switch (op) {
        case NCL_OP_ADD: return gv_d(__dadd_rn(x, y)); case NCL_OP_SUB: return gv_d(__dsub_rn(x, y));
        case NCL_OP_MUL: return gv_d(__dmul_rn(x, y)); case NCL_OP_DIV: return gv_d(__ddiv_rn(x, y));
        case NCL_OP_MAX: return gv_d(x >= y ? x : y);  case NCL_OP_MIN: return gv_d(x <= y ? x : y);
        case NCL_OP_EQ: return gv_i(x == y); case NCL_OP_NE: return gv_i(x != y);
        case NCL_OP_LE: return gv_i(x <= y); case NCL_OP_GE: return gv_i(x >= y);
        case NCL_OP_LT: return gv_i(x < y);  default: return gv_i(x > y);
      }
    }
  }
  // unary
  int ca = nodes[nd.a].cls; GV a = val[nd.a];
  if (is_cx(ca)) {
    if (ca == CLS_CF) {
      const float re = gf(a), im = gfi(a);
      switch (op) {
        case NCL_OP_NEG: return gv_cf(-re, -im);
        case NCL_OP_CONJUGATE: return gv_cf(re, -im);
        case NCL_OP_REALPART: return gv_f(re);
        case NCL_OP_IMAGPART: return gv_f(im);
        case NCL_OP_ABS: return gv_f(hypotf(re, im));                       // %hypot: tolerance only
        case NCL_OP_SQUARE: { float rr, ri; cx_binary<float>(NCL_OP_MUL, true, re, im, true, re, im, &rr, &ri); return gv_cf(rr, ri); }
        default: return a;
      }
    }
    const double re = gd(a), im = gdi(a);
    switch (op) {
      case NCL_OP_NEG: return gv_cd(-re, -im);
      case NCL_OP_CONJUGATE: return gv_cd(re, -im);
      case NCL_OP_REALPART: return gv_d(re);
      case NCL_OP_IMAGPART: return gv_d(im);
      case NCL_OP_ABS: return gv_d(hypot(re, im));
      case NCL_OP_SQUARE: { double rr, ri; cx_binary<double>(NCL_OP_MUL, true, re, im, true, re, im, &rr, &ri); return gv_cd(rr, ri); }
      default: return a;
    }
  }
  switch (op) {
    case NCL_OP_CONJUGATE: case NCL_OP_REALPART: return a;               // of a real: the number itself
    case NCL_OP_IMAGPART: return ca == CLS_I ? gv_i(0) : ca == CLS_F ? gv_f(__fmul_rn(0.0f, gf(a))) : gv_d(__dmul_rn(0.0, gd(a)));   // (imagpart x) = (* 0 x)
    case NCL_OP_IDENTITY: return a;
    case NCL_OP_LOGNOT: return gv_i(~a.i);
    case NCL_OP_ISQRT: return gv_i(int_isqrt(a.i, flt));
    case NCL_OP_LOGCOUNT: return gv_i(int_logcount(a.i));
    case NCL_OP_INTEGER_LENGTH: return gv_i(int_length(a.i));
    case NCL_OP_NEG:
      if (ca == CLS_I) return gv_i(Ops<long long>::neg(a.i));
      if (ca == CLS_F) return gv_f(-gf(a));
      return gv_d(-gd(a));
    case NCL_OP_ABS:
      if (ca == CLS_I) return gv_i(Ops<long long>::abs_(a.i));
      if (ca == CLS_F) return gv_f(fabsf(gf(a)));
      return gv_d(fabs(gd(a)));
    case NCL_OP_SQUARE:
      if (ca == CLS_I) return gv_i(Ops<long long>::mul(a.i, a.i));
      if (ca == CLS_F) return gv_f(__fmul_rn(gf(a), gf(a)));
      return gv_d(__dmul_rn(gd(a), gd(a)));
    case NCL_OP_SIGNUM:
      if (ca == CLS_I) return gv_i(a.i > 0 ? 1 : (a.i < 0 ? -1 : 0));
      if (ca == CLS_F) { float x = gf(a); return gv_f(x > 0 ? 1.0f : (x < 0 ? -1.0f : x)); }
      { double x = gd(a); return gv_d(x > 0 ? 1.0 : (x < 0 ? -1.0 : x)); }
    default: break;
  }
  // irrational functions: rationals are substituted by single-floats (src/1type.lisp:358-416);
  // SBCL calls libm, so these are tolerance-checked, never bit-checked
  if (nd.cls == CLS_D) return gv_d(float_unary<double>(op, gv_to_f64(a, ca), flt));
  return gv_f(float_unary<float>(op, gv_to_f32(a, ca), flt));
}

__device__ inline GV load_gv(const char* base, long long idx, int lk, int cls) {
  if (cls == CLS_CF) { const float2 v = *reinterpret_cast<const float2*>(base + idx * 8); return gv_cf(v.x, v.y); }
  if (cls == CLS_CD) { const double2 v = *reinterpret_cast<const double2*>(base + idx * 16); return gv_cd(v.x, v.y); }
  if (cls == CLS_F) return gv_f(load_scalar<float>(base, idx, lk));
  if (cls == CLS_D) return gv_d(load_scalar<double>(base, idx, lk));
  return gv_i(load_scalar<long long>(base, idx, lk));
}

__global__ void __launch_bounds__(128) generic_nest_kernel(const __grid_constant__ GenericParams p) {
  long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= p.total) return;
  long long pos[NCL_MAX_ARR];
  for (int a = 0; a < p.n_arr; a++) pos[a] = p.off[a];
  {
    long long rem = tid;
    for (int l = p.n_kept - 1; l >= 0; l--) {
      long long c = rem % p.ext[l]; rem /= p.ext[l];
      for (int a = 0; a < p.n_arr; a++) pos[a] += c * p.str[a][l];
    }
  }
  GV acc;
  if (!p.fresh) acc = load_gv(p.base[0], pos[0], p.lk[0], p.out_cls);
  else if (p.out_cls == CLS_CF) acc = gv_cf(p.has_identity ? (float)p.ident_f : 0.0f, 0.0f);
  else if (p.out_cls == CLS_CD) acc = gv_cd(p.has_identity ? p.ident_f : 0.0, 0.0);
  else if (p.has_identity) acc = (p.out_cls == CLS_I) ? gv_i(p.ident_i) : (p.out_cls == CLS_F) ? gv_f((float)p.ident_f) : gv_d(p.ident_f);
  else acc = (p.out_cls == CLS_I) ? gv_i(0) : (p.out_cls == CLS_F) ? gv_f(0.0f) : gv_d(0.0);

  long long cnt[NCL_MAX_LOOPS];
  long long iters = 1;
  for (int l = 0; l < p.n_con; l++) { cnt[l] = 0; iters *= p.ext[p.n_kept + l]; }
  GV val[NCL_MAX_EXPR];
  unsigned int flt_local = 0;
  for (long long it = 0; it < iters; it++) {
    for (int k = 0; k < p.n_nodes; k++) {
      const GNode& nd = p.node[k];
      switch (nd.op) {
        case NCL_X_INPUT:  val[k] = load_gv(p.base[nd.a], pos[nd.a], p.lk[nd.a], p.cls[nd.a]); break;
        case NCL_X_OUTPUT: val[k] = acc; break;
        case NCL_X_CONST_INT: val[k] = gv_i(nd.c); break;
        case NCL_X_CONST_F32: val[k] = gv_f((float)__longlong_as_double(nd.c)); break;
        case NCL_X_CONST_F64: val[k] = gv_d(__longlong_as_double(nd.c)); break;
        default: val[k] = eval_node(nd, val, p.node, &flt_local);
      }
    }
    acc = coerce_out(val[p.n_nodes - 1], p.node[p.n_nodes - 1].cls, p);     // (setf @k (%coerce transform type))
    // odometer over the contracted loops, innermost (last) fastest: the nest's own order
    for (int l = p.n_con - 1; l >= 0; l--) {
      int L = p.n_kept + l;
      cnt[l]++;
      for (int a = 0; a < p.n_arr; a++) pos[a] += p.str[a][L];
      if (cnt[l] < p.ext[L]) break;
      for (int a = 0; a < p.n_arr; a++) pos[a] -= cnt[l] * p.str[a][L];
      cnt[l] = 0;
    }
  }
  ncl_fault_flush(p.fault, flt_local);
  if (p.out_cls == CLS_CF) *reinterpret_cast<float2*>(p.base[0] + pos[0] * 8) = make_float2(gf(acc), gfi(acc));
  else if (p.out_cls == CLS_CD) *reinterpret_cast<double2*>(p.base[0] + pos[0] * 16) = make_double2(gd(acc), gdi(acc));
  else if (p.out_cls == CLS_F) store_scalar<float>(p.base[0], p.off[0] + (pos[0] - p.off[0]), p.st, gf(acc));
  else if (p.out_cls == CLS_D) store_scalar<double>(p.base[0], pos[0], p.st, gd(acc));
  else store_scalar<long long>(p.base[0], pos[0], p.st, acc.i);
}

static bool expr_uses_only_output(const Expr& e, int self_out) {
  for (int k = 0; k < e.n; k++) if (e.node[k].op == NCL_X_OUTPUT && e.node[k].a != self_out + 1) return false;
  return true;
}

int launch_generic(const Nest& n) {
  Context& c = ctx();
  if (n.n_in + 1 > NCL_MAX_ARR) return set_error(NCL_ERR_UNSUPPORTED, "too many arrays");
  for (int o = 0; o < n.n_out; o++) {
    const Expr& e = n.transform[o];
    if (!expr_uses_only_output(e, o))
      return set_error(NCL_ERR_UNSUPPORTED, "transform %d reads another output's @ variable: outputs are computed independently", o + 1);
    GenericParams* pp = new GenericParams; GenericParams& p = *pp; memset(&p, 0, sizeof p);
    const Operand& out = n.out[o];
    // kept loops: those that move this output; sort by output stride descending so the last is the fastest
    int kept[NCL_MAX_LOOPS], nk = 0, con[NCL_MAX_LOOPS], nc = 0;
    for (int l = 0; l < n.nloops; l++) {
      if (n.extent[l] == 1) continue;
      if (out.stride[l] != 0) kept[nk++] = l; else con[nc++] = l;
    }
    for (int i = 1; i < nk; i++) { int k = kept[i], j = i - 1; while (j >= 0 && out.stride[kept[j]] < out.stride[k]) { kept[j + 1] = kept[j]; j--; } kept[j + 1] = k; }
    p.n_kept = nk; p.n_con = nc; p.n_arr = 1 + n.n_in; p.total = 1;
    // an empty KEPT loop means there is no output element; an empty CONTRACTED loop still leaves the
    // accumulator's start value (zeros / the identity) to be stored
    bool empty = false;
    for (int l = 0; l < n.nloops; l++) if (n.extent[l] == 0 && out.stride[l] != 0) empty = true;
    for (int i = 0; i < nk; i++) { p.ext[i] = n.extent[kept[i]]; p.total *= p.ext[i]; }
    for (int i = 0; i < nc; i++) p.ext[nk + i] = n.extent[con[i]];
    auto fill_arr = [&](int slot, const Operand& a) {
      p.base[slot] = a.base; p.off[slot] = a.offset; p.lk[slot] = dtype_load_kind(a.dtype); p.cls[slot] = g_dt[a.dtype].cls;
      for (int i = 0; i < nk; i++) p.str[slot][i] = a.stride[kept[i]];
      for (int i = 0; i < nc; i++) p.str[slot][nk + i] = a.stride[con[i]];
    };
    fill_arr(0, out);
    for (int i = 0; i < n.n_in; i++) fill_arr(1 + i, n.in[i]);
    p.st = dtype_store_spec(out.dtype); p.out_cls = g_dt[out.dtype].cls; p.out_bits = g_dt[out.dtype].value_bits; p.out_signed = g_dt[out.dtype].is_signed;
    p.fresh = n.fresh ? 1 : 0; p.has_identity = n.has_identity ? 1 : 0; p.ident_i = n.ident_i; p.ident_f = n.ident_f;
    p.n_nodes = e.n; p.fault = c.fault;
    for (int k = 0; k < e.n; k++) if (op_may_fault(e.node[k].op)) c.fault_possible = true;
    for (int k = 0; k < e.n; k++) {
      const XNode& x = e.node[k]; GNode& g = p.node[k];
      g.op = (short)x.op; g.a = (short)x.a; g.b = (short)x.b; g.cls = (short)x.cls;
      if (x.op == NCL_X_CONST_INT) g.c = x.ival;
      else if (x.op == NCL_X_CONST_F32 || x.op == NCL_X_CONST_F64) { double d = x.fval; memcpy(&g.c, &d, 8); }
      if (x.op == NCL_X_INPUT) g.a = (short)x.a;        // array slot: inputs are 1-based == slot index
    }
    int rc = NCL_OK;
    if (!empty && p.total > 0) {
      long long blocks = (p.total + 127) / 128;
      if (blocks > 0x7fffffffLL) { delete pp; return set_error(NCL_ERR_UNSUPPORTED, "generic nest: output too large"); }
      generic_nest_kernel<<<(unsigned)blocks, 128, 0, c.stream>>>(p);
      note_launch("generic_nest");
      rc = cuda_fail(cudaGetLastError(), "generic_nest_kernel");
    }
    delete pp;
    if (rc != NCL_OK) return rc;
  }
  return NCL_OK;
}
