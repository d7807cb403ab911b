// k_elementwise.cu -- K1/K2/K6: broadcast elementwise and same-shape map kernels.
//
// Replaces the inner loops of broadcast-core (src/4numeric.lisp:228-293), map-array-into
// (:78-113) and einsum nests without a contracted index (common-lisp.lisp:108-224, e.g.
// `(- - -> (+ $1 $2) -> -)`, `(i j -> ij)`, `(i -> (f $1) -> i)`).
//
// HBM-bound: no reuse, so no shared memory.  Layout of the work:
//   * the nest is collapsed on the host to <= 6 dims with the output's contiguous dim innermost;
//   * one thread handles a CHUNK of E consecutive output elements = 16 bytes of output
//     (E = 16 / sizeof(out slot); E = 32 when the output is a packed bit array), so stores are
//     full 128-bit vectors and every warp store instruction writes 512 contiguous bytes;
//   * inputs whose innermost stride is 1 are read with the matching vector load, inputs whose
//     innermost stride is 0 (numpy broadcasting: `b` of (numcl:+ a b), scalars) are read once and
//     splat -- stride-0 indexing, not a materialised copy;
//   * each thread issues U independent chunks' loads before any math for memory-level parallelism;
//     persistent warps walk work items of 32*U chunks (one resident wave of CTAs, or four waves when
//     nothing is hoisted out of the item loop -- see try_elementwise);
//   * the widening integer shape (narrow int array x 64-bit operand -> 64-bit) has its own
//     kind-specialised family in k_ew_widen.cu.
// Arithmetic follows ncl_device.cuh (IEEE rn, no FMA contraction, CL max/min, exact integers).
#pragma once
#include "ncl_fn.cuh"

#define EW_MAXD 6
#define EW_THREADS 256

struct EwParams {
  int nd, n_in;
  unsigned ext[EW_MAXD];               // ext[0] counts chunks, ext[d>0] elements
  unsigned mul[EW_MAXD], shr[EW_MAXD]; // q / ext[d] = umulhi(q, mul) >> shr for q < 2^31 (mul == 0: ext is 1)
  int st[4][EW_MAXD];                  // [0] = output, [1..3] = inputs; elements (|stride| < 2^31)
  char* base[4]; long long off[4];
  int lk[3]; StoreSpec sk;
  unsigned period[3], pmul[3], pshr[3]; // flat mode: input k repeats every period[k] chunks (0 = not periodic)
  unsigned segs_per_row, seg_mul, seg_shr;   // a row of ext[0] chunks is cut into warp segments of 32*U chunks
  unsigned nitems;                     // rows * segs_per_row, < 2^31
  // element-granular flat mode: the output is one contiguous run over a (rows x eper) nest whose row length is not a
  // multiple of the chunk; operand k with gather[k] is addressed per element at (f / eper) * gst[k] + (f % eper) * st[1+k][0]
  // (a row vector: gst 0, st 1; one value per row: gst s, st 0), the others stream as flat vectors
  unsigned eper, epmul, epshr; int gather[3]; int gst[3];
  int rt_op; unsigned int* fault;       // runtime-selected function (FnRt*: an ncl_op) and the sticky DOMAIN flag it may raise
  // bit k: input k is a row vector whose piece is the same in EVERY item of a warp -- the host made the number of warps a
  // multiple of the number of segments the vector spans, so item mod span = warp mod span -- and is loaded once, before the loop
  unsigned whoist;
};
__device__ __forceinline__ unsigned ew_div(unsigned q, unsigned mul, unsigned shr) { return mul ? (__umulhi(q, mul) >> shr) : q; }

enum Shape { S_NONE = 0, S_ADD, S_SUB, S_MUL, S_DIV, S_MAX, S_MIN, S_ADD0, S_MUL0, S_IDIV, S_EQ, S_NE, S_LE, S_GE, S_LT, S_GT,
             S_AND, S_IOR, S_XOR, S_COPY, S_ADD0U, S_NEG, S_ABS, S_SQUARE, S_NOT, S_SQRT, S_EXP, S_LOG, S_SIN, S_COS, S_TAN, S_TANH,
             S_ADDMUL, S_RT1, S_RT2 };

// ---- functors ------------------------------------------------------------------------------------------------
template <class T> struct FnAdd { typedef T R; static __device__ __forceinline__ T f(T a, T b) { return Ops<T>::add(a, b); } };
template <class T> struct FnSub { typedef T R; static __device__ __forceinline__ T f(T a, T b) { return Ops<T>::sub(a, b); } };
template <class T> struct FnMul { typedef T R; static __device__ __forceinline__ T f(T a, T b) { return Ops<T>::mul(a, b); } };
template <class T> struct FnDiv { typedef T R; static __device__ __forceinline__ T f(T a, T b) { return Ops<T>::div(a, b); } };
template <class T> struct FnMax { typedef T R; static __device__ __forceinline__ T f(T a, T b) { return Ops<T>::mx(a, b); } };
template <class T> struct FnMin { typedef T R; static __device__ __forceinline__ T f(T a, T b) { return Ops<T>::mn(a, b); } };
// (+ (+ 0 a) b): the identity-seeded fold of numcl:+ (src/4numeric.lisp:528); 0.0 + -0.0 = +0.0
template <class T> struct FnAdd0 { typedef T R; static __device__ __forceinline__ T f(T a, T b) { return Ops<T>::add(Ops<T>::add(Ops<T>::zero(), a), b); } };
// (+ 0 (* a b)): default transform into a fresh output
template <class T> struct FnMul0 { typedef T R; static __device__ __forceinline__ T f(T a, T b) { return Ops<T>::add(Ops<T>::zero(), Ops<T>::mul(a, b)); } };
struct FnIDiv { typedef float R; static __device__ __forceinline__ float f(long long a, long long b) { return ratio_to_f32(a, b); } };
// comparisons produce 0 / 1 for a packed bit array: 32-bit results (32 of them per thread) keep the register count down
template <class T> struct FnEq { typedef int R; static __device__ __forceinline__ int f(T a, T b) { return a == b; } };
template <class T> struct FnNe { typedef int R; static __device__ __forceinline__ int f(T a, T b) { return a != b; } };
template <class T> struct FnLe { typedef int R; static __device__ __forceinline__ int f(T a, T b) { return a <= b; } };
template <class T> struct FnGe { typedef int R; static __device__ __forceinline__ int f(T a, T b) { return a >= b; } };
template <class T> struct FnLt { typedef int R; static __device__ __forceinline__ int f(T a, T b) { return a < b; } };
template <class T> struct FnGt { typedef int R; static __device__ __forceinline__ int f(T a, T b) { return a > b; } };
template <class T> struct FnAndT { typedef T R; static __device__ __forceinline__ T f(T a, T b) { return a & b; } };
template <class T> struct FnIorT { typedef T R; static __device__ __forceinline__ T f(T a, T b) { return a | b; } };
template <class T> struct FnXorT { typedef T R; static __device__ __forceinline__ T f(T a, T b) { return a ^ b; } };
typedef FnAndT<long long> FnAnd; typedef FnIorT<long long> FnIor; typedef FnXorT<long long> FnXor;
// unary
template <class T> struct FnCopy { typedef T R; static __device__ __forceinline__ T f(T a) { return a; } };
template <class T> struct FnAdd0U { typedef T R; static __device__ __forceinline__ T f(T a) { return Ops<T>::add(Ops<T>::zero(), a); } };
template <class T> struct FnNeg { typedef T R; static __device__ __forceinline__ T f(T a) { return Ops<T>::neg(a); } };
template <class T> struct FnAbs { typedef T R; static __device__ __forceinline__ T f(T a) { return Ops<T>::abs_(a); } };
template <class T> struct FnSquare { typedef T R; static __device__ __forceinline__ T f(T a) { return Ops<T>::mul(a, a); } };
template <class T> struct FnNotT { typedef T R; static __device__ __forceinline__ T f(T a) { return ~a; } };
typedef FnNotT<long long> FnNot;
template <class T> struct FnSqrt;
template <> struct FnSqrt<float> { typedef float R; static __device__ __forceinline__ float f(float a) { return __fsqrt_rn(a); } };
template <> struct FnSqrt<double> { typedef double R; static __device__ __forceinline__ double f(double a) { return __dsqrt_rn(a); } };
#define LIBM_FN(NAME, FF, DF) template <class T> struct NAME; \
  template <> struct NAME<float> { typedef float R; static __device__ __forceinline__ float f(float a) { return FF(a); } }; \
  template <> struct NAME<double> { typedef double R; static __device__ __forceinline__ double f(double a) { return DF(a); } };
LIBM_FN(FnExp, expf, exp) LIBM_FN(FnLog, logf, log) LIBM_FN(FnSin, sinf, sin) LIBM_FN(FnCos, cosf, cos)
LIBM_FN(FnTan, tanf, tan) LIBM_FN(FnTanh, tanhf, tanh)
// ---- runtime-selected functions (ncl_fn.cuh): the rest of numcl's scalar-function surface (src/4numeric.lisp:466-519,568-602).
// ONE kernel per (value class, E) covers a whole family; the function code is a kernel parameter and the switch on it is
// warp-uniform.  These are HBM-bound like everything here, so the few extra instructions per element are free, and the
// family costs 18 instantiations instead of ~150.
// calls g(integral_constant<OP>) for the one OP of the list that equals the runtime code: taken once per work item, so that
// inside the element loops the function is a compile-time constant (a per-element switch kept libm calls out of line:
// sqrt ran at 0.05 of the HBM peak)
template <int... OPS, class G> __device__ __forceinline__ void rt_switch(int op, G g) {
  ((op == OPS ? (g(std::integral_constant<int, OPS>{}), 0) : 0), ...);
}
struct FnRtBinI {          // integer x integer -> integer: bitwise family, floor ceiling round truncate mod rem
  typedef long long R; static constexpr bool RT = true;
  template <class G> static __device__ __forceinline__ void dispatch(int op, G g) {
    rt_switch<NCL_OP_LOGANDC1, NCL_OP_LOGANDC2, NCL_OP_LOGEQV, NCL_OP_LOGNAND, NCL_OP_LOGNOR, NCL_OP_LOGORC1, NCL_OP_LOGORC2,
              NCL_OP_FLOOR, NCL_OP_CEILING, NCL_OP_ROUND, NCL_OP_TRUNCATE, NCL_OP_MOD, NCL_OP_REM>(op, g);
  }
  template <int OP> static __device__ __forceinline__ long long f(unsigned int* flt, long long a, long long b) {
    if constexpr (OP >= NCL_OP_LOGAND && OP <= NCL_OP_LOGORC2) return int_bitwise(OP, a, b);
    else { long long r; const long long q = int_divide(OP == NCL_OP_REM ? NCL_OP_TRUNCATE : OP, a, b, &r, flt); return (OP == NCL_OP_MOD || OP == NCL_OP_REM) ? r : q; }
  }
};
struct FnRtUnI {           // integer -> integer: isqrt logcount integer-length
  typedef long long R; static constexpr bool RT = true;
  template <class G> static __device__ __forceinline__ void dispatch(int op, G g) { rt_switch<NCL_OP_ISQRT, NCL_OP_LOGCOUNT, NCL_OP_INTEGER_LENGTH>(op, g); }
  template <int OP> static __device__ __forceinline__ long long f(unsigned int* flt, long long a) {
    if constexpr (OP == NCL_OP_ISQRT) return int_isqrt(a, flt); else if constexpr (OP == NCL_OP_LOGCOUNT) return int_logcount(a); else return int_length(a);
  }
};
template <class T> struct FnRtBinF {   // float x float -> float: mod rem
  typedef T R; static constexpr bool RT = true;
  template <class G> static __device__ __forceinline__ void dispatch(int op, G g) { rt_switch<NCL_OP_MOD, NCL_OP_REM>(op, g); }
  template <int OP> static __device__ __forceinline__ T f(unsigned int* flt, T a, T b) {
    T r; float_divide<T>(OP == NCL_OP_REM ? NCL_OP_TRUNCATE : NCL_OP_MOD, a, b, &r); (void)flt; return r;
  }
};
template <class T> struct FnRtUnF {    // float -> float: sqrt log log2 and the inverse / hyperbolic functions, with the DOMAIN flag
  typedef T R; static constexpr bool RT = true;
  template <class G> static __device__ __forceinline__ void dispatch(int op, G g) {
    rt_switch<NCL_OP_SQRT, NCL_OP_LOG, NCL_OP_LOG2, NCL_OP_ASIN, NCL_OP_ACOS, NCL_OP_ATAN, NCL_OP_SINH, NCL_OP_COSH, NCL_OP_ASINH, NCL_OP_ACOSH, NCL_OP_ATANH>(op, g);
  }
  template <int OP> static __device__ __forceinline__ T f(unsigned int* flt, T a) { return float_unary<T>(OP, a, flt); }
};
template <class F, class = void> struct fn_is_rt : std::false_type {};
template <class F> struct fn_is_rt<F, std::void_t<decltype(F::RT)>> : std::true_type {};

// ternary: a + b*c, each op rounded on its own (the einsum leaf with a live accumulator)
template <class T> struct FnAddMul { typedef T R; static __device__ __forceinline__ T f(T a, T b, T c) { return Ops<T>::add(a, Ops<T>::mul(b, c)); } };

// ---- the kernel -------------------------------------------------------------------------------------------------
// All U chunks of one operand for one work item.  The element-kind switch is taken ONCE per operand
// per item (it is warp-uniform); inside a case the U loads are unrolled with constant offsets.  An
// operand that is broadcast along the innermost loop (stride 0) is read once and splat.
template <class T, int E, int U, int BITS, bool SIGNED, bool TAGGED>
__device__ __forceinline__ void load_op_int(const char* p, unsigned q0, unsigned ext0, T (&x)[U][E]) {
#pragma unroll
  for (int u = 0; u < U; u++) if (q0 + u * 32 < ext0) load_int_chunk<T, E, BITS, SIGNED, TAGGED>(p + (size_t)u * 32 * E * (BITS / 8), 0, x[u]);
}
// periodic operand (numpy row-vector broadcast in flat mode): chunk q reads chunk q mod period
template <class T, int E, int U>
__device__ __forceinline__ void load_operand_periodic(const char* base, long long elem0, int lk, unsigned q0, unsigned ext0,
                                                      unsigned period, unsigned pmul, unsigned pshr, T (&x)[U][E]) {
#pragma unroll
  for (int u = 0; u < U; u++) {
    unsigned q = q0 + u * 32;
    if (q < ext0) { unsigned qm = q - ew_div(q, pmul, pshr) * period; load_elems<T, E>(base, elem0 + (long long)qm * E, lk, true, x[u]); }
  }
}
// operand addressed per element in element-granular flat mode: chunk q covers flat elements q*E .. q*E+E-1 of a
// (rows x P) nest; (row, col) are decoded once per chunk and stepped with a wrap
template <class T, int E, int U>
__device__ __forceinline__ void load_operand_gather(const char* base, long long elem0, int lk, int s_in, int s_out, unsigned q0, unsigned ext0,
                                                    unsigned P, unsigned mul, unsigned shr, T (&x)[U][E]) {
  dispatch_lk(lk, [&](auto K) {
#pragma unroll
    for (int u = 0; u < U; u++) {
      const unsigned q = q0 + u * 32;
      if (q < ext0) {
        const unsigned f = q * E, r = ew_div(f, mul, shr);
        unsigned e = f - r * P;
        long long idx = elem0 + (long long)r * s_out + (long long)e * s_in;
        const long long wrap = (long long)s_out - (long long)P * s_in;
#pragma unroll
        for (int i = 0; i < E; i++) {
          x[u][i] = load_scalar_k<T, decltype(K)::value>(base, idx);
          e++; idx += s_in;
          if (e >= P) { e = 0; idx += wrap; }
        }
      }
    }
  });
}
template <class T, int E, int U>
__device__ __forceinline__ void load_operand(const char* base, long long elem0, int st0, int lk, unsigned q0, unsigned ext0, T (&x)[U][E]) {
  if (st0 == 0) {
    T s = load_scalar<T>(base, elem0, lk);
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
      for (int i = 0; i < E; i++) x[u][i] = s;
    return;
  }
  if constexpr (E == 1) {
#pragma unroll
    for (int u = 0; u < U; u++) if (q0 + u * 32 < ext0) x[u][0] = load_scalar<T>(base, elem0 + (long long)(q0 + u * 32) * st0, lk);
  } else {
    const long long e0 = elem0 + (long long)q0 * E;          // st0 == 1 on the vector path
    switch (lk) {
      case LK_U8:  load_op_int<T, E, U, 8, false, false>(base + e0, q0, ext0, x); break;
      case LK_S8:  load_op_int<T, E, U, 8, true, false>(base + e0, q0, ext0, x); break;
      case LK_U16: load_op_int<T, E, U, 16, false, false>(base + e0 * 2, q0, ext0, x); break;
      case LK_S16: load_op_int<T, E, U, 16, true, false>(base + e0 * 2, q0, ext0, x); break;
      case LK_U32: load_op_int<T, E, U, 32, false, false>(base + e0 * 4, q0, ext0, x); break;
      case LK_S32: load_op_int<T, E, U, 32, true, false>(base + e0 * 4, q0, ext0, x); break;
      case LK_W64: load_op_int<T, E, U, 64, true, false>(base + e0 * 8, q0, ext0, x); break;
      case LK_T64: load_op_int<T, E, U, 64, true, true>(base + e0 * 8, q0, ext0, x); break;
      case LK_P1:
        if constexpr (E <= 32) {
          // every word of the item first (chunks past the end of the row re-read the row's first word instead of branching around the
          // load: with a branch per chunk the four chunks' loads went out one after the other, mask * array at 0.11 of the peak)
          unsigned raw[U];
#pragma unroll
          for (int u = 0; u < U; u++) raw[u] = load_bits_raw<E>(base, (q0 + u * 32 < ext0) ? e0 + (long long)u * 32 * E : elem0);
#pragma unroll
          for (int u = 0; u < U; u++) if (q0 + u * 32 < ext0) bits_to_elems<T, E>(raw[u], x[u]);
        }
        break;
      case LK_P2:
        if constexpr (E <= 16) {
          unsigned raw[U];
#pragma unroll
          for (int u = 0; u < U; u++) raw[u] = load_bits_raw<E, 2>(base, (q0 + u * 32 < ext0) ? e0 + (long long)u * 32 * E : elem0);
#pragma unroll
          for (int u = 0; u < U; u++) if (q0 + u * 32 < ext0) bits_to_elems<T, E, 2>(raw[u], x[u]);
        }
        break;
      case LK_P4:
        if constexpr (E <= 8) {
          unsigned raw[U];
#pragma unroll
          for (int u = 0; u < U; u++) raw[u] = load_bits_raw<E, 4>(base, (q0 + u * 32 < ext0) ? e0 + (long long)u * 32 * E : elem0);
#pragma unroll
          for (int u = 0; u < U; u++) if (q0 + u * 32 < ext0) bits_to_elems<T, E, 4>(raw[u], x[u]);
        }
        break;
      case LK_F32: {
        const char* p = base + e0 * 4;
#pragma unroll
        for (int u = 0; u < U; u++) if (q0 + u * 32 < ext0) {
          Chunk<E * 4> c; ld_chunk(c, p + (size_t)u * 32 * E * 4);
#pragma unroll
          for (int i = 0; i < E; i++) x[u][i] = from_f32<T>(__uint_as_float(c.w[i]));
        }
        break; }
      default: {
        const char* p = base + e0 * 8;
#pragma unroll
        for (int u = 0; u < U; u++) if (q0 + u * 32 < ext0) {
          Chunk<E * 8> c; ld_chunk(c, p + (size_t)u * 32 * E * 8);
#pragma unroll
          for (int i = 0; i < E; i++) x[u][i] = from_f64<T>(__longlong_as_double((long long)chunk_get<64>(c, i)));
        }
        break; }
    }
  }
}

// all U chunks of the output for one item: the store-kind switch is taken once, offsets are constants
template <class TO, int E, int U, int BITS>
__device__ __forceinline__ void store_op_int(char* p, unsigned q0, unsigned ext0, const StoreSpec& sk, const TO (&y)[U][E]) {
#pragma unroll
  for (int u = 0; u < U; u++) if (q0 + u * 32 < ext0) {
    Chunk<E * BITS / 8> c;
#pragma unroll
    for (int i = 0; i < E; i++) chunk_put<BITS>(c, i, (((unsigned long long)y[u][i]) & sk.mask) << sk.shift);
    st_chunk(c, p + (size_t)u * 32 * E * (BITS / 8));
  }
}
template <class TO, int E, int U>
__device__ __forceinline__ void store_operand(char* base, long long elem0, unsigned q0, unsigned ext0, const StoreSpec& sk, const TO (&y)[U][E]) {
  if constexpr (E == 1 || E == 32) {
#pragma unroll
    for (int u = 0; u < U; u++) if (q0 + u * 32 < ext0) store_elems<TO, E>(base, elem0 + (long long)(q0 + u * 32) * E, sk, y[u]);
  } else if constexpr (!std::is_integral<TO>::value) {
#pragma unroll
    for (int u = 0; u < U; u++) if (q0 + u * 32 < ext0) store_elems<TO, E>(base, elem0 + (long long)(q0 + u * 32) * E, sk, y[u]);
  } else {
    const long long e0 = elem0 + (long long)q0 * E;
    switch (sk.kind) {
      case SK_P2: case SK_P4:                                   // ub2 / ub4 results: a chunk is one 32-bit word (store_elems)
#pragma unroll
        for (int u = 0; u < U; u++) if (q0 + u * 32 < ext0) store_elems<TO, E>(base, elem0 + (long long)(q0 + u * 32) * E, sk, y[u]);
        break;
      case SK_8:  store_op_int<TO, E, U, 8>(base + e0, q0, ext0, sk, y); break;
      case SK_16: store_op_int<TO, E, U, 16>(base + e0 * 2, q0, ext0, sk, y); break;
      case SK_32: store_op_int<TO, E, U, 32>(base + e0 * 4, q0, ext0, sk, y); break;
      default:    store_op_int<TO, E, U, 64>(base + e0 * 8, q0, ext0, sk, y); break;
    }
  }
}

// Persistent: one resident wave of CTAs; every WARP repeatedly takes a work item = one segment of
// 32*U chunks of one innermost row.  The (row, segment) -> offsets decode is warp-uniform and paid
// once per item; lane l then handles chunks l, l+32, ... of the segment, so each load / store
// instruction of the warp covers 32 consecutive chunks (512 contiguous bytes for 16-byte chunks).
// Flat mode (nd == 1 with periodic operands): a row-vector operand whose period divides the segment
// gives every lane the SAME elements in every item, so it is loaded once, before the loop.
template <class TI, class F, int NIN, int E, int U>
__global__ void __launch_bounds__(EW_THREADS) ew_kernel(const __grid_constant__ EwParams p) {
  typedef typename F::R TO;
  const int lane = threadIdx.x & 31;
  const unsigned nwarps = gridDim.x * (EW_THREADS / 32);
  constexpr unsigned SEG = 32 * U;
  TI x[NIN][U][E];
  unsigned int flt_local = 0;                       // DOMAIN faults of this thread (runtime-selected families), flushed once at the end
  // programmatic dependent launch (ncl_internal.h: PdlTracker): the next kernel of the stream may start while this one drains;
  // this one reads its first item while its predecessor drains and writes nothing before griddepcontrol.wait
  pdl_trigger();
  bool waited = false;
  bool hoisted[NIN], stepped[NIN];
#pragma unroll
  for (int k = 0; k < NIN; k++) {
    hoisted[k] = p.period[k] != 0 && (SEG % p.period[k]) == 0;        // item-invariant
    stepped[k] = p.period[k] != 0 && !hoisted[k] && (p.period[k] % SEG) == 0;   // no wrap inside an item
    if (hoisted[k]) load_operand_periodic<TI, E, U>(p.base[1 + k], p.off[1 + k], p.lk[k], (unsigned)lane, 0xffffffffu, p.period[k], p.pmul[k], p.pshr[k], x[k]);
  }
  // Row vectors longer than a segment (C1's 4096 floats, `outer`'s second operand): every item used to fetch its piece again
  // -- L2 traffic as large as the output, and under a write-dominated stream the vector's lines kept being evicted (ncu on
  // outer 16384^2: 433 MB of DRAM READS for two 64 KB inputs, 0.75 of the peak).  With the warp count a multiple of the span
  // the piece is warp-invariant.
  {
    const unsigned warp0 = blockIdx.x * (EW_THREADS / 32) + (threadIdx.x >> 5);
#pragma unroll
    for (int k = 0; k < NIN; k++) {
      if ((p.whoist >> k) & 1u) {
        hoisted[k] = true;                                                         // (one flag for both kinds: nothing more is live in the loop)
        if (p.nd == 1) {                                                           // flat mode, period a multiple of the segment
          const unsigned s0 = warp0 * SEG, base_chunk = s0 - ew_div(s0, p.pmul[k], p.pshr[k]) * p.period[k];
          load_operand<TI, E, U>(p.base[1 + k], p.off[1 + k] + (long long)base_chunk * E, 1, p.lk[k], (unsigned)lane, 0xffffffffu, x[k]);
        } else {                                                                   // rows x segments, the vector moves with the segment only
          const unsigned seg0 = warp0 - ew_div(warp0, p.seg_mul, p.seg_shr) * p.segs_per_row;
          load_operand<TI, E, U>(p.base[1 + k], p.off[1 + k], 1, p.lk[k], seg0 * SEG + lane, p.ext[0], x[k]);
        }
      }
    }
  }
  for (unsigned item = blockIdx.x * (EW_THREADS / 32) + (threadIdx.x >> 5); item < p.nitems; item += nwarps) {
    unsigned row = ew_div(item, p.seg_mul, p.seg_shr);
    unsigned seg = item - row * p.segs_per_row;
    long long rowoff[4];
#pragma unroll
    for (int k = 0; k < 4; k++) rowoff[k] = p.off[k];
    {
      unsigned r = row;
      for (int d = 1; d < p.nd; d++) {
        unsigned r2 = ew_div(r, p.mul[d], p.shr[d]);
        int c = (int)(r - r2 * p.ext[d]); r = r2;
#pragma unroll
        for (int k = 0; k < 4; k++) rowoff[k] += (long long)c * p.st[k][d];
      }
    }
    const unsigned q0 = seg * SEG + lane;
#pragma unroll
    for (int k = 0; k < NIN; k++) {
      if (hoisted[k]) continue;
      if (stepped[k]) {
        unsigned s0 = seg * SEG;
        unsigned base_chunk = s0 - ew_div(s0, p.pmul[k], p.pshr[k]) * p.period[k];
        load_operand<TI, E, U>(p.base[1 + k], rowoff[1 + k] + ((long long)base_chunk - (long long)s0) * E, 1, p.lk[k], q0, p.ext[0], x[k]);
      } else if (p.gather[k]) load_operand_gather<TI, E, U>(p.base[1 + k], p.off[1 + k], p.lk[k], p.st[1 + k][0], p.gst[k], q0, p.ext[0], p.eper, p.epmul, p.epshr, x[k]);
      else if (p.period[k]) load_operand_periodic<TI, E, U>(p.base[1 + k], rowoff[1 + k], p.lk[k], q0, p.ext[0], p.period[k], p.pmul[k], p.pshr[k], x[k]);
      else load_operand<TI, E, U>(p.base[1 + k], rowoff[1 + k], p.st[1 + k][0], p.lk[k], q0, p.ext[0], x[k]);
    }
    if (p.nd == 1) {
      // flat mode: pull the streaming operands of this warp's NEXT item into L1 now, so that its loads
      // pay L1 latency instead of DRAM latency (one prefetch instruction per 512-byte warp row; with the
      // L2-only prefetch 31 % of the stall samples still sat on the first use of the loaded registers)
      const unsigned nq0 = q0 + nwarps * SEG;
#pragma unroll
      for (int k = 0; k < NIN; k++) {
        if (p.period[k] == 0 && p.st[1 + k][0] == 1 && !p.gather[k] && lk_bits(p.lk[k]) >= 8) {      // (packed operands: 64 bytes per item, not worth a prefetch)
          const int bytes = lk_bits(p.lk[k]) >> 3;
          const char* a = p.base[1 + k] + (p.off[1 + k] + (long long)nq0 * E) * bytes;
#pragma unroll
          for (int u = 0; u < U; u++)
            if (nq0 + u * 32 < p.ext[0]) asm volatile("prefetch.global.L1 [%0];" ::"l"(a + (size_t)u * 32 * E * bytes));
        }
      }
    }
    TO y[U][E];
    if constexpr (fn_is_rt<F>::value) {
      F::dispatch(p.rt_op, [&](auto OPC) {
        constexpr int OP = decltype(OPC)::value;
#pragma unroll
        for (int u = 0; u < U; u++) {
          // chunks past the end of the row were not loaded: their lanes hold junk, which must not raise the DOMAIN flag
          const bool valid = q0 + u * 32 < p.ext[0];
#pragma unroll
          for (int i = 0; i < E; i++) {
            if (!valid) y[u][i] = (TO)0;
            else if constexpr (NIN == 1) y[u][i] = F::template f<OP>(&flt_local, x[0][u][i]);
            else y[u][i] = F::template f<OP>(&flt_local, x[0][u][i], x[1][u][i]);
          }
        }
      });
    } else {
#pragma unroll
      for (int u = 0; u < U; u++)
#pragma unroll
        for (int i = 0; i < E; i++) {
          if constexpr (NIN == 1) y[u][i] = F::f(x[0][u][i]);
          else if constexpr (NIN == 2) y[u][i] = F::f(x[0][u][i], x[1][u][i]);
          else y[u][i] = F::f(x[0][u][i], x[1][u][i], x[2][u][i]);
        }
    }
    if (!waited) { pdl_wait(); waited = true; }
    store_operand<TO, E, U>(p.base[0], rowoff[0], q0, p.ext[0], p.sk, y);
  }
  if constexpr (fn_is_rt<F>::value) ncl_fault_flush(p.fault, flt_local);
}

typedef void (*EwKernel)(const EwParams);
// chunks per thread and item.  The integer class computes in 64-bit lanes: 4 chunks of 8 or 16 narrow elements took 244
// registers / spilled, so wide chunks get fewer of them.
template <class TI, int E> constexpr int ew_u() {
  return E == 32 ? 1 : !std::is_integral<TI>::value ? 4 : sizeof(TI) == 4 ? (E >= 8 ? 2 : 4) : (E >= 16 ? 1 : E >= 4 ? 2 : 4);
}
// lanes: 0 = float / double, 1 = 64-bit integer lanes, 2 = 32-bit integer lanes
inline int ew_u_host(int lanes, int E) { return E == 32 ? 1 : lanes == 0 ? 4 : lanes == 2 ? (E >= 8 ? 2 : 4) : (E >= 16 ? 1 : E >= 4 ? 2 : 4); }
template <class TI, class F, int NIN, int E> static EwKernel kern() { return (EwKernel)ew_kernel<TI, F, NIN, E, ew_u<TI, E>()>; }

template <class T, int E> static EwKernel pick_arith(int shape) {     // T-class in, T-class out
  switch (shape) {
    case S_ADD: return kern<T, FnAdd<T>, 2, E>(); case S_SUB: return kern<T, FnSub<T>, 2, E>(); case S_MUL: return kern<T, FnMul<T>, 2, E>();
    case S_MAX: return kern<T, FnMax<T>, 2, E>(); case S_MIN: return kern<T, FnMin<T>, 2, E>(); case S_ADD0: return kern<T, FnAdd0<T>, 2, E>();
    case S_MUL0: return kern<T, FnMul0<T>, 2, E>(); case S_COPY: return kern<T, FnCopy<T>, 1, E>(); case S_ADD0U: return kern<T, FnAdd0U<T>, 1, E>();
    case S_NEG: return kern<T, FnNeg<T>, 1, E>(); case S_ABS: return kern<T, FnAbs<T>, 1, E>(); case S_SQUARE: return kern<T, FnSquare<T>, 1, E>();
    case S_ADDMUL: return kern<T, FnAddMul<T>, 3, E>();
    default: return nullptr;
  }
}
template <class T, int E> static EwKernel pick_float(int shape) {
  if (EwKernel k = pick_arith<T, E>(shape)) return k;
  switch (shape) {
    case S_DIV: return kern<T, FnDiv<T>, 2, E>(); case S_EXP: return kern<T, FnExp<T>, 1, E>();
    case S_SIN: return kern<T, FnSin<T>, 1, E>(); case S_COS: return kern<T, FnCos<T>, 1, E>();
    case S_TAN: return kern<T, FnTan<T>, 1, E>(); case S_TANH: return kern<T, FnTanh<T>, 1, E>();
    default: return nullptr;
  }
}
template <int E> static EwKernel pick_int(int shape) {
  if (EwKernel k = pick_arith<long long, E>(shape)) return k;
  switch (shape) {
    case S_AND: return kern<long long, FnAnd, 2, E>(); case S_IOR: return kern<long long, FnIor, 2, E>(); case S_XOR: return kern<long long, FnXor, 2, E>();
    case S_NOT: return kern<long long, FnNot, 1, E>();
    default: return nullptr;
  }
}
template <int E> static EwKernel pick_int32(int shape) {
  if (EwKernel k = pick_arith<int, E>(shape)) return k;
  switch (shape) {
    case S_AND: return kern<int, FnAndT<int>, 2, E>(); case S_IOR: return kern<int, FnIorT<int>, 2, E>(); case S_XOR: return kern<int, FnXorT<int>, 2, E>();
    case S_NOT: return kern<int, FnNotT<int>, 1, E>();
    default: return nullptr;
  }
}
template <class T, int E> static EwKernel pick_cmp(int shape) {
  switch (shape) {
    case S_EQ: return kern<T, FnEq<T>, 2, E>(); case S_NE: return kern<T, FnNe<T>, 2, E>(); case S_LE: return kern<T, FnLe<T>, 2, E>();
    case S_GE: return kern<T, FnGe<T>, 2, E>(); case S_LT: return kern<T, FnLt<T>, 2, E>(); case S_GT: return kern<T, FnGt<T>, 2, E>();
    default: return nullptr;
  }
}


// one translation unit per value class keeps the build parallel
EwKernel ew_pick_f32(int shape, int E);
EwKernel ew_pick_f64(int shape, int E);
EwKernel ew_pick_int(int shape, int E);
EwKernel ew_pick_i32(int shape, int E);      // 32-bit lanes, E = 4 / 8 / 16
EwKernel ew_pick_cmp(int shape, int cls);
EwKernel ew_pick_idiv(int E);
EwKernel ew_pick_cmppack(int lanes, int EI, int* U);      // k_ew_cmppack.cu
EwKernel ew_pick_cmppack_swar(int bits, int* U);
enum { RT_BIN_I = 0, RT_UN_I, RT_BIN_F, RT_UN_F, RT_BIN_D, RT_UN_D };
EwKernel ew_pick_rt(int family, int E);      // k_ew_rt.cu
