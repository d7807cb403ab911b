// k_ew_widen.cu -- K1/K2 fast family: a NARROW integer array (8/16/32-bit slots) combined with a 64-bit integer
// operand into a 64-bit result:  (numcl:* image-ub8 weights-fixnum), (numcl:+ counts-ub16 offsets) ...
//
// numcl narrows integer arrays to the smallest element type (asarray / arange: src/3array.lisp:62-118), while
// products and sums with fixnum data come back as fixnum (src/1type.lisp:575-621), so this widening shape is
// the common integer broadcast (config C5a).  It is write-dominated -- 1..4 input bytes per 8 output bytes --
// and on the generic engine (k_elementwise.cuh) it was instruction-bound: 241..276 warp instructions per
// 256-element item (run-time element-kind dispatch, generic 64x64-bit products, guarded chunks) against
// ~85 here, 0.73 of the HBM peak against 0.84..0.90 for a specialised kernel (tools/micro/c5a_ceiling.cu).
// Same layout as the engine's flat mode: persistent warps, item = 128 chunks of 16 output bytes, lane l
// owns chunks l, l+32, l+64, l+96, a row-vector operand whose period divides the item is read once per warp.
// Same arithmetic: exact in 64 bits, then (value & mask) << shift = %coerce's wrap + SBCL's fixnum tag.
#include "ncl_device.cuh"

enum { WB_FLAT = 0, WB_HOISTED = 1, WB_STEPPED = 2, WB_SCALAR = 3 };
enum { WOP_ADD = 0, WOP_SUB = 1, WOP_MUL = 2, WOP_MAX = 3, WOP_MIN = 4 };

struct WidenParams {
  const char* a;            // narrow operand, first element (bytes)
  const char* b;            // 64-bit operand, first element
  char* out;
  unsigned nchunks;         // 16-byte output chunks (2 elements each), < 2^31
  int bmode; unsigned period, pmul, pshr;      // chunks; q mod period = q - (umulhi(q, pmul) >> pshr) * period
  int btag;                 // 1: b holds tagged fixnums (value << 1)
  int swap;                 // 1: result = op(b, a)   (a is always the narrow operand here)
  unsigned long long mask; int shift;          // store: (value & mask) << shift
};

template <int BITS, bool SGN> __device__ __forceinline__ void widen_load(const char* p, long long& x0, long long& x1) {
  // two consecutive BITS-wide elements
  if constexpr (BITS == 8) {
    unsigned v = __ldg((const unsigned short*)p);
    if constexpr (SGN) { x0 = (long long)(signed char)(v & 0xff); x1 = (long long)(signed char)(v >> 8); }
    else { x0 = (long long)(v & 0xff); x1 = (long long)(v >> 8); }
  } else if constexpr (BITS == 16) {
    unsigned v = __ldg((const unsigned int*)p);
    if constexpr (SGN) { x0 = (long long)(short)(v & 0xffff); x1 = (long long)(short)(v >> 16); }
    else { x0 = (long long)(v & 0xffff); x1 = (long long)(v >> 16); }
  } else {
    uint2 v = __ldg((const uint2*)p);
    if constexpr (SGN) { x0 = (long long)(int)v.x; x1 = (long long)(int)v.y; }
    else { x0 = (long long)v.x; x1 = (long long)v.y; }
  }
}
__device__ __forceinline__ void wide_load(const char* p, int tag, long long& y0, long long& y1) {
  const uint4 v = __ldg((const uint4*)p);
  y0 = (long long)(((unsigned long long)v.y << 32) | v.x) >> tag;       // arithmetic shift: untag
  y1 = (long long)(((unsigned long long)v.w << 32) | v.z) >> tag;
}
template <int OP> __device__ __forceinline__ long long widen_op(long long a, long long b) {
  typedef Ops<long long> O;
  if constexpr (OP == WOP_ADD) return O::add(a, b);
  else if constexpr (OP == WOP_SUB) return O::sub(a, b);
  else if constexpr (OP == WOP_MUL) return O::mul(a, b);
  else if constexpr (OP == WOP_MAX) return O::mx(a, b);
  else return O::mn(a, b);
}

template <int BITS, bool SGN, int OP>
__global__ void __launch_bounds__(256) ew_widen_kernel(const __grid_constant__ WidenParams p) {
  constexpr int U = 4;
  constexpr unsigned SEG = 32 * U;
  const unsigned lane = threadIdx.x & 31;
  const unsigned nwarps = gridDim.x * 8, warp0 = blockIdx.x * 8 + (threadIdx.x >> 5);
  const unsigned nitems = (p.nchunks + SEG - 1) / SEG;
  long long y[U][2];
  pdl_trigger();                                            // programmatic dependent launch: see ew_kernel
  bool waited = false;
  if (p.bmode == WB_HOISTED) {
#pragma unroll
    for (int u = 0; u < U; u++) {
      unsigned q = lane + 32 * u;
      q = p.pmul ? q - (__umulhi(q, p.pmul) >> p.pshr) * p.period : 0u;        // q mod period (period 1: always chunk 0)
      wide_load(p.b + (size_t)q * 16, p.btag, y[u][0], y[u][1]);
    }
  } else if (p.bmode == WB_SCALAR) {
    long long s = (long long)__ldg((const unsigned long long*)p.b) >> p.btag;
#pragma unroll
    for (int u = 0; u < U; u++) { y[u][0] = s; y[u][1] = s; }
  }
  for (unsigned item = warp0; item < nitems; item += nwarps) {
    const unsigned q0 = item * SEG + lane;
    long long x[U][2];
    const char* ap = p.a + (size_t)q0 * (BITS / 4);
#pragma unroll
    for (int u = 0; u < U; u++) if (q0 + 32 * u < p.nchunks) widen_load<BITS, SGN>(ap + (size_t)u * 32 * (BITS / 4), x[u][0], x[u][1]);
    if (p.bmode == WB_FLAT) {
#pragma unroll
      for (int u = 0; u < U; u++) if (q0 + 32 * u < p.nchunks) wide_load(p.b + (size_t)(q0 + 32 * u) * 16, p.btag, y[u][0], y[u][1]);
    } else if (p.bmode == WB_STEPPED) {                      // the item lies inside one period
      const unsigned s0 = item * SEG;
      const unsigned b0 = s0 - (__umulhi(s0, p.pmul) >> p.pshr) * p.period + lane;
#pragma unroll
      for (int u = 0; u < U; u++) if (q0 + 32 * u < p.nchunks) wide_load(p.b + (size_t)(b0 + 32 * u) * 16, p.btag, y[u][0], y[u][1]);
    }
    char* op = p.out + (size_t)q0 * 16;
    if (!waited) { pdl_wait(); waited = true; }
#pragma unroll
    for (int u = 0; u < U; u++) if (q0 + 32 * u < p.nchunks) {
      long long r0 = p.swap ? widen_op<OP>(y[u][0], x[u][0]) : widen_op<OP>(x[u][0], y[u][0]);
      long long r1 = p.swap ? widen_op<OP>(y[u][1], x[u][1]) : widen_op<OP>(x[u][1], y[u][1]);
      const unsigned long long w0 = ((unsigned long long)r0 & p.mask) << p.shift, w1 = ((unsigned long long)r1 & p.mask) << p.shift;
      *(uint4*)(op + (size_t)u * 32 * 16) = make_uint4((unsigned)w0, (unsigned)(w0 >> 32), (unsigned)w1, (unsigned)(w1 >> 32));
    }
  }
}

typedef void (*WidenKernel)(const WidenParams);
template <int BITS, bool SGN> static WidenKernel pick_op(int op) {
  switch (op) {
    case WOP_ADD: return ew_widen_kernel<BITS, SGN, WOP_ADD>;
    case WOP_SUB: return ew_widen_kernel<BITS, SGN, WOP_SUB>;
    case WOP_MUL: return ew_widen_kernel<BITS, SGN, WOP_MUL>;
    case WOP_MAX: return ew_widen_kernel<BITS, SGN, WOP_MAX>;
    default: return ew_widen_kernel<BITS, SGN, WOP_MIN>;
  }
}

// a_lk: LK_U8 .. LK_S32 of the narrow operand.  Returns NCL_ERR_UNSUPPORTED when the kinds are not covered.
int ew_widen_launch(int a_lk, int op, const char* a, const char* b, char* out, unsigned nchunks, int bmode, unsigned period, int btag, int swap,
                    unsigned long long mask, int shift, const Operand* oa, const Operand* ob, const Operand* oout) {
  Context& c = ctx();
  WidenKernel k = nullptr;
  switch (a_lk) {
    case LK_U8:  k = pick_op<8, false>(op); break;
    case LK_S8:  k = pick_op<8, true>(op); break;
    case LK_U16: k = pick_op<16, false>(op); break;
    case LK_S16: k = pick_op<16, true>(op); break;
    case LK_U32: k = pick_op<32, false>(op); break;
    case LK_S32: k = pick_op<32, true>(op); break;
    default: return NCL_ERR_UNSUPPORTED;
  }
  WidenParams p; memset(&p, 0, sizeof p);
  p.a = a; p.b = b; p.out = out; p.nchunks = nchunks; p.bmode = bmode; p.period = period; p.btag = btag; p.swap = swap; p.mask = mask; p.shift = shift;
  if (period > 1) { unsigned s2 = 0; while ((1ull << s2) < period) s2++; p.pmul = (unsigned)(((1ull << (31 + s2)) + period - 1) / period); p.pshr = s2 - 1; }
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)k, 256, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
  unsigned long long nitems = ((unsigned long long)nchunks + 127) / 128, blocks = (nitems + 7) / 8;
  // eight waves' worth of CTAs measured best on C5a (29.9 us with one resident wave, 25.8 us here): the kernel is
  // light enough that re-reading a hoisted row vector per warp costs less than the better tail balance gains
  unsigned long long cap = (unsigned long long)c.sm_count * per_sm * 8;
  if (blocks > cap) blocks = cap;
  if (blocks == 0) return NCL_OK;
  PdlRange reads[2] = {pdl_range(*oa), pdl_range(*ob)};
  const bool early = pdl_eligible(reads, 2);
  void* args[1] = {(void*)&p};
  NCL_TRY(pdl_launch((const void*)k, dim3((unsigned)blocks), dim3(256), args, c.stream, early));
  note_launch("bcast_widen");
  PdlRange w = pdl_range(*oout); pdl_commit(early, &w, 1);
  return NCL_OK;
}
