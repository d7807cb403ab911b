// k_ew_cmppack.cu -- comparisons into packed bit arrays (numcl:= /= <= >= < >, src/4numeric.lisp:609-655) at stream speed.
//
// Every numcl comparison returns a BIT array, so the output is 1/32 .. 1/256 of the traffic and the kernel is a pure read
// stream.  The general engine (k_elementwise.cuh) is organised around the OUTPUT chunk: for a packed result a lane owned
// one 32-bit word = 32 input elements = 128 contiguous bytes of a single-float operand, so a warp-level load instruction
// touched 32 different 128-byte lines (measured 0.35-0.42 of the HBM peak, profiles/r01_cliff_hunt.txt).  Here the layout
// follows the INPUT: a lane loads 16-byte chunks (EI = 16 / widest slot = 2 / 4 / 8 / 16 elements), consecutive lanes take
// consecutive chunks (512 contiguous bytes per warp load, U chunks in flight per operand), each lane forms EI result bits and
// the 32 / EI lanes that share a result word combine them with a masked warp OR-reduction; one lane per word stores it.
// Rows must be whole result words (innermost extent a multiple of 32) -- the engine's own condition for packed outputs.
#include "k_elementwise.cuh"

template <class T> __device__ __forceinline__ unsigned cmp_bit(int op, T a, T b) {
  switch (op) {
    case NCL_OP_EQ: return a == b; case NCL_OP_NE: return a != b; case NCL_OP_LE: return a <= b;
    case NCL_OP_GE: return a >= b; case NCL_OP_LT: return a < b; default: return a > b;
  }
}

template <class T, int EI, int U>
__global__ void __launch_bounds__(EW_THREADS) cmp_pack_kernel(const __grid_constant__ EwParams p) {
  constexpr int G = 32 / EI;                       // lanes per result word
  constexpr unsigned SEG = 32 * U;
  const int lane = threadIdx.x & 31;
  const unsigned nwarps = gridDim.x * (EW_THREADS / 32);
  const int op = p.rt_op;
  for (unsigned item = blockIdx.x * (EW_THREADS / 32) + (threadIdx.x >> 5); item < p.nitems; item += nwarps) {
    const unsigned row = ew_div(item, p.seg_mul, p.seg_shr);
    const unsigned seg = item - row * p.segs_per_row;
    long long rowoff[3];
#pragma unroll
    for (int k = 0; k < 3; k++) rowoff[k] = p.off[k];
    {
      unsigned r = row;
      for (int d = 1; d < p.nd; d++) {
        const unsigned r2 = ew_div(r, p.mul[d], p.shr[d]);
        const int c = (int)(r - r2 * p.ext[d]); r = r2;
#pragma unroll
        for (int k = 0; k < 3; k++) rowoff[k] += (long long)c * p.st[k][d];
      }
    }
    const unsigned q0 = seg * SEG + lane;
    T x[2][U][EI];
#pragma unroll
    for (int k = 0; k < 2; k++) load_operand<T, EI, U>(p.base[1 + k], rowoff[1 + k], p.st[1 + k][0], p.lk[k], q0, p.ext[0], x[k]);
#pragma unroll
    for (int u = 0; u < U; u++) {
      const unsigned q = q0 + u * 32;
      unsigned m = 0;
      if (q < p.ext[0]) {
#pragma unroll
        for (int i = 0; i < EI; i++) m |= cmp_bit<T>(op, x[0][u][i], x[1][u][i]) << i;
      }
      unsigned w = m << (EI * (lane & (G - 1)));
      // OR over the G lanes of one word with full-warp shuffles (a masked redux per group serialises the groups); the G lanes
      // of one word are all in range or all out of it (rows are whole words)
#pragma unroll
      for (int o = 1; o < G; o <<= 1) w |= __shfl_xor_sync(0xffffffffu, w, o);
      if ((lane & (G - 1)) == 0 && q < p.ext[0])
        *(unsigned*)(p.base[0] + ((rowoff[0] + (long long)q * EI) >> 3)) = w;
    }
  }
}

// ---- 8- and 16-bit operands of one kind: SIMD within a register -------------------------------------------------------------
// Unpacking every byte into a 32-bit lane costs ~6 instructions per element and made `ub8 < ub8` issue-bound (0.17 of the HBM
// peak with the kernel above).  Here a 32-bit word holds 4 (2) fields and the comparison is done on all of them at once with
// carry-isolated arithmetic: H marks the top bit of every field, (a | H) - (b & ~H) compares the low bits of each field without
// borrowing across fields, the top bits are compared separately, and the result masks are gathered to 4 (2) adjacent bits with
// one multiplication.  Signed fields are biased by H first.  ~2.5 instructions per element instead of 6.
template <int BITS> __device__ __forceinline__ unsigned swar_lt(unsigned a, unsigned b) {     // top bit of each field: a < b (unsigned fields)
  constexpr unsigned H = BITS == 8 ? 0x80808080u : 0x80008000u;
  const unsigned t = (a | H) - (b & ~H);                 // top bit set <=> low(a) >= low(b)
  return ((~a & b) | (~(a ^ b) & ~t)) & H;
}
template <int BITS> __device__ __forceinline__ unsigned swar_eq(unsigned a, unsigned b) {
  constexpr unsigned H = BITS == 8 ? 0x80808080u : 0x80008000u;
  const unsigned x = a ^ b;
  return ~(((x & ~H) + ~H) | x) & H;
}
template <int BITS> __device__ __forceinline__ unsigned swar_cmp_bits(int op, unsigned a, unsigned b, unsigned bias) {
  constexpr unsigned H = BITS == 8 ? 0x80808080u : 0x80008000u;
  a ^= bias; b ^= bias;
  unsigned r;
  switch (op) {
    case NCL_OP_EQ: r = swar_eq<BITS>(a, b); break;          case NCL_OP_NE: r = ~swar_eq<BITS>(a, b) & H; break;
    case NCL_OP_LT: r = swar_lt<BITS>(a, b); break;          case NCL_OP_GT: r = swar_lt<BITS>(b, a); break;
    case NCL_OP_LE: r = ~swar_lt<BITS>(b, a) & H; break;     default: r = ~swar_lt<BITS>(a, b) & H; break;
  }
  if constexpr (BITS == 8) return (r * 0x204081u) >> 28;     // bits 7, 15, 23, 31 -> 0..3
  else return ((r >> 15) & 1u) | ((r >> 30) & 2u);
}

template <int BITS, int U>
__global__ void __launch_bounds__(EW_THREADS) cmp_pack_swar_kernel(const __grid_constant__ EwParams p) {
  constexpr int EI = 128 / BITS, G = 32 / EI, PER_WORD = 32 / BITS;
  constexpr unsigned SEG = 32 * U;
  const int lane = threadIdx.x & 31;
  const unsigned nwarps = gridDim.x * (EW_THREADS / 32);
  const int op = p.rt_op;
  // signed fields: bias by H (same order as unsigned after the flip); a broadcast scalar is replicated into every field
  const bool sg = (lk_flags(p.st[1][0] ? p.lk[0] : p.lk[1]) & 1u) != 0;
  const unsigned bias = sg ? (BITS == 8 ? 0x80808080u : 0x80008000u) : 0u;
  for (unsigned item = blockIdx.x * (EW_THREADS / 32) + (threadIdx.x >> 5); item < p.nitems; item += nwarps) {
    const unsigned row = ew_div(item, p.seg_mul, p.seg_shr);
    const unsigned seg = item - row * p.segs_per_row;
    long long rowoff[3];
#pragma unroll
    for (int k = 0; k < 3; k++) rowoff[k] = p.off[k];
    {
      unsigned r = row;
      for (int d = 1; d < p.nd; d++) {
        const unsigned r2 = ew_div(r, p.mul[d], p.shr[d]);
        const int c = (int)(r - r2 * p.ext[d]); r = r2;
#pragma unroll
        for (int k = 0; k < 3; k++) rowoff[k] += (long long)c * p.st[k][d];
      }
    }
    const unsigned q0 = seg * SEG + lane;
    uint4 x[2][U];
#pragma unroll
    for (int k = 0; k < 2; k++) {
      if (p.st[1 + k][0] == 0) {
        const unsigned v = (unsigned)load_scalar<long long>(p.base[1 + k], rowoff[1 + k], p.lk[k]) & (BITS == 8 ? 0xffu : 0xffffu);
        const unsigned w = BITS == 8 ? v * 0x01010101u : v * 0x00010001u;
#pragma unroll
        for (int u = 0; u < U; u++) x[k][u] = make_uint4(w, w, w, w);
      } else {
        const char* base = p.base[1 + k] + (rowoff[1 + k] + (long long)q0 * EI) * (BITS / 8);
#pragma unroll
        for (int u = 0; u < U; u++) x[k][u] = (q0 + u * 32 < p.ext[0]) ? load_raw16(base + (size_t)u * 32 * 16) : make_uint4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const unsigned q = q0 + u * 32;
      unsigned m = 0;
      if (q < p.ext[0]) {
        m = swar_cmp_bits<BITS>(op, x[0][u].x, x[1][u].x, bias) | (swar_cmp_bits<BITS>(op, x[0][u].y, x[1][u].y, bias) << PER_WORD)
          | (swar_cmp_bits<BITS>(op, x[0][u].z, x[1][u].z, bias) << (2 * PER_WORD)) | (swar_cmp_bits<BITS>(op, x[0][u].w, x[1][u].w, bias) << (3 * PER_WORD));
      }
      unsigned w = m << (EI * (lane & (G - 1)));
#pragma unroll
      for (int o = 1; o < G; o <<= 1) w |= __shfl_xor_sync(0xffffffffu, w, o);
      if ((lane & (G - 1)) == 0 && q < p.ext[0])
        *(unsigned*)(p.base[0] + ((rowoff[0] + (long long)q * EI) >> 3)) = w;
    }
  }
}
EwKernel ew_pick_cmppack_swar(int bits, int* U) { *U = 4; return bits == 8 ? (EwKernel)cmp_pack_swar_kernel<8, 4> : bits == 16 ? (EwKernel)cmp_pack_swar_kernel<16, 4> : nullptr; }

template <class T, int EI> static EwKernel ck() { return (EwKernel)cmp_pack_kernel<T, EI, (sizeof(T) * EI >= 64 ? 2 : 4)>; }
// lanes: CLS_F float, CLS_D double, CLS_I 64-bit integer lanes, 4 = 32-bit integer lanes; EI = elements per 16-byte chunk of the widest operand
EwKernel ew_pick_cmppack(int lanes, int EI, int* U) {
  *U = 4;
  switch (lanes) {
    case CLS_F: return EI == 4 ? ck<float, 4>() : EI == 8 ? ck<float, 8>() : EI == 16 ? (*U = 2, ck<float, 16>()) : nullptr;
    case CLS_D: return EI == 2 ? ck<double, 2>() : EI == 4 ? ck<double, 4>() : nullptr;
    case 4: return EI == 4 ? ck<int, 4>() : EI == 8 ? ck<int, 8>() : EI == 16 ? (*U = 2, ck<int, 16>()) : nullptr;
    default: return EI == 2 ? ck<long long, 2>() : EI == 4 ? ck<long long, 4>() : nullptr;
  }
}
