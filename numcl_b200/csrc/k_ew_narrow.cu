// k_ew_narrow.cu -- 8-bit integer arrays, SIMD within a register.
//
// numcl narrows every integer array to the smallest element type, so (unsigned-byte 8) data is everywhere, and its type
// inference widens results only as far as needed: ub8 + ub8 and ub8 * 3 are (unsigned-byte 15) arrays, ub8 - ub8 a
// (signed-byte 16) array, (max a b) / (min a b) / logand / logior / logxor of ub8 arrays stay ub8.  The general engine
// unpacks every byte into a 32- or 64-bit lane (~7 instructions per element): 0.42-0.55 of the HBM peak, issue-bound
// (profiles/r01_cliff_hunt.txt).  Here a 32-bit word carries 4 bytes (or 2 halfwords) through the whole computation:
//   same width   max / min : carry-isolated unsigned compare -> byte mask -> select        (~3 instructions per element)
//                and / ior / xor : one instruction per 4 elements
//   8 -> 16 bit  two PRMT spread the 4 bytes of a word into 2 x 2 halfword fields; + and * (by a scalar whose products fit a
//                field) never carry across fields; - is done on fields biased by 0x8000               (~1.5 per element)
// One thread handles 16-byte input chunks of a flat run (same-shape operands or a broadcast scalar), grid-stride.
#include "ncl_device.cuh"

enum { NW_ADD = 0, NW_SUB, NW_MUL, NW_MAX, NW_MIN, NW_AND, NW_IOR, NW_XOR };
struct NarrowParams {
  const char* a; const char* b; char* out;
  long long nchunks;              // 16-byte chunks of the 8-bit operand(s)
  int op, wide;                   // wide: 16-bit result fields
  int b_scalar; long long b_off; int b_lk;   // second operand broadcast: a rank-0 array, read by the kernel (element b_off of b, kind b_lk)
  int swap;                       // the streamed operand is the SECOND argument (matters for -)
  unsigned bias;                  // 0x80808080 when both operands are signed bytes (max / min)
  unsigned mask16;                // result mask per 16-bit field, replicated ((unsigned-byte 15): 0x7fff7fff)
};

__device__ __forceinline__ unsigned nw_lt8(unsigned a, unsigned b) {            // 0xff in every byte where a < b (unsigned)
  const unsigned H = 0x80808080u;
  const unsigned t = (a | H) - (b & ~H);
  const unsigned h = ((~a & b) | (~(a ^ b) & ~t)) & H;
  return (h >> 7) * 0xffu;
}
__device__ __forceinline__ unsigned nw_same8(int op, unsigned a, unsigned b, unsigned bias) {
  switch (op) {
    case NW_AND: return a & b; case NW_IOR: return a | b; case NW_XOR: return a ^ b;
    case NW_MAX: { const unsigned m = nw_lt8(a ^ bias, b ^ bias); return (a & ~m) | (b & m); }
    default:     { const unsigned m = nw_lt8(b ^ bias, a ^ bias); return (a & ~m) | (b & m); }      // min
  }
}
// two halfword fields from the low / high pair of bytes of w
__device__ __forceinline__ unsigned nw_lo16(unsigned w) { return __byte_perm(w, 0u, 0x4140); }
__device__ __forceinline__ unsigned nw_hi16(unsigned w) { return __byte_perm(w, 0u, 0x4342); }
__device__ __forceinline__ unsigned nw_wide(int op, unsigned x, unsigned y) {      // x, y: 2 halfword fields each
  switch (op) {
    case NW_ADD: return x + y;
    case NW_MUL: return x * y;                                                     // y = the scalar (a single field): products fit
    default:     return ((x | 0x80008000u) - y) ^ 0x80008000u;                     // x - y per field, two's complement in 16 bits
  }
}

__global__ void __launch_bounds__(256) ew_narrow_kernel(const __grid_constant__ NarrowParams p) {
  const long long nthr = (long long)gridDim.x * 256;
  unsigned bval = 0;
  if (p.b_scalar) {
    // replicated into every field of a word: 8-bit fields for the same-width functions, ONE 16-bit field for *, both for + -
    const unsigned v = (unsigned)load_scalar<long long>(p.b, p.b_off, p.b_lk);
    bval = !p.wide ? (v & 0xffu) * 0x01010101u : (p.op == NW_MUL ? (v & 0xffffu) : (v & 0xffffu) * 0x00010001u);
  }
  const uint4 bs = make_uint4(bval, bval, bval, bval);
  for (long long c = (long long)blockIdx.x * 256 + threadIdx.x; c < p.nchunks; c += 2 * nthr) {
    const bool two = c + nthr < p.nchunks;
    uint4 a0 = load_raw16(p.a + c * 16), a1 = two ? load_raw16(p.a + (c + nthr) * 16) : a0;
    uint4 b0 = bs, b1 = bs;
    if (!p.b_scalar) { b0 = load_raw16(p.b + c * 16); if (two) b1 = load_raw16(p.b + (c + nthr) * 16); }
#pragma unroll
    for (int h = 0; h < 2; h++) {
      if (h == 1 && !two) break;
      const uint4 a = h ? a1 : a0, b = h ? b1 : b0;
      const long long cc = h ? c + nthr : c;
      const unsigned aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
      if (!p.wide) {
        unsigned r[4];
#pragma unroll
        for (int i = 0; i < 4; i++) r[i] = nw_same8(p.op, aw[i], bw[i], p.bias);
        *(uint4*)(p.out + cc * 16) = make_uint4(r[0], r[1], r[2], r[3]);
      } else {
        unsigned r[8];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          unsigned xl = nw_lo16(aw[i]), xh = nw_hi16(aw[i]);
          unsigned yl = p.b_scalar ? bw[i] : nw_lo16(bw[i]), yh = p.b_scalar ? bw[i] : nw_hi16(bw[i]);
          if (p.swap) { const unsigned t0 = xl; xl = yl; yl = t0; const unsigned t1 = xh; xh = yh; yh = t1; }
          r[2 * i] = nw_wide(p.op, xl, yl) & p.mask16; r[2 * i + 1] = nw_wide(p.op, xh, yh) & p.mask16;
        }
        *(uint4*)(p.out + cc * 32) = make_uint4(r[0], r[1], r[2], r[3]);
        *(uint4*)(p.out + cc * 32 + 16) = make_uint4(r[4], r[5], r[6], r[7]);
      }
    }
  }
}

// a, b, out: first byte of each flat run; n: elements (a multiple of 16).  See try_elementwise for the conditions.
int ew_narrow_launch(int op, bool wide, const char* a, const char* b, char* out, long long n, bool b_scalar, long long b_off, int b_lk, bool swap,
                     bool signed8, unsigned mask16) {
  Context& c = ctx();
  NarrowParams p; memset(&p, 0, sizeof p);
  p.a = a; p.b = b; p.out = out; p.nchunks = n / 16; p.op = op; p.wide = wide ? 1 : 0;
  p.b_scalar = b_scalar ? 1 : 0; p.b_off = b_off; p.b_lk = b_lk; p.swap = swap ? 1 : 0; p.bias = signed8 ? 0x80808080u : 0u; p.mask16 = mask16;
  long long blocks = (p.nchunks + 511) / 512, cap = (long long)c.sm_count * 16;
  if (blocks > cap) blocks = cap; if (blocks < 1) blocks = 1;
  ew_narrow_kernel<<<(unsigned)blocks, 256, 0, c.stream>>>(p);
  note_launch("bcast_narrow");
  return cuda_fail(cudaGetLastError(), "ew_narrow_kernel");
}
