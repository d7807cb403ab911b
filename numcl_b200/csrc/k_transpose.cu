// k_transpose.cu -- K5: N-D axis permutation, the einsum nests with no contracted index whose input
// and output are contiguous along DIFFERENT loops: `transpose` (src/4linear-algebra2.lisp:38-45),
// `(ij -> ji)`, `(abcd -> dbca)` (config C5b).
//
// HBM-bound copy.  A direct kernel would be coalesced on one side only, so the (input-contiguous,
// output-contiguous) plane is tiled through shared memory: a 256-thread block reads a square tile
// along the input's contiguous loop (64 x 64 for words of up to 4 bytes: 16 independent loads per
// thread; 32 x 32 for 8-byte words), parks it in a padded tile (odd row pitch: conflict-free for both
// phases) and writes it out along the output's contiguous loop.  All remaining loops are batch
// coordinates decoded from the block index.  Elements pass through the value class in the converting
// variant, so element-type changing copies (ub8 -> fixnum: an integer einsum output is always fixnum,
// 3einsum.lisp:517-518) and the `0 + x` of the default transform into a fresh output (-0.0 -> +0.0)
// cost nothing extra.  Planes with one short side: transpose_reg_kernel ((N,C) <-> (C,N), C <= 8, in
// registers) and transpose_thin_kernel (C <= 16, rectangular tiles).
#include "ncl_device.cuh"

#define TR_MAXB 6

struct TrParams {
  const char* in; char* out; long long in_off, out_off; int lk; StoreSpec sk;
  long long ni, no;                 // extents of the input-contiguous loop (i) and output-contiguous loop (o)
  long long in_so, out_si;          // input stride along o, output stride along i (elements)
  int nb; long long bext[TR_MAXB], bin[TR_MAXB], bout[TR_MAXB];
  long long tiles_i, tiles_o;
  int add_zero;                     // default transform into a fresh output: x + 0
};

template <class T, int TILE>
__global__ void __launch_bounds__(256) transpose_kernel(const __grid_constant__ TrParams p) {
  __shared__ T tile[TILE][TILE + 1];
  long long b = blockIdx.x;
  long long ti = b % p.tiles_i; b /= p.tiles_i;
  long long to = b % p.tiles_o; b /= p.tiles_o;
  long long ibase = p.in_off, obase = p.out_off;
  for (int k = p.nb - 1; k >= 0; k--) { long long c = b % p.bext[k]; b /= p.bext[k]; ibase += c * p.bin[k]; obase += c * p.bout[k]; }
  const int tx = threadIdx.x % TILE, ty = threadIdx.x / TILE;
  constexpr int ROWS = 256 / TILE;
  long long i0 = ti * TILE, o0 = to * TILE;
  // load: rows of the tile run along o, threads along i (input-contiguous)
#pragma unroll
  for (int r = 0; r < TILE; r += ROWS) {
    long long o = o0 + r + ty, i = i0 + tx;
    if (o < p.no && i < p.ni) tile[r + ty][tx] = load_scalar<T>(p.in, ibase + o * p.in_so + i, p.lk);
  }
  __syncthreads();
  // store: rows run along i, threads along o (output-contiguous)
#pragma unroll
  for (int r = 0; r < TILE; r += ROWS) {
    long long i = i0 + r + ty, o = o0 + tx;
    if (i < p.ni && o < p.no) {
      T v = tile[tx][r + ty];
      if (p.add_zero) v = Ops<T>::add(Ops<T>::zero(), v);
      store_scalar<T>(p.out, obase + i * p.out_si + o, p.sk, v);
    }
  }
}

// Fast path: same element type on both sides (or an integer type, for which `0 + x` is the identity):
// the elements move as raw W-byte words, no element-kind dispatch in the loops.
// FZ = 1 / 2: single / double floats through the default transform into a fresh output: x + 0.0.
template <class WT, int TILE, int FZ>
__global__ void __launch_bounds__(256) transpose_raw_kernel(const __grid_constant__ TrParams p) {
  __shared__ WT tile[TILE][TILE + 1];
  pdl_trigger();                          // programmatic dependent launch: the tile is read while the previous kernel drains, stored after it
  long long b = blockIdx.x;
  long long ti = b % p.tiles_i; b /= p.tiles_i;
  long long to = b % p.tiles_o; b /= p.tiles_o;
  long long ibase = p.in_off, obase = p.out_off;
  for (int k = p.nb - 1; k >= 0; k--) { long long c = b % p.bext[k]; b /= p.bext[k]; ibase += c * p.bin[k]; obase += c * p.bout[k]; }
  const int tx = threadIdx.x % TILE, ty = threadIdx.x / TILE;
  constexpr int ROWS = 256 / TILE;
  const long long i0 = ti * TILE, o0 = to * TILE;
  const WT* src = (const WT*)p.in + ibase + (o0 + ty) * p.in_so + i0 + tx;
  WT* dst = (WT*)p.out + obase + (i0 + ty) * p.out_si + o0 + tx;
  const bool full = (i0 + TILE <= p.ni) && (o0 + TILE <= p.no);
  if (full) {
    WT v[TILE / ROWS];
#pragma unroll
    for (int r = 0; r < TILE / ROWS; r++) v[r] = __ldg(src + (long long)(r * ROWS) * p.in_so);
#pragma unroll
    for (int r = 0; r < TILE / ROWS; r++) tile[r * ROWS + ty][tx] = v[r];
    __syncthreads();
    pdl_wait();
#pragma unroll
    for (int r = 0; r < TILE / ROWS; r++) {
      WT w = tile[tx][r * ROWS + ty];
      if constexpr (FZ == 1) w = (WT)__float_as_uint(__fadd_rn(0.0f, __uint_as_float((unsigned)w)));
      if constexpr (FZ == 2) w = (WT)__double_as_longlong(__dadd_rn(0.0, __longlong_as_double((long long)w)));
      dst[(long long)(r * ROWS) * p.out_si] = w;
    }
  } else {
#pragma unroll
    for (int r = 0; r < TILE; r += ROWS) if (o0 + r + ty < p.no && i0 + tx < p.ni) tile[r + ty][tx] = __ldg(src + (long long)r * p.in_so);
    __syncthreads();
    pdl_wait();
#pragma unroll
    for (int r = 0; r < TILE; r += ROWS) if (i0 + r + ty < p.ni && o0 + tx < p.no) {
      WT w = tile[tx][r + ty];
      if constexpr (FZ == 1) w = (WT)__float_as_uint(__fadd_rn(0.0f, __uint_as_float((unsigned)w)));
      if constexpr (FZ == 2) w = (WT)__double_as_longlong(__dadd_rn(0.0, __longlong_as_double((long long)w)));
      dst[(long long)r * p.out_si] = w;
    }
  }
}

// Thin planes: one of the two contiguous loops is short (a trailing axis of 3: (N,3) -> (3,N)).  Square tiles would be
// mostly empty, so the tile is TI x TO with TI*TO = 4096 (16 independent loads per thread: a CTA's only latency hiding is
// what it has in flight before its barrier) and the short side covering the whole short loop; the long side
// keeps every warp access on contiguous memory (rows of the short loop are adjacent when the array is contiguous).
// MODE 0: raw words, 1 / 2: raw single / double floats through x + 0.0, 3: through the value class V (type-changing).
template <class V, int TI, int TO, int MODE>
__global__ void __launch_bounds__(256) transpose_thin_kernel(const __grid_constant__ TrParams p) {
  constexpr bool I_SHORT = TI <= TO;
  constexpr int PITCH = I_SHORT ? TO + 32 / TI : TI + 32 / TO;      // conflict-free for both phases
  __shared__ V tile[(I_SHORT ? TI : TO) * PITCH];
  long long b = blockIdx.x;
  long long ti = b % p.tiles_i; b /= p.tiles_i;
  long long to = b % p.tiles_o; b /= p.tiles_o;
  long long ibase = p.in_off, obase = p.out_off;
  for (int k = p.nb - 1; k >= 0; k--) { long long c = b % p.bext[k]; b /= p.bext[k]; ibase += c * p.bin[k]; obase += c * p.bout[k]; }
  const long long i0 = ti * TI, o0 = to * TO;
  auto slot = [](int i, int o) { return I_SHORT ? i * PITCH + o : o * PITCH + i; };
  // load: threads run along i (input-contiguous) first.  All of a thread's loads are issued before the first shared-memory
  // store: out-of-range positions read a valid dummy address instead of branching around the load (with a branch per element
  // ptxas kept ONE load in flight per thread).  Element e = tid + 256 k of the tile is (i, o) = (e % TI, e / TI): the part
  // that depends on k is a compile-time constant, so an address is one multiply-add from a per-thread base (ncu on the first
  // version: 84 % issue-slot utilisation -- index arithmetic, not memory, was the limit).
  constexpr int N = TI * TO / 256;
  const int tid = threadIdx.x;
  V v[N];
  {
    constexpr bool NARROW = TI <= 256;                      // one i per thread, o advances by 256 / TI per step
    const int ti0 = NARROW ? tid % TI : tid, to0 = NARROW ? tid / TI : 0;
    const long long base = ibase + (o0 + to0) * p.in_so + i0 + ti0;
#pragma unroll
    for (int k = 0; k < N; k++) {
      const int di = NARROW ? 0 : 256 * (k % (TI / 256 > 0 ? TI / 256 : 1)), dO = NARROW ? k * (256 / (TI < 256 ? TI : 256)) : k / (TI / 256 > 0 ? TI / 256 : 1);
      const bool ok = i0 + ti0 + di < p.ni && o0 + to0 + dO < p.no;
      const long long idx = ok ? base + (long long)dO * p.in_so + di : p.in_off;
      if constexpr (MODE == 3) v[k] = load_scalar<V>(p.in, idx, p.lk);
      else v[k] = __ldg((const V*)p.in + idx);
    }
#pragma unroll
    for (int k = 0; k < N; k++) {
      const int di = NARROW ? 0 : 256 * (k % (TI / 256 > 0 ? TI / 256 : 1)), dO = NARROW ? k * (256 / (TI < 256 ? TI : 256)) : k / (TI / 256 > 0 ? TI / 256 : 1);
      tile[slot(ti0 + di, to0 + dO)] = v[k];
    }
  }
  __syncthreads();
  // store: threads run along o (output-contiguous) first
  {
    constexpr bool NARROW = TO <= 256;
    const int so0 = NARROW ? tid % TO : tid, si0 = NARROW ? tid / TO : 0;
    const long long base = obase + (i0 + si0) * p.out_si + o0 + so0;
#pragma unroll
    for (int k = 0; k < N; k++) {
      const int dO = NARROW ? 0 : 256 * (k % (TO / 256 > 0 ? TO / 256 : 1)), di = NARROW ? k * (256 / (TO < 256 ? TO : 256)) : k / (TO / 256 > 0 ? TO / 256 : 1);
      v[k] = tile[slot(si0 + di, so0 + dO)];
    }
#pragma unroll
    for (int k = 0; k < N; k++) {
      const int dO = NARROW ? 0 : 256 * (k % (TO / 256 > 0 ? TO / 256 : 1)), di = NARROW ? k * (256 / (TO < 256 ? TO : 256)) : k / (TO / 256 > 0 ? TO / 256 : 1);
      if (i0 + si0 + di < p.ni && o0 + so0 + dO < p.no) {
        V w = v[k];
        const long long idx = base + (long long)di * p.out_si + dO;
        if constexpr (MODE == 1) w = (V)__float_as_uint(__fadd_rn(0.0f, __uint_as_float((unsigned)w)));
        if constexpr (MODE == 2) w = (V)__double_as_longlong(__dadd_rn(0.0, __longlong_as_double((long long)w)));
        if constexpr (MODE == 3) { if (p.add_zero) w = Ops<V>::add(Ops<V>::zero(), w); store_scalar<V>(p.out, idx, p.sk, w); }
        else ((V*)p.out)[idx] = w;
      }
    }
  }
}
// (N, C) <-> (C, N) with C <= 8, entirely in registers: a thread owns E = 16 / sizeof(V) consecutive positions of the long
// loop, i.e. C whole chunks on the side where the short loop is contiguous and one chunk of each of the C rows on the other
// side; every index below is a compile-time constant.  No shared memory, no barrier, C 128-bit loads in flight per thread and
// fully coalesced 128-bit accesses on both sides.  FWD: input (N, C) rows back to back -> output C rows of length N.
template <class V, int C, bool FWD, int MODE>
__global__ void __launch_bounds__(256) transpose_reg_kernel(const __grid_constant__ TrParams p) {
  constexpr int E = 16 / (int)sizeof(V);
  const long long q = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long n_long = FWD ? p.no : p.ni;
  if (q * E >= n_long) return;
  auto fix = [](V w) {
    if constexpr (MODE == 1) return (V)__float_as_uint(__fadd_rn(0.0f, __uint_as_float((unsigned)w)));
    else if constexpr (MODE == 2) return (V)__double_as_longlong(__dadd_rn(0.0, __longlong_as_double((long long)w)));
    else return w;
  };
  const V* in = (const V*)p.in + p.in_off; V* out = (V*)p.out + p.out_off;
  if (q * E + E <= n_long) {
    V a[C][E], b[C][E];
    if constexpr (FWD) {
#pragma unroll
      for (int k = 0; k < C; k++) { Chunk<16> c; ld_chunk(c, (const char*)(in + (q * E * C + k * E))); 
#pragma unroll
        for (int j = 0; j < E; j++) a[k][j] = (V)chunk_get<8 * (int)sizeof(V)>(c, j); }
#pragma unroll
      for (int f = 0; f < C * E; f++) b[f % C][f / C] = fix(a[f / E][f % E]);           // flat f = o_local * C + i
#pragma unroll
      for (int i = 0; i < C; i++) { Chunk<16> c;
#pragma unroll
        for (int j = 0; j < E; j++) chunk_put<8 * (int)sizeof(V)>(c, j, (unsigned long long)b[i][j]);
        st_chunk(c, (char*)(out + (i * p.out_si + q * E))); }
    } else {
#pragma unroll
      for (int o = 0; o < C; o++) { Chunk<16> c; ld_chunk(c, (const char*)(in + (o * p.in_so + q * E)));
#pragma unroll
        for (int j = 0; j < E; j++) a[o][j] = (V)chunk_get<8 * (int)sizeof(V)>(c, j); }
#pragma unroll
      for (int f = 0; f < C * E; f++) b[f / E][f % E] = fix(a[f % C][f / C]);           // flat f = i_local * C + o
#pragma unroll
      for (int k = 0; k < C; k++) { Chunk<16> c;
#pragma unroll
        for (int j = 0; j < E; j++) chunk_put<8 * (int)sizeof(V)>(c, j, (unsigned long long)b[k][j]);
        st_chunk(c, (char*)(out + (q * E * C + k * E))); }
    }
  } else {                                                  // fewer than E positions left on the long loop
    for (long long l = q * E; l < n_long; l++)
      for (int s = 0; s < C; s++) {
        if constexpr (FWD) out[s * p.out_si + l] = fix(in[l * C + s]);
        else out[l * C + s] = fix(in[s * p.in_so + l]);
      }
  }
}

// Strips: (N, C) <-> (C, N) with a short side of up to 128 whose rows lie back to back.  Square tiles waste the lanes that fall
// outside a 48-wide plane (transpose (4M,48): 0.57 of the HBM peak) and the rectangular tiles of transpose_thin_kernel stop
// at C = 16.  Here a CTA takes R consecutive positions of the long loop (R a power of two <= 256, the tile of R x C words in up
// to 72 KB of shared memory) and ALL of the short loop: on the side where the short loop is contiguous that is ONE contiguous
// run of R * C words, moved with fully coalesced accesses whatever C is (a thread's flat index advances by 256: its (row,
// column) is stepped, never divided); on the other side it is C runs of R words, a thread staying on one position of the run
// and stepping through the rows with a constant pointer increment.  The tile has an odd pitch: lanes running along either loop
// hit 32 different banks.  Words travel global -> shared with cp.async (LDGSTS): nothing is staged in registers, every copy of
// a thread is in flight at once at 32 registers.  (First version: 12 elements per thread behind a prologue of 64-bit
// divisions -- ncu: 2.1 warp instructions per ELEMENT, issue slots 82 % busy, 0.45 of the peak.)
#define STRIP_SMEM_BYTES (72 * 1024)
template <int BYTES> __device__ __forceinline__ void cp_async_word(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(s), "l"(gmem), "n"(BYTES) : "memory");
}
template <class WT, bool FWD, int FZ>
__global__ void __launch_bounds__(256) transpose_strip_kernel(const __grid_constant__ TrParams p, const int C, const int pitch, const int rshift) {
  extern __shared__ __align__(16) unsigned char strip_smem[];
  WT* tile = reinterpret_cast<WT*>(strip_smem);
  const int R = 1 << rshift, tid = threadIdx.x;
  long long strip = blockIdx.x, ibase = p.in_off, obase = p.out_off;
  if (p.nb > 0) {
    long long b = blockIdx.x;
    strip = b % p.tiles_o; b /= p.tiles_o;
    for (int k = p.nb - 1; k >= 0; k--) { long long c = b % p.bext[k]; b /= p.bext[k]; ibase += c * p.bin[k]; obase += c * p.bout[k]; }
  }
  const long long nlong = FWD ? p.no : p.ni, n0 = strip * R;
  const int rows = (int)(nlong - n0 < R ? nlong - n0 : R), total = rows * C;
  auto fix = [](WT w) {
    if constexpr (FZ == 1) return (WT)__float_as_uint(__fadd_rn(0.0f, __uint_as_float((unsigned)w)));
    else if constexpr (FZ == 2) return (WT)__double_as_longlong(__dadd_rn(0.0, __longlong_as_double((long long)w)));
    else return w;
  };
  // flat side: index tid + 256 k -> (row, column), stepped;  row side: position r2 of the run, rows c2 = cb, cb + cstep, ...
  const int r0 = tid / C, c0 = tid - r0 * C, dr = 256 / C, dc = 256 - dr * C;
  const int r2 = tid & (R - 1), cb = tid >> rshift, cstep = 256 >> rshift;
  if constexpr (FWD) {
    const WT* src = (const WT*)p.in + ibase + n0 * C + tid;
    int r = r0, c = c0;
#pragma unroll 4
    for (int e = tid; e < total; e += 256) {
      cp_async_word<sizeof(WT)>(&tile[r * pitch + c], src);
      src += 256; r += dr; c += dc; if (c >= C) { c -= C; r++; }
    }
    asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (r2 < rows) {
      WT* dst = (WT*)p.out + obase + n0 + r2 + (long long)cb * p.out_si;
      const long long dstep = (long long)cstep * p.out_si;
      const WT* t = tile + r2 * pitch + cb;
#pragma unroll 4
      for (int c2 = cb; c2 < C; c2 += cstep) { *dst = fix(*t); dst += dstep; t += cstep; }
    }
  } else {
    if (r2 < rows) {
      const WT* src = (const WT*)p.in + ibase + n0 + r2 + (long long)cb * p.in_so;
      const long long sstep = (long long)cstep * p.in_so;
      WT* t = tile + r2 * pitch + cb;
#pragma unroll 4
      for (int c2 = cb; c2 < C; c2 += cstep) { cp_async_word<sizeof(WT)>(t, src); src += sstep; t += cstep; }
    }
    asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    WT* dst = (WT*)p.out + obase + n0 * C + tid;
    int r = r0, c = c0;
#pragma unroll 4
    for (int e = tid; e < total; e += 256) {
      *dst = fix(tile[r * pitch + c]);
      dst += 256; r += dr; c += dc; if (c >= C) { c -= C; r++; }
    }
  }
}
typedef void (*StripKernel)(const TrParams, int, int, int);
template <class WT> static StripKernel strip_pick(bool fwd, int fz) {
  if (fwd) return fz == 1 ? transpose_strip_kernel<WT, true, 1> : fz == 2 ? transpose_strip_kernel<WT, true, 2> : transpose_strip_kernel<WT, true, 0>;
  return fz == 1 ? transpose_strip_kernel<WT, false, 1> : fz == 2 ? transpose_strip_kernel<WT, false, 2> : transpose_strip_kernel<WT, false, 0>;
}
typedef void (*TrKernel)(const TrParams);
template <class V, bool FWD, int MODE> static TrKernel reg_pick(int c) {
  switch (c) {
    case 2: return transpose_reg_kernel<V, 2, FWD, MODE>; case 3: return transpose_reg_kernel<V, 3, FWD, MODE>;
    case 4: return transpose_reg_kernel<V, 4, FWD, MODE>; case 5: return transpose_reg_kernel<V, 5, FWD, MODE>;
    case 6: return transpose_reg_kernel<V, 6, FWD, MODE>; case 7: return transpose_reg_kernel<V, 7, FWD, MODE>;
    case 8: return transpose_reg_kernel<V, 8, FWD, MODE>; default: return nullptr;
  }
}
template <class V, int MODE> static TrKernel thin_pick(int ti) {
  switch (ti) {
    case 4: return transpose_thin_kernel<V, 4, 1024, MODE>;  case 8: return transpose_thin_kernel<V, 8, 512, MODE>;
    case 16: return transpose_thin_kernel<V, 16, 256, MODE>; case 256: return transpose_thin_kernel<V, 256, 16, MODE>;
    case 512: return transpose_thin_kernel<V, 512, 8, MODE>; default: return transpose_thin_kernel<V, 1024, 4, MODE>;
  }
}

int try_transpose(const Nest& n) {
  if (n.n_out != 1 || n.n_in != 1) return NCL_ERR_UNSUPPORTED;
  const Operand& in = n.in[0]; const Operand& out = n.out[0];
  const Expr& e = n.transform[0];
  int add_zero = 0;
  const XNode& root = e.node[e.n - 1];
  if (root.op == NCL_X_INPUT) add_zero = 0;
  else if (e.n == 3 && root.op == NCL_OP_ADD && e.node[root.a].op == NCL_X_OUTPUT && e.node[root.b].op == NCL_X_INPUT && n.fresh) add_zero = 1;
  else return NCL_ERR_UNSUPPORTED;
  int icls = g_dt[in.dtype].cls, ocls = g_dt[out.dtype].cls;
  if (icls != ocls) return NCL_ERR_UNSUPPORTED;
  if (g_dt[in.dtype].slot_bits < 8 || g_dt[out.dtype].slot_bits < 8) return NCL_ERR_UNSUPPORTED;
  int li = -1, lo = -1;
  for (int l = 0; l < n.nloops; l++) {
    if (n.extent[l] == 1) continue;
    if (out.stride[l] == 0 || in.stride[l] == 0) return NCL_ERR_UNSUPPORTED;   // contraction or broadcast
    if (in.stride[l] == 1 && li < 0) li = l;
    if (out.stride[l] == 1 && lo < 0) lo = l;
  }
  if (li < 0 || lo < 0 || li == lo) return NCL_ERR_UNSUPPORTED;
  TrParams p; memset(&p, 0, sizeof p);
  p.in = in.base; p.out = out.base; p.in_off = in.offset; p.out_off = out.offset;
  p.lk = dtype_load_kind(in.dtype); p.sk = dtype_store_spec(out.dtype); p.add_zero = add_zero;
  p.ni = n.extent[li]; p.no = n.extent[lo]; p.in_so = in.stride[lo]; p.out_si = out.stride[li];
  // batch loops, merged where both sides allow
  struct B { long long ext, sin, sout; } bd[NCL_MAX_LOOPS]; int nb = 0;
  for (int l = 0; l < n.nloops; l++) { if (l == li || l == lo || n.extent[l] == 1) continue; bd[nb++] = {n.extent[l], in.stride[l], out.stride[l]}; }
  for (int i = 1; i < nb; i++) { B t = bd[i]; int j = i - 1; while (j >= 0 && bd[j].sout < t.sout) { bd[j + 1] = bd[j]; j--; } bd[j + 1] = t; }
  for (int i = nb - 1; i > 0; i--)
    if (bd[i - 1].sin == bd[i].sin * bd[i].ext && bd[i - 1].sout == bd[i].sout * bd[i].ext) {
      bd[i - 1].ext *= bd[i].ext; bd[i - 1].sin = bd[i].sin; bd[i - 1].sout = bd[i].sout;
      for (int j = i; j + 1 < nb; j++) bd[j] = bd[j + 1]; nb--;
    }
  if (nb > TR_MAXB) return NCL_ERR_UNSUPPORTED;
  p.nb = nb; long long batch = 1;
  for (int i = 0; i < nb; i++) { p.bext[i] = bd[i].ext; p.bin[i] = bd[i].sin; p.bout[i] = bd[i].sout; batch *= bd[i].ext; }
  Context& c = ctx();
  // raw-word fast path: identical storage on both sides
  bool raw = in.dtype == out.dtype || (icls == CLS_I && g_dt[in.dtype].slot_bits == g_dt[out.dtype].slot_bits &&
             g_dt[in.dtype].tagged == g_dt[out.dtype].tagged && g_dt[out.dtype].value_bits >= g_dt[in.dtype].value_bits &&
             g_dt[in.dtype].is_signed == g_dt[out.dtype].is_signed);
  int wbytes = g_dt[in.dtype].slot_bits / 8;
  if (raw && nb == 0 && (wbytes == 4 || wbytes == 8) && ((p.ni >= 2 && p.ni <= 8 && p.no >= 1024) || (p.no >= 2 && p.no <= 8 && p.ni >= 1024))) {
    // (N, C) <-> (C, N), C <= 8: the register kernel, when both sides allow 128-bit accesses
    const bool fwd = p.ni <= 8;
    auto al = [&](const Operand& o, long long off) { return (((uintptr_t)o.base + off * wbytes) % 16) == 0; };
    const bool ok = al(in, in.offset) && al(out, out.offset) &&
                    (fwd ? (p.in_so == p.ni && (p.out_si * wbytes) % 16 == 0) : (p.out_si == p.no && (p.in_so * wbytes) % 16 == 0));
    if (ok) {
      const int fz = (add_zero && icls == CLS_F) ? 1 : (add_zero && icls == CLS_D) ? 2 : 0;
      const int cshort = (int)(fwd ? p.ni : p.no); const long long nlong = fwd ? p.no : p.ni;
      TrKernel k;
      if (wbytes == 4) k = fwd ? (fz ? reg_pick<unsigned int, true, 1>(cshort) : reg_pick<unsigned int, true, 0>(cshort))
                               : (fz ? reg_pick<unsigned int, false, 1>(cshort) : reg_pick<unsigned int, false, 0>(cshort));
      else k = fwd ? (fz ? reg_pick<unsigned long long, true, 2>(cshort) : reg_pick<unsigned long long, true, 0>(cshort))
                   : (fz ? reg_pick<unsigned long long, false, 2>(cshort) : reg_pick<unsigned long long, false, 0>(cshort));
      const long long threads = (nlong + 16 / wbytes - 1) / (16 / wbytes), blocks = (threads + 255) / 256;
      if (k && blocks <= 0x7fffffffLL) {
        k<<<(unsigned)blocks, 256, 0, c.stream>>>(p);
        note_launch("transpose_reg");
        return cuda_fail(cudaGetLastError(), "transpose_reg_kernel");
      }
    }
  }
  if (raw && (wbytes == 4 || wbytes == 8)) {
    // strips: the short side (<= 128) lies back to back on the side where it is contiguous
    const bool fwd = p.ni <= p.no;
    const long long cs = fwd ? p.ni : p.no, nl = fwd ? p.no : p.ni;
    if (cs >= 2 && cs <= 128 && nl >= 1024 && (fwd ? p.in_so == p.ni : p.out_si == p.no)) {
      const int C = (int)cs, pitch = C | 1;
      int rshift = 5; while (rshift < 8 && (size_t)(2 << rshift) * pitch * wbytes <= STRIP_SMEM_BYTES) rshift++;      // R = 32 .. 256
      const size_t smem = ((size_t)(1 << rshift) * pitch * wbytes + 15) & ~(size_t)15;
      p.tiles_i = 1; p.tiles_o = (nl + (1 << rshift) - 1) >> rshift;
      const long long blocks = p.tiles_o * batch;
      if (blocks > 0 && blocks <= 0x7fffffffLL) {
        const int fz = (add_zero && icls == CLS_F) ? 1 : (add_zero && icls == CLS_D) ? 2 : 0;
        StripKernel k = wbytes == 4 ? strip_pick<unsigned int>(fwd, fz) : strip_pick<unsigned long long>(fwd, fz);
        static bool attr_set[2][2][3] = {};
        bool& done = attr_set[wbytes == 8][fwd][fz];
        if (!done) { NCL_CUDA(cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, STRIP_SMEM_BYTES)); done = true; }
        k<<<(unsigned)blocks, 256, smem, c.stream>>>(p, C, pitch, rshift);
        note_launch("transpose_strip");
        return cuda_fail(cudaGetLastError(), "transpose_strip_kernel");
      }
    }
  }
  if ((p.ni <= 16 || p.no <= 16) && (p.ni >= 128 || p.no >= 128)) {
    // thin plane: rectangular tiles, short side first
    const long long sh = p.ni <= p.no ? p.ni : p.no;
    const int ts = sh <= 4 ? 4 : sh <= 8 ? 8 : 16, tl = 4096 / ts;
    const int TI = p.ni <= p.no ? ts : tl, TO = p.ni <= p.no ? tl : ts;
    p.tiles_i = (p.ni + TI - 1) / TI; p.tiles_o = (p.no + TO - 1) / TO;
    long long blocks = p.tiles_i * p.tiles_o * batch;
    if (blocks > 0x7fffffffLL || blocks <= 0) return NCL_ERR_UNSUPPORTED;
    TrKernel k;
    if (raw) {
      int fz = (add_zero && icls == CLS_F) ? 1 : (add_zero && icls == CLS_D) ? 2 : 0;
      if (wbytes == 8) k = fz == 2 ? thin_pick<unsigned long long, 2>(TI) : thin_pick<unsigned long long, 0>(TI);
      else if (wbytes == 4) k = fz == 1 ? thin_pick<unsigned int, 1>(TI) : thin_pick<unsigned int, 0>(TI);
      else if (wbytes == 2) k = thin_pick<unsigned short, 0>(TI);
      else k = thin_pick<unsigned char, 0>(TI);
    } else k = icls == CLS_F ? thin_pick<float, 3>(TI) : icls == CLS_D ? thin_pick<double, 3>(TI) : thin_pick<long long, 3>(TI);
    k<<<(unsigned)blocks, 256, 0, c.stream>>>(p);
    note_launch("transpose_thin");
    return cuda_fail(cudaGetLastError(), "transpose_thin_kernel");
  }
  if (raw) {
    int tile = wbytes >= 8 ? 32 : 64;      // 4-byte words: 64 x 64 (16 loads in flight per thread: abc -> cab 0.66 -> 1.0 of the HBM peak)
    p.tiles_i = (p.ni + tile - 1) / tile; p.tiles_o = (p.no + tile - 1) / tile;
    long long blocks = p.tiles_i * p.tiles_o * batch;
    if (blocks > 0x7fffffffLL || blocks <= 0) return NCL_ERR_UNSUPPORTED;
    int fz = (add_zero && icls == CLS_F) ? 1 : (add_zero && icls == CLS_D) ? 2 : 0;
    TrKernel k;
    if (wbytes == 8) k = fz == 2 ? (TrKernel)transpose_raw_kernel<unsigned long long, 32, 2> : (TrKernel)transpose_raw_kernel<unsigned long long, 32, 0>;
    else if (wbytes == 4) k = fz == 1 ? (TrKernel)transpose_raw_kernel<unsigned int, 64, 1> : (TrKernel)transpose_raw_kernel<unsigned int, 64, 0>;
    else if (wbytes == 2) k = (TrKernel)transpose_raw_kernel<unsigned short, 64, 0>;
    else k = (TrKernel)transpose_raw_kernel<unsigned char, 64, 0>;
    PdlRange rd = pdl_range(in);
    const bool early = pdl_eligible(&rd, 1);
    void* args[1] = {(void*)&p};
    NCL_TRY(pdl_launch((const void*)k, dim3((unsigned)blocks), dim3(256), args, c.stream, early));
    note_launch("transpose_tile");
    PdlRange wr = pdl_range(out); pdl_commit(early, &wr, 1);
    return NCL_OK;
  }
  int tile = 64;
  p.tiles_i = (p.ni + tile - 1) / tile; p.tiles_o = (p.no + tile - 1) / tile;
  long long blocks = p.tiles_i * p.tiles_o * batch;
  if (blocks > 0x7fffffffLL || blocks <= 0) return NCL_ERR_UNSUPPORTED;
  if (icls == CLS_F) transpose_kernel<float, 64><<<(unsigned)blocks, 256, 0, c.stream>>>(p);
  else if (icls == CLS_D) transpose_kernel<double, 64><<<(unsigned)blocks, 256, 0, c.stream>>>(p);
  else transpose_kernel<long long, 64><<<(unsigned)blocks, 256, 0, c.stream>>>(p);
  note_launch("transpose_convert");
  return cuda_fail(cudaGetLastError(), "transpose_kernel");
}
