// k_reduce_rot.cu -- column reduce of a matrix whose row pitch is NOT a multiple of 16 bytes ((sum a :axes 0) of a
// (32767, 8193) single-float array; src/5reduce.lisp:67-87).  The one-element-per-thread path such an array used to take moves
// 4 bytes per load and ran at 0.55 of the HBM peak (ncu: 0.6 warp instructions per element, issue slots 62 % busy, DRAM 45 %).
//
// Rows r, r + E, r + 2E, ... (E = elements per 16 bytes) all start at the SAME offset within a 16-byte chunk, because E rows
// are a whole number of chunks.  So the rows are dealt into E classes by r mod E and every class is an ordinary vector column
// reduce over rows E * pitch apart: a thread owns one ALIGNED chunk position and adds the chunks of its class's rows into E
// accumulators with 128-bit loads; the chunk's first element is column E * k - A, where A (the class's misalignment) is a
// per-CTA constant.  Each (row segment, class) writes one partial row, indexed by column; columns < 0 or >= ncols, which
// belong to the neighbouring rows, are simply not written.  The finish kernel folds the partial rows per column in a fixed
// order (deterministic) and performs the final store, fused statistics included.
#include "k_reduce.cuh"

template <class T, int OP, int E>
__global__ void __launch_bounds__(256) reduce_col_rot_kernel(const __grid_constant__ RedParams p) {
  constexpr int SB = 16 / E;                                               // bytes per slot
  const int phi = blockIdx.y % E; const long long seg = blockIdx.y / E;
  const long long rbeg = seg * p.seg + phi, rend = (seg + 1) * p.seg < p.r_ext ? (seg + 1) * p.seg : p.r_ext;
  if (rbeg >= rend) return;
  const unsigned long long first = (unsigned long long)(uintptr_t)p.in / SB + (unsigned long long)(p.in_off + rbeg * p.r_stride);
  const int A = (int)(first % E);                                          // elements of the previous row in front of column 0
  const long long k = (long long)blockIdx.x * 256 + threadIdx.x;
  if (k * E - A >= p.ncols) return;
  const char* src = p.in + (p.in_off + rbeg * p.r_stride - A + k * E) * SB;                    // 16-byte aligned
  const long long step = p.r_stride * E * SB;                                                 // bytes between rows of one class
  const long long nrows = (rend - rbeg + E - 1) / E;
  const unsigned fl = lk_flags(p.lk);
  T acc[E];
  convert_raw16<T, E>(load_raw16(src), fl, acc);
  long long t = 1;
  constexpr int UR = 8;
  for (; t + UR - 1 < nrows; t += UR) {
    uint4 w[UR];
#pragma unroll
    for (int u = 0; u < UR; u++) w[u] = load_raw16(src + (t + u) * step);
#pragma unroll
    for (int u = 0; u < UR; u++) {
      T v[E]; convert_raw16<T, E>(w[u], fl, v);
#pragma unroll
      for (int i = 0; i < E; i++) acc[i] = Red<T, OP>::f(acc[i], v[i]);
    }
  }
  if (t < nrows) {                                                         // fewer than UR rows left: issued together
    uint4 w[UR - 1]; bool ok[UR - 1];
#pragma unroll
    for (int u = 0; u < UR - 1; u++) { ok[u] = t + u < nrows; w[u] = load_raw16(src + (ok[u] ? t + u : t) * step); }
#pragma unroll
    for (int u = 0; u < UR - 1; u++) if (ok[u]) {
      T v[E]; convert_raw16<T, E>(w[u], fl, v);
#pragma unroll
      for (int i = 0; i < E; i++) acc[i] = Red<T, OP>::f(acc[i], v[i]);
    }
  }
  T* part = (T*)p.partial + (long long)blockIdx.y * p.ncols;
#pragma unroll
  for (int i = 0; i < E; i++) { const long long c = k * E - A + i; if (c >= 0 && c < p.ncols) part[c] = acc[i]; }
}

template <class T, int OP, int E>
__global__ void __launch_bounds__(256) reduce_col_rot_finish_kernel(const __grid_constant__ RedParams p) {
  const long long c = (long long)blockIdx.x * 256 + threadIdx.x;
  if (c >= p.ncols) return;
  const T* part = (const T*)p.partial + c;
  // partial rows in flight eight at a time (one dependent load per row cost 9.5 us for 33 CTAs in the aligned kernel's finish)
  T a; bool have = false;
  const int nrow = p.nseg * E;
  for (int y0 = 0; y0 < nrow; y0 += 8) {
    T v[8]; bool ok[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int y = y0 + u;
      ok[u] = y < nrow && (long long)(y / E) * p.seg + (y % E) < ((long long)(y / E + 1) * p.seg < p.r_ext ? (long long)(y / E + 1) * p.seg : p.r_ext);
      v[u] = part[(long long)(ok[u] ? y : 0) * p.ncols];
    }
#pragma unroll
    for (int u = 0; u < 8; u++) if (ok[u]) { a = have ? Red<T, OP>::f(a, v[u]) : v[u]; have = true; }
  }
  const T init = p.fresh ? ident_of<T>(p) : load_scalar<T>(p.out, p.out_off + c * p.kout[0], p.out_lk);
  store_result<T, T>(p, p.out_off + c * p.kout[0], have ? Red<T, OP>::f(init, a) : init);
}

template <class T, int E> static void rot_pick(int op, RedKernel* main, RedKernel* fin) {
  switch (op) {
    case R_ADD: *main = (RedKernel)reduce_col_rot_kernel<T, R_ADD, E>; *fin = (RedKernel)reduce_col_rot_finish_kernel<T, R_ADD, E>; break;
    case R_MUL: *main = (RedKernel)reduce_col_rot_kernel<T, R_MUL, E>; *fin = (RedKernel)reduce_col_rot_finish_kernel<T, R_MUL, E>; break;
    case R_MAX: *main = (RedKernel)reduce_col_rot_kernel<T, R_MAX, E>; *fin = (RedKernel)reduce_col_rot_finish_kernel<T, R_MAX, E>; break;
    default:    *main = (RedKernel)reduce_col_rot_kernel<T, R_MIN, E>; *fin = (RedKernel)reduce_col_rot_finish_kernel<T, R_MIN, E>; break;
  }
}
void red_col_rot_pick(int cls, int slot_bits, int op, RedKernel* main, RedKernel* fin) {
  if (cls == CLS_F) rot_pick<float, 4>(op, main, fin);
  else if (cls == CLS_D) rot_pick<double, 2>(op, main, fin);
  else if (slot_bits == 32) rot_pick<long long, 4>(op, main, fin);
  else rot_pick<long long, 2>(op, main, fin);
}
