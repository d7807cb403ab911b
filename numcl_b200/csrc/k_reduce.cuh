// k_reduce.cuh -- K3 (kernels; the host side is k_reduce.cu, the instantiations are split over k_reduce_*.cu): axis reductions (replaces reduce-lambda's nest, src/5reduce.lisp:67-87, and
// einsum specs whose only work is a reduction, e.g. `(ij -> i)`, `(i ->)`), plus the contractions that are
// really reductions (`(i i ->)` vdot / inner, `(ij ij -> i)`, `(ij j -> i)`: DOT variants, product formed on load).
//
// All of them are HBM-bound streams over the input (1 GiB at config C2), read exactly once, with 128-bit loads and
// several independent loads in flight per thread.  The basic forms:
//   reduce_row  : kept x reduced with the reduced axis contiguous   ((sum a :axes 1))
//                 256 / 32 / 4 threads per row by row length, warp-shuffle + shared-memory tree, one store per row;
//   reduce_all  : everything reduced ((sum a)): persistent grid, per-thread accumulators -> float64/int64 block
//                 partials -> the last block to finish (ticket) folds the partials in a fixed order: deterministic;
//   reduce_col  : reduced x kept with the kept axis contiguous ((sum a :axes 0)): one thread per 16-byte column
//                 chunk walking down the reduced axis; when there are too few columns to fill the GPU the reduced
//                 axis is split into one resident wave of slabs and a second pass folds them.
// Shapes those forms handle badly have their own paths ("Shapes off the headline configs"):
//   reduce_rows_reg     rows of 2..8 elements, back to back: E rows per thread, entirely in registers;
//   reduce_row_thread   rows of at most 4 chunks / 16 elements: one thread per row, four rows per thread;
//   reduce_rows_staged  ragged rows up to 2 KB, back to back: 32 KB tiles through shared memory;
//   few long rows       host split [rows x S x L] -> scratch [rows x S] -> result (three launches);
//   reduce_fewcols      <= 128 columns, back to back: the array as ONE stream, accumulators with fixed column phase;
//   dot_matvec          matrix x vector: four rows per CTA, every chunk of the vector loaded once per four rows;
//   8-bit `+`           a chunk's 16 bytes are summed with four dp4a.
// Accumulation order differs from the reference's sequential row-major scan (5reduce.lisp:78-87) except where one
// thread folds a whole row: bit-exact for integers (wrap-around add/mul are associative) and for exactly representable
// float sums, within tolerance otherwise (north_star: 1e-5 single, 1e-12 double).
#pragma once
#include "ncl_device.cuh"

enum { R_ADD = 0, R_MUL, R_MAX, R_MIN };

template <class T, int OP> struct Red {
  static __device__ __forceinline__ T f(T a, T b) {
    if constexpr (OP == R_ADD) return Ops<T>::add(a, b);
    else if constexpr (OP == R_MUL) return Ops<T>::mul(a, b);
    else if constexpr (OP == R_MAX) return Ops<T>::mx(a, b);
    else return Ops<T>::mn(a, b);
  }
};
// neutral element of the fold, for accumulators that have seen nothing yet.  max / min: the lowest / highest value of the lane
// type ((max -inf x) = x for every x, NaN included).  They used to be seeded with the row's FIRST ELEMENT: one scalar load,
// behind an element-kind switch, in front of the chunk loads -- two memory latencies in a row for a row that is a single
// unrolled step long (amax over rows of 512 floats: 0.79 of the HBM peak where the sum of the same rows runs at 1.06).
template <class T, int OP> __device__ __forceinline__ T neutral() {
  if constexpr (OP == R_ADD) return (T)0;
  else if constexpr (OP == R_MUL) return (T)1;
  else if constexpr (std::is_same<T, float>::value) return OP == R_MAX ? -__int_as_float(0x7f800000) : __int_as_float(0x7f800000);
  else if constexpr (std::is_same<T, double>::value) return OP == R_MAX ? -__longlong_as_double(0x7ff0000000000000LL) : __longlong_as_double(0x7ff0000000000000LL);
  else if constexpr (sizeof(T) == 8) return OP == R_MAX ? (T)(-0x7fffffffffffffffLL - 1) : (T)0x7fffffffffffffffLL;
  else return OP == R_MAX ? (T)(-0x7fffffff - 1) : (T)0x7fffffff;
}

struct RedParams {
  const char* in; long long in_off; int lk;
  char* out; long long out_off; StoreSpec sk; int out_lk;
  int fresh; long long ident_i; double ident_f;
  int stat; long long count;                       // fused mean / variance / stdev epilogue (store_result): 0 none, 1 / count, 3 sqrt(/ count)
  // row: rows = prod(kext), cols contiguous
  int nk; long long kext[4], kin[4], kout[4];      // kept dims (outer -> inner): extents, input strides, output strides
  long long rows, cols;
  int vec;                                         // 1: every row start is 16B aligned and cols % E == 0
  // col: [kout...][R][kin chunk]
  long long r_ext, r_stride; long long ncols;      // ncols = contiguous kept extent
  long long seg; int nseg; void* partial;          // split of the reduced axis
  // all
  void* block_partials; unsigned int* ticket; int nblocks;
  void* raw_partial;                               // when set: write the f64/i64 partial here and skip the final store
  // fused all-reduce over NVLink peer memory (sharded full reduce): every rank's mailbox is mapped
  // into every other rank; the LAST block of this kernel pushes the local partial into all peers'
  // mailboxes, waits for theirs, folds them in rank order and stores the final value itself.
  unsigned long long* const* peers; int rank, world; unsigned long long seq;
  // second operand of a contraction that is really a reduction (DOT kernels: r[kept] += a[..] * b[..]; vdot, inner,
  // row-wise dot, matrix x vector).  b may be broadcast (stride 0) along any kept loop and along the reduced one.
  const char* in2; long long in2_off; int lk2;
  long long kin2[4];                               // strides of b along the outer kept dims
  long long r_stride2;                             // stride of b along the reduced loop (row form: 0 or 1)
  long long c_stride2;                             // col form: stride of b along the contiguous kept loop (0 or 1)
};

// element fetch of the reductions: a[i], or the product a[i] * b[i] of the einsum leaf (+ @1 (* $1 $2)) -- each product
// rounded on its own, then added (common-lisp.lisp:226-240; no FMA contraction)
// E > 1: raw 16-byte loads first (all of a step's loads can be issued back to back), conversion afterwards, no branch.
// one element of the slot width that goes with (T, E) -- the kind is again two flags, no switch around the load (the broadcast
// operand of a column-form dot product, `(i ij -> j)`: its switch-wrapped load serialised the loads of the whole step)
template <class T, int E>
__device__ __forceinline__ T load_one_bf(const char* base, long long idx, unsigned kf) {
  if constexpr (std::is_same<T, float>::value) return __ldg((const float*)base + idx);
  else if constexpr (std::is_same<T, double>::value) return __ldg((const double*)base + idx);
  else {
    const bool sg = (kf & 1u) != 0; const int tg = (int)((kf >> 1) & 1u);
    if constexpr (E == 2) return (T)(__ldg((const long long*)base + idx) >> tg);
    else if constexpr (E == 4) { const unsigned w = __ldg((const unsigned*)base + idx); return sg ? (T)(long long)(int)w : (T)(long long)w; }
    else if constexpr (E == 8) { const unsigned short w = __ldg((const unsigned short*)base + idx); return sg ? (T)(long long)(short)w : (T)(long long)w; }
    else { const unsigned char w = __ldg((const unsigned char*)base + idx); return sg ? (T)(long long)(signed char)w : (T)(long long)w; }
  }
}
template <class T, int E> struct RawPair { uint4 a, b; T bs; };
// BC: is the second operand contiguous along the chunk?  -1 = decided at run time (a branch per fetch), 0 / 1 = compile time.
// The streaming loops dispatch ONCE on the run-time flag and run a BC-specialised copy: with the branch inside an unrolled step
// the loads of later chunks could not be hoisted above it and a column-form dot product kept one or two loads in flight
// (`(i ij -> j)`: DRAM 44 % busy, long-scoreboard stalls, profiles/r02_ncu_slow_paths.md).
template <class T, int E, bool DOT, int BC = -1>
__device__ __forceinline__ void fetch_raw(const RedParams& p, long long ia, long long ib, bool b_contig, RawPair<T, E>& r) {
  r.a = load_raw16(p.in + ia * (16 / E));
  if constexpr (DOT) {
    if (BC == 1 || (BC == -1 && b_contig)) r.b = load_raw16(p.in2 + ib * (16 / E)); else r.bs = load_one_bf<T, E>(p.in2, ib, lk_flags(p.lk2));
  }
}
template <class T, int E, bool DOT, int BC = -1>
__device__ __forceinline__ void fetch_convert(const RedParams& p, const RawPair<T, E>& r, bool b_contig, T (&v)[E]) {
  convert_raw16<T, E>(r.a, lk_flags(p.lk), v);
  if constexpr (DOT) {
    T w[E];
    if (BC == 1 || (BC == -1 && b_contig)) convert_raw16<T, E>(r.b, lk_flags(p.lk2), w);
    else {
#pragma unroll
      for (int i = 0; i < E; i++) w[i] = r.bs;
    }
#pragma unroll
    for (int i = 0; i < E; i++) v[i] = Ops<T>::mul(v[i], w[i]);
  }
}
template <class T, int E, bool DOT, int BC = -1>
__device__ __forceinline__ void fetch_elems(const RedParams& p, long long ia, long long ib, bool b_contig, T (&v)[E]) {
  if constexpr (E > 1) {
    RawPair<T, E> r; fetch_raw<T, E, DOT, BC>(p, ia, ib, b_contig, r); fetch_convert<T, E, DOT, BC>(p, r, b_contig, v);
  } else {
    load_elems<T, E>(p.in, ia, p.lk, false, v);
    if constexpr (DOT) {
      T w[E]; load_elems<T, E>(p.in2, ib, p.lk2, false, w);
      v[0] = Ops<T>::mul(v[0], w[0]);
    }
  }
}
// runs body(integral_constant<int, BC>) with BC = 1 (no second operand, or a contiguous one) or 0 (a broadcast one)
template <bool DOT, class F> __device__ __forceinline__ void with_bc(bool b_contig, F body) {
  if constexpr (!DOT) body(std::integral_constant<int, 1>{});
  else { if (b_contig) body(std::integral_constant<int, 1>{}); else body(std::integral_constant<int, 0>{}); }
}
template <class T, int E>
__device__ __forceinline__ void load_vec(const char* base, long long idx, int lk, T (&v)[E]) {
  if constexpr (E > 1) convert_raw16<T, E>(load_raw16(base + idx * (16 / E)), lk_flags(lk), v);
  else v[0] = load_scalar<T>(base, idx, lk);
}
template <class T, bool DOT>
__device__ __forceinline__ T fetch_scalar(const RedParams& p, long long ia, long long ib) {
  T x = load_scalar<T>(p.in, ia, p.lk);
  if constexpr (DOT) x = Ops<T>::mul(x, load_scalar<T>(p.in2, ib, p.lk2));
  return x;
}

typedef void (*RedKernel)(const RedParams);
template <class T> __device__ __forceinline__ T ident_of(const RedParams& p) {
  if constexpr (std::is_integral<T>::value) return (T)p.ident_i; else return (T)p.ident_f;
}

// The final store of a reduction.  stat == 0: through %coerce into the result's element type.  Otherwise the fused
// epilogue of numcl:mean / variance / standard-deviation (src/5reduce.lisp:128-167), which the reference computes as
// (numcl:/ (numcl:sum ...) count) [then numcl:sqrt]: first the rounding the SUM array's own store would have performed
// (single-float / double-float; fixnum = wrap to 63 bits), then ONE division by the element count exactly as broadcast '/
// performs it (float / integer: the count converted by contagion; integer / integer: the exact ratio rounded once to
// single-float), then sqrt.  Bit-identical to the two- and three-pass composition, in one launch.
template <class T, class V>
__device__ __forceinline__ void store_result(const RedParams& p, long long o, V v) {
  if (p.stat == 0) { store_scalar<V>(p.out, o, p.sk, v); return; }
  if constexpr (std::is_same<T, float>::value) {
    float m = __fdiv_rn(to_f32(v), __ll2float_rn(p.count));
    if (p.stat == 3) m = __fsqrt_rn(m);
    store_scalar<float>(p.out, o, p.sk, m);
  } else if constexpr (std::is_same<T, double>::value) {
    double m = __ddiv_rn(to_f64(v), __ll2double_rn(p.count));
    if (p.stat == 3) m = __dsqrt_rn(m);
    store_scalar<double>(p.out, o, p.sk, m);
  } else {
    const long long s = (long long)((unsigned long long)round_to_i64(v) << 1) >> 1;      // the fixnum sum array: 63 bits
    float m = ratio_to_f32(s, p.count);
    if (p.stat == 3) m = __fsqrt_rn(m);
    store_scalar<float>(p.out, o, p.sk, m);
  }
}

// Sum of the 16 bytes of one chunk with four dot-product instructions (8-bit element types under +): the generic path
// unpacks every byte into a 64-bit lane, ~5 instructions per element, and ran at 0.13 of the HBM peak.
template <bool BYTESUM> __device__ __forceinline__ long long chunk_bytesum_raw(const uint4 w, bool is_signed) {
  if constexpr (BYTESUM) {
    // unsigned sum of the 16 bytes with four dp4a; signed bytes: every byte with its top bit set counts 256 too much.
    // No branch: a branch around the dot products kept the loads of an unrolled step from being issued together.
    const unsigned u = __dp4a(w.w, 0x01010101u, __dp4a(w.z, 0x01010101u, __dp4a(w.y, 0x01010101u, __dp4a(w.x, 0x01010101u, 0u))));
    const unsigned H = 0x80808080u;
    const int neg = __popc(w.x & H) + __popc(w.y & H) + __popc(w.z & H) + __popc(w.w & H);
    return (long long)u - (is_signed ? 256ll * neg : 0ll);
  } else return 0;
}
template <bool BYTESUM> __device__ __forceinline__ long long chunk_bytesum(const char* ptr, bool is_signed) {
  if constexpr (BYTESUM) return chunk_bytesum_raw<true>(__ldg((const uint4*)ptr), is_signed); else return 0;
}
template <class T, int OP, int E, bool DOT> constexpr bool use_bytesum() { return std::is_integral<T>::value && OP == R_ADD && E == 16 && !DOT; }

template <class T, int OP, int E>
__device__ __forceinline__ T fold_elems(T acc, const T (&v)[E]) {
#pragma unroll
  for (int i = 0; i < E; i++) acc = Red<T, OP>::f(acc, v[i]);
  return acc;
}

// block-wide combine of one value per thread; result valid in thread 0
template <class T, int OP, int THREADS>
__device__ __forceinline__ T block_combine(T v) {
  __shared__ T sm[THREADS / 32];
  v = warp_reduce(v, [](T a, T b) { return Red<T, OP>::f(a, b); });
  if constexpr (THREADS == 32) return v;
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sm[w] = v;
  __syncthreads();
  if (w == 0) {
    T x = sm[l < THREADS / 32 ? l : 0];
    if (l >= THREADS / 32) x = sm[0];
    // only THREADS/32 lanes hold distinct data; duplicates would corrupt ADD/MUL, so mask them
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      T y = __shfl_down_sync(0xffffffffu, x, o);
      if (l + o < THREADS / 32) x = Red<T, OP>::f(x, y);
    }
    v = x;
  }
  return v;
}

// ---- row reduce: TPR threads cooperate on one row -----------------------------------------------------
// TPR = 1 / 4: short rows, one thread (four threads) per row -- a warp then covers 32 (8) consecutive rows, and with one
// thread per row the fold is the reference's own left-to-right order.  vec == 2: rows whose length or start is not a
// multiple of the 16-byte chunk: scalar head up to the first aligned element, vector body, scalar tail.
template <class T, int OP, int E, int TPR, bool DOT>
__device__ __forceinline__ void reduce_one_row(const RedParams& p, const long long row, const int t) {
  bool live = row < p.rows;
  long long ibase = p.in_off, obase = p.out_off, ibase2 = p.in2_off;
  if (live) {
    if (p.nk == 1) { ibase += row * p.kin[0]; obase += row * p.kout[0]; if constexpr (DOT) ibase2 += row * p.kin2[0]; }
    else { long long r = row; for (int k = p.nk - 1; k >= 0; k--) { long long c = r % p.kext[k]; r /= p.kext[k]; ibase += c * p.kin[k]; obase += c * p.kout[k]; if constexpr (DOT) ibase2 += c * p.kin2[k]; } }
  }
  const long long rs2 = DOT ? p.r_stride2 : 0; const bool bc = rs2 == 1;
  T acc[4];
  const T seed = neutral<T, OP>();
#pragma unroll
  for (int u = 0; u < 4; u++) acc[u] = seed;
  if (live) {
    if (p.vec && E > 1) {
      long long head = 0;
      if (p.vec == 2) { head = (E - (ibase & (E - 1))) & (E - 1); if (head > p.cols) head = p.cols; }   // the base vector is 16-byte aligned
      const long long nch = (p.cols - head) / E, vbase = ibase + head, vbase2 = ibase2 + head * rs2;
      long long c = t;
      if constexpr (use_bytesum<T, OP, E, DOT>()) {
        const bool sg = p.lk == LK_S8;
        for (; c + 3 * TPR < nch; c += 4 * TPR) {
          uint4 w[4];
#pragma unroll
          for (int u = 0; u < 4; u++) w[u] = load_raw16(p.in + vbase + (c + u * TPR) * E);
#pragma unroll
          for (int u = 0; u < 4; u++) acc[u] += (T)chunk_bytesum_raw<true>(w[u], sg);
        }
        for (; c < nch; c += TPR) acc[0] += (T)chunk_bytesum<true>(p.in + vbase + c * E, sg);
      } else {
      // (plain macro, not a lambda: capturing the accumulators by reference cost the short-row variants half their speed)
#define NCL_ROW_MAIN_LOOP(BC)                                                                                                   \
      for (; c + 3 * TPR < nch; c += 4 * TPR) {                                                                                \
        T v[4][E];                                                                                                             \
        _Pragma("unroll") for (int u = 0; u < 4; u++)                                                                          \
          fetch_elems<T, E, DOT, BC>(p, vbase + (c + u * TPR) * E, vbase2 + (c + u * TPR) * E * rs2, bc, v[u]);                \
        _Pragma("unroll") for (int u = 0; u < 4; u++) acc[u] = fold_elems<T, OP, E>(acc[u], v[u]);                             \
      }
      if constexpr (!DOT) { NCL_ROW_MAIN_LOOP(1) } else if (bc) { NCL_ROW_MAIN_LOOP(1) } else { NCL_ROW_MAIN_LOOP(0) }
#undef NCL_ROW_MAIN_LOOP
      if constexpr (DOT) {                                  // (two operands: the batch below doubled the register count)
        for (; c < nch; c += TPR) { T v[E]; fetch_elems<T, E, DOT>(p, vbase + c * E, vbase2 + c * E * rs2, bc, v); acc[0] = fold_elems<T, OP, E>(acc[0], v); }
      } else if (c < nch) {                                 // up to three chunks left for this thread: issued together
        T v[3][E]; bool ok[3];
#pragma unroll
        for (int u = 0; u < 3; u++) { ok[u] = c + u * TPR < nch; const long long cc = ok[u] ? c + u * TPR : c; fetch_elems<T, E, DOT>(p, vbase + cc * E, vbase2 + cc * E * rs2, bc, v[u]); }
#pragma unroll
        for (int u = 0; u < 3; u++) if (ok[u]) acc[u] = fold_elems<T, OP, E>(acc[u], v[u]);
      }
      }
      if (p.vec == 2) {
        const long long t0 = head + nch * E, extra = head + (p.cols - t0);       // fewer than 2E scalar elements
        for (long long j = t; j < extra; j += TPR) { const long long e = j < head ? j : t0 + (j - head); acc[1] = Red<T, OP>::f(acc[1], fetch_scalar<T, DOT>(p, ibase + e, ibase2 + e * rs2)); }
      }
    } else {
      for (long long c = t; c < p.cols; c += TPR) acc[0] = Red<T, OP>::f(acc[0], fetch_scalar<T, DOT>(p, ibase + c, ibase2 + c * rs2));
    }
  }
  T a = Red<T, OP>::f(Red<T, OP>::f(acc[0], acc[1]), Red<T, OP>::f(acc[2], acc[3]));
  if constexpr (TPR <= 32) {
#pragma unroll
    for (int o = TPR / 2; o > 0; o >>= 1) a = Red<T, OP>::f(a, __shfl_xor_sync(0xffffffffu, a, o));
  } else {
    a = block_combine<T, OP, 256>(a);
  }
  pdl_wait();                                                                // nothing is written before every earlier kernel has completed
  if (live && t == 0) {
    T init = p.fresh ? ident_of<T>(p) : load_scalar<T>(p.out, obase, p.out_lk);
    store_result<T, T>(p, obase, Red<T, OP>::f(init, a));                  // r <- fn(r, partial)
  }
}
// 256 / TPR rows per CTA.  (Persistent warps walking the short rows with a grid stride were tried for TPR <= 32: no gain,
// 0.64 -> 0.61 on rows of 512 floats.)
// DOT: two operand streams doubled the live registers (125 -> two CTAs per SM, 0.53 of the HBM peak on `(ij ij -> i)`,
// profiles/r02_ncu_slow_paths.md); four resident CTAs are worth a few spills outside the streaming loop.
// (0 = unspecified for the single-operand variants: with min-blocks 1 ptxas grew them from 46 to 96 registers, 0.96 -> 0.62 on ragged rows.)
template <class T, int OP, int E, int TPR, bool DOT>
__global__ void __launch_bounds__(256, DOT ? 4 : 0) reduce_row_kernel(const __grid_constant__ RedParams p) {
  pdl_trigger();                                                             // the next kernel may start filling the SMs this one leaves
  reduce_one_row<T, OP, E, TPR, DOT>(p, (long long)blockIdx.x * (256 / TPR) + threadIdx.x / TPR, threadIdx.x % TPR);
}

// ---- short rows: one thread per row, four rows per thread ---------------------------------------------------
// A row is at most 16 elements (or 4 chunks).  Thread t of a block folds rows t, t+256, t+512, t+768 of the block's 1024,
// left to right (the reference's own order); the four rows' loads are issued together, and a warp's loads of one step
// cover 32 adjacent rows, so every sector fetched is consumed through L1 by the following steps.
template <class T, int OP, int E, bool DOT>
__global__ void __launch_bounds__(256) reduce_row_thread_kernel(const __grid_constant__ RedParams p) {
  long long ib[4], ob[4], ib2[4]; bool live[4]; T acc[4];
  const long long rs2 = DOT ? p.r_stride2 : 0; const bool bc = rs2 == 1;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const long long row = (long long)blockIdx.x * 1024 + j * 256 + threadIdx.x;
    live[j] = row < p.rows; ib[j] = p.in_off; ob[j] = p.out_off; ib2[j] = p.in2_off;      // dead rows read row 0: no branch around the loads
    if (live[j]) {
      if (p.nk == 1) { ib[j] += row * p.kin[0]; ob[j] += row * p.kout[0]; if constexpr (DOT) ib2[j] += row * p.kin2[0]; }
      else { long long r = row; for (int k = p.nk - 1; k >= 0; k--) { long long c = r % p.kext[k]; r /= p.kext[k]; ib[j] += c * p.kin[k]; ob[j] += c * p.kout[k]; if constexpr (DOT) ib2[j] += c * p.kin2[k]; } }
    }
    acc[j] = (OP == R_ADD || OP == R_MUL) ? neutral<T, OP>() : (p.cols > 0 ? load_scalar<T>(p.in, ib[j], p.lk) : (T)0);
  }
  if (p.vec && E > 1) {
    const int nch = (int)(p.cols / E);
    for (int c = 0; c < nch; c++) {
      T v[4][E];
#pragma unroll
      for (int j = 0; j < 4; j++) fetch_elems<T, E, DOT>(p, ib[j] + (long long)c * E, ib2[j] + (long long)c * E * rs2, bc, v[j]);
#pragma unroll
      for (int j = 0; j < 4; j++) acc[j] = fold_elems<T, OP, E>(acc[j], v[j]);
    }
  } else {
    const int nc = (int)p.cols;
#pragma unroll 4
    for (int c = 0; c < nc; c++) {
      T v[4];
#pragma unroll
      for (int j = 0; j < 4; j++) v[j] = fetch_scalar<T, DOT>(p, ib[j] + c, ib2[j] + c * rs2);
#pragma unroll
      for (int j = 0; j < 4; j++) acc[j] = Red<T, OP>::f(acc[j], v[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; j++) if (live[j]) {
    T init = p.fresh ? ident_of<T>(p) : load_scalar<T>(p.out, ob[j], p.out_lk);
    store_result<T, T>(p, ob[j], Red<T, OP>::f(init, acc[j]));
  }
}

// ---- matrix x vector: r[i] += sum_j a[i,j] * y[j] ------------------------------------------------------------------
// One CTA per row re-reads the whole of y for every row: twice the bytes through L2 (measured 0.38 of the HBM peak where
// the plain row sum reaches 0.94).  Here a CTA takes RB = 4 rows at once and every y chunk it loads is used for all four.
template <class T, int E>
__global__ void __launch_bounds__(256) dot_matvec_kernel(const __grid_constant__ RedParams p) {
  constexpr int RB = 4;
  __shared__ T sm[8][RB];
  const long long row0 = (long long)blockIdx.x * RB;
  long long ib[RB];
#pragma unroll
  for (int j = 0; j < RB; j++) ib[j] = p.in_off + (row0 + j < p.rows ? row0 + j : row0) * p.kin[0];    // dead rows re-read row0
  T acc[RB];
#pragma unroll
  for (int j = 0; j < RB; j++) acc[j] = (T)0;
  const long long nch = p.cols / E;
  long long c = threadIdx.x;
  for (; c + 256 < nch; c += 512) {
    T y[2][E], a[2][RB][E];
#pragma unroll
    for (int u = 0; u < 2; u++) {
      load_vec<T, E>(p.in2, p.in2_off + (c + u * 256) * E, p.lk2, y[u]);
#pragma unroll
      for (int j = 0; j < RB; j++) load_vec<T, E>(p.in, ib[j] + (c + u * 256) * E, p.lk, a[u][j]);
    }
#pragma unroll
    for (int u = 0; u < 2; u++)
#pragma unroll
      for (int j = 0; j < RB; j++)
#pragma unroll
        for (int i = 0; i < E; i++) acc[j] = Ops<T>::add(acc[j], Ops<T>::mul(a[u][j][i], y[u][i]));
  }
  for (; c < nch; c += 256) {
    T y[E], a[RB][E];
    load_vec<T, E>(p.in2, p.in2_off + c * E, p.lk2, y);
#pragma unroll
    for (int j = 0; j < RB; j++) load_vec<T, E>(p.in, ib[j] + c * E, p.lk, a[j]);
#pragma unroll
    for (int j = 0; j < RB; j++)
#pragma unroll
      for (int i = 0; i < E; i++) acc[j] = Ops<T>::add(acc[j], Ops<T>::mul(a[j][i], y[i]));
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < RB; j++) {
    T v = warp_reduce(acc[j], [](T x, T z) { return Ops<T>::add(x, z); });
    if (l == 0) sm[w][j] = v;
  }
  __syncthreads();
  if (threadIdx.x < RB && row0 + threadIdx.x < p.rows) {
    T tot = sm[0][threadIdx.x];
#pragma unroll
    for (int k = 1; k < 8; k++) tot = Ops<T>::add(tot, sm[k][threadIdx.x]);
    const long long ob = p.out_off + (row0 + threadIdx.x) * p.kout[0];
    T init = p.fresh ? ident_of<T>(p) : load_scalar<T>(p.out, ob, p.out_lk);
    store_result<T, T>(p, ob, Ops<T>::add(init, tot));
  }
}

// ---- rows of 2..8 elements, back to back: entirely in registers ---------------------------------------------------------
// A thread owns E = (elements per 16-byte chunk) consecutive rows = C whole chunks: C 128-bit loads in flight, every element
// index a compile-time constant, each row folded left to right (the reference's own order), E adjacent results stored.
// No shared memory, no barrier, no idle lanes whatever C is ((sum a :axes 1) of an (N,3) array: 0.56 staged -> this).
template <class T, int OP, int E, int C>
__global__ void __launch_bounds__(256) reduce_rows_reg_kernel(const __grid_constant__ RedParams p) {
  const long long row0 = ((long long)blockIdx.x * 256 + threadIdx.x) * E;
  if (row0 >= p.rows) return;
  if (row0 + E <= p.rows) {
    T a[C][E];
#pragma unroll
    for (int k = 0; k < C; k++) load_vec<T, E>(p.in, p.in_off + row0 * C + k * E, p.lk, a[k]);
#pragma unroll
    for (int j = 0; j < E; j++) {
      T r = a[(j * C) / E][(j * C) % E];
#pragma unroll
      for (int i = 1; i < C; i++) r = Red<T, OP>::f(r, a[(j * C + i) / E][(j * C + i) % E]);
      const long long ob = p.out_off + (row0 + j) * p.kout[0];
      T init = p.fresh ? ident_of<T>(p) : load_scalar<T>(p.out, ob, p.out_lk);
      store_result<T, T>(p, ob, Red<T, OP>::f(init, r));
    }
  } else {                                                  // fewer than E rows left
    for (long long row = row0; row < p.rows; row++) {
      T r = load_scalar<T>(p.in, p.in_off + row * C, p.lk);
      for (int i = 1; i < C; i++) r = Red<T, OP>::f(r, load_scalar<T>(p.in, p.in_off + row * C + i, p.lk));
      const long long ob = p.out_off + row * p.kout[0];
      T init = p.fresh ? ident_of<T>(p) : load_scalar<T>(p.out, ob, p.out_lk);
      store_result<T, T>(p, ob, Red<T, OP>::f(init, r));
    }
  }
}
template <class T, int OP, int E> static RedKernel rows_reg_pick_c(int c) {
  switch (c) {
    case 2: return (RedKernel)reduce_rows_reg_kernel<T, OP, E, 2>; case 3: return (RedKernel)reduce_rows_reg_kernel<T, OP, E, 3>;
    case 4: return (RedKernel)reduce_rows_reg_kernel<T, OP, E, 4>; case 5: return (RedKernel)reduce_rows_reg_kernel<T, OP, E, 5>;
    case 6: return (RedKernel)reduce_rows_reg_kernel<T, OP, E, 6>; case 7: return (RedKernel)reduce_rows_reg_kernel<T, OP, E, 7>;
    case 8: return (RedKernel)reduce_rows_reg_kernel<T, OP, E, 8>; default: return nullptr;
  }
}
template <class T, int E> static RedKernel rows_reg_pick(int op, int c) {
  switch (op) { case R_ADD: return rows_reg_pick_c<T, R_ADD, E>(c); case R_MUL: return rows_reg_pick_c<T, R_MUL, E>(c);
    case R_MAX: return rows_reg_pick_c<T, R_MAX, E>(c); default: return rows_reg_pick_c<T, R_MIN, E>(c); }
}

// ---- short / ragged rows of a contiguous array: staged through shared memory -----------------------------------------
// When the rows lie back to back the whole input is one stream, whatever the row length.  A CTA copies a tile of p.seg
// consecutive rows (<= 32 KB, a whole number of 16-byte chunks) into shared memory with coalesced 128-bit loads -- up to
// eight in flight per thread, all issued before the first is used -- and then folds the rows out of shared memory: p.nseg
// (a power of two <= 32) lanes per row read consecutive elements, so rows of 3, 48 or 100 elements all keep the memory
// system busy (one warp per 400-byte row reached 0.26 of the HBM peak).  One lane per row folds left to right, the
// reference's own order.
template <class T, int OP>
__global__ void __launch_bounds__(256) reduce_rows_staged_kernel(const __grid_constant__ RedParams p) {
  __shared__ uint4 tile[2048];
  const int sbytes = lk_bits(p.lk) >> 3, tpr = p.nseg, groups = 256 / tpr;
  const int gid = threadIdx.x / tpr, l = threadIdx.x % tpr;
  const long long RT = p.seg, ntiles = (p.rows + RT - 1) / RT;
  const uint4* src0 = (const uint4*)(p.in + p.in_off * sbytes);          // 16-byte aligned (host check)
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long long r0 = t * RT;
    const int nrows = (int)(p.rows - r0 < RT ? p.rows - r0 : RT);
    const int nch = (int)(((long long)nrows * p.cols * sbytes + 15) >> 4);
    const uint4* src = src0 + ((r0 * p.cols * sbytes) >> 4);               // RT rows are a whole number of chunks
    uint4 v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { const int c = threadIdx.x + k * 256; v[k] = __ldg(src + (c < nch ? c : 0)); }
#pragma unroll
    for (int k = 0; k < 8; k++) { const int c = threadIdx.x + k * 256; if (c < nch) tile[c] = v[k]; }
    __syncthreads();
    const char* sm = (const char*)tile;
    const int ncols = (int)p.cols;
    dispatch_lk(p.lk, [&](auto K) {                        // element kind resolved once, outside the element loops
      constexpr int LK = decltype(K)::value;
      for (int it = 0; it * groups < nrows; it++) {
        const int row = it * groups + gid;
        const bool live = row < nrows;
        const int e0 = (live ? row : 0) * ncols;
        T acc = (OP == R_ADD || OP == R_MUL) ? neutral<T, OP>() : load_scalar_k<T, LK>(sm, e0);
#pragma unroll 4
        for (int c = l; c < ncols; c += tpr) acc = Red<T, OP>::f(acc, load_scalar_k<T, LK>(sm, e0 + c));
        for (int o = tpr >> 1; o > 0; o >>= 1) acc = Red<T, OP>::f(acc, __shfl_xor_sync(0xffffffffu, acc, o));
        if (live && l == 0) {
          const long long ob = p.out_off + (r0 + row) * p.kout[0];
          T init = p.fresh ? ident_of<T>(p) : load_scalar<T>(p.out, ob, p.out_lk);
          store_result<T, T>(p, ob, Red<T, OP>::f(init, acc));
        }
      }
    });
    __syncthreads();
  }
}
template <class T> static RedKernel staged_pick(int op) {
  switch (op) { case R_ADD: return (RedKernel)reduce_rows_staged_kernel<T, R_ADD>; case R_MUL: return (RedKernel)reduce_rows_staged_kernel<T, R_MUL>;
    case R_MAX: return (RedKernel)reduce_rows_staged_kernel<T, R_MAX>; default: return (RedKernel)reduce_rows_staged_kernel<T, R_MIN>; }
}

// ---- full reduce -------------------------------------------------------------------------------------------
template <class T> struct Wide { typedef double W; };
template <> struct Wide<long long> { typedef long long W; };

// 48 registers -> five resident CTAs per SM: the stream needs the occupancy (0.93 -> 1.00 of the measured HBM peak
// on C2b); narrow integer inputs unpack 4..16 elements per chunk into 64-bit lanes and keep the default budget.
template <class T, int E> constexpr int reduce_all_min_blocks() { return (sizeof(T) * E <= 16) ? 5 : 1; }
template <class T, int OP, int E, bool DOT>
__global__ void __launch_bounds__(256, reduce_all_min_blocks<T, E>()) reduce_all_kernel(const __grid_constant__ RedParams p) {
  typedef typename Wide<T>::W W;
  pdl_trigger();
  const T seed = neutral<T, OP>();                   // also what an EMPTY slab (a rank that owns no rows of a sharded array) contributes
  T acc[4];
#pragma unroll
  for (int u = 0; u < 4; u++) acc[u] = seed;
  long long tid = (long long)blockIdx.x * 256 + threadIdx.x, nthr = (long long)gridDim.x * 256;
  if (p.vec && E > 1) {
    long long nch = p.cols / E;
    long long c = tid;
    if constexpr (use_bytesum<T, OP, E, DOT>()) {
      const bool sg = p.lk == LK_S8;
      for (; c + 7 * nthr < nch; c += 8 * nthr) {                   // a byte stream needs twice the chunks in flight of a float stream
        uint4 w[8];
#pragma unroll
        for (int u = 0; u < 8; u++) w[u] = load_raw16(p.in + p.in_off + (c + u * nthr) * E);
#pragma unroll
        for (int u = 0; u < 8; u++) acc[u & 3] += (T)chunk_bytesum_raw<true>(w[u], sg);
      }
      for (; c < nch; c += nthr) acc[0] += (T)chunk_bytesum<true>(p.in + p.in_off + c * E, sg);
    } else {
    for (; c + 3 * nthr < nch; c += 4 * nthr) {
      T v[4][E];
#pragma unroll
      for (int u = 0; u < 4; u++) fetch_elems<T, E, DOT>(p, p.in_off + (c + u * nthr) * E, p.in2_off + (c + u * nthr) * E, true, v[u]);
#pragma unroll
      for (int u = 0; u < 4; u++) acc[u] = fold_elems<T, OP, E>(acc[u], v[u]);
    }
    for (; c < nch; c += nthr) { T v[E]; fetch_elems<T, E, DOT>(p, p.in_off + c * E, p.in2_off + c * E, true, v); acc[0] = fold_elems<T, OP, E>(acc[0], v); }
    }
    // fewer than E trailing elements when the length is not a multiple of the chunk (odd-sized arrays)
    const long long t0 = nch * E;
    if (blockIdx.x == 0 && t0 + threadIdx.x < p.cols) acc[1] = Red<T, OP>::f(acc[1], fetch_scalar<T, DOT>(p, p.in_off + t0 + threadIdx.x, p.in2_off + t0 + threadIdx.x));
  } else {
    for (long long c = tid; c < p.cols; c += nthr) acc[0] = Red<T, OP>::f(acc[0], fetch_scalar<T, DOT>(p, p.in_off + c, p.in2_off + c));
  }
  // thread partials leave the element type here: float sums continue in float64
  W a = Red<W, OP>::f(Red<W, OP>::f((W)acc[0], (W)acc[1]), Red<W, OP>::f((W)acc[2], (W)acc[3]));
  a = block_combine<W, OP, 256>(a);
  __shared__ bool last;
  W* partials = (W*)p.block_partials;
  pdl_wait();                                          // scratch, ticket and the result are shared with earlier kernels: they must be done
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = a;
    __threadfence();
    unsigned int done = atomicAdd(p.ticket, 1u);
    last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  W s = (OP == R_ADD) ? (W)0 : (OP == R_MUL) ? (W)1 : ((volatile W*)partials)[0];
  for (int i = threadIdx.x; i < (int)gridDim.x; i += 256) s = Red<W, OP>::f(s, ((volatile W*)partials)[i]);
  s = block_combine<W, OP, 256>(s);
  if (p.world > 1) {
    // ---- compute + collective push in ONE kernel: P2P stores into NVLink-mapped mailboxes -----------------
    // Every rank's mailbox (layout: mbox_* helpers in ncl_device.cuh) is mapped into every other rank.  The last block
    // takes the next sequence number from the rank's own device-resident counter (so the launch is identical every time
    // and can sit in a CUDA graph), and W threads store the partial into all W mailboxes.  A message is two 8-byte words,
    // each {sequence : 32 | half of the value : 32}: every word is published atomically by its own store, so there is no
    // fence and no second "flag" store on the critical path -- the kernel ends right after issuing the stores.
    // Nobody waits here: mailbox_finalize_kernel (k_reduce.cu), one small CTA on the communication stream, collects the W
    // messages, folds them in rank order and stores the result while the compute stream already runs the next kernel.
    __shared__ unsigned long long xseq, xbits;
    if (threadIdx.x == 0) {                                  // the combined partial is valid in thread 0 only
      *p.ticket = 0; xseq = mbox_push_seq(p.peers[p.rank], p.world) + 1ull;
      xbits = (unsigned long long)(std::is_integral<T>::value ? (long long)s : __double_as_longlong((double)s));
    }
    __syncthreads();
    const unsigned long long seq = xseq, bits = xbits;
    if (threadIdx.x < p.world) mbox_send(p.peers[threadIdx.x], p.world, seq, p.rank, bits);
    if (threadIdx.x == 0) mbox_set_push_seq(p.peers[p.rank], p.world, seq);
    return;
  }
  if (threadIdx.x == 0) {
    *p.ticket = 0;                                   // re-arm for the next launch on this stream
    if (p.raw_partial) { *(W*)p.raw_partial = s; return; }
    W init = p.fresh ? (W)ident_of<T>(p) : (W)load_scalar<T>(p.out, p.out_off, p.out_lk);
    W r = Red<W, OP>::f(init, s);
    store_result<T, W>(p, p.out_off, r);                            // one rounding to the element type
  }
}

// ---- few columns: (sum a :axes 0) of an (N,3) array ------------------------------------------------------------------
// The array is ONE contiguous stream of R*ncols elements and a row is far shorter than a warp's worth of chunks, so the
// column form below would keep 3 lanes busy.  Here the stream is read exactly like the full reduce (persistent grid,
// 4 x 16 bytes in flight per thread), with the grid sized so that (threads x E) is a multiple of ncols: every chunk a
// thread visits then starts at the same column, and its E accumulators stay with fixed columns.  Thread partials widen to
// float64 / int64, fold per column inside the block through shared memory, block partials per column go to global
// memory and the last block (ticket) folds them in block order: deterministic, one rounding.
template <class T, int OP, int E>
__global__ void __launch_bounds__(256, reduce_all_min_blocks<T, E>()) reduce_fewcols_kernel(const __grid_constant__ RedParams p) {
  typedef typename Wide<T>::W W;
  __shared__ W S[256 * E];
  __shared__ W S2[256];
  __shared__ bool H2[256];
  __shared__ bool last;
  const int nc = (int)p.ncols;
  const long long tid = (long long)blockIdx.x * 256 + threadIdx.x, nthr = (long long)gridDim.x * 256;
  const long long nch = p.cols / E;                        // p.cols: elements in the stream, a multiple of E
  const int ph0 = (int)((tid * E) % nc);
  T acc[4][E];
#pragma unroll
  for (int i = 0; i < E; i++) {
    // max / min start from the row-0 element of the accumulator's own column: no artificial neutral value
    const T seed = neutral<T, OP>();
#pragma unroll
    for (int u = 0; u < 4; u++) acc[u][i] = seed;
  }
  long long c = tid;
  for (; c + 3 * nthr < nch; c += 4 * nthr) {
    T v[4][E];
#pragma unroll
    for (int u = 0; u < 4; u++) load_vec<T, E>(p.in, p.in_off + (c + u * nthr) * E, p.lk, v[u]);
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int i = 0; i < E; i++) acc[u][i] = Red<T, OP>::f(acc[u][i], v[u][i]);
  }
  for (; c < nch; c += nthr) {
    T v[E]; load_vec<T, E>(p.in, p.in_off + c * E, p.lk, v);
#pragma unroll
    for (int i = 0; i < E; i++) acc[0][i] = Red<T, OP>::f(acc[0][i], v[i]);
  }
#pragma unroll
  for (int i = 0; i < E; i++)
    S[threadIdx.x * E + i] = Red<W, OP>::f(Red<W, OP>::f((W)acc[0][i], (W)acc[1][i]), Red<W, OP>::f((W)acc[2][i], (W)acc[3][i]));
  __syncthreads();
  // fold per column: thread -> (column j, group g); group g takes every ng-th of the block's positions of that column
  const int j = threadIdx.x % nc, g = threadIdx.x / nc, ng = 256 / nc;
  {
    int f0 = j - (int)(((long long)blockIdx.x * 256 * E) % nc); if (f0 < 0) f0 += nc;
    W part = (W)0; bool has = false;
    if (g < ng) for (int f = f0 + g * nc; f < 256 * E; f += ng * nc) { part = has ? Red<W, OP>::f(part, S[f]) : S[f]; has = true; }
    S2[threadIdx.x] = part; H2[threadIdx.x] = has;
  }
  __syncthreads();
  W* partials = (W*)p.block_partials;
  if (threadIdx.x < nc) {
    W tot = S2[j];                                          // group 0 always holds at least one position
    for (int gg = 1; gg < ng; gg++) if (H2[gg * nc + j]) tot = Red<W, OP>::f(tot, S2[gg * nc + j]);
    partials[(long long)blockIdx.x * nc + j] = tot;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) { unsigned int done = atomicAdd(p.ticket, 1u); last = (done == gridDim.x - 1); }
  __syncthreads();
  if (!last) return;
  __threadfence();
  {
    W part = (W)0; bool has = false;
    if (g < ng) for (int b = g; b < (int)gridDim.x; b += ng) { W x = ((volatile W*)partials)[(long long)b * nc + j]; part = has ? Red<W, OP>::f(part, x) : x; has = true; }
    S2[threadIdx.x] = part; H2[threadIdx.x] = has;
  }
  __syncthreads();
  if (threadIdx.x == 0) *p.ticket = 0;                      // re-arm for the next launch on this stream
  if (threadIdx.x < nc) {
    W tot = S2[j];
    for (int gg = 1; gg < ng; gg++) if (H2[gg * nc + j]) tot = Red<W, OP>::f(tot, S2[gg * nc + j]);
    const long long o = p.out_off + (long long)j * p.kout[0];
    W init = p.fresh ? (W)ident_of<T>(p) : (W)load_scalar<T>(p.out, o, p.out_lk);
    W r = Red<W, OP>::f(init, tot);
    store_result<T, W>(p, o, r);                            // one rounding to the element type
  }
}

// ---- column reduce ---------------------------------------------------------------------------------------------
// thread -> one chunk of E contiguous kept columns of one outer-kept coordinate; walks rows [r0, r1)
template <class T, int OP, int E, bool DOT>
__global__ void __launch_bounds__(256) reduce_col_kernel(const __grid_constant__ RedParams p) {
  long long chunks_per_row = p.ncols / E;
  long long q = (long long)blockIdx.x * 256 + threadIdx.x;
  long long nq = p.rows * chunks_per_row;              // rows = prod of the outer kept dims
  if (q >= nq) return;
  // a plain matrix (no outer kept loops) needs no 64-bit division: with a reduced axis of a few rows the two divisions below were
  // most of a thread's instructions ((sum a :axes 0) of a (3, N) array: 48 bytes of loads per thread)
  long long cc = q, orow = 0;
  if (p.nk > 0) { cc = q % chunks_per_row; orow = q / chunks_per_row; }
  long long ibase = p.in_off + cc * E, obase = p.out_off + cc * E, ibase2 = DOT ? p.in2_off + cc * E * p.c_stride2 : 0;
  { long long r = orow; for (int k = p.nk - 1; k >= 0; k--) { long long c = r % p.kext[k]; r /= p.kext[k]; ibase += c * p.kin[k]; obase += c * p.kout[k]; if constexpr (DOT) ibase2 += c * p.kin2[k]; } }
  const long long rs2 = DOT ? p.r_stride2 : 0; const bool bc = DOT && p.c_stride2 == 1;
  long long r0 = (long long)blockIdx.y * p.seg, r1 = r0 + p.seg; if (r1 > p.r_ext) r1 = p.r_ext;
  T acc[E];
  if (r0 >= r1) return;
  // (the first row used to seed the accumulators with a load of its own: for a reduced axis of 2 .. 8 rows that was a second
  // memory latency in front of the batch below -- (sum a :axes 0) of a (3, N) array)
#pragma unroll
  for (int i = 0; i < E; i++) acc[i] = neutral<T, OP>();
  long long r = r0;
  // rows in flight per thread: 4 x 16 bytes on the vector path; the scalar path (ragged column count) moves only one
  // element per load and needs 16 rows to keep a comparable number of bytes in flight (0.17 -> of the HBM peak with 4)
  // (8 rows in flight for 16-byte chunks was tried after ncu showed the kernel latency-bound at 49 % of DRAM bandwidth: worse everywhere,
  // 0.96 -> 0.73 on (1024,512,512) axis 1 and 0.57 -> 0.25 on 8192^2 fixnum -- more rows per thread means more DRAM pages open at once)
  constexpr int UR = E == 1 ? 16 : 4;
#define NCL_COL_MAIN_LOOP(BC)                                                                                                   \
  for (; r + UR - 1 < r1; r += UR) {                                                                                            \
    T v[UR][E];                                                                                                                 \
    _Pragma("unroll") for (int u = 0; u < UR; u++)                                                                              \
      fetch_elems<T, E, DOT, BC>(p, ibase + (r + u) * p.r_stride, ibase2 + (r + u) * rs2, bc, v[u]);                            \
    _Pragma("unroll") for (int u = 0; u < UR; u++)                                                                              \
      _Pragma("unroll") for (int i = 0; i < E; i++) acc[i] = Red<T, OP>::f(acc[i], v[u][i]);                                    \
  }
  if constexpr (!DOT) { NCL_COL_MAIN_LOOP(1) } else if (bc) { NCL_COL_MAIN_LOOP(1) } else { NCL_COL_MAIN_LOOP(0) }
#undef NCL_COL_MAIN_LOOP
  if constexpr (DOT) {
    for (; r < r1; r++) {
      T v[E]; fetch_elems<T, E, DOT>(p, ibase + r * p.r_stride, ibase2 + r * rs2, bc, v);
#pragma unroll
      for (int i = 0; i < E; i++) acc[i] = Red<T, OP>::f(acc[i], v[i]);
    }
  } else if (r < r1) {                                      // fewer than UR rows left: issued together, not one latency each
    constexpr int TR = E == 1 ? 15 : 3;
    T v[TR][E]; bool ok[TR];
#pragma unroll
    for (int u = 0; u < TR; u++) { ok[u] = r + u < r1; const long long rr = ok[u] ? r + u : r; fetch_elems<T, E, DOT>(p, ibase + rr * p.r_stride, ibase2 + rr * rs2, bc, v[u]); }
#pragma unroll
    for (int u = 0; u < TR; u++)
      if (ok[u]) {
#pragma unroll
        for (int i = 0; i < E; i++) acc[i] = Red<T, OP>::f(acc[i], v[u][i]);
      }
  }
  if (p.nseg > 1) {
    T* part = (T*)p.partial + ((long long)blockIdx.y * nq + q) * E;
#pragma unroll
    for (int i = 0; i < E; i++) part[i] = acc[i];
    return;
  }
#pragma unroll
  for (int i = 0; i < E; i++) {
    T init = p.fresh ? ident_of<T>(p) : load_scalar<T>(p.out, obase + i, p.out_lk);
    store_result<T, T>(p, obase + i, Red<T, OP>::f(init, acc[i]));
  }
}
template <class T, int OP, int E>
__global__ void __launch_bounds__(256) reduce_col_finish_kernel(const __grid_constant__ RedParams p) {
  // ONE column per thread (not one chunk of E columns: a thread then walked E x nseg dependent loads -- (sum a :axes 0) of an
  // 8192 x 8192 (unsigned-byte 8) array spent 300 of its 415 us here, 512 threads folding 16 x 147 partials each)
  long long chunks_per_row = p.ncols / E;
  long long c = (long long)blockIdx.x * 256 + threadIdx.x;
  long long nq = p.rows * chunks_per_row;
  if (c >= nq * E) return;
  const long long q = c / E; const int i = (int)(c - q * E);
  long long cc = q, orow = 0;
  if (p.nk > 0) { cc = q % chunks_per_row; orow = q / chunks_per_row; }
  long long obase = p.out_off + cc * E + i;
  { long long r = orow; for (int k = p.nk - 1; k >= 0; k--) { long long c2 = r % p.kext[k]; r /= p.kext[k]; obase += c2 * p.kout[k]; } }
  const T* part = (const T*)p.partial + c;
  const long long pitch = nq * E;
  T a = part[0];
  int s = 1;
  for (; s + 7 < p.nseg; s += 8) {                        // eight segment partials in flight, folded in segment order
    T t[8];
#pragma unroll
    for (int k = 0; k < 8; k++) t[k] = part[(long long)(s + k) * pitch];
#pragma unroll
    for (int k = 0; k < 8; k++) a = Red<T, OP>::f(a, t[k]);
  }
  for (; s < p.nseg; s++) a = Red<T, OP>::f(a, part[(long long)s * pitch]);
  T init = p.fresh ? ident_of<T>(p) : load_scalar<T>(p.out, obase, p.out_lk);
  store_result<T, T>(p, obase, Red<T, OP>::f(init, a));
}

// ---- kernel tables ----------------------------------------------------------------------------------------------
enum { KROW1, KROW4, KROW32, KROW256, KALL, KCOL, KCOLFIN, KFEWCOLS };

template <class T, int OP, int E> static RedKernel kget(int which) {
  switch (which) {
    case KROW1: return (RedKernel)reduce_row_thread_kernel<T, OP, E, false>;
    case KFEWCOLS: if constexpr (E > 1) return (RedKernel)reduce_fewcols_kernel<T, OP, E>; else return nullptr;
    case KROW4: return (RedKernel)reduce_row_kernel<T, OP, E, 4, false>;
    case KROW32: return (RedKernel)reduce_row_kernel<T, OP, E, 32, false>;
    case KROW256: return (RedKernel)reduce_row_kernel<T, OP, E, 256, false>;
    case KALL: return (RedKernel)reduce_all_kernel<T, OP, E, false>;
    case KCOL: return (RedKernel)reduce_col_kernel<T, OP, E, false>;
    default: return (RedKernel)reduce_col_finish_kernel<T, OP, E>;
  }
}
template <class T, int E> static RedKernel kget_dot(int which) {       // r += a * b: the sum only
  switch (which) {
    case KROW1: return (RedKernel)reduce_row_thread_kernel<T, R_ADD, E, true>;
    case KROW4: return (RedKernel)reduce_row_kernel<T, R_ADD, E, 4, true>;
    case KROW32: return (RedKernel)reduce_row_kernel<T, R_ADD, E, 32, true>;
    case KROW256: return (RedKernel)reduce_row_kernel<T, R_ADD, E, 256, true>;
    case KALL: return (RedKernel)reduce_all_kernel<T, R_ADD, E, true>;
    case KCOL: return (RedKernel)reduce_col_kernel<T, R_ADD, E, true>;
    case KCOLFIN: return (RedKernel)reduce_col_finish_kernel<T, R_ADD, E>;
    default: return nullptr;
  }
}
template <class T, int E> static RedKernel kget_op(int op, int which) {
  switch (op) { case R_ADD: return kget<T, R_ADD, E>(which); case R_MUL: return kget<T, R_MUL, E>(which);
    case R_MAX: return kget<T, R_MAX, E>(which); default: return kget<T, R_MIN, E>(which); }
}

// one translation unit per group of instantiations keeps the build parallel (k_reduce_f.cu, k_reduce_i.cu, k_reduce_in.cu)
RedKernel red_pick_fd(int cls, int E, int op, int which);          // float (E = 4, 1) and double (E = 2, 1)
RedKernel red_pick_fd_dot(int cls, int E, int which);
RedKernel red_pick_int_wide(int E, int op, int which);             // 64-bit lanes over 64-bit slots: E = 2, 1
RedKernel red_pick_int_wide_dot(int E, int which);
RedKernel red_pick_int_narrow(int E, int op, int which);           // 64-bit lanes over 8/16/32-bit slots: E = 16, 8, 4
RedKernel red_pick_int_narrow_dot(int E, int which);
RedKernel red_staged_pick(int cls, int op);
RedKernel red_rows_reg_pick(int cls, int slot_bits, int op, int c);
void red_dot_matvec_launch(int cls, unsigned blocks, cudaStream_t stream, const RedParams& p);
void red_col_rot_pick(int cls, int slot_bits, int op, RedKernel* main, RedKernel* fin);     // k_reduce_rot.cu: row pitch not a multiple of 16 bytes
