// k_reduce.cu -- K3: axis reductions (replaces reduce-lambda's nest, src/5reduce.lisp:67-87, and
// einsum specs whose only work is a reduction, e.g. `(ij -> i)`, `(i ->)`).
//
// All three shapes are HBM-bound streams over the input (1 GiB at config C2), read exactly once
// with 128-bit loads, several independent loads in flight per thread:
//   reduce_row  : kept x reduced with the reduced axis contiguous   ((sum a :axes 1))
//                 one row per warp (short rows) or per 256-thread block (long rows), warp-shuffle +
//                 shared-memory tree, one store per row;
//   reduce_all  : everything reduced ((sum a)): persistent grid of sm_count*8 blocks, per-thread
//                 accumulators -> float64/int64 block partials -> the last block to finish (ticket)
//                 folds the partials in a fixed order, so the result is deterministic;
//   reduce_col  : reduced x kept with the kept axis contiguous ((sum a :axes 0)): one thread per
//                 16-byte column chunk walking down the reduced axis; when there are too few
//                 columns to fill the GPU the reduced axis is split and a second pass folds the slabs.
// Accumulation order differs from the reference's sequential row-major scan (5reduce.lisp:78-87):
// bit-exact for integers (wrap-around add/mul are associative) and for exactly representable
// float sums, within tolerance otherwise (north_star: 1e-5 single, 1e-12 double).
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
// neutral element for padding lanes; for max/min it is the first real element seen (passed in)
template <class T, int OP> __device__ __forceinline__ T neutral() {
  if constexpr (OP == R_ADD) return (T)0; else return (T)1;
}

struct RedParams {
  const char* in; long long in_off; int lk;
  char* out; long long out_off; StoreSpec sk; int out_lk;
  int fresh; long long ident_i; double ident_f;
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
};

template <class T> __device__ __forceinline__ T ident_of(const RedParams& p) {
  if constexpr (std::is_integral<T>::value) return (T)p.ident_i; else return (T)p.ident_f;
}

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
template <class T, int OP, int E, int TPR>
__global__ void __launch_bounds__(256) reduce_row_kernel(const __grid_constant__ RedParams p) {
  constexpr int RPB = 256 / TPR;
  long long row = (long long)blockIdx.x * RPB + threadIdx.x / TPR;
  int t = threadIdx.x % TPR;
  bool live = row < p.rows;
  long long ibase = p.in_off, obase = p.out_off;
  if (live) { long long r = row; for (int k = p.nk - 1; k >= 0; k--) { long long c = r % p.kext[k]; r /= p.kext[k]; ibase += c * p.kin[k]; obase += c * p.kout[k]; } }
  // first element seeds max/min so that no artificial neutral value is needed
  T acc[4];
  T seed = (OP == R_ADD || OP == R_MUL) ? neutral<T, OP>() : (live && p.cols > 0 ? load_scalar<T>(p.in, ibase, p.lk) : (T)0);
#pragma unroll
  for (int u = 0; u < 4; u++) acc[u] = seed;
  if (live) {
    if (p.vec && E > 1) {
      long long nch = p.cols / E;
      long long c = t;
      for (; c + 3 * TPR < nch; c += 4 * TPR) {
        T v[4][E];
#pragma unroll
        for (int u = 0; u < 4; u++) load_elems<T, E>(p.in, ibase + (c + u * TPR) * E, p.lk, true, v[u]);
#pragma unroll
        for (int u = 0; u < 4; u++) acc[u] = fold_elems<T, OP, E>(acc[u], v[u]);
      }
      for (; c < nch; c += TPR) { T v[E]; load_elems<T, E>(p.in, ibase + c * E, p.lk, true, v); acc[0] = fold_elems<T, OP, E>(acc[0], v); }
    } else {
      for (long long c = t; c < p.cols; c += TPR) acc[0] = Red<T, OP>::f(acc[0], load_scalar<T>(p.in, ibase + c, p.lk));
    }
  }
  T a = Red<T, OP>::f(Red<T, OP>::f(acc[0], acc[1]), Red<T, OP>::f(acc[2], acc[3]));
  if constexpr (TPR == 32) {
    a = warp_reduce(a, [](T x, T y) { return Red<T, OP>::f(x, y); });
  } else {
    a = block_combine<T, OP, 256>(a);
  }
  if (live && t == 0) {
    T init = p.fresh ? ident_of<T>(p) : load_scalar<T>(p.out, obase, p.out_lk);
    store_scalar<T>(p.out, obase, p.sk, Red<T, OP>::f(init, a));          // r <- fn(r, partial)
  }
}

// ---- full reduce -------------------------------------------------------------------------------------------
template <class T> struct Wide { typedef double W; };
template <> struct Wide<long long> { typedef long long W; };

// 48 registers -> five resident CTAs per SM: the stream needs the occupancy (0.93 -> 1.00 of the measured HBM peak
// on C2b); narrow integer inputs unpack 4..16 elements per chunk into 64-bit lanes and keep the default budget.
template <class T, int E> constexpr int reduce_all_min_blocks() { return (sizeof(T) * E <= 16) ? 5 : 1; }
template <class T, int OP, int E>
__global__ void __launch_bounds__(256, reduce_all_min_blocks<T, E>()) reduce_all_kernel(const __grid_constant__ RedParams p) {
  typedef typename Wide<T>::W W;
  T seed = (OP == R_ADD || OP == R_MUL) ? neutral<T, OP>() : load_scalar<T>(p.in, p.in_off, p.lk);
  T acc[4];
#pragma unroll
  for (int u = 0; u < 4; u++) acc[u] = seed;
  long long tid = (long long)blockIdx.x * 256 + threadIdx.x, nthr = (long long)gridDim.x * 256;
  if (p.vec && E > 1) {
    long long nch = p.cols / E;
    long long c = tid;
    for (; c + 3 * nthr < nch; c += 4 * nthr) {
      T v[4][E];
#pragma unroll
      for (int u = 0; u < 4; u++) load_elems<T, E>(p.in, p.in_off + (c + u * nthr) * E, p.lk, true, v[u]);
#pragma unroll
      for (int u = 0; u < 4; u++) acc[u] = fold_elems<T, OP, E>(acc[u], v[u]);
    }
    for (; c < nch; c += nthr) { T v[E]; load_elems<T, E>(p.in, p.in_off + c * E, p.lk, true, v); acc[0] = fold_elems<T, OP, E>(acc[0], v); }
    // fewer than E trailing elements when the length is not a multiple of the chunk (odd-sized arrays)
    const long long t0 = nch * E;
    if (blockIdx.x == 0 && t0 + threadIdx.x < p.cols) acc[1] = Red<T, OP>::f(acc[1], load_scalar<T>(p.in, p.in_off + t0 + threadIdx.x, p.lk));
  } else {
    for (long long c = tid; c < p.cols; c += nthr) acc[0] = Red<T, OP>::f(acc[0], load_scalar<T>(p.in, p.in_off + c, p.lk));
  }
  // thread partials leave the element type here: float sums continue in float64
  W a = Red<W, OP>::f(Red<W, OP>::f((W)acc[0], (W)acc[1]), Red<W, OP>::f((W)acc[2], (W)acc[3]));
  a = block_combine<W, OP, 256>(a);
  __shared__ bool last;
  W* partials = (W*)p.block_partials;
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
    // ---- compute + collective in ONE kernel: P2P stores / loads on NVLink-mapped mailboxes -------
    // mailbox layout per rank: [parity 2][source rank world]{value bits, sequence}; the parity slot of
    // call n cannot be overwritten by call n+2 before every rank has consumed call n (each rank only
    // reaches n+2 after receiving everybody's n+1, which is sent after n was read).
    __shared__ unsigned long long xch[64];
    if (threadIdx.x == 0) { *p.ticket = 0; xch[63] = (unsigned long long)(std::is_integral<T>::value ? (long long)s : __double_as_longlong((double)s)); }
    __syncthreads();
    const int t = threadIdx.x;
    if (t < p.world) {
      const size_t slot = (((size_t)(p.seq & 1) * p.world) + p.rank) * 2;
      volatile unsigned long long* remote = p.peers[t] + slot;
      remote[0] = xch[63];
      __threadfence_system();
      remote[1] = p.seq;                                                   // publish
      volatile unsigned long long* mine = p.peers[p.rank] + (((size_t)(p.seq & 1) * p.world) + t) * 2;
      long long t0 = clock64();
      unsigned long long seen;
      do {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(mine + 1) : "memory");
        if (seen != p.seq && clock64() - t0 > 20000000000LL) { printf("numcl-cuda: peer %d did not arrive (rank %d)\n", t, p.rank); __trap(); }
      } while (seen != p.seq);
      xch[t] = mine[0];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      W tot;
      if constexpr (std::is_integral<T>::value) tot = (W)(long long)xch[0]; else tot = (W)__longlong_as_double((long long)xch[0]);
      for (int r = 1; r < p.world; r++) {
        W v; if constexpr (std::is_integral<T>::value) v = (W)(long long)xch[r]; else v = (W)__longlong_as_double((long long)xch[r]);
        tot = Red<W, OP>::f(tot, v);                                       // rank order: every rank gets the same bits
      }
      W init = p.fresh ? (W)ident_of<T>(p) : (W)load_scalar<T>(p.out, p.out_off, p.out_lk);
      W r = Red<W, OP>::f(init, tot);
      if constexpr (std::is_integral<T>::value) store_scalar<long long>(p.out, p.out_off, p.sk, r);
      else store_scalar<double>(p.out, p.out_off, p.sk, r);
    }
    return;
  }
  if (threadIdx.x == 0) {
    *p.ticket = 0;                                   // re-arm for the next launch on this stream
    if (p.raw_partial) { *(W*)p.raw_partial = s; return; }
    W init = p.fresh ? (W)ident_of<T>(p) : (W)load_scalar<T>(p.out, p.out_off, p.out_lk);
    W r = Red<W, OP>::f(init, s);
    if constexpr (std::is_integral<T>::value) store_scalar<long long>(p.out, p.out_off, p.sk, r);
    else store_scalar<double>(p.out, p.out_off, p.sk, r);           // one rounding to the element type
  }
}

// ---- column reduce ---------------------------------------------------------------------------------------------
// thread -> one chunk of E contiguous kept columns of one outer-kept coordinate; walks rows [r0, r1)
template <class T, int OP, int E>
__global__ void __launch_bounds__(256) reduce_col_kernel(const __grid_constant__ RedParams p) {
  long long chunks_per_row = p.ncols / E;
  long long q = (long long)blockIdx.x * 256 + threadIdx.x;
  long long nq = p.rows * chunks_per_row;              // rows = prod of the outer kept dims
  if (q >= nq) return;
  long long cc = q % chunks_per_row, orow = q / chunks_per_row;
  long long ibase = p.in_off + cc * E, obase = p.out_off + cc * E;
  { long long r = orow; for (int k = p.nk - 1; k >= 0; k--) { long long c = r % p.kext[k]; r /= p.kext[k]; ibase += c * p.kin[k]; obase += c * p.kout[k]; } }
  long long r0 = (long long)blockIdx.y * p.seg, r1 = r0 + p.seg; if (r1 > p.r_ext) r1 = p.r_ext;
  T acc[E];
  if (r0 >= r1) return;
  load_elems<T, E>(p.in, ibase + r0 * p.r_stride, p.lk, E > 1, acc);     // the first row seeds the accumulators
  long long r = r0 + 1;
  // rows in flight per thread: 4 x 16 bytes on the vector path; the scalar path (ragged column count) moves only one
  // element per load and needs 16 rows to keep a comparable number of bytes in flight (0.17 -> of the HBM peak with 4)
  constexpr int UR = E == 1 ? 16 : 4;
  for (; r + UR - 1 < r1; r += UR) {
    T v[UR][E];
#pragma unroll
    for (int u = 0; u < UR; u++) load_elems<T, E>(p.in, ibase + (r + u) * p.r_stride, p.lk, E > 1, v[u]);
#pragma unroll
    for (int u = 0; u < UR; u++)
#pragma unroll
      for (int i = 0; i < E; i++) acc[i] = Red<T, OP>::f(acc[i], v[u][i]);
  }
  for (; r < r1; r++) {
    T v[E]; load_elems<T, E>(p.in, ibase + r * p.r_stride, p.lk, E > 1, v);
#pragma unroll
    for (int i = 0; i < E; i++) acc[i] = Red<T, OP>::f(acc[i], v[i]);
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
    store_scalar<T>(p.out, obase + i, p.sk, Red<T, OP>::f(init, acc[i]));
  }
}
template <class T, int OP, int E>
__global__ void __launch_bounds__(256) reduce_col_finish_kernel(const __grid_constant__ RedParams p) {
  long long chunks_per_row = p.ncols / E;
  long long q = (long long)blockIdx.x * 256 + threadIdx.x;
  long long nq = p.rows * chunks_per_row;
  if (q >= nq) return;
  long long cc = q % chunks_per_row, orow = q / chunks_per_row;
  long long obase = p.out_off + cc * E;
  { long long r = orow; for (int k = p.nk - 1; k >= 0; k--) { long long c = r % p.kext[k]; r /= p.kext[k]; obase += c * p.kout[k]; } }
  const T* part = (const T*)p.partial;
#pragma unroll
  for (int i = 0; i < E; i++) {
    T a = part[q * E + i];
    for (int s = 1; s < p.nseg; s++) a = Red<T, OP>::f(a, part[((long long)s * nq + q) * E + i]);
    T init = p.fresh ? ident_of<T>(p) : load_scalar<T>(p.out, obase + i, p.out_lk);
    store_scalar<T>(p.out, obase + i, p.sk, Red<T, OP>::f(init, a));
  }
}

// ---- host ---------------------------------------------------------------------------------------------------------
typedef void (*RedKernel)(const RedParams);
enum { KROW32, KROW256, KALL, KCOL, KCOLFIN };

template <class T, int OP, int E> static RedKernel kget(int which) {
  switch (which) {
    case KROW32: return (RedKernel)reduce_row_kernel<T, OP, E, 32>;
    case KROW256: return (RedKernel)reduce_row_kernel<T, OP, E, 256>;
    case KALL: return (RedKernel)reduce_all_kernel<T, OP, E>;
    case KCOL: return (RedKernel)reduce_col_kernel<T, OP, E>;
    default: return (RedKernel)reduce_col_finish_kernel<T, OP, E>;
  }
}
template <class T, int E> static RedKernel kget_op(int op, int which) {
  switch (op) { case R_ADD: return kget<T, R_ADD, E>(which); case R_MUL: return kget<T, R_MUL, E>(which);
    case R_MAX: return kget<T, R_MAX, E>(which); default: return kget<T, R_MIN, E>(which); }
}
static RedKernel kernel_for(int cls, int E, int op, int which) {
  if (cls == CLS_F) return E == 4 ? kget_op<float, 4>(op, which) : kget_op<float, 1>(op, which);
  if (cls == CLS_D) return E == 2 ? kget_op<double, 2>(op, which) : kget_op<double, 1>(op, which);
  switch (E) { case 16: return kget_op<long long, 16>(op, which); case 8: return kget_op<long long, 8>(op, which);
    case 4: return kget_op<long long, 4>(op, which); case 2: return kget_op<long long, 2>(op, which); default: return kget_op<long long, 1>(op, which); }
}

static bool match_reduce(const Expr& e, int* op) {
  if (e.n != 3) return false;
  const XNode& r = e.node[2];
  int o;
  switch (r.op) { case NCL_OP_ADD: o = R_ADD; break; case NCL_OP_MUL: o = R_MUL; break; case NCL_OP_MAX: o = R_MAX; break; case NCL_OP_MIN: o = R_MIN; break; default: return false; }
  const XNode& a = e.node[r.a]; const XNode& b = e.node[r.b];
  bool ok = (a.op == NCL_X_OUTPUT && a.a == 1 && b.op == NCL_X_INPUT && b.a == 1) ||
            (o <= R_MUL && b.op == NCL_X_OUTPUT && b.a == 1 && a.op == NCL_X_INPUT && a.a == 1);
  if (!ok) return false;
  *op = o; return true;
}

struct RDim { long long ext, sin, sout; };
static void sort_merge(RDim* d, int& n, bool kept) {
  for (int i = 1; i < n; i++) { RDim t = d[i]; int j = i - 1; while (j >= 0 && llabs(d[j].sin) < llabs(t.sin)) { d[j + 1] = d[j]; j--; } d[j + 1] = t; }
  for (int i = n - 1; i > 0; i--) {
    bool ok = d[i - 1].sin == d[i].sin * d[i].ext && (!kept || d[i - 1].sout == d[i].sout * d[i].ext);
    if (ok) { d[i - 1].ext *= d[i].ext; d[i - 1].sin = d[i].sin; d[i - 1].sout = d[i].sout; for (int j = i; j + 1 < n; j++) d[j] = d[j + 1]; n--; }
  }
}

static int fill_common(RedParams& p, const Nest& n, int op) {
  const Operand& in = n.in[0]; const Operand& out = n.out[0];
  memset(&p, 0, sizeof p);
  p.in = in.base; p.in_off = in.offset; p.lk = dtype_load_kind(in.dtype);
  p.out = out.base; p.out_off = out.offset; p.sk = dtype_store_spec(out.dtype); p.out_lk = dtype_load_kind(out.dtype);
  p.fresh = n.fresh ? 1 : 0;
  if (n.fresh) {
    if (n.has_identity) { p.ident_i = n.ident_i; p.ident_f = n.ident_f; }
    else { p.ident_i = 0; p.ident_f = 0.0; }     // einsum into a fresh `zeros` output
  }
  (void)op; return NCL_OK;
}

// one resident wave: sm_count x (blocks of 256 threads that fit one SM for this kernel)
static long long persistent_blocks(const void* kfn) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, 256, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
  return (long long)ctx().sm_count * per_sm;
}

static bool aligned16(const Operand& o, long long elem_off) {
  int sb = g_dt[o.dtype].slot_bits;
  return (((uintptr_t)o.base + (elem_off * sb) / 8) % 16) == 0;
}

int try_reduce(const Nest& n) {
  if (n.n_out != 1 || n.n_in != 1) return NCL_ERR_UNSUPPORTED;
  int op;
  if (!match_reduce(n.transform[0], &op)) return NCL_ERR_UNSUPPORTED;
  const Operand& in = n.in[0]; const Operand& out = n.out[0];
  int icls = g_dt[in.dtype].cls, ocls = g_dt[out.dtype].cls;
  if (icls != ocls) return NCL_ERR_UNSUPPORTED;                       // e.g. integer sum into a float array: generic path
  if (g_dt[in.dtype].slot_bits < 8) return NCL_ERR_UNSUPPORTED;       // packed inputs: generic path
  if (n.fresh && !n.has_identity && op != R_ADD) return NCL_ERR_UNSUPPORTED;   // (* @1 $1) into zeros etc.: generic defines it
  RDim kd[NCL_MAX_LOOPS], rd[NCL_MAX_LOOPS]; int nk = 0, nr = 0;
  for (int l = 0; l < n.nloops; l++) {
    if (n.extent[l] == 1) continue;
    RDim d = {n.extent[l], in.stride[l], out.stride[l]};
    if (out.stride[l] != 0) kd[nk++] = d; else rd[nr++] = d;
  }
  if (nr == 0) return NCL_ERR_UNSUPPORTED;
  for (int i = 0; i < nk; i++) if (kd[i].sin == 0) return NCL_ERR_UNSUPPORTED;
  for (int i = 0; i < nr; i++) if (rd[i].sin == 0) return NCL_ERR_UNSUPPORTED;   // broadcast input under a reduction: generic
  sort_merge(kd, nk, true); sort_merge(rd, nr, false);
  if (nr != 1) return NCL_ERR_UNSUPPORTED;
  Context& c = ctx();
  int sb = g_dt[in.dtype].slot_bits;
  int Evec = 128 / sb;
  RedParams p; fill_common(p, n, op);
  long long R = rd[0].ext, sR = rd[0].sin;
  int cls = icls;

  if (sR == 1 && nk == 0) {
    // ---- full reduce --------------------------------------------------------------------------
    p.cols = R;
    p.vec = R >= Evec && aligned16(in, in.offset);          // a ragged tail (< Evec elements) is folded by block 0
    int E = p.vec ? Evec : 1;
    long long work = p.vec ? R / E : R;
    RedKernel kfn = kernel_for(cls, E, op, KALL);
    long long blocks = (work + 256 * 4 - 1) / (256 * 4);
    long long cap = persistent_blocks((const void*)kfn); if (blocks > cap) blocks = cap; if (blocks < 1) blocks = 1;
    NCL_TRY(ensure_scratch((size_t)blocks * 8));
    p.block_partials = c.scratch; p.ticket = c.ticket; p.nblocks = (int)blocks;
    kfn<<<(unsigned)blocks, 256, 0, c.stream>>>(p);
    note_launch("reduce_all");
    return cuda_fail(cudaGetLastError(), "reduce_all_kernel");
  }
  // the contiguous kept dim must be the innermost of the kept group for the column form
  bool inner_reduced = (sR == 1);
  if (inner_reduced) {
    // ---- row reduce: kept dims (any strides) x contiguous reduced axis ------------------------------
    if (nk > 4) return NCL_ERR_UNSUPPORTED;
    p.nk = nk; p.rows = 1;
    for (int i = 0; i < nk; i++) { p.kext[i] = kd[i].ext; p.kin[i] = kd[i].sin; p.kout[i] = kd[i].sout; p.rows *= kd[i].ext; }
    p.cols = R;
    bool vec = (R % Evec == 0) && aligned16(in, in.offset);
    for (int i = 0; i < nk && vec; i++) if (((kd[i].sin * sb) / 8) % 16) vec = false;
    p.vec = vec; int E = vec ? Evec : 1;
    long long chunks = vec ? R / E : R;
    bool wide = chunks >= 512;
    long long blocks = wide ? p.rows : (p.rows + 7) / 8;
    if (blocks > 0x7fffffffLL) return NCL_ERR_UNSUPPORTED;
    kernel_for(cls, E, op, wide ? KROW256 : KROW32)<<<(unsigned)blocks, 256, 0, c.stream>>>(p);
    note_launch(wide ? "reduce_row_block" : "reduce_row_warp");
    return cuda_fail(cudaGetLastError(), "reduce_row_kernel");
  }
  // ---- column reduce: [outer kept...][R][contiguous kept] ---------------------------------------------
  if (nk < 1 || kd[nk - 1].sin != 1 || kd[nk - 1].sout != 1) return NCL_ERR_UNSUPPORTED;
  if (nk - 1 > 4) return NCL_ERR_UNSUPPORTED;
  long long ncols = kd[nk - 1].ext;
  p.ncols = ncols; p.r_ext = R; p.r_stride = sR;
  p.nk = nk - 1; p.rows = 1;
  for (int i = 0; i < nk - 1; i++) { p.kext[i] = kd[i].ext; p.kin[i] = kd[i].sin; p.kout[i] = kd[i].sout; p.rows *= kd[i].ext; }
  bool vec = (ncols % Evec == 0) && aligned16(in, in.offset) && ((sR * sb) / 8) % 16 == 0;
  for (int i = 0; i < nk - 1 && vec; i++) if (((kd[i].sin * sb) / 8) % 16) vec = false;
  int E = vec ? Evec : 1; p.vec = vec;
  long long nq = p.rows * (ncols / E);
  // split R so that the grid is ONE resident wave of equal segments (2048 threads per SM assumed here before: two
  // uneven waves at 60 registers)
  long long want = persistent_blocks((const void*)kernel_for(cls, E, op, KCOL)) * 256;
  int nseg = 1;
  if (nq < want && R >= 128) { nseg = (int)((want + nq - 1) / nq); long long maxseg = R / 32; if (nseg > maxseg) nseg = (int)maxseg; if (nseg > 1024) nseg = 1024; if (nseg < 1) nseg = 1; }
  p.nseg = nseg; p.seg = (R + nseg - 1) / nseg;
  p.nseg = (int)((R + p.seg - 1) / p.seg);
  long long bx = (nq + 255) / 256;
  if (bx > 0x7fffffffLL) return NCL_ERR_UNSUPPORTED;
  if (p.nseg > 1) { NCL_TRY(ensure_scratch((size_t)p.nseg * nq * E * 8)); p.partial = c.scratch; }
  dim3 grid((unsigned)bx, (unsigned)p.nseg);
  kernel_for(cls, E, op, KCOL)<<<grid, 256, 0, c.stream>>>(p);
  note_launch("reduce_col");
  NCL_TRY(cuda_fail(cudaGetLastError(), "reduce_col_kernel"));
  if (p.nseg > 1) {
    kernel_for(cls, E, op, KCOLFIN)<<<(unsigned)bx, 256, 0, c.stream>>>(p);
    note_launch("reduce_col");
    NCL_TRY(cuda_fail(cudaGetLastError(), "reduce_col_finish_kernel"));
  }
  return NCL_OK;
}

// sharded full reduce with the exchange fused into the kernel (peer mailboxes over NVLink)
int reduce_all_p2p(int op, const Operand& a, int64_t count, const Operand& r, unsigned long long* const* peers, int rank, int world,
                   unsigned long long seq) {
  Context& c = ctx();
  int cls = g_dt[a.dtype].cls; int sb = g_dt[a.dtype].slot_bits;
  if (sb < 8 || world > 63) return NCL_ERR_UNSUPPORTED;
  int rop = op == NCL_OP_ADD ? R_ADD : op == NCL_OP_MUL ? R_MUL : op == NCL_OP_MAX ? R_MAX : op == NCL_OP_MIN ? R_MIN : -1;
  if (rop < 0) return NCL_ERR_UNSUPPORTED;
  RedParams p; memset(&p, 0, sizeof p);
  p.in = a.base; p.in_off = a.offset; p.lk = dtype_load_kind(a.dtype); p.cols = count;
  p.out = r.base; p.out_off = r.offset; p.sk = dtype_store_spec(r.dtype); p.out_lk = dtype_load_kind(r.dtype);
  p.fresh = 1; p.ident_i = rop == R_MUL ? 1 : 0; p.ident_f = rop == R_MUL ? 1.0 : 0.0;
  if (rop == R_MAX || rop == R_MIN) {
    const DtInfo& d = g_dt[r.dtype]; bool mx = rop == R_MAX;
    if (d.cls == CLS_F) p.ident_f = mx ? -3.4028234663852886e38 : 3.4028234663852886e38;
    else if (d.cls == CLS_D) p.ident_f = mx ? -1.7976931348623157e308 : 1.7976931348623157e308;
    else if (d.is_signed) p.ident_i = mx ? (d.value_bits >= 64 ? INT64_MIN : -(1ll << (d.value_bits - 1))) : (d.value_bits >= 64 ? INT64_MAX : (1ll << (d.value_bits - 1)) - 1);
    else p.ident_i = mx ? 0 : (d.value_bits >= 63 ? INT64_MAX : (1ll << d.value_bits) - 1);
  }
  int Evec = 128 / sb;
  p.vec = count >= Evec && aligned16(a, a.offset);
  int E = p.vec ? Evec : 1;
  long long work = p.vec ? count / E : count;
  RedKernel kfn = kernel_for(cls, E, rop, KALL);
  long long blocks = (work + 1023) / 1024; long long cap = persistent_blocks((const void*)kfn); if (blocks > cap) blocks = cap; if (blocks < 1) blocks = 1;
  NCL_TRY(ensure_scratch((size_t)blocks * 8));
  p.block_partials = c.scratch; p.ticket = c.ticket; p.nblocks = (int)blocks;
  p.peers = peers; p.rank = rank; p.world = world; p.seq = seq;
  kfn<<<(unsigned)blocks, 256, 0, c.stream>>>(p);
  note_launch("reduce_all_p2p");
  return cuda_fail(cudaGetLastError(), "reduce_all_kernel (p2p)");
}

// local phase of a sharded full reduce: contiguous a -> one float64 / int64 partial on the device
int reduce_all_to_partial(int op, const Operand& a, int64_t count, void* partial8) {
  Context& c = ctx();
  int cls = g_dt[a.dtype].cls; int sb = g_dt[a.dtype].slot_bits;
  if (sb < 8) return set_error(NCL_ERR_UNSUPPORTED, "sharded reduce of packed arrays");
  int rop = op == NCL_OP_ADD ? R_ADD : op == NCL_OP_MUL ? R_MUL : op == NCL_OP_MAX ? R_MAX : op == NCL_OP_MIN ? R_MIN : -1;
  if (rop < 0) return set_error(NCL_ERR_UNSUPPORTED, "reduce op");
  RedParams p; memset(&p, 0, sizeof p);
  p.in = a.base; p.in_off = a.offset; p.lk = dtype_load_kind(a.dtype); p.cols = count;
  int Evec = 128 / sb;
  p.vec = count >= Evec && aligned16(a, a.offset);
  int E = p.vec ? Evec : 1;
  long long work = p.vec ? count / E : count;
  RedKernel kfn = kernel_for(cls, E, rop, KALL);
  long long blocks = (work + 1023) / 1024; long long cap = persistent_blocks((const void*)kfn); if (blocks > cap) blocks = cap; if (blocks < 1) blocks = 1;
  NCL_TRY(ensure_scratch((size_t)blocks * 8));
  p.block_partials = c.scratch; p.ticket = c.ticket; p.nblocks = (int)blocks; p.raw_partial = partial8; p.fresh = 1;
  kfn<<<(unsigned)blocks, 256, 0, c.stream>>>(p);
  note_launch("reduce_all");
  return cuda_fail(cudaGetLastError(), "reduce_all_kernel");
}
