// k_gemm_tf32.cu -- K4, single-float: `(ij jk -> ik)` / `(bij bjk -> bik)` as 3xTF32 on the
// 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM, operands fetched by the
// TMA bulk-copy engine).
//
// Why 3xTF32: one TF32 MMA keeps 10 mantissa bits (~5e-4 relative), far outside the 1e-5 bar.
// Each operand is split  x = hi + lo  with hi = rna_tf32(x), lo = rna_tf32(x - hi)  and
//   A*B ~= A_lo*B_hi + A_hi*B_lo + A_hi*B_hi          (lo*lo ~ 2^-22 relative, dropped)
// accumulated in fp32 in TMEM: ~5e-7 relative, about as good as the reference's own sequential
// single-float accumulation.
//
// Three variants, chosen by gemm_f32_tf32x3 (bottom of this file) from measured crossovers:
//   * pack + gemm_tf32x3_kernel            large matrices (operands are re-used by many tiles)
//   * gemm_tf32x3_fused_kernel             small tile-aligned matrices: the split happens inside the GEMM
//   * gemm_tf32x3_fused2_kernel            the same on CTA pairs (tcgen05 cta_group::2), M % 256 == 0
// The first one is two kernels:
//  1. pack (HBM-bound pre-pass): splits A and B and writes them as PACKED TILES, byte-for-byte
//     the shared-memory image tcgen05 wants -- K-major, 128-byte rows of 32 k-values, 8-row
//     swizzle atoms (16-byte chunk c of row r stored at chunk c ^ (r & 7): UMMA SWIZZLE_128B) --
//     with hi and lo tiles adjacent.  B is transposed on the way so both operands are K-major.
//     A pipeline stage is then two plain contiguous blobs, fetched with two cp.async.bulk
//     (SASS UBLKCP) instructions: no tensor maps, no address math in the hot loop.
//  2. gemm: persistent, one CTA per SM looping over 128 x 256 output tiles, 10 warps, warp-specialised:
//       warp 0      TMA producer  (one lane): empty[s] -> expect_tx -> 2 bulk copies -> full[s]
//       warp 1      MMA issuer    (one lane): full[s] -> 4 k-steps x 3 tcgen05.mma -> commit -> empty[s]
//       warps 2..9  epilogue: tmem_full[b] -> tcgen05.ld 32 lanes x 32 columns -> running fp32 sum -> global
//     K is consumed in chunks of 256 that alternate between two TMEM accumulators (see the kernel).
#include "k_gemm.cuh"

#define TF_BM 128
#define TF_BN 256
#define TF_BK 32
#define TF_STAGES 2
#define TF_A_TILE (TF_BM * TF_BK * 4)            // 16 KB
#define TF_B_TILE (TF_BN * TF_BK * 4)            // 32 KB
#define TF_STAGE_BYTES (2 * TF_A_TILE + 2 * TF_B_TILE)     // hi+lo of both: 96 KB
#define TF_SMEM (TF_STAGES * TF_STAGE_BYTES + 2048)

// ---- pack ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  unsigned h;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  hi = __uint_as_float(h);
  float r = __fsub_rn(x, hi);
  if (!isfinite(hi)) r = 0.0f;                       // inf/nan stay in hi; inf - inf must not poison lo
  unsigned l;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(r));
  lo = __uint_as_float(l);
}
__device__ __forceinline__ int swz(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

// A (M x K, k contiguous) -> tiles [b][mt][kt]{hi 16KB, lo 16KB}
__global__ void __launch_bounds__(256) pack_a_kernel(const float* __restrict__ A, char* __restrict__ out, int M, int K, long long lda, long long sA,
                                                     int MT, int KT) {
  // one block per (kt, mt, b): 128 rows x 8 chunks = 1024 chunks, 4 per thread
  const int kt = blockIdx.x, mt = blockIdx.y, b = blockIdx.z;
  char* tile = out + (((long long)b * MT + mt) * KT + kt) * (2LL * TF_A_TILE);
  const float* Ab = A + (long long)b * sA;
  const bool vec = ((lda & 3) == 0) && ((sA & 3) == 0) && (((uintptr_t)A & 15) == 0);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int idx = threadIdx.x + i * 256;
    int row = idx >> 3, c = idx & 7;
    int m = mt * TF_BM + row, k = kt * TF_BK + c * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (m < M) {
      const float* src = Ab + (long long)m * lda + k;
      if (vec && k + 3 < K) { float4 t = __ldg((const float4*)src); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
      else { for (int j = 0; j < 4; j++) if (k + j < K) v[j] = __ldg(src + j); }
    }
    float h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; j++) split_tf32(v[j], h[j], l[j]);
    int o = swz(row, c);
    *(float4*)(tile + o) = make_float4(h[0], h[1], h[2], h[3]);
    *(float4*)(tile + TF_A_TILE + o) = make_float4(l[0], l[1], l[2], l[3]);
  }
}

// B (K x N, n contiguous) -> transposed tiles [b][nt][kt]{hi 32KB, lo 32KB}, rows = n, 32 k-values per row
__global__ void __launch_bounds__(256) pack_b_kernel(const float* __restrict__ B, char* __restrict__ out, int K, int N, long long ldb, long long sB,
                                                     int NT, int KT) {
  // one block per (kt, 64-column slab of n, b)
  __shared__ float t[TF_BK][65];
  const int kt = blockIdx.x, ns = blockIdx.y, b = blockIdx.z;
  const int n_base = ns * 64;
  const int nt = n_base / TF_BN, row_base = n_base % TF_BN;
  char* tile = out + (((long long)b * NT + nt) * KT + kt) * (2LL * TF_B_TILE);
  const float* Bb = B + (long long)b * sB;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;      // 64 columns x 4 rows per pass
#pragma unroll
  for (int r = 0; r < TF_BK; r += 4) {
    int k = kt * TF_BK + r + ty, n = n_base + tx;
    t[r + ty][tx] = (k < K && n < N) ? __ldg(Bb + (long long)k * ldb + n) : 0.0f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 2; i++) {
    int idx = threadIdx.x + i * 256;                            // 64 rows x 8 chunks
    int rn = idx >> 3, c = idx & 7;
    float h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; j++) split_tf32(t[c * 4 + j][rn], h[j], l[j]);
    int o = swz(row_base + rn, c);
    *(float4*)(tile + o) = make_float4(h[0], h[1], h[2], h[3]);
    *(float4*)(tile + TF_B_TILE + o) = make_float4(l[0], l[1], l[2], l[3]);
  }
}

// ---- PTX wrappers -------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity) {
  unsigned ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug must trap, not hang the GPU
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) { printf("numcl-cuda: mbarrier timeout (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x); __trap(); }
  }
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(unsigned bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(unsigned d_tmem, unsigned long long adesc, unsigned long long bdesc, unsigned idesc, unsigned accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
// K-major SWIZZLE_128B operand tile at `addr` (1024B aligned + k-step offset): SBO = 1024 B between 8-row atoms
__device__ __forceinline__ unsigned long long umma_desc(unsigned addr) {
  return (unsigned long long)((addr >> 4) & 0x3fffu) | ((unsigned long long)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ void mbar_arrive(unsigned bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }

#define TF_CHUNK_KB 8            // k-blocks (of 32) accumulated inside TMEM before the sum leaves the tensor core
#define TF_THREADS 320           // warp 0 producer, warp 1 MMA, warps 2..9 epilogue

// The tensor core ADDS into the TMEM accumulator with truncation (measured on B200: relative error
// 7e-9 * K, always toward zero -- tools/tf32_error.py), which at K = 8192 would be 6e-5, outside the
// 1e-5 bar.  So K is cut into chunks of TF_CHUNK_KB*32 = 256: each chunk is accumulated in TMEM
// from zero (<= 1.8e-6), then the eight epilogue warps pull it out with tcgen05.ld and add it to a
// running round-to-nearest fp32 sum held in registers (128 per thread).  Two TMEM accumulators
// (2 x 256 columns) alternate so that draining chunk c overlaps the MMAs of chunk c+1.
__global__ void __launch_bounds__(TF_THREADS, 1) gemm_tf32x3_kernel(const char* __restrict__ Ap, const char* __restrict__ Bp, float* __restrict__ C,
                                                             int M, int N, int KT, long long ldc, long long sC, int MT, int NT, int batch, int accumulate) {
  // PERSISTENT: one CTA per SM walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...  The smem ring
  // and the two TMEM accumulators keep rolling across tile boundaries, so the MMAs of the next tile
  // run while the epilogue warps are still storing the previous one.
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long bars[2 * TF_STAGES + 4];
  __shared__ unsigned tmem_slot;
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[TF_STAGES]);
  const unsigned tfull0 = smem_u32(&bars[2 * TF_STAGES]), tempty0 = smem_u32(&bars[2 * TF_STAGES + 2]);
  const int NC = (KT + TF_CHUNK_KB - 1) / TF_CHUNK_KB;
  const int tiles_per_batch = MT * NT;
  const int total_tiles = tiles_per_batch * batch;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TF_STAGES; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    for (int t = 0; t < 2; t++) { mbar_init(tfull0 + 8 * t, 1); mbar_init(tempty0 + 8 * t, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {                                    // TMEM: 2 x 256 columns = two 128 x 256 fp32 accumulators
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_base = tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      unsigned it = 0;                                // k-blocks issued so far (all tiles)
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int b = t / tiles_per_batch, tm, tn; raster(t - b * tiles_per_batch, MT, NT, 8, tm, tn);
        const char* a_src = Ap + (((long long)b * MT + tm) * KT) * (2LL * TF_A_TILE);
        const char* b_src = Bp + (((long long)b * NT + tn) * KT) * (2LL * TF_B_TILE);
        for (int kt = 0; kt < KT; kt++, it++) {
          int s = it % TF_STAGES; unsigned ph = (it / TF_STAGES) & 1;
          mbar_wait(empty0 + 8 * s, ph ^ 1);
          mbar_expect_tx(full0 + 8 * s, TF_STAGE_BYTES);
          unsigned dst = smem_u32(smem + s * TF_STAGE_BYTES);
          bulk_g2s(dst, a_src + (long long)kt * (2 * TF_A_TILE), 2 * TF_A_TILE, full0 + 8 * s);
          bulk_g2s(dst + 2 * TF_A_TILE, b_src + (long long)kt * (2 * TF_B_TILE), 2 * TF_B_TILE, full0 + 8 * s);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D=f32 (bits 4-5 = 1), A=B=tf32 (bits 7-9, 10-12 = 2), K-major both, N>>3 at 17, M>>4 at 24
      const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(TF_BN >> 3) << 17) | ((unsigned)(TF_BM >> 4) << 24);
      unsigned it = 0, g = 0;                         // k-blocks / chunks consumed so far (all tiles)
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int kt = 0;
        for (int c = 0; c < NC; c++, g++) {
          const int buf = g & 1; const unsigned use = g >> 1;
          mbar_wait(tempty0 + 8 * buf, (use & 1) ^ 1);          // the epilogue has drained this accumulator
          tc_fence_after();
          const unsigned d_tmem = tmem_base + (unsigned)(buf * TF_BN);
          const int kend = (kt + TF_CHUNK_KB < KT) ? kt + TF_CHUNK_KB : KT;
          for (int first = 1; kt < kend; kt++, it++) {
            int s = it % TF_STAGES; unsigned ph = (it / TF_STAGES) & 1;
            mbar_wait(full0 + 8 * s, ph);
            tc_fence_after();
            unsigned a_hi = smem_u32(smem + s * TF_STAGE_BYTES), a_lo = a_hi + TF_A_TILE;
            unsigned b_hi = a_hi + 2 * TF_A_TILE, b_lo = b_hi + TF_B_TILE;
#pragma unroll
            for (int ks = 0; ks < TF_BK / 8; ks++) {     // one MMA covers K = 8 tf32 = 32 bytes of every row
              unsigned off = ks * 32;
              tc_mma_tf32(d_tmem, umma_desc(a_lo + off), umma_desc(b_hi + off), idesc, first ? 0u : 1u);   // small terms first
              first = 0;
              tc_mma_tf32(d_tmem, umma_desc(a_hi + off), umma_desc(b_lo + off), idesc, 1u);
              tc_mma_tf32(d_tmem, umma_desc(a_hi + off), umma_desc(b_hi + off), idesc, 1u);
            }
            tc_commit(empty0 + 8 * s);                    // stage reusable once these MMAs have read it
          }
          tc_commit(tfull0 + 8 * buf);                    // chunk complete in TMEM
        }
      }
    }
  } else {
    // epilogue warps 2..9: TMEM lane quarter q = warp % 4, column half h; thread owns 128 outputs of one row
    const int q = warp & 3, h = (warp - 2) >> 2;
    const bool vec = ((ldc & 3) == 0) && ((sC & 3) == 0) && (((uintptr_t)C & 15) == 0);
    unsigned g = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int b = t / tiles_per_batch, tm, tn; raster(t - b * tiles_per_batch, MT, NT, 8, tm, tn);
      float acc[128];
#pragma unroll
      for (int j = 0; j < 128; j++) acc[j] = 0.0f;
      for (int c = 0; c < NC; c++, g++) {
        const int buf = g & 1; const unsigned use = g >> 1;
        mbar_wait(tfull0 + 8 * buf, use & 1);
        tc_fence_after();
#pragma unroll
        for (int cc = 0; cc < 4; cc++) {
          unsigned r[32];
          unsigned taddr = tmem_base + ((unsigned)(q * 32) << 16) + (unsigned)(buf * TF_BN + h * 128 + cc * 32);
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                       "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                       "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                       : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                         "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                         "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                         "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                       : "r"(taddr) : "memory");
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 32; j++) acc[cc * 32 + j] = __fadd_rn(acc[cc * 32 + j], __uint_as_float(r[j]));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
      }
      const int m = tm * TF_BM + q * 32 + lane;
      const int nbase = tn * TF_BN + h * 128;
      if (m < M) {
        float* crow = C + (long long)b * sC + (long long)m * ldc + nbase;
        if (vec && nbase + 127 < N) {
#pragma unroll
          for (int j = 0; j < 128; j += 4) {
            float4 o = accumulate ? *(const float4*)(crow + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 v = make_float4(__fadd_rn(o.x, acc[j]), __fadd_rn(o.y, acc[j + 1]), __fadd_rn(o.z, acc[j + 2]), __fadd_rn(o.w, acc[j + 3]));
            *(float4*)(crow + j) = v;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 128; j++) if (nbase + j < N) {
            float o = accumulate ? crow[j] : 0.0f;
            crow[j] = __fadd_rn(o, acc[j]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}


// ===== fused variant: the hi/lo split happens INSIDE the GEMM kernel ================================================
//
// For small matrices (C4: 256 x 512^3) the pack pre-pass and the doubled operand stream cost more HBM
// time than the MMAs take (DESIGN.md "C4").  Here the raw fp32 operands are read once by CONVERTER
// warps straight from global memory (L2 serves the re-reads of a batch), split in registers and
// stored to shared memory as the four swizzled tf32 tiles the tensor core wants:
//   warps 0..3    convert A   128 x 32 block  -> A_hi, A_lo   K-major  SWIZZLE_128B (as in the pack kernel)
//   warps 4..11   convert B    32 x 256 block -> B_hi, B_lo   MN-major: B is n-contiguous in memory and
//                 tcgen05 takes it as is (instruction-descriptor bit 16), so no transpose.  The only
//                 MN-major layout 32-bit operands have is SWIZZLE_128B_BASE32B (descriptor layout type 1):
//                 atom = 4 k-rows x 128 B (32 n), 32-byte units of a row XORed with (k & 3), i.e. byte
//                 address bits [5,7) ^= bits [7,9).  Stage layout [n-atom (8)][32 k-rows x 128 B]:
//                 LBO (next n-atom) = 4096 B, SBO (next 4-row group) = 512 B; a K = 8 MMA spans two groups.
//   warp 12       MMA issuer (one lane), as above
//   warps 13..16  epilogue (one per TMEM lane quarter).  K is still cut into chunks of 256 (TMEM truncation, see above) but the
//                 running sum lives in C itself: chunk 0 stores, later chunks add to what this thread
//                 stored (same thread, same addresses, L1/L2 hits) -- no 128-register accumulator, so
//                 17 warps fit in the register file.  Used for K <= TF_FUSED_MAX_K only.
// (Tried and dropped: an 18th warp prefetching the raw rows into L2 with cp.async.bulk.prefetch -- 2.5x slower.)
// Every converter thread keeps the NEXT k-block's 8 x 16 B in registers while it converts the current
// one, so global latency overlaps conversion + the wait for a free stage.
// Shapes: M % 128 == 0, N % 256 == 0, K % 32 == 0, 16-byte aligned rows; anything else takes the packed path.
#define TFF_CONV_WARPS 12
#define TFF_MMA_WARP 12
#define TFF_EPI_WARP0 13
#define TFF_EPI_WARPS 4
#define TFF_THREADS ((TFF_EPI_WARP0 + TFF_EPI_WARPS) * 32)     // 544: 17 warps, at most 5 per SM sub-partition -> 96 registers each
#define TF_FUSED_MAX_K 1024
#define TFF_STG_PITCH 36                          // floats per staged row (32 + 4: 16-byte aligned, bank-conflict free)
#define TFF_STG_BYTES (32 * TFF_STG_PITCH * 4)
#define TFF_SMEM (TF_SMEM + TFF_EPI_WARPS * TFF_STG_BYTES)

// MN-major SWIZZLE_128B_BASE32B operand (layout type 1): LBO = 4096 B to the next 32-column atom, SBO = 512 B
// to the next group of 4 k-rows
__device__ __forceinline__ unsigned long long umma_desc_mn(unsigned addr) {
  return (unsigned long long)((addr >> 4) & 0x3fffu) | ((unsigned long long)(4096u >> 4) << 16) |
         ((unsigned long long)(512u >> 4) << 32) | (1ull << 46) | (1ull << 61);
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// hi = rna_tf32(x), lo = rna_tf32(x - hi) with integer rounding: ptxas expands cvt.rna.tf32 into a
// guarded add/mask anyway; sharing one finiteness test between hi, lo and the inf/NaN rule of
// split_tf32 brings the split from ~12 to 8 instructions per element (the converters are ALU-bound).
__device__ __forceinline__ void split_fast(float x, unsigned& hi, unsigned& lo) {
  const unsigned b = __float_as_uint(x);
  const bool fin = fabsf(x) < __int_as_float(0x7f800000);
  unsigned h = (b + 0x1000u) & 0xffffe000u;             // round to nearest, ties away, 10 mantissa bits kept
  h = fin ? h : b;                                       // inf / NaN pass through untouched
  const float r = __fsub_rn(x, __uint_as_float(h));      // exact
  const unsigned l = (__float_as_uint(r) + 0x1000u) & 0xffffe000u;
  hi = h; lo = fin ? l : 0u;
}
__device__ __forceinline__ void split4(const float4& v, uint4& h, uint4& l) {
  split_fast(v.x, h.x, l.x); split_fast(v.y, h.y, l.y); split_fast(v.z, h.z, l.z); split_fast(v.w, h.w, l.w);
}
__device__ __forceinline__ void sts128(unsigned addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(unsigned addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

__global__ void __launch_bounds__(TFF_THREADS, 1) gemm_tf32x3_fused_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C,
                                                                            int KT, long long lda, long long ldb, long long ldc,
                                                                            long long sA, long long sB, long long sC,
                                                                            int MT, int NT, int batch, int accumulate) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long bars[2 * TF_STAGES + 4];
  __shared__ unsigned tmem_slot;
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[TF_STAGES]);
  const unsigned tfull0 = smem_u32(&bars[2 * TF_STAGES]), tempty0 = smem_u32(&bars[2 * TF_STAGES + 2]);
  const int NC = (KT + TF_CHUNK_KB - 1) / TF_CHUNK_KB;
  const int tiles_per_batch = MT * NT;
  const int total_tiles = tiles_per_batch * batch;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TF_STAGES; s++) { mbar_init(full0 + 8 * s, TFF_CONV_WARPS); mbar_init(empty0 + 8 * s, 1); }
    for (int t = 0; t < 2; t++) { mbar_init(tfull0 + 8 * t, 1); mbar_init(tempty0 + 8 * t, TFF_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == TFF_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const unsigned tmem_base = tmem_slot;

  if (warp < TFF_CONV_WARPS) {
    // ---- converters -------------------------------------------------------------------------------------
    const bool isA = warp < 4;
    const int ct = isA ? (int)threadIdx.x : (int)threadIdx.x - 128;
    // A thread: chunk i -> row (ct>>3) + 16 i, 16-byte column ct&7.   B thread: chunk i -> k (ct>>6) + 4 i, 16-byte column ct&63.
    const int r0 = isA ? (ct >> 3) : (ct >> 6), c0 = isA ? (ct & 7) : (ct & 63);
    const long long step = isA ? 16 * lda : 4 * ldb;                 // elements between this thread's chunks
    const long long kstep = isA ? (long long)TF_BK : (long long)TF_BK * ldb;   // elements between k-blocks
    // smem offset of this thread's hi chunk i inside a stage: (i odd ? off_odd : off_even) + (i >> 1) * pair_stride
    unsigned off_even, off_odd, pair_stride;
    if (isA) { off_even = (unsigned)swz(r0, c0); off_odd = off_even + 16 * 128; pair_stride = 32 * 128; }   // (row + 16 i) & 7 == row & 7
    else {
      // k = r0 + 4 i: k & 3 = r0 for every i, so the swizzled 16-byte column is fixed and the offset is linear in i
      off_even = (unsigned)(2 * TF_A_TILE + (c0 >> 3) * 4096 + r0 * 128 + (((c0 & 7) ^ (r0 << 1)) << 4));
      off_odd = off_even + 4 * 128;
      pair_stride = 8 * 128;
    }
    const unsigned lo_off = isA ? TF_A_TILE : TF_B_TILE;
    auto tile_ptr = [&](int t) -> const float* {
      int b = t / tiles_per_batch, tm, tn; raster(t - b * tiles_per_batch, MT, NT, 8, tm, tn);
      return isA ? A + (long long)b * sA + ((long long)tm * TF_BM + r0) * lda + c0 * 4
                 : B + (long long)b * sB + (long long)r0 * ldb + (long long)tn * TF_BN + c0 * 4;
    };
    int t = blockIdx.x, kt = 0;
    bool have = t < total_tiles;
    const float* src = have ? tile_ptr(t) : nullptr;
    float4 cur[8], nxt[8];
    if (have) {
#pragma unroll
      for (int i = 0; i < 8; i++) cur[i] = __ldg((const float4*)(src + i * step));
    }
    unsigned it = 0;
    while (have) {
      int nkt = kt + 1, ntile = t; bool nhave = true;
      if (nkt == KT) { nkt = 0; ntile = t + gridDim.x; nhave = ntile < total_tiles; if (nhave) src = tile_ptr(ntile); }
      if (nhave) {
        const float* s2 = src + nkt * kstep;
#pragma unroll
        for (int i = 0; i < 8; i++) nxt[i] = __ldg((const float4*)(s2 + i * step));
      }
      const int s = it % TF_STAGES; const unsigned ph = (it / TF_STAGES) & 1;
      mbar_wait(empty0 + 8 * s, ph ^ 1);
      const unsigned st = smem_u32(smem) + s * TF_STAGE_BYTES;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        uint4 h, l; split4(cur[i], h, l);
        const unsigned d = st + ((i & 1) ? off_odd : off_even) + (i >> 1) * pair_stride;
        sts128(d, h);
        sts128(d + lo_off, l);
      }
      fence_proxy_async_smem();                       // generic-proxy stores -> visible to the tensor core's async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(full0 + 8 * s);
#pragma unroll
      for (int i = 0; i < 8; i++) cur[i] = nxt[i];
      t = ntile; kt = nkt; have = nhave; it++;
    }
  } else if (warp == TFF_MMA_WARP) {
    if (lane == 0) {
      // as the packed kernel, plus bit 16: B is MN-major
      const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((unsigned)(TF_BN >> 3) << 17) | ((unsigned)(TF_BM >> 4) << 24);
      unsigned it = 0, g = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int kt = 0;
        for (int c = 0; c < NC; c++, g++) {
          const int buf = g & 1; const unsigned use = g >> 1;
          mbar_wait(tempty0 + 8 * buf, (use & 1) ^ 1);
          tc_fence_after();
          const unsigned d_tmem = tmem_base + (unsigned)(buf * TF_BN);
          const int kend = (kt + TF_CHUNK_KB < KT) ? kt + TF_CHUNK_KB : KT;
          for (int first = 1; kt < kend; kt++, it++) {
            int s = it % TF_STAGES; unsigned ph = (it / TF_STAGES) & 1;
            mbar_wait(full0 + 8 * s, ph);
            tc_fence_after();
            unsigned a_hi = smem_u32(smem + s * TF_STAGE_BYTES), a_lo = a_hi + TF_A_TILE;
            unsigned b_hi = a_hi + 2 * TF_A_TILE, b_lo = b_hi + TF_B_TILE;
#pragma unroll
            for (int ks = 0; ks < TF_BK / 8; ks++) {
              unsigned aoff = ks * 32, boff = ks * 1024;           // A: 32 B along the row; B: the next 8 k-rows
              tc_mma_tf32(d_tmem, umma_desc(a_lo + aoff), umma_desc_mn(b_hi + boff), idesc, first ? 0u : 1u);
              first = 0;
              tc_mma_tf32(d_tmem, umma_desc(a_hi + aoff), umma_desc_mn(b_lo + boff), idesc, 1u);
              tc_mma_tf32(d_tmem, umma_desc(a_hi + aoff), umma_desc_mn(b_hi + boff), idesc, 1u);
            }
            tc_commit(empty0 + 8 * s);
          }
          tc_commit(tfull0 + 8 * buf);
        }
      }
    }
  } else {
    // ---- epilogue: TMEM lane quarter q = warp % 4 (hardware rule), all 256 columns ----------------------------
    // tcgen05.ld hands every lane one ROW (32 consecutive columns); stored like that a warp instruction
    // would touch 32 different 128-byte lines.  Each 32 x 32 slab is therefore turned through a private
    // 4.5 KB staging buffer (pitch 36 floats: conflict-free both ways) so that one instruction moves
    // 4 complete 128-byte row segments of C.
    const int q = warp & 3;
    const unsigned stg = smem_u32(smem) + TF_STAGES * TF_STAGE_BYTES + (warp - TFF_EPI_WARP0) * TFF_STG_BYTES;
    const unsigned stg_w = stg + lane * (TFF_STG_PITCH * 4);                       // this lane's row when writing
    const int rr = lane >> 3, c4 = lane & 7;                                       // when reading: rows rr + 4 p, 16-byte column c4
    const unsigned stg_r = stg + rr * (TFF_STG_PITCH * 4) + c4 * 16;
    unsigned g = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int b = t / tiles_per_batch, tm, tn; raster(t - b * tiles_per_batch, MT, NT, 8, tm, tn);
      float* cbase = C + (long long)b * sC + ((long long)tm * TF_BM + q * 32 + rr) * ldc + (long long)tn * TF_BN + c4 * 4;
      for (int c = 0; c < NC; c++, g++) {
        const int buf = g & 1; const unsigned use = g >> 1;
        const bool add = (c > 0) || accumulate;
        mbar_wait(tfull0 + 8 * buf, use & 1);
        tc_fence_after();
#pragma unroll 1
        for (int cc = 0; cc < TF_BN / 32; cc++) {
          unsigned r[32];
          unsigned taddr = tmem_base + ((unsigned)(q * 32) << 16) + (unsigned)(buf * TF_BN + cc * 32);
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                       "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                       "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                       : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                         "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                         "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                         "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                       : "r"(taddr) : "memory");
          float* dst = cbase + cc * 32;
          float4 old[8];
          if (add) {                        // the running sum so far, fetched while the TMEM load is in flight
#pragma unroll
            for (int p = 0; p < 8; p++) old[p] = *(const float4*)(dst + (long long)(4 * p) * ldc);
          }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (cc == TF_BN / 32 - 1) {                     // everything is out of TMEM: hand the accumulator back early
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
          }
#pragma unroll
          for (int j = 0; j < 32; j += 4) sts128(stg_w + j * 4, make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]));
          __syncwarp();
#pragma unroll
          for (int p = 0; p < 8; p++) {
            uint4 v = lds128(stg_r + p * (4 * TFF_STG_PITCH * 4));
            float4 o = add ? old[p] : make_float4(0.f, 0.f, 0.f, 0.f);
            *(float4*)(dst + (long long)(4 * p) * ldc) = make_float4(__fadd_rn(o.x, __uint_as_float(v.x)), __fadd_rn(o.y, __uint_as_float(v.y)),
                                                                       __fadd_rn(o.z, __uint_as_float(v.z)), __fadd_rn(o.w, __uint_as_float(v.w)));
          }
          __syncwarp();                                   // staging buffer is rewritten by the next slab
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == TFF_MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ===== fused variant on CTA PAIRS (tcgen05 cta_group::2) ============================================================
//
// The single-CTA fused kernel is bound by the SM's shared-memory data path (per k-block: 48 KB LDG fill +
// 96 KB converter stores + 144 KB tensor-core operand reads).  Two CTAs of a cluster (one TPC) share one
// 256 x 256 output tile: UMMA M = 256, each CTA owns 128 rows of A and D and only HALF of B (128 of the
// 256 columns); the tensor cores fetch the other half from the peer's shared memory.  Per CTA and k-block
// that is 32 KB LDG + 64 KB stores + 96 KB operand reads, and a 64 KB stage, so three stages fit.
//   warps 0..3  convert A (128 x 32), warps 4..7 convert this CTA's B half (32 x 128)
//   warp 8      leader CTA only: issues tcgen05.mma.cta_group::2; commits are multicast to both CTAs
//   warps 9..12 epilogue of this CTA's 128 rows (TMEM lanes) x 256 columns
// full[] and tmem_empty[] live in the LEADER's shared memory; the peer's warps arrive on them remotely
// (mapa + mbarrier.arrive.release.cluster).  Needs M % 256 == 0 on top of the fused kernel's conditions.
#define F2_CONV_WARPS 8
#define F2_MMA_WARP 8
#define F2_EPI_WARP0 9
#define F2_EPI_WARPS 4
#define F2_THREADS ((F2_EPI_WARP0 + F2_EPI_WARPS) * 32)          // 416: 13 warps, <= 128 registers
#define F2_CPT 8                                                  // 16-byte chunks per converter thread and k-block
#define F2_STAGES 3
#define F2_B_TILE (TF_BK * 128 * 4)                               // this CTA's half of B: 16 KB
#define F2_STAGE_BYTES (2 * TF_A_TILE + 2 * F2_B_TILE)            // 64 KB
#define F2_SMEM (F2_STAGES * F2_STAGE_BYTES + 2048 + F2_EPI_WARPS * TFF_STG_BYTES)

__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned mapa_u32(unsigned addr, unsigned rank) {
  unsigned r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void mbar_arrive_remote(unsigned cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(unsigned bar, unsigned parity) {
  unsigned ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(unsigned bar, unsigned parity) {
  long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) { printf("numcl-cuda: cluster mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
  }
}
__device__ __forceinline__ void tc_commit2(unsigned bar) {          // arrives on the barrier at this offset in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((unsigned short)3) : "memory");
}
__device__ __forceinline__ void tc_mma2_tf32(unsigned d_tmem, unsigned long long adesc, unsigned long long bdesc, unsigned idesc, unsigned accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__global__ void __launch_bounds__(F2_THREADS, 1) gemm_tf32x3_fused2_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C,
                                                                            int KT, long long lda, long long ldb, long long ldc,
                                                                            long long sA, long long sB, long long sC,
                                                                            int MT2, int NT, int batch, int accumulate) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long bars[3 * F2_STAGES + 4];
  __shared__ unsigned tmem_slot;
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const unsigned full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[F2_STAGES]);
  const unsigned tfull0 = smem_u32(&bars[2 * F2_STAGES]), tempty0 = smem_u32(&bars[2 * F2_STAGES + 2]);
  const unsigned pfull0 = smem_u32(&bars[2 * F2_STAGES + 4]);          // leader only: "the peer's half of stage s is converted"
  const unsigned pfull0_leader = mapa_u32(pfull0, 0), tempty0_leader = mapa_u32(tempty0, 0);
  const int NC = (KT + TF_CHUNK_KB - 1) / TF_CHUNK_KB;
  const int tiles_per_batch = MT2 * NT;
  const int total_tiles = tiles_per_batch * batch;

  if (threadIdx.x == 0) {
    for (int s = 0; s < F2_STAGES; s++) { mbar_init(full0 + 8 * s, F2_CONV_WARPS); mbar_init(empty0 + 8 * s, 1); mbar_init(pfull0 + 8 * s, 1); }
    for (int t = 0; t < 2; t++) { mbar_init(tfull0 + 8 * t, 1); mbar_init(tempty0 + 8 * t, 2 * F2_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == F2_MMA_WARP) {                          // one warp of EACH CTA: the pair's TMEM is allocated collectively
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();                                  // barriers initialised and TMEM allocated in both CTAs
  tc_fence_after();
  const unsigned tmem_base = tmem_slot;

  if (warp < F2_CONV_WARPS) {
    // ---- converters -------------------------------------------------------------------------------------
    const bool isA = warp < 4;
    const int ct = isA ? (int)threadIdx.x : (int)threadIdx.x - 128;
    // A thread: chunk i -> row (ct>>3) + 16 i, 16-byte column ct&7.   B thread: chunk i -> k (ct>>5) + 4 i, 16-byte column ct&31.
    // Both swizzles depend on the row only through bits that i does not change, so the smem offset is linear in i.
    const int r0 = isA ? (ct >> 3) : (ct >> 5), c0 = isA ? (ct & 7) : (ct & 31);
    const long long step = isA ? 16 * lda : 4 * ldb;
    const long long kstep = isA ? (long long)TF_BK : (long long)TF_BK * ldb;
    const unsigned off0 = isA ? (unsigned)swz(r0, c0)
                              : (unsigned)(2 * TF_A_TILE + (c0 >> 3) * 4096 + r0 * 128 + (((c0 & 7) ^ (r0 << 1)) << 4));
    const unsigned istride = isA ? 16 * 128 : 4 * 128;
    const unsigned lo_off = isA ? TF_A_TILE : F2_B_TILE;
    auto tile_ptr = [&](int t) -> const float* {
      int b = t / tiles_per_batch, tm, tn; raster(t - b * tiles_per_batch, MT2, NT, 8, tm, tn);
      return isA ? A + (long long)b * sA + ((long long)(2 * tm + (int)rank) * TF_BM + r0) * lda + c0 * 4
                 : B + (long long)b * sB + (long long)r0 * ldb + (long long)tn * TF_BN + (int)rank * 128 + c0 * 4;
    };
    // One k-block of look-ahead in registers, and an L2 prefetch two k-blocks beyond that: about half of the
    // loads are first touches that come from HBM (> 1.5 us under load, longer than one iteration), and there
    // are no registers for a second look-ahead set.  One lane per 128-byte line issues the prefetch.
    int lt = pair, lkt = 0;                                   // load cursor
    const float* lsrc = lt < total_tiles ? tile_ptr(lt) : nullptr;
    const float* nsrc = lt + npairs < total_tiles ? tile_ptr(lt + npairs) : nullptr;     // the tile after the cursor's
    const bool pf_lane = (c0 & 7) == 0;
    auto load_next = [&](float4 (&dst)[F2_CPT]) {
      if (lt >= total_tiles) return;
      const float* s2 = lsrc + lkt * kstep;
#pragma unroll
      for (int i = 0; i < F2_CPT; i++) dst[i] = __ldg((const float4*)(s2 + i * step));
      if (pf_lane) {
        const int pk = lkt + 2;
        const float* ps = pk < KT ? lsrc + pk * kstep : ((nsrc && pk - KT < KT) ? nsrc + (pk - KT) * kstep : nullptr);
        if (ps) {
#pragma unroll
          for (int i = 0; i < F2_CPT; i++) asm volatile("prefetch.global.L2 [%0];" ::"l"(ps + i * step));
        }
      }
      if (++lkt == KT) { lkt = 0; lt += npairs; lsrc = nsrc; nsrc = lt + npairs < total_tiles ? tile_ptr(lt + npairs) : nullptr; }
    };
    float4 cur[F2_CPT], n1[F2_CPT];
    load_next(cur);
    long long nblocks = 0;
    for (int t = pair; t < total_tiles; t += npairs) nblocks += KT;
    for (unsigned it = 0; it < (unsigned)nblocks; it++) {
      load_next(n1);
      const int s = it % F2_STAGES; const unsigned ph = (it / F2_STAGES) & 1;
      mbar_wait(empty0 + 8 * s, ph ^ 1);              // own CTA's copy: the multicast commit arrives on both
      const unsigned st = smem_u32(smem) + s * F2_STAGE_BYTES + off0;
#pragma unroll
      for (int i = 0; i < F2_CPT; i++) {
        uint4 h, l; split4(cur[i], h, l);
        sts128(st + i * istride, h);
        sts128(st + i * istride + lo_off, l);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(full0 + 8 * s);      // local and cheap; the peer's forwarder relays it (see warp F2_MMA_WARP)
#pragma unroll
      for (int i = 0; i < F2_CPT; i++) cur[i] = n1[i];
    }
  } else if (warp == F2_MMA_WARP) {
    if (rank == 1 && lane == 0) {
      // Forwarder.  A cluster-scope release costs a full memory barrier (MEMBAR.ALL.GPU + ERRBAR: ~0.7 us,
      // 36 % of the converters' time when every converter warp arrived remotely).  So the converters
      // arrive on their own CTA's full[] and this otherwise idle thread relays ONE arrival per k-block
      // to the leader, off the converters' critical path.
      long long nblocks = 0;
      for (int t = pair; t < total_tiles; t += npairs) nblocks += KT;
      for (unsigned it = 0; it < (unsigned)nblocks; it++) {
        const int s = it % F2_STAGES; const unsigned ph = (it / F2_STAGES) & 1;
        mbar_wait(full0 + 8 * s, ph);
        mbar_arrive_remote(pfull0_leader + 8 * s);
      }
    }
    if (rank == 0) {
      // The whole warp walks the loop (uniform control flow, so descriptor and barrier arithmetic can stay in
      // uniform registers) and one lane issues.  Descriptors are the stage-0 descriptor plus a constant: with
      // the loop under `if (lane == 0)` ptxas rebuilt every 64-bit descriptor in vector registers and moved
      // it over with R2UR, ~20 instructions per MMA, and ncu showed this warp busy issuing, not waiting.
      // D f32, A/B tf32, B MN-major (bit 16), N = 256, M = 256 over the pair
      const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((unsigned)(256 >> 3) << 17) | ((unsigned)(256 >> 4) << 24);
      const unsigned long long adesc0 = umma_desc(smem_u32(smem));                      // stage 0: A_hi
      const unsigned long long bdesc0 = umma_desc_mn(smem_u32(smem) + 2 * TF_A_TILE);   // stage 0: B_hi
      unsigned it = 0, g = 0;
      for (int t = pair; t < total_tiles; t += npairs) {
        int kt = 0;
        for (int c = 0; c < NC; c++, g++) {
          const int buf = g & 1; const unsigned use = g >> 1;
          mbar_wait_cluster(tempty0 + 8 * buf, (use & 1) ^ 1);
          tc_fence_after();
          const unsigned d_tmem = tmem_base + (unsigned)(buf * TF_BN);
          const int kend = (kt + TF_CHUNK_KB < KT) ? kt + TF_CHUNK_KB : KT;
          for (unsigned first = 1; kt < kend; kt++, it++, first = 0) {
            const int s = it % F2_STAGES; const unsigned ph = (it / F2_STAGES) & 1;
            mbar_wait(full0 + 8 * s, ph);                  // this CTA's converters
            mbar_wait_cluster(pfull0 + 8 * s, ph);         // the peer's, relayed
            tc_fence_after();
            const unsigned so = (unsigned)s * (F2_STAGE_BYTES >> 4);      // descriptor address field counts 16-byte units
            if (lane == 0) {
#pragma unroll
              for (int ks = 0; ks < TF_BK / 8; ks++) {
                const unsigned long long a_hi = adesc0 + so + ks * (32 >> 4), a_lo = a_hi + (TF_A_TILE >> 4);       // A: 32 B along the row
                const unsigned long long b_hi = bdesc0 + so + ks * (1024 >> 4), b_lo = b_hi + (F2_B_TILE >> 4);     // B: the next 8 k-rows
                tc_mma2_tf32(d_tmem, a_lo, b_hi, idesc, (ks == 0 && first) ? 0u : 1u);
                tc_mma2_tf32(d_tmem, a_hi, b_lo, idesc, 1u);
                tc_mma2_tf32(d_tmem, a_hi, b_hi, idesc, 1u);
              }
              tc_commit2(empty0 + 8 * s);
            }
            __syncwarp();
          }
          if (lane == 0) tc_commit2(tfull0 + 8 * buf);
          __syncwarp();
        }
      }
    }
  } else {
    // ---- epilogue: this CTA's 128 rows, TMEM lane quarter q = warp % 4, all 256 columns ------------------------
    // (eight warps on 32 x 16 slabs were tried: slower, 0.56 vs 0.43 ms on C4)
    const int q = warp & 3;
    const unsigned stg = smem_u32(smem) + F2_STAGES * F2_STAGE_BYTES + (warp - F2_EPI_WARP0) * TFF_STG_BYTES;
    const unsigned stg_w = stg + lane * (TFF_STG_PITCH * 4);
    const int rr = lane >> 3, c4 = lane & 7;
    const unsigned stg_r = stg + rr * (TFF_STG_PITCH * 4) + c4 * 16;
    unsigned g = 0;
    for (int t = pair; t < total_tiles; t += npairs) {
      int b = t / tiles_per_batch, tm, tn; raster(t - b * tiles_per_batch, MT2, NT, 8, tm, tn);
      float* cbase = C + (long long)b * sC + ((long long)(2 * tm + (int)rank) * TF_BM + q * 32 + rr) * ldc + (long long)tn * TF_BN + c4 * 4;
      for (int c = 0; c < NC; c++, g++) {
        const int buf = g & 1; const unsigned use = g >> 1;
        const bool add = (c > 0) || accumulate;
        mbar_wait(tfull0 + 8 * buf, use & 1);
        tc_fence_after();
#pragma unroll 1
        for (int cc = 0; cc < TF_BN / 32; cc++) {
          unsigned r[32];
          unsigned taddr = tmem_base + ((unsigned)(q * 32) << 16) + (unsigned)(buf * TF_BN + cc * 32);
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                       "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                       "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                       : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                         "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                         "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                         "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                       : "r"(taddr) : "memory");
          float* dst = cbase + cc * 32;
          float4 old[8];
          if (add) {                                      // the running sum so far, requested while the TMEM load is in flight
#pragma unroll
            for (int p = 0; p < 8; p++) old[p] = *(const float4*)(dst + (long long)(4 * p) * ldc);
          }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (cc == TF_BN / 32 - 1) {                     // everything is out of TMEM: hand the accumulator back early
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(tempty0_leader + 8 * buf);
          }
#pragma unroll
          for (int j = 0; j < 32; j += 4) sts128(stg_w + j * 4, make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]));
          __syncwarp();
#pragma unroll
          for (int p = 0; p < 8; p++) {
            uint4 v = lds128(stg_r + p * (4 * TFF_STG_PITCH * 4));
            float4 o = add ? old[p] : make_float4(0.f, 0.f, 0.f, 0.f);
            *(float4*)(dst + (long long)(4 * p) * ldc) = make_float4(__fadd_rn(o.x, __uint_as_float(v.x)), __fadd_rn(o.y, __uint_as_float(v.y)),
                                                                       __fadd_rn(o.z, __uint_as_float(v.z)), __fadd_rn(o.w, __uint_as_float(v.w)));
          }
          __syncwarp();
        }
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();                                  // the peer's smem and TMEM stay alive until both CTAs are done
  if (warp == F2_MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// (The pack + GEMM path was also tried on CTA pairs -- half of the packed B tile per CTA, three 64 KB stages:
// identical accuracy, no gain (8192^3: 4.66 vs 4.68 ms): that path is bound by the tensor pipe and the
// power-capped clock, not by shared memory.  Not kept.)
// NUMCL_TF32_MODE = auto | packed | fused  (debug/benchmark knob; `fused` still requires an eligible shape)
static int tf32_mode() {
  const char* e = getenv("NUMCL_TF32_MODE");
  if (!e) return 0;
  return !strcmp(e, "packed") ? 1 : !strcmp(e, "fused") ? 2 : !strcmp(e, "fused2") ? 3 : 0;
}

static int gemm_f32_tf32x3_fused(const GemmParams& p) {
  Context& c = ctx();
  int MT = p.M / TF_BM, NT = p.N / TF_BN, KT = p.K / TF_BK;
  long long total_tiles = (long long)MT * NT * p.batch;
  if (total_tiles > 0x7fffffffLL) return NCL_ERR_UNSUPPORTED;
  static bool attr_set = false;
  if (!attr_set) { NCL_CUDA(cudaFuncSetAttribute(gemm_tf32x3_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TFF_SMEM)); attr_set = true; }
  unsigned grid = (unsigned)(total_tiles < c.sm_count ? total_tiles : c.sm_count);
  gemm_tf32x3_fused_kernel<<<grid, TFF_THREADS, TFF_SMEM, c.stream>>>((const float*)p.A, (const float*)p.B, (float*)p.C, KT, p.lda, p.ldb, p.ldc,
                                                                       p.sA, p.sB, p.sC, MT, NT, p.batch, p.accumulate);
  note_launch("gemm_tf32x3_fused_tcgen05", 1);
  return cuda_fail(cudaGetLastError(), "gemm_tf32x3_fused_kernel");
}

static int gemm_f32_tf32x3_fused2(const GemmParams& p) {
  Context& c = ctx();
  int MT2 = p.M / 256, NT = p.N / TF_BN, KT = p.K / TF_BK;
  long long total_tiles = (long long)MT2 * NT * p.batch;
  if (total_tiles > 0x7fffffffLL) return NCL_ERR_UNSUPPORTED;
  static bool attr_set = false;
  if (!attr_set) { NCL_CUDA(cudaFuncSetAttribute(gemm_tf32x3_fused2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F2_SMEM)); attr_set = true; }
  long long pairs = c.sm_count / 2;
  if (total_tiles < pairs) pairs = total_tiles;
  cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3((unsigned)(2 * pairs)); cfg.blockDim = dim3(F2_THREADS); cfg.dynamicSmemBytes = F2_SMEM; cfg.stream = c.stream;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tf32x3_fused2_kernel, (const float*)p.A, (const float*)p.B, (float*)p.C, KT, p.lda, p.ldb, p.ldc,
                                     p.sA, p.sB, p.sC, MT2, NT, p.batch, p.accumulate);
  if (e != cudaSuccess) { cudaGetLastError(); return NCL_ERR_UNSUPPORTED; }     // no CTA pairs on this device/partition: single-CTA kernel
  note_launch("gemm_tf32x3_fused_2cta_tcgen05", 1);
  return cuda_fail(cudaGetLastError(), "gemm_tf32x3_fused2_kernel");
}

static bool fused_eligible(const GemmParams& p) {
  if (p.M % TF_BM || p.N % TF_BN || p.K % TF_BK || p.K > TF_FUSED_MAX_K) return false;
  if ((p.lda & 3) || (p.ldb & 3) || (p.ldc & 3) || (p.sA & 3) || (p.sB & 3) || (p.sC & 3)) return false;
  if (((uintptr_t)p.A & 15) || ((uintptr_t)p.B & 15) || ((uintptr_t)p.C & 15)) return false;
  return true;
}

int gemm_f32_tf32x3(const GemmParams& p) {
  Context& c = ctx();
  if (p.lkA != LK_F32 || p.lkB != LK_F32 || p.lkC != LK_F32) return NCL_ERR_UNSUPPORTED;
  if (((uintptr_t)p.C & 3) || ((uintptr_t)p.A & 3) || ((uintptr_t)p.B & 3)) return NCL_ERR_UNSUPPORTED;
  {
    // the pack pre-pass costs 3x the operand bytes; per output tile that is worth it only when the
    // operands are re-used by many tiles (large M and N).  Small matrices: split inside the kernel.
    int mode = tf32_mode();
    const long long mn = (long long)p.M * p.N;
    if (mode != 1 && fused_eligible(p)) {
      // measured crossovers (tools/gemm_fused_check.py): CTA pairs win up to 2048^2 outputs per matrix, the
      // single-CTA kernel up to 1024^2, the pack pre-pass beyond
      if (p.M % 256 == 0 && (mode == 3 || (mode == 0 && mn <= 2048LL * 2048LL))) {
        int rc = gemm_f32_tf32x3_fused2(p);
        if (rc != NCL_ERR_UNSUPPORTED) return rc;
      }
      if (mode == 2 || mode == 3 || mn <= 1024LL * 1024LL) return gemm_f32_tf32x3_fused(p);
    }
  }
  int MT = (p.M + TF_BM - 1) / TF_BM, NT = (p.N + TF_BN - 1) / TF_BN, KT = (p.K + TF_BK - 1) / TF_BK;
  if ((long long)MT * NT > 0x7fffffffLL || p.batch > 65535 || MT > 65535) return NCL_ERR_UNSUPPORTED;
  size_t a_bytes = (size_t)p.batch * MT * KT * 2 * TF_A_TILE, b_bytes = (size_t)p.batch * NT * KT * 2 * TF_B_TILE;
  NCL_TRY(ensure_scratch(a_bytes + b_bytes + 2048));
  char* Ap = (char*)(((uintptr_t)c.scratch + 1023) & ~(uintptr_t)1023);
  char* Bp = Ap + a_bytes;
  static bool attr_set = false;
  if (!attr_set) { NCL_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TF_SMEM)); attr_set = true; }
  pack_a_kernel<<<dim3(KT, MT, p.batch), 256, 0, c.stream>>>((const float*)p.A, Ap, p.M, p.K, p.lda, p.sA, MT, KT);
  int nslabs = NT * (TF_BN / 64);
  pack_b_kernel<<<dim3(KT, nslabs, p.batch), 256, 0, c.stream>>>((const float*)p.B, Bp, p.K, p.N, p.ldb, p.sB, NT, KT);
  NCL_TRY(cuda_fail(cudaGetLastError(), "pack kernels"));
  long long total_tiles = (long long)MT * NT * p.batch;
  if (total_tiles > 0x7fffffffLL) return NCL_ERR_UNSUPPORTED;
  unsigned grid = (unsigned)(total_tiles < c.sm_count ? total_tiles : c.sm_count);
  gemm_tf32x3_kernel<<<grid, TF_THREADS, TF_SMEM, c.stream>>>(Ap, Bp, (float*)p.C, p.M, p.N, KT, p.ldc, p.sC, MT, NT, p.batch, p.accumulate);
  note_launch("gemm_tf32x3_tcgen05", 3);
  return cuda_fail(cudaGetLastError(), "gemm_tf32x3_kernel");
}
