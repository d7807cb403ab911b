// k_gemm_f64.cu -- K4, double-float: `(ij jk -> ik)` / `(bij bjk -> bik)` on the FP64 tensor path.
//
// Replaces the i,j,k nest  C[i,k] <- C[i,k] + A[i,j]*B[j,k]  (common-lisp.lisp:108-174, 226-240).
// tcgen05 has no FP64 kind, so the tensor path for doubles is the warp-level DMMA
// (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4; on sm_100a the wider m16n8k* shapes lower to the same
// instruction).  Compute-bound: 2*M*N*K flops over 8*(MK+KN+MN) bytes.
//
// CTA tile 128x128x16, 8 warps as 2(M) x 4(N), warp tile 64x32 = 8x4 DMMA tiles with the 64
// accumulator doubles of a thread in registers (=> one CTA per SM, 4-stage cp.async ring of
// 37 KB stages to keep the pipe fed, synchronised per stage with mbarriers -- no CTA-wide barrier in
// the k-loop: 0.914 -> 0.956 of cuBLAS f64 at 8192^3).  Shared-memory pitches are padded (A: 16+4, B: 128+4
// doubles) so the 64-bit fragment loads of a half-warp hit 16 distinct 8-byte banks.
// Accumulation is FMA inside the tensor pipe and k-blocked, i.e. not the reference's order:
// held to 1e-12 relative Frobenius error (north_star), not bit-exactness.
#include "k_gemm.cuh"

#define DG_BM 128
#define DG_BN 128
#define DG_BK 16
#define DG_STAGES 4
#define DG_PA (DG_BK + 4)          // A pitch in doubles
#define DG_PB (DG_BN + 4)          // B pitch in doubles
#define DG_STAGE_DOUBLES (DG_BM * DG_PA + DG_BK * DG_PB)
#define DG_SMEM (DG_STAGES * DG_STAGE_DOUBLES * 8)

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, int src_bytes) {     // rows that start on odd doubles
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// mbarrier helpers: the k-loop synchronises per STAGE (full: this stage's cp.asyncs have landed; empty: all eight
// warps have read it) instead of with one CTA-wide barrier per k-tile, so warps may drift by a k-tile
__device__ __forceinline__ unsigned dg_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dg_mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void dg_mbar_arrive(unsigned bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void dg_cp_async_arrive(unsigned bar) { asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void dg_mbar_wait(unsigned bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <bool A16>
__global__ void __launch_bounds__(256, 1) dgemm_dmma_kernel(const __grid_constant__ GemmParams p, int tiles_m, int tiles_n, int cvec) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;           // 2 x 4 warps
  const int g = lane >> 2, t4 = lane & 3;
  int tile = blockIdx.x, b = blockIdx.y;
  int tm, tn; raster(tile, tiles_m, tiles_n, 8, tm, tn);
  const int m0 = tm * DG_BM, n0 = tn * DG_BN;
  const double* A = (const double*)p.A + (long long)b * p.sA;
  const double* B = (const double*)p.B + (long long)b * p.sB;
  double* C = (double*)p.C + (long long)b * p.sC;

  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int ktiles = (p.K + DG_BK - 1) / DG_BK;
  auto load_tile = [&](int stage, int kt) {
    double* sA = smem + stage * DG_STAGE_DOUBLES;
    double* sB = sA + DG_BM * DG_PA;
    const int k0 = kt * DG_BK;
    if constexpr (A16) {
#pragma unroll
      for (int i = 0; i < 4; i++) {                                   // A: 128 rows x 8 chunks of 2 doubles
        int idx = tid + i * 256; int row = idx >> 3, c = idx & 7;
        int m = m0 + row, k = k0 + c * 2;
        int bytes = (m < p.M && kt < ktiles) ? ((p.K - k) >= 2 ? 16 : ((p.K - k) == 1 ? 8 : 0)) : 0;
        const double* src = bytes ? A + (long long)m * p.lda + k : A;
        cp_async16(sA + row * DG_PA + c * 2, src, bytes);
      }
#pragma unroll
      for (int i = 0; i < 4; i++) {                                   // B: 16 rows x 64 chunks of 2 doubles
        int idx = tid + i * 256; int row = idx >> 6, c = idx & 63;
        int k = k0 + row, n = n0 + c * 2;
        int bytes = (k < p.K && kt < ktiles) ? ((p.N - n) >= 2 ? 16 : ((p.N - n) == 1 ? 8 : 0)) : 0;
        const double* src = bytes ? B + (long long)k * p.ldb + n : B;
        cp_async16(sB + row * DG_PB + c * 2, src, bytes);
      }
    } else {
      // leading dimensions / bases that are only 8-byte aligned (odd-sized matrices): one double per copy
#pragma unroll
      for (int i = 0; i < 8; i++) {                                   // A: 128 rows x 16 doubles
        int idx = tid + i * 256; int row = idx >> 4, c = idx & 15;
        int m = m0 + row, k = k0 + c;
        int bytes = (m < p.M && k < p.K && kt < ktiles) ? 8 : 0;
        const double* src = bytes ? A + (long long)m * p.lda + k : A;
        cp_async8(sA + row * DG_PA + c, src, bytes);
      }
#pragma unroll
      for (int i = 0; i < 8; i++) {                                   // B: 16 rows x 128 doubles
        int idx = tid + i * 256; int row = idx >> 7, c = idx & 127;
        int k = k0 + row, n = n0 + c;
        int bytes = (k < p.K && n < p.N && kt < ktiles) ? 8 : 0;
        const double* src = bytes ? B + (long long)k * p.ldb + n : B;
        cp_async8(sB + row * DG_PB + c, src, bytes);
      }
    }
  };

  __shared__ __align__(8) unsigned long long bars[2 * DG_STAGES];
  const unsigned full0 = dg_smem_u32(&bars[0]), empty0 = dg_smem_u32(&bars[DG_STAGES]);
  if (tid == 0) {
    for (int s = 0; s < DG_STAGES; s++) { dg_mbar_init(full0 + 8 * s, 256); dg_mbar_init(empty0 + 8 * s, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
#pragma unroll
  for (int s = 0; s < DG_STAGES - 1; s++) { load_tile(s, s); dg_cp_async_arrive(full0 + 8 * s); }

  for (int kt = 0; kt < ktiles; kt++) {
    {
      // refill the stage k-tile kt-1 was read from (with k-tile kt+3) once every warp has released it
      const int ls = (kt + DG_STAGES - 1) % DG_STAGES;
      if (kt > 0) dg_mbar_wait(empty0 + 8 * ls, ((kt - 1) / DG_STAGES) & 1);
      load_tile(ls, kt + DG_STAGES - 1);
      dg_cp_async_arrive(full0 + 8 * ls);
    }
    const int cs = kt % DG_STAGES;
    dg_mbar_wait(full0 + 8 * cs, (kt / DG_STAGES) & 1);
    const double* sA = smem + cs * DG_STAGE_DOUBLES;
    const double* sB = sA + DG_BM * DG_PA;
    const double* pa = sA + (wm * 64 + g) * DG_PA + t4;
    const double* pb = sB + t4 * DG_PB + wn * 32 + g;
#pragma unroll
    for (int kk = 0; kk < DG_BK / 4; kk++) {
      double a[8], bb[4];
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = pa[i * 8 * DG_PA + kk * 4];
#pragma unroll
      for (int j = 0; j < 4; j++) bb[j] = pb[kk * 4 * DG_PB + j * 8];
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], bb[j]);
    }
    __syncwarp();
    if (lane == 0) dg_mbar_arrive(empty0 + 8 * cs);
  }
  cp_async_wait<0>();

  // epilogue: thread owns C[m0 + wm*64 + i*8 + g][n0 + wn*32 + j*8 + t4*2 + {0,1}]
#pragma unroll
  for (int i = 0; i < 8; i++) {
    int m = m0 + wm * 64 + i * 8 + g;
    if (m >= p.M) continue;
    double* crow = C + (long long)m * p.ldc;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int n = n0 + wn * 32 + j * 8 + t4 * 2;
      if (n + 1 < p.N && cvec) {
        double2 v = make_double2(acc[i][j][0], acc[i][j][1]);
        if (p.accumulate) { double2 o = *(const double2*)(crow + n); v.x = __dadd_rn(o.x, v.x); v.y = __dadd_rn(o.y, v.y); }
        else { v.x = __dadd_rn(0.0, v.x); v.y = __dadd_rn(0.0, v.y); }       // (+ 0 acc): -0.0 -> +0.0 like the reference
        *(double2*)(crow + n) = v;
      } else {
#pragma unroll
        for (int e = 0; e < 2; e++) if (n + e < p.N) {
          double v = acc[i][j][e];
          crow[n + e] = p.accumulate ? __dadd_rn(crow[n + e], v) : __dadd_rn(0.0, v);
        }
      }
    }
  }
}

int gemm_f64_dmma(const GemmParams& p) {
  // cp.async moves 16-byte chunks when every row starts 16B aligned, single doubles otherwise (odd-sized matrices)
  if (((uintptr_t)p.A & 7) || ((uintptr_t)p.B & 7) || ((uintptr_t)p.C & 7)) return NCL_ERR_UNSUPPORTED;
  const bool a16 = !((p.lda & 1) || (p.ldb & 1) || (p.sA & 1) || (p.sB & 1) || ((uintptr_t)p.A & 15) || ((uintptr_t)p.B & 15));
  int cvec = ((p.ldc & 1) == 0 && (p.sC & 1) == 0 && ((uintptr_t)p.C & 15) == 0) ? 1 : 0;   // double2 stores allowed
  Context& c = ctx();
  static bool attr_set = false;
  if (!attr_set) {
    NCL_CUDA(cudaFuncSetAttribute(dgemm_dmma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DG_SMEM));
    NCL_CUDA(cudaFuncSetAttribute(dgemm_dmma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DG_SMEM));
    attr_set = true;
  }
  int tiles_m = (p.M + DG_BM - 1) / DG_BM, tiles_n = (p.N + DG_BN - 1) / DG_BN;
  if ((long long)tiles_m * tiles_n > 0x7fffffffLL || p.batch > 65535) return NCL_ERR_UNSUPPORTED;
  dim3 grid((unsigned)(tiles_m * tiles_n), (unsigned)p.batch);
  if (a16) dgemm_dmma_kernel<true><<<grid, 256, DG_SMEM, c.stream>>>(p, tiles_m, tiles_n, cvec);
  else dgemm_dmma_kernel<false><<<grid, 256, DG_SMEM, c.stream>>>(p, tiles_m, tiles_n, cvec);
  note_launch("gemm_f64_dmma");
  return cuda_fail(cudaGetLastError(), "dgemm_dmma_kernel");
}
