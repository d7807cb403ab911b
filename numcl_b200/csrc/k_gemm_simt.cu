// k_gemm_simt.cu -- K4 fallback on the ordinary FMA / IMAD pipes: skinny or unaligned float shapes
// and every INTEGER contraction (the reference's matmul golden vector is fixnum,
// t/package.lisp:514-526; an integer einsum output is always fixnum, 3einsum.lisp:517-518).
//
// 64x64x16 CTA tile, 256 threads, 4x4 register micro-tile per thread.  Elements are read through
// the value class, so ub8 / fixnum(tagged) / sb32 ... operands need no conversion pass.  Integers
// accumulate exactly modulo 2^64 and wrap to the output width once at the end, which equals the
// reference's wrap-after-every-step (modular arithmetic); floats use FMA and are held to the
// contraction tolerance.
#include "k_gemm.cuh"

#define SG_BM 64
#define SG_BN 64
#define SG_BK 16

template <class T> __device__ __forceinline__ T mad(T a, T b, T c);
template <> __device__ __forceinline__ float mad<float>(float a, float b, float c) { return __fmaf_rn(a, b, c); }
template <> __device__ __forceinline__ double mad<double>(double a, double b, double c) { return __fma_rn(a, b, c); }
template <> __device__ __forceinline__ long long mad<long long>(long long a, long long b, long long c) {
  return (long long)((unsigned long long)a * (unsigned long long)b + (unsigned long long)c);
}

template <class T>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const __grid_constant__ GemmParams p, int tiles_m, int tiles_n) {
  __shared__ T As[SG_BK][SG_BM + 4];
  __shared__ T Bs[SG_BK][SG_BN + 4];
  int tm, tn; raster(blockIdx.x, tiles_m, tiles_n, 8, tm, tn);
  const int b = blockIdx.y;
  const int m0 = tm * SG_BM, n0 = tn * SG_BN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const long long aoff = (long long)b * p.sA, boff = (long long)b * p.sB, coff = (long long)b * p.sC;
  T acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = (T)0;
  for (int k0 = 0; k0 < p.K; k0 += SG_BK) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      int idx = threadIdx.x + i * 256;
      int am = idx >> 4, ak = idx & 15;                 // A tile: 64 rows x 16 k (k fastest: contiguous reads)
      T v = (T)0;
      if (m0 + am < p.M && k0 + ak < p.K) v = load_scalar<T>(p.A, aoff + (long long)(m0 + am) * p.lda + k0 + ak, p.lkA);
      As[ak][am] = v;
      int bk = idx >> 6, bn = idx & 63;                 // B tile: 16 k x 64 cols (n fastest)
      T w = (T)0;
      if (k0 + bk < p.K && n0 + bn < p.N) w = load_scalar<T>(p.B, boff + (long long)(k0 + bk) * p.ldb + n0 + bn, p.lkB);
      Bs[bk][bn] = w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SG_BK; k++) {
      T a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; j++) bb[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = mad<T>(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      long long ci = coff + (long long)m * p.ldc + n;
      T init = p.accumulate ? load_scalar<T>(p.C, ci, p.lkC) : (T)0;
      store_scalar<T>(p.C, ci, p.skC, Ops<T>::add(init, acc[i][j]));
    }
  }
}

int gemm_simt(const GemmParams& p, int cls) {
  Context& c = ctx();
  int tiles_m = (p.M + SG_BM - 1) / SG_BM, tiles_n = (p.N + SG_BN - 1) / SG_BN;
  if ((long long)tiles_m * tiles_n > 0x7fffffffLL || p.batch > 65535) return NCL_ERR_UNSUPPORTED;
  dim3 grid((unsigned)(tiles_m * tiles_n), (unsigned)p.batch);
  if (cls == CLS_F) gemm_simt_kernel<float><<<grid, 256, 0, c.stream>>>(p, tiles_m, tiles_n);
  else if (cls == CLS_D) gemm_simt_kernel<double><<<grid, 256, 0, c.stream>>>(p, tiles_m, tiles_n);
  else gemm_simt_kernel<long long><<<grid, 256, 0, c.stream>>>(p, tiles_m, tiles_n);
  note_launch(cls == CLS_I ? "gemm_int_simt" : "gemm_fma_simt");
  return cuda_fail(cudaGetLastError(), "gemm_simt_kernel");
}
