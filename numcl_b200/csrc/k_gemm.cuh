// k_gemm.cuh -- shared description of a (batched) contraction C[b,m,n] (+)= sum_k A[b,m,k] * B[b,k,n]
// after try_gemm (k_gemm.cu) has mapped an einsum nest onto it.
#pragma once
#include "ncl_device.cuh"

struct GemmParams {
  const char* A; const char* B; char* C;      // base vectors with the nest origin applied (bytes)
  int M, N, K, batch;
  long long lda, ldb, ldc;                     // element strides between rows (A: m, B: k, C: m); k / n contiguous
  long long sA, sB, sC;                        // batch strides (elements)
  int accumulate;                              // 1: C holds caller data (src/3einsum.lisp:468); 0: fresh output
  int lkA, lkB, lkC; StoreSpec skC;            // element kinds (SIMT path)
};

int gemm_f64_dmma(const GemmParams& p);        // k_gemm_f64.cu
int gemm_f32_tf32x3(const GemmParams& p);      // k_gemm_tf32.cu
int gemm_simt(const GemmParams& p, int cls);   // k_gemm_simt.cu

// L2-friendly rasterisation: tiles are walked in column-major order inside groups of GROUP tile rows,
// so concurrently resident CTAs share A row panels and B column panels in the 126 MB L2.
__device__ __forceinline__ void raster(int t, int tiles_m, int tiles_n, int group, int& tm, int& tn) {
  int per_group = group * tiles_n;
  int g = t / per_group;
  int first_m = g * group;
  int size = tiles_m - first_m < group ? tiles_m - first_m : group;
  int r = t - g * per_group;
  tm = first_m + r % size;
  tn = r / size;
}
