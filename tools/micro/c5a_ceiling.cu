// Micro-benchmark: what a hand-specialised kernel gets on C5a (u8 x fixnum row vector -> tagged fixnum) and C1
// (f32 + f32 row vector).  Scratch tool: measures the ceiling the generic elementwise engine is compared with.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/c5a_ceiling tools/micro/c5a_ceiling.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <algorithm>
#include <vector>

// out[i] = tag( x[i] * untag(y[i % 256]) ), 2 elements (16 bytes of output) per thread per step
__global__ void c5a_simple(const unsigned char* __restrict__ x, const long long* __restrict__ y, long long* __restrict__ out, long long n2) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x, s = (long long)gridDim.x * blockDim.x;
  for (; i < n2; i += s) {
    unsigned short v = __ldg((const unsigned short*)x + i);
    long long y0 = __ldg(y + ((2 * i) & 255)) >> 1, y1 = __ldg(y + ((2 * i + 1) & 255)) >> 1;
    longlong2 o; o.x = ((long long)(v & 0xff) * y0) << 1; o.y = ((long long)(v >> 8) * y1) << 1;
    ((longlong2*)out)[i] = o;
  }
}
// 16 input bytes -> 8 x 16-byte stores per thread per step (more bytes in flight per thread)
__global__ void c5a_wide(const unsigned char* __restrict__ x, const long long* __restrict__ y, long long* __restrict__ out, long long n16) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x, s = (long long)gridDim.x * blockDim.x;
  for (; i < n16; i += s) {
    uint4 v = __ldg((const uint4*)x + i);
    unsigned w[4] = {v.x, v.y, v.z, v.w};
    const int yb = (int)((16 * i) & 255);
#pragma unroll
    for (int k = 0; k < 8; k++) {
      unsigned b0 = (w[k >> 1] >> ((k & 1) * 16)) & 0xff, b1 = (w[k >> 1] >> ((k & 1) * 16 + 8)) & 0xff;
      long long y0 = __ldg(y + yb + 2 * k) >> 1, y1 = __ldg(y + yb + 2 * k + 1) >> 1;
      longlong2 o; o.x = ((long long)b0 * y0) << 1; o.y = ((long long)b1 * y1) << 1;
      ((longlong2*)out)[i * 8 + k] = o;
    }
  }
}
// warp-coalesced variant of the above: lane handles 16-byte output chunks l, l+32, .. of a 128-chunk segment (the engine's layout)
__global__ void c5a_seg(const unsigned char* __restrict__ x, const long long* __restrict__ y, long long* __restrict__ out, long long nseg) {
  const int lane = threadIdx.x & 31;
  long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
  long long yv[4][2];
#pragma unroll
  for (int u = 0; u < 4; u++) { yv[u][0] = __ldg(y + (((lane + 32 * u) * 2) & 255)) >> 1; yv[u][1] = __ldg(y + (((lane + 32 * u) * 2 + 1) & 255)) >> 1; }
  for (; w < nseg; w += nw) {
    unsigned short v[4];
#pragma unroll
    for (int u = 0; u < 4; u++) v[u] = __ldg((const unsigned short*)x + w * 128 + u * 32 + lane);
#pragma unroll
    for (int u = 0; u < 4; u++) {
      longlong2 o; o.x = ((long long)(v[u] & 0xff) * yv[u][0]) << 1; o.y = ((long long)(v[u] >> 8) * yv[u][1]) << 1;
      ((longlong2*)out)[w * 128 + u * 32 + lane] = o;
    }
  }
}
// as c5a_seg, but the NEXT segment's input is requested before the current segment's stores are issued
__global__ void c5a_seg_pipe(const unsigned char* __restrict__ x, const long long* __restrict__ y, long long* __restrict__ out, long long nseg) {
  const int lane = threadIdx.x & 31;
  long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
  long long yv[4][2];
#pragma unroll
  for (int u = 0; u < 4; u++) { yv[u][0] = __ldg(y + (((lane + 32 * u) * 2) & 255)) >> 1; yv[u][1] = __ldg(y + (((lane + 32 * u) * 2 + 1) & 255)) >> 1; }
  unsigned short v[4], nv[4];
  if (w < nseg) {
#pragma unroll
    for (int u = 0; u < 4; u++) v[u] = __ldg((const unsigned short*)x + w * 128 + u * 32 + lane);
  }
  for (; w < nseg; w += nw) {
    if (w + nw < nseg) {
#pragma unroll
      for (int u = 0; u < 4; u++) nv[u] = __ldg((const unsigned short*)x + (w + nw) * 128 + u * 32 + lane);
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      longlong2 o; o.x = ((long long)(v[u] & 0xff) * yv[u][0]) << 1; o.y = ((long long)(v[u] >> 8) * yv[u][1]) << 1;
      ((longlong2*)out)[w * 128 + u * 32 + lane] = o;
    }
#pragma unroll
    for (int u = 0; u < 4; u++) v[u] = nv[u];
  }
}
__global__ void c1_simple(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ out, long long n4) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x, s = (long long)gridDim.x * blockDim.x;
  for (; i < n4; i += s) {
    float4 p = __ldg(a + i), q = __ldg(b + (i & 1023));
    out[i] = make_float4(__fadd_rn(__fadd_rn(0.f, p.x), q.x), __fadd_rn(__fadd_rn(0.f, p.y), q.y), __fadd_rn(__fadd_rn(0.f, p.z), q.z), __fadd_rn(__fadd_rn(0.f, p.w), q.w));
  }
}

int main() {
  const long long n = 64ll * 32 * 32 * 256;          // 16.7M elements
  const int SETS = 6;                                // rotate > L2
  unsigned char* x[SETS]; long long* o[SETS]; long long* y;
  for (int s = 0; s < SETS; s++) { cudaMalloc(&x[s], n); cudaMalloc(&o[s], n * 8); cudaMemset(x[s], 3 + s, n); }
  cudaMalloc(&y, 256 * 8); cudaMemset(y, 1, 256 * 8);
  float *a[SETS], *c[SETS], *b; const long long n1 = 4096ll * 4096;
  for (int s = 0; s < SETS; s++) { cudaMalloc(&a[s], n1 * 4); cudaMalloc(&c[s], n1 * 4); cudaMemset(a[s], 0, n1 * 4); }
  cudaMalloc(&b, 4096 * 4); cudaMemset(b, 0, 4096 * 4);
  cudaStream_t st; cudaStreamCreate(&st);
  auto run = [&](const char* name, auto launch, double bytes) {
    // CUDA-graph replay of 24 back-to-back launches over rotating buffer sets (same protocol as bench.py)
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
    for (int i = 0; i < 24; i++) launch(i % SETS);
    cudaStreamEndCapture(st, &g); cudaGraphInstantiate(&ge, g, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaGraphLaunch(ge, st); cudaStreamSynchronize(st);
    std::vector<float> t;
    for (int r = 0; r < 5; r++) { cudaEventRecord(e0, st); cudaGraphLaunch(ge, st); cudaEventRecord(e1, st); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); t.push_back(ms / 24); }
    std::sort(t.begin(), t.end());
    printf("%-28s %7.2f us  %6.0f GB/s  (%.3f of 6450)\n", name, t[2] * 1e3, bytes / t[2] / 1e6, bytes / t[2] / 1e6 / 6450.3);
  };
  const double b5 = 150996992.0, b1 = 134234112.0;
  for (int bps : {4, 8}) {
    int grid = 148 * bps; char nm[64];
    snprintf(nm, 64, "c5a_simple grid=%d", grid); run(nm, [&](int s) { c5a_simple<<<grid, 256, 0, st>>>(x[s], y, o[s], n / 2); }, b5);
    snprintf(nm, 64, "c5a_wide   grid=%d", grid); run(nm, [&](int s) { c5a_wide<<<grid, 256, 0, st>>>(x[s], y, o[s], n / 16); }, b5);
    snprintf(nm, 64, "c5a_seg    grid=%d", grid); run(nm, [&](int s) { c5a_seg<<<grid, 256, 0, st>>>(x[s], y, o[s], n / 256); }, b5);
    snprintf(nm, 64, "c5a_segpipe grid=%d", grid); run(nm, [&](int s) { c5a_seg_pipe<<<grid, 256, 0, st>>>(x[s], y, o[s], n / 256); }, b5);
    snprintf(nm, 64, "c1_simple  grid=%d", grid); run(nm, [&](int s) { c1_simple<<<grid, 256, 0, st>>>((const float4*)a[s], (const float4*)b, (float4*)c[s], n1 / 4); }, b1);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
