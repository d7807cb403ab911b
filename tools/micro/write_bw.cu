// Micro-benchmark: HBM write-only bandwidth for different store shapes (scratch tool, not part of the library).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/write_bw tools/micro/write_bw.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <algorithm>
#include <vector>

__global__ void st16(uint4* p, long long n16, unsigned v) {           // 16-byte stores, grid-stride
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x, s = (long long)gridDim.x * blockDim.x;
  uint4 x = make_uint4(v, v + 1, v + 2, v + 3);
  for (; i < n16; i += s) p[i] = x;
}
__global__ void st16cs(uint4* p, long long n16, unsigned v) {         // streaming hint
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x, s = (long long)gridDim.x * blockDim.x;
  uint4 x = make_uint4(v, v + 1, v + 2, v + 3);
  for (; i < n16; i += s) __stcs(p + i, x);
}
__global__ void st32(char* p, long long n32, unsigned v) {            // 32-byte stores (st.global.v8.b32)
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x, s = (long long)gridDim.x * blockDim.x;
  for (; i < n32; i += s)
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p + i * 32), "r"(v), "r"(v + 1), "r"(v + 2), "r"(v + 3),
                 "r"(v + 4), "r"(v + 5), "r"(v + 6), "r"(v + 7) : "memory");
}
__global__ void st16x4(uint4* p, long long n16, unsigned v) {         // each warp writes 4 consecutive 512-byte rows per iteration
  long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
  int lane = threadIdx.x & 31;
  uint4 x = make_uint4(v, v + 1, v + 2, v + 3);
  for (long long b = warp * 128; b < n16; b += nw * 128) {
#pragma unroll
    for (int u = 0; u < 4; u++) if (b + u * 32 + lane < n16) p[b + u * 32 + lane] = x;
  }
}
__global__ void cp16(const uint4* __restrict__ a, uint4* p, long long n16) {   // copy: read + write
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x, s = (long long)gridDim.x * blockDim.x;
  for (; i < n16; i += s) p[i] = a[i];
}

template <class F> float timeit(F f, int iters = 20) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; i++) f(i);
  std::vector<float> t;
  for (int i = 0; i < iters; i++) { cudaEventRecord(e0); f(i); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); t.push_back(ms); }
  std::sort(t.begin(), t.end());
  return t[t.size() / 2];
}

int main() {
  const long long bytes = 1ll << 30;
  char *a, *b; cudaMalloc(&a, bytes); cudaMalloc(&b, bytes);
  cudaMemset(a, 1, bytes);
  int sms = 148;
  for (int bpsm : {2, 4, 8, 16, 32}) {
    int grid = sms * bpsm;
    float t;
    t = timeit([&](int i) { st16<<<grid, 256>>>((uint4*)b, bytes / 16, i); });       printf("grid %5d st16    %7.0f GB/s\n", grid, bytes / t / 1e6);
    t = timeit([&](int i) { st16cs<<<grid, 256>>>((uint4*)b, bytes / 16, i); });     printf("grid %5d st16.cs %7.0f GB/s\n", grid, bytes / t / 1e6);
    t = timeit([&](int i) { st32<<<grid, 256>>>(b, bytes / 32, i); });               printf("grid %5d st32    %7.0f GB/s\n", grid, bytes / t / 1e6);
    t = timeit([&](int i) { st16x4<<<grid, 256>>>((uint4*)b, bytes / 16, i); });     printf("grid %5d st16x4  %7.0f GB/s\n", grid, bytes / t / 1e6);
    t = timeit([&](int i) { cp16<<<grid, 256>>>((const uint4*)a, (uint4*)b, bytes / 32); });  printf("grid %5d copy    %7.0f GB/s (r+w, half buffer)\n", grid, bytes / t / 1e6);
  }
  float t = timeit([&](int i) { cudaMemsetAsync(b, 7 + i, bytes); });                printf("cudaMemset(nonzero) %7.0f GB/s\n", bytes / t / 1e6);
  t = timeit([&](int i) { cudaMemcpyAsync(b, a, bytes / 2, cudaMemcpyDeviceToDevice); }); printf("cudaMemcpy D2D      %7.0f GB/s (r+w)\n", bytes / t / 1e6);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
