// Scratch microbenchmark: random 512-byte row gathers from local HBM / peer HBM (same process, peer access enabled),
// with ld.global.nc (__ldg) and plain ld.global.  nvcc -O3 -arch=sm_100a tools/p2p_gather.cu -o tools/p2p_gather
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <bool NC, int ROWS>
__global__ void gather(const float4* __restrict__ src, uint32_t nrows, uint32_t iters, float* out) {
  const int lane = threadIdx.x & 31, t = lane & 7, grp = lane >> 3;
  uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) / 8 * 2654435761u + 12345u;
  float acc = 0.f;
  for (uint32_t it = 0; it < iters; ++it) {
    float4 v[ROWS][4];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      s = s * 1664525u + 1013904223u;
      const uint32_t row = __shfl_sync(0xffffffffu, s, grp * 8) % nrows;
      const float4* p = src + (size_t)row * 32 + t;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        if (NC) v[r][b] = __ldg(p + 8 * b);
        else asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[r][b].x), "=f"(v[r][b].y), "=f"(v[r][b].z), "=f"(v[r][b].w) : "l"(p + 8 * b));
      }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc += v[r][b].x + v[r][b].w;
  }
  if (acc == 123.456f) out[0] = acc;
}

template <bool NC, int ROWS>
int run(const char* name, const float4* src, uint32_t nrows, int blocks) {
  float* out; CK(cudaMalloc(&out, 4));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const uint32_t iters = 200;
  gather<NC, ROWS><<<blocks, 128>>>(src, nrows, 10, out);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(a);
  gather<NC, ROWS><<<blocks, 128>>>(src, nrows, iters, out);
  cudaEventRecord(b);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double bytes = (double)blocks * 4 * 4 * ROWS * iters * 512.0;
  printf("%-28s rows/pass %d blocks %5d: %8.1f GB/s\n", name, ROWS, blocks, bytes / ms / 1e6);
  cudaFree(out);
  return 0;
}

int main() {
  const uint32_t nrows = 8u << 20;  // 4 GiB of 512-byte rows
  float4 *loc, *rem;
  CK(cudaSetDevice(1)); CK(cudaMalloc(&rem, (size_t)nrows * 512)); CK(cudaMemset(rem, 0, (size_t)nrows * 512));
  CK(cudaSetDevice(0)); CK(cudaMalloc(&loc, (size_t)nrows * 512)); CK(cudaMemset(loc, 0, (size_t)nrows * 512));
  CK(cudaDeviceEnablePeerAccess(1, 0));
  for (int blocks : {148 * 4, 148 * 8}) {
    run<true, 2>("local  ld.global.nc", loc, nrows, blocks);
    run<true, 4>("local  ld.global.nc", loc, nrows, blocks);
    run<true, 2>("peer   ld.global.nc", rem, nrows, blocks);
    run<true, 4>("peer   ld.global.nc", rem, nrows, blocks);
    run<false, 2>("peer   ld.global", rem, nrows, blocks);
    run<false, 4>("peer   ld.global", rem, nrows, blocks);
  }
  return 0;
}
