// Scratch: random 512-byte row gathers from a peer GPU's memory mapped through CUDA IPC, K chunks of S bytes.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <unistd.h>
#include <sys/wait.h>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void gather(const float4* const* chunks, uint32_t nchunks, uint32_t rows_per_chunk, uint32_t iters, float* out) {
  const int lane = threadIdx.x & 31, t = lane & 7, grp = lane >> 3;
  uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) / 8 * 2654435761u + 12345u;
  float acc = 0.f;
  for (uint32_t it = 0; it < iters; ++it) {
    float4 v[2][4];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      s = s * 1664525u + 1013904223u;
      const uint32_t x = __shfl_sync(0xffffffffu, s, grp * 8);
      const float4* p = chunks[(x >> 8) % nchunks] + (size_t)(x % rows_per_chunk) * 32 + t;
#pragma unroll
      for (int b = 0; b < 4; ++b) v[r][b] = __ldg(p + 8 * b);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc += v[r][b].x + v[r][b].w;
  }
  if (acc == 123.456f) out[0] = acc;
}

int main(int argc, char** argv) {
  const int K = atoi(argv[1]);
  const size_t S = (size_t)atof(argv[2]) * (1u << 20);  // MiB per chunk
  int to_child[2], to_parent[2];
  pipe(to_child); pipe(to_parent);
  if (fork() == 0) {  // exporter on GPU 1
    CK(cudaSetDevice(1));
    const size_t dummy_gb = argc > 3 ? atoi(argv[3]) : 0;
    void* dummy = nullptr;
    if (dummy_gb) { CK(cudaMalloc(&dummy, dummy_gb << 30)); CK(cudaMemset(dummy, 1, dummy_gb << 30)); }
    const bool free_dummy = argc > 4;
    for (int k = 0; k < K; ++k) {
      void* p; CK(cudaMalloc(&p, S)); CK(cudaMemset(p, 0, S));
      cudaIpcMemHandle_t h; CK(cudaIpcGetMemHandle(&h, p));
      write(to_parent[1], &h, sizeof h);
    }
    if (dummy && free_dummy) CK(cudaFree(dummy));
    CK(cudaDeviceSynchronize());
    char c; read(to_child[0], &c, 1);
    return 0;
  }
  CK(cudaSetDevice(0));
  const float4* ptrs[64];
  for (int k = 0; k < K; ++k) {
    cudaIpcMemHandle_t h; read(to_parent[0], &h, sizeof h);
    void* p; CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ptrs[k] = (const float4*)p;
  }
  const float4** d_ptrs; CK(cudaMalloc(&d_ptrs, sizeof ptrs)); CK(cudaMemcpy(d_ptrs, ptrs, sizeof ptrs, cudaMemcpyHostToDevice));
  float* out; CK(cudaMalloc(&out, 4));
  const uint32_t rows = (uint32_t)(S / 512), iters = 200;
  const int blocks = 148 * 8;
  gather<<<blocks, 128>>>(d_ptrs, K, rows, 10, out); CK(cudaDeviceSynchronize());
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a); gather<<<blocks, 128>>>(d_ptrs, K, rows, iters, out); cudaEventRecord(b); CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, a, b);
  printf("IPC peer gather: %d chunk(s) x %.0f MiB (exporter dummy %s GiB%s): %8.1f GB/s\n", K, S / 1048576.0, argc > 3 ? argv[3] : "0", argc > 4 ? ", freed" : "", (double)blocks * 16 * 2 * iters * 512.0 / ms / 1e6);
  write(to_child[1], "x", 1);
  wait(nullptr);
  return 0;
}
