// layout.cu — natural row order <-> the blocked order rows are stored in (graph.h row_pos).
#include "engine.h"

namespace shn {
namespace {

__global__ void to_layout_kernel(const float* __restrict__ src, float* __restrict__ dst, uint64_t n, uint32_t dim,
                                 uint32_t row_floats) {
  // one thread per stored float: coalesced writes, gathered reads (within one row: L1/L2 hits)
  const uint64_t total = n * row_floats;
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint64_t row = i / row_floats;
    const uint32_t pos = static_cast<uint32_t>(i - row * row_floats);
    // invert row_pos: blocks first, then the tail
    const uint32_t d16 = dim & ~15u, nblk = row_blocks(dim);
    float v = 0.f;
    if (pos < 32u * nblk) {
      const uint32_t e = row_elem(pos);
      if (e < d16) v = src[row * dim + e];
    } else if (pos - 32u * nblk < (dim & 15u)) {
      v = src[row * dim + d16 + (pos - 32u * nblk)];
    }
    dst[i] = v;
  }
}

__global__ void from_layout_kernel(const float* __restrict__ src, float* __restrict__ dst, uint64_t n, uint32_t dim,
                                   uint32_t row_floats) {
  const uint64_t total = n * dim;
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint64_t row = i / dim;
    const uint32_t e = static_cast<uint32_t>(i - row * dim);
    dst[i] = src[row * row_floats + row_pos(dim, e)];
  }
}

}  // namespace

cudaError_t rows_to_layout(const float* d_src, float* d_dst, uint64_t n, uint32_t dim, uint32_t row_f4, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  const uint64_t total = n * row_f4 * 4ull;
  const int grid = static_cast<int>(std::min<uint64_t>((total + 255) / 256, 148ull * 32));
  to_layout_kernel<<<grid, 256, 0, stream>>>(d_src, d_dst, n, dim, row_f4 * 4);
  return cudaGetLastError();
}

cudaError_t rows_from_layout(const float* d_src, float* d_dst, uint64_t n, uint32_t dim, uint32_t row_f4, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  const uint64_t total = n * dim;
  const int grid = static_cast<int>(std::min<uint64_t>((total + 255) / 256, 148ull * 32));
  from_layout_kernel<<<grid, 256, 0, stream>>>(d_src, d_dst, n, dim, row_f4 * 4);
  return cudaGetLastError();
}

}  // namespace shn
