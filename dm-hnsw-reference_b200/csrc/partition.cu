// partition.cu — split a full index (one GPU) into what one GPU of `world` keeps (graph.h DeviceGraph, flat numbering):
// rows are renumbered [hot set | share 0 | share 1 | ...] with granule-aligned pieces; this GPU materialises the hot set
// and its own share.
#include "engine.h"

namespace shn {
namespace {

__global__ void invert_perm_kernel(const uint32_t* __restrict__ new_of_old, uint32_t* __restrict__ old_of_new, uint32_t n) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) old_of_new[new_of_old[i]] = i;
}

// dst[i] = src[old_of_new[first + i]], rows of row_f4 float4; a warp per destination row.  Flat ids nobody was given (the pads
// between the pieces) stay zero.
__global__ void gather_rows_kernel(const float4* __restrict__ src, const uint32_t* __restrict__ old_of_new, uint32_t first,
                                   uint32_t count, uint32_t row_f4, float4* __restrict__ dst) {
  const uint32_t warps = gridDim.x * (blockDim.x >> 5);
  const int lane = threadIdx.x & 31;
  for (uint32_t i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < count; i += warps) {
    const uint32_t old = old_of_new[first + i];
    if (old == kInvalid) continue;
    for (uint32_t f = lane; f < row_f4; f += 32) dst[static_cast<size_t>(i) * row_f4 + f] = src[static_cast<size_t>(old) * row_f4 + f];
  }
}

// dst[i][s] = remap(src[old_of_new[first + i]][s]); 0xFFFFFFFF stays
__global__ void gather_lists_kernel(const uint32_t* __restrict__ src, uint32_t width, const uint32_t* __restrict__ old_of_new,
                                    const uint32_t* __restrict__ new_of_old, uint32_t first, uint32_t count,
                                    uint32_t* __restrict__ dst) {
  const uint64_t total = static_cast<uint64_t>(count) * width;
  for (uint64_t t = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; t < total;
       t += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint32_t i = static_cast<uint32_t>(t / width), s = static_cast<uint32_t>(t % width);
    const uint32_t old = old_of_new[first + i];
    const uint32_t nb = old == kInvalid ? kInvalid : src[static_cast<size_t>(old) * width + s];
    dst[t] = nb == kInvalid ? kInvalid : new_of_old[nb];
  }
}

__global__ void remap_kernel(const uint32_t* __restrict__ src, const uint32_t* __restrict__ new_of_old, uint64_t total,
                             uint32_t* __restrict__ dst) {
  for (uint64_t t = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; t < total;
       t += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint32_t nb = src[t];
    dst[t] = nb == kInvalid ? kInvalid : new_of_old[nb];
  }
}

__global__ void gather_u32_kernel(const uint32_t* __restrict__ src, const uint32_t* __restrict__ old_of_new, uint32_t count,
                                  uint32_t* __restrict__ dst) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
    const uint32_t old = old_of_new[i];
    dst[i] = old == kInvalid ? kInvalid : src[old];
  }
}

// diagnostic: random whole-row reads from one share (local or peer-mapped)
__global__ void probe_gather_kernel(const float4* __restrict__ src, uint32_t nrows, uint32_t row_f4, uint32_t iters, float* out) {
  const int lane = threadIdx.x & 31, t = lane & 7, grp = lane >> 3;
  uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) / 8 * 2654435761u + 12345u;
  float acc = 0.f;
  for (uint32_t it = 0; it < iters; ++it) {
    s = s * 1664525u + 1013904223u;
    const uint32_t row = __shfl_sync(0xffffffffu, s, grp * 8) % nrows;
    const float4* p = src + static_cast<size_t>(row) * row_f4 + t;
    for (uint32_t b = 0; b < row_f4 / 8; ++b) { const float4 v = __ldg(p + 8 * b); acc += v.x + v.w; }
  }
  if (acc == 123.456f) out[0] = acc;
}

// halo slot i <- row rows[i] of the partitioned graph, wherever it lives (a peer's share over NVLink, normally); a warp per row
__global__ void halo_gather_kernel(const DeviceGraph g, const uint32_t* __restrict__ rows, uint32_t count, float4* __restrict__ dst_vec,
                                   uint32_t* __restrict__ dst_l0) {
  const uint32_t warps = gridDim.x * (blockDim.x >> 5);
  const int lane = threadIdx.x & 31;
  for (uint32_t i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < count; i += warps) {
    const uint32_t row = rows[i];
    const float4* v = vec_row<false>(g, row);   // flat numbering: the row is where its id says, on whichever GPU
    const uint32_t* l = l0_row<false>(g, row);
    for (uint32_t f = lane; f < g.row_f4; f += 32) dst_vec[static_cast<size_t>(i) * g.row_f4 + f] = v[f];
    for (uint32_t f = lane; f < g.m0; f += 32) dst_l0[static_cast<size_t>(i) * g.m0 + f] = l[f];
  }
}

inline int grid_for(uint64_t work, int threads) {
  return static_cast<int>(std::max<uint64_t>(1, std::min<uint64_t>((work + threads - 1) / threads, 148ull * 16)));
}

}  // namespace

cudaError_t partition_arrays(const PartitionJob& j, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(j.old_of_new, 0xFF, static_cast<size_t>(j.n_flat) * sizeof(uint32_t), s);
  if (e != cudaSuccess) return e;
  invert_perm_kernel<<<grid_for(j.n, 256), 256, 0, s>>>(j.new_of_old, j.old_of_new, j.n);
  // replicated hot set: flat ids [0, hot)
  if (j.hot) {
    gather_rows_kernel<<<grid_for(static_cast<uint64_t>(j.hot) * 32, 256), 256, 0, s>>>(j.src_vec, j.old_of_new, 0, j.hot, j.row_f4, j.hot_vec);
    gather_lists_kernel<<<grid_for(static_cast<uint64_t>(j.hot) * j.m0, 256), 256, 0, s>>>(j.src_l0, j.m0, j.old_of_new, j.new_of_old, 0, j.hot, j.hot_l0);
    gather_u32_kernel<<<grid_for(j.hot, 256), 256, 0, s>>>(j.src_up_base, j.old_of_new, j.hot, j.hot_up_base);
  }
  if (j.n_up) remap_kernel<<<grid_for(j.n_up * j.m, 256), 256, 0, s>>>(j.src_up, j.new_of_old, j.n_up * j.m, j.up);
  gather_u32_kernel<<<grid_for(j.n_flat, 256), 256, 0, s>>>(j.src_ext_id, j.old_of_new, j.n_flat, j.ext_id);
  // this GPU's share: flat ids [own_first, own_first + own)
  if (j.own) {
    gather_rows_kernel<<<grid_for(static_cast<uint64_t>(j.own) * 32, 256), 256, 0, s>>>(j.src_vec, j.old_of_new, j.own_first, j.own, j.row_f4, j.own_vec);
    gather_lists_kernel<<<grid_for(static_cast<uint64_t>(j.own) * j.m0, 256), 256, 0, s>>>(j.src_l0, j.m0, j.old_of_new, j.new_of_old, j.own_first, j.own, j.own_l0);
  }
  return cudaGetLastError();
}


// dst[i] = src[rows[i]] (whole rows of row_f4 float4): one launch instead of one copy per row
cudaError_t gather_rows(const float4* src, const uint32_t* d_rows, uint32_t count, uint32_t row_f4, float4* dst, cudaStream_t s) {
  if (count == 0) return cudaSuccess;
  gather_rows_kernel<<<grid_for(static_cast<uint64_t>(count) * 32, 256), 256, 0, s>>>(src, d_rows, 0, count, row_f4, dst);
  return cudaGetLastError();
}

cudaError_t halo_gather(const DeviceGraph& g, const uint32_t* d_rows, uint32_t count, float4* dst_vec, uint32_t* dst_l0, cudaStream_t s) {
  if (count == 0) return cudaSuccess;
  halo_gather_kernel<<<grid_for(static_cast<uint64_t>(count) * 32, 256), 256, 0, s>>>(g, d_rows, count, dst_vec, dst_l0);
  return cudaGetLastError();
}

cudaError_t probe_gather(const float4* src, uint32_t nrows, uint32_t row_f4, double* gbs, cudaStream_t s) {
  float* out = nullptr;
  cudaError_t e = cudaMalloc(&out, 4);
  if (e != cudaSuccess) return e;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  const int blocks = 148 * 8;
  const uint32_t iters = 400;
  probe_gather_kernel<<<blocks, 128, 0, s>>>(src, nrows, row_f4, 20, out);
  cudaEventRecord(a, s);
  probe_gather_kernel<<<blocks, 128, 0, s>>>(src, nrows, row_f4, iters, out);
  cudaEventRecord(b, s);
  e = cudaStreamSynchronize(s);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, a, b);
  *gbs = static_cast<double>(blocks) * 16 * iters * row_f4 * 16.0 / ms / 1e6;
  cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(out);
  return e;
}

}  // namespace shn
