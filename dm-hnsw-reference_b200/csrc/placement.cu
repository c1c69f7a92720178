// placement.cu — which GPU owns which node, and which GPU a query should run on.
//
// The reference places nodes on memory nodes uniformly at random (src/compute_thread.hh:57) and gets locality only from
// the compute-node caches: Placement (src/cache/placement.hh:22-106) fetches the upper-level nodes, runs a balanced
// k-means with k = #compute nodes (src/cache/kmeans.hh:24-377, k-means++ seeded with 1234 :163-169) and the query router
// sends a query to the compute node of its nearest centroid unless that node is over its per-batch limit
// (src/router/query_router.hh:356-368).  With the index itself living in the GPUs' HBM the same centroids can also
// decide where a node is STORED: a node goes to the GPU of its nearest centroid (balanced), a query to the GPU of its
// nearest centroid — most hops then stay in local HBM.
#include <cmath>
#include <cstring>
#include <random>
#include <vector>

#include "engine.h"

namespace shn {
namespace {

constexpr int kMaxParts = 8;

// out[row][c] = || v_row - centroid_c ||^2 or -<v_row, centroid_c>, on the STORED row order (a permutation of the
// natural order with zero padding, so sums are unaffected); a warp per row
__global__ void centroid_dist_kernel(const float4* __restrict__ vec, uint32_t n, uint32_t row_f4, const float* __restrict__ cent,
                                     int k, bool ip, float* __restrict__ out) {
  extern __shared__ float s_cent[];
  for (uint32_t i = threadIdx.x; i < static_cast<uint32_t>(k) * row_f4 * 4; i += blockDim.x) s_cent[i] = cent[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const uint32_t warps = gridDim.x * (blockDim.x >> 5);
  for (uint32_t row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < n; row += warps) {
    float acc[kMaxParts];
#pragma unroll
    for (int c = 0; c < kMaxParts; ++c) acc[c] = 0.f;
    for (uint32_t f = lane; f < row_f4; f += 32) {
      const float4 v = __ldg(vec + static_cast<size_t>(row) * row_f4 + f);
#pragma unroll
      for (int c = 0; c < kMaxParts; ++c) {
        if (c < k) {
          const float4 m = reinterpret_cast<const float4*>(s_cent)[static_cast<size_t>(c) * row_f4 + f];
          if (ip) acc[c] -= v.x * m.x + v.y * m.y + v.z * m.z + v.w * m.w;
          else { const float a = v.x - m.x, b = v.y - m.y, d = v.z - m.z, e = v.w - m.w; acc[c] += a * a + b * b + d * d + e * e; }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < kMaxParts; ++c) {
      if (c < k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xFFFFFFFFu, acc[c], o);
        if (lane == 0) out[static_cast<size_t>(row) * k + c] = acc[c];
      }
    }
  }
}

// owner[row] = argmin_c (dist[row][c] + bias[c]); counts[c] += 1
__global__ void biased_assign_kernel(const float* __restrict__ dist, uint32_t n, int k, const float* __restrict__ bias,
                                     uint8_t* __restrict__ owner, unsigned int* __restrict__ counts) {
  __shared__ unsigned int s_cnt[kMaxParts];
  if (threadIdx.x < kMaxParts) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  for (uint32_t row = blockIdx.x * blockDim.x + threadIdx.x; row < n; row += gridDim.x * blockDim.x) {
    int best = 0;
    float bd = dist[static_cast<size_t>(row) * k] + bias[0];
    for (int c = 1; c < k; ++c) {
      const float d = dist[static_cast<size_t>(row) * k + c] + bias[c];
      if (d < bd) { bd = d; best = c; }
    }
    owner[row] = static_cast<uint8_t>(best);
    atomicAdd(&s_cnt[best], 1u);
  }
  __syncthreads();
  if (threadIdx.x < kMaxParts && s_cnt[threadIdx.x]) atomicAdd(&counts[threadIdx.x], s_cnt[threadIdx.x]);
}

float host_dist(const float* a, const float* b, uint32_t d, bool ip) {
  double s = 0;
  if (ip) { for (uint32_t i = 0; i < d; ++i) s -= static_cast<double>(a[i]) * b[i]; }
  else { for (uint32_t i = 0; i < d; ++i) { const double t = static_cast<double>(a[i]) - b[i]; s += t * t; } }
  return static_cast<float>(s);
}

}  // namespace

// k-means on the host over a sample (rows of `d` floats): k-means++ (mt19937(seed), kmeans.hh:163-169) + Lloyd.
void kmeans_host(const std::vector<float>& sample, uint32_t count, uint32_t d, int k, uint32_t seed, bool ip,
                 std::vector<float>& centroids) {
  centroids.assign(static_cast<size_t>(k) * d, 0.f);
  std::mt19937 gen(seed);
  std::vector<double> best(count, 1e300);
  uint32_t first = std::uniform_int_distribution<uint32_t>(0, count - 1)(gen);
  std::memcpy(centroids.data(), sample.data() + static_cast<size_t>(first) * d, d * sizeof(float));
  for (int c = 1; c < k; ++c) {
    double total = 0;
    for (uint32_t i = 0; i < count; ++i) {
      double dd = host_dist(sample.data() + static_cast<size_t>(i) * d, centroids.data() + static_cast<size_t>(c - 1) * d, d, false);
      if (dd < best[i]) best[i] = dd;
      total += best[i];
    }
    double pick = std::uniform_real_distribution<double>(0., total)(gen);
    uint32_t chosen = count - 1;
    for (uint32_t i = 0; i < count; ++i) { pick -= best[i]; if (pick <= 0) { chosen = i; break; } }
    std::memcpy(centroids.data() + static_cast<size_t>(c) * d, sample.data() + static_cast<size_t>(chosen) * d, d * sizeof(float));
  }
  std::vector<int> assign(count, 0);
  for (int iter = 0; iter < 25; ++iter) {
    bool changed = false;
    for (uint32_t i = 0; i < count; ++i) {
      int b = 0;
      float bd = host_dist(sample.data() + static_cast<size_t>(i) * d, centroids.data(), d, ip);
      for (int c = 1; c < k; ++c) {
        const float dd = host_dist(sample.data() + static_cast<size_t>(i) * d, centroids.data() + static_cast<size_t>(c) * d, d, ip);
        if (dd < bd) { bd = dd; b = c; }
      }
      if (assign[i] != b) { assign[i] = b; changed = true; }
    }
    std::vector<double> sum(static_cast<size_t>(k) * d, 0.);
    std::vector<uint32_t> cnt(k, 0);
    for (uint32_t i = 0; i < count; ++i) {
      ++cnt[assign[i]];
      for (uint32_t j = 0; j < d; ++j) sum[static_cast<size_t>(assign[i]) * d + j] += sample[static_cast<size_t>(i) * d + j];
    }
    for (int c = 0; c < k; ++c) {
      if (!cnt[c]) continue;  // an empty cluster keeps its centroid
      for (uint32_t j = 0; j < d; ++j) centroids[static_cast<size_t>(c) * d + j] = static_cast<float>(sum[static_cast<size_t>(c) * d + j] / cnt[c]);
    }
    if (!changed && iter > 0) break;
  }
}

// Balanced nearest-centroid assignment of all n rows: argmin(dist + bias) with the biases of over-full parts raised
// until every part holds at most (1 + slack) * n / k rows.
cudaError_t balanced_assign(const float4* d_vec, uint32_t n, uint32_t row_f4, const float* d_cent_stored, int k, bool ip,
                            double slack, uint8_t* d_owner, std::vector<uint32_t>& sizes, cudaStream_t s) {
  float* d_dist = nullptr;
  float* d_bias = nullptr;
  unsigned int* d_cnt = nullptr;
  cudaError_t e = cudaMalloc(&d_dist, static_cast<size_t>(n) * k * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&d_bias, kMaxParts * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&d_cnt, kMaxParts * sizeof(unsigned int));
  if (e != cudaSuccess) { cudaFree(d_dist); cudaFree(d_bias); cudaFree(d_cnt); return e; }
  const size_t smem = static_cast<size_t>(k) * row_f4 * 16;
  cudaFuncSetAttribute(centroid_dist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  centroid_dist_kernel<<<148 * 8, 256, smem, s>>>(d_vec, n, row_f4, d_cent_stored, k, ip, d_dist);
  std::vector<float> bias(kMaxParts, 0.f);
  sizes.assign(k, 0);
  // scale of the distances, for the bias step: mean nearest distance of a sample
  std::vector<float> probe(static_cast<size_t>(std::min<uint32_t>(n, 4096)) * k);
  e = cudaMemcpyAsync(probe.data(), d_dist, probe.size() * sizeof(float), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  double scale = 0;
  for (size_t i = 0; i < probe.size() / k; ++i) {
    float lo = probe[i * k], hi = probe[i * k];
    for (int c = 1; c < k; ++c) { lo = std::min(lo, probe[i * k + c]); hi = std::max(hi, probe[i * k + c]); }
    scale += hi - lo;
  }
  scale = std::max(1e-12, scale / std::max<size_t>(1, probe.size() / k));
  const double target = static_cast<double>(n) / k;
  for (int iter = 0; iter < 60 && e == cudaSuccess; ++iter) {
    e = cudaMemcpyAsync(d_bias, bias.data(), kMaxParts * sizeof(float), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_cnt, 0, kMaxParts * sizeof(unsigned int), s);
    if (e != cudaSuccess) break;
    biased_assign_kernel<<<148 * 8, 256, 0, s>>>(d_dist, n, k, d_bias, d_owner, d_cnt);
    unsigned int cnt[kMaxParts];
    e = cudaMemcpyAsync(cnt, d_cnt, sizeof cnt, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) break;
    bool ok = true;
    for (int c = 0; c < k; ++c) {
      sizes[c] = cnt[c];
      if (cnt[c] > (1.0 + slack) * target + 1) ok = false;
    }
    if (ok) break;
    for (int c = 0; c < k; ++c) bias[c] += static_cast<float>(0.15 * scale * (cnt[c] - target) / target);
  }
  cudaFree(d_dist); cudaFree(d_bias); cudaFree(d_cnt);
  return e;
}

}  // namespace shn
