// bruteforce.cu — exact top-k by exhaustive search: the ground truth the reference only ever READS
// (queries/groundtruth-<suffix>.bin, src/compute_node.cc:317,588) and that has to be produced for the synthetic
// 10M / 100M sets.  fp32, distances accumulated with one fma per element in natural element order
// (acc = fma(q-v, q-v, acc) resp. acc = fma(q, v, acc), distance 1 - acc), ties broken by the lower row id.
//
// Kernel 1: a CTA owns 64 queries x one slice of the base; it streams the slice in tiles of 128 rows, forms the
// 64 x 128 distance tile with a register-tiled SIMT loop (4 queries x 8 rows per thread, operands staged through
// shared memory in chunks of 16 dimensions), and folds the tile into per-query sorted top-k lists kept in shared memory.
// Kernel 2: a warp per query merges the per-slice lists.
// (Round-1 implementation on the fp32 pipes; the tcgen05 candidate-generation + fp32 re-rank variant is the next step.)
#include <cfloat>

#include "engine.h"

namespace shn {
namespace {

constexpr int QT = 64;    // queries per CTA
constexpr int BT = 128;   // base rows per tile
constexpr int KC = 16;    // dimensions per staged chunk
constexpr int THREADS = 256;
constexpr uint32_t kFullMask = 0xFFFFFFFFu;

struct BfParams {
  const float* base;
  const float* queries;
  uint64_t n;
  uint32_t nq, dim, k, slices;
  uint64_t rows_per_slice;
  float* part_d;      // [slices][nq][k]
  uint32_t* part_i;
};

// Insert (d, id) into the ascending list (ld, li) of length k; ids arrive in increasing order inside one CTA, so
// "after equal distances" keeps the lower id first.  Warp-cooperative.
__device__ __forceinline__ void list_insert(float* ld, uint32_t* li, uint32_t k, float d, uint32_t id, int lane) {
  uint32_t pos = 0;
  for (uint32_t b = 0; b < k; b += 32) {
    const uint32_t j = b + lane;
    pos += __popc(__ballot_sync(kFullMask, j < k && ld[j] <= d));
  }
  if (pos >= k) return;
  int hi = static_cast<int>(k) - 1;
  while (hi > static_cast<int>(pos)) {
    const int lo = max(static_cast<int>(pos), hi - 32);
    const int j = lo + lane;
    const bool act = j < hi;
    float td = 0.f;
    uint32_t ti = 0;
    if (act) { td = ld[j]; ti = li[j]; }
    __syncwarp();
    if (act) { ld[j + 1] = td; li[j + 1] = ti; }
    __syncwarp();
    hi = lo;
  }
  if (lane == 0) { ld[pos] = d; li[pos] = id; }
  __syncwarp();
}

template <bool IP>
__global__ void __launch_bounds__(THREADS) bf_tile_kernel(const BfParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* Qs = reinterpret_cast<float*>(smem_raw);            // [KC][QT+1]
  float* Bs = Qs + KC * (QT + 1);                              // [KC][BT+1]
  float* Ds = Bs + KC * (BT + 1);                              // [QT][BT+1]
  float* Ld = Ds + QT * (BT + 1);                              // [QT][k]
  uint32_t* Li = reinterpret_cast<uint32_t*>(Ld + QT * p.k);   // [QT][k]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ty = tid >> 4, tx = tid & 15;  // thread computes queries ty*4..+3, rows tx + 16*j (j < 8)
  const uint32_t q0 = blockIdx.x * QT;
  const uint32_t slice = blockIdx.y;
  const uint64_t r_begin = slice * p.rows_per_slice;
  const uint64_t r_end = min(p.n, r_begin + p.rows_per_slice);

  for (uint32_t i = tid; i < QT * p.k; i += THREADS) { Ld[i] = FLT_MAX; Li[i] = kInvalid; }
  __syncthreads();

  for (uint64_t r0 = r_begin; r0 < r_end; r0 += BT) {
    float acc[4][8];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;

    for (uint32_t k0 = 0; k0 < p.dim; k0 += KC) {
      // stage the chunk, transposed (element-major) so that the inner loop reads are conflict-free / broadcast
      for (uint32_t i = tid; i < QT * KC; i += THREADS) {
        const uint32_t q = i / KC, e = i % KC;
        const bool ok = q0 + q < p.nq && k0 + e < p.dim;
        Qs[e * (QT + 1) + q] = ok ? __ldg(p.queries + static_cast<size_t>(q0 + q) * p.dim + k0 + e) : 0.f;
      }
      for (uint32_t i = tid; i < BT * KC; i += THREADS) {
        const uint32_t r = i / KC, e = i % KC;
        const bool ok = r0 + r < r_end && k0 + e < p.dim;
        Bs[e * (BT + 1) + r] = ok ? __ldg(p.base + (r0 + r) * p.dim + k0 + e) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int e = 0; e < KC; ++e) {
        float qv[4], bv[8];
#pragma unroll
        for (int a = 0; a < 4; ++a) qv[a] = Qs[e * (QT + 1) + ty * 4 + a];
#pragma unroll
        for (int b = 0; b < 8; ++b) bv[b] = Bs[e * (BT + 1) + tx + 16 * b];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 8; ++b) {
            if (IP) acc[a][b] = __fmaf_rn(qv[a], bv[b], acc[a][b]);
            else { const float d = __fsub_rn(qv[a], bv[b]); acc[a][b] = __fmaf_rn(d, d, acc[a][b]); }
          }
      }
      __syncthreads();
    }
    // zero padding beyond dim contributed fma(0,0,acc) = acc: harmless
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 8; ++b) Ds[(ty * 4 + a) * (BT + 1) + tx + 16 * b] = IP ? __fsub_rn(1.0f, acc[a][b]) : acc[a][b];
    __syncthreads();

    // fold the tile into the per-query lists: warp w owns queries w*8 .. w*8+7
    for (int qq = 0; qq < QT / (THREADS / 32); ++qq) {
      const int q = warp * (QT / (THREADS / 32)) + qq;
      if (q0 + q >= p.nq) break;
      float* ld = Ld + q * p.k;
      uint32_t* li = Li + q * p.k;
      for (int c = 0; c < BT; c += 32) {
        const int r = c + lane;
        const float d = Ds[q * (BT + 1) + r];
        const bool cand = r0 + r < r_end && d < ld[p.k - 1];
        uint32_t mask = __ballot_sync(kFullMask, cand);
        while (mask) {
          const int src = __ffs(mask) - 1;
          mask &= mask - 1;
          const float dd = __shfl_sync(kFullMask, d, src);
          if (dd < ld[p.k - 1]) list_insert(ld, li, p.k, dd, static_cast<uint32_t>(r0 + c + src), lane);
        }
      }
    }
    __syncthreads();
  }

  for (uint32_t i = tid; i < QT * p.k; i += THREADS) {
    const uint32_t q = i / p.k, j = i % p.k;
    if (q0 + q < p.nq) {
      const size_t o = (static_cast<size_t>(slice) * p.nq + q0 + q) * p.k + j;
      p.part_d[o] = Ld[i];
      p.part_i[o] = Li[i];
    }
  }
}

// Merge the per-slice lists of one query (each ascending, ties by id): a warp repeatedly takes the smallest head.
__global__ void bf_merge_kernel(const float* part_d, const uint32_t* part_i, uint32_t nq, uint32_t k, uint32_t slices,
                                uint32_t* out_i, float* out_d) {
  const uint32_t q = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= nq) return;
  // lane s (+32, +64 ...) tracks the head of slice s
  uint32_t head[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) head[t] = 0;
  for (uint32_t j = 0; j < k; ++j) {
    float bd = FLT_MAX;
    uint32_t bi = kInvalid, bs = kInvalid;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const uint32_t s = lane + 32 * t;
      if (s < slices && head[t] < k) {
        const size_t o = (static_cast<size_t>(s) * nq + q) * k + head[t];
        const float d = part_d[o];
        const uint32_t id = part_i[o];
        if (id != kInvalid && (d < bd || (d == bd && id < bi))) { bd = d; bi = id; bs = s; }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float od = __shfl_xor_sync(kFullMask, bd, o);
      const uint32_t oi = __shfl_xor_sync(kFullMask, bi, o);
      const uint32_t os = __shfl_xor_sync(kFullMask, bs, o);
      if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; bs = os; }
    }
    if (lane == 0) {
      out_i[static_cast<size_t>(q) * k + j] = bi;
      if (out_d) out_d[static_cast<size_t>(q) * k + j] = bi == kInvalid ? __int_as_float(0x7f800000) : bd;
    }
    if (bs != kInvalid && (bs & 31u) == static_cast<uint32_t>(lane)) ++head[bs >> 5];
  }
}

}  // namespace

cudaError_t bruteforce_launch(const float* d_base, uint64_t n, const float* d_queries, uint32_t nq, uint32_t dim, bool ip,
                              uint32_t k, uint32_t* d_ids, float* d_dists, int num_sms, cudaStream_t stream) {
  if (nq == 0) return cudaSuccess;
  if (k == 0 || k > 256) return cudaErrorInvalidValue;
  const uint32_t qtiles = (nq + QT - 1) / QT;
  uint32_t slices = std::max<uint32_t>(1, (2u * num_sms + qtiles - 1) / qtiles);
  slices = std::min<uint32_t>(slices, 256);
  slices = static_cast<uint32_t>(std::min<uint64_t>(slices, (n + BT - 1) / BT));
  if (slices == 0) slices = 1;
  uint64_t rows_per_slice = (n + slices - 1) / slices;
  rows_per_slice = (rows_per_slice + BT - 1) / BT * BT;
  slices = static_cast<uint32_t>((n + rows_per_slice - 1) / rows_per_slice);
  if (slices == 0) slices = 1;

  float* part_d = nullptr;
  uint32_t* part_i = nullptr;
  const size_t part = static_cast<size_t>(slices) * nq * k;
  cudaError_t e = cudaMalloc(&part_d, part * sizeof(float));
  if (e != cudaSuccess) return e;
  e = cudaMalloc(&part_i, part * sizeof(uint32_t));
  if (e != cudaSuccess) { cudaFree(part_d); return e; }

  BfParams p;
  p.base = d_base; p.queries = d_queries; p.n = n; p.nq = nq; p.dim = dim; p.k = k; p.slices = slices;
  p.rows_per_slice = rows_per_slice; p.part_d = part_d; p.part_i = part_i;
  const size_t smem = sizeof(float) * (KC * (QT + 1) + KC * (BT + 1) + QT * (BT + 1) + 2ull * QT * k);
  const dim3 grid(qtiles, slices);
  if (ip) {
    e = cudaFuncSetAttribute(bf_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e == cudaSuccess) bf_tile_kernel<true><<<grid, THREADS, smem, stream>>>(p);
  } else {
    e = cudaFuncSetAttribute(bf_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e == cudaSuccess) bf_tile_kernel<false><<<grid, THREADS, smem, stream>>>(p);
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e == cudaSuccess) {
    bf_merge_kernel<<<(nq + 3) / 4, 128, 0, stream>>>(part_d, part_i, nq, k, slices, d_ids, d_dists);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  cudaFree(part_d);
  cudaFree(part_i);
  return e;
}

}  // namespace shn
