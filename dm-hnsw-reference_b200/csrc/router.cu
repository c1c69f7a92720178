// router.cu — query routing between the GPUs of a partitioned index, fused with the exchange.
//
// The reference: with --cache --routing every compute node runs Placement (k-means over the upper-level nodes, k = #CN,
// src/cache/placement.hh:22-106, src/cache/kmeans.hh) and its QueryRouter sends each query to the compute node of the
// nearest centroid unless that node is over its per-batch limit (src/router/query_router.hh:356-368, limits
// :106-151); queries travel as SEND/RECV messages relayed through a memory node (:83-104,195-210) and results stay on
// the compute node that processed them.
//
// Here: one exchange block per GPU (CUDA VMM, mapped by every peer) = [inbox | tags | counts | landing ids | landing
// dists].  A routed step is three kernels per GPU and no copy engine, host loop or library collective on the data path:
//   1. route_pref_kernel    — distances of every query to the `world` centroids -> preference order (3 bits per rank)
//      route_assign_kernel  — the router's rule, exactly as a sequential pass over the batch would apply it (nearest
//                             centroid whose rank is still under `limit`), computed by one CTA in chunks of 1024 queries
//                             with ballots/prefix sums; gives dest[q] and the slot of q in dest's inbox segment
//   2. route_scatter_kernel — writes each query, already in the stored row order of graph.h, straight into the inbox of
//                             its destination over NVLink (plain stores to the peer mapping), plus its home slot as tag
//                             and the per-destination counts
//   3. search_kernel (search.cu, routed I/O) — reads its queries from the inbox segments and writes each result row
//                             straight into the landing buffer of the query's home GPU at the home slot.
// Between 2 and 3 and after 3 the ranks must pass a barrier (the caller's: a stream-ordered NCCL collective in bench.py,
// cross-stream events in the single-process host binary).
#include <algorithm>
#include <cstring>
#include <vector>

#include "handle.h"

using namespace shn;

namespace shn {
namespace {

constexpr int kMaxWorld = 8;
constexpr int kChunk = 1024;

// pref[q]: ranks ordered by distance of query q to their centroid, nearest first (ties: lower rank), 3 bits each.
// Arithmetic: plain fp32 sums (routing decides where a query runs, never what it returns).
__global__ void route_pref_kernel(const float* __restrict__ queries, uint32_t nq, uint32_t dim, const float* __restrict__ cent,
                                  int world, bool ip, uint32_t* __restrict__ pref) {
  extern __shared__ float s_cent[];  // [world][dim]
  for (uint32_t i = threadIdx.x; i < static_cast<uint32_t>(world) * dim; i += blockDim.x) s_cent[i] = cent[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const uint32_t warps = gridDim.x * (blockDim.x >> 5);
  for (uint32_t q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); q < nq; q += warps) {
    float acc[kMaxWorld];
#pragma unroll
    for (int c = 0; c < kMaxWorld; ++c) acc[c] = 0.f;
    const float* row = queries + static_cast<size_t>(q) * dim;
    for (uint32_t j = lane; j < dim; j += 32) {
      const float v = __ldg(row + j);
#pragma unroll
      for (int c = 0; c < kMaxWorld; ++c) {
        if (c < world) {
          const float m = s_cent[c * dim + j];
          if (ip) acc[c] -= v * m;
          else { const float d = v - m; acc[c] += d * d; }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < kMaxWorld; ++c) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xFFFFFFFFu, acc[c], o);
    }
    if (lane == 0) {
      uint32_t packed = 0, used = 0;
      for (int r = 0; r < world; ++r) {
        int best = -1;
#pragma unroll
        for (int c = 0; c < kMaxWorld; ++c) {
          if (c < world && !(used & (1u << c)) && (best < 0 || acc[c] < acc[best])) best = c;
        }
        used |= 1u << best;
        packed |= static_cast<uint32_t>(best) << (3 * r);
      }
      pref[q] = packed;
    }
  }
}

// The router's rule (query_router.hh:356-368) applied in query order: the nearest centroid whose rank has received fewer
// than `limit` queries of this batch; if every rank is at its limit, the farthest one.  One CTA walks the batch in chunks of
// 1024.  A chunk in which no rank crosses its limit is resolved in parallel (every query takes its first choice among the
// ranks that were open when the chunk began — identical to the sequential outcome); the at most `world` chunks in which
// a rank fills up are resolved by one thread in order.  slot[q] = how many earlier queries of the batch go to dest[q].
__global__ void __launch_bounds__(kChunk) route_assign_kernel(const uint32_t* __restrict__ pref, uint32_t nq, int world,
                                                              uint32_t limit, uint8_t* __restrict__ dest,
                                                              uint32_t* __restrict__ slot, uint32_t* __restrict__ hist_out) {
  __shared__ uint32_t s_hist[kMaxWorld], s_cnt[kMaxWorld];
  __shared__ uint32_t s_wcnt[kChunk / 32][kMaxWorld];
  __shared__ uint32_t s_pref[kChunk], s_slot[kChunk];
  __shared__ uint8_t s_dest[kChunk];
  __shared__ int s_slow;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < kMaxWorld) s_hist[tid] = 0;
  __syncthreads();
  for (uint32_t base = 0; base < nq; base += kChunk) {
    const uint32_t q = base + tid;
    const bool valid = q < nq;
    const uint32_t p = valid ? pref[q] : 0u;
    s_pref[tid] = p;
    int c = -1;
    if (valid) {
      for (int r = 0; r < world; ++r) {
        const int cand = (p >> (3 * r)) & 7;
        if (s_hist[cand] < limit) { c = cand; break; }
      }
      if (c < 0) c = (p >> (3 * (world - 1))) & 7;
    }
    uint32_t my_rank = 0;
    for (int d = 0; d < world; ++d) {
      const uint32_t m = __ballot_sync(0xFFFFFFFFu, c == d);
      if (lane == 0) s_wcnt[warp][d] = __popc(m);
      if (c == d) my_rank = __popc(m & ((1u << lane) - 1));
    }
    __syncthreads();
    if (tid < world) {
      uint32_t run = 0;
      for (int w = 0; w < kChunk / 32; ++w) { const uint32_t t = s_wcnt[w][tid]; s_wcnt[w][tid] = run; run += t; }
      s_cnt[tid] = run;
    }
    __syncthreads();
    if (tid == 0) {
      int slow = 0;
      for (int d = 0; d < world; ++d) if (s_hist[d] < limit && s_hist[d] + s_cnt[d] > limit) slow = 1;
      s_slow = slow;
    }
    __syncthreads();
    if (!s_slow) {
      if (valid) { dest[q] = static_cast<uint8_t>(c); slot[q] = s_hist[c] + s_wcnt[warp][c] + my_rank; }
      __syncthreads();
      if (tid < world) s_hist[tid] += s_cnt[tid];
    } else {
      if (tid == 0) {
        const uint32_t cnt = min(static_cast<uint32_t>(kChunk), nq - base);
        for (uint32_t i = 0; i < cnt; ++i) {
          const uint32_t pp = s_pref[i];
          int pick = -1;
          for (int r = 0; r < world; ++r) {
            const int cand = (pp >> (3 * r)) & 7;
            if (s_hist[cand] < limit) { pick = cand; break; }
          }
          if (pick < 0) pick = (pp >> (3 * (world - 1))) & 7;
          s_dest[i] = static_cast<uint8_t>(pick);
          s_slot[i] = s_hist[pick]++;
        }
      }
      __syncthreads();
      if (valid) { dest[q] = s_dest[tid]; slot[q] = s_slot[tid]; }
    }
    __syncthreads();
  }
  if (tid < world) hist_out[tid] = s_hist[tid];
}

// Query q -> inbox segment `rank` of GPU dest[q], slot slot[q], in the stored row order (graph.h row_pos; the search
// kernel then stages it with straight 128-bit copies); tag = q (the home slot the result has to come back to).  A thread per
// stored float4: gathered reads inside the query's own row (L1), coalesced 16-byte stores into local or peer memory.
__global__ void route_scatter_kernel(const float* __restrict__ queries, uint32_t nq, uint32_t dim, uint32_t row_f4,
                                     const uint8_t* __restrict__ dest, const uint32_t* __restrict__ slot, uint32_t rank,
                                     uint32_t cap, char* const* __restrict__ peer_base, uint64_t tags_off) {
  const uint32_t d16 = dim & ~15u, nblk = row_blocks(dim), ntail = dim & 15u;
  const uint64_t total = static_cast<uint64_t>(nq) * row_f4;
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint32_t q = static_cast<uint32_t>(i / row_f4), f = static_cast<uint32_t>(i - static_cast<uint64_t>(q) * row_f4);
    const float* src = queries + static_cast<size_t>(q) * dim;
    float v[4];
#pragma unroll
    for (int e4 = 0; e4 < 4; ++e4) {
      const uint32_t pos = 4 * f + e4;
      float x = 0.f;
      if (pos < 32u * nblk) {  // inverse of row_pos (layout.cu)
        const uint32_t e = row_elem(pos);
        if (e < d16) x = __ldg(src + e);
      } else if (pos - 32u * nblk < ntail) {
        x = __ldg(src + d16 + (pos - 32u * nblk));
      }
      v[e4] = x;
    }
    char* base = peer_base[dest[q]];
    const size_t cell = static_cast<size_t>(rank) * cap + slot[q];
    reinterpret_cast<float4*>(base)[cell * row_f4 + f] = make_float4(v[0], v[1], v[2], v[3]);
    if (f == 0) reinterpret_cast<uint32_t*>(base + tags_off)[cell] = q;
  }
}

__global__ void route_publish_kernel(const uint32_t* __restrict__ hist, int world, uint32_t rank, char* const* __restrict__ peer_base,
                                     uint64_t counts_off) {
  if (threadIdx.x < static_cast<unsigned>(world))
    reinterpret_cast<uint32_t*>(peer_base[threadIdx.x] + counts_off)[rank] = hist[threadIdx.x];
  __threadfence_system();
}

}  // namespace
}  // namespace shn

// ---------------------------------------------------------------------------------------------------------------------

struct shn_router {
  shn_index* ix = nullptr;  // the partitioned (or, world == 1, plain) index this router feeds; not owned
  int gpu = 0;
  uint32_t world = 1, rank = 0, dim = 0, row_f4 = 0, k_max = 0;
  uint64_t max_batch = 0;
  uint32_t cap = 0;          // slots per (source, destination) inbox segment = the router's per-batch limit
  double slack = 0.;
  uint64_t tags_off = 0, counts_off = 0, ids_off = 0, dists_off = 0, block_bytes = 0;
  VmmBlock own, peer[8];
  char* base[8] = {nullptr};  // exchange blocks as this GPU addresses them
  uint32_t attached = 1;
  char** d_base = nullptr;    // device copy of base[]
  float* d_cent = nullptr;
  uint32_t *d_pref = nullptr, *d_slot = nullptr, *d_hist = nullptr;
  uint8_t* d_dest = nullptr;
  uint32_t** d_out_ids = nullptr;  // device tables [world]: landing buffers of every rank
  float** d_out_dists = nullptr;
  cudaStream_t stream = nullptr;
  uint64_t last_nq = 0;
};

static int router_publish_tables(shn_router* r);

extern "C" {

int shn_router_create(shn_router** out, shn_index* ix, const float* centroids, double slack, uint64_t max_batch, uint32_t k_max) {
  if (!out || !ix || !centroids) return fail(SHN_ERR_ARG, "null argument");
  if (max_batch == 0 || max_batch >= (1ull << 31) || k_max == 0 || k_max > 4096) return fail(SHN_ERR_ARG, "need 1 <= max_batch < 2^31 and 1 <= k_max <= 4096");
  if (slack < 0. || slack > 8.) return fail(SHN_ERR_ARG, "slack must be in [0, 8]");
  CU(cudaSetDevice(ix->gpu));
  shn_router* r = new shn_router();
  r->ix = ix; r->gpu = ix->gpu; r->world = ix->world; r->rank = ix->rank; r->dim = ix->dim; r->row_f4 = ix->row_f4;
  r->k_max = k_max; r->max_batch = max_batch; r->slack = slack;
  // query_router.hh:106-151: a compute node takes at most its share of the batch plus the slack; with slack >= 0 the limits
  // of all ranks together exceed the batch, so "every rank is at its limit" cannot occur
  r->cap = static_cast<uint32_t>(std::min<double>(static_cast<double>(max_batch), (1.0 + slack) * static_cast<double>(max_batch) / r->world + 1.0));
  const uint64_t cells = static_cast<uint64_t>(r->world) * r->cap;
  auto up = [](uint64_t v) { return (v + 255) / 256 * 256; };
  r->tags_off = up(cells * r->row_f4 * 16ull);
  r->counts_off = r->tags_off + up(cells * 4);
  r->ids_off = r->counts_off + 256;
  r->dists_off = r->ids_off + up(max_batch * k_max * 4ull);
  r->block_bytes = r->dists_off + up(max_batch * k_max * 4ull);
  auto bail = [&](int code) { shn_router_free(r); return code; };
  const char* why = "";
  if (vmm_alloc(r->own, r->block_bytes, r->gpu, &why) != cudaSuccess) return bail(fail(SHN_ERR_CUDA, "allocating the exchange block (%llu bytes): %s failed", static_cast<unsigned long long>(r->block_bytes), why));
  r->base[r->rank] = static_cast<char*>(r->own.ptr);
#define CUB(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return bail(fail(SHN_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__))); } while (0)
  CUB(cudaMemset(r->own.ptr, 0, r->block_bytes));
  CUB(cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking));
  CUB(cudaMalloc(&r->d_base, 8 * sizeof(char*)));
  CUB(cudaMalloc(&r->d_out_ids, 8 * sizeof(void*)));
  CUB(cudaMalloc(&r->d_out_dists, 8 * sizeof(void*)));
  CUB(cudaMalloc(&r->d_cent, static_cast<size_t>(r->world) * r->dim * sizeof(float)));
  CUB(cudaMemcpy(r->d_cent, centroids, static_cast<size_t>(r->world) * r->dim * sizeof(float), cudaMemcpyHostToDevice));
  CUB(cudaMalloc(&r->d_pref, max_batch * sizeof(uint32_t)));
  CUB(cudaMalloc(&r->d_slot, max_batch * sizeof(uint32_t)));
  CUB(cudaMalloc(&r->d_dest, max_batch));
  CUB(cudaMalloc(&r->d_hist, 8 * sizeof(uint32_t)));
#undef CUB
  {
    const int rc = router_publish_tables(r);
    if (rc != SHN_OK) return bail(rc);
  }
  if (cudaDeviceSynchronize() != cudaSuccess) return bail(fail(SHN_ERR_CUDA, "router setup failed"));  // the memset above is asynchronous
  *out = r;
  return SHN_OK;
}

void shn_router_free(shn_router* r) {
  if (!r) return;
  cudaSetDevice(r->gpu);
  cudaDeviceSynchronize();
  for (auto& b : r->peer) vmm_free(b);
  vmm_free(r->own);
  cudaFree(r->d_base); cudaFree(r->d_out_ids); cudaFree(r->d_out_dists); cudaFree(r->d_cent); cudaFree(r->d_pref);
  cudaFree(r->d_slot); cudaFree(r->d_dest); cudaFree(r->d_hist);
  if (r->stream) cudaStreamDestroy(r->stream);
  delete r;
}

int shn_router_export(const shn_router* r, int* fd, uint64_t* size, uint64_t* raw_ptr) {
  if (!r) return fail(SHN_ERR_ARG, "null router");
  CU(cudaSetDevice(r->gpu));
  if (fd) {
    const char* why = "";
    if (vmm_export_fd(r->own, fd, &why) != cudaSuccess) return fail(SHN_ERR_CUDA, "exporting the exchange block: %s failed", why);
  }
  if (size) *size = r->own.size;
  if (raw_ptr) *raw_ptr = reinterpret_cast<uint64_t>(r->own.ptr);
  return SHN_OK;
}

int shn_router_attach(shn_router* r, int peer, int fd, uint64_t size, uint64_t raw_ptr) {
  if (!r) return fail(SHN_ERR_ARG, "null router");
  if (peer < 0 || peer >= static_cast<int>(r->world) || peer == static_cast<int>(r->rank)) return fail(SHN_ERR_ARG, "bad peer rank %d", peer);
  if (r->base[peer]) return fail(SHN_ERR_STATE, "rank %d is already attached", peer);
  CU(cudaSetDevice(r->gpu));
  if (raw_ptr) {
    r->base[peer] = reinterpret_cast<char*>(raw_ptr);
  } else {
    if (size < r->block_bytes) return fail(SHN_ERR_ARG, "the peer's exchange block is smaller than this rank's (different batch / k / slack?)");
    const char* why = "";
    if (vmm_import_fd(r->peer[peer], fd, size, r->gpu, &why) != cudaSuccess) return fail(SHN_ERR_CUDA, "mapping the exchange block of rank %d: %s failed", peer, why);
    r->base[peer] = static_cast<char*>(r->peer[peer].ptr);
  }
  ++r->attached;
  return router_publish_tables(r);
}

// Phase 1+2: route d_queries[nq][dim] (device, this GPU) and write every query into its destination's inbox.
int shn_router_scatter(shn_router* r, const float* d_queries, uint64_t nq, void* stream) {
  if (!r || (!d_queries && nq)) return fail(SHN_ERR_ARG, "null argument");
  if (nq > r->max_batch) return fail(SHN_ERR_ARG, "batch of %llu queries exceeds the router's max_batch %llu", static_cast<unsigned long long>(nq), static_cast<unsigned long long>(r->max_batch));
  if (r->attached != r->world) return fail(SHN_ERR_STATE, "router: %u of %u exchange blocks attached", r->attached, r->world);
  CU(cudaSetDevice(r->gpu));
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : r->stream;
  r->last_nq = nq;
  const int world = static_cast<int>(r->world);
  // the per-batch limit scales with the batch actually routed (query_router.hh:106-151), never above the segment size
  const uint32_t limit = static_cast<uint32_t>(std::min<double>(r->cap, (1.0 + r->slack) * static_cast<double>(nq) / world + 1.0));
  const size_t smem = static_cast<size_t>(world) * r->dim * sizeof(float);
  if (nq) {
    CU(cudaFuncSetAttribute(route_pref_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    const int grid = static_cast<int>(std::min<uint64_t>((nq + 7) / 8, 148ull * 8));
    route_pref_kernel<<<grid, 256, smem, s>>>(d_queries, static_cast<uint32_t>(nq), r->dim, r->d_cent, world, r->ix->metric == SHN_IP, r->d_pref);
  }
  route_assign_kernel<<<1, kChunk, 0, s>>>(r->d_pref, static_cast<uint32_t>(nq), world, limit, r->d_dest, r->d_slot, r->d_hist);
  if (nq) {
    const uint64_t total = nq * r->row_f4;
    const int grid = static_cast<int>(std::min<uint64_t>((total + 255) / 256, 148ull * 16));
    route_scatter_kernel<<<grid, 256, 0, s>>>(d_queries, static_cast<uint32_t>(nq), r->dim, r->row_f4, r->d_dest, r->d_slot, r->rank,
                                              r->cap, r->d_base, r->tags_off);
  }
  route_publish_kernel<<<1, 32, 0, s>>>(r->d_hist, world, r->rank, r->d_base, r->counts_off);
  CU(cudaGetLastError());
  return SHN_OK;
}


// Phase 3: search every query that arrived in this GPU's inbox (after the barrier that follows every rank's scatter) and
// write each result row into the landing buffer of its home GPU.  Asynchronous unless stats != NULL.
int shn_router_search(shn_router* r, uint32_t k, uint32_t ef, void* stream, shn_stats* stats) {
  if (!r) return fail(SHN_ERR_ARG, "null router");
  if (k > r->k_max) return fail(SHN_ERR_ARG, "k=%u exceeds the router's k_max=%u", k, r->k_max);
  if (r->attached != r->world) return fail(SHN_ERR_STATE, "router: %u of %u exchange blocks attached", r->attached, r->world);
  const uint64_t bound = static_cast<uint64_t>(r->world) * r->cap;  // upper bound on what the inbox can hold
  int rc = check_search_args(r->ix, bound, k, ef);
  if (rc != SHN_OK) return rc;
  if (stats) std::memset(stats, 0, sizeof *stats);
  CU(cudaSetDevice(r->gpu));
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : r->stream;
  char* mine = r->base[r->rank];
  RoutedIo io;
  io.in_rows = reinterpret_cast<const float4*>(mine);
  io.in_tags = reinterpret_cast<const uint32_t*>(mine + r->tags_off);
  io.in_counts = reinterpret_cast<const uint32_t*>(mine + r->counts_off);
  io.cap = r->cap; io.world = r->world;
  io.out_ids = r->d_out_ids; io.out_dists = r->d_out_dists;
  rc = run_search(r->ix, nullptr, bound, k, ef, nullptr, nullptr, nullptr, s, stats != nullptr, &io);
  if (rc != SHN_OK) return rc;
  if (stats) {
    unsigned long long t[kNumTotals];
    CU(cudaMemcpyAsync(t, r->ix->ws.totals, sizeof t, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, r->ix->ev[1], r->ix->ev[2]));
    fill_stats(r->ix, t, t[kProcessed], stats);
    stats->kernel_ms = ms;
    if (t[kFailedQueries]) return fail(SHN_ERR_CAPACITY, "%llu queries overflowed the visited set (ef=%u)", t[kFailedQueries], ef);
  }
  return SHN_OK;
}

// The routing rule alone, for callers that move the queries themselves (parallel.RoutedExchange: one all-to-all each way):
// dest[q] = the rank query q runs on.  The same two kernels as shn_router_scatter — preference order per query, then the
// limit-aware assignment in query order — and one copy of nq bytes back to the host.
int shn_route_queries(const float* centroids, int world, uint32_t dim, shn_metric metric, const float* d_queries, uint64_t nq,
                      double slack, uint8_t* dest, int gpu_id) {
  if (!centroids || !d_queries || !dest) return fail(SHN_ERR_ARG, "null argument");
  if (world < 1 || world > 8 || dim == 0) return fail(SHN_ERR_ARG, "need 1 <= world <= 8");
  if (slack < 0.) return fail(SHN_ERR_ARG, "slack must be >= 0");
  if (nq == 0) return SHN_OK;
  if (nq >= (1ull << 31)) return fail(SHN_ERR_ARG, "too many queries");
  int sms = 0;
  int rc = select_device(gpu_id, &sms);
  if (rc != SHN_OK) return rc;
  // Workspace kept for the life of the process, one per GPU and host thread: cudaMalloc / cudaFree in the per-batch path
  // were measured at ~0.5 s per call once peers' shares are mapped into the address space.
  struct RouteWs { int gpu = -1; DevBuf<float> cent; DevBuf<uint32_t> pref, slot, hist; DevBuf<uint8_t> dest; cudaStream_t stream = nullptr; };
  thread_local RouteWs ws;
  if (ws.gpu != gpu_id) {
    ws.cent.release(); ws.pref.release(); ws.slot.release(); ws.hist.release(); ws.dest.release();
    if (ws.stream) cudaStreamDestroy(ws.stream);
    ws.stream = nullptr;
    ws.gpu = gpu_id;
  }
  if (!ws.stream) CU(cudaStreamCreateWithFlags(&ws.stream, cudaStreamNonBlocking));
  const size_t cent_floats = static_cast<size_t>(world) * dim;
  cudaError_t e = ws.cent.ensure(cent_floats);
  if (e == cudaSuccess) e = ws.pref.ensure(nq);
  if (e == cudaSuccess) e = ws.slot.ensure(nq);
  if (e == cudaSuccess) e = ws.hist.ensure(8);
  if (e == cudaSuccess) e = ws.dest.ensure(nq);
  if (e == cudaSuccess) e = cudaMemcpyAsync(ws.cent.p, centroids, cent_floats * sizeof(float), cudaMemcpyHostToDevice, ws.stream);
  const size_t smem = cent_floats * sizeof(float);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(route_pref_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return fail(SHN_ERR_CUDA, "routing: %s", cudaGetErrorString(e));
  const int grid = static_cast<int>(std::min<uint64_t>((nq + 7) / 8, static_cast<uint64_t>(sms) * 8));
  route_pref_kernel<<<grid, 256, smem, ws.stream>>>(d_queries, static_cast<uint32_t>(nq), dim, ws.cent.p, world, metric == SHN_IP, ws.pref.p);
  // query_router.hh:356-368: the nearest centroid whose compute node is under its limit for this batch
  const uint32_t limit = static_cast<uint32_t>(std::min<double>(static_cast<double>(nq), (1.0 + slack) * static_cast<double>(nq) / world + 1.0));
  route_assign_kernel<<<1, kChunk, 0, ws.stream>>>(ws.pref.p, static_cast<uint32_t>(nq), world, limit, ws.dest.p, ws.slot.p, ws.hist.p);
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(dest, ws.dest.p, nq, cudaMemcpyDeviceToHost, ws.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ws.stream);
  if (e != cudaSuccess) return fail(SHN_ERR_CUDA, "routing: %s", cudaGetErrorString(e));
  return SHN_OK;
}

// This rank's landing buffers (device pointers): row q = the k results of the q-th query this rank handed to
// shn_router_scatter, valid after the barrier that follows every rank's shn_router_search; row stride = the k of that search.
int shn_router_results(const shn_router* r, uint32_t** d_ids, float** d_dists) {
  if (!r) return fail(SHN_ERR_ARG, "null router");
  if (d_ids) *d_ids = reinterpret_cast<uint32_t*>(r->base[r->rank] + r->ids_off);
  if (d_dists) *d_dists = reinterpret_cast<float*>(r->base[r->rank] + r->dists_off);
  return SHN_OK;
}

// Host copies of the last scatter's per-destination counts and of the inbox's per-source counts (synchronises `stream`).
int shn_router_counts(shn_router* r, uint32_t* sent, uint32_t* received, void* stream) {
  if (!r) return fail(SHN_ERR_ARG, "null router");
  CU(cudaSetDevice(r->gpu));
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : r->stream;
  if (sent) CU(cudaMemcpyAsync(sent, r->d_hist, r->world * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  if (received) CU(cudaMemcpyAsync(received, r->base[r->rank] + r->counts_off, r->world * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return SHN_OK;
}

// dest[q] of the last scatter (device pointer, u8 per query) — for tests and the router statistics.
int shn_router_destinations(const shn_router* r, const uint8_t** d_dest) {
  if (!r || !d_dest) return fail(SHN_ERR_ARG, "null argument");
  *d_dest = r->d_dest;
  return SHN_OK;
}

}  // extern "C"

static int router_publish_tables(shn_router* r) {
  uint32_t* ids[8] = {nullptr};
  float* dists[8] = {nullptr};
  for (uint32_t p = 0; p < r->world; ++p) {
    if (!r->base[p]) continue;
    ids[p] = reinterpret_cast<uint32_t*>(r->base[p] + r->ids_off);
    dists[p] = reinterpret_cast<float*>(r->base[p] + r->dists_off);
  }
  CU(cudaMemcpy(r->d_base, r->base, sizeof r->base, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(r->d_out_ids, ids, sizeof ids, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(r->d_out_dists, dists, sizeof dists, cudaMemcpyHostToDevice));
  return SHN_OK;
}
