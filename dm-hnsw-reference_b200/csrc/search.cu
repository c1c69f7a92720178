// search.cu — batched k-NN search: one warp per query, persistent warps pulling queries from a global cursor.
//
// Restates HNSW::knn (src/hnsw/hnsw.hh:253-307): entry point -> greedy descent through the upper layers
// (search_for_one<without_lock>, :332-393) -> ef-bounded best-first search on layer 0
// (search_level<without_lock>, :407-476) -> trim to k (:296-298).  What the reference does with one RDMA READ per
// neighbour and a coroutine switch (src/rdma/rdma_reads.hh:9,40) is here a wave of 128-bit loads from HBM.
// Semantics kept exactly: neighbours are admitted in stored list order against the running
// farthest distance, strict '<', visited marked before the distance is computed; distances are summed in the
// reference's own order (search.cuh), so ids and distances are bit-identical except on exact distance ties.
#include <cstdio>
#include <cstdlib>

#include "engine.h"
#include "search.cuh"

namespace shn {
namespace {

// CTA shape: 4 warps, 5 CTAs per SM -> 20 query warps per SM at 96 registers.  Measured on B200 (SIFT10M-shaped, 1 M queries,
// ef 16/64/128/256, M q/s): 4x5 18.3/6.65/2.92/1.35, 4x6 (80 registers, spills) 18.2/5.76/2.46/1.15, 8x3 18.3/5.74/2.41/1.10,
// 8x4 (64 registers) 12.7/3.55/1.95/0.85 — the kernel is bound by memory latency x bytes in flight per warp, and squeezing
// the row registers costs more than the extra warps bring.
#ifndef SHN_WARPS_PER_BLOCK
#define SHN_WARPS_PER_BLOCK 4
#endif
#ifndef SHN_MIN_BLOCKS
#define SHN_MIN_BLOCKS 5
#endif
constexpr int kWarpsPerBlock = SHN_WARPS_PER_BLOCK;

struct SearchParams {
  DeviceGraph g;
  const float* queries;
  uint32_t nq, k, ef;
  uint32_t* out_ids;
  float* out_dists;
  uint32_t* per_query;
  uint32_t* counter;
  unsigned long long* totals;
  uint32_t* ovf;
  uint32_t vis_cap, vis_limit, ovf_cap, ovf_limit;
  uint32_t vis_compact;                 // 16-bit keys in the shared visited table (search.cuh visited_compact)
  uint32_t q_floats, ef_cap, list_cap;  // shared-memory strides
  RoutedIo io;
};

// per warp: launch totals | read ids of one list (partitioned handles only, graph.h read_id) | query | row/distance staging (one list) |
// queue distances | queue ids | visited table
constexpr uint32_t kTotalsBytes = (kNumTotals * 8 + 15) / 16 * 16;
__host__ __device__ inline size_t warp_smem_bytes(uint32_t q_floats, uint32_t ef_cap, uint32_t list_cap, uint32_t vis_cap, bool part) {
  return kTotalsBytes + (part ? 4ull * list_cap : 0ull) + 4ull * (q_floats + 2 * list_cap + 2 * ef_cap + vis_cap);
}

// WIDE: the large-ef configuration — 16 rows in flight per warp (4 passes, 128 registers) on 4 CTAs per SM instead of 8 rows on
// 5.  Measured on B200 (M q/s, narrow / wide): 10 M x 128 at ef 64 / 72 / 100 / 128 / 200: 7.07/6.76, 5.49/5.71, 3.89/4.03, 3.02/3.13,
// 1.82/1.89; 20 M x 96 at ef 64 / 100 / 128: 7.60/7.11, 4.44/4.33, 3.42/3.33.  The wider waves pay when a row spans four lines and
// the expansions are long; SHN_WIDE_FROM_EF overrides the crossover (tuning).
bool use_wide(uint32_t dim, uint32_t ef) {
  static const int env = [] { const char* e = getenv("SHN_WIDE_FROM_EF"); return e ? atoi(e) : -1; }();
  if (env >= 0) return ef >= static_cast<uint32_t>(env);
  return dim >= 128 && ef >= 72;
}
template <bool IP, int NCHUNK, bool PART, bool WIDE>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, WIDE ? SHN_MIN_BLOCKS - 1 : SHN_MIN_BLOCKS) search_kernel(const SearchParams p) {
  constexpr int PASSES = WIDE ? 2 * SHN_PASSES : SHN_PASSES;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const DeviceGraph& g = p.g;

  const uint32_t q_floats = NCHUNK > 0 ? row_stride_f4(NCHUNK) * 4u : p.q_floats;
  unsigned char* base = smem_raw + warp * warp_smem_bytes(q_floats, p.ef_cap, p.list_cap, p.vis_cap, PART);
  // the launch totals live in shared memory (11 x 64 bit per warp would otherwise sit in registers for the whole kernel)
  unsigned long long* s_tot = reinterpret_cast<unsigned long long*>(base);
  uint32_t* s_read = reinterpret_cast<uint32_t*>(base + kTotalsBytes);
  float* s_q = reinterpret_cast<float*>(base + kTotalsBytes + (PART ? 4u * p.list_cap : 0u));
  uint32_t* s_rows = reinterpret_cast<uint32_t*>(s_q + q_floats);
  float* s_dist = reinterpret_cast<float*>(s_rows + p.list_cap);
  float* qd = s_dist + p.list_cap;
  uint32_t* qi = reinterpret_cast<uint32_t*>(qd + p.ef_cap);
  VisitedSet vis;
  vis.tab = qi + p.ef_cap;
  vis.cap = p.vis_cap; vis.limit = p.vis_limit;
  vis.ovf = p.ovf + static_cast<size_t>(blockIdx.x * kWarpsPerBlock + warp) * p.ovf_cap;
  vis.ovf_cap = p.ovf_cap; vis.ovf_limit = p.ovf_limit;
  vis.count = 0; vis.ovf_count = 0; vis.failed = false;
  vis.compact = p.vis_compact != 0;
  if (lane < kNumTotals) s_tot[lane] = 0ull;
  __syncwarp();

  for (;;) {
    uint32_t q = 0;
    if (lane == 0) q = atomicAdd(p.counter, 1u);
    q = __shfl_sync(kFull, q, 0);
    if (q >= p.nq) break;

    if (p.io.in_rows) {
      // routed: the q-th query of this GPU's inbox = slot j of source segment `src`; it arrived in the stored order
      uint32_t src = 0, j = q;
      while (src < p.io.world) {
        const uint32_t c = __ldg(p.io.in_counts + src);
        if (j < c) break;
        j -= c; ++src;
      }
      if (src == p.io.world) break;  // past the last received query
      const float4* row = p.io.in_rows + (static_cast<size_t>(src) * p.io.cap + j) * g.row_f4;
      for (uint32_t f = lane; f < g.row_f4; f += 32) reinterpret_cast<float4*>(s_q)[f] = row[f];
    } else {
      // stage the query (database slot components, io/database.hh:17-21)
      const float* gq = p.queries + static_cast<size_t>(q) * g.dim;
      for (uint32_t j = lane; j < q_floats; j += 32) s_q[j] = 0.f;
      __syncwarp();
      for (uint32_t j = lane; j < g.dim; j += 32) s_q[row_pos(g.dim, j)] = __ldg(gq + j);  // stored order (graph.h)
    }
    visited_reset(vis, lane);

    // distcomps is not tracked separately: every visited node costs one distance computation, the entry point two
    // (hnsw.hh:272 and :286) -> distcomps = visited_nodes + visited_nodes_l0 + 1
    uint32_t c_unused = 0, c_vup = 0, c_vl0 = 0, c_l0 = 0, c_lup = 0;

    // hnsw.hh:261-272 — the entry point and its distance
    uint32_t cur = g.ep_row;
    if (g.ep_level > 0) ++c_vup; else ++c_vl0;
    if (lane == 0) s_rows[0] = cur;
    __syncwarp();
    eval_rows<IP, NCHUNK, SHN_SMALL_PASSES>(g, s_q, s_rows, 1, s_dist, lane);
    float closest = s_dist[0];
    __syncwarp();

    // search_for_one<without_lock>: levels ep.level .. 1 (hnsw.hh:341-391)
    for (uint32_t level = g.ep_level; level > 0; --level) {
      while (greedy_step<IP, NCHUNK>(g, s_q, level, cur, closest, s_rows, s_dist, c_unused, c_vup, c_lup, lane)) {}
    }

    // hnsw.hh:285-288 — distance recomputed (same bits), seed of the level-0 search
    if (lane == 0) { qd[0] = closest; qi[0] = cur; }
    uint32_t qsize = 1;
    visited_test_and_set(vis, cur, lane == 0, lane);

    // search_level<without_lock>(ef, level 0) (hnsw.hh:407-476)
    uint32_t c_hot = 0, c_local = 0, c_halo = 0;
    const uint32_t l0_before = c_vl0;
    beam_search<IP, NCHUNK, PART, PASSES>(g, s_q, 0, p.ef, qd, qi, qsize, s_rows, s_dist, vis, c_unused, c_vl0, c_l0, c_hot, c_local, c_halo,
                                  s_read, lane);

    // trim to k (:296-298); the reference reports ids in heap-array order, here ascending by distance
    {
      uint32_t* ids_out = p.out_ids;
      float* dists_out = p.out_dists;
      size_t out_slot = q;
      if (p.io.in_rows) {  // routed: the result row goes to the landing buffer of the query's home GPU, at its home slot
        uint32_t src = 0, j = q;
        for (;;) {
          const uint32_t c = __ldg(p.io.in_counts + src);
          if (j < c) break;
          j -= c; ++src;
        }
        out_slot = p.io.in_tags[static_cast<size_t>(src) * p.io.cap + j];
        ids_out = p.io.out_ids[src];
        dists_out = p.io.out_dists ? p.io.out_dists[src] : nullptr;
      }
      for (uint32_t j = lane; j < p.k; j += 32) {
        const bool ok = j < qsize;
        ids_out[out_slot * p.k + j] = ok ? __ldg(g.ext_id + (qi[j] & ~kExpanded)) : kInvalid;
        if (dists_out) dists_out[out_slot * p.k + j] = ok ? qd[j] : __int_as_float(0x7f800000);
      }
    }
    if (lane == 0) {
      const uint32_t c_dist = c_vup + c_vl0 + 1;
      const uint32_t overflowed = vis.ovf_count ? 1u : 0u;
      if (p.per_query) {
        uint32_t* o = p.per_query + static_cast<size_t>(q) * kPerQueryWords;
        o[0] = c_dist; o[1] = c_vup; o[2] = c_vl0; o[3] = c_l0; o[4] = c_lup; o[5] = overflowed | (vis.failed ? 2u : 0u);
      }
      s_tot[kDistcomps] += c_dist; s_tot[kVisitedUpper] += c_vup; s_tot[kVisitedL0] += c_vl0;
      s_tot[kListsL0] += c_l0; s_tot[kListsUpper] += c_lup; s_tot[kOverflowQueries] += overflowed;
      s_tot[kFailedQueries] += vis.failed ? 1u : 0u;
      s_tot[kProcessed] += 1;
      if (PART && g.world > 1) {
        s_tot[kRowsHot] += c_hot; s_tot[kRowsLocal] += c_local; s_tot[kRowsHalo] += c_halo;
        s_tot[kRowsRemote] += (c_vl0 - l0_before) - c_hot - c_local - c_halo;
      }
    }
    __syncwarp();
  }

  if (vis.ovf_count) visited_reset(vis, lane);  // leave the HBM table clean for the next launch
  __syncwarp();
  if (lane < kNumTotals && s_tot[lane]) atomicAdd(p.totals + lane, s_tot[lane]);
}

uint32_t next_pow2(uint32_t v) {
  uint32_t p = 1;
  while (p < v) p <<= 1;
  return p;
}

// Shared visited table: sized for the typical query (about 24 x ef nodes are touched at m = 16), the tail spills.  2048 keys is
// the measured optimum on B200: 4096 / 8192 keys cost occupancy (ef=64: 4.0 / 1.4 M q/s against 6.0), and even filling exactly
// the shared memory that 5 CTAs/SM leave (2500 keys, no power of two needed) lost 7 % at ef=64 although it ended the spills
// there (33 instead of 84 612 overflowing queries per million).
uint32_t pick_vis_cap(uint32_t ef, uint32_t m0) {
  const uint32_t want = ef * (m0 > 32 ? 40u : 24u);
  uint32_t cap = next_pow2(want);
  if (cap < 1024) cap = 1024;
  if (cap > 2048) cap = 2048;
  return cap;
}

template <bool IP, int NCHUNK, bool PART>
cudaError_t launch_t(bool wide, const SearchParams& p, int grid, size_t smem, cudaStream_t stream) {
  auto kernel = wide ? search_kernel<IP, NCHUNK, PART, true> : search_kernel<IP, NCHUNK, PART, false>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  kernel<<<grid, kWarpsPerBlock * 32, smem, stream>>>(p);
  return cudaGetLastError();
}

template <bool IP, int NCHUNK, bool PART>
cudaError_t occupancy_t(bool wide, size_t smem, int* blocks) {
  auto kernel = wide ? search_kernel<IP, NCHUNK, PART, true> : search_kernel<IP, NCHUNK, PART, false>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, kernel, kWarpsPerBlock * 32, smem);
}

// dimensions the kernels are instantiated for at compile time (0 = any dim)
int chunk_variant(uint32_t dim) { return (dim == 96 || dim == 128 || dim == 200 || dim == 960) ? static_cast<int>(dim) : 0; }
#define DISPATCH_V(fn, ip, part, v, ...)                                                                                \
  ((v) == 128 ? fn<ip, 128, part>(__VA_ARGS__) : (v) == 96 ? fn<ip, 96, part>(__VA_ARGS__)                              \
   : (v) == 200 ? fn<ip, 200, part>(__VA_ARGS__) : (v) == 960 ? fn<ip, 960, part>(__VA_ARGS__) : fn<ip, 0, part>(__VA_ARGS__))
// PART: the handle is a partition (rows may live on peers) or counts visits for the hot set; otherwise the plain kernel
#define DISPATCH(fn, ip, part, v, ...)                                                                                 \
  ((ip) ? ((part) ? DISPATCH_V(fn, true, true, v, __VA_ARGS__) : DISPATCH_V(fn, true, false, v, __VA_ARGS__))          \
        : ((part) ? DISPATCH_V(fn, false, true, v, __VA_ARGS__) : DISPATCH_V(fn, false, false, v, __VA_ARGS__)))

}  // namespace

cudaError_t search_plan(const DeviceGraph& g, const SearchConfig& cfg, uint32_t nq, int* grid, int* block, size_t* smem,
                        uint32_t* vis_cap_out) {
  const uint32_t q_floats = g.row_f4 * 4;
  const uint32_t ef_cap = (cfg.ef + 31u) & ~31u;
  const uint32_t list_cap = g.m0 <= 32 ? 32u : 64u;
  const bool part = g.world > 1 || g.visit_count != nullptr;
  const bool wide = use_wide(g.dim, cfg.ef);
  uint32_t vis_cap = cfg.vis_cap ? (cfg.vis_cap + 3) / 4 * 4 : pick_vis_cap(cfg.ef, g.m0);
  size_t bytes = kWarpsPerBlock * warp_smem_bytes(q_floats, ef_cap, list_cap, vis_cap, part);
  while (bytes > 227 * 1024 && vis_cap > 1024) {
    vis_cap = vis_cap / 2 / 4 * 4;
    bytes = kWarpsPerBlock * warp_smem_bytes(q_floats, ef_cap, list_cap, vis_cap, part);
  }
  if (bytes > 227 * 1024) return cudaErrorInvalidValue;
  int blocks = 0;
  const int v = chunk_variant(g.dim);
  cudaError_t e = DISPATCH(occupancy_t, cfg.ip, part, v, wide, bytes, &blocks);
  if (e != cudaSuccess) return e;
  if (blocks < 1) return cudaErrorInvalidConfiguration;
  if (cfg.warps_per_sm > 0) {
    const int want = (cfg.warps_per_sm + kWarpsPerBlock - 1) / kWarpsPerBlock;
    if (want < blocks) blocks = want;
  }
  long long gsz = static_cast<long long>(blocks) * cfg.num_sms;
  const long long need = (static_cast<long long>(nq) + kWarpsPerBlock - 1) / kWarpsPerBlock;
  if (gsz > need) gsz = need;
  if (gsz < 1) gsz = 1;
  *grid = static_cast<int>(gsz);
  *block = kWarpsPerBlock * 32;
  *smem = bytes;
  *vis_cap_out = vis_cap;
  return cudaSuccess;
}

cudaError_t search_launch(const DeviceGraph& g, const SearchConfig& cfg, const float* d_queries, uint32_t nq,
                          uint32_t* d_ids, float* d_dists, uint32_t* d_per_query, SearchWorkspace& ws,
                          cudaStream_t stream, const RoutedIo* io) {
  int grid, block;
  size_t smem;
  uint32_t vis_cap;
  cudaError_t e = search_plan(g, cfg, nq, &grid, &block, &smem, &vis_cap);
  if (e != cudaSuccess) return e;
  if (static_cast<uint32_t>(grid) * kWarpsPerBlock > ws.ovf_slots) return cudaErrorInvalidValue;

  SearchParams p;
  p.g = g;
  p.queries = d_queries; p.nq = nq; p.k = cfg.k; p.ef = cfg.ef;
  p.out_ids = d_ids; p.out_dists = d_dists; p.per_query = d_per_query;
  p.counter = ws.counter; p.totals = ws.totals; p.ovf = ws.ovf;
  p.vis_cap = vis_cap; p.vis_limit = vis_cap / 4 * 3;
  p.vis_compact = (cfg.vis_compact && vis_cap == 2048 && g.n <= (1u << 24)) ? 1u : 0u;
  p.ovf_cap = ws.ovf_cap; p.ovf_limit = ws.ovf_cap / 4 * 3;
  p.q_floats = g.row_f4 * 4; p.ef_cap = (cfg.ef + 31u) & ~31u; p.list_cap = g.m0 <= 32 ? 32u : 64u;
  if (io) p.io = *io;

  e = cudaMemsetAsync(ws.counter, 0, sizeof(uint32_t), stream);
  if (e != cudaSuccess) return e;
  if (!ws.keep_totals) {
    e = cudaMemsetAsync(ws.totals, 0, kNumTotals * sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return e;
  }

  const int v = chunk_variant(g.dim);
  const bool part = g.world > 1 || g.visit_count != nullptr;
  return DISPATCH(launch_t, cfg.ip, part, v, use_wide(g.dim, cfg.ef), p, grid, smem, stream);
}

}  // namespace shn
