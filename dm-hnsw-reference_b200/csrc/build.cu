// build.cu — bulk construction of the HNSW graph on the GPU, in batches.
//
// What the reference does one insert at a time under per-node RDMA spin-locks (HNSW::insert, src/hnsw/hnsw.hh:40-251;
// rdma_atomics.hh:13-132) is done here for a whole batch of new nodes per step, without locks:
//   1. insert_search_kernel — one warp per new node, on the graph as it stood before the batch: greedy descent
//      (search_for_one, :129-143), then per level ef_construction beam search (search_level, :153) and the
//      neighbour-selection heuristic (select_heuristic, :482-522, :163); the node's own lists are written and one
//      back-link request per selected neighbour is emitted into a fixed slot (no atomics).
//   2. the requests are sorted by (level, target, source) — a radix sort, so the result is deterministic;
//   3. link_kernel — one warp per (level, target): append the new sources if the list has room, otherwise re-run the
//      heuristic over old + new neighbours with distances to the target (:180-225).
// Nodes of one batch do not see each other; batches grow with the graph (at most 1/64 of it) so this stays a small
// perturbation.  The graph is therefore not the reference's graph — the bar is equal recall at equal M / efC / ef
// (tests/test_build.py) — but every distance is computed in the reference's arithmetic and the selection rule is
// the reference's, and the result is written in the reference's dump format by shn_index_store.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cub/device/device_radix_sort.cuh>
#include <random>
#include <vector>

#include "engine.h"
#include "search.cuh"

namespace shn {
namespace {

constexpr int kBuildWarps = 4;
constexpr int kCand = 128;  // old + new neighbours of one target in the link step
constexpr uint64_t kNoRequest = ~0ull;
constexpr uint64_t kMask30 = (1ull << 30) - 1;

struct BuildParams {
  DeviceGraph g;
  uint32_t* l0_w;
  uint32_t* up_w;
  const uint32_t* level;
  uint32_t batch_begin, batch_size, efc;
  const uint32_t* req_base;     // [batch_size] first request slot of each new node
  unsigned long long* req_key;  // level << 60 | target << 30 | source
  float* req_dist;
  uint32_t n_req;
  uint32_t* counter;
  unsigned long long* totals;   // [0] distcomps
  uint32_t* ovf;
  uint32_t vis_cap, vis_limit, ovf_cap, ovf_limit;
  uint32_t q_floats, ef_cap;
};

__host__ __device__ inline size_t insert_warp_smem(uint32_t q_floats, uint32_t ef_cap, uint32_t vis_cap) {
  // s_q, s_c, queue (d,i), s_rows/s_dist [64], sel rows/dist [64], tmp [64], visited
  return 4ull * (2 * q_floats + 2 * ef_cap + 5 * kMaxList + vis_cap);
}
__host__ __device__ inline size_t link_warp_smem(uint32_t q_floats) {
  // s_q, s_c, raw rows/dist [128], sorted rows/dist [128], sel rows/dist [64], tmp [64]
  return 4ull * (2 * q_floats + 4 * kCand + 3 * kMaxList);
}

// select_heuristic (hnsw.hh:482-522) over candidates sorted ascending by distance to the query: the nearest is kept,
// then a candidate is kept iff it is not closer to an already kept one than to the query.  Fewer than m_target
// candidates are all kept (:483).  Returns the number kept (sel_rows / sel_dist).
template <bool IP, int NCHUNK>
__device__ __forceinline__ uint32_t select_neighbors(const DeviceGraph& g, const uint32_t* c_rows, const float* c_dist,
                                                     uint32_t n_cand, uint32_t m_target, float* s_c, uint32_t* sel_rows,
                                                     float* sel_dist, float* s_tmp, unsigned long long& distcomps, int lane) {
  if (n_cand < m_target) {
    for (uint32_t j = lane; j < n_cand; j += 32) { sel_rows[j] = c_rows[j] & ~kExpanded; sel_dist[j] = c_dist[j]; }
    __syncwarp();
    return n_cand;
  }
  if (lane == 0) { sel_rows[0] = c_rows[0] & ~kExpanded; sel_dist[0] = c_dist[0]; }
  __syncwarp();
  uint32_t ns = 1;
  float4* s_c4 = reinterpret_cast<float4*>(s_c);
  for (uint32_t ci = 1; ci < n_cand && ns < m_target; ++ci) {
    const uint32_t row = c_rows[ci] & ~kExpanded;
    const float dc = c_dist[ci];
    for (uint32_t f = lane; f < g.row_f4; f += 32) s_c4[f] = ldg_f4(g.vec + static_cast<size_t>(row) * g.row_f4 + f);
    __syncwarp();
    eval_rows<IP, NCHUNK, SHN_SMALL_PASSES>(g, s_c, sel_rows, ns, s_tmp, lane);
    bool closer = false;
    for (uint32_t j = lane; j < ns; j += 32) closer |= s_tmp[j] < dc;  // :506
    distcomps += ns;
    if (!__any_sync(kFull, closer)) {
      if (lane == 0) { sel_rows[ns] = row; sel_dist[ns] = dc; }
      ++ns;
    }
    __syncwarp();
  }
  return ns;
}

template <bool IP, int NCHUNK>
__global__ void __launch_bounds__(kBuildWarps * 32, 4) insert_search_kernel(const BuildParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const DeviceGraph& g = p.g;

  unsigned char* base = smem_raw + warp * insert_warp_smem(p.q_floats, p.ef_cap, p.vis_cap);
  float* s_q = reinterpret_cast<float*>(base);
  float* s_c = s_q + p.q_floats;
  float* qd = s_c + p.q_floats;
  uint32_t* qi = reinterpret_cast<uint32_t*>(qd + p.ef_cap);
  uint32_t* s_rows = qi + p.ef_cap;
  float* s_dist = reinterpret_cast<float*>(s_rows + kMaxList);
  uint32_t* sel_rows = reinterpret_cast<uint32_t*>(s_dist + kMaxList);
  float* sel_dist = reinterpret_cast<float*>(sel_rows + kMaxList);
  float* s_tmp = sel_dist + kMaxList;
  VisitedSet vis;
  vis.tab = reinterpret_cast<uint32_t*>(s_tmp + kMaxList);
  vis.cap = p.vis_cap; vis.limit = p.vis_limit;
  vis.ovf = p.ovf + static_cast<size_t>(blockIdx.x * kBuildWarps + warp) * p.ovf_cap;
  vis.ovf_cap = p.ovf_cap; vis.ovf_limit = p.ovf_limit;
  vis.count = 0; vis.ovf_count = 0; vis.failed = false; vis.compact = false;

  unsigned long long t_dist = 0;
  const uint32_t m = g.m;

  for (;;) {
    uint32_t t = 0;
    if (lane == 0) t = atomicAdd(p.counter, 1u);
    t = __shfl_sync(kFull, t, 0);
    if (t >= p.batch_size) break;
    const uint32_t node = p.batch_begin + t;
    const uint32_t node_level = __ldg(p.level + node);

    // the new node's components are the query
    float4* s_q4 = reinterpret_cast<float4*>(s_q);
    for (uint32_t f = lane; f < g.row_f4; f += 32) s_q4[f] = ldg_f4(g.vec + static_cast<size_t>(node) * g.row_f4 + f);
    if (lane == 0) s_rows[0] = g.ep_row;
    __syncwarp();
    uint32_t c_dist = 0, c_vis = 0, c_lists = 0;
    eval_rows<IP, NCHUNK, SHN_SMALL_PASSES>(g, s_q, s_rows, 1, s_dist, lane);
    uint32_t cur = g.ep_row;
    float closest = s_dist[0];
    ++c_dist;
    __syncwarp();

    // hnsw.hh:129-143 — greedy descent through the levels above the node's own
    uint32_t level = g.ep_level;
    for (; level > node_level; --level) {
      while (greedy_step<IP, NCHUNK>(g, s_q, level, cur, closest, s_rows, s_dist, c_dist, c_vis, c_lists, lane)) {}
    }

    // hnsw.hh:151-231 — connect on every level from min(node_level, top) down to 0
    const uint32_t req0 = __ldg(p.req_base + t);
    bool any_failed = false;  // visited_reset clears vis.failed: remember an overflow on any level
    for (int lv = static_cast<int>(level); lv >= 0; --lv) {
      any_failed |= vis.failed;
      visited_reset(vis, lane);
      if (lane == 0) { qd[0] = closest; qi[0] = cur; }
      uint32_t qsize = 1;
      visited_test_and_set(vis, cur, lane == 0, lane);
      uint32_t c_hot = 0, c_local = 0;
      beam_search<IP, NCHUNK>(g, s_q, lv, p.efc, qd, qi, qsize, s_rows, s_dist, vis, c_dist, c_vis, c_lists, c_hot, c_local, c_hot, nullptr, lane);

      const uint32_t ns = select_neighbors<IP, NCHUNK>(g, qi, qd, qsize, m, s_c, sel_rows, sel_dist, s_tmp, t_dist, lane);

      // the node's own list (:165-175) and one back-link request per selected neighbour (:180)
      uint32_t* own = lv == 0 ? p.l0_w + static_cast<size_t>(node) * g.m0
                              : p.up_w + (static_cast<size_t>(__ldg(g.up_base + node)) + (lv - 1)) * m;
      const uint32_t slot = req0 + static_cast<uint32_t>(lv) * m;
      for (uint32_t j = lane; j < ns; j += 32) {
        const uint32_t nb = sel_rows[j];
        own[j] = nb;
        p.req_key[slot + j] = (static_cast<unsigned long long>(lv) << 60) | (static_cast<unsigned long long>(nb) << 30) | node;
        p.req_dist[slot + j] = sel_dist[j];
      }
      // only the nearest selected node seeds the next level (:228-230); select keeps candidates ascending, so [0]
      cur = sel_rows[0];
      closest = sel_dist[0];
      __syncwarp();
    }
    t_dist += c_dist;
    if ((any_failed || vis.failed) && lane == 0) atomicAdd(p.totals + 1, 1ull);
  }
  if (vis.ovf_count) visited_reset(vis, lane);
  if (lane == 0) atomicAdd(p.totals, t_dist);
}

template <bool IP, int NCHUNK>
__global__ void __launch_bounds__(kBuildWarps * 32, 4) link_kernel(const BuildParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const DeviceGraph& g = p.g;

  unsigned char* base = smem_raw + warp * link_warp_smem(p.q_floats);
  float* s_q = reinterpret_cast<float*>(base);
  float* s_c = s_q + p.q_floats;
  uint32_t* raw_rows = reinterpret_cast<uint32_t*>(s_c + p.q_floats);
  float* raw_dist = reinterpret_cast<float*>(raw_rows + kCand);
  uint32_t* c_rows = reinterpret_cast<uint32_t*>(raw_dist + kCand);
  float* c_dist = reinterpret_cast<float*>(c_rows + kCand);
  uint32_t* sel_rows = reinterpret_cast<uint32_t*>(c_dist + kCand);
  float* sel_dist = reinterpret_cast<float*>(sel_rows + kMaxList);
  float* s_tmp = sel_dist + kMaxList;

  unsigned long long t_dist = 0;
  const uint32_t m = g.m;

  for (;;) {
    uint32_t r = 0;
    if (lane == 0) r = atomicAdd(p.counter, 1u);
    r = __shfl_sync(kFull, r, 0);
    if (r >= p.n_req) break;
    const unsigned long long key = p.req_key[r];
    if (key == kNoRequest) break;  // sorted: nothing but empty slots from here on
    const unsigned long long group = key >> 30;
    if (r > 0 && (p.req_key[r - 1] >> 30) == group) continue;  // not the head of its (level, target) group

    const uint32_t lv = static_cast<uint32_t>(key >> 60);
    const uint32_t target = static_cast<uint32_t>(group & kMask30);
    const uint32_t cap = lv == 0 ? g.m0 : m;
    uint32_t* list = lv == 0 ? p.l0_w + static_cast<size_t>(target) * g.m0
                             : p.up_w + (static_cast<size_t>(__ldg(g.up_base + target)) + (lv - 1)) * m;

    // the group's requests, ascending by source (= insertion order); at most 64 are honoured
    uint32_t gsz = 0;
    for (uint32_t b = 0; b < 64; b += 32) {
      const uint32_t i = r + b + lane;
      const bool in = i < p.n_req && (p.req_key[i] >> 30) == group;
      const uint32_t mask = __ballot_sync(kFull, in);
      const uint32_t run = mask == kFull ? 32u : static_cast<uint32_t>(__ffs(~mask) - 1);
      gsz += run;
      if (run < 32) break;
    }
    // the current list
    uint32_t cnt_old = 0;
    for (uint32_t j0 = 0; j0 < cap; j0 += 32) {
      const uint32_t j = j0 + lane;
      const uint32_t nb = j < cap ? list[j] : kInvalid;
      if (nb != kInvalid) raw_rows[j] = nb;
      cnt_old += __popc(__ballot_sync(kFull, nb != kInvalid));
    }
    __syncwarp();

    if (cnt_old + gsz <= cap) {  // room: append (:193-196)
      for (uint32_t i = lane; i < gsz; i += 32) list[cnt_old + i] = static_cast<uint32_t>(p.req_key[r + i] & kMask30);
      continue;
    }

    // shrink: distances from the target to its old neighbours, then the heuristic over old + new (:197-222)
    float4* s_q4 = reinterpret_cast<float4*>(s_q);
    for (uint32_t f = lane; f < g.row_f4; f += 32) s_q4[f] = ldg_f4(g.vec + static_cast<size_t>(target) * g.row_f4 + f);
    __syncwarp();
    eval_rows<IP, NCHUNK, SHN_SMALL_PASSES>(g, s_q, raw_rows, cnt_old, raw_dist, lane);
    t_dist += cnt_old;
    for (uint32_t i = lane; i < gsz; i += 32) {
      raw_rows[cnt_old + i] = static_cast<uint32_t>(p.req_key[r + i] & kMask30);
      raw_dist[cnt_old + i] = p.req_dist[r + i];
    }
    __syncwarp();
    const uint32_t n_cand = cnt_old + gsz;
    // sort_ascending (heap.hh:53-57): by distance, ties by id — rank by counting
    for (uint32_t x = lane; x < n_cand; x += 32) {
      const float dx = raw_dist[x];
      const uint32_t rx = raw_rows[x];
      uint32_t rank = 0;
      for (uint32_t y = 0; y < n_cand; ++y) {
        const float dy = raw_dist[y];
        rank += (dy < dx || (dy == dx && raw_rows[y] < rx)) ? 1u : 0u;
      }
      c_rows[rank] = rx;
      c_dist[rank] = dx;
    }
    __syncwarp();
    const uint32_t ns = select_neighbors<IP, NCHUNK>(g, c_rows, c_dist, n_cand, cap, s_c, sel_rows, sel_dist, s_tmp, t_dist, lane);
    for (uint32_t j = lane; j < cap; j += 32) list[j] = j < ns ? sel_rows[j] : kInvalid;
    __syncwarp();
  }
  if (lane == 0) atomicAdd(p.totals, t_dist);
}

// Test hook: select_neighbors on one candidate set (rows of the index, ascending by distance), one warp.
template <bool IP, int NCHUNK>
__global__ void select_probe_kernel(const DeviceGraph g, const uint32_t* cand_rows, const float* cand_dist, uint32_t n_cand,
                                    uint32_t m_target, uint32_t* out_rows, uint32_t* out_n, unsigned long long* out_distcomps) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  float* s_c = reinterpret_cast<float*>(smem_raw);
  uint32_t* sel_rows = reinterpret_cast<uint32_t*>(s_c + g.row_f4 * 4);
  float* sel_dist = reinterpret_cast<float*>(sel_rows + kMaxList);
  float* s_tmp = sel_dist + kMaxList;
  unsigned long long dc = 0;
  const uint32_t ns = select_neighbors<IP, NCHUNK>(g, cand_rows, cand_dist, n_cand, m_target, s_c, sel_rows, sel_dist, s_tmp, dc, lane);
  for (uint32_t j = lane; j < ns; j += 32) out_rows[j] = sel_rows[j];
  if (lane == 0) { *out_n = ns; *out_distcomps = dc; }
}
template <bool IP, int NCHUNK>
cudaError_t launch_select_probe(const DeviceGraph& g, const uint32_t* cand_rows, const float* cand_dist, uint32_t n_cand,
                                uint32_t m_target, uint32_t* out_rows, uint32_t* out_n, unsigned long long* out_dc, cudaStream_t s) {
  const size_t smem = 4ull * (g.row_f4 * 4 + 3 * kMaxList);
  select_probe_kernel<IP, NCHUNK><<<1, 32, smem, s>>>(g, cand_rows, cand_dist, n_cand, m_target, out_rows, out_n, out_dc);
  return cudaGetLastError();
}

template <bool IP, int NCHUNK>
cudaError_t launch_insert(const BuildParams& p, int grid, size_t smem, cudaStream_t s) {
  insert_search_kernel<IP, NCHUNK><<<grid, kBuildWarps * 32, smem, s>>>(p);
  return cudaGetLastError();
}
template <bool IP, int NCHUNK>
cudaError_t launch_link(const BuildParams& p, int grid, size_t smem, cudaStream_t s) {
  link_kernel<IP, NCHUNK><<<grid, kBuildWarps * 32, smem, s>>>(p);
  return cudaGetLastError();
}
template <bool IP, int NCHUNK>
cudaError_t setup_t(size_t smem_insert, size_t smem_link, int* occ_insert, int* occ_link) {
  cudaError_t e = cudaFuncSetAttribute(insert_search_kernel<IP, NCHUNK>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_insert));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(link_kernel<IP, NCHUNK>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_link));
  if (e != cudaSuccess) return e;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ_insert, insert_search_kernel<IP, NCHUNK>, kBuildWarps * 32, smem_insert);
  if (e != cudaSuccess) return e;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ_link, link_kernel<IP, NCHUNK>, kBuildWarps * 32, smem_link);
}

#define DISPATCH_V(fn, ip, v, ...)                                                                   \
  ((v) == 128 ? fn<ip, 128>(__VA_ARGS__) : (v) == 96 ? fn<ip, 96>(__VA_ARGS__) : (v) == 200 ? fn<ip, 200>(__VA_ARGS__) \
   : (v) == 960 ? fn<ip, 960>(__VA_ARGS__) : fn<ip, 0>(__VA_ARGS__))
#define DISPATCH(fn, ip, v, ...) ((ip) ? DISPATCH_V(fn, true, v, __VA_ARGS__) : DISPATCH_V(fn, false, v, __VA_ARGS__))

uint32_t next_pow2(uint32_t v) {
  uint32_t p = 1;
  while (p < v) p <<= 1;
  return p;
}

}  // namespace

// Levels as HNSW::insert draws them (hnsw.hh:48,563-564): floor(-ln(U) / ln(m)) with U from
// uniform_real_distribution<double>(0,1) over mt19937(seed), never more than one above the current top (:106);
// one draw per insert, in insertion order, as a single-coroutine reference build consumes them.
void draw_levels(uint64_t n, uint32_t m, uint32_t seed, std::vector<uint32_t>& level) {
  std::mt19937 prng(seed);
  std::uniform_real_distribution<double> uniform(0., 1.);
  const double norm = 1. / std::log(static_cast<double>(m));
  level.resize(n);
  uint32_t top = 0;
  for (uint64_t i = 0; i < n; ++i) {
    uint32_t l = static_cast<uint32_t>(std::floor(-std::log(uniform(prng)) * norm));
    if (i == 0) l = 0;  // the node that initialises the index is written at level 0 (:61-63)
    if (l > top) { l = top + 1; top = l; }
    level[i] = l;
  }
}

cudaError_t select_probe(const DeviceGraph& g, bool ip, const uint32_t* d_cand_rows, const float* d_cand_dist, uint32_t n_cand,
                         uint32_t m_target, uint32_t* d_out_rows, uint32_t* d_out_n, unsigned long long* d_out_dc, cudaStream_t s) {
  const uint32_t d_rt = g.dim;
  const int v = (d_rt == 96 || d_rt == 128 || d_rt == 200 || d_rt == 960) ? static_cast<int>(d_rt) : 0;
  return DISPATCH(launch_select_probe, ip, v, g, d_cand_rows, d_cand_dist, n_cand, m_target, d_out_rows, d_out_n, d_out_dc, s);
}

cudaError_t build_graph(BuildJob& job, cudaStream_t stream) {
  const uint32_t n = job.n, m = job.g.m;
  // a back-link request packs its level into the 4 bits above target (30) and source (30): m = 2 or 3 with very many nodes
  // could reach level 16
  for (uint32_t l : *job.level_host) if (l > 15) return cudaErrorInvalidValue;
  const bool ip = job.ip;
  const uint32_t d_rt = job.g.dim;
  const int v = (d_rt == 96 || d_rt == 128 || d_rt == 200 || d_rt == 960) ? static_cast<int>(d_rt) : 0;
  const uint32_t q_floats = job.g.row_f4 * 4;
  const uint32_t ef_cap = (job.efc + 31u) & ~31u;
  uint32_t vis_cap = next_pow2(std::min<uint32_t>(std::max<uint32_t>(job.efc * 12u, 1024u), 4096u));
  size_t smem_insert = kBuildWarps * insert_warp_smem(q_floats, ef_cap, vis_cap);
  while (smem_insert > 200 * 1024 && vis_cap > 1024) { vis_cap >>= 1; smem_insert = kBuildWarps * insert_warp_smem(q_floats, ef_cap, vis_cap); }
  const size_t smem_link = kBuildWarps * link_warp_smem(q_floats);
  if (smem_insert > 227 * 1024 || smem_link > 227 * 1024) return cudaErrorInvalidValue;
  int occ_i = 0, occ_l = 0;
  cudaError_t e = DISPATCH(setup_t, ip, v, smem_insert, smem_link, &occ_i, &occ_l);
  if (e != cudaSuccess) return e;
  if (occ_i < 1 || occ_l < 1) return cudaErrorInvalidConfiguration;
  const int grid_i_max = occ_i * job.num_sms, grid_l_max = occ_l * job.num_sms;

  const uint32_t batch_max = job.batch_max ? job.batch_max : 16384;
  const uint32_t batch_div = job.batch_div ? job.batch_div : 64;  // a batch is at most 1/batch_div of the graph it is inserted into
  // request slots: (levels linked + 1) * m per new node, laid out per batch
  const std::vector<uint32_t>& level = *job.level_host;
  uint32_t max_levels_sum = 0;
  {
    // upper bound over any window of batch_max nodes: count levels (cheap exact scan)
    uint64_t run = 0;
    for (uint32_t i = 0; i < n; ++i) {
      run += level[i] + 1;
      if (i >= batch_max) run -= level[i - batch_max] + 1;
      max_levels_sum = std::max<uint32_t>(max_levels_sum, static_cast<uint32_t>(run));
    }
  }
  const size_t req_cap = static_cast<size_t>(max_levels_sum) * m;
  unsigned long long *key_a = nullptr, *key_b = nullptr;
  float *val_a = nullptr, *val_b = nullptr;
  uint32_t *d_req_base = nullptr, *d_counter = nullptr, *d_ovf = nullptr;
  unsigned long long* d_totals = nullptr;
  void* d_temp = nullptr;
  size_t temp_bytes = 0;
  const uint32_t ovf_cap = next_pow2(std::min<uint32_t>(std::max<uint32_t>(64u * job.efc, 1024u), 1u << 16));
  const size_t ovf_words = static_cast<size_t>(grid_i_max) * kBuildWarps * ovf_cap;

#define CK(x) do { e = (x); if (e != cudaSuccess) goto done; } while (0)
  CK(cudaMalloc(&key_a, req_cap * sizeof(unsigned long long)));
  CK(cudaMalloc(&key_b, req_cap * sizeof(unsigned long long)));
  CK(cudaMalloc(&val_a, req_cap * sizeof(float)));
  CK(cudaMalloc(&val_b, req_cap * sizeof(float)));
  CK(cudaMalloc(&d_req_base, batch_max * sizeof(uint32_t)));
  CK(cudaMalloc(&d_counter, sizeof(uint32_t)));
  CK(cudaMalloc(&d_totals, 2 * sizeof(unsigned long long)));
  CK(cudaMalloc(&d_ovf, ovf_words * sizeof(uint32_t)));
  CK(cudaMemsetAsync(d_ovf, 0xFF, ovf_words * sizeof(uint32_t), stream));
  CK(cudaMemsetAsync(d_totals, 0, 2 * sizeof(unsigned long long), stream));
  {
    cub::DoubleBuffer<unsigned long long> kb(key_a, key_b);
    cub::DoubleBuffer<float> vb(val_a, val_b);
    CK(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, kb, vb, static_cast<int>(req_cap), 0, 64, stream));
    CK(cudaMalloc(&d_temp, temp_bytes));
  }

  {
    BuildParams p;
    p.g = job.g;
    p.l0_w = job.l0_w; p.up_w = job.up_w; p.level = job.level_dev; p.efc = job.efc;
    p.req_base = d_req_base; p.counter = d_counter; p.totals = d_totals; p.ovf = d_ovf;
    p.vis_cap = vis_cap; p.vis_limit = vis_cap / 4 * 3; p.ovf_cap = ovf_cap; p.ovf_limit = ovf_cap / 4 * 3;
    p.q_floats = q_floats; p.ef_cap = ef_cap;

    uint32_t ep = 0, ep_level = level[0];
    std::vector<uint32_t> req_base_host(batch_max);
    uint32_t inserted = 1;  // the first node is the first entry point (:56-85)
    while (inserted < n) {
      uint32_t bsz = std::max<uint32_t>(1, inserted / batch_div);
      bsz = std::min<uint32_t>(bsz, std::min<uint32_t>(batch_max, n - inserted));
      uint32_t slots = 0;
      uint32_t batch_top = 0, batch_top_node = 0;
      for (uint32_t t = 0; t < bsz; ++t) {
        const uint32_t lv = level[inserted + t];
        req_base_host[t] = slots;
        slots += (std::min(lv, ep_level) + 1) * m;
        if (lv > batch_top) { batch_top = lv; batch_top_node = inserted + t; }
      }
      CK(cudaMemcpyAsync(d_req_base, req_base_host.data(), bsz * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
      CK(cudaMemsetAsync(key_a, 0xFF, slots * sizeof(unsigned long long), stream));
      CK(cudaMemsetAsync(d_counter, 0, sizeof(uint32_t), stream));
      p.g.ep_row = ep; p.g.ep_level = ep_level;
      p.batch_begin = inserted; p.batch_size = bsz; p.n_req = slots;
      p.req_key = key_a; p.req_dist = val_a;
      const int grid_i = std::max(1, std::min<int>(grid_i_max, (bsz + kBuildWarps - 1) / kBuildWarps));
      CK(DISPATCH(launch_insert, ip, v, p, grid_i, smem_insert, stream));

      cub::DoubleBuffer<unsigned long long> kb(key_a, key_b);
      cub::DoubleBuffer<float> vb(val_a, val_b);
      CK(cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, kb, vb, static_cast<int>(slots), 0, 64, stream));
      CK(cudaMemsetAsync(d_counter, 0, sizeof(uint32_t), stream));
      p.req_key = kb.Current(); p.req_dist = vb.Current();
      const int grid_l = std::max(1, std::min<int>(grid_l_max, (slots + kBuildWarps - 1) / kBuildWarps));
      CK(DISPATCH(launch_link, ip, v, p, grid_l, smem_link, stream));
      // the copy of req_base for the next batch must not overtake this batch's kernels
      CK(cudaStreamSynchronize(stream));

      if (batch_top > ep_level) { ep = batch_top_node; ep_level = batch_top; }  // new top level (:237-247)
      inserted += bsz;
      if (job.progress && (inserted == n || (inserted / bsz) % 64 == 0)) job.progress(inserted, n);
    }
    job.ep_row = ep; job.ep_level = ep_level;
    unsigned long long totals[2];
    CK(cudaMemcpyAsync(totals, d_totals, sizeof totals, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    job.distcomps = totals[0];
    job.failed = totals[1];
  }
done:
#undef CK
  cudaFree(key_a); cudaFree(key_b); cudaFree(val_a); cudaFree(val_b); cudaFree(d_req_base); cudaFree(d_counter);
  cudaFree(d_totals); cudaFree(d_ovf); cudaFree(d_temp);
  return e;
}

}  // namespace shn
