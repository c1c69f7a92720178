// engine.h — internal C++ interface between the C ABI (capi.cu) and the CUDA kernels.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <vector>

#include "graph.h"

namespace shn {

// Indices into the device-side totals array (u64 each), summed over the queries of one launch.
enum Total { kDistcomps = 0, kVisitedUpper, kVisitedL0, kListsL0, kListsUpper, kOverflowQueries, kFailedQueries,
             kRowsHot, kRowsLocal, kRowsRemote, kRowsHalo, kProcessed, kNumTotals };

// Per-query counter record written when the caller asks for it (shn_search_device per_query_counters).
constexpr int kPerQueryWords = 6;  // distcomps, visited_nodes, visited_nodes_l0, lists_l0, lists_upper, overflow

struct SearchWorkspace {
  uint32_t* counter = nullptr;            // work cursor
  unsigned long long* totals = nullptr;   // [kNumTotals]
  uint32_t* ovf = nullptr;                // [ovf_slots][ovf_cap], all kInvalid at rest
  uint32_t ovf_cap = 0, ovf_slots = 0;
  bool keep_totals = false;               // the launch adds to the totals instead of starting from zero (chunked calls)
};

// Routed I/O (router.cu): the queries of a launch come from this GPU's inbox segments (one per source GPU, already in the
// stored row order) and every result row goes straight to the landing buffer of the query's home GPU, at the home slot.
struct RoutedIo {
  const float4* in_rows = nullptr;      // [world][cap][row_f4]; nullptr = plain I/O
  const uint32_t* in_tags = nullptr;    // [world][cap]: home slot of each query
  const uint32_t* in_counts = nullptr;  // [world]: queries received from each source GPU
  uint32_t cap = 0, world = 0;
  uint32_t* const* out_ids = nullptr;   // device tables [world] (local or peer-mapped)
  float* const* out_dists = nullptr;
};

struct SearchConfig {
  uint32_t k, ef;
  bool ip;
  int warps_per_sm;  // 0 = auto
  uint32_t vis_cap;  // entries of the shared-memory visited table per warp; 0 = auto
  bool vis_compact = false;  // 16-bit keys where the graph allows it (<= 2^24 ids, 2048-word table)
  int num_sms;
};

// Number of warp slots the launch will use (so the caller can size the overflow tables), given k/ef/dim.
cudaError_t search_plan(const DeviceGraph& g, const SearchConfig& cfg, uint32_t nq, int* grid, int* block, size_t* smem,
                        uint32_t* vis_cap);

// Enqueue the search of nq queries on `stream`.  d_per_query may be null.  ws.totals/ws.counter are reset on the
// stream first.  No synchronisation.
cudaError_t search_launch(const DeviceGraph& g, const SearchConfig& cfg, const float* d_queries, uint32_t nq,
                          uint32_t* d_ids, float* d_dists, uint32_t* d_per_query, SearchWorkspace& ws,
                          cudaStream_t stream, const RoutedIo* io = nullptr);

// Exact top-k by brute force over base[n][dim] (row stride = dim floats): ids are row numbers.  Synchronises.
cudaError_t bruteforce_launch(const float* d_base, uint64_t n, const float* d_queries, uint32_t nq, uint32_t dim, bool ip,
                              uint32_t k, uint32_t* d_ids, float* d_dists, int num_sms, cudaStream_t stream);

// Tensor-core variant (bruteforce_tc.cu): tcgen05 candidate generation (bf16 split operands) + exact fp32 re-rank.
bool bruteforce_tc_supported(uint32_t dim, uint32_t k);
unsigned long long bruteforce_tc_last_fallbacks();  // queries of the last launch that lacked the exactness certificate
cudaError_t bruteforce_tc_launch(const float* d_base, uint64_t n, const float* d_queries, uint32_t nq, uint32_t dim, bool ip,
                                 uint32_t k, uint32_t* d_ids, float* d_dists, int num_sms, cudaStream_t stream);

// ---- row layout (layout.cu): natural order [n][dim] <-> stored order [n][row_f4*4] (graph.h row_pos), both on device
cudaError_t rows_to_layout(const float* d_src, float* d_dst, uint64_t n, uint32_t dim, uint32_t row_f4, cudaStream_t stream);
cudaError_t rows_from_layout(const float* d_src, float* d_dst, uint64_t n, uint32_t dim, uint32_t row_f4, cudaStream_t stream);

// ---- partitioning (partition.cu)
struct PartitionJob {
  uint32_t n = 0;        // rows of the source index
  uint32_t n_flat = 0;   // size of the flat id space (graph.h)
  uint32_t hot = 0;      // flat ids [0, hot): the replicated hot set
  uint32_t own_first = 0, own = 0;  // flat ids of this GPU's share
  uint32_t row_f4 = 0, m = 0, m0 = 0;
  uint64_t n_up = 0;
  const uint32_t* new_of_old = nullptr;  // [n] device
  uint32_t* old_of_new = nullptr;        // [n_flat] device scratch
  const float4* src_vec = nullptr;
  const uint32_t *src_l0 = nullptr, *src_up_base = nullptr, *src_up = nullptr, *src_ext_id = nullptr;
  float4 *hot_vec = nullptr, *own_vec = nullptr;
  uint32_t *hot_l0 = nullptr, *own_l0 = nullptr, *hot_up_base = nullptr, *up = nullptr, *ext_id = nullptr;
};
cudaError_t partition_arrays(const PartitionJob& job, cudaStream_t stream);
cudaError_t gather_rows(const float4* src, const uint32_t* d_rows, uint32_t count, uint32_t row_f4, float4* dst, cudaStream_t s);
// halo (graph.h): copy the rows d_rows[count] of the partitioned graph g (g.halo_dir must be null) into dst_vec / dst_l0
cudaError_t halo_gather(const DeviceGraph& g, const uint32_t* d_rows, uint32_t count, float4* dst_vec, uint32_t* dst_l0, cudaStream_t s);
cudaError_t probe_gather(const float4* src, uint32_t nrows, uint32_t row_f4, double* gbs, cudaStream_t stream);

// ---- placement and routing (placement.cu)
void kmeans_host(const std::vector<float>& sample, uint32_t count, uint32_t d, int k, uint32_t seed, bool ip,
                 std::vector<float>& centroids);
cudaError_t balanced_assign(const float4* d_vec, uint32_t n, uint32_t row_f4, const float* d_cent_stored, int k, bool ip,
                            double slack, uint8_t* d_owner, std::vector<uint32_t>& sizes, cudaStream_t s);

// ---- construction (build.cu) ------------------------------------------------------------------------------------
struct BuildJob {
  DeviceGraph g;             // vec / l0 / up_base / up / ext_id / n / dim / m / m0 / row_f4 (entry point is filled per batch)
  uint32_t* l0_w = nullptr;  // writable aliases of g.l0 / g.up, all kInvalid on entry
  uint32_t* up_w = nullptr;
  const uint32_t* level_dev = nullptr;
  const std::vector<uint32_t>* level_host = nullptr;
  uint32_t n = 0, efc = 0, batch_max = 0, batch_div = 0;
  bool ip = false;
  int num_sms = 0;
  void (*progress)(uint64_t done, uint64_t total) = nullptr;
  // results
  uint32_t ep_row = 0, ep_level = 0;
  unsigned long long distcomps = 0, failed = 0;
};

void draw_levels(uint64_t n, uint32_t m, uint32_t seed, std::vector<uint32_t>& level);
cudaError_t build_graph(BuildJob& job, cudaStream_t stream);
// test hook: the neighbour-selection heuristic of the builder on one candidate set (<= 4096 rows of the index, ascending)
cudaError_t select_probe(const DeviceGraph& g, bool ip, const uint32_t* d_cand_rows, const float* d_cand_dist, uint32_t n_cand,
                         uint32_t m_target, uint32_t* d_out_rows, uint32_t* d_out_n, unsigned long long* d_out_dc, cudaStream_t s);

}  // namespace shn
