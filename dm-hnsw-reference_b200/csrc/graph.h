// graph.h — host- and device-side views of an HNSW index (internal to libshn_b200.so).
//
// The reference keeps one variable-size record per node (header | uid | level | components | list0 | lists 1..L,
// src/node/node.hh:10-19) addressed by 64-bit RemotePtrs.  That layout is hostile to a GPU (8-byte pointers at
// 4-byte alignment, variable stride), so nodes are renumbered to dense rows in dump-scan order and stored SoA:
//
//   vec      [n][row_f4] float4   components in the blocked order of row_pos() below, row stride a multiple of 128 B
//   l0       [n][2m]     u32      level-0 neighbour rows, 0xFFFFFFFF-padded (m=16: exactly one 128 B line)
//   up_base  [n]         u32      first row of the node's upper lists in `up`, 0xFFFFFFFF when level == 0
//   up       [n_up][m]   u32      upper lists; the list of level l (>=1) is row up_base + (l-1)
//   ext_id   [n]         u32      the node's uid (node/node.hh:81) = what knn() reports (hnsw.hh:302)
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

namespace shn {

constexpr uint32_t kInvalid = 0xFFFFFFFFu;

// Reference-format sizes (src/node/node.hh:45-53, src/rdma/rdma_atomics.hh:88-95).
inline uint64_t ref_node_bytes(uint32_t dim) { return 16 + 4ull * dim; }
inline uint64_t ref_list0_bytes(uint32_t m) { return 4 + 8ull * (2 * m); }
inline uint64_t ref_listu_bytes(uint32_t m) { return 4 + 8ull * m; }
inline uint64_t ref_alloc_bytes(uint32_t dim, uint32_t m, uint32_t level) {
  uint64_t s = ref_node_bytes(dim) + ref_list0_bytes(m) + level * ref_listu_bytes(m);
  while (s % 8 != 0) s += 4;
  return s;
}

// Stored order of a row's components ("AVX-lane-major blocks", see search.cuh): the 16*(dim/16) leading elements are
// grouped in blocks of 32 floats = two 16-element chunks of the reference's distance loop (distance.hh:88-110); inside
// a block the 16-byte piece t (0..7) holds {v[t], v[16+t], v[8+t], v[24+t]} — everything AVX lane t needs from the two
// chunks, arranged as two aligned pairs (first halves of chunk c and c+1 | second halves of chunk c and c+1) so that the
// two chunks go through the packed fp32x2 pipe together (search.cuh block_accumulate).  An odd last chunk fills half
// of its block (the rest is zero).  The dim%16 tail elements follow in natural order.  The row stride is rounded up to
// whole 128-byte lines.
#ifdef __CUDACC__
#define SHN_HD __host__ __device__
#else
#define SHN_HD
#endif
SHN_HD inline uint32_t row_blocks(uint32_t dim) { return ((dim >> 4) + 1) >> 1; }
SHN_HD inline uint32_t row_stride_f4(uint32_t dim) { return ((32u * row_blocks(dim) + (dim & 15u) + 31u) / 32u) * 8u; }
SHN_HD inline uint32_t row_pos(uint32_t dim, uint32_t e) {
  const uint32_t d16 = dim & ~15u;
  if (e >= d16) return 32u * row_blocks(dim) + (e - d16);
  const uint32_t w = e & 31u, q = w >> 3;  // q: quarter of the block; quarters 1 and 2 swap places inside a piece
  return (e & ~31u) + 4u * (w & 7u) + (((q & 1u) << 1) | (q >> 1));
}
// inverse of row_pos inside the blocks: stored position -> element
SHN_HD inline uint32_t row_elem(uint32_t pos) {
  const uint32_t w = pos & 31u, s = w & 3u;
  return (pos & ~31u) + 8u * (((s & 1u) << 1) | (s >> 1)) + (w >> 2);
}

struct HostGraph {
  uint32_t n = 0, dim = 0, m = 0;
  uint32_t ep_row = kInvalid, max_level = 0;
  uint64_t n_up = 0;
  std::vector<float> vec;         // [n][dim], unpadded
  std::vector<uint32_t> uid;      // [n]
  std::vector<uint32_t> level;    // [n]
  std::vector<uint32_t> l0;       // [n][2m]
  std::vector<uint32_t> up_base;  // [n]
  std::vector<uint32_t> up;       // [n_up][m]
};

// Parse reference dumps.  Returns false and fills err on malformed input.
bool parse_dumps(const void* const* dumps, const uint64_t* sizes, int n_parts, uint32_t dim, uint32_t m, HostGraph& g,
                 std::string& err);
// Byte size of each part when storing g as n_parts dumps; nodes go to part (row % n_parts).
void dump_sizes(const HostGraph& g, int n_parts, uint64_t* sizes);
// Emit dumps; each dumps[i] must have sizes[i] bytes.
void emit_dumps(const HostGraph& g, int n_parts, void* const* dumps);

// Device view, passed to kernels by value.
//
// One GPU: every row is in vec / l0 (hot == n, world == 1).
//
// Partitioned over `world` GPUs (capi.cu shn_index_partition): rows are renumbered ("flat numbering")
//     [ hot set | pad | share of GPU 0 | pad | share of GPU 1 | pad | ... ]            ids < n_flat
//     [ halo of THIS GPU ]                                                              ids >= n_flat, never stored in a list
// and vec / l0 are ONE address range per GPU in which every piece is a separate physical allocation (vmm.h VmmSpace): the
// hot set — all nodes with level > 0 plus the most visited level-0 nodes, the HBM-resident stand-in for the reference's
// compute-node cache (src/cache/cache.hh) — is this GPU's own copy, the GPU's share is local, a peer's share is the
// peer's physical memory mapped here (what was an RDMA READ, src/rdma/rdma_reads.hh, is a load over NVLink), and the pads
// make every piece start on an allocation-granule boundary.  A row therefore lives at vec + row * row_f4 wherever it
// is; the MMU does what a placement lookup would otherwise do in the instruction stream.  Nodes are placed by owner
// (shn_placement_fit: k-means cluster; or round-robin as the reference scatters them, src/compute_thread.hh:57).
//
// Halo (shn_index_partition_build_halo): local copies of the peer-owned rows THIS GPU's queries read most — with query
// routing a GPU's queries stay near its own cluster, and what they read remotely are the same border rows again and
// again.  The copies sit behind the shares in the same address range (ids n_flat + slot); the directory holds one
// {bits, prefix} pair per 32 flat ids (bit i: row 32 w + i is in the halo; prefix: halo rows before word w), n/4 bytes in
// all, so it stays in L2, and only rows of a peer's share ever consult it.
struct DeviceGraph {
  const float4* vec;
  const uint32_t* l0;
  const uint32_t* up_base;   // rows < hot only (every node with level > 0 is hot)
  const uint32_t* up;
  const uint32_t* ext_id;
  uint32_t n, dim, m, m0;    // n: size of the id space (partitions: n_flat)
  uint32_t row_f4;           // row stride of vec in float4
  uint32_t ep_row, ep_level;
  uint32_t hot, world, rank;
  uint32_t own_lo, own_hi;   // this GPU's share: flat ids [own_lo, own_hi)
  uint32_t halo_first;       // flat id of halo slot 0 (= n_flat)
  const uint2* halo_dir;     // nullptr = no halo
  uint32_t* visit_count;     // optional [n]: +1 per level-0 distance computation (warm-up passes that pick hot set and halo)
};

#ifdef __CUDACC__
// Where a level-0 row is read from.
enum RowClass : uint32_t { kRowHot = 0, kRowOwn = 1, kRowHalo = 2, kRowPeer = 3 };
// The id under which `row` is read on this GPU: itself, or its halo copy.  PART = false: the handle holds every row (one
// GPU, or the builder) and nothing of this is in the instruction stream.
template <bool PART>
__device__ __forceinline__ uint32_t read_id(const DeviceGraph& g, uint32_t row, uint32_t& cls) {
  if (!PART || row < g.hot) { cls = kRowHot; return row; }
  if (row >= g.own_lo && row < g.own_hi) { cls = kRowOwn; return row; }
  cls = kRowPeer;
  if (g.halo_dir) {
    const uint2 e = __ldg(g.halo_dir + (row >> 5));
    const uint32_t bit = 1u << (row & 31u);
    if (e.x & bit) { cls = kRowHalo; return g.halo_first + e.y + __popc(e.x & (bit - 1u)); }
  }
  return row;
}
template <bool PART>
__device__ __forceinline__ const float4* vec_row(const DeviceGraph& g, uint32_t row) {
  uint32_t cls;
  return g.vec + static_cast<size_t>(read_id<PART>(g, row, cls)) * g.row_f4;
}
template <bool PART>
__device__ __forceinline__ const uint32_t* l0_row(const DeviceGraph& g, uint32_t row) {
  uint32_t cls;
  return g.l0 + static_cast<size_t>(read_id<PART>(g, row, cls)) * g.m0;
}
#endif

}  // namespace shn
