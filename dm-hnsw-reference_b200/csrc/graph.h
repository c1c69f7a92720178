// graph.h — host- and device-side views of an HNSW index (internal to libshn_b200.so).
//
// The reference keeps one variable-size record per node (header | uid | level | components | list0 | lists 1..L,
// src/node/node.hh:10-19) addressed by 64-bit RemotePtrs.  That layout is hostile to a GPU (8-byte pointers at
// 4-byte alignment, variable stride), so nodes are renumbered to dense rows in dump-scan order and stored SoA:
//
//   vec      [n][row_f4] float4   components, row padded to 32 B so a row is whole sectors
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
struct DeviceGraph {
  const float4* vec;
  const uint32_t* l0;
  const uint32_t* up_base;
  const uint32_t* up;
  const uint32_t* ext_id;
  uint32_t n, dim, m, m0;
  uint32_t row_f4;    // row stride of vec in float4
  uint32_t ep_row, ep_level;
};

}  // namespace shn
