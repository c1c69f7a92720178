// dump_io.cu — host code: reference index dumps <-> HostGraph.
//
// File format (one file per memory node, little-endian; written verbatim from the memory node's buffer by
// src/memory_node.hh:187-195): u64 free_ptr (== bytes used), u64 ep_ptr (entry RemotePtr, memory node 0 only,
// src/rdma/rdma_reads.hh:78-86), then nodes from byte 16 in allocation order, each
//   u64 header | u32 uid | u32 level | f32 components[dim] | list0 | list[1..level] | pad to 8
// with list = u32 count + 8-byte RemotePtr slots (2m at level 0, m above; src/node/node.hh:10-19,45-46,
// src/node/node.cc:18-27), RemotePtr = memory_node << 48 | byte offset (src/remote_pointer.hh:9-22).
#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>

#include "graph.h"

namespace shn {
namespace {

inline uint64_t rd64(const uint8_t* p) { uint64_t v; std::memcpy(&v, p, 8); return v; }
inline uint32_t rd32(const uint8_t* p) { uint32_t v; std::memcpy(&v, p, 4); return v; }
inline void wr64(uint8_t* p, uint64_t v) { std::memcpy(p, &v, 8); }
inline void wr32(uint8_t* p, uint32_t v) { std::memcpy(p, &v, 4); }

template <class F>
void parallel_for(uint64_t n, F&& f) {
  unsigned nt = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
  if (n < 1u << 16) nt = 1;
  std::vector<std::thread> ts;
  const uint64_t chunk = (n + nt - 1) / nt;
  for (unsigned t = 0; t < nt; ++t) {
    const uint64_t b = t * chunk, e = std::min<uint64_t>(n, b + chunk);
    if (b >= e) break;
    if (nt == 1) { f(b, e); return; }
    ts.emplace_back([=, &f] { f(b, e); });
  }
  for (auto& t : ts) t.join();
}

}  // namespace

bool parse_dumps(const void* const* dumps, const uint64_t* sizes, int n_parts, uint32_t dim, uint32_t m, HostGraph& g,
                 std::string& err) {
  if (n_parts < 1 || dim == 0 || m == 0) { err = "parse_dumps: bad arguments"; return false; }
  g = HostGraph{};
  g.dim = dim; g.m = m;
  const uint64_t min_node = ref_node_bytes(dim) + ref_list0_bytes(m);

  // pass 1: enumerate nodes by linear scan, using `level` to size each record
  std::vector<std::vector<uint64_t>> offsets(n_parts);
  std::vector<uint64_t> first_row(n_parts + 1, 0);
  for (int p = 0; p < n_parts; ++p) {
    const uint8_t* d = static_cast<const uint8_t*>(dumps[p]);
    if (sizes[p] < 16) { err = "dump part " + std::to_string(p) + " is shorter than its 16-byte prologue"; return false; }
    uint64_t end = rd64(d);
    if (end > sizes[p] || end < 16) { err = "dump part " + std::to_string(p) + ": free_ptr does not match the file size"; return false; }
    uint64_t off = 16;
    offsets[p].reserve((end - 16) / min_node + 1);
    while (off + min_node <= end) {
      const uint32_t lvl = rd32(d + off + 12);
      if (lvl > 64 || off + ref_alloc_bytes(dim, m, lvl) > end + 4) {
        err = "dump part " + std::to_string(p) + ": record at byte " + std::to_string(off) + " is not a node for dim=" +
              std::to_string(dim) + " m=" + std::to_string(m);
        return false;
      }
      offsets[p].push_back(off);
      off += ref_alloc_bytes(dim, m, lvl);
    }
    if (off != end) { err = "dump part " + std::to_string(p) + ": trailing bytes do not form a node (wrong dim or m?)"; return false; }
    first_row[p + 1] = first_row[p] + offsets[p].size();
  }
  const uint64_t n = first_row[n_parts];
  if (n == 0) { err = "index dump holds no nodes"; return false; }
  if (n >= kInvalid) { err = "index dump holds more than 2^32-2 nodes"; return false; }
  g.n = static_cast<uint32_t>(n);
  g.vec.resize(n * dim);
  g.uid.resize(n); g.level.resize(n); g.up_base.assign(n, kInvalid);
  g.l0.assign(n * 2ull * m, kInvalid);

  // pass 2: fixed-size fields + upper-list bases
  uint64_t n_up = 0;
  for (int p = 0; p < n_parts; ++p) {
    const uint8_t* d = static_cast<const uint8_t*>(dumps[p]);
    for (size_t i = 0; i < offsets[p].size(); ++i) {
      const uint64_t r = first_row[p] + i;
      const uint8_t* nd = d + offsets[p][i];
      g.uid[r] = rd32(nd + 8);
      g.level[r] = rd32(nd + 12);
      if (g.level[r] > 0) { g.up_base[r] = static_cast<uint32_t>(n_up); n_up += g.level[r]; }
      g.max_level = std::max(g.max_level, g.level[r]);
    }
  }
  g.n_up = n_up;
  g.up.assign(n_up * m, kInvalid);

  auto resolve = [&](uint64_t rptr) -> uint32_t {
    const uint32_t mn = static_cast<uint32_t>(rptr >> 48);
    const uint64_t off = (rptr << 16) >> 16;
    if (mn >= static_cast<uint32_t>(n_parts)) return kInvalid;
    const auto& v = offsets[mn];
    auto it = std::lower_bound(v.begin(), v.end(), off);
    if (it == v.end() || *it != off) return kInvalid;
    return static_cast<uint32_t>(first_row[mn] + (it - v.begin()));
  };

  // pass 3: components + adjacency (RemotePtr -> row), in parallel over rows
  std::atomic<bool> bad{false};
  for (int p = 0; p < n_parts; ++p) {
    const uint8_t* d = static_cast<const uint8_t*>(dumps[p]);
    parallel_for(offsets[p].size(), [&](uint64_t b, uint64_t e) {
      for (uint64_t i = b; i < e; ++i) {
        const uint64_t r = first_row[p] + i;
        const uint8_t* nd = d + offsets[p][i];
        std::memcpy(&g.vec[r * dim], nd + 16, 4ull * dim);
        const uint8_t* l0 = nd + ref_node_bytes(dim);
        const uint32_t c0 = rd32(l0);
        if (c0 > 2 * m) { bad = true; continue; }
        for (uint32_t j = 0; j < c0; ++j) {
          const uint32_t row = resolve(rd64(l0 + 4 + 8ull * j));
          if (row == kInvalid) bad = true;
          g.l0[r * 2ull * m + j] = row;
        }
        for (uint32_t l = 1; l <= g.level[r]; ++l) {
          const uint8_t* lu = l0 + ref_list0_bytes(m) + (l - 1) * ref_listu_bytes(m);
          const uint32_t cu = rd32(lu);
          if (cu > m) { bad = true; continue; }
          const uint64_t u = static_cast<uint64_t>(g.up_base[r]) + (l - 1);
          for (uint32_t j = 0; j < cu; ++j) {
            const uint32_t row = resolve(rd64(lu + 4 + 8ull * j));
            if (row == kInvalid) bad = true;
            g.up[u * m + j] = row;
          }
        }
      }
    });
  }
  if (bad) { err = "index dump has a neighbour list that is over-full or points outside the dump"; return false; }
  g.ep_row = resolve(rd64(static_cast<const uint8_t*>(dumps[0]) + 8));
  if (g.ep_row == kInvalid) { err = "entry-point pointer (byte 8 of memory node 1's dump) does not address a node"; return false; }
  return true;
}

void dump_sizes(const HostGraph& g, int n_parts, uint64_t* sizes) {
  for (int p = 0; p < n_parts; ++p) sizes[p] = 16;
  for (uint32_t r = 0; r < g.n; ++r) sizes[r % n_parts] += ref_alloc_bytes(g.dim, g.m, g.level[r]);
}

void emit_dumps(const HostGraph& g, int n_parts, void* const* dumps) {
  const uint32_t dim = g.dim, m = g.m;
  std::vector<uint64_t> rptr(g.n);
  std::vector<uint64_t> cursor(n_parts, 16);
  for (uint32_t r = 0; r < g.n; ++r) {
    const int p = r % n_parts;
    rptr[r] = (static_cast<uint64_t>(p) << 48) | cursor[p];
    cursor[p] += ref_alloc_bytes(dim, m, g.level[r]);
  }
  for (int p = 0; p < n_parts; ++p) {
    uint8_t* d = static_cast<uint8_t*>(dumps[p]);
    wr64(d, cursor[p]);                                // free_ptr (memory_node.hh:61)
    wr64(d + 8, p == 0 ? rptr[g.ep_row] : 0);          // ep_ptr lives on memory node 0 only
  }
  parallel_for(g.n, [&](uint64_t b, uint64_t e) {
    for (uint64_t r = b; r < e; ++r) {
      uint8_t* nd = static_cast<uint8_t*>(dumps[r % n_parts]) + ((rptr[r] << 16) >> 16);
      const uint64_t total = ref_alloc_bytes(dim, m, g.level[r]);
      std::memset(nd, 0, total);
      wr64(nd, r == g.ep_row ? (1ull << 16) : 0);      // HEADER_ENTRY_NODE (node/node.hh:30); locks clear at rest
      wr32(nd + 8, g.uid[r]);
      wr32(nd + 12, g.level[r]);
      std::memcpy(nd + 16, &g.vec[r * dim], 4ull * dim);
      uint8_t* l0 = nd + ref_node_bytes(dim);
      uint32_t c0 = 0;
      for (uint32_t j = 0; j < 2 * m; ++j) {
        const uint32_t nb = g.l0[r * 2ull * m + j];
        if (nb == kInvalid) continue;
        wr64(l0 + 4 + 8ull * c0, rptr[nb]);
        ++c0;
      }
      wr32(l0, c0);
      for (uint32_t l = 1; l <= g.level[r]; ++l) {
        uint8_t* lu = l0 + ref_list0_bytes(m) + (l - 1) * ref_listu_bytes(m);
        const uint64_t u = static_cast<uint64_t>(g.up_base[r]) + (l - 1);
        uint32_t cu = 0;
        for (uint32_t j = 0; j < m; ++j) {
          const uint32_t nb = g.up[u * m + j];
          if (nb == kInvalid) continue;
          wr64(lu + 4 + 8ull * cu, rptr[nb]);
          ++cu;
        }
        wr32(lu, cu);
      }
    }
  });
}

}  // namespace shn
