// oracle/shim: replaces src/router/query_router.hh for the oracle build.
// TEST INFRASTRUCTURE ONLY.  hnsw::schedule<D,false> (src/hnsw/scheduler.hh:66-75) only
// reads `done`, `queue_size` and `query_queue`; with one compute node and no --routing
// the reference fills them exactly like this (src/compute_node.cc:225-232).
#pragma once
#include <atomic>

#include "common/types.hh"
#include "io/database.hh"

namespace query_router {
template <class Distance>
class QueryRouter {
public:
  explicit QueryRouter(size_t num_slots) {
    queue_size += static_cast<i32>(num_slots);
    for (idx_t slot = 0; slot < num_slots; ++slot) query_queue.enqueue(slot);
    done = true;
  }
  concurrent_queue<idx_t> query_queue;
  std::atomic<i32> queue_size{0};
  std::atomic<bool> done{false};
};
}  // namespace query_router
