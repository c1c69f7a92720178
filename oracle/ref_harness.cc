// oracle/ref_harness.cc — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Drives the reference's OWN hot-path code — src/hnsw/hnsw.hh (insert :40, knn :253),
// src/hnsw/scheduler.hh:20, src/hnsw/distance.hh, src/hnsw/heap.hh, src/node/*, src/rdma/*.hh,
// src/cache/*.hh, src/compute_thread.hh, src/shared_context.hh, src/buffer_allocator.hh —
// compiled UNMODIFIED from /root/reference by oracle/Makefile, with oracle/shim/ placed first
// on the include path so that ibverbs, oneTBB and huge pages resolve to in-process stand-ins.
// What the reference's WorkerPool does in src/worker_pool.hh:34-89 (allocate threads, register
// them with a SharedContext, run hnsw::schedule) is restated here for one process that also
// plays the memory node(s): a memory node is a buffer whose first word is free_ptr = 16
// (src/memory_node.hh:61) and whose [0, free_ptr) bytes ARE the index dump
// (src/memory_node.hh:187-195).
//
// Exposed as a small C API (libshine_ref.so) for tests/, bench.py --impl reference and the
// cpu_baseline leg.  Results returned per query: the ids the reference stores in
// ComputeThread::query_results (heap-array order, src/hnsw/hnsw.hh:300-303) and, since the
// reference discards distances, the distance of each returned id recomputed with the
// reference's own Distance::dist on the node's components found in the dump.
#include <sys/mman.h>

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>
#include <unordered_map>

#include "cache/cache.hh"
#include "compute_thread.hh"
// HNSW<D>::select_heuristic (hnsw.hh:482-522) is a private static member; the harness reaches it by relaxing access for
// this one include — the reference source itself stays untouched.
#define private public
#include "hnsw/hnsw.hh"
#undef private
#include "hnsw/scheduler.hh"
#include "io/database.hh"

namespace {

struct Stats {  // mirrors statistics::ThreadStatistics (src/common/statistics.hh:148-176), summed over threads
  uint64_t distcomps, rdma_reads_in_bytes, rdma_writes_in_bytes, processed, remote_allocations, allocation_size,
    visited_nodes, visited_nodes_l0, visited_neighborlists, max_level, cache_hits, cache_misses;
};

void accumulate(Stats& s, const statistics::ThreadStatistics& t) {
  s.distcomps += t.distcomps;
  s.rdma_reads_in_bytes += t.rdma_reads_in_bytes;
  s.rdma_writes_in_bytes += t.rdma_writes_in_bytes;
  s.processed += t.processed;
  s.remote_allocations += t.remote_allocations;
  s.allocation_size += t.allocation_size;
  s.visited_nodes += t.visited_nodes;
  s.visited_nodes_l0 += t.visited_nodes_l0;
  s.visited_neighborlists += t.visited_neighborlists;
  s.max_level = std::max<uint64_t>(s.max_level, t.max_level);
  s.cache_hits += t.cache_hits;
  s.cache_misses += t.cache_misses;
}

// One "compute node" worth of state (src/worker_pool.hh:13-102 restated without latches).
struct Pool {
  configuration::Configuration cfg;
  Context channel{cfg};
  ClientConnectionManager cm;
  MemoryRegionTokens tokens;
  BufferAllocator allocator;
  cache::Cache cache;
  vec<u_ptr<SharedContext<ComputeThread>>> contexts;
  vec<u_ptr<ComputeThread>> threads;

  Pool(u32 num_threads, u32 num_coroutines, const vec<byte_t*>& memory_nodes, size_t cache_bytes, bool use_cache)
      : cm(memory_nodes.size()),
        allocator(num_threads),
        cache(cache_bytes,
              cache_bytes / Node::size_until_components(),
              static_cast<size_t>(std::ceil(static_cast<f64>(cache_bytes / Node::size_until_components()) /
                                            cache::COOLING_TABLE_BUCKET_ENTRIES * cache::COOLING_TABLE_RATIO)),
              num_threads,
              use_cache) {
    for (byte_t* mn : memory_nodes) {
      tokens.push_back(std::make_unique<MemoryRegionToken>(MemoryRegionToken{reinterpret_cast<u64>(mn), 0, 0}));
    }
    for (u32 id = 0; id < num_threads; ++id) {
      // one context (= one completion list) per thread: no cross-thread polling, no lock
      contexts.push_back(std::make_unique<SharedContext<ComputeThread>>(channel, cm, allocator.get_raw_buffer(), tokens));
      threads.push_back(std::make_unique<ComputeThread>(
        id, 0, cfg.max_send_queue_wr, allocator, cache, static_cast<u32>(memory_nodes.size()), num_coroutines));
    }
    for (u32 id = 0; id < num_threads; ++id) contexts[id]->register_thread(threads[id].get());
  }
};

void fill_database(io::Database<element_t>& db, const float* rows, size_t n, u32 dim, const u32* ids) {
  db.dim = dim;
  db.num_vectors_total = std::max<size_t>(n, 10);  // progress print divides by total/10 (scheduler.hh:26-30)
  db.num_vectors_read = n;
  db.allocate();
  for (size_t i = 0; i < n; ++i) {
    std::memcpy(db.get_components(i).data(), rows + i * dim, dim * sizeof(float));
    db.set_id(i, ids ? ids[i] : static_cast<u32>(i));
  }
  db.max_slot = n;
}

template <class Fn>
double run_threads(u32 num_threads, Fn&& fn) {
  vec<std::thread> ts;
  const auto t0 = std::chrono::steady_clock::now();
  for (u32 t = 1; t < num_threads; ++t) ts.emplace_back(fn, t);
  fn(0u);  // worker 0 is the calling thread, as in compute_node.cc:348-350,383-385
  for (auto& t : ts) t.join();
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

struct SilenceStderr {  // the scheduler prints progress to stderr
  int saved{-1};
  explicit SilenceStderr(bool on) {
    if (!on) return;
    fflush(stderr);
    saved = dup(2);
    int nul = open("/dev/null", O_WRONLY);
    dup2(nul, 2);
    close(nul);
  }
  ~SilenceStderr() {
    if (saved < 0) return;
    fflush(stderr);
    dup2(saved, 2);
    close(saved);
  }
};

template <class D>
int build_impl(const float* base, u32 n, u32 dim, u32 m, u32 efc, u32 seed, u32 num_threads, u32 num_coroutines,
               u32 num_mn, uint8_t** dumps, uint64_t* dump_sizes, Stats* out_stats, double* seconds) {
  hnsw::HNSW<D> index{m, efc, 1, 1, seed, dim, false};  // sets Node::DIM etc. (hnsw.hh:37)

  // worst case: every node in one MN, plus upper lists (P(level>=l) = m^-l) with slack
  const size_t per_node = Node::total_size(0) + 8;
  const size_t cap = 4096 + static_cast<size_t>(n) * per_node + static_cast<size_t>(n) * Node::NEIGHBORLIST_SIZE / 2;
  vec<byte_t*> mns;
  for (u32 i = 0; i < num_mn; ++i) {
    void* p = mmap(nullptr, cap, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (p == MAP_FAILED) return -1;
    *static_cast<u64*>(p) = 16;  // free_ptr (memory_node.hh:61); ep_ptr at +8 stays null
    mns.push_back(static_cast<byte_t*>(p));
  }

  {
    Pool pool(num_threads, num_coroutines, mns, 0, false);
    io::Database<element_t> db;
    fill_database(db, base, n, dim, nullptr);
    db.num_vectors_total = std::max<size_t>(n, 10);
    std::atomic<idx_t> next{0};
    const double s = run_threads(num_threads, [&](u32 t) {
      hnsw::schedule<D, true>(index, next, db, num_coroutines, pool.threads[t]);
    });
    if (seconds) *seconds = s;
    if (out_stats) {
      std::memset(out_stats, 0, sizeof(Stats));
      for (auto& t : pool.threads) accumulate(*out_stats, t->stats);
    }
    for (auto& t : pool.threads) {
      for (auto& b : t->post_balances) {
        if (b != 0) return -2;  // "incomplete READs" (compute_node.cc:399-401)
      }
      t->reset();
    }
  }

  for (u32 i = 0; i < num_mn; ++i) {
    const u64 free_ptr = *reinterpret_cast<u64*>(mns[i]);
    dumps[i] = static_cast<uint8_t*>(std::malloc(free_ptr));
    std::memcpy(dumps[i], mns[i], free_ptr);  // memory_node.hh:187-195 writes exactly [0, free_ptr)
    dump_sizes[i] = free_ptr;
    munmap(mns[i], cap);
  }
  return 0;
}

// uid -> pointer to the node's components inside the dump (linear scan)
void index_nodes(const vec<byte_t*>& mns, const uint64_t* sizes, std::unordered_map<u32, const float*>& out) {
  for (size_t i = 0; i < mns.size(); ++i) {
    u64 off = 16;
    while (off + Node::size_until_components() <= sizes[i]) {
      const byte_t* p = mns[i] + off;
      const u32 uid = *reinterpret_cast<const u32*>(p + 8);
      const u32 level = *reinterpret_cast<const u32*>(p + 12);
      out[uid] = reinterpret_cast<const float*>(p + 16);
      size_t sz = Node::total_size(level);
      while (sz % 8 != 0) sz += 4;  // rdma_atomics.hh:92-95
      off += sz;
    }
  }
}

template <class D>
int search_impl(const uint8_t* const* dumps, const uint64_t* sizes, u32 num_mn, u32 dim, u32 m, u32 k, u32 ef,
                const float* queries, u32 nq, u32 num_threads, u32 num_coroutines, u32 cache_ratio_pct,
                int per_query_stats, u32* out_ids, float* out_dists, u32* out_counts, Stats* stats, double* seconds) {
  hnsw::HNSW<D> index{m, 200, k, ef, 1234, dim, cache_ratio_pct > 0};
  vec<byte_t*> mns;
  size_t total = 0;
  for (u32 i = 0; i < num_mn; ++i) {
    mns.push_back(const_cast<byte_t*>(dumps[i]));  // the query path never writes (hnsw.hh:279,290 without_lock)
    total += sizes[i];
  }
  std::unordered_map<u32, const float*> components;
  if (out_dists) index_nodes(mns, sizes, components);

  const size_t cache_bytes = static_cast<size_t>(static_cast<f64>(total) / 100. * cache_ratio_pct);
  Pool pool(num_threads, num_coroutines, mns, cache_bytes, cache_ratio_pct > 0);

  auto collect = [&](u32 q_begin, u32 q_end) {
    for (auto& t : pool.threads) {
      for (auto& [q_id, result] : t->query_results) {
        if (q_id < q_begin || q_id >= q_end) continue;
        u32 c = 0;
        for (const u32 uid : result) {
          if (c >= k) break;
          out_ids[static_cast<size_t>(q_id) * k + c] = uid;
          if (out_dists) {
            const float* comp = components.at(uid);
            out_dists[static_cast<size_t>(q_id) * k + c] =
              D::dist(span<const f32>(queries + static_cast<size_t>(q_id) * dim, dim), span<const f32>(comp, dim), dim);
          }
          ++c;
        }
        if (out_counts) out_counts[q_id] = c;
        for (; c < k; ++c) {
          out_ids[static_cast<size_t>(q_id) * k + c] = 0xFFFFFFFFu;
          if (out_dists) out_dists[static_cast<size_t>(q_id) * k + c] = std::numeric_limits<float>::infinity();
        }
      }
    }
  };

  if (!per_query_stats) {
    io::Database<element_t> db;
    fill_database(db, queries, nq, dim, nullptr);
    query_router::QueryRouter<D> router(nq);
    std::atomic<idx_t> next{0};
    const double s = run_threads(num_threads, [&](u32 t) {
      hnsw::schedule<D, false>(index, next, db, num_coroutines, pool.threads[t], &router);
    });
    if (seconds) *seconds = s;
    collect(0, nq);
    if (stats) {
      std::memset(stats, 0, sizeof(Stats));
      for (auto& t : pool.threads) accumulate(*stats, t->stats);
    }
  } else {
    // one query at a time on thread 0 so that the reference's counters can be read per query;
    // `stats` is then an array of nq records
    double s = 0;
    for (u32 q = 0; q < nq; ++q) {
      io::Database<element_t> db;
      const u32 id = q;
      fill_database(db, queries + static_cast<size_t>(q) * dim, 1, dim, &id);
      query_router::QueryRouter<D> router(1);
      std::atomic<idx_t> next{0};
      pool.threads[0]->reset();
      const auto t0 = std::chrono::steady_clock::now();
      hnsw::schedule<D, false>(index, next, db, 1, pool.threads[0], &router);
      s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      collect(q, q + 1);
      if (stats) {
        std::memset(&stats[q], 0, sizeof(Stats));
        accumulate(stats[q], pool.threads[0]->stats);
      }
    }
    if (seconds) *seconds = s;
  }
  for (auto& t : pool.threads) {
    for (auto& b : t->post_balances) {
      if (b != 0) return -2;
    }
    t->reset();
  }
  return 0;
}

// HNSW<D>::select_heuristic (hnsw.hh:482-522) on a candidate set given as (uid, distance to the query, components):
// `selected` receives the candidate indices left in top_candidates, in heap-array order after make_heap (:521).
// Returns the number selected; *distcomps the reference's counter.
template <class D>
uint32_t select_impl(const uint32_t* uids, const float* dists, const float* vectors, uint32_t c, uint32_t dim, uint32_t m,
                            uint32_t* selected, uint64_t* distcomps) {
  hnsw::HNSW<D> index{m, 200, 1, 1, 1234, dim, false};
  vec<byte_t*> none;
  byte_t* fake_mn = static_cast<byte_t*>(std::calloc(64, 1));
  none.push_back(fake_mn);
  uint32_t n_sel = 0;
  {
    Pool pool(1, 1, none, 0, false);
    auto& thread = pool.threads[0];
    MaxHeap heap;
    std::unordered_map<u32, uint32_t> index_of;
    for (uint32_t i = 0; i < c; ++i) {
      byte_t* buf = thread->buffer_allocator.allocate_node(0);  // returned to the freelist by ~Node
      std::memset(buf, 0, Node::size_until_components());
      *reinterpret_cast<u32*>(buf + 8) = uids[i];
      std::memcpy(buf + 16, vectors + static_cast<size_t>(i) * dim, dim * sizeof(float));
      heap.push({std::make_shared<Node>(buf, RemotePtr{}, thread.get()), dists[i]});
      index_of[uids[i]] = i;
    }
    const uint64_t before = thread->stats.distcomps;
    hnsw::HNSW<D>::select_heuristic(heap, m, thread);
    if (distcomps) *distcomps = thread->stats.distcomps - before;
    for (const auto& e : heap.heap) selected[n_sel++] = index_of.at(e.node->id());
    heap.clear();
  }
  std::free(fake_mn);
  return n_sel;
}

}  // namespace

extern "C" {

// Build an index over base[n][dim] with the reference's HNSW<D>::insert; dumps[i] (malloc'ed, free with
// shine_ref_free) is what memory node i+1 would write to dump/index_m<M>_efc<efC>_node<i+1>_of<num_mn>.dat.
int shine_ref_build(const float* base, uint32_t n, uint32_t dim, uint32_t m, uint32_t efc, uint32_t seed, int ip,
                    uint32_t num_threads, uint32_t num_coroutines, uint32_t num_mn, uint8_t** dumps,
                    uint64_t* dump_sizes, void* stats, double* seconds, int quiet) {
  SilenceStderr hush(quiet != 0);
  if (ip) {
    return build_impl<IPDistance>(base, n, dim, m, efc, seed, num_threads, num_coroutines, num_mn, dumps, dump_sizes,
                                  static_cast<Stats*>(stats), seconds);
  }
  return build_impl<L2Distance>(base, n, dim, m, efc, seed, num_threads, num_coroutines, num_mn, dumps, dump_sizes,
                                static_cast<Stats*>(stats), seconds);
}

// Run HNSW<D>::knn for every query through hnsw::schedule<D,false>.
int shine_ref_search(const uint8_t* const* dumps, const uint64_t* sizes, uint32_t num_mn, uint32_t dim, uint32_t m,
                     uint32_t k, uint32_t ef, int ip, const float* queries, uint32_t nq, uint32_t num_threads,
                     uint32_t num_coroutines, uint32_t cache_ratio_pct, int per_query_stats, uint32_t* out_ids,
                     float* out_dists, uint32_t* out_counts, void* stats, double* seconds, int quiet) {
  SilenceStderr hush(quiet != 0);
  if (ip) {
    return search_impl<IPDistance>(dumps, sizes, num_mn, dim, m, k, ef, queries, nq, num_threads, num_coroutines,
                                   cache_ratio_pct, per_query_stats, out_ids, out_dists, out_counts,
                                   static_cast<Stats*>(stats), seconds);
  }
  return search_impl<L2Distance>(dumps, sizes, num_mn, dim, m, k, ef, queries, nq, num_threads, num_coroutines,
                                 cache_ratio_pct, per_query_stats, out_ids, out_dists, out_counts,
                                 static_cast<Stats*>(stats), seconds);
}

// The reference's own distance functions (src/hnsw/distance.hh:153-161) for pinning the restatement.
float shine_ref_dist(const float* a, const float* b, uint32_t dim, int ip) {
  return ip ? IPDistance::dist(span<const f32>(a, dim), span<const f32>(b, dim), dim)
            : L2Distance::dist(span<const f32>(a, dim), span<const f32>(b, dim), dim);
}

uint32_t shine_ref_select_heuristic(const uint32_t* uids, const float* dists, const float* vectors, uint32_t c, uint32_t dim,
                                    uint32_t m, int ip, uint32_t* selected, uint64_t* distcomps) {
  SilenceStderr hush(true);  // the cache prints its histograms on teardown
  return ip ? select_impl<IPDistance>(uids, dists, vectors, c, dim, m, selected, distcomps)
            : select_impl<L2Distance>(uids, dists, vectors, c, dim, m, selected, distcomps);
}

void shine_ref_free(void* p) { std::free(p); }

uint32_t shine_ref_stats_words(void) { return sizeof(Stats) / sizeof(uint64_t); }

}  // extern "C"
