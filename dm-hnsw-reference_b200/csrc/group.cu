// group.cu — one process driving several GPUs: the multi-GPU fan-out of the search path behind include/shn.h.
//
// What the reference spreads over compute nodes and memory nodes (src/compute_node.cc:110-131 placement + warm-up,
// :191-245 routed query phase, src/io/read_data.hh:58 round-robin query split) happens here inside one handle:
//   shn_group_create : every GPU's full index -> its partition (hot set + own share, shn_index_partition), shares attached
//                      through raw peer pointers, optional placement by k-means cluster and one router per GPU
//   shn_group_search : queries dealt round-robin to the GPUs (query id % n), one host thread per GPU; with routing each
//                      GPU scatters its shard into the peers' inboxes, a host barrier, every GPU searches what arrived
//                      and writes the results into the home GPU's landing buffer, a second barrier, results copied out
//                      in query order.  No library collective: the exchange is peer stores over NVLink.
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "handle.h"

using namespace shn;

struct shn_group {
  int n = 0;
  bool routing = false;
  uint64_t max_batch = 0;
  uint32_t k_max = 0;
  std::vector<shn_index*> part;
  std::vector<shn_router*> router;
  std::vector<float*> d_q;        // per GPU: staging for its shard of a batch
  std::vector<uint32_t*> d_ids;   // unrouted mode: device results
  std::vector<float*> d_dists;
  std::vector<cudaStream_t> stream;
  double fit_ms = 0., partition_ms = 0.;
};

namespace {
double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
// all `n` host threads (one per GPU) meet here; reusable
class HostBarrier {
 public:
  explicit HostBarrier(int n) : n_(n) {}
  void arrive_and_wait() {
    std::unique_lock<std::mutex> lk(m_);
    const unsigned gen = gen_;
    if (++count_ == n_) { count_ = 0; ++gen_; cv_.notify_all(); }
    else cv_.wait(lk, [&] { return gen_ != gen; });
  }
 private:
  std::mutex m_;
  std::condition_variable cv_;
  int n_, count_ = 0;
  unsigned gen_ = 0;
};
}  // namespace

extern "C" {

void shn_group_free(shn_group* g) {
  if (!g) return;
  for (auto* r : g->router) shn_router_free(r);
  for (int i = 0; i < g->n; ++i) {
    if (i < static_cast<int>(g->part.size()) && g->part[i]) {
      cudaSetDevice(g->part[i]->gpu);
      if (i < static_cast<int>(g->d_q.size())) { cudaFree(g->d_q[i]); cudaFree(g->d_ids[i]); cudaFree(g->d_dists[i]); }
      if (i < static_cast<int>(g->stream.size()) && g->stream[i]) cudaStreamDestroy(g->stream[i]);
    }
  }
  for (auto* p : g->part) shn_index_free(p);
  delete g;
}

int shn_group_create(shn_group** out, shn_index* const* full, int n_gpus, uint32_t cache_ratio_pct, int placement_by_cluster,
                     int routing, double slack, uint64_t max_batch, uint32_t k_max, uint32_t seed) {
  if (!out || !full) return fail(SHN_ERR_ARG, "null argument");
  if (n_gpus < 2 || n_gpus > 8) return fail(SHN_ERR_ARG, "a group spans 2 to 8 GPUs");
  if (routing && !placement_by_cluster) return fail(SHN_ERR_ARG, "routing needs placement by cluster (the centroids are the routing table)");
  if (max_batch == 0 || k_max == 0) return fail(SHN_ERR_ARG, "max_batch and k_max must be positive");
  for (int g = 0; g < n_gpus; ++g) {
    if (!full[g]) return fail(SHN_ERR_ARG, "null index handle");
    if (full[g]->world > 1) return fail(SHN_ERR_STATE, "the handles must be full indexes");
    for (int o = 0; o < g; ++o) if (full[o]->gpu == full[g]->gpu) return fail(SHN_ERR_ARG, "two handles on GPU %d", full[g]->gpu);
    // the graphs must be the same graph: a share is addressed with the numbering of the GPU that reads it
    if (full[g]->n != full[0]->n || full[g]->n_up != full[0]->n_up || full[g]->ep_row != full[0]->ep_row || full[g]->dim != full[0]->dim ||
        full[g]->m != full[0]->m || full[g]->metric != full[0]->metric)
      return fail(SHN_ERR_ARG, "the index on GPU %d is not the index on GPU %d (different size, levels or entry point)", full[g]->gpu, full[0]->gpu);
  }
  shn_group* grp = new shn_group();
  grp->n = n_gpus; grp->routing = routing != 0; grp->max_batch = max_batch; grp->k_max = k_max;
  grp->part.assign(n_gpus, nullptr);
  auto bail = [&](int code) { shn_group_free(grp); return code; };
  const uint32_t n = full[0]->n, dim = full[0]->dim;
  std::vector<float> centroids(static_cast<size_t>(n_gpus) * dim);
  std::vector<uint8_t> owner_host;
  if (placement_by_cluster) {
    const double t0 = now_ms();
    uint8_t* d_owner0 = nullptr;
    if (cudaSetDevice(full[0]->gpu) != cudaSuccess || cudaMalloc(&d_owner0, n) != cudaSuccess) return bail(fail(SHN_ERR_CUDA, "allocating the placement"));
    int rc = shn_placement_fit(full[0], n_gpus, seed, 0.05, centroids.data(), d_owner0, nullptr);
    owner_host.resize(n);
    if (rc == SHN_OK && cudaMemcpy(owner_host.data(), d_owner0, n, cudaMemcpyDeviceToHost) != cudaSuccess) rc = fail(SHN_ERR_CUDA, "copying the placement");
    cudaFree(d_owner0);
    if (rc != SHN_OK) return bail(rc);
    grp->fit_ms = now_ms() - t0;
  }
  const double t1 = now_ms();
  for (int g = 0; g < n_gpus; ++g) {
    uint8_t* d_owner = nullptr;
    if (placement_by_cluster) {
      if (cudaSetDevice(full[g]->gpu) != cudaSuccess || cudaMalloc(&d_owner, n) != cudaSuccess ||
          cudaMemcpy(d_owner, owner_host.data(), n, cudaMemcpyHostToDevice) != cudaSuccess)
        return bail(fail(SHN_ERR_CUDA, "staging the placement on GPU %d", full[g]->gpu));
    }
    const int rc = shn_index_partition(&grp->part[g], full[g], g, n_gpus, cache_ratio_pct, d_owner);
    if (d_owner) { cudaSetDevice(full[g]->gpu); cudaFree(d_owner); }
    if (rc != SHN_OK) return bail(rc);
  }
  for (int g = 0; g < n_gpus; ++g) {
    uint64_t raw[2];
    int rc = shn_index_partition_export(grp->part[g], nullptr, nullptr, raw);
    for (int o = 0; o < n_gpus && rc == SHN_OK; ++o) if (o != g) rc = shn_index_partition_attach(grp->part[o], g, nullptr, nullptr, raw);
    if (rc != SHN_OK) return bail(rc);
  }
  grp->partition_ms = now_ms() - t1;
  if (grp->routing) {
    grp->router.assign(n_gpus, nullptr);
    // a shard of a batch holds at most ceil(max_batch / n) queries
    const uint64_t shard = (max_batch + n_gpus - 1) / n_gpus;
    for (int g = 0; g < n_gpus; ++g) {
      const int rc = shn_router_create(&grp->router[g], grp->part[g], centroids.data(), slack, shard, k_max);
      if (rc != SHN_OK) return bail(rc);
    }
    for (int g = 0; g < n_gpus; ++g) {
      uint64_t raw = 0;
      int rc = shn_router_export(grp->router[g], nullptr, nullptr, &raw);
      for (int o = 0; o < n_gpus && rc == SHN_OK; ++o) if (o != g) rc = shn_router_attach(grp->router[o], g, -1, 0, raw);
      if (rc != SHN_OK) return bail(rc);
    }
  }
  const uint64_t shard = (max_batch + n_gpus - 1) / n_gpus;
  grp->d_q.assign(n_gpus, nullptr); grp->d_ids.assign(n_gpus, nullptr); grp->d_dists.assign(n_gpus, nullptr);
  grp->stream.assign(n_gpus, nullptr);
  for (int g = 0; g < n_gpus; ++g) {
    if (cudaSetDevice(grp->part[g]->gpu) != cudaSuccess || cudaMalloc(&grp->d_q[g], shard * dim * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&grp->d_ids[g], shard * k_max * sizeof(uint32_t)) != cudaSuccess ||
        cudaMalloc(&grp->d_dists[g], shard * k_max * sizeof(float)) != cudaSuccess ||
        cudaStreamCreateWithFlags(&grp->stream[g], cudaStreamNonBlocking) != cudaSuccess)
      return bail(fail(SHN_ERR_CUDA, "allocating the batch buffers on GPU %d", grp->part[g]->gpu));
  }
  *out = grp;
  return SHN_OK;
}

int shn_group_size(const shn_group* g) { return g ? g->n : 0; }
shn_index* shn_group_partition(shn_group* g, int i) { return (g && i >= 0 && i < g->n) ? g->part[i] : nullptr; }

int shn_group_timings(const shn_group* g, double* placement_kmeans_ms, double* placement_partition_ms) {
  if (!g) return fail(SHN_ERR_ARG, "null group");
  if (placement_kmeans_ms) *placement_kmeans_ms = g->fit_ms;
  if (placement_partition_ms) *placement_partition_ms = g->partition_ms;
  return SHN_OK;
}

int shn_group_search(shn_group* grp, const float* queries, uint64_t nq, uint32_t k, uint32_t ef, uint32_t* out_ids,
                     float* out_dists, shn_stats* per_gpu, double* routing_ms) {
  if (!grp || (!queries && nq) || (!out_ids && nq)) return fail(SHN_ERR_ARG, "null argument");
  if (nq > grp->max_batch) return fail(SHN_ERR_ARG, "batch of %llu queries exceeds the group's max_batch", static_cast<unsigned long long>(nq));
  if (k == 0 || k > grp->k_max) return fail(SHN_ERR_ARG, "k must be in [1, k_max]");
  if (ef < k) return fail(SHN_ERR_ARG, "ef_search must be >= k (hnsw.hh:36): ef=%u k=%u", ef, k);
  const int n = grp->n;
  const uint32_t dim = grp->part[0]->dim;
  std::vector<std::string> errors(n);
  std::vector<double> route_ms(n, 0.);
  HostBarrier sync(n);
  std::vector<std::thread> workers;
  for (int g = 0; g < n; ++g) {
    workers.emplace_back([&, g] {
      // GPU g takes the queries with id % n == g (io/read_data.hh:58)
      const uint64_t cnt = nq > static_cast<uint64_t>(g) ? (nq - g + n - 1) / n : 0;
      std::vector<float> shard(cnt * dim);
      for (uint64_t i = 0; i < cnt; ++i) std::memcpy(shard.data() + i * dim, queries + (g + i * n) * dim, dim * sizeof(float));
      std::vector<uint32_t> ids(cnt * k);
      std::vector<float> dists(out_dists ? cnt * k : 0);
      shn_stats st;
      std::memset(&st, 0, sizeof st);
      bool ok = true;
      auto check = [&](int rc) { if (rc != SHN_OK && ok) { ok = false; errors[g] = shn_last_error(); } };
      auto cuda = [&](cudaError_t e) { if (e != cudaSuccess && ok) { ok = false; errors[g] = cudaGetErrorString(e); } };
      cuda(cudaSetDevice(grp->part[g]->gpu));
      cudaStream_t s = grp->stream[g];
      if (!grp->routing) {
        if (cnt) check(shn_search(grp->part[g], shard.data(), cnt, k, ef, ids.data(), out_dists ? dists.data() : nullptr, &st));
      } else {
        // every thread passes every barrier, failed or not (a missing arrival would hang the others)
        const double t0 = now_ms();
        if (cnt) cuda(cudaMemcpyAsync(grp->d_q[g], shard.data(), cnt * dim * sizeof(float), cudaMemcpyHostToDevice, s));
        if (ok) check(shn_router_scatter(grp->router[g], grp->d_q[g], cnt, s));
        cuda(cudaStreamSynchronize(s));
        route_ms[g] = now_ms() - t0;
        sync.arrive_and_wait();   // every inbox is complete
        if (ok) check(shn_router_search(grp->router[g], k, ef, s, &st));
        cuda(cudaStreamSynchronize(s));
        sync.arrive_and_wait();   // every landing buffer is complete
        uint32_t* l_ids = nullptr;
        float* l_d = nullptr;
        if (ok) check(shn_router_results(grp->router[g], &l_ids, &l_d));
        if (ok && cnt) cuda(cudaMemcpyAsync(ids.data(), l_ids, cnt * k * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        if (ok && cnt && out_dists) cuda(cudaMemcpyAsync(dists.data(), l_d, cnt * k * sizeof(float), cudaMemcpyDeviceToHost, s));
        cuda(cudaStreamSynchronize(s));
      }
      if (ok) {
        for (uint64_t i = 0; i < cnt; ++i) {
          std::memcpy(out_ids + (g + i * n) * k, ids.data() + i * k, k * sizeof(uint32_t));
          if (out_dists) std::memcpy(out_dists + (g + i * n) * k, dists.data() + i * k, k * sizeof(float));
        }
      }
      if (per_gpu) per_gpu[g] = st;
    });
  }
  for (auto& w : workers) w.join();
  for (int g = 0; g < n; ++g) if (!errors[g].empty()) return fail(SHN_ERR_CUDA, "GPU %d: %s", grp->part[g]->gpu, errors[g].c_str());
  if (routing_ms) { *routing_ms = 0.; for (double v : route_ms) *routing_ms = std::max(*routing_ms, v); }
  return SHN_OK;
}

int shn_group_warmup(shn_group* grp, const float* queries, uint64_t nq, uint32_t k, uint32_t ef, uint32_t halo_ratio_pct,
                     uint64_t* halo_rows) {
  if (!grp || (!queries && nq)) return fail(SHN_ERR_ARG, "null argument");
  if (halo_rows) *halo_rows = 0;
  if (halo_ratio_pct) {
    for (int g = 0; g < grp->n; ++g) {
      const int rc = shn_index_count_visits(grp->part[g], 1);
      if (rc != SHN_OK) return rc;
    }
  }
  std::vector<uint32_t> ids(nq * k);
  int rc = shn_group_search(grp, queries, nq, k, ef, ids.data(), nullptr, nullptr, nullptr);
  if (rc != SHN_OK) return rc;
  if (halo_ratio_pct) {
    for (int g = 0; g < grp->n; ++g) {
      uint64_t rows = 0;
      rc = shn_index_partition_build_halo(grp->part[g], halo_ratio_pct, &rows);
      if (rc != SHN_OK) return rc;
      if (halo_rows) *halo_rows += rows;
    }
  }
  return SHN_OK;
}

}  // extern "C"
