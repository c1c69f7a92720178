// capi.cu — the C ABI of libshn_b200.so (include/shn.h).  Owns the index handle: device arrays, stream, scratch.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "handle.h"

using namespace shn;

namespace {

thread_local std::string g_err;
// construction knobs (shn_set_build_option): process-wide, read when shn_index_build* starts
uint32_t g_build_batch_max = 0, g_build_batch_div = 0;

}  // namespace

namespace shn {
int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
}  // namespace shn

namespace shn {
int select_device(int gpu_id, int* num_sms) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(SHN_ERR_CUDA, "no CUDA device is usable (%s); libshn_b200 has no CPU path", cudaGetErrorString(e));
  if (gpu_id < 0 || gpu_id >= count) return fail(SHN_ERR_ARG, "gpu_id %d out of range (%d devices)", gpu_id, count);
  CU(cudaSetDevice(gpu_id));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, gpu_id));
  if (prop.major != 10)
    return fail(SHN_ERR_CUDA, "device %d is sm_%d%d; libshn_b200 is built for sm_100a only", gpu_id, prop.major, prop.minor);
  *num_sms = prop.multiProcessorCount;
  return SHN_OK;
}
}  // namespace shn

namespace {

// Move a parsed graph into HBM in the layout of graph.h.
int upload(shn_index* ix, const HostGraph& g) {
  ix->n = g.n; ix->dim = g.dim; ix->m = g.m; ix->ep_row = g.ep_row; ix->max_level = g.max_level; ix->n_up = g.n_up;
  ix->max_level_of_ep = g.level[g.ep_row];
  ix->row_f4 = row_stride_f4(g.dim);
  const size_t row_floats = static_cast<size_t>(ix->row_f4) * 4;
  const size_t m0 = 2ull * g.m;
  CU(cudaMalloc(&ix->d_vec, g.n * row_floats * sizeof(float)));
  CU(cudaMalloc(&ix->d_l0, g.n * m0 * sizeof(uint32_t)));
  CU(cudaMalloc(&ix->d_up_base, g.n * sizeof(uint32_t)));
  CU(cudaMalloc(&ix->d_up, std::max<size_t>(g.n_up, 1) * g.m * sizeof(uint32_t)));
  CU(cudaMalloc(&ix->d_ext_id, g.n * sizeof(uint32_t)));
  CU(cudaMalloc(&ix->d_level, g.n * sizeof(uint32_t)));
  ix->hbm_bytes = g.n * (row_floats * 4 + m0 * 4 + 12) + std::max<size_t>(g.n_up, 1) * g.m * 4;
  {  // components: natural order on the host -> stored order in HBM, in slabs through a staging buffer
    const uint64_t slab = std::min<uint64_t>(g.n, 1u << 20);
    float* stage = nullptr;
    CU(cudaMalloc(&stage, slab * g.dim * sizeof(float)));
    for (uint64_t r0 = 0; r0 < g.n; r0 += slab) {
      const uint64_t cnt = std::min<uint64_t>(slab, g.n - r0);
      cudaError_t e = cudaMemcpyAsync(stage, g.vec.data() + r0 * g.dim, cnt * g.dim * sizeof(float), cudaMemcpyHostToDevice, ix->stream);
      if (e == cudaSuccess) e = rows_to_layout(stage, reinterpret_cast<float*>(ix->d_vec) + r0 * row_floats, cnt, g.dim, ix->row_f4, ix->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(ix->stream);
      if (e != cudaSuccess) { cudaFree(stage); return fail(SHN_ERR_CUDA, "uploading components: %s", cudaGetErrorString(e)); }
    }
    cudaFree(stage);
  }
  CU(cudaMemcpyAsync(ix->d_l0, g.l0.data(), g.n * m0 * sizeof(uint32_t), cudaMemcpyHostToDevice, ix->stream));
  CU(cudaMemcpyAsync(ix->d_up_base, g.up_base.data(), g.n * sizeof(uint32_t), cudaMemcpyHostToDevice, ix->stream));
  if (g.n_up) CU(cudaMemcpyAsync(ix->d_up, g.up.data(), g.n_up * g.m * sizeof(uint32_t), cudaMemcpyHostToDevice, ix->stream));
  CU(cudaMemcpyAsync(ix->d_ext_id, g.uid.data(), g.n * sizeof(uint32_t), cudaMemcpyHostToDevice, ix->stream));
  CU(cudaMemcpyAsync(ix->d_level, g.level.data(), g.n * sizeof(uint32_t), cudaMemcpyHostToDevice, ix->stream));
  CU(cudaStreamSynchronize(ix->stream));
  uint64_t db = 0;
  for (uint32_t r = 0; r < g.n; ++r) db += ref_alloc_bytes(g.dim, g.m, g.level[r]);
  ix->dump_bytes = db;
  return SHN_OK;
}

int download(const shn_index* ix, HostGraph& g) {
  if (ix->world > 1) return fail(SHN_ERR_STATE, "a partitioned handle holds only its share of the index and cannot be stored");
  g = HostGraph{};
  g.n = ix->n; g.dim = ix->dim; g.m = ix->m; g.ep_row = ix->ep_row; g.max_level = ix->max_level; g.n_up = ix->n_up;
  const size_t row_floats = static_cast<size_t>(ix->row_f4) * 4, m0 = 2ull * ix->m;
  g.vec.resize(static_cast<size_t>(g.n) * g.dim);
  g.uid.resize(g.n); g.level.resize(g.n); g.up_base.resize(g.n);
  g.l0.resize(g.n * m0);
  g.up.resize(g.n_up * g.m);
  CU(cudaSetDevice(ix->gpu));
  CU(cudaStreamSynchronize(ix->stream));
  {
    const uint64_t slab = std::min<uint64_t>(g.n, 1u << 20);
    float* stage = nullptr;
    CU(cudaMalloc(&stage, slab * g.dim * sizeof(float)));
    for (uint64_t r0 = 0; r0 < g.n; r0 += slab) {
      const uint64_t cnt = std::min<uint64_t>(slab, g.n - r0);
      cudaError_t e = rows_from_layout(reinterpret_cast<const float*>(ix->d_vec) + r0 * row_floats, stage, cnt, g.dim, ix->row_f4, ix->stream);
      if (e == cudaSuccess) e = cudaMemcpyAsync(g.vec.data() + r0 * g.dim, stage, cnt * g.dim * sizeof(float), cudaMemcpyDeviceToHost, ix->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(ix->stream);
      if (e != cudaSuccess) { cudaFree(stage); return fail(SHN_ERR_CUDA, "downloading components: %s", cudaGetErrorString(e)); }
    }
    cudaFree(stage);
  }
  CU(cudaMemcpy(g.l0.data(), ix->d_l0, g.n * m0 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(g.up_base.data(), ix->d_up_base, g.n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  if (g.n_up) CU(cudaMemcpy(g.up.data(), ix->d_up, g.n_up * g.m * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(g.uid.data(), ix->d_ext_id, g.n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(g.level.data(), ix->d_level, g.n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  return SHN_OK;
}

}  // namespace
int shn::default_vis_compact() {
  static const int v = [] { const char* e = std::getenv("SHN_VIS_COMPACT"); return e ? std::atoi(e) : 1; }();
  return v;
}
namespace {
int new_handle(shn_index** out, int gpu_id, shn_metric metric) {
  int sms = 0;
  int rc = select_device(gpu_id, &sms);
  if (rc != SHN_OK) return rc;
  shn_index* ix = new shn_index();
  ix->gpu = gpu_id; ix->num_sms = sms; ix->metric = metric;
  cudaError_t e = cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking);
  for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaEventCreate(&ix->ev[i]);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ix->ev_last, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaMalloc(&ix->ws.counter, sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMalloc(&ix->ws.totals, kNumTotals * sizeof(unsigned long long));
  if (e != cudaSuccess) {
    shn_index_free(ix);
    return fail(SHN_ERR_CUDA, "handle setup: %s", cudaGetErrorString(e));
  }
  *out = ix;
  return SHN_OK;
}

bool read_file(const char* path, std::vector<uint8_t>& out, std::string& err) {
  FILE* f = std::fopen(path, "rb");
  if (!f) { err = std::string("cannot open ") + path; return false; }
  std::fseek(f, 0, SEEK_END);
  const long sz = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  if (sz < 0) { std::fclose(f); err = std::string("cannot size ") + path; return false; }
  out.resize(static_cast<size_t>(sz));
  const size_t got = sz ? std::fread(out.data(), 1, out.size(), f) : 0;
  std::fclose(f);
  if (got != out.size()) { err = std::string("short read on ") + path; return false; }
  return true;
}

}  // namespace

namespace shn {

// Overflow tables: one per warp slot of the launch, sized from ef.
int prepare_workspace(shn_index* ix, const SearchConfig& cfg, uint32_t nq) {
  int grid, block;
  size_t smem;
  uint32_t vis_cap;
  const DeviceGraph g = ix->view();
  cudaError_t e = search_plan(g, cfg, nq, &grid, &block, &smem, &vis_cap);
  if (e == cudaErrorInvalidValue) return fail(SHN_ERR_ARG, "ef=%u / dim=%u do not fit the per-warp shared-memory budget", cfg.ef, g.dim);
  if (e != cudaSuccess) return fail(SHN_ERR_CUDA, "search_plan: %s", cudaGetErrorString(e));
  uint32_t ovf_cap = 1024;
  while (ovf_cap < 64u * cfg.ef) ovf_cap <<= 1;
  if (ovf_cap > (1u << 17)) ovf_cap = 1u << 17;
  if (ovf_cap > 4 * ix->n) { ovf_cap = 1024; while (ovf_cap < 4 * ix->n) ovf_cap <<= 1; }
  const uint32_t slots = static_cast<uint32_t>(grid) * (block / 32);
  if (ix->ws.ovf_cap != ovf_cap || ix->ws.ovf_slots < slots) {
    const size_t words = static_cast<size_t>(slots) * ovf_cap;
    if (ix->ovf.ensure(words) != cudaSuccess) return fail(SHN_ERR_CUDA, "cannot allocate %zu bytes of visited-set overflow", words * 4);
    CU(cudaMemsetAsync(ix->ovf.p, 0xFF, ix->ovf.n * sizeof(uint32_t), ix->stream));
    CU(cudaStreamSynchronize(ix->stream));
    ix->ws.ovf = ix->ovf.p; ix->ws.ovf_cap = ovf_cap; ix->ws.ovf_slots = static_cast<uint32_t>(ix->ovf.n / ovf_cap);
  }
  return SHN_OK;
}

void fill_stats(const shn_index* ix, const unsigned long long* t, uint64_t nq, shn_stats* s) {
  s->distcomps = t[kDistcomps];
  s->visited_nodes = t[kVisitedUpper];
  s->visited_nodes_l0 = t[kVisitedL0];
  s->lists_l0 = t[kListsL0];
  s->lists_upper = t[kListsUpper];
  s->visited_neighborlists = t[kListsL0] + t[kListsUpper];
  s->algorithmic_bytes = 4ull * ix->dim * t[kDistcomps] + 4ull * (2 * ix->m) * t[kListsL0] + 4ull * ix->m * t[kListsUpper];
  // node READs = distcomps - 1 per query (hnsw.hh:271,285 share one READ); lists at their fixed slot sizes
  s->reference_layout_bytes = ref_node_bytes(ix->dim) * (t[kDistcomps] - nq) + ref_list0_bytes(ix->m) * t[kListsL0] +
                              ref_listu_bytes(ix->m) * t[kListsUpper];
  s->overflow_queries = t[kOverflowQueries];
  s->rows_hot = t[kRowsHot]; s->rows_local = t[kRowsLocal]; s->rows_remote = t[kRowsRemote]; s->rows_halo = t[kRowsHalo];
  s->processed = nq;
}

int run_search(shn_index* ix, const float* d_queries, uint64_t nq, uint32_t k, uint32_t ef, uint32_t* d_ids, float* d_dists,
               uint32_t* d_per_query, cudaStream_t stream, bool timed, const RoutedIo* io) {
  SearchConfig cfg;
  cfg.k = k; cfg.ef = ef; cfg.ip = ix->metric == SHN_IP; cfg.warps_per_sm = ix->warps_per_sm; cfg.num_sms = ix->num_sms; cfg.vis_cap = ix->vis_cap; cfg.vis_compact = ix->vis_compact != 0;
  int rc = prepare_workspace(ix, cfg, static_cast<uint32_t>(nq));
  if (rc != SHN_OK) return rc;
  // One launch per handle at a time: the work cursor, the totals and the overflow tables are per handle.  A launch on
  // another stream first waits for the previous one (a caller that alternates streams is serialised, not corrupted).
  if (ix->launched) CU(cudaStreamWaitEvent(stream, ix->ev_last, 0));
  if (timed) CU(cudaEventRecord(ix->ev[1], stream));
  cudaError_t e = search_launch(ix->view(), cfg, d_queries, static_cast<uint32_t>(nq), d_ids, d_dists, d_per_query, ix->ws, stream, io);
  if (e != cudaSuccess) return fail(SHN_ERR_CUDA, "search_launch: %s", cudaGetErrorString(e));
  if (timed) CU(cudaEventRecord(ix->ev[2], stream));
  CU(cudaEventRecord(ix->ev_last, stream));
  ix->launched = true;
  return SHN_OK;
}

int check_search_args(const shn_index* ix, uint64_t nq, uint32_t k, uint32_t ef) {
  if (!ix) return fail(SHN_ERR_ARG, "null index handle");
  if (k == 0) return fail(SHN_ERR_ARG, "k must be positive");
  if (ef < k) return fail(SHN_ERR_ARG, "ef_search must be >= k (hnsw.hh:36): ef=%u k=%u", ef, k);
  if (ef > 4096) return fail(SHN_ERR_ARG, "ef_search %u exceeds the supported maximum 4096", ef);
  if (nq >= kInvalid) return fail(SHN_ERR_ARG, "too many queries in one call");
  if (ix->world > 1 && ix->attached != ix->world) return fail(SHN_ERR_STATE, "partitioned handle: %u of %u partitions attached", ix->attached, ix->world);
  return SHN_OK;
}

}  // namespace shn

namespace {
}  // namespace

extern "C" {

int shn_index_load_mem(shn_index** out, const void* const* dumps, const uint64_t* sizes, int n_parts, uint32_t dim,
                       uint32_t m, shn_metric metric, int gpu_id) {
  if (!out || !dumps || !sizes) return fail(SHN_ERR_ARG, "null argument");
  if (n_parts < 1 || dim == 0 || m == 0 || m > 32) return fail(SHN_ERR_ARG, "need n_parts >= 1, dim >= 1, 1 <= m <= 32");
  HostGraph g;
  std::string err;
  if (!parse_dumps(dumps, sizes, n_parts, dim, m, g, err)) return fail(SHN_ERR_IO, "%s", err.c_str());
  if (g.n >= (1u << 31)) return fail(SHN_ERR_ARG, "more than 2^31 - 1 nodes (the top bit of a row id marks expanded queue entries)");
  shn_index* ix = nullptr;
  int rc = new_handle(&ix, gpu_id, metric);
  if (rc != SHN_OK) return rc;
  rc = upload(ix, g);
  if (rc != SHN_OK) { shn_index_free(ix); return rc; }
  *out = ix;
  return SHN_OK;
}

int shn_index_load(shn_index** out, const char* const* dump_paths, int n_parts, uint32_t dim, uint32_t m,
                   shn_metric metric, int gpu_id) {
  if (!out || !dump_paths || n_parts < 1) return fail(SHN_ERR_ARG, "null argument");
  std::vector<std::vector<uint8_t>> bufs(n_parts);
  std::vector<const void*> ptrs(n_parts);
  std::vector<uint64_t> sizes(n_parts);
  for (int i = 0; i < n_parts; ++i) {
    std::string err;
    if (!read_file(dump_paths[i], bufs[i], err)) return fail(SHN_ERR_IO, "%s", err.c_str());
    ptrs[i] = bufs[i].data();
    sizes[i] = bufs[i].size();
  }
  return shn_index_load_mem(out, ptrs.data(), sizes.data(), n_parts, dim, m, metric, gpu_id);
}

int shn_index_build_device(shn_index** out, const float* d_base, const uint32_t* d_ids, uint64_t n, uint32_t dim,
                           uint32_t m, uint32_t ef_construction, shn_metric metric, uint32_t seed, int gpu_id) {
  if (!out || !d_base) return fail(SHN_ERR_ARG, "null argument");
  if (n == 0 || n >= (1ull << 30)) return fail(SHN_ERR_ARG, "n must be in [1, 2^30)");
  if (dim == 0 || m < 2 || m > 32) return fail(SHN_ERR_ARG, "need dim >= 1 and 2 <= m <= 32");
  if (ef_construction == 0 || ef_construction > 4096) return fail(SHN_ERR_ARG, "ef_construction must be in [1, 4096]");
  shn_index* ix = nullptr;
  int rc = new_handle(&ix, gpu_id, metric);
  if (rc != SHN_OK) return rc;
  auto bail = [&](int code) { shn_index_free(ix); return code; };
#define CUB(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return bail(fail(SHN_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__))); } while (0)
  std::vector<uint32_t> level;
  draw_levels(n, m, seed, level);
  std::vector<uint32_t> up_base(n, kInvalid);
  uint64_t n_up = 0;
  uint32_t max_level = 0;
  for (uint64_t i = 0; i < n; ++i) {
    if (level[i]) { up_base[i] = static_cast<uint32_t>(n_up); n_up += level[i]; }
    max_level = std::max(max_level, level[i]);
  }
  ix->n = static_cast<uint32_t>(n); ix->dim = dim; ix->m = m; ix->n_up = n_up; ix->max_level = max_level;
  ix->row_f4 = row_stride_f4(dim);
  const size_t row_floats = static_cast<size_t>(ix->row_f4) * 4, m0 = 2ull * m;
  CUB(cudaMalloc(&ix->d_vec, n * row_floats * sizeof(float)));
  CUB(cudaMalloc(&ix->d_l0, n * m0 * sizeof(uint32_t)));
  CUB(cudaMalloc(&ix->d_up_base, n * sizeof(uint32_t)));
  CUB(cudaMalloc(&ix->d_up, std::max<size_t>(n_up, 1) * m * sizeof(uint32_t)));
  CUB(cudaMalloc(&ix->d_ext_id, n * sizeof(uint32_t)));
  CUB(cudaMalloc(&ix->d_level, n * sizeof(uint32_t)));
  ix->hbm_bytes = n * (row_floats * 4 + m0 * 4 + 12) + std::max<size_t>(n_up, 1) * m * 4;
  cudaStream_t s = ix->stream;
  CUB(cudaEventRecord(ix->ev[0], s));
  CUB(rows_to_layout(d_base, reinterpret_cast<float*>(ix->d_vec), n, dim, ix->row_f4, s));
  CUB(cudaMemsetAsync(ix->d_l0, 0xFF, n * m0 * sizeof(uint32_t), s));
  CUB(cudaMemsetAsync(ix->d_up, 0xFF, std::max<size_t>(n_up, 1) * m * sizeof(uint32_t), s));
  CUB(cudaMemcpyAsync(ix->d_up_base, up_base.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
  CUB(cudaMemcpyAsync(ix->d_level, level.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
  if (d_ids) {
    CUB(cudaMemcpyAsync(ix->d_ext_id, d_ids, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
  } else {
    std::vector<uint32_t> ids(n);
    for (uint64_t i = 0; i < n; ++i) ids[i] = static_cast<uint32_t>(i);
    CUB(cudaMemcpyAsync(ix->d_ext_id, ids.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    CUB(cudaStreamSynchronize(s));
  }
  BuildJob job;
  job.g = ix->view();
  job.l0_w = ix->d_l0; job.up_w = ix->d_up; job.level_dev = ix->d_level; job.level_host = &level;
  job.n = ix->n; job.efc = ef_construction; job.batch_max = g_build_batch_max; job.batch_div = g_build_batch_div; job.ip = metric == SHN_IP; job.num_sms = ix->num_sms;
  for (uint32_t l : level)
    if (l > 15) return bail(fail(SHN_ERR_ARG, "a node drew level %u: the builder supports at most 16 levels (m = %u with this many nodes)", l, m));
  cudaError_t e = build_graph(job, s);
  if (e != cudaSuccess) return bail(fail(SHN_ERR_CUDA, "build_graph: %s", cudaGetErrorString(e)));
  if (job.failed) return bail(fail(SHN_ERR_CAPACITY, "%llu inserts overflowed the visited set", job.failed));
  CUB(cudaEventRecord(ix->ev[1], s));
  CUB(cudaStreamSynchronize(s));
  float ms = 0.f;
  CUB(cudaEventElapsedTime(&ms, ix->ev[0], ix->ev[1]));
  ix->ep_row = job.ep_row; ix->max_level_of_ep = job.ep_level;
  uint64_t db = 0;
  for (uint64_t r = 0; r < n; ++r) db += ref_alloc_bytes(dim, m, level[r]);
  ix->dump_bytes = db;
  ix->built = true;
  ix->build_stats.distcomps = job.distcomps;
  ix->build_stats.processed = n;
  ix->build_stats.kernel_ms = ms;
  *out = ix;
  return SHN_OK;
#undef CUB
}

int shn_index_build(shn_index** out, const float* base, const uint32_t* ids, uint64_t n, uint32_t dim, uint32_t m,
                    uint32_t ef_construction, shn_metric metric, uint32_t seed, int gpu_id) {
  if (!out || !base) return fail(SHN_ERR_ARG, "null argument");
  if (n == 0 || dim == 0) return fail(SHN_ERR_ARG, "empty base");
  int sms = 0;
  int rc = select_device(gpu_id, &sms);
  if (rc != SHN_OK) return rc;
  float* d_base = nullptr;
  uint32_t* d_ids = nullptr;
  CU(cudaMalloc(&d_base, n * dim * sizeof(float)));
  cudaError_t e = cudaMemcpy(d_base, base, n * dim * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && ids) {
    e = cudaMalloc(&d_ids, n * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemcpy(d_ids, ids, n * sizeof(uint32_t), cudaMemcpyHostToDevice);
  }
  if (e != cudaSuccess) { cudaFree(d_base); cudaFree(d_ids); return fail(SHN_ERR_CUDA, "staging the base: %s", cudaGetErrorString(e)); }
  rc = shn_index_build_device(out, d_base, d_ids, n, dim, m, ef_construction, metric, seed, gpu_id);
  cudaFree(d_base); cudaFree(d_ids);
  return rc;
}

int shn_draw_levels(uint64_t n, uint32_t m, uint32_t seed, uint32_t* levels) {
  if (!levels || m < 2) return fail(SHN_ERR_ARG, "need an output buffer and m >= 2");
  std::vector<uint32_t> l;
  draw_levels(n, m, seed, l);
  std::memcpy(levels, l.data(), n * sizeof(uint32_t));
  return SHN_OK;
}

int shn_set_build_option(const char* key, int64_t value) {
  if (!key) return fail(SHN_ERR_ARG, "null argument");
  if (std::strcmp(key, "batch_max") == 0) {
    if (value < 0 || value > (1 << 20)) return fail(SHN_ERR_ARG, "batch_max must be in [0, 2^20]");
    g_build_batch_max = static_cast<uint32_t>(value);
    return SHN_OK;
  }
  if (std::strcmp(key, "batch_div") == 0) {
    if (value < 0 || value > (1 << 20)) return fail(SHN_ERR_ARG, "batch_div must be in [0, 2^20]");
    g_build_batch_div = static_cast<uint32_t>(value);
    return SHN_OK;
  }
  return fail(SHN_ERR_ARG, "unknown build option '%s'", key);
}

int shn_index_build_stats(const shn_index* ix, shn_stats* out) {
  if (!ix || !out) return fail(SHN_ERR_ARG, "null argument");
  if (!ix->built) return fail(SHN_ERR_STATE, "the index was loaded, not built");
  *out = ix->build_stats;
  return SHN_OK;
}

int shn_index_store_mem(const shn_index* ix, void* const* dumps, uint64_t* sizes, int n_parts) {
  if (!ix || !sizes || n_parts < 1 || n_parts > 65535) return fail(SHN_ERR_ARG, "bad arguments");
  HostGraph g;
  int rc = download(ix, g);
  if (rc != SHN_OK) return rc;
  dump_sizes(g, n_parts, sizes);
  if (dumps) emit_dumps(g, n_parts, dumps);
  return SHN_OK;
}

int shn_index_store(const shn_index* ix, const char* const* dump_paths, int n_parts) {
  if (!ix || !dump_paths || n_parts < 1 || n_parts > 65535) return fail(SHN_ERR_ARG, "bad arguments");
  HostGraph g;
  int rc = download(ix, g);
  if (rc != SHN_OK) return rc;
  std::vector<uint64_t> sizes(n_parts);
  dump_sizes(g, n_parts, sizes.data());
  std::vector<std::vector<uint8_t>> bufs(n_parts);
  std::vector<void*> ptrs(n_parts);
  for (int i = 0; i < n_parts; ++i) { bufs[i].resize(sizes[i]); ptrs[i] = bufs[i].data(); }
  emit_dumps(g, n_parts, ptrs.data());
  for (int i = 0; i < n_parts; ++i) {
    FILE* f = std::fopen(dump_paths[i], "wb");
    if (!f) return fail(SHN_ERR_IO, "cannot create %s", dump_paths[i]);
    const size_t put = std::fwrite(bufs[i].data(), 1, bufs[i].size(), f);
    if (std::fclose(f) != 0 || put != bufs[i].size()) return fail(SHN_ERR_IO, "short write on %s", dump_paths[i]);
  }
  return SHN_OK;
}

int shn_dump_repartition(const void* const* dumps, const uint64_t* sizes, int n_parts_in, uint32_t dim, uint32_t m,
                         int n_parts_out, void* const* out_dumps, uint64_t* out_sizes) {
  if (!dumps || !sizes || !out_sizes || n_parts_in < 1 || n_parts_out < 1 || n_parts_out > 65535 || dim == 0 || m == 0 || m > 32)
    return fail(SHN_ERR_ARG, "bad arguments");
  HostGraph g;
  std::string err;
  if (!parse_dumps(dumps, sizes, n_parts_in, dim, m, g, err)) return fail(SHN_ERR_IO, "%s", err.c_str());
  dump_sizes(g, n_parts_out, out_sizes);
  if (out_dumps) emit_dumps(g, n_parts_out, out_dumps);
  return SHN_OK;
}

void shn_index_free(shn_index* ix) {
  if (!ix) return;
  cudaSetDevice(ix->gpu);
  if (ix->stream) cudaStreamSynchronize(ix->stream);
  if (ix->world > 1) {
    // a partition: d_vec / d_l0 are address ranges; unmap every piece, drop the allocations, give the ranges back
    const size_t row_bytes = static_cast<size_t>(ix->row_f4) * 16, list_bytes = 2ull * ix->m * sizeof(uint32_t);
    if (ix->hot_vec_blk.handle) { vmm_unplace(ix->vec_space, 0, ix->hot_vec_blk.size); vmm_unplace(ix->l0_space, 0, ix->hot_l0_blk.size); }
    if (ix->own_vec_blk.handle) {
      vmm_unplace(ix->vec_space, ix->part_begin[ix->rank] * row_bytes, ix->own_vec_blk.size);
      vmm_unplace(ix->l0_space, ix->part_begin[ix->rank] * list_bytes, ix->own_l0_blk.size);
    }
    for (uint32_t p = 0; p < 8; ++p) {
      if (ix->peer_placed[p]) {
        vmm_unplace(ix->vec_space, ix->part_begin[p] * row_bytes, ix->peer_vec_bytes[p]);
        vmm_unplace(ix->l0_space, ix->part_begin[p] * list_bytes, ix->peer_l0_bytes[p]);
      }
      vmm_drop(ix->peer_vec_blk[p]); vmm_drop(ix->peer_l0_blk[p]);
    }
    if (ix->halo_vec_blk.handle) {
      vmm_unplace(ix->vec_space, ix->n_ids * row_bytes, ix->halo_vec_blk.size);
      vmm_unplace(ix->l0_space, ix->n_ids * list_bytes, ix->halo_l0_blk.size);
    }
    vmm_drop(ix->hot_vec_blk); vmm_drop(ix->hot_l0_blk); vmm_drop(ix->own_vec_blk); vmm_drop(ix->own_l0_blk);
    vmm_drop(ix->halo_vec_blk); vmm_drop(ix->halo_l0_blk);
    vmm_release(ix->vec_space); vmm_release(ix->l0_space);
  } else {
    cudaFree(ix->d_vec); cudaFree(ix->d_l0);
  }
  cudaFree(ix->d_up_base); cudaFree(ix->d_up); cudaFree(ix->d_ext_id);
  cudaFree(ix->d_level);
  cudaFree(ix->d_visits);
  cudaFree(ix->d_halo_dir);
  cudaFree(ix->ws.counter); cudaFree(ix->ws.totals);
  ix->ovf.release(); ix->q_stage.release(); ix->dist_stage.release(); ix->id_stage.release();
  for (auto& e : ix->ev) if (e) cudaEventDestroy(e);
  for (auto& c : ix->ev_chunk) for (auto& e : c) if (e) cudaEventDestroy(e);
  if (ix->s_in) cudaStreamDestroy(ix->s_in);
  if (ix->s_out) cudaStreamDestroy(ix->s_out);
  if (ix->ev_last) cudaEventDestroy(ix->ev_last);
  if (ix->stream) cudaStreamDestroy(ix->stream);
  delete ix;
}

uint64_t shn_index_size(const shn_index* ix) { return ix ? ix->n : 0; }
uint32_t shn_index_dim(const shn_index* ix) { return ix ? ix->dim : 0; }
uint32_t shn_index_m(const shn_index* ix) { return ix ? ix->m : 0; }
uint32_t shn_index_max_level(const shn_index* ix) { return ix ? ix->max_level : 0; }
uint64_t shn_index_hbm_bytes(const shn_index* ix) { return ix ? ix->hbm_bytes : 0; }
int shn_index_partition_info(const shn_index* ix, uint32_t* hot, uint32_t* own, uint32_t* entry_row) {
  if (!ix) return fail(SHN_ERR_ARG, "null index");
  if (ix->world < 2) return fail(SHN_ERR_STATE, "the handle is not a partition");
  if (hot) *hot = ix->hot;
  if (own) *own = ix->own;
  if (entry_row) *entry_row = ix->ep_row;
  return SHN_OK;
}
uint64_t shn_index_dump_bytes(const shn_index* ix) { return ix ? ix->dump_bytes : 0; }

int shn_set_option(shn_index* ix, const char* key, int64_t value) {
  if (!ix || !key) return fail(SHN_ERR_ARG, "null argument");
  if (std::strcmp(key, "warps_per_sm") == 0) {
    if (value < 0 || value > 64) return fail(SHN_ERR_ARG, "warps_per_sm must be in [0, 64]");
    ix->warps_per_sm = static_cast<int>(value);
    return SHN_OK;
  }
  if (std::strcmp(key, "visited_smem_entries") == 0) {
    if (value < 0 || value > 32768) return fail(SHN_ERR_ARG, "visited_smem_entries must be in [0, 32768]");
    ix->vis_cap = static_cast<uint32_t>(value);
    return SHN_OK;
  }
  if (std::strcmp(key, "visited_compact") == 0) {
    if (value < 0 || value > 1) return fail(SHN_ERR_ARG, "visited_compact is 0 or 1");
    ix->vis_compact = static_cast<int>(value);
    return SHN_OK;
  }
  return fail(SHN_ERR_ARG, "unknown option '%s'", key);
}

int shn_search_device(shn_index* ix, const float* d_queries, uint64_t nq, uint32_t k, uint32_t ef, uint32_t* d_out_ids,
                      float* d_out_dists, uint32_t* d_per_query_counters, void* stream, shn_stats* stats) {
  int rc = check_search_args(ix, nq, k, ef);
  if (rc != SHN_OK) return rc;
  if (stats) std::memset(stats, 0, sizeof *stats);
  if (nq == 0) return SHN_OK;
  if (!d_queries || !d_out_ids) return fail(SHN_ERR_ARG, "null buffer");
  CU(cudaSetDevice(ix->gpu));
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : ix->stream;
  rc = run_search(ix, d_queries, nq, k, ef, d_out_ids, d_out_dists, d_per_query_counters, s, stats != nullptr, nullptr);
  if (rc != SHN_OK) return rc;
  if (stats) {
    unsigned long long t[kNumTotals];
    CU(cudaMemcpyAsync(t, ix->ws.totals, sizeof t, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, ix->ev[1], ix->ev[2]));
    fill_stats(ix, t, nq, stats);
    stats->kernel_ms = ms;
    if (t[kFailedQueries]) return fail(SHN_ERR_CAPACITY, "%llu queries overflowed the visited set (ef=%u)", t[kFailedQueries], ef);
  }
  return SHN_OK;
}

int shn_search(shn_index* ix, const float* queries, uint64_t nq, uint32_t k, uint32_t ef, uint32_t* out_ids,
               float* out_dists, shn_stats* stats) {
  int rc = check_search_args(ix, nq, k, ef);
  if (rc != SHN_OK) return rc;
  if (stats) std::memset(stats, 0, sizeof *stats);
  if (nq == 0) return SHN_OK;
  if (!queries || !out_ids) return fail(SHN_ERR_ARG, "null buffer");
  CU(cudaSetDevice(ix->gpu));
  if (ix->q_stage.ensure(nq * ix->dim) != cudaSuccess || ix->id_stage.ensure(nq * k) != cudaSuccess ||
      ix->dist_stage.ensure(nq * k) != cudaSuccess)
    return fail(SHN_ERR_CUDA, "cannot allocate staging buffers for %llu queries", static_cast<unsigned long long>(nq));
  // Large batches are cut into chunks: the queries of chunk c+1 travel to the device and the results of chunk c-1 travel
  // back (each on its own stream, i.e. its own copy engine) while chunk c is searched.  SHN_SEARCH_CHUNKS overrides the count.
  uint32_t chunks = nq >= (1u << 17) ? 4u : 1u;
  if (const char* env = std::getenv("SHN_SEARCH_CHUNKS")) chunks = static_cast<uint32_t>(std::max(1, std::min(8, std::atoi(env))));
  if (chunks > nq) chunks = 1;
  if (chunks > 1 && !ix->s_in) {
    CU(cudaStreamCreateWithFlags(&ix->s_in, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&ix->s_out, cudaStreamNonBlocking));
  }
  for (uint32_t c = 0; c < chunks; ++c) for (auto& e : ix->ev_chunk[c]) if (!e) CU(cudaEventCreate(&e));
  cudaStream_t s = ix->stream;
  cudaStream_t s_in = chunks > 1 ? ix->s_in : s, s_out = chunks > 1 ? ix->s_out : s;
  const uint64_t per = (nq + chunks - 1) / chunks;
  chunks = static_cast<uint32_t>((nq + per - 1) / per);  // chunks that actually hold queries (a forced count on a tiny batch)
  CU(cudaEventRecord(ix->ev[0], s));
  if (chunks > 1) CU(cudaStreamWaitEvent(s_in, ix->ev[0], 0));  // after whatever the handle's stream was doing with the staging buffers
  for (uint32_t c = 0; c < chunks; ++c) {
    const uint64_t q0 = c * per, cnt = std::min<uint64_t>(per, nq - q0);
    CU(cudaMemcpyAsync(ix->q_stage.p + q0 * ix->dim, queries + q0 * ix->dim, cnt * ix->dim * sizeof(float), cudaMemcpyHostToDevice, s_in));
    CU(cudaEventRecord(ix->ev_chunk[c][0], s_in));
  }
  for (uint32_t c = 0; c < chunks; ++c) {
    const uint64_t q0 = c * per, cnt = std::min<uint64_t>(per, nq - q0);
    if (chunks > 1) CU(cudaStreamWaitEvent(s, ix->ev_chunk[c][0], 0));
    CU(cudaEventRecord(ix->ev_chunk[c][1], s));
    ix->ws.keep_totals = c > 0;
    rc = run_search(ix, ix->q_stage.p + q0 * ix->dim, cnt, k, ef, ix->id_stage.p + q0 * k, ix->dist_stage.p + q0 * k, nullptr, s, false, nullptr);
    ix->ws.keep_totals = false;
    if (rc != SHN_OK) { cudaStreamSynchronize(s_in); cudaStreamSynchronize(s); cudaStreamSynchronize(s_out); return rc; }
    CU(cudaEventRecord(ix->ev_chunk[c][2], s));
    if (chunks > 1) CU(cudaStreamWaitEvent(s_out, ix->ev_chunk[c][2], 0));
    CU(cudaMemcpyAsync(out_ids + q0 * k, ix->id_stage.p + q0 * k, cnt * k * sizeof(uint32_t), cudaMemcpyDeviceToHost, s_out));
    if (out_dists) CU(cudaMemcpyAsync(out_dists + q0 * k, ix->dist_stage.p + q0 * k, cnt * k * sizeof(float), cudaMemcpyDeviceToHost, s_out));
  }
  unsigned long long t[kNumTotals];
  CU(cudaMemcpyAsync(t, ix->ws.totals, sizeof t, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  if (chunks > 1) {
    CU(cudaEventRecord(ix->ev[3], s_out));
    CU(cudaStreamSynchronize(s_out));
    CU(cudaStreamSynchronize(s_in));
  } else {
    CU(cudaEventRecord(ix->ev[3], s));
    CU(cudaStreamSynchronize(s));
  }
  if (stats) {
    // kernel_ms: the search launches; h2d_ms / d2h_ms: what the copies add in front of the first and behind the last launch
    float ker = 0.f, h2d = 0.f, d2h = 0.f;
    for (uint32_t c = 0; c < chunks; ++c) {
      float ms = 0.f;
      CU(cudaEventElapsedTime(&ms, ix->ev_chunk[c][1], ix->ev_chunk[c][2]));
      ker += ms;
    }
    CU(cudaEventElapsedTime(&h2d, ix->ev[0], ix->ev_chunk[0][1]));
    CU(cudaEventElapsedTime(&d2h, ix->ev_chunk[chunks - 1][2], ix->ev[3]));
    fill_stats(ix, t, nq, stats);
    stats->kernel_ms = ker; stats->h2d_ms = h2d; stats->d2h_ms = d2h;
  }
  if (t[kFailedQueries]) return fail(SHN_ERR_CAPACITY, "%llu queries overflowed the visited set (ef=%u)", t[kFailedQueries], ef);
  return SHN_OK;
}

int shn_index_count_visits(shn_index* ix, int enable) {
  if (!ix) return fail(SHN_ERR_ARG, "null index handle");
  if (ix->world > 1 && ix->d_halo_dir && enable) return fail(SHN_ERR_STATE, "this partition already has its halo");
  CU(cudaSetDevice(ix->gpu));
  CU(cudaStreamSynchronize(ix->stream));
  if (!enable) { cudaFree(ix->d_visits); ix->d_visits = nullptr; return SHN_OK; }
  const size_t ids = ix->world > 1 ? ix->n_ids : ix->n;
  if (!ix->d_visits) CU(cudaMalloc(&ix->d_visits, ids * sizeof(uint32_t)));
  CU(cudaMemsetAsync(ix->d_visits, 0, ids * sizeof(uint32_t), ix->stream));
  CU(cudaStreamSynchronize(ix->stream));
  return SHN_OK;
}

int shn_index_visit_counts(shn_index* ix, uint32_t* d_counts, int write_back) {
  if (!ix || !d_counts) return fail(SHN_ERR_ARG, "null argument");
  if (!ix->d_visits) return fail(SHN_ERR_STATE, "visit counting is off");
  if (ix->world > 1) return fail(SHN_ERR_STATE, "the counters of a partition are consumed by shn_index_partition_build_halo");
  CU(cudaSetDevice(ix->gpu));
  CU(cudaStreamSynchronize(ix->stream));
  // on the handle's stream and complete on return (a device-to-device cudaMemcpy on the legacy stream is neither ordered
  // with the caller's non-blocking streams nor finished when it returns)
  if (write_back) CU(cudaMemcpyAsync(ix->d_visits, d_counts, static_cast<size_t>(ix->n) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ix->stream));
  else CU(cudaMemcpyAsync(d_counts, ix->d_visits, static_cast<size_t>(ix->n) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ix->stream));
  CU(cudaStreamSynchronize(ix->stream));
  return SHN_OK;
}

int shn_index_partition_build_halo(shn_index* ix, uint32_t ratio_pct, uint64_t* halo_rows) {
  if (!ix) return fail(SHN_ERR_ARG, "null index");
  if (ix->world < 2) return fail(SHN_ERR_STATE, "the handle is not a partition");
  if (ix->attached != ix->world) return fail(SHN_ERR_STATE, "partitioned handle: %u of %u partitions attached", ix->attached, ix->world);
  if (ix->d_halo_dir) return fail(SHN_ERR_STATE, "this partition already has its halo");
  if (!ix->d_visits) return fail(SHN_ERR_STATE, "visit counting is off: the halo is chosen from the counts of a warm-up pass");
  if (ratio_pct > 100) return fail(SHN_ERR_ARG, "the halo budget is a percentage of the nodes");
  CU(cudaSetDevice(ix->gpu));
  CU(cudaStreamSynchronize(ix->stream));
  const uint32_t ids = ix->n_ids;
  std::vector<uint32_t> visits(ids);
  CU(cudaMemcpy(visits.data(), ix->d_visits, static_cast<size_t>(ids) * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  cudaFree(ix->d_visits); ix->d_visits = nullptr;
  // candidates: rows of the peers' shares that this GPU's warm-up queries read at least once
  std::vector<uint32_t> cand;
  for (uint32_t p = 0; p < ix->world; ++p) {
    if (p == ix->rank) continue;
    for (uint32_t r = ix->part_begin[p]; r < ix->part_begin[p] + ix->part_rows[p]; ++r) if (visits[r] > 0) cand.push_back(r);
  }
  const uint64_t want = std::min<uint64_t>(static_cast<uint64_t>(ix->n) * ratio_pct / 100, cand.size());
  auto hotter = [&](uint32_t a, uint32_t b) { return visits[a] != visits[b] ? visits[a] > visits[b] : a < b; };
  if (want < cand.size()) std::nth_element(cand.begin(), cand.begin() + want, cand.end(), hotter);
  cand.resize(want);
  std::sort(cand.begin(), cand.end());  // slots in row order: slot = prefix + popc(bits below)
  if (halo_rows) *halo_rows = want;
  if (want == 0) return SHN_OK;
  const size_t words = (static_cast<size_t>(ids) + 31) / 32;
  std::vector<uint2> dir(words, make_uint2(0u, 0u));
  for (uint32_t r : cand) dir[r >> 5].x |= 1u << (r & 31u);
  uint32_t run = 0;
  for (size_t w = 0; w < words; ++w) { dir[w].y = run; run += static_cast<uint32_t>(__builtin_popcount(dir[w].x)); }
  const size_t row_bytes = static_cast<size_t>(ix->row_f4) * 16, list_bytes = 2ull * ix->m * sizeof(uint32_t);
  const char* why = "";
  // the copies go behind the shares, in the same two address ranges: halo slot s is flat id n_ids + s
  if (vmm_create(ix->halo_vec_blk, want * row_bytes, ix->gpu, &why) != cudaSuccess ||
      vmm_create(ix->halo_l0_blk, want * list_bytes, ix->gpu, &why) != cudaSuccess ||
      vmm_place(ix->vec_space, ids * row_bytes, ix->halo_vec_blk.size, ix->halo_vec_blk.handle, ix->gpu, &why) != cudaSuccess) {
    vmm_drop(ix->halo_vec_blk); vmm_drop(ix->halo_l0_blk);
    return fail(SHN_ERR_CUDA, "allocating the halo (%llu rows): %s failed", static_cast<unsigned long long>(want), why);
  }
  if (vmm_place(ix->l0_space, ids * list_bytes, ix->halo_l0_blk.size, ix->halo_l0_blk.handle, ix->gpu, &why) != cudaSuccess) {
    vmm_unplace(ix->vec_space, ids * row_bytes, ix->halo_vec_blk.size);
    vmm_drop(ix->halo_vec_blk); vmm_drop(ix->halo_l0_blk);
    return fail(SHN_ERR_CUDA, "mapping the halo: %s failed", why);
  }
  uint32_t* d_rows = nullptr;
  uint2* d_dir = nullptr;
  cudaError_t e = cudaMalloc(&d_rows, want * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMalloc(&d_dir, words * sizeof(uint2));
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_rows, cand.data(), want * sizeof(uint32_t), cudaMemcpyHostToDevice, ix->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_dir, dir.data(), words * sizeof(uint2), cudaMemcpyHostToDevice, ix->stream);
  if (e == cudaSuccess) e = halo_gather(ix->view(), d_rows, static_cast<uint32_t>(want), ix->d_vec + static_cast<size_t>(ids) * ix->row_f4,
                                        ix->d_l0 + static_cast<size_t>(ids) * 2 * ix->m, ix->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ix->stream);
  cudaFree(d_rows);
  if (e != cudaSuccess) {
    cudaFree(d_dir);
    vmm_unplace(ix->vec_space, ids * row_bytes, ix->halo_vec_blk.size); vmm_unplace(ix->l0_space, ids * list_bytes, ix->halo_l0_blk.size);
    vmm_drop(ix->halo_vec_blk); vmm_drop(ix->halo_l0_blk);
    return fail(SHN_ERR_CUDA, "building the halo (%llu rows): %s", static_cast<unsigned long long>(want), cudaGetErrorString(e));
  }
  ix->d_halo_dir = d_dir; ix->halo = static_cast<uint32_t>(want);
  ix->hbm_bytes += ix->halo_vec_blk.size + ix->halo_l0_blk.size + words * sizeof(uint2);
  return SHN_OK;
}

int shn_index_partition(shn_index** out, const shn_index* full, int rank, int world, uint32_t cache_ratio_pct,
                        const uint8_t* d_owner) {
  if (!out || !full) return fail(SHN_ERR_ARG, "null argument");
  if (full->world > 1) return fail(SHN_ERR_STATE, "the handle is already a partition");
  if (world < 2 || world > 8 || rank < 0 || rank >= world) return fail(SHN_ERR_ARG, "need 2 <= world <= 8 and 0 <= rank < world");
  if (cache_ratio_pct > 100) return fail(SHN_ERR_ARG, "cache ratio is a percentage");
  CU(cudaSetDevice(full->gpu));
  CU(cudaStreamSynchronize(full->stream));
  const uint32_t n = full->n;
  // hot set: every node with level > 0, then the most visited level-0 nodes (ties: lower row) up to the cache budget
  std::vector<uint32_t> level(n), visits;
  CU(cudaMemcpy(level.data(), full->d_level, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  if (full->d_visits) {
    visits.resize(n);
    CU(cudaMemcpy(visits.data(), full->d_visits, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  }
  std::vector<uint8_t> is_hot(n, 0);
  uint64_t hot = 0;
  for (uint32_t r = 0; r < n; ++r) if (level[r] > 0) { is_hot[r] = 1; ++hot; }
  if (!is_hot[full->ep_row]) { is_hot[full->ep_row] = 1; ++hot; }  // a graph with one level: the entry point is read by every query
  const uint64_t budget = std::min<uint64_t>(n, std::max<uint64_t>(hot, static_cast<uint64_t>(n) * cache_ratio_pct / 100));
  if (hot < budget && !visits.empty()) {
    std::vector<uint32_t> cand;
    cand.reserve(n - hot);
    for (uint32_t r = 0; r < n; ++r) if (!is_hot[r] && visits[r] > 0) cand.push_back(r);
    const uint64_t want = std::min<uint64_t>(budget - hot, cand.size());
    auto hotter = [&](uint32_t a, uint32_t b) { return visits[a] != visits[b] ? visits[a] > visits[b] : a < b; };
    if (want < cand.size()) std::nth_element(cand.begin(), cand.begin() + want, cand.end(), hotter);
    for (uint64_t i = 0; i < want; ++i) is_hot[cand[i]] = 1;
    hot += want;
  }
  const uint32_t H = static_cast<uint32_t>(hot);
  const size_t row_bytes = static_cast<size_t>(full->row_f4) * 16, m0 = 2ull * full->m, list_bytes = m0 * sizeof(uint32_t);
  // flat numbering (graph.h): every piece starts on a multiple of `align` rows, so that it starts on an allocation granule in
  // both address ranges
  size_t gran = 0;
  {
    const char* why = "";
    if (vmm_granularity(full->gpu, &gran, &why) != cudaSuccess) return fail(SHN_ERR_CUDA, "%s failed", why);
  }
  auto gcd = [](size_t a, size_t b) { while (b) { const size_t t = a % b; a = b; b = t; } return a; };
  const size_t align = std::max(gran / gcd(gran, row_bytes), gran / gcd(gran, list_bytes));
  auto up = [&](uint64_t v) { return (v + align - 1) / align * align; };
  // owner of every cold row: the placement (k-means cluster), or round-robin as the reference scatters nodes over the
  // memory nodes (src/compute_thread.hh:57)
  std::vector<uint8_t> owner(n);
  if (d_owner) {
    CU(cudaMemcpy(owner.data(), d_owner, n, cudaMemcpyDeviceToHost));
  } else {
    uint32_t cold = 0;
    for (uint32_t r = 0; r < n; ++r) if (!is_hot[r]) owner[r] = static_cast<uint8_t>(cold++ % world);
  }
  uint64_t rows_of[8] = {0};
  for (uint32_t r = 0; r < n; ++r) {
    if (is_hot[r]) continue;
    if (owner[r] >= world) return fail(SHN_ERR_ARG, "owner[%u] = %u is not a rank below %d", r, owner[r], world);
    ++rows_of[owner[r]];
  }
  uint64_t begins[9];
  begins[0] = up(std::max<uint64_t>(H, 1));
  for (int g = 0; g < world; ++g) begins[g + 1] = begins[g] + up(std::max<uint64_t>(rows_of[g], 1));
  const uint64_t n_flat = begins[world];
  const uint64_t halo_cap = up(n);
  if (n_flat + halo_cap >= (1ull << 31)) return fail(SHN_ERR_ARG, "the flat id space of the partition exceeds 2^31 - 1");
  std::vector<uint32_t> new_of_old(n);
  {
    std::vector<uint32_t> cursor(world);
    for (int g = 0; g < world; ++g) cursor[g] = static_cast<uint32_t>(begins[g]);
    uint32_t next_hot = 0;
    for (uint32_t r = 0; r < n; ++r) new_of_old[r] = is_hot[r] ? next_hot++ : cursor[owner[r]]++;
  }
  const uint32_t own = static_cast<uint32_t>(rows_of[rank]);

  shn_index* ix = nullptr;
  int rc = new_handle(&ix, full->gpu, full->metric);
  if (rc != SHN_OK) return rc;
  auto bail = [&](int code) { shn_index_free(ix); return code; };
#define CUB(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return bail(fail(SHN_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__))); } while (0)
  ix->n = n; ix->dim = full->dim; ix->m = full->m; ix->row_f4 = full->row_f4; ix->n_up = full->n_up; ix->max_level = full->max_level;
  ix->max_level_of_ep = full->max_level_of_ep; ix->ep_row = new_of_old[full->ep_row];
  ix->hot = H; ix->world = world; ix->rank = rank; ix->own = own; ix->attached = 1;
  ix->n_ids = static_cast<uint32_t>(n_flat); ix->row_align = align;
  for (int i = 0; i < 9; ++i) ix->part_begin[i] = i <= world ? static_cast<uint32_t>(begins[i]) : kInvalid;
  for (int i = 0; i < 8; ++i) ix->part_rows[i] = i < world ? static_cast<uint32_t>(rows_of[i]) : 0u;
  ix->warps_per_sm = full->warps_per_sm; ix->vis_cap = full->vis_cap; ix->dump_bytes = full->dump_bytes;
  {
    const char* why = "";
    if (vmm_reserve(ix->vec_space, (n_flat + halo_cap) * row_bytes, gran, &why) != cudaSuccess ||
        vmm_reserve(ix->l0_space, (n_flat + halo_cap) * list_bytes, gran, &why) != cudaSuccess)
      return bail(fail(SHN_ERR_CUDA, "reserving the partition's address ranges: %s failed", why));
    ix->d_vec = static_cast<float4*>(ix->vec_space.base);
    ix->d_l0 = static_cast<uint32_t*>(ix->l0_space.base);
    if (vmm_create(ix->hot_vec_blk, begins[0] * row_bytes, ix->gpu, &why) != cudaSuccess ||
        vmm_create(ix->hot_l0_blk, begins[0] * list_bytes, ix->gpu, &why) != cudaSuccess ||
        vmm_create(ix->own_vec_blk, up(std::max<uint64_t>(own, 1)) * row_bytes, ix->gpu, &why) != cudaSuccess ||
        vmm_create(ix->own_l0_blk, up(std::max<uint64_t>(own, 1)) * list_bytes, ix->gpu, &why) != cudaSuccess)
      return bail(fail(SHN_ERR_CUDA, "allocating the hot set and this GPU's share: %s failed", why));
    if (vmm_place(ix->vec_space, 0, ix->hot_vec_blk.size, ix->hot_vec_blk.handle, ix->gpu, &why) != cudaSuccess ||
        vmm_place(ix->l0_space, 0, ix->hot_l0_blk.size, ix->hot_l0_blk.handle, ix->gpu, &why) != cudaSuccess ||
        vmm_place(ix->vec_space, begins[rank] * row_bytes, ix->own_vec_blk.size, ix->own_vec_blk.handle, ix->gpu, &why) != cudaSuccess ||
        vmm_place(ix->l0_space, begins[rank] * list_bytes, ix->own_l0_blk.size, ix->own_l0_blk.handle, ix->gpu, &why) != cudaSuccess)
      return bail(fail(SHN_ERR_CUDA, "mapping the hot set and this GPU's share: %s failed", why));
  }
  uint32_t *d_new_of_old = nullptr, *d_old_of_new = nullptr;
  CUB(cudaMalloc(&d_new_of_old, n * sizeof(uint32_t)));
  CUB(cudaMalloc(&d_old_of_new, n_flat * sizeof(uint32_t)));
  CUB(cudaMemcpy(d_new_of_old, new_of_old.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice));
  CUB(cudaMalloc(&ix->d_up_base, begins[0] * sizeof(uint32_t)));
  CUB(cudaMalloc(&ix->d_up, std::max<size_t>(ix->n_up, 1) * ix->m * sizeof(uint32_t)));
  CUB(cudaMalloc(&ix->d_ext_id, n_flat * sizeof(uint32_t)));
  // the pads read as empty rows with empty lists
  CUB(cudaMemsetAsync(ix->d_vec, 0, ix->hot_vec_blk.size, ix->stream));
  CUB(cudaMemsetAsync(ix->d_l0, 0xFF, ix->hot_l0_blk.size, ix->stream));
  CUB(cudaMemsetAsync(ix->d_vec + begins[rank] * ix->row_f4, 0, ix->own_vec_blk.size, ix->stream));
  CUB(cudaMemsetAsync(ix->d_l0 + begins[rank] * m0, 0xFF, ix->own_l0_blk.size, ix->stream));
  ix->hbm_bytes = ix->hot_vec_blk.size + ix->hot_l0_blk.size + ix->own_vec_blk.size + ix->own_l0_blk.size + begins[0] * 4ull +
                  n_flat * 4ull + std::max<size_t>(ix->n_up, 1) * ix->m * 4;
  PartitionJob job;
  job.n = n; job.n_flat = static_cast<uint32_t>(n_flat); job.hot = H; job.own_first = static_cast<uint32_t>(begins[rank]); job.own = own;
  job.row_f4 = ix->row_f4; job.m = ix->m; job.m0 = 2 * ix->m;
  job.n_up = ix->n_up; job.new_of_old = d_new_of_old; job.old_of_new = d_old_of_new;
  job.src_vec = full->d_vec; job.src_l0 = full->d_l0; job.src_up_base = full->d_up_base; job.src_up = full->d_up; job.src_ext_id = full->d_ext_id;
  job.hot_vec = ix->d_vec; job.own_vec = ix->d_vec + begins[rank] * ix->row_f4;
  job.hot_l0 = ix->d_l0; job.own_l0 = ix->d_l0 + begins[rank] * m0;
  job.hot_up_base = ix->d_up_base; job.up = ix->d_up; job.ext_id = ix->d_ext_id;
  cudaError_t e = partition_arrays(job, ix->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ix->stream);
  cudaFree(d_new_of_old); cudaFree(d_old_of_new);
  if (e != cudaSuccess) return bail(fail(SHN_ERR_CUDA, "partition_arrays: %s", cudaGetErrorString(e)));
#undef CUB
  *out = ix;
  return SHN_OK;
}

int shn_placement_fit(const shn_index* full, int world, uint32_t seed, double slack, float* centroids, uint8_t* d_owner,
                      uint64_t* part_sizes) {
  if (!full || !centroids || !d_owner) return fail(SHN_ERR_ARG, "null argument");
  if (full->world > 1) return fail(SHN_ERR_STATE, "placement is fitted on the full index");
  if (world < 2 || world > 8) return fail(SHN_ERR_ARG, "need 2 <= world <= 8");
  CU(cudaSetDevice(full->gpu));
  CU(cudaStreamSynchronize(full->stream));
  const uint32_t n = full->n, dimS = full->row_f4 * 4;
  // sample: the upper-level nodes (placement.hh:78-106 fetches the top levels), topped up with level-0 rows if scarce
  std::vector<uint32_t> level(n);
  CU(cudaMemcpy(level.data(), full->d_level, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  std::vector<uint32_t> rows;
  for (uint32_t r = 0; r < n; ++r) if (level[r] > 0) rows.push_back(r);
  if (rows.size() > 65536) {
    std::vector<uint32_t> thin;
    const double step = static_cast<double>(rows.size()) / 65536;
    for (uint32_t i = 0; i < 65536; ++i) thin.push_back(rows[static_cast<size_t>(i * step)]);
    rows.swap(thin);
  }
  if (rows.size() < 500) {
    const uint32_t want = std::min<uint32_t>(n, 4096);
    const uint32_t stride = std::max<uint32_t>(1, n / want);
    for (uint32_t r = 0; r < n && rows.size() < want + 500; r += stride) if (level[r] == 0) rows.push_back(r);
  }
  if (rows.size() < static_cast<size_t>(world)) return fail(SHN_ERR_ARG, "fewer nodes than partitions");
  std::vector<float> sample(rows.size() * dimS);
  {  // one gather kernel + one copy
    uint32_t* d_rows = nullptr;
    float4* d_sample = nullptr;
    cudaError_t ge = cudaMalloc(&d_rows, rows.size() * sizeof(uint32_t));
    if (ge == cudaSuccess) ge = cudaMalloc(&d_sample, sample.size() * sizeof(float));
    if (ge == cudaSuccess) ge = cudaMemcpyAsync(d_rows, rows.data(), rows.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, full->stream);
    if (ge == cudaSuccess) ge = gather_rows(full->d_vec, d_rows, static_cast<uint32_t>(rows.size()), full->row_f4, d_sample, full->stream);
    if (ge == cudaSuccess) ge = cudaMemcpyAsync(sample.data(), d_sample, sample.size() * sizeof(float), cudaMemcpyDeviceToHost, full->stream);
    if (ge == cudaSuccess) ge = cudaStreamSynchronize(full->stream);
    cudaFree(d_rows); cudaFree(d_sample);
    if (ge != cudaSuccess) return fail(SHN_ERR_CUDA, "fetching the k-means sample: %s", cudaGetErrorString(ge));
  }
  std::vector<float> cent;
  kmeans_host(sample, static_cast<uint32_t>(rows.size()), dimS, world, seed, full->metric == SHN_IP, cent);
  float* d_cent = nullptr;
  CU(cudaMalloc(&d_cent, cent.size() * sizeof(float)));
  CU(cudaMemcpy(d_cent, cent.data(), cent.size() * sizeof(float), cudaMemcpyHostToDevice));
  std::vector<uint32_t> sizes;
  cudaError_t e = balanced_assign(full->d_vec, n, full->row_f4, d_cent, world, full->metric == SHN_IP, slack, d_owner, sizes, full->stream);
  cudaFree(d_cent);
  if (e != cudaSuccess) return fail(SHN_ERR_CUDA, "balanced_assign: %s", cudaGetErrorString(e));
  for (int c = 0; c < world; ++c) {
    if (part_sizes) part_sizes[c] = sizes[c];
    for (uint32_t j = 0; j < full->dim; ++j) centroids[static_cast<size_t>(c) * full->dim + j] = cent[static_cast<size_t>(c) * dimS + row_pos(full->dim, j)];
  }
  return SHN_OK;
}

int shn_index_partition_export(const shn_index* ix, int* fds, uint64_t* sizes, uint64_t* raw_ptrs) {
  if (!ix || ix->world < 2) return fail(SHN_ERR_STATE, "not a partitioned handle");
  CU(cudaSetDevice(ix->gpu));
  if (fds) {
    const char* why = "";
    if (vmm_export_fd(ix->own_vec_blk, &fds[0], &why) != cudaSuccess || vmm_export_fd(ix->own_l0_blk, &fds[1], &why) != cudaSuccess)
      return fail(SHN_ERR_CUDA, "exporting the share: %s failed", why);
  }
  if (sizes) { sizes[0] = ix->own_vec_blk.size; sizes[1] = ix->own_l0_blk.size; }
  // inside one process the share travels as the handle that owns it (its allocations are mapped a second time, into the
  // attaching handle's address ranges)
  if (raw_ptrs) { raw_ptrs[0] = reinterpret_cast<uint64_t>(ix); raw_ptrs[1] = ~reinterpret_cast<uint64_t>(ix); }
  return SHN_OK;
}

int shn_index_partition_attach(shn_index* ix, int peer, const int* fds, const uint64_t* sizes, const uint64_t* raw_ptrs) {
  if (!ix || ix->world < 2) return fail(SHN_ERR_STATE, "not a partitioned handle");
  if (peer < 0 || peer >= static_cast<int>(ix->world) || peer == static_cast<int>(ix->rank)) return fail(SHN_ERR_ARG, "bad peer rank %d", peer);
  if (!raw_ptrs && !(fds && sizes)) return fail(SHN_ERR_ARG, "need file descriptors with sizes, or the in-process tokens of shn_index_partition_export");
  if (ix->peer_placed[peer]) return fail(SHN_ERR_STATE, "rank %d is already attached", peer);
  CU(cudaSetDevice(ix->gpu));
  const size_t row_bytes = static_cast<size_t>(ix->row_f4) * 16, list_bytes = 2ull * ix->m * sizeof(uint32_t);
  const size_t room_rows = ix->part_begin[peer + 1] - ix->part_begin[peer];
  unsigned long long hv = 0, hl = 0;
  size_t sv = 0, sl = 0;
  const char* why = "";
  if (raw_ptrs) {  // same process (tests, shn_group): the peer's handle itself
    const shn_index* other = reinterpret_cast<const shn_index*>(raw_ptrs[0]);
    if (!other || raw_ptrs[1] != ~raw_ptrs[0] || other->world != ix->world || static_cast<int>(other->rank) != peer || other->n_ids != ix->n_ids ||
        other->hot != ix->hot || other->part_begin[peer] != ix->part_begin[peer])
      return fail(SHN_ERR_ARG, "the token is not rank %d of this partitioned index", peer);
    hv = other->own_vec_blk.handle; hl = other->own_l0_blk.handle; sv = other->own_vec_blk.size; sl = other->own_l0_blk.size;
  } else {         // another process: import its physical allocations (loads will go over NVLink)
    if (vmm_import_handle(ix->peer_vec_blk[peer], fds[0], sizes[0], &why) != cudaSuccess ||
        vmm_import_handle(ix->peer_l0_blk[peer], fds[1], sizes[1], &why) != cudaSuccess)
      return fail(SHN_ERR_CUDA, "importing the share of rank %d: %s failed", peer, why);
    hv = ix->peer_vec_blk[peer].handle; hl = ix->peer_l0_blk[peer].handle; sv = sizes[0]; sl = sizes[1];
  }
  if (sv > room_rows * row_bytes || sl > room_rows * list_bytes) return fail(SHN_ERR_ARG, "the share of rank %d does not fit its place in the id space (another partitioning?)", peer);
  if (vmm_place(ix->vec_space, ix->part_begin[peer] * row_bytes, sv, hv, ix->gpu, &why) != cudaSuccess)
    return fail(SHN_ERR_CUDA, "mapping the share of rank %d: %s failed", peer, why);
  if (vmm_place(ix->l0_space, ix->part_begin[peer] * list_bytes, sl, hl, ix->gpu, &why) != cudaSuccess) {
    vmm_unplace(ix->vec_space, ix->part_begin[peer] * row_bytes, sv);
    return fail(SHN_ERR_CUDA, "mapping the share of rank %d: %s failed", peer, why);
  }
  ix->peer_placed[peer] = true; ix->peer_vec_bytes[peer] = sv; ix->peer_l0_bytes[peer] = sl;
  ++ix->attached;
  return SHN_OK;
}

// Test hook (not in include/shn.h): the builder's select_neighbors (= HNSW::select_heuristic, hnsw.hh:482-522) on one
// candidate set: rows of the index with their distances to some query, ascending.  Host buffers.
int shn_debug_select_neighbors(shn_index* ix, const uint32_t* cand_rows, const float* cand_dist, uint32_t n_cand, uint32_t m_target,
                               uint32_t* out_rows, uint32_t* out_n, uint64_t* out_distcomps) {
  if (!ix || !cand_rows || !cand_dist || !out_rows || !out_n || n_cand == 0 || n_cand > 4096 || m_target == 0 || m_target > 64)
    return fail(SHN_ERR_ARG, "bad arguments");
  if (ix->world > 1) return fail(SHN_ERR_STATE, "not on a partitioned handle");
  CU(cudaSetDevice(ix->gpu));
  uint32_t *d_rows = nullptr, *d_out = nullptr, *d_n = nullptr;
  float* d_dist = nullptr;
  unsigned long long* d_dc = nullptr;
  CU(cudaMalloc(&d_rows, n_cand * 4)); CU(cudaMalloc(&d_dist, n_cand * 4)); CU(cudaMalloc(&d_out, 64 * 4));
  CU(cudaMalloc(&d_n, 4)); CU(cudaMalloc(&d_dc, 8));
  CU(cudaMemcpy(d_rows, cand_rows, n_cand * 4, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(d_dist, cand_dist, n_cand * 4, cudaMemcpyHostToDevice));
  cudaError_t e = select_probe(ix->view(), ix->metric == SHN_IP, d_rows, d_dist, n_cand, m_target, d_out, d_n, d_dc, ix->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ix->stream);
  unsigned long long dc = 0;
  if (e == cudaSuccess) e = cudaMemcpy(out_n, d_n, 4, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(out_rows, d_out, 64 * 4, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(&dc, d_dc, 8, cudaMemcpyDeviceToHost);
  cudaFree(d_rows); cudaFree(d_dist); cudaFree(d_out); cudaFree(d_n); cudaFree(d_dc);
  if (e != cudaSuccess) return fail(SHN_ERR_CUDA, "select probe: %s", cudaGetErrorString(e));
  if (out_distcomps) *out_distcomps = dc;
  return SHN_OK;
}

// Diagnostic (not in include/shn.h): bandwidth of random whole-row reads from partition `part` as this GPU sees it.
double shn_debug_partition_gather_gbs(shn_index* ix, int part) {
  if (!ix || ix->world < 2 || part < 0 || part >= static_cast<int>(ix->world)) return -1.0;
  if (part != static_cast<int>(ix->rank) && !ix->peer_placed[part]) return -1.0;
  cudaSetDevice(ix->gpu);
  const uint32_t rows = ix->part_rows[part];
  double gbs = -1.0;
  if (rows == 0 || probe_gather(ix->d_vec + static_cast<size_t>(ix->part_begin[part]) * ix->row_f4, rows, ix->row_f4, &gbs, ix->stream) != cudaSuccess) return -1.0;
  return gbs;
}

int shn_bruteforce_topk_device(const float* d_base, uint64_t n, const float* d_queries, uint64_t nq, uint32_t dim,
                               shn_metric metric, uint32_t k, uint32_t* d_out_ids, float* d_out_dists, int gpu_id,
                               void* stream) {
  if (!d_base || !d_queries || !d_out_ids) return fail(SHN_ERR_ARG, "null buffer");
  if (n == 0 || dim == 0 || k == 0 || k > 256) return fail(SHN_ERR_ARG, "need n >= 1, dim >= 1, 1 <= k <= 256");
  if (n >= kInvalid || nq >= kInvalid) return fail(SHN_ERR_ARG, "n and nq must be below 2^32 - 1");
  int sms = 0;
  int rc = select_device(gpu_id, &sms);
  if (rc != SHN_OK) return rc;
  // SHN_BRUTEFORCE=simt forces the fp32-pipe kernel, =tc the tensor-core one; default: tensor cores when the shape allows
  const char* mode = std::getenv("SHN_BRUTEFORCE");
  const bool want_tc = !(mode && std::strcmp(mode, "simt") == 0) && bruteforce_tc_supported(dim, k);
  if (mode && std::strcmp(mode, "tc") == 0 && !want_tc) return fail(SHN_ERR_ARG, "tensor-core brute force needs k <= 16 and dim <= 4096");
  cudaError_t e = want_tc ? bruteforce_tc_launch(d_base, n, d_queries, static_cast<uint32_t>(nq), dim, metric == SHN_IP, k, d_out_ids,
                                                 d_out_dists, sms, static_cast<cudaStream_t>(stream))
                          : bruteforce_launch(d_base, n, d_queries, static_cast<uint32_t>(nq), dim, metric == SHN_IP, k, d_out_ids,
                                              d_out_dists, sms, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail(SHN_ERR_CUDA, "bruteforce: %s", cudaGetErrorString(e));
  return SHN_OK;
}

// Diagnostic (not in include/shn.h): queries of the last tensor-core brute-force launch that lacked the exactness certificate
// and were answered by the fp32 kernel.
unsigned long long shn_debug_bruteforce_fallbacks() { return bruteforce_tc_last_fallbacks(); }

int shn_bruteforce_topk(const float* base, uint64_t n, const float* queries, uint64_t nq, uint32_t dim, shn_metric metric,
                        uint32_t k, uint32_t* out_ids, float* out_dists, int gpu_id) {
  if (!base || !queries || !out_ids) return fail(SHN_ERR_ARG, "null buffer");
  int sms = 0;
  int rc = select_device(gpu_id, &sms);
  if (rc != SHN_OK) return rc;
  if (nq == 0) return SHN_OK;
  float *d_base = nullptr, *d_q = nullptr, *d_d = nullptr;
  uint32_t* d_i = nullptr;
  cudaError_t e = cudaMalloc(&d_base, n * dim * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&d_q, nq * dim * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&d_i, nq * k * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMalloc(&d_d, nq * k * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpy(d_base, base, n * dim * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_q, queries, nq * dim * sizeof(float), cudaMemcpyHostToDevice);
  rc = SHN_OK;
  if (e != cudaSuccess) rc = fail(SHN_ERR_CUDA, "staging: %s", cudaGetErrorString(e));
  if (rc == SHN_OK) rc = shn_bruteforce_topk_device(d_base, n, d_q, nq, dim, metric, k, d_i, d_d, gpu_id, nullptr);
  if (rc == SHN_OK) {
    e = cudaMemcpy(out_ids, d_i, nq * k * sizeof(uint32_t), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && out_dists) e = cudaMemcpy(out_dists, d_d, nq * k * sizeof(float), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = fail(SHN_ERR_CUDA, "copy back: %s", cudaGetErrorString(e));
  }
  cudaFree(d_base); cudaFree(d_q); cudaFree(d_i); cudaFree(d_d);
  return rc;
}

const char* shn_last_error(void) { return g_err.c_str(); }
const char* shn_version(void) { return "shn_b200 0.1 (sm_100a)"; }

}  // extern "C"
