// handle.h — the index handle behind include/shn.h and the helpers shared by the files that implement the C ABI
// (capi.cu, router.cu).  Internal to libshn_b200.so.
#pragma once
#include <cstdint>
#include <string>

#include "../../include/shn.h"
#include "engine.h"
#include "vmm.h"

namespace shn {

// records the thread-local message returned by shn_last_error() and hands `code` back
int fail(int code, const char* fmt, ...);
// cudaSetDevice + "is this an sm_100 device" (there is no CPU path)
int select_device(int gpu_id, int* num_sms);

#define CU(call)                                                                                        \
  do {                                                                                                  \
    cudaError_t e__ = (call);                                                                           \
    if (e__ != cudaSuccess) return shn::fail(SHN_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__));  \
  } while (0)

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaError_t ensure(size_t want) {
    if (want <= n) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; n = 0;
    cudaError_t e = cudaMalloc(&p, want * sizeof(T));
    if (e == cudaSuccess) n = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

}  // namespace shn

struct shn_index;
namespace shn {
// shared by shn_search* (capi.cu) and shn_router_search (router.cu)
int check_search_args(const shn_index* ix, uint64_t nq, uint32_t k, uint32_t ef);
int run_search(shn_index* ix, const float* d_queries, uint64_t nq, uint32_t k, uint32_t ef, uint32_t* d_ids, float* d_dists,
               uint32_t* d_per_query, cudaStream_t stream, bool timed, const RoutedIo* io);
void fill_stats(const shn_index* ix, const unsigned long long* totals, uint64_t nq, shn_stats* s);
}  // namespace shn

struct shn_index {
  int gpu = 0;
  int num_sms = 0;
  shn_metric metric = SHN_L2;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_last = nullptr;  // recorded after every search launch: the next launch waits for it
  bool launched = false;

  // graph in HBM (graph.h)
  uint32_t n = 0, dim = 0, m = 0, row_f4 = 0, ep_row = shn::kInvalid, max_level = 0;
  uint64_t n_up = 0;
  float4* d_vec = nullptr;
  uint32_t* d_l0 = nullptr;
  uint32_t* d_up_base = nullptr;
  uint32_t* d_up = nullptr;
  uint32_t* d_ext_id = nullptr;
  uint32_t* d_level = nullptr;
  uint64_t hbm_bytes = 0, dump_bytes = 0;

  // scratch
  shn::SearchWorkspace ws;
  shn::DevBuf<uint32_t> ovf;
  shn::DevBuf<float> q_stage, dist_stage;
  shn::DevBuf<uint32_t> id_stage;
  int warps_per_sm = 0;
  uint32_t vis_cap = 0;
  shn_stats build_stats{};
  // partitioned handle (shn_index_partition): d_vec / d_l0 / d_up_base hold the replicated hot set
  uint32_t hot = 0, world = 1, rank = 0, own = 0, attached = 1, clustered = 0;
  uint32_t part_begin[9] = {0};
  float4* d_own_vec = nullptr;
  uint32_t* d_own_l0 = nullptr;
  const float4* part_vec[8] = {nullptr};   // the shares as this GPU addresses them (own, or peer-mapped); travel to the
  const uint32_t* part_l0[8] = {nullptr};  // kernels inside DeviceGraph, i.e. in the constant bank
  shn::VmmBlock own_vec_blk, own_l0_blk;      // this GPU's share (exportable as POSIX fds)
  shn::VmmBlock peer_vec_blk[8], peer_l0_blk[8];  // shares of other processes mapped here
  uint32_t* d_visits = nullptr;          // [n] when visit counting is on
  uint2* d_halo_dir = nullptr;           // halo (graph.h): directory, rows, lists; halo = number of rows
  float4* d_halo_vec = nullptr;
  uint32_t* d_halo_l0 = nullptr;
  uint32_t halo = 0;
  bool built = false;

  shn::DeviceGraph view() const {
    shn::DeviceGraph g;
    g.vec = d_vec; g.l0 = d_l0; g.up_base = d_up_base; g.up = d_up; g.ext_id = d_ext_id;
    g.n = n; g.dim = dim; g.m = m; g.m0 = 2 * m; g.row_f4 = row_f4; g.ep_row = ep_row; g.ep_level = max_level_of_ep;
    g.hot = world > 1 ? hot : n; g.world = world; g.rank = rank;
    for (int i = 0; i < 8; ++i) { g.part_vec[i] = part_vec[i]; g.part_l0[i] = part_l0[i]; }
    g.visit_count = d_visits;
    g.clustered = clustered;
    g.halo_dir = d_halo_dir; g.halo_vec = d_halo_vec; g.halo_l0 = d_halo_l0;
    for (int i = 0; i < 9; ++i) g.part_begin[i] = part_begin[i];
    return g;
  }
  uint32_t max_level_of_ep = 0;
};

