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
int default_vis_compact();  // SHN_VIS_COMPACT=0/1 overrides the built-in default (tuning)
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
  // shn_search with host buffers, large batches: the batch is cut into chunks whose copies run on their own streams
  cudaStream_t s_in = nullptr, s_out = nullptr;
  cudaEvent_t ev_chunk[8][3] = {};  // per chunk: queries arrived | kernel started | kernel finished
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
  int vis_compact = shn::default_vis_compact();  // 16-bit keys in the shared visited table where the graph allows it
  shn_stats build_stats{};
  // partitioned handle (shn_index_partition; graph.h "flat numbering"): d_vec / d_l0 are the bases of two address ranges
  // into which the hot set, this GPU's share, the peers' shares and the halo are mapped; n_ids = size of the id space
  uint32_t hot = 0, world = 1, rank = 0, own = 0, attached = 1;
  uint32_t n_ids = 0;                    // flat ids in use (n for a full index)
  uint32_t part_begin[9] = {0};          // share p = flat ids [part_begin[p], part_begin[p] + part_rows[p])
  uint32_t part_rows[8] = {0};
  size_t row_align = 0;                  // pieces start at multiples of this many rows
  shn::VmmSpace vec_space, l0_space;
  shn::VmmBlock hot_vec_blk, hot_l0_blk;      // replicated hot set (this GPU's copy)
  shn::VmmBlock own_vec_blk, own_l0_blk;      // this GPU's share (exportable as POSIX fds)
  shn::VmmBlock peer_vec_blk[8], peer_l0_blk[8];  // shares of other processes imported here (handles of this process are mapped directly)
  bool peer_placed[8] = {false};
  size_t peer_vec_bytes[8] = {0}, peer_l0_bytes[8] = {0};
  shn::VmmBlock halo_vec_blk, halo_l0_blk;    // halo (graph.h): local copies of peer-owned rows, mapped behind the shares
  uint2* d_halo_dir = nullptr;
  uint32_t halo = 0;
  uint32_t* d_visits = nullptr;          // [n_ids] when visit counting is on
  bool built = false;

  shn::DeviceGraph view() const {
    shn::DeviceGraph g;
    g.vec = d_vec; g.l0 = d_l0; g.up_base = d_up_base; g.up = d_up; g.ext_id = d_ext_id;
    g.n = world > 1 ? n_ids : n; g.dim = dim; g.m = m; g.m0 = 2 * m; g.row_f4 = row_f4; g.ep_row = ep_row; g.ep_level = max_level_of_ep;
    g.hot = world > 1 ? hot : n; g.world = world; g.rank = rank;
    g.own_lo = part_begin[rank]; g.own_hi = part_begin[rank] + part_rows[rank];
    g.halo_first = n_ids; g.halo_dir = d_halo_dir;
    g.visit_count = d_visits;
    return g;
  }
  uint32_t max_level_of_ep = 0;
};

