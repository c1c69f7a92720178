/* include/shn.h — C ABI of libshn_b200.so: the B200-native HNSW search / construction engine that replaces
 * the compute-node hot path of SHINE (TianqiLiu-777/DM-HNSW-reference).
 *
 * The reference has no FFI; its hot path is reached through C++ templates inside the `shine` binary.  Each
 * entry point below names the reference interface it stands in for (paths relative to the reference tree).
 * A host program keeps the reference's process contract (CLI, dataset directory, index dumps, JSON on
 * stdout — see dm-hnsw-reference_b200/host/ and INTEGRATION.md) and calls this ABI at the two places where
 * the reference calls run_inserts (src/compute_node.cc:80) and run_queries (src/compute_node.cc:238).
 *
 * Conventions: plain pointers and sizes only; every int return is 0 on success, <0 on error, with a
 * thread-local message available from shn_last_error().  The caller owns every buffer it passes; the handle
 * owns all device memory and streams.  One handle may be used by one host thread at a time.  There is no CPU
 * fallback: every call fails with SHN_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef SHN_H
#define SHN_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct shn_index shn_index; /* opaque */

typedef enum { SHN_L2 = 0, SHN_IP = 1 } shn_metric; /* src/hnsw/distance.hh:153 L2Distance, :157 IPDistance */

enum {
  SHN_OK = 0,
  SHN_ERR_ARG = -1,      /* invalid argument (e.g. ef < k: src/hnsw/hnsw.hh:36) */
  SHN_ERR_IO = -2,       /* file missing / unreadable / not a dump for this dim and m */
  SHN_ERR_CUDA = -3,     /* CUDA error or no usable device */
  SHN_ERR_CAPACITY = -4, /* a per-query structure overflowed (visited set); results of that call are invalid */
  SHN_ERR_STATE = -5     /* operation not valid for this handle */
};

/* Counters with the meaning of statistics::ThreadStatistics (src/common/statistics.hh:148-176), summed over
 * the queries (or inserts) of one call, plus device timing. */
typedef struct {
  uint64_t distcomps;             /* hnsw.hh:272,286,376,459 */
  uint64_t visited_nodes;         /* levels > 0  (hnsw.hh:270,365) */
  uint64_t visited_nodes_l0;      /* level 0     (hnsw.hh:270,442) */
  uint64_t visited_neighborlists; /* hnsw.hh:359,438 */
  uint64_t lists_l0;              /* the level-0 share of visited_neighborlists */
  uint64_t lists_upper;           /* the upper-level share */
  uint64_t algorithmic_bytes;     /* 4*dim*distcomps + 4*(2m)*lists_l0 + 4*m*lists_upper   */
  uint64_t reference_layout_bytes;/* what the reference counts in rdma_reads_in_bytes for the same work (rdma_reads.hh:12,46) */
  uint64_t overflow_queries;      /* queries whose visited set spilled to the HBM table (still exact) */
  uint64_t processed;             /* statistics.hh processed */
  double kernel_ms;               /* CUDA-event time of the device work of this call */
  double h2d_ms, d2h_ms;          /* what the host<->device copies add in front of the first / behind the last kernel of this call
                                   * (shn_search cuts large batches into chunks whose copies overlap the search; 0 for *_device) */
  /* partitioned handles only: where the level-0 rows of this call were read from (cache.hits_total = rows_hot +
   * rows_local, cache.misses_total = rows_remote in the reference's JSON) */
  uint64_t rows_hot;              /* replicated hot set in local HBM (the compute-node cache of src/cache/cache.hh) */
  uint64_t rows_local;            /* this GPU's own partition */
  uint64_t rows_remote;           /* a peer's partition, over NVLink (what the reference READs over RDMA) */
  uint64_t rows_halo;             /* this GPU's halo: local copies of peer-owned rows (shn_index_partition_build_halo); also cache hits */
} shn_stats;

/* ---- index lifetime -------------------------------------------------------------------------------------- */

/* Load an index from reference-format dumps dump/index_m<M>_efc<efC>_node<i>_of<n>.dat, one per memory node
 * (written by src/memory_node.hh:187-195, path built in src/compute_node.cc:428-430; format: SURVEY App. B).
 * Replaces MemoryNode::store_or_load_index load branch (memory_node.hh:182) + the RDMA READ path: nodes of
 * memory node i become rows in HBM.  dim and m are not stored in the file (io/read_data.hh:35-40). */
int shn_index_load(shn_index** out, const char* const* dump_paths, int n_parts, uint32_t dim, uint32_t m,
                   shn_metric metric, int gpu_id);
/* Same, from dumps already in host memory. */
int shn_index_load_mem(shn_index** out, const void* const* dumps, const uint64_t* sizes, int n_parts, uint32_t dim,
                       uint32_t m, shn_metric metric, int gpu_id);

/* Build an index over base[n][dim] (host, row-major fp32; ids[i] = external id of row i, NULL = i) on the GPU.
 * Replaces ComputeNode::run_inserts -> hnsw::schedule<D,true> -> HNSW::insert (compute_node.cc:322,
 * hnsw/scheduler.hh:20, hnsw/hnsw.hh:40-251).  Levels are drawn with the reference's recipe
 * floor(-ln(U)/ln(m)) from mt19937(seed) (hnsw.hh:48,563).  The graph is not the reference's graph (insertion
 * is batched); it meets the same recall at equal m / ef_construction / ef. */
int shn_index_build(shn_index** out, const float* base, const uint32_t* ids, uint64_t n, uint32_t dim, uint32_t m,
                    uint32_t ef_construction, shn_metric metric, uint32_t seed, int gpu_id);
/* Same with base already in device memory on gpu_id (row stride = dim floats). */
int shn_index_build_device(shn_index** out, const float* d_base, const uint32_t* d_ids, uint64_t n, uint32_t dim,
                           uint32_t m, uint32_t ef_construction, shn_metric metric, uint32_t seed, int gpu_id);
/* Host-only: the level of every node as shn_index_build* will draw it — HNSW::insert's recipe (hnsw.hh:48,563-564:
 * floor(-ln(U)/ln(m)), U from uniform_real_distribution<double> over mt19937(seed)), the first node at level 0 (:61),
 * never more than one above the current top (:106).  Identical to a single-coroutine build of the reference. */
int shn_draw_levels(uint64_t n, uint32_t m, uint32_t seed, uint32_t* levels);
/* Construction knobs, process-wide, read when shn_index_build* starts: "batch_max" (nodes inserted per step, default
 * 16384) and "batch_div" (a step inserts at most 1/batch_div of the current graph, default 64); 0 = default. */
int shn_set_build_option(const char* key, int64_t value);
/* Build-time counters (distcomps, processed, kernel_ms) of a handle created by shn_index_build*. */
int shn_index_build_stats(const shn_index*, shn_stats* out);

/* Write the index as n_parts reference-format dumps (nodes dealt round-robin to parts when n_parts > 1).
 * Replaces MemoryNode::store_or_load_index store branch (memory_node.hh:187-195). */
int shn_index_store(const shn_index*, const char* const* dump_paths, int n_parts);
/* Size query + in-memory variant: sizes[i] receives the byte size of part i; if dumps != NULL, dumps[i] must
 * point to sizes[i] writable bytes. */
int shn_index_store_mem(const shn_index*, void* const* dumps, uint64_t* sizes, int n_parts);

/* Host-only utility (no GPU needed): re-partition reference-format dumps written for n_parts_in memory nodes into
 * n_parts_out parts (a dump is otherwise only loadable with the memory-node count it was written with,
 * compute_node.cc:428-430 "_of<n>").  Nodes keep their scan order and are dealt round-robin to the output parts.
 * Two-call protocol as shn_index_store_mem: out_dumps == NULL fills out_sizes only. */
int shn_dump_repartition(const void* const* dumps, const uint64_t* sizes, int n_parts_in, uint32_t dim, uint32_t m,
                         int n_parts_out, void* const* out_dumps, uint64_t* out_sizes);

void shn_index_free(shn_index*);

/* ---- introspection ----------------------------------------------------------------------------------------- */
uint64_t shn_index_size(const shn_index*);       /* number of nodes */
uint32_t shn_index_dim(const shn_index*);
uint32_t shn_index_m(const shn_index*);
uint32_t shn_index_max_level(const shn_index*);  /* build.max_level in the reference's JSON */
uint64_t shn_index_hbm_bytes(const shn_index*);  /* device memory held by the handle */
uint64_t shn_index_dump_bytes(const shn_index*); /* "index_size": bytes the reference would have allocated (rdma_atomics.hh:98) */
/* Options: "warps_per_sm" (0 = auto): cap on resident query warps per SM; "visited_smem_entries" (0 = auto): size of
 * the per-warp visited table in shared memory (the rest spills to HBM, still exact); "visited_compact" (default 1): 16-bit
 * keys in that table where the graph allows it (at most 2^24 ids, default table size) — twice the keys per byte, exact.  Distances are always summed in the
 * reference's own order (src/hnsw/distance.hh as compiled, see oracle/hnsw_oracle.c), so results are bit-identical
 * to the reference's except on exact distance ties. */
int shn_set_option(shn_index*, const char* key, int64_t value);

/* ---- multi-GPU: one global graph partitioned over the GPUs of a node ---------------------------------------------
 * The reference scatters nodes over memory nodes (RemotePtr = memory node + offset, src/remote_pointer.hh:7-22,
 * src/compute_thread.hh:57) and every hop is an RDMA READ (src/rdma/rdma_reads.hh) unless the compute-node cache holds
 * the node (src/cache/cache.hh, admission src/hnsw/hnsw.hh:447-448).  Here: one process per GPU; each builds or loads the
 * full index, (optionally) runs warm-up queries with visit counting, all-reduces the counts, and keeps
 *   - the hot set (all nodes with level > 0 + the most visited level-0 nodes, cache_ratio_pct % of the nodes) replicated,
 *   - its share of the remaining rows (round-robin, or the rows shn_placement_fit gave it);
 * the other shares are its peers' physical memory mapped into the same address range (CUDA virtual memory management):
 * a row is read at base + row * stride wherever it lives, over NVLink if it is a peer's.  Results do not depend on the
 * partitioning. */
int shn_index_count_visits(shn_index*, int enable);   /* on: allocate + zero the per-node counters; searches then count */
/* copy the counters to (write_back = 0) or from (write_back = 1) a device buffer of n u32 — the caller all-reduces.
 * Complete on return; with write_back the caller's writes to d_counts must have completed (the call waits for no stream
 * but the handle's own). */
int shn_index_visit_counts(shn_index*, uint32_t* d_counts, int write_back);
/* d_owner == NULL: cold rows dealt round-robin (the reference's uniform scatter).  d_owner = device array of n bytes,
 * owner[row] < world (from shn_placement_fit): rank p keeps exactly the cold rows it owns. */
int shn_index_partition(shn_index** out, const shn_index* full, int rank, int world, uint32_t cache_ratio_pct,
                        const uint8_t* d_owner);
/* Placement by cluster + query routing (the reference: Placement src/cache/placement.hh:22-106, balanced k-means
 * src/cache/kmeans.hh:24-377 with k = #compute nodes, router src/router/query_router.hh:280-387).  Fit: k-means
 * (k-means++ from mt19937(seed)) over the upper-level nodes, then every node goes to the GPU of its nearest centroid with
 * no GPU more than (1 + slack) * n / world nodes.  centroids: host [world][dim] out; d_owner: device [n] out. */
int shn_placement_fit(const shn_index* full, int world, uint32_t seed, double slack, float* centroids, uint8_t* d_owner,
                      uint64_t* part_sizes /*[world], may be NULL*/);
/* dest[q] = the rank query q should run on: its nearest centroid whose rank is still under (1 + slack) * nq / world
 * queries of this batch, in query order (query_router.hh:356-368).  d_queries: device [nq][dim]; dest: host [nq].
 * Computed on the GPU (the two routing kernels of csrc/router.cu); only the nq destination bytes travel to the host. */
int shn_route_queries(const float* centroids, int world, uint32_t dim, shn_metric metric, const float* d_queries, uint64_t nq,
                      double slack, uint8_t* dest, int gpu_id);
/* Halo: a second, per-GPU cache on top of the replicated hot set.  With query routing a GPU's queries stay near its own
 * cluster and what they still read from peers are the same border rows over and over (the reference's compute-node cache
 * fills with exactly those, src/cache/cache.hh:232-311 admission on miss).  Protocol: shn_index_count_visits(partition, 1),
 * run warm-up queries the way they will be served (routed), then this call: the ratio_pct % of the nodes that this GPU read
 * most often among the rows its peers own are copied over NVLink into local HBM together with their level-0 lists; later
 * searches read them locally (stats.rows_halo).  Results are unchanged.  halo_rows (may be NULL) receives the row count. */
int shn_index_partition_build_halo(shn_index*, uint32_t ratio_pct, uint64_t* halo_rows);
/* How the handle was cut: size of the replicated hot set, rows of this GPU's own share, the entry point in the partition's
 * numbering.  Every rank of one partitioned index must report the same hot and entry_row (the ranks address each other's
 * shares with their own numbering); any output may be NULL. */
int shn_index_partition_info(const shn_index*, uint32_t* hot, uint32_t* own, uint32_t* entry_row);
/* This GPU's share as two POSIX file descriptors (vectors, level-0 lists; CUDA virtual-memory-management export — the
 * caller passes them to the other processes over a Unix socket and closes them), their mapped sizes, and — for peers
 * inside the same process — two opaque tokens (valid while this handle lives).  Any of the three outputs may be NULL. */
int shn_index_partition_export(const shn_index*, int* fds /*2*/, uint64_t* sizes /*2*/, uint64_t* tokens /*2*/);
/* Attach rank `peer`'s share: fds + sizes received from another process, or the tokens of a handle in this process.  The
 * share is mapped into this handle's own address ranges at the place the flat numbering gives it, so that a row of a
 * peer is read like any other row (over NVLink).  Searching needs every peer attached. */
int shn_index_partition_attach(shn_index*, int peer, const int* fds, const uint64_t* sizes, const uint64_t* tokens);

/* ---- query routing between the GPUs of a partitioned index, fused with the exchange (csrc/router.cu) ---------------
 * The reference: QueryRouter (src/router/query_router.hh:280-387) sends a query to the compute node of its nearest
 * k-means centroid (src/cache/placement.hh:22-106) unless that node is over its per-batch limit (:356-368, limits
 * :106-151); queries are relayed with SEND/RECV through a memory node (:83-104,195-210).  Here every GPU owns one
 * exchange block (inbox + landing buffers) that all peers map; a routed batch is
 *     every rank: shn_router_scatter  -> barrier -> shn_router_search -> barrier -> results in shn_router_results
 * scatter routes on the GPU (no host loop, nothing copied to the host) and stores each query straight into the inbox
 * of its destination over NVLink; search writes each result row straight into the landing buffer of the query's home
 * GPU.  The barriers are the caller's (a stream-ordered collective between processes, events between the streams of
 * one process).  world == 1 handles work too (everything stays local).
 * centroids: host [world][dim] from shn_placement_fit.  slack: a rank takes at most (1 + slack) * nq / world + 1
 * queries of a batch.  max_batch / k_max size the exchange block and must be the same on every rank. */
typedef struct shn_router shn_router;
int shn_router_create(shn_router** out, shn_index* partition, const float* centroids, double slack, uint64_t max_batch,
                      uint32_t k_max);
void shn_router_free(shn_router*);
/* The exchange block as a POSIX fd (CUDA VMM export; the caller passes it to the other processes and closes it), its
 * mapped size, and the raw device pointer (for peers inside the same process).  Any output may be NULL. */
int shn_router_export(const shn_router*, int* fd, uint64_t* size, uint64_t* raw_ptr);
/* Attach rank `peer`'s block: raw_ptr != 0 (same process) or fd + size (received from another process). */
int shn_router_attach(shn_router*, int peer, int fd, uint64_t size, uint64_t raw_ptr);
/* Route d_queries[nq][dim] (device memory of this GPU) and deliver them.  Asynchronous on `stream` (NULL = the router's). */
int shn_router_scatter(shn_router*, const float* d_queries, uint64_t nq, void* stream);
/* Search what arrived (HNSW::knn per query, as shn_search_device) and deliver the results.  Asynchronous unless stats. */
int shn_router_search(shn_router*, uint32_t k, uint32_t ef, void* stream, shn_stats* stats);
/* This rank's landing buffers: row q = results of the q-th query it scattered (row stride = k of the search). */
int shn_router_results(const shn_router*, uint32_t** d_ids, float** d_dists);
/* Host copies of how many queries the last scatter sent to each rank / the inbox holds from each rank (synchronises). */
int shn_router_counts(shn_router*, uint32_t* sent /*[world]*/, uint32_t* received /*[world]*/, void* stream);
/* dest[q] of the last scatter (device pointer, one byte per query). */
int shn_router_destinations(const shn_router*, const uint8_t** d_dest);

/* ---- one process, several GPUs: the whole multi-GPU fan-out behind one handle (csrc/group.cu) ----------------------
 * Replaces what ComputeNode does around the hot path with >1 compute node: placement + warm-up (src/compute_node.cc:
 * 110-131), the round-robin query split (src/io/read_data.hh:58) and the routed query phase (:191-245).
 * full[g]: one full index per GPU, the SAME graph on every GPU (built with the same seed or loaded from the same dumps;
 * checked).  If the caller ran warm-up queries with shn_index_count_visits on them (the same queries on every GPU), the
 * most visited nodes join the replicated hot set.  The group takes over nothing: the caller frees full[g] afterwards.
 * placement_by_cluster: nodes stored on the GPU of their nearest k-means centroid (shn_placement_fit) instead of dealt
 * round-robin; routing (needs placement_by_cluster): queries run on the GPU of their nearest centroid (shn_router_*). */
typedef struct shn_group shn_group;
int shn_group_create(shn_group** out, shn_index* const* full, int n_gpus, uint32_t cache_ratio_pct, int placement_by_cluster,
                     int routing, double slack, uint64_t max_batch, uint32_t k_max, uint32_t seed);
void shn_group_free(shn_group*);
int shn_group_size(const shn_group*);
shn_index* shn_group_partition(shn_group*, int i); /* GPU i's partition handle (introspection; owned by the group) */
int shn_group_timings(const shn_group*, double* placement_kmeans_ms, double* placement_partition_ms);
/* HNSW::knn for nq host queries, dealt to the GPUs by query id % n_gpus; results in query order (as shn_search).
 * per_gpu (may be NULL): n_gpus shn_stats, `processed` = queries that GPU answered; routing_ms (may be NULL): the
 * slowest GPU's route + scatter time. */
int shn_group_search(shn_group*, const float* queries, uint64_t nq, uint32_t k, uint32_t ef, uint32_t* out_ids,
                     float* out_dists, shn_stats* per_gpu, double* routing_ms);
/* The warm-up pass of compute_node.cc:116-131 through the group (results discarded), served the way the queries will be
 * (routed or round-robin).  halo_ratio_pct > 0: the pass counts what every GPU reads from its peers and each GPU then
 * caches the halo_ratio_pct % most-read of those rows locally (shn_index_partition_build_halo); halo_rows (may be NULL)
 * receives the total over the GPUs.  Once per group. */
int shn_group_warmup(shn_group*, const float* queries, uint64_t nq, uint32_t k, uint32_t ef, uint32_t halo_ratio_pct,
                     uint64_t* halo_rows);

/* ---- search ------------------------------------------------------------------------------------------------ */

/* k-NN for nq queries: HNSW::knn (hnsw/hnsw.hh:253-307) for every slot that hnsw::schedule<D,false>
 * (hnsw/scheduler.hh:66-75) would dequeue.  queries/out_* are HOST buffers; the call returns when the results
 * are in host memory.  out_ids[q*k + i] = external id (node uid, node/node.hh:81) of the i-th nearest result,
 * ascending by distance (the reference returns the same set in heap-array order without distances,
 * hnsw.hh:300-303), padded with 0xFFFFFFFF / +inf when fewer than k nodes are reachable.
 * out_dists may be NULL.  stats may be NULL. */
int shn_search(shn_index*, const float* queries, uint64_t nq, uint32_t k, uint32_t ef, uint32_t* out_ids,
               float* out_dists, shn_stats* stats);

/* Same with every buffer resident in the HBM of the index's GPU; `stream` is a cudaStream_t (NULL = the
 * handle's own stream).  Asynchronous unless stats != NULL.  per_query_counters (device, may be NULL) receives 6
 * u32 per query: distcomps, visited_nodes, visited_nodes_l0, lists_l0, lists_upper, flags (bit 0: the visited set
 * spilled to HBM, still exact; bit 1: it overflowed — that row of the output is invalid).  With stats == NULL that flag
 * is the only report of SHN_ERR_CAPACITY.  One launch per handle is in flight at a time: a call on another stream first
 * waits (on the device) for the handle's previous launch. */
int shn_search_device(shn_index*, const float* d_queries, uint64_t nq, uint32_t k, uint32_t ef, uint32_t* d_out_ids,
                      float* d_out_dists, uint32_t* d_per_query_counters, void* stream, shn_stats* stats);

/* ---- ground truth ------------------------------------------------------------------------------------------ */

/* Exact top-k by brute force (the reference only reads ground truth produced offline, compute_node.cc:317,588).
 * Host buffers; ids are row numbers of base; distances are fp32 squared L2 or 1 - dot accumulated with one fma per
 * element in element order; ascending, ties broken by the lower id; k <= 256.  The *_device variant takes device
 * pointers on gpu_id and synchronises the stream before it returns. */
int shn_bruteforce_topk(const float* base, uint64_t n, const float* queries, uint64_t nq, uint32_t dim,
                        shn_metric metric, uint32_t k, uint32_t* out_ids, float* out_dists, int gpu_id);
int shn_bruteforce_topk_device(const float* d_base, uint64_t n, const float* d_queries, uint64_t nq, uint32_t dim,
                               shn_metric metric, uint32_t k, uint32_t* d_out_ids, float* d_out_dists, int gpu_id,
                               void* stream);

const char* shn_last_error(void);
const char* shn_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SHN_H */
