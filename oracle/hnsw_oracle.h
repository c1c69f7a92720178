/* oracle/hnsw_oracle.h — CPU restatement of the reference's search path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library; it is the
 * checker, never the product.  Every function cites the reference file:line it restates
 * (paths relative to /root/reference).  Parity status: PINNED — checked against the reference's own
 * code (oracle/_ref/libshine_ref.so, built from /root/reference unmodified) by tests/test_oracle_pin.py
 * and against the committed fixtures in tests/golden/ that were generated from it: knn (ids in heap-array order,
 * distance bits, counters), distances, select_heuristic (selected set, distcomps) and insert (the whole graph of a
 * single-coroutine build, lists in stored order, distcomps).
 */
#ifndef HNSW_ORACLE_H
#define HNSW_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_index orc_index;

typedef struct {
  uint64_t distcomps;             /* stats.distcomps           hnsw.hh:272,286,376,459 */
  uint64_t visited_nodes;         /* stats.visited_nodes       (levels > 0) hnsw.hh:270,365 */
  uint64_t visited_nodes_l0;      /* stats.visited_nodes_l0    hnsw.hh:270,442 */
  uint64_t lists_l0;              /* visited_neighborlists at level 0   hnsw.hh:438 */
  uint64_t lists_upper;           /* visited_neighborlists at levels>0  hnsw.hh:359 */
  uint64_t rdma_reads_in_bytes;   /* rdma_reads.hh:12,46 (the ep-ptr READ :75 is once per coroutine, not counted here) */
  uint64_t tie;                   /* 1 if some comparison in this query was decided on exactly equal distances */
} orc_counters;

/* src/hnsw/distance.hh:80-151 as compiled by g++ 13.3 -O2 -ffast-math -mavx2 -mfma (see hnsw_oracle.c) */
float orc_dist(const float* a, const float* b, uint32_t dim, int ip);

/* Parse n_parts index dumps (src/memory_node.hh:14-26,187-195).  The dump bytes are copied. */
orc_index* orc_load(const uint8_t* const* dumps, const uint64_t* sizes, uint32_t n_parts, uint32_t dim, uint32_t m);
void orc_free(orc_index*);

uint32_t orc_num_nodes(const orc_index*);
uint32_t orc_entry_row(const orc_index*);
uint32_t orc_max_level(const orc_index*);
/* Row-major exports in dump scan order (memory node 1 first, ascending byte offset): */
void orc_export(const orc_index*, uint32_t* uid, uint32_t* level, float* vectors /*[n][dim]*/,
                uint32_t* l0_cnt, uint32_t* l0_adj /*[n][2m], rows, 0xFFFFFFFF padded*/);
/* neighbours (rows) of `row` at `level`; returns count, or 0xFFFFFFFF if level > node level */
uint32_t orc_neighbors(const orc_index*, uint32_t row, uint32_t level, uint32_t* out_rows);

/* HNSW::knn (src/hnsw/hnsw.hh:253-307).  out_ids/out_dists hold k entries per query in the reference's
 * heap-array order, padded with 0xFFFFFFFF / +inf; out_counts[q] = number of valid entries. */
int orc_knn(const orc_index*, const float* queries, uint32_t nq, uint32_t k, uint32_t ef, int ip,
            uint32_t* out_ids, float* out_dists, uint32_t* out_counts, orc_counters* per_query /*[nq] or NULL*/,
            int track_ties, int num_threads);

/* HNSW::select_heuristic (src/hnsw/hnsw.hh:482-522) over candidates given as (uid, dist_to_query, vector):
 * returns the number selected; selected[] receives indices into the candidate arrays, in the reference's
 * post-selection array order (before make_heap). */
uint32_t orc_select_heuristic(const uint32_t* uids, const float* dists, const float* vectors /*[c][dim]*/,
                              uint32_t c, uint32_t dim, uint32_t m, int ip, uint32_t* selected,
                              uint64_t* distcomps);

/* HNSW::insert (src/hnsw/hnsw.hh:40-251) for ids 0..n-1 in order, one thread, one coroutine, one memory node — the
 * configuration in which the reference's build is deterministic.  *dump (free with orc_free_buffer) is what memory node 1
 * would write (src/memory_node.hh:187-195); lists are in the reference's stored order. */
int orc_build(const float* base, uint32_t n, uint32_t dim, uint32_t m, uint32_t efc, uint32_t seed, int ip, uint8_t** dump,
              uint64_t* dump_size, uint64_t* distcomps);
void orc_free_buffer(void* p);

#ifdef __cplusplus
}
#endif
#endif
