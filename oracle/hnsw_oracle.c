/* oracle/hnsw_oracle.c — CPU restatement of the reference's search path.  TEST INFRASTRUCTURE ONLY.
 *
 * Plain C, no SIMD, no fast-math: every rounding step is written out so that the result is the same on
 * any host.  Parity status: PINNED (see hnsw_oracle.h).  Citations are into /root/reference.
 *
 * Arithmetic note.  The reference compiles src/hnsw/distance.hh with -O2 -ffast-math -mavx2 (CMakeLists.txt:16),
 * which lets g++ contract and re-associate the intrinsics.  The order restated in orc_dist() is the one
 * g++ 13.3 emits for those sources with -march=x86-64-v3 (read off the disassembly of
 * oracle/_ref/libshine_ref.so, where both metrics exist exactly once and every call site calls them):
 *   L2  per 16 elements, lane j of 8:  s_j += fma(d0_j, d0_j, d1_j*d1_j),  d0 = a[0..7]-b[0..7], d1 = a[8..15]-b[8..15]
 *       horizontal:  x_j = s_j + s_{j+4};  r = (x_1 + x_3) + (x_0 + x_2);   tail: r = r + (a-b)*(a-b)  (mul, add)
 *   IP  per 16 elements:               s_j += fma(a_j, b_j, a_{8+j}*b_{8+j})
 *       horizontal:  r = ((1 - (s_4+s_5)) - ((s_0+s_1)+(s_2+s_3))) - (s_6+s_7);  tail: t += a*b (mul, add); r = r - t
 * tests/test_oracle_pin.py checks this bit-for-bit against the reference build.
 */
#include "hnsw_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

/* ------------------------------------------------------------------ distances */

static float dist_l2(const float* a, const float* b, uint32_t dim) {
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const uint32_t d16 = dim & ~15u; /* distance.hh:88 qty16 */
  for (uint32_t c = 0; c < d16; c += 16) {
    for (int j = 0; j < 8; ++j) {
      const float d0 = a[c + j] - b[c + j];
      const float d1 = a[c + 8 + j] - b[c + 8 + j];
      const float t = d1 * d1;
      s[j] = s[j] + fmaf(d0, d0, t);
    }
  }
  const float x0 = s[0] + s[4], x1 = s[1] + s[5], x2 = s[2] + s[6], x3 = s[3] + s[7];
  float r = (x1 + x3) + (x0 + x2);
  for (uint32_t i = d16; i < dim; ++i) { /* distance.hh:112-115 */
    const float d = a[i] - b[i];
    const float sq = d * d;
    r = r + sq;
  }
  return r;
}

static float dist_ip(const float* a, const float* b, uint32_t dim) {
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const uint32_t d16 = dim & ~15u; /* distance.hh:127 */
  for (uint32_t c = 0; c < d16; c += 16) {
    for (int j = 0; j < 8; ++j) {
      const float t = a[c + 8 + j] * b[c + 8 + j];
      s[j] = s[j] + fmaf(a[c + j], b[c + j], t);
    }
  }
  float r = 1.0f - (s[4] + s[5]);
  r = r - ((s[0] + s[1]) + (s[2] + s[3]));
  r = r - (s[6] + s[7]);
  if (dim & 15u) { /* distance.hh:134-139 */
    float t = 0.0f;
    for (uint32_t i = d16; i < dim; ++i) {
      const float p = a[i] * b[i];
      t = t + p;
    }
    r = r - t;
  }
  return r;
}

float orc_dist(const float* a, const float* b, uint32_t dim, int ip) {
  return ip ? dist_ip(a, b, dim) : dist_l2(a, b, dim);
}

/* ------------------------------------------------------------------ dump parsing */

struct orc_index {
  uint32_t dim, m, n, n_parts, ep_row, max_level;
  uint8_t** dump;       /* owned copies */
  uint64_t* dump_size;
  uint32_t* part_first; /* first row of each part, n_parts+1 */
  uint64_t* offset;     /* [n] byte offset of row within its part (ascending per part) */
  uint32_t* part_of;    /* [n] */
  uint32_t* uid;        /* [n] node/node.hh:81 */
  uint32_t* level;      /* [n] node/node.hh:82 */
  uint32_t* l0_cnt;     /* [n] */
  uint32_t* l0_adj;     /* [n][2m] rows */
  uint64_t* up_base;    /* [n] index into up_cnt/up_adj rows, valid if level>0 */
  uint32_t* up_cnt;     /* [n_up] */
  uint32_t* up_adj;     /* [n_up][m] rows */
};

static size_t node_header_bytes(uint32_t dim) { return 16 + 4 * (size_t)dim; }        /* node/node.hh:50 */
static size_t list0_bytes(uint32_t m) { return 4 + 8 * (size_t)(2 * m); }              /* node/node.hh:45 */
static size_t listu_bytes(uint32_t m) { return 4 + 8 * (size_t)m; }                    /* node/node.hh:46 */
static size_t node_alloc_bytes(uint32_t dim, uint32_t m, uint32_t level) {             /* rdma_atomics.hh:88-95 */
  size_t s = node_header_bytes(dim) + list0_bytes(m) + (size_t)level * listu_bytes(m);
  while (s % 8 != 0) s += 4;
  return s;
}

static uint64_t rd_u64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static uint32_t rd_u32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }

/* RemotePtr -> row: memory node in the top 16 bits, byte offset below (remote_pointer.hh:9-22) */
static uint32_t resolve(const orc_index* ix, uint64_t rptr) {
  const uint32_t mn = (uint32_t)(rptr >> 48);
  const uint64_t off = (rptr << 16) >> 16;
  if (mn >= ix->n_parts) return 0xFFFFFFFFu;
  uint32_t lo = ix->part_first[mn], hi = ix->part_first[mn + 1];
  while (lo < hi) {
    const uint32_t mid = lo + (hi - lo) / 2;
    if (ix->offset[mid] < off) lo = mid + 1; else hi = mid;
  }
  if (lo < ix->part_first[mn + 1] && ix->offset[lo] == off) return lo;
  return 0xFFFFFFFFu;
}

orc_index* orc_load(const uint8_t* const* dumps, const uint64_t* sizes, uint32_t n_parts, uint32_t dim, uint32_t m) {
  orc_index* ix = (orc_index*)calloc(1, sizeof(orc_index));
  ix->dim = dim; ix->m = m; ix->n_parts = n_parts;
  ix->dump = (uint8_t**)calloc(n_parts, sizeof(uint8_t*));
  ix->dump_size = (uint64_t*)calloc(n_parts, sizeof(uint64_t));
  ix->part_first = (uint32_t*)calloc(n_parts + 1, sizeof(uint32_t));
  /* pass 1: count nodes; file = u64 free_ptr, u64 ep_ptr, nodes from offset 16 (memory_node.hh:14-26,61) */
  uint64_t n = 0, n_up = 0;
  for (uint32_t p = 0; p < n_parts; ++p) {
    ix->dump[p] = (uint8_t*)malloc(sizes[p]);
    memcpy(ix->dump[p], dumps[p], sizes[p]);
    uint64_t end = rd_u64(ix->dump[p]); /* free_ptr == bytes in use */
    if (end > sizes[p]) end = sizes[p];
    ix->dump_size[p] = end;
    ix->part_first[p] = (uint32_t)n;
    uint64_t off = 16;
    while (off + node_header_bytes(dim) + list0_bytes(m) <= end) {
      const uint32_t lvl = rd_u32(ix->dump[p] + off + 12);
      if (lvl > 64) { orc_free(ix); return NULL; } /* not a dump for this dim/m */
      ++n; n_up += lvl;
      off += node_alloc_bytes(dim, m, lvl);
    }
  }
  ix->part_first[n_parts] = (uint32_t)n;
  ix->n = (uint32_t)n;
  ix->offset = (uint64_t*)malloc(n * sizeof(uint64_t));
  ix->part_of = (uint32_t*)malloc(n * sizeof(uint32_t));
  ix->uid = (uint32_t*)malloc(n * sizeof(uint32_t));
  ix->level = (uint32_t*)malloc(n * sizeof(uint32_t));
  ix->l0_cnt = (uint32_t*)malloc(n * sizeof(uint32_t));
  ix->l0_adj = (uint32_t*)malloc(n * 2 * m * sizeof(uint32_t));
  ix->up_base = (uint64_t*)malloc(n * sizeof(uint64_t));
  ix->up_cnt = (uint32_t*)malloc((n_up + 1) * sizeof(uint32_t));
  ix->up_adj = (uint32_t*)malloc((n_up + 1) * m * sizeof(uint32_t));
  /* pass 2: offsets, uid, level */
  uint64_t row = 0, up = 0;
  for (uint32_t p = 0; p < n_parts; ++p) {
    uint64_t off = 16;
    const uint64_t end = ix->dump_size[p];
    while (off + node_header_bytes(dim) + list0_bytes(m) <= end) {
      const uint8_t* nd = ix->dump[p] + off;
      ix->offset[row] = off; ix->part_of[row] = p;
      ix->uid[row] = rd_u32(nd + 8); ix->level[row] = rd_u32(nd + 12);
      ix->up_base[row] = up; up += ix->level[row];
      if (ix->level[row] > ix->max_level) ix->max_level = ix->level[row];
      off += node_alloc_bytes(dim, m, ix->level[row]);
      ++row;
    }
  }
  /* pass 3: adjacency (lists: u32 count then 8-byte RemotePtrs at +4, node/neighborlist.hh:27-38) */
  for (row = 0; row < n; ++row) {
    const uint8_t* nd = ix->dump[ix->part_of[row]] + ix->offset[row];
    const uint8_t* l0 = nd + node_header_bytes(dim);
    uint32_t c = rd_u32(l0);
    if (c > 2 * m) c = 2 * m;
    ix->l0_cnt[row] = c;
    for (uint32_t i = 0; i < 2 * m; ++i)
      ix->l0_adj[row * 2 * m + i] = i < c ? resolve(ix, rd_u64(l0 + 4 + 8 * (size_t)i)) : 0xFFFFFFFFu;
    for (uint32_t l = 1; l <= ix->level[row]; ++l) { /* node/node.cc:18-27 */
      const uint8_t* lu = l0 + list0_bytes(m) + (size_t)(l - 1) * listu_bytes(m);
      uint32_t cu = rd_u32(lu);
      if (cu > m) cu = m;
      const uint64_t u = ix->up_base[row] + (l - 1);
      ix->up_cnt[u] = cu;
      for (uint32_t i = 0; i < m; ++i)
        ix->up_adj[u * m + i] = i < cu ? resolve(ix, rd_u64(lu + 4 + 8 * (size_t)i)) : 0xFFFFFFFFu;
    }
  }
  /* entry point: RemotePtr at byte 8 of memory node 0 (rdma_reads.hh:78-86) */
  ix->ep_row = n ? resolve(ix, rd_u64(ix->dump[0] + 8)) : 0xFFFFFFFFu;
  return ix;
}

void orc_free(orc_index* ix) {
  if (!ix) return;
  if (ix->dump) for (uint32_t p = 0; p < ix->n_parts; ++p) free(ix->dump[p]);
  free(ix->dump); free(ix->dump_size); free(ix->part_first); free(ix->offset); free(ix->part_of);
  free(ix->uid); free(ix->level); free(ix->l0_cnt); free(ix->l0_adj); free(ix->up_base); free(ix->up_cnt);
  free(ix->up_adj); free(ix);
}

uint32_t orc_num_nodes(const orc_index* ix) { return ix->n; }
uint32_t orc_entry_row(const orc_index* ix) { return ix->ep_row; }
uint32_t orc_max_level(const orc_index* ix) { return ix->max_level; }

static const float* row_vec(const orc_index* ix, uint32_t row) {
  return (const float*)(ix->dump[ix->part_of[row]] + ix->offset[row] + 16);
}

void orc_export(const orc_index* ix, uint32_t* uid, uint32_t* level, float* vectors, uint32_t* l0_cnt, uint32_t* l0_adj) {
  for (uint32_t r = 0; r < ix->n; ++r) {
    if (uid) uid[r] = ix->uid[r];
    if (level) level[r] = ix->level[r];
    if (vectors) memcpy(vectors + (size_t)r * ix->dim, row_vec(ix, r), 4 * (size_t)ix->dim);
    if (l0_cnt) l0_cnt[r] = ix->l0_cnt[r];
  }
  if (l0_adj) memcpy(l0_adj, ix->l0_adj, (size_t)ix->n * 2 * ix->m * sizeof(uint32_t));
}

uint32_t orc_neighbors(const orc_index* ix, uint32_t row, uint32_t level, uint32_t* out) {
  if (row >= ix->n || level > ix->level[row]) return 0xFFFFFFFFu;
  if (level == 0) {
    memcpy(out, ix->l0_adj + (size_t)row * 2 * ix->m, ix->l0_cnt[row] * sizeof(uint32_t));
    return ix->l0_cnt[row];
  }
  const uint64_t u = ix->up_base[row] + (level - 1);
  memcpy(out, ix->up_adj + u * ix->m, ix->up_cnt[u] * sizeof(uint32_t));
  return ix->up_cnt[u];
}

/* ------------------------------------------------------------------ heaps (src/hnsw/heap.hh:24-63)
 * std::push_heap / std::pop_heap restated (libstdc++ bits/stl_heap.h __push_heap / __adjust_heap) so that the
 * array order — which is the order knn() reports ids in, hnsw.hh:300-303 — is reproduced, ties included. */

typedef struct { uint32_t row; float dist; } entry;
typedef struct { entry* a; size_t n, cap; int is_max; } heap;

static int hcomp(const heap* h, entry l, entry r) { /* heap.hh:16-22 */
  return h->is_max ? (l.dist < r.dist) : (l.dist > r.dist);
}
static void sift_up(heap* h, size_t hole, size_t top, entry v) {
  while (hole > top) {
    const size_t parent = (hole - 1) / 2;
    if (!hcomp(h, h->a[parent], v)) break;
    h->a[hole] = h->a[parent];
    hole = parent;
  }
  h->a[hole] = v;
}
static void heap_push(heap* h, entry v) { /* heap.hh:43-46 */
  if (h->n == h->cap) { h->cap = h->cap ? 2 * h->cap : 64; h->a = (entry*)realloc(h->a, h->cap * sizeof(entry)); }
  h->n++;
  sift_up(h, h->n - 1, 0, v);
}
static void heap_pop(heap* h) { /* heap.hh:48-51 */
  if (h->n > 1) {
    const entry v = h->a[h->n - 1];
    h->a[h->n - 1] = h->a[0];
    const size_t len = h->n - 1;
    size_t hole = 0, child = 0;
    while (child < (len - 1) / 2) {
      child = 2 * (child + 1);
      if (hcomp(h, h->a[child], h->a[child - 1])) child--;
      h->a[hole] = h->a[child];
      hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
      child = 2 * (child + 1);
      h->a[hole] = h->a[child - 1];
      hole = child - 1;
    }
    sift_up(h, hole, 0, v);
  }
  h->n--;
}
static void heap_push_k(heap* h, entry v, size_t k) { /* heap.hh:34-41 */
  if (h->n < k) heap_push(h, v);
  else if (hcomp(h, v, h->a[0])) { heap_pop(h); heap_push(h, v); }
}

/* ------------------------------------------------------------------ knn */

typedef struct { uint8_t* seen; uint32_t* touched; size_t n_touched, cap; } visited_set;

static void vs_insert(visited_set* v, uint32_t row) {
  v->seen[row] = 1;
  if (v->n_touched == v->cap) { v->cap = v->cap ? 2 * v->cap : 1024; v->touched = (uint32_t*)realloc(v->touched, v->cap * 4); }
  v->touched[v->n_touched++] = row;
}
static void vs_clear(visited_set* v) {
  for (size_t i = 0; i < v->n_touched; ++i) v->seen[v->touched[i]] = 0;
  v->n_touched = 0;
}

static int tie_with(const heap* h, float d) {
  for (size_t i = 0; i < h->n; ++i) if (h->a[i].dist == d) return 1;
  return 0;
}

static void knn_one(const orc_index* ix, const float* q, uint32_t k, uint32_t ef, int ip, heap* top, heap* next,
                    visited_set* vis, uint32_t* out_ids, float* out_dists, uint32_t* out_count, orc_counters* ct,
                    int track_ties) {
  const uint32_t dim = ix->dim, m = ix->m;
  const uint64_t node_bytes = node_header_bytes(dim);
  orc_counters c; memset(&c, 0, sizeof c);
  top->n = 0; next->n = 0;

  /* hnsw.hh:261-272: fetch the entry point, count it at its own top level, first distance */
  uint32_t cur = ix->ep_row;
  c.rdma_reads_in_bytes += node_bytes;
  if (ix->level[cur] > 0) c.visited_nodes++; else c.visited_nodes_l0++;
  float closest = orc_dist(q, row_vec(ix, cur), dim, ip);
  c.distcomps++;

  /* search_for_one<without_lock>, hnsw.hh:332-393: levels ep.level .. 1 */
  for (uint32_t level = ix->level[ix->ep_row]; level > 0; --level) {
    int changed;
    do {
      changed = 0;
      if (level > ix->level[cur]) break; /* cannot happen in a consistent dump */
      const uint64_t u = ix->up_base[cur] + (level - 1);
      c.lists_upper++; c.rdma_reads_in_bytes += listu_bytes(m);     /* hnsw.hh:358-359, rdma_reads.hh:41-46 */
      uint32_t best = cur;
      for (uint32_t i = 0; i < ix->up_cnt[u]; ++i) {
        const uint32_t nb = ix->up_adj[u * m + i];
        c.visited_nodes++;                                          /* hnsw.hh:365 (level > 0) */
        c.rdma_reads_in_bytes += node_bytes;
        const float d = orc_dist(q, row_vec(ix, nb), dim, ip);
        c.distcomps++;                                              /* hnsw.hh:375-376 */
        if (track_ties && d == closest) c.tie = 1;
        if (d < closest) { closest = d; best = nb; changed = 1; }   /* hnsw.hh:378-382 */
      }
      cur = best;
    } while (changed);
  }

  /* hnsw.hh:285-286: the distance of the chosen node is computed once more */
  entry e0 = {cur, orc_dist(q, row_vec(ix, cur), dim, ip)};
  c.distcomps++;
  heap_push(top, e0);

  /* search_level<without_lock>(ef, level 0), hnsw.hh:407-476 */
  for (size_t i = 0; i < top->n; ++i) { heap_push(next, top->a[i]); vs_insert(vis, top->a[i].row); }
  while (next->n > 0) {
    const entry cand = next->a[0];
    heap_pop(next);
    float far = top->a[0].dist;
    if (cand.dist > far) break;                                      /* hnsw.hh:424 */
    c.lists_l0++; c.rdma_reads_in_bytes += list0_bytes(m);           /* hnsw.hh:437-438 */
    const uint32_t* adj = ix->l0_adj + (size_t)cand.row * 2 * m;
    for (uint32_t i = 0; i < ix->l0_cnt[cand.row]; ++i) {
      const uint32_t nb = adj[i];
      if (vis->seen[nb]) continue;                                   /* hnsw.hh:441 */
      c.visited_nodes_l0++;
      vs_insert(vis, nb);                                            /* marked before the distance, :443 */
      c.rdma_reads_in_bytes += node_bytes;
      far = top->a[0].dist;                                          /* re-read every time, :456 */
      const float d = orc_dist(q, row_vec(ix, nb), dim, ip);
      c.distcomps++;
      if (track_ties && (tie_with(top, d) || tie_with(next, d))) c.tie = 1;
      if (d < far || top->n < ef) {                                  /* hnsw.hh:461 */
        const entry e = {nb, d};
        heap_push(next, e);
        heap_push_k(top, e, ef);
      }
    }
  }
  vs_clear(vis);

  while (top->n > k) heap_pop(top);                                  /* hnsw.hh:296-298 */
  for (uint32_t i = 0; i < k; ++i) {                                 /* hnsw.hh:300-303: heap-array order */
    out_ids[i] = i < top->n ? ix->uid[top->a[i].row] : 0xFFFFFFFFu;
    out_dists[i] = i < top->n ? top->a[i].dist : INFINITY;
  }
  *out_count = (uint32_t)top->n;
  if (ct) *ct = c;
}

typedef struct {
  const orc_index* ix; const float* queries; uint32_t nq, k, ef; int ip, track_ties;
  uint32_t* out_ids; float* out_dists; uint32_t* out_counts; orc_counters* per_query;
  volatile int64_t* next; /* shared cursor, 16 queries per grab */
} knn_job;

static void* knn_worker(void* arg) {
  knn_job* j = (knn_job*)arg;
  const orc_index* ix = j->ix;
  heap top = {0, 0, 0, 1}, next = {0, 0, 0, 0};
  visited_set vis = {(uint8_t*)calloc(ix->n, 1), 0, 0, 0};
  for (;;) {
    const int64_t begin = __atomic_fetch_add(j->next, 16, __ATOMIC_RELAXED);
    if (begin >= (int64_t)j->nq) break;
    const int64_t end = begin + 16 < (int64_t)j->nq ? begin + 16 : (int64_t)j->nq;
    for (int64_t qi = begin; qi < end; ++qi) {
      uint32_t cnt;
      knn_one(ix, j->queries + (size_t)qi * ix->dim, j->k, j->ef, j->ip, &top, &next, &vis,
              j->out_ids + (size_t)qi * j->k, j->out_dists + (size_t)qi * j->k, &cnt,
              j->per_query ? &j->per_query[qi] : NULL, j->track_ties);
      if (j->out_counts) j->out_counts[qi] = cnt;
    }
  }
  free(top.a); free(next.a); free(vis.seen); free(vis.touched);
  return NULL;
}

int orc_knn(const orc_index* ix, const float* queries, uint32_t nq, uint32_t k, uint32_t ef, int ip,
            uint32_t* out_ids, float* out_dists, uint32_t* out_counts, orc_counters* per_query, int track_ties,
            int num_threads) {
  if (!ix || ix->n == 0 || ix->ep_row == 0xFFFFFFFFu || ef < k) return -1; /* hnsw.hh:36 */
  if (num_threads < 1) num_threads = 1;
  if (num_threads > 256) num_threads = 256;
  volatile int64_t cursor = 0;
  knn_job job = {ix, queries, nq, k, ef, ip, track_ties, out_ids, out_dists, out_counts, per_query, &cursor};
  pthread_t th[256];
  for (int t = 1; t < num_threads; ++t) pthread_create(&th[t], NULL, knn_worker, &job);
  knn_worker(&job);
  for (int t = 1; t < num_threads; ++t) pthread_join(th[t], NULL);
  return 0;
}

/* ------------------------------------------------------------------ select_heuristic (hnsw.hh:482-522) */

typedef struct { uint32_t idx, uid; float dist; } sel_entry;
static int sel_cmp(const void* a, const void* b) { /* heap.hh:53-57: ascending distance, ties by id */
  const sel_entry* l = (const sel_entry*)a; const sel_entry* r = (const sel_entry*)b;
  if (l->dist == r->dist) return l->uid < r->uid ? -1 : (l->uid > r->uid ? 1 : 0);
  return l->dist < r->dist ? -1 : 1;
}

uint32_t orc_select_heuristic(const uint32_t* uids, const float* dists, const float* vectors, uint32_t c,
                              uint32_t dim, uint32_t m, int ip, uint32_t* selected, uint64_t* distcomps) {
  uint64_t dc = 0;
  if (c < m) { /* hnsw.hh:483: fewer than m candidates are all kept, untouched */
    for (uint32_t i = 0; i < c; ++i) selected[i] = i;
    if (distcomps) *distcomps = 0;
    return c;
  }
  sel_entry* h = (sel_entry*)malloc((c ? c : 1) * sizeof(sel_entry));
  for (uint32_t i = 0; i < c; ++i) { h[i].idx = i; h[i].uid = uids[i]; h[i].dist = dists[i]; }
  qsort(h, c, sizeof(sel_entry), sel_cmp);
  uint32_t n_sel = c ? 1 : 0, consumed = 1;
  while (n_sel < m && consumed < c) {
    int keep = 1;
    const sel_entry cand = h[consumed];
    for (uint32_t i = 0; i < n_sel; ++i) {
      const float d = orc_dist(vectors + (size_t)h[i].idx * dim, vectors + (size_t)cand.idx * dim, dim, ip);
      ++dc;
      if (d < cand.dist) { keep = 0; break; } /* hnsw.hh:506 */
    }
    if (keep) { const sel_entry t = h[n_sel]; h[n_sel] = h[consumed]; h[consumed] = t; ++n_sel; } /* :513 */
    ++consumed;
  }
  for (uint32_t i = 0; i < n_sel; ++i) selected[i] = h[i].idx;
  free(h);
  if (distcomps) *distcomps = dc;
  return n_sel;
}

/* ------------------------------------------------------------------ construction (HNSW::insert, hnsw.hh:40-251)
 * Single thread, single coroutine, one memory node: the configuration in which the reference's build is deterministic
 * (SURVEY App. D).  The graph lists come out in the reference's stored order, which search parity depends on:
 * lists are written in heap-ARRAY order after select_heuristic's make_heap (hnsw.hh:165-175,521), so libstdc++'s
 * __make_heap / __adjust_heap are restated as well. */

/* std::mt19937 + std::uniform_real_distribution<double>(0,1) as libstdc++ implements them (generate_canonical<double,53>
 * draws two 32-bit words: (lo + hi * 2^32) / 2^64) */
typedef struct { uint32_t mt[624]; int idx; } mt19937;
static void mt_seed(mt19937* g, uint32_t seed) {
  g->mt[0] = seed;
  for (int i = 1; i < 624; ++i) g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
  g->idx = 624;
}
static uint32_t mt_next(mt19937* g) {
  if (g->idx >= 624) {
    for (int i = 0; i < 624; ++i) {
      const uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
      g->mt[i] = g->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    g->idx = 0;
  }
  uint32_t y = g->mt[g->idx++];
  y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
  return y;
}
static double mt_canonical(mt19937* g) {
  const double lo = (double)mt_next(g), hi = (double)mt_next(g);
  double r = (lo + hi * 4294967296.0) / 18446744073709551616.0;
  if (r >= 1.0) r = nextafter(1.0, 0.0);
  return r;
}

static void adjust_heap(heap* h, size_t hole, size_t len, entry v) { /* bits/stl_heap.h __adjust_heap */
  const size_t top = hole;
  size_t child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (hcomp(h, h->a[child], h->a[child - 1])) child--;
    h->a[hole] = h->a[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    h->a[hole] = h->a[child - 1];
    hole = child - 1;
  }
  sift_up(h, hole, top, v);
}
static void heap_make(heap* h) { /* std::make_heap */
  const size_t len = h->n;
  if (len < 2) return;
  size_t parent = (len - 2) / 2;
  for (;;) {
    const entry v = h->a[parent];
    adjust_heap(h, parent, len, v);
    if (parent == 0) return;
    parent--;
  }
}

typedef struct {
  uint32_t n, dim, m, m0;
  const float* base;
  uint32_t* level;
  uint32_t* l0_cnt; uint32_t* l0;          /* [n][m0] */
  uint64_t* up_base; uint32_t* up_cnt; uint32_t* up; uint64_t n_up, up_cap; /* lists of levels 1..level, m wide */
  uint32_t ep; int have_ep;
  uint64_t distcomps;
  int ip;
} builder;

static uint32_t* b_list(builder* b, uint32_t row, uint32_t lvl, uint32_t** cnt) {
  if (lvl == 0) { *cnt = &b->l0_cnt[row]; return b->l0 + (size_t)row * b->m0; }
  const uint64_t u = b->up_base[row] + (lvl - 1);
  *cnt = &b->up_cnt[u];
  return b->up + u * b->m;
}
static float b_dist(builder* b, const float* q, uint32_t row) { b->distcomps++; return orc_dist(q, b->base + (size_t)row * b->dim, b->dim, b->ip); }

static int entry_cmp(const void* x, const void* y) { /* heap.hh:53-57 */
  const entry* l = (const entry*)x; const entry* r = (const entry*)y;
  if (l->dist == r->dist) return l->row < r->row ? -1 : (l->row > r->row ? 1 : 0);
  return l->dist < r->dist ? -1 : 1;
}

/* select_heuristic on a max-heap of (row, distance to the query), hnsw.hh:482-522 */
static void b_select(builder* b, heap* top, uint32_t m) {
  if (top->n < m) return;
  qsort(top->a, top->n, sizeof(entry), entry_cmp);
  const size_t initial = top->n;
  size_t selected = 1, consumed = 1;
  while (selected < m && consumed < initial) {
    int keep = 1;
    const entry c = top->a[consumed];
    for (size_t i = 0; i < selected; ++i) {
      const float d = b_dist(b, b->base + (size_t)top->a[i].row * b->dim, c.row);
      if (d < c.dist) { keep = 0; break; }
    }
    if (keep) { const entry t = top->a[selected]; top->a[selected] = top->a[consumed]; top->a[consumed] = t; ++selected; }
    ++consumed;
  }
  top->n = selected;
  heap_make(top);
}

/* search_level<with_lock>, hnsw.hh:407-476, on `lvl` */
static void b_search_level(builder* b, const float* q, uint32_t ef, uint32_t lvl, heap* top, heap* next, visited_set* vis) {
  for (size_t i = 0; i < top->n; ++i) { heap_push(next, top->a[i]); vs_insert(vis, top->a[i].row); }
  while (next->n > 0) {
    const entry cand = next->a[0];
    heap_pop(next);
    float far = top->a[0].dist;
    if (cand.dist > far) break;
    uint32_t* cnt;
    const uint32_t* list = b_list(b, cand.row, lvl, &cnt);
    for (uint32_t i = 0; i < *cnt; ++i) {
      const uint32_t nb = list[i];
      if (vis->seen[nb]) continue;
      vs_insert(vis, nb);
      far = top->a[0].dist;
      const float d = b_dist(b, q, nb);
      if (d < far || top->n < ef) { const entry e = {nb, d}; heap_push(next, e); heap_push_k(top, e, ef); }
    }
  }
  next->n = 0;
  vs_clear(vis);
}

static void b_insert(builder* b, uint32_t id, uint32_t drawn_level, uint32_t efc, heap* top, heap* next, heap* tmp, visited_set* vis) {
  const float* q = b->base + (size_t)id * b->dim;
  const uint32_t m = b->m;
  if (!b->have_ep) { /* hnsw.hh:56-85: the node that initialises the index is written at level 0 */
    b->level[id] = 0; b->ep = id; b->have_ep = 1;
    return;
  }
  uint32_t lvl = drawn_level;
  const uint32_t top_level = b->level[b->ep];
  const int is_new_level = lvl > top_level;
  if (is_new_level) lvl = top_level + 1;               /* :106 */
  b->level[id] = lvl;
  if (lvl > 0) {                                        /* allocate the upper lists of the new node */
    if (b->n_up + lvl > b->up_cap) {
      b->up_cap = (b->n_up + lvl) * 2 + 64;
      b->up_cnt = (uint32_t*)realloc(b->up_cnt, b->up_cap * sizeof(uint32_t));
      b->up = (uint32_t*)realloc(b->up, b->up_cap * m * sizeof(uint32_t));
    }
    b->up_base[id] = b->n_up;
    for (uint32_t l = 0; l < lvl; ++l) b->up_cnt[b->n_up + l] = 0;
    b->n_up += lvl;
  }
  const float ep_dist = b_dist(b, q, b->ep);            /* :125-126 */
  top->n = 0;
  if (lvl < top_level) {                                /* search_for_one, :129-143 / :332-393 */
    uint32_t cur = b->ep;
    float closest = ep_dist;
    for (uint32_t l = top_level; l > lvl; --l) {
      int changed;
      do {
        changed = 0;
        uint32_t* cnt;
        const uint32_t* list = b_list(b, cur, l, &cnt);
        uint32_t best = cur;
        for (uint32_t i = 0; i < *cnt; ++i) {
          const float d = b_dist(b, q, list[i]);
          if (d < closest) { closest = d; best = list[i]; changed = 1; }
        }
        cur = best;
      } while (changed);
    }
    const entry e = {cur, b_dist(b, q, cur)};
    heap_push(top, e);
  } else {
    const entry e = {b->ep, ep_dist};
    heap_push(top, e);
  }
  uint32_t connect_from = lvl;
  if (is_new_level) --connect_from;                     /* :146-148 */
  for (int cl = (int)connect_from; cl >= 0; --cl) {
    b_search_level(b, q, efc, (uint32_t)cl, top, next, vis);
    b_select(b, top, m);                                /* :163 */
    uint32_t* own_cnt;
    uint32_t* own = b_list(b, id, (uint32_t)cl, &own_cnt);
    *own_cnt = 0;
    for (size_t i = 0; i < top->n; ++i) own[(*own_cnt)++] = top->a[i].row;   /* heap-array order, :170-172 */
    const uint32_t m_max = cl == 0 ? b->m0 : m;
    for (size_t i = 0; i < top->n; ++i) {               /* :180-225 */
      const uint32_t nb = top->a[i].row;
      const float nb_dist = top->a[i].dist;
      uint32_t* cnt;
      uint32_t* list = b_list(b, nb, (uint32_t)cl, &cnt);
      if (*cnt < m_max) {
        list[(*cnt)++] = id;
      } else {
        tmp->n = 0;
        const entry self = {id, nb_dist};
        heap_push(tmp, self);
        for (uint32_t j = 0; j < *cnt; ++j) {
          const entry o = {list[j], b_dist(b, b->base + (size_t)nb * b->dim, list[j])};
          heap_push(tmp, o);
        }
        b_select(b, tmp, m_max);
        *cnt = 0;
        for (size_t j = 0; j < tmp->n; ++j) list[(*cnt)++] = tmp->a[j].row;
      }
    }
    while (cl > 0 && top->n > 1) heap_pop(top);         /* :228-230 */
  }
  if (is_new_level) b->ep = id;                         /* :237-247 */
}

int orc_build(const float* base, uint32_t n, uint32_t dim, uint32_t m, uint32_t efc, uint32_t seed, int ip, uint8_t** dump,
              uint64_t* dump_size, uint64_t* distcomps) {
  if (!base || n == 0 || m < 2 || !dump || !dump_size) return -1;
  builder b; memset(&b, 0, sizeof b);
  b.n = n; b.dim = dim; b.m = m; b.m0 = 2 * m; b.base = base; b.ip = ip;
  b.level = (uint32_t*)calloc(n, sizeof(uint32_t));
  b.l0_cnt = (uint32_t*)calloc(n, sizeof(uint32_t));
  b.l0 = (uint32_t*)malloc((size_t)n * b.m0 * sizeof(uint32_t));
  b.up_base = (uint64_t*)calloc(n, sizeof(uint64_t));
  mt19937 gen; mt_seed(&gen, seed);
  const double norm = 1.0 / log((double)m);
  heap top = {0, 0, 0, 1}, next = {0, 0, 0, 0}, tmp = {0, 0, 0, 1};
  visited_set vis = {(uint8_t*)calloc(n, 1), 0, 0, 0};
  for (uint32_t id = 0; id < n; ++id) {
    const uint32_t lvl = (uint32_t)floor(-log(mt_canonical(&gen)) * norm);   /* :48 — drawn for every insert, the first too */
    b_insert(&b, id, lvl, efc, &top, &next, &tmp, &vis);
  }
  /* emit what memory node 1 would hold: nodes in allocation (= insertion) order (memory_node.hh:14-26,187-195) */
  uint64_t* off = (uint64_t*)malloc((size_t)n * sizeof(uint64_t));
  uint64_t cur = 16;
  for (uint32_t r = 0; r < n; ++r) { off[r] = cur; cur += node_alloc_bytes(dim, m, b.level[r]); }
  uint8_t* d = (uint8_t*)calloc(cur, 1);
  memcpy(d, &cur, 8);
  memcpy(d + 8, &off[b.ep], 8);                          /* RemotePtr: memory node 0, byte offset */
  for (uint32_t r = 0; r < n; ++r) {
    uint8_t* nd = d + off[r];
    const uint64_t header = r == b.ep ? (1ull << 16) : 0;
    memcpy(nd, &header, 8); memcpy(nd + 8, &r, 4); memcpy(nd + 12, &b.level[r], 4);
    memcpy(nd + 16, base + (size_t)r * dim, 4 * (size_t)dim);
    uint8_t* l0 = nd + node_header_bytes(dim);
    memcpy(l0, &b.l0_cnt[r], 4);
    for (uint32_t j = 0; j < b.l0_cnt[r]; ++j) memcpy(l0 + 4 + 8 * (size_t)j, &off[b.l0[(size_t)r * b.m0 + j]], 8);
    for (uint32_t l = 1; l <= b.level[r]; ++l) {
      uint8_t* lu = l0 + list0_bytes(m) + (size_t)(l - 1) * listu_bytes(m);
      const uint64_t u = b.up_base[r] + (l - 1);
      memcpy(lu, &b.up_cnt[u], 4);
      for (uint32_t j = 0; j < b.up_cnt[u]; ++j) memcpy(lu + 4 + 8 * (size_t)j, &off[b.up[u * m + j]], 8);
    }
  }
  *dump = d; *dump_size = cur;
  if (distcomps) *distcomps = b.distcomps;
  free(off); free(b.level); free(b.l0_cnt); free(b.l0); free(b.up_base); free(b.up_cnt); free(b.up);
  free(top.a); free(next.a); free(tmp.a); free(vis.seen); free(vis.touched);
  return 0;
}

void orc_free_buffer(void* p) { free(p); }
