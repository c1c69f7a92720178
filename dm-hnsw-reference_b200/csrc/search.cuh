// search.cuh — device-side building blocks shared by the query kernel (search.cu) and the construction kernels
// (build.cu): warp-cooperative distance evaluation of a list of rows, the per-warp sorted candidate queue and the
// exact visited set.  One warp owns one query (or one insert); nothing here synchronises beyond the warp.
//
// Reference semantics restated: HNSW::search_for_one (src/hnsw/hnsw.hh:332-393),
// HNSW::search_level (:407-476), heap::Heap::push_k (src/hnsw/heap.hh:34-41), L2/IP distances
// (src/hnsw/distance.hh:80-151).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "graph.h"

namespace shn {

#ifndef SHN_STEP
#define SHN_STEP 8
#endif
#ifndef SHN_GSTEP
#define SHN_GSTEP 4
#endif
constexpr uint32_t kFull = 0xFFFFFFFFu;
constexpr uint32_t kExpanded = 0x80000000u;  // flag bit in a queue entry's row id
constexpr int kMaxList = 64;                  // 2m <= 64

#ifndef SHN_ROW_L1
// rows are read once per query: keep them out of L1 so that the adjacency lists and the pointer tables stay there
__device__ __forceinline__ float4 ldg_f4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
#else
__device__ __forceinline__ float4 ldg_f4(const float4* p) { return __ldg(p); }
#endif

// Ask the L2 for the line at p (per-lane address; nothing to wait for).  The bulk form (cp.async.bulk.prefetch.L2, SASS UBLKPF)
// takes ONE address per warp from a uniform register, so a list of rows would be walked lane by lane; this one covers 32
// lines per instruction.
__device__ __forceinline__ void prefetch_l2_line(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p) : "memory");
}
#ifdef SHN_ROW_PREFETCH_BULK
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
#endif

// ---------------------------------------------------------------------------------------------------------------
// Distances, in the reference's summation order (bit-identical results; see oracle/hnsw_oracle.c for how the
// order was read off the reference build).  The reference accumulates 8 AVX lanes j over 16-element chunks c:
//   L2:  S_j += fma(d0_j, d0_j, d1_j * d1_j)      d0 = q[16c+j] - v[16c+j],  d1 = q[16c+8+j] - v[16c+8+j]
//   IP:  S_j += fma(q[16c+j], v[16c+j], q[16c+8+j] * v[16c+8+j])
// then x_j = S_j + S_{j+4}, r = (x1+x3)+(x0+x2) (L2) resp. r = ((1-(S4+S5)) - ((S0+S1)+(S2+S3))) - (S6+S7) (IP), then a
// scalar tail.  Rows are stored in HBM so that this order is also the coalesced order (graph.h row_pos): EIGHT lanes
// own a row, lane t IS the AVX lane j = t, and the 16-byte piece t of every 128-byte block holds exactly the four
// floats lane j needs from two consecutive chunks.  One warp-wide 128-bit load therefore reads four whole cache
// lines (4 rows x 128 B) — a quarter of the L1TEX wavefronts of a sector-per-lane-pair layout — and the
// accumulation needs no shuffles until the final horizontal step.  With the default SHN_PASSES = 2, 8 rows are in flight
// per warp (4 row groups x 2 passes), i.e. 8 independent 128-bit loads per lane at d = 128 (96 registers, 5 CTAs/SM —
// measured better than 16 rows in flight at 128 registers and 4 CTAs/SM).
// ---------------------------------------------------------------------------------------------------------------

// One 16-byte piece q / a = {first half of chunk c, first half of chunk c+1, second half of chunk c, second half of
// chunk c+1} for AVX lane t.  The two chunks go through the packed fp32x2 pipe together (sub.f32x2 / mul.f32x2 / fma.f32x2,
// sm_100: each half is an ordinary IEEE fp32 operation, so the bits are those of the scalar sequence) and are then
// added to the running sum one after the other, as the reference does.
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
template <bool IP>
__device__ __forceinline__ void block_accumulate(const float4& q, const float4& a, float& s, bool second_chunk) {
  const unsigned long long q0 = pack2(q.x, q.y), q1 = pack2(q.z, q.w), a0 = pack2(a.x, a.y), a1 = pack2(a.z, a.w);
  unsigned long long f;
  if (IP) {
    unsigned long long m;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(m) : "l"(q1), "l"(a1));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(f) : "l"(q0), "l"(a0), "l"(m));
  } else {
    unsigned long long d0, d1, m;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d0) : "l"(q0), "l"(a0));
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d1) : "l"(q1), "l"(a1));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(m) : "l"(d1), "l"(d1));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(f) : "l"(d0), "l"(d0), "l"(m));
  }
  float f_lo, f_hi;
  unpack2(f, f_lo, f_hi);
  s = __fadd_rn(s, f_lo);
  if (second_chunk) s = __fadd_rn(s, f_hi);
}

// Horizontal step (distance.hh:40 / :134-139); the result is valid in lane t == 0 of the row group.
template <bool IP>
__device__ __forceinline__ float horizontal(float s) {
  if (IP) {
    const float p = __fadd_rn(s, __shfl_xor_sync(kFull, s, 1));        // t=0: S0+S1, 2: S2+S3, 4: S4+S5, 6: S6+S7
    const float a = __fadd_rn(p, __shfl_xor_sync(kFull, p, 2));        // t=0: (S0+S1)+(S2+S3)
    const float b = __shfl_xor_sync(kFull, p, 4);                      // t=0: S4+S5
    const float c = __shfl_xor_sync(kFull, p, 6);                      // t=0: S6+S7
    return __fsub_rn(__fsub_rn(__fsub_rn(1.0f, b), a), c);
  } else {
    const float x = __fadd_rn(s, __shfl_xor_sync(kFull, s, 4));        // t=j<4: x_j = S_j + S_{j+4}
    const float y = __fadd_rn(x, __shfl_xor_sync(kFull, x, 2));        // t=0: x0+x2, t=1: x1+x3
    return __fadd_rn(__shfl_xor_sync(kFull, y, 1), y);                 // t=0: (x1+x3)+(x0+x2)
  }
}

// Scalar tail (distance.hh:112-115 / :134-139): the dim%16 trailing elements are added one by one, in order, to the
// finished sum.  They follow the blocks in natural order (graph.h), lane t < 4 of the row group holds tail elements
// 4t..4t+3 (loaded together with the first wave of blocks), and the running value is handed from lane to lane.
template <bool IP>
__device__ __forceinline__ float tail_chain(float r, const float4& q, const float4& v, uint32_t ntail, int t) {
  float acc = IP ? 0.f : r;  // IP: t = sum q*v, then r -= t
  for (int hop = 0; hop < 4; ++hop) {
    if (4u * hop >= ntail) break;  // warp-uniform
    if (hop > 0) acc = __shfl_up_sync(kFull, acc, 1);  // lane hop takes over from lane hop-1
    if (t == hop) {
      const float qs[4] = {q.x, q.y, q.z, q.w}, vs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (4u * hop + e < ntail) {
          if (IP) acc = __fadd_rn(acc, __fmul_rn(qs[e], vs[e]));
          else { const float d = __fsub_rn(qs[e], vs[e]); acc = __fadd_rn(acc, __fmul_rn(d, d)); }
        }
      }
    }
  }
  // the finished value sits in lane (ntail-1)/4; bring it back to lane 0
  const int last = static_cast<int>((ntail - 1) >> 2);
  acc = __shfl_down_sync(kFull, acc, last);
  return IP ? __fsub_rn(r, acc) : acc;
}

#ifdef SHN_EVAL_NOINLINE
#define SHN_EVAL_ATTR __noinline__
#else
#define SHN_EVAL_ATTR __forceinline__
#endif
// dist(query, row) for rows s_rows[0..cnt) -> s_out[0..cnt).  s_q: the query in shared memory in the stored layout
// (16-byte aligned).  NCHUNK > 0: the dimension, fixed at compile time (96, 128, 200, 960: the shapes BASELINE.json
// names); NCHUNK == 0: any dim.
// PASSES: row groups of 4 evaluated together (SHN_PASSES = 2: 8 rows in flight per warp, the beam-search setting; 1 = the compact
// variant for the entry point, the greedy descent and the selection heuristic, where register pressure matters more).
#ifndef SHN_PASSES
#define SHN_PASSES 2
#endif
#ifndef SHN_SMALL_PASSES
#define SHN_SMALL_PASSES 1
#endif
// On a partitioned handle s_rows holds the ids the rows are READ under (graph.h read_id: the row itself, or its halo copy).
template <bool IP, int NCHUNK, int PASSES = SHN_PASSES>
__device__ SHN_EVAL_ATTR void eval_rows(const DeviceGraph& g, const float* s_q, const uint32_t* s_rows, uint32_t cnt,
                                          float* s_out, int lane) {
  const int t = lane & 7, grp = lane >> 3;
  const float4* s_q4 = reinterpret_cast<const float4*>(s_q);
  // NCHUNK > 0 is the dimension itself, known at compile time: block count, odd last chunk and tail length fold to
  // constants and the wave loop unrolls completely
  const uint32_t dim = NCHUNK > 0 ? static_cast<uint32_t>(NCHUNK) : g.dim;
  const uint32_t nblk = row_blocks(dim);
  const bool odd = (dim >> 4) & 1u;
  const uint32_t ntail = dim & 15u;
  constexpr int WAVE = 4;  // blocks per wave of loads: 4 rows x 4 blocks = 16 loads in flight per lane
  for (uint32_t base = 0; base < cnt; base += 4 * PASSES) {
    const float4* rp[PASSES];
    float s[PASSES];
#pragma unroll
    for (int p = 0; p < PASSES; ++p) {
      s[p] = 0.f;
      const uint32_t i = base + 4 * p + grp;
      const uint32_t row = s_rows[i < cnt ? i : cnt - 1];  // clamp: a redundant load instead of a divergent branch
      rp[p] = vec_row<false>(g, row) + t;
    }
    const uint32_t npass = min(static_cast<uint32_t>(PASSES), (cnt - base + 3) >> 2);  // warp-uniform: passes that hold at least one row
    float4 tv[PASSES];
    if (ntail) {
#pragma unroll
      for (int p = 0; p < PASSES; ++p) {
        tv[p] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (static_cast<uint32_t>(p) < npass && 4u * t < ntail) tv[p] = ldg_f4(rp[p] + 8 * nblk);  // rp already includes + t
      }
    }
    for (uint32_t b0 = 0; b0 < nblk; b0 += WAVE) {
      float4 v[PASSES][WAVE];
#pragma unroll
      for (int p = 0; p < PASSES; ++p) {
        if (static_cast<uint32_t>(p) < npass) {
#pragma unroll
          for (int w = 0; w < WAVE; ++w) {
            if (b0 + w < nblk) v[p][w] = ldg_f4(rp[p] + 8 * (b0 + w));
          }
        }
      }
#pragma unroll
      for (int w = 0; w < WAVE; ++w) {
        if (b0 + w < nblk) {
          const float4 q = s_q4[8 * (b0 + w) + t];
          const bool second = !(odd && b0 + w == nblk - 1);
#pragma unroll
          for (int p = 0; p < PASSES; ++p) {
            if (static_cast<uint32_t>(p) < npass) block_accumulate<IP>(q, v[p][w], s[p], second);
          }
        }
      }
    }
#pragma unroll
    for (int p = 0; p < PASSES; ++p) {
      if (static_cast<uint32_t>(p) < npass) {
        float r = horizontal<IP>(s[p]);
        if (ntail) {
          r = __shfl_sync(kFull, r, lane & ~7);  // every lane of the group starts from the group's sum
          r = tail_chain<IP>(r, s_q4[8 * nblk + (t & 3)], tv[p], ntail, t);
        }
        const uint32_t i = base + 4 * p + grp;
        if (t == 0 && i < cnt) s_out[i] = r;
      }
    }
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------------------------
// Sorted candidate queue: the reference's top_candidates (max-heap, <= ef) and next_candidates (min-heap) folded
// into ONE ascending array of <= ef entries; an entry's kExpanded bit says it has left next_candidates.
// An entry evicted from top_candidates has dist >= the new farthest, so the reference would only still expand
// it on an exact distance tie (hnsw.hh:424 breaks on strictly greater) — see DESIGN.md "ties".
// ---------------------------------------------------------------------------------------------------------------

// Admit the candidates (s_rows[i], s_dist[i]), i < cnt <= 64, given in stored list order, in ONE merge.
// Equivalent to the reference's one-by-one admission (hnsw.hh:456-465) whenever no two distances are equal: the
// running farthest distance only shrinks, so what survives is exactly the ef smallest of old and new entries;
// old entries stay ahead of new ones at equal distance and new ones keep their list order.
// Returns the lowest queue position that received a new entry (kInvalid if none).
// Steps: compact the survivors (closer than the current farthest entry), rank them among themselves by counting, find
// each one's slot among the old entries by binary search, then move the old entries at or above the lowest slot up by
// the number of survivors ahead of them, highest block first.  (A formulation that does all the counting in one pass over
// registers-held queue blocks was measured 4-5 % slower at every ef on B200 and removed.)
__device__ __forceinline__ uint32_t queue_merge(float* qd, uint32_t* qi, uint32_t& qsize, uint32_t ef, uint32_t* s_rows,
                                                float* s_dist, uint32_t cnt, int lane) {
  const bool full = qsize == ef;
  const float far = full ? qd[ef - 1] : __int_as_float(0x7f800000);
  float d0 = 0.f, d1 = 0.f;
  uint32_t r0 = 0, r1 = 0;
  bool k0 = false, k1 = false;
  if (static_cast<uint32_t>(lane) < cnt) { d0 = s_dist[lane]; r0 = s_rows[lane]; k0 = !full || d0 < far; }
  if (static_cast<uint32_t>(lane) + 32 < cnt) { d1 = s_dist[lane + 32]; r1 = s_rows[lane + 32]; k1 = !full || d1 < far; }
  const uint32_t m0 = __ballot_sync(kFull, k0), m1 = __ballot_sync(kFull, k1);
  const uint32_t c0 = __popc(m0), c = c0 + __popc(m1);
  if (c == 0) return kInvalid;
  const uint32_t below = (1u << lane) - 1;
  const uint32_t x0 = __popc(m0 & below), x1 = c0 + __popc(m1 & below);  // compacted index, list order
  __syncwarp();
  if (k0) s_dist[x0] = d0;
  if (k1) s_dist[x1] = d1;
  __syncwarp();
  uint32_t rank0 = 0, rank1 = 0;
  for (uint32_t y = 0; y < c; ++y) {
    const float dy = s_dist[y];
    rank0 += (dy < d0 || (dy == d0 && y < x0)) ? 1u : 0u;
    rank1 += (dy < d1 || (dy == d1 && y < x1)) ? 1u : 0u;
  }
  __syncwarp();
  if (k0) s_dist[rank0] = d0;  // survivors, ascending
  if (k1) s_dist[rank1] = d1;
  __syncwarp();
  // position among the old entries: those with distance <= d stay ahead
  uint32_t f0 = kInvalid, f1 = kInvalid;
  if (k0) {
    uint32_t lo = 0, hi = qsize;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (qd[mid] <= d0) lo = mid + 1; else hi = mid; }
    f0 = lo + rank0;
  }
  if (k1) {
    uint32_t lo = 0, hi = qsize;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (qd[mid] <= d1) lo = mid + 1; else hi = mid; }
    f1 = lo + rank1;
  }
  uint32_t pmin = min(f0, f1);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) pmin = min(pmin, __shfl_xor_sync(kFull, pmin, o));
  if (pmin >= ef) return kInvalid;
  // old entries at or above pmin move up by the number of survivors strictly closer, highest block first
  int hi = static_cast<int>(qsize);
  while (hi > static_cast<int>(pmin)) {
    const int lo = max(static_cast<int>(pmin), hi - 32);
    const int j = lo + lane;
    const bool act = j < hi;
    float td = 0.f;
    uint32_t ti = 0, shift = 0;
    if (act) {
      td = qd[j]; ti = qi[j];
      for (uint32_t y = 0; y < c; ++y) shift += s_dist[y] < td ? 1u : 0u;
    }
    __syncwarp();
    if (act && j + shift < ef) { qd[j + shift] = td; qi[j + shift] = ti; }
    __syncwarp();
    hi = lo;
  }
  if (k0 && f0 < ef) { qd[f0] = d0; qi[f0] = r0; }
  if (k1 && f1 < ef) { qd[f1] = d1; qi[f1] = r1; }
  __syncwarp();
  qsize = min(ef, qsize + c);
  return pmin;
}

// ---------------------------------------------------------------------------------------------------------------
// Exact visited set (the reference's unordered_set<RemotePtr>, hnsw.hh:408,441-443): open addressing in shared
// memory; once `limit` keys are in, further keys spill to a per-warp table in HBM.  Never a false positive.
// ---------------------------------------------------------------------------------------------------------------

struct VisitedSet {
  uint32_t* tab;     // shared, cap entries
  uint32_t* ovf;     // global, ovf_cap entries (kept all-kInvalid between queries)
  uint32_t cap, limit, ovf_cap, ovf_limit;
  uint32_t count, ovf_count;  // warp-uniform
  bool failed;
  bool compact;      // 16-bit keys (visited_compact below); needs cap == 2048 and ids below 2^24
};

__device__ __forceinline__ uint32_t hash_row(uint32_t h) {  // murmur3 finaliser
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h;
}

__device__ __forceinline__ void visited_reset(VisitedSet& v, int lane) {
  for (uint32_t j = lane; j < v.cap; j += 32) v.tab[j] = kInvalid;
  if (v.ovf_count) {
    for (uint32_t j = lane; j < v.ovf_cap; j += 32) v.ovf[j] = kInvalid;
  }
  v.count = 0; v.ovf_count = 0; v.failed = false;
  __syncwarp();
}

// One 128-bit shared-memory read that the compiler may not cache across the CAS of another lane.
__device__ __forceinline__ uint4 lds_bucket(const uint32_t* p) {
  uint4 k;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(k.x), "=r"(k.y), "=r"(k.z), "=r"(k.w)
               : "r"(static_cast<uint32_t>(__cvta_generic_to_shared(p)))
               : "memory");
  return k;
}

// Compact form of the shared table, for graphs of at most 2^24 ids: the same 8 KB hold 4096 keys instead of 2048.  The id goes
// through a BIJECTION of the 24-bit integers (odd multiplications and xor-shifts, each invertible); the top 9 bits of the result
// name the home bucket (512 buckets of 16 bytes), the other 15 are the key stored there — bucket and key together still identify
// the id exactly, so there is no false positive.  A bucket holds 8 keys; a key whose home bucket is full goes to the next bucket
// with the top bit set ("displaced by one"), and only if that is full as well to the table in HBM.  Buckets only ever fill, so
// "home (or next) still has room" proves that the key was never pushed further, and the HBM table is consulted only by the
// few keys whose two buckets are both full.  One 128-bit read + four packed 16-bit compares answer a bucket.
__device__ __forceinline__ uint32_t mix24(uint32_t h) {
  h = (h * 0x9E3779u) & 0xFFFFFFu; h ^= h >> 12;
  h = (h * 0x85EBCBu) & 0xFFFFFFu; h ^= h >> 13;
  h = (h * 0xC2B2AFu) & 0xFFFFFFu; h ^= h >> 11;
  return h;
}
__device__ __forceinline__ bool visited_compact(VisitedSet& v, uint32_t id, bool active, int lane) {
  bool is_new = false, spill = false;
  if (active) {
    const uint32_t h = mix24(id);
    const uint32_t home = h >> 15, rem = h & 0x7FFFu;
    bool done = false;
#pragma unroll
    for (uint32_t step = 0; step < 2 && !done; ++step) {
      const uint32_t key = rem | (step << 15);
      if (key == 0xFFFFu) break;                      // the one key that would read as "empty": straight to HBM
      const uint32_t pat = key * 0x10001u;
      uint32_t* bucket = v.tab + 4u * ((home + step) & 511u);
      for (;;) {
        const uint4 k = lds_bucket(bucket);
        if (__vcmpeq2(k.x, pat) | __vcmpeq2(k.y, pat) | __vcmpeq2(k.z, pat) | __vcmpeq2(k.w, pat)) { done = true; break; }
        const uint32_t ex = __vcmpeq2(k.x, 0xFFFFFFFFu), ey = __vcmpeq2(k.y, 0xFFFFFFFFu), ez = __vcmpeq2(k.z, 0xFFFFFFFFu),
                       ew = __vcmpeq2(k.w, 0xFFFFFFFFu);
        if (!(ex | ey | ez | ew)) break;              // full: try the next bucket (or HBM)
        const uint32_t w = ex ? 0u : (ey ? 1u : (ez ? 2u : 3u));
        const uint32_t old = w == 0 ? k.x : (w == 1 ? k.y : (w == 2 ? k.z : k.w));
        const uint32_t e = w == 0 ? ex : (w == 1 ? ey : (w == 2 ? ez : ew));
        const uint32_t want = (e & 0xFFFFu) ? ((old & 0xFFFF0000u) | key) : ((old & 0x0000FFFFu) | (key << 16));
        if (atomicCAS(bucket + w, old, want) == old) { is_new = true; done = true; break; }
        // another lane changed the word: look at the bucket again
      }
    }
    spill = !done;
  }
  v.count += __popc(__ballot_sync(kFull, is_new));
  if (__any_sync(kFull, spill)) {  // warp-uniform
    if (v.ovf_count + 32 > v.ovf_limit) {
      v.failed = true;  // caller reports SHN_ERR_CAPACITY
    } else {
      bool added = false;
      if (spill) {
        uint32_t s = (hash_row(id) >> 7) & (v.ovf_cap - 1);
        for (;;) {
          const uint32_t old = atomicCAS(&v.ovf[s], kInvalid, id);
          if (old == kInvalid) { added = true; break; }
          if (old == id) break;
          s = (s + 1) & (v.ovf_cap - 1);
        }
      }
      is_new |= added;
      v.ovf_count += __popc(__ballot_sync(kFull, added));
    }
  }
  __syncwarp();
  return is_new;
}

// Each lane with active==true offers one id (ids offered together are distinct); returns true for lanes whose
// id was not in the set (and now is).  The shared table is organised in 16-byte buckets of four keys: one 128-bit
// read answers "already visited" (about half of all offers) without an atomic, an insert costs that read plus one
// CAS on the first free slot of the bucket.
__device__ __forceinline__ bool visited_test_and_set(VisitedSet& v, uint32_t id, bool active, int lane) {
  if (v.compact) return visited_compact(v, id, active, lane);  // warp-uniform
  bool is_new = false;
  // any number of buckets: the home bucket is the high part of hash x buckets (no power-of-two constraint, so the table can
  // take exactly the shared memory the launch has left)
  const uint32_t nbuckets = v.cap >> 2;
  if (v.count + 32 <= v.limit) {  // warp-uniform: room for every lane
    if (active) {
      uint32_t b = __umulhi(hash_row(id), nbuckets);
      for (;;) {
        const uint4 k = lds_bucket(v.tab + 4 * b);
        if (k.x == id || k.y == id || k.z == id || k.w == id) break;
        const int e = k.x == kInvalid ? 0 : (k.y == kInvalid ? 1 : (k.z == kInvalid ? 2 : (k.w == kInvalid ? 3 : -1)));
        if (e < 0) { b = b + 1 == nbuckets ? 0u : b + 1; continue; }
        const uint32_t old = atomicCAS(&v.tab[4 * b + e], kInvalid, id);
        if (old == kInvalid) { is_new = true; break; }
        if (old == id) break;
        // another lane took the slot: look at the same bucket again
      }
    }
    v.count += __popc(__ballot_sync(kFull, is_new));
  } else {
    bool found = false;
    if (active) {  // the shared table is closed for inserts but still answers lookups
      uint32_t b = __umulhi(hash_row(id), nbuckets);
      for (;;) {
        const uint4 k = reinterpret_cast<const uint4*>(v.tab)[b];
        if (k.x == id || k.y == id || k.z == id || k.w == id) { found = true; break; }
        if (k.w == kInvalid) break;  // buckets fill front to back: a free last slot ends the probe sequence
        b = b + 1 == nbuckets ? 0u : b + 1;
      }
    }
    if (v.ovf_count + 32 > v.ovf_limit) {
      v.failed = true;  // caller reports SHN_ERR_CAPACITY
    } else {
      if (active && !found) {
        uint32_t s = (hash_row(id) >> 7) & (v.ovf_cap - 1);
        for (;;) {
          const uint32_t old = atomicCAS(&v.ovf[s], kInvalid, id);
          if (old == kInvalid) { is_new = true; break; }
          if (old == id) break;
          s = (s + 1) & (v.ovf_cap - 1);
        }
      }
      v.ovf_count += __popc(__ballot_sync(kFull, is_new));
    }
  }
  __syncwarp();
  return is_new;
}

// ---------------------------------------------------------------------------------------------------------------
// ef-bounded best-first search on one level (HNSW::search_level, hnsw.hh:407-476).  On entry the queue holds the
// seed entries (unexpanded) and the visited set holds their rows.  Level 0 reads the 2m-wide lists, upper levels the
// m-wide ones.  Counters: distance computations / nodes visited / lists read on this level.
// ---------------------------------------------------------------------------------------------------------------
template <bool IP, int NCHUNK, bool PART = false, int PASSES = SHN_PASSES>
__device__ __forceinline__ void beam_search(const DeviceGraph& g, const float* s_q, uint32_t level, uint32_t ef, float* qd,
                                            uint32_t* qi, uint32_t& qsize, uint32_t* s_rows, float* s_dist, VisitedSet& vis,
                                            uint32_t& c_dist, uint32_t& c_vis, uint32_t& c_lists, uint32_t& c_hot,
                                            uint32_t& c_local, uint32_t& c_halo, uint32_t* s_read, int lane) {
  const uint32_t width = level == 0 ? g.m0 : g.m;
  uint32_t lb = 0;  // every entry below lb is expanded
  for (;;) {
    // next_candidates.pop(): the closest entry not yet expanded
    uint32_t pos = kInvalid;
    for (uint32_t b = lb & ~31u; b < qsize; b += 32) {
      const uint32_t j = b + lane;
      const bool un = j < qsize && j >= lb && !(qi[j] & kExpanded);
      const uint32_t mask = __ballot_sync(kFull, un);
      if (mask) { pos = b + __ffs(mask) - 1; break; }
    }
    if (pos == kInvalid) break;  // what is left of next_candidates is farther than top_candidates.top() (:424)
    const uint32_t cand = qi[pos];
    __syncwarp();
    if (lane == 0) qi[pos] = cand | kExpanded;
    lb = pos + 1;
    ++c_lists;

    // read_neighborlist (:437)
    const uint32_t* list = level == 0 ? l0_row<PART>(g, cand)
                                      : g.up + (static_cast<size_t>(__ldg(g.up_base + cand)) + (level - 1)) * g.m;
    const uint32_t nb0 = static_cast<uint32_t>(lane) < width ? __ldg(list + lane) : kInvalid;
    uint32_t nb1 = kInvalid;
    if (width > 32) nb1 = static_cast<uint32_t>(lane) + 32 < width ? __ldg(list + lane + 32) : kInvalid;

    // the visited filter, in stored order (:440-443)
    uint32_t cnt = 0;
    {
      const bool fresh = visited_test_and_set(vis, nb0, nb0 != kInvalid, lane);
      const uint32_t mask = __ballot_sync(kFull, fresh);
      if (fresh) s_rows[__popc(mask & ((1u << lane) - 1))] = nb0;
      cnt = __popc(mask);
    }
    if (width > 32) {
      const bool fresh = visited_test_and_set(vis, nb1, nb1 != kInvalid, lane);
      const uint32_t mask = __ballot_sync(kFull, fresh);
      if (fresh) s_rows[cnt + __popc(mask & ((1u << lane) - 1))] = nb1;
      cnt += __popc(mask);
    }
    __syncwarp();
    if (cnt == 0) continue;
    c_vis += cnt; c_dist += cnt;
    if (PART) {
      // the ids the rows are read under (a halo copy replaces a row of a peer's share), where they live, and — during a
      // warm-up pass — how often each is read
      for (uint32_t base = 0; base < cnt; base += 32) {
        const uint32_t i = base + lane;
        const bool in = i < cnt;
        uint32_t cls = kRowPeer + 1, row = 0, rd = 0;
        if (in) {
          row = s_rows[i];
          if (g.visit_count) atomicAdd(g.visit_count + row, 1u);
          rd = read_id<true>(g, row, cls);
        }
        c_hot += __popc(__ballot_sync(kFull, cls == kRowHot));
        c_local += __popc(__ballot_sync(kFull, cls == kRowOwn));
        c_halo += __popc(__ballot_sync(kFull, cls == kRowHalo));
        // Rows behind NVLink go to the front of the list (both groups keep their stored order): a wave of loads takes as long
        // as its slowest row, so the remote rows should share one wave.  What reaches queue_merge is a permutation of the
        // same candidates; only the order of two candidates of this list at the exact same distance can differ.
        const uint32_t in_mask = __ballot_sync(kFull, in);
        const uint32_t rem_mask = __ballot_sync(kFull, cls == kRowPeer);
        uint32_t pos = i;
        if (rem_mask != 0u && rem_mask != in_mask) {  // warp-uniform
          const uint32_t below = (1u << lane) - 1u;
          pos = base + ((rem_mask >> lane) & 1u ? __popc(rem_mask & below) : __popc(rem_mask) + __popc(in_mask & ~rem_mask & below));
          __syncwarp();
          if (in) s_rows[pos] = row;
        }
        if (in) s_read[pos] = rd;
      }
      __syncwarp();
    }
#ifndef SHN_NO_ROW_PREFETCH
    // The rows are evaluated in waves of 4 * PASSES; every wave waits out a full DRAM round trip.  All rows of the list are
    // known now, so the ones behind the first wave are requested into L2 at once (one prefetch per 128-byte line, a lane each):
    // when their wave comes they are an L2 hit.  No speculation, no extra DRAM traffic.
    if (PASSES <= 2 && cnt > 4u * PASSES) {  // measured: +2.4 % at ef=64 (8-row waves), nothing to gain with 16-row waves
      const uint32_t* rd = PART ? s_read : s_rows;
#ifdef SHN_ROW_PREFETCH_BULK
      for (uint32_t i = 4u * PASSES + lane; i < cnt; i += 32) prefetch_l2_bulk(vec_row<false>(g, rd[i]), g.row_f4 * 16u);
#else
      const uint32_t lines = g.row_f4 >> 3;  // 128-byte lines per row
      for (uint32_t j = lane; j < (cnt - 4u * PASSES) * lines; j += 32) {
        const uint32_t i = 4u * PASSES + j / lines, l = j - (j / lines) * lines;
        prefetch_l2_line(vec_row<false>(g, rd[i]) + 8u * l);
      }
#endif
    }
#endif
    eval_rows<IP, NCHUNK, PASSES>(g, s_q, PART ? s_read : s_rows, cnt, s_dist, lane);

    // (Requesting the NEXT expansion's list here, before the merge — the entry is known exactly: the closer of the first
    // unexpanded queue entry and the closest candidate of this list — was measured on B200: no gain at any ef; the ~60 extra
    // instructions per expansion cost what the hidden round trip saves.  The kernel is bound by instruction issue
    // latency x resident warps as much as by memory latency.)
    // admission against the running farthest distance (:456-465, heap.hh:34-41), all neighbours in one merge
    const uint32_t at = queue_merge(qd, qi, qsize, ef, s_rows, s_dist, cnt, lane);
    if (at < lb) lb = at;
  }
}

// One step of search_for_one (hnsw.hh:342-391) on `level`: scan the whole list of `cur`, move to the list minimum if
// it is strictly closer.  Returns true if cur changed.
template <bool IP, int NCHUNK>
__device__ __forceinline__ bool greedy_step(const DeviceGraph& g, const float* s_q, uint32_t level, uint32_t& cur,
                                            float& closest, uint32_t* s_rows, float* s_dist, uint32_t& c_dist,
                                            uint32_t& c_vis, uint32_t& c_lists, int lane) {
  const uint32_t* list = g.up + (static_cast<size_t>(__ldg(g.up_base + cur)) + (level - 1)) * g.m;
  ++c_lists;
  const uint32_t nb = static_cast<uint32_t>(lane) < g.m ? __ldg(list + lane) : kInvalid;
  const uint32_t cnt = __popc(__ballot_sync(kFull, nb != kInvalid));  // lists are stored compacted
  if (nb != kInvalid) s_rows[lane] = nb;
  __syncwarp();
  if (cnt == 0) return false;
  eval_rows<IP, NCHUNK, SHN_SMALL_PASSES>(g, s_q, s_rows, cnt, s_dist, lane);
  c_vis += cnt; c_dist += cnt;
  // sequential scan with strict '<' (hnsw.hh:378-382) == first index of the list minimum, if it beats closest
  float bd = static_cast<uint32_t>(lane) < cnt ? s_dist[lane] : __int_as_float(0x7f800000);
  uint32_t bi = lane;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float od = __shfl_xor_sync(kFull, bd, o);
    const uint32_t oi = __shfl_xor_sync(kFull, bi, o);
    if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
  }
  bool changed = false;
  if (bi < cnt && bd < closest) {
    closest = bd;
    cur = s_rows[bi];
    changed = true;
  }
  __syncwarp();
  return changed;
}

}  // namespace shn
