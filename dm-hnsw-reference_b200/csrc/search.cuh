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

constexpr uint32_t kFull = 0xFFFFFFFFu;
constexpr uint32_t kExpanded = 0x80000000u;  // flag bit in a queue entry's row id
constexpr int kMaxList = 64;                  // 2m <= 64

__device__ __forceinline__ float4 ldg_f4(const float4* p) { return __ldg(p); }

// ---------------------------------------------------------------------------------------------------------------
// Distances.
//
// FAST: lane l owns float4 #(l + 32 t) of the row; partial sums are combined across the warp.  For a batch of
// U rows the U partials per lane are reduced with a halving exchange (U/2 + U/4 + ... shuffles instead of 5 U).
// EXACT: reproduces the reference's summation order bit for bit (see oracle/hnsw_oracle.c): a lane pair owns one
// row, lane h of the pair accumulates AVX lanes 4h..4h+3 of the reference's 8-lane loop.
// ---------------------------------------------------------------------------------------------------------------

template <bool IP>
__device__ __forceinline__ float fast_partial(const float4& q, const float4& v, float acc) {
  if (IP) {
    acc = fmaf(q.x, v.x, acc); acc = fmaf(q.y, v.y, acc); acc = fmaf(q.z, v.z, acc); acc = fmaf(q.w, v.w, acc);
  } else {
    float d;
    d = q.x - v.x; acc = fmaf(d, d, acc);
    d = q.y - v.y; acc = fmaf(d, d, acc);
    d = q.z - v.z; acc = fmaf(d, d, acc);
    d = q.w - v.w; acc = fmaf(d, d, acc);
  }
  return acc;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// Reduce U (power of two, <= 8) per-lane partials across the warp.  Afterwards value index
// idx(lane) = bits of lane above log2(32/U) (see code) is complete in every lane of its group.
template <int U>
__device__ __forceinline__ float multi_reduce(float (&p)[U], int lane, int& owner_idx) {
  int idx = 0;
  int width = U;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    if (width > 1) {
      const int half = width / 2;
      const bool upper = (lane & o) != 0;
#pragma unroll
      for (int i = 0; i < half; ++i) {
        const float send = upper ? p[i] : p[i + half];
        const float keep = upper ? p[i + half] : p[i];
        p[i] = keep + __shfl_xor_sync(kFull, send, o);
      }
      idx = idx * 2 + (upper ? 1 : 0);
      width = half;
    } else {
      p[0] += __shfl_xor_sync(kFull, p[0], o);
    }
  }
  // idx was built most-significant-first over the halving rounds, but value i of round r came from p[i + half]
  // for the upper lanes: the surviving value index is sum(upper_r * half_r).
  owner_idx = idx;
  return p[0];
}

// Evaluate dist(query, row) for rows s_rows[0..cnt) -> s_out[0..cnt).  FAST path.
// qreg: the query as NV float4 per lane (float4 #(lane + 32 t)), zero beyond dim.
template <int NV, bool IP>
__device__ __forceinline__ void eval_rows_fast(const DeviceGraph& g, const float4 (&qreg)[NV], const uint32_t* s_rows,
                                               uint32_t cnt, float* s_out, int lane) {
  constexpr int U = NV == 1 ? 8 : (NV == 2 ? 4 : 2);
  const uint32_t dim4 = g.row_f4;
  for (uint32_t base = 0; base < cnt; base += U) {
    float4 v[U][NV];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t i = base + u;
      const uint32_t row = s_rows[i < cnt ? i : cnt - 1];  // clamp: redundant load instead of a branch
      const float4* rp = g.vec + static_cast<size_t>(row) * dim4;
#pragma unroll
      for (int t = 0; t < NV; ++t) {
        const uint32_t f = lane + 32 * t;
        v[u][t] = (NV * 32 == dim4 || f < dim4) ? ldg_f4(rp + f) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    float p[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float acc = 0.f;
#pragma unroll
      for (int t = 0; t < NV; ++t) acc = fast_partial<IP>(qreg[t], v[u][t], acc);
      p[u] = acc;
    }
    int idx;
    float r = multi_reduce<U>(p, lane, idx);
    if (IP) r = 1.0f - r;
    constexpr int group = 32 / U;  // lanes sharing one value
    if ((lane & (group - 1)) == 0 && base + idx < cnt) s_out[base + idx] = r;
  }
  __syncwarp();
}

// EXACT path: s_q = query in shared memory (dim floats, 16-byte aligned).
template <bool IP>
__device__ __forceinline__ void eval_rows_exact(const DeviceGraph& g, const float* s_q, const uint32_t* s_rows,
                                                uint32_t cnt, float* s_out, int lane) {
  const uint32_t dim = g.dim;
  const uint32_t nchunk = dim >> 4;  // distance.hh:88 qty16
  const int h = lane & 1;
  const float4* s_q4 = reinterpret_cast<const float4*>(s_q);
  for (uint32_t base = 0; base < cnt; base += 16) {
    const uint32_t i = base + (lane >> 1);
    const uint32_t row = s_rows[i < cnt ? i : cnt - 1];
    const float4* rp = g.vec + static_cast<size_t>(row) * g.row_f4;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;  // AVX lanes 4h .. 4h+3
    for (uint32_t c0 = 0; c0 < nchunk; c0 += 4) {
      float4 a[4], b[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c0 + c < nchunk) {
          a[c] = ldg_f4(rp + (c0 + c) * 4 + h);      // elements 16c + 4h ..      (first 8 of the 16)
          b[c] = ldg_f4(rp + (c0 + c) * 4 + 2 + h);  // elements 16c + 8 + 4h ..  (second 8)
        }
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c0 + c < nchunk) {
          const float4 qa = s_q4[(c0 + c) * 4 + h];
          const float4 qb = s_q4[(c0 + c) * 4 + 2 + h];
          if (IP) {  // s_j += fma(a_j, b_j, a_{8+j} * b_{8+j})   (query is the first operand, hnsw.hh:271,375,458)
            s0 = __fadd_rn(s0, __fmaf_rn(qa.x, a[c].x, __fmul_rn(qb.x, b[c].x)));
            s1 = __fadd_rn(s1, __fmaf_rn(qa.y, a[c].y, __fmul_rn(qb.y, b[c].y)));
            s2 = __fadd_rn(s2, __fmaf_rn(qa.z, a[c].z, __fmul_rn(qb.z, b[c].z)));
            s3 = __fadd_rn(s3, __fmaf_rn(qa.w, a[c].w, __fmul_rn(qb.w, b[c].w)));
          } else {   // s_j += fma(d0_j, d0_j, d1_j * d1_j)
            float d0, d1;
            d0 = __fsub_rn(qa.x, a[c].x); d1 = __fsub_rn(qb.x, b[c].x); s0 = __fadd_rn(s0, __fmaf_rn(d0, d0, __fmul_rn(d1, d1)));
            d0 = __fsub_rn(qa.y, a[c].y); d1 = __fsub_rn(qb.y, b[c].y); s1 = __fadd_rn(s1, __fmaf_rn(d0, d0, __fmul_rn(d1, d1)));
            d0 = __fsub_rn(qa.z, a[c].z); d1 = __fsub_rn(qb.z, b[c].z); s2 = __fadd_rn(s2, __fmaf_rn(d0, d0, __fmul_rn(d1, d1)));
            d0 = __fsub_rn(qa.w, a[c].w); d1 = __fsub_rn(qb.w, b[c].w); s3 = __fadd_rn(s3, __fmaf_rn(d0, d0, __fmul_rn(d1, d1)));
          }
        }
      }
    }
    float r;
    if (IP) {
      // r = ((1 - (S4+S5)) - ((S0+S1)+(S2+S3))) - (S6+S7)
      const float t01 = __fadd_rn(s0, s1), t23 = __fadd_rn(s2, s3);
      const float p01 = __shfl_xor_sync(kFull, t01, 1), p23 = __shfl_xor_sync(kFull, t23, 1);
      const float A = h ? __fadd_rn(p01, p23) : __fadd_rn(t01, t23);
      const float B = h ? t01 : p01;
      const float C = h ? t23 : p23;
      r = __fsub_rn(__fsub_rn(__fsub_rn(1.0f, B), A), C);
      if (dim & 15u) {
        float t = 0.f;
        const float* rs = reinterpret_cast<const float*>(rp);
        for (uint32_t e = nchunk * 16; e < dim; ++e) t = __fadd_rn(t, __fmul_rn(s_q[e], __ldg(rs + e)));
        r = __fsub_rn(r, t);
      }
    } else {
      // x_j = S_j + S_{j+4};  r = (x1 + x3) + (x0 + x2)
      const float x0 = __fadd_rn(s0, __shfl_xor_sync(kFull, s0, 1));
      const float x1 = __fadd_rn(s1, __shfl_xor_sync(kFull, s1, 1));
      const float x2 = __fadd_rn(s2, __shfl_xor_sync(kFull, s2, 1));
      const float x3 = __fadd_rn(s3, __shfl_xor_sync(kFull, s3, 1));
      r = __fadd_rn(__fadd_rn(x1, x3), __fadd_rn(x0, x2));
      const float* rs = reinterpret_cast<const float*>(rp);
      for (uint32_t e = nchunk * 16; e < dim; ++e) {
        const float d = __fsub_rn(s_q[e], __ldg(rs + e));
        r = __fadd_rn(r, __fmul_rn(d, d));
      }
    }
    if (h == 0 && i < cnt) s_out[i] = r;
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------------------------
// Sorted candidate queue: the reference's top_candidates (max-heap, <= ef) and next_candidates (min-heap) folded
// into ONE ascending array of <= ef entries; an entry's kExpanded bit says it has left next_candidates.
// An entry evicted from top_candidates has dist >= the new farthest, so the reference would only still expand
// it on an exact distance tie (hnsw.hh:424 breaks on strictly greater) — see DESIGN.md "ties".
// ---------------------------------------------------------------------------------------------------------------

// Insert (d, id) keeping ascending order, stable (after equal distances), dropping the last entry when full.
// Returns the insert position, or 0xFFFFFFFF if the entry fell off the end.
__device__ __forceinline__ uint32_t queue_insert(float* qd, uint32_t* qi, uint32_t& qsize, uint32_t ef, float d,
                                                 uint32_t id, int lane) {
  uint32_t pos = 0;
  for (uint32_t b = 0; b < qsize; b += 32) {
    const uint32_t j = b + lane;
    const bool le = (j < qsize) && (qd[j] <= d);
    pos += __popc(__ballot_sync(kFull, le));
  }
  const uint32_t nsize = qsize < ef ? qsize + 1 : ef;
  if (pos >= nsize) return kInvalid;
  int hi = static_cast<int>(nsize) - 1;  // source range is [pos, hi)
  while (hi > static_cast<int>(pos)) {
    const int lo = max(static_cast<int>(pos), hi - 32);
    const int j = lo + lane;
    const bool act = j < hi;
    float td = 0.f;
    uint32_t ti = 0;
    if (act) { td = qd[j]; ti = qi[j]; }
    __syncwarp();
    if (act) { qd[j + 1] = td; qi[j + 1] = ti; }
    __syncwarp();
    hi = lo;
  }
  if (lane == 0) { qd[pos] = d; qi[pos] = id; }
  __syncwarp();
  qsize = nsize;
  return pos;
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
};

__device__ __forceinline__ uint32_t hash_row(uint32_t h) {  // murmur3 finaliser: every input bit reaches every output bit
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

// Each lane with active==true offers one id; returns true for lanes whose id was not in the set (and now is).
__device__ __forceinline__ bool visited_test_and_set(VisitedSet& v, uint32_t id, bool active, int lane) {
  bool is_new = false;
  if (v.count + 32 <= v.limit) {  // warp-uniform: room for every lane
    if (active) {
      uint32_t s = hash_row(id) & (v.cap - 1);
      for (;;) {
        const uint32_t old = atomicCAS(&v.tab[s], kInvalid, id);
        if (old == kInvalid) { is_new = true; break; }
        if (old == id) break;
        s = (s + 1) & (v.cap - 1);
      }
    }
    v.count += __popc(__ballot_sync(kFull, is_new));
  } else {
    bool found = false;
    if (active) {  // shared table is closed for inserts but still answers lookups
      uint32_t s = hash_row(id) & (v.cap - 1);
      for (;;) {
        const uint32_t old = v.tab[s];
        if (old == kInvalid) break;
        if (old == id) { found = true; break; }
        s = (s + 1) & (v.cap - 1);
      }
    }
    if (v.ovf_count + 32 > v.ovf_limit) {
      v.failed = true;  // caller reports SHN_ERR_CAPACITY
    } else {
      if (active && !found) {
        uint32_t s = (hash_row(id) >> 11) & (v.ovf_cap - 1);
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

}  // namespace shn
