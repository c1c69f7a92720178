// bruteforce_tc.cu — exhaustive top-k on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), any dimension (K is
// zero-padded to a multiple of 64), k <= 16.
//
// The dense part of ground-truth generation is the nq x n matrix of inner products.  It is formed by tcgen05.mma in
// bf16 with fp32 accumulation in TMEM; to keep fp32-grade accuracy every operand is split x = hi + lo (two bf16) and
// three products are accumulated (hi*hi + hi*lo + lo*hi, error ~2^-16 |q||x|).  The tensor-core pass only GENERATES
// candidates — per query the 32 best approximate scores of every slice of the base — and an fp32 pass re-ranks them
// exactly (same arithmetic and tie rule as csrc/bruteforce.cu).  Exactness is CERTIFIED per query, not assumed: a row that
// is not a candidate has an approximate score >= the 32nd of its slice, so its true distance is at least that threshold
// minus the error bound 2 (2^-16 + d 2^-23) |q| max|x|; the k-th exact distance must lie strictly below that for every full slice, and a
// query for which it does not (a slice that holds more than 32 - k of the true neighbours' near-ties) is answered again
// by the exact fp32 kernel of csrc/bruteforce.cu.
//
// CTA = 128 queries x a slice of base rows, 6 warps: warp 0 issues TMA loads (A = queries hi/lo, B = base rows hi/lo,
// 128-byte-swizzled K-major tiles of 64 bf16, 2 stages of 96 KB), warp 1 allocates TMEM (2 x 256 columns: the epilogue of
// tile j overlaps the MMAs of tile j+1) and issues the MMAs (M=128, N=256, K=16 per instruction) from one thread, warps
// 2-5 (one TMEM lane quadrant each; thread = query row) read the accumulators with tcgen05.ld and keep a sorted
// candidate list per query.
#include <cuda.h>
#include <cuda_bf16.h>

#include <cfloat>
#include <vector>

#include "engine.h"

namespace shn {
namespace {

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 2, CAND = 32;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;          // one bf16 tile
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;               // hi + lo of both operands
constexpr int TC_THREADS = 192;

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 32 consecutive columns of fp32 -> 32 registers per thread (thread = TMEM lane of the warp's quadrant)
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);          // start address
  d |= static_cast<uint64_t>(1) << 16;                             // leading byte offset (unused with swizzle): 1
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                     // stride byte offset: 8 rows x 128 B
  d |= static_cast<uint64_t>(1) << 46;                             // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;                             // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, N = 256, M = 128
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(BN >> 3) << 17) | (static_cast<uint32_t>(BM >> 4) << 24);

struct TcParams {
  uint32_t nq, slab_rows, rows_per_slice, k_blocks, list0;  // list0: first candidate-list slot of this slab
  uint64_t row0;                                            // global id of the slab's first row
  const float* xnorm;                                       // [slab_rows] |x|^2
  float* part_d;                                            // [lists][nq][CAND]
  uint32_t* part_i;
  int ip;
};

__global__ void __launch_bounds__(TC_THREADS, 1)
bf_tc_kernel(const __grid_constant__ CUtensorMap map_qh, const __grid_constant__ CUtensorMap map_ql,
             const __grid_constant__ CUtensorMap map_bh, const __grid_constant__ CUtensorMap map_bl, const TcParams p) {
  extern __shared__ unsigned char smem_raw[];
  // 128-byte-swizzled tiles need 1024-byte alignment
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* tiles = smem;                                  // STAGES x {A_hi, A_lo, B_hi, B_lo}
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* s_norm = reinterpret_cast<float*>(tmem_base_slot + 2);  // [2][BN]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t q0 = blockIdx.x * BM;
  const uint32_t r_begin = blockIdx.y * p.rows_per_slice;
  const uint32_t r_end = min(p.slab_rows, r_begin + p.rows_per_slice);
  const uint32_t n_tiles = r_end > r_begin ? (r_end - r_begin + BN - 1) / BN : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_base_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===== TMA producer (one thread) =====
    if (lane == 0) {
      uint32_t it = 0;
      for (uint32_t t = 0; t < n_tiles; ++t) {
        const int row = static_cast<int>(r_begin + t * BN);
        for (uint32_t kb = 0; kb < p.k_blocks; ++kb, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          unsigned char* st = tiles + s * STAGE_BYTES;
          mbar_expect_tx(&full[s], STAGE_BYTES);
          tma_load_2d(st, &map_qh, &full[s], static_cast<int>(kb * BK), static_cast<int>(q0));
          tma_load_2d(st + A_BYTES, &map_ql, &full[s], static_cast<int>(kb * BK), static_cast<int>(q0));
          tma_load_2d(st + 2 * A_BYTES, &map_bh, &full[s], static_cast<int>(kb * BK), row);
          tma_load_2d(st + 2 * A_BYTES + B_BYTES, &map_bl, &full[s], static_cast<int>(kb * BK), row);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      uint32_t it = 0;
      for (uint32_t t = 0; t < n_tiles; ++t) {
        const uint32_t a = t & 1, aph = (t >> 1) & 1;
        mbar_wait(&acc_empty[a], aph ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + a * BN;
        for (uint32_t kb = 0; kb < p.k_blocks; ++kb, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t base = smem_u32(tiles + s * STAGE_BYTES);
          const uint64_t a_hi = umma_desc(base), a_lo = umma_desc(base + A_BYTES);
          const uint64_t b_hi = umma_desc(base + 2 * A_BYTES), b_lo = umma_desc(base + 2 * A_BYTES + B_BYTES);
#pragma unroll
          for (uint32_t kk = 0; kk < BK / 16; ++kk) {
            const uint64_t adv = static_cast<uint64_t>((kk * 16 * 2) >> 4);  // 32 bytes per K step inside the swizzle atom
            tc_mma(tmem_d, a_hi + adv, b_hi + adv, kIdesc, (kb | kk) != 0 ? 1u : 0u);
            tc_mma(tmem_d, a_hi + adv, b_lo + adv, kIdesc, 1u);
            tc_mma(tmem_d, a_lo + adv, b_hi + adv, kIdesc, 1u);
          }
          tc_commit(&empty[s]);  // the stage is free once these MMAs have read it
        }
        tc_commit(&acc_full[a]);  // accumulator complete
      }
    }
  } else {
    // ===== epilogue: warps 2..5, TMEM lane quadrant = warp % 4, thread = query row =====
    const uint32_t quad = warp & 3;
    const uint32_t qrow = quad * 32 + lane;
    const uint32_t q = q0 + qrow;
    float cd[CAND];
    uint32_t ci[CAND];
#pragma unroll
    for (int j = 0; j < CAND; ++j) { cd[j] = FLT_MAX; ci[j] = kInvalid; }
    const int ep_tid = threadIdx.x - 64;  // 0..127
    for (uint32_t t = 0; t < n_tiles; ++t) {
      const uint32_t a = t & 1, aph = (t >> 1) & 1;
      const uint32_t row0 = r_begin + t * BN;
      // stage the |x|^2 of this tile (read by all 128 epilogue threads)
      if (!p.ip) {
        for (int j = ep_tid; j < BN; j += 128) s_norm[a * BN + j] = row0 + j < r_end ? __ldg(p.xnorm + row0 + j) : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(&acc_full[a], aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + a * BN + ((quad * 32) << 16);
      for (uint32_t c = 0; c < BN / 32; ++c) {
        float v[32];
        tc_ld32(taddr + c * 32, v);
        if (q < p.nq) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const uint32_t r = row0 + c * 32 + j;
            const float dist = p.ip ? -v[j] : s_norm[a * BN + c * 32 + j] - 2.f * v[j];
            if (r < r_end && dist < cd[CAND - 1]) {
              int pos = CAND - 1;
              while (pos > 0 && cd[pos - 1] > dist) { cd[pos] = cd[pos - 1]; ci[pos] = ci[pos - 1]; --pos; }
              cd[pos] = dist;
              ci[pos] = r;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[a]);
    }
    if (q < p.nq) {
      const size_t o = (static_cast<size_t>(p.list0 + blockIdx.y) * p.nq + q) * CAND;
      for (int j = 0; j < CAND; ++j) {
        p.part_d[o + j] = cd[j];
        p.part_i[o + j] = ci[j] == kInvalid ? kInvalid : static_cast<uint32_t>(p.row0 + ci[j]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// x = hi + lo in bf16 (rows zero-padded to dim_pad), |x|^2 per row, and the largest |x|^2 seen (for the error bound)
__global__ void split_bf16_kernel(const float* __restrict__ src, uint64_t rows, uint32_t dim, uint32_t dim_pad,
                                  __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, float* __restrict__ norm,
                                  float* __restrict__ max_norm2) {
  const int lane = threadIdx.x & 31;
  const uint64_t warps = static_cast<uint64_t>(gridDim.x) * (blockDim.x >> 5);
  float biggest = 0.f;
  for (uint64_t r = blockIdx.x * static_cast<uint64_t>(blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += warps) {
    float acc = 0.f;
    for (uint32_t e = lane; e < dim_pad; e += 32) {
      const float x = e < dim ? src[r * dim + e] : 0.f;
      const __nv_bfloat16 h = __float2bfloat16_rn(x);
      hi[r * dim_pad + e] = h;
      lo[r * dim_pad + e] = __float2bfloat16_rn(x - __bfloat162float(h));
      acc = fmaf(x, x, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    if (norm && lane == 0) norm[r] = acc;
    biggest = fmaxf(biggest, acc);
  }
  // non-negative floats order like their bit patterns
  if (lane == 0 && biggest > 0.f) atomicMax(reinterpret_cast<int*>(max_norm2), __float_as_int(biggest));
}

// Exact fp32 re-rank of the candidates of one query (a warp): distance with one fma per element in element order, ties
// by lower id — the arithmetic of bruteforce.cu — then the k best.
// Then the certificate (file header): uncertain[q] = 1 unless the k-th exact distance lies strictly below what any
// non-candidate row can have.
__global__ void rerank_kernel(const float* __restrict__ base, const float* __restrict__ queries, uint32_t nq, uint32_t dim, int ip,
                              const uint32_t* __restrict__ part_i, const float* __restrict__ part_d, uint32_t lists, uint32_t k,
                              const float* __restrict__ max_xnorm2, const float* __restrict__ qnorm2, uint32_t* __restrict__ out_i,
                              float* __restrict__ out_d, uint8_t* __restrict__ uncertain) {
  extern __shared__ unsigned char sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t q = blockIdx.x * (blockDim.x >> 5) + warp;
  float* ld = reinterpret_cast<float*>(sm) + static_cast<size_t>(warp) * 2 * k;
  uint32_t* li = reinterpret_cast<uint32_t*>(ld + k);
  if (q >= nq) return;
  for (uint32_t j = lane; j < k; j += 32) { ld[j] = FLT_MAX; li[j] = kInvalid; }
  __syncwarp();
  const float* qv = queries + static_cast<size_t>(q) * dim;
  const uint32_t total = lists * CAND;
  // per warp: a 32 x 32 transpose tile so that the 32 candidate rows are read with coalesced 128-byte loads while
  // every lane still sums ITS candidate in element order
  float* tile = reinterpret_cast<float*>(sm) + static_cast<size_t>(blockDim.x >> 5) * 2 * k + static_cast<size_t>(warp) * (32 * 33 + 32);
  float* qchunk = tile + 32 * 33;
  for (uint32_t c0 = 0; c0 < total; c0 += 32) {
    const uint32_t c = c0 + lane;
    uint32_t id = kInvalid;
    if (c < total) id = part_i[(static_cast<size_t>(c / CAND) * nq + q) * CAND + c % CAND];
    float acc = 0.f;
    for (uint32_t e0 = 0; e0 < dim; e0 += 32) {
      const uint32_t e = e0 + lane;
      qchunk[lane] = e < dim ? qv[e] : 0.f;
#pragma unroll 4
      for (int j = 0; j < 32; ++j) {
        const uint32_t idj = __shfl_sync(0xFFFFFFFFu, id, j);
        tile[j * 33 + lane] = (idj != kInvalid && e < dim) ? base[static_cast<size_t>(idj) * dim + e] : 0.f;
      }
      __syncwarp();
      const uint32_t lim = min(32u, dim - e0);
      if (ip) { for (uint32_t j = 0; j < lim; ++j) acc = __fmaf_rn(qchunk[j], tile[lane * 33 + j], acc); }
      else { for (uint32_t j = 0; j < lim; ++j) { const float t = __fsub_rn(qchunk[j], tile[lane * 33 + j]); acc = __fmaf_rn(t, t, acc); } }
      __syncwarp();
    }
    const float d = id == kInvalid ? FLT_MAX : (ip ? __fsub_rn(1.0f, acc) : acc);
    uint32_t mask = __ballot_sync(0xFFFFFFFFu, id != kInvalid);
    while (mask) {
      const int src = __ffs(mask) - 1;
      mask &= mask - 1;
      const float dd = __shfl_sync(0xFFFFFFFFu, d, src);
      const uint32_t ii = __shfl_sync(0xFFFFFFFFu, id, src);
      // position among (distance, id) pairs
      uint32_t pos = 0;
      for (uint32_t b = 0; b < k; b += 32) {
        const uint32_t j = b + lane;
        pos += __popc(__ballot_sync(0xFFFFFFFFu, j < k && (ld[j] < dd || (ld[j] == dd && li[j] < ii))));
      }
      if (pos >= k) continue;
      int hi = static_cast<int>(k) - 1;
      while (hi > static_cast<int>(pos)) {
        const int lo = max(static_cast<int>(pos), hi - 32);
        const int j = lo + lane;
        const bool act = j < hi;
        float td = 0.f; uint32_t ti = 0;
        if (act) { td = ld[j]; ti = li[j]; }
        __syncwarp();
        if (act) { ld[j + 1] = td; li[j + 1] = ti; }
        __syncwarp();
        hi = lo;
      }
      if (lane == 0) { ld[pos] = dd; li[pos] = ii; }
      __syncwarp();
    }
  }
  for (uint32_t j = lane; j < k; j += 32) {
    out_i[static_cast<size_t>(q) * k + j] = li[j];
    if (out_d) out_d[static_cast<size_t>(q) * k + j] = li[j] == kInvalid ? __int_as_float(0x7f800000) : ld[j];
  }
  // certificate: score = dist - |q|^2 (L2) resp. dist - 1 (IP); a full slice list hides only rows with score >= its last
  // entry, up to the error of the split-bf16 products
  __syncwarp();
  const float qn2 = qnorm2[q];
  // |score error| <= 2 (2^-16 + dim 2^-23) |q||x|: the terms the bf16 split drops, plus dim fp32 accumulation steps (worst case)
  const float eps = (6.1035156e-5f + static_cast<float>(dim) * 2.3841858e-7f) * sqrtf(qn2) * sqrtf(*max_xnorm2) + 1e-30f;
  const float kth = li[k - 1] == kInvalid ? __int_as_float(0x7f800000) : ld[k - 1];
  bool bad = false;
  for (uint32_t l = lane; l < lists; l += 32) {
    const size_t o = (static_cast<size_t>(l) * nq + q) * CAND;
    if (part_i[o + CAND - 1] == kInvalid) continue;  // the slice had fewer than CAND rows: all of them are candidates
    const float floor_d = (ip ? 1.0f : qn2) + part_d[o + CAND - 1] - eps;
    bad |= !(kth < floor_d);
  }
  bad = __any_sync(0xFFFFFFFFu, bad);
  if (lane == 0) uncertain[q] = bad ? 1 : 0;
}

__global__ void gather_queries_kernel(const float* __restrict__ src, const uint32_t* __restrict__ which, uint32_t count, uint32_t dim,
                                      float* __restrict__ dst) {
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < static_cast<uint64_t>(count) * dim;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x)
    dst[i] = src[static_cast<uint64_t>(which[i / dim]) * dim + i % dim];
}
__global__ void scatter_results_kernel(const uint32_t* __restrict__ src_i, const float* __restrict__ src_d, const uint32_t* __restrict__ which,
                                       uint32_t count, uint32_t k, uint32_t* __restrict__ dst_i, float* __restrict__ dst_d) {
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < static_cast<uint64_t>(count) * k;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint64_t o = static_cast<uint64_t>(which[i / k]) * k + i % k;
    dst_i[o] = src_i[i];
    if (dst_d) dst_d[o] = src_d[i];
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// [rows][dim] bf16, K-major; box = 64 elements x box_rows, 128-byte swizzle, out-of-bounds rows read as zero
bool make_map(CUtensorMap* m, const void* ptr, uint64_t rows, uint32_t dim, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {dim, rows};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(dim) * 2};
  const cuuint32_t box[2] = {BK, box_rows};
  const cuuint32_t elem[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

static unsigned long long g_tc_fallback_queries = 0;
unsigned long long bruteforce_tc_last_fallbacks() { return g_tc_fallback_queries; }

bool bruteforce_tc_supported(uint32_t dim, uint32_t k) { return dim >= 1 && dim <= 4096 && k <= CAND / 2 && encode_fn() != nullptr; }

cudaError_t bruteforce_tc_launch(const float* d_base, uint64_t n, const float* d_queries, uint32_t nq, uint32_t dim, bool ip,
                                 uint32_t k, uint32_t* d_ids, float* d_dists, int num_sms, cudaStream_t stream) {
  if (nq == 0) return cudaSuccess;
  if (!bruteforce_tc_supported(dim, k)) return cudaErrorNotSupported;
  g_tc_fallback_queries = 0;
  const uint32_t dim_pad = (dim + BK - 1) / BK * BK;  // zero-padded K: the padding adds exact zeros to every product
  const uint64_t slab_max = 4ull << 20;  // rows converted to bf16 hi/lo at a time
  const uint32_t qtiles = (nq + BM - 1) / BM;
  const uint32_t n_slabs = static_cast<uint32_t>((n + slab_max - 1) / slab_max);
  // One CTA per SM (199 KB of shared memory).  Slices cost epilogue time (every slice re-warms its candidate lists), so
  // take as many as fit in ONE wave and no more (measured: 13 slices at 79 query tiles were 2x slower than 1-2).
  const uint32_t slices = std::max<uint32_t>(1, std::min<uint32_t>(64, static_cast<uint32_t>(num_sms) / qtiles));
  const uint32_t lists = n_slabs * slices;

  __nv_bfloat16 *qh = nullptr, *ql = nullptr, *bh = nullptr, *bl = nullptr;
  float *xn = nullptr, *qn = nullptr, *part_d = nullptr, *maxn = nullptr;
  uint32_t* part_i = nullptr;
  uint8_t* uncertain = nullptr;
  const uint64_t slab_alloc = std::min<uint64_t>(n, slab_max);
  cudaError_t e = cudaMalloc(&qh, static_cast<size_t>(nq) * dim_pad * 2);
  if (e == cudaSuccess) e = cudaMalloc(&ql, static_cast<size_t>(nq) * dim_pad * 2);
  if (e == cudaSuccess) e = cudaMalloc(&bh, slab_alloc * dim_pad * 2);
  if (e == cudaSuccess) e = cudaMalloc(&bl, slab_alloc * dim_pad * 2);
  if (e == cudaSuccess) e = cudaMalloc(&xn, slab_alloc * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&qn, static_cast<size_t>(nq) * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&maxn, 2 * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&uncertain, nq);
  if (e == cudaSuccess) e = cudaMalloc(&part_d, static_cast<size_t>(lists) * nq * CAND * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&part_i, static_cast<size_t>(lists) * nq * CAND * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMemsetAsync(part_i, 0xFF, static_cast<size_t>(lists) * nq * CAND * sizeof(uint32_t), stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(maxn, 0, 2 * sizeof(float), stream);
  const size_t smem = STAGES * STAGE_BYTES + 1024 /*alignment*/ + 256 /*barriers*/ + 2 * BN * sizeof(float);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(bf_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e == cudaSuccess) {
    split_bf16_kernel<<<num_sms * 4, 256, 0, stream>>>(d_queries, nq, dim, dim_pad, qh, ql, qn, maxn + 1);
    CUtensorMap mqh, mql;
    if (!make_map(&mqh, qh, nq, dim_pad, BM) || !make_map(&mql, ql, nq, dim_pad, BM)) e = cudaErrorInvalidValue;
    for (uint32_t sl = 0; sl < n_slabs && e == cudaSuccess; ++sl) {
      const uint64_t r0 = sl * slab_max, rows = std::min<uint64_t>(slab_max, n - r0);
      split_bf16_kernel<<<num_sms * 8, 256, 0, stream>>>(d_base + r0 * dim, rows, dim, dim_pad, bh, bl, xn, maxn);
      CUtensorMap mbh, mbl;
      if (!make_map(&mbh, bh, rows, dim_pad, BN) || !make_map(&mbl, bl, rows, dim_pad, BN)) { e = cudaErrorInvalidValue; break; }
      TcParams p;
      p.nq = nq; p.slab_rows = static_cast<uint32_t>(rows); p.k_blocks = dim_pad / BK; p.list0 = sl * slices; p.row0 = r0;
      uint32_t per = static_cast<uint32_t>((rows + slices - 1) / slices);
      p.rows_per_slice = (per + BN - 1) / BN * BN;
      p.xnorm = xn; p.part_d = part_d; p.part_i = part_i; p.ip = ip ? 1 : 0;
      bf_tc_kernel<<<dim3(qtiles, slices), TC_THREADS, smem, stream>>>(mqh, mql, mbh, mbl, p);
      e = cudaGetLastError();
      if (e == cudaSuccess) e = cudaStreamSynchronize(stream);  // the slab buffers are reused
    }
  }
  std::vector<uint8_t> flags(nq);
  if (e == cudaSuccess) {
    const int warps = 4;
    rerank_kernel<<<(nq + warps - 1) / warps, warps * 32, warps * (2 * k + 32 * 33 + 32) * sizeof(float), stream>>>(
        d_base, d_queries, nq, dim, ip ? 1 : 0, part_i, part_d, lists, k, maxn, qn, d_ids, d_dists, uncertain);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(flags.data(), uncertain, nq, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  }
  cudaFree(qh); cudaFree(ql); cudaFree(bh); cudaFree(bl); cudaFree(xn); cudaFree(qn); cudaFree(maxn); cudaFree(uncertain);
  cudaFree(part_d); cudaFree(part_i);
  if (e != cudaSuccess) return e;
  // queries without a certificate: the exact fp32 kernel answers them again
  std::vector<uint32_t> which;
  for (uint32_t q = 0; q < nq; ++q) if (flags[q]) which.push_back(q);
  g_tc_fallback_queries = which.size();
  if (!which.empty()) {
    const uint32_t cnt = static_cast<uint32_t>(which.size());
    uint32_t *d_which = nullptr, *sub_i = nullptr;
    float *sub_q = nullptr, *sub_d = nullptr;
    e = cudaMalloc(&d_which, cnt * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&sub_q, static_cast<size_t>(cnt) * dim * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&sub_i, static_cast<size_t>(cnt) * k * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&sub_d, static_cast<size_t>(cnt) * k * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_which, which.data(), cnt * sizeof(uint32_t), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) {
      gather_queries_kernel<<<std::min<uint32_t>(cnt, 1024), 256, 0, stream>>>(d_queries, d_which, cnt, dim, sub_q);
      e = bruteforce_launch(d_base, n, sub_q, cnt, dim, ip, k, sub_i, sub_d, num_sms, stream);
    }
    if (e == cudaSuccess) {
      scatter_results_kernel<<<std::min<uint32_t>(cnt, 1024), 256, 0, stream>>>(sub_i, sub_d, d_which, cnt, k, d_ids, d_dists);
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(d_which); cudaFree(sub_q); cudaFree(sub_i); cudaFree(sub_d);
  }
  return e;
}

}  // namespace shn
