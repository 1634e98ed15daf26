// Flash-style masked attention for the CFM estimator (CausalConditionalDecoder transformer blocks:
// 8 heads x 64, softmax(q k^T / 8 + mask) v).  The mask is never materialised: it is generated from
// integers (valid length per sequence; optional block-causal chunk) inside the kernel.
//
// One CTA = 128 query rows of one (sequence, head), key tiles of 64.  warp 8: TMA producer (Q once, K/V tiles through
// a 2-stage ring), warp 9: tcgen05.mma issuer (S = Q K^T into TMEM, O += P V), warps 0-7: online softmax, two threads per row
// (32 logits each, in registers), P written 16-bit into 128B-swizzled smem as the A operand of the PV MMA;
// O accumulates in TMEM and is rescaled in place when the running max moves.
// The logits tile is double-buffered in TMEM so S(j+1) is computed while softmax j runs (the softmax warps otherwise
// spend most of their time waiting for the S MMA round trip); two CTAs per SM (100 KB smem, 256 TMEM columns each).
// The kernel is MUFU-bound in the limit (head_dim 64: one exp2 per 256 MMA FLOPs).
#include <stdlib.h>

#include "attention.cuh"
#include "common.cuh"
#include "epi_util.cuh"
#include "host_util.h"

namespace cv2 {

static constexpr int kKT = 64;                          // keys per tile
static constexpr int kQBytes = 128 * 64 * 2;            // 16 KB
static constexpr int kKBytes = kKT * 64 * 2;            // 64 keys x 64 d   (B operand of S, K-major)
static constexpr int kVBytes = 64 * kKT * 2;            // 64 d x 64 keys   (B operand of PV, K-major: v stored transposed)
static constexpr int kKVStages = 4;
static constexpr int kPBytes = 128 * kKT * 2;           // 128 rows x 64 keys, one swizzle atom column
static constexpr int kOffK = kQBytes;
static constexpr int kOffV = kOffK + kKVStages * kKBytes;
static constexpr int kOffP = kOffV + kKVStages * kVBytes;
static constexpr int kOffBar = kOffP + kPBytes;
static constexpr int kOffXch = kOffBar + 256;           // softmax exchange: 2x2x128 maxima + 2x128 row sums (floats)
static constexpr int kAttnSmem = kOffXch + 3 * 1024;    // 67.25 KB
static constexpr int kAttnThreads = 10 * 32;            // 8 softmax warps + TMA warp + MMA warp
static constexpr uint32_t kTmemS = 0, kTmemO = 128;     // S double-buffered (2 x 64 columns), O 64 columns (256 allocated)

__global__ void __launch_bounds__(kAttnThreads, 2)
flash_attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  const int t0 = blockIdx.x * 128;
  const int h = blockIdx.y;
  const int s = blockIdx.z;
  const int len = p.lens ? p.lens[s] : p.len_all;
  if (t0 >= len + p.halo) return;
  const int sh = s * p.heads + h;
  // number of key tiles this query tile can see
  int kv_end = len;
  if (p.chunk > 0) kv_end = min(len, ((t0 + 127) / p.chunk + 1) * p.chunk);
  const int nkt = (kv_end + kKT - 1) / kKT;

  extern __shared__ __align__(1024) uint8_t smem[];   // swizzle-128B tiles need 1024 B alignment
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* q_full = bars;              // 1
  uint64_t* kv_full = bars + 1;                 // [kKVStages]  (K and V of a stage land on the same barrier)
  uint64_t* kv_empty = kv_full + kKVStages;     // [kKVStages]
  uint64_t* s_full = kv_empty + kKVStages;      // [2]
  uint64_t* p_full = s_full + 2;                // 1 (256 arrivals)
  uint64_t* pv_done = p_full + 1;               // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < kKVStages; i++) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(&s_full[0], 1);
    mbar_init(&s_full[1], 1);
    mbar_init(p_full, 256);
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    if (lane == 0) {
      mbar_expect_tx(q_full, kQBytes);
      tma_load_3d(smem, &tmQ, q_full, 0, t0, sh);
      for (int j = 0; j < nkt; j++) {
        const int st = j % kKVStages;
        const uint32_t ph = (j / kKVStages) & 1;
        mbar_wait(&kv_empty[st], ph ^ 1);
        mbar_expect_tx(&kv_full[st], kKBytes + kVBytes);
        tma_load_3d(smem + kOffK + st * kKBytes, &tmK, &kv_full[st], 0, j * kKT, sh);
        tma_load_3d(smem + kOffV + st * kVBytes, &tmV, &kv_full[st], j * kKT, 0, sh);
      }
    }
  } else if (warp == 9) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_f16(128, kKT, 0);
      constexpr uint32_t idesc_o = umma_idesc_f16(128, 64, 0);
      const uint64_t q_desc = umma_smem_desc_sw128(smem_u32(smem));
      const uint64_t p_desc = umma_smem_desc_sw128(smem_u32(smem + kOffP));
      auto issue_s = [&](int j) {
        const int st = j % kKVStages;
        mbar_wait(&kv_full[st], (j / kKVStages) & 1);
        tc_fence_after();
        const uint64_t k_desc = umma_smem_desc_sw128(smem_u32(smem + kOffK + st * kKBytes));
#pragma unroll
        for (int k = 0; k < 4; k++)
          umma_f16(tmem_base + kTmemS + (j & 1) * 64, q_desc + (uint64_t)(k * 2), k_desc + (uint64_t)(k * 2), idesc_s, k != 0);
        umma_commit(&s_full[j & 1]);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      if (nkt > 1) issue_s(1);      // logits run one tile ahead of the softmax (double-buffered S)
      for (int j = 0; j < nkt; j++) {
        mbar_wait(p_full, j & 1);   // softmax j has consumed S_j and written P_j
        tc_fence_after();
        const int st = j % kKVStages;
        const uint64_t v_desc = umma_smem_desc_sw128(smem_u32(smem + kOffV + st * kVBytes));
#pragma unroll
        for (int k = 0; k < kKT / 16; k++)
          umma_f16(tmem_base + kTmemO, p_desc + (uint64_t)(k * 2), v_desc + (uint64_t)(k * 2), idesc_o, (j | k) != 0);
        umma_commit(&kv_empty[st]);
        umma_commit(pv_done);
        if (j + 2 < nkt) issue_s(j + 2);   // S buffer j&1 is free again
      }
    }
  } else {
    // ---- softmax: 8 warps, TWO threads per query row (warp w and w+4 share TMEM lane quarter w&3); thread `half`
    //      owns keys [32*half, 32*half+32) of every tile and columns [32*half, +32) of O ----
    const int q = warp & 3;
    const int half = warp >> 2;
    const int r = q * 32 + lane;
    const int t = t0 + r;
    int kv_lim = len;
    if (p.chunk > 0) kv_lim = min(len, (t / p.chunk + 1) * p.chunk);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const float LOG2E = 1.4426950408889634f;
    float m = -INFINITY, l = 0.f;
    float* xch = reinterpret_cast<float*>(smem + kOffXch);       // [2 parity][2 half][128 rows] partial maxima
    uint8_t* prow = smem + kOffP + r * 128;
    for (int j = 0; j < nkt; j++) {
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      const int kbase = j * kKT + half * 32;
      uint32_t sr[32];
      tmem_ld32(lane_addr + kTmemS + (j & 1) * 64 + half * 32, sr);
      tmem_ld_wait();
      if (kbase + 32 > kv_lim) {   // masking needed inside this half tile for this row
#pragma unroll
        for (int i = 0; i < 32; i++)
          if (kbase + i >= kv_lim) sr[i] = 0xff800000u;   // -inf
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        mx0 = fmaxf(mx0, __uint_as_float(sr[i]));
        mx1 = fmaxf(mx1, __uint_as_float(sr[i + 1]));
        mx2 = fmaxf(mx2, __uint_as_float(sr[i + 2]));
        mx3 = fmaxf(mx3, __uint_as_float(sr[i + 3]));
      }
      const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      float* xc = xch + (j & 1) * 256;
      xc[half * 128 + r] = mx;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");   // the two warps of this lane quarter
      const float m_new = fmaxf(m, fmaxf(mx, xc[(half ^ 1) * 128 + r]));
      const float alpha = (m == -INFINITY) ? 0.f : fast_exp2((m - m_new) * LOG2E);
      const float mscaled = (m_new == -INFINITY) ? 0.f : m_new * LOG2E;
      // probabilities in place (exp2(-inf) = 0 for masked keys), packed to fp16 pairs
      float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float e0 = fast_exp2(fmaf(__uint_as_float(sr[i]), LOG2E, -mscaled));
        const float e1 = fast_exp2(fmaf(__uint_as_float(sr[i + 1]), LOG2E, -mscaled));
        const float e2 = fast_exp2(fmaf(__uint_as_float(sr[i + 2]), LOG2E, -mscaled));
        const float e3 = fast_exp2(fmaf(__uint_as_float(sr[i + 3]), LOG2E, -mscaled));
        l0 += e0; l1 += e1; l2 += e2; l3 += e3;
        __half2 h0 = __floats2half2_rn(e0, e1), h1 = __floats2half2_rn(e2, e3);
        sr[i >> 1] = *reinterpret_cast<uint32_t*>(&h0);
        sr[(i >> 1) + 1] = *reinterpret_cast<uint32_t*>(&h1);
      }
      l = l * alpha + ((l0 + l1) + (l2 + l3));
      // PV of the previous tile must be complete before P is overwritten or O rescaled
      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1);
        tc_fence_after();
      }
      // P half tile, 16-bit, swizzled K-major: 4 x 16-byte groups per row
#pragma unroll
      for (int g = 0; g < 4; g++) {
        uint4 u;
        u.x = sr[g * 4 + 0]; u.y = sr[g * 4 + 1]; u.z = sr[g * 4 + 2]; u.w = sr[g * 4 + 3];
        *reinterpret_cast<uint4*>(prow + (((half * 4 + g) ^ (r & 7)) << 4)) = u;
      }
      if (j > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {   // rescale this thread's 32 columns of O
        tmem_ld32(lane_addr + kTmemO + half * 32, sr);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i++) sr[i] = __float_as_uint(__uint_as_float(sr[i]) * alpha);
        tmem_st32(lane_addr + kTmemO + half * 32, sr);
        tmem_st_wait();
      }
      m = m_new;
      fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // final: O / l -> 16-bit [S, T_alloc, heads*64]; the row sum is split over the two threads of the row
    float* lx = xch + 512;
    lx[half * 128 + r] = l;
    asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
    const float inv = 1.f / (l + lx[(half ^ 1) * 128 + r]);
    mbar_wait(pv_done, (nkt - 1) & 1);
    tc_fence_after();
    __half* dst = p.out + ((long long)s * p.T_alloc + t) * (p.heads * 64) + h * 64 + half * 32;
    const bool valid = t < len;
    uint32_t raw[32];
    tmem_ld32(lane_addr + kTmemO + half * 32, raw);
    tmem_ld_wait();
    uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      float f[8];
#pragma unroll
      for (int e = 0; e < 8; e++) f[e] = valid ? __uint_as_float(raw[i * 8 + e]) * inv : 0.f;
      __half2 h0 = __floats2half2_rn(f[0], f[1]);
      __half2 h1 = __floats2half2_rn(f[2], f[3]);
      __half2 h2 = __floats2half2_rn(f[4], f[5]);
      __half2 h3 = __floats2half2_rn(f[6], f[7]);
      uint4 u;
      u.x = *reinterpret_cast<uint32_t*>(&h0);
      u.y = *reinterpret_cast<uint32_t*>(&h1);
      u.z = *reinterpret_cast<uint32_t*>(&h2);
      u.w = *reinterpret_cast<uint32_t*>(&h3);
      d4[i] = u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<256>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------
// v8: one softmax thread per query row, P kept in tensor memory, logits double-buffered.
//   * each of the 128 softmax threads owns one row of the 128 x 64 logits tile: no cross-thread max / sum exchange;
//   * P (16-bit) overwrites the first 32 columns of its own S buffer in place and is the A operand of the PV MMA straight
//     from TMEM (tcgen05.mma with a tensor-memory A operand): no shared-memory P tile, no generic->async proxy fence, and
//     the P store never waits for the previous PV;
//   * S is double-buffered in TMEM and issued one tile ahead (S(j+2) right behind PV(j)), so the softmax threads do not wait
//     for the S MMA round trip;
//   * the running max is only raised when the tile max exceeds it by more than 2^8 (P stays < 2^8 in fp16; the row sum and
//     O use the same stale reference, so the result is exact), which removes almost all O rescales;
//   * 192 threads (4 softmax warps + TMA + MMA), 80 KB smem, 256 TMEM columns: two CTAs per SM.
//   (Measured and rejected: row sums from a ones-row appended to V^T (PV with N = 80) -- 5 % slower than 64 FADDs per tile.)
// ---------------------------------------------------------------------------------------------------------
static constexpr int k8Stages = 4;
static constexpr int k8OffK = kQBytes;
static constexpr int k8OffV = k8OffK + k8Stages * kKBytes;
static constexpr int k8VStage = kVBytes;
static constexpr int k8OffBar = k8OffV + k8Stages * k8VStage;
static constexpr int k8Smem = k8OffBar + 256;
static constexpr int k8Threads = 6 * 32;
static constexpr uint32_t k8TmemS = 0, k8TmemO = 128;

// MODE: 0 = every exp2 on the MUFU; 2 = every exp2 as an FMA-pipe polynomial; 3 / 4 = every 2nd / 4th group of four
// elements on the polynomial (splits the work between the XU and FMA pipes); 1 = no exp at all (timing experiments only)
template <int MODE>
__global__ void __launch_bounds__(k8Threads, 2)
flash_attn_v8_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  const int t0 = blockIdx.x * 128;
  const int h = blockIdx.y;
  const int s = blockIdx.z;
  const int len = p.lens ? p.lens[s] : p.len_all;
  if (t0 >= len + p.halo) return;
  const int sh = s * p.heads + h;
  int kv_end = len;
  if (p.chunk > 0) kv_end = min(len, ((t0 + 127) / p.chunk + 1) * p.chunk);
  const int nkt = (kv_end + kKT - 1) / kKT;

  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + k8OffBar);
  uint64_t* q_full = bars;                      // 1
  uint64_t* kv_full = bars + 1;                 // [k8Stages]
  uint64_t* kv_empty = kv_full + k8Stages;      // [k8Stages]
  uint64_t* s_full = kv_empty + k8Stages;       // [2]
  uint64_t* p_full = s_full + 2;                // 1 (128 arrivals)
  uint64_t* pv_done = p_full + 1;               // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < k8Stages; i++) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(&s_full[0], 1);
    mbar_init(&s_full[1], 1);
    mbar_init(p_full, 128);
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      mbar_expect_tx(q_full, kQBytes);
      tma_load_3d(smem, &tmQ, q_full, 0, t0, sh);
      for (int j = 0; j < nkt; j++) {
        const int st = j % k8Stages;
        const uint32_t ph = (j / k8Stages) & 1;
        mbar_wait(&kv_empty[st], ph ^ 1);
        mbar_expect_tx(&kv_full[st], kKBytes + kVBytes);
        tma_load_3d(smem + k8OffK + st * kKBytes, &tmK, &kv_full[st], 0, j * kKT, sh);
        tma_load_3d(smem + k8OffV + st * k8VStage, &tmV, &kv_full[st], j * kKT, 0, sh);
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_f16(128, kKT, 0);
      constexpr uint32_t idesc_o = umma_idesc_f16(128, 64, 0);
      const uint64_t q_desc = umma_smem_desc_sw128(smem_u32(smem));
      auto issue_s = [&](int j) {
        const int st = j % k8Stages;
        mbar_wait(&kv_full[st], (j / k8Stages) & 1);
        tc_fence_after();
        const uint64_t k_desc = umma_smem_desc_sw128(smem_u32(smem + k8OffK + st * kKBytes));
#pragma unroll
        for (int k = 0; k < 4; k++)
          umma_f16(tmem_base + k8TmemS + (j & 1) * 64, q_desc + (uint64_t)(k * 2), k_desc + (uint64_t)(k * 2), idesc_s, k != 0);
        umma_commit(&s_full[j & 1]);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      if (nkt > 1) issue_s(1);
      for (int j = 0; j < nkt; j++) {
        mbar_wait(p_full, j & 1);   // softmax j has replaced S_j by P_j in tensor memory
        tc_fence_after();
        const int st = j % k8Stages;
        const uint64_t v_desc = umma_smem_desc_sw128(smem_u32(smem + k8OffV + st * k8VStage));
#pragma unroll
        for (int k = 0; k < kKT / 16; k++)
          umma_f16_ts(tmem_base + k8TmemO, tmem_base + k8TmemS + (j & 1) * 64 + k * 8, v_desc + (uint64_t)(k * 2), idesc_o, (j | k) != 0);
        umma_commit(&kv_empty[st]);
        umma_commit(pv_done);
        if (j + 2 < nkt) issue_s(j + 2);   // in order behind PV(j): overwrites P_j only after it was consumed
      }
    }
  } else {
    const int r = warp * 32 + lane;
    const int t = t0 + r;
    int kv_lim = len;
    if (p.chunk > 0) kv_lim = min(len, (t / p.chunk + 1) * p.chunk);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const float LOG2E = 1.4426950408889634f;
    float mref = 0.f, l = 0.f;     // reference max in log2 units
    for (int j = 0; j < nkt; j++) {
      const uint32_t s_addr = lane_addr + k8TmemS + (j & 1) * 64;
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      const int kbase = j * kKT;
      uint32_t sr[64];
      tmem_ld32(s_addr, sr);
      tmem_ld32(s_addr + 32, sr + 32);
      tmem_ld_wait();
      if (kbase + kKT > kv_lim) {
#pragma unroll
        for (int i = 0; i < 64; i++)
          if (kbase + i >= kv_lim) sr[i] = 0xff800000u;   // -inf
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 64; i += 8) {
        mx0 = fmaxf(mx0, fmaxf(__uint_as_float(sr[i]), __uint_as_float(sr[i + 4])));
        mx1 = fmaxf(mx1, fmaxf(__uint_as_float(sr[i + 1]), __uint_as_float(sr[i + 5])));
        mx2 = fmaxf(mx2, fmaxf(__uint_as_float(sr[i + 2]), __uint_as_float(sr[i + 6])));
        mx3 = fmaxf(mx3, fmaxf(__uint_as_float(sr[i + 3]), __uint_as_float(sr[i + 7])));
      }
      const float mxl = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * LOG2E;
      float alpha = 1.f;
      if (j == 0) {
        mref = mxl;                       // tile 0 always holds a visible key for every row
      } else if (mxl > mref + 8.f) {      // lazy: P <= 2^8 otherwise
        alpha = fast_exp2(mref - mxl);
        mref = mxl;
      }
      bool pv_seen = j == 0;
      if (j > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
        mbar_wait(pv_done, (j - 1) & 1);
        tc_fence_after();
        pv_seen = true;
        uint32_t o[32];
#pragma unroll
        for (int c = 0; c < 2; c++) {
          tmem_ld32(lane_addr + k8TmemO + c * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i++) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st32(lane_addr + k8TmemO + c * 32, o);
        }
        l *= alpha;
      }
      float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll
      for (int c = 0; c < 2; c++) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float x0 = fmaf(__uint_as_float(sr[c * 32 + i]), LOG2E, -mref);
          const float x1 = fmaf(__uint_as_float(sr[c * 32 + i + 1]), LOG2E, -mref);
          const float x2 = fmaf(__uint_as_float(sr[c * 32 + i + 2]), LOG2E, -mref);
          const float x3 = fmaf(__uint_as_float(sr[c * 32 + i + 3]), LOG2E, -mref);
          const bool poly = MODE == 2 || (MODE == 3 && (i & 4)) || (MODE == 4 && (i & 12) == 12);
          float e0, e1, e2, e3;
          if (MODE == 1) {
            e0 = x0; e1 = x1; e2 = x2; e3 = x3;
          } else if (poly) {
            e0 = poly_exp2(x0); e1 = poly_exp2(x1); e2 = poly_exp2(x2); e3 = poly_exp2(x3);
          } else {
            e0 = fast_exp2(x0); e1 = fast_exp2(x1); e2 = fast_exp2(x2); e3 = fast_exp2(x3);
          }
          l0 += e0; l1 += e1; l2 += e2; l3 += e3;
          __half2 h0 = __floats2half2_rn(e0, e1), h1 = __floats2half2_rn(e2, e3);
          pk[i >> 1] = *reinterpret_cast<uint32_t*>(&h0);
          pk[(i >> 1) + 1] = *reinterpret_cast<uint32_t*>(&h1);
        }
        tmem_st16(s_addr + c * 16, pk);   // P columns [16c, 16c+16) of this S buffer (all 64 logits are in registers)
      }
      l += (l0 + l1) + (l2 + l3);
      if (!pv_seen) mbar_wait(pv_done, (j - 1) & 1);   // every thread observes every phase of pv_done (PV(j-1) is long done here)
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // final: O / l -> 16-bit [S, T_alloc, heads*64]
    const float inv = l > 0.f ? 1.f / l : 0.f;
    mbar_wait(pv_done, (nkt - 1) & 1);
    tc_fence_after();
    __half* dst = p.out + ((long long)s * p.T_alloc + t) * (p.heads * 64) + h * 64;
    const bool valid = t < len;
#pragma unroll
    for (int c = 0; c < 2; c++) {
      uint32_t raw[32];
      tmem_ld32(lane_addr + k8TmemO + c * 32, raw);
      tmem_ld_wait();
      uint4* d4 = reinterpret_cast<uint4*>(dst + c * 32);
#pragma unroll
      for (int i = 0; i < 4; i++) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; e++) f[e] = valid ? __uint_as_float(raw[i * 8 + e]) * inv : 0.f;
        __half2 h0 = __floats2half2_rn(f[0], f[1]);
        __half2 h1 = __floats2half2_rn(f[2], f[3]);
        __half2 h2 = __floats2half2_rn(f[4], f[5]);
        __half2 h3 = __floats2half2_rn(f[6], f[7]);
        uint4 u;
        u.x = *reinterpret_cast<uint32_t*>(&h0);
        u.y = *reinterpret_cast<uint32_t*>(&h1);
        u.z = *reinterpret_cast<uint32_t*>(&h2);
        u.w = *reinterpret_cast<uint32_t*>(&h3);
        d4[i] = u;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<256>(tmem_base);
}

static void launch_flash_attn_v8(const AttnParams& p, cudaStream_t stream) {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("CV2_ATTN_MODE");
    mode = e ? atoi(e) : 0;
    CV2_CUDA(cudaFuncSetAttribute(flash_attn_v8_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, k8Smem));
    CV2_CUDA(cudaFuncSetAttribute(flash_attn_v8_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, k8Smem));
    CV2_CUDA(cudaFuncSetAttribute(flash_attn_v8_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, k8Smem));
    CV2_CUDA(cudaFuncSetAttribute(flash_attn_v8_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, k8Smem));
    CV2_CUDA(cudaFuncSetAttribute(flash_attn_v8_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, k8Smem));
  }
  const uint64_t SH = (uint64_t)p.S * p.heads;
  uint64_t dq[3] = {64, (uint64_t)p.T_alloc, SH};
  uint64_t sq[2] = {128, (uint64_t)p.T_alloc * 128};
  uint32_t bq[3] = {64, 128, 1};
  uint32_t bk[3] = {64, (uint32_t)kKT, 1};
  CUtensorMap tmQ = make_tmap_16b(p.q, 3, dq, sq, bq);
  CUtensorMap tmK = make_tmap_16b(p.k, 3, dq, sq, bk);
  uint64_t dv[3] = {(uint64_t)p.T_alloc, 64, SH};
  uint64_t sv[2] = {(uint64_t)p.T_alloc * 2, (uint64_t)p.T_alloc * 128};
  uint32_t bv[3] = {(uint32_t)kKT, 64, 1};
  CUtensorMap tmV = make_tmap_16b(p.vt, 3, dv, sv, bv);
  dim3 grid(p.T_alloc / 128, p.heads, p.S);
  switch (mode) {
    case 1: flash_attn_v8_kernel<1><<<grid, k8Threads, k8Smem, stream>>>(tmQ, tmK, tmV, p); break;
    case 2: flash_attn_v8_kernel<2><<<grid, k8Threads, k8Smem, stream>>>(tmQ, tmK, tmV, p); break;
    case 3: flash_attn_v8_kernel<3><<<grid, k8Threads, k8Smem, stream>>>(tmQ, tmK, tmV, p); break;
    case 4: flash_attn_v8_kernel<4><<<grid, k8Threads, k8Smem, stream>>>(tmQ, tmK, tmV, p); break;
    default: flash_attn_v8_kernel<0><<<grid, k8Threads, k8Smem, stream>>>(tmQ, tmK, tmV, p); break;
  }
  CV2_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------------------
// v9: the v8 softmax (one thread per row, P in tensor memory, lazy rescale) at FOUR CTAs per SM.
//   The v8 profile shows two softmax warps per scheduler, each latency bound (MUFU issue spacing, TMEM round trips) and in
//   phase with each other; more resident warps is what fills the XU pipe.  To fit four CTAs: 128 TMEM columns per CTA
//   (S single-buffered with P aliased on top, O), 48 KB smem (2 K/V stages), and <= 80 registers per thread by walking
//   the 64 logits of a row in two 32-column passes (max, then probabilities) instead of holding the row.  The per-CTA
//   chain S(j) -> softmax(j) -> PV(j) -> S(j+1) is serial; the other three CTAs of the SM fill the gaps.
// ---------------------------------------------------------------------------------------------------------
static constexpr int k9Stages = 2;
static constexpr int k9OffK = kQBytes;
static constexpr int k9OffV = k9OffK + k9Stages * kKBytes;
static constexpr int k9OffBar = k9OffV + k9Stages * kVBytes;
static constexpr int k9Smem = k9OffBar + 128;
static constexpr int k9Threads = 6 * 32;
static constexpr uint32_t k9TmemS = 0, k9TmemO = 64;

__global__ void __launch_bounds__(k9Threads, 4)
flash_attn_v9_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  const int t0 = blockIdx.x * 128;
  const int h = blockIdx.y;
  // blocks are dispatched in blockIdx.z order: walking the sequences backwards puts the longest ones first when the caller
  // sorted the batch by ascending length (length bucketing), so the last wave is made of short CTAs
  const int s = p.reverse_seq ? p.S - 1 - (int)blockIdx.z : (int)blockIdx.z;
  const int len = p.lens ? p.lens[s] : p.len_all;
  if (t0 >= len + p.halo) return;
  const int sh = s * p.heads + h;
  int kv_end = len;
  if (p.chunk > 0) kv_end = min(len, ((t0 + 127) / p.chunk + 1) * p.chunk);
  const int nkt = (kv_end + kKT - 1) / kKT;

  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + k9OffBar);
  uint64_t* q_full = bars;                      // 1
  uint64_t* kv_full = bars + 1;                 // [k9Stages]
  uint64_t* kv_empty = kv_full + k9Stages;      // [k9Stages]
  uint64_t* s_full = kv_empty + k9Stages;       // 1
  uint64_t* p_full = s_full + 1;                // 1 (128 arrivals)
  uint64_t* o_done = p_full + 1;                // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < k9Stages; i++) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_done, 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc<128>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      mbar_expect_tx(q_full, kQBytes);
      tma_load_3d(smem, &tmQ, q_full, 0, t0, sh);
      for (int j = 0; j < nkt; j++) {
        const int st = j % k9Stages;
        const uint32_t ph = (j / k9Stages) & 1;
        mbar_wait(&kv_empty[st], ph ^ 1);
        mbar_expect_tx(&kv_full[st], kKBytes + kVBytes);
        tma_load_3d(smem + k9OffK + st * kKBytes, &tmK, &kv_full[st], 0, j * kKT, sh);
        tma_load_3d(smem + k9OffV + st * kVBytes, &tmV, &kv_full[st], j * kKT, 0, sh);
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_f16(128, kKT, 0);
      constexpr uint32_t idesc_o = umma_idesc_f16(128, 64, 0);
      const uint64_t q_desc = umma_smem_desc_sw128(smem_u32(smem));
      auto issue_s = [&](int j) {
        const int st = j % k9Stages;
        mbar_wait(&kv_full[st], (j / k9Stages) & 1);
        tc_fence_after();
        const uint64_t k_desc = umma_smem_desc_sw128(smem_u32(smem + k9OffK + st * kKBytes));
#pragma unroll
        for (int k = 0; k < 4; k++)
          umma_f16(tmem_base + k9TmemS, q_desc + (uint64_t)(k * 2), k_desc + (uint64_t)(k * 2), idesc_s, k != 0);
        umma_commit(s_full);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < nkt; j++) {
        mbar_wait(p_full, j & 1);   // softmax j has replaced S_j by P_j in tensor memory
        tc_fence_after();
        const int st = j % k9Stages;
        const uint64_t v_desc = umma_smem_desc_sw128(smem_u32(smem + k9OffV + st * kVBytes));
#pragma unroll
        for (int k = 0; k < kKT / 16; k++)
          umma_f16_ts(tmem_base + k9TmemO, tmem_base + k9TmemS + k * 8, v_desc + (uint64_t)(k * 2), idesc_o, (j | k) != 0);
        umma_commit(&kv_empty[st]);
        if (j + 1 < nkt) issue_s(j + 1);   // in order behind PV(j): S_{j+1} overwrites P_j only after it was consumed
        else umma_commit(o_done);
      }
    }
  } else {
    const int r = warp * 32 + lane;
    const int t = t0 + r;
    int kv_lim = len;
    if (p.chunk > 0) kv_lim = min(len, (t / p.chunk + 1) * p.chunk);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const float LOG2E = 1.4426950408889634f;
    float mref = 0.f, l = 0.f;     // reference max in log2 units
    for (int j = 0; j < nkt; j++) {
      mbar_wait(s_full, j & 1);     // also implies PV(j-1) has completed (same in-order pipe, commit covers prior MMAs)
      tc_fence_after();
      const int kbase = j * kKT;
      const bool edge = kbase + kKT > kv_lim;
      uint32_t sa[32];
      // pass 1: row maximum
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int c = 0; c < 2; c++) {
        tmem_ld32(lane_addr + k9TmemS + c * 32, sa);
        tmem_ld_wait();
        if (edge) {
#pragma unroll
          for (int i = 0; i < 32; i++)
            if (kbase + c * 32 + i >= kv_lim) sa[i] = 0xff800000u;
        }
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          mx0 = fmaxf(mx0, fmaxf(__uint_as_float(sa[i]), __uint_as_float(sa[i + 4])));
          mx1 = fmaxf(mx1, fmaxf(__uint_as_float(sa[i + 1]), __uint_as_float(sa[i + 5])));
          mx2 = fmaxf(mx2, fmaxf(__uint_as_float(sa[i + 2]), __uint_as_float(sa[i + 6])));
          mx3 = fmaxf(mx3, fmaxf(__uint_as_float(sa[i + 3]), __uint_as_float(sa[i + 7])));
        }
      }
      const float mxl = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * LOG2E;
      float alpha = 1.f;
      if (j == 0) {
        mref = mxl;
      } else if (mxl > mref + 8.f) {
        alpha = fast_exp2(mref - mxl);
        mref = mxl;
      }
      if (j > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll
        for (int c = 0; c < 2; c++) {
          tmem_ld32(lane_addr + k9TmemO + c * 32, sa);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i++) sa[i] = __float_as_uint(__uint_as_float(sa[i]) * alpha);
          tmem_st32(lane_addr + k9TmemO + c * 32, sa);
        }
        l *= alpha;
      }
      // pass 2: probabilities; P chunk c (16 columns) lands on S columns [16c, 16c+16), all consumed by then
      float2 la = make_float2(0.f, 0.f), lb = make_float2(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < 2; c++) {
        tmem_ld32(lane_addr + k9TmemS + c * 32, sa);
        tmem_ld_wait();
        if (edge) {
#pragma unroll
          for (int i = 0; i < 32; i++)
            if (kbase + c * 32 + i >= kv_lim) sa[i] = 0xff800000u;
        }
        uint32_t pk[16];
        const float2 sc2 = make_float2(LOG2E, LOG2E), nm2 = make_float2(-mref, -mref);
#pragma unroll
        for (int i = 0; i < 32; i += 4) {   // packed fp32 pairs: one FFMA2 / FADD2 per two logits
          const float2 x01 = ffma2(make_float2(__uint_as_float(sa[i]), __uint_as_float(sa[i + 1])), sc2, nm2);
          const float2 x23 = ffma2(make_float2(__uint_as_float(sa[i + 2]), __uint_as_float(sa[i + 3])), sc2, nm2);
          const float2 e01 = make_float2(fast_exp2(x01.x), fast_exp2(x01.y));
          const float2 e23 = make_float2(fast_exp2(x23.x), fast_exp2(x23.y));
          la = fadd2(la, e01);
          lb = fadd2(lb, e23);
          __half2 h0 = __floats2half2_rn(e01.x, e01.y), h1 = __floats2half2_rn(e23.x, e23.y);
          pk[i >> 1] = *reinterpret_cast<uint32_t*>(&h0);
          pk[(i >> 1) + 1] = *reinterpret_cast<uint32_t*>(&h1);
        }
        tmem_st16(lane_addr + k9TmemS + c * 16, pk);
      }
      l += (la.x + la.y) + (lb.x + lb.y);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    const float inv = l > 0.f ? 1.f / l : 0.f;
    mbar_wait(o_done, 0);
    tc_fence_after();
    __half* dst = p.out + ((long long)s * p.T_alloc + t) * (p.heads * 64) + h * 64;
    const bool valid = t < len;
#pragma unroll
    for (int c = 0; c < 2; c++) {
      uint32_t raw[32];
      tmem_ld32(lane_addr + k9TmemO + c * 32, raw);
      tmem_ld_wait();
      // lane-pair transposed 256-bit stores: 16 rows x 64 contiguous bytes per instruction (row-per-lane 16-byte stores drain at
      // half the SM's store-path rate, profiles/micro/st_path.cu)
      float w[32];
#pragma unroll
      for (int e = 0; e < 32; e++) w[e] = valid ? __uint_as_float(raw[e]) * inv : 0.f;
      tile_store_f16(dst - (long long)lane * (p.heads * 64) + c * 32, (long long)p.heads * 64, nullptr, lane, w);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<128>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------
// v11: v9 as CTA pairs with 2-SM MMAs (tcgen05.mma.cta_group::2, M = 256).
//   In cta_group::1 form the issue path costs ~115 cycles per MMA whatever its N (profiles/micro/mma_bubble.cu), so v9's eight
//   N = 64 MMAs per 64-key tile and four CTAs per SM keep the SM's one tensor-issue path ~75 % busy: it, not the MUFU, was the
//   limiter (s_full wait = a third of the softmax warps' time).  Here two CTAs of a cluster take two adjacent query tiles of
//   the same (sequence, head); ONE thread issues S and PV for both (half the instructions per row), each CTA loads its own Q
//   and HALF of every K / V^T tile (the pair shares the B operand), TMA bytes are credited to the leader's barriers, commits
//   are multicast, and the softmax threads of both CTAs arrive on the leader's p_full.
//   MEASURED: correct, but 364 us vs v9's 315 us at the bench shape -- the pair-wide p_full (256 threads on two SMs), the
//   cross-SM barrier latencies and the filler tiles of odd tile counts cost more than the halved MMA count saves.  Kept as an
//   experiment (CV2_ATTN_V11); v9 stays the default.
// ---------------------------------------------------------------------------------------------------------
static constexpr int k11Stages = 4;
static constexpr int k11StageBytes = 8192;              // 32 keys x 64 d (K half) + 32 d x 64 keys (V^T half)
static constexpr int k11OffKV = kQBytes;
static constexpr int k11OffBar = k11OffKV + k11Stages * k11StageBytes;
static constexpr int k11Smem = k11OffBar + 128;

__global__ void __launch_bounds__(k9Threads, 4)
flash_attn_v11_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  const uint32_t rank = cluster_ctarank();          // == blockIdx.x & 1
  const bool leader = rank == 0;
  const int t0_pair = (blockIdx.x >> 1) * 256;
  const int t0 = t0_pair + (int)rank * 128;
  const int h = blockIdx.y;
  // blocks are dispatched in blockIdx.z order: walking the sequences backwards puts the longest ones first when the caller
  // sorted the batch by ascending length (length bucketing), so the last wave is made of short CTAs
  const int s = p.reverse_seq ? p.S - 1 - (int)blockIdx.z : (int)blockIdx.z;
  const int len = p.lens ? p.lens[s] : p.len_all;
  if (t0_pair >= len + p.halo) return;               // pair-uniform: both CTAs leave together
  const int sh = s * p.heads + h;
  int kv_end = len;
  if (p.chunk > 0) kv_end = min(len, ((t0_pair + 255) / p.chunk + 1) * p.chunk);   // both CTAs walk the same key tiles
  const int nkt = (kv_end + kKT - 1) / kKT;

  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + k11OffBar);
  uint64_t* q_full = bars;                      // 1
  uint64_t* kv_full = bars + 1;                 // [k11Stages]
  uint64_t* kv_empty = kv_full + k11Stages;      // [k11Stages]
  uint64_t* s_full = kv_empty + k11Stages;       // 1
  uint64_t* p_full = s_full + 1;                // 1 (128 arrivals)
  uint64_t* o_done = p_full + 1;                // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < k11Stages; i++) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(s_full, 1);      // multicast commit
    mbar_init(p_full, 256);    // leader's: the softmax threads of both CTAs
    mbar_init(o_done, 1);      // multicast commit
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc2<128>(tmem_slot);
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      // both CTAs: own Q tile, own half of every K / V^T tile; bytes credited to the leader's barriers
      if (leader) mbar_expect_tx(q_full, 2 * kQBytes);
      tma2_load_3d(smem, &tmQ, q_full, 0, t0, sh);
      for (int j = 0; j < nkt; j++) {
        const int st = j % k11Stages;
        const uint32_t ph = (j / k11Stages) & 1;
        mbar_wait(&kv_empty[st], ph ^ 1);
        if (leader) mbar_expect_tx(&kv_full[st], 2 * k11StageBytes);
        tma2_load_3d(smem + k11OffKV + st * k11StageBytes, &tmK, &kv_full[st], 0, j * kKT + (int)rank * 32, sh);        // 32 keys x 64 d
        tma2_load_3d(smem + k11OffKV + st * k11StageBytes + 4096, &tmV, &kv_full[st], j * kKT, (int)rank * 32, sh);     // 32 d x 64 keys
      }
    }
  } else if (warp == 5) {
    if (lane == 0 && leader) {
      constexpr uint32_t idesc_s = umma_idesc_f16(256, kKT, 0);   // both q tiles x 64 keys
      constexpr uint32_t idesc_o = umma_idesc_f16(256, 64, 0);
      const uint64_t q_desc = umma_smem_desc_sw128(smem_u32(smem));
      auto issue_s = [&](int j) {
        const int st = j % k11Stages;
        mbar_wait(&kv_full[st], (j / k11Stages) & 1);
        const uint64_t k_desc = umma_smem_desc_sw128(smem_u32(smem + k11OffKV + st * k11StageBytes));
#pragma unroll
        for (int k = 0; k < 4; k++)
          umma2_f16(tmem_base + k9TmemS, q_desc + (uint64_t)(k * 2), k_desc + (uint64_t)(k * 2), idesc_s, k != 0);
        umma2_commit(s_full);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < nkt; j++) {
        mbar_wait(p_full, j & 1);   // the softmax threads of both CTAs have replaced S_j by P_j in tensor memory
        tc_fence_after();
        const int st = j % k11Stages;
        const uint64_t v_desc = umma_smem_desc_sw128(smem_u32(smem + k11OffKV + st * k11StageBytes + 4096));
#pragma unroll
        for (int k = 0; k < kKT / 16; k++)
          umma2_f16_ts(tmem_base + k9TmemO, tmem_base + k9TmemS + k * 8, v_desc + (uint64_t)(k * 2), idesc_o, (j | k) != 0);
        umma2_commit(&kv_empty[st]);
        if (j + 1 < nkt) issue_s(j + 1);   // in order behind PV(j): S_{j+1} overwrites P_j only after it was consumed
        else umma2_commit(o_done);
      }
    }
  } else {
    const int r = warp * 32 + lane;
    const int t = t0 + r;
    int kv_lim = len;
    if (p.chunk > 0) kv_lim = min(len, (t / p.chunk + 1) * p.chunk);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const float LOG2E = 1.4426950408889634f;
    float mref = 0.f, l = 0.f;     // reference max in log2 units
    for (int j = 0; j < nkt; j++) {
      mbar_wait(s_full, j & 1);     // also implies PV(j-1) has completed (same in-order pipe, commit covers prior MMAs)
      tc_fence_after();
      const int kbase = j * kKT;
      const bool edge = kbase + kKT > kv_lim;
      uint32_t sa[32];
      // pass 1: row maximum
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int c = 0; c < 2; c++) {
        tmem_ld32(lane_addr + k9TmemS + c * 32, sa);
        tmem_ld_wait();
        if (edge) {
#pragma unroll
          for (int i = 0; i < 32; i++)
            if (kbase + c * 32 + i >= kv_lim) sa[i] = 0xff800000u;
        }
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          mx0 = fmaxf(mx0, fmaxf(__uint_as_float(sa[i]), __uint_as_float(sa[i + 4])));
          mx1 = fmaxf(mx1, fmaxf(__uint_as_float(sa[i + 1]), __uint_as_float(sa[i + 5])));
          mx2 = fmaxf(mx2, fmaxf(__uint_as_float(sa[i + 2]), __uint_as_float(sa[i + 6])));
          mx3 = fmaxf(mx3, fmaxf(__uint_as_float(sa[i + 3]), __uint_as_float(sa[i + 7])));
        }
      }
      const float mxl = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * LOG2E;
      float alpha = 1.f;
      if (j == 0) {
        mref = mxl;
      } else if (mxl > mref + 8.f) {
        alpha = fast_exp2(mref - mxl);
        mref = mxl;
      }
      if (j > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll
        for (int c = 0; c < 2; c++) {
          tmem_ld32(lane_addr + k9TmemO + c * 32, sa);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i++) sa[i] = __float_as_uint(__uint_as_float(sa[i]) * alpha);
          tmem_st32(lane_addr + k9TmemO + c * 32, sa);
        }
        l *= alpha;
      }
      // pass 2: probabilities; P chunk c (16 columns) lands on S columns [16c, 16c+16), all consumed by then
      float2 la = make_float2(0.f, 0.f), lb = make_float2(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < 2; c++) {
        tmem_ld32(lane_addr + k9TmemS + c * 32, sa);
        tmem_ld_wait();
        if (edge) {
#pragma unroll
          for (int i = 0; i < 32; i++)
            if (kbase + c * 32 + i >= kv_lim) sa[i] = 0xff800000u;
        }
        uint32_t pk[16];
        const float2 sc2 = make_float2(LOG2E, LOG2E), nm2 = make_float2(-mref, -mref);
#pragma unroll
        for (int i = 0; i < 32; i += 4) {   // packed fp32 pairs: one FFMA2 / FADD2 per two logits
          const float2 x01 = ffma2(make_float2(__uint_as_float(sa[i]), __uint_as_float(sa[i + 1])), sc2, nm2);
          const float2 x23 = ffma2(make_float2(__uint_as_float(sa[i + 2]), __uint_as_float(sa[i + 3])), sc2, nm2);
          const float2 e01 = make_float2(fast_exp2(x01.x), fast_exp2(x01.y));
          const float2 e23 = make_float2(fast_exp2(x23.x), fast_exp2(x23.y));
          la = fadd2(la, e01);
          lb = fadd2(lb, e23);
          __half2 h0 = __floats2half2_rn(e01.x, e01.y), h1 = __floats2half2_rn(e23.x, e23.y);
          pk[i >> 1] = *reinterpret_cast<uint32_t*>(&h0);
          pk[(i >> 1) + 1] = *reinterpret_cast<uint32_t*>(&h1);
        }
        tmem_st16(lane_addr + k9TmemS + c * 16, pk);
      }
      l += (la.x + la.y) + (lb.x + lb.y);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive_leader(p_full);
    }
    const float inv = l > 0.f ? 1.f / l : 0.f;
    mbar_wait(o_done, 0);
    tc_fence_after();
    __half* dst = p.out + ((long long)s * p.T_alloc + t) * (p.heads * 64) + h * 64;
    const bool valid = t < len;
    const bool in_tensor = t < p.T_alloc;   // (an odd tile count is padded with a filler CTA)
#pragma unroll
    for (int c = 0; c < 2; c++) {
      uint32_t raw[32];
      tmem_ld32(lane_addr + k9TmemO + c * 32, raw);
      tmem_ld_wait();
      uint4* d4 = reinterpret_cast<uint4*>(dst + c * 32);
#pragma unroll
      for (int i = 0; i < 4; i++) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; e++) f[e] = valid ? __uint_as_float(raw[i * 8 + e]) * inv : 0.f;
        __half2 h0 = __floats2half2_rn(f[0], f[1]);
        __half2 h1 = __floats2half2_rn(f[2], f[3]);
        __half2 h2 = __floats2half2_rn(f[4], f[5]);
        __half2 h3 = __floats2half2_rn(f[6], f[7]);
        uint4 u;
        u.x = *reinterpret_cast<uint32_t*>(&h0);
        u.y = *reinterpret_cast<uint32_t*>(&h1);
        u.z = *reinterpret_cast<uint32_t*>(&h2);
        u.w = *reinterpret_cast<uint32_t*>(&h3);
        if (in_tensor) d4[i] = u;
      }
    }
  }

  tc_fence_before();
  cluster_sync();
  if (warp == 5) tmem_dealloc2<128>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------
// v10: v9 with the logits double-buffered inside the same 128 TMEM columns.
//   The v9 profile has the softmax warps waiting for S a third of the time (S(j+1) sits behind PV(j) because P_j lives
//   on top of S_j).  Here the softmax step is 32 keys: S0 [0,32), S1 [32,64), O [64,128); S(jj+2) is issued right behind
//   PV(jj), so the logits of step jj+1 are ready before softmax(jj) ends.  K / V^T still move as 64-key TMA tiles; a
//   step uses rows [32h, 32h+32) of the K tile (descriptor + 4 KB) and k-steps {2h, 2h+1} of the V^T tile.  One 32-column
//   pass per step (32 logits in registers), so the TMEM read traffic is half of v9's two-pass walk.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(k9Threads, 4)
flash_attn_v10_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  const int t0 = blockIdx.x * 128;
  const int h = blockIdx.y;
  const int s = blockIdx.z;
  const int len = p.lens ? p.lens[s] : p.len_all;
  if (t0 >= len + p.halo) return;
  const int sh = s * p.heads + h;
  int kv_end = len;
  if (p.chunk > 0) kv_end = min(len, ((t0 + 127) / p.chunk + 1) * p.chunk);
  const int nh = (kv_end + 31) / 32;            // 32-key softmax steps
  const int nkt = (nh + 1) / 2;                 // 64-key K / V^T tiles

  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + k9OffBar);
  uint64_t* q_full = bars;                      // 1
  uint64_t* kv_full = bars + 1;                 // [k9Stages]
  uint64_t* kv_empty = kv_full + k9Stages;      // [k9Stages]
  uint64_t* s_full = kv_empty + k9Stages;       // [2]
  uint64_t* p_full = s_full + 2;                // 1 (128 arrivals)
  uint64_t* pv_done = p_full + 1;               // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < k9Stages; i++) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(&s_full[0], 1);
    mbar_init(&s_full[1], 1);
    mbar_init(p_full, 128);
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc<128>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      mbar_expect_tx(q_full, kQBytes);
      tma_load_3d(smem, &tmQ, q_full, 0, t0, sh);
      for (int j = 0; j < nkt; j++) {
        const int st = j % k9Stages;
        const uint32_t ph = (j / k9Stages) & 1;
        mbar_wait(&kv_empty[st], ph ^ 1);
        mbar_expect_tx(&kv_full[st], kKBytes + kVBytes);
        tma_load_3d(smem + k9OffK + st * kKBytes, &tmK, &kv_full[st], 0, j * kKT, sh);
        tma_load_3d(smem + k9OffV + st * kVBytes, &tmV, &kv_full[st], j * kKT, 0, sh);
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_f16(128, 32, 0);
      constexpr uint32_t idesc_o = umma_idesc_f16(128, 64, 0);
      const uint64_t q_desc = umma_smem_desc_sw128(smem_u32(smem));
      auto issue_s = [&](int jj) {
        const int j = jj >> 1, hf = jj & 1;
        const int st = j % k9Stages;
        if (hf == 0) {
          mbar_wait(&kv_full[st], (j / k9Stages) & 1);
          tc_fence_after();
        }
        const uint64_t k_desc = umma_smem_desc_sw128(smem_u32(smem + k9OffK + st * kKBytes + hf * 4096));
#pragma unroll
        for (int k = 0; k < 4; k++)
          umma_f16(tmem_base + hf * 32, q_desc + (uint64_t)(k * 2), k_desc + (uint64_t)(k * 2), idesc_s, k != 0);
        umma_commit(&s_full[hf]);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      if (nh > 1) issue_s(1);
      for (int jj = 0; jj < nh; jj++) {
        mbar_wait(p_full, jj & 1);   // softmax jj has replaced S_jj by P_jj in tensor memory
        tc_fence_after();
        const int j = jj >> 1, hf = jj & 1;
        const int st = j % k9Stages;
        const uint64_t v_desc = umma_smem_desc_sw128(smem_u32(smem + k9OffV + st * kVBytes));
#pragma unroll
        for (int k = 0; k < 2; k++)
          umma_f16_ts(tmem_base + k9TmemO, tmem_base + hf * 32 + k * 8, v_desc + (uint64_t)((hf * 2 + k) * 2), idesc_o, (jj | k) != 0);
        if (hf == 1 || jj == nh - 1) umma_commit(&kv_empty[st]);
        umma_commit(pv_done);
        if (jj + 2 < nh) issue_s(jj + 2);   // in order behind PV(jj): overwrites P_jj only after it was consumed
      }
    }
  } else {
    const int r = warp * 32 + lane;
    const int t = t0 + r;
    int kv_lim = len;
    if (p.chunk > 0) kv_lim = min(len, (t / p.chunk + 1) * p.chunk);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const float LOG2E = 1.4426950408889634f;
    float mref = 0.f, l = 0.f;     // reference max in log2 units
    for (int jj = 0; jj < nh; jj++) {
      const uint32_t s_addr = lane_addr + (jj & 1) * 32;
      mbar_wait(&s_full[jj & 1], (jj >> 1) & 1);
      tc_fence_after();
      const int kbase = jj * 32;
      uint32_t sa[32];
      tmem_ld32(s_addr, sa);
      tmem_ld_wait();
      if (kbase + 32 > kv_lim) {
#pragma unroll
        for (int i = 0; i < 32; i++)
          if (kbase + i >= kv_lim) sa[i] = 0xff800000u;
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        mx0 = fmaxf(mx0, fmaxf(__uint_as_float(sa[i]), __uint_as_float(sa[i + 4])));
        mx1 = fmaxf(mx1, fmaxf(__uint_as_float(sa[i + 1]), __uint_as_float(sa[i + 5])));
        mx2 = fmaxf(mx2, fmaxf(__uint_as_float(sa[i + 2]), __uint_as_float(sa[i + 6])));
        mx3 = fmaxf(mx3, fmaxf(__uint_as_float(sa[i + 3]), __uint_as_float(sa[i + 7])));
      }
      const float mxl = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * LOG2E;
      float alpha = 1.f;
      if (jj == 0) {
        mref = mxl;                        // step 0 always holds a visible key for every row
      } else if (mxl > mref + 8.f) {
        alpha = fast_exp2(mref - mxl);
        mref = mxl;
      }
      bool pv_seen = jj == 0;
      if (jj > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
        mbar_wait(pv_done, (jj - 1) & 1);
        tc_fence_after();
        pv_seen = true;
        uint32_t o[16];
#pragma unroll 1
        for (int c = 0; c < 4; c++) {     // 16 columns at a time: the 32 logits stay in registers
          tmem_ld16(lane_addr + k9TmemO + c * 16, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; i++) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st16(lane_addr + k9TmemO + c * 16, o);
        }
        l *= alpha;
      }
      float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float e0 = fast_exp2(fmaf(__uint_as_float(sa[i]), LOG2E, -mref));
        const float e1 = fast_exp2(fmaf(__uint_as_float(sa[i + 1]), LOG2E, -mref));
        const float e2 = fast_exp2(fmaf(__uint_as_float(sa[i + 2]), LOG2E, -mref));
        const float e3 = fast_exp2(fmaf(__uint_as_float(sa[i + 3]), LOG2E, -mref));
        l0 += e0; l1 += e1; l2 += e2; l3 += e3;
        __half2 h0 = __floats2half2_rn(e0, e1), h1 = __floats2half2_rn(e2, e3);
        pk[i >> 1] = *reinterpret_cast<uint32_t*>(&h0);
        pk[(i >> 1) + 1] = *reinterpret_cast<uint32_t*>(&h1);
      }
      tmem_st16(s_addr, pk);
      l += (l0 + l1) + (l2 + l3);
      if (!pv_seen) mbar_wait(pv_done, (jj - 1) & 1);   // every thread observes every phase of pv_done
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    const float inv = l > 0.f ? 1.f / l : 0.f;
    mbar_wait(pv_done, (nh - 1) & 1);
    tc_fence_after();
    __half* dst = p.out + ((long long)s * p.T_alloc + t) * (p.heads * 64) + h * 64;
    const bool valid = t < len;
#pragma unroll
    for (int c = 0; c < 2; c++) {
      uint32_t raw[32];
      tmem_ld32(lane_addr + k9TmemO + c * 32, raw);
      tmem_ld_wait();
      uint4* d4 = reinterpret_cast<uint4*>(dst + c * 32);
#pragma unroll
      for (int i = 0; i < 4; i++) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; e++) f[e] = valid ? __uint_as_float(raw[i * 8 + e]) * inv : 0.f;
        __half2 h0 = __floats2half2_rn(f[0], f[1]);
        __half2 h1 = __floats2half2_rn(f[2], f[3]);
        __half2 h2 = __floats2half2_rn(f[4], f[5]);
        __half2 h3 = __floats2half2_rn(f[6], f[7]);
        uint4 u;
        u.x = *reinterpret_cast<uint32_t*>(&h0);
        u.y = *reinterpret_cast<uint32_t*>(&h1);
        u.z = *reinterpret_cast<uint32_t*>(&h2);
        u.w = *reinterpret_cast<uint32_t*>(&h3);
        d4[i] = u;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<128>(tmem_base);
}

static void launch_flash_attn_v11(const AttnParams& p, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    CV2_CUDA(cudaFuncSetAttribute(flash_attn_v11_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, k11Smem));
    configured = true;
  }
  const uint64_t SH = (uint64_t)p.S * p.heads;
  uint64_t dq[3] = {64, (uint64_t)p.T_alloc, SH};
  uint64_t sq[2] = {128, (uint64_t)p.T_alloc * 128};
  uint32_t bq[3] = {64, 128, 1};
  uint32_t bk[3] = {64, 32, 1};                          // this CTA's 32 of the tile's 64 keys
  CUtensorMap tmQ = make_tmap_16b(p.q, 3, dq, sq, bq);
  CUtensorMap tmK = make_tmap_16b(p.k, 3, dq, sq, bk);
  uint64_t dv[3] = {(uint64_t)p.T_alloc, 64, SH};
  uint64_t sv[2] = {(uint64_t)p.T_alloc * 2, (uint64_t)p.T_alloc * 128};
  uint32_t bv[3] = {(uint32_t)kKT, 32, 1};               // this CTA's 32 of the 64 head-dim rows of V^T
  CUtensorMap tmV = make_tmap_16b(p.vt, 3, dv, sv, bv);
  const int qt = p.T_alloc / 128;
  cudaLaunchConfig_t q = {};
  q.gridDim = dim3((qt + 1) / 2 * 2, p.heads, p.S);
  q.blockDim = dim3(k9Threads);
  q.dynamicSmemBytes = k11Smem;
  q.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  q.attrs = at;
  q.numAttrs = 1;
  CV2_CUDA(cudaLaunchKernelEx(&q, flash_attn_v11_kernel, tmQ, tmK, tmV, p));
  CV2_LAUNCH_CHECK();
}

static void launch_flash_attn_v9(const AttnParams& p, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    CV2_CUDA(cudaFuncSetAttribute(flash_attn_v9_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, k9Smem));
    configured = true;
  }
  const uint64_t SH = (uint64_t)p.S * p.heads;
  uint64_t dq[3] = {64, (uint64_t)p.T_alloc, SH};
  uint64_t sq[2] = {128, (uint64_t)p.T_alloc * 128};
  uint32_t bq[3] = {64, 128, 1};
  uint32_t bk[3] = {64, (uint32_t)kKT, 1};
  CUtensorMap tmQ = make_tmap_16b(p.q, 3, dq, sq, bq);
  CUtensorMap tmK = make_tmap_16b(p.k, 3, dq, sq, bk);
  uint64_t dv[3] = {(uint64_t)p.T_alloc, 64, SH};
  uint64_t sv[2] = {(uint64_t)p.T_alloc * 2, (uint64_t)p.T_alloc * 128};
  uint32_t bv[3] = {(uint32_t)kKT, 64, 1};
  CUtensorMap tmV = make_tmap_16b(p.vt, 3, dv, sv, bv);
  dim3 grid(p.T_alloc / 128, p.heads, p.S);
  static const bool use_v10 = getenv("CV2_ATTN_V10") != nullptr;   // measured 7 % slower than v9: twice the barrier round trips per key
  if (use_v10) {
    static bool configured10 = false;
    if (!configured10) {
      CV2_CUDA(cudaFuncSetAttribute(flash_attn_v10_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, k9Smem));
      configured10 = true;
    }
    flash_attn_v10_kernel<<<grid, k9Threads, k9Smem, stream>>>(tmQ, tmK, tmV, p);
  } else {
    flash_attn_v9_kernel<<<grid, k9Threads, k9Smem, stream>>>(tmQ, tmK, tmV, p);
  }
  CV2_LAUNCH_CHECK();
}

void launch_flash_attn(const AttnParams& p, cudaStream_t stream) {
  CV2_CHECK(p.T_alloc % 128 == 0, "attention: T_alloc %d not a multiple of 128", p.T_alloc);
  static const bool use_v6 = getenv("CV2_ATTN_V6") != nullptr;   // two threads per row, P through smem, 2 CTAs/SM
  static const bool use_v8 = getenv("CV2_ATTN_V8") != nullptr;   // one thread per row, P in TMEM, S double-buffered, 2 CTAs/SM
  static const bool use_v11 = getenv("CV2_ATTN_V11") != nullptr;   // 2-SM pairing of v9: measured 16 % slower (364 vs 315 us)
  if (use_v8) return launch_flash_attn_v8(p, stream);
  if (use_v11) return launch_flash_attn_v11(p, stream);
  if (!use_v6) return launch_flash_attn_v9(p, stream);
  static bool configured = false;
  if (!configured) {
    CV2_CUDA(cudaFuncSetAttribute(flash_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
    configured = true;
  }
  const uint64_t SH = (uint64_t)p.S * p.heads;
  uint64_t dq[3] = {64, (uint64_t)p.T_alloc, SH};
  uint64_t sq[2] = {128, (uint64_t)p.T_alloc * 128};
  uint32_t bq[3] = {64, 128, 1};
  uint32_t bk[3] = {64, (uint32_t)kKT, 1};
  CUtensorMap tmQ = make_tmap_16b(p.q, 3, dq, sq, bq);
  CUtensorMap tmK = make_tmap_16b(p.k, 3, dq, sq, bk);
  uint64_t dv[3] = {(uint64_t)p.T_alloc, 64, SH};
  uint64_t sv[2] = {(uint64_t)p.T_alloc * 2, (uint64_t)p.T_alloc * 128};
  uint32_t bv[3] = {(uint32_t)kKT, 64, 1};
  CUtensorMap tmV = make_tmap_16b(p.vt, 3, dv, sv, bv);
  dim3 grid(p.T_alloc / 128, p.heads, p.S);
  flash_attn_kernel<<<grid, kAttnThreads, kAttnSmem, stream>>>(tmQ, tmK, tmV, p);
  CV2_LAUNCH_CHECK();
}

}  // namespace cv2
