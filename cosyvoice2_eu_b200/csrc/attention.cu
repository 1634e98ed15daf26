// Flash-style masked attention for the CFM estimator (CausalConditionalDecoder transformer blocks:
// 8 heads x 64, softmax(q k^T / 8 + mask) v; reference: matcha/models/components/transformer.py:262-271 with the
// additive mask of cosyvoice/flow/decoder.py:439-443).  The mask is never materialised: it is generated from
// integers (valid length per sequence; optional block-causal chunk) inside the kernel.
//
// One CTA = 128 query rows of one (sequence, head), key tiles of 64.  warp 4: TMA producer (Q once, K / V^T tiles through a
// 2-stage ring), warp 5: tcgen05.mma issuer (warp-uniform code, one elected lane: S = Q K^T into TMEM, O += P V with P as the TMEM
// A operand; a single lane of divergent code pays ~100 cycles per MMA, profiles/r2/micro_mma_issue.txt), warps 0-3: online
// softmax, one thread per row.  Four CTAs per SM.  (Round 1 also carried four measured-and-rejected variants -- P through
// shared memory, double-buffered S at two CTAs per SM, 32-key steps, 2-SM pairs; their numbers are in profiles/README.md,
// the code is in the history before round 2.)
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

// measurement aid (profiles/attn_trace.py), compiled in only with -DCV2_ATTN_TRACE: the kernel is sensitive to every register
// (a run-time-null trace pointer alone cost 273 -> 347 us through 52 B of spills at the 80-register ceiling of four CTAs per SM)
#ifdef CV2_ATTN_TRACE
__device__ __forceinline__ void atrace(long long* tb, int& ti, int code) {
  if (tb && ti < 2040) {
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
    tb[ti++] = (t << 8) | code;
  }
}
#define ATRACE(code) atrace(tb, ti, code)
#else
#define ATRACE(code) do { } while (0)
#endif

__global__ void __launch_bounds__(k9Threads, 4)
flash_attn_v9_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  const int t0 = blockIdx.x * 128;
  const int h = blockIdx.y;
  // blocks are dispatched in blockIdx.z order: walking the sequences backwards puts the longest ones first when the caller
  // sorted the batch by ascending length (length bucketing), so the last wave is made of short CTAs
  const int s = p.reverse_seq ? p.S - 1 - (int)blockIdx.z : (int)blockIdx.z;
  const int len = p.lens ? p.lens[s] : p.len_all;
  if (t0 >= len + p.halo || len <= 0) return;
  if (p.lo && t0 < (p.lo[s] / 128) * 128) return;
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
  // (trace: a CTA in the middle of the grid, so that its SM is shared with three others in steady state)
#ifdef CV2_ATTN_TRACE
  long long* tb = (p.trace && blockIdx.x == 2 && blockIdx.y == 3 && blockIdx.z == (unsigned)(p.S / 3) && lane == 0 && (warp == 0 || warp == 5))
                      ? p.trace + (warp == 0 ? 0 : 2048) : nullptr;
  int ti = 0;
#endif

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
    mbar_init(p_full, 4);          // one arrival per softmax warp
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
    // warp-uniform control flow (descriptor arithmetic on the uniform datapath), one elected lane issues
    constexpr uint32_t idesc_s = umma_idesc_f16(128, kKT, 0);
    constexpr uint32_t idesc_o = umma_idesc_f16(128, 64, 0);
    const uint64_t q_desc = umma_smem_desc_sw128(smem_u32(smem));
    auto s_mmas = [&](int j) {   // S_j = Q K_j^T (the caller has waited for kv_full of tile j); elected lane only
      const uint64_t k_desc = umma_smem_desc_sw128(smem_u32(smem + k9OffK + (j % k9Stages) * kKBytes));
#pragma unroll
      for (int k = 0; k < 4; k++)
        umma_f16(tmem_base + k9TmemS, q_desc + (uint64_t)(k * 2), k_desc + (uint64_t)(k * 2), idesc_s, k != 0);
      umma_commit(s_full);
    };
    mbar_wait(q_full, 0);
    mbar_wait(&kv_full[0], 0);
    tc_fence_after();
    if (elect_one()) s_mmas(0);
    __syncwarp();
    for (int j = 0; j < nkt; j++) {
      // K_{j+1} / V_{j+1} only depend on PV(j-1): polled here, outside the softmax -> PV -> S chain
      if (j + 1 < nkt) mbar_wait(&kv_full[(j + 1) % k9Stages], ((j + 1) / k9Stages) & 1);
      ATRACE(20);
      mbar_wait(p_full, j & 1);   // softmax j has replaced S_j by P_j in tensor memory
      ATRACE(21);
      tc_fence_after();
      const int st = j % k9Stages;
      const uint64_t v_desc = umma_smem_desc_sw128(smem_u32(smem + k9OffV + st * kVBytes));
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kKT / 16; k++)
          umma_f16_ts(tmem_base + k9TmemO, tmem_base + k9TmemS + k * 8, v_desc + (uint64_t)(k * 2), idesc_o, (j | k) != 0);
        // S_{j+1} in order behind PV(j) (it overwrites P_j only after that was consumed), and first in line: the softmax chain
        // waits for it, the stage release below does not
        if (j + 1 < nkt) s_mmas(j + 1);
        else umma_commit(o_done);
        umma_commit(&kv_empty[st]);
      }
      __syncwarp();
      ATRACE(22);
    }
  } else {
    const int r = warp * 32 + lane;
    const int t = t0 + r;
    int kv_lim = len;
    if (p.chunk > 0) kv_lim = min(len, (t / p.chunk + 1) * p.chunk);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const float LOG2E = 1.4426950408889634f;
    // lazy rescale: the reference only moves when a tile's maximum exceeds it by 2^kLazy; probabilities then reach 2^kLazy, which a
    // 16-bit P (max 65504, relative precision independent of magnitude) and the fp32 accumulators hold without loss
    constexpr float kLazy = 13.f;
    float mref = 0.f, l = 0.f;     // reference max in log2 units
    for (int j = 0; j < nkt; j++) {
      ATRACE(1);
      mbar_wait(s_full, j & 1);     // also implies PV(j-1) has completed (same in-order pipe, commit covers prior MMAs)
      ATRACE(2);
      tc_fence_after();
      const int kbase = j * kKT;
      const bool edge = kbase + kKT > kv_lim;
      uint32_t sa[32];
      // pass 1: row maximum, chunk 1 first: chunk 0 (masked) is still in registers when pass 2 starts -- one tensor-memory round
      // trip less per key tile.  (Both 32-column loads in flight at once was measured no faster: 301.5 vs 297.0 us.)
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int c = 1; c >= 0; c--) {
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
      ATRACE(3);
      float alpha = 1.f;
      if (j == 0) {
        mref = mxl;
      } else if (mxl > mref + kLazy) {
        alpha = fast_exp2(mref - mxl);
        mref = mxl;
      }
      bool have0 = true;             // chunk 0 of S is in `sa` (warp-uniform)
      if (j > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
        have0 = false;               // (sa is the scratch of the rescale)
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
        if (c == 1 || !have0) {
          tmem_ld32(lane_addr + k9TmemS + c * 32, sa);
          tmem_ld_wait();
          if (edge) {
#pragma unroll
            for (int i = 0; i < 32; i++)
              if (kbase + c * 32 + i >= kv_lim) sa[i] = 0xff800000u;
          }
        }
        uint32_t pk[16];
        const float2 sc2 = make_float2(LOG2E, LOG2E), nm2 = make_float2(-mref, -mref);
#pragma unroll
        for (int i = 0; i < 32; i += 4) {   // packed fp32 pairs: one FFMA2 / FADD2 per two logits
          const float2 x01 = ffma2(make_float2(__uint_as_float(sa[i]), __uint_as_float(sa[i + 1])), sc2, nm2);
          const float2 x23 = ffma2(make_float2(__uint_as_float(sa[i + 2]), __uint_as_float(sa[i + 3])), sc2, nm2);
          const float2 e01 = make_float2(fast_exp2(x01.x), fast_exp2(x01.y));
          // (one pair of every 2nd / 4th group on the FMA pipe instead -- a cubic 2^x -- measured 276.2 / 274.5 against 273.5 us: the
          // exponentials do not bound this kernel, profiles/README.md)
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
      ATRACE(4);
      tmem_st_wait();
      ATRACE(5);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      ATRACE(6);
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

extern long long* g_ffn_trace_ptr();

void launch_flash_attn(const AttnParams& p_in, cudaStream_t stream) {
  AttnParams p = p_in;
  static const bool tr = getenv("CV2_TRACE_ATTN") != nullptr;
  p.trace = tr ? g_ffn_trace_ptr() : nullptr;
  CV2_CHECK(p.T_alloc % 128 == 0, "attention: T_alloc %d not a multiple of 128", p.T_alloc);
  static PerDeviceOnce once;
  once.run([] { CV2_CUDA(cudaFuncSetAttribute(flash_attn_v9_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, k9Smem)); });
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
  flash_attn_v9_kernel<<<grid, k9Threads, k9Smem, stream>>>(tmQ, tmK, tmV, p);
  CV2_LAUNCH_CHECK();
}

}  // namespace cv2
