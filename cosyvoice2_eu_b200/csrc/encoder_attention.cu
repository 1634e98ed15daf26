// Relative-position multi-head attention of the UpsampleConformerEncoder on tcgen05
// (reference: cosyvoice/transformer/attention.py:249-330, rel_shift :225-247).
//   score[i,j] = ((q_i + u) . k_j + (q_i + v) . P[i - j]) / 8 ;  masked softmax ; . v
// The [T, 2T-1] `bd` matrix and its rel_shift are never materialised in memory.  Per (128-query, 64-key) tile the
// kernel issues two MMAs into TMEM:
//   AC = Qu K^T                       [128 x 64]
//   BD = Qv Pband^T                   [128 x 192], Pband = the 191 table rows (relative positions) the tile can see
// and the rel_shift becomes a per-row skew of BD: row r needs BD[r][r + 63 - jj] for key jj.  The warp-uniform part of
// that offset (32 * lane quarter) goes into the TMEM column address of the tcgen05.ld, the per-lane part (0..31) is a
// 5-stage register barrel shift.  Everything else is the flash-attention pipeline of attention.cu (online softmax, two
// threads per row, P through swizzled smem, O rescaled in TMEM).
//
// warp 8: TMA producer (Qu, Qv once; K, V^T and the position band through a 3-stage ring), warp 9: MMA issuer,
// warps 0-7: softmax.  TMEM: AC [0,64), BD [64,256), O [256,320) -> 512 columns, one CTA per SM.
#include "attention.cuh"
#include "common.cuh"
#include "host_util.h"

namespace cv2 {

static constexpr int kRKT = 64;                          // keys per tile
static constexpr int kRBand = 192;                       // table rows per tile (191 needed)
static constexpr int kRQBytes = 128 * 64 * 2;            // 16 KB each for Qu, Qv
static constexpr int kRKBytes = kRKT * 64 * 2;           // 8 KB
static constexpr int kRVBytes = 64 * kRKT * 2;           // 8 KB
static constexpr int kRPosBytes = kRBand * 64 * 2;       // 24 KB
static constexpr int kRStageBytes = kRKBytes + kRVBytes + kRPosBytes;
static constexpr int kRStages = 3;
static constexpr int kROffStage = 2 * kRQBytes;
static constexpr int kROffP = kROffStage + kRStages * kRStageBytes;
static constexpr int kRPBytes = 128 * kRKT * 2;
static constexpr int kROffBar = kROffP + kRPBytes;
static constexpr int kROffXch = kROffBar + 256;
static constexpr int kRelSmem = kROffXch + 3 * 1024;
static constexpr int kRelThreads = 10 * 32;
static constexpr uint32_t kRTmemAC = 0, kRTmemBD = 64, kRTmemO = 256;

__global__ void __launch_bounds__(kRelThreads, 1)
rel_attn_kernel(const __grid_constant__ CUtensorMap tmQu, const __grid_constant__ CUtensorMap tmQv,
                const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                const __grid_constant__ CUtensorMap tmPos, const RelAttnParams p) {
  const int t0 = blockIdx.x * 128;
  const int h = blockIdx.y;
  const int s = blockIdx.z;
  const int len = p.lens ? p.lens[s] : p.len_all;
  if (t0 >= len + p.halo || len <= 0) return;   // (an empty sequence -- an idle streaming slot -- has no key tile at all)
  const int sh = s * 8 + h;
  int kv_end = len;
  if (p.chunk > 0) kv_end = min(len, ((t0 + 127) / p.chunk + 1) * p.chunk);
  const int nkt = (kv_end + kRKT - 1) / kRKT;

  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kROffBar);
  uint64_t* q_full = bars;                      // 1
  uint64_t* kv_full = bars + 1;                 // [kRStages]
  uint64_t* kv_empty = kv_full + kRStages;      // [kRStages]
  uint64_t* s_full = kv_empty + kRStages;       // 1: AC and BD of a tile are in TMEM
  uint64_t* s_free = s_full + 1;                // 1 (256 arrivals): softmax holds AC/BD in registers
  uint64_t* p_full = s_free + 1;                // 1 (256 arrivals)
  uint64_t* pv_done = p_full + 1;               // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQu);
    tma_prefetch_desc(&tmQv);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmPos);
    mbar_init(q_full, 1);
    for (int i = 0; i < kRStages; i++) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 256);
    mbar_init(p_full, 256);
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    if (lane == 0) {
      mbar_expect_tx(q_full, 2 * kRQBytes);
      tma_load_3d(smem, &tmQu, q_full, 0, t0, sh);
      tma_load_3d(smem + kRQBytes, &tmQv, q_full, 0, t0, sh);
      for (int j = 0; j < nkt; j++) {
        const int st = j % kRStages;
        const uint32_t ph = (j / kRStages) & 1;
        uint8_t* stage = smem + kROffStage + st * kRStageBytes;
        mbar_wait(&kv_empty[st], ph ^ 1);
        mbar_expect_tx(&kv_full[st], kRStageBytes);
        tma_load_3d(stage, &tmK, &kv_full[st], 0, j * kRKT, sh);
        tma_load_3d(stage + kRKBytes, &tmV, &kv_full[st], j * kRKT, 0, sh);
        // table row c <-> relative position c - (Tmax - 1); the tile sees rel in [t0 - j0 - 63, t0 - j0 + 127]
        tma_load_2d(stage + kRKBytes + kRVBytes, &tmPos, &kv_full[st], h * 64, t0 - j * kRKT - (kRKT - 1) + p.Tmax - 1);
      }
    }
  } else if (warp == 9) {
    if (lane == 0) {
      constexpr uint32_t idesc_ac = umma_idesc_f16(128, kRKT, 0);
      constexpr uint32_t idesc_bd = umma_idesc_f16(128, kRBand, 0);
      constexpr uint32_t idesc_o = umma_idesc_f16(128, 64, 0);
      const uint64_t qu_desc = umma_smem_desc_sw128(smem_u32(smem));
      const uint64_t qv_desc = umma_smem_desc_sw128(smem_u32(smem + kRQBytes));
      const uint64_t p_desc = umma_smem_desc_sw128(smem_u32(smem + kROffP));
      auto issue_s = [&](int j) {
        const int st = j % kRStages;
        mbar_wait(&kv_full[st], (j / kRStages) & 1);
        tc_fence_after();
        const uint32_t stage = smem_u32(smem + kROffStage + st * kRStageBytes);
        const uint64_t k_desc = umma_smem_desc_sw128(stage);
        const uint64_t pos_desc = umma_smem_desc_sw128(stage + kRKBytes + kRVBytes);
#pragma unroll
        for (int k = 0; k < 4; k++)
          umma_f16(tmem_base + kRTmemAC, qu_desc + (uint64_t)(k * 2), k_desc + (uint64_t)(k * 2), idesc_ac, k != 0);
#pragma unroll
        for (int k = 0; k < 4; k++)
          umma_f16(tmem_base + kRTmemBD, qv_desc + (uint64_t)(k * 2), pos_desc + (uint64_t)(k * 2), idesc_bd, k != 0);
        umma_commit(s_full);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < nkt; j++) {
        if (j + 1 < nkt) {
          mbar_wait(s_free, j & 1);   // softmax j holds its logits in registers: AC/BD can be overwritten
          tc_fence_after();
          issue_s(j + 1);
        }
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        const int st = j % kRStages;
        const uint64_t v_desc = umma_smem_desc_sw128(smem_u32(smem + kROffStage + st * kRStageBytes + kRKBytes));
#pragma unroll
        for (int k = 0; k < kRKT / 16; k++)
          umma_f16(tmem_base + kRTmemO, p_desc + (uint64_t)(k * 2), v_desc + (uint64_t)(k * 2), idesc_o, (j | k) != 0);
        umma_commit(&kv_empty[st]);
        umma_commit(pv_done);
      }
    }
  } else {
    const int q = warp & 3;
    const int half = warp >> 2;
    const int r = q * 32 + lane;
    const int t = t0 + r;
    int kv_lim = len;
    if (p.chunk > 0) kv_lim = min(len, (t / p.chunk + 1) * p.chunk);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t bd_addr = lane_addr + kRTmemBD + (uint32_t)(32 * q + 32 - 32 * half);
    const float LOG2E = 1.4426950408889634f;
    float m = -INFINITY, l = 0.f;
    float* xch = reinterpret_cast<float*>(smem + kROffXch);
    uint8_t* prow = smem + kROffP + r * 128;
    for (int j = 0; j < nkt; j++) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int kbase = j * kRKT + half * 32;
      uint32_t sr[32], c[64];
      tmem_ld32(lane_addr + kRTmemAC + half * 32, sr);
      tmem_ld32(bd_addr, c);
      tmem_ld32(bd_addr + 32, c + 32);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(s_free);
      // skew: key i of this thread's half tile needs BD column (lane + 31 - i) of the 64 loaded ones
#pragma unroll
      for (int b = 16; b >= 1; b >>= 1) {
        const bool on = (lane & b) != 0;
#pragma unroll
        for (int k = 0; k < 32 + b - 1; k++) c[k] = on ? c[k + b] : c[k];
      }
#pragma unroll
      for (int i = 0; i < 32; i++) sr[i] = __float_as_uint(__uint_as_float(sr[i]) + __uint_as_float(c[31 - i]));
      if (kbase + 32 > kv_lim) {
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
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
      const float m_new = fmaxf(m, fmaxf(mx, xc[(half ^ 1) * 128 + r]));
      const float alpha = (m == -INFINITY) ? 0.f : fast_exp2((m - m_new) * LOG2E);
      const float mscaled = (m_new == -INFINITY) ? 0.f : m_new * LOG2E;
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
      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1);
        tc_fence_after();
      }
#pragma unroll
      for (int g = 0; g < 4; g++) {
        uint4 u;
        u.x = sr[g * 4 + 0]; u.y = sr[g * 4 + 1]; u.z = sr[g * 4 + 2]; u.w = sr[g * 4 + 3];
        *reinterpret_cast<uint4*>(prow + (((half * 4 + g) ^ (r & 7)) << 4)) = u;
      }
      if (j > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
        tmem_ld32(lane_addr + kRTmemO + half * 32, sr);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i++) sr[i] = __float_as_uint(__uint_as_float(sr[i]) * alpha);
        tmem_st32(lane_addr + kRTmemO + half * 32, sr);
        tmem_st_wait();
      }
      m = m_new;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    float* lx = xch + 512;
    lx[half * 128 + r] = l;
    asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
    const float lsum = l + lx[(half ^ 1) * 128 + r];
    const float inv = lsum > 0.f ? 1.f / lsum : 0.f;
    mbar_wait(pv_done, (nkt - 1) & 1);
    tc_fence_after();
    __half* dst = p.out + ((long long)s * p.T_alloc + t) * 512 + h * 64 + half * 32;
    const bool valid = t < len;
    uint32_t raw[32];
    tmem_ld32(lane_addr + kRTmemO + half * 32, raw);
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
  if (warp == 9) tmem_dealloc<512>(tmem_base);
}

void launch_rel_attn(const RelAttnParams& p, cudaStream_t stream) {
  CV2_CHECK(p.T_alloc % 128 == 0, "rel_attn: T_alloc %d not a multiple of 128", p.T_alloc);
  CV2_CHECK(p.R_alloc >= 2 * p.Tmax - 1, "rel_attn: position table has %d rows, need %d", p.R_alloc, 2 * p.Tmax - 1);
  static PerDeviceOnce once;
  once.run([] { CV2_CUDA(cudaFuncSetAttribute(rel_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRelSmem)); });
  const uint64_t SH = (uint64_t)p.S * 8;
  uint64_t dq[3] = {64, (uint64_t)p.T_alloc, SH};
  uint64_t sq[2] = {128, (uint64_t)p.T_alloc * 128};
  uint32_t bq[3] = {64, 128, 1};
  uint32_t bk[3] = {64, (uint32_t)kRKT, 1};
  CUtensorMap tmQu = make_tmap_16b(p.qu, 3, dq, sq, bq);
  CUtensorMap tmQv = make_tmap_16b(p.qv, 3, dq, sq, bq);
  CUtensorMap tmK = make_tmap_16b(p.k, 3, dq, sq, bk);
  uint64_t dv[3] = {(uint64_t)p.T_alloc, 64, SH};
  uint64_t sv[2] = {(uint64_t)p.T_alloc * 2, (uint64_t)p.T_alloc * 128};
  uint32_t bv[3] = {(uint32_t)kRKT, 64, 1};
  CUtensorMap tmV = make_tmap_16b(p.vt, 3, dv, sv, bv);
  uint64_t dp[2] = {512, (uint64_t)p.R_alloc};
  uint64_t sp[1] = {1024};
  uint32_t bp[2] = {64, (uint32_t)kRBand};
  CUtensorMap tmPos = make_tmap_16b(p.pos, 2, dp, sp, bp);
  dim3 grid(p.T_alloc / 128, 8, p.S);
  rel_attn_kernel<<<grid, kRelThreads, kRelSmem, stream>>>(tmQu, tmQv, tmK, tmV, tmPos, p);
  CV2_LAUNCH_CHECK();
}

}  // namespace cv2
