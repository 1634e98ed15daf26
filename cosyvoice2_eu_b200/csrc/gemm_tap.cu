// Tap GEMM kernel (see gemm_tap.cuh).  Persistent, warp-specialised:
//   warps 0..E-1  epilogue (E = 16, 8 for BN = 64): see below
//   warp E        TMA producer: A/B k-blocks into a 192 KB 128B-swizzled smem ring (mbarrier full/empty)
//   warp E+1      tcgen05.mma issuer (kind::f16, fp32 accumulate) into a DOUBLE-BUFFERED TMEM accumulator
//   (the two issuing warps carry the highest warp ids: the scheduler favours high ids, and an issuer that loses arbitration
//    against the math-heavy epilogue warps paces the tensor pipe)
//   epilogue: 8 warps, two per TMEM lane quarter (each thread owns half a row), so the epilogue of tile i
//               overlaps the MMAs of tile i+1; LayerNorm statistics are combined between the two half-row threads
//               through shared memory.
// Each CTA walks tiles blockIdx.x, +gridDim.x, ... in (n fastest, then t, then sequence) order and skips tiles that lie
// entirely in a sequence's padding.
#include "common.cuh"
#include "epi_util.cuh"
#include "gemm_tap.cuh"
#include "host_util.h"

namespace cv2 {

static constexpr int kTileM = 128;
static constexpr int kKBlock = 64;                       // 64 x 16-bit = one 128B swizzle row
static constexpr int kABytes = kTileM * kKBlock * 2;     // 16 KB

template <int BN, int CS = 1>
struct GemmSmem {
  static constexpr int kParts = BN >= 128 ? 4 : 2;        // column parts of a tile = epilogue warps per TMEM lane quarter
  static constexpr int kEpiWarps = 4 * kParts;            // 16 (8 for BN = 64): enough warps to hide TMEM / global / MUFU latency
  static constexpr int kThreads = 64 + kEpiWarps * 32;    // + TMA warp + MMA warp
  // BN = 256: the 16 epilogue warps work as two independent groups of 8, group g draining accumulator buffer g (tiles of
  // parity g).  One SM's store path drains ~32 B/clk however the stores are issued (profiles/micro/st_path.cu), and with all 16
  // warps on one tile their accumulator-read / convert phases and their store phases alternated; two groups on two tiles run
  // out of phase, so the store path stays busy while the other group computes.
  static constexpr int kGroups = BN >= 256 ? 2 : 1;
  static constexpr int kBBytes = (BN / CS) * kKBlock * 2;   // CS = 2 (2-SM MMA): every CTA stages half of the weight tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  // the GEMMs of this path are short-K (K = 256..1536) and TMA-latency bound: what matters is bytes in flight.  The epilogue
  // needs no shared memory (256-bit global accesses), so the ring takes all of it: 192 KB = 4 / 6 / 8 stages.
  static constexpr int kStages = (192 * 1024) / kStageBytes;
  static constexpr int kRedOff = kStages * kStageBytes;   // 2 x 4 x [kParts][128] floats for LayerNorm exchanges
  static constexpr int kBarOff = kRedOff + 2 * 4 * kParts * 128 * 4;   // double-buffered by tile parity (one barrier per LayerNorm)
  // per-column parameter vectors of the epilogue, staged once per CTA (the streaming residual / output traffic keeps evicting
  // them from the small L1 that is left next to a 192 KB ring: in the clock64 trace every re-read cost an L2 round trip)
  static constexpr int kVecBias = 2048, kVecLn = 256;     // bias [N <= 2048]; LayerNorm gamma/beta + emitted-LN gamma/beta [N <= 256]
  static constexpr int kVecOff = kBarOff + 512;           // + barriers
  static constexpr int kTotal = kVecOff + (kVecBias + 4 * kVecLn) * 4;
};

// conv mode (GemmParams::tmA_halo): ring of [192 x 64] A boxes + the resident weight matrix
template <int BN>
struct GemmSmemConv {
  static constexpr int kParts = BN >= 128 ? 4 : 2;
  static constexpr int kEpiWarps = 4 * kParts;
  static constexpr int kThreads = 64 + kEpiWarps * 32;
  static constexpr int kGroups = 1;   // (two epilogue groups, one per accumulator buffer, were measured neutral here: 22.0 vs 22.3 ms)
  static constexpr int kStageBytes = kConvHaloRows * kKBlock * 2;   // 24 KB
  static constexpr int kStages = 4;
  static constexpr int kWOff = kStages * kStageBytes;               // resident weights: [ntaps * kb_per_tap][BN x 64]
  static constexpr int kWBytes = 96 * 1024;
  static constexpr int kRedOff = kWOff + kWBytes;
  static constexpr int kBarOff = kRedOff + 2 * 4 * kParts * 128 * 4;
  static constexpr int kVecBias = 2048, kVecLn = 256;
  static constexpr int kVecOff = kBarOff + 512;
  static constexpr int kTotal = kVecOff + (kVecBias + 4 * kVecLn) * 4;
};

template <int NT>
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }

struct TileCoord {
  int n0, t0, s, len;
};
// Work unit u -> coordinates.  A unit is one n tile x CS row tiles (CS = cluster size): CTA `rank` of the cluster takes row tile
// rg * CS + rank, all CTAs of the cluster share the n tile (= the multicast weight operand).  Returns false when the tile lies
// entirely in the padding (CS = 1, no compact list); `valid` is false for the filler tile of an odd tail (computed, not stored).
template <int CS>
__device__ __forceinline__ bool tile_coord(const GemmParams& p, int u, int n_tiles, int t_tiles, int BN, int rank, int row_count,
                                           TileCoord& c, bool& valid) {
  const int n = u % n_tiles;
  const int rg = u / n_tiles;
  valid = true;
  if (p.tile_list) {   // compacted list: every entry is active
    int idx = rg * CS + rank;
    if (CS > 1 && idx >= row_count) {
      valid = false;
      idx = 0;
    }
    c.s = __ldg(p.tile_list + 2 * idx);
    c.t0 = __ldg(p.tile_list + 2 * idx + 1);
    c.n0 = n * BN;
    c.len = p.lens ? __ldg(p.lens + c.s) : p.len_all;
    return true;
  }
  const int tt = rg % t_tiles;
  c.s = rg / t_tiles;
  c.n0 = n * BN;
  c.t0 = tt * kTileM;
  c.len = p.lens ? __ldg(p.lens + c.s) : p.len_all;
  return c.t0 < c.len + p.halo;
}

__device__ __forceinline__ void gtrace(long long* tb, int& ti, int code, float dep = 0.f) {
  if (tb && ti < 4000) {
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) : "f"(dep) : "memory");
    tb[ti++] = (t << 8) | code;
  }
}

// Compile-time epilogue configuration.  Every field is -1 (decided at run time from GemmParams: the generic kernel) or a
// fixed value; fixing them removes the untaken branches from the instruction stream (the generic epilogue is ~160 KB of
// SASS and was instruction-cache bound).  Hot estimator / HiFT epilogues get their own instantiation (see kSpecs).
template <int ACT_, int BIAS_, int LN_, int ROWVEC_, int MASK_, int RES_, int RES2_, int OUT32_, int ACCUM_, int FLAT_,
          int SCALE_, int EMIT0_, int EMIT1_, int EMIT2_, int QKV_>
struct EpiCfg {
  static constexpr int ACT = ACT_, BIAS = BIAS_, LN = LN_, ROWVEC = ROWVEC_, MASK = MASK_, RES = RES_, RES2 = RES2_,
                       OUT32 = OUT32_, ACCUM = ACCUM_, FLAT = FLAT_, SCALE = SCALE_, EMIT0 = EMIT0_, EMIT1 = EMIT1_,
                       EMIT2 = EMIT2_, QKV = QKV_;
};
using EpiGeneric = EpiCfg<-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1>;
#define CFGB(field, rt) (Cfg::field < 0 ? (rt) : (Cfg::field != 0))

// CS = 2: 2-SM form.  A cluster of two CTAs works on two row tiles of the same n tile; ONE thread (CTA rank 0) issues every MMA
// for both SMs with tcgen05.mma.cta_group::2 (M = 256), each CTA stages its own A tile and HALF of every weight k-block (the
// pair shares B), TMA bytes of both CTAs are credited to the leader's full barrier, the leader's commits are multicast to both
// CTAs (stage free, accumulator full) and the peer's epilogue warps release accumulators on the leader's barrier.
// Why: in cta_group::1 form these short-K GEMMs run the tensor pipe at ~55 % (issue path: profiles/micro/mma_bubble.cu); the
// 2-SM form does twice the work per instruction and sustains the nominal rate (profiles/micro/mma_2cta.cu).
template <int BN, int CS, bool CONV>
struct SmemOf { using type = GemmSmem<BN, CS>; };
template <int BN, int CS>
struct SmemOf<BN, CS, true> { using type = GemmSmemConv<BN>; };

template <int BN, class Cfg, int CS, bool CONV = false>
__global__ void __launch_bounds__(GemmSmem<BN, CS>::kThreads, 1)
gemm_tap_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using SM = typename SmemOf<BN, CS, CONV>::type;
  static_assert(!CONV || CS == 1, "conv mode is a 1-SM form (its 2-SM form was built and measured no faster: 23.0 vs 22.3 ms)");
  constexpr int kStages = SM::kStages;
  constexpr int kEpiWarps = SM::kEpiWarps;
  constexpr int kParts = SM::kParts;
  extern __shared__ __align__(1024) uint8_t smem[];   // 1024 B alignment for the 128B-swizzle atoms; keeps STS/LDS addressing
  float* red = reinterpret_cast<float*>(smem + SM::kRedOff);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SM::kBarOff);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint64_t* w_bar = reinterpret_cast<uint64_t*>(smem + SM::kBarOff + 256);   // conv mode: resident weights have landed

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_it = p.ntaps * p.kb_per_tap;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int t_tiles = p.T_alloc / kTileM;
  const int row_count = p.tile_list ? __ldg(p.tile_count) : t_tiles * p.S;
  const int total_tiles = n_tiles * ((row_count + CS - 1) / CS);    // work units
  const int rank = CS > 1 ? (int)cluster_ctarank() : 0;
  const int unit0 = blockIdx.x / CS, unit_step = gridDim.x / CS;
  const bool leader = rank == 0;
  constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;   // 128 / 256 / 512

  float* vec_bias = reinterpret_cast<float*>(smem + SM::kVecOff);
  float* vec_lng = vec_bias + SM::kVecBias;
  float* vec_lnb = vec_lng + SM::kVecLn;
  float* vec_eg = vec_lnb + SM::kVecLn;
  float* vec_eb = vec_eg + SM::kVecLn;
  const bool c_bias = p.bias != nullptr && p.N <= SM::kVecBias;
  const bool c_ln = p.ln != 0 && p.N <= SM::kVecLn;
  int e_ln = -1;
#pragma unroll
  for (int e = 2; e >= 0; e--)
    if ((e == 0 ? (Cfg::EMIT0 < 0 ? p.emit[0].kind : Cfg::EMIT0) : e == 1 ? (Cfg::EMIT1 < 0 ? p.emit[1].kind : Cfg::EMIT1)
                                                                           : (Cfg::EMIT2 < 0 ? p.emit[2].kind : Cfg::EMIT2)) == EMIT_LN)
      e_ln = e;
  const bool c_eln = e_ln >= 0 && p.N <= SM::kVecLn;
  for (int i = threadIdx.x; i < p.N; i += blockDim.x) {
    if (c_bias) vec_bias[i] = __ldg(p.bias + i);
    if (c_ln) {
      vec_lng[i] = __ldg(p.ln_g + i);
      vec_lnb[i] = __ldg(p.ln_b + i);
    }
    if (c_eln) {
      vec_eg[i] = __ldg(p.emit[e_ln].a + i);
      vec_eb[i] = __ldg(p.emit[e_ln].b + i);
    }
  }
  if (warp == kEpiWarps && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < kStages; i++) {
      mbar_init(&full_bar[i], 1);    // CS = 2: only the leader's is used (its producer expects the bytes of both CTAs)
      mbar_init(&empty_bar[i], 1);   // CS = 2: multicast commit
    }
    for (int i = 0; i < 2; i++) {
      mbar_init(&tmem_full[i], 1);               // CS = 2: multicast commit
      mbar_init(&tmem_empty[i], CS * kEpiWarps / SM::kGroups); // released by the warps of the owning group (CS = 2: of both CTAs, on the leader's)
    }
    if (CONV) mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == kEpiWarps + 1) {
    if (CS == 1) tmem_alloc<kTmemCols>(tmem_slot);
    else tmem_alloc2<kTmemCols>(tmem_slot);
  }
  tc_fence_before();
  if (CS > 1) cluster_sync();   // the peer signals this CTA's barriers: all must be initialised cluster-wide
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kEpiWarps) {
    if (lane == 0) {
      // ------------------------------- TMA producer -------------------------------
      int kit = 0;
      if constexpr (CONV) {
        // the whole weight matrix once (N <= BN: one n tile), then per tile and k-block ONE [192 x 64] box of A rows
        // t0 + min_off .. t0 + min_off + 191 that serves every tap
        constexpr int kWTile = BN * kKBlock * 2;   // one [BN x 64] weight tile
        mbar_expect_tx(w_bar, (uint32_t)(num_it * kWTile));
        for (int it = 0; it < num_it; it++) tma_load_2d(smem + SM::kWOff + it * kWTile, &tmB, w_bar, it * kKBlock, 0);
        for (int tile = unit0; tile < total_tiles; tile += unit_step) {
          TileCoord c;
          bool tvalid;
          if (!tile_coord<CS>(p, tile, n_tiles, t_tiles, BN, rank, row_count, c, tvalid)) continue;
          for (int kb = 0; kb < p.kb_per_tap; kb++, kit++) {
            const int st = kit % kStages;
            mbar_wait(&empty_bar[st], ((kit / kStages) & 1) ^ 1);
            mbar_expect_tx(&full_bar[st], SM::kStageBytes);
            tma_load_3d(smem + st * SM::kStageBytes, &tmA, &full_bar[st], kb * kKBlock, c.t0 + p.conv_min_off, c.s);
          }
        }
      } else
      for (int tile = unit0; tile < total_tiles; tile += unit_step) {
        TileCoord c;
        bool tvalid;
        if (!tile_coord<CS>(p, tile, n_tiles, t_tiles, BN, rank, row_count, c, tvalid)) continue;
        for (int it = 0; it < num_it; it++, kit++) {
          const int st = kit % kStages;
          const uint32_t ph = (kit / kStages) & 1;
          mbar_wait(&empty_bar[st], ph ^ 1);
          const int tap = it / p.kb_per_tap;
          const int kb = it - tap * p.kb_per_tap;
          uint8_t* a_dst = smem + st * SM::kStageBytes;
          uint8_t* b_dst = a_dst + kABytes;
          if (CS == 1) {
            mbar_expect_tx(&full_bar[st], SM::kStageBytes);
            tma_load_3d(a_dst, &tmA, &full_bar[st], kb * kKBlock, c.t0 + p.tap_off[tap], c.s + p.tap_seq[tap]);
            tma_load_2d(b_dst, &tmB, &full_bar[st], it * kKBlock, c.n0);
          } else {   // own A tile + own half of the weight k-block, bytes credited to the leader's barrier
            if (leader) mbar_expect_tx(&full_bar[st], 2 * SM::kStageBytes);
            tma2_load_3d(a_dst, &tmA, &full_bar[st], kb * kKBlock, c.t0 + p.tap_off[tap], c.s + p.tap_seq[tap]);
            tma2_load_2d(b_dst, &tmB, &full_bar[st], it * kKBlock, c.n0 + rank * (BN / CS));
          }
        }
      }
    }
  } else if (warp == kEpiWarps + 1) {
    constexpr bool kUni = BN == 256;   // measured on one box: BN = 256 107.4 vs 109.0 ms per step, BN = 128 16.5 vs 15.6 (worse), BN = 64 equal
    if (leader && (kUni || lane == 0)) {
      // ------------------------------- MMA issuer (CS = 2: for both SMs) ----------
      // BN = 256: warp-uniform control flow (descriptors and barrier polls on the uniform datapath), one elected lane issues; the
      // narrower tiles keep the single-lane form
      constexpr uint32_t idesc = umma_idesc_f16(CS * kTileM, BN, 0);
      int kit = 0, lt = 0;
      long long* tb = (p.trace && blockIdx.x == 0 && lane == 0) ? p.trace + 4096 : nullptr;
      int ti = 0;
      for (int tile = unit0; tile < total_tiles; tile += unit_step) {
        TileCoord c;
        bool tvalid;
        if (!tile_coord<CS>(p, tile, n_tiles, t_tiles, BN, rank, row_count, c, tvalid)) continue;
        const int acc = lt & 1;
        gtrace(tb, ti, 1);
        mbar_wait(&tmem_empty[acc], ((lt >> 1) & 1) ^ 1);   // epilogue has drained this accumulator
        tc_fence_after();
        gtrace(tb, ti, 2);
        const uint32_t d_tmem = tmem_base + acc * BN;
        if constexpr (CONV) {
          if (lt == 0) mbar_wait(w_bar, 0);
          const uint32_t w_addr = smem_u32(smem + SM::kWOff);
          for (int kb = 0; kb < p.kb_per_tap; kb++, kit++) {
            const int st = kit % kStages;
            mbar_wait(&full_bar[st], (kit / kStages) & 1);
            tc_fence_after();
            const uint32_t box = smem_u32(smem + st * SM::kStageBytes);
            if (kUni ? elect_one() : true) {
              for (int tap = 0; tap < p.ntaps; tap++) {
                // rows of this tap start (off - min_off) rows into the box: a descriptor whose start is not 1024-byte aligned
                const uint32_t a_addr = box + (uint32_t)(p.tap_off[tap] - p.conv_min_off) * 128u;
                // (the tensor core applies the 128B swizzle to ABSOLUTE shared-memory address bits, exactly like the TMA write did:
                // measured -- with the descriptor's "matrix base offset" field set to (start >> 7) & 7 the result is wrong)
                const uint64_t a_desc = umma_smem_desc_sw128(a_addr);
                const uint64_t b_desc = umma_smem_desc_sw128(w_addr + (uint32_t)(tap * p.kb_per_tap + kb) * (BN * kKBlock * 2));
#pragma unroll
                for (int k = 0; k < kKBlock / 16; k++)
                  umma_f16(d_tmem, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (kb | tap | k) != 0 ? 1u : 0u);
              }
              umma_commit(&empty_bar[st]);
              if (kb + 1 == p.kb_per_tap) umma_commit(&tmem_full[acc]);   // accumulator complete
            }
            if (kUni) __syncwarp();
          }
        } else
        for (int it = 0; it < num_it; it++, kit++) {
          const int st = kit % kStages;
          const uint32_t ph = (kit / kStages) & 1;
          mbar_wait(&full_bar[st], ph);
          tc_fence_after();
          gtrace(tb, ti, 3);
          const uint32_t a_addr = smem_u32(smem + st * SM::kStageBytes);
          const uint64_t a_desc = umma_smem_desc_sw128(a_addr);
          const uint64_t b_desc = umma_smem_desc_sw128(a_addr + kABytes);
          if (kUni ? elect_one() : true) {
#pragma unroll
            for (int k = 0; k < kKBlock / 16; k++) {
              if (CS == 1) umma_f16(d_tmem, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (it | k) != 0 ? 1u : 0u);
              else umma2_f16(d_tmem, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (it | k) != 0 ? 1u : 0u);
            }
            if (CS == 1) umma_commit(&empty_bar[st]);
            else umma2_commit(&empty_bar[st]);
            if (it + 1 == num_it) {      // accumulator complete
              if (CS == 1) umma_commit(&tmem_full[acc]);
              else umma2_commit(&tmem_full[acc]);
            }
          }
          if (kUni) __syncwarp();
        }
        gtrace(tb, ti, 4);
        lt++;
      }
    }
  } else {
    // --------------------------------- epilogue -----------------------------------
    constexpr int kGroups = SM::kGroups;
    constexpr int kGWarps = kEpiWarps / kGroups;   // warps per group: 8
    constexpr int kGParts = kParts / kGroups;      // column parts of a tile inside a group: 2 (BN = 256), 4 (128), 2 (64)
    const int grp = warp / kGWarps;      // this warp's group = the accumulator buffer it drains
    const int ew = warp % kGWarps;
    const int q = warp & 3;              // TMEM lane quarter accessible to this warp (kGWarps is a multiple of 4)
    const int half = ew >> 2;            // which part of the tile's columns this thread owns
    const int r = q * 32 + lane;
    constexpr int kHalfCols = BN / kGParts;   // 128 / 32 / 32
    constexpr int kChunks = kHalfCols / 32;   // 4 / 1 / 1
    float* stg = nullptr;                // (epilogue I/O is direct 256-bit global access: no staging tile)
    float* red_g = red + grp * 2 * 4 * kGParts * 128;
    float *red_a, *red_b, *red_c, *red_d;   // [kGParts][128] each, per group; set per tile (two sets, alternating: with ONE barrier per
                                            // LayerNorm a fast thread may already write the next tile's sums while a slow one reads)
    auto red_sum = [&](const float* a) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < kGParts; k++) t += a[k * 128 + r];
      return t;
    };
    auto group_bar = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(kGWarps * 32) : "memory"); };
    const int ek0 = Cfg::EMIT0 < 0 ? p.emit[0].kind : Cfg::EMIT0;
    const int ek1 = Cfg::EMIT1 < 0 ? p.emit[1].kind : Cfg::EMIT1;
    const int ek2 = Cfg::EMIT2 < 0 ? p.emit[2].kind : Cfg::EMIT2;
    const bool stat2 = ek0 == EMIT_LN || ek1 == EMIT_LN || ek2 == EMIT_LN;
    const bool has_bias = CFGB(BIAS, p.bias != nullptr);
    const bool has_ln = CFGB(LN, p.ln != 0);
    const int act = Cfg::ACT < 0 ? p.act : Cfg::ACT;
    const bool has_rv = CFGB(ROWVEC, p.rowvec != nullptr);
    const bool has_mask = CFGB(MASK, p.mask_pre_res != 0);
    const bool has_res = CFGB(RES, p.res != nullptr);
    const bool has_res2 = CFGB(RES2, p.res2 != nullptr);
    const bool has_out32 = CFGB(OUT32, p.out32 != nullptr);
    const bool has_accum = CFGB(ACCUM, p.out32_accum != 0);
    const bool is_flat = CFGB(FLAT, p.flat != 0);
    const bool has_scale = CFGB(SCALE, p.out_scale != 1.f);
    const bool has_qkv = CFGB(QKV, p.q != nullptr);
    // 32 consecutive per-column parameters: from the staged copy (broadcast LDS.128) when there is one
    auto vload = [&](const float* sv, bool cached, const float* gv, int col, float* d, bool full, int nv) {
      if (cached && full) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const float4 f = *reinterpret_cast<const float4*>(sv + col + i * 4);
          d[i * 4 + 0] = f.x; d[i * 4 + 1] = f.y; d[i * 4 + 2] = f.z; d[i * 4 + 3] = f.w;
        }
      } else {
        load32(gv + col, d, full, nv);
      }
    };

    int lt = 0;
    long long* tb = (p.trace && blockIdx.x == 0 && (warp == 0 || warp == 13) && lane == 0) ? p.trace + (warp == 0 ? 0 : 2048) : nullptr;
    int ti = 0;
    // Tile coordinates come from two dependent global loads (tile list entry, then the sequence length): ~1.5 k cycles of latency
    // that used to sit between two tiles of an epilogue group (clock64 trace of the QKV GEMM).  With a compact tile list every
    // unit is a tile, so (i) the other group's tiles are skipped without loading anything and (ii) the coordinates of this group's
    // NEXT tile are fetched while the current tile's epilogue runs.
    TileCoord c_pref;
    bool tv_pref = false, have_pref = false;
    const int pref_step = unit_step * (p.tile_list ? kGroups : 1);
    for (int tile = unit0; tile < total_tiles; tile += unit_step) {
      if (p.tile_list && kGroups > 1 && (lt & 1) != grp) {   // the other group's tile (compact list: every unit counts)
        lt++;
        continue;
      }
      TileCoord c;
      bool tvalid;
      if (have_pref) {
        c = c_pref;
        tvalid = tv_pref;
        have_pref = false;
      } else if (!tile_coord<CS>(p, tile, n_tiles, t_tiles, BN, rank, row_count, c, tvalid)) {
        continue;
      }
      if (!p.tile_list && kGroups > 1 && (lt & 1) != grp) {   // the other group's tile
        lt++;
        continue;
      }
      if (p.tile_list && tile + pref_step < total_tiles)
        have_pref = tile_coord<CS>(p, tile + pref_step, n_tiles, t_tiles, BN, rank, row_count, c_pref, tv_pref);
      if (CS > 1 && !tvalid) {   // filler tile of an odd tail: drain the accumulator, store nothing
        mbar_wait(&tmem_full[lt & 1], (lt >> 1) & 1);
        tc_fence_after();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&tmem_empty[lt & 1]);
        lt++;
        continue;
      }
      const int acc = lt & 1;
      {
        float* rb = red_g + ((lt / kGroups) & 1) * 4 * kGParts * 128;
        red_a = rb;
        red_b = rb + kGParts * 128;
        red_c = rb + 2 * kGParts * 128;
        red_d = rb + 3 * kGParts * 128;
      }
      const int t = c.t0 + r;
      const bool valid = t < c.len;
      const int ncols = min(BN, p.N - c.n0);                         // valid columns of this tile
      const int my_c0 = half * kHalfCols;                            // first column (within tile) of this thread
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + my_c0;
      const long long row = (long long)c.s * p.T_alloc + t;
      const long long row0 = row - lane;                              // first row of this warp's 32-row block
      gtrace(tb, ti, 10);
      mbar_wait(&tmem_full[acc], (lt >> 1) & 1);
      tc_fence_after();
      gtrace(tb, ti, 11);
      if (p.dbg_skip_epi) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CS > 1) mbar_arrive_leader(&tmem_empty[acc]);
          else mbar_arrive(&tmem_empty[acc]);
        }
        lt++;
        continue;
      }

      if constexpr (Cfg::QKV == 1 && Cfg::BIAS == 0 && Cfg::LN == 0 && Cfg::ACT == 0 && BN == 256) {
        // q / k / v^T split of the estimator blocks (nothing but the split): software-pipelined over the four 32-column chunks.  The
        // tensor-memory read of chunk c+1 is in flight while chunk c is converted, and the accumulator goes back to the MMA warp as
        // soon as the LAST chunk has been read -- before the stores of chunks 2 and 3 -- so the next tile's MMAs overlap them.
        if (!p.q2) {
          const int hd = p.heads * 64;
          const int cb0 = c.n0 + my_c0;                 // first global column of this thread's 128
          const int which = cb0 / hd;                   // 0 q, 1 k, 2 v (hd is a multiple of 128: the thread's columns never straddle)
          const int cc0 = cb0 - which * hd;
          const float sc = valid ? (which == 0 ? p.q_scale : 1.f) : 0.f;
          const bool plain = __all_sync(0xffffffffu, sc == 1.f);   // all 32 rows valid and unscaled (the usual case): no multiplies
          uint32_t rawq[32], pk[16];
          tmem_ld32(taddr, rawq);
          tmem_ld_wait();
#pragma unroll
          for (int ch = 0; ch < kChunks; ch++) {
            if (plain) {
#pragma unroll
              for (int i = 0; i < 16; i++) {
                __half2 h2 = __floats2half2_rn(__uint_as_float(rawq[2 * i]), __uint_as_float(rawq[2 * i + 1]));
                pk[i] = *reinterpret_cast<uint32_t*>(&h2);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; i++) {
                __half2 h2 = __floats2half2_rn(__uint_as_float(rawq[2 * i]) * sc, __uint_as_float(rawq[2 * i + 1]) * sc);
                pk[i] = *reinterpret_cast<uint32_t*>(&h2);
              }
            }
            if (ch + 1 < kChunks) {
              tmem_ld32(taddr + (ch + 1) * 32, rawq);
              tmem_ld_wait();
              if (ch + 2 == kChunks) {   // that was the last read of this accumulator
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                  if (CS == 1) mbar_arrive(&tmem_empty[acc]);
                  else mbar_arrive_leader(&tmem_empty[acc]);
                }
              }
            }
            const int cc = cc0 + ch * 32;
            const int h = cc >> 6, d0 = cc & 63;
            gtrace(tb, ti, 30 + which);
            if (which < 2) {   // q / k: row-major per head; lane-pair transposed 256-bit stores (tile_store_f16 on packed data)
              __half* g0 = (which == 0 ? p.q : p.k) + (((long long)c.s * p.heads + h) * p.T_alloc + (t - lane)) * 64 + d0;
              lane_group_transpose<8, 1>(pk, lane);
              __half* pp = g0 + (lane & ~1) * 64 + (lane & 1) * 16;
              stg256(pp, pk);
              stg256(pp + 64, pk + 8);
            } else {           // v: transposed per head.  Lane pairs exchange so that every lane holds (t, t+1) for 16 of the 32 d rows
                               // of the chunk: 16 four-byte stores per lane instead of 32 two-byte ones (the epilogue is issue bound)
              const bool odd = lane & 1;
              uint32_t out[16];
#pragma unroll
              for (int i = 0; i < 16; i++) {
                // pk[i] = (d = 2i, d = 2i+1) of row t.  Even lane keeps d = 2i of both rows, odd lane d = 2i+1.
                const uint32_t other = __shfl_xor_sync(0xffffffffu, pk[i], 1);
                const uint32_t lo = odd ? other : pk[i], hi = odd ? pk[i] : other;      // lo: row t_even, hi: row t_even + 1
                out[i] = odd ? __byte_perm(lo, hi, 0x7632) : __byte_perm(lo, hi, 0x5410);
              }
              __half* dst = p.vt + (((long long)c.s * p.heads + h) * 64 + d0 + (odd ? 1 : 0)) * p.T_alloc + (t & ~1);
#pragma unroll
              for (int i = 0; i < 16; i++) *reinterpret_cast<uint32_t*>(dst + (long long)(2 * i) * p.T_alloc) = out[i];
            }
          }
          gtrace(tb, ti, 20);
          lt++;
          continue;
        }
      }
      float mean = 0.f, rstd = 1.f;
      uint32_t raw[32];
      float v[32];
      if (has_ln) {  // LayerNorm over the N columns of the row (whole row lives in this tile).  One pass over the accumulator:
                     // sum and sum of squares together, var = E[x^2] - mean^2 in fp32 (<= 256 values of O(1): the cancellation
                     // costs ~1e-5 relative on var, far inside the 1e-2 mel budget) -- one TMEM sweep and one barrier less per tile
        float sum = 0.f, sq = 0.f;
#pragma unroll 1
        for (int ch = 0; ch < kChunks; ch++) {
          const int cl = my_c0 + ch * 32;
          const int nv = min(32, ncols - cl);
          if (nv <= 0) break;
          tmem_ld32(taddr + ch * 32, raw);
          float bch[32];
          vload(vec_bias, c_bias, p.bias, c.n0 + cl, bch, nv == 32, nv);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i++)
            if (i < nv) {
              const float x = __uint_as_float(raw[i]) + bch[i];
              sum += x;
              sq = fmaf(x, x, sq);
            }
        }
        red_a[half * 128 + r] = sum;
        red_b[half * 128 + r] = sq;
        group_bar();
        mean = red_sum(red_a) / (float)ncols;
        rstd = rsqrtf(fmaxf(red_sum(red_b) / (float)ncols - mean * mean, 0.f) + p.ln_eps);
      }

      float sum2 = 0.f, sq2 = 0.f;
      bool released = false;   // (warp-uniform: all lanes of a warp own the same columns)
      const float* rv = has_rv ? p.rowvec + (long long)c.s * p.rowvec_ld : nullptr;
#pragma unroll 1
      for (int ch = 0; ch < kChunks; ch++) {
        const int cl = my_c0 + ch * 32;          // column within the tile
        const int nv = min(32, ncols - cl);
        if (nv <= 0) break;
        const bool full = nv == 32;
        const int cbase = c.n0 + cl;              // global column
        tmem_ld32(taddr + ch * 32, raw);
        float tmp[32];
        if (has_bias) vload(vec_bias, c_bias, p.bias, cbase, tmp, full, nv);
        tmem_ld_wait();
        if (!stat2 && (ch == kChunks - 1 || cl + 32 >= ncols)) {
          released = true;
          // last read of this accumulator (no LayerNorm-emit sweep follows): hand it back to the MMA warp NOW, not after this chunk's
          // math and stores have drained -- the next tile's MMAs then overlap them (QKV trace: ~1.3 k cycles per tile)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CS == 1) mbar_arrive(&tmem_empty[acc]);
            else mbar_arrive_leader(&tmem_empty[acc]);
          }
        }
#pragma unroll
        for (int i = 0; i < 32; i++) v[i] = __uint_as_float(raw[i]);
        gtrace(tb, ti, 12, v[0] + tmp[0]);
        if (Cfg::SCALE < 0 && p.acc_scale != 0.f) {   // generic kernel only
#pragma unroll
          for (int i = 0; i < 32; i++) v[i] *= p.acc_scale;
        }
        if (has_bias) {
#pragma unroll
          for (int i = 0; i < 32; i++) v[i] += tmp[i];
        }
        if (has_ln) {
          vload(vec_lng, c_ln, p.ln_g, cbase, tmp, full, nv);
#pragma unroll
          for (int i = 0; i < 32; i++) v[i] = (v[i] - mean) * rstd * tmp[i];
          vload(vec_lnb, c_ln, p.ln_b, cbase, tmp, full, nv);
#pragma unroll
          for (int i = 0; i < 32; i++) v[i] += tmp[i];
        }
        switch (act) {
          case ACT_MISH:
#pragma unroll
            for (int i = 0; i < 32; i++) v[i] = fast_mish(v[i]);
            break;
          case ACT_GELU:
#pragma unroll
            for (int i = 0; i < 32; i++) v[i] = fast_gelu_erf(v[i]);
            break;
          case ACT_SILU:
#pragma unroll
            for (int i = 0; i < 32; i++) v[i] = fast_silu(v[i]);
            break;
          case ACT_ELU:
#pragma unroll
            for (int i = 0; i < 32; i++) v[i] = elu_f(v[i]);
            break;
          case ACT_LRELU:
#pragma unroll
            for (int i = 0; i < 32; i++) v[i] = v[i] > 0.f ? v[i] : v[i] * p.act_f;
            break;
          case ACT_SNAKE:
            load32(p.act_a + cbase, tmp, full, nv);
#pragma unroll
            for (int i = 0; i < 32; i++) v[i] = fast_snake(v[i], tmp[i]);
            break;
          default: break;
        }
        if (has_rv) {
          load32(rv + cbase, tmp, full, nv);
#pragma unroll
          for (int i = 0; i < 32; i++) v[i] += tmp[i];
        }
        if (has_mask && !valid) {
#pragma unroll
          for (int i = 0; i < 32; i++) v[i] = 0.f;
        }
        if (has_res) {
          if (full) tile_load_f32(p.res + row0 * p.res_ld + cbase, p.res_ld, stg, lane, tmp);
          else load32(p.res + row * p.res_ld + cbase, tmp, full, nv);
#pragma unroll
          for (int i = 0; i < 32; i++) v[i] += tmp[i];
          gtrace(tb, ti, 13, v[0] + v[31]);
        }
        if (has_res2) {
          if (full) tile_load_f32(p.res2 + row0 * p.res2_ld + cbase, p.res2_ld, stg, lane, tmp);
          else load32(p.res2 + row * p.res2_ld + cbase, tmp, full, nv);
#pragma unroll
          for (int i = 0; i < 32; i++) v[i] += tmp[i];
        }
        if (has_scale) {
#pragma unroll
          for (int i = 0; i < 32; i++) v[i] *= p.out_scale;
        }
        if (!full) {
#pragma unroll
          for (int i = 0; i < 32; i++)
            if (i >= nv) v[i] = 0.f;
        }
        if (has_out32) {
          if (is_flat) {
            // transposed conv: row t holds stride*Cout consecutive output elements of the sequence slab
            const long long e0 = (long long)t * p.out32_ld + cbase + p.flat_off;
            const long long hi = (long long)c.len * p.flat_hi_per_len + p.flat_hi_add;
            if (e0 >= p.flat_lo && e0 + nv <= hi) store32_f32(p.out32 + (long long)c.s * p.flat_seq_elems + e0, v, full, nv);
          } else {
            float* op = p.out32 + row * p.out32_ld + cbase;
            float* op0 = p.out32 + row0 * p.out32_ld + cbase;
            if (has_accum) {
              if (full) tile_load_f32(op0, p.out32_ld, stg, lane, tmp);
              else load32(op, tmp, full, nv);
#pragma unroll
              for (int i = 0; i < 32; i++) v[i] += tmp[i];
            }
            if (full) tile_store_f32(op0, p.out32_ld, stg, lane, v);
            else store32_f32(op, v, full, nv);
          }
        }
        gtrace(tb, ti, 14);
        // 16-bit emits (the next contraction's A operand); padded rows are written as zeros
#pragma unroll
        for (int e = 0; e < 3; e++) {
          const Emit& em = p.emit[e];
          const int ek = e == 0 ? ek0 : (e == 1 ? ek1 : ek2);
          if (ek == EMIT_NONE || ek == EMIT_LN) continue;
          float w[32];
          if (ek == EMIT_SNAKE) {
            load32(em.a + cbase, tmp, full, nv);
#pragma unroll
            for (int i = 0; i < 32; i++) w[i] = fast_snake(v[i], tmp[i]);
          } else if (ek == EMIT_LRELU) {
#pragma unroll
            for (int i = 0; i < 32; i++) w[i] = v[i] > 0.f ? v[i] : v[i] * em.f;
          } else if (ek == EMIT_LO) {
#pragma unroll
            for (int i = 0; i < 32; i++) w[i] = v[i] - __half2float(__float2half_rn(v[i]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; i++) w[i] = v[i];
          }
          if (em.scale != 1.f) {
#pragma unroll
            for (int i = 0; i < 32; i++) w[i] *= em.scale;
          }
          if (!valid) {
#pragma unroll
            for (int i = 0; i < 32; i++) w[i] = 0.f;
          }
          if (full) tile_store_f16(em.ptr + row0 * em.ld + em.col_off + cbase, em.ld, stg, lane, w);
          else store32_f16(em.ptr + row * em.ld + em.col_off + cbase, w, full, nv);
        }
        if (has_qkv) {  // attention operand split: q (pre-scaled) and k row-major per head, v transposed per head
          const int hd = p.heads * 64;
          int which = cbase / hd;                  // 0 q, 1 k, 2 v (a 32-col chunk never straddles); with q2: 0 q, 1 q2, 2 k, 3 v
          const int cc = cbase - which * hd;
          const int h = cc >> 6, d0 = cc & 63;
          __half* qdst = p.q;
          if (p.q2) {
            if (which == 1) qdst = p.q2;
            which = which < 2 ? 0 : which - 1;
          }
          gtrace(tb, ti, 30 + which);
          if (which < 2) {
            const float sc = valid ? (which == 0 ? p.q_scale : 1.f) : 0.f;
            float w[32];
#pragma unroll
            for (int i = 0; i < 32; i++) w[i] = v[i] * sc;
            tile_store_f16((which == 0 ? qdst : p.k) + (((long long)c.s * p.heads + h) * p.T_alloc + (t - lane)) * 64 + d0, 64, stg, lane, w);
          } else {
            __half* dst = p.vt + (((long long)c.s * p.heads + h) * 64 + d0) * p.T_alloc + t;
#pragma unroll
            for (int i = 0; i < 32; i++) dst[(long long)i * p.T_alloc] = __float2half_rn(valid ? v[i] : 0.f);
          }
        }
        if (stat2) {
#pragma unroll
          for (int i = 0; i < 32; i++) {
            sum2 += v[i];
            sq2 = fmaf(v[i], v[i], sq2);
            raw[i] = __float_as_uint(v[i]);
          }
          tmem_st32(taddr + ch * 32, raw);
        }
      }
      gtrace(tb, ti, 15);
      if (stat2) {  // LayerNorm of the final row value (pre-norm of the next sub-block), emitted as 16-bit
        tmem_st_wait();
        red_c[half * 128 + r] = sum2;
        red_d[half * 128 + r] = sq2;
        gtrace(tb, ti, 16);
        group_bar();
        gtrace(tb, ti, 17);
        const float mean2 = red_sum(red_c) / (float)ncols;
        const float var2 = fmaxf(red_sum(red_d) / (float)ncols - mean2 * mean2, 0.f);   // one pass, see has_ln above
#pragma unroll 1
        for (int ch = 0; ch < kChunks; ch++) {
          const int cl = my_c0 + ch * 32;
          const int nv = min(32, ncols - cl);
          if (nv <= 0) break;
          const int cbase = c.n0 + cl;
          tmem_ld32(taddr + ch * 32, raw);
          tmem_ld_wait();
          if (ch == kChunks - 1 || cl + 32 >= ncols) {   // last read of the accumulator: release it before this chunk's math and stores
            released = true;
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (CS == 1) mbar_arrive(&tmem_empty[acc]);
              else mbar_arrive_leader(&tmem_empty[acc]);
            }
          }
#pragma unroll
          for (int e = 0; e < 3; e++) {
            const Emit& em = p.emit[e];
            const int ek = e == 0 ? ek0 : (e == 1 ? ek1 : ek2);
            if (ek != EMIT_LN) continue;
            const float rstd2 = rsqrtf(var2 + em.f);
            float g[32], w[32];
            vload(vec_eg, c_eln && e == e_ln, em.a, cbase, g, nv == 32, nv);
#pragma unroll
            for (int i = 0; i < 32; i++) w[i] = (__uint_as_float(raw[i]) - mean2) * rstd2 * g[i];
            vload(vec_eb, c_eln && e == e_ln, em.b, cbase, g, nv == 32, nv);
            const float sc = valid ? em.scale : 0.f;
#pragma unroll
            for (int i = 0; i < 32; i++) w[i] = (w[i] + g[i]) * sc;
            if (nv == 32) tile_store_f16(em.ptr + row0 * em.ld + em.col_off + cbase, em.ld, stg, lane, w);
            else store32_f16(em.ptr + row * em.ld + em.col_off + cbase, w, false, nv);
          }
        }
      }
      gtrace(tb, ti, 20);
      if (!released) {   // (normally released right after the last accumulator read, see the chunk loops)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CS == 1) mbar_arrive(&tmem_empty[acc]);
          else mbar_arrive_leader(&tmem_empty[acc]);
        }
      }
      lt++;
    }
  }

  tc_fence_before();
  if (CS > 1) cluster_sync();   // no CTA may exit while a peer can still multicast into it or arrive on its barriers
  else __syncthreads();
  if (warp == kEpiWarps + 1) {
    if (CS == 1) tmem_dealloc<kTmemCols>(tmem_base);
    else tmem_dealloc2<kTmemCols>(tmem_base);
  }
}

// Compact list of the (sequence, t0) row tiles that contain at least one row < len + halo, in (s, t) order.
__global__ void build_tile_list_kernel(const int* __restrict__ lens, int S, int T_alloc, int halo, int* __restrict__ list,
                                       int* __restrict__ count, const int* __restrict__ lo) {
  extern __shared__ int offs[];   // [S + 1]
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const int rows = lens[s] > 0 ? min(lens[s] + halo, T_alloc) : 0;   // an empty sequence (idle streaming slot) has no tile
    const int first = lo ? lo[s] / kTileM : 0;
    const int n = rows > 0 ? (rows + kTileM - 1) / kTileM : 0;
    offs[s + 1] = n > first ? n - first : 0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    offs[0] = 0;
    for (int s = 0; s < S; s++) offs[s + 1] += offs[s];
    *count = offs[S];
  }
  __syncthreads();
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const int n = offs[s + 1] - offs[s];
    const int first = lo ? lo[s] / kTileM : 0;
    for (int k = 0; k < n; k++) {
      list[2 * (offs[s] + k)] = s;
      list[2 * (offs[s] + k) + 1] = (first + k) * kTileM;
    }
  }
}
void launch_build_tile_list(const int* lens, int S, int T_alloc, int halo, int* list, int* count, cudaStream_t stream, const int* lo) {
  build_tile_list_kernel<<<1, 256, (S + 1) * sizeof(int), stream>>>(lens, S, T_alloc, halo, list, count, lo);
  CV2_LAUNCH_CHECK();
}

static int g_num_sms = 0;

template <int BN, class Cfg, int CS>
static void launch_cfg_cs(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream) {
  static PerDeviceOnce once;
  static int max_clusters = 0;   // every device of a box is the same part
  if (g_num_sms == 0) {
    int dev = 0;
    CV2_CUDA(cudaGetDevice(&dev));
    CV2_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  once.run([&] {
    CV2_CUDA(cudaFuncSetAttribute(gemm_tap_kernel<BN, Cfg, CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmSmem<BN, CS>::kTotal));
    if (CS > 1) {
      cudaLaunchConfig_t q = {};
      q.gridDim = dim3(g_num_sms / CS * CS);
      q.blockDim = dim3(GemmSmem<BN, CS>::kThreads);
      q.dynamicSmemBytes = GemmSmem<BN, CS>::kTotal;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      q.attrs = at;
      q.numAttrs = 1;
      CV2_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, gemm_tap_kernel<BN, Cfg, CS>, &q));
      CV2_CHECK(max_clusters > 0, "gemm_tap: no cluster of %d CTAs fits", CS);
    }
  });
  const int rows = (p.T_alloc / kTileM) * p.S;                         // upper bound on row tiles (the list may hold fewer)
  const int total = ((p.N + BN - 1) / BN) * ((rows + CS - 1) / CS);    // work units
  if (CS == 1) {
    const int grid = total < g_num_sms ? total : g_num_sms;   // persistent: one CTA per SM
    gemm_tap_kernel<BN, Cfg, 1><<<grid, GemmSmem<BN>::kThreads, GemmSmem<BN>::kTotal, stream>>>(tmA, tmB, p);
  } else {
    const int clusters = total < max_clusters ? total : max_clusters;
    cudaLaunchConfig_t q = {};
    q.gridDim = dim3(clusters * CS);
    q.blockDim = dim3(GemmSmem<BN, CS>::kThreads);
    q.dynamicSmemBytes = GemmSmem<BN, CS>::kTotal;
    q.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    q.attrs = at;
    q.numAttrs = 1;
    CV2_CUDA(cudaLaunchKernelEx(&q, gemm_tap_kernel<BN, Cfg, CS>, tmA, tmB, p));
  }
  CV2_LAUNCH_CHECK();
}

// conv mode (GemmParams::tmA_halo): persistent 1-SM CTAs, [192 x 64] A boxes, resident weights
template <int BN, class Cfg>
static void launch_conv(const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream) {
  using SM = GemmSmemConv<BN>;
  static PerDeviceOnce once;
  if (g_num_sms == 0) {
    int dev = 0;
    CV2_CUDA(cudaGetDevice(&dev));
    CV2_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  once.run([] { CV2_CUDA(cudaFuncSetAttribute(gemm_tap_kernel<BN, Cfg, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kTotal)); });
  const int rows = (p.T_alloc / kTileM) * p.S;
  const int grid = rows < g_num_sms ? rows : g_num_sms;
  gemm_tap_kernel<BN, Cfg, 1, true><<<grid, SM::kThreads, SM::kTotal, stream>>>(*p.tmA_halo, tmB, p);
  CV2_LAUNCH_CHECK();
}

bool gemm_tap_conv_eligible(int bn, const GemmParams& p) {
  if (bn != 64 && bn != 128) return false;
  if (p.N > bn || p.ntaps < 2 || p.q || p.S_map > 0) return false;
  if ((long long)p.ntaps * p.kb_per_tap * bn * kKBlock * 2 > GemmSmemConv<64>::kWBytes) return false;
  int lo = p.tap_off[0], hi = p.tap_off[0];
  for (int j = 0; j < p.ntaps; j++) {
    if (p.tap_seq[j] != 0) return false;
    lo = p.tap_off[j] < lo ? p.tap_off[j] : lo;
    hi = p.tap_off[j] > hi ? p.tap_off[j] : hi;
  }
  return hi - lo <= kConvHaloRows - kTileM;
}

// tmB2: weight map with a {64, BN/2} box for the 2-CTA multicast path (null: single-CTA path)
template <int BN, class Cfg>
static void launch_cfg(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream) {
  if constexpr (BN != 256) {
    if (p.tmA_halo) {
      launch_conv<BN, Cfg>(tmB, p, stream);
      return;
    }
  }
  if constexpr (BN == 256) {
    if (p.tmB_half && p.tile_list) {
      launch_cfg_cs<BN, Cfg, 2>(tmA, *p.tmB_half, p, stream);
      return;
    }
  }
  launch_cfg_cs<BN, Cfg, 1>(tmA, tmB, p, stream);
}

// Specialised epilogues.               ACT       B LN RV MK RS R2 O32 AC FL SC  EMIT0       EMIT1      EMIT2     QKV
using EpiQkv      = EpiCfg<ACT_NONE,  0, 0, 0, 0, 0, 0, 0, 0, 0, 0, EMIT_NONE,  EMIT_NONE, EMIT_NONE, 1>;
using EpiQkvB     = EpiCfg<ACT_NONE,  1, 0, 0, 0, 0, 0, 0, 0, 0, 0, EMIT_NONE,  EMIT_NONE, EMIT_NONE, 1>;   // encoder q|q2|k|v
using EpiResLn    = EpiCfg<ACT_NONE,  1, 0, 0, 0, 1, 0, 1, 0, 0, 0, EMIT_LN,    EMIT_NONE, EMIT_NONE, 0>;   // out-proj / FF2
using EpiResPlain = EpiCfg<ACT_NONE,  1, 0, 0, 0, 1, 0, 1, 0, 0, 0, EMIT_PLAIN, EMIT_NONE, EMIT_NONE, 0>;   // last FF2 of a group
using EpiGelu     = EpiCfg<ACT_GELU,  1, 0, 0, 0, 0, 0, 0, 0, 0, 0, EMIT_PLAIN, EMIT_NONE, EMIT_NONE, 0>;   // FF1
using EpiConv1    = EpiCfg<ACT_MISH,  1, 1, 1, 0, 0, 0, 0, 0, 0, 0, EMIT_PLAIN, EMIT_NONE, EMIT_NONE, 0>;   // resnet block1
using EpiConv2    = EpiCfg<ACT_MISH,  1, 1, 0, 1, 1, 0, 1, 0, 0, 0, EMIT_LN,    EMIT_NONE, EMIT_NONE, 0>;   // resnet block2
using EpiOut32    = EpiCfg<ACT_NONE,  1, 0, 0, 0, 0, 0, 1, 0, 0, 0, EMIT_NONE,  EMIT_NONE, EMIT_NONE, 0>;   // res_conv, projections
using EpiPlain    = EpiCfg<ACT_NONE,  1, 0, 0, 0, 0, 0, 0, 0, 0, 0, EMIT_PLAIN, EMIT_NONE, EMIT_NONE, 0>;   // plain conv -> 16-bit
using EpiSnake    = EpiCfg<ACT_SNAKE, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, EMIT_PLAIN, EMIT_NONE, EMIT_NONE, 0>;   // ResBlock conv1
using EpiResSnake = EpiCfg<ACT_NONE,  1, 0, 0, 0, 1, 0, 1, 0, 0, 0, EMIT_SNAKE, EMIT_NONE, EMIT_NONE, 0>;   // ResBlock conv2
using EpiResSum   = EpiCfg<ACT_NONE,  1, 0, 0, 0, 1, 0, 1, -1, 0, 1, -1,        EMIT_NONE, EMIT_NONE, 0>;   // ResBlock tail (x/3 sum)
using EpiSilu     = EpiCfg<ACT_SILU,  1, 0, 0, 0, 0, 0, 0, 0, 0, 0, EMIT_PLAIN, EMIT_NONE, EMIT_NONE, 0>;   // encoder FFN w_1
using EpiRes      = EpiCfg<ACT_NONE,  1, 0, 0, 0, 1, 0, 1, 0, 0, 0, EMIT_NONE,  EMIT_NONE, EMIT_NONE, 0>;   // encoder residual adds

template <class Cfg>
static bool cfg_matches(const GemmParams& p) {
  auto ok = [](int want, int have) { return want < 0 || want == have; };
  return ok(Cfg::ACT, p.act) && ok(Cfg::BIAS, p.bias != nullptr) && ok(Cfg::LN, p.ln != 0) && ok(Cfg::ROWVEC, p.rowvec != nullptr) &&
         ok(Cfg::MASK, p.mask_pre_res != 0) && ok(Cfg::RES, p.res != nullptr) && ok(Cfg::RES2, p.res2 != nullptr) &&
         ok(Cfg::OUT32, p.out32 != nullptr) && ok(Cfg::ACCUM, p.out32_accum != 0) && ok(Cfg::FLAT, p.flat != 0) &&
         ok(Cfg::SCALE, p.out_scale != 1.f) && ok(Cfg::EMIT0, p.emit[0].kind) && ok(Cfg::EMIT1, p.emit[1].kind) &&
         ok(Cfg::EMIT2, p.emit[2].kind) && ok(Cfg::QKV, p.q != nullptr);
}

template <int BN>
static void launch_bn(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream) {
  if (p.acc_scale != 0.f && p.acc_scale != 1.f) {   // only the generic epilogue implements the accumulator pre-scale
    launch_cfg<BN, EpiGeneric>(tmA, tmB, p, stream);
    return;
  }
#define TRY(CFG)                                  \
  if (cfg_matches<CFG>(p)) {                      \
    launch_cfg<BN, CFG>(tmA, tmB, p, stream);     \
    return;                                       \
  }
  if constexpr (BN == 256) {
    TRY(EpiQkv) TRY(EpiResLn) TRY(EpiGelu) TRY(EpiConv1) TRY(EpiConv2) TRY(EpiResPlain) TRY(EpiSilu) TRY(EpiRes) TRY(EpiQkvB)
  }
  TRY(EpiOut32) TRY(EpiPlain) TRY(EpiSnake) TRY(EpiResSnake)
  if (p.emit[0].kind == EMIT_NONE || p.emit[0].kind == EMIT_LRELU) { TRY(EpiResSum) }
#undef TRY
  launch_cfg<BN, EpiGeneric>(tmA, tmB, p, stream);
}

// index of the epilogue specialisation launch_gemm_tap would pick (profiling labels only)
int gemm_tap_spec(int bn, const GemmParams& p) {
  if (p.acc_scale != 0.f && p.acc_scale != 1.f) return 0;
  int i = 1;
#define TRYS(CFG)                 \
  if (cfg_matches<CFG>(p)) return i; \
  i++;
  if (bn == 256) {
    TRYS(EpiQkv) TRYS(EpiResLn) TRYS(EpiGelu) TRYS(EpiConv1) TRYS(EpiConv2) TRYS(EpiResPlain) TRYS(EpiSilu) TRYS(EpiRes)
  } else {
    i += 8;
  }
  TRYS(EpiOut32) TRYS(EpiPlain) TRYS(EpiSnake) TRYS(EpiResSnake)
  if (p.emit[0].kind == EMIT_NONE || p.emit[0].kind == EMIT_LRELU) { TRYS(EpiResSum) }
#undef TRYS
  return 0;
}

void launch_gemm_tap(int bn, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream) {
  CV2_CHECK(p.T_alloc % kTileM == 0, "gemm_tap: T_alloc %d not a multiple of 128", p.T_alloc);
  CV2_CHECK(p.ntaps >= 1 && p.ntaps <= 16 && p.kb_per_tap >= 1, "gemm_tap: bad taps %d / kb %d", p.ntaps, p.kb_per_tap);
  bool wants_row = p.ln;
  for (int e = 0; e < 3; e++) wants_row |= (p.emit[e].kind == EMIT_LN);
  CV2_CHECK(!wants_row || p.N <= bn, "gemm_tap: LayerNorm epilogue needs the full row in one tile (N=%d, BN=%d)", p.N, bn);
  CV2_CHECK(!p.ln || p.bias, "gemm_tap: LayerNorm epilogue expects a bias");
  CV2_CHECK(!p.ln || p.acc_scale == 0.f || p.acc_scale == 1.f, "gemm_tap: acc_scale is not implemented together with the LayerNorm epilogue");
  CV2_CHECK(!p.q || (p.N == (p.q2 ? 4 : 3) * p.heads * 64), "gemm_tap: qkv split needs N == 3*heads*64 (4*heads*64 with q2)");
  switch (bn) {
    case 64: launch_bn<64>(tmA, tmB, p, stream); break;
    case 128: launch_bn<128>(tmA, tmB, p, stream); break;
    case 256: launch_bn<256>(tmA, tmB, p, stream); break;
    default: fail("gemm_tap: unsupported BN %d", bn);
  }
}

}  // namespace cv2
