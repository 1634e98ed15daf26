// Two-SM form of the fused estimator feed-forward (see ffn_fused.cu for the algorithm): a cluster of two CTAs works on two
// 128-row tiles and ONE thread (CTA rank 0) issues every MMA for both SMs with tcgen05.mma.cta_group::2 (M = 256).
// Why: with one CTA per MMA the tensor pipe ran at ~55 % of its rate in this kernel (112 cycles per 128x128x16 MMA instead of
// 64; profiles/ffn_trace.py), and the micro-benchmarks in profiles/micro show the cause is the issue path -- descriptors that
// change every few MMAs cost ~120 cycles per instruction in cta_group::1 form, while the cta_group::2 form sustains the nominal
// 64 cycles with twice the work per instruction -- and each SM only stages HALF of every weight tile (the pair shares B).
//   * each CTA: its own H tile, its own 16 epilogue warps, its own TMEM (acc1 2 x 128, acc2 256), its own output epilogue;
//   * weights: every CTA TMA-loads its half of each tile (W1 chunk: 64 of 128 rows; W2 k-block: 128 of 256 rows) into its own
//     ring and signals the LEADER's full barrier (cp.async.bulk.tensor ... .cta_group::2, peer bit cleared);
//   * the leader's commits are multicast to both CTAs' barriers (slot free, acc1_full, acc2_full); the peer's epilogue warps
//     arrive remotely on the leader's f_full / acc2_empty (count 32); the leader relays f_seen to the peer.
#include "common.cuh"
#include "epi_util.cuh"
#include "ffn_fused.cuh"
#include "host_util.h"

namespace cv2 {

static constexpr int kHBytes = 4 * 16384;       // [128 x 256] 16-bit, four 64-column swizzle atoms
static constexpr int kSlots = 8;                 // the hidden chunk lives in TMEM and the epilogue needs no staging: all the
static constexpr int kSlotBytes = 16384;        // remaining shared memory is weight ring (one [128 x 64] tile per slot)
static constexpr int kOffW = kHBytes;
static constexpr int kEpiW = 16;                 // epilogue warps: four per TMEM lane quarter
static constexpr int kOffRed = kOffW + kSlots * kSlotBytes;
static constexpr int kOffRed2 = kOffRed + 2 * 2 * 4 * 128 * 4;  // (sum, sum of squares) x [4][128], double-buffered by tile parity
static constexpr int kOffBar = kOffRed2 + 2 * 2 * 4 * 128 * 4;  // second set: LayerNorm3 of the chained out-projection
static constexpr int kOffVec = kOffBar + 256;       // b1 [1024] | b2 [256] | next pre-norm gamma [256] | beta [256] | bo | ln3 gamma | beta
static constexpr int kFfnSmem = kOffVec + (1024 + 6 * 256) * 4;
static constexpr int kFfnThreads = 64 + kEpiW * 32;
static constexpr uint32_t kAcc1 = 0, kAcc2 = 256;   // TMEM columns: acc1 = 2 x 128, acc2 = 256

__device__ __forceinline__ void ffn_trace(long long* buf, int& idx, int code, float dep = 0.f) {
  if (buf && idx < 4095) {
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) : "f"(dep) : "memory");   // `dep` orders the read after the value exists
    buf[idx++] = (t << 8) | code;
  }
}
__device__ __forceinline__ void ffn_bar() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

__device__ __forceinline__ bool ffn_tile(const FfnParams& p, int tile, int t_tiles, int& s, int& t0, int& len) {
  if (p.tile_list) {
    s = __ldg(p.tile_list + 2 * tile);
    t0 = __ldg(p.tile_list + 2 * tile + 1);
    len = __ldg(p.lens + s);
    return true;
  }
  s = tile / t_tiles;
  t0 = (tile % t_tiles) * 128;
  len = p.lens ? __ldg(p.lens + s) : p.len_all;
  return t0 < len + p.halo;
}

// pair unit -> this CTA's row tile; false for the filler tile of an odd tail (computed on a copy of tile 0, nothing stored)
__device__ __forceinline__ bool ffn2_tile(const FfnParams& p, int unit, uint32_t rank, int row_tiles, int& s, int& t0, int& len) {
  int idx = 2 * unit + (int)rank;
  const bool ok = idx < row_tiles;
  if (!ok) idx = 0;
  s = __ldg(p.tile_list + 2 * idx);
  t0 = __ldg(p.tile_list + 2 * idx + 1);
  len = __ldg(p.lens + s);
  return ok;
}

// op o of a tile's schedule: FF1(0), FF1(1), FF2(0), FF1(2), FF2(1), ..., FF1(7), FF2(6), FF2(7)
__device__ __forceinline__ void ffn_op(int o, bool& is_ff2, int& c) {
  if (o < 2) { is_ff2 = false; c = o; }
  else if (o == 15) { is_ff2 = true; c = 7; }
  else if (o & 1) { is_ff2 = false; c = (o + 1) >> 1; }
  else { is_ff2 = true; c = (o >> 1) - 1; }
}

// 32 consecutive staged parameters (every lane reads the same address: broadcast LDS.128)
__device__ __forceinline__ void lds32(const float* sv, float* d) {
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const float4 f = *reinterpret_cast<const float4*>(sv + i * 4);
    d[i * 4 + 0] = f.x; d[i * 4 + 1] = f.y; d[i * 4 + 2] = f.z; d[i * 4 + 3] = f.w;
  }
}

// release at cluster scope: the peer's generic-proxy writes to ITS shared memory (the H tile it built) must be visible to the
// MMA the leader issues for both SMs after observing the barrier
__device__ __forceinline__ void mbar_arrive_leader_release(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}

// CHAIN: tmH is the attention output [S, T_alloc, 512] (A operand of the out-projection), tmWo the out-proj weight halves.
template <bool CHAIN, int HS = 1>
__global__ void __launch_bounds__(kFfnThreads, 1)
ffn_fused2_kernel(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmWo,
                  const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2, const FfnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];   // 1024 B alignment for the 128B-swizzle atoms
  float* red = reinterpret_cast<float*>(smem + kOffRed);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* h_full = bars + 0;
  uint64_t* h_empty = bars + 1;
  uint64_t* w_full = bars + 2;             // [kSlots]
  uint64_t* w_empty = w_full + kSlots;     // [kSlots]
  uint64_t* acc1_full = w_empty + kSlots;  // [2]
  uint64_t* acc1_empty = acc1_full + 2;    // [2]
  uint64_t* f_full = acc1_empty + 2;
  uint64_t* f_seen = f_full + 1;           // MMA thread has observed f_full of a chunk (keeps f_full at most one phase ahead)
  uint64_t* acc2_full = f_seen + 1;
  uint64_t* acc2_empty = acc2_full + 1;
  uint64_t* op_full = acc2_empty + 1;      // CHAIN: out-projection accumulator complete (multicast commit)
  uint64_t* h_ready = op_full + 1;         // CHAIN: leader only -- both CTAs' epilogue warps have built their H tile in smem
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(h_ready + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  static_assert(HS == 1 || (CHAIN && 8 % HS == 0), "the hidden split exists for the chained kernel only");
  constexpr int kNC = 8 / HS;                                 // hidden chunks per unit
  const int row_tiles = __ldg(p.tile_count);                 // (the 2-SM path requires the compact tile list)
  const int total_tiles = ((row_tiles + 1) / 2) * HS;         // units: tile pairs x hidden splits
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int unit0 = blockIdx.x >> 1, unit_step = gridDim.x >> 1;

  float* vec_b1 = reinterpret_cast<float*>(smem + kOffVec);
  float* vec_b2 = vec_b1 + 1024;
  float* vec_g = vec_b2 + 256;
  float* vec_b = vec_g + 256;
  float* vec_bo = vec_b + 256;
  float* vec_g3 = vec_bo + 256;
  float* vec_b3 = vec_g3 + 256;
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
    vec_b1[i] = __ldg(p.b1 + i);
    if (i < 256) {
      vec_b2[i] = __ldg(p.b2 + i);
      if (p.emit_ln.ptr) {
        vec_g[i] = __ldg(p.emit_ln.a + i);
        vec_b[i] = __ldg(p.emit_ln.b + i);
      }
      if (CHAIN) {
        vec_bo[i] = __ldg(p.bo + i);
        vec_g3[i] = __ldg(p.ln3_g + i);
        vec_b3[i] = __ldg(p.ln3_b + i);
      }
    }
  }
  if (warp == kEpiW && lane == 0) {
    tma_prefetch_desc(&tmH);
    if (CHAIN) tma_prefetch_desc(&tmWo);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    mbar_init(h_full, 1);          // leader only: its producer's arrive.expect_tx (bytes of BOTH CTAs)
    mbar_init(h_empty, 1);         // multicast commit
    for (int i = 0; i < kSlots; i++) {
      mbar_init(&w_full[i], 1);    // leader only
      mbar_init(&w_empty[i], 1);   // multicast commit
    }
    for (int i = 0; i < 2; i++) {
      mbar_init(&acc1_full[i], 1); // multicast commit
      mbar_init(&acc1_empty[i], 1);
    }
    mbar_init(f_full, 2 * kEpiW);  // leader only: the epilogue warps of both CTAs
    mbar_init(f_seen, 1);          // leader's MMA thread arrives in both CTAs
    mbar_init(acc2_full, 1);       // multicast commit
    mbar_init(acc2_empty, 2 * kEpiW);   // leader only
    mbar_init(op_full, 1);              // multicast commit
    mbar_init(h_ready, 2 * kEpiW);      // leader only
    fence_barrier_init();
  }
  if (warp == kEpiW + 1) {         // both CTAs, same warp id, same destination
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync();                  // the peer signals this CTA's barriers: everything must be initialised cluster-wide
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == kEpiW) {
    if (lane == 0) {
      // ------------------------------- TMA producer (both CTAs) --------------------
      // Each CTA loads ITS H tile and ITS half of every weight tile into its own shared memory; the bytes are credited to the
      // leader's full barriers, on which only the leader's producer arrives (expecting the bytes of both CTAs).
      int lt = 0, wit = 0;
      for (int unit = unit0; unit < total_tiles; unit += unit_step) {
        int s, t0, len;
        ffn2_tile(p, unit / HS, rank, row_tiles, s, t0, len);
        const int c0 = (unit % HS) * kNC;
        prefetch_l2_bulk(p.x32 + ((long long)s * p.T_alloc + t0) * 256, 128 * 256 * 4);
        if (!CHAIN) {
          mbar_wait(h_empty, (lt & 1) ^ 1);
          if (leader) mbar_expect_tx(h_full, 2 * kHBytes);
#pragma unroll
          for (int kb = 0; kb < 4; kb++) tma2_load_3d(smem + kb * 16384, &tmH, h_full, kb * 64, t0, s);
        } else {
          // out-projection operands through the weight ring, two slots per 64-wide k-block: this CTA's [128 rows x 64] slice
          // of the attention output, and its 128 of the 256 output rows of Wo
          for (int kb = 0; kb < 8; kb++) {
#pragma unroll
            for (int i = 0; i < 2; i++, wit++) {
              const int st = wit % kSlots;
              mbar_wait(&w_empty[st], ((wit / kSlots) & 1) ^ 1);
              if (leader) mbar_expect_tx(&w_full[st], 2 * kSlotBytes);
              uint8_t* dst = smem + kOffW + st * kSlotBytes;
              if (i == 0) tma2_load_3d(dst, &tmH, &w_full[st], kb * 64, t0, s);
              else tma2_load_2d(dst, &tmWo, &w_full[st], kb * 64, (int)rank * 128);
            }
          }
        }
        for (int o = 0; o < 2 * kNC; o++) {
          bool is_ff2;
          int c;
          if (HS == 1) ffn_op(o, is_ff2, c);
          else { is_ff2 = o >= kNC; c = c0 + (o % kNC); }     // split: FF1(c0) .. FF1(c0+n-1), FF2(c0) .. FF2(c0+n-1)
          for (int i = 0; i < 2; i++, wit++) {     // two 16 KB slots per op and CTA
            const int st = wit % kSlots;
            mbar_wait(&w_empty[st], ((wit / kSlots) & 1) ^ 1);
            if (leader) mbar_expect_tx(&w_full[st], 2 * kSlotBytes);
            uint8_t* dst = smem + kOffW + st * kSlotBytes;
            if (!is_ff2) {   // W1 chunk c, k-blocks 2i and 2i+1: this CTA's 64 of the 128 hidden rows, [64 x 64] each (8 KB)
              tma2_load_2d(dst, &tmW1, &w_full[st], (2 * i) * 64, c * 128 + (int)rank * 64);
              tma2_load_2d(dst + 8192, &tmW1, &w_full[st], (2 * i + 1) * 64, c * 128 + (int)rank * 64);
            } else {         // W2 k-block i of chunk c: this CTA's 128 of the 256 output rows, [128 x 64] (16 KB)
              tma2_load_2d(dst, &tmW2, &w_full[st], c * 128 + i * 64, (int)rank * 128);
            }
          }
        }
        lt++;
      }
    }
  } else if (warp == kEpiW + 1) {
    if (leader) {
      // ------------------------------- MMA issuer (leader CTA, for both SMs) -------
      // The whole warp runs this code (warp-uniform control flow: descriptor arithmetic and barrier polls on the uniform datapath)
      // and ONE elected lane issues the tcgen05 instructions: a single lane of divergent code pays ~100 cycles per MMA
      // (profiles/r2/micro_mma_issue.txt), which for FF1's sixteen N = 128 MMAs per chunk (64 cycles each on the pipe) was the
      // chunk phase's bound (r2/ffn_trace_chained_2sm.txt: 1 700 cycles per FF1 against 1 024 nominal).
      // The 2-SM pipe retires these MMAs at the nominal rate (8 per 513 cycles in the trace), but its queue is shallow and
      // the issuing thread blocks on it, so every cycle the thread spends elsewhere between bursts is an idle tensor pipe: a
      // satisfied mbarrier poll alone costs ~120 cycles.  Hence (i) the op schedule is straight-line code per chunk, and
      // (ii) the barriers of the NEXT burst (weight slot, GELU chunk) are polled in the MIDDLE of the current burst, where
      // the thread would be blocked anyway; the blocking wait at the start of a burst is skipped when that poll succeeded.
      constexpr uint32_t idesc1 = umma_idesc_f16(256, 128, 0);   // FF1: 256 rows (2 x 128) x 128 hidden columns
      constexpr uint32_t idesc2 = umma_idesc_f16(256, 256, 0);   // FF2: 256 rows x 256 outputs
      const uint32_t h_addr = smem_u32(smem);
      const uint32_t w_base = smem_u32(smem + kOffW);
      int lt = 0, wit = 0, fcnt = 0;
      bool w_ready = false, f_ready = false;
      long long* tb = (blockIdx.x == 0 && lane == 0) ? p.trace : nullptr;
      int ti = 0;
      auto poll_next_slot = [&]() {
        const int n = wit + 1;
        w_ready = mbar_try_wait(&w_full[n % kSlots], (n / kSlots) & 1);
      };
      auto poll_f = [&]() { f_ready = mbar_try_wait(f_full, fcnt & 1); };
      auto slot_begin = [&]() -> uint32_t {   // returns the smem address of the current slot, data landed
        const int st = wit % kSlots;
        if (!w_ready) mbar_wait(&w_full[st], (wit / kSlots) & 1);
        w_ready = false;
        return w_base + st * kSlotBytes;
      };
      auto ff1 = [&](int c, bool poll_f_after) {
        const int b = c & 1;   // buffer b was last read by FF2(c-2), issued earlier on the same in-order pipe
#pragma unroll
        for (int i = 0; i < 2; i++) {
          const uint32_t w_addr = slot_begin();
          const uint64_t a_desc0 = umma_smem_desc_sw128(h_addr + (2 * i) * 16384), a_desc1 = umma_smem_desc_sw128(h_addr + (2 * i + 1) * 16384);
          const uint64_t b_desc0 = umma_smem_desc_sw128(w_addr), b_desc1 = umma_smem_desc_sw128(w_addr + 8192);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; k++)
              umma2_f16(tmem_base + kAcc1 + b * 128, a_desc0 + (uint64_t)(k * 2), b_desc0 + (uint64_t)(k * 2), idesc1, (i | k) != 0);
          }
          __syncwarp();
          poll_next_slot();               // middle of the burst: look ahead
          if (i == 1 && poll_f_after) poll_f();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; k++)
              umma2_f16(tmem_base + kAcc1 + b * 128, a_desc1 + (uint64_t)(k * 2), b_desc1 + (uint64_t)(k * 2), idesc1, 1u);
            umma2_commit(&w_empty[wit % kSlots]);
            if (i == 1) umma2_commit(&acc1_full[b]);
          }
          __syncwarp();
          wit++;
        }
        ffn_trace(tb, ti, 3);
      };
      int c_first = 0;   // first chunk of the unit: its FF2 overwrites the output accumulator
      auto ff2 = [&](int c, bool ff2_next = false) {
        ffn_trace(tb, ti, 4);
        if (!f_ready) mbar_wait(f_full, fcnt & 1);      // GELU chunk c (16-bit) is in TMEM in both CTAs
        f_ready = false;
        fcnt++;
        ffn_trace(tb, ti, 5);
        tc_fence_after();
        const uint32_t a_tmem = tmem_base + kAcc1 + (c & 1) * 128;
#pragma unroll
        for (int kb = 0; kb < 2; kb++) {
          const uint64_t b_desc = umma_smem_desc_sw128(slot_begin());
          if (elect_one()) {
            if (kb == 0) {
              mbar_arrive(f_seen);                      // back-pressure (see ffn_fused.cu), relayed to the peer
              mbar_arrive_rank(f_seen, 1);
            }
#pragma unroll
            for (int k = 0; k < 2; k++)
              umma2_f16_ts(tmem_base + kAcc2, a_tmem + kb * 32 + k * 8, b_desc + (uint64_t)(k * 2), idesc2, (CHAIN && HS == 1) ? 1u : (uint32_t)(((c - c_first) | kb | k) != 0));
          }
          __syncwarp();
          poll_next_slot();
          if (kb == 1 && (c == 6 || ff2_next)) poll_f();          // FF2(7) follows FF2(6) directly
          if (elect_one()) {
#pragma unroll
            for (int k = 2; k < 4; k++)
              umma2_f16_ts(tmem_base + kAcc2, a_tmem + kb * 32 + k * 8, b_desc + (uint64_t)(k * 2), idesc2, 1u);
            umma2_commit(&w_empty[wit % kSlots]);
          }
          __syncwarp();
          wit++;
        }
        ffn_trace(tb, ti, 6);
      };
      for (int unit = unit0; unit < total_tiles; unit += unit_step) {
        ffn_trace(tb, ti, 1);
        if (!CHAIN) {
          mbar_wait(h_full, lt & 1);
        } else {
          // out-projection into the 256 columns of the two FF1 accumulators: they are free -- the last readers, FF2(6) / FF2(7) of
          // the previous unit, were issued earlier on the same in-order pipe
          w_ready = false;
#pragma unroll 1
          for (int kb = 0; kb < 8; kb++) {
            const int stA = wit % kSlots;
            const uint64_t a_desc = umma_smem_desc_sw128(slot_begin());
            wit++;
            const int stB = wit % kSlots;
            const uint64_t b_desc = umma_smem_desc_sw128(slot_begin());
            wit++;
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; k++)
                umma2_f16(tmem_base + kAcc1, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc2, (kb | k) != 0);
              umma2_commit(&w_empty[stA]);
              umma2_commit(&w_empty[stB]);
              if (kb == 7) umma2_commit(op_full);
            }
            __syncwarp();
          }
          w_ready = false;
          mbar_wait(h_ready, lt & 1);       // both CTAs have added the residual and written LayerNorm3(x) as their H tile
          asm volatile("fence.acq_rel.cluster;" ::: "memory");
          tc_fence_after();
        }
        ffn_trace(tb, ti, 2);
        if constexpr (HS == 1) {
          // schedule: FF1(0) FF1(1) | FF2(0) FF1(2) | FF2(1) FF1(3) | ... | FF2(5) FF1(7) | FF2(6) FF2(7)
          ff1(0, false);
          ff1(1, true);
          mbar_wait(acc2_empty, (lt & 1) ^ 1);            // both CTAs' output epilogues have drained acc2
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < 6; c++) {
            ff2(c);
            ff1(c + 2, true);
          }
          if (elect_one()) umma2_commit(h_empty);         // all FF1 MMAs of this unit issued: the H tiles may be refilled
          __syncwarp();
          ff2(6);
          ff2(7);
        } else {
          static_assert(HS == 1 || kNC == 2, "split schedule is written for two chunks per unit");
          c_first = (unit % HS) * kNC;                    // FF1(c0) FF1(c0+1) | FF2(c0) FF2(c0+1)
          ff1(c_first, false);
          ff1(c_first + 1, true);
          if (elect_one()) umma2_commit(h_empty);
          __syncwarp();
          mbar_wait(acc2_empty, (lt & 1) ^ 1);
          tc_fence_after();
          ff2(c_first, true);
          ff2(c_first + 1);
        }
        if (elect_one()) umma2_commit(acc2_full);
        __syncwarp();
        lt++;
      }
    }
  } else {
    // --------------------------------- epilogue -----------------------------------
    const int ew = warp;
    const int q = warp & 3;
    const int part = ew >> 2;                // 0..3: which quarter of the columns
    const int r = q * 32 + lane;
    float* stg = nullptr;                    // (epilogue I/O is direct 256-bit global access: no staging tile)
    float *red_c, *red_d;                    // [4][128] each; two sets alternating by tile (one barrier per LayerNorm)
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    int lt = 0, use1[2] = {0, 0}, g = 0;
    long long* tb = (blockIdx.x == 0 && warp == 0 && lane == 0 && p.trace) ? p.trace + 4096 : nullptr;
    int ti = 0;
    for (int unit = unit0; unit < total_tiles; unit += unit_step) {
      int s, t0, len;
      const bool tile_ok = ffn2_tile(p, unit / HS, rank, row_tiles, s, t0, len);
      const int hs = unit % HS;
      const int t = t0 + r;
      const bool valid = t < len;
      const long long row = (long long)s * p.T_alloc + t;
      const long long row0 = row - lane;
      if (CHAIN) {
        // ---- chained out-projection: x' = acc + bo + x (kept in tensor memory as the FF2 accumulator's initial value), H = LayerNorm3(x') as the swizzled 16-bit A tile ----
        ffn_trace(tb, ti, 30);
        mbar_wait(op_full, lt & 1);
        tc_fence_after();
        mbar_wait(h_empty, (lt & 1) ^ 1);      // FF1(7) of the previous unit has read the H tile
        ffn_trace(tb, ti, 31);
        const uint32_t oaddr = lane_addr + kAcc1 + part * 64;
        float sum3 = 0.f, sq3 = 0.f;
        uint32_t raw[32];
        float v[32], tmp[32];
#pragma unroll 1
        for (int ch = 0; ch < 2; ch++) {
          const int cbase = part * 64 + ch * 32;
          tmem_ld32(oaddr + ch * 32, raw);
          lds32(vec_bo + cbase, tmp);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i++) v[i] = __uint_as_float(raw[i]) + tmp[i];
          tile_load_f32_h16(p.x32 + row0 * 256 + cbase, 256, stg, lane, tmp);
#pragma unroll
          for (int i = 0; i < 32; i++) v[i] += tmp[i];
          if (HS > 1 && tile_ok && hs == 0) {     // split: x32 stays untouched until the reduction; x' goes to its own buffer
            tile_store_f32_h16(p.xprime + row0 * 256 + cbase, 256, stg, lane, v);
          }
#pragma unroll
          for (int i = 0; i < 32; i++) {
            sum3 += v[i];
            sq3 = fmaf(v[i], v[i], sq3);
            raw[i] = __float_as_uint(v[i]);
          }
          tmem_st32(oaddr + ch * 32, raw);
          // x' is not written back to global memory: it becomes the INITIAL VALUE of the FF2 accumulator (same rows, same 256
          // columns; the previous unit's output epilogue -- these very warps -- has drained it), every FF2 MMA accumulates, and the
          // output epilogue finds x' + FF2 there: no fp32 store here, no residual load there (256 KB per tile, and their latency)
          if (HS == 1) tmem_st32(lane_addr + kAcc2 + cbase, raw);
        }
        tmem_st_wait();
        ffn_trace(tb, ti, 32);
        float* r3c = reinterpret_cast<float*>(smem + kOffRed2) + (lt & 1) * 1024;
        float* r3d = r3c + 512;
        r3c[part * 128 + r] = sum3;
        r3d[part * 128 + r] = sq3;
        ffn_bar();
        ffn_trace(tb, ti, 33);
        const float mean3 = (r3c[r] + r3c[128 + r] + r3c[256 + r] + r3c[384 + r]) * (1.f / 256.f);
        const float rstd3 = rsqrtf(fmaxf((r3d[r] + r3d[128 + r] + r3d[256 + r] + r3d[384 + r]) * (1.f / 256.f) - mean3 * mean3, 0.f) + p.ln3_eps);
        // K-major 128B-swizzled atom `part` (columns 64 part .. 64 part + 63): row r at r * 128 B, 16-byte chunk j at (j ^ (r & 7))
        uint8_t* hrow = smem + part * 16384 + r * 128;
#pragma unroll 1
        for (int ch = 0; ch < 2; ch++) {
          const int cbase = part * 64 + ch * 32;
          tmem_ld32(oaddr + ch * 32, raw);
          float gg[32], w[32];
          lds32(vec_g3 + cbase, gg);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i++) w[i] = (__uint_as_float(raw[i]) - mean3) * rstd3 * gg[i];
          lds32(vec_b3 + cbase, gg);
#pragma unroll
          for (int i = 0; i < 32; i++) w[i] = valid ? (w[i] + gg[i]) : 0.f;
#pragma unroll
          for (int j = 0; j < 4; j++) {
            uint4 u;
            __half2 h0 = __floats2half2_rn(w[j * 8 + 0], w[j * 8 + 1]), h1 = __floats2half2_rn(w[j * 8 + 2], w[j * 8 + 3]);
            __half2 h2 = __floats2half2_rn(w[j * 8 + 4], w[j * 8 + 5]), h3 = __floats2half2_rn(w[j * 8 + 6], w[j * 8 + 7]);
            u.x = *reinterpret_cast<uint32_t*>(&h0);
            u.y = *reinterpret_cast<uint32_t*>(&h1);
            u.z = *reinterpret_cast<uint32_t*>(&h2);
            u.w = *reinterpret_cast<uint32_t*>(&h3);
            *reinterpret_cast<uint4*>(hrow + (((ch * 4 + j) ^ (r & 7)) << 4)) = u;
          }
        }
        fence_proxy_async_smem();            // generic-proxy writes of H -> visible to the tensor core's async-proxy reads
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader_release(h_ready);
        ffn_trace(tb, ti, 34);
      }
      // ---- 8 hidden chunks: bias + GELU -> 16-bit chunk in shared memory (A operand of FF2) ----
      for (int c = hs * kNC; c < hs * kNC + kNC; c++, g++) {
        const int b = c & 1;
        float bv[32];
        lds32(vec_b1 + c * 128 + part * 32, bv);
        ffn_trace(tb, ti, 10);
        mbar_wait(&acc1_full[b], use1[b] & 1);
        ffn_trace(tb, ti, 11);
        use1[b]++;
        tc_fence_after();
        uint32_t raw[32];
        tmem_ld32(lane_addr + kAcc1 + b * 128 + part * 32, raw);
        tmem_ld_wait();
        // the 16-bit chunk lands on columns [16*part, 16*part+16) of this buffer = fp32 columns of part/2: the four warps
        // of a lane quarter must all hold their accumulator slice in registers before any of them writes
        asm volatile("bar.sync %0, 128;" ::"r"(2 + q) : "memory");
        ffn_trace(tb, ti, 12);
#pragma unroll
        for (int i = 0; i < 32; i += 2) {   // bias + GELU on packed fp32 pairs, straight to 16-bit pairs
          const float2 g2 = fast_gelu_erf2(fadd2(make_float2(__uint_as_float(raw[i]), __uint_as_float(raw[i + 1])),
                                                 make_float2(bv[i], bv[i + 1])));
          __half2 h2 = __floats2half2_rn(g2.x, g2.y);
          raw[i >> 1] = *reinterpret_cast<uint32_t*>(&h2);
        }
        tmem_st16(lane_addr + kAcc1 + b * 128 + part * 16, raw);
        tmem_st_wait();
        tc_fence_before();
        ffn_trace(tb, ti, 13);
        if (g > 0) mbar_wait(f_seen, (g - 1) & 1);              // never two unobserved phases of f_full (robust to any warp skew)
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(f_full);
        ffn_trace(tb, ti, 14);
      }
      ffn_trace(tb, ti, 20);
      // ---- output tile: + b2 + residual -> X32 ; LayerNorm / plain emits ----
      mbar_wait(acc2_full, lt & 1);
      tc_fence_after();
      ffn_trace(tb, ti, 21);
      const uint32_t taddr = lane_addr + kAcc2 + part * 64;
      const bool want_ln = tile_ok && p.emit_ln.ptr != nullptr;
      if (!tile_ok) {   // filler tile of an odd tail: nothing to store, just hand acc2 back
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(acc2_empty);
        lt++;
        continue;
      }
      if constexpr (HS > 1) {   // split: this pair's partial FF2 sum goes to its slab; bias, residual and LayerNorm happen in ffn_reduce
        float* slab = p.slabs + (long long)hs * p.S * p.T_alloc * 256;
#pragma unroll 1
        for (int ch = 0; ch < 2; ch++) {
          const int cbase = part * 64 + ch * 32;
          uint32_t raw[32];
          tmem_ld32(taddr + ch * 32, raw);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; i++) v[i] = __uint_as_float(raw[i]);
          if (tile_ok) tile_store_f32_h16(slab + row0 * 256 + cbase, 256, stg, lane, v);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(acc2_empty);
        lt++;
        continue;
      }
      float sum2 = 0.f, sq2 = 0.f;
      uint32_t raw[32];
      float v[32], tmp[32];
#pragma unroll 1
      for (int ch = 0; ch < 2; ch++) {
        const int cbase = part * 64 + ch * 32;
        tmem_ld32(taddr + ch * 32, raw);
        lds32(vec_b2 + cbase, tmp);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i++) v[i] = __uint_as_float(raw[i]) + tmp[i];
        ffn_trace(tb, ti, 23);
        if (!(CHAIN && HS == 1)) {   // (chained: the accumulator was initialised with x', see the mid-epilogue)
          tile_load_f32_h16(p.x32 + row0 * 256 + cbase, 256, stg, lane, tmp);
#pragma unroll
          for (int i = 0; i < 32; i++) v[i] += tmp[i];
        }
        ffn_trace(tb, ti, 24, v[0]);
        tile_store_f32_h16(p.x32 + row0 * 256 + cbase, 256, stg, lane, v);
        ffn_trace(tb, ti, 25);
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const Emit& em = p.emit_plain[e];
          if (!em.ptr) continue;
          float w[32];
#pragma unroll
          for (int i = 0; i < 32; i++) w[i] = valid ? v[i] : 0.f;
          tile_store_f16_h16(em.ptr + row0 * em.ld + em.col_off + cbase, em.ld, stg, lane, w);
        }
        if (want_ln) {
#pragma unroll
          for (int i = 0; i < 32; i++) {
            sum2 += v[i];
            sq2 = fmaf(v[i], v[i], sq2);
            raw[i] = __float_as_uint(v[i]);
          }
          tmem_st32(taddr + ch * 32, raw);
        }
      }
      if (want_ln) {
        tmem_st_wait();
        // one sweep: sum and sum of squares were taken while the row was written back; var = E[x^2] - mean^2 (fp32, 256 values)
        red_c = red + (lt & 1) * 1024;
        red_d = red_c + 512;
        red_c[part * 128 + r] = sum2;
        red_d[part * 128 + r] = sq2;
        ffn_trace(tb, ti, 26);
        ffn_bar();
        ffn_trace(tb, ti, 27);
        const float mean2 = (red_c[r] + red_c[128 + r] + red_c[256 + r] + red_c[384 + r]) * (1.f / 256.f);
        const float rstd2 = rsqrtf(fmaxf((red_d[r] + red_d[128 + r] + red_d[256 + r] + red_d[384 + r]) * (1.f / 256.f) - mean2 * mean2, 0.f) + p.emit_ln.f);
#pragma unroll 1
        for (int ch = 0; ch < 2; ch++) {
          const int cbase = part * 64 + ch * 32;
          tmem_ld32(taddr + ch * 32, raw);
          float gg[32], w[32];
          lds32(vec_g + cbase, gg);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i++) w[i] = (__uint_as_float(raw[i]) - mean2) * rstd2 * gg[i];
          lds32(vec_b + cbase, gg);
#pragma unroll
          for (int i = 0; i < 32; i++) w[i] = valid ? (w[i] + gg[i]) : 0.f;
          tile_store_f16_h16(p.emit_ln.ptr + row0 * p.emit_ln.ld + p.emit_ln.col_off + cbase, p.emit_ln.ld, stg, lane, w);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(acc2_empty);
      ffn_trace(tb, ti, 22);
      lt++;
    }
  }

  tc_fence_before();
  cluster_sync();   // no CTA may exit (or free TMEM) while the leader can still issue MMAs into it or signal its barriers
  if (warp == kEpiW + 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

extern long long* g_ffn_trace_ptr();

template <bool CHAIN>
static void launch_ffn2(const CUtensorMap& tmH, const CUtensorMap& tmWo, const CUtensorMap& tmW1h, const CUtensorMap& tmW2h,
                        const FfnParams& p_in, cudaStream_t stream) {
  static PerDeviceOnce once;
  static int max_clusters = 0;   // every device of a box is the same part
  FfnParams p = p_in;
  p.trace = g_ffn_trace_ptr();
  CV2_CHECK(p.tile_list && p.tile_count && p.lens, "ffn_fused2: the 2-SM path needs the compact tile list");
  CV2_CHECK(!CHAIN || (p.bo && p.ln3_g && p.ln3_b), "ffn_fused2: chained out-projection needs its bias and LayerNorm3 vectors");
  cudaLaunchConfig_t q = {};
  q.blockDim = dim3(kFfnThreads);
  q.dynamicSmemBytes = kFfnSmem;
  q.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  q.attrs = at;
  q.numAttrs = 1;
  once.run([&] {
    CV2_CUDA(cudaFuncSetAttribute(ffn_fused2_kernel<CHAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFfnSmem));
    int dev = 0, sms = 0;
    CV2_CUDA(cudaGetDevice(&dev));
    CV2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    q.gridDim = dim3(sms / 2 * 2);
    CV2_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, ffn_fused2_kernel<CHAIN>, &q));
    CV2_CHECK(max_clusters > 0, "ffn_fused2: no 2-CTA cluster fits");
  });
  CV2_CHECK(p.T_alloc % 128 == 0, "ffn_fused2: T_alloc %d not a multiple of 128", p.T_alloc);
  const int units = ((p.T_alloc / 128) * p.S + 1) / 2;
  const int clusters = units < max_clusters ? units : max_clusters;
  q.gridDim = dim3(2 * clusters);
  CV2_CUDA(cudaLaunchKernelEx(&q, ffn_fused2_kernel<CHAIN>, tmH, tmWo, tmW1h, tmW2h, p));
  CV2_LAUNCH_CHECK();
}

void launch_ffn_fused2(const CUtensorMap& tmH, const CUtensorMap& tmW1h, const CUtensorMap& tmW2h, const FfnParams& p,
                       cudaStream_t stream) {
  launch_ffn2<false>(tmH, tmW1h /* unused */, tmW1h, tmW2h, p, stream);
}

void launch_ffn_fused2_chain(const CUtensorMap& tmATT, const CUtensorMap& tmWoh, const CUtensorMap& tmW1h, const CUtensorMap& tmW2h,
                             const FfnParams& p, cudaStream_t stream) {
  launch_ffn2<true>(tmATT, tmWoh, tmW1h, tmW2h, p, stream);
}

// ---------------------------------------------------------------------------------------------------------
// Hidden split for small launches (FfnParams::hsplit): the reduction of the partial FF2 sums.
//   x = x' + sum_h slab_h + b2 -> x32 ; masked 16-bit emits / LayerNorm emit of the result (what the unsplit kernel's output
//   epilogue does).  One warp per row (8 consecutive columns per lane: 2 x 128-bit per operand), 8 rows per CTA: a batch-1 call
//   has ~2 300 rows, and the kernel is a latency chain per row, so it wants every row in flight at once.
// ---------------------------------------------------------------------------------------------------------
template <int HS>
__global__ void __launch_bounds__(256) ffn_reduce_kernel(const FfnParams p) {
  const int tile = blockIdx.x >> 4;
  if (tile >= __ldg(p.tile_count)) return;
  const int s = __ldg(p.tile_list + 2 * tile), t0 = __ldg(p.tile_list + 2 * tile + 1);
  const int len = __ldg(p.lens + s);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long slab_stride = (long long)p.S * p.T_alloc * 256;
  float bias[8], g[8], be[8];
#pragma unroll
  for (int i = 0; i < 8; i++) bias[i] = __ldg(p.b2 + lane * 8 + i);
  const bool want_ln = p.emit_ln.ptr != nullptr;
  if (want_ln) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      g[i] = __ldg(p.emit_ln.a + lane * 8 + i);
      be[i] = __ldg(p.emit_ln.b + lane * 8 + i);
    }
  }
  {
    const int t = t0 + (blockIdx.x & 15) * 8 + warp;
    const long long row = (long long)s * p.T_alloc + t;
    const bool valid = t < len;
    float v[8];
    {
      const float4* src = reinterpret_cast<const float4*>(p.xprime + row * 256 + lane * 8);
      const float4 a = src[0], b = src[1];
      v[0] = a.x + bias[0]; v[1] = a.y + bias[1]; v[2] = a.z + bias[2]; v[3] = a.w + bias[3];
      v[4] = b.x + bias[4]; v[5] = b.y + bias[5]; v[6] = b.z + bias[6]; v[7] = b.w + bias[7];
    }
#pragma unroll
    for (int h = 0; h < HS; h++) {
      const float4* src = reinterpret_cast<const float4*>(p.slabs + h * slab_stride + row * 256 + lane * 8);
      const float4 a = src[0], b = src[1];
      v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w;
      v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
    }
    float4* dst = reinterpret_cast<float4*>(p.x32 + row * 256 + lane * 8);
    dst[0] = make_float4(v[0], v[1], v[2], v[3]);
    dst[1] = make_float4(v[4], v[5], v[6], v[7]);
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const Emit& em = p.emit_plain[e];
      if (!em.ptr) continue;
      uint32_t pk[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        __half2 hh = __floats2half2_rn(valid ? v[2 * i] : 0.f, valid ? v[2 * i + 1] : 0.f);
        pk[i] = *reinterpret_cast<uint32_t*>(&hh);
      }
      *reinterpret_cast<uint4*>(em.ptr + row * em.ld + em.col_off + lane * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
    if (want_ln) {
      float sum = 0.f, sq = 0.f;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        sum += v[i];
        sq = fmaf(v[i], v[i], sq);
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, d);
        sq += __shfl_xor_sync(0xffffffffu, sq, d);
      }
      const float mean = sum * (1.f / 256.f);
      const float rstd = rsqrtf(fmaxf(sq * (1.f / 256.f) - mean * mean, 0.f) + p.emit_ln.f);
      uint32_t pk[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const float w0 = valid ? (v[2 * i] - mean) * rstd * g[2 * i] + be[2 * i] : 0.f;
        const float w1 = valid ? (v[2 * i + 1] - mean) * rstd * g[2 * i + 1] + be[2 * i + 1] : 0.f;
        __half2 hh = __floats2half2_rn(w0, w1);
        pk[i] = *reinterpret_cast<uint32_t*>(&hh);
      }
      *reinterpret_cast<uint4*>(p.emit_ln.ptr + row * p.emit_ln.ld + p.emit_ln.col_off + lane * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
  }
}

static int g_ffn2_max_pairs = 0;
int ffn_fused2_max_pairs() { return g_ffn2_max_pairs; }

void launch_ffn_fused2_chain_split(const CUtensorMap& tmATT, const CUtensorMap& tmWoh, const CUtensorMap& tmW1h, const CUtensorMap& tmW2h,
                                   const FfnParams& p_in, cudaStream_t stream) {
  constexpr int HS = 4;
  static PerDeviceOnce once;
  FfnParams p = p_in;
  p.trace = nullptr;
  CV2_CHECK(p.hsplit == HS && p.slabs && p.xprime, "ffn_fused2 split: hsplit must be %d with slab / x' scratch", HS);
  CV2_CHECK(p.tile_list && p.tile_count && p.lens && p.bo && p.ln3_g && p.ln3_b, "ffn_fused2 split: chained form with a compact tile list only");
  CV2_CHECK(p.T_alloc % 128 == 0, "ffn_fused2: T_alloc %d not a multiple of 128", p.T_alloc);
  cudaLaunchConfig_t q = {};
  q.blockDim = dim3(kFfnThreads);
  q.dynamicSmemBytes = kFfnSmem;
  q.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  q.attrs = at;
  q.numAttrs = 1;
  once.run([&] {
    CV2_CUDA(cudaFuncSetAttribute(ffn_fused2_kernel<true, HS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFfnSmem));
    int dev = 0, sms = 0;
    CV2_CUDA(cudaGetDevice(&dev));
    CV2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    q.gridDim = dim3(sms / 2 * 2);
    CV2_CUDA(cudaOccupancyMaxActiveClusters(&g_ffn2_max_pairs, ffn_fused2_kernel<true, HS>, &q));
    CV2_CHECK(g_ffn2_max_pairs > 0, "ffn_fused2: no 2-CTA cluster fits");
  });
  const int row_tiles = (p.T_alloc / 128) * p.S;
  const int units = ((row_tiles + 1) / 2) * HS;
  const int clusters = units < g_ffn2_max_pairs ? units : g_ffn2_max_pairs;
  q.gridDim = dim3(2 * clusters);
  CV2_CUDA(cudaLaunchKernelEx(&q, ffn_fused2_kernel<true, HS>, tmATT, tmWoh, tmW1h, tmW2h, p));
  CV2_LAUNCH_CHECK();
  ffn_reduce_kernel<HS><<<row_tiles * 16, 256, 0, stream>>>(p);
  CV2_LAUNCH_CHECK();
}

}  // namespace cv2
