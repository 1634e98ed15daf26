// Fused transformer feed-forward of the CFM estimator (BasicTransformerBlock FeedForward, matcha transformer.py:83-134,
// 291-314):   x <- x + W2 * gelu(W1 * h + b1) + b2 ,  h = LayerNorm3(x) (16-bit, produced by the out-proj epilogue),
// then emit the next block's pre-norm LayerNorm1'(x) (or the masked plain block output) as the next 16-bit A operand.
//
// The 1024-wide GELU activation never leaves the SM: per 128-row tile the hidden dimension is processed in 8 chunks of
// 128; FF1(c) accumulates into a double-buffered 128-column TMEM tile, the epilogue warps apply bias + GELU and write the
// 16-bit chunk back into the first 64 columns of the SAME TMEM buffer, where it is the A operand of FF2(c) (tcgen05.mma with
// a tensor-memory A operand), which accumulates the 256-column output tile in TMEM across the 8 chunks.  No shared-memory
// hand-off, no generic->async proxy fence, and the hidden chunk is double-buffered with its accumulator.  Buffer reuse needs
// no barrier: FF1(c+2) is issued behind FF2(c) on the in-order tensor pipe.  Weights stream through a 9-slot TMA ring.
//   warps 0-15: epilogue (four per TMEM lane quarter)   warp 16: TMA (H tile + weight stream)   warp 17: tcgen05.mma issuer.
// The issuing warps carry the HIGHEST warp ids: the scheduler favours high ids, and a single issuing thread that loses
// arbitration against 16 math-heavy warps paces the tensor pipe (measured with the clock64 trace, profiles/ffn_trace.py:
// 112-144 cycles per MMA issued instead of 64 when the issuer was warp 1).
// Saves the [rows,1024] 16-bit round trip through HBM (4 KB/row of the 18 KB/row a transformer block moves) and one launch.
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
static constexpr int kOffBar = kOffRed + 2 * 2 * 4 * 128 * 4;   // (sum, sum of squares) x [4][128], double-buffered by tile parity
static constexpr int kFfnSmem = kOffBar + 256;
static constexpr int kFfnThreads = 64 + kEpiW * 32;
static constexpr uint32_t kAcc1 = 0, kAcc2 = 256;   // TMEM columns: acc1 = 2 x 128, acc2 = 256

__device__ __forceinline__ void ffn_trace(long long* buf, int& idx, int code) {
  if (buf && idx < 4095) buf[idx++] = (clock64() << 8) | code;
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

// op o of a tile's schedule: FF1(0), FF1(1), FF2(0), FF1(2), FF2(1), ..., FF1(7), FF2(6), FF2(7)
__device__ __forceinline__ void ffn_op(int o, bool& is_ff2, int& c) {
  if (o < 2) { is_ff2 = false; c = o; }
  else if (o == 15) { is_ff2 = true; c = 7; }
  else if (o & 1) { is_ff2 = false; c = (o + 1) >> 1; }
  else { is_ff2 = true; c = (o >> 1) - 1; }
}

__global__ void __launch_bounds__(kFfnThreads, 1)
ffn_fused_kernel(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmW2, const FfnParams p) {
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
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc2_empty + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int t_tiles = p.T_alloc / 128;
  const int total_tiles = p.tile_list ? __ldg(p.tile_count) : t_tiles * p.S;

  if (warp == kEpiW && lane == 0) {
    tma_prefetch_desc(&tmH);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    mbar_init(h_full, 1);
    mbar_init(h_empty, 1);
    for (int i = 0; i < kSlots; i++) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    for (int i = 0; i < 2; i++) {
      mbar_init(&acc1_full[i], 1);
      mbar_init(&acc1_empty[i], kEpiW);
    }
    mbar_init(f_full, kEpiW);
    mbar_init(f_seen, 1);
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, kEpiW);
    fence_barrier_init();
  }
  if (warp == kEpiW + 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kEpiW) {
    if (lane == 0) {
      // ------------------------------- TMA producer -------------------------------
      int lt = 0, wit = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int s, t0, len;
        if (!ffn_tile(p, tile, t_tiles, s, t0, len)) continue;
        // the output epilogue adds the fp32 residual tile (128 rows x 1 KB, contiguous): pull it into L2 now, ~60 us early
        prefetch_l2_bulk(p.x32 + ((long long)s * p.T_alloc + t0) * 256, 128 * 256 * 4);
        mbar_wait(h_empty, (lt & 1) ^ 1);
        mbar_expect_tx(h_full, kHBytes);
#pragma unroll
        for (int kb = 0; kb < 4; kb++) tma_load_3d(smem + kb * 16384, &tmH, h_full, kb * 64, t0, s);
        for (int o = 0; o < 16; o++) {
          bool is_ff2;
          int c;
          ffn_op(o, is_ff2, c);
          for (int i = 0; i < 4; i++, wit++) {
            const int st = wit % kSlots;
            mbar_wait(&w_empty[st], ((wit / kSlots) & 1) ^ 1);
            mbar_expect_tx(&w_full[st], kSlotBytes);
            uint8_t* dst = smem + kOffW + st * kSlotBytes;
            if (!is_ff2) tma_load_2d(dst, &tmW1, &w_full[st], i * 64, c * 128);                       // W1[c*128.., kb*64..]
            else tma_load_2d(dst, &tmW2, &w_full[st], c * 128 + (i >> 1) * 64, (i & 1) * 128);       // W2[half*128.., c*128+kb*64..]
          }
        }
        lt++;
      }
    }
  } else if (warp == kEpiW + 1) {
    if (lane == 0) {
      // ------------------------------- MMA issuer ---------------------------------
      constexpr uint32_t idesc = umma_idesc_f16(128, 128, 0);
      const uint32_t h_addr = smem_u32(smem);
      int lt = 0, wit = 0, fcnt = 0;
      long long* tb = blockIdx.x == 0 ? p.trace : nullptr;
      int ti = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int s, t0, len;
        if (!ffn_tile(p, tile, t_tiles, s, t0, len)) continue;
        ffn_trace(tb, ti, 1);
        mbar_wait(h_full, lt & 1);
        tc_fence_after();
        ffn_trace(tb, ti, 2);
        for (int o = 0; o < 16; o++) {
          bool is_ff2;
          int c;
          ffn_op(o, is_ff2, c);
          if (!is_ff2) {
            const int b = c & 1;   // buffer b was last read by FF2(c-2), issued earlier on the same in-order pipe
            // The tensor-pipe queue is shallow: whatever the issuing thread does between the last MMA of one slot and the first
            // MMA of the next is a bubble (measured 450 cycles per 4-MMA slot instead of 256, profiles/micro/mma_bubble.cu).  So
            // the bookkeeping for slot i+1 (full-barrier poll, descriptor arithmetic) is done right after the FIRST MMA of slot
            // i, while the pipe is busy and the thread would be blocked on issue anyway.
            int st = wit % kSlots;
            mbar_wait(&w_full[st], (wit / kSlots) & 1);   // (TMA -> MMA through smem: the mbarrier wait orders it, no tcgen05 fence)
            uint64_t b_desc = umma_smem_desc_sw128(smem_u32(smem + kOffW + st * kSlotBytes));
            for (int kb = 0; kb < 4; kb++, wit++) {
              const uint64_t a_desc = umma_smem_desc_sw128(h_addr + kb * 16384);
              umma_f16(tmem_base + kAcc1 + b * 128, a_desc, b_desc, idesc, kb != 0);
              int st_n = st;
              uint64_t b_next = b_desc;
              if (kb < 3) {
                st_n = (wit + 1) % kSlots;
                mbar_wait(&w_full[st_n], ((wit + 1) / kSlots) & 1);
                b_next = umma_smem_desc_sw128(smem_u32(smem + kOffW + st_n * kSlotBytes));
              }
#pragma unroll
              for (int k = 1; k < 4; k++)
                umma_f16(tmem_base + kAcc1 + b * 128, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, 1);
              umma_commit(&w_empty[st]);
              st = st_n;
              b_desc = b_next;
            }
            umma_commit(&acc1_full[b]);
            ffn_trace(tb, ti, 3);
            if (c == 7) umma_commit(h_empty);                  // all FF1 MMAs of this tile issued: H tile may be refilled
          } else {
            if (c == 0) {
              mbar_wait(acc2_empty, (lt & 1) ^ 1);              // previous tile's output epilogue has drained acc2
              tc_fence_after();
            }
            ffn_trace(tb, ti, 4);
            mbar_wait(f_full, fcnt & 1);                        // GELU chunk c (16-bit) is in TMEM, on top of its accumulator
            ffn_trace(tb, ti, 5);
            mbar_arrive(f_seen);                                // back-pressure: chunk c+1 may only be signalled after this observation
            tc_fence_after();
            const uint32_t a_tmem = tmem_base + kAcc1 + (c & 1) * 128;
            // W2 tile of a k-block = two adjacent ring slots (output rows 0-127 | 128-255; every op takes 4 slots and kSlots is
            // even, so the pair never wraps): ONE N = 256 MMA per k-step reads the TMEM A operand once for all 256 outputs
            constexpr uint32_t idesc256 = umma_idesc_f16(128, 256, 0);
            int st = wit % kSlots;
            mbar_wait(&w_full[st], (wit / kSlots) & 1);
            mbar_wait(&w_full[st + 1], ((wit + 1) / kSlots) & 1);
            uint64_t b_desc = umma_smem_desc_sw128(smem_u32(smem + kOffW + st * kSlotBytes));
            for (int kb = 0; kb < 2; kb++, wit += 2) {
              umma_f16_ts(tmem_base + kAcc2, a_tmem + kb * 32, b_desc, idesc256, (c | kb) != 0);
              int st_n = st;
              uint64_t b_next = b_desc;
              if (kb == 0) {   // next pair's readiness + descriptor while the pipe works on this one
                st_n = (wit + 2) % kSlots;
                mbar_wait(&w_full[st_n], ((wit + 2) / kSlots) & 1);
                mbar_wait(&w_full[st_n + 1], ((wit + 3) / kSlots) & 1);
                b_next = umma_smem_desc_sw128(smem_u32(smem + kOffW + st_n * kSlotBytes));
              }
#pragma unroll
              for (int k = 1; k < 4; k++)
                umma_f16_ts(tmem_base + kAcc2, a_tmem + kb * 32 + k * 8, b_desc + (uint64_t)(k * 2), idesc256, 1);
              umma_commit(&w_empty[st]);
              umma_commit(&w_empty[st + 1]);
              st = st_n;
              b_desc = b_next;
            }
            fcnt++;
            ffn_trace(tb, ti, 6);
            if (c == 7) umma_commit(acc2_full);
          }
        }
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
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int s, t0, len;
      if (!ffn_tile(p, tile, t_tiles, s, t0, len)) continue;
      const int t = t0 + r;
      const bool valid = t < len;
      const long long row = (long long)s * p.T_alloc + t;
      const long long row0 = row - lane;
      // ---- 8 hidden chunks: bias + GELU -> 16-bit chunk in shared memory (A operand of FF2) ----
      for (int c = 0; c < 8; c++, g++) {
        const int b = c & 1;
        float bv[32];
        load32(p.b1 + c * 128 + part * 32, bv, true, 32);      // bias first: its latency hides behind the accumulator wait
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
        if (lane == 0) mbar_arrive(f_full);
        ffn_trace(tb, ti, 14);
      }
      ffn_trace(tb, ti, 20);
      // ---- output tile: + b2 + residual -> X32 ; LayerNorm / plain emits ----
      mbar_wait(acc2_full, lt & 1);
      tc_fence_after();
      ffn_trace(tb, ti, 21);
      const uint32_t taddr = lane_addr + kAcc2 + part * 64;
      const bool want_ln = p.emit_ln.ptr != nullptr;
      float sum2 = 0.f, sq2 = 0.f;
      uint32_t raw[32];
      float v[32], tmp[32];
#pragma unroll 1
      for (int ch = 0; ch < 2; ch++) {
        const int cbase = part * 64 + ch * 32;
        tmem_ld32(taddr + ch * 32, raw);
        load32(p.b2 + cbase, tmp, true, 32);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i++) v[i] = __uint_as_float(raw[i]) + tmp[i];
        tile_load_f32_h16(p.x32 + row0 * 256 + cbase, 256, stg, lane, tmp);
#pragma unroll
        for (int i = 0; i < 32; i++) v[i] += tmp[i];
        tile_store_f32_h16(p.x32 + row0 * 256 + cbase, 256, stg, lane, v);
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
        ffn_bar();
        const float mean2 = (red_c[r] + red_c[128 + r] + red_c[256 + r] + red_c[384 + r]) * (1.f / 256.f);
        const float rstd2 = rsqrtf(fmaxf((red_d[r] + red_d[128 + r] + red_d[256 + r] + red_d[384 + r]) * (1.f / 256.f) - mean2 * mean2, 0.f) + p.emit_ln.f);
#pragma unroll 1
        for (int ch = 0; ch < 2; ch++) {
          const int cbase = part * 64 + ch * 32;
          tmem_ld32(taddr + ch * 32, raw);
          float gg[32], w[32];
          load32(p.emit_ln.a + cbase, gg, true, 32);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i++) w[i] = (__uint_as_float(raw[i]) - mean2) * rstd2 * gg[i];
          load32(p.emit_ln.b + cbase, gg, true, 32);
#pragma unroll
          for (int i = 0; i < 32; i++) w[i] = valid ? (w[i] + gg[i]) : 0.f;
          tile_store_f16_h16(p.emit_ln.ptr + row0 * p.emit_ln.ld + p.emit_ln.col_off + cbase, p.emit_ln.ld, stg, lane, w);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc2_empty);
      ffn_trace(tb, ti, 22);
      lt++;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kEpiW + 1) tmem_dealloc<512>(tmem_base);
}

static long long* g_ffn_trace = nullptr;
void ffn_set_trace(long long* dev_buf) { g_ffn_trace = dev_buf; }
long long* g_ffn_trace_ptr() { return g_ffn_trace; }

void launch_ffn_fused(const CUtensorMap& tmH, const CUtensorMap& tmW1, const CUtensorMap& tmW2, const FfnParams& p_in,
                      cudaStream_t stream) {
  FfnParams p = p_in;
  p.trace = g_ffn_trace;
  static PerDeviceOnce once;
  static int num_sms = 0;      // every device of a box is the same part
  once.run([] {
    CV2_CUDA(cudaFuncSetAttribute(ffn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFfnSmem));
    int dev = 0;
    CV2_CUDA(cudaGetDevice(&dev));
    CV2_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  });
  CV2_CHECK(p.T_alloc % 128 == 0, "ffn_fused: T_alloc %d not a multiple of 128", p.T_alloc);
  const int total = (p.T_alloc / 128) * p.S;
  const int grid = total < num_sms ? total : num_sms;
  ffn_fused_kernel<<<grid, kFfnThreads, kFfnSmem, stream>>>(tmH, tmW1, tmW2, p);
  CV2_LAUNCH_CHECK();
}

}  // namespace cv2
