// Tap GEMM kernel (see gemm_tap.cuh).  tcgen05.mma (kind::f16, fp32 accumulate in TMEM), operands staged by
// TMA into 128B-swizzled shared memory through a 4-stage mbarrier ring.
#include "common.cuh"
#include "gemm_tap.cuh"
#include "host_util.h"

namespace cv2 {

static constexpr int kStages = 4;
static constexpr int kTileM = 128;
static constexpr int kKBlock = 64;                       // 64 x 16-bit = one 128B swizzle row
static constexpr int kABytes = kTileM * kKBlock * 2;     // 16 KB

template <int BN>
struct GemmSmem {
  static constexpr int kBBytes = BN * kKBlock * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOff = kStages * kStageBytes;
  static constexpr int kTotal = kBarOff + 128 + 1024;    // barriers + alignment slack
};

__device__ __forceinline__ float apply_act(float v, int act, float f, float a) {
  switch (act) {
    case ACT_MISH: return mish_f(v);
    case ACT_GELU: return gelu_erf_f(v);
    case ACT_SILU: return silu_f(v);
    case ACT_ELU: return elu_f(v);
    case ACT_LRELU: return v > 0.f ? v : v * f;
    case ACT_SNAKE: return snake_f(v, a);
    default: return v;
  }
}

__device__ __forceinline__ void store_h32(__half* dst, const float* v, bool full, int nvalid) {
  if (full) {
    uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      __half2 h0 = __floats2half2_rn(v[i * 8 + 0], v[i * 8 + 1]);
      __half2 h1 = __floats2half2_rn(v[i * 8 + 2], v[i * 8 + 3]);
      __half2 h2 = __floats2half2_rn(v[i * 8 + 4], v[i * 8 + 5]);
      __half2 h3 = __floats2half2_rn(v[i * 8 + 6], v[i * 8 + 7]);
      uint4 u;
      u.x = *reinterpret_cast<uint32_t*>(&h0);
      u.y = *reinterpret_cast<uint32_t*>(&h1);
      u.z = *reinterpret_cast<uint32_t*>(&h2);
      u.w = *reinterpret_cast<uint32_t*>(&h3);
      d4[i] = u;
    }
  } else {
    for (int i = 0; i < nvalid; i++) dst[i] = __float2half_rn(v[i]);
  }
}

template <int BN>
__global__ void __launch_bounds__(192, 1)
gemm_tap_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using SM = GemmSmem<BN>;
  const int n0 = blockIdx.x * BN;
  const int t0 = blockIdx.y * kTileM;
  const int s = blockIdx.z;
  const int len = p.lens ? p.lens[s] : p.len_all;
  if (t0 >= len + p.halo) return;  // tile entirely in the padding: nothing reads it

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SM::kBarOff);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_it = p.ntaps * p.kb_per_tap;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < kStages; i++) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<(BN < 32 ? 32 : BN)>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------- TMA producer -------------------------------
      for (int it = 0; it < num_it; it++) {
        const int st = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        mbar_wait(&empty_bar[st], ph ^ 1);
        const int tap = it / p.kb_per_tap;
        const int kb = it - tap * p.kb_per_tap;
        uint8_t* a_dst = smem + st * SM::kStageBytes;
        uint8_t* b_dst = a_dst + kABytes;
        mbar_expect_tx(&full_bar[st], SM::kStageBytes);
        tma_load_3d(a_dst, &tmA, &full_bar[st], kb * kKBlock, t0 + p.tap_off[tap], s);
        tma_load_2d(b_dst, &tmB, &full_bar[st], it * kKBlock, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------- MMA issuer ---------------------------------
      constexpr uint32_t idesc = umma_idesc_f16(kTileM, BN, 0);
      for (int it = 0; it < num_it; it++) {
        const int st = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        mbar_wait(&full_bar[st], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + st * SM::kStageBytes);
        const uint64_t a_desc = umma_smem_desc_sw128(a_addr);
        const uint64_t b_desc = umma_smem_desc_sw128(a_addr + kABytes);
#pragma unroll
        for (int k = 0; k < kKBlock / 16; k++) {
          // advance 16 elements (32 B) along K inside the swizzle atom: +2 in the (addr >> 4) field
          umma_f16(tmem_base, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (it | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[st]);  // frees this smem stage when the MMAs have read it
      }
      umma_commit(tmem_full);  // accumulator complete
    }
  } else {
    // --------------------------------- epilogue -----------------------------------
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;       // row inside the tile
    const int t = t0 + r;
    const bool valid = t < len;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int ncols = min(BN, p.N - n0);           // valid columns in this tile
    const int nchunks = (ncols + 31) / 32;
    mbar_wait(tmem_full, 0);
    tc_fence_after();

    float mean = 0.f, rstd = 1.f;
    uint32_t raw[32];
    if (p.ln) {  // LayerNorm over the N columns of this row (requires the whole row in one tile)
      float sum = 0.f;
      for (int c = 0; c < nchunks; c++) {
        tmem_ld32(taddr + c * 32, raw);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i++) {
          const int col = c * 32 + i;
          if (col < ncols) sum += __uint_as_float(raw[i]) + (p.bias ? __ldg(p.bias + n0 + col) : 0.f);
        }
      }
      mean = sum / (float)ncols;
      float sq = 0.f;
      for (int c = 0; c < nchunks; c++) {
        tmem_ld32(taddr + c * 32, raw);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i++) {
          const int col = c * 32 + i;
          if (col < ncols) {
            const float d = __uint_as_float(raw[i]) + (p.bias ? __ldg(p.bias + n0 + col) : 0.f) - mean;
            sq += d * d;
          }
        }
      }
      rstd = rsqrtf(sq / (float)ncols + p.ln_eps);
    }

    bool stat2 = false;
#pragma unroll
    for (int e = 0; e < 3; e++) stat2 |= (p.emit[e].kind == EMIT_LN);
    float sum2 = 0.f;
    const long long row = (long long)s * p.T_alloc + t;
    const float* rv = p.rowvec ? p.rowvec + (long long)s * p.rowvec_ld : nullptr;

    for (int c = 0; c < nchunks; c++) {
      tmem_ld32(taddr + c * 32, raw);
      tmem_ld_wait();
      const int cbase = n0 + c * 32;
      const int nv = min(32, ncols - c * 32);
      const bool full = nv == 32;
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; i++) {
        const int col = cbase + i;
        float x = __uint_as_float(raw[i]);
        if (i < nv) {
          if (p.bias) x += __ldg(p.bias + col);
          if (p.ln) x = (x - mean) * rstd * __ldg(p.ln_g + col) + __ldg(p.ln_b + col);
          if (p.act) x = apply_act(x, p.act, p.act_f, p.act == ACT_SNAKE ? __ldg(p.act_a + col) : 0.f);
          if (rv) x += __ldg(rv + col);
          if (p.mask_pre_res && !valid) x = 0.f;
        } else {
          x = 0.f;
        }
        v[i] = x;
      }
      if (p.res) {
        const float* rp = p.res + row * p.res_ld + cbase;
        if (full) {
#pragma unroll
          for (int i = 0; i < 8; i++) {
            const float4 f = __ldg(reinterpret_cast<const float4*>(rp) + i);
            v[i * 4 + 0] += f.x; v[i * 4 + 1] += f.y; v[i * 4 + 2] += f.z; v[i * 4 + 3] += f.w;
          }
        } else {
          for (int i = 0; i < nv; i++) v[i] += __ldg(rp + i);
        }
      }
      if (p.res2) {
        const float* rp = p.res2 + row * p.res2_ld + cbase;
        if (full) {
#pragma unroll
          for (int i = 0; i < 8; i++) {
            const float4 f = __ldg(reinterpret_cast<const float4*>(rp) + i);
            v[i * 4 + 0] += f.x; v[i * 4 + 1] += f.y; v[i * 4 + 2] += f.z; v[i * 4 + 3] += f.w;
          }
        } else {
          for (int i = 0; i < nv; i++) v[i] += __ldg(rp + i);
        }
      }
      if (p.out_scale != 1.f) {
#pragma unroll
        for (int i = 0; i < 32; i++) v[i] *= p.out_scale;
      }
      if (p.out32) {
        if (p.flat) {
          // transposed conv: row t holds stride*Cout consecutive output elements of the sequence slab
          const long long e0 = (long long)t * p.out32_ld + cbase + p.flat_off;
          const long long hi = (long long)len * p.flat_hi_per_len + p.flat_hi_add;
          if (e0 >= p.flat_lo && e0 + nv <= hi) {
            float* op = p.out32 + (long long)s * p.flat_seq_elems + e0;
            if (full) {
#pragma unroll
              for (int i = 0; i < 8; i++)
                reinterpret_cast<float4*>(op)[i] = make_float4(v[i * 4], v[i * 4 + 1], v[i * 4 + 2], v[i * 4 + 3]);
            } else {
              for (int i = 0; i < nv; i++) op[i] = v[i];
            }
          }
        } else {
          float* op = p.out32 + row * p.out32_ld + cbase;
          if (p.out32_accum) {
            if (full) {
#pragma unroll
              for (int i = 0; i < 8; i++) {
                const float4 f = reinterpret_cast<const float4*>(op)[i];
                v[i * 4 + 0] += f.x; v[i * 4 + 1] += f.y; v[i * 4 + 2] += f.z; v[i * 4 + 3] += f.w;
              }
            } else {
              for (int i = 0; i < nv; i++) v[i] += op[i];
            }
          }
          if (full) {
#pragma unroll
            for (int i = 0; i < 8; i++)
              reinterpret_cast<float4*>(op)[i] = make_float4(v[i * 4], v[i * 4 + 1], v[i * 4 + 2], v[i * 4 + 3]);
          } else {
            for (int i = 0; i < nv; i++) op[i] = v[i];
          }
        }
      }
      // 16-bit emits (the next contraction's A operand); padded rows are written as zeros
#pragma unroll
      for (int e = 0; e < 3; e++) {
        const Emit& em = p.emit[e];
        if (em.kind == EMIT_NONE || em.kind == EMIT_LN) continue;
        float w[32];
#pragma unroll
        for (int i = 0; i < 32; i++) {
          float x = v[i];
          if (em.kind == EMIT_SNAKE) x = (i < nv) ? snake_f(x, __ldg(em.a + cbase + i)) : 0.f;
          else if (em.kind == EMIT_LRELU) x = x > 0.f ? x : x * em.f;
          w[i] = valid ? x * em.scale : 0.f;
        }
        store_h32(em.ptr + row * em.ld + em.col_off + cbase, w, full, nv);
      }
      if (p.q) {  // attention operand split
        const int hd = p.heads * 64;
        const int which = cbase / hd;            // 0 q, 1 k, 2 v (a 32-col chunk never straddles)
        const int cc = cbase - which * hd;
        const int h = cc >> 6, d0 = cc & 63;
        if (which < 2) {
          float w[32];
          const float sc = which == 0 ? p.q_scale : 1.f;
#pragma unroll
          for (int i = 0; i < 32; i++) w[i] = valid ? v[i] * sc : 0.f;
          __half* dst = (which == 0 ? p.q : p.k) + (((long long)s * p.heads + h) * p.T_alloc + t) * 64 + d0;
          store_h32(dst, w, true, 32);
        } else {
          __half* dst = p.vt + (((long long)s * p.heads + h) * 64 + d0) * p.T_alloc + t;
#pragma unroll
          for (int i = 0; i < 32; i++) dst[(long long)i * p.T_alloc] = __float2half_rn(valid ? v[i] : 0.f);
        }
      }
      if (stat2) {
#pragma unroll
        for (int i = 0; i < 32; i++) {
          sum2 += v[i];
          raw[i] = __float_as_uint(v[i]);
        }
        tmem_st32(taddr + c * 32, raw);
      }
    }
    if (stat2) {  // LayerNorm of the final row value (pre-norm of the next sub-block), emitted as 16-bit
      tmem_st_wait();
      const float mean2 = sum2 / (float)ncols;
      float sq2 = 0.f;
      for (int c = 0; c < nchunks; c++) {
        tmem_ld32(taddr + c * 32, raw);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i++) {
          if (c * 32 + i < ncols) {
            const float d = __uint_as_float(raw[i]) - mean2;
            sq2 += d * d;
          }
        }
      }
      for (int c = 0; c < nchunks; c++) {
        tmem_ld32(taddr + c * 32, raw);
        tmem_ld_wait();
        const int cbase = n0 + c * 32;
        const int nv = min(32, ncols - c * 32);
#pragma unroll
        for (int e = 0; e < 3; e++) {
          const Emit& em = p.emit[e];
          if (em.kind != EMIT_LN) continue;
          const float rstd2 = rsqrtf(sq2 / (float)ncols + em.f);
          float w[32];
#pragma unroll
          for (int i = 0; i < 32; i++) {
            float x = 0.f;
            if (i < nv && valid)
              x = ((__uint_as_float(raw[i]) - mean2) * rstd2 * __ldg(em.a + cbase + i) + __ldg(em.b + cbase + i)) * em.scale;
            w[i] = x;
          }
          store_h32(em.ptr + row * em.ld + em.col_off + cbase, w, nv == 32, nv);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<(BN < 32 ? 32 : BN)>(tmem_base);
}

template <int BN>
static void launch_bn(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    CV2_CUDA(cudaFuncSetAttribute(gemm_tap_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmSmem<BN>::kTotal));
    configured = true;
  }
  dim3 grid((p.N + BN - 1) / BN, p.T_alloc / kTileM, p.S);
  gemm_tap_kernel<BN><<<grid, 192, GemmSmem<BN>::kTotal, stream>>>(tmA, tmB, p);
  CV2_LAUNCH_CHECK();
}

void launch_gemm_tap(int bn, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream) {
  CV2_CHECK(p.T_alloc % kTileM == 0, "gemm_tap: T_alloc %d not a multiple of 128", p.T_alloc);
  CV2_CHECK(p.ntaps >= 1 && p.ntaps <= 16 && p.kb_per_tap >= 1, "gemm_tap: bad taps %d / kb %d", p.ntaps, p.kb_per_tap);
  bool wants_row = p.ln;
  for (int e = 0; e < 3; e++) wants_row |= (p.emit[e].kind == EMIT_LN);
  CV2_CHECK(!wants_row || p.N <= bn, "gemm_tap: LayerNorm epilogue needs the full row in one tile (N=%d, BN=%d)", p.N, bn);
  CV2_CHECK(!p.q || (p.N == 3 * p.heads * 64), "gemm_tap: qkv split needs N == 3*heads*64");
  switch (bn) {
    case 64: launch_bn<64>(tmA, tmB, p, stream); break;
    case 128: launch_bn<128>(tmA, tmB, p, stream); break;
    case 256: launch_bn<256>(tmA, tmB, p, stream); break;
    default: fail("gemm_tap: unsupported BN %d", bn);
  }
}

}  // namespace cv2
