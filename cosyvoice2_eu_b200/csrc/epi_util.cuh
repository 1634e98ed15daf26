// Epilogue helpers shared by the tensor-core kernels (gemm_tap.cu, ffn_fused.cu): fast fp32 activation math, 32-wide
// register-row loads/stores, and coalesced global I/O through a per-warp shared-memory transpose.
#pragma once
#include "common.cuh"

namespace cv2 {

static constexpr int kStgFloats = 32 * 36;   // per-warp transpose staging: 32 rows x (32 + 4 pad) floats

// ---- fast epilogue math (fp32) ----
__device__ __forceinline__ float fast_mish(float x) {
  // x * tanh(softplus(x)) = x * n / (n + 2), n = e^x (e^x + 2)
  const float e = __expf(fminf(x, 20.f));
  const float n = e * (e + 2.f);
  const float y = __fdividef(x * n, n + 2.f);
  return x > 20.f ? x : y;
}
__device__ __forceinline__ float fast_gelu_erf(float x) {
  // exact-erf GELU = x * Phi(x), Phi(-|x|) = 0.5 erfc(|x|/sqrt 2) = 2^(-g(z) - 1), g = log2(e) * (-ln erfc(z)) fitted by a
  // degree-4 polynomial on z in [0,4] (|gelu err| < 2.6e-5, an order below the fp16 rounding of the emitted value).
  // 10 instructions, one MUFU (ex2), no division: gelu = max(x,0) - |x| * 2^(-g-1).
  const float z = fminf(fabsf(x) * 0.70710678118654752440f, 4.0f);
  float g = fmaf(-0.0192909595f, z, 0.136979282f);
  g = fmaf(g, z, 0.923393071f);
  g = fmaf(g, z, 1.6273005f);
  const float e = fast_exp2(fmaf(-g, z, -1.f));
  return fmaf(-fabsf(x), e, fmaxf(x, 0.f));
}
// the same GELU on a packed pair: polynomial and scaling on FFMA2 / FMUL2 (half the FMA-pipe issue slots)
__device__ __forceinline__ float2 fast_gelu_erf2(float2 x) {
  const float2 xc = fmul2(x, make_float2(0.70710678118654752440f, 0.70710678118654752440f));
  const float2 z = make_float2(fminf(fabsf(xc.x), 4.0f), fminf(fabsf(xc.y), 4.0f));
  float2 h = ffma2(make_float2(0.0192909595f, 0.0192909595f), z, make_float2(-0.136979282f, -0.136979282f));   // h = -g
  h = ffma2(h, z, make_float2(-0.923393071f, -0.923393071f));
  h = ffma2(h, z, make_float2(-1.6273005f, -1.6273005f));
  const float2 t = ffma2(h, z, make_float2(-1.f, -1.f));
  return make_float2(fmaf(-fabsf(x.x), fast_exp2(t.x), fmaxf(x.x, 0.f)), fmaf(-fabsf(x.y), fast_exp2(t.y), fmaxf(x.y, 0.f)));
}
__device__ __forceinline__ float fast_silu(float x) { return __fdividef(x, 1.f + __expf(-x)); }
__device__ __forceinline__ float fast_snake(float x, float a) {
  const float s = __sinf(x * a);
  return fmaf(s * s, __fdividef(1.f, a + 1e-9f), x);
}

// 32 consecutive floats starting at p (p 16B aligned when full)
__device__ __forceinline__ void load32(const float* __restrict__ p, float* d, bool full, int nv) {
  if (full) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const float4 f = __ldg(reinterpret_cast<const float4*>(p) + i);
      d[i * 4 + 0] = f.x; d[i * 4 + 1] = f.y; d[i * 4 + 2] = f.z; d[i * 4 + 3] = f.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; i++) d[i] = (i < nv) ? __ldg(p + i) : 0.f;
  }
}
__device__ __forceinline__ void store32_f32(float* __restrict__ p, const float* v, bool full, int nv) {
  if (full) {
#pragma unroll
    for (int i = 0; i < 8; i++) reinterpret_cast<float4*>(p)[i] = make_float4(v[i * 4], v[i * 4 + 1], v[i * 4 + 2], v[i * 4 + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < 32; i++)
      if (i < nv) p[i] = v[i];
  }
}
__device__ __forceinline__ void store32_f16(__half* __restrict__ p, const float* v, bool full, int nv) {
  if (full) {
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
      reinterpret_cast<uint4*>(p)[i] = u;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; i++)
      if (i < nv) p[i] = __float2half_rn(v[i]);
  }
}

// ---- global I/O for the thread-per-row epilogue ----
// A warp owns 32 consecutive rows x 32 columns and each thread its own row.  Blackwell's 256-bit global accesses
// (ld/st.global.v8.b32 -> LDG.E.256 / STG.E.256) move one full 32-byte sector per lane per instruction, so the per-thread
// row segment (128 B of fp32 = 4 instructions, 64 B of 16-bit = 2) is sector-exact without any shared-memory transpose:
// no staging tile (72 KB in the GEMM kernel, spent on a 4th TMA stage instead), no syncwarp / LDS / STS in the epilogue.
// (The round-1 kernels used 128-bit pieces 32 rows apart -- half-used sectors -- and then a smem transpose to fix that.)
// Rows must be 32-byte aligned: ld a multiple of 8 floats / 16 halves, chunk offsets multiples of 32 elements.
__device__ __forceinline__ void ldg256(const void* p, uint32_t* r) {
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p)
               : "memory");
}
__device__ __forceinline__ void stg256(void* p, const uint32_t* r) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]),
               "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
// Epilogue tiles: the accumulator comes out of TMEM one row per lane (tcgen05.ld 32x32b), so a direct global access has every
// lane in a different 128-byte line (32 tag look-ups per instruction; the clock64 trace of the out-proj epilogue showed the
// LSU queue, not HBM, pacing it).  A 4x4 (2x2 for 16-bit) register transpose inside each group of 4 (2) lanes makes every
// instruction touch 8 (16) full rows of 128 (64) contiguous bytes instead.
template <int U, int M>   // U registers per unit, groups of M + 1 lanes... M = 3: 4x4 units, M = 1: 2x2 units
__device__ __forceinline__ void lane_group_transpose(uint32_t* a, int lane) {
#pragma unroll
  for (int m = 1; m <= (M + 1) / 2; m <<= 1) {
    const bool up = (lane & m) != 0;
#pragma unroll
    for (int u = 0; u <= M; u++) {
      if (u & m) continue;
#pragma unroll
      for (int j = 0; j < U; j++) {
        const uint32_t send = up ? a[u * U + j] : a[(u | m) * U + j];
        const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, m);
        if (up) a[u * U + j] = recv;
        else a[(u | m) * U + j] = recv;
      }
    }
  }
}
__device__ __forceinline__ void tile_load_f32(const float* __restrict__ g0, long long ld, float*, int lane, float* d) {
  // g0 -> element (first row of the warp, first column of the chunk); lane 4g+s fetches columns 8s..8s+7 of rows 4g..4g+3
  const float* p = g0 + (lane & ~3) * ld + (lane & 3) * 8;
  uint32_t* r = reinterpret_cast<uint32_t*>(d);
#pragma unroll
  for (int k = 0; k < 4; k++) ldg256(p + k * ld, r + k * 8);
  lane_group_transpose<8, 3>(r, lane);
}
__device__ __forceinline__ void tile_store_f32(float* __restrict__ g0, long long ld, float*, int lane, const float* v) {
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; i++) r[i] = __float_as_uint(v[i]);
  lane_group_transpose<8, 3>(r, lane);
  float* p = g0 + (lane & ~3) * ld + (lane & 3) * 8;
#pragma unroll
  for (int k = 0; k < 4; k++) stg256(p + k * ld, r + k * 8);
}
__device__ __forceinline__ void tile_store_f16(__half* __restrict__ g0, long long ld, float*, int lane, const float* v) {
  uint32_t pk[16];
#pragma unroll
  for (int i = 0; i < 16; i++) {
    __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    pk[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  lane_group_transpose<8, 1>(pk, lane);      // lane 2g+s: columns 16s..16s+15 of rows 2g, 2g+1
  __half* p = g0 + (lane & ~1) * ld + (lane & 1) * 16;
  stg256(p, pk);
  stg256(p + ld, pk + 8);
}
// (the former half-height staging variants of the FFN kernel are the same direct accesses now)
__device__ __forceinline__ void tile_load_f32_h16(const float* __restrict__ g0, long long ld, float* stg, int lane, float* d) {
  tile_load_f32(g0, ld, stg, lane, d);
}
__device__ __forceinline__ void tile_store_f32_h16(float* __restrict__ g0, long long ld, float* stg, int lane, const float* v) {
  tile_store_f32(g0, ld, stg, lane, v);
}
__device__ __forceinline__ void tile_store_f16_h16(__half* __restrict__ g0, long long ld, float* stg, int lane, const float* v) {
  tile_store_f16(g0, ld, stg, lane, v);
}

}  // namespace cv2
