// Pipe-rate micro-benchmarks behind the attention kernel's softmax: what one SM can do per clock with
//   (a) tcgen05.ld 32x32b.x32 (TMEM -> registers) from W warps,
//   (b) MUFU.EX2 alone, (c) MUFU.EX2 + the fp32 -> f16x2 pack that follows it, (d) the FMA-pipe cubic exp2 in packed two-lane form, (e) a 3 : 1 mix of (c) and (d).
// 148 CTAs (one per SM) x W warps; prints elements per clock per SM.  Build: make softmax_pipes; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "common.cuh"

using namespace cv2;

__global__ void __launch_bounds__(512, 1) tmem_rd_kernel(int iters, uint32_t* sink, long long* cyc) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    uint32_t r[32];
    tmem_ld32(base + ((it + warp) & 15) * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; i += 8) acc ^= r[i];
  }
  const long long t1 = clock64();
  if (acc == 0x12345678u) sink[0] = acc;
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  if (warp == 0) tmem_dealloc<512>(slot);
}

// two loads in flight per warp
__global__ void __launch_bounds__(512, 1) tmem_rd2_kernel(int iters, uint32_t* sink, long long* cyc) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; it += 2) {
    uint32_t r[32], q[32];
    tmem_ld32(base + ((it + warp) & 15) * 32, r);
    tmem_ld32(base + ((it + 1 + warp) & 15) * 32, q);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; i += 8) acc ^= r[i] ^ q[i];
  }
  const long long t1 = clock64();
  if (acc == 0x12345678u) sink[0] = acc;
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  if (warp == 0) tmem_dealloc<512>(slot);
}

__device__ __forceinline__ float2 poly_exp2x2(float2 x) {
  // Cody-Waite split with the round-down magic add, cubic minimax for 2^f on [0,1) (max relative error 1.03e-4), exponent spliced in
  // with one shift-add; no clamp (inputs in [-100, 8])
  float2 t, fl, f, p;
  const float2 magic = make_float2(12582912.f, 12582912.f);
  asm("add.rm.ftz.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<unsigned long long*>(&t)) : "l"(*reinterpret_cast<unsigned long long*>(&x)), "l"(*reinterpret_cast<const unsigned long long*>(&magic)));
  const float2 nmagic = make_float2(-12582912.f, -12582912.f);
  fl = fadd2(t, nmagic);
  f = ffma2(fl, make_float2(-1.f, -1.f), x);
  p = ffma2(f, make_float2(0.07826797f, 0.07826797f), make_float2(0.22630768f, 0.22630768f));
  p = ffma2(p, f, make_float2(0.69542435f, 0.69542435f));
  p = ffma2(p, f, make_float2(1.f, 1.f));
  float2 r;
  r.x = __uint_as_float(__float_as_uint(p.x) + (__float_as_uint(t.x) << 23));
  r.y = __uint_as_float(__float_as_uint(p.y) + (__float_as_uint(t.y) << 23));
  return r;
}

// MODE 0: ex2 only; 1: ex2 + pack; 2: poly + pack; 3: 3 MUFU pairs : 1 poly pair, + pack; 4: 1 : 1
template <int MODE>
__global__ void __launch_bounds__(512, 1) exp_kernel(int iters, uint32_t* sink, long long* cyc, float seed) {
  float2 x[16];
#pragma unroll
  for (int i = 0; i < 16; i++) x[i] = make_float2(-seed * (threadIdx.x + i), -seed * (threadIdx.x + 2 * i + 1));
  uint32_t acc = 0;
  float2 s = make_float2(0.f, 0.f);
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const float2 a = ffma2(x[i], make_float2(1.0001f, 1.0001f), make_float2(-0.001f * it, -0.001f * it));
      float2 e;
      const bool poly = MODE == 2 || (MODE == 3 && (i & 3) == 3) || (MODE == 4 && (i & 1));
      if (poly) e = poly_exp2x2(a);
      else e = make_float2(fast_exp2(a.x), fast_exp2(a.y));
      s = fadd2(s, e);
      if (MODE >= 1) {
        __half2 h = __floats2half2_rn(e.x, e.y);
        acc ^= *reinterpret_cast<uint32_t*>(&h);
      }
    }
  }
  const long long t1 = clock64();
  if (acc == 0x12345678u || s.x + s.y == 1.2345f) sink[0] = acc;
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  uint32_t* sink; long long* cyc;
  cudaMalloc(&sink, 64); cudaMalloc(&cyc, 8);
  const int iters = 4000;
  for (int warps : {4, 8, 16}) {
    for (int two = 0; two < 2; two++) {
      for (int rep = 0; rep < 2; rep++) {
        if (two) tmem_rd2_kernel<<<148, warps * 32>>>(iters, sink, cyc);
        else tmem_rd_kernel<<<148, warps * 32>>>(iters, sink, cyc);
        cudaDeviceSynchronize();
      }
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      printf("tcgen05.ld x32, %2d warps, %d in flight: %7.1f B/clk/SM (%lld cycles)  err=%s\n", warps, two + 1, (double)warps * 4096 * iters / c, c,
             cudaGetErrorString(cudaGetLastError()));
    }
  }
  const char* names[5] = {"ex2", "ex2 + pack", "poly + pack", "3 ex2 : 1 poly + pack", "1 ex2 : 1 poly + pack"};
  for (int warps : {4, 8, 16}) {
    for (int mode = 0; mode < 5; mode++) {
      for (int rep = 0; rep < 2; rep++) {
        switch (mode) {
          case 0: exp_kernel<0><<<148, warps * 32>>>(iters, sink, cyc, 0.01f); break;
          case 1: exp_kernel<1><<<148, warps * 32>>>(iters, sink, cyc, 0.01f); break;
          case 2: exp_kernel<2><<<148, warps * 32>>>(iters, sink, cyc, 0.01f); break;
          case 3: exp_kernel<3><<<148, warps * 32>>>(iters, sink, cyc, 0.01f); break;
          default: exp_kernel<4><<<148, warps * 32>>>(iters, sink, cyc, 0.01f); break;
        }
        cudaDeviceSynchronize();
      }
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      printf("%-24s %2d warps: %6.2f elements/clk/SM (%lld cycles)  err=%s\n", names[mode], warps, (double)warps * 32 * 32 * iters / c, c,
             cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
