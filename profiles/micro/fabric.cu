// SM <-> L2 fabric micro-benchmark: do the per-SM load path (L2 -> SM) and store path (SM -> L2) share bandwidth?
// 148 CTAs x 16 warps.  Loads: coalesced ld.global.v8.b32 (L1 bypassed with .cg) over a buffer that stays in L2 (48 MB);
// stores: coalesced st.global.v8.b32 over another L2-sized buffer.  Modes: loads only, stores only, 8 warps each at once.
// Prints bytes per clock per SM for each direction.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void stg256(void* p, uint32_t v) {
  asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ldg256(const void* p) {
  uint32_t r[8];
  asm volatile("ld.global.cg.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p) : "memory");
  return r[0] ^ r[1] ^ r[2] ^ r[3] ^ r[4] ^ r[5] ^ r[6] ^ r[7];
}

// mode 0: all 16 warps load; 1: all store; 2: warps 0-7 load, 8-15 store
__global__ void __launch_bounds__(512, 1) fabric_kernel(const uint8_t* src, uint8_t* dst, int iters, int mode, size_t span, long long* cyc,
                                                       uint32_t* sink) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool loader = mode == 0 || (mode == 2 && warp < 8);
  uint32_t acc = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    const size_t off = (((size_t)it * gridDim.x + blockIdx.x) * 16 + warp) * 4096 % span;
    if (loader) {
#pragma unroll
      for (int i = 0; i < 4; i++) acc ^= ldg256(src + off + i * 1024 + lane * 32);
    } else {
#pragma unroll
      for (int i = 0; i < 4; i++) stg256(dst + off + i * 1024 + lane * 32, it);
    }
  }
  if (acc == 0x12345678u) sink[0] = acc;
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = clock64() - t0;
}

int main() {
  const size_t span = (size_t)48 << 20;      // stays in the 126 MB L2 together with the store buffer
  uint8_t *src, *dst; long long* cyc; uint32_t* sink;
  cudaMalloc(&src, span); cudaMalloc(&dst, span); cudaMalloc(&cyc, 8); cudaMalloc(&sink, 4);
  cudaMemset(src, 1, span); cudaMemset(dst, 0, span);
  const int iters = 400;
  const char* names[3] = {"loads only (16 warps)", "stores only (16 warps)", "8 warps load + 8 warps store"};
  for (int mode = 0; mode < 3; mode++) {
    fabric_kernel<<<148, 512>>>(src, dst, iters, mode, span, cyc, sink);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    fabric_kernel<<<148, 512>>>(src, dst, iters, mode, span, cyc, sink);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_warp = 4096.0 * iters;
    const double ld = (mode == 0 ? 16 : mode == 2 ? 8 : 0) * per_warp, st = (mode == 1 ? 16 : mode == 2 ? 8 : 0) * per_warp;
    printf("%-32s %7.3f ms  in %6.1f B/clk/SM (%5.2f TB/s)  out %6.1f B/clk/SM (%5.2f TB/s)  err=%s\n", names[mode], ms, ld / c,
           148 * ld / ms / 1e9, st / c, 148 * st / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
