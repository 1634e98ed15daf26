// How much does one SM pull through the TMA engine as a function of the bytes it keeps in flight?  148 CTAs, one thread each
// issues cp.async.bulk (global -> shared, mbarrier complete_tx) of 16 KB into a ring of D slots and re-issues a slot as soon as
// it has landed (no consumer work).  Source: a 96 MB buffer (L2 resident after the warm-up pass), every CTA on its own stream of
// 16 KB blocks.  Prints B/clk/SM and the implied latency = bytes in flight / rate for D = 1..12 (16..192 KB in flight).
#include <cstdio>
#include <cstdint>
#include "common.cuh"
using namespace cv2;

__global__ void __launch_bounds__(64, 1) depth_kernel(const uint8_t* src, size_t span, int depth, int iters, long long* cyc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[16];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; i++) mbar_init(&full[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    constexpr uint32_t kBytes = 16384;
    auto issue = [&](int it) {
      const int slot = it % depth;
      const size_t off = (((size_t)it * gridDim.x + blockIdx.x) * kBytes) % span;
      mbar_expect_tx(&full[slot], kBytes);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem + slot * kBytes)),
                   "l"(src + off), "r"(kBytes), "r"(smem_u32(&full[slot]))
                   : "memory");
    };
    const long long t0 = clock64();
    for (int it = 0; it < depth && it < iters; it++) issue(it);
    for (int it = depth; it < iters + depth; it++) {
      const int slot = it % depth;
      mbar_wait(&full[slot], ((it / depth) - 1) & 1);
      if (it < iters) issue(it);
    }
    if (blockIdx.x == 0) cyc[0] = clock64() - t0;
  }
}

int main() {
  uint8_t* src; long long* cyc;
  cudaMalloc(&src, (size_t)96 << 20); cudaMalloc(&cyc, 8);
  cudaMemset(src, 1, (size_t)96 << 20);
  cudaFuncSetAttribute(depth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 600;
  for (size_t mb : {24, 48, 96}) {
    const size_t span = mb << 20;
    for (int depth : {2, 4, 8, 12}) {
      depth_kernel<<<148, 64, 196608 + 1024>>>(src, span, depth, iters, cyc);
      cudaDeviceSynchronize();
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      depth_kernel<<<148, 64, 196608 + 1024>>>(src, span, depth, iters, cyc);
      cudaEventRecord(e1);
      cudaDeviceSynchronize();
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      const double rate = 16384.0 * iters / (double)c;
      printf("span %3zu MB, in flight %3d KB: %7.3f ms  %6.1f B/clk/SM  %5.2f TB/s aggregate  implied latency %6.0f cycles  err=%s\n", mb, depth * 16, ms,
             rate, 148.0 * 16384.0 * iters / ms / 1e9, depth * 16384.0 / rate, cudaGetErrorString(cudaGetLastError()));
    }
  }
  // per-SM or chip-wide?  the same at 128 KB in flight with fewer CTAs (one per SM)
  for (int grid : {18, 37, 74, 148}) {
    const size_t span = (size_t)24 << 20;
    depth_kernel<<<grid, 64, 196608 + 1024>>>(src, span, 8, iters, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    depth_kernel<<<grid, 64, 196608 + 1024>>>(src, span, 8, iters, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("grid %3d CTAs, 128 KB in flight: %7.3f ms  %6.1f B/clk/SM  %5.2f TB/s aggregate\n", grid, ms, 16384.0 * iters / (double)c,
           grid * 16384.0 * iters / ms / 1e9);
  }
  return 0;
}
