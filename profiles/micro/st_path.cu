// Store-path micro-benchmark: how fast can one SM drain epilogue output?  148 CTAs x 16 warps, every warp writes 4 KB per "tile"
// (32 rows x 128 B, contiguous) for `iters` tiles, either with 4 fully coalesced st.global.v8.b32 per lane (LSU path) or by
// staging the 4 KB in shared memory and handing them to one cp.async.bulk.global.shared::cta (TMA path), or the row-per-lane
// pattern the epilogues used before the lane-group transposes.  Prints bytes per clock per SM and aggregate TB/s.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void stg256(void* p, uint32_t v) {
  asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) st_kernel(uint8_t* out, int iters, long long* cyc) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* stage = smem + warp * 4096;
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    uint8_t* dst = out + ((size_t)(it * gridDim.x + blockIdx.x) * 16 + warp) * 4096;
    if (MODE == 0) {            // coalesced: instruction i covers bytes [1024 i, 1024 i + 1024) of the warp's 4 KB
#pragma unroll
      for (int i = 0; i < 4; i++) stg256(dst + i * 1024 + lane * 32, it);
    } else if (MODE == 1) {     // row per lane: lane owns 128 contiguous bytes
#pragma unroll
      for (int i = 0; i < 4; i++) stg256(dst + lane * 128 + i * 32, it);
    } else {                    // smem stage + one bulk copy
      if (lane == 0 && it > 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; i++)
        *reinterpret_cast<uint4*>(stage + i * 512 + lane * 16) = make_uint4(it, it, it, it);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 4096;" ::"l"(dst), "r"((uint32_t)__cvta_generic_to_shared(stage)) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
  }
  if (MODE == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = clock64() - t0;
}

template <int MODE>
void run(const char* name, uint8_t* buf, int iters, long long* cyc) {
  cudaFuncSetAttribute(st_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  st_kernel<MODE><<<148, 512, 65536>>>(buf, iters, cyc);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  st_kernel<MODE><<<148, 512, 65536>>>(buf, iters, cyc);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double bytes = 148.0 * 16 * 4096 * iters;
  printf("%-34s %8.3f ms  %6.2f TB/s  %6.1f B/clk/SM (CTA 0: %lld cycles)  err=%s\n", name, ms, bytes / ms / 1e9, 16.0 * 4096 * iters / (double)c, c,
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const int iters = 200;                      // 148 x 64 KB x 200 = 1.94 GB per launch (far beyond L2)
  uint8_t* buf; long long* cyc;
  cudaMalloc(&buf, (size_t)148 * 65536 * iters);
  cudaMalloc(&cyc, 8);
  run<0>("st.global.v8 coalesced", buf, iters, cyc);
  run<1>("st.global.v8 row per lane", buf, iters, cyc);
  run<2>("smem stage + cp.async.bulk 4 KB", buf, iters, cyc);
  const int small = 6;                        // 58 MB per launch: stays in L2 (write-back), shows the SM-side path alone
  run<0>("L2-resident: coalesced", buf, small, cyc);
  run<1>("L2-resident: row per lane", buf, small, cyc);
  run<2>("L2-resident: bulk", buf, small, cyc);
  return 0;
}
