// Issue-rate micro-benchmark of tcgen05.mma kind::f16 (M = 128, K = 16 per instruction) for N in {64, 128, 256}, with the A
// operand in shared memory (SS) or in tensor memory (TS).  One CTA per SM, one thread issues `reps` dependent-accumulate MMAs
// back to back, commit -> mbarrier, clock64 around it.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../cosyvoice2_eu_b200/csrc
#include <cstdio>
#include "common.cuh"
using namespace cv2;

template <int N, bool TS>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int reps, int ctas_active) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(&slot);
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0 && (int)blockIdx.x < ctas_active) {
    constexpr uint32_t idesc = umma_idesc_f16(128, N, 0);
    const uint64_t a_desc = umma_smem_desc_sw128(smem_u32(smem));
    const uint64_t b_desc = umma_smem_desc_sw128(smem_u32(smem + 16384));
    const long long t0 = clock64();
    for (int r = 0; r < reps; r++) {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        if (TS) umma_f16_ts(tm, tm + 256 + k * 8, b_desc + (uint64_t)(k * 2), idesc, 1);
        else umma_f16(tm, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, 1);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tm);
}

template <int N, bool TS>
void run(const char* name, long long* d, int reps) {
  cudaFuncSetAttribute(rate_kernel<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int warm = 0; warm < 2; warm++) rate_kernel<N, TS><<<148, 128, 65536>>>(d, reps, 148);
  cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%s N=%3d: %.1f cycles per MMA (nominal %d), %.0f FLOP/clk/SM\n", name, N, (double)h / (4.0 * reps), N / 2,
         2.0 * 128 * N * 16 * 4.0 * reps / (double)h);
}
int main() {
  long long* d;
  cudaMalloc(&d, 8);
  const int reps = 2000;
  run<64, false>("SS", d, reps);
  run<128, false>("SS", d, reps);
  run<256, false>("SS", d, reps);
  run<64, true>("TS", d, reps);
  run<128, true>("TS", d, reps);
  run<256, true>("TS", d, reps);
  // sustained: ~0.3 s launches back to back -- does the rate hold once the board reaches its power cap?
  cudaFuncSetAttribute(rate_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int it = 0; it < 12; it++) {
    rate_kernel<128, false><<<148, 128, 65536>>>(d, 1000000, 148);
    cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("sustained SS N=128 launch %d: %.1f cycles per MMA\n", it, (double)h / 4.0e6);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
