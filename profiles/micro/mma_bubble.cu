// Does per-k-block bookkeeping in the issuing thread (tcgen05.commit to a stage barrier, mbarrier try_wait on the next stage,
// descriptor arithmetic) open bubbles in the tensor pipe?  SS N=128 MMAs in groups of G with optional commit / wait between groups.
#include <cstdio>
#include "common.cuh"
using namespace cv2;

template <int N, int G, bool UNI>
__global__ void __launch_bounds__(128, 1) k(long long* out, int groups, int do_commit, int do_wait, int do_fence) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, sbar[8], fbar[8];
  __shared__ uint32_t slot;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    for (int i = 0; i < 8; i++) { mbar_init(&sbar[i], 1); mbar_init(&fbar[i], 1); }
    fence_barrier_init();
  }
  if (threadIdx.x < 32) tmem_alloc<512>(&slot);
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0)
    for (int i = 0; i < 8; i++) mbar_arrive(&fbar[i]);     // phase 0 of every "full" barrier is complete
  __syncthreads();
  if (UNI && threadIdx.x < 32) {   // warp-uniform control flow: descriptors live in uniform registers, one elected lane issues
    constexpr uint32_t idesc = umma_idesc_f16(128, N, 0);
    const long long t0 = clock64();
    for (int r = 0; r < groups; r++) {
      const int st = r & 7;
      if (do_wait) mbar_wait(&fbar[st], 0);
      const uint64_t a_desc = umma_smem_desc_sw128(smem_u32(smem + (st & 3) * 4096));
      const uint64_t b_desc = umma_smem_desc_sw128(smem_u32(smem + 16384 + (st & 1) * 16384));
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < G; kk++) umma_f16(tm, a_desc + (uint64_t)((kk & 3) * 2), b_desc + (uint64_t)((kk & 3) * 2), idesc, 1);
        if (do_commit) umma_commit(&sbar[st]);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
  }
  if (!UNI && threadIdx.x == 0) {
    constexpr uint32_t idesc = umma_idesc_f16(128, N, 0);
    const long long t0 = clock64();
    for (int r = 0; r < groups; r++) {
      const int st = r & 7;
      if (do_wait) mbar_wait(&fbar[st], 0);                 // always already complete: pure polling cost
      if (do_fence) tc_fence_after();
      const uint64_t a_desc = umma_smem_desc_sw128(smem_u32(smem + (st & 3) * 4096));
      const uint64_t b_desc = umma_smem_desc_sw128(smem_u32(smem + 16384 + (st & 1) * 16384));
#pragma unroll
      for (int kk = 0; kk < G; kk++) umma_f16(tm, a_desc + (uint64_t)((kk & 3) * 2), b_desc + (uint64_t)((kk & 3) * 2), idesc, 1);
      if (do_commit) umma_commit(&sbar[st]);
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
template <int N, int G, bool UNI>
void run(long long* d, int c, int w, int f) {
  cudaFuncSetAttribute(k<N, G, UNI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  const int groups = 20000;
  for (int i = 0; i < 2; i++) k<N, G, UNI><<<148, 128, 65536>>>(d, groups, c, w, f);
  cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%s N=%3d group=%d commit=%d wait=%d fence=%d: %.1f cycles per MMA (nominal %d), %.0f per group [%s]\n", UNI ? "warp-uniform" : "lane0-only  ", N, G, c, w, f,
         (double)h / ((double)G * groups), N / 2, (double)h / groups, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  long long* d;
  cudaMalloc(&d, 8);
  run<128, 4, false>(d, 1, 1, 0);
  run<128, 4, true>(d, 1, 1, 0);
  run<128, 4, true>(d, 0, 0, 0);
  run<256, 4, true>(d, 1, 1, 0);
  run<64, 4, false>(d, 1, 1, 0);
  run<64, 4, true>(d, 1, 1, 0);
  run<128, 16, true>(d, 1, 1, 0);
  return 0;
}
