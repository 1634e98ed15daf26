// tcgen05.mma.cta_group::2 (one issuing thread drives the tensor cores of a CTA pair, M = 256): issue rate with descriptors
// that change every group of 4 MMAs (the realistic case; with fixed descriptors the 1-CTA form already reaches nominal rate,
// mma_rate.cu), next to the same loop in cta_group::1 form (mma_bubble.cu).  Cluster (2,1,1), one CTA per SM.
#include <cstdio>
#include "common.cuh"
using namespace cv2;

__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}

template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k2(long long* out, int groups) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  const uint32_t tm = slot;
  if (rank == 0 && threadIdx.x == 0) {
    constexpr uint32_t idesc = umma_idesc_f16(256, N, 0);
    const long long t0 = clock64();
    for (int r = 0; r < groups; r++) {
      const int st = r & 7;
      const uint64_t a_desc = umma_smem_desc_sw128(smem_u32(smem + (st & 3) * 4096));
      const uint64_t b_desc = umma_smem_desc_sw128(smem_u32(smem + 16384 + (st & 1) * 16384));
#pragma unroll
      for (int kk = 0; kk < 4; kk++) umma2_f16(tm, a_desc + (uint64_t)(kk * 2), b_desc + (uint64_t)(kk * 2), idesc, 1);
    }
    umma2_commit_mc(&bar, 3);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  } else if (rank == 1 && threadIdx.x == 0) {
    mbar_wait(&bar, 0);
  }
  tc_fence_before();
  cluster_sync();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}
template <int N>
void run(long long* d) {
  cudaFuncSetAttribute(k2<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  const int groups = 20000;
  for (int i = 0; i < 2; i++) k2<N><<<148, 128, 65536>>>(d, groups);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("cta_group::2 M=256 N=%3d group=4 (descriptors change per group): %.1f cycles per MMA (nominal %d), %.0f FLOP/clk/SM [%s]\n", N,
         (double)h / (4.0 * groups), N / 2, 2.0 * 128 * N * 16 * 4.0 * groups / (double)h, cudaGetErrorString(e));
}
int main() {
  long long* d;
  cudaMalloc(&d, 8);
  run<128>(d);
  run<256>(d);
  run<64>(d);
  return 0;
}
