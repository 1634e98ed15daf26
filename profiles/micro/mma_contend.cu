// What slows tcgen05.mma inside the FFN kernel (112-230 cycles per instruction instead of 64-128)?  Same MMA chain as mma_rate.cu
// with optional concurrent traffic: (a) a second thread streaming weight-sized bulk copies global -> smem (cp.async.bulk) into a
// separate ring, (b) four warps doing tcgen05.ld / tcgen05.st on other TMEM columns.
#include <cstdio>
#include "common.cuh"
using namespace cv2;

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int N, bool TS>
__global__ void __launch_bounds__(192, 1) k(long long* out, int reps, const uint8_t* gsrc, int do_tma, int do_tmem) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, lbar[4];
  __shared__ uint32_t slot;
  __shared__ volatile int done;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 4; i++) mbar_init(&lbar[i], 1); fence_barrier_init(); done = 0; }
  if (threadIdx.x < 32) tmem_alloc<512>(&slot);
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 192) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = umma_idesc_f16(128, N, 0);
    const uint64_t a_desc = umma_smem_desc_sw128(smem_u32(smem));
    const uint64_t b_desc = umma_smem_desc_sw128(smem_u32(smem + 16384));
    const long long t0 = clock64();
    for (int r = 0; r < reps; r++) {
#pragma unroll
      for (int kk = 0; kk < 4; kk++) {
        if (TS) umma_f16_ts(tm, tm + 256 + kk * 8, b_desc + (uint64_t)(kk * 2), idesc, 1);
        else umma_f16(tm, a_desc + (uint64_t)(kk * 2), b_desc + (uint64_t)(kk * 2), idesc, 1);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    done = 1;
    if (blockIdx.x == 0) out[0] = t1 - t0;
  } else if (threadIdx.x == 32 && do_tma) {
    // stream 16 KB copies into a 4 x 16 KB ring at smem + 64 KB as fast as they complete
    uint8_t* ring = smem + 65536;
    int it = 0;
    long long bytes = 0;
    while (!done) {
      const int s = it & 3;
      if (it >= 4) mbar_wait(&lbar[s], ((it >> 2) - 1) & 1);
      mbar_expect_tx(&lbar[s], 16384);
      bulk_g2s(ring + s * 16384, gsrc + ((size_t)(blockIdx.x * 64 + (it & 63)) * 16384), 16384, &lbar[s]);
      it++;
      bytes += 16384;
    }
    for (int j = (it > 4 ? it - 4 : 0); j < it; j++) mbar_wait(&lbar[j & 3], (j >> 2) & 1);
    if (blockIdx.x == 0) out[1] = bytes;
  } else if (warp >= 2 && do_tmem) {
    const uint32_t la = tm + ((uint32_t)((warp & 3) * 32) << 16) + 384;
    uint32_t r[32];
    while (!done) {
      tmem_ld32(la, r);
      tmem_ld_wait();
      tmem_st32(la + 32, r);
      tmem_st_wait();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tm);
}

template <int N, bool TS>
void run(const char* name, long long* d, const uint8_t* g, int tma, int tmem) {
  cudaFuncSetAttribute(k<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 65536);
  const int reps = 20000;
  for (int w = 0; w < 2; w++) k<N, TS><<<148, 192, 131072>>>(d, reps, g, tma, tmem);
  cudaDeviceSynchronize();
  long long h[2] = {0, 0};
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%s N=%3d tma=%d tmem=%d: %.1f cycles per MMA (nominal %d)", name, N, tma, tmem, (double)h[0] / (4.0 * reps), N / 2);
  if (tma) printf(", TMA-in %.1f B/clk", (double)h[1] / (double)h[0]);
  printf("  [%s]\n", cudaGetErrorString(cudaGetLastError()));
}
int main() {
  long long* d;
  cudaMalloc(&d, 16);
  uint8_t* g;
  cudaMalloc(&g, (size_t)148 * 64 * 16384);
  cudaMemset(g, 0, (size_t)148 * 64 * 16384);
  for (int tma = 0; tma < 2; tma++)
    for (int tmem = 0; tmem < 2; tmem++) {
      run<128, false>("SS", d, g, tma, tmem);
      run<128, true>("TS", d, g, tma, tmem);
      run<256, true>("TS", d, g, tma, tmem);
      run<64, false>("SS", d, g, tma, tmem);
      run<64, true>("TS", d, g, tma, tmem);
    }
  return 0;
}
