// What does the MMA-issuing thread of the attention kernel cost per key tile, and do the issuing threads of several CTAs on one SM
// serialise?  Every CTA runs the attention MMA schedule without any softmax: per "tile" one group of GA SS MMAs with N = NA
// (S = Q K^T), tcgen05.commit, one group of GB TS MMAs with N = 64 (O += P V), tcgen05.commit -- optionally with an (always
// complete) mbarrier wait + tcgen05 fence in front of each group, as the real loop has.  CPS CTAs per SM.
// Prints cycles per tile per CTA, and the tensor-pipe nominal (M=128: N/2 cycles per K=16 instruction; N = 64 measured 45-48).
#include <cstdio>
#include "common.cuh"
using namespace cv2;

template <int NA, int GA, int GB, int CPS, int COLS, bool UNI>
__global__ void __launch_bounds__(64, CPS) k(long long* out, int tiles, int do_commit, int do_wait) {
  extern __shared__ __align__(1024) uint8_t smem[];   // Q 16 KB | K / V 32 KB
  __shared__ uint64_t bar, sbar[4], fbar[4];
  __shared__ uint32_t slot;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    for (int i = 0; i < 4; i++) { mbar_init(&sbar[i], 1); mbar_init(&fbar[i], 1); }
    fence_barrier_init();
  }
  if (threadIdx.x < 32) tmem_alloc<COLS>(&slot);
  for (int i = threadIdx.x; i < 49152 / 4; i += 64) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0)
    for (int i = 0; i < 4; i++) mbar_arrive(&fbar[i]);
  __syncthreads();
  constexpr uint32_t idesc_a = umma_idesc_f16(128, NA, 0);
  constexpr uint32_t idesc_b = umma_idesc_f16(128, 64, 0);
  constexpr uint32_t colO = COLS - 64;       // O accumulator at the top; S / P at 0
  if (threadIdx.x < 32 && (UNI || threadIdx.x == 0)) {
    const bool issuer = UNI ? elect_one() : true;
    const long long t0 = clock64();
    for (int r = 0; r < tiles; r++) {
      const int st = r & 1;
      if (do_wait) { mbar_wait(&fbar[st], 0); tc_fence_after(); }
      const uint64_t q_desc = umma_smem_desc_sw128(smem_u32(smem));
      const uint64_t k_desc = umma_smem_desc_sw128(smem_u32(smem + 16384 + st * 8192));
      if (issuer) {
#pragma unroll
        for (int kk = 0; kk < GA; kk++) umma_f16(tm, q_desc + (uint64_t)((kk & 3) * 2), k_desc + (uint64_t)((kk & 3) * 2), idesc_a, kk != 0);
        if (do_commit) umma_commit(&sbar[st]);
      }
      if (UNI) __syncwarp();
      if (do_wait) { mbar_wait(&fbar[2 + st], 0); tc_fence_after(); }
      const uint64_t v_desc = umma_smem_desc_sw128(smem_u32(smem + 32768 + st * 8192));
      if (issuer) {
#pragma unroll
        for (int kk = 0; kk < GB; kk++) umma_f16_ts(tm + colO, tm + (kk & 7) * 8, v_desc + (uint64_t)((kk & 3) * 2), idesc_b, 1);
        if (do_commit) umma_commit(&sbar[2 + st]);
      }
      if (UNI) __syncwarp();
    }
    if (issuer) umma_commit(&bar);
    if (UNI) __syncwarp();
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<COLS>(tm);
}

template <int NA, int GA, int GB, int CPS, int COLS, bool UNI>
void run(long long* d, int c, int w) {
  cudaFuncSetAttribute(k<NA, GA, GB, CPS, COLS, UNI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152);
  const int tiles = 20000;
  for (int i = 0; i < 2; i++) k<NA, GA, GB, CPS, COLS, UNI><<<148 * CPS, 64, 49152>>>(d, tiles, c, w);
  cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  const double nominal = GA * NA / 2.0 + GB * 32.0;
  printf("%s S: %2d x N=%3d, PV: %2d x N=64, %d CTA/SM, commit=%d wait=%d: %7.1f cycles per tile per CTA (pipe nominal %5.0f; x CTAs %5.0f) -> %6.1f per (128 x 64) unit per SM [%s]\n",
         UNI ? "warp-uniform" : "lane0-only  ", GA, NA, GB, CPS, c, w, (double)h / tiles, nominal, nominal * CPS,
         (double)h / tiles / (NA / 64.0) / CPS * 1.0, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  // the v9 shape: 64-key tiles, 4 + 4 MMAs
  run<64, 4, 4, 1, 128, false>(d, 1, 1);
  run<64, 4, 4, 1, 128, false>(d, 1, 0);
  run<64, 4, 4, 1, 128, false>(d, 0, 0);
  run<64, 4, 4, 1, 128, true>(d, 1, 1);
  run<64, 4, 4, 1, 128, true>(d, 0, 0);
  run<64, 4, 4, 2, 128, false>(d, 1, 1);
  run<64, 4, 4, 4, 128, false>(d, 1, 1);
  run<64, 4, 4, 4, 128, true>(d, 1, 1);
  // 128-key tiles
  run<128, 4, 8, 1, 256, false>(d, 1, 1);
  run<128, 4, 8, 1, 256, true>(d, 1, 1);
  run<128, 4, 8, 2, 256, false>(d, 1, 1);
  run<128, 4, 8, 2, 256, true>(d, 1, 1);
  // 256-key tiles
  run<256, 4, 16, 1, 512, false>(d, 1, 1);
  run<256, 4, 16, 1, 512, true>(d, 1, 1);
  return 0;
}
