// HiFT vocoder kernels that are not tensor-core contractions (reference: cosyvoice/hifigan/generator.py,
// f0_predictor.py).  The F0 predictor stays fp32 (its output drives a chaotic phase integrator);
// NSF source, source STFT and the iSTFT head are bandwidth-bound fp32 kernels.
#include "hift_kernels.cuh"
#include "common.cuh"
#include "host_util.h"

namespace cv2 {

// ---------------------------------------------------------------------------------------------------------
// fp32 conv1d k=3 pad=1 + ELU, channels-last (ConvRNNF0Predictor.condnet, f0_predictor.py:31-52).
// fp32 on purpose (see engine.h: f0_split): 128 x 128 register-tiled SGEMM over K = 3*Cin, 8 x 8 outputs per thread as
// packed FFMA2 (one issue slot per two FMAs), BK = 16, global loads of step k+1 in flight while step k is computed
// (double-buffered shared tiles); rows beyond len read as 0 (tensor-edge semantics).
// w: [3][Cin][Cout] (Cout contiguous), weight-norm folded on the host.  Cin % 16 == 0, Cout % 128 == 0.
// ---------------------------------------------------------------------------------------------------------
static constexpr int kF0BM = 128, kF0BN = 128, kF0BK = 16, kF0LD = 132;
__global__ void __launch_bounds__(256, 1) conv3_elu_f32_kernel(const float* __restrict__ x, int Cin, const float* __restrict__ w,
                                                               const float* __restrict__ bias, float* __restrict__ y, int Cout,
                                                               const int* __restrict__ lens, int len_all, int T_alloc) {
  __shared__ __align__(16) float xs[2][kF0BK][kF0LD];   // [k][t]
  __shared__ __align__(16) float ws[2][kF0BK][kF0LD];   // [k][co]
  const int b = blockIdx.z, t0 = blockIdx.x * kF0BM, c0 = blockIdx.y * kF0BN;
  const int len = lens ? lens[b] : len_all;
  if (t0 >= len) return;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;   // 16 x 16 threads; rows {4ty..4ty+3, 64+4ty..}, cols {4tx..4tx+3, 64+4tx..}
  float2 acc[8][4];                          // [row][col pair]
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = make_float2(0.f, 0.f);
  const float* xb = x + (long long)b * T_alloc * Cin;
  const int ksteps = Cin / kF0BK, nsteps = 3 * ksteps;
  // this thread's slice of a step's global loads: x: rows xr, xr + 64, 4 consecutive k (float4); w: rows wk, wk + 8, float4 of co
  const int xr = tid >> 2, xk = (tid & 3) * 4;
  const int wk = tid >> 5, wc = (tid & 31) * 4;
  float4 gx[2], gw[2];
  auto gload = [&](int step) {
    const int tap = step / ksteps, k0 = (step - tap * ksteps) * kF0BK;
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int ts = t0 + xr + 64 * h + tap - 1;
      gx[h] = (ts >= 0 && ts < len) ? __ldg(reinterpret_cast<const float4*>(xb + (long long)ts * Cin + k0 + xk)) : make_float4(0.f, 0.f, 0.f, 0.f);
      gw[h] = __ldg(reinterpret_cast<const float4*>(w + ((long long)tap * Cin + k0 + wk + 8 * h) * Cout + c0 + wc));
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; h++) {
      xs[buf][xk + 0][xr + 64 * h] = gx[h].x;
      xs[buf][xk + 1][xr + 64 * h] = gx[h].y;
      xs[buf][xk + 2][xr + 64 * h] = gx[h].z;
      xs[buf][xk + 3][xr + 64 * h] = gx[h].w;
      *reinterpret_cast<float4*>(&ws[buf][wk + 8 * h][wc]) = gw[h];
    }
  };
  gload(0);
  sstore(0);
  __syncthreads();
  for (int step = 0; step < nsteps; step++) {
    const int buf = step & 1;
    if (step + 1 < nsteps) gload(step + 1);
#pragma unroll
    for (int kk = 0; kk < kF0BK; kk++) {
      const float4 a0 = *reinterpret_cast<const float4*>(&xs[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&xs[buf][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&ws[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&ws[buf][kk][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float2 bp[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const float2 aa = make_float2(av[i], av[i]);
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = ffma2(aa, bp[j], acc[i][j]);
      }
    }
    if (step + 1 < nsteps) {
      sstore(buf ^ 1);     // the other buffer was last read in step-1, which every thread left at the barrier below
      __syncthreads();
    }
  }
  float bv[8];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    bv[j] = bias[c0 + tx * 4 + j];
    bv[4 + j] = bias[c0 + 64 + tx * 4 + j];
  }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int t = t0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (t >= T_alloc) continue;
    const bool valid = t < len;
    float4 o0, o1;
    o0.x = valid ? elu_f(acc[i][0].x + bv[0]) : 0.f;
    o0.y = valid ? elu_f(acc[i][0].y + bv[1]) : 0.f;
    o0.z = valid ? elu_f(acc[i][1].x + bv[2]) : 0.f;
    o0.w = valid ? elu_f(acc[i][1].y + bv[3]) : 0.f;
    o1.x = valid ? elu_f(acc[i][2].x + bv[4]) : 0.f;
    o1.y = valid ? elu_f(acc[i][2].y + bv[5]) : 0.f;
    o1.z = valid ? elu_f(acc[i][3].x + bv[6]) : 0.f;
    o1.w = valid ? elu_f(acc[i][3].y + bv[7]) : 0.f;
    float* yr = y + ((long long)b * T_alloc + t) * Cout + c0;
    *reinterpret_cast<float4*>(yr + tx * 4) = o0;
    *reinterpret_cast<float4*>(yr + 64 + tx * 4) = o1;
  }
}
void launch_conv3_elu_f32(const float* x, int Cin, const float* w, const float* bias, float* y, int Cout, const int* lens,
                          int len_all, int B, int T_alloc, cudaStream_t st) {
  CV2_CHECK(Cin % kF0BK == 0 && Cout % kF0BN == 0, "conv3_elu_f32: Cin %d / Cout %d not tileable", Cin, Cout);
  dim3 grid((T_alloc + kF0BM - 1) / kF0BM, Cout / kF0BN, B);
  conv3_elu_f32_kernel<<<grid, 256, 0, st>>>(x, Cin, w, bias, y, Cout, lens, len_all, T_alloc);
  CV2_LAUNCH_CHECK();
}

// classifier Linear 512 -> 1 + abs (f0_predictor.py:58); warp per frame
__global__ void f0_head_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                               float* __restrict__ f0, const int* __restrict__ lens, int len_all, int T_alloc, int f0_stride) {
  const int bi = blockIdx.y;
  const int t = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int len = lens ? lens[bi] : len_all;
  if (t >= T_alloc) return;
  float acc = 0.f;
  if (t < len) {
    const float* xr = x + ((long long)bi * T_alloc + t) * 512;
    for (int i = lane; i < 512; i += 32) acc += xr[i] * w[i];
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  }
  if (lane == 0 && t < f0_stride) f0[(long long)bi * f0_stride + t] = (t < len) ? fabsf(acc + b[0]) : 0.f;
}
void launch_f0_head(const float* x, const float* w, const float* b, float* f0, const int* lens, int len_all, int B, int T_alloc,
                    int f0_stride, cudaStream_t st) {
  f0_head_kernel<<<dim3((T_alloc + 7) / 8, B), 256, 0, st>>>(x, w, b, f0, lens, len_all, T_alloc, f0_stride);
  CV2_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------------------
// NSF source (SineGen2 + SourceModuleHnNSF2, generator.py:261-283, 314-339, 375-389).
// Step 1 (frame rate): per harmonic, rad = fp32(f0*h / 24000) mod 1; inclusive scan with fp64 accumulation and
//   fp32 outputs (what ATen's CPU cumsum does); P = ((C*2)*pi)*480 with fp32 roundings in that order.
// Step 2 (sample rate): linear interpolation of P (ATen upsample_linear1d, align_corners=False, scale 1/480),
//   sin, voiced gate, additive Gaussian noise, Linear 9->1, tanh.
// ---------------------------------------------------------------------------------------------------------
__global__ void nsf_phase_kernel(const float* __restrict__ f0, int f0_stride, const int* __restrict__ lens, int len_all,
                                 float* __restrict__ P, int T_alloc) {
  const int b = blockIdx.x, h = threadIdx.x;  // 9 threads
  if (h >= 9) return;
  const int len = lens ? lens[b] : len_all;
  const float hm = (float)(h + 1);
  double c = 0.0;
  const float PI32 = 3.14159274101257324f;  // float(np.pi)
  for (int t = 0; t < len; t++) {
    const float fn = __fmul_rn(f0[(long long)b * f0_stride + t], hm);
    const float rad = fmodf(__fdiv_rn(fn, 24000.f), 1.f);
    c += (double)rad;
    const float cf = (float)c;
    const float ph = __fmul_rn(__fmul_rn(__fmul_rn(cf, 2.f), PI32), 480.f);
    P[((long long)b * T_alloc + t) * 9 + h] = ph;
  }
}

// counter-based N(0,1) pair: one 64-bit mix -> two uniforms -> both Box-Muller outputs (production mode only; parity mode
// injects the reference's noise tensor)
__device__ __forceinline__ void gauss_pair(unsigned long long seed, unsigned long long idx, float& n0, float& n1) {
  unsigned long long z = seed + idx * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  const float u1 = ((unsigned)(z >> 40) + 1.f) * (1.f / 16777217.f);
  const float u2 = (unsigned)((z >> 8) & 0xFFFFFF) * (1.f / 16777216.f);
  const float rr = sqrtf(-2.f * __logf(u1));
  float sn, cs;
  __sincosf(6.283185307f * u2, &sn, &cs);
  n0 = rr * cs;
  n1 = rr * sn;
}
// sin of a large fp32 phase (2 pi * 480 * cumulative cycles reaches 1e5..1e6): reduce modulo 2 pi in fp64 (exact to 1e-10),
// then the fp32 fast path; avoids sinf's Payne-Hanek slow path, which dominated this kernel
__device__ __forceinline__ float sin_big(float ph) {
  const double x = (double)ph;
  const double k = rint(x * 0.15915494309189535);
  const float r = (float)fma(-k, 6.283185307179586, x);
  return sinf(r);
}

__global__ void nsf_source_kernel(const float* __restrict__ f0, int f0_stride, const float* __restrict__ P, int T_alloc,
                                  const int* __restrict__ lens, int len_all, const float* __restrict__ noise,
                                  long long noise_bstride, unsigned long long seed, const unsigned long long* __restrict__ seed_ptr,
                                  const float* __restrict__ lw,
                                  const float* __restrict__ lb, const float* __restrict__ cache, int cache_len,
                                  long long cache_bstride, float* __restrict__ src, long long src_bstride) {
  const int b = blockIdx.y;
  const int len = lens ? lens[b] : len_all;
  const long long L = (long long)len * 480;
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= L) return;
  if (j < cache_len) {  // inference(): s[:, :, :cache_len] = cache_source  (generator.py:579-580)
    src[(long long)b * src_bstride + j] = cache[(long long)b * cache_bstride + j];
    return;
  }
  // ATen area_pixel_compute_source_index: scale*(dst+0.5)-0.5 clamped at 0, scale = float(1/480)
  const float scale = (float)(1.0 / 480.0);
  float sidx = __fsub_rn(__fmul_rn(scale, __fadd_rn((float)j, 0.5f)), 0.5f);
  sidx = fmaxf(sidx, 0.f);
  const int i0 = min((int)sidx, len - 1);
  const int i1 = min(i0 + 1, len - 1);
  const float l1 = __fsub_rn(sidx, (float)i0);
  const float l0 = __fsub_rn(1.f, l1);
  const int frame = (int)(j / 480);
  const float uv = f0[(long long)b * f0_stride + frame] > 10.f ? 1.f : 0.f;
  const float namp = uv * 0.003f + (1.f - uv) * 0.1f / 3.f;
  const float* p0 = P + ((long long)b * T_alloc + i0) * 9;
  const float* p1 = P + ((long long)b * T_alloc + i1) * 9;
  const float* nz = noise ? noise + (long long)b * noise_bstride + j * 9 : nullptr;
  if (seed_ptr) seed = *seed_ptr;   // device-resident seed: a captured CUDA graph still draws fresh noise per replay
  float nv[10];
  if (nz) {
#pragma unroll
    for (int h = 0; h < 9; h++) nv[h] = nz[h];
  } else {
#pragma unroll
    for (int h = 0; h < 10; h += 2)
      gauss_pair(seed + (unsigned long long)b * 0x1000003ull, (unsigned long long)j * 5 + (h >> 1), nv[h], nv[h + 1]);
  }
  float acc = lb[0];
#pragma unroll
  for (int h = 0; h < 9; h++) {
    const float ph = __fadd_rn(__fmul_rn(l0, p0[h]), __fmul_rn(l1, p1[h]));
    const float sine = sin_big(ph) * 0.1f;
    acc += lw[h] * (sine * uv + namp * nv[h]);
  }
  src[(long long)b * src_bstride + j] = tanhf(acc);
}
void launch_nsf_source(const float* f0, int f0_stride, float* P, int T_alloc, const int* lens, int len_all, const float* noise,
                       long long noise_bstride, unsigned long long seed, const unsigned long long* seed_ptr, const float* lw,
                       const float* lb, const float* cache,
                       int cache_len, long long cache_bstride, float* src, long long src_bstride, int B, int max_len,
                       cudaStream_t st) {
  nsf_phase_kernel<<<B, 32, 0, st>>>(f0, f0_stride, lens, len_all, P, T_alloc);
  CV2_LAUNCH_CHECK();
  const long long Lmax = (long long)max_len * 480;
  nsf_source_kernel<<<dim3((unsigned)((Lmax + 255) / 256), B), 256, 0, st>>>(f0, f0_stride, P, T_alloc, lens, len_all, noise,
                                                                             noise_bstride, seed, seed_ptr, lw, lb, cache, cache_len,
                                                                             cache_bstride, src, src_bstride);
  CV2_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------------------
// Source STFT (generator.py:504-510): n_fft 16, hop 4, periodic Hann, center=True (reflect pad 8).
// Output channels-last fp32 [B, F_alloc, 18] = [Re X0..8 | Im X0..8]; frames >= L/4+1 are zero.
// ---------------------------------------------------------------------------------------------------------
__constant__ float c_cos16[16];
__constant__ float c_sin16[16];
__constant__ float c_hann16[16];
__constant__ float c_inv_env4[4];     // 1 / (sum of the squared window of the four frames covering a hop): interior hops of the iSTFT
static PerDeviceOnce g_tables_once;   // __constant__ memory is per device
static void upload_tables() {
  float c[16], s[16], w[16];
  for (int i = 0; i < 16; i++) {
    c[i] = (float)cos(2.0 * M_PI * i / 16.0);
    s[i] = (float)sin(2.0 * M_PI * i / 16.0);
    w[i] = (float)(0.5 * (1.0 - cos(2.0 * M_PI * i / 16.0)));
  }
  CV2_CUDA(cudaMemcpyToSymbol(c_cos16, c, sizeof(c)));
  CV2_CUDA(cudaMemcpyToSymbol(c_sin16, s, sizeof(s)));
  CV2_CUDA(cudaMemcpyToSymbol(c_hann16, w, sizeof(w)));
  float inv[4];
  for (int i = 0; i < 4; i++) {
    float e = 0.f;
    for (int kp0 = 12; kp0 >= 0; kp0 -= 4) e = fmaf(w[kp0 + i], w[kp0 + i], e);   // the kernel's own order (df = -1 .. 2)
    inv[i] = 1.f / e;
  }
  CV2_CUDA(cudaMemcpyToSymbol(c_inv_env4, inv, sizeof(inv)));
}
static void ensure_tables() { g_tables_once.run(upload_tables); }
void hift_init_tables() { ensure_tables(); }

// One block = 256 consecutive frames of one utterance: the 1036 source samples they cover are staged in shared memory
// with coalesced loads (reflect indexing only matters for the first / last block), every thread computes the 18 outputs of
// its frame into a shared [256][18] tile, and the tile -- one contiguous 18 KB span of the channels-last output -- is
// written back with coalesced 16-byte stores.  HBM traffic = the algorithmic 4 B read + 18 B written per sample.
__global__ void __launch_bounds__(256) source_stft_kernel(const float* __restrict__ src, long long src_bstride,
                                                          const int* __restrict__ lens, int len_all, float* __restrict__ out,
                                                          int F_alloc) {
  __shared__ __align__(16) float xs[4 * 256 + 16];
  __shared__ __align__(16) float os[256 * 18];
  const int b = blockIdx.y;
  const int f0 = blockIdx.x * 256;
  const int nf = min(256, F_alloc - f0);        // frames of this block that exist in the output tensor
  const int len = lens ? lens[b] : len_all;     // mel frames
  const int L = len * 480;
  const int F = L / 4 + 1;
  float* ob = out + ((long long)b * F_alloc + f0) * 18;
  if (f0 < F) {
    const float* sb = src + (long long)b * src_bstride;
    const int m0 = 4 * f0 - 8;
    for (int e = threadIdx.x; e < 4 * 256 + 12; e += 256) {
      int m = m0 + e;
      if (m < 0) m = -m;
      if (m >= L) m = 2 * (L - 1) - m;
      xs[e] = (m >= 0 && m < L) ? sb[m] : 0.f;
    }
  }
  __syncthreads();
  const int f = f0 + threadIdx.x;
  float* o = os + threadIdx.x * 18;
  if (f < F) {
    float x[16];
#pragma unroll
    for (int q4 = 0; q4 < 4; q4++) {   // 16-byte aligned, conflict-free shared loads
      const float4 v4 = *reinterpret_cast<const float4*>(xs + 4 * threadIdx.x + 4 * q4);
      x[4 * q4 + 0] = v4.x * c_hann16[4 * q4 + 0];
      x[4 * q4 + 1] = v4.y * c_hann16[4 * q4 + 1];
      x[4 * q4 + 2] = v4.z * c_hann16[4 * q4 + 2];
      x[4 * q4 + 3] = v4.w * c_hann16[4 * q4 + 3];
    }
    // real-input symmetry: with a[n] = x[n] + x[16-n], b[n] = x[n] - x[16-n] (n = 1..7),
    //   Re X[k] = x[0] + (-1)^k x[8] + sum_n a[n] cos(2 pi k n / 16),   Im X[k] = - sum_n b[n] sin(2 pi k n / 16)
    // 112 multiply-adds per frame instead of the direct form's 288
    float a[8], bb[8];
#pragma unroll
    for (int n = 1; n < 8; n++) {
      a[n] = x[n] + x[16 - n];
      bb[n] = x[n] - x[16 - n];
    }
#pragma unroll
    for (int k = 0; k < 9; k++) {
      float re = x[0] + ((k & 1) ? -x[8] : x[8]);
      float im = 0.f;
#pragma unroll
      for (int n = 1; n < 8; n++) {
        const int idx = (k * n) & 15;
        re = fmaf(a[n], c_cos16[idx], re);
        if (k > 0 && k < 8) im = fmaf(-bb[n], c_sin16[idx], im);
      }
      o[k] = re;
      o[9 + k] = im;
    }
  } else {
#pragma unroll
    for (int k = 0; k < 18; k++) o[k] = 0.f;
  }
  __syncthreads();
  const int nflt = nf * 18;
  if ((reinterpret_cast<uintptr_t>(ob) & 15) == 0 && (nflt & 3) == 0) {
    float4* o4 = reinterpret_cast<float4*>(ob);
    const float4* s4 = reinterpret_cast<const float4*>(os);
    for (int e = threadIdx.x; e < nflt / 4; e += 256) o4[e] = s4[e];
  } else {
    for (int e = threadIdx.x; e < nflt; e += 256) ob[e] = os[e];
  }
}
void launch_source_stft(const float* src, long long src_bstride, const int* lens, int len_all, float* out, int F_alloc, int B,
                        cudaStream_t st) {
  ensure_tables();
  source_stft_kernel<<<dim3((F_alloc + 255) / 256, B), 256, 0, st>>>(src, src_bstride, lens, len_all, out, F_alloc);
  CV2_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------------------
// source_downs[i]: Conv1d(18 -> C, k, stride r, pad r/2) on the source spectrum (generator.py:455-466, 533).
// fp32 SIMT (K = 18*k is tiny).  w: [k][18][C] (C contiguous).  Emits the fp32 result and snake(alpha)(result)
// in 16-bit as the A operand of the source ResBlock's first conv.
// ---------------------------------------------------------------------------------------------------------
// Register-blocked: each thread owns 4 channels x 4 output frames (16 FMAs per weight float4 + 4 broadcast spectrum loads).
__global__ void __launch_bounds__(256) source_down_kernel(const float* __restrict__ stft, int F_alloc, const float* __restrict__ w,
                                   const float* __restrict__ bias, int k, int stride, int pad, int C, const int* __restrict__ lens,
                                   int len_all, int frames_per_len, int frames_add, float* __restrict__ out32,
                                   __half* __restrict__ out16, const float* __restrict__ alpha, int T_alloc) {
  const int b = blockIdx.y;
  const int c4 = threadIdx.x * 4;
  const int t0 = (blockIdx.x * blockDim.y + threadIdx.y) * 4;
  const int len = lens ? lens[b] : len_all;
  const int T_out = len * frames_per_len + frames_add;       // output frames of this stage
  const int F = len * 120 + 1;                               // stft frames
  if (t0 >= T_alloc || c4 >= C) return;
  float2 acc[4][2];   // packed channel pairs: one FFMA2 per two multiply-adds
  const float4 bv = *reinterpret_cast<const float4*>(bias + c4);
#pragma unroll
  for (int f = 0; f < 4; f++) { acc[f][0] = make_float2(bv.x, bv.y); acc[f][1] = make_float2(bv.z, bv.w); }
  if (t0 < T_out) {
    const float* sb = stft + (long long)b * F_alloc * 18;
    for (int j = 0; j < k; j++) {
      int fr[4];
      bool ok[4];
#pragma unroll
      for (int f = 0; f < 4; f++) {
        fr[f] = (t0 + f) * stride + j - pad;
        ok[f] = fr[f] >= 0 && fr[f] < F;
      }
      const float* wp = w + (long long)j * 18 * C + c4;
#pragma unroll 6
      for (int ch = 0; ch < 18; ch++) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(wp + (long long)ch * C));
#pragma unroll
        for (int f = 0; f < 4; f++) {
          const float sv = ok[f] ? __ldg(sb + (long long)fr[f] * 18 + ch) : 0.f;
          const float2 s2 = make_float2(sv, sv);
          acc[f][0] = ffma2(s2, make_float2(wv.x, wv.y), acc[f][0]);
          acc[f][1] = ffma2(s2, make_float2(wv.z, wv.w), acc[f][1]);
        }
      }
    }
  }
  const float4 av = *reinterpret_cast<const float4*>(alpha + c4);
#pragma unroll
  for (int f = 0; f < 4; f++) {
    const int t = t0 + f;
    if (t >= T_alloc) break;
    const bool valid = t < T_out;
    const long long o = ((long long)b * T_alloc + t) * C + c4;
    float4 v = valid ? make_float4(acc[f][0].x, acc[f][0].y, acc[f][1].x, acc[f][1].y) : make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(out32 + o) = v;
    __half2 h0 = __floats2half2_rn(valid ? snake_f(v.x, av.x) : 0.f, valid ? snake_f(v.y, av.y) : 0.f);
    __half2 h1 = __floats2half2_rn(valid ? snake_f(v.z, av.z) : 0.f, valid ? snake_f(v.w, av.w) : 0.f);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0);
    u.y = *reinterpret_cast<uint32_t*>(&h1);
    *reinterpret_cast<uint2*>(out16 + o) = u;
  }
}
void launch_source_down(const float* stft, int F_alloc, const float* w, const float* bias, int k, int stride, int pad, int C,
                        const int* lens, int len_all, int frames_per_len, int frames_add, float* out32, __half* out16,
                        const float* alpha, int B, int T_alloc, cudaStream_t st) {
  const int tx = C / 4;            // 64 / 32 / 16 threads along channels
  const int ty = 256 / tx;         // frames-of-4 per block
  const int frames_per_block = ty * 4;
  source_down_kernel<<<dim3((T_alloc + frames_per_block - 1) / frames_per_block, B), dim3(tx, ty), 0, st>>>(
      stft, F_alloc, w, bias, k, stride, pad, C, lens, len_all, frames_per_len, frames_add, out32, out16, alpha, T_alloc);
  CV2_LAUNCH_CHECK();
}

// reflection pad (1,0) of the last upsampler output (generator.py:529-530): row 0 <- row 2 (= x[1])
__global__ void reflect_row0_kernel(float* __restrict__ x, int T_alloc, int C) {
  const int b = blockIdx.x;
  float* base = x + (long long)b * T_alloc * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) base[c] = base[2 * C + c];
}
void launch_reflect_row0(float* x, int B, int T_alloc, int C, cudaStream_t st) {
  reflect_row0_kernel<<<B, 64, 0, st>>>(x, T_alloc, C);
  CV2_LAUNCH_CHECK();
}

// ---------------------------------------------------------------------------------------------------------
// iSTFT head (generator.py:546-551, 512-518): conv_post output [B, F_alloc, 18] ->
// mag = min(exp(c[0:9]), 100), phi = sin(c[9:18]), X = mag (cos phi + i sin phi), 16-point inverse real DFT,
// Hann-windowed overlap-add of the 4 frames covering each output sample, divide by the window envelope,
// trim 8 samples per side, clamp to +-0.99.  One pass: 18 B read + 4 B written per output sample... per hop.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) istft_kernel(const float* __restrict__ cp, int F_alloc, int ld, const int* __restrict__ lens,
                                                    int len_all, float* __restrict__ wav, long long wav_bstride,
                                                    short* __restrict__ pcm) {
  // Per frame ONCE: spectrum -> all 16 windowed samples of its inverse real DFT, using y[n] = A[n] - B[n], y[16-n] = A[n] + B[n]
  // (A = cosine part, B = sine part) so that the (A[n], B[n]) pairs of n = 1..7 are 49 packed FFMA2 on (re_k, im_k) x (2 cos, 2 sin);
  // the samples go through shared memory (rows padded to 80 B: conflict-free 16-byte accesses), then every thread overlap-adds the
  // four frames covering its hop.  (The previous form re-evaluated 7 complex terms for each of the 16 (frame, sample) pairs of a
  // hop: 504 instructions per thread, issue bound at 83 %.)
  __shared__ __align__(16) float ys[259][20];
  const int b = blockIdx.y;
  const int len = lens ? lens[b] : len_all;
  const int F = len * 120 + 1;
  const int q0 = blockIdx.x * 256;          // first hop of this block; hop q covers samples n = 4q..4q+3
  if (q0 >= F - 1) return;
  // frames q0-1 .. q0+257; a frame's 18 fp32 values are 72 contiguous bytes (8-byte aligned when ld is even).  (Fetching the block's
  // rows as one contiguous span into shared memory first was measured twice and is slower: 162 -> 270 us.)
  for (int e = threadIdx.x; e < 259; e += 256) {
    const int f = q0 - 1 + e;
    float4* out4 = reinterpret_cast<float4*>(&ys[e][0]);
    if (f < 0 || f >= F) {
#pragma unroll
      for (int i = 0; i < 4; i++) out4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      continue;
    }
    const float* c = cp + ((long long)b * F_alloc + f) * ld;
    float cv[18];
    if ((ld & 1) == 0 && (reinterpret_cast<uintptr_t>(cp) & 7) == 0) {
#pragma unroll
      for (int k = 0; k < 9; k++) {
        const float2 t2 = __ldg(reinterpret_cast<const float2*>(c) + k);
        cv[2 * k] = t2.x;
        cv[2 * k + 1] = t2.y;
      }
    } else {
#pragma unroll
      for (int k = 0; k < 18; k++) cv[k] = c[k];
    }
    float2 X[9];                             // (re, im)
#pragma unroll
    for (int k = 0; k < 9; k++) {
      const float mag = fminf(__expf(cv[k]), 100.f);
      const float ph = __sinf(cv[9 + k]);    // MUFU sine: abs error ~1e-6 on the conv_post phase logits (|x| of a few units),
                                             // six orders below the 35 dB parity budget; |ph| <= 1 for the sincos below
      // |ph| <= 1: sine and cosine as Taylor polynomials on the FMA pipe (packed pair; truncation error 2.8e-8 / 2.8e-7 at |ph| = 1)
      // instead of two more MUFU ops per bin -- the kernel was bound by its 36 transcendentals per 72-byte frame
      const float ph2 = ph * ph;
      float2 sc = ffma2(make_float2(ph2, ph2), make_float2(2.7557319e-6f, 2.4801587e-5f), make_float2(-1.9841270e-4f, -1.3888889e-3f));
      sc = ffma2(sc, make_float2(ph2, ph2), make_float2(8.3333333e-3f, 4.1666667e-2f));
      sc = ffma2(sc, make_float2(ph2, ph2), make_float2(-1.6666667e-1f, -0.5f));
      sc = ffma2(sc, make_float2(ph2, ph2), make_float2(1.f, 1.f));      // (sin(ph) / ph, cos(ph))
      X[k] = make_float2(mag * sc.y, mag * (sc.x * ph));
    }
    // n = 0 and n = 8: only cosines (+-1)
    const float ev = X[2].x + X[4].x + X[6].x, od = X[1].x + X[3].x + X[5].x + X[7].x;
    float y[16];
    y[0] = X[0].x + X[8].x + 2.f * (ev + od);
    y[8] = X[0].x + X[8].x + 2.f * (ev - od);
#pragma unroll
    for (int n = 1; n < 8; n++) {
      float2 acc = make_float2(X[0].x + ((n & 1) ? -X[8].x : X[8].x), 0.f);
#pragma unroll
      for (int k = 1; k < 8; k++) {
        const int idx = (k * n) & 15;
        acc = ffma2(X[k], make_float2(2.f * c_cos16[idx], 2.f * c_sin16[idx]), acc);
      }
      y[n] = acc.x - acc.y;
      y[16 - n] = acc.x + acc.y;
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
      out4[i] = make_float4(y[4 * i] * c_hann16[4 * i] * (1.f / 16.f), y[4 * i + 1] * c_hann16[4 * i + 1] * (1.f / 16.f),
                            y[4 * i + 2] * c_hann16[4 * i + 2] * (1.f / 16.f), y[4 * i + 3] * c_hann16[4 * i + 3] * (1.f / 16.f));
  }
  __syncthreads();
  const int q = q0 + threadIdx.x;
  if (q >= F - 1) return;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float env[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int df = -1; df <= 2; df++) {
    const int f = q + df;
    if (f < 0 || f >= F) continue;
    const int kp0 = 8 - 4 * df;             // position of the hop's first sample inside frame f: 12, 8, 4, 0
    const float4 v = *reinterpret_cast<const float4*>(&ys[threadIdx.x + 1 + df][kp0]);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
#pragma unroll
    for (int i = 0; i < 4; i++) env[i] = fmaf(c_hann16[kp0 + i], c_hann16[kp0 + i], env[i]);
  }
  float4 o;
  if (q >= 1 && q + 2 < F) {   // all four frames present: the envelope is the same four numbers for every interior hop (no divisions)
    o.x = fminf(fmaxf(acc.x * c_inv_env4[0], -0.99f), 0.99f);
    o.y = fminf(fmaxf(acc.y * c_inv_env4[1], -0.99f), 0.99f);
    o.z = fminf(fmaxf(acc.z * c_inv_env4[2], -0.99f), 0.99f);
    o.w = fminf(fmaxf(acc.w * c_inv_env4[3], -0.99f), 0.99f);
  } else {
    o.x = fminf(fmaxf(acc.x / env[0], -0.99f), 0.99f);
    o.y = fminf(fmaxf(acc.y / env[1], -0.99f), 0.99f);
    o.z = fminf(fmaxf(acc.z / env[2], -0.99f), 0.99f);
    o.w = fminf(fmaxf(acc.w / env[3], -0.99f), 0.99f);
  }
  *reinterpret_cast<float4*>(wav + (long long)b * wav_bstride + 4 * (long long)q) = o;
  if (pcm) {   // the servers' wire format: (speech * 2**15).astype(int16), i.e. truncation toward zero (fastapi/server.py:42)
    short4 s4;
    s4.x = (short)__float2int_rz(o.x * 32768.f);
    s4.y = (short)__float2int_rz(o.y * 32768.f);
    s4.z = (short)__float2int_rz(o.z * 32768.f);
    s4.w = (short)__float2int_rz(o.w * 32768.f);
    *reinterpret_cast<short4*>(pcm + (long long)b * wav_bstride + 4 * (long long)q) = s4;
  }
}
void launch_istft(const float* cp, int F_alloc, int ld, const int* lens, int len_all, float* wav, long long wav_bstride, int B,
                  int max_len, cudaStream_t st, short* pcm) {
  ensure_tables();
  const int hops = max_len * 120;
  istft_kernel<<<dim3((hops + 255) / 256, B), 256, 0, st>>>(cp, F_alloc, ld, lens, len_all, wav, wav_bstride, pcm);
  CV2_LAUNCH_CHECK();
}

// streaming crossfade (common.py:142-150): float64 Hamming window maths, stored as fp32
__global__ void crossfade_kernel(float* __restrict__ speech, const float* __restrict__ old_tail, const double* __restrict__ window,
                                 int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) speech[i] = (float)((double)speech[i] * window[i] + (double)old_tail[i] * window[n + i]);
}
void launch_crossfade(float* speech, const float* old_tail, const double* window, int n, cudaStream_t st) {
  crossfade_kernel<<<(n + 255) / 256, 256, 0, st>>>(speech, old_tail, window, n);
  CV2_LAUNCH_CHECK();
}

// speed change of an offline utterance (cosyvoice/cli/model.py:325-327): F.interpolate(mel, size = int(T / speed), mode = 'linear'),
// align_corners = False: source position (j + 0.5) * T_in / T_out - 0.5 clamped at 0, neighbours blended linearly.
__global__ void mel_time_stretch_kernel(const float* __restrict__ x, int T_in, float* __restrict__ y, int T_out, int rows) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (j >= T_out || r >= rows) return;
  const float scale = (float)T_in / (float)T_out;
  float src = ((float)j + 0.5f) * scale - 0.5f;
  src = src < 0.f ? 0.f : src;
  const int i0 = min((int)src, T_in - 1);
  const int i1 = min(i0 + 1, T_in - 1);
  const float w1 = src - (float)i0, w0 = 1.f - w1;
  const float* xr = x + (long long)r * T_in;
  y[(long long)r * T_out + j] = w0 * xr[i0] + w1 * xr[i1];
}
void launch_mel_time_stretch(const float* x, int T_in, float* y, int T_out, int rows, cudaStream_t st) {
  mel_time_stretch_kernel<<<dim3((T_out + 127) / 128, rows), 128, 0, st>>>(x, T_in, y, T_out, rows);
  CV2_LAUNCH_CHECK();
}

}  // namespace cv2
