// Prompt log-mel front end: the reference's `mel_spectrogram`
// (third_party/Matcha-TTS/matcha/utils/audio.py:45-82) at the cosyvoice2.yaml:152-160 settings -- n_fft = win = 1920, hop 480,
// 80 Slaney mels over [0, 8000] Hz at 24 kHz, reflect pad 720, centre = False, log(clamp(., 1e-5)) -- for a batch of prompts.
//
// fp32 throughout (the result is a log spectrum; 16-bit operands are not an option).  The 1920-point real DFT of a frame is
// evaluated as a dense fp32 contraction with K halved by the symmetry of the (periodic Hann) window, w[n] = w[N - n]:
//     Re X[k] = sum_{n=0}^{960} e[n] cos(2 pi k n / N),   Im X[k] = -sum_{n=1}^{959} o[n] sin(2 pi k n / N),
//     e[n] = w[n] (x[n] + x[N - n]),  o[n] = w[n] (x[n] - x[N - n])   (n = 1..959),   e[0] = w[0] x[0],  e[960] = w[960] x[960],
// so one frame costs 2 x 961 x 961 FMAs instead of 2 x 1920 x 961.  `pm_dft_mag_kernel` is a 128 frames x 64 bins register-tiled
// SGEMM whose A tile is folded on the fly from the waveform (reflect indexing + window), whose B tile is an exact (cos, sin)
// table (argument reduced in integers, evaluated in fp64 once per device) and whose inner product is packed
// `fma.rn.f32x2` on (re, im) pairs; its epilogue writes |X| = sqrt(re^2 + im^2 + 1e-9).  `pm_mel_log_kernel` applies the 80
// triangular filters (compact per-band weights) and the clamped log, writing [B, T, 80] -- the layout `flow.inference` takes.
#include <math.h>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "host_util.h"
#include "prompt_mel.cuh"

namespace cv2 {

namespace {

constexpr int kNfft = 1920, kHop = 480, kPad = 720, kBins = 961, kMels = 80;
constexpr int kKfold = 961;        // folded contraction length (n = 0..960)
constexpr int kKpad = 976;         // padded to a multiple of the k step
constexpr int kBinsPad = 1024;     // table / spectrum row length
constexpr int kBandMax = 64;       // widest mel band in bins (top band: 49)
constexpr int kTM = 128, kTN = 64, kTK = 16;
constexpr int kDftSmemA = 2 * kTK * (kTM + 2) * 8, kDftSmem = kDftSmemA + 2 * kTK * kTN * 8;

struct Tables {
  float2* cs = nullptr;     // [kKpad][kBinsPad] (cos, sin)(2 pi n k / N); zero outside n < 961, k < 961
  float* win = nullptr;     // [kKpad] periodic Hann, zero for n > 960
  float* mel_w = nullptr;   // [kBandMax][kMels] band weights, j-th bin of band m at [j][m]
  int* mel_lo = nullptr;    // [kMels] first bin of the band
  int* mel_cnt = nullptr;   // [kMels] bins in the band
};

__global__ void pm_table_kernel(float2* cs, float* win) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (k >= kBinsPad) return;
  float2 v = make_float2(0.f, 0.f);
  if (n < kKfold && k < kBins) {
    const int m = (int)(((long long)n * k) % kNfft);           // exact argument reduction
    double s, c;
    sincospi(2.0 * (double)m / (double)kNfft, &s, &c);
    v = make_float2((float)c, (float)s);
  }
  cs[(size_t)n * kBinsPad + k] = v;
  if (k == 0) win[n] = n < kKfold ? (float)(0.5 - 0.5 * cospi(2.0 * (double)n / (double)kNfft)) : 0.f;
}

// reflect padding needs more than kPad samples (the reference raises below that): such a row yields no frames
__device__ __forceinline__ int frames_of(int n_samples) { return n_samples <= kPad ? 0 : (n_samples - kHop) / kHop + 1; }

__device__ __forceinline__ float wave_reflect(const float* __restrict__ w, int i, int L) {
  i = i < 0 ? -i : i;                  // left reflect pad (no edge repeat)
  i = i >= L ? 2 * (L - 1) - i : i;    // right reflect pad
  return __ldg(w + i);
}

// grid (frame tiles, bin tiles, B), 256 threads; thread (ty, tx) owns frames ty*8..+7 and bins tx*2, tx*2+1, 32+tx*2, 33+tx*2 of the 128 x 64 tile
// (32 packed accumulators; per k: 6 LDS.128 feed 32 FFMA2).
__global__ void __launch_bounds__(256, 2) pm_dft_mag_kernel(const float* __restrict__ wav, long long wav_stride,
                                                         const int* __restrict__ n_samples, const float2* __restrict__ cs,
                                                         const float* __restrict__ win, float* __restrict__ spec, int T_alloc) {
  extern __shared__ __align__(16) uint8_t pm_smem[];
  // (e, o) tile, rows padded by 16 B: the k-fastest stores below are 2-way conflicts, not 16-way; then the (cos, sin) tile
  float2 (*As)[kTK][kTM + 2] = reinterpret_cast<float2 (*)[kTK][kTM + 2]>(pm_smem);
  float2 (*Bs)[kTK][kTN] = reinterpret_cast<float2 (*)[kTK][kTN]>(pm_smem + kDftSmemA);
  const int b = blockIdx.z;
  const int L = __ldg(n_samples + b);
  const int T = frames_of(L);
  const int t0 = blockIdx.x * kTM;
  if (t0 >= T) return;
  const int k0 = blockIdx.y * kTN;
  const float* w = wav + (long long)b * wav_stride;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int lj = tid & 15, lfs = tid >> 4;    // A loader: n fastest -- a warp reads two frames' 64-byte runs (forward and mirrored), 2-4 lines
                                              // per instruction instead of 32 (ncu: the frame-fastest mapping was L1TEX bound at 92 %)
  const int bk = tid >> 4, bn = (tid & 15) * 4;  // B loader: one k row, 4 bins

  float2 acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = make_float2(0.f, 0.f);

  float2 ra[8];
  float4 rb[2];
  auto fetch = [&](int kb) {
    const int n = kb + lj;
    const float wn = n < kKfold ? __ldg(win + n) : 0.f;
    const bool single = n == 0 || n == kNfft / 2;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int f = lfs + 16 * i;
      const int base = (t0 + f) * kHop - kPad;      // first sample of the frame in the unpadded signal
      float e = 0.f, o = 0.f;
      if (t0 + f < T && n < kKfold) {
        const float a = wave_reflect(w, base + n, L) * wn;
        if (single) {
          e = a;
        } else {
          const float c = wave_reflect(w, base + kNfft - n, L) * wn;
          e = a + c;
          o = a - c;
        }
      }
      ra[i] = make_float2(e, o);
    }
    const float4* src = reinterpret_cast<const float4*>(cs + (size_t)(kb + bk) * kBinsPad + k0 + bn);
    rb[0] = __ldg(src);
    rb[1] = __ldg(src + 1);
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 8; i++) As[buf][lj][lfs + 16 * i] = ra[i];
    float4* dst = reinterpret_cast<float4*>(&Bs[buf][bk][bn]);
    dst[0] = rb[0];
    dst[1] = rb[1];
  };

  fetch(0);
  stash(0);
  __syncthreads();
  constexpr int kSteps = kKpad / kTK;
  for (int s = 0; s < kSteps; s++) {
    const int buf = s & 1;
    if (s + 1 < kSteps) fetch((s + 1) * kTK);
#pragma unroll
    for (int kk = 0; kk < kTK; kk++) {
      float2 a[8], bb[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const float4 f = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8 + 2 * i]);
        a[2 * i] = make_float2(f.x, f.y);
        a[2 * i + 1] = make_float2(f.z, f.w);
      }
#pragma unroll
      for (int j = 0; j < 2; j++) {
        const float4 f = *reinterpret_cast<const float4*>(&Bs[buf][kk][32 * j + tx * 2]);   // 16 lanes x 16 B contiguous
        bb[2 * j] = make_float2(f.x, f.y);
        bb[2 * j + 1] = make_float2(f.z, f.w);
      }
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = ffma2(a[i], bb[j], acc[i][j]);
    }
    if (s + 1 < kSteps) {
      stash(buf ^ 1);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int t = t0 + ty * 8 + i;
    if (t >= T) continue;
    float* row = spec + ((size_t)b * T_alloc + t) * kBinsPad + k0 + tx * 2;    // bins tx*2, tx*2+1 and 32 + the same
#pragma unroll
    for (int j = 0; j < 2; j++) {
      float2 m;
      m.x = sqrtf(acc[i][2 * j].x * acc[i][2 * j].x + acc[i][2 * j].y * acc[i][2 * j].y + 1e-9f);
      m.y = sqrtf(acc[i][2 * j + 1].x * acc[i][2 * j + 1].x + acc[i][2 * j + 1].y * acc[i][2 * j + 1].y + 1e-9f);
      *reinterpret_cast<float2*>(row + 32 * j) = m;
    }
  }
}

// one block of 128 threads per 4 frames: the spectrum rows go through shared memory, thread m < 80 owns one band per frame
__global__ void __launch_bounds__(128) pm_mel_log_kernel(const float* __restrict__ spec, int T_alloc, const int* __restrict__ n_samples,
                                                         const float* __restrict__ mel_w, const int* __restrict__ mel_lo,
                                                         const int* __restrict__ mel_cnt, float* __restrict__ mel, int T_out,
                                                         int* __restrict__ mel_len) {
  __shared__ float row[4][kBinsPad];
  const int b = blockIdx.y;
  const int T = frames_of(__ldg(n_samples + b));
  const int t0 = blockIdx.x * 4;
  if (blockIdx.x == 0 && threadIdx.x == 0 && mel_len) mel_len[b] = T;
  if (t0 >= T_out) return;
  for (int f = 0; f < 4; f++) {
    const int t = t0 + f;
    if (t < T)
      for (int k = threadIdx.x; k < kBinsPad; k += 128) row[f][k] = spec[((size_t)b * T_alloc + t) * kBinsPad + k];
  }
  __syncthreads();
  const int m = threadIdx.x;
  if (m >= kMels) return;
  const int lo = __ldg(mel_lo + m), cnt = __ldg(mel_cnt + m);
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  for (int j = 0; j < cnt; j++) {
    const float wj = __ldg(mel_w + j * kMels + m);
#pragma unroll
    for (int f = 0; f < 4; f++) s[f] = fmaf(wj, row[f][lo + j], s[f]);
  }
#pragma unroll
  for (int f = 0; f < 4; f++) {
    const int t = t0 + f;
    if (t >= T_out) break;
    mel[((size_t)b * T_out + t) * kMels + m] = t < T ? logf(fmaxf(s[f], 1e-5f)) : 0.f;   // audio.py:22-23; padding rows are zero
  }
}

// 16 kHz -> 24 kHz: torchaudio.transforms.Resample(16000, 24000) (cosyvoice/cli/frontend.py:495,541): polyphase FIR, 3 phases x 16
// taps of a Hann-windowed sinc (lowpass_filter_width 6, rolloff 0.99): y[3 i + p] = sum_k x[2 i + k - 7] h[p][k], zero outside.
// HBM bound: 8 B read + 12 B written per input pair.
__constant__ float c_rs[3][16];

__global__ void __launch_bounds__(256) pm_resample_kernel(const float* __restrict__ x, long long x_stride, const int* __restrict__ n_in,
                                                          float* __restrict__ y, long long y_stride, int* __restrict__ n_out,
                                                          int max_out) {
  const int b = blockIdx.y;
  const int L = __ldg(n_in + b);
  const int Lo = (3 * L + 1) / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && n_out) n_out[b] = Lo;
  if (3 * i >= max_out) return;
  const float* xr = x + (long long)b * x_stride;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  if (3 * i < Lo) {
#pragma unroll
    for (int k = 0; k < 16; k++) {
      const int j = 2 * i + k - 7;
      const float v = (j >= 0 && j < L) ? __ldg(xr + j) : 0.f;
      a0 = fmaf(v, c_rs[0][k], a0);
      a1 = fmaf(v, c_rs[1][k], a1);
      a2 = fmaf(v, c_rs[2][k], a2);
    }
  }
  float* yr = y + (long long)b * y_stride + 3 * i;
  const float o[3] = {a0, a1, a2};
#pragma unroll
  for (int p_ = 0; p_ < 3; p_++)
    if (3 * i + p_ < max_out) yr[p_] = 3 * i + p_ < Lo ? o[p_] : 0.f;    // rows are zero-padded to max_out
}

void build_resample(float h[3][16]) {
  const double base = 2.0 * 0.99, lpw = 6.0, pi = 3.14159265358979323846;
  for (int p_ = 0; p_ < 3; p_++) {
    const double ph = (double)((float)(-p_) / 3.0f);     // torch forms the phase offset in float32, the rest in float64
    for (int k = 0; k < 16; k++) {
      double t = (ph + (double)(k - 7) / 2.0) * base;
      t = t < -lpw ? -lpw : (t > lpw ? lpw : t);
      const double c = cos(t * pi / lpw / 2.0);
      const double tp = t * pi;
      const double sinc = tp == 0.0 ? 1.0 : sin(tp) / tp;
      h[p_][k] = (float)(sinc * (c * c) * (base / 2.0));
    }
  }
}

// librosa.filters.mel(sr=24000, n_fft=1920, n_mels=80, fmin=0, fmax=8000), Slaney scale + Slaney norm, float32 (audio.py:52)
double hz_to_mel(double f) {
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
  return f >= min_log_hz ? min_log_mel + log(f / min_log_hz) / logstep : f / f_sp;
}
double mel_to_hz(double m) {
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
  return m >= min_log_mel ? min_log_hz * exp(logstep * (m - min_log_mel)) : f_sp * m;
}

void build_mel(std::vector<float>& w, std::vector<int>& lo, std::vector<int>& cnt) {
  const double sr = 24000.0, f_lo = 0.0, f_hi = 8000.0;
  std::vector<double> mel_f(kMels + 2);
  const double m0 = hz_to_mel(f_lo), m1 = hz_to_mel(f_hi);
  for (int i = 0; i < kMels + 2; i++) mel_f[i] = mel_to_hz(m0 + (m1 - m0) * (double)i / (double)(kMels + 1));
  w.assign((size_t)kBandMax * kMels, 0.f);
  lo.assign(kMels, 0);
  cnt.assign(kMels, 0);
  for (int m = 0; m < kMels; m++) {
    const double enorm = 2.0 / (mel_f[m + 2] - mel_f[m]);
    int first = -1, last = -1;
    std::vector<float> full(kBins);
    for (int k = 0; k < kBins; k++) {
      const double f = (sr / 2.0) * (double)k / (double)(kBins - 1);
      const double lower = (f - mel_f[m]) / (mel_f[m + 1] - mel_f[m]);
      const double upper = (mel_f[m + 2] - f) / (mel_f[m + 2] - mel_f[m + 1]);
      const double lu = lower < upper ? lower : upper;
      const float tri = (float)(lu > 0.0 ? lu : 0.0);   // float32 array, then *= enorm in place
      full[k] = (float)((double)tri * enorm);
      if (full[k] != 0.f) {
        if (first < 0) first = k;
        last = k;
      }
    }
    if (first < 0) continue;
    CV2_CHECK(last - first + 1 <= kBandMax, "mel band %d spans %d bins", m, last - first + 1);
    lo[m] = first;
    cnt[m] = last - first + 1;
    for (int j = 0; j < cnt[m]; j++) w[(size_t)j * kMels + m] = full[first + j];
  }
}

Tables& tables_for_device() {
  static std::mutex mu;
  static Tables tab[64];
  int dev = 0;
  CV2_CUDA(cudaGetDevice(&dev));
  CV2_CHECK(dev >= 0 && dev < 64, "device index %d", dev);
  std::lock_guard<std::mutex> g(mu);
  Tables& t = tab[dev];
  if (t.cs) return t;
  std::vector<float> w;
  std::vector<int> lo, cnt;
  build_mel(w, lo, cnt);
  CV2_CUDA(cudaMalloc(&t.win, kKpad * sizeof(float)));
  CV2_CUDA(cudaMalloc(&t.mel_w, w.size() * sizeof(float)));
  CV2_CUDA(cudaMalloc(&t.mel_lo, kMels * sizeof(int)));
  CV2_CUDA(cudaMalloc(&t.mel_cnt, kMels * sizeof(int)));
  CV2_CUDA(cudaMemcpy(t.mel_w, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice));
  CV2_CUDA(cudaMemcpy(t.mel_lo, lo.data(), kMels * sizeof(int), cudaMemcpyHostToDevice));
  CV2_CUDA(cudaMemcpy(t.mel_cnt, cnt.data(), kMels * sizeof(int), cudaMemcpyHostToDevice));
  float h[3][16];
  build_resample(h);
  CV2_CUDA(cudaMemcpyToSymbol(c_rs, h, sizeof(h)));
  float2* cs = nullptr;
  CV2_CUDA(cudaMalloc(&cs, (size_t)kKpad * kBinsPad * sizeof(float2)));
  pm_table_kernel<<<dim3(kBinsPad / 256, kKpad), 256>>>(cs, t.win);
  CV2_CUDA(cudaGetLastError());
  CV2_CUDA(cudaDeviceSynchronize());
  CV2_CUDA(cudaFuncSetAttribute(pm_dft_mag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDftSmem));
  t.cs = cs;
  return t;
}

}  // namespace

int prompt_mel_frames(int n_samples) { return n_samples <= kPad ? 0 : (n_samples - kHop) / kHop + 1; }

static int t_alloc_of(int max_samples) { return (prompt_mel_frames(max_samples) + kTM - 1) / kTM * kTM; }

size_t prompt_mel_workspace_bytes(int B, int max_samples) {
  return (size_t)B * t_alloc_of(max_samples) * kBinsPad * sizeof(float);
}

void launch_prompt_mel(const float* wav, long long wav_stride, const int* n_samples, int B, int max_samples, float* mel,
                       int* mel_len, void* ws, size_t ws_bytes, cudaStream_t st) {
  CV2_CHECK(B > 0 && max_samples > kPad, "prompt_mel: reflect padding needs more than %d samples (got %d)", kPad, max_samples);
  CV2_CHECK(ws_bytes >= prompt_mel_workspace_bytes(B, max_samples), "prompt_mel: workspace too small");
  Tables& t = tables_for_device();
  const int T_out = prompt_mel_frames(max_samples);
  const int T_alloc = t_alloc_of(max_samples);
  float* spec = static_cast<float*>(ws);
  pm_dft_mag_kernel<<<dim3(T_alloc / kTM, kBinsPad / kTN, B), 256, kDftSmem, st>>>(wav, wav_stride, n_samples, t.cs, t.win, spec, T_alloc);
  CV2_CUDA(cudaGetLastError());
  pm_mel_log_kernel<<<dim3((T_out + 3) / 4, B), 128, 0, st>>>(spec, T_alloc, n_samples, t.mel_w, t.mel_lo, t.mel_cnt, mel, T_out, mel_len);
  CV2_CUDA(cudaGetLastError());
}

int resample_16k_24k_len(int n_in) { return (3 * n_in + 1) / 2; }

void launch_resample_16k_24k(const float* wav16, long long in_stride, const int* n_in, int B, int max_in, float* wav24,
                             long long out_stride, int* n_out, cudaStream_t st) {
  CV2_CHECK(B > 0 && max_in > 0, "resample: empty input");
  const int max_out = resample_16k_24k_len(max_in);
  CV2_CHECK(out_stride >= max_out, "resample: output rows of %lld samples cannot hold %d", out_stride, max_out);
  tables_for_device();
  const int n_i = (max_out + 2) / 3;
  pm_resample_kernel<<<dim3((n_i + 255) / 256, B), 256, 0, st>>>(wav16, in_stride, n_in, wav24, out_stride, n_out, max_out);
  CV2_CUDA(cudaGetLastError());
}

}  // namespace cv2
