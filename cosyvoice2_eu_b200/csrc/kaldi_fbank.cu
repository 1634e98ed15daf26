// x-vector features on the GPU (second half): what CosyVoiceFrontEnd._extract_spk_embedding feeds
// the CAM++ ONNX session (cosyvoice/cli/frontend.py:276-278):
//     feat = torchaudio.compliance.kaldi.fbank(speech, num_mel_bins=80, dither=0, sample_frequency=16000)
//     feat = feat - feat.mean(dim=0, keepdim=True)
// i.e. kaldi's compute-fbank-feats with its defaults: 400-sample frames every 160 samples (snip_edges), per-frame DC removal,
// pre-emphasis 0.97 (first sample replicated), povey window, zero padding to 512, power spectrum, 80 triangular filters on the
// kaldi mel scale between 20 Hz and Nyquist, log with the fp32 epsilon floor; then the per-utterance column mean is removed.
// One block per frame: the 256 useful bins of the 512-point real DFT are 256 dot products of length 400 against an exact
// (cos, sin) table indexed by (n k) mod 512 (the Nyquist bin carries filter weight 0 and is never formed), fp32 accumulation;
// the band sums and the log follow in the same block; a second launch removes the column means (fp64 accumulation).
// Runs once per prompt / speaker (the reference caches the result per speaker, cosyvoice/cli/cosyvoice.py:70-76): it is not a
// throughput kernel, it removes a CPU round trip from the request path.
#include <math.h>

#include <mutex>
#include <vector>

#include "host_util.h"
#include "kaldi_fbank.cuh"

namespace cv2 {

namespace {

constexpr int kWin = 400, kShift = 160, kNfft = 512, kBins = 256, kMels = 80, kMaxBand = 48;

struct FbTables {
  float2* cs = nullptr;    // [512] (cos, sin)(2 pi j / 512)
  float* win = nullptr;    // [400] povey window
  float* mel_w = nullptr;  // [80][kMaxBand]
  int* mel_lo = nullptr;   // [80] first bin of the band
  int* mel_cnt = nullptr;  // [80]
};

FbTables& fb_tables() {
  static std::mutex mu;
  static FbTables tab[64];
  int dev = 0;
  CV2_CUDA(cudaGetDevice(&dev));
  CV2_CHECK(dev >= 0 && dev < 64, "device index %d", dev);
  std::lock_guard<std::mutex> g(mu);
  FbTables& t = tab[dev];
  if (t.cs) return t;
  std::vector<float2> cs(kNfft);
  for (int j = 0; j < kNfft; j++) cs[j] = make_float2((float)cos(2.0 * M_PI * j / kNfft), (float)sin(2.0 * M_PI * j / kNfft));
  std::vector<float> win(kWin);
  for (int n = 0; n < kWin; n++) win[n] = (float)pow(0.5 - 0.5 * cos(2.0 * M_PI * n / (kWin - 1)), 0.85);
  // torchaudio.compliance.kaldi.get_mel_banks(80, 512, 16000, 20, 0 (-> Nyquist), vtln_warp = 1)
  auto mel = [](double f) { return 1127.0 * log(1.0 + f / 700.0); };
  const double lo = mel(20.0), hi = mel(8000.0), delta = (hi - lo) / (kMels + 1);
  std::vector<float> w((size_t)kMels * kMaxBand, 0.f);
  std::vector<int> first(kMels), cnt(kMels);
  for (int m = 0; m < kMels; m++) {
    const double left = lo + m * delta, center = lo + (m + 1) * delta, right = lo + (m + 2) * delta;
    int f0 = -1, n = 0;
    for (int k = 0; k < kBins; k++) {
      const double mk = mel(16000.0 / kNfft * k);
      const double v = fmax(0.0, fmin((mk - left) / (center - left), (right - mk) / (right - center)));
      if (v > 0.0) {
        if (f0 < 0) f0 = k;
        CV2_CHECK(k - f0 < kMaxBand, "fbank: band %d wider than %d bins", m, kMaxBand);
        w[(size_t)m * kMaxBand + (k - f0)] = (float)v;
        n = k - f0 + 1;
      }
    }
    first[m] = f0 < 0 ? 0 : f0;
    cnt[m] = n;
  }
  CV2_CUDA(cudaMalloc(&t.win, kWin * sizeof(float)));
  CV2_CUDA(cudaMalloc(&t.mel_w, w.size() * sizeof(float)));
  CV2_CUDA(cudaMalloc(&t.mel_lo, kMels * sizeof(int)));
  CV2_CUDA(cudaMalloc(&t.mel_cnt, kMels * sizeof(int)));
  CV2_CUDA(cudaMemcpy(t.win, win.data(), kWin * sizeof(float), cudaMemcpyHostToDevice));
  CV2_CUDA(cudaMemcpy(t.mel_w, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice));
  CV2_CUDA(cudaMemcpy(t.mel_lo, first.data(), kMels * sizeof(int), cudaMemcpyHostToDevice));
  CV2_CUDA(cudaMemcpy(t.mel_cnt, cnt.data(), kMels * sizeof(int), cudaMemcpyHostToDevice));
  float2* d = nullptr;
  CV2_CUDA(cudaMalloc(&d, kNfft * sizeof(float2)));
  CV2_CUDA(cudaMemcpy(d, cs.data(), kNfft * sizeof(float2), cudaMemcpyHostToDevice));
  t.cs = d;
  return t;
}

__device__ __forceinline__ int fb_frames(int n) { return n < kWin ? 0 : 1 + (n - kWin) / kShift; }

__global__ void __launch_bounds__(256)
kaldi_fbank_kernel(const float* __restrict__ wav, long long wav_stride, const int* __restrict__ n_samples,
                   const float2* __restrict__ cs_g, const float* __restrict__ win, const float* __restrict__ mel_w,
                   const int* __restrict__ mel_lo, const int* __restrict__ mel_cnt, float* __restrict__ out, int T_out,
                   int* __restrict__ out_len) {
  __shared__ float y[kWin];
  __shared__ float2 cs[kNfft];
  __shared__ float power[kBins];
  __shared__ float red[8];
  const int f = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int frames = fb_frames(n_samples[b]);
  if (f == 0 && tid == 0 && out_len) out_len[b] = frames;
  float* o = out + ((long long)b * T_out + f) * kMels;
  if (f >= frames) {
    if (tid < kMels) o[tid] = 0.f;
    return;
  }
  const float* x = wav + (long long)b * wav_stride + (long long)f * kShift;
  for (int i = tid; i < kNfft; i += 256) cs[i] = cs_g[i];
  float part = 0.f;
  for (int i = tid; i < kWin; i += 256) {
    const float v = x[i];
    y[i] = v;
    part += v;
  }
  for (int s = 16; s > 0; s >>= 1) part += __shfl_xor_sync(0xffffffffu, part, s);
  if ((tid & 31) == 0) red[tid >> 5] = part;
  __syncthreads();
  float mean = 0.f;
#pragma unroll
  for (int i = 0; i < 8; i++) mean += red[i];
  mean *= 1.f / kWin;
  // pre-emphasis on the DC-free frame, first sample against itself; then the povey window.  Two steps: every thread needs its
  // left neighbour's ORIGINAL value.
  float e[2];
  int cntv = 0;
  for (int i = tid; i < kWin; i += 256) {
    const float cur = y[i] - mean, prev = y[i > 0 ? i - 1 : 0] - mean;
    e[cntv++] = (cur - 0.97f * prev) * win[i];
  }
  __syncthreads();
  cntv = 0;
  for (int i = tid; i < kWin; i += 256) y[i] = e[cntv++];
  __syncthreads();
  // bin k = tid of the 512-point DFT of the zero-padded frame
  {
    const int k = tid;
    float re0 = 0.f, im0 = 0.f, re1 = 0.f, im1 = 0.f;
    int j = 0;   // (n * k) mod 512, advanced incrementally
    for (int n = 0; n < kWin; n += 2) {
      const float2 c0 = cs[j];
      j = (j + k) & (kNfft - 1);
      const float2 c1 = cs[j];
      j = (j + k) & (kNfft - 1);
      const float a0 = y[n], a1 = y[n + 1];
      re0 = fmaf(a0, c0.x, re0);
      im0 = fmaf(a0, c0.y, im0);
      re1 = fmaf(a1, c1.x, re1);
      im1 = fmaf(a1, c1.y, im1);
    }
    const float re = re0 + re1, im = im0 + im1;
    power[k] = re * re + im * im;
  }
  __syncthreads();
  if (tid < kMels) {
    const int lo = mel_lo[tid], n = mel_cnt[tid];
    float acc = 0.f;
    for (int j = 0; j < n; j++) acc = fmaf(power[lo + j], mel_w[tid * kMaxBand + j], acc);
    o[tid] = logf(fmaxf(acc, 1.1920928955078125e-07f));
  }
}

// feat[b, :, m] -= mean over the utterance's own frames (frontend.py:278)
__global__ void fbank_cmn_kernel(float* __restrict__ feat, const int* __restrict__ n_samples, int T_out) {
  const int b = blockIdx.x, m = threadIdx.x;
  const int frames = fb_frames(n_samples[b]);
  if (m >= kMels || frames == 0) return;
  float* col = feat + (long long)b * T_out * kMels + m;
  double s = 0.0;
  for (int f = 0; f < frames; f++) s += (double)col[(long long)f * kMels];
  const float mean = (float)(s / frames);
  for (int f = 0; f < frames; f++) col[(long long)f * kMels] -= mean;
}

}  // namespace

int kaldi_fbank_frames(int n_samples) { return n_samples < kWin ? 0 : 1 + (n_samples - kWin) / kShift; }

void launch_kaldi_fbank(const float* wav16, long long wav_stride, const int* n_samples, int B, int max_samples, float* feat,
                        int* feat_len, int subtract_mean, cudaStream_t st) {
  CV2_CHECK(B > 0 && max_samples >= kWin, "kaldi_fbank: a frame needs %d samples (got %d)", kWin, max_samples);
  FbTables& t = fb_tables();
  const int T_out = kaldi_fbank_frames(max_samples);
  kaldi_fbank_kernel<<<dim3(T_out, B), 256, 0, st>>>(wav16, wav_stride, n_samples, t.cs, t.win, t.mel_w, t.mel_lo, t.mel_cnt, feat, T_out,
                                                      feat_len);
  CV2_CUDA(cudaGetLastError());
  if (subtract_mean) {
    fbank_cmn_kernel<<<B, 96, 0, st>>>(feat, n_samples, T_out);
    CV2_CUDA(cudaGetLastError());
  }
}

}  // namespace cv2
