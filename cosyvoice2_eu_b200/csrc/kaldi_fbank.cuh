// x-vector features (kaldi_fbank.cu): torchaudio.compliance.kaldi.fbank(num_mel_bins=80, dither=0, 16 kHz) minus its column mean.
#pragma once
#include <cuda_runtime.h>

namespace cv2 {

int kaldi_fbank_frames(int n_samples);     // 1 + (n - 400) / 160, or 0
// wav16 [B, wav_stride] fp32 at 16 kHz, n_samples [B] (device) -> feat [B, kaldi_fbank_frames(max_samples), 80] fp32 (rows past an
// utterance's own frames are zero) and feat_len [B] (device, may be null); subtract_mean: remove each utterance's column means
void launch_kaldi_fbank(const float* wav16, long long wav_stride, const int* n_samples, int B, int max_samples, float* feat,
                        int* feat_len, int subtract_mean, cudaStream_t st);

}  // namespace cv2
