// Prompt log-mel front end (prompt_mel.cu): matcha `mel_spectrogram` at the CosyVoice2 settings, batched.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace cv2 {

int prompt_mel_frames(int n_samples);                         // frames of one prompt (0 when it is too short to reflect-pad)
size_t prompt_mel_workspace_bytes(int B, int max_samples);    // magnitude spectrum scratch
// wav [B, wav_stride] fp32 in [-1, 1] at 24 kHz, n_samples [B] (device) -> mel [B, prompt_mel_frames(max_samples), 80] fp32
// (rows past a prompt's own frame count are zero) and mel_len [B] (device, may be null)
void launch_prompt_mel(const float* wav, long long wav_stride, const int* n_samples, int B, int max_samples, float* mel,
                       int* mel_len, void* ws, size_t ws_bytes, cudaStream_t st);

// torchaudio Resample(16000, 24000): wav16 [B, in_stride] -> wav24 [B, out_stride] (rows zero-padded to ceil(3 max_in / 2)), n_out [B]
int resample_16k_24k_len(int n_in);
void launch_resample_16k_24k(const float* wav16, long long in_stride, const int* n_in, int B, int max_in, float* wav24,
                             long long out_stride, int* n_out, cudaStream_t st);

}  // namespace cv2
