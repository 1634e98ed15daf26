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

}  // namespace cv2
