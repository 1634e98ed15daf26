// Launchers of the HiFT non-GEMM kernels (hift_kernels.cu).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace cv2 {

void launch_conv3_elu_f32(const float* x, int Cin, const float* w, const float* bias, float* y, int Cout, const int* lens, int len_all, int B, int T_alloc, cudaStream_t st);
void launch_f0_head(const float* x, const float* w, const float* b, float* f0, const int* lens, int len_all, int B, int T_alloc, int f0_stride, cudaStream_t st);
void launch_nsf_source(const float* f0, int f0_stride, float* P, int T_alloc, const int* lens, int len_all, const float* noise, long long noise_bstride, unsigned long long seed, const unsigned long long* seed_ptr, const float* lw, const float* lb, const float* cache, int cache_len, long long cache_bstride, float* src, long long src_bstride, int B, int max_len, cudaStream_t st);
void hift_init_tables();
void launch_source_stft(const float* src, long long src_bstride, const int* lens, int len_all, float* out, int F_alloc, int B, cudaStream_t st);
void launch_source_down(const float* stft, int F_alloc, const float* w, const float* bias, int k, int stride, int pad, int C, const int* lens, int len_all, int frames_per_len, int frames_add, float* out32, __half* out16, const float* alpha, int B, int T_alloc, cudaStream_t st);
void launch_reflect_row0(float* x, int B, int T_alloc, int C, cudaStream_t st);
void launch_istft(const float* cp, int F_alloc, int ld, const int* lens, int len_all, float* wav, long long wav_bstride, int B, int max_len, cudaStream_t st,
                  short* pcm = nullptr);   // optional int16 PCM copy of the waveform (same layout)
void launch_crossfade(float* speech, const float* old_tail, const double* window, int n, cudaStream_t st);
void launch_mel_time_stretch(const float* x, int T_in, float* y, int T_out, int rows, cudaStream_t st);

}  // namespace cv2
