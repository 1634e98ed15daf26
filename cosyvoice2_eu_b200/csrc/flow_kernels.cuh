// Launchers of the flow-side glue kernels (flow_kernels.cu).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace cv2 {

void launch_embed_tokens(const int* prompt_tok, const int* prompt_len, int prompt_stride, const int* tok, const int* tok_len, int tok_stride, const float* table, __half* out, int B, int T_alloc, int vocab, cudaStream_t st);
void launch_layernorm512(const float* x, const float* g, const float* b, float eps, __half* out16, float* out32, const int* lens, int len_all, int S, int T_alloc, cudaStream_t st);
void launch_repeat2(const __half* in, __half* out, const int* lens, int S, int T_in_alloc, int T_out_alloc, int C, cudaStream_t st);
void launch_pos_table(__half* out, int Tmax, cudaStream_t st);
void launch_spk_affine(const float* emb, const float* w, const float* bias, float* out, int B, cudaStream_t st);
void launch_time_mlp(const float* t, int nt, const float* w1, const float* b1, const float* w2, const float* b2, float* h1, float* temb, cudaStream_t st);
void launch_resnet_time_proj(const float* temb, int nt, const float* w, const float* b, float* out, cudaStream_t st);
void launch_pack_cond(const float* mu, const float* spks, const float* cond, __half* xin, const int* lens, int B, int T_alloc, cudaStream_t st);
void launch_euler_pack(float* x, const float* v, const float* noise, int noise_stride, __half* xin, const int* lens, int B, int T_alloc, float dt, float cfg, int init, cudaStream_t st);
void launch_nct_to_ntc(const float* in, long long in_bstride, int in_T, float* out32, __half* out16, const int* lens, int len_all, int B, int T_alloc, int C, int ldc, int col_off, int t_src_off, cudaStream_t st);
void launch_ntc_to_nct(const float* in, int T_alloc, int ldc, int t_off, const int* t_offs, float* out, long long out_bstride, int T_out, int C, const int* in_lens, int B, cudaStream_t st);
void launch_build_cond(const float* prompt_feat, long long pf_bstride, const int* pf_len, float* cond, int B, int T_alloc, cudaStream_t st);
void launch_lens_affine(const int* a, const int* b, int mul, int add, int* out, int n, cudaStream_t st);
void launch_mask_to_lens(const float* mask, int T, int* lens, int S, cudaStream_t st);
void launch_bcast_rows16(const float* v, int C, __half* out, int ldc, int col_off, const int* lens, int S, int T_alloc, cudaStream_t st);
void launch_f32_to_f16(const float* in, __half* out, long long n, cudaStream_t st);

void launch_split16(const float* x, int C, __half* out, int ld, int lo_off, long long rows, cudaStream_t st);

// encoder slot: fp32 embeddings [B, T_in, 512] (+ optional 3-row look-ahead context) -> 16-bit A operand rows, zero padded
void launch_embed_rows(const float* xs, int T_in, const int* lens, const float* ctx, int n_ctx, __half* out, int B, int T_alloc,
                       cudaStream_t st);
// incremental streaming (engine.h StreamState): all per-slot lengths of a non-final chunk call in one launch.
// active slot (token_len > 0): len_ctx = prompt + token, len_enc = len_ctx - 3, len_mel = 2 len_enc (both CFG rows),
// t_lo = floor(t_done / 128) * 128, then t_done = len_mel; inactive slot: every length 0, t_done untouched.
void launch_stream_lens(const int* prompt_len, const int* token_len, int* t_done, int* len_ctx, int* len_enc, int* len_mel, int* t_lo,
                        int B, cudaStream_t st);
// Causal-conv input rows across 128-row tile boundaries: for boundary b (row 128 b) of sequence s, restore rows 128 b - 2, 128 b - 1
// of `buf` from the cache if the call starts there (t_lo[s] == 128 b), save them if the call has just produced them
// (t_lo[s] < 128 b <= lens[s]).  buf: 16-bit [S, T_alloc, ld] (first C columns), cache: [S][T_alloc / 128][2][512].
void launch_tail_swap(__half* buf, int S, int T_alloc, int C, long long ld, const int* lens, const int* t_lo, __half* cache,
                      cudaStream_t st);
// max |x| of a strided 16-bit matrix into *slot (fp32 bit pattern, atomicMax; NaN counts as +inf)
void launch_absmax16(const __half* p, long long rows, int cols, long long ld, unsigned* slot, cudaStream_t st);
// out[i] = min(a[i], hi) + add
void launch_lens_clamp(const int* a, int hi, int add, int* out, int n, cudaStream_t st);
}  // namespace cv2
