// Bandwidth-bound glue kernels of flow.inference (reference: cosyvoice/flow/flow.py:235-283,
// flow_matching.py:71-123, matcha decoder.py:14-29,73-117).  Everything here is coalesced,
// vectorised where rows allow, and does no host synchronisation.
#include "flow_kernels.cuh"
#include "common.cuh"
#include "host_util.h"

namespace cv2 {

// ---- token embedding: clamp(tok, 0) -> Embedding(6561, 512) * pad-mask -> 16-bit A operand -------------
__global__ void embed_tokens_kernel(const int* __restrict__ prompt_tok, const int* __restrict__ prompt_len, int prompt_stride,
                                    const int* __restrict__ tok, const int* __restrict__ tok_len, int tok_stride,
                                    const float* __restrict__ table, __half* __restrict__ out, int T_alloc, int vocab) {
  const int t = blockIdx.x, b = blockIdx.y;
  const int pl = prompt_len[b], tl = tok_len[b];
  __half* o = out + ((long long)b * T_alloc + t) * 512;
  int id = -1;
  if (t < pl) id = prompt_tok[(long long)b * prompt_stride + t];
  else if (t < pl + tl) id = tok[(long long)b * tok_stride + (t - pl)];
  if (t < pl + tl) {
    id = max(id, 0);
    id = min(id, vocab - 1);
    const float* row = table + (long long)id * 512;
    for (int c = threadIdx.x; c < 512; c += blockDim.x) o[c] = __float2half_rn(row[c]);
  } else {
    for (int c = threadIdx.x; c < 512; c += blockDim.x) o[c] = __float2half_rn(0.f);
  }
}
void launch_embed_tokens(const int* prompt_tok, const int* prompt_len, int prompt_stride, const int* tok, const int* tok_len,
                         int tok_stride, const float* table, __half* out, int B, int T_alloc, int vocab, cudaStream_t st) {
  embed_tokens_kernel<<<dim3(T_alloc, B), 128, 0, st>>>(prompt_tok, prompt_len, prompt_stride, tok, tok_len, tok_stride, table,
                                                         out, T_alloc, vocab);
  CV2_LAUNCH_CHECK();
}

// ---- encoder slot (upsample_encoder.py:243-251): the caller hands over input_embedding(token) * mask and, for a non-final
//      streaming chunk, the 3 look-ahead embeddings as `context`; rows [len, len + n_ctx) of the A operand take the context
__global__ void embed_rows_kernel(const float* __restrict__ xs, int T_in, const int* __restrict__ lens, const float* __restrict__ ctx,
                                  int n_ctx, __half* __restrict__ out, int T_alloc) {
  const int t = blockIdx.x, b = blockIdx.y;
  const int len = lens[b];
  const float* row = nullptr;
  if (t < len) row = xs + ((long long)b * T_in + t) * 512;
  else if (ctx && t < len + n_ctx) row = ctx + ((long long)b * n_ctx + (t - len)) * 512;
  __half* o = out + ((long long)b * T_alloc + t) * 512;
  for (int c = threadIdx.x; c < 512; c += blockDim.x) o[c] = __float2half_rn(row ? row[c] : 0.f);
}
void launch_embed_rows(const float* xs, int T_in, const int* lens, const float* ctx, int n_ctx, __half* out, int B, int T_alloc,
                       cudaStream_t st) {
  embed_rows_kernel<<<dim3(T_alloc, B), 128, 0, st>>>(xs, T_in, lens, ctx, n_ctx, out, T_alloc);
  CV2_LAUNCH_CHECK();
}
__global__ void stream_lens_kernel(const int* __restrict__ prompt_len, const int* __restrict__ token_len, int* __restrict__ t_done,
                                   int* __restrict__ len_ctx, int* __restrict__ len_enc, int* __restrict__ len_mel,
                                   int* __restrict__ t_lo, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const bool active = token_len[b] > 0;
  const int lc = active ? prompt_len[b] + token_len[b] : 0;
  const int le = max(lc - 3, 0);
  const int lm = 2 * le;
  const int lo = active ? (t_done[b] / 128) * 128 : 0;
  len_ctx[b] = lc;
  len_enc[b] = le;
  len_mel[b] = len_mel[B + b] = lm;
  t_lo[b] = t_lo[B + b] = min(lo, lm);
  if (active) t_done[b] = lm;
}
void launch_stream_lens(const int* prompt_len, const int* token_len, int* t_done, int* len_ctx, int* len_enc, int* len_mel, int* t_lo,
                        int B, cudaStream_t st) {
  stream_lens_kernel<<<(B + 127) / 128, 128, 0, st>>>(prompt_len, token_len, t_done, len_ctx, len_enc, len_mel, t_lo, B);
  CV2_LAUNCH_CHECK();
}

__global__ void tail_swap_kernel(__half* __restrict__ buf, int T_alloc, int C8, long long ld, const int* __restrict__ lens,
                                 const int* __restrict__ t_lo, __half* __restrict__ cache) {
  const int nb = T_alloc / 128;
  const int b = blockIdx.x + 1, s = blockIdx.y;
  const int edge = b * 128, lo = t_lo[s], len = lens[s];
  const bool restore = lo == edge && len > lo;
  const bool save = edge > lo && edge <= len;
  if (!restore && !save) return;
  uint4* c = reinterpret_cast<uint4*>(cache + ((long long)(s * nb + b) * 2) * 512);
  for (int i = threadIdx.x; i < 2 * C8; i += blockDim.x) {
    const int r = i / C8, k = i - r * C8;
    uint4* g = reinterpret_cast<uint4*>(buf + ((long long)s * T_alloc + edge - 2 + r) * ld) + k;
    if (restore) *g = c[r * 64 + k];
    else c[r * 64 + k] = *g;
  }
}
void launch_tail_swap(__half* buf, int S, int T_alloc, int C, long long ld, const int* lens, const int* t_lo, __half* cache,
                      cudaStream_t st) {
  if (T_alloc / 128 < 2) return;
  tail_swap_kernel<<<dim3(T_alloc / 128 - 1, S), 64, 0, st>>>(buf, T_alloc, C / 8, ld, lens, t_lo, cache);
  CV2_LAUNCH_CHECK();
}

__global__ void absmax16_kernel(const __half* __restrict__ p, long long rows, int cols, long long ld, unsigned* __restrict__ slot) {
  float m = 0.f;
  const long long n = rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols;
    const float v = fabsf(__half2float(p[r * ld + (i - r * cols)]));
    m = fmaxf(m, v == v ? v : INFINITY);
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(slot, __float_as_uint(m));
}
void launch_absmax16(const __half* p, long long rows, int cols, long long ld, unsigned* slot, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return;
  const long long n = rows * cols;
  const unsigned grid = (unsigned)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
  absmax16_kernel<<<grid, 256, 0, st>>>(p, rows, cols, ld, slot);
  CV2_LAUNCH_CHECK();
}
__global__ void lens_clamp_kernel(const int* a, int hi, int add, int* out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = min(a[i], hi) + add;
}
void launch_lens_clamp(const int* a, int hi, int add, int* out, int n, cudaStream_t st) {
  lens_clamp_kernel<<<(n + 127) / 128, 128, 0, st>>>(a, hi, add, out, n);
  CV2_LAUNCH_CHECK();
}

// ---- LayerNorm over 512 channels, fp32 in -> 16-bit out (warp per row) ---------------------------------
__global__ void layernorm512_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ bta,
                                    float eps, __half* __restrict__ out16, float* __restrict__ out32, const int* __restrict__ lens,
                                    int len_all, int T_alloc, long long rows) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int s = (int)(row / T_alloc), t = (int)(row % T_alloc);
  const int len = lens ? lens[s] : len_all;
  const float4* xr = reinterpret_cast<const float4*>(x + row * 512);
  float4 v[4];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    v[i] = xr[lane + 32 * i];
    sum += v[i].x + v[i].y + v[i].z + v[i].w;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum * (1.f / 512.f);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    sq += a * a + b * b + c * c + d * d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq * (1.f / 512.f) + eps);
  const bool valid = t < len;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int c = (lane + 32 * i) * 4;
    const float4 gg = *reinterpret_cast<const float4*>(g + c);
    const float4 bb = *reinterpret_cast<const float4*>(bta + c);
    float4 y;
    y.x = (v[i].x - mean) * rstd * gg.x + bb.x;
    y.y = (v[i].y - mean) * rstd * gg.y + bb.y;
    y.z = (v[i].z - mean) * rstd * gg.z + bb.z;
    y.w = (v[i].w - mean) * rstd * gg.w + bb.w;
    if (out32) *reinterpret_cast<float4*>(out32 + row * 512 + c) = y;
    if (out16) {
      if (!valid) y = make_float4(0.f, 0.f, 0.f, 0.f);
      __half2 h0 = __floats2half2_rn(y.x, y.y), h1 = __floats2half2_rn(y.z, y.w);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&h0);
      u.y = *reinterpret_cast<uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(out16 + row * 512 + c) = u;
    }
  }
}
void launch_layernorm512(const float* x, const float* g, const float* b, float eps, __half* out16, float* out32, const int* lens,
                         int len_all, int S, int T_alloc, cudaStream_t st) {
  const long long rows = (long long)S * T_alloc;
  layernorm512_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(x, g, b, eps, out16, out32, lens, len_all, T_alloc, rows);
  CV2_LAUNCH_CHECK();
}

// ---- nearest x2 upsample along time of a 16-bit [S, T_in_alloc, C] tensor (Upsample1D, upsample_encoder.py:59-63)
__global__ void repeat2_kernel(const __half* __restrict__ in, __half* __restrict__ out, const int* __restrict__ lens, int T_in_alloc,
                               int T_out_alloc, int C8) {
  const int t = blockIdx.x, s = blockIdx.y;  // output row
  const int len2 = lens[s] * 2;
  const uint4* src = reinterpret_cast<const uint4*>(in + ((long long)s * T_in_alloc + (t >> 1)) * C8 * 8);
  uint4* dst = reinterpret_cast<uint4*>(out + ((long long)s * T_out_alloc + t) * C8 * 8);
  for (int c = threadIdx.x; c < C8; c += blockDim.x) dst[c] = (t < len2) ? src[c] : make_uint4(0, 0, 0, 0);
}
void launch_repeat2(const __half* in, __half* out, const int* lens, int S, int T_in_alloc, int T_out_alloc, int C,
                    cudaStream_t st) {
  repeat2_kernel<<<dim3(T_out_alloc, S), 64, 0, st>>>(in, out, lens, T_in_alloc, T_out_alloc, C / 8);
  CV2_LAUNCH_CHECK();
}

// ---- relative positional table by relative position (embedding.py:228-253): row (rel + Tmax - 1) -------
__global__ void pos_table_kernel(__half* __restrict__ out, int Tmax) {
  const int r = blockIdx.x;
  const float rel = (float)(r - (Tmax - 1));
  for (int c = threadIdx.x; c < 256; c += blockDim.x) {
    const float div = expf((float)(2 * c) * (float)(-(9.210340371976184 / 512.0)));  // -ln(10000)/d_model, rounded once like torch
    const float a = rel * div;
    out[(long long)r * 512 + 2 * c] = __float2half_rn(sinf(a));
    out[(long long)r * 512 + 2 * c + 1] = __float2half_rn(cosf(a));
  }
}
void launch_pos_table(__half* out, int Tmax, cudaStream_t st) {
  pos_table_kernel<<<2 * Tmax - 1, 128, 0, st>>>(out, Tmax);
  CV2_LAUNCH_CHECK();
}

// ---- x-vector: L2-normalise (F.normalize eps 1e-12) then Linear 192 -> 80 (flow.py:248-249) --------------
__global__ void spk_affine_kernel(const float* __restrict__ emb, const float* __restrict__ w, const float* __restrict__ bias,
                                  float* __restrict__ out) {
  const int b = blockIdx.x;
  __shared__ float e[192];
  __shared__ float nrm;
  const float* x = emb + (long long)b * 192;
  if (threadIdx.x < 32) {
    float s = 0.f;
    for (int i = threadIdx.x; i < 192; i += 32) s += x[i] * x[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) nrm = fmaxf(sqrtf(s), 1e-12f);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 192; i += blockDim.x) e[i] = x[i] / nrm;
  __syncthreads();
  for (int o = threadIdx.x; o < 80; o += blockDim.x) {
    float acc = bias[o];
    for (int i = 0; i < 192; i++) acc += w[o * 192 + i] * e[i];
    out[(long long)b * 80 + o] = acc;
  }
}
void launch_spk_affine(const float* emb, const float* w, const float* bias, float* out, int B, cudaStream_t st) {
  spk_affine_kernel<<<B, 96, 0, st>>>(emb, w, bias, out);
  CV2_LAUNCH_CHECK();
}

// ---- time embedding: sinusoid(320, x1000) -> Linear 320->1024 -> SiLU -> Linear 1024->1024, then per resnet
//      Mish -> Linear 1024->256 (matcha decoder.py:14-29, 73-117, 49).  fp32 GEMVs, warp per output.
__global__ void time_mlp1_kernel(const float* __restrict__ t, const float* __restrict__ w1, const float* __restrict__ b1,
                                 float* __restrict__ h1) {  // grid (1024/8, nt)
  const int it = blockIdx.y;
  const int o = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const float tv = t[it];
  float acc = 0.f;
  for (int i = lane; i < 320; i += 32) {
    const int k = i < 160 ? i : i - 160;
    const float f = expf((float)k * (float)(-(9.210340371976184 / 159.0)));
    const float a = 1000.f * tv * f;
    const float e = i < 160 ? sinf(a) : cosf(a);
    acc += w1[(long long)o * 320 + i] * e;
  }
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if (lane == 0) h1[(long long)it * 1024 + o] = silu_f(acc + b1[o]);
}
__global__ void gemv_rows_kernel(const float* __restrict__ x, int K, const float* __restrict__ w, const float* __restrict__ b,
                                 float* __restrict__ y, int N, int mish_in) {  // grid (N/8, rows)
  const int row = blockIdx.y;
  const int o = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (o >= N) return;
  const float* xr = x + (long long)row * K;
  float acc = 0.f;
  for (int i = lane; i < K; i += 32) {
    float xv = xr[i];
    if (mish_in) xv = mish_f(xv);
    acc += w[(long long)o * K + i] * xv;
  }
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if (lane == 0) y[(long long)row * N + o] = acc + b[o];
}
void launch_time_mlp(const float* t, int nt, const float* w1, const float* b1, const float* w2, const float* b2, float* h1,
                     float* temb, cudaStream_t st) {
  time_mlp1_kernel<<<dim3(128, nt), 256, 0, st>>>(t, w1, b1, h1);
  CV2_LAUNCH_CHECK();
  gemv_rows_kernel<<<dim3(128, nt), 256, 0, st>>>(h1, 1024, w2, b2, temb, 1024, 0);
  CV2_LAUNCH_CHECK();
}
void launch_resnet_time_proj(const float* temb, int nt, const float* w, const float* b, float* out, cudaStream_t st) {
  gemv_rows_kernel<<<dim3(32, nt), 256, 0, st>>>(temb, 1024, w, b, out, 256, 1);
  CV2_LAUNCH_CHECK();
}

// ---- estimator input packing: [x | mu | spks | cond] (320 ch) for the conditional row, [x | 0 | 0 | 0] for the
//      unconditional row (flow_matching.py:94-107, decoder.py:425-431).  cond part written once per solve.
__global__ void pack_cond_kernel(const float* __restrict__ mu, const float* __restrict__ spks, const float* __restrict__ cond,
                                 __half* __restrict__ xin, const int* __restrict__ lens, int B, int T_alloc) {
  const int t = blockIdx.x, b = blockIdx.y;
  const bool valid = t < lens[b];
  __half* oc = xin + ((long long)b * T_alloc + t) * 320;
  __half* ou = xin + ((long long)(B + b) * T_alloc + t) * 320;
  const long long r = (long long)b * T_alloc + t;
  for (int c = threadIdx.x; c < 240; c += blockDim.x) {
    float v = 0.f;
    if (valid) v = c < 80 ? mu[r * 80 + c] : (c < 160 ? spks[b * 80 + c - 80] : cond[r * 80 + c - 160]);
    oc[80 + c] = __float2half_rn(v);
    ou[80 + c] = __float2half_rn(0.f);
  }
}
void launch_pack_cond(const float* mu, const float* spks, const float* cond, __half* xin, const int* lens, int B, int T_alloc,
                      cudaStream_t st) {
  pack_cond_kernel<<<dim3(T_alloc, B), 128, 0, st>>>(mu, spks, cond, xin, lens, B, T_alloc);
  CV2_LAUNCH_CHECK();
}

// ---- CFG combine + Euler update (flow_matching.py:115-121) + re-pack of x into both estimator rows ------
//   step < 0 : x = z (initial noise, rand_noise[:, :, :T] sliced by frame index), no update
//   else     : x += dt * ((1 + cfg) * v[b] - cfg * v[B + b])
__global__ void euler_pack_kernel(float* __restrict__ x, const float* __restrict__ v, const float* __restrict__ noise,
                                  int noise_stride, __half* __restrict__ xin, const int* __restrict__ lens, int B, int T_alloc,
                                  float dt, float cfg, int init) {
  const int t = blockIdx.x, b = blockIdx.y;
  const int c = threadIdx.x;  // 80 threads (rounded to 96)
  if (c >= 80) return;
  const long long r = (long long)b * T_alloc + t;
  const bool valid = t < lens[b];
  float xv;
  if (init) {
    xv = valid ? noise[(long long)c * noise_stride + t] : 0.f;
  } else {
    const float vc = v[r * 80 + c];
    const float vu = v[((long long)(B + b) * T_alloc + t) * 80 + c];
    xv = x[r * 80 + c] + dt * ((1.f + cfg) * vc - cfg * vu);
    if (!valid) xv = 0.f;
  }
  x[r * 80 + c] = xv;
  const __half hv = __float2half_rn(xv);
  xin[r * 320 + c] = hv;
  xin[((long long)(B + b) * T_alloc + t) * 320 + c] = hv;
}
void launch_euler_pack(float* x, const float* v, const float* noise, int noise_stride, __half* xin, const int* lens, int B,
                       int T_alloc, float dt, float cfg, int init, cudaStream_t st) {
  euler_pack_kernel<<<dim3(T_alloc, B), 96, 0, st>>>(x, v, noise, noise_stride, xin, lens, B, T_alloc, dt, cfg, init);
  CV2_LAUNCH_CHECK();
}

// fp32 channels-last [rows, C] -> two-term 16-bit split [rows, ld]: hi = fp16(x) at columns [0, C), lo = fp16(x - hi) at
// [lo_off, lo_off + C); the columns in between stay zero (K padding of the split-precision GEMM)
__global__ void split16_kernel(const float* __restrict__ x, int C, __half* __restrict__ out, int ld, int lo_off, long long rows) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * C) return;
  const long long r = i / C;
  const int c = (int)(i - r * C);
  const float v = x[i];
  const __half h = __float2half_rn(v);
  out[r * ld + c] = h;
  out[r * ld + lo_off + c] = __float2half_rn(v - __half2float(h));
}
void launch_split16(const float* x, int C, __half* out, int ld, int lo_off, long long rows, cudaStream_t st) {
  split16_kernel<<<(unsigned)((rows * C + 255) / 256), 256, 0, st>>>(x, C, out, ld, lo_off, rows);
  CV2_LAUNCH_CHECK();
}

// ---- generic layout movers ------------------------------------------------------------------------------
// fp32 NCT [B, C, T] (reference layout) -> fp32 / 16-bit channels-last [B, T_alloc, ldc]
__global__ void nct_to_ntc_kernel(const float* __restrict__ in, long long in_bstride, int in_T, float* __restrict__ out32,
                                  __half* __restrict__ out16, const int* __restrict__ lens, int len_all, int T_alloc, int C, int ldc,
                                  int col_off, int t_src_off) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int len = lens ? lens[b] : len_all;
  {
    const int c = c0 + threadIdx.y, t = t0 + threadIdx.x;
    float v = 0.f;
    if (c < C && t < len && t + t_src_off < in_T) v = in[(long long)b * in_bstride + (long long)c * in_T + t + t_src_off];
    tile[threadIdx.y][threadIdx.x] = v;
  }
  __syncthreads();
  const int t = t0 + threadIdx.y, c = c0 + threadIdx.x;
  if (t < T_alloc && c < C) {
    const float v = tile[threadIdx.x][threadIdx.y];
    const long long o = ((long long)b * T_alloc + t) * ldc + col_off + c;
    if (out32) out32[o] = v;
    if (out16) out16[o] = __float2half_rn(v);
  }
}
void launch_nct_to_ntc(const float* in, long long in_bstride, int in_T, float* out32, __half* out16, const int* lens, int len_all,
                       int B, int T_alloc, int C, int ldc, int col_off, int t_src_off, cudaStream_t st) {
  dim3 grid((T_alloc + 31) / 32, (C + 31) / 32, B);
  nct_to_ntc_kernel<<<grid, dim3(32, 32), 0, st>>>(in, in_bstride, in_T, out32, out16, lens, len_all, T_alloc, C, ldc, col_off,
                                                   t_src_off);
  CV2_LAUNCH_CHECK();
}
// fp32 channels-last [B, T_alloc, ldc] rows [t_off, t_off + T_out) -> fp32 NCT [B, C, T_out]
__global__ void ntc_to_nct_kernel(const float* __restrict__ in, int T_alloc, int ldc, int t_off, const int* __restrict__ t_offs,
                                  float* __restrict__ out, long long out_bstride, int T_out, int C,
                                  const int* __restrict__ in_lens) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  if (t_offs) t_off = t_offs[b];
  const int lim = in_lens ? in_lens[b] - t_off : T_out;   // rows of `in` beyond in_lens[b] are padding -> zeros
  {
    const int t = t0 + threadIdx.y, c = c0 + threadIdx.x;
    float v = 0.f;
    if (t < lim && c < C) v = in[((long long)b * T_alloc + t + t_off) * ldc + c];
    tile[threadIdx.y][threadIdx.x] = v;
  }
  __syncthreads();
  const int c = c0 + threadIdx.y, t = t0 + threadIdx.x;
  if (c < C && t < T_out) out[(long long)b * out_bstride + (long long)c * T_out + t] = tile[threadIdx.x][threadIdx.y];
}
void launch_ntc_to_nct(const float* in, int T_alloc, int ldc, int t_off, const int* t_offs, float* out, long long out_bstride,
                       int T_out, int C, const int* in_lens, int B, cudaStream_t st) {
  dim3 grid((T_out + 31) / 32, (C + 31) / 32, B);
  ntc_to_nct_kernel<<<grid, dim3(32, 32), 0, st>>>(in, T_alloc, ldc, t_off, t_offs, out, out_bstride, T_out, C, in_lens);
  CV2_LAUNCH_CHECK();
}

// cond[b, t, :] = prompt_feat[b, t, :] for t < prompt_feat_len[b], else 0 (flow.py:268-270); channels-last fp32
__global__ void build_cond_kernel(const float* __restrict__ prompt_feat, long long pf_bstride, const int* __restrict__ pf_len,
                                  float* __restrict__ cond, int T_alloc) {
  const int t = blockIdx.x, b = blockIdx.y, c = threadIdx.x;
  if (c >= 80) return;
  float v = 0.f;
  if (t < pf_len[b]) v = prompt_feat[(long long)b * pf_bstride + (long long)t * 80 + c];
  cond[((long long)b * T_alloc + t) * 80 + c] = v;
}
void launch_build_cond(const float* prompt_feat, long long pf_bstride, const int* pf_len, float* cond, int B, int T_alloc,
                       cudaStream_t st) {
  build_cond_kernel<<<dim3(T_alloc, B), 96, 0, st>>>(prompt_feat, pf_bstride, pf_len, cond, T_alloc);
  CV2_LAUNCH_CHECK();
}

// lens arithmetic on device (no host sync): out[i] = a[i]*mul + (b ? b[i] : 0) + add
__global__ void lens_affine_kernel(const int* a, const int* b, int mul, int add, int* out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] * mul + (b ? b[i] : 0) + add;
}
void launch_lens_affine(const int* a, const int* b, int mul, int add, int* out, int n, cudaStream_t st) {
  lens_affine_kernel<<<(n + 127) / 128, 128, 0, st>>>(a, b, mul, add, out, n);
  CV2_LAUNCH_CHECK();
}

// lens[b] = number of non-zero entries of mask[b, 0, :] (estimator C-ABI takes the reference's float mask)
__global__ void mask_to_lens_kernel(const float* __restrict__ mask, int T, int* __restrict__ lens) {
  const int b = blockIdx.x;
  int cnt = 0;
  for (int t = threadIdx.x; t < T; t += 32) cnt += mask[(long long)b * T + t] != 0.f;
  for (int s = 16; s > 0; s >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, s);
  if (threadIdx.x == 0) lens[b] = cnt;
}
void launch_mask_to_lens(const float* mask, int T, int* lens, int S, cudaStream_t st) {
  mask_to_lens_kernel<<<S, 32, 0, st>>>(mask, T, lens);
  CV2_LAUNCH_CHECK();
}
// xin[s, t, col_off + c] = v[s, c] for valid rows (speaker vector broadcast along time)
__global__ void bcast_rows16_kernel(const float* __restrict__ v, int C, __half* __restrict__ out, int ldc, int col_off,
                                    const int* __restrict__ lens, int T_alloc) {
  const int t = blockIdx.x, s = blockIdx.y;
  const bool valid = t < lens[s];
  for (int c = threadIdx.x; c < C; c += blockDim.x)
    out[((long long)s * T_alloc + t) * ldc + col_off + c] = __float2half_rn(valid ? v[(long long)s * C + c] : 0.f);
}
void launch_bcast_rows16(const float* v, int C, __half* out, int ldc, int col_off, const int* lens, int S, int T_alloc,
                         cudaStream_t st) {
  bcast_rows16_kernel<<<dim3(T_alloc, S), 96, 0, st>>>(v, C, out, ldc, col_off, lens, T_alloc);
  CV2_LAUNCH_CHECK();
}

__global__ void f32_to_f16_kernel(const float* __restrict__ in, __half* __restrict__ out, long long n) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(in + i);
    __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0);
    u.y = *reinterpret_cast<uint32_t*>(&h1);
    *reinterpret_cast<uint2*>(out + i) = u;
  } else {
    for (long long j = i; j < n; j++) out[j] = __float2half_rn(in[j]);
  }
}
void launch_f32_to_f16(const float* in, __half* out, long long n, cudaStream_t st) {
  f32_to_f16_kernel<<<(unsigned)((n / 4 + 255) / 256 + 1), 256, 0, st>>>(in, out, n);
  CV2_LAUNCH_CHECK();
}

}  // namespace cv2
