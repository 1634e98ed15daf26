// Attention kernels of the path (see attention.cu / encoder_attention.cu).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace cv2 {

struct AttnParams {
  const __half* q;   // [S, H, T_alloc, 64], pre-scaled by 1/sqrt(64)
  const __half* k;   // [S, H, T_alloc, 64]
  const __half* vt;  // [S, H, 64, T_alloc]
  __half* out;       // [S, T_alloc, H*64]
  const int* lens;   // [S] or null
  int len_all;
  int S, heads, T_alloc;
  int chunk;         // 0: every valid key visible; >0: key j visible to query i iff j < (i/chunk+1)*chunk
  int halo;
};
void launch_flash_attn(const AttnParams& p, cudaStream_t stream);

// Encoder: Transformer-XL relative-position attention, fp32 SIMT (1.5 % of the path's FLOPs).
struct RelAttnParams {
  const float* qkv;     // [S, T_alloc, 1536] (q | k | v, bias already added)
  const float* pos;     // [2*Tmax-1, 512] = linear_pos(PE), row (rel + Tmax - 1) <-> relative position rel = i - j
  const float* bias_u;  // [8, 64]
  const float* bias_v;  // [8, 64]
  __half* out;          // [S, T_alloc, 512]
  const int* lens;
  int len_all;
  int S, T_alloc, Tmax;
  int chunk;
};
void launch_rel_attn(const RelAttnParams& p, cudaStream_t stream);

}  // namespace cv2
