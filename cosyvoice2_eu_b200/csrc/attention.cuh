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
  int reverse_seq;   // dispatch sequences S-1 .. 0 (longest first for an ascending length-sorted batch)
  long long* trace;  // measurement aid (profiles/attn_trace.py): one mid-grid CTA logs (clock64 << 8 | event) for its first softmax
                     // warp [0, 2048) and its MMA warp [2048, 4096); null in production
  const int* lo;     // optional [S]: query tiles with t0 < floor(lo[s] / 128) * 128 are skipped (incremental streaming)
};
void launch_flash_attn(const AttnParams& p, cudaStream_t stream);

// Encoder: Transformer-XL relative-position attention on tcgen05 (encoder_attention.cu).
struct RelAttnParams {
  const __half* qu;     // [S, 8, T_alloc, 64]  (q + bias + pos_bias_u) / 8
  const __half* qv;     // [S, 8, T_alloc, 64]  (q + bias + pos_bias_v) / 8
  const __half* k;      // [S, 8, T_alloc, 64]
  const __half* vt;     // [S, 8, 64, T_alloc]
  const __half* pos;    // [R_alloc, 512] = linear_pos(PE), row (rel + Tmax - 1) <-> relative position rel = i - j
  __half* out;          // [S, T_alloc, 512]
  const int* lens;
  int len_all;
  int S, T_alloc, Tmax, R_alloc;
  int chunk;
  int halo;
};
void launch_rel_attn(const RelAttnParams& p, cudaStream_t stream);

}  // namespace cv2
