// Fused FF1 -> GELU -> FF2 -> residual -> LayerNorm/plain emit of one estimator transformer block (ffn_fused.cu).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include "gemm_tap.cuh"

namespace cv2 {

struct FfnParams {
  int S, T_alloc;
  const int* lens;          // [S] valid rows (null: len_all)
  int len_all;
  int halo;
  const int* tile_list;     // optional compact (s, t0) list + count (device)
  const int* tile_count;
  const float* b1;          // [1024]
  const float* b2;          // [256]
  float* x32;               // [S*T_alloc, 256] residual stream, updated in place
  Emit emit_ln;             // ptr != null: LayerNorm(x; a = gamma, b = beta, f = eps) -> 16-bit
  Emit emit_plain[2];       // ptr != null: masked x -> 16-bit (ld / col_off honoured)
  // chained attention out-projection (2-SM kernel only): the tile first computes x <- x + Wo att + bo into the (still free)
  // FF1 accumulators, writes x back once and builds h = LayerNorm3(x) as the 128B-swizzled 16-bit H tile directly in shared
  // memory -- no separate out-proj launch, no 16-bit h round trip through HBM (transformer.py:274, 291-314)
  const float* bo;          // [256] out-proj bias; null: H comes from global memory (tmH) as before
  const float* ln3_g;       // [256] LayerNorm3 gamma / beta / eps
  const float* ln3_b;
  float ln3_eps;
  // hidden split (chained 2-SM kernel, small launches only): a tile pair is worked on by `hsplit` CTA pairs, each taking 8 / hsplit
  // of the eight hidden chunks.  Every pair recomputes the out-projection and LayerNorm3 (it needs the whole H tile), pair 0 stores
  // x' = x + Wo att + bo to `xprime`, every pair stores its partial FF2 sum to its slab, and ffn_reduce adds them up, applies b2,
  // updates x32 and emits the next pre-norm.  One launch more, but the launch's critical path -- ONE unit -- shrinks from eight
  // serial GELU chunks to two: this is what a batch-1 call (12 row tiles for 148 SMs) is bound by.
  int hsplit;               // 0 / 1: off
  float* slabs;             // [hsplit][S * T_alloc][256] fp32
  float* xprime;            // [S * T_alloc][256] fp32
  long long* trace;         // debug: CTA 0 logs (clock64 << 8 | event code) for its MMA thread [0, 4096) and first epilogue warp
                            // [4096, 8192) (profiles/ffn_trace.py); null in production
};
void ffn_set_trace(long long* dev_buf);   // the next launches log into dev_buf (null: off)

// tmH: 3-D {256, T_alloc, S} box {64,128,1};  tmW1: 2-D {256, 1024} box {64,128};  tmW2: 2-D {1024, 256} box {64,128}
void launch_ffn_fused(const CUtensorMap& tmH, const CUtensorMap& tmW1, const CUtensorMap& tmW2, const FfnParams& p,
                      cudaStream_t stream);

// 2-SM form (ffn_fused2.cu): cluster of two CTAs, tcgen05.mma.cta_group::2.  tmW1h: {256, 1024} box {64, 64};
// tmW2h: {1024, 256} box {64, 128}.  Requires the compact tile list.
void launch_ffn_fused2(const CUtensorMap& tmH, const CUtensorMap& tmW1h, const CUtensorMap& tmW2h, const FfnParams& p,
                       cudaStream_t stream);
// chained form (p.bo != null): tmATT: 3-D {512, T_alloc, S} box {64,128,1} over the attention output; tmWoh: {512, 256} box {64, 128}
void launch_ffn_fused2_chain(const CUtensorMap& tmATT, const CUtensorMap& tmWoh, const CUtensorMap& tmW1h, const CUtensorMap& tmW2h,
                             const FfnParams& p, cudaStream_t stream);
// p.hsplit == 4: the same kernel with the hidden dimension of every tile pair divided among four CTA pairs, followed by the reduction
void launch_ffn_fused2_chain_split(const CUtensorMap& tmATT, const CUtensorMap& tmWoh, const CUtensorMap& tmW1h, const CUtensorMap& tmW2h,
                                   const FfnParams& p, cudaStream_t stream);
int ffn_fused2_max_pairs();   // CTA pairs that are resident at once (after the first launch on this device; 0 before)

}  // namespace cv2
