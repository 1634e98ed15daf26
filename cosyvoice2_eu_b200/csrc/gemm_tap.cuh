// "Tap GEMM": the one tensor-core kernel family behind every dense contraction of the path
// (linear layers, k-tap (dilated / causal / same-pad) conv1d as implicit GEMM, transposed conv
// as phase-concatenated GEMM).
//
//   acc[s, t, n] = sum_{tap j} sum_{k < Kc} A[s, t + off_j, k] * W[n, j*Kc_pad + k]
//
// A: 16-bit activations, channels-last [S, T_alloc, Kc] (TMA 3-D map, OOB rows/cols read as 0, which
//    is exactly the conv's zero padding at the tensor edges);  W: 16-bit [N, ntaps*Kc_pad] (TMA 2-D map).
// One CTA = one 128 x BN output tile: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (+TMEM owner),
// warps 2-5 = epilogue (thread-per-row over the fp32 accumulator in TMEM).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace cv2 {

enum Act { ACT_NONE = 0, ACT_MISH = 1, ACT_GELU = 2, ACT_SILU = 3, ACT_ELU = 4, ACT_LRELU = 5, ACT_SNAKE = 6 };
enum EmitKind { EMIT_NONE = 0, EMIT_PLAIN = 1, EMIT_SNAKE = 2, EMIT_LN = 3, EMIT_LRELU = 4,
                EMIT_LO = 5 };   // v - fp16(v): the low half of a two-term 16-bit split (with EMIT_PLAIN as the high half)

struct Emit {
  __half* ptr;         // 16-bit output [S, T_alloc, ld]; rows >= len are written as 0
  long long ld;        // elements per row
  int col_off;         // first column inside the row
  int kind;            // EmitKind
  const float* a;      // snake: alpha[N]; LN: gamma[N]
  const float* b;      // LN: beta[N]
  float f;             // LN: eps; lrelu: slope
  float scale;         // multiplied into the emitted value (e.g. sqrt(d) of the encoder embed)
};

struct GemmParams {
  // problem
  int S, T_alloc, N;
  int kb_per_tap;      // Kc_pad / 64
  int ntaps;
  int tap_off[16];     // row offset of each tap
  int tap_seq[16];     // sequence offset of each tap (A row block read from sequence s + tap_seq[j]; CFG pair = two "taps")
  int S_map;           // sequences covered by the A tensor map when it differs from S (0: S)
  const int* lens;     // [S] valid rows per sequence (nullptr: len_all)
  int len_all;
  int halo;            // a tile is computed iff t0 < len + halo
  const int* tile_list;   // optional compact list of active (s, t0) pairs (device), balanced round-robin over CTAs
  const int* tile_count;  // number of pairs in tile_list (device)
  long long* trace;              // clock64 timeline of CTA 0 (measurement aid)
  int dbg_skip_epi;              // measurement aid: epilogue drains the accumulator without computing or storing
  const CUtensorMap* tmB_half;   // host pointer: weight map with a {64, BN/2} box -> 2-CTA cluster with TMA multicast of the
                                 // weight operand (BN = 256 and a compact tile list only); null: one CTA per tile
  // "conv mode" (k-tap convs of the vocoder's narrow stages, BN = 64 / 128, N <= BN): instead of one [128 x 64] A tile PER TAP
  // (the same rows shifted by the tap offset: 11 x 16 KB for a k = 11 conv) ONE [192 x 64] box per k-block carries the rows of
  // all taps and every tap reads it through a row-shifted shared-memory descriptor; the whole weight matrix (<= 96 KB) is
  // loaded once per CTA and stays resident.  Per-tile TMA traffic: 24 KB per k-block instead of (16 + BN / 8) KB per tap.
  const CUtensorMap* tmA_halo;   // host pointer: A map with a {64, 192, 1} box; null: normal mode
  int conv_min_off;              // smallest tap offset (first row of the box = t0 + conv_min_off)
  // epilogue program: v = acc * acc_scale + bias -> LN -> act -> + rowvec[s] -> (mask) -> + res + res2 -> *scale (+= out32) -> store / emit
  float acc_scale;     // 0 / 1: none; else the raw accumulator is multiplied by it before the bias (weights stored pre-scaled to
                       // keep a 16-bit low half out of the subnormal range); generic epilogue only, not with ln
  const float* bias;   // [N] or null
  int ln;
  const float* ln_g;
  const float* ln_b;
  float ln_eps;
  int act;
  float act_f;         // lrelu slope
  const float* act_a;  // snake alpha[N]
  const float* rowvec; // [S, rowvec_ld] or null
  int rowvec_ld;
  int mask_pre_res;    // zero rows >= len before the residual add
  const float* res;    // fp32 [S*T_alloc, res_ld] or null
  long long res_ld;
  const float* res2;
  long long res2_ld;
  float out_scale;
  float* out32;        // fp32 output or null
  long long out32_ld;
  int out32_accum;     // out32 += v*scale (emits then see the total)
  // flat mode (transposed conv): element index e = t*out32_ld + col + flat_off inside a sequence slab of
  // flat_seq_elems; stored iff flat_lo <= e < len*flat_hi_per_len + flat_hi_add
  int flat;
  long long flat_off, flat_lo, flat_hi_per_len, flat_hi_add, flat_seq_elems;
  Emit emit[3];
  // attention split (N = 3*heads*64): q*q_scale -> [S,H,T_alloc,64], k -> same, v -> transposed [S,H,64,T_alloc];
  // with q2 != null N = 4*heads*64 and the column blocks are q | q2 | k | v (both q blocks scaled)
  __half* q;
  __half* q2;
  __half* k;
  __half* vt;
  int heads;
  float q_scale;
};

// Launch; tmA = 3-D map {Kc, T_alloc, S} box {64,128,1}; tmB = 2-D map {Ktot_pad, N} box {64, BN}.
// bn in {64, 128, 256}.
int gemm_tap_spec(int bn, const GemmParams& p);
// can this launch run in conv mode (see GemmParams::tmA_halo)?  bn / N / taps / k-blocks as launch_gemm_tap gets them
bool gemm_tap_conv_eligible(int bn, const GemmParams& p);
static constexpr int kConvHaloRows = 192;
// list: [2 * S * (T_alloc/128)] ints, count: 1 int (device)
// lo (optional, [S]): only tiles with t0 >= floor(lo[s] / 128) * 128 are listed (incremental streaming: earlier rows are final)
void launch_build_tile_list(const int* lens, int S, int T_alloc, int halo, int* list, int* count, cudaStream_t stream,
                            const int* lo = nullptr);
void launch_gemm_tap(int bn, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream);

}  // namespace cv2
