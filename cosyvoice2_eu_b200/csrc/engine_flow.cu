// flow.inference on the device: UpsampleConformerEncoder -> encoder_proj -> 10-step CFG Euler solve of the
// CausalConditionalDecoder estimator (reference: cosyvoice/flow/flow.py:235-283, flow_matching.py:71-123,
// decoder.py:405-494, transformer/upsample_encoder.py:243-306).  Batched over utterances with per-utterance
// lengths (B=1 semantics per utterance: every kernel treats a sequence end like a tensor edge).
#include <math.h>

#include "attention.cuh"
#include "engine.h"
#include "flow_kernels.cuh"

namespace cv2 {

static const int kHalo = 32;
// The estimator is causal end to end (CausalConv1d taps -2..0, attention masked by length), so no row >= len is ever read by a
// valid row: its tiles need no halo.  (The encoder looks 3 tokens ahead and the vocoder's same-pad convs up to 25 frames.)
static const int kEstHalo = 0;

static GemmParams base_params(const int* lens, int halo = kHalo) {
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.lens = lens;
  p.halo = halo;
  p.out_scale = 1.f;
  return p;
}
static Emit emit_plain(__half* ptr, long long ld, int col_off = 0) {
  Emit e;
  memset(&e, 0, sizeof(e));
  e.ptr = ptr; e.ld = ld; e.col_off = col_off; e.kind = EMIT_PLAIN; e.scale = 1.f;
  return e;
}
static Emit emit_ln(__half* ptr, long long ld, const LN& ln, float eps) {
  Emit e;
  memset(&e, 0, sizeof(e));
  e.ptr = ptr; e.ld = ld; e.kind = EMIT_LN; e.a = ln.g; e.b = ln.b; e.f = eps; e.scale = 1.f;
  return e;
}

// ---------------------------------------------------------------------------------------------------------
// Estimator core: xin16 [S, T_alloc, 320] -> v32 [S, T_alloc, 80].  temb_res: per-resnet time vectors
// [14][nt][256]; row `trow` (stride trow_ld: 0 = shared by all sequences, 256 = one per sequence).
// ---------------------------------------------------------------------------------------------------------
static constexpr int kFfnSplitMaxTiles = 36;   // 18 tile pairs x 4 hidden splits <= 74 resident CTA pairs
struct EstBuffers {
  float *X32, *R32, *V32;
  float* SP32;   // small launches: x' + 4 partial FF2 sums of the hidden-split FFN (null otherwise)
  __half *H16, *Q16, *K16, *VT16, *ATT16, *F16, *C16, *CAT16, *M16, *N16;
};
static EstBuffers est_alloc(Arena& ws, int S, int T) {
  EstBuffers b;
  const size_t rows = (size_t)S * T;
  b.X32 = ws.get<float>(rows * 256);
  b.R32 = ws.get<float>(rows * 256);
  b.V32 = ws.get<float>(rows * 80);
  b.H16 = ws.get<__half>(rows * 256);
  b.Q16 = ws.get<__half>(rows * 512);
  b.K16 = ws.get<__half>(rows * 512);
  b.VT16 = ws.get<__half>(rows * 512);
  b.ATT16 = ws.get<__half>(rows * 512);
  b.F16 = ws.get<__half>(rows * 1024);
  b.C16 = ws.get<__half>(rows * 256);
  b.CAT16 = ws.get<__half>(rows * 512);
  b.M16 = ws.get<__half>(rows * 256);
  b.N16 = ws.get<__half>(rows * 256);
  b.SP32 = (long long)S * (T / 128) <= kFfnSplitMaxTiles ? ws.get<float>(rows * 256 * 5) : nullptr;
  return b;
}

// Classifier-free-guidance combine + Euler update fused into final_proj (flow_matching.py:104-121): the projection is
// linear, so (1+cfg) W h_cond - cfg W h_uncond is ONE GEMM over the K-concatenated pair of rows (weights [(1+cfg) W | -cfg W],
// the second "tap" reads sequence s + B), and the epilogue does x += dt * (v + b) * mask in fp32 and re-emits x as the
// 16-bit input channels of both CFG rows for the next step.
struct EulerFuse {
  float* x32;        // [B, T_alloc, 80] Euler state
  __half* xin16;     // [2B, T_alloc, 320]: channels 0..79 of both rows are rewritten
  float dt;
  int B;
};

// Incremental streaming: the estimator call of Euler step `step` of a chunk that only computes the row tiles from t_lo on.
struct EstStream {
  const StreamState* ss;
  int step;
  const int* t_lo;     // [S] device
};

static void est_core(Engine& e, cudaStream_t st, const EstBuffers& b, const __half* xin16, const int* lens, int S, int T,
                     const float* temb_res, int nt, int trow, int trow_ld, int streaming, bool dry,
                     const EulerFuse* ef = nullptr, const EstStream* es = nullptr) {
  static const int causal3[3] = {-2, -1, 0};
  static const int tap1[1] = {0};
  // a causal k=3 conv reads two rows in front of its first tile: with a stream state they come from / go to the tail cache
  auto tails = [&](int conv, const __half* buf, int C, long long ld) {
    if (!es) return;
    e.launches++;
    if (dry) return;
    launch_tail_swap(const_cast<__half*>(buf), S, T, C, ld, lens, es->t_lo,
                     es->ss->tails + ((size_t)es->step * kEstConvs + conv) * es->ss->tail_block(), st);
  };
  const __half* cur = xin16;
  int cur_c = 320;
  long long cur_ld = 320;
  for (int r = 0; r < 14; r++) {
    const std::string rp = "est.res." + std::to_string(r);
    if (r == 13) {  // up block: cat[x, skip]
      cur = b.CAT16; cur_c = 512; cur_ld = 512;
    }
    tails(2 * r, cur, cur_c, cur_ld);
    // res_conv (1x1) -> R32
    {
      GemmParams p = base_params(lens, kEstHalo);
      p.out32 = b.R32; p.out32_ld = 256;
      e.gemm(st, cur, S, T, cur_c, cur_ld, e.W(rp + ".res"), 256, 1, tap1, p, dry);
    }
    // block1: causal conv k3 -> LN -> Mish -> + time vector -> C16
    {
      GemmParams p = base_params(lens, kEstHalo);
      LN ln = e.ln(rp + ".ln1");
      p.ln = 1; p.ln_g = ln.g; p.ln_b = ln.b; p.ln_eps = 1e-5f;
      p.act = ACT_MISH;
      p.rowvec = temb_res + ((size_t)r * nt + trow) * 256; p.rowvec_ld = trow_ld;
      p.emit[0] = emit_plain(b.C16, 256);
      e.gemm(st, cur, S, T, cur_c, cur_ld, e.W(rp + ".c1"), 256, 3, causal3, p, dry);
    }
    tails(2 * r + 1, b.C16, 256, 256);
    // block2: conv -> LN -> Mish -> mask -> + R32 -> X32 ; emit LN(norm1 of tfm 0) -> H16
    {
      GemmParams p = base_params(lens, kEstHalo);
      LN ln = e.ln(rp + ".ln2");
      p.ln = 1; p.ln_g = ln.g; p.ln_b = ln.b; p.ln_eps = 1e-5f;
      p.act = ACT_MISH;
      p.mask_pre_res = 1;
      p.res = b.R32; p.res_ld = 256;
      p.out32 = b.X32; p.out32_ld = 256;
      p.emit[0] = emit_ln(b.H16, 256, e.ln("est.tfm." + std::to_string(r) + ".0.ln1"), 1e-5f);
      e.gemm(st, b.C16, S, T, 256, 256, e.W(rp + ".c2"), 256, 3, causal3, p, dry);
    }
    for (int j = 0; j < 4; j++) {
      const std::string tp = "est.tfm." + std::to_string(r) + "." + std::to_string(j);
      __half *k16 = b.K16, *vt16 = b.VT16;
      if (es) {   // k / v^T of this (Euler step, block) live in the stream state: earlier chunks' rows are already there
        const size_t blk = ((size_t)es->step * kEstBlocks + r * 4 + j) * es->ss->kv_block();
        k16 = es->ss->kcache + blk;
        vt16 = es->ss->vcache + blk;
      }
      {  // q,k,v (no bias); q pre-scaled by 1/sqrt(64) through its weights
        GemmParams p = base_params(lens, kEstHalo);
        p.q = b.Q16; p.k = k16; p.vt = vt16; p.heads = 8; p.q_scale = 1.f;   // (1/sqrt(64) is folded into the packed q weights, pack.py)
        e.gemm(st, b.H16, S, T, 256, 256, e.W(tp + ".qkv"), 256, 1, tap1, p, dry);
      }
      {
        AttnParams ap;
        memset(&ap, 0, sizeof(ap));
        ap.q = b.Q16; ap.k = k16; ap.vt = vt16; ap.out = b.ATT16;
        ap.lens = lens; ap.S = S; ap.heads = 8; ap.T_alloc = T; ap.chunk = streaming ? 50 : 0; ap.halo = kEstHalo;
        ap.reverse_seq = 1;
        ap.lo = es ? es->t_lo : nullptr;
        e.launches++;
        if (!dry) {
          e.prof_begin(st, Engine::F_FLASH_ATTN);
          launch_flash_attn(ap, st);
          e.prof_end(st);
          e.range_scan(st, Engine::R_ATTN_OUT, b.ATT16, (long long)S * T, 512, 512);
        }
      }
      if (!e.fuse_ffn) {  // out-proj + bias + residual ; emit LN(norm3)  (with the fused FFN it belongs to e.ffn below)
        GemmParams p = base_params(lens, kEstHalo);
        p.res = b.X32; p.res_ld = 256;
        p.out32 = b.X32; p.out32_ld = 256;
        p.emit[0] = emit_ln(b.H16, 256, e.ln(tp + ".ln3"), 1e-5f);
        e.gemm(st, b.ATT16, S, T, 512, 512, e.W(tp + ".o"), 256, 1, tap1, p, dry);
      }
      if (e.fuse_ffn) {  // FF1 -> GELU -> FF2 -> + residual -> next pre-norm / masked block output, one kernel
        FfnParams fp;
        memset(&fp, 0, sizeof(fp));
        fp.lens = lens; fp.halo = kEstHalo; fp.x32 = b.X32;
        if (b.SP32 && !es) {
          fp.hsplit = 4;
          fp.xprime = b.SP32;
          fp.slabs = b.SP32 + (size_t)S * T * 256;
        }
        if (j < 3) {
          fp.emit_ln = emit_ln(b.H16, 256, e.ln("est.tfm." + std::to_string(r) + "." + std::to_string(j + 1) + ".ln1"), 1e-5f);
        } else if (r == 0) {
          fp.emit_plain[0] = emit_plain(b.CAT16, 512, 256);
          fp.emit_plain[1] = emit_plain(b.M16, 256);
        } else if (r == 12) {
          fp.emit_plain[0] = emit_plain(b.CAT16, 512, 0);
        } else {
          fp.emit_plain[0] = emit_plain(b.M16, 256);
        }
        Engine::OutProj op{b.ATT16, &e.W(tp + ".o"), e.ln(tp + ".ln3"), 1e-5f};
        e.ffn(st, b.H16, S, T, e.W(tp + ".ff1"), e.W(tp + ".ff2"), fp, dry, &op);
        continue;
      }
      {  // FF1 + exact GELU
        GemmParams p = base_params(lens, kEstHalo);
        p.act = ACT_GELU;
        p.emit[0] = emit_plain(b.F16, 1024);
        e.gemm(st, b.H16, S, T, 256, 256, e.W(tp + ".ff1"), 256, 1, tap1, p, dry);
      }
      {  // FF2 + bias + residual ; emit next pre-norm or the (masked) block output
        GemmParams p = base_params(lens, kEstHalo);
        p.res = b.X32; p.res_ld = 256;
        p.out32 = b.X32; p.out32_ld = 256;
        if (j < 3) {
          p.emit[0] = emit_ln(b.H16, 256, e.ln("est.tfm." + std::to_string(r) + "." + std::to_string(j + 1) + ".ln1"), 1e-5f);
        } else if (r == 0) {
          p.emit[0] = emit_plain(b.CAT16, 512, 256);  // skip connection (hiddens.append)
          p.emit[1] = emit_plain(b.M16, 256);         // input of the down "CausalConv1d"
        } else if (r == 12) {
          p.emit[0] = emit_plain(b.CAT16, 512, 0);
        } else {
          p.emit[0] = emit_plain(b.M16, 256);
        }
        e.gemm(st, b.F16, S, T, 1024, 1024, e.W(tp + ".ff2"), 256, 1, tap1, p, dry);
      }
    }
    if (r == 0) {  // down_blocks.0.2 : CausalConv1d(256,256,3) on x*mask
      tails(28, b.M16, 256, 256);
      GemmParams p = base_params(lens, kEstHalo);
      p.emit[0] = emit_plain(b.N16, 256);
      e.gemm(st, b.M16, S, T, 256, 256, e.W("est.down_conv"), 256, 3, causal3, p, dry);
      cur = b.N16; cur_c = 256; cur_ld = 256;
    } else {
      cur = b.M16; cur_c = 256; cur_ld = 256;
    }
  }
  tails(29, b.M16, 256, 256);
  {  // up_blocks.0.2
    GemmParams p = base_params(lens, kEstHalo);
    p.emit[0] = emit_plain(b.N16, 256);
    e.gemm(st, b.M16, S, T, 256, 256, e.W("est.up_conv"), 256, 3, causal3, p, dry);
  }
  tails(30, b.N16, 256, 256);
  {  // final_block
    GemmParams p = base_params(lens, kEstHalo);
    LN ln = e.ln("est.final.ln");
    p.ln = 1; p.ln_g = ln.g; p.ln_b = ln.b; p.ln_eps = 1e-5f;
    p.act = ACT_MISH;
    p.emit[0] = emit_plain(b.M16, 256);
    e.gemm(st, b.N16, S, T, 256, 256, e.W("est.final.c"), 256, 3, causal3, p, dry);
  }
  if (ef) {  // final_proj + CFG combine + Euler update, one launch over the B utterances
    static const int pair[2] = {0, 0};
    GemmParams p = base_params(lens, kEstHalo);
    p.tap_seq[1] = ef->B;
    p.S_map = S;
    p.mask_pre_res = 1;
    p.out_scale = ef->dt;
    p.out32 = ef->x32; p.out32_ld = 80; p.out32_accum = 1;
    p.emit[0] = emit_plain(ef->xin16, 320, 0);
    p.emit[1] = emit_plain(ef->xin16 + (size_t)ef->B * T * 320, 320, 0);
    e.gemm(st, b.M16, ef->B, T, 256, 256, e.W("est.proj_cfg"), 128, 2, pair, p, dry);
    return;
  }
  {  // final_proj (1x1, 256 -> 80), output * mask
    GemmParams p = base_params(lens, kEstHalo);
    p.mask_pre_res = 1;
    p.out32 = b.V32; p.out32_ld = 80;
    e.gemm(st, b.M16, S, T, 256, 256, e.W("est.proj"), 128, 1, tap1, p, dry);
  }
}

// time embedding for `nt` time values (device) -> temb_res [14][nt][256]
static float* est_time(Engine& e, cudaStream_t st, Arena& ws, const float* t_dev, int nt, bool dry) {
  float* h1 = ws.get<float>((size_t)nt * 1024);
  float* temb = ws.get<float>((size_t)nt * 1024);
  float* tres = ws.get<float>((size_t)14 * nt * 256);
  e.launches += 2 + 14;
  if (!dry) {
    launch_time_mlp(t_dev, nt, e.f32("est.time.w1"), e.f32("est.time.b1"), e.f32("est.time.w2"), e.f32("est.time.b2"), h1, temb, st);
    for (int r = 0; r < 14; r++) {
      const std::string rp = "est.res." + std::to_string(r);
      launch_resnet_time_proj(temb, nt, e.f32(rp + ".mlp_w"), e.f32(rp + ".mlp_b"), tres + (size_t)r * nt * 256, st);
    }
  }
  return tres;
}

// ---------------------------------------------------------------------------------------------------------
// Estimator C-ABI entry (reference boundary: ConditionalCFM.forward_estimator, flow_matching.py:125-150; TRT I/O
// contract of cosyvoice/bin/export_onnx.py:89-109): NCT fp32 x/mu/cond [B2,80,T], mask [B2,1,T], t [B2], spks [B2,80].
// ---------------------------------------------------------------------------------------------------------
size_t estimator_forward(Engine& e, cudaStream_t st, const EstArgs& a, Arena& ws) {
  const bool dry = ws.measuring();
  const int S = a.B2, T = round_up(a.T, 128);
  int* lens = ws.get<int>(S);
  e.tile_lists.clear();
  __half* xin = ws.get<__half>((size_t)S * T * 320);
  EstBuffers b = est_alloc(ws, S, T);
  float* tres = est_time(e, st, ws, a.t, S, dry);
  e.launches += 6;
  if (!dry) launch_mask_to_lens(a.mask, a.T, lens, S, st);
  e.make_tile_list(st, ws, lens, S, T, kEstHalo, dry);
  if (!dry) {
    launch_nct_to_ntc(a.x, (long long)80 * a.T, a.T, nullptr, xin, lens, 0, S, T, 80, 320, 0, 0, st);
    launch_nct_to_ntc(a.mu, (long long)80 * a.T, a.T, nullptr, xin, lens, 0, S, T, 80, 320, 80, 0, st);
    launch_bcast_rows16(a.spks, 80, xin, 320, 160, lens, S, T, st);
    launch_nct_to_ntc(a.cond, (long long)80 * a.T, a.T, nullptr, xin, lens, 0, S, T, 80, 320, 240, 0, st);
  }
  est_core(e, st, b, xin, lens, S, T, tres, S, 0, 256, a.streaming, dry);
  if (!dry) launch_ntc_to_nct(b.V32, T, 80, 0, nullptr, a.out, (long long)80 * a.T, a.T, 80, nullptr, S, st);
  return ws.peak;
}

// ---------------------------------------------------------------------------------------------------------
// Encoder layer (ConformerEncoderLayer without macaron / conv module, encoder_layer.py:160-236)
// ---------------------------------------------------------------------------------------------------------
struct EncBuffers {
  float* X32;
  __half *H16, *ATT16, *F16, *PE16, *POS16, *QU16, *QV16, *K16, *VT16;
  int R_alloc;
};

static void enc_layer(Engine& e, cudaStream_t st, const EncBuffers& b, const std::string& lp, const int* lens, int S, int T,
                      int chunk, const LN* next_norm, __half* plain_out, bool dry) {
  static const int tap1[1] = {0};
  {  // (q + u)/8 | (q + v)/8 | k | v with bias -> per-head 16-bit attention operands (pos_bias_u / v folded into the bias)
    GemmParams p = base_params(lens);
    p.q = b.QU16; p.q2 = b.QV16; p.k = b.K16; p.vt = b.VT16; p.heads = 8; p.q_scale = 0.125f;
    e.gemm(st, b.H16, S, T, 512, 512, e.W(lp + ".qkv"), 256, 1, tap1, p, dry);
  }
  {  // linear_pos on the relative-position table (no bias)
    GemmParams p = base_params(nullptr);
    p.len_all = 2 * T - 1;
    p.emit[0] = emit_plain(b.POS16, 512);
    e.gemm(st, b.PE16, 1, b.R_alloc, 512, 512, e.W(lp + ".pos"), 256, 1, tap1, p, dry);
  }
  {
    RelAttnParams ap;
    memset(&ap, 0, sizeof(ap));
    ap.qu = b.QU16; ap.qv = b.QV16; ap.k = b.K16; ap.vt = b.VT16; ap.pos = b.POS16;
    ap.out = b.ATT16; ap.lens = lens; ap.S = S; ap.T_alloc = T; ap.Tmax = T; ap.R_alloc = b.R_alloc; ap.chunk = chunk; ap.halo = kHalo;
    e.launches++;
    if (!dry) {
      e.prof_begin(st, Engine::F_REL_ATTN);
      launch_rel_attn(ap, st);
      e.prof_end(st);
      e.range_scan(st, Engine::R_ATTN_OUT, b.ATT16, (long long)S * T, 512, 512);
    }
  }
  {  // linear_out + residual
    GemmParams p = base_params(lens);
    p.res = b.X32; p.res_ld = 512;
    p.out32 = b.X32; p.out32_ld = 512;
    e.gemm(st, b.ATT16, S, T, 512, 512, e.W(lp + ".o"), 256, 1, tap1, p, dry);
  }
  e.launches++;
  if (!dry) {
    LN ln = e.ln(lp + ".ln_ff");
    launch_layernorm512(b.X32, ln.g, ln.b, 1e-12f, b.H16, nullptr, lens, 0, S, T, st);
  }
  {  // w_1 + SiLU
    GemmParams p = base_params(lens);
    p.act = ACT_SILU;
    p.emit[0] = emit_plain(b.F16, 2048);
    e.gemm(st, b.H16, S, T, 512, 512, e.W(lp + ".ff1"), 256, 1, tap1, p, dry);
  }
  {  // w_2 + residual
    GemmParams p = base_params(lens);
    p.res = b.X32; p.res_ld = 512;
    p.out32 = b.X32; p.out32_ld = 512;
    if (plain_out) p.emit[0] = emit_plain(plain_out, 512);
    e.gemm(st, b.F16, S, T, 2048, 2048, e.W(lp + ".ff2"), 256, 1, tap1, p, dry);
  }
  if (next_norm) {
    e.launches++;
    if (!dry) launch_layernorm512(b.X32, next_norm->g, next_norm->b, 1e-12f, b.H16, nullptr, lens, 0, S, T, st);
  }
}

size_t flow_forward(Engine& e, cudaStream_t st, const FlowArgs& a, Arena& ws) {
  const bool dry = ws.measuring();
  e.in_hift = false;
  const int B = a.B;
  const StreamState* ss = a.stream_state;
  if (ss) {
    CV2_CHECK(B == ss->n_slots && a.n_steps == ss->n_steps, "stream state was made for %d slots / %d steps, call has %d / %d", ss->n_slots,
              ss->n_steps, B, a.n_steps);
    CV2_CHECK(a.streaming && !a.finalize && !a.enc_only, "the stream state serves non-final streaming chunks only");
    CV2_CHECK(2 * (a.max_tok_total - 3) <= ss->T_cap, "chunk of %d mel frames exceeds the stream state's capacity %d",
              2 * (a.max_tok_total - 3), ss->T_cap);
  }
  // with a stream state the row strides are the state's (its k / v^T caches are laid out for T_cap rows)
  const int Tt = ss ? round_up(ss->T_cap / 2 + 8, 128) : round_up(a.max_tok_total + 4, 128);   // token-rate rows (+ lookahead reads)
  const int Tm = ss ? ss->T_cap : round_up(2 * a.max_tok_total, 128);                         // mel-rate rows
  static const int tap1[1] = {0};

  // ---- lengths on the device (no host sync) ----
  int* len_ctx = ws.get<int>(B);       // prompt + token
  int* len_enc = ws.get<int>(B);       // tokens the encoder keeps (minus the 3 lookahead tokens when not final)
  int* len_mel = ws.get<int>(2 * B);   // mel frames, duplicated for the two CFG rows
  int* t_lo = ss ? ws.get<int>(2 * B) : nullptr;   // first row of the first tile the estimator computes, per CFG row
  e.launches += 4;
  if (!dry && ss) {
    launch_stream_lens(a.prompt_len, a.token_len, ss->t_done, len_ctx, len_enc, len_mel, t_lo, B, st);
  } else if (!dry && a.enc_only) {
    launch_lens_clamp(a.enc_lens, a.enc_T, 0, len_enc, B, st);
    launch_lens_clamp(a.enc_lens, a.enc_T, a.enc_ctx ? 3 : 0, len_ctx, B, st);
  } else if (!dry) {
    launch_lens_affine(a.prompt_len, a.token_len, 1, 0, len_ctx, B, st);
    launch_lens_affine(len_ctx, nullptr, 1, a.finalize ? 0 : -3, len_enc, B, st);
  }
  if (!dry && !ss) {
    launch_lens_affine(len_enc, nullptr, 2, 0, len_mel, B, st);
    launch_lens_affine(len_enc, nullptr, 2, 0, len_mel + B, B, st);
  }
  e.tile_lists.clear();
  e.make_tile_list(st, ws, len_ctx, B, Tt, kHalo, dry, len_enc);        // token-rate encoder GEMMs
  e.make_tile_list(st, ws, len_mel, 2 * B, Tm, kEstHalo, dry, nullptr, nullptr, t_lo);   // estimator (both CFG rows)
  e.make_tile_list(st, ws, len_mel, B, Tm, kHalo, dry);                 // mel-rate encoder GEMMs

  // ---- encoder, token rate ----
  EncBuffers eb;
  __half* A0 = ws.get<__half>((size_t)B * Tt * 512);
  float* E32 = ws.get<float>((size_t)B * Tm * 512);
  __half* X16 = ws.get<__half>((size_t)B * Tm * 512);
  __half* Y16 = ws.get<__half>((size_t)B * Tm * 512);
  eb.X32 = ws.get<float>((size_t)B * Tm * 512);
  eb.QU16 = ws.get<__half>((size_t)B * Tm * 512);
  eb.QV16 = ws.get<__half>((size_t)B * Tm * 512);
  eb.K16 = ws.get<__half>((size_t)B * Tm * 512);
  eb.VT16 = ws.get<__half>((size_t)B * Tm * 512);
  eb.R_alloc = round_up(2 * Tm - 1, 128);
  eb.POS16 = ws.get<__half>((size_t)eb.R_alloc * 512);
  eb.PE16 = ws.get<__half>((size_t)eb.R_alloc * 512);
  __half* PE16_t = ws.get<__half>((size_t)round_up(2 * Tt - 1, 128) * 512);
  eb.H16 = ws.get<__half>((size_t)B * Tm * 512);
  eb.ATT16 = ws.get<__half>((size_t)B * Tm * 512);
  eb.F16 = ws.get<__half>((size_t)B * Tm * 2048);
  float *MU32 = nullptr, *COND32 = nullptr, *SPK32 = nullptr, *Xst = nullptr, *tres = nullptr;
  __half* XIN16 = nullptr;
  EstBuffers sb;
  memset(&sb, 0, sizeof(sb));
  if (!a.enc_only) {
    MU32 = ws.get<float>((size_t)B * Tm * 80);
    COND32 = ws.get<float>((size_t)B * Tm * 80);
    SPK32 = ws.get<float>((size_t)B * 80);
    Xst = ws.get<float>((size_t)B * Tm * 80);
    XIN16 = ws.get<__half>((size_t)2 * B * Tm * 320);
    sb = est_alloc(ws, 2 * B, Tm);
    tres = est_time(e, st, ws, a.t_steps, a.n_steps, dry);
  }

  e.launches += 1;
  if (!dry) {
    if (a.enc_only)
      launch_embed_rows(a.enc_xs, a.enc_T, len_enc, a.enc_ctx, 3, A0, B, Tt, st);
    else
      launch_embed_tokens(a.prompt_token, a.prompt_len, a.prompt_stride, a.token, a.token_len, a.token_stride,
                          e.f32("flow.embedding"), A0, B, Tt, 6561, st);
  }
  {  // embed: Linear -> (LN * sqrt(512) folded into gamma/beta)
    GemmParams p = base_params(len_ctx);
    p.out32 = E32; p.out32_ld = 512;
    e.gemm(st, A0, B, Tt, 512, 512, e.W("enc.embed"), 256, 1, tap1, p, dry);
  }
  e.launches += 2;
  if (!dry) {
    LN ln = e.ln("enc.embed.ln");
    launch_layernorm512(E32, ln.g, ln.b, 1e-5f, X16, eb.X32, len_ctx, 0, B, Tt, st);
    launch_pos_table(PE16_t, Tt, st);
  }
  {  // pre-lookahead conv1 (k=4, looks 3 tokens ahead) + leaky_relu(0.01)
    static const int look4[4] = {0, 1, 2, 3};
    GemmParams p = base_params(len_enc);
    p.act = ACT_LRELU; p.act_f = 0.01f;
    p.emit[0] = emit_plain(Y16, 512);
    e.gemm(st, X16, B, Tt, 512, 512, e.W("enc.pre.conv1"), 256, 4, look4, p, dry);
  }
  {  // conv2 (k=3 causal) + residual
    static const int causal3[3] = {-2, -1, 0};
    GemmParams p = base_params(len_enc);
    p.res = eb.X32; p.res_ld = 512;
    p.out32 = eb.X32; p.out32_ld = 512;
    e.gemm(st, Y16, B, Tt, 512, 512, e.W("enc.pre.conv2"), 256, 3, causal3, p, dry);
  }
  e.launches++;
  if (!dry) {
    LN ln = e.ln("enc.layers.0.ln_mha");
    launch_layernorm512(eb.X32, ln.g, ln.b, 1e-12f, eb.H16, nullptr, len_enc, 0, B, Tt, st);
  }
  {
    EncBuffers tb = eb;
    tb.PE16 = PE16_t;
    tb.R_alloc = round_up(2 * Tt - 1, 128);
    for (int i = 0; i < 6; i++) {
      LN nn;
      if (i < 5) nn = e.ln("enc.layers." + std::to_string(i + 1) + ".ln_mha");
      enc_layer(e, st, tb, "enc.layers." + std::to_string(i), len_enc, B, Tt, a.streaming ? 25 : 0, i < 5 ? &nn : nullptr,
                i == 5 ? X16 : nullptr, dry);
    }
  }
  // ---- upsample x2 + causal conv k5, up_embed ----
  e.launches += 1;
  if (!dry) launch_repeat2(X16, Y16, len_enc, B, Tt, Tm, 512, st);
  {
    static const int causal5[5] = {-4, -3, -2, -1, 0};
    GemmParams p = base_params(len_mel);
    p.emit[0] = emit_plain(X16, 512);
    e.gemm(st, Y16, B, Tm, 512, 512, e.W("enc.up.conv"), 256, 5, causal5, p, dry);
  }
  {
    GemmParams p = base_params(len_mel);
    p.out32 = E32; p.out32_ld = 512;
    e.gemm(st, X16, B, Tm, 512, 512, e.W("enc.up_embed"), 256, 1, tap1, p, dry);
  }
  e.launches += 3;
  if (!dry) {
    LN ln = e.ln("enc.up_embed.ln");
    launch_layernorm512(E32, ln.g, ln.b, 1e-5f, nullptr, eb.X32, len_mel, 0, B, Tm, st);
    LN l0 = e.ln("enc.up_layers.0.ln_mha");
    launch_layernorm512(eb.X32, l0.g, l0.b, 1e-12f, eb.H16, nullptr, len_mel, 0, B, Tm, st);
    launch_pos_table(eb.PE16, Tm, st);
  }
  for (int i = 0; i < 4; i++) {
    LN nn;
    if (i < 3) nn = e.ln("enc.up_layers." + std::to_string(i + 1) + ".ln_mha");
    enc_layer(e, st, eb, "enc.up_layers." + std::to_string(i), len_mel, B, Tm, a.streaming ? 50 : 0, i < 3 ? &nn : nullptr, nullptr,
              dry);
  }
  e.launches++;
  if (!dry) {
    LN ln = e.ln("enc.after_norm");
    launch_layernorm512(eb.X32, ln.g, ln.b, 1e-5f, eb.H16, a.enc_out ? E32 : nullptr, len_mel, 0, B, Tm, st);
    const size_t enc_rows = a.enc_only ? (size_t)2 * a.enc_T : (size_t)2 * a.max_tok_total;
    if (a.enc_out)
      CV2_CUDA(cudaMemcpy2DAsync(a.enc_out, enc_rows * 512 * 4, E32, (size_t)Tm * 512 * 4, enc_rows * 512 * 4, B,
                                 cudaMemcpyDeviceToDevice, st));
  }
  if (a.enc_only) return ws.peak;
  {  // encoder_proj 512 -> 80
    GemmParams p = base_params(len_mel);
    p.out32 = MU32; p.out32_ld = 80;
    e.gemm(st, eb.H16, B, Tm, 512, 512, e.W("enc.proj"), 128, 1, tap1, p, dry);
  }
  // ---- conditioning + Euler solve ----
  e.launches += 4;
  if (!dry) {
    launch_build_cond(a.prompt_feat, a.prompt_feat_bstride, a.prompt_feat_len, COND32, B, Tm, st);
    launch_spk_affine(a.embedding, e.f32("flow.spk.w"), e.f32("flow.spk.b"), SPK32, B, st);
    launch_pack_cond(MU32, SPK32, COND32, XIN16, len_mel, B, Tm, st);
    launch_euler_pack(Xst, nullptr, a.rand_noise, a.noise_stride, XIN16, len_mel, B, Tm, 0.f, a.cfg, 1, st);
    if (a.mu_out) launch_ntc_to_nct(MU32, Tm, 80, 0, nullptr, a.mu_out, (long long)80 * 2 * a.max_tok_total, 2 * a.max_tok_total, 80,
                                    len_mel, B, st);
  }
  // the fused weights are packed for the model's inference_cfg_rate (0.7, cosyvoice2.yaml:75); any other rate takes the
  // separate CFG-combine + Euler kernel
  const bool fuse_euler = e.fuse_euler && fabsf(a.cfg - 0.7f) < 1e-6f && e.tensors.count("est.proj_cfg.w");
  for (int step = 0; step < a.n_steps; step++) {
    EstStream es{ss, step, t_lo};
    if (fuse_euler) {
      EulerFuse ef{Xst, XIN16, dry ? 0.f : a.dt_steps[step], B};
      est_core(e, st, sb, XIN16, len_mel, 2 * B, Tm, tres, a.n_steps, step, 0, a.streaming, dry, &ef, ss ? &es : nullptr);
      continue;
    }
    est_core(e, st, sb, XIN16, len_mel, 2 * B, Tm, tres, a.n_steps, step, 0, a.streaming, dry, nullptr, ss ? &es : nullptr);
    e.launches++;
    if (!dry) launch_euler_pack(Xst, sb.V32, nullptr, 0, XIN16, len_mel, B, Tm, a.dt_steps[step], a.cfg, 0, st);
  }
  e.launches++;
  if (!dry)
    launch_ntc_to_nct(Xst, Tm, 80, 0, a.prompt_feat_len, a.mel_out, (long long)80 * a.mel_out_T, a.mel_out_T, 80, len_mel, B, st);
  return ws.peak;
}

size_t stream_state_bytes(int n_slots, int T_cap, int n_steps) {
  CV2_CHECK(n_slots >= 1 && T_cap >= 128 && T_cap % 128 == 0 && n_steps >= 1, "stream state: bad sizes (%d slots, T_cap %d, %d steps)",
            n_slots, T_cap, n_steps);
  StreamState s;
  s.n_slots = n_slots; s.T_cap = T_cap; s.n_steps = n_steps;
  return 256 + ((size_t)n_slots * 4 + 255) / 256 * 256 + 2 * (size_t)n_steps * kEstBlocks * s.kv_block() * sizeof(__half) +
         (size_t)n_steps * kEstConvs * s.tail_block() * sizeof(__half) + 768;
}

StreamState stream_state_carve(void* base, size_t bytes, int n_slots, int T_cap, int n_steps) {
  CV2_CHECK(base && bytes >= stream_state_bytes(n_slots, T_cap, n_steps), "stream state buffer too small");
  StreamState s;
  s.n_slots = n_slots; s.T_cap = T_cap; s.n_steps = n_steps;
  Arena a;
  a.base = static_cast<uint8_t*>(base);
  a.cap = bytes;
  s.t_done = a.get<int>(n_slots);
  s.kcache = a.get<__half>((size_t)n_steps * kEstBlocks * s.kv_block());
  s.vcache = a.get<__half>((size_t)n_steps * kEstBlocks * s.kv_block());
  s.tails = a.get<__half>((size_t)n_steps * kEstConvs * s.tail_block());
  return s;
}

}  // namespace cv2
