// hift.inference on the device (reference: cosyvoice/hifigan/generator.py:570-582 inference, :520-552 decode,
// f0_predictor.py:55-58).  mel [B,80,T] -> f0 -> NSF source -> source STFT -> conv_pre -> 3 x (transposed conv,
// source fusion, 3 Snake ResBlocks) -> conv_post -> iSTFT -> waveform.  All convs are tap GEMMs on tensor cores
// with Snake / leaky-relu fused into the producing epilogue; F0 predictor, source and (i)STFT are fp32.
#include "engine.h"
#include "flow_kernels.cuh"
#include "hift_kernels.cuh"

namespace cv2 {

static const int kHalo = 32;

static GemmParams hp(const int* lens) {
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.lens = lens;
  p.halo = kHalo;
  p.out_scale = 1.f;
  return p;
}
static Emit mk_emit(int kind, __half* ptr, long long ld, const float* a = nullptr, float f = 0.f) {
  Emit e;
  memset(&e, 0, sizeof(e));
  e.kind = kind; e.ptr = ptr; e.ld = ld; e.a = a; e.f = f; e.scale = 1.f;
  return e;
}
static void conv_taps(int k, int d, int* taps) {
  for (int j = 0; j < k; j++) taps[j] = (j - (k - 1) / 2) * d;
}

size_t hift_forward(Engine& e, cudaStream_t st, const HiftArgs& a, Arena& ws) {
  const bool dry = ws.measuring();
  const int B = a.B;
  const int up_rate[3] = {8, 5, 3};
  const int up_taps[3] = {2, 3, 3};
  const int up_pad[3] = {4, 3, 2};
  const int chans[4] = {512, 256, 128, 64};
  const int sd_k[3] = {30, 6, 1}, sd_s[3] = {15, 3, 1}, sd_p[3] = {7, 1, 0};
  const int sd_fpl[3] = {8, 40, 120}, sd_add[3] = {0, 0, 1};
  const int rb_k[3] = {3, 7, 11}, srb_k[3] = {7, 7, 11}, dil[3] = {1, 3, 5};

  // per-stage lengths (frames) on the device
  int Tmax[4];
  Tmax[0] = a.mel_T; Tmax[1] = 8 * a.mel_T; Tmax[2] = 40 * a.mel_T; Tmax[3] = 120 * a.mel_T + 1;
  int Ta[4];
  for (int i = 0; i < 4; i++) Ta[i] = round_up(Tmax[i] + kHalo, 128);
  int* lens[4];
  for (int i = 0; i < 4; i++) lens[i] = ws.get<int>(B);
  int* lens0_src = ws.get<int>(B);
  e.launches += 5;
  if (!dry) {
    if (a.lens) launch_lens_affine(a.lens, nullptr, 1, 0, lens[0], B, st);
    else launch_lens_affine(lens0_src, nullptr, 0, a.mel_T, lens[0], B, st);
    launch_lens_affine(lens[0], nullptr, 8, 0, lens[1], B, st);
    launch_lens_affine(lens[0], nullptr, 40, 0, lens[2], B, st);
    launch_lens_affine(lens[0], nullptr, 120, 1, lens[3], B, st);
  }
  e.tile_lists.clear();
  e.in_hift = true;
  for (int i = 0; i < 4; i++) e.make_tile_list(st, ws, lens[i], B, Ta[i], kHalo, dry);

  // ---- mel -> channels-last (fp32 for the F0 predictor, 16-bit for conv_pre) ----
  float* MEL32 = ws.get<float>((size_t)B * Ta[0] * 80);
  __half* MEL16 = ws.get<__half>((size_t)B * Ta[0] * 80);
  float* FA = ws.get<float>((size_t)B * Ta[0] * 512);
  float* FB = ws.get<float>((size_t)B * Ta[0] * 512);
  float* F0 = ws.get<float>((size_t)B * Ta[0]);
  float* PH = ws.get<float>((size_t)B * Ta[0] * 9);
  float* STFT = ws.get<float>((size_t)B * Ta[3] * 18);
  // F0 predictor (f0_predictor.py:55-58).  Its output drives a phase accumulator (0.1 Hz decorrelates the harmonics within
  // a second), so plain 16-bit operands are not enough -- but fp32 SIMT was 2 % of the step.  Split precision on the tensor
  // cores: every activation and weight is carried as hi + lo 16-bit halves and the conv is hi*W_hi + lo*W_hi + hi*W_lo in one
  // tap GEMM (6 "taps": the 3 conv taps over [hi | lo] against [W_hi | W_hi], then again against [W_lo | 0]); products are
  // exact to ~2^-21, accumulation is fp32.
  const bool f0_split = e.f0_split && e.tensors.count("f0.t0.w");
  __half* MS16 = f0_split ? ws.get<__half>((size_t)B * Ta[0] * 256) : nullptr;
  __half* FS[2] = {nullptr, nullptr};
  if (f0_split)
    for (int i = 0; i < 2; i++) FS[i] = ws.get<__half>((size_t)B * Ta[0] * 1024);
  e.launches += 1 + 5 + 1 + 2 + 1;
  if (!dry) launch_nct_to_ntc(a.mel, (long long)80 * a.mel_T, a.mel_T, MEL32, MEL16, lens[0], 0, B, Ta[0], 80, 80, 0, 0, st);
  if (f0_split) {
    static const int taps6[6] = {-1, 0, 1, -1, 0, 1};
    e.launches++;
    if (!dry) launch_split16(MEL32, 80, MS16, 256, 128, (long long)B * Ta[0], st);
    const __half* cur = MS16;
    int kc = 256;
    for (int l = 0; l < 5; l++) {
      GemmParams p = hp(lens[0]);
      p.act = ACT_ELU;
      p.acc_scale = 1.f / 256.f;   // pack.py stores these weights x256 so that W_lo is a normal fp16 number
      if (l < 4) {
        p.emit[0] = mk_emit(EMIT_PLAIN, FS[l & 1], 1024);
        p.emit[1] = mk_emit(EMIT_LO, FS[l & 1], 1024);
        p.emit[1].col_off = 512;
      } else {
        p.out32 = FA; p.out32_ld = 512;
      }
      e.gemm(st, cur, B, Ta[0], kc, kc, e.W("f0.t" + std::to_string(l)), 256, 6, taps6, p, dry);
      cur = FS[l & 1];
      kc = 1024;
    }
  }
  if (!dry) {
    const float* cur = f0_split ? FA : MEL32;
    int cin = 80;
    for (int l = 0; l < 5 && !f0_split; l++) {
      const std::string n = "f0.c" + std::to_string(l);
      float* out = (l & 1) ? FB : FA;
      e.prof_begin(st, Engine::F_F0_CONV);
      launch_conv3_elu_f32(cur, cin, e.f32(n + ".w"), e.f32(n + ".b"), out, 512, lens[0], 0, B, Ta[0], st);
      e.prof_end(st);
      cur = out;
      cin = 512;
    }
    launch_f0_head(cur, e.f32("f0.cls.w"), e.f32("f0.cls.b"), F0, lens[0], 0, B, Ta[0], Ta[0], st);
    if (a.f0_out) CV2_CUDA(cudaMemcpy2DAsync(a.f0_out, (size_t)a.mel_T * 4, F0, (size_t)Ta[0] * 4, (size_t)a.mel_T * 4, B,
                                             cudaMemcpyDeviceToDevice, st));
    e.prof_begin(st, Engine::F_NSF);
    launch_nsf_source(F0, Ta[0], PH, Ta[0], lens[0], 0, a.noise, (long long)480 * a.mel_T * 9, a.seed, e.seed_dev, e.f32("hift.src.lw"),
                      e.f32("hift.src.lb"), a.cache_source, a.cache_len, a.cache_len, a.source, (long long)480 * a.mel_T, B,
                      a.mel_T, st);
    e.prof_end(st);
    e.prof_begin(st, Engine::F_STFT);
    launch_source_stft(a.source, (long long)480 * a.mel_T, lens[0], 0, STFT, Ta[3], B, st);
    e.prof_end(st);
  }

  // ---- conv_pre (k7) + leaky_relu(0.1) -> U16 ----
  __half* U16 = ws.get<__half>((size_t)B * Ta[0] * 512);
  {
    int taps[7];
    conv_taps(7, 1, taps);
    GemmParams p = hp(lens[0]);
    p.emit[0] = mk_emit(EMIT_LRELU, U16, 512, nullptr, 0.1f);
    e.gemm(st, MEL16, B, Ta[0], 80, 80, e.W("hift.conv_pre"), 256, 7, taps, p, dry);
  }

  // The three upsampling stages run one after the other and only hand a 16-bit tensor (NEXT16) to the next one, so their
  // scratch tensors share one region of the workspace (sized by the largest stage) instead of 12 tensors per stage side by side:
  // 16.7 GB instead of 32 GB for 64 utterances of up to 20 s.  Stale contents are harmless: every kernel writes the rows of
  // its active tiles (zeros past a sequence's end) before anything reads them, which is also what makes reusing the whole
  // workspace from call to call valid.
  __half* NEXT16s[3];
  for (int i = 0; i < 3; i++) NEXT16s[i] = ws.get<__half>((size_t)B * Ta[i + 1] * chans[i + 1]);
  const size_t stage_mark = ws.mark();
  const __half* up_in = U16;
  for (int i = 0; i < 3; i++) {
    const int Cin = chans[i], C = chans[i + 1], T = Ta[i + 1];
    const int* ln = lens[i + 1];
    const size_t rows = (size_t)B * T;
    ws.rewind(stage_mark);
    float* XU32 = ws.get<float>(rows * C);
    float* SI32 = ws.get<float>(rows * C);
    float* XF32 = ws.get<float>(rows * C);
    float* XR32 = ws.get<float>(rows * C);
    float* XS32 = ws.get<float>(rows * C);
    __half* S16 = ws.get<__half>(rows * C);
    __half* TMP16 = ws.get<__half>(rows * C);
    __half* RB16[3];
    for (int j = 0; j < 3; j++) RB16[j] = ws.get<__half>(rows * C);
    __half* A16 = ws.get<__half>(rows * C);
    __half* NEXT16 = NEXT16s[i];

    {  // transposed conv as a phase-concatenated GEMM over input frames
      int taps[3] = {0, -1, -2};
      GemmParams p = hp(lens[i]);
      p.out32 = XU32; p.out32_ld = (long long)up_rate[i] * C;
      p.flat = 1;
      p.flat_seq_elems = (long long)T * C;
      const int shift = (i == 2) ? 1 : 0;                       // ReflectionPad1d((1,0)) shifts the last stage by one row
      p.flat_off = (long long)(shift - up_pad[i]) * C;
      p.flat_lo = (long long)shift * C;
      p.flat_hi_per_len = (long long)up_rate[i] * C;
      p.flat_hi_add = (long long)shift * C;
      const int bn = (up_rate[i] * C) >= 2048 ? 256 : ((up_rate[i] * C) >= 512 ? 128 : 64);
      e.gemm(st, up_in, B, Ta[i], Cin, Cin, e.W("hift.ups." + std::to_string(i)), bn, up_taps[i], taps, p, dry);
    }
    if (i == 2) {
      e.launches++;
      if (!dry) launch_reflect_row0(XU32, B, T, C, st);
    }
    const std::string sp = "hift.srb." + std::to_string(i);
    e.launches++;
    if (!dry) {
      const std::string dn = "hift.sd." + std::to_string(i);
      e.prof_begin(st, Engine::F_SRC_DOWN);
      launch_source_down(STFT, Ta[3], e.f32(dn + ".w"), e.f32(dn + ".b"), sd_k[i], sd_s[i], sd_p[i], C, lens[0], 0, sd_fpl[i],
                         sd_add[i], SI32, S16, e.f32(sp + ".a1.0"), B, T, st);
      e.prof_end(st);
    }
    const int bnc = C >= 256 ? 256 : C;
    // source ResBlock; its last conv also adds the upsampled main path and emits the three ResBlock inputs
    for (int d = 0; d < 3; d++) {
      int taps[16];
      {
        conv_taps(srb_k[i], dil[d], taps);
        GemmParams p = hp(ln);
        p.act = ACT_SNAKE; p.act_a = e.f32(sp + ".a2." + std::to_string(d));
        p.emit[0] = mk_emit(EMIT_PLAIN, TMP16, C);
        e.gemm(st, S16, B, T, C, C, e.W(sp + ".c1." + std::to_string(d)), bnc, srb_k[i], taps, p, dry);
      }
      {
        conv_taps(srb_k[i], 1, taps);
        GemmParams p = hp(ln);
        p.res = SI32; p.res_ld = C;
        if (d < 2) {
          p.out32 = SI32; p.out32_ld = C;
          p.emit[0] = mk_emit(EMIT_SNAKE, S16, C, e.f32(sp + ".a1." + std::to_string(d + 1)));
        } else {
          p.res2 = XU32; p.res2_ld = C;
          p.out32 = XF32; p.out32_ld = C;
          for (int j = 0; j < 3; j++)
            p.emit[j] = mk_emit(EMIT_SNAKE, RB16[j], C, e.f32("hift.rb." + std::to_string(i * 3 + j) + ".a1.0"));
        }
        e.gemm(st, TMP16, B, T, C, C, e.W(sp + ".c2." + std::to_string(d)), bnc, srb_k[i], taps, p, dry);
      }
    }
    // three parallel ResBlocks, averaged
    for (int j = 0; j < 3; j++) {
      const std::string rp = "hift.rb." + std::to_string(i * 3 + j);
      const int k = rb_k[j];
      const __half* ain = RB16[j];
      const float* xin = XF32;
      for (int d = 0; d < 3; d++) {
        int taps[16];
        {
          conv_taps(k, dil[d], taps);
          GemmParams p = hp(ln);
          p.act = ACT_SNAKE; p.act_a = e.f32(rp + ".a2." + std::to_string(d));
          p.emit[0] = mk_emit(EMIT_PLAIN, TMP16, C);
          e.gemm(st, ain, B, T, C, C, e.W(rp + ".c1." + std::to_string(d)), bnc, k, taps, p, dry);
        }
        {
          conv_taps(k, 1, taps);
          GemmParams p = hp(ln);
          p.res = xin; p.res_ld = C;
          if (d < 2) {
            p.out32 = XR32; p.out32_ld = C;
            p.emit[0] = mk_emit(EMIT_SNAKE, A16, C, e.f32(rp + ".a1." + std::to_string(d + 1)));
          } else {
            p.out_scale = 1.f / 3.f;
            p.out32 = XS32; p.out32_ld = C;
            p.out32_accum = j > 0;
            if (j == 2) p.emit[0] = mk_emit(EMIT_LRELU, NEXT16, C, nullptr, i < 2 ? 0.1f : 0.01f);
          }
          e.gemm(st, TMP16, B, T, C, C, e.W(rp + ".c2." + std::to_string(d)), bnc, k, taps, p, dry);
        }
        ain = A16;
        xin = XR32;
      }
    }
    up_in = NEXT16;
  }

  // ---- conv_post (k7, 64 -> 18) + iSTFT head ----
  ws.rewind(stage_mark);
  float* CP32 = ws.get<float>((size_t)B * Ta[3] * 18);
  {
    int taps[7];
    conv_taps(7, 1, taps);
    GemmParams p = hp(lens[3]);
    p.out32 = CP32; p.out32_ld = 18;
    e.gemm(st, up_in, B, Ta[3], 64, 64, e.W("hift.conv_post"), 64, 7, taps, p, dry);
  }
  e.launches++;
  if (!dry) {
    e.prof_begin(st, Engine::F_ISTFT);
    launch_istft(CP32, Ta[3], 18, lens[0], 0, a.speech, (long long)480 * a.mel_T, B, a.mel_T, st, a.pcm16);
    e.prof_end(st);
  }
  e.in_hift = false;
  return ws.peak;
}

}  // namespace cv2
