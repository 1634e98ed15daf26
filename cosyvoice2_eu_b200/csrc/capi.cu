// extern "C" surface of libcv2eu_b200.so (declared in include/cv2eu_b200.h).
#include <stdlib.h>
#include "../../include/cv2eu_b200.h"

#include "attention.cuh"
#include "engine.h"
#include "flow_kernels.cuh"
#include "hift_kernels.cuh"
#include "kaldi_fbank.cuh"
#include "prompt_mel.cuh"

using namespace cv2;

struct cv2_engine {
  Engine e;
};

#define CV2_API_BEGIN try {
#define CV2_API_END                                \
  }                                                \
  catch (const std::exception& ex) {               \
    last_error_ref() = ex.what();                  \
    return 1;                                      \
  }                                                \
  catch (...) {                                    \
    last_error_ref() = "unknown error";            \
    return 1;                                      \
  }                                                \
  return 0;

// every forward runs on the engine's own device whatever the caller's current device is
static void use_device(const cv2_engine* h) { CV2_CUDA(cudaSetDevice(h->e.device)); }

static void require_sm100(int device) {
  cudaDeviceProp prop;
  CV2_CUDA(cudaGetDeviceProperties(&prop, device));
  CV2_CHECK(prop.major == 10, "cv2eu_b200 needs an sm_100-class GPU (B200); device %d is sm_%d%d -- there is no fallback path",
            device, prop.major, prop.minor);
}

extern "C" {

const char* cv2_last_error(void) { return last_error_ref().c_str(); }
int cv2_version(void) { return 1; }

int cv2_engine_create(cv2_engine** out, int device) {
  CV2_API_BEGIN
  CV2_CHECK(out != nullptr, "null out");
  int n = 0;
  cudaError_t err = cudaGetDeviceCount(&n);
  CV2_CHECK(err == cudaSuccess && n > 0, "no CUDA device available (%s) -- cv2eu_b200 has no CPU fallback",
            cudaGetErrorString(err));
  CV2_CUDA(cudaSetDevice(device));
  require_sm100(device);
  hift_init_tables();
  cv2_engine* h = new cv2_engine();
  h->e.device = device;
  *out = h;
  CV2_API_END
}

void cv2_engine_destroy(cv2_engine* e) {
  if (e && e->e.range_dev) cudaFree(e->e.range_dev);
  delete e;
}

int cv2_engine_set_tensor(cv2_engine* h, const char* name, const void* dptr, int dtype, int ndim, const int64_t* shape) {
  CV2_API_BEGIN
  CV2_CHECK(h && name && dptr, "null argument");
  TensorRef t;
  t.ptr = dptr;
  t.dtype = dtype;
  t.shape.assign(shape, shape + ndim);
  h->e.tensors[name] = t;
  h->e.weights.erase(std::string(name).substr(0, std::string(name).rfind('.')));
  CV2_API_END
}

int cv2_engine_finalize(cv2_engine* h, int need_flow, int need_hift) {
  CV2_API_BEGIN
  CV2_CHECK(h, "null engine");
  Engine& e = h->e;
  // A dry run of each forward touches every tensor it needs and reports the first missing name.
  if (need_flow) {
    Arena ws;
    FlowArgs a;
    memset(&a, 0, sizeof(a));
    a.B = 1; a.max_tok_total = 64; a.n_steps = 1; a.finalize = 1;
    flow_forward(e, nullptr, a, ws);
    e.has_flow = true;
  }
  if (need_hift) {
    Arena ws;
    HiftArgs a;
    memset(&a, 0, sizeof(a));
    a.B = 1; a.mel_T = 16;
    hift_forward(e, nullptr, a, ws);
    e.has_hift = true;
  }
  e.finalized = true;
  CV2_API_END
}

long long cv2_engine_last_launches(cv2_engine* h) { return h ? h->e.launches : -1; }

int cv2_engine_set_seed_ptr(cv2_engine* h, const unsigned long long* seed_dev) {
  CV2_API_BEGIN
  CV2_CHECK(h, "null engine");
  h->e.seed_dev = seed_dev;
  CV2_API_END
}

int cv2_engine_set_option(cv2_engine* h, const char* name, int value) {
  CV2_API_BEGIN
  CV2_CHECK(h && name, "null argument");
  const std::string n(name);
  if (n == "fuse_euler") h->e.fuse_euler = value != 0;
  else if (n == "fuse_ffn") h->e.fuse_ffn = value != 0;
  else if (n == "f0_split") h->e.f0_split = value != 0;
  else if (n == "cluster_mc") h->e.cluster_mc = value != 0;
  else if (n == "ffn_2cta") h->e.ffn_2cta = value != 0;
  else if (n == "min_2sm_tiles") h->e.min_2sm_tiles = value;
  else if (n == "min_2sm_tiles_ffn") h->e.min_2sm_tiles_ffn = value;
  else if (n == "chain_outproj") h->e.chain_outproj = value != 0;
  else if (n == "ffn_hsplit") h->e.ffn_hsplit = value != 0;
  else if (n == "conv_mode") h->e.conv_mode = value;
  else if (n == "range_check") {
    use_device(h);
    if (value && !h->e.range_dev) CV2_CUDA(cudaMalloc(&h->e.range_dev, Engine::R_COUNT * sizeof(unsigned)));
    if (h->e.range_dev) CV2_CUDA(cudaMemset(h->e.range_dev, 0, Engine::R_COUNT * sizeof(unsigned)));
    h->e.range_check = value != 0;
  }
  else fail("unknown engine option '%s'", name);
  CV2_API_END
}

int cv2_engine_read_ranges(cv2_engine* h, float* max_abs, int n) {
  CV2_API_BEGIN
  CV2_CHECK(h && max_abs && n >= 1, "null argument");
  CV2_CHECK(h->e.range_dev, "range_check was never enabled (cv2_engine_set_option(e, \"range_check\", 1))");
  use_device(h);
  unsigned bits[Engine::R_COUNT];
  CV2_CUDA(cudaDeviceSynchronize());
  CV2_CUDA(cudaMemcpy(bits, h->e.range_dev, sizeof(bits), cudaMemcpyDeviceToHost));
  for (int i = 0; i < n; i++) {
    float f = 0.f;
    if (i < Engine::R_COUNT) memcpy(&f, &bits[i], 4);
    max_abs[i] = f;
  }
  CV2_CUDA(cudaMemset(h->e.range_dev, 0, sizeof(bits)));
  CV2_API_END
}

int cv2_debug_set_ffn_trace(long long* dev_buf) {
  CV2_API_BEGIN
  ffn_set_trace(dev_buf);
  CV2_API_END
}

int cv2_engine_set_profiling(cv2_engine* h, int on) {
  CV2_API_BEGIN
  CV2_CHECK(h, "null engine");
  h->e.profiling = on != 0;
  h->e.prof.clear();
  h->e.ev_used = 0;
  CV2_API_END
}

int cv2_engine_read_profile(cv2_engine* h, double* ms_per_family, long long* launches_per_family, int n_families) {
  CV2_API_BEGIN
  CV2_CHECK(h && ms_per_family && launches_per_family, "null argument");
  for (int i = 0; i < n_families; i++) {
    ms_per_family[i] = 0.0;
    launches_per_family[i] = 0;
  }
  for (auto& r : h->e.prof) {
    CV2_CUDA(cudaEventSynchronize(r.b));
    float ms = 0.f;
    CV2_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
    if (r.family < n_families) {
      ms_per_family[r.family] += ms;
      launches_per_family[r.family]++;
    }
  }
  h->e.prof.clear();
  h->e.ev_used = 0;
  CV2_API_END
}

size_t cv2_estimator_workspace_bytes(cv2_engine* h, int B2, int T) {
  try {
    Arena ws;
    EstArgs a;
    memset(&a, 0, sizeof(a));
    a.B2 = B2; a.T = T;
    return estimator_forward(h->e, nullptr, a, ws) + 4096;
  } catch (const std::exception& ex) {
    last_error_ref() = ex.what();
    return 0;
  }
}

int cv2_estimator_forward(cv2_engine* h, void* stream, const float* x, const float* mask, const float* mu, const float* t,
                          const float* spks, const float* cond, float* out, int B2, int T, int streaming, void* workspace,
                          size_t workspace_bytes) {
  CV2_API_BEGIN
  CV2_CHECK(h && h->e.has_flow, "engine not finalized for flow");
  CV2_CHECK(x && mask && mu && t && spks && cond && out && workspace, "null argument");
  use_device(h);
  Arena ws;
  ws.base = static_cast<uint8_t*>(workspace);
  ws.cap = workspace_bytes;
  EstArgs a{x, mask, mu, t, spks, cond, out, B2, T, streaming};
  h->e.launches = 0;
  estimator_forward(h->e, (cudaStream_t)stream, a, ws);
  CV2_API_END
}

size_t cv2_flow_workspace_bytes(cv2_engine* h, int B, int max_tok_total, int n_steps) {
  try {
    Arena ws;
    FlowArgs a;
    memset(&a, 0, sizeof(a));
    a.B = B; a.max_tok_total = max_tok_total; a.n_steps = n_steps; a.finalize = 1;
    return flow_forward(h->e, nullptr, a, ws) + 4096;
  } catch (const std::exception& ex) {
    last_error_ref() = ex.what();
    return 0;
  }
}

int cv2_flow_forward(cv2_engine* h, void* stream, const int32_t* token, int token_stride, const int32_t* token_len,
                     const int32_t* prompt_token, int prompt_stride, const int32_t* prompt_len, const float* prompt_feat,
                     long long prompt_feat_bstride, const int32_t* prompt_feat_len, const float* embedding,
                     const float* rand_noise, int noise_stride, int B, int max_tok_total, int streaming, int finalize,
                     const float* t_steps_dev, const float* dt_steps_host, int n_steps, float cfg_rate, float* mel_out,
                     int mel_out_T, float* mu_out, float* enc_out, void* workspace, size_t workspace_bytes) {
  CV2_API_BEGIN
  CV2_CHECK(h && h->e.has_flow, "engine not finalized for flow");
  CV2_CHECK(token && token_len && prompt_token && prompt_len && prompt_feat && prompt_feat_len && embedding && rand_noise &&
                t_steps_dev && dt_steps_host && mel_out && workspace,
            "null argument");
  CV2_CHECK(B >= 1 && max_tok_total >= (finalize ? 1 : 4) && n_steps >= 1,   // a non-final chunk keeps 3 look-ahead tokens back
            "bad sizes B=%d max_tok_total=%d n_steps=%d", B, max_tok_total, n_steps);
  CV2_CHECK(2 * max_tok_total <= noise_stride, "sequence of %d mel frames exceeds the CFM noise buffer (%d)", 2 * max_tok_total,
            noise_stride);
  use_device(h);
  Arena ws;
  ws.base = static_cast<uint8_t*>(workspace);
  ws.cap = workspace_bytes;
  FlowArgs a;
  memset(&a, 0, sizeof(a));
  a.token = token; a.token_stride = token_stride; a.token_len = token_len;
  a.prompt_token = prompt_token; a.prompt_stride = prompt_stride; a.prompt_len = prompt_len;
  a.prompt_feat = prompt_feat; a.prompt_feat_bstride = prompt_feat_bstride; a.prompt_feat_len = prompt_feat_len;
  a.embedding = embedding; a.rand_noise = rand_noise; a.noise_stride = noise_stride;
  a.B = B; a.max_tok_total = max_tok_total; a.streaming = streaming; a.finalize = finalize;
  a.t_steps = t_steps_dev; a.dt_steps = dt_steps_host; a.n_steps = n_steps; a.cfg = cfg_rate;
  a.mel_out = mel_out; a.mel_out_T = mel_out_T; a.mu_out = mu_out; a.enc_out = enc_out;
  h->e.launches = 0;
  flow_forward(h->e, (cudaStream_t)stream, a, ws);
  CV2_API_END
}

size_t cv2_stream_state_bytes(int n_slots, int T_cap, int n_steps) {
  try {
    return stream_state_bytes(n_slots, T_cap, n_steps);
  } catch (const std::exception& ex) {
    last_error_ref() = ex.what();
    return 0;
  }
}

int cv2_stream_state_reset_slot(void* stream, void* state, size_t state_bytes, int n_slots, int T_cap, int n_steps, int slot) {
  CV2_API_BEGIN
  StreamState ss = stream_state_carve(state, state_bytes, n_slots, T_cap, n_steps);
  CV2_CHECK(slot >= 0 && slot < n_slots, "slot %d out of range", slot);
  CV2_CUDA(cudaMemsetAsync(ss.t_done + slot, 0, sizeof(int), (cudaStream_t)stream));
  CV2_API_END
}

size_t cv2_flow_stream_workspace_bytes(cv2_engine* h, int n_slots, int T_cap, int n_steps) {
  try {
    Arena ws;
    StreamState ss;
    memset(&ss, 0, sizeof(ss));
    ss.n_slots = n_slots; ss.T_cap = T_cap; ss.n_steps = n_steps;
    CV2_CHECK(T_cap >= 128 && T_cap % 128 == 0, "T_cap must be a multiple of 128");
    FlowArgs a;
    memset(&a, 0, sizeof(a));
    a.B = n_slots; a.max_tok_total = T_cap / 2 + 3; a.n_steps = n_steps; a.streaming = 1; a.stream_state = &ss;
    return flow_forward(h->e, nullptr, a, ws) + 4096;
  } catch (const std::exception& ex) {
    last_error_ref() = ex.what();
    return 0;
  }
}

int cv2_flow_forward_stream(cv2_engine* h, void* stream, const int32_t* token, int token_stride, const int32_t* token_len,
                            const int32_t* prompt_token, int prompt_stride, const int32_t* prompt_len, const float* prompt_feat,
                            long long prompt_feat_bstride, const int32_t* prompt_feat_len, const float* embedding,
                            const float* rand_noise, int noise_stride, int n_slots, int max_tok_total, const float* t_steps_dev,
                            const float* dt_steps_host, int n_steps, float cfg_rate, float* mel_out, int mel_out_T, void* state,
                            size_t state_bytes, int T_cap, void* workspace, size_t workspace_bytes) {
  CV2_API_BEGIN
  CV2_CHECK(h && h->e.has_flow, "engine not finalized for flow");
  CV2_CHECK(token && token_len && prompt_token && prompt_len && prompt_feat && prompt_feat_len && embedding && rand_noise &&
                t_steps_dev && dt_steps_host && mel_out && workspace && state,
            "null argument");
  CV2_CHECK(n_slots >= 1 && max_tok_total >= 4 && n_steps >= 1, "bad sizes");
  CV2_CHECK(2 * max_tok_total <= noise_stride, "sequence of %d mel frames exceeds the CFM noise buffer (%d)", 2 * max_tok_total,
            noise_stride);
  use_device(h);
  StreamState ss = stream_state_carve(state, state_bytes, n_slots, T_cap, n_steps);
  Arena ws;
  ws.base = static_cast<uint8_t*>(workspace);
  ws.cap = workspace_bytes;
  FlowArgs a;
  memset(&a, 0, sizeof(a));
  a.token = token; a.token_stride = token_stride; a.token_len = token_len;
  a.prompt_token = prompt_token; a.prompt_stride = prompt_stride; a.prompt_len = prompt_len;
  a.prompt_feat = prompt_feat; a.prompt_feat_bstride = prompt_feat_bstride; a.prompt_feat_len = prompt_feat_len;
  a.embedding = embedding; a.rand_noise = rand_noise; a.noise_stride = noise_stride;
  a.B = n_slots; a.max_tok_total = max_tok_total; a.streaming = 1; a.finalize = 0;
  a.t_steps = t_steps_dev; a.dt_steps = dt_steps_host; a.n_steps = n_steps; a.cfg = cfg_rate;
  a.mel_out = mel_out; a.mel_out_T = mel_out_T;
  a.stream_state = &ss;
  h->e.launches = 0;
  flow_forward(h->e, (cudaStream_t)stream, a, ws);
  CV2_API_END
}

size_t cv2_encoder_workspace_bytes(cv2_engine* h, int B, int T, int with_context) {
  try {
    Arena ws;
    FlowArgs a;
    memset(&a, 0, sizeof(a));
    a.B = B; a.enc_only = 1; a.enc_T = T; a.max_tok_total = T + (with_context ? 3 : 0); a.n_steps = 1; a.finalize = !with_context;
    return flow_forward(h->e, nullptr, a, ws) + 4096;
  } catch (const std::exception& ex) {
    last_error_ref() = ex.what();
    return 0;
  }
}

int cv2_encoder_forward(cv2_engine* h, void* stream, const float* xs, int T, const int32_t* xs_lens, const float* context,
                        int streaming, float* out, int B, void* workspace, size_t workspace_bytes) {
  CV2_API_BEGIN
  CV2_CHECK(h && h->e.has_flow, "engine not finalized for flow");
  CV2_CHECK(xs && xs_lens && out && workspace, "null argument");
  CV2_CHECK(B >= 1 && T >= 1, "bad sizes B=%d T=%d", B, T);
  use_device(h);
  Arena ws;
  ws.base = static_cast<uint8_t*>(workspace);
  ws.cap = workspace_bytes;
  FlowArgs a;
  memset(&a, 0, sizeof(a));
  a.B = B; a.enc_only = 1; a.enc_xs = xs; a.enc_T = T; a.enc_lens = xs_lens; a.enc_ctx = context;
  a.max_tok_total = T + (context ? 3 : 0);
  a.streaming = streaming; a.finalize = context == nullptr; a.n_steps = 1; a.enc_out = out;
  h->e.launches = 0;
  flow_forward(h->e, (cudaStream_t)stream, a, ws);
  CV2_API_END
}

size_t cv2_hift_workspace_bytes(cv2_engine* h, int B, int mel_T) {
  try {
    Arena ws;
    HiftArgs a;
    memset(&a, 0, sizeof(a));
    a.B = B; a.mel_T = mel_T;
    return hift_forward(h->e, nullptr, a, ws) + 4096;
  } catch (const std::exception& ex) {
    last_error_ref() = ex.what();
    return 0;
  }
}

static int hift_forward_impl(cv2_engine* h, void* stream, const float* mel, int mel_T, const int32_t* lens, const float* cache_source,
                             int cache_len, const float* noise, unsigned long long seed, float* speech, float* source, float* f0_out,
                             int16_t* pcm16, int B, void* workspace, size_t workspace_bytes) {
  CV2_API_BEGIN
  CV2_CHECK(h && h->e.has_hift, "engine not finalized for hift");
  CV2_CHECK(mel && speech && source && workspace, "null argument");
  CV2_CHECK(B >= 1 && mel_T >= 1, "bad sizes");
  use_device(h);
  Arena ws;
  ws.base = static_cast<uint8_t*>(workspace);
  ws.cap = workspace_bytes;
  HiftArgs a;
  memset(&a, 0, sizeof(a));
  a.mel = mel; a.mel_T = mel_T; a.lens = lens; a.cache_source = cache_source; a.cache_len = cache_source ? cache_len : 0;
  a.noise = noise; a.seed = seed; a.speech = speech; a.source = source; a.f0_out = f0_out; a.pcm16 = pcm16; a.B = B;
  h->e.launches = 0;
  hift_forward(h->e, (cudaStream_t)stream, a, ws);
  CV2_API_END
}

int cv2_hift_forward(cv2_engine* h, void* stream, const float* mel, int mel_T, const int32_t* lens, const float* cache_source,
                     int cache_len, const float* noise, unsigned long long seed, float* speech, float* source, float* f0_out,
                     int B, void* workspace, size_t workspace_bytes) {
  return hift_forward_impl(h, stream, mel, mel_T, lens, cache_source, cache_len, noise, seed, speech, source, f0_out, nullptr, B,
                           workspace, workspace_bytes);
}

int cv2_hift_forward_pcm16(cv2_engine* h, void* stream, const float* mel, int mel_T, const int32_t* lens, const float* cache_source,
                           int cache_len, const float* noise, unsigned long long seed, float* speech, float* source, float* f0_out,
                           int16_t* pcm16, int B, void* workspace, size_t workspace_bytes) {
  return hift_forward_impl(h, stream, mel, mel_T, lens, cache_source, cache_len, noise, seed, speech, source, f0_out, pcm16, B,
                           workspace, workspace_bytes);
}

int cv2_crossfade(void* stream, float* speech, const float* old_tail, const double* window, int n) {
  CV2_API_BEGIN
  launch_crossfade(speech, old_tail, window, n, (cudaStream_t)stream);
  CV2_API_END
}

int cv2_mel_time_stretch(void* stream, const float* mel, int T_in, float* out, int T_out, int rows) {
  CV2_API_BEGIN
  CV2_CHECK(mel && out && T_in >= 1 && T_out >= 1 && rows >= 1, "cv2_mel_time_stretch: bad arguments");
  launch_mel_time_stretch(mel, T_in, out, T_out, rows, (cudaStream_t)stream);
  CV2_API_END
}

int cv2_prompt_mel_frames(int n_samples) { return prompt_mel_frames(n_samples); }

size_t cv2_prompt_mel_workspace_bytes(int B, int max_samples) { return prompt_mel_workspace_bytes(B, max_samples); }

int cv2_prompt_mel(void* stream, const float* wav, long long wav_stride, const int32_t* n_samples, int B, int max_samples,
                   float* mel, int32_t* mel_len, void* workspace, size_t workspace_bytes) {
  CV2_API_BEGIN
  CV2_CHECK(wav && n_samples && mel && workspace, "cv2_prompt_mel: null pointer");
  launch_prompt_mel(wav, wav_stride, n_samples, B, max_samples, mel, mel_len, workspace, workspace_bytes, (cudaStream_t)stream);
  CV2_API_END
}

int cv2_kaldi_fbank_frames(int n_samples) { return kaldi_fbank_frames(n_samples); }

int cv2_kaldi_fbank(void* stream, const float* wav16, long long wav_stride, const int32_t* n_samples, int B, int max_samples,
                    float* feat, int32_t* feat_len, int subtract_mean) {
  CV2_API_BEGIN
  CV2_CHECK(wav16 && n_samples && feat, "cv2_kaldi_fbank: null pointer");
  launch_kaldi_fbank(wav16, wav_stride, n_samples, B, max_samples, feat, feat_len, subtract_mean, (cudaStream_t)stream);
  CV2_API_END
}

int cv2_resample_16k_24k_len(int n_in) { return resample_16k_24k_len(n_in); }

int cv2_resample_16k_24k(void* stream, const float* wav16, long long in_stride, const int32_t* n_in, int B, int max_in, float* wav24,
                         long long out_stride, int32_t* n_out) {
  CV2_API_BEGIN
  CV2_CHECK(wav16 && n_in && wav24, "cv2_resample_16k_24k: null pointer");
  launch_resample_16k_24k(wav16, in_stride, n_in, B, max_in, wav24, out_stride, n_out, (cudaStream_t)stream);
  CV2_API_END
}

int cv2_op_gemm_tap(void* stream, const void* A, int S, int T_alloc, int Kc, long long ldA, const void* W, int N, int Ktot,
                    const float* bias, int bn, int ntaps, const int* tap_off_host, const int32_t* lens, int len_all,
                    const float* ln_g, const float* ln_b, float ln_eps, int act, float act_f, const float* act_a,
                    const float* rowvec, int rowvec_ld, int mask_pre_res, const float* res, float out_scale, float* out32,
                    int out32_accum, void* out16, const float* ln2_g, const float* ln2_b, void* out16_ln) {
  CV2_API_BEGIN
  Engine e;
  Weight w;
  w.w = static_cast<const __half*>(W);
  w.b = bias;
  w.N = N;
  w.Ktot = Ktot;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.lens = lens; p.len_all = len_all; p.halo = 32;
  p.ln = ln_g != nullptr; p.ln_g = ln_g; p.ln_b = ln_b; p.ln_eps = ln_eps;
  p.act = act; p.act_f = act_f; p.act_a = act_a;
  p.rowvec = rowvec; p.rowvec_ld = rowvec_ld;
  p.mask_pre_res = mask_pre_res;
  p.res = res; p.res_ld = N;
  p.out_scale = out_scale;
  p.out32 = out32; p.out32_ld = N; p.out32_accum = out32_accum;
  int ne = 0;
  if (out16) {
    p.emit[ne].kind = EMIT_PLAIN; p.emit[ne].ptr = static_cast<__half*>(out16); p.emit[ne].ld = N; p.emit[ne].scale = 1.f;
    ne++;
  }
  if (out16_ln) {
    p.emit[ne].kind = EMIT_LN; p.emit[ne].ptr = static_cast<__half*>(out16_ln); p.emit[ne].ld = N; p.emit[ne].a = ln2_g;
    p.emit[ne].b = ln2_b; p.emit[ne].f = 1e-5f; p.emit[ne].scale = 1.f;
    ne++;
  }
  e.gemm((cudaStream_t)stream, static_cast<const __half*>(A), S, T_alloc, Kc, ldA, w, bn, ntaps, tap_off_host, p, false);
  CV2_API_END
}

int cv2_op_flash_attn(void* stream, const void* q, const void* k, const void* vt, void* out, const int32_t* lens, int len_all,
                      int S, int heads, int T_alloc, int chunk) {
  CV2_API_BEGIN
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.q = static_cast<const __half*>(q); p.k = static_cast<const __half*>(k); p.vt = static_cast<const __half*>(vt);
  p.out = static_cast<__half*>(out); p.lens = lens; p.len_all = len_all; p.S = S; p.heads = heads; p.T_alloc = T_alloc;
  p.chunk = chunk; p.halo = 32;
  p.reverse_seq = getenv("CV2_ATTN_FWD_ORDER") == nullptr;
  launch_flash_attn(p, (cudaStream_t)stream);
  CV2_API_END
}

int cv2_op_rel_attn(void* stream, const void* qu, const void* qv, const void* k, const void* vt, const void* pos, void* out,
                    const int32_t* lens, int len_all, int S, int T_alloc, int Tmax, int R_alloc, int chunk) {
  CV2_API_BEGIN
  RelAttnParams p;
  memset(&p, 0, sizeof(p));
  p.qu = static_cast<const __half*>(qu); p.qv = static_cast<const __half*>(qv); p.k = static_cast<const __half*>(k);
  p.vt = static_cast<const __half*>(vt); p.pos = static_cast<const __half*>(pos); p.out = static_cast<__half*>(out); p.lens = lens;
  p.len_all = len_all; p.S = S; p.T_alloc = T_alloc; p.Tmax = Tmax; p.R_alloc = R_alloc; p.chunk = chunk; p.halo = 32;
  launch_rel_attn(p, (cudaStream_t)stream);
  CV2_API_END
}

int cv2_op_source_stft(void* stream, const float* src, int mel_T, const int32_t* lens, float* out, int F_alloc, int B) {
  CV2_API_BEGIN
  launch_source_stft(src, (long long)480 * mel_T, lens, mel_T, out, F_alloc, B, (cudaStream_t)stream);
  CV2_API_END
}

int cv2_op_istft(void* stream, const float* cp, int F_alloc, const int32_t* lens, int mel_T, float* wav, int B) {
  CV2_API_BEGIN
  launch_istft(cp, F_alloc, 18, lens, mel_T, wav, (long long)480 * mel_T, B, mel_T, (cudaStream_t)stream);
  CV2_API_END
}

int cv2_op_nsf_source(void* stream, const float* f0, int mel_T, const int32_t* lens, const float* noise, unsigned long long seed,
                      const float* lw, const float* lb, float* phase_ws, float* src, int B) {
  CV2_API_BEGIN
  launch_nsf_source(f0, mel_T, phase_ws, mel_T, lens, mel_T, noise, (long long)480 * mel_T * 9, seed, nullptr, lw, lb, nullptr, 0, 0, src,
                    (long long)480 * mel_T, B, mel_T, (cudaStream_t)stream);
  CV2_API_END
}

}  // extern "C"
