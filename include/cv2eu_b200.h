/* cv2eu_b200 -- C ABI of the B200-native token2wav engine for CosyVoice2-EU.
 *
 * Plain C, raw device pointers + a CUDA stream, no torch types: this is the same convention as the reference's one
 * existing FFI on this path, the TensorRT estimator slot (cosyvoice/flow/flow_matching.py:129-150: set_tensor_address
 * on caller-owned contiguous NCT fp32 buffers, execute_async_v3(current_stream); I/O names and shapes from
 * cosyvoice/bin/export_onnx.py:89-109; context pool cosyvoice/utils/common.py:171-186).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; cv2_last_error() gives the message (thread local);
 *   - the library never allocates caller-visible device memory: weights are registered by name as device pointers that
 *     the caller keeps alive, activations live in a caller-provided, ZERO-INITIALISED workspace whose size is queried
 *     with the *_workspace_bytes functions (same arguments as the forward);
 *   - all work is enqueued on `stream` (a cudaStream_t); no host synchronisation, CUDA-graph capturable;
 *   - one in-flight call per (engine, workspace); concurrency = one workspace per stream.
 *   - there is no CPU fallback: without an sm_100a device every forward fails.
 */
#ifndef CV2EU_B200_H_
#define CV2EU_B200_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cv2_engine cv2_engine;

enum { CV2_F32 = 0, CV2_F16 = 1, CV2_I32 = 2 };

const char* cv2_last_error(void);
int cv2_version(void);

/* ---- engine / weights (replaces flow.load_state_dict / hift.load_state_dict, cosyvoice/cli/model.py:85-90) ---- */
int cv2_engine_create(cv2_engine** out, int device);
void cv2_engine_destroy(cv2_engine* e);
/* register one packed tensor (names: cosyvoice2_eu_b200/pack.py) */
int cv2_engine_set_tensor(cv2_engine* e, const char* name, const void* dptr, int dtype, int ndim, const int64_t* shape);
/* validates that every tensor the flow and/or hift forward needs is present */
int cv2_engine_finalize(cv2_engine* e, int need_flow, int need_hift);
/* kernels launched by the most recent forward on this engine */
long long cv2_engine_last_launches(cv2_engine* e);

/* per-launch CUDA-event timing on the launching stream, summed per kernel family (bench.py roofline):
 * 0 gemm<64> 1 gemm<128> 2 gemm<256> 3 flash_attn 4 rel_attn 5 f0_conv 6 nsf_source 7 source_stft 8 source_down 9 istft 10 layernorm */
/* optional device-resident seed of the in-kernel NSF noise generator (read at kernel run time, so a captured CUDA graph
 * draws fresh noise on every replay); NULL restores the by-value `seed` argument of cv2_hift_forward */
int cv2_engine_set_seed_ptr(cv2_engine* e, const unsigned long long* seed_dev);
/* engine switches (tests / A-B measurements): "fuse_euler" (CFG combine + Euler update inside final_proj, default 1),
 * "fuse_ffn" (FF1+GELU+FF2 in one kernel, default 1) */
int cv2_engine_set_option(cv2_engine* e, const char* name, int value);
/* fp16 range telemetry (opt in with cv2_engine_set_option(e, "range_check", 1); costs one scan kernel per launch): the MMA
 * operands are fp16, which saturates at 65504 -- after any forward(s), max_abs[i] = largest |x| written to a 16-bit tensor of
 * category i since the last read (0 GEMM emits of the flow, 1 q, 2 k, 3 v, 4 attention output, 5 FFN emits, 6 vocoder emits);
 * the read synchronises the device and resets the maxima.  Other options: "min_2sm_tiles" (row tiles at which the
 * cta_group::2 kernels take over, default 148). */
int cv2_engine_read_ranges(cv2_engine* e, float* max_abs, int n);
/* debug: CTA 0 of the following ffn_fused launches logs (clock64 << 8 | event) records into dev_buf[8192] (profiles/ffn_trace.py) */
int cv2_debug_set_ffn_trace(long long* dev_buf);
int cv2_engine_set_profiling(cv2_engine* e, int on);
int cv2_engine_read_profile(cv2_engine* e, double* ms_per_family, long long* launches_per_family, int n_families);

/* ---- estimator: CausalConditionalDecoder.forward (cosyvoice/flow/decoder.py:405-494), the reference's TRT slot
 *      (flow_matching.py:125-150).  x, mu, cond: [B2,80,T]; mask: [B2,1,T] (prefix of ones); t: [B2]; spks: [B2,80];
 *      out: [B2,80,T] (may alias x).  All fp32, contiguous, on the device. ---- */
size_t cv2_estimator_workspace_bytes(cv2_engine* e, int B2, int T);
int cv2_estimator_forward(cv2_engine* e, void* stream, const float* x, const float* mask, const float* mu, const float* t,
                          const float* spks, const float* cond, float* out, int B2, int T, int streaming, void* workspace,
                          size_t workspace_bytes);

/* ---- flow: CausalMaskedDiffWithXvec.inference (cosyvoice/flow/flow.py:235-283), batched over B utterances.
 *      token [B,token_stride] i32, token_len [B] i32, prompt_token [B,prompt_stride] i32, prompt_len [B] i32,
 *      prompt_feat [B, prompt_feat_bstride] f32 (each utterance [2P,80] row-major), prompt_feat_len [B] i32,
 *      embedding [B,192] f32, rand_noise [80, noise_stride] f32 (CausalConditionalCFM.rand_noise),
 *      t_steps [n_steps] f32 ON THE DEVICE (running t of solve_euler), dt_steps [n_steps] f32 ON THE HOST,
 *      mel_out [B,80,mel_out_T] f32: frames after the prompt, zero padded.  max_tok_total = max_b(prompt_len+token_len)
 *      is the only host-side length (grid sizing); everything else stays on the device.
 *      mu_out (optional, [B,80,2*max_tok_total]) and enc_out (optional, [B,2*max_tok_total,512]) expose
 *      intermediates for parity tests. ---- */
size_t cv2_flow_workspace_bytes(cv2_engine* e, int B, int max_tok_total, int n_steps);
int cv2_flow_forward(cv2_engine* e, void* stream, const int32_t* token, int token_stride, const int32_t* token_len,
                     const int32_t* prompt_token, int prompt_stride, const int32_t* prompt_len, const float* prompt_feat,
                     long long prompt_feat_bstride, const int32_t* prompt_feat_len, const float* embedding,
                     const float* rand_noise, int noise_stride, int B, int max_tok_total, int streaming, int finalize,
                     const float* t_steps_dev, const float* dt_steps_host, int n_steps, float cfg_rate, float* mel_out,
                     int mel_out_T, float* mu_out, float* enc_out, void* workspace, size_t workspace_bytes);

/* ---- incremental streaming (replaces the prefix recomputation of CosyVoice2Model.tts,
 *      cosyvoice/cli/model.py:351-381, for the NON-FINAL chunks; the final chunk runs with full attention through cv2_flow_forward
 *      exactly as the reference does, model.py:373-380).  A stream state serves `n_slots` concurrent sessions; it is caller-owned
 *      device memory of cv2_stream_state_bytes(), ZERO-INITIALISED once.  It holds, per Euler step and transformer block, the
 *      k / v^T rows of everything computed so far and, per Euler step and causal conv, the two input rows in front of every
 *      128-row tile boundary; a chunk call then computes only the row tiles from floor(rows_done / 128) * 128 on.
 *      cv2_flow_forward_stream takes the same per-utterance inputs as cv2_flow_forward with B = n_slots (row b = slot b; a slot
 *      with token_len 0 sits this call out).  T_cap (multiple of 128) bounds the mel frames of a non-final chunk call
 *      (2 * (prompt + visible tokens - 3) <= T_cap).  mel_out [n_slots,80,mel_out_T]: frames after the prompt; frames in
 *      front of those that are NEW in this call (everything before 2 * token_offset, which token2wav drops anyway,
 *      model.py:311) are unspecified.  cv2_stream_state_reset_slot starts a new session in a slot. ---- */
size_t cv2_stream_state_bytes(int n_slots, int T_cap, int n_steps);
int cv2_stream_state_reset_slot(void* stream, void* state, size_t state_bytes, int n_slots, int T_cap, int n_steps, int slot);
size_t cv2_flow_stream_workspace_bytes(cv2_engine* e, int n_slots, int T_cap, int n_steps);
int cv2_flow_forward_stream(cv2_engine* e, void* stream, const int32_t* token, int token_stride, const int32_t* token_len,
                            const int32_t* prompt_token, int prompt_stride, const int32_t* prompt_len, const float* prompt_feat,
                            long long prompt_feat_bstride, const int32_t* prompt_feat_len, const float* embedding,
                            const float* rand_noise, int noise_stride, int n_slots, int max_tok_total, const float* t_steps_dev,
                            const float* dt_steps_host, int n_steps, float cfg_rate, float* mel_out, int mel_out_T, void* state,
                            size_t state_bytes, int T_cap, void* workspace, size_t workspace_bytes);

/* ---- encoder slot (boundary #5): UpsampleConformerEncoder.forward (cosyvoice/transformer/upsample_encoder.py:243-306), the
 *      module CosyVoice2Model.load_jit swaps in as `flow.encoder` (cosyvoice/cli/model.py:285-287), called by flow.inference as
 *      `encoder(token, token_len, context=..., streaming=...)` (cosyvoice/flow/flow.py:258-263).
 *      xs [B,T,512] f32 = input_embedding(token) * mask; xs_lens [B] i32 on the device (values above T are clamped, which is
 *      what make_pad_mask(xs_lens, T) does to flow.py's `token_len` that still counts the 3 context tokens);
 *      context [B,3,512] f32 = the look-ahead embeddings of a non-final chunk, or NULL; out [B,2T,512] f32.
 *      Rows of `out` at or beyond 2*xs_lens[b] are padding (unspecified, finite). ---- */
size_t cv2_encoder_workspace_bytes(cv2_engine* e, int B, int T, int with_context);
int cv2_encoder_forward(cv2_engine* e, void* stream, const float* xs, int T, const int32_t* xs_lens, const float* context,
                        int streaming, float* out, int B, void* workspace, size_t workspace_bytes);

/* ---- hift: HiFTGenerator.inference (cosyvoice/hifigan/generator.py:570-582), batched.
 *      mel [B,80,mel_T] f32; lens [B] i32 valid frames (NULL: all mel_T); cache_source [B,1,cache_len] or NULL;
 *      noise [B,480*mel_T,9] f32 = the N(0,1) draw of SineGen2 (generator.py:334) for parity runs, NULL -> in-kernel
 *      counter-based generator seeded with `seed`; speech [B,480*mel_T]; source [B,1,480*mel_T]; f0_out optional [B,mel_T]. */
size_t cv2_hift_workspace_bytes(cv2_engine* e, int B, int mel_T);
int cv2_hift_forward(cv2_engine* e, void* stream, const float* mel, int mel_T, const int32_t* lens, const float* cache_source,
                     int cache_len, const float* noise, unsigned long long seed, float* speech, float* source, float* f0_out,
                     int B, void* workspace, size_t workspace_bytes);

/* Same, plus the servers' wire format fused into the iSTFT epilogue: pcm16 [B,480*mel_T] int16 = (speech * 2**15) truncated
 * toward zero, bit-exact with `(tts_speech.numpy() * (2 ** 15)).astype(np.int16)` (runtime/python/fastapi/server.py:42,
 * runtime/python/grpc/server.py:68). */
int cv2_hift_forward_pcm16(cv2_engine* e, void* stream, const float* mel, int mel_T, const int32_t* lens, const float* cache_source,
                           int cache_len, const float* noise, unsigned long long seed, float* speech, float* source, float* f0_out,
                           int16_t* pcm16, int B, void* workspace, size_t workspace_bytes);

/* ---- streaming glue: fade_in_out (cosyvoice/utils/common.py:142-150) on the device.  window: [2n] f64 device. ---- */
int cv2_crossfade(void* stream, float* speech, const float* old_tail, const double* window, int n);

/* speed change of an offline utterance (cosyvoice/cli/model.py:325-327): out[r, :] = F.interpolate(mel[r, :], size = T_out,
 * mode = 'linear') for `rows` = B * 80 rows; T_out = int(T_in / speed).  mel / out: fp32 device, contiguous rows. */
int cv2_mel_time_stretch(void* stream, const float* mel, int T_in, float* out, int T_out, int rows);

/* ---- prompt features: the 24 kHz log-mel of the prompt waveform, replacing
 * matcha.utils.audio.mel_spectrogram (third_party/Matcha-TTS/matcha/utils/audio.py:45-82) at the cosyvoice2.yaml:152-160
 * settings as called by CosyVoiceFrontEnd._extract_speech_feat (cosyvoice/cli/frontend.py:285-289), for a batch of prompts.
 * wav: [B, wav_stride] fp32 device, n_samples: [B] int32 device, max_samples: the largest of them (host).
 * mel: [B, cv2_prompt_mel_frames(max_samples), 80] fp32 device (the [1, T, 80] layout flow.inference takes; rows past a
 * prompt's own frames are zero), mel_len: [B] int32 device or NULL.  A prompt of <= 720 samples cannot be reflect-padded
 * (the reference raises): cv2_prompt_mel rejects max_samples <= 720, shorter rows of a batch yield mel_len 0. ---- */
int cv2_prompt_mel_frames(int n_samples);
size_t cv2_prompt_mel_workspace_bytes(int B, int max_samples);
int cv2_prompt_mel(void* stream, const float* wav, long long wav_stride, const int32_t* n_samples, int B, int max_samples,
                   float* mel, int32_t* mel_len, void* workspace, size_t workspace_bytes);

/* x-vector features: what CosyVoiceFrontEnd._extract_spk_embedding hands the CAM++ session
 * (cosyvoice/cli/frontend.py:276-278): torchaudio.compliance.kaldi.fbank(speech, num_mel_bins=80, dither=0,
 * sample_frequency=16000) and, with subtract_mean != 0, `feat - feat.mean(dim=0)`.  wav16: [B, wav_stride] fp32 device at
 * 16 kHz, n_samples: [B] int32 device, max_samples: the largest (host, >= 400); feat: [B, cv2_kaldi_fbank_frames(max_samples), 80]
 * fp32 device (rows past an utterance's own frames are zero), feat_len: [B] int32 device or NULL. */
int cv2_kaldi_fbank_frames(int n_samples);
int cv2_kaldi_fbank(void* stream, const float* wav16, long long wav_stride, const int32_t* n_samples, int B, int max_samples,
                    float* feat, int32_t* feat_len, int subtract_mean);

/* 16 kHz -> 24 kHz resampling of the prompt, replacing torchaudio.transforms.Resample(orig_freq=16000, new_freq=24000)
 * (cosyvoice/cli/frontend.py:495,541; sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99).  wav16: [B, in_stride] fp32 device,
 * n_in: [B] int32 device, max_in: the largest (host); wav24: [B, out_stride] fp32 device with out_stride >=
 * cv2_resample_16k_24k_len(max_in) = ceil(3 max_in / 2) (rows zero-padded up to that), n_out: [B] int32 device or NULL. */
int cv2_resample_16k_24k_len(int n_in);
int cv2_resample_16k_24k(void* stream, const float* wav16, long long in_stride, const int32_t* n_in, int B, int max_in, float* wav24,
                         long long out_stride, int32_t* n_out);

/* ---- single-kernel entry points (parity tests of the individual kernels) ---- */
/* out = epilogue(sum_taps A[s, t+off, :] W^T): A 16-bit [S,T_alloc,ldA], W 16-bit [N, ntaps*ceil64(Kc)], see gemm_tap.cuh.
 * act: 0 none 1 mish 2 gelu 3 silu 4 elu 5 lrelu(act_f) 6 snake(act_a).  Optional outputs: out32 [S*T_alloc,N] fp32,
 * out16 [S*T_alloc,N] 16-bit of the same value, out16_ln = LayerNorm(value; ln2_g, ln2_b, eps 1e-5) 16-bit. */
int cv2_op_gemm_tap(void* stream, const void* A, int S, int T_alloc, int Kc, long long ldA, const void* W, int N, int Ktot,
                    const float* bias, int bn, int ntaps, const int* tap_off_host, const int32_t* lens, int len_all,
                    const float* ln_g, const float* ln_b, float ln_eps, int act, float act_f, const float* act_a,
                    const float* rowvec, int rowvec_ld, int mask_pre_res, const float* res, float out_scale, float* out32,
                    int out32_accum, void* out16, const float* ln2_g, const float* ln2_b, void* out16_ln);
/* q,k [S,H,T_alloc,64], vt [S,H,64,T_alloc] 16-bit -> out [S,T_alloc,H*64] 16-bit */
int cv2_op_flash_attn(void* stream, const void* q, const void* k, const void* vt, void* out, const int32_t* lens, int len_all,
                      int S, int heads, int T_alloc, int chunk);
/* encoder relative-position attention (RelPositionMultiHeadedAttention, cosyvoice/transformer/attention.py:249-330):
 * 16-bit qu = (q+pos_bias_u)/8, qv = (q+pos_bias_v)/8, k [S,8,T_alloc,64], vt [S,8,64,T_alloc],
 * pos [R_alloc,512] = linear_pos(table), row (rel + Tmax - 1) <-> relative position rel = i - j  -> out [S,T_alloc,512] */
int cv2_op_rel_attn(void* stream, const void* qu, const void* qv, const void* k, const void* vt, const void* pos, void* out,
                    const int32_t* lens, int len_all, int S, int T_alloc, int Tmax, int R_alloc, int chunk);
/* source STFT: src [B, 480*mel_T] -> out [B, F_alloc, 18] */
int cv2_op_source_stft(void* stream, const float* src, int mel_T, const int32_t* lens, float* out, int F_alloc, int B);
/* iSTFT head: conv_post output [B, F_alloc, 18] -> wav [B, 480*mel_T] */
int cv2_op_istft(void* stream, const float* cp, int F_alloc, const int32_t* lens, int mel_T, float* wav, int B);
/* NSF source: f0 [B, mel_T] -> src [B, 480*mel_T]; phase_ws [B, mel_T, 9] scratch */
int cv2_op_nsf_source(void* stream, const float* f0, int mel_T, const int32_t* lens, const float* noise, unsigned long long seed,
                      const float* lw, const float* lb, float* phase_ws, float* src, int B);

#ifdef __cplusplus
}
#endif
#endif /* CV2EU_B200_H_ */
