// Host-side engine: owns nothing on the device.  Weights are registered by name as raw device pointers
// (torch-owned), activations live in a caller-provided workspace carved by a bump allocator, and every
// forward is a pure sequence of kernel launches on the caller's stream (CUDA-graph capturable, no host sync).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "ffn_fused.cuh"
#include "gemm_tap.cuh"
#include "host_util.h"

namespace cv2 {

enum DType { DT_F32 = 0, DT_F16 = 1, DT_I32 = 2 };

struct TensorRef {
  const void* ptr = nullptr;
  int dtype = 0;
  std::vector<int64_t> shape;
  int64_t numel() const {
    int64_t n = 1;
    for (auto d : shape) n *= d;
    return n;
  }
};

// Bump allocator over the caller's workspace.  With base == nullptr it only measures.
struct Arena {
  uint8_t* base = nullptr;
  size_t cap = 0, off = 0, peak = 0;
  void* alloc(size_t bytes) {
    off = (off + 255) & ~size_t(255);
    size_t o = off;
    off += bytes;
    if (off > peak) peak = off;
    if (base && off > cap) fail("workspace too small: need > %zu bytes, have %zu", off, cap);
    return base ? base + o : reinterpret_cast<void*>(uintptr_t(256) + o);  // fake non-null pointers when measuring
  }
  template <class T>
  T* get(size_t n) { return reinterpret_cast<T*>(alloc(n * sizeof(T))); }
  bool measuring() const { return base == nullptr; }
  // scoped reuse: everything allocated after mark() is handed back by rewind(mark) (peak keeps the high-water mark)
  size_t mark() const { return off; }
  void rewind(size_t m) { off = m; }
};

// A 16-bit weight matrix [N, Ktot] prepared for the tap GEMM (+ fp32 bias).
struct Weight {
  const __half* w = nullptr;
  const float* b = nullptr;
  int N = 0, Ktot = 0;
  CUtensorMap map[3];       // per BN in {64,128,256}
  bool map_ok[3] = {false, false, false};
};

struct LN {
  const float* g = nullptr;
  const float* b = nullptr;
};

struct Engine {
  int device = 0;
  std::unordered_map<std::string, TensorRef> tensors;
  bool finalized = false;
  bool has_flow = false, has_hift = false;
  bool fuse_ffn = getenv("CV2_NO_FFN_FUSION") == nullptr;   // estimator FF1+GELU+FF2 in one kernel (ffn_fused.cu)
  // F0 predictor as split-precision (hi + lo 16-bit) tensor-core GEMMs: 10x faster than the fp32 SIMT convs but measured f0
  // error 4.8e-3 Hz (fp32: 4.5e-4) -- the tensor core adds each K=16 group to the fp32 accumulator with truncation, ~2^-16
  // relative over K = 6144 -- which the phase accumulator turns into 8.7 dB waveform SNR at 8 s.  Off by default.
  bool f0_split = getenv("CV2_F0_SPLIT") != nullptr;
  bool ffn_2cta = getenv("CV2_NO_FFN_2CTA") == nullptr;      // fused FFN as CTA pairs with tcgen05.mma.cta_group::2 (ffn_fused2.cu)
  bool cluster_mc = getenv("CV2_NO_CLUSTER") == nullptr;      // BN=256 GEMMs as 2-CTA clusters sharing the weight tile by TMA multicast
  bool fuse_euler = getenv("CV2_NO_EULER_FUSION") == nullptr; // CFG combine + Euler update inside final_proj's epilogue
  const unsigned long long* seed_dev = nullptr;   // optional device-resident NSF noise seed (CUDA-graph replays)
  std::unordered_map<std::string, Weight> weights;      // lazily built from tensors
  std::unordered_map<uint64_t, CUtensorMap> amap_cache;  // activation tensor maps
  long long launches = 0;                                // kernels launched by the last forward (claims for bench)
  // the 2-SM (cta_group::2) GEMM / FFN forms switch on at this many row tiles
  int min_2sm_tiles = getenv("CV2_MIN_2SM_TILES") ? atoi(getenv("CV2_MIN_2SM_TILES")) : 148;
  // ... and the fused FFN (whose 2-SM form also carries the chained out-projection: one launch less per transformer block,
  // which is what a launch-latency-bound small batch needs) already at this many
  int min_2sm_tiles_ffn = getenv("CV2_MIN_2SM_TILES_FFN") ? atoi(getenv("CV2_MIN_2SM_TILES_FFN")) : 2;

  // ---- opt-in fp16 range telemetry (option "range_check"): operands are fp16 (saturates at 65504), so a trained checkpoint
  //      can be checked for headroom: after every launch the 16-bit tensors it wrote are scanned for their max |x| ----
  enum RangeCat { R_GEMM_EMIT = 0, R_Q, R_K, R_V, R_ATTN_OUT, R_FFN_EMIT, R_HIFT_EMIT, R_COUNT };
  bool range_check = false;
  bool in_hift = false;
  unsigned* range_dev = nullptr;   // [R_COUNT] fp32 bit patterns of the running max |x| (device, owned by the engine)
  void range_scan(cudaStream_t st, int cat, const __half* ptr, long long rows, int cols, long long ld);
  void range_scan_emit(cudaStream_t st, int cat, const Emit& em, long long rows, int cols) {
    if (em.kind != EMIT_NONE && em.ptr) range_scan(st, cat, em.ptr + em.col_off, rows, cols, em.ld);
  }

  // compact active-tile lists keyed by (lens pointer, T_alloc): gemm() attaches them automatically
  struct TileList { const int* list; const int* count; };
  std::unordered_map<uint64_t, TileList> tile_lists;
  static uint64_t tl_key(const int* lens, int T_alloc, int S) {
    return ((uint64_t)(uintptr_t)lens * 1315423911ull + (uint64_t)T_alloc) * 2654435761ull + (uint64_t)S;
  }
  // builds the list for `lens` (S sequences) and registers it under lens and every alias pointer (same or shorter lengths)
  void make_tile_list(cudaStream_t st, Arena& ws, const int* lens, int S, int T_alloc, int halo, bool dry,
                      const int* alias0 = nullptr, const int* alias1 = nullptr, const int* lo = nullptr);

  // ---- optional per-launch timing (CUDA events on the launching stream), grouped by kernel family ----
  enum Family { F_GEMM64 = 0, F_GEMM128, F_GEMM256, F_FLASH_ATTN, F_REL_ATTN, F_F0_CONV, F_NSF, F_STFT, F_SRC_DOWN, F_ISTFT,
                F_FFN_FUSED, F_COUNT };
  bool profiling = false;
  struct ProfRec { int family; cudaEvent_t a, b; };
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  cudaEvent_t next_event() {
    if (ev_used == ev_pool.size()) {
      cudaEvent_t ev;
      CV2_CUDA(cudaEventCreate(&ev));
      ev_pool.push_back(ev);
    }
    return ev_pool[ev_used++];
  }
  void prof_begin(cudaStream_t st, int family) {
    if (!profiling) return;
    ProfRec r{family, next_event(), next_event()};
    CV2_CUDA(cudaEventRecord(r.a, st));
    prof.push_back(r);
  }
  void prof_end(cudaStream_t st) {
    if (!profiling) return;
    CV2_CUDA(cudaEventRecord(prof.back().b, st));
  }

  const TensorRef& T(const std::string& name) const {
    auto it = tensors.find(name);
    if (it == tensors.end()) fail("engine: tensor '%s' was not registered", name.c_str());
    return it->second;
  }
  const float* f32(const std::string& name) const {
    const TensorRef& t = T(name);
    if (t.dtype != DT_F32) fail("engine: tensor '%s' must be fp32", name.c_str());
    return static_cast<const float*>(t.ptr);
  }
  Weight& W(const std::string& name);   // expects '<name>.w' (fp16 2-D) and optional '<name>.b'
  LN ln(const std::string& name) const { return LN{f32(name + "_g"), f32(name + "_b")}; }

  // ---- tap GEMM front-end ------------------------------------------------------------------------
  // A: 16-bit [S, T_alloc, ld] (first Kc columns used).  Remaining epilogue fields come in through `p`.
  void gemm(cudaStream_t st, const __half* A, int S, int T_alloc, int Kc, long long ldA, Weight& w, int bn, int ntaps,
            const int* tap_off, GemmParams p, bool dry);
  // fused FF1 -> GELU -> FF2 -> +residual -> emits (H: 16-bit [S, T_alloc, 256]).  With `op` the attention out-projection
  // (x += Wo att + bo, H = LayerNorm3(x)) belongs to the call: chained into the 2-SM kernel when that one runs, else launched
  // as the separate GEMM in front of the 1-SM kernel.
  struct OutProj {
    const __half* att;   // [S, T_alloc, 512]
    Weight* wo;          // [256, 512] (+ bias)
    LN ln3;
    float eps;
  };
  void ffn(cudaStream_t st, __half* H, int S, int T_alloc, Weight& w1, Weight& w2, FfnParams p, bool dry, const OutProj* op = nullptr);
  bool chain_outproj = getenv("CV2_NO_OUTPROJ_CHAIN") == nullptr;
  bool ffn_hsplit = getenv("CV2_NO_FFN_HSPLIT") == nullptr;   // small launches: hidden split of the chained FFN (ffn_fused2.cu)
  const CUtensorMap& amap(const __half* A, int S, int T_alloc, int Kc, long long ldA, int box_rows = 128);
  // conv mode of the tap GEMM for the vocoder's narrow stages (gemm_tap.cuh: GemmParams::tmA_halo): 0 off, 1 on
  int conv_mode = getenv("CV2_CONV_MODE") ? atoi(getenv("CV2_CONV_MODE")) : 1;
  const CUtensorMap& wmap(Weight& w, int bn);
};

// forward passes (engine_flow.cu / engine_hift.cu)
struct StreamState;
struct FlowArgs {
  const int* token; int token_stride; const int* token_len;
  const int* prompt_token; int prompt_stride; const int* prompt_len;
  const float* prompt_feat; long long prompt_feat_bstride; const int* prompt_feat_len;
  const float* embedding;       // [B,192]
  const float* rand_noise;      // [80, noise_stride]
  int noise_stride;
  int B;
  int max_tok_total;            // host-known max over b of prompt_len + token_len
  int streaming, finalize;
  const float* t_steps;         // host [n_steps]
  const float* dt_steps;        // host [n_steps]
  int n_steps;
  float cfg;
  float* mel_out;               // [B, 80, mel_out_T] NCT; frames after the prompt; zero padded
  int mel_out_T;
  float* mu_out;                // optional [B, 80, 2*T] (debug / tests) or null
  float* enc_out;               // optional [B, 2*max_tok_total, 512] (encoder slot: [B, 2*enc_T, 512]) or null
  // encoder slot (boundary #5, cv2_encoder_forward): UpsampleConformerEncoder.forward alone on caller-made embeddings
  int enc_only;
  const float* enc_xs;          // [B, enc_T, 512] fp32 (input_embedding(token) * mask), replaces the table lookup
  int enc_T;
  const int* enc_lens;          // [B] valid rows of enc_xs (values above enc_T are clamped, like make_pad_mask(xs_lens, T))
  const float* enc_ctx;         // [B, 3, 512] look-ahead context (flow.py:262-263) or null
  const StreamState* stream_state;   // non-null: incremental non-final streaming chunk over the state's slots (B == n_slots)
};
size_t flow_forward(Engine& e, cudaStream_t st, const FlowArgs& a, Arena& ws);

// ---- incremental streaming state of the CFM estimator for a group of `n_slots` concurrent sessions ----
// A non-final streaming chunk runs the estimator with block-causal attention (chunk 50) and causal convolutions only, so every
// mel row it produced is final.  The state keeps, per Euler step and transformer block, the k / v^T of all rows computed so far,
// and per Euler step and causal conv the two input rows in front of every 128-row tile boundary; the next chunk then computes
// the row tiles from floor(t_done / 128) * 128 on, instead of the whole prefix (what the reference does, model.py:351-381).
// Caller-owned, zero-initialised device memory of stream_state_bytes().
struct StreamState {
  int n_slots, T_cap, n_steps;
  int* t_done;        // [n_slots] mel rows already final per slot (0 = fresh session)
  __half* kcache;     // [n_steps][56][2*n_slots][8][T_cap][64]
  __half* vcache;     // [n_steps][56][2*n_slots][8][64][T_cap]
  __half* tails;      // [n_steps][31 convs][2*n_slots][T_cap/128 boundaries][2 rows][512]
  size_t kv_block() const { return (size_t)2 * n_slots * 8 * T_cap * 64; }
  size_t tail_block() const { return (size_t)2 * n_slots * (T_cap / 128) * 2 * 512; }
};
static const int kEstConvs = 31, kEstBlocks = 56;
size_t stream_state_bytes(int n_slots, int T_cap, int n_steps);
StreamState stream_state_carve(void* base, size_t bytes, int n_slots, int T_cap, int n_steps);
// flow_forward with FlowArgs::stream_state set: a.B must equal n_slots (inactive slots: token_len = 0); mel_out rows in front of
// the rows that are new in this call are unspecified

struct EstArgs {
  const float* x; const float* mask; const float* mu; const float* t; const float* spks; const float* cond;
  float* out;
  int B2, T, streaming;
};
size_t estimator_forward(Engine& e, cudaStream_t st, const EstArgs& a, Arena& ws);

struct HiftArgs {
  const float* mel;            // [B, 80, mel_T] NCT fp32
  int mel_T;                   // stride / max frames
  const int* lens;             // device [B] valid frames, or null (all mel_T)
  const float* cache_source;   // [B, 1, cache_len] or null
  int cache_len;
  const float* noise;          // [B, 480*mel_T, 9] or null (-> in-kernel generator with `seed`)
  unsigned long long seed;
  float* speech;               // [B, 480*mel_T]
  float* source;               // [B, 1, 480*mel_T]
  float* f0_out;               // optional [B, mel_T]
  short* pcm16;                // optional [B, 480*mel_T] int16 PCM of `speech` (servers' wire format), written by the iSTFT kernel
  int B;
};
size_t hift_forward(Engine& e, cudaStream_t st, const HiftArgs& a, Arena& ws);

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

}  // namespace cv2
