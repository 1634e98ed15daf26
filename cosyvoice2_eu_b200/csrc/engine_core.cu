// Engine plumbing shared by the flow and HiFT forwards: weight lookup, tensor-map caches, GEMM front-end.
#include "engine.h"
#include "flow_kernels.cuh"

namespace cv2 {

void Engine::range_scan(cudaStream_t st, int cat, const __half* ptr, long long rows, int cols, long long ld) {
  if (!range_check || !range_dev || !ptr) return;
  launch_absmax16(ptr, rows, cols, ld, range_dev + (in_hift && cat == R_GEMM_EMIT ? (int)R_HIFT_EMIT : cat), st);
}

Weight& Engine::W(const std::string& name) {
  auto it = weights.find(name);
  if (it != weights.end()) return it->second;
  const TensorRef& t = T(name + ".w");
  if (t.dtype != DT_F16 || t.shape.size() != 2) fail("engine: weight '%s.w' must be 16-bit 2-D", name.c_str());
  Weight w;
  w.w = static_cast<const __half*>(t.ptr);
  w.N = (int)t.shape[0];
  w.Ktot = (int)t.shape[1];
  if (w.Ktot % 64 != 0) fail("engine: weight '%s.w' K=%d is not a multiple of 64", name.c_str(), w.Ktot);
  auto bi = tensors.find(name + ".b");
  if (bi != tensors.end()) {
    if (bi->second.dtype != DT_F32 || bi->second.numel() != w.N) fail("engine: bias '%s.b' must be fp32 [%d]", name.c_str(), w.N);
    w.b = static_cast<const float*>(bi->second.ptr);
  }
  return weights.emplace(name, w).first->second;
}

static inline uint64_t mix(uint64_t h, uint64_t v) {
  h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
  return h;
}

void Engine::make_tile_list(cudaStream_t st, Arena& ws, const int* lens, int S, int T_alloc, int halo, bool dry,
                            const int* alias0, const int* alias1, const int* lo) {
  int* list = ws.get<int>((size_t)2 * S * (T_alloc / 128));
  int* count = ws.get<int>(1);
  launches++;
  if (dry) return;
  launch_build_tile_list(lens, S, T_alloc, halo, list, count, st, lo);
  TileList tl{list, count};
  tile_lists[tl_key(lens, T_alloc, S)] = tl;
  if (alias0) tile_lists[tl_key(alias0, T_alloc, S)] = tl;
  if (alias1) tile_lists[tl_key(alias1, T_alloc, S)] = tl;
}

const CUtensorMap& Engine::wmap(Weight& w, int bn) {
  const int bi = bn == 64 ? 0 : (bn == 128 ? 1 : 2);
  if (!w.map_ok[bi]) {
    uint64_t dims[2] = {(uint64_t)w.Ktot, (uint64_t)w.N};
    uint64_t strides[1] = {(uint64_t)w.Ktot * 2};
    uint32_t box[2] = {64, (uint32_t)bn};
    w.map[bi] = make_tmap_16b(w.w, 2, dims, strides, box);
    w.map_ok[bi] = true;
  }
  return w.map[bi];
}

const CUtensorMap& Engine::amap(const __half* A, int S, int T_alloc, int Kc, long long ldA, int box_rows) {
  uint64_t key = mix(mix(mix(mix(mix(mix(0x1234, (uint64_t)(uintptr_t)A), (uint64_t)Kc), (uint64_t)ldA), (uint64_t)T_alloc), (uint64_t)S),
                     (uint64_t)box_rows);
  auto it = amap_cache.find(key);
  if (it == amap_cache.end()) {
    uint64_t dims[3] = {(uint64_t)Kc, (uint64_t)T_alloc, (uint64_t)S};
    uint64_t strides[2] = {(uint64_t)ldA * 2, (uint64_t)T_alloc * (uint64_t)ldA * 2};
    uint32_t box[3] = {64, (uint32_t)box_rows, 1};
    if (amap_cache.size() > 8192) amap_cache.clear();
    it = amap_cache.emplace(key, make_tmap_16b(A, 3, dims, strides, box)).first;
  }
  return it->second;
}

void Engine::ffn(cudaStream_t st, __half* H, int S, int T_alloc, Weight& w1, Weight& w2, FfnParams p, bool dry, const OutProj* op) {
  CV2_CHECK(w1.N == 1024 && w1.Ktot == 256 && w2.N == 256 && w2.Ktot == 1024, "ffn: unexpected weight shapes");
  p.S = S;
  p.T_alloc = T_alloc;
  p.b1 = w1.b;
  p.b2 = w2.b;
  const bool two_sm = ffn_2cta && S > 1 && p.lens && (long long)S * (T_alloc / 128) >= (min_2sm_tiles < min_2sm_tiles_ffn ? min_2sm_tiles : min_2sm_tiles_ffn);
  const bool chain = op && two_sm && chain_outproj && op->wo->b;
  if (op && !chain) {   // the out-projection as its own GEMM: + bias + residual -> x32 ; emit LayerNorm3 -> H
    static const int tap1[1] = {0};
    GemmParams g;
    memset(&g, 0, sizeof(g));
    g.lens = p.lens; g.len_all = p.len_all; g.halo = p.halo; g.out_scale = 1.f;
    g.res = p.x32; g.res_ld = 256;
    g.out32 = p.x32; g.out32_ld = 256;
    g.emit[0].ptr = H; g.emit[0].ld = 256; g.emit[0].kind = EMIT_LN; g.emit[0].a = op->ln3.g; g.emit[0].b = op->ln3.b;
    g.emit[0].f = op->eps; g.emit[0].scale = 1.f;
    gemm(st, op->att, S, T_alloc, 512, 512, *op->wo, 256, 1, tap1, g, dry);
  }
  launches++;
  if (dry) return;
  if (p.lens && !p.tile_list && S > 1) {
    auto tl = tile_lists.find(tl_key(p.lens, T_alloc, S));
    if (tl != tile_lists.end()) {
      p.tile_list = tl->second.list;
      p.tile_count = tl->second.count;
    }
  }
  prof_begin(st, F_FFN_FUSED);
  if (two_sm && p.tile_list) {
    if (chain) {   // out-projection chained in front: H never exists in global memory
      p.bo = op->wo->b; p.ln3_g = op->ln3.g; p.ln3_b = op->ln3.b; p.ln3_eps = op->eps;
      if (ffn_hsplit && p.hsplit == 4 && p.slabs)   // small launch: the hidden dimension of every tile pair divided among four CTA pairs
        launch_ffn_fused2_chain_split(amap(op->att, S, T_alloc, 512, 512), wmap(*op->wo, 128), wmap(w1, 64), wmap(w2, 128), p, st);
      else {
        p.hsplit = 0;
        launch_ffn_fused2_chain(amap(op->att, S, T_alloc, 512, 512), wmap(*op->wo, 128), wmap(w1, 64), wmap(w2, 128), p, st);
      }
    } else {
      p.hsplit = 0;
      launch_ffn_fused2(amap(H, S, T_alloc, 256, 256), wmap(w1, 64), wmap(w2, 128), p, st);   // 2-SM MMAs, half of every weight tile per CTA
    }
  } else {
    CV2_CHECK(!chain, "ffn: chained out-projection without a tile list");
    launch_ffn_fused(amap(H, S, T_alloc, 256, 256), wmap(w1, 128), wmap(w2, 128), p, st);
  }
  prof_end(st);
  if (range_check) {
    const long long rows = (long long)S * T_alloc;
    range_scan_emit(st, R_FFN_EMIT, p.emit_ln, rows, 256);
    range_scan_emit(st, R_FFN_EMIT, p.emit_plain[0], rows, 256);
    range_scan_emit(st, R_FFN_EMIT, p.emit_plain[1], rows, 256);
  }
}

void Engine::gemm(cudaStream_t st, const __half* A, int S, int T_alloc, int Kc, long long ldA, Weight& w, int bn, int ntaps,
                  const int* tap_off, GemmParams p, bool dry) {
  const int kb = (Kc + 63) / 64;
  CV2_CHECK(w.Ktot == ntaps * kb * 64, "gemm: weight K=%d does not match taps=%d x Kc_pad=%d", w.Ktot, ntaps, kb * 64);
  CV2_CHECK(bn == 64 || bn == 128 || bn == 256, "gemm: bad BN %d", bn);
  p.S = S;
  p.T_alloc = T_alloc;
  p.N = w.N;
  p.kb_per_tap = kb;
  p.ntaps = ntaps;
  for (int i = 0; i < ntaps; i++) p.tap_off[i] = tap_off[i];
  if (!p.bias) p.bias = w.b;
  if (p.out_scale == 0.f) p.out_scale = 1.f;
  for (int e = 0; e < 3; e++)
    if (p.emit[e].kind != EMIT_NONE && p.emit[e].scale == 0.f) p.emit[e].scale = 1.f;
  launches++;
  if (dry) return;
  if (p.lens && !p.tile_list) {
    auto tl = tile_lists.find(tl_key(p.lens, T_alloc, S));
    if (tl != tile_lists.end() && S > 1) {
      p.tile_list = tl->second.list;
      p.tile_count = tl->second.count;
    }
  }
  const int bi = bn == 64 ? 0 : (bn == 128 ? 1 : 2);
  // 2-CTA clusters with TMA multicast of the weight operand: big launches only (>= 2 row tiles per SM pair), compact list required
  { extern long long* g_ffn_trace_ptr(); static const int tq = getenv("CV2_TRACE_QKV") != nullptr; static const int tc = getenv("CV2_TRACE_CONV") != nullptr; p.trace = (tc ? (p.ntaps == 3 && p.res != nullptr && bn == 256 && p.ln != 0) : tq ? p.q != nullptr : (p.res != nullptr && p.N == 256 && p.emit[0].kind != 0 && bn == 256 && p.ntaps == 1)) && g_ffn_trace_ptr() ? g_ffn_trace_ptr() + 8192 : nullptr; }
  { static const int dbg = getenv("CV2_DBG_SKIP_EPI") ? atoi(getenv("CV2_DBG_SKIP_EPI")) : 0; p.dbg_skip_epi = (dbg == 1) || (dbg == 2 && p.q != nullptr) || (dbg == 3 && p.res != nullptr); }
  if (cluster_mc && bn == 256 && p.tile_list && (long long)S * (T_alloc / 128) >= min_2sm_tiles) p.tmB_half = &wmap(w, 128);
  const CUtensorMap& tb = wmap(w, bn);
  const CUtensorMap ta = amap(A, p.S_map > 0 ? p.S_map : S, T_alloc, Kc, ldA);   // (copies: the cache may be flushed by the next lookup)
  CUtensorMap ta_halo;
  if (conv_mode && gemm_tap_conv_eligible(bn, p)) {   // k-tap conv of a narrow vocoder stage: one A box per k-block, resident weights
    ta_halo = amap(A, S, T_alloc, Kc, ldA, kConvHaloRows);
    p.tmA_halo = &ta_halo;
    p.conv_min_off = p.tap_off[0];
    for (int i = 1; i < ntaps; i++) p.conv_min_off = p.tap_off[i] < p.conv_min_off ? p.tap_off[i] : p.conv_min_off;
  }
  prof_begin(st, bn == 256 ? F_COUNT + gemm_tap_spec(bn, p) : bi);
  launch_gemm_tap(bn, ta, tb, p, st);
  prof_end(st);
  if (range_check && !p.flat) {
    const long long rows = (long long)S * T_alloc;
    for (int i = 0; i < 3; i++) range_scan_emit(st, R_GEMM_EMIT, p.emit[i], rows, p.N);
    const long long hd = (long long)S * p.heads * T_alloc;     // rows of 64 per head tensor
    if (p.q) range_scan(st, R_Q, p.q, hd, 64, 64);
    if (p.q2) range_scan(st, R_Q, p.q2, hd, 64, 64);
    if (p.k) range_scan(st, R_K, p.k, hd, 64, 64);
    if (p.vt) range_scan(st, R_V, p.vt, hd, 64, 64);
  }
}

}  // namespace cv2
