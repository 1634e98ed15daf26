// Relative-position multi-head attention of the UpsampleConformerEncoder
// (reference: cosyvoice/transformer/attention.py:249-330, rel_shift :225-247).
//   score[i,j] = ((q_i + u) . k_j + (q_i + v) . P[i - j]) / 8 ;  masked softmax ; . v
// The [T, 2T-1] `bd` matrix and its rel_shift are never materialised: after the shift, column j of
// row i is the positional projection of relative position (i - j), which is read straight from the
// table.  fp32 SIMT, flash-style online softmax: the encoder is ~1.5 % of the path's FLOPs and its
// output feeds a chaotic 10-step ODE, so it is kept at full precision rather than on tensor cores.
#include "attention.cuh"
#include "host_util.h"

namespace cv2 {

static constexpr int QT = 32, KT = 64, LD = 68;  // LD: 16B-aligned rows, conflict-free LDS.128
static constexpr int PR = QT + KT - 1;           // 95 relative positions per (q tile, k tile)
static constexpr int kRelSmemFloats = 2 * QT * LD + 2 * KT * LD + 96 * LD + QT * LD;
static constexpr int kRelSmem = kRelSmemFloats * 4;

__global__ void __launch_bounds__(256) rel_attn_kernel(const RelAttnParams p) {
  const int i0 = blockIdx.x * QT;
  const int h = blockIdx.y;
  const int s = blockIdx.z;
  const int len = p.lens ? p.lens[s] : p.len_all;
  if (i0 >= len) {
    // padded query rows: emit zeros so downstream A operands stay clean
    if (i0 < p.T_alloc) {
      for (int e = threadIdx.x; e < QT * 64; e += 256) {
        const int r = e >> 6, d = e & 63;
        p.out[((long long)s * p.T_alloc + i0 + r) * 512 + h * 64 + d] = __float2half_rn(0.f);
      }
    }
    return;
  }
  extern __shared__ float sm[];
  float* qu = sm;
  float* qv = qu + QT * LD;
  float* kt = qv + QT * LD;
  float* vt = kt + KT * LD;
  float* pt = vt + KT * LD;
  float* sc = pt + 96 * LD;

  const int tid = threadIdx.x;
  const int ti = tid >> 3, tj = tid & 7;
  const float* base = p.qkv + (long long)s * p.T_alloc * 1536;

  for (int e = tid; e < QT * 64; e += 256) {
    const int r = e >> 6, d = e & 63;
    const float qq = base[(long long)(i0 + r) * 1536 + h * 64 + d];
    qu[r * LD + d] = qq + p.bias_u[h * 64 + d];
    qv[r * LD + d] = qq + p.bias_v[h * 64 + d];
  }
  int kv_end = len;
  if (p.chunk > 0) kv_end = min(len, ((i0 + QT - 1) / p.chunk + 1) * p.chunk);
  const int nkt = (kv_end + KT - 1) / KT;
  const int i = i0 + ti;
  int kv_lim = len;
  if (p.chunk > 0) kv_lim = min(len, (i / p.chunk + 1) * p.chunk);

  float m = -INFINITY, l = 0.f;
  float o[8];
#pragma unroll
  for (int d = 0; d < 8; d++) o[d] = 0.f;

  for (int jt = 0; jt < nkt; jt++) {
    const int j0 = jt * KT;
    __syncthreads();
    for (int e = tid; e < KT * 64; e += 256) {
      const int r = e >> 6, d = e & 63;
      const float* row = base + (long long)(j0 + r) * 1536 + h * 64 + d;
      kt[r * LD + d] = row[512];
      vt[r * LD + d] = row[1024];
    }
    // relative positions rel = (i0 - j0 - (KT-1)) + lr, lr in [0, PR)
    for (int e = tid; e < PR * 64; e += 256) {
      const int lr = e >> 6, d = e & 63;
      const int prow = (i0 - j0 - (KT - 1) + lr) + p.Tmax - 1;
      pt[lr * LD + d] = (prow >= 0 && prow < 2 * p.Tmax - 1) ? p.pos[(long long)prow * 512 + h * 64 + d] : 0.f;
    }
    __syncthreads();
    float sv[8];
#pragma unroll
    for (int jj = 0; jj < 8; jj++) sv[jj] = 0.f;
    for (int d = 0; d < 64; d += 4) {
      const float4 a = *reinterpret_cast<const float4*>(qu + ti * LD + d);
      const float4 b = *reinterpret_cast<const float4*>(qv + ti * LD + d);
#pragma unroll
      for (int jj = 0; jj < 8; jj++) {
        const int jl = tj + 8 * jj;
        const float4 kk = *reinterpret_cast<const float4*>(kt + jl * LD + d);
        const float4 pp = *reinterpret_cast<const float4*>(pt + (ti - jl + KT - 1) * LD + d);
        sv[jj] += a.x * kk.x + a.y * kk.y + a.z * kk.z + a.w * kk.w + b.x * pp.x + b.y * pp.y + b.z * pp.z + b.w * pp.w;
      }
    }
    float mx = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < 8; jj++) {
      const int j = j0 + tj + 8 * jj;
      sv[jj] = (j < kv_lim) ? sv[jj] * 0.125f : -INFINITY;
      mx = fmaxf(mx, sv[jj]);
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
    const float m_new = fmaxf(m, mx);
    const float alpha = (m == -INFINITY) ? 0.f : __expf(m - m_new);
    float ls = 0.f;
#pragma unroll
    for (int jj = 0; jj < 8; jj++) {
      const float e = (m_new == -INFINITY || sv[jj] == -INFINITY) ? 0.f : __expf(sv[jj] - m_new);
      sc[ti * LD + tj + 8 * jj] = e;
      ls += e;
    }
    ls += __shfl_xor_sync(0xffffffffu, ls, 1);
    ls += __shfl_xor_sync(0xffffffffu, ls, 2);
    ls += __shfl_xor_sync(0xffffffffu, ls, 4);
    l = l * alpha + ls;
    m = m_new;
#pragma unroll
    for (int d = 0; d < 8; d++) o[d] *= alpha;
    __syncwarp();  // the 8 threads of a row live in one warp
    for (int j = 0; j < KT; j++) {
      const float pj = sc[ti * LD + j];
      const float4 v0 = *reinterpret_cast<const float4*>(vt + j * LD + tj * 8);
      const float4 v1 = *reinterpret_cast<const float4*>(vt + j * LD + tj * 8 + 4);
      o[0] += pj * v0.x; o[1] += pj * v0.y; o[2] += pj * v0.z; o[3] += pj * v0.w;
      o[4] += pj * v1.x; o[5] += pj * v1.y; o[6] += pj * v1.z; o[7] += pj * v1.w;
    }
  }
  const bool valid = i < len;
  const float inv = (l > 0.f) ? 1.f / l : 0.f;
  __half* dst = p.out + ((long long)s * p.T_alloc + i) * 512 + h * 64 + tj * 8;
#pragma unroll
  for (int d = 0; d < 8; d++) dst[d] = __float2half_rn(valid ? o[d] * inv : 0.f);
}

void launch_rel_attn(const RelAttnParams& p, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    CV2_CUDA(cudaFuncSetAttribute(rel_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRelSmem));
    configured = true;
  }
  CV2_CHECK(p.T_alloc % 64 == 0, "rel_attn: T_alloc %d not a multiple of 64", p.T_alloc);
  dim3 grid(p.T_alloc / QT, 8, p.S);
  rel_attn_kernel<<<grid, 256, kRelSmem, stream>>>(p);
  CV2_LAUNCH_CHECK();
}

}  // namespace cv2
