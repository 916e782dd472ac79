// Non-GEMM encoder kernels (norms, modality fusion, attention, output head). Templated on the
// activation storage type: float (reference-precision mode) or __half (tensor-core mode); all
// arithmetic is fp32. One warp owns one 256-wide row (8 channels per lane, 16/32-byte accesses).
//
// Replaces (reference model.py): GroupNorm(1,256) :32/:40; F.layer_norm no-affine :175; MinimalPerFrameFusion
// :79-98 (kv-LN, constant-query logits, temperature/bias, softmax over modalities, weighted sum);
// cls/positional embedding :187-188; nn.TransformerEncoderLayer self-attention + LayerNorms :145-146;
// output normalisation :190-192 and eval.py:218-224 per-window temporal coherence.
#include "common.cuh"
#include "kernels.h"
#include <math_constants.h>

namespace {

constexpr int kD = TAG_D_MODEL;
constexpr float kLnEps = 1e-5f;

// ---------------------------------------------------------------- GroupNorm(1 group) over (T x 256) per window
template <typename TA>
__global__ void __launch_bounds__(256) k_groupnorm(const TA* __restrict__ z, const float* __restrict__ gamma,
                                                   const float* __restrict__ beta, TA* __restrict__ out, int T) {
  __shared__ float red[33];
  const int64_t w = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const TA* zw = z + w * (int64_t)T * kD;
  TA* ow = out + w * (int64_t)T * kD;
  float s = 0.f;
  for (int t = warp; t < T; t += 8) {
    float v[8];
    Row8<TA>::load(zw + (int64_t)t * kD + lane * 8, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) s += v[k];
  }
  const float n = (float)T * (float)kD;
  const float mean = block_sum(s, red) / n;
  float q = 0.f;
  for (int t = warp; t < T; t += 8) {
    float v[8];
    Row8<TA>::load(zw + (int64_t)t * kD + lane * 8, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) { const float d = v[k] - mean; q += d * d; }
  }
  const float var = block_sum(q, red) / n;          // biased, as torch group_norm
  const float rstd = 1.0f / sqrtf(var + kLnEps);
  float g[8], b[8];
  Row8<float>::load(gamma + lane * 8, g);
  Row8<float>::load(beta + lane * 8, b);
  for (int t = warp; t < T; t += 8) {
    float v[8];
    Row8<TA>::load(zw + (int64_t)t * kD + lane * 8, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = (v[k] - mean) * rstd * g[k] + b[k];
    Row8<TA>::store(ow + (int64_t)t * kD + lane * 8, v);
  }
}

// row LayerNorm statistics across a warp (two-pass)
__device__ __forceinline__ void row_norm(float (&v)[8]) {
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += v[k];
  const float mean = warp_sum(s) * (1.0f / kD);
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) { v[k] -= mean; q += v[k] * v[k]; }
  const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / kD) + kLnEps);
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] *= rstd;
}

// ---------------------------------------------------------------- A11 + A12 (front half)
// s = state + motion -> F.layer_norm (no affine, model.py:175) -> fusion.kv_ln (affine, model.py:81) -> logit ->
// softmax over modalities -> sum_m A_m kv_m. The second LayerNorm needs no reductions of its own: its input
// u = (s - mean) * rstd1 has mean 0 and variance v / (v + eps) (v = var(s)), so
// kv = (s - mean) * rstd1 * rstd2 * gamma + beta with rstd2 = 1 / sqrt(v / (v + eps) + eps).
// (The reference evaluates mean(u) numerically; it is 0 up to fp32 round-off, ~1e-8.)
// kv_m is held as packed fp16 pairs between the logit pass and the mixing pass when activations are fp16 (the mix is
// stored as fp16 anyway), which keeps the kernel at 3+ CTAs per SM.
template <typename TA> struct KvHold;
template <> struct KvHold<float> {
  float v[8];
  __device__ __forceinline__ void put(const float (&x)[8]) {
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = x[k];
  }
  __device__ __forceinline__ void get(float (&x)[8]) const {
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = v[k];
  }
};
template <> struct KvHold<__half> {
  __half2 v[4];
  __device__ __forceinline__ void put(const float (&x)[8]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = __floats2half2_rn(x[2 * k], x[2 * k + 1]);
  }
  __device__ __forceinline__ void get(float (&x)[8]) const {
#pragma unroll
    for (int k = 0; k < 4; ++k) { const float2 f = __half22float2(v[k]); x[2 * k] = f.x; x[2 * k + 1] = f.y; }
  }
};

template <typename TA>
__global__ void __launch_bounds__(256) k_merge_fusion(const MergeParams p) {
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= p.R) return;
  float g[8], b[8], qk[8];
  Row8<float>::load(p.kv_gamma + lane * 8, g);
  Row8<float>::load(p.kv_beta + lane * 8, b);
  Row8<float>::load(p.qk + lane * 8, qk);
  KvHold<TA> kv[TAG_MAX_MODALITIES];
  float logit[TAG_MAX_MODALITIES];
  float mx = -CUDART_INF_F;
#pragma unroll
  for (int m = 0; m < TAG_MAX_MODALITIES; ++m) {
    if (m < p.M) {
      float v[8];
      Row8<TA>::load(reinterpret_cast<const TA*>(p.ps[m]) + r * kD + lane * 8, v);
      if (p.pm[m] != nullptr) {
        float u[8];
        Row8<TA>::load(reinterpret_cast<const TA*>(p.pm[m]) + r * kD + lane * 8, u);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] += u[k];             // s = state + motion  (model.py:174)
      }
      float sm = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) sm += v[k];
      const float mean = warp_sum(sm) * (1.0f / kD);
      float sq = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) { v[k] -= mean; sq = fmaf(v[k], v[k], sq); }
      const float var = warp_sum(sq) * (1.0f / kD);
      const float rstd1 = 1.0f / sqrtf(var + kLnEps);
      const float var_u = var * rstd1 * rstd1;                // variance of the first LayerNorm's output
      const float sc = rstd1 / sqrtf(var_u + kLnEps);
      float d = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) { v[k] = fmaf(v[k] * sc, g[k], b[k]); d = fmaf(v[k], qk[k], d); }
      kv[m].put(v);
      d = warp_sum(d);                                        // Q . (Wk kv) / sqrt(D)
      logit[m] = d * p.inv_tau[m] + p.lbias[m];               // model.py:89-91
      mx = fmaxf(mx, logit[m]);
    }
  }
  float den = 0.f;
#pragma unroll
  for (int m = 0; m < TAG_MAX_MODALITIES; ++m)
    if (m < p.M) { logit[m] = expf(logit[m] - mx); den += logit[m]; }
  const float inv_den = 1.0f / den;
  float mix[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) mix[k] = 0.f;
#pragma unroll
  for (int m = 0; m < TAG_MAX_MODALITIES; ++m)
    if (m < p.M) {
      const float a = logit[m] * inv_den;
      float x[8];
      kv[m].get(x);
#pragma unroll
      for (int k = 0; k < 8; ++k) mix[k] = fmaf(a, x[k], mix[k]);
      if (p.attn != nullptr && lane == 0) p.attn[r * p.M + m] = a;
    }
  Row8<TA>::store(reinterpret_cast<TA*>(p.mix) + r * kD + lane * 8, mix);
}

// fp16-activation form of the same step, built for memory-level parallelism: the row's 2*M 16-byte loads per lane are
// all issued before any arithmetic (the generic kernel above chained load -> three dependent shuffle reductions per
// modality and sat at 31 % of HBM bandwidth), the shuffle reductions of all modalities run interleaved (two sequences
// in total instead of three per modality), and kv is never materialised:
//   logit_m = sc_m * sum_k d_m[k] g[k] qk[k] + sum_k b[k] qk[k],   mix[k] = g[k] * sum_m (a_m sc_m) d_m[k] + b[k]
// with d_m = s_m - mean_m, sc_m = rstd1 * rstd2 (see above), a = softmax(logit) (sum_m a_m = 1).
// HM = false: no separate motion tensors (the encoder already produced s = state + motion in one projection GEMM)
template <int MM, bool HM>
__global__ void __launch_bounds__(256) k_merge_fusion_h(const MergeParams p) {
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= p.R) return;
  uint4 us[MM], um[MM];
#pragma unroll
  for (int m = 0; m < MM; ++m) {
    us[m] = make_uint4(0u, 0u, 0u, 0u); um[m] = us[m];
    if (m < p.M) {
      us[m] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(p.ps[m]) + r * kD + lane * 8));
      if (HM && p.pm[m] != nullptr) um[m] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(p.pm[m]) + r * kD + lane * 8));
    }
  }
  float g[8], b[8], gq[8];
  Row8<float>::load(p.kv_gamma + lane * 8, g);
  Row8<float>::load(p.kv_beta + lane * 8, b);
  Row8<float>::load(p.qk + lane * 8, gq);
  float bq = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) { bq = fmaf(b[k], gq[k], bq); gq[k] *= g[k]; }
  float d[MM][8], s1[MM];
#pragma unroll
  for (int m = 0; m < MM; ++m) {
    const __half2* hs = reinterpret_cast<const __half2*>(&us[m]);
    const __half2* hm = reinterpret_cast<const __half2*>(&um[m]);
    s1[m] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 a = __half22float2(hs[i]);
      const float2 c = HM ? __half22float2(hm[i]) : make_float2(0.f, 0.f);
      d[m][2 * i] = a.x + c.x; d[m][2 * i + 1] = a.y + c.y;            // s = state + motion  (model.py:174)
      s1[m] += d[m][2 * i] + d[m][2 * i + 1];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int m = 0; m < MM; ++m) s1[m] += __shfl_xor_sync(FULL_MASK, s1[m], o);
  }
  bq = warp_sum(bq);
  float s2[MM], sq[MM];
#pragma unroll
  for (int m = 0; m < MM; ++m) {
    const float mean = s1[m] * (1.0f / kD);
    s2[m] = 0.f; sq[m] = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { d[m][k] -= mean; s2[m] = fmaf(d[m][k], d[m][k], s2[m]); sq[m] = fmaf(d[m][k], gq[k], sq[m]); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int m = 0; m < MM; ++m) { s2[m] += __shfl_xor_sync(FULL_MASK, s2[m], o); sq[m] += __shfl_xor_sync(FULL_MASK, sq[m], o); }
  }
  float logit[MM], sc[MM];
  float mx = -CUDART_INF_F;
#pragma unroll
  for (int m = 0; m < MM; ++m) {
    const float var = s2[m] * (1.0f / kD);
    const float rstd1 = rsqrtf(var + kLnEps);                 // MUFU.RSQ (2^-22 relative): far below the fp16 rounding of the result
    const float var_u = var * rstd1 * rstd1;                  // variance of the first LayerNorm's output
    sc[m] = rstd1 * rsqrtf(var_u + kLnEps);
    logit[m] = -CUDART_INF_F;
    if (m < p.M) {
      logit[m] = fmaf(sc[m], sq[m], bq) * p.inv_tau[m] + p.lbias[m];    // Q . (Wk kv) / sqrt(D), model.py:89-91
      mx = fmaxf(mx, logit[m]);
    }
  }
  float den = 0.f;
#pragma unroll
  for (int m = 0; m < MM; ++m) { logit[m] = (m < p.M) ? __expf(logit[m] - mx) : 0.f; den += logit[m]; }
  const float inv_den = __fdividef(1.0f, den);
  float mix[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) mix[k] = 0.f;
#pragma unroll
  for (int m = 0; m < MM; ++m) {
    const float a = logit[m] * inv_den;
    const float wgt = a * sc[m];
#pragma unroll
    for (int k = 0; k < 8; ++k) mix[k] = fmaf(wgt, d[m][k], mix[k]);
    if (p.attn != nullptr && lane == 0 && m < p.M) p.attn[r * p.M + m] = a;
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) mix[k] = fmaf(mix[k], g[k], b[k]);
  Row8<__half>::store(reinterpret_cast<__half*>(p.mix) + r * kD + lane * 8, mix);
}

// Two rows per warp (round 2): the kernel above is instruction-issue bound (ncu: 737 warp instructions per row, 61 % of the
// issue slots, 0.38 of the HBM peak) and a third of those are warp-wide shuffle reductions and per-row scalar work that a
// 32-lane row pays once per ROW. Here a row is 16 lanes x 16 columns (two 16-byte loads per lane and modality: columns
// [8s, 8s+8) and [128+8s, 128+8s+8) of half-warp lane s, so a half-warp still reads two contiguous 256-byte runs), the
// reductions are 4 xor-steps that never leave the half-warp, and every shuffle / softmax instruction serves two rows.
// gamma / beta / gamma*qk come from shared memory (loaded once per CTA). No separate motion tensors (the tensor-core path
// always feeds s = state + motion from the fused projection GEMM).
template <int MM>
__global__ void __launch_bounds__(256) k_merge_fusion_h2(const MergeParams p) {
  __shared__ __align__(16) float s_g[kD], s_b[kD], s_gq[kD];
  __shared__ float s_bq;
  for (int i = threadIdx.x; i < kD; i += blockDim.x) {
    const float g = __ldg(p.kv_gamma + i), b = __ldg(p.kv_beta + i), q = __ldg(p.qk + i);
    s_g[i] = g; s_b[i] = b; s_gq[i] = g * q;
  }
  if (threadIdx.x < 32) {
    float bq = 0.f;
    for (int i = threadIdx.x; i < kD; i += 32) bq = fmaf(__ldg(p.kv_beta + i), __ldg(p.qk + i), bq);
    bq = warp_sum(bq);
    if (threadIdx.x == 0) s_bq = bq;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, sub = lane & 15;
  const int64_t r = (int64_t)blockIdx.x * 16 + (threadIdx.x >> 5) * 2 + (lane >> 4);
  const bool live = r < p.R;
  const int c0 = sub * 8, c1 = 128 + sub * 8;
  // all per-element arithmetic on packed fp32 pairs (FADD2 / FFMA2): the kernel is bound by instruction issue, and 56 % of its
  // instructions were scalar FFMA / FADD / FMUL (ncu, profiles/r2_ncu_pass_summary.md)
  f32x2 d[MM][8];
  float s1[MM];
#pragma unroll
  for (int m = 0; m < MM; ++m) {
    uint4 ua = make_uint4(0u, 0u, 0u, 0u), ub = ua;
    if (live && m < p.M) {
      const __half* src = reinterpret_cast<const __half*>(p.ps[m]) + r * kD;
      ua = __ldg(reinterpret_cast<const uint4*>(src + c0));
      ub = __ldg(reinterpret_cast<const uint4*>(src + c1));
    }
    const __half2* ha = reinterpret_cast<const __half2*>(&ua);
    const __half2* hb = reinterpret_cast<const __half2*>(&ub);
    f32x2 acc = pk2(0.f);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 a = __half22float2(ha[i]), b = __half22float2(hb[i]);
      d[m][i] = pk2(a.x, a.y); d[m][4 + i] = pk2(b.x, b.y);
      acc = add2(acc, add2(d[m][i], d[m][4 + i]));
    }
    float lo, hi;
    upk2(acc, lo, hi);
    s1[m] = lo + hi;
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) {
#pragma unroll
    for (int m = 0; m < MM; ++m) s1[m] += __shfl_xor_sync(FULL_MASK, s1[m], o);
  }
  f32x2 s2p[MM], sqp[MM];
#pragma unroll
  for (int m = 0; m < MM; ++m) { s2p[m] = pk2(0.f); sqp[m] = pk2(0.f); }
#pragma unroll
  for (int hblk = 0; hblk < 2; ++hblk) {
    const float4 q0 = *reinterpret_cast<const float4*>(s_gq + (hblk ? c1 : c0));
    const float4 q1 = *reinterpret_cast<const float4*>(s_gq + (hblk ? c1 : c0) + 4);
    const f32x2 gq[4] = {pk2(q0.x, q0.y), pk2(q0.z, q0.w), pk2(q1.x, q1.y), pk2(q1.z, q1.w)};
#pragma unroll
    for (int m = 0; m < MM; ++m) {
      const f32x2 nmean = pk2(s1[m] * (-1.0f / kD));
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const f32x2 x = add2(d[m][hblk * 4 + k], nmean);
        d[m][hblk * 4 + k] = x;
        s2p[m] = fma2(x, x, s2p[m]);
        sqp[m] = fma2(x, gq[k], sqp[m]);
      }
    }
  }
  float s2[MM], sq[MM];
#pragma unroll
  for (int m = 0; m < MM; ++m) {
    float lo, hi;
    upk2(s2p[m], lo, hi); s2[m] = lo + hi;
    upk2(sqp[m], lo, hi); sq[m] = lo + hi;
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) {
#pragma unroll
    for (int m = 0; m < MM; ++m) { s2[m] += __shfl_xor_sync(FULL_MASK, s2[m], o); sq[m] += __shfl_xor_sync(FULL_MASK, sq[m], o); }
  }
  const float bq = s_bq;
  float logit[MM], sc[MM];
  float mx = -CUDART_INF_F;
#pragma unroll
  for (int m = 0; m < MM; ++m) {
    const float var = s2[m] * (1.0f / kD);
    const float rstd1 = rsqrtf(var + kLnEps);
    const float var_u = var * rstd1 * rstd1;
    sc[m] = rstd1 * rsqrtf(var_u + kLnEps);
    logit[m] = -CUDART_INF_F;
    if (m < p.M) {
      logit[m] = fmaf(sc[m], sq[m], bq) * p.inv_tau[m] + p.lbias[m];
      mx = fmaxf(mx, logit[m]);
    }
  }
  float den = 0.f;
#pragma unroll
  for (int m = 0; m < MM; ++m) { logit[m] = (m < p.M) ? __expf(logit[m] - mx) : 0.f; den += logit[m]; }
  const float inv_den = __fdividef(1.0f, den);
  f32x2 wgt[MM];
#pragma unroll
  for (int m = 0; m < MM; ++m) {
    const float a = logit[m] * inv_den;
    wgt[m] = pk2(a * sc[m]);
    if (p.attn != nullptr && sub == 0 && live && m < p.M) p.attn[r * p.M + m] = a;
  }
#pragma unroll
  for (int hblk = 0; hblk < 2; ++hblk) {
    const int c = hblk ? c1 : c0;
    const float4 g0 = *reinterpret_cast<const float4*>(s_g + c), g1 = *reinterpret_cast<const float4*>(s_g + c + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(s_b + c), b1 = *reinterpret_cast<const float4*>(s_b + c + 4);
    const f32x2 g2[4] = {pk2(g0.x, g0.y), pk2(g0.z, g0.w), pk2(g1.x, g1.y), pk2(g1.z, g1.w)};
    const f32x2 b2[4] = {pk2(b0.x, b0.y), pk2(b0.z, b0.w), pk2(b1.x, b1.y), pk2(b1.z, b1.w)};
    float mix[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      f32x2 a = mul2(wgt[0], d[0][hblk * 4 + k]);
#pragma unroll
      for (int m = 1; m < MM; ++m) a = fma2(wgt[m], d[m][hblk * 4 + k], a);
      upk2(fma2(a, g2[k], b2[k]), mix[2 * k], mix[2 * k + 1]);
    }
    if (live) Row8<__half>::store(reinterpret_cast<__half*>(p.mix) + r * kD + c, mix);
  }
}

// ---------------------------------------------------------------- tokens = [cls ; frames] + PE
template <typename TA>
__global__ void __launch_bounds__(256) k_build_tokens(const TA* __restrict__ fused, const float* __restrict__ cls,
                                                      const float* __restrict__ pe, float* __restrict__ x32,
                                                      TA* __restrict__ x16, int64_t rows, int T) {
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const int S = T + 1;
  const int64_t n = r / S;
  const int s = (int)(r - n * S);
  float v[8], e[8];
  if (s == 0) Row8<float>::load(cls + lane * 8, v);
  else Row8<TA>::load(fused + (n * T + (s - 1)) * kD + lane * 8, v);
  Row8<float>::load(pe + (int64_t)s * kD + lane * 8, e);
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] += e[k];
  Row8<float>::store(x32 + r * kD + lane * 8, v);
  if (x16 != nullptr) Row8<TA>::store(x16 + r * kD + lane * 8, v);
}

// ---------------------------------------------------------------- self-attention, head_dim 32
// One CTA per (window, group of HPC heads); one thread per (head, query). K/V of the group are staged in shared
// memory as fp32 [S][HPC][36] (the 4-float pad keeps two heads read by one warp on different banks); every thread
// walks the keys with 128-bit shared loads (a warp reads one address -> broadcast) and an online softmax in the
// exp2 domain. For S = 33 and 8 heads, 264 of 288 threads are busy (one query per lane would leave 31/64 idle).
constexpr int kHeadPad = 36;

template <typename TA> struct Vec8;
template <> struct Vec8<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[8]) { Row8<float>::load(p, v); }
};
template <> struct Vec8<__half> {
  static __device__ __forceinline__ void load(const __half* p, float (&v)[8]) { Row8<__half>::load(p, v); }
};

template <typename TA, int HPC>
__global__ void __launch_bounds__(640) k_attention(const TA* __restrict__ qkv, TA* __restrict__ out, int S, int n_heads) {
  extern __shared__ __align__(16) float sm[];
  const int rowf = HPC * kHeadPad;
  float* sK = sm;                               // [S][HPC][36]
  float* sV = sm + (size_t)S * rowf;
  const int groups = n_heads / HPC;
  const int64_t n = blockIdx.x / groups;
  const int h0 = (blockIdx.x % groups) * HPC;
  const TA* base = qkv + n * (int64_t)S * (3 * kD);
  // stage K and V: units of 8 elements
  for (int u = threadIdx.x; u < S * HPC * 4; u += blockDim.x) {
    const int j = u / (HPC * 4), c = u - j * (HPC * 4);
    const int h = c >> 2, part = c & 3;
    float k8[8], v8[8];
    Vec8<TA>::load(base + (int64_t)j * (3 * kD) + kD + (h0 + h) * 32 + part * 8, k8);
    Vec8<TA>::load(base + (int64_t)j * (3 * kD) + 2 * kD + (h0 + h) * 32 + part * 8, v8);
    float* kd = sK + j * rowf + h * kHeadPad + part * 8;
    float* vd = sV + j * rowf + h * kHeadPad + part * 8;
    *reinterpret_cast<float4*>(kd) = make_float4(k8[0], k8[1], k8[2], k8[3]);
    *reinterpret_cast<float4*>(kd + 4) = make_float4(k8[4], k8[5], k8[6], k8[7]);
    *reinterpret_cast<float4*>(vd) = make_float4(v8[0], v8[1], v8[2], v8[3]);
    *reinterpret_cast<float4*>(vd + 4) = make_float4(v8[4], v8[5], v8[6], v8[7]);
  }
  __syncthreads();
  const float scale = 0.17677669529663688110f * 1.4426950408889634f;      // log2(e) / sqrt(32)
  for (int idx = threadIdx.x; idx < HPC * S; idx += blockDim.x) {
    const int h = idx / S, i = idx - h * S;
    float q[32], acc[32];
    const TA* qp = base + (int64_t)i * (3 * kD) + (h0 + h) * 32;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float t8[8];
      Vec8<TA>::load(qp + c * 8, t8);
#pragma unroll
      for (int d = 0; d < 8; ++d) { q[c * 8 + d] = t8[d] * scale; acc[c * 8 + d] = 0.f; }
    }
    float m = -CUDART_INF_F, l = 0.f;
    const float* kp = sK + h * kHeadPad;
    const float* vp = sV + h * kHeadPad;
    for (int j = 0; j < S; ++j, kp += rowf, vp += rowf) {
      float sc = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 k4 = *reinterpret_cast<const float4*>(kp + c * 4);
        sc = fmaf(q[c * 4], k4.x, sc); sc = fmaf(q[c * 4 + 1], k4.y, sc);
        sc = fmaf(q[c * 4 + 2], k4.z, sc); sc = fmaf(q[c * 4 + 3], k4.w, sc);
      }
      const float mn = fmaxf(m, sc);
      const float corr = exp2f(m - mn);
      const float pj = exp2f(sc - mn);
      l = fmaf(l, corr, pj);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 v4 = *reinterpret_cast<const float4*>(vp + c * 4);
        acc[c * 4] = fmaf(pj, v4.x, acc[c * 4] * corr); acc[c * 4 + 1] = fmaf(pj, v4.y, acc[c * 4 + 1] * corr);
        acc[c * 4 + 2] = fmaf(pj, v4.z, acc[c * 4 + 2] * corr); acc[c * 4 + 3] = fmaf(pj, v4.w, acc[c * 4 + 3] * corr);
      }
      m = mn;
    }
    const float inv = 1.0f / l;
    TA* o = out + (n * S + i) * kD + (h0 + h) * 32;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float t8[8];
#pragma unroll
      for (int d = 0; d < 8; ++d) t8[d] = acc[c * 8 + d] * inv;
      Row8<TA>::store(o + c * 8, t8);
    }
  }
}

// ---------------------------------------------------------------- self-attention on mma.sync fragments (S <= 48)
// One warp per (window, head): Q, K, V (row-major, 48 rows x 40 halfs each) are fetched with 16-byte cp.async — all
// ~400 requests of a warp in flight at once, nothing passes through registers — S = Q K^T as 3 x 6 tiles of m16n8k16
// (fp16 in, fp32 accumulate), softmax on the accumulator fragments (row statistics via quad shuffles, exp2 domain),
// P re-used in registers as the A operand of P V (3 x 4 tiles) whose B fragments come from row-major V through
// ldmatrix.trans; the output tile is staged in shared memory and leaves as 16-byte stores (64 B per row).
// 72 tensor instructions replace ~4.7k scalar-FMA warp instructions per (window, head). Attention is 0.03 % of the
// encoder's FLOPs, so the legacy warp-level MMA is the right tool here; the GEMM-shaped 99.9 % runs on tcgen05.
constexpr int kAttS = 48, kQStride = 40;
constexpr int kAttWarpHalfs = 3 * kAttS * kQStride;                      // Q + K + V per warp

// 2^x on the MUFU (softmax arguments are <= 0 and scores far from the denormal range; -inf -> 0)
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}

__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__global__ void __launch_bounds__(128) k_attention_mma(const __half* __restrict__ qkv, __half* __restrict__ out,
                                                       int64_t n_pairs, int S, int n_heads) {
  extern __shared__ __align__(16) __half smh[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t pair = (int64_t)blockIdx.x * 4 + warp;          // (window, head)
  if (pair >= n_pairs) return;
  const int64_t n = pair / n_heads;
  const int h = (int)(pair - n * n_heads);
  __half* sQ = smh + (size_t)warp * kAttWarpHalfs;
  __half* sK = sQ + kAttS * kQStride;
  __half* sV = sK + kAttS * kQStride;
  const __half* base = qkv + n * (int64_t)S * (3 * kD) + h * 32;
  for (int u = lane; u < S * 4; u += 32) {
    const int row = u >> 2, part = u & 3;
    const __half* src = base + (int64_t)row * (3 * kD) + part * 8;
    cp_async16(sQ + row * kQStride + part * 8, src);
    cp_async16(sK + row * kQStride + part * 8, src + kD);
    cp_async16(sV + row * kQStride + part * 8, src + 2 * kD);
  }
  // value rows of the padding keys must be exact zeros (P is 0 there, but 0 x NaN would poison the sum); padding rows
  // of Q / K only produce scores that are masked or rows that are never stored
  for (int u = lane; u < (kAttS - S) * 4; u += 32)
    *reinterpret_cast<uint4*>(sV + (S + (u >> 2)) * kQStride + (u & 3) * 8) = make_uint4(0u, 0u, 0u, 0u);
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
  __syncwarp();

  const int g = lane >> 2, tig = lane & 3;
  // ---- S = Q K^T : acc[mt][nt][4]
  float sc[3][6][4];
#pragma unroll
  for (int mt = 0; mt < 3; ++mt)
#pragma unroll
    for (int nt = 0; nt < 6; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) sc[mt][nt][e] = 0.f;
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    uint32_t bk[6][2];
#pragma unroll
    for (int nt = 0; nt < 6; ++nt) {
      const __half* kp = sK + (nt * 8 + g) * kQStride + ks * 16 + 2 * tig;
      bk[nt][0] = *reinterpret_cast<const uint32_t*>(kp);
      bk[nt][1] = *reinterpret_cast<const uint32_t*>(kp + 8);
    }
#pragma unroll
    for (int mt = 0; mt < 3; ++mt) {
      uint32_t a[4];
      const __half* qp = sQ + (mt * 16 + g) * kQStride + ks * 16 + 2 * tig;
      a[0] = *reinterpret_cast<const uint32_t*>(qp);
      a[1] = *reinterpret_cast<const uint32_t*>(qp + 8 * kQStride);
      a[2] = *reinterpret_cast<const uint32_t*>(qp + 8);
      a[3] = *reinterpret_cast<const uint32_t*>(qp + 8 * kQStride + 8);
#pragma unroll
      for (int nt = 0; nt < 6; ++nt) mma_16816(sc[mt][nt], a, bk[nt]);
    }
  }
  // ---- softmax over keys (columns): row g uses elements [0],[1]; row g+8 uses [2],[3]
  const float scale = 0.17677669529663688110f * 1.4426950408889634f;      // log2(e) / sqrt(32)
  uint32_t pa[3][3][4];        // P as A fragments: [mt][kt][4]
  float inv_sum[3][2];
#pragma unroll
  for (int mt = 0; mt < 3; ++mt) {
    float mx0 = -CUDART_INF_F, mx1 = -CUDART_INF_F;
#pragma unroll
    for (int nt = 0; nt < 6; ++nt) {
      if (nt * 8 + 8 > S) {                              // only the tiles that contain padding keys need the mask
        const int c0 = nt * 8 + 2 * tig;
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (c0 + (e & 1) >= S) sc[mt][nt][e] = -CUDART_INF_F;
      }
      mx0 = fmaxf(mx0, fmaxf(sc[mt][nt][0], sc[mt][nt][1]));
      mx1 = fmaxf(mx1, fmaxf(sc[mt][nt][2], sc[mt][nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(FULL_MASK, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(FULL_MASK, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(FULL_MASK, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(FULL_MASK, mx1, 2));
    // softmax(q.k / sqrt(32)): p = 2^(s * scale - max * scale) — scale > 0, so the raw maximum is the scaled one's argmax
    const float m0s = -mx0 * scale, m1s = -mx1 * scale;
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 6; ++nt) {
      sc[mt][nt][0] = ex2_fast(fmaf(sc[mt][nt][0], scale, m0s)); sc[mt][nt][1] = ex2_fast(fmaf(sc[mt][nt][1], scale, m0s));
      sc[mt][nt][2] = ex2_fast(fmaf(sc[mt][nt][2], scale, m1s)); sc[mt][nt][3] = ex2_fast(fmaf(sc[mt][nt][3], scale, m1s));
      s0 += sc[mt][nt][0] + sc[mt][nt][1];
      s1 += sc[mt][nt][2] + sc[mt][nt][3];
    }
    s0 += __shfl_xor_sync(FULL_MASK, s0, 1); s0 += __shfl_xor_sync(FULL_MASK, s0, 2);
    s1 += __shfl_xor_sync(FULL_MASK, s1, 1); s1 += __shfl_xor_sync(FULL_MASK, s1, 2);
    inv_sum[mt][0] = 1.0f / s0; inv_sum[mt][1] = 1.0f / s1;
#pragma unroll
    for (int kt = 0; kt < 3; ++kt) {
      __half2 t;
      t = __floats2half2_rn(sc[mt][2 * kt][0], sc[mt][2 * kt][1]);         pa[mt][kt][0] = *reinterpret_cast<uint32_t*>(&t);
      t = __floats2half2_rn(sc[mt][2 * kt][2], sc[mt][2 * kt][3]);         pa[mt][kt][1] = *reinterpret_cast<uint32_t*>(&t);
      t = __floats2half2_rn(sc[mt][2 * kt + 1][0], sc[mt][2 * kt + 1][1]); pa[mt][kt][2] = *reinterpret_cast<uint32_t*>(&t);
      t = __floats2half2_rn(sc[mt][2 * kt + 1][2], sc[mt][2 * kt + 1][3]); pa[mt][kt][3] = *reinterpret_cast<uint32_t*>(&t);
    }
  }
  // ---- O = P V : B[k = key][n = d] = V[key][d] = Vt[d][key]
  float o[3][4][4];
#pragma unroll
  for (int mt = 0; mt < 3; ++mt)
#pragma unroll
    for (int nd = 0; nd < 4; ++nd)
#pragma unroll
      for (int e = 0; e < 4; ++e) o[mt][nd][e] = 0.f;
#pragma unroll
  for (int kt = 0; kt < 3; ++kt) {
#pragma unroll
    for (int np = 0; np < 2; ++np) {                 // two 8-wide d tiles per ldmatrix.x4
      // matrices: (keys kt*16 + 0..7, d tile 2np), (keys +8..15, d tile 2np), (keys 0..7, d tile 2np+1), (keys +8..15, 2np+1)
      const int mi = lane >> 3, rr = lane & 7;
      const __half* vp = sV + (kt * 16 + (mi & 1) * 8 + rr) * kQStride + (2 * np + (mi >> 1)) * 8;
      uint32_t b0, b1, b2, b3;
      asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                   : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3) : "r"((uint32_t)__cvta_generic_to_shared(vp)));
      const uint32_t bva[2] = {b0, b1}, bvb[2] = {b2, b3};
#pragma unroll
      for (int mt = 0; mt < 3; ++mt) { mma_16816(o[mt][2 * np], pa[mt][kt], bva); mma_16816(o[mt][2 * np + 1], pa[mt][kt], bvb); }
    }
  }
  // ---- output: stage the [48 x 32] tile over Q (dead after the score MMAs), then 16 bytes per lane, 64 B per row
  __syncwarp();
#pragma unroll
  for (int mt = 0; mt < 3; ++mt) {
    const int r0 = mt * 16 + g, r1 = r0 + 8;
#pragma unroll
    for (int nd = 0; nd < 4; ++nd) {
      const int col = nd * 8 + 2 * tig;
      *reinterpret_cast<__half2*>(sQ + r0 * kQStride + col) = __floats2half2_rn(o[mt][nd][0] * inv_sum[mt][0], o[mt][nd][1] * inv_sum[mt][0]);
      *reinterpret_cast<__half2*>(sQ + r1 * kQStride + col) = __floats2half2_rn(o[mt][nd][2] * inv_sum[mt][1], o[mt][nd][3] * inv_sum[mt][1]);
    }
  }
  __syncwarp();
  __half* ob = out + n * (int64_t)S * kD + h * 32;
  for (int u = lane; u < S * 4; u += 32) {
    const int row = u >> 2, part = u & 3;
    *reinterpret_cast<uint4*>(ob + (int64_t)row * kD + part * 8) = *reinterpret_cast<const uint4*>(sQ + row * kQStride + part * 8);
  }
}

// ---------------------------------------------------------------- self-attention, any S > 48 (long clips: S = 65 ... 257 ...)
// One CTA (4 warps) per (window, head): Q, K, V [S_pad x 40 halfs] of the head are fetched once with cp.async; each warp
// takes 16-query blocks round-robin and walks the keys in chunks of 32 with an online softmax (flash-attention recurrence in
// the exp2 domain): 8 score MMAs, rescale, 8 P.V MMAs per chunk, V fragments through ldmatrix.trans. The finished 16 x 32
// output block is staged over the warp's own (dead) Q rows and leaves as 16-byte stores.
__global__ void __launch_bounds__(128) k_attention_flash(const __half* __restrict__ qkv, __half* __restrict__ out, int S, int S_pad,
                                                         int n_heads) {
  extern __shared__ __align__(16) __half smh[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n = blockIdx.x / n_heads;
  const int h = (int)(blockIdx.x - n * n_heads);
  __half* sQ = smh;
  __half* sK = sQ + (size_t)S_pad * kQStride;
  __half* sV = sK + (size_t)S_pad * kQStride;
  const __half* base = qkv + n * (int64_t)S * (3 * kD) + h * 32;
  for (int u = threadIdx.x; u < S * 4; u += 128) {
    const int row = u >> 2, part = u & 3;
    const __half* src = base + (int64_t)row * (3 * kD) + part * 8;
    cp_async16(sQ + row * kQStride + part * 8, src);
    cp_async16(sK + row * kQStride + part * 8, src + kD);
    cp_async16(sV + row * kQStride + part * 8, src + 2 * kD);
  }
  // padding rows: V must be exact zeros (P = 0 there, but 0 x NaN poisons the sum); Q / K zeroed too so that masked scores
  // and never-stored rows stay finite
  for (int u = threadIdx.x; u < (S_pad - S) * 4; u += 128) {
    const int off = (S + (u >> 2)) * kQStride + (u & 3) * 8;
    *reinterpret_cast<uint4*>(sQ + off) = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(sK + off) = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(sV + off) = make_uint4(0u, 0u, 0u, 0u);
  }
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  const int g = lane >> 2, tig = lane & 3;
  const float scale = 0.17677669529663688110f * 1.4426950408889634f;      // log2(e) / sqrt(32)
  __half* ob = out + n * (int64_t)S * kD + h * 32;
  for (int qb = warp; qb * 16 < S; qb += 4) {
    uint32_t aq[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const __half* qp = sQ + (qb * 16 + g) * kQStride + ks * 16 + 2 * tig;
      aq[ks][0] = *reinterpret_cast<const uint32_t*>(qp);
      aq[ks][1] = *reinterpret_cast<const uint32_t*>(qp + 8 * kQStride);
      aq[ks][2] = *reinterpret_cast<const uint32_t*>(qp + 8);
      aq[ks][3] = *reinterpret_cast<const uint32_t*>(qp + 8 * kQStride + 8);
    }
    float m0 = -CUDART_INF_F, m1 = -CUDART_INF_F, l0 = 0.f, l1 = 0.f;
    float o[4][4];
#pragma unroll
    for (int nd = 0; nd < 4; ++nd)
#pragma unroll
      for (int e = 0; e < 4; ++e) o[nd][e] = 0.f;
    for (int kc = 0; kc < S_pad; kc += 32) {
      float sc[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) sc[nt][e] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          uint32_t bk[2];
          const __half* kp = sK + (kc + nt * 8 + g) * kQStride + ks * 16 + 2 * tig;
          bk[0] = *reinterpret_cast<const uint32_t*>(kp);
          bk[1] = *reinterpret_cast<const uint32_t*>(kp + 8);
          mma_16816(sc[nt], aq[ks], bk);
        }
      }
      float mx0 = m0, mx1 = m1;                           // running maxima of the RAW scores (scale > 0)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        if (kc + nt * 8 + 8 > S) {                         // only tiles with padding keys need the mask
          const int c0 = kc + nt * 8 + 2 * tig;
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (c0 + (e & 1) >= S) sc[nt][e] = -CUDART_INF_F;
        }
        mx0 = fmaxf(mx0, fmaxf(sc[nt][0], sc[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(sc[nt][2], sc[nt][3]));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(FULL_MASK, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(FULL_MASK, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(FULL_MASK, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(FULL_MASK, mx1, 2));
      // every chunk holds at least one real key (S_pad - S < 32), so mx is finite from the first chunk on
      const float a0 = ex2_fast((m0 - mx0) * scale), a1 = ex2_fast((m1 - mx1) * scale);
      m0 = mx0; m1 = mx1;
      const float m0s = -mx0 * scale, m1s = -mx1 * scale;
      l0 *= a0; l1 *= a1;
#pragma unroll
      for (int nd = 0; nd < 4; ++nd) { o[nd][0] *= a0; o[nd][1] *= a0; o[nd][2] *= a1; o[nd][3] *= a1; }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        sc[nt][0] = ex2_fast(fmaf(sc[nt][0], scale, m0s)); sc[nt][1] = ex2_fast(fmaf(sc[nt][1], scale, m0s));
        sc[nt][2] = ex2_fast(fmaf(sc[nt][2], scale, m1s)); sc[nt][3] = ex2_fast(fmaf(sc[nt][3], scale, m1s));
        l0 += sc[nt][0] + sc[nt][1];
        l1 += sc[nt][2] + sc[nt][3];
      }
#pragma unroll
      for (int kt = 0; kt < 2; ++kt) {
        uint32_t pa[4];
        __half2 t;
        t = __floats2half2_rn(sc[2 * kt][0], sc[2 * kt][1]);         pa[0] = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2half2_rn(sc[2 * kt][2], sc[2 * kt][3]);         pa[1] = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2half2_rn(sc[2 * kt + 1][0], sc[2 * kt + 1][1]); pa[2] = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2half2_rn(sc[2 * kt + 1][2], sc[2 * kt + 1][3]); pa[3] = *reinterpret_cast<uint32_t*>(&t);
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          const int mi = lane >> 3, rr = lane & 7;
          const __half* vp = sV + (kc + kt * 16 + (mi & 1) * 8 + rr) * kQStride + (2 * np + (mi >> 1)) * 8;
          uint32_t b0, b1, b2, b3;
          asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                       : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3) : "r"((uint32_t)__cvta_generic_to_shared(vp)));
          const uint32_t bva[2] = {b0, b1}, bvb[2] = {b2, b3};
          mma_16816(o[2 * np], pa, bva);
          mma_16816(o[2 * np + 1], pa, bvb);
        }
      }
    }
    l0 += __shfl_xor_sync(FULL_MASK, l0, 1); l0 += __shfl_xor_sync(FULL_MASK, l0, 2);
    l1 += __shfl_xor_sync(FULL_MASK, l1, 1); l1 += __shfl_xor_sync(FULL_MASK, l1, 2);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    // stage over this warp's own Q rows (their A fragments are in registers), then 16 bytes per lane
    __syncwarp();
#pragma unroll
    for (int nd = 0; nd < 4; ++nd) {
      const int col = nd * 8 + 2 * tig;
      *reinterpret_cast<__half2*>(sQ + (qb * 16 + g) * kQStride + col) = __floats2half2_rn(o[nd][0] * i0, o[nd][1] * i0);
      *reinterpret_cast<__half2*>(sQ + (qb * 16 + g + 8) * kQStride + col) = __floats2half2_rn(o[nd][2] * i1, o[nd][3] * i1);
    }
    __syncwarp();
#pragma unroll
    for (int u = lane; u < 64; u += 32) {
      const int row = qb * 16 + (u >> 2), part = u & 3;
      if (row < S) *reinterpret_cast<uint4*>(ob + (int64_t)row * kD + part * 8) = *reinterpret_cast<const uint4*>(sQ + row * kQStride + part * 8);
    }
  }
}

// ---------------------------------------------------------------- LayerNorm with affine
template <typename TA>
__global__ void __launch_bounds__(256) k_layernorm(const float* __restrict__ x, const float* __restrict__ gamma,
                                                   const float* __restrict__ beta, float* __restrict__ y32,
                                                   TA* __restrict__ y16, int64_t rows) {
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  float v[8], g[8], b[8];
  Row8<float>::load(x + r * kD + lane * 8, v);
  Row8<float>::load(gamma + lane * 8, g);
  Row8<float>::load(beta + lane * 8, b);
  row_norm(v);
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = v[k] * g[k] + b[k];
  Row8<float>::store(y32 + r * kD + lane * 8, v);
  if (y16 != nullptr) Row8<TA>::store(y16 + r * kD + lane * 8, v);
}

// ---------------------------------------------------------------- output head: one warp per window
// Rows are taken four at a time, and the NEXT four are loaded before these four are processed: 8 KB per warp in flight, the
// warp-shuffle chains of the four row norms (then of the four consecutive-frame distances) interleaved, so the kernel streams
// instead of waiting on one load -> shuffle -> sqrt chain per token (round 1: 3.0 TB/s; the token stream is read once, 1 KB
// per token). 128-thread CTAs (4 windows) keep the last wave short.
template <bool NORMALIZE>
__global__ void __launch_bounds__(128) k_finalize(const float* __restrict__ tokens, int64_t n_windows, int S,
                                                  float* __restrict__ seq, float* __restrict__ frame_embeds,
                                                  float* __restrict__ tokens_out, float* __restrict__ tc_window) {
  const int64_t n = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= n_windows) return;
  constexpr int U = 4;
  const float* base = tokens + n * S * kD + lane * 8;
  float prev[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) prev[k] = 0.f;
  float tsum = 0.f;
  float nxt[U][8];                                             // the NEXT four rows are already in flight while these four are processed
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (u < S) Row8<float>::load(base + (int64_t)u * kD, nxt[u]);
    else {
#pragma unroll
      for (int k = 0; k < 8; ++k) nxt[u][k] = 0.f;
    }
  }
  for (int s0 = 0; s0 < S; s0 += U) {
    float v[U][8];
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[u][k] = nxt[u][k];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (s0 + U + u < S) Row8<float>::load(base + (int64_t)(s0 + U + u) * kD, nxt[u]);
    }
    if (tokens_out != nullptr) {
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (s0 + u < S) Row8<float>::store(tokens_out + (n * S + s0 + u) * kD + lane * 8, v[u]);
    }
    if (NORMALIZE) {
      float ss[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        ss[u] = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) ss[u] = fmaf(v[u][k], v[u][k], ss[u]);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int u = 0; u < U; ++u) ss[u] += __shfl_xor_sync(FULL_MASK, ss[u], o);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        // F.normalize (model.py:191-192): x / max(|x|, 1e-12) as ONE correctly rounded reciprocal per row and eight multiplies (<= 1 ulp
        // from the eight IEEE divisions, which made this kernel instruction-issue bound: ncu 61 % of the issue slots)
        const float inv = 1.0f / fmaxf(sqrtf(ss[u]), 1e-12f);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[u][k] = v[u][k] * inv;
      }
    }
    float d2[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      d2[u] = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) { const float d = v[u][k] - (u == 0 ? prev[k] : v[u - 1][k]); d2[u] = fmaf(d, d, d2[u]); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int u = 0; u < U; ++u) d2[u] += __shfl_xor_sync(FULL_MASK, d2[u], o);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int s = s0 + u;
      if (s < S) {
        if (frame_embeds != nullptr) Row8<float>::store(frame_embeds + (n * S + s) * kD + lane * 8, v[u]);
        if (s == 0 && seq != nullptr) Row8<float>::store(seq + n * kD + lane * 8, v[u]);
        if (s >= 2) tsum += sqrtf(d2[u]);                         // eval.py:218-224: frames only (drop CLS)
      }
    }
    constexpr int L = U - 1;
#pragma unroll
    for (int k = 0; k < 8; ++k) prev[k] = v[L][k];
  }
  if (tc_window != nullptr && lane == 0) tc_window[n] = (S >= 3) ? tsum / (float)(S - 2) : CUDART_NAN_F;
}

}  // namespace

// ---------------------------------------------------------------- launchers
template <typename TA>
cudaError_t launch_groupnorm(const TA* z, const float* gamma, const float* beta, TA* out, int64_t n_windows, int T,
                             cudaStream_t s) {
  if (n_windows <= 0) return cudaSuccess;
  k_groupnorm<TA><<<(unsigned)n_windows, 256, 0, s>>>(z, gamma, beta, out, T);
  return cudaGetLastError();
}
template cudaError_t launch_groupnorm<float>(const float*, const float*, const float*, float*, int64_t, int, cudaStream_t);
template cudaError_t launch_groupnorm<__half>(const __half*, const float*, const float*, __half*, int64_t, int, cudaStream_t);

template <typename TA>
cudaError_t launch_merge_fusion(const MergeParams& p, cudaStream_t s) {
  if (p.R <= 0) return cudaSuccess;
  const unsigned grid = (unsigned)((p.R + 7) / 8);
  if constexpr (sizeof(TA) == 2) {
    bool hm = false;
    for (int m = 0; m < p.M; ++m) hm = hm || p.pm[m] != nullptr;
    if (p.M <= 5 && !hm) { k_merge_fusion_h2<5><<<(unsigned)((p.R + 15) / 16), 256, 0, s>>>(p); return cudaGetLastError(); }
    if (p.M <= 5) { if (hm) k_merge_fusion_h<5, true><<<grid, 256, 0, s>>>(p); else k_merge_fusion_h<5, false><<<grid, 256, 0, s>>>(p); }
    else { if (hm) k_merge_fusion_h<TAG_MAX_MODALITIES, true><<<grid, 256, 0, s>>>(p); else k_merge_fusion_h<TAG_MAX_MODALITIES, false><<<grid, 256, 0, s>>>(p); }
    return cudaGetLastError();
  }
  k_merge_fusion<TA><<<grid, 256, 0, s>>>(p);
  return cudaGetLastError();
}
template cudaError_t launch_merge_fusion<float>(const MergeParams&, cudaStream_t);
template cudaError_t launch_merge_fusion<__half>(const MergeParams&, cudaStream_t);

template <typename TA>
cudaError_t launch_build_tokens(const TA* fused, const float* cls, const float* pe, float* x32, TA* x16,
                                int64_t n_windows, int T, cudaStream_t s) {
  const int64_t rows = n_windows * (T + 1);
  if (rows <= 0) return cudaSuccess;
  k_build_tokens<TA><<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(fused, cls, pe, x32, x16, rows, T);
  return cudaGetLastError();
}
template cudaError_t launch_build_tokens<float>(const float*, const float*, const float*, float*, float*, int64_t, int, cudaStream_t);
template cudaError_t launch_build_tokens<__half>(const __half*, const float*, const float*, float*, __half*, int64_t, int, cudaStream_t);

template <typename TA, int HPC>
static cudaError_t launch_attention_hpc(const TA* qkv, TA* out, int64_t n_windows, int S, int n_heads, cudaStream_t s) {
  const size_t smem = (size_t)2 * S * HPC * kHeadPad * sizeof(float);
  int threads = ((HPC * S + 31) / 32) * 32;
  if (threads > 640) threads = 640;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_attention<TA, HPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  k_attention<TA, HPC><<<(unsigned)(n_windows * (n_heads / HPC)), threads, smem, s>>>(qkv, out, S, n_heads);
  return cudaGetLastError();
}

static cudaError_t launch_attention_mma(const __half* qkv, __half* out, int64_t n_windows, int S, int n_heads, cudaStream_t s) {
  const int64_t pairs = n_windows * n_heads;
  const size_t smem = (size_t)4 * kAttWarpHalfs * sizeof(__half);
  k_attention_mma<<<(unsigned)((pairs + 3) / 4), 128, smem, s>>>(qkv, out, pairs, S, n_heads);
  return cudaGetLastError();
}

template <typename TA>
cudaError_t launch_attention(const TA* qkv, TA* out, int64_t n_windows, int S, int n_heads, cudaStream_t s) {
  if (n_windows <= 0) return cudaSuccess;
  if constexpr (sizeof(TA) == 2) {
    if (S <= kAttS) return launch_attention_mma(qkv, out, n_windows, S, n_heads, s);
    const int S_pad = (S + 31) / 32 * 32;
    const size_t smem = (size_t)3 * S_pad * kQStride * sizeof(__half);
    if (smem <= 200 * 1024 && n_windows * n_heads < (1ll << 31)) {
      static size_t configured[64];                     // per device: function attributes are device state
      int dev = 0;
      cudaGetDevice(&dev);
      size_t& cfg = configured[dev & 63];
      if (smem > 48 * 1024 && smem > cfg) {
        cudaError_t e = cudaFuncSetAttribute(k_attention_flash, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        cfg = smem;
      }
      k_attention_flash<<<(unsigned)(n_windows * n_heads), 128, smem, s>>>(qkv, out, S, S_pad, n_heads);
      return cudaGetLastError();
    }
  }
  // largest head group whose K/V fit comfortably in shared memory
  auto fits = [&](int hpc) { return n_heads % hpc == 0 && (size_t)2 * S * hpc * kHeadPad * sizeof(float) <= 160 * 1024; };
  if (fits(8)) return launch_attention_hpc<TA, 8>(qkv, out, n_windows, S, n_heads, s);
  if (fits(4)) return launch_attention_hpc<TA, 4>(qkv, out, n_windows, S, n_heads, s);
  if (fits(2)) return launch_attention_hpc<TA, 2>(qkv, out, n_windows, S, n_heads, s);
  if (fits(1)) return launch_attention_hpc<TA, 1>(qkv, out, n_windows, S, n_heads, s);
  return cudaErrorInvalidValue;       // S > ~560 tokens per window: not built
}
template cudaError_t launch_attention<float>(const float*, float*, int64_t, int, int, cudaStream_t);
template cudaError_t launch_attention<__half>(const __half*, __half*, int64_t, int, int, cudaStream_t);

template <typename TA>
cudaError_t launch_layernorm(const float* x, const float* gamma, const float* beta, float* y32, TA* y16, int64_t rows,
                             cudaStream_t s) {
  if (rows <= 0) return cudaSuccess;
  k_layernorm<TA><<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(x, gamma, beta, y32, y16, rows);
  return cudaGetLastError();
}
template cudaError_t launch_layernorm<float>(const float*, const float*, const float*, float*, float*, int64_t, cudaStream_t);
template cudaError_t launch_layernorm<__half>(const float*, const float*, const float*, float*, __half*, int64_t, cudaStream_t);

cudaError_t launch_finalize(const float* tokens, int64_t n_windows, int S, float* seq, float* frame_embeds,
                            float* tokens_out, float* tc_window, cudaStream_t s) {
  if (n_windows <= 0) return cudaSuccess;
  k_finalize<true><<<(unsigned)((n_windows + 3) / 4), 128, 0, s>>>(tokens, n_windows, S, seq, frame_embeds, tokens_out, tc_window);
  return cudaGetLastError();
}

// eval.py:218-224 on already-normalised frame embeddings [N, S, 256] (S = T+1, row 0 = CLS)
cudaError_t launch_window_tc(const float* frame_embeds, int64_t n_windows, int S, float* tc_window, cudaStream_t s) {
  if (n_windows <= 0) return cudaSuccess;
  k_finalize<false><<<(unsigned)((n_windows + 3) / 4), 128, 0, s>>>(frame_embeds, n_windows, S, nullptr, nullptr, nullptr, tc_window);
  return cudaGetLastError();
}
