// Shared device helpers for the TAG scoring kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

#define TAG_WARP 32
#define FULL_MASK 0xffffffffu

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL_MASK, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
  return v;
}

// Sum over a whole CTA (blockDim.x multiple of 32, <= 1024). `red` = >= 33 floats of smem.
// Every thread gets the result. Contains two __syncthreads.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();               // protect `red` from a previous use
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : 0.f;
  t = warp_sum(t);
  return t;
}

// exact (erf) GELU, as nn.GELU() default (reference model.py:27, :31)
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

// Branch-free GELU for the tensor-core epilogues: erf by Abramowitz-Stegun 7.1.26 (|abs err| < 1.5e-7, below
// fp32 round-off of the surrounding math and far below the fp16 rounding of the stored activation) with
// MUFU rcp/ex2 — ~16 instructions per element instead of erff's ~30, which is what lets the epilogue warps keep
// pace with the MMA pipe (budget: 40 thread-instructions per output element at K = 1280).
__device__ __forceinline__ float gelu_fast(float x) {
  // GELU(x) = max(x, 0) - 0.5 |x| erfc(|x| / sqrt 2), erfc(u) = t (a1 + t (a2 + t (a3 + t (a4 + t a5)))) exp(-u^2),
  // t = 1 / (1 + 0.3275911 u). With z = |x| sqrt(log2(e) / 2): exp(-u^2) = 2^(-z^2), so the exponent needs no extra scaling.
  constexpr float kZ = 0.84932180028801904272f;              // sqrt(log2(e) / 2)
  constexpr float kT = 0.3275911f * 0.70710678118654752440f / kZ;   // 0.3275911 u = kT z
  constexpr float kH = 0.5f / kZ;                            // 0.5 |x| = kH z
  const float z = fabsf(x) * kZ;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(kT, z, 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * -z));
  float p = fmaf(t, 1.061405429f, -1.453152027f);
  p = fmaf(t, p, 1.421413741f);
  p = fmaf(t, p, -0.284496736f);
  p = fmaf(t, p, 0.254829592f);
  return fmaf(-kH * (p * t), z * e, fmaxf(x, 0.0f));
}

// ---- packed fp32 pairs (sm_100: FFMA2 / FMUL2 / FADD2 take one issue slot for two lanes of fp32 math). The tensor-core
// epilogues are bounded by the instruction issue of their 16 warps, so the per-element arithmetic runs on register pairs.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ f32x2 pk2(float a) { return pk2(a, a); }
__device__ __forceinline__ void upk2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// gelu_fast on a register pair: the polynomial and the products run packed (4.5 packed + 5 scalar instructions per element
// instead of 14 scalar); |x| only appears where a scalar instruction takes it as an operand modifier.
__device__ __forceinline__ f32x2 gelu_fast2(f32x2 x) {
  constexpr float kZ = 0.84932180028801904272f;              // sqrt(log2(e) / 2), as in gelu_fast
  constexpr float kT = 0.3275911f * 0.70710678118654752440f / kZ;
  constexpr float kH = 0.5f / kZ;
  const f32x2 s = mul2(x, pk2(kZ));                          // signed z
  const f32x2 q = mul2(s, s);
  float x0, x1, s0, s1, q0, q1, t0, t1, e0, e1;
  upk2(x, x0, x1); upk2(s, s0, s1); upk2(q, q0, q1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(fmaf(kT, fabsf(s0), 1.0f)));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(fmaf(kT, fabsf(s1), 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(-q0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(-q1));
  const f32x2 t = pk2(t0, t1);
  f32x2 p = fma2(t, pk2(1.061405429f), pk2(-1.453152027f));
  p = fma2(t, p, pk2(1.421413741f));
  p = fma2(t, p, pk2(-0.284496736f));
  p = fma2(t, p, pk2(0.254829592f));
  const f32x2 c = mul2(mul2(p, t), mul2(pk2(e0, e1), s));    // z t p(t) 2^(-z^2), carrying the sign of x
  float c0, c1;
  upk2(c, c0, c1);
  return pk2(fmaf(-kH, fabsf(c0), fmaxf(x0, 0.0f)), fmaf(-kH, fabsf(c1), fmaxf(x1, 0.0f)));
}

// activation storage type helpers: the encoder keeps activations as float (fp32 mode) or
// __half (tensor-core mode); all arithmetic is fp32.
template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<__half>(const __half* p) { return __half2float(*p); }
template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<__half>(__half* p, float v) { *p = __float2half_rn(v); }

// 8 consecutive activations (one lane's share of a 256-wide row: lane*8 .. lane*8+7)
template <typename T> struct Row8;
template <> struct Row8<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};
template <> struct Row8<__half> {
  static __device__ __forceinline__ void load(const __half* p, float (&v)[8]) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
  static __device__ __forceinline__ void store(__half* p, const float (&v)[8]) {
    uint4 u;
    __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};
