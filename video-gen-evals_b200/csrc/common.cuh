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
