// fp32 CUDA-core GEMM with temporal-conv taps: the reference-precision mode of the encoder.
//
//   C[r, n] = act( sum_{j<taps} sum_{k<K} A[r + (j - taps/2)*dil, k] * W[n, j*K + k]  + bias[n] + res[r, n] )
//
// rows r = (window, t) with T frames per window; a shifted row outside its window contributes zero
// (nn.Conv1d zero "same" padding, reference model.py:24-30). taps == 1 is a plain x @ W^T (nn.Linear /
// 1x1 conv, model.py:46, :50). 128x128x16 tiles, 8x8 register micro-tile per thread, double-buffered
// shared memory. W rows must be 16-byte aligned with ldw % 4 == 0 (packed by tag_api.cu); A may be
// arbitrary (the stem reads unaligned column slices of the feats tensor).
#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;

template <bool A_VEC>
__global__ void __launch_bounds__(256) k_gemm_f32(const GemmF32 g) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int kt_per_tap = (g.K + BK - 1) / BK;
  const int n_kt = g.taps * kt_per_tap;

  // loader coordinates: 2 float4 per thread per operand
  int lrow[2], lk[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) { const int idx = tid + i * 256; lrow[i] = idx >> 2; lk[i] = (idx & 3) * 4; }
  int64_t arow[2]; int at[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    arow[i] = m0 + lrow[i];
    at[i] = (g.taps > 1) ? (int)(arow[i] % g.T) : 0;
  }

  float4 ra[2], rb[2];
  auto load_tile = [&](int kt) {
    const int j = kt / kt_per_tap;
    const int k0 = (kt - j * kt_per_tap) * BK;
    const int shift = (g.taps > 1) ? (j - g.taps / 2) * g.dil : 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      // ---- A
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      const int ts = at[i] + shift;
      const bool rv = arow[i] < g.M && (g.taps == 1 || (ts >= 0 && ts < g.T));
      const int k = k0 + lk[i];
      if (rv) {
        const float* p = g.A + (arow[i] + shift) * (int64_t)g.lda + k;
        if (A_VEC) {
          if (k + 3 < g.K) v = *reinterpret_cast<const float4*>(p);
          else {
            if (k < g.K) v.x = p[0];
            if (k + 1 < g.K) v.y = p[1];
            if (k + 2 < g.K) v.z = p[2];
          }
        } else {
          if (k < g.K) v.x = __ldg(p);
          if (k + 1 < g.K) v.y = __ldg(p + 1);
          if (k + 2 < g.K) v.z = __ldg(p + 2);
          if (k + 3 < g.K) v.w = __ldg(p + 3);
        }
      }
      ra[i] = v;
      // ---- W  (row n0+lrow, columns j*K + k .. +3; padded rows are zero beyond K up to ldw)
      float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
      const int n = n0 + lrow[i];
      if (n < g.N && k < g.K) {
        const int col = j * g.K + k;
        if (col + 3 < g.ldw) u = *reinterpret_cast<const float4*>(g.W + (int64_t)n * g.ldw + col);
        if (k + 1 >= g.K) u.y = 0.f;
        if (k + 2 >= g.K) u.z = 0.f;
        if (k + 3 >= g.K) u.w = 0.f;
      }
      rb[i] = u;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      As[buf][lk[i] + 0][lrow[i]] = ra[i].x; As[buf][lk[i] + 1][lrow[i]] = ra[i].y;
      As[buf][lk[i] + 2][lrow[i]] = ra[i].z; As[buf][lk[i] + 3][lrow[i]] = ra[i].w;
      Bs[buf][lk[i] + 0][lrow[i]] = rb[i].x; Bs[buf][lk[i] + 1][lrow[i]] = rb[i].y;
      Bs[buf][lk[i] + 2][lrow[i]] = rb[i].z; Bs[buf][lk[i] + 3][lrow[i]] = rb[i].w;
    }
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  load_tile(0);
  store_tile(0);
  __syncthreads();
  for (int kt = 0; kt < n_kt; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < n_kt) load_tile(kt + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < n_kt) {
      store_tile(buf ^ 1);
      __syncthreads();
    }
  }

  // ---- epilogue: v = act(acc + bias + res)
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t r = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (r >= g.M) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = n0 + (h == 0 ? tx * 4 : 64 + tx * 4);
      if (n >= g.N) continue;
      float v[4] = {acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]};
      if (g.bias) {
        const float4 b = *reinterpret_cast<const float4*>(g.bias + n);
        v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
      }
      if (g.res) {
        const float4 q = *reinterpret_cast<const float4*>(g.res + r * (int64_t)g.ldr + n);
        v[0] += q.x; v[1] += q.y; v[2] += q.z; v[3] += q.w;
      }
      if (g.act == 1) {
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = gelu_erf(v[q]);
      } else if (g.act == 2) {
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = fmaxf(v[q], 0.f);
      }
      *reinterpret_cast<float4*>(g.C + r * (int64_t)g.ldc + n) = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
}

}  // namespace

cudaError_t launch_gemm_f32(const GemmF32& g, cudaStream_t s) {
  if (g.M <= 0) return cudaSuccess;
  if ((g.N & 3) || (g.ldc & 3) || (g.ldw & 3) || (g.res && (g.ldr & 3))) return cudaErrorInvalidValue;
  if (g.taps > 1 && (g.K % BK) != 0) return cudaErrorInvalidValue;
  dim3 grid((unsigned)((g.M + BM - 1) / BM), (g.N + BN - 1) / BN);
  const bool a_vec = ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0) && (g.lda % 4 == 0);
  if (a_vec) k_gemm_f32<true><<<grid, 256, 0, s>>>(g);
  else k_gemm_f32<false><<<grid, 256, 0, s>>>(g);
  return cudaGetLastError();
}
