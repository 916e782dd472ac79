// K3 centroid segmented reduction, K4 AC/TC scoring, N2 stats sums, N1 TCL forward.
//
// Replaces (reference):
//   utils.py:1035-1043  sums.index_add_(y, z); counts.index_add_(y, 1); normalize(sums / max(counts,1))
//   eval.py:235-255     per-video mean of window embeddings -> normalise -> L2 distance to class centroid
//   eval.py:226         per-video mean of per-window temporal coherence
//   utils.py:589-593    float64 per-column sum / sum of squares
//   losses.py:14-34     TCL forward
// All HBM-bound streaming kernels: one warp owns a 256-float embedding row (8 floats / lane, two
// 16-byte loads), rows are walked in label runs so the running sum stays in registers and is flushed
// once per run (shared-memory atomics), then once per CTA to global.
#include "common.cuh"
#include "kernels.h"
#include <math_constants.h>

namespace {

constexpr int kD = TAG_D_MODEL;

// ---------------------------------------------------------------- K3
// grid-stride over row blocks; each warp takes a contiguous slice of rows so equal labels (videos are
// stored class-contiguous or at least window-contiguous) form runs.
__global__ void __launch_bounds__(256) k_centroid_accumulate(const float* __restrict__ z,
                                                             const int32_t* __restrict__ labels, int64_t n, int C,
                                                             float* __restrict__ sums_counts, int rows_per_warp) {
  extern __shared__ float s_acc[];   // [C][257]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int i = threadIdx.x; i < C * (kD + 1); i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();

  const int64_t gw = (int64_t)blockIdx.x * nwarp + warp;
  const int64_t r0 = gw * rows_per_warp;
  const int64_t r1 = min(n, r0 + (int64_t)rows_per_warp);
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  int cur = -1;
  float cnt = 0.f;
  for (int64_t r = r0; r < r1; ++r) {
    const int y = labels[r];
    if (y != cur) {
      if (cur >= 0 && cur < C) {
#pragma unroll
        for (int k = 0; k < 8; ++k) atomicAdd(&s_acc[cur * (kD + 1) + lane * 8 + k], acc[k]);
        if (lane == 0) atomicAdd(&s_acc[cur * (kD + 1) + kD], cnt);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = 0.f;
      cnt = 0.f;
      cur = y;
    }
    if (y >= 0 && y < C) {
      float v[8];
      Row8<float>::load(z + r * kD + lane * 8, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += v[k];
      cnt += 1.f;
    }
  }
  if (cur >= 0 && cur < C) {
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(&s_acc[cur * (kD + 1) + lane * 8 + k], acc[k]);
    if (lane == 0) atomicAdd(&s_acc[cur * (kD + 1) + kD], cnt);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * (kD + 1); i += blockDim.x) {
    const float v = s_acc[i];
    if (v != 0.f) atomicAdd(&sums_counts[i], v);
  }
}

__global__ void k_centroid_finalize(const float* __restrict__ sc, int C, float* __restrict__ cen,
                                    float* __restrict__ counts) {
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= C) return;
  const float cnt = sc[c * (kD + 1) + kD];
  const float d = fmaxf(cnt, 1.0f);                 // counts.clamp_min(1.0)
  float v[8];
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) { v[k] = sc[c * (kD + 1) + lane * 8 + k] / d; ss += v[k] * v[k]; }
  ss = warp_sum(ss);
  const float dn = fmaxf(sqrtf(ss), 1e-12f);        // F.normalize
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = v[k] / dn;
  Row8<float>::store(cen + c * kD + lane * 8, v);
  if (counts != nullptr && lane == 0) counts[c] = cnt;
}

// ---------------------------------------------------------------- K4: one warp per video
__global__ void __launch_bounds__(256) k_score(const float* __restrict__ seq, const float* __restrict__ tcw,
                                               const int64_t* __restrict__ seg, const int32_t* __restrict__ label,
                                               const float* __restrict__ cen, int C, int64_t V,
                                               float* __restrict__ ac, float* __restrict__ tc) {
  const int64_t v = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (v >= V) return;
  const int64_t a = seg[v], b = seg[v + 1];
  const int64_t nwin = b - a;
  float m[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) m[k] = 0.f;
  float tsum = 0.f;
  if (seq != nullptr) {
    for (int64_t r = a; r < b; ++r) {
      float x[8];
      Row8<float>::load(seq + r * kD + lane * 8, x);
#pragma unroll
      for (int k = 0; k < 8; ++k) m[k] += x[k];
    }
  }
  if (tcw != nullptr) {
    for (int64_t r = a + lane; r < b; r += 32) tsum += tcw[r];
    tsum = warp_sum(tsum);
  }
  const int y = (seq != nullptr) ? label[v] : -1;
  float out_ac = CUDART_NAN_F;
  if (nwin > 0 && y >= 0 && y < C) {
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { m[k] = m[k] / (float)nwin; ss += m[k] * m[k]; }
    ss = warp_sum(ss);
    const float dn = fmaxf(sqrtf(ss), 1e-12f);
    float c[8];
    Row8<float>::load(cen + (int64_t)y * kD + lane * 8, c);
    float d2 = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { const float d = m[k] / dn - c[k]; d2 += d * d; }
    d2 = warp_sum(d2);
    out_ac = sqrtf(d2);
  }
  if (lane == 0) {
    if (ac != nullptr) ac[v] = out_ac;
    if (tcw != nullptr && tc != nullptr) tc[v] = nwin > 0 ? tsum / (float)nwin : CUDART_NAN_F;
  }
}

// ---------------------------------------------------------------- N2: per-column double sums
// block = 256 threads = 256 consecutive columns; grid.y splits the rows.
__global__ void __launch_bounds__(256) k_stats_accumulate(const float* __restrict__ x, int64_t rows, int D,
                                                          double* __restrict__ sum, double* __restrict__ sumsq,
                                                          int rows_per_block) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= D) return;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = min(rows, r0 + (int64_t)rows_per_block);
  double s = 0.0, q = 0.0;
  for (int64_t r = r0; r < r1; ++r) {
    const double v = (double)x[r * D + col];
    s += v;
    q += v * v;
  }
  atomicAdd(sum + col, s);
  atomicAdd(sumsq + col, q);
}

// ---------------------------------------------------------------- N1: TCL forward, one warp per anchor row
// loss_i = log(den_i) - mean_{j in pos(i)} S_ij / temperature,
// den_i = sum_pos exp(S/temp) + k1 sum_pos exp(-S) + k2 sum_neg exp(S/temp)      (losses.py:18-31)
__global__ void __launch_bounds__(256) k_tcl_forward(const float* __restrict__ z, const int32_t* __restrict__ y,
                                                     int64_t B, float inv_temp, float k1, float k2,
                                                     float* __restrict__ loss_rows) {
  const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= B) return;
  float a[8];
  Row8<float>::load(z + i * kD + lane * 8, a);
  const int yi = y[i];
  float e_pos = 0.f, en_pos = 0.f, e_neg = 0.f, s_pos = 0.f, n_pos = 0.f;
  for (int64_t j = 0; j < B; ++j) {
    float b[8];
    Row8<float>::load(z + j * kD + lane * 8, b);
    float d = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) d += a[k] * b[k];
    d = warp_sum(d);
    const float e = expf(d * inv_temp);
    if (y[j] == yi) {
      if (j != i) { e_pos += e; en_pos += expf(-d); s_pos += d * inv_temp; n_pos += 1.f; }
    } else {
      e_neg += e;
    }
  }
  if (lane == 0) {
    const float den = e_pos + k1 * en_pos + k2 * e_neg;
    loss_rows[i] = (n_pos * logf(den) - s_pos) / n_pos;      // 0/0 = NaN when a row has no positives, as the reference
  }
}

}  // namespace

cudaError_t launch_centroid_accumulate(const float* z, const int32_t* labels, int64_t n, int C, float* sums_counts,
                                       cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  const int warps = 8;
  int rows_per_warp = 32;
  // keep the grid around a few waves of 148 SMs
  int64_t blocks = (n + (int64_t)warps * rows_per_warp - 1) / ((int64_t)warps * rows_per_warp);
  while (blocks > 148 * 16) { rows_per_warp *= 2; blocks = (n + (int64_t)warps * rows_per_warp - 1) / ((int64_t)warps * rows_per_warp); }
  const size_t smem = (size_t)C * (kD + 1) * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_centroid_accumulate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  k_centroid_accumulate<<<(unsigned)blocks, warps * 32, smem, s>>>(z, labels, n, C, sums_counts, rows_per_warp);
  return cudaGetLastError();
}

cudaError_t launch_centroid_finalize(const float* sums_counts, int C, float* centroids, float* counts, cudaStream_t s) {
  if (C <= 0) return cudaSuccess;
  k_centroid_finalize<<<(C + 3) / 4, 128, 0, s>>>(sums_counts, C, centroids, counts);
  return cudaGetLastError();
}

cudaError_t launch_score(const float* seq, const float* tcw, const int64_t* seg, const int32_t* label, const float* cen,
                         int C, int64_t V, float* ac, float* tc, cudaStream_t s) {
  if (V <= 0) return cudaSuccess;
  k_score<<<(unsigned)((V + 7) / 8), 256, 0, s>>>(seq, tcw, seg, label, cen, C, V, ac, tc);
  return cudaGetLastError();
}

cudaError_t launch_stats_accumulate(const float* x, int64_t rows, int D, double* sum, double* sumsq, cudaStream_t s) {
  if (rows <= 0 || D <= 0) return cudaSuccess;
  const int gx = (D + 255) / 256;
  int gy = (int)min((int64_t)((148 * 8 + gx - 1) / gx), rows);
  if (gy < 1) gy = 1;
  const int rpb = (int)((rows + gy - 1) / gy);
  gy = (int)((rows + rpb - 1) / rpb);
  k_stats_accumulate<<<dim3(gx, gy), 256, 0, s>>>(x, rows, D, sum, sumsq, rpb);
  return cudaGetLastError();
}

cudaError_t launch_tcl_forward(const float* z, const int32_t* y, int64_t B, float temperature, float k1, float k2,
                               float* loss_rows, cudaStream_t s) {
  if (B <= 0) return cudaSuccess;
  k_tcl_forward<<<(unsigned)((B + 7) / 8), 256, 0, s>>>(z, y, B, 1.0f / temperature, k1, k2, loss_rows);
  return cudaGetLastError();
}
