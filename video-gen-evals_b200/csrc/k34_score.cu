// K3 centroid segmented reduction, K4 AC/TC scoring, N2 stats sums, N1 TCL forward.
//
// Replaces (reference):
//   utils.py:1035-1043  sums.index_add_(y, z); counts.index_add_(y, 1); normalize(sums / max(counts,1))
//   eval.py:235-255     per-video mean of window embeddings -> normalise -> L2 distance to class centroid
//   eval.py:226         per-video mean of per-window temporal coherence
//   utils.py:589-593    float64 per-column sum / sum of squares
//   losses.py:14-34     TCL forward
// All HBM-bound streaming kernels: one warp owns a 256-float embedding row (8 floats / lane, two
// 16-byte loads); K3 is a deterministic two-pass segmented reduction (see below).
#include "common.cuh"
#include "kernels.h"
#include <math_constants.h>

namespace {

constexpr int kD = TAG_D_MODEL;

// ---------------------------------------------------------------- K3
// Deterministic segmented reduction (utils.py:1035-1041: sums.index_add_(y, z); counts.index_add_(y, 1)).
// Pass 1: every CTA owns a contiguous block of rows and every warp a contiguous slice of it. A warp fetches the labels of
// 32 rows with one coalesced load and walks the rows with the label broadcast by warp shuffle; rows of one label run
// (a video's windows are adjacent) are summed in registers (8 columns per lane, four rows being summed while the next four are in flight) and a finished run is
// added to the warp's PRIVATE [C][257] accumulator in shared memory (plain read-modify-write: no atomics anywhere). The
// warps' accumulators are combined in warp order into one partial per CTA. Pass 2 adds the CTA partials in CTA order into
// sums_counts. The order of every floating-point addition is a function of (n, C) only: results are bit-identical run to run.
constexpr int kRows = 4;   // rows (1 KB each) per batch; two batches in flight
__device__ __forceinline__ int k3_slot(int col) { return (col & 7) * 32 + (col >> 3); }   // lane-major layout: conflict-free flushes

__global__ void __launch_bounds__(256) k_centroid_partial(const float* __restrict__ z, const int32_t* __restrict__ labels,
                                                          int64_t n, int C, float* __restrict__ partial, int64_t rows_per_cta) {
  extern __shared__ float s_acc[];   // [nwarp][C][257]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int E = C * (kD + 1);
  for (int i = threadIdx.x; i < nwarp * E; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  float* mine = s_acc + warp * E;
  const int64_t c0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t c1 = min(n, c0 + rows_per_cta);
  const int64_t per = ((c1 - c0 + nwarp - 1) / nwarp + 31) & ~(int64_t)31;      // whole label groups per warp
  const int64_t r0 = min(c1, c0 + warp * per), r1 = min(c1, r0 + per);
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  int cur = -1;
  float cnt = 0.f;
  auto flush = [&]() {
    if (cur >= 0 && cur < C) {
      float* row = mine + cur * (kD + 1);
#pragma unroll
      for (int k = 0; k < 8; ++k) row[k * 32 + lane] += acc[k];
      if (lane == 0) row[kD] += cnt;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    cnt = 0.f;
  };
  for (int64_t base = r0; base < r1; base += 32) {
    const int m = (int)min((int64_t)32, r1 - base);
    const int my_label = lane < m ? __ldg(labels + base + lane) : -1;
    // four rows are summed while the next four are already in flight (4-8 KB per warp outstanding at any time)
    float nxt[kRows][8];
#pragma unroll
    for (int j = 0; j < kRows; ++j)
      if (j < m) Row8<float>::load(z + (base + j) * kD + lane * 8, nxt[j]);
    for (int i0 = 0; i0 < m; i0 += kRows) {
      float v[kRows][8];
      int y[kRows];
#pragma unroll
      for (int j = 0; j < kRows; ++j) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[j][k] = nxt[j][k];
        y[j] = __shfl_sync(FULL_MASK, my_label, (i0 + j) & 31);
        if (i0 + j >= m) y[j] = -1;
      }
#pragma unroll
      for (int j = 0; j < kRows; ++j)
        if (i0 + kRows + j < m) Row8<float>::load(z + (base + i0 + kRows + j) * kD + lane * 8, nxt[j]);
#pragma unroll
      for (int j = 0; j < kRows; ++j) {
        if (i0 + j >= m) break;
        if (y[j] != cur) { flush(); cur = y[j]; }
        if (y[j] >= 0 && y[j] < C) {
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] += v[j][k];
          cnt += 1.f;
        }
      }
    }
  }
  flush();
  __syncthreads();
  float* out = partial + (size_t)blockIdx.x * E;
  for (int i = threadIdx.x; i < E; i += blockDim.x) {
    const int c = i / (kD + 1), col = i - c * (kD + 1);
    const int slot = c * (kD + 1) + (col < kD ? k3_slot(col) : kD);
    float t = 0.f;
    for (int w = 0; w < nwarp; ++w) t += s_acc[w * E + slot];
    out[i] = t;
  }
}

__global__ void __launch_bounds__(256) k_centroid_combine(const float* __restrict__ partial, int G, int E,
                                                          float* __restrict__ sums_counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= E) return;
  float t = 0.f;
  for (int g = 0; g < G; ++g) t += partial[(size_t)g * E + i];
  sums_counts[i] += t;
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


// ---------------------------------------------------------------- N1: tensor-core TCL helpers
// z (fp32, unit rows) -> hi = half(z), lo = half(z - hi): A = [hi | lo | hi], W = [hi | hi | lo]; rows >= B are zero
__global__ void __launch_bounds__(256) k_tcl_split(const float* __restrict__ z, int64_t B, int64_t Bp, __half* __restrict__ A,
                                                   __half* __restrict__ W) {
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= Bp) return;
  float v[8];
  if (r < B) Row8<float>::load(z + r * kD + lane * 8, v);
  else {
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = 0.f;
  }
  float hi[8], lo[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { hi[k] = __half2float(__float2half_rn(v[k])); lo[k] = v[k] - hi[k]; }
  __half* a = A + r * 3 * kD + lane * 8;
  __half* w = W + r * 3 * kD + lane * 8;
  Row8<__half>::store(a, hi); Row8<__half>::store(a + kD, lo); Row8<__half>::store(a + 2 * kD, hi);
  Row8<__half>::store(w, hi); Row8<__half>::store(w + kD, hi); Row8<__half>::store(w + 2 * kD, lo);
}

// loss_i = log(den_i) - mean_pos S_ij / t from the per-slice partial sums (fixed summation order: deterministic)
__global__ void __launch_bounds__(256) k_tcl_finish(const float* __restrict__ part, int64_t B, int slices, float k1, float k2,
                                                    float* __restrict__ loss_rows) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const float* p = part + (size_t)i * slices * 5;
  float e_pos = 0.f, en_pos = 0.f, e_neg = 0.f, s_pos = 0.f, n_pos = 0.f;
  for (int s = 0; s < slices; ++s) { e_pos += p[0]; en_pos += p[1]; e_neg += p[2]; s_pos += p[3]; n_pos += p[4]; p += 5; }
  const float den = e_pos + k1 * en_pos + k2 * e_neg;
  loss_rows[i] = (n_pos * logf(den) - s_pos) / n_pos;        // 0/0 = NaN when a row has no positives, as the reference
}

// ---------------------------------------------------------------- N1: SupConWithHardNegatives forward (losses.py:37-56)
// CrossEntropy over logits [a.p/t, a.h/t] with the positive as the target = softplus((a.h - a.p)/t); warp per row
__global__ void __launch_bounds__(256) k_supcon_hard(const float* __restrict__ a, const float* __restrict__ pos,
                                                     const float* __restrict__ neg, int64_t B, float inv_temp,
                                                     float* __restrict__ loss_rows) {
  const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= B) return;
  float x[8], p[8], h[8];
  Row8<float>::load(a + i * kD + lane * 8, x);
  Row8<float>::load(pos + i * kD + lane * 8, p);
  Row8<float>::load(neg + i * kD + lane * 8, h);
  float sp = 0.f, sh = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) { sp = fmaf(x[k], p[k], sp); sh = fmaf(x[k], h[k], sh); }
  sp = warp_sum(sp) * inv_temp; sh = warp_sum(sh) * inv_temp;
  if (lane == 0) {
    const float m = fmaxf(sp, sh);                            // log-sum-exp as nn.CrossEntropyLoss evaluates it
    loss_rows[i] = m + logf(expf(sp - m) + expf(sh - m)) - sp;
  }
}

// ---------------------------------------------------------------- N1: frame gather (shuffle / reverse / static windows)
__global__ void __launch_bounds__(256) k_gather_frames(const float4* __restrict__ x, const int32_t* __restrict__ idx, int64_t rows,
                                                       int T, int D4, float4* __restrict__ out) {
  const int64_t r = blockIdx.x;
  if (r >= rows) return;
  const int64_t b = r / T;
  int src = idx[r];
  src = src < 0 ? 0 : (src >= T ? T - 1 : src);
  const float4* in = x + (b * T + src) * D4;
  float4* o = out + r * D4;
  for (int i = threadIdx.x; i < D4; i += blockDim.x) o[i] = __ldg(in + i);
}

}  // namespace

// grid of the deterministic K3 (a function of n and C only) — also the size of the scratch buffer it needs
static void centroid_plan(int64_t n, int C, int* warps, int* ctas, int64_t* rows_per_cta) {
  const size_t per_warp = (size_t)C * (kD + 1) * sizeof(float);
  int w = (int)((200u * 1024u) / per_warp);
  w = w < 1 ? 1 : (w > 8 ? 8 : w);
  const int per_sm = (int)((200u * 1024u) / (per_warp * w)) < 1 ? 1 : (int)((200u * 1024u) / (per_warp * w));
  int64_t g = (n + (int64_t)w * 32 - 1) / ((int64_t)w * 32);
  const int64_t cap = (int64_t)148 * (per_sm > 2 ? 2 : per_sm);
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  int64_t rpc = (n + g - 1) / g;
  rpc = (rpc + 31) & ~(int64_t)31;
  g = (n + rpc - 1) / rpc;
  *warps = w; *ctas = (int)g; *rows_per_cta = rpc;
}

size_t centroid_scratch_floats(int64_t n, int C) {
  int w, g; int64_t rpc;
  centroid_plan(n, C, &w, &g, &rpc);
  return (size_t)g * C * (kD + 1);
}

cudaError_t launch_centroid_accumulate(const float* z, const int32_t* labels, int64_t n, int C, float* sums_counts,
                                       float* scratch, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  int w, g; int64_t rpc;
  centroid_plan(n, C, &w, &g, &rpc);
  const size_t smem = (size_t)w * C * (kD + 1) * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_centroid_partial, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  k_centroid_partial<<<g, w * 32, smem, s>>>(z, labels, n, C, scratch, rpc);
  const int E = C * (kD + 1);
  k_centroid_combine<<<(E + 255) / 256, 256, 0, s>>>(scratch, g, E, sums_counts);
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

cudaError_t launch_tcl_split(const float* z, int64_t B, int64_t Bp, __half* A, __half* W, cudaStream_t s) {
  if (Bp <= 0) return cudaSuccess;
  k_tcl_split<<<(unsigned)((Bp + 7) / 8), 256, 0, s>>>(z, B, Bp, A, W);
  return cudaGetLastError();
}

cudaError_t launch_tcl_finish(const float* part, int64_t B, int slices, float k1, float k2, float* loss_rows, cudaStream_t s) {
  if (B <= 0) return cudaSuccess;
  k_tcl_finish<<<(unsigned)((B + 255) / 256), 256, 0, s>>>(part, B, slices, k1, k2, loss_rows);
  return cudaGetLastError();
}

cudaError_t launch_supcon_hard(const float* anchor, const float* positive, const float* hard, int64_t B, float temperature,
                               float* loss_rows, cudaStream_t s) {
  if (B <= 0) return cudaSuccess;
  k_supcon_hard<<<(unsigned)((B + 7) / 8), 256, 0, s>>>(anchor, positive, hard, B, 1.0f / temperature, loss_rows);
  return cudaGetLastError();
}

cudaError_t launch_gather_frames(const float* x, const int32_t* idx, int64_t B, int T, int D, float* out, cudaStream_t s) {
  if (B <= 0 || T <= 0) return cudaSuccess;
  k_gather_frames<<<(unsigned)(B * T), 256, 0, s>>>(reinterpret_cast<const float4*>(x), idx, B * T, T, D / 4,
                                                    reinterpret_cast<float4*>(out));
  return cudaGetLastError();
}
