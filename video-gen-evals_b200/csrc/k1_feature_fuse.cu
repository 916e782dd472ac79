// K1 feature fuse — one pass from packed per-frame SMPL / keypoint / appearance arrays to the
// encoder input [raw blocks || diff blocks], z-scored.
//
// Replaces (reference, /root/reference):
//   utils.py:366-381  WindowDataset._slice_or_pad     (frame gather with nearest-repeat padding)
//   utils.py:396-404  raw flatten
//   utils.py:142-147  _vit_delta        (cosine: L2-normalise, first difference, row 0 = 0)
//   utils.py:165-174  _rotmat_delta  +  :130-140 _log_so3
//   utils.py:161-163  _betas_delta
//   utils.py:177-217  _procrustes_kp_delta (closed form of `Vh @ U.T` for det(H) > 0; det(H) < 0 frames
//                     are counted in flags[0] — SURVEY.md §8a A6)
//   utils.py:472-514  z-score (x-mean)/(std+1e-6) and concat
//
// HBM-bound: per (window, frame) reads sum(raw_dims)*4 B and writes D*4 B (fp32 feats) and/or D16*2 B
// (padded fp16 operand for the tensor-core encoder). One CTA per window walks its T frames; the
// previous frame is re-read through L1/L2 (same CTA touched it one iteration earlier), so DRAM sees
// each source frame once per window.
#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int kThreads = 256;
constexpr float kEpsStd = 1e-6f;

__device__ __forceinline__ int src_frame(int start, int t, int L) {
  // utils.py:371-381: start outside [0,L) repeats frame 0 / L-1; short tail repeats the last frame
  if (start < 0) return 0;
  int f = start + t;
  return f < L - 1 ? f : L - 1;
}

struct Norm {
  const float* mean;
  const float* stdv;
  __device__ __forceinline__ float operator()(float x, int col) const {
    if (mean == nullptr) return x;
    return (x - __ldg(mean + col)) / (__ldg(stdv + col) + kEpsStd);
  }
};

__global__ void __launch_bounds__(kThreads) k_feature_fuse(const FuseParams p) {
  __shared__ float red[33];
  __shared__ float s_dn[TAG_MAX_MODALITIES];       // previous frame's cosine denominators
  __shared__ float s_kp[2][128];                   // normalised centred keypoints (double buffer)

  const int64_t w = blockIdx.x;
  const int tid = threadIdx.x;
  const int vid = p.win_video[w];
  const int start = p.win_start[w];
  const int64_t f0 = p.frame_offset[vid];
  const int L = (int)(p.frame_offset[vid + 1] - f0);
  const Norm nz{p.mean, p.stdv};
  int n_reflect = 0;

  for (int t = 0; t < p.T; ++t) {
    const int64_t cur = f0 + src_frame(start, t, L);
    const int64_t prv = (t == 0) ? cur : f0 + src_frame(start, t - 1, L);
    float* out = p.feats ? p.feats + ((int64_t)w * p.T + t) * p.D : nullptr;
    __half* out16 = p.feats16 ? p.feats16 + ((int64_t)w * p.T + t) * p.D16 : nullptr;

#pragma unroll 1
    for (int m = 0; m < p.M; ++m) {
      const int dim = p.raw_dim[m];
      const float* xc = p.src[m] + cur * dim;
      const float* xp = p.src[m] + prv * dim;
      const int ro = p.raw_off[m], dofs = p.diff_off[m];
      const int ro16 = p.raw_off16[m], do16 = p.diff_off16[m];
      const int kind = p.kind[m];

      if (kind == TAG_KIND_COSINE) {
        // ---- raw + sum of squares (dim <= 4*kThreads, checked on the host)
        float x[4], xq[4];
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i = tid + j * kThreads;
          x[j] = (i < dim) ? __ldg(xc + i) : 0.f;
          ss += x[j] * x[j];
        }
        ss = block_sum(ss, red);
        const float dn = fmaxf(sqrtf(ss), 1e-12f);          // F.normalize eps
        const float dnp = (t == 0) ? dn : s_dn[m];
        const bool has_diff = p.diff_dim[m] > 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i = tid + j * kThreads;
          xq[j] = (i < dim && has_diff && t > 0) ? __ldg(xp + i) : x[j];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i = tid + j * kThreads;
          if (i < dim) {
            const float r = nz(x[j], ro + i);
            if (out) out[ro + i] = r;
            if (out16) out16[ro16 + i] = __float2half_rn(r);
            if (has_diff) {
              const float d = nz(x[j] / dn - xq[j] / dnp, dofs + i);
              if (out) out[dofs + i] = d;
              if (out16) out16[do16 + i] = __float2half_rn(d);
            }
          }
        }
        __syncthreads();                                    // everyone has read s_dn[m]
        if (tid == 0) s_dn[m] = dn;
      } else if (kind == TAG_KIND_ROTMAT) {
        for (int i = tid; i < dim; i += kThreads) {
          const float r = nz(__ldg(xc + i), ro + i);
          if (out) out[ro + i] = r;
          if (out16) out16[ro16 + i] = __float2half_rn(r);
        }
        const int J = dim / 9;
        if (p.diff_dim[m] > 0 && tid < J) {
          float R[9], Q[9];
#pragma unroll
          for (int k = 0; k < 9; ++k) { R[k] = __ldg(xc + tid * 9 + k); Q[k] = __ldg(xp + tid * 9 + k); }
          // Rrel = Q^T R  (utils.py:172), entries [i][j] = sum_k Q[k][i] R[k][j]
          float E[9];
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
              E[i * 3 + j] = Q[0 * 3 + i] * R[0 * 3 + j] + Q[1 * 3 + i] * R[1 * 3 + j] + Q[2 * 3 + i] * R[2 * 3 + j];
          float tr = E[0] + E[4] + E[8];
          tr = fminf(fmaxf(tr, -1.f + 1e-6f), 3.f - 1e-6f);
          const float theta = acosf((tr - 1.f) / 2.f);
          const float den = fmaxf(2.f * sinf(theta), 1e-6f);
          const float v0 = (E[7] - E[5]) / den, v1 = (E[2] - E[6]) / den, v2 = (E[3] - E[1]) / den;
          const float wv[3] = {theta * v0, theta * v1, theta * v2};
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const float d = nz(wv[k], dofs + tid * 3 + k);
            if (out) out[dofs + tid * 3 + k] = d;
            if (out16) out16[do16 + tid * 3 + k] = __float2half_rn(d);
          }
        }
      } else if (kind == TAG_KIND_PLAIN) {
        for (int i = tid; i < dim; i += kThreads) {
          const float x = __ldg(xc + i);
          const float r = nz(x, ro + i);
          if (out) out[ro + i] = r;
          if (out16) out16[ro16 + i] = __float2half_rn(r);
          if (p.diff_dim[m] > 0) {
            const float d = nz(x - __ldg(xp + i), dofs + i);
            if (out) out[dofs + i] = d;
            if (out16) out16[do16 + i] = __float2half_rn(d);
          }
        }
      } else {  // TAG_KIND_PROCRUSTES: K = dim/2 <= 64 points, handled by warp 0
        for (int i = tid; i < dim; i += kThreads) {
          const float r = nz(__ldg(xc + i), ro + i);
          if (out) out[ro + i] = r;
          if (out16) out16[ro16 + i] = __float2half_rn(r);
        }
        if (p.diff_dim[m] > 0 && tid < 32) {
          const int K = dim / 2;
          const int k0 = tid, k1 = tid + 32;
          const bool a0 = k0 < K, a1 = k1 < K;
          float x0 = a0 ? __ldg(xc + 2 * k0) : 0.f, y0 = a0 ? __ldg(xc + 2 * k0 + 1) : 0.f;
          float x1 = a1 ? __ldg(xc + 2 * k1) : 0.f, y1 = a1 ? __ldg(xc + 2 * k1 + 1) : 0.f;
          const float mx = warp_sum(x0 + x1) / (float)K, my = warp_sum(y0 + y1) / (float)K;   // utils.py:192
          x0 = a0 ? x0 - mx : 0.f; y0 = a0 ? y0 - my : 0.f;
          x1 = a1 ? x1 - mx : 0.f; y1 = a1 ? y1 - my : 0.f;
          const float sc = fmaxf(sqrtf(warp_sum(x0 * x0 + y0 * y0 + x1 * x1 + y1 * y1)), 1e-6f);   // :195
          x0 /= sc; y0 /= sc; x1 /= sc; y1 /= sc;
          float* cb = s_kp[t & 1];
          const float* pb = s_kp[(t & 1) ^ 1];
          cb[2 * k0] = x0; cb[2 * k0 + 1] = y0; cb[2 * k1] = x1; cb[2 * k1 + 1] = y1;
          float d00 = 0.f, d01 = 0.f, d10 = 0.f, d11 = 0.f;
          if (t > 0) {
            const float px0 = pb[2 * k0], py0 = pb[2 * k0 + 1], px1 = pb[2 * k1], py1 = pb[2 * k1 + 1];
            // H = X^T Y (utils.py:207), X = previous frame, Y = current frame
            const float h00 = warp_sum(px0 * x0 + px1 * x1), h01 = warp_sum(px0 * y0 + px1 * y1);
            const float h10 = warp_sum(py0 * x0 + py1 * x1), h11 = warp_sum(py0 * y0 + py1 * y1);
            if (h00 * h11 - h01 * h10 < 0.f) ++n_reflect;
            const float ang = atan2f(h10 - h01, h00 + h11);
            float sn, cs;
            sincosf(ang, &sn, &cs);
            // X @ R with R = [[c, s], [-s, c]]  (== Vh @ U.T for det(H) > 0)
            d00 = x0 - (px0 * cs - py0 * sn); d01 = y0 - (px0 * sn + py0 * cs);
            d10 = x1 - (px1 * cs - py1 * sn); d11 = y1 - (px1 * sn + py1 * cs);
          }
          if (a0) {
            const float e0 = nz(d00, dofs + 2 * k0), e1 = nz(d01, dofs + 2 * k0 + 1);
            if (out) { out[dofs + 2 * k0] = e0; out[dofs + 2 * k0 + 1] = e1; }
            if (out16) { out16[do16 + 2 * k0] = __float2half_rn(e0); out16[do16 + 2 * k0 + 1] = __float2half_rn(e1); }
          }
          if (a1) {
            const float e0 = nz(d10, dofs + 2 * k1), e1 = nz(d11, dofs + 2 * k1 + 1);
            if (out) { out[dofs + 2 * k1] = e0; out[dofs + 2 * k1 + 1] = e1; }
            if (out16) { out16[do16 + 2 * k1] = __float2half_rn(e0); out16[do16 + 2 * k1 + 1] = __float2half_rn(e1); }
          }
          __syncwarp();
        }
      }
    }
  }
  if (p.flags != nullptr && tid == 0 && n_reflect > 0) atomicAdd(p.flags, n_reflect);
}

}  // namespace

cudaError_t launch_feature_fuse(const FuseParams& p, cudaStream_t s) {
  if (p.n_windows <= 0) return cudaSuccess;
  k_feature_fuse<<<(unsigned)p.n_windows, kThreads, 0, s>>>(p);
  return cudaGetLastError();
}
