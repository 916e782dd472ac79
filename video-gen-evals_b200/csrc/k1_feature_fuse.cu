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
// (padded fp16 operand for the tensor-core encoder). One warp per (window, 8 consecutive frames); overlapping
// windows re-read source frames through L2, so DRAM sees each source frame about once or twice per batch.
#include <stdlib.h>

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

// z-score as one FMA per element: (x - mean) / (std + 1e-6) == x * scale + shift with scale = 1/(std+1e-6) and
// shift = -mean*scale, tabulated once per call by k_zscore_table (an IEEE division sequence per output element made
// this HBM-bound kernel instruction-bound: ncu counted 1.9e9 warp instructions per 12.5k windows).
struct Norm {
  const float* scale;
  const float* shift;
  __device__ __forceinline__ float operator()(float x, int col) const {
    if (scale == nullptr) return x;
    return fmaf(x, __ldg(scale + col), __ldg(shift + col));
  }
};

__global__ void k_zscore_table(const float* __restrict__ mean, const float* __restrict__ stdv, float* __restrict__ scale,
                               float* __restrict__ shift, int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < D) {
    const float sc = 1.0f / (stdv[i] + kEpsStd);
    scale[i] = sc;
    shift[i] = -mean[i] * sc;
  }
}



// One warp per (window, block of kF = 8 consecutive frames). Every reduction (cosine norms, keypoint centre / scale /
// 2x2 correlation) is a warp shuffle: no block barrier, no shared memory. For the wide cosine modalities (vit / clip /
// dino: 70 % of the bytes) the warp first computes the kF+1 row norms, then walks the columns in chunks of 64: the four
// z-score table entries of a column pair are loaded ONCE per chunk and the previous frame's values are carried in
// registers from one frame to the next, so per output element the kernel issues one 8-byte load, ~5 FMAs and two
// 4-byte stores. The small modalities (rotations, betas, keypoints: 346 of 1370 input floats) are handled per frame.
// Two launches per call: the cosine modalities with KF = 8 frames per warp, everything else with KF = 2 (measured:
// run together in one warp the two halves take 3.1 ms per 12.5k windows, separately 1.3 + 0.9 ms — the small
// modalities are a chain of dependent load -> shuffle -> store phases that wants many short warps).

template <bool O32, bool O16>
__device__ __forceinline__ void put1(float* out, __half* out16, int c32, int c16, float v) {
  if (O32) out[c32] = v;
  if (O16) out16[c16] = __float2half_rn(v);
}

template <bool O32, bool O16, bool COSINE_PART, int kF>
__global__ void __launch_bounds__(kThreads, 4) k_feature_fuse(const FuseParams p) {
  const int lane = threadIdx.x & 31;
  const int blocks_per_win = (p.T + kF - 1) / kF;
  const int64_t gw = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (gw >= p.n_windows * blocks_per_win) return;
  const int64_t w = gw / blocks_per_win;
  const int t0 = (int)(gw - w * blocks_per_win) * kF;
  const int nf = min(kF, p.T - t0);                    // frames of this block
  const int vid = p.win_video[w];
  const int start = p.win_start[w];
  const int64_t f0 = p.frame_offset[vid];
  const int L = (int)(p.frame_offset[vid + 1] - f0);
  const Norm nz{p.mean, p.stdv};          // (scale, shift) tables when stats are given
  // source row of window frame t (t = t0-1 .. t0+nf-1); frame -1 of the window pairs with itself (zero delta)
  auto row_of = [&](int t) -> int64_t { return f0 + src_frame(start, t < 0 ? 0 : t, L); };
  float* outw = O32 ? p.feats + ((int64_t)w * p.T + t0) * p.D : nullptr;
  __half* outw16 = O16 ? p.feats16 + ((int64_t)w * p.T + t0) * p.D16 : nullptr;

#pragma unroll 1
  for (int m = 0; m < p.M; ++m) {
    const int dim = p.raw_dim[m];
    const float* src = p.src[m];
    const int ro = p.raw_off[m], dofs = p.diff_off[m];
    const int ro16 = p.raw_off16[m], do16 = p.diff_off16[m];
    const int kind = p.kind[m];
    const bool has_diff = p.diff_dim[m] > 0;
    if ((kind == TAG_KIND_COSINE) != COSINE_PART) continue;

    if (kind == TAG_KIND_COSINE) {
      // ---- pass 1: 1 / max(||row||, 1e-12) of rows t0-1 .. t0+nf-1  (F.normalize eps)
      float inv[kF + 1];
#pragma unroll
      for (int f = 0; f <= kF; ++f) {
        inv[f] = 0.f;
        if (f <= nf && (f > 0 || has_diff)) {
          const float* x = src + row_of(t0 + f - 1) * dim;
          float ss = 0.f;
#pragma unroll 4
          for (int i = 2 * lane; i < dim; i += 64) {
            const float2 a = __ldg(reinterpret_cast<const float2*>(x + i));
            ss = fmaf(a.x, a.x, ss); ss = fmaf(a.y, a.y, ss);
          }
          inv[f] = 1.0f / fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
        }
      }
      // ---- pass 2: column chunks; tables once per chunk, previous frame carried in registers
      const bool vec32 = O32 && ((p.D | ro | dofs) & 1) == 0;
#pragma unroll 1
      for (int i = 2 * lane; i < dim; i += 64) {
        float2 sr = make_float2(1.f, 1.f), hr = make_float2(0.f, 0.f), sd = sr, hd = hr;
        if (nz.scale != nullptr) {
          sr = make_float2(__ldg(nz.scale + ro + i), __ldg(nz.scale + ro + i + 1));
          hr = make_float2(__ldg(nz.shift + ro + i), __ldg(nz.shift + ro + i + 1));
          if (has_diff) {
            sd = make_float2(__ldg(nz.scale + dofs + i), __ldg(nz.scale + dofs + i + 1));
            hd = make_float2(__ldg(nz.shift + dofs + i), __ldg(nz.shift + dofs + i + 1));
          }
        }
        float2 prev = make_float2(0.f, 0.f);
        if (has_diff) {
          prev = __ldg(reinterpret_cast<const float2*>(src + row_of(t0 - 1) * dim + i));
          prev.x *= inv[0]; prev.y *= inv[0];
        }
#pragma unroll
        for (int f = 0; f < kF; ++f) {
          if (f < nf) {
            const float2 a = __ldg(reinterpret_cast<const float2*>(src + row_of(t0 + f) * dim + i));
            const float r0 = fmaf(a.x, sr.x, hr.x), r1 = fmaf(a.y, sr.y, hr.y);
            float* o = outw + (int64_t)f * p.D;
            __half* o16 = outw16 + (int64_t)f * p.D16;
            if (O32) {
              if (vec32) *reinterpret_cast<float2*>(o + ro + i) = make_float2(r0, r1);
              else { o[ro + i] = r0; o[ro + i + 1] = r1; }
            }
            if (O16) *reinterpret_cast<__half2*>(o16 + ro16 + i) = __floats2half2_rn(r0, r1);
            if (has_diff) {
              const float2 cur = make_float2(a.x * inv[f + 1], a.y * inv[f + 1]);
              const float d0 = fmaf(cur.x - prev.x, sd.x, hd.x), d1 = fmaf(cur.y - prev.y, sd.y, hd.y);
              if (O32) {
                if (vec32) *reinterpret_cast<float2*>(o + dofs + i) = make_float2(d0, d1);
                else { o[dofs + i] = d0; o[dofs + i + 1] = d1; }
              }
              if (O16) *reinterpret_cast<__half2*>(o16 + do16 + i) = __floats2half2_rn(d0, d1);
              prev = cur;
            }
          }
        }
      }
      continue;
    }

    // ---- small modalities: per frame
#pragma unroll 1
    for (int f = 0; f < nf; ++f) {
      const int t = t0 + f;
      const float* xc = src + row_of(t) * dim;
      const float* xp = src + row_of(t - 1) * dim;
      float* out = O32 ? outw + (int64_t)f * p.D : nullptr;
      __half* out16 = O16 ? outw16 + (int64_t)f * p.D16 : nullptr;
      if (kind == TAG_KIND_ROTMAT) {
        for (int i = lane; i < dim; i += 32) put1<O32, O16>(out, out16, ro + i, ro16 + i, nz(__ldg(xc + i), ro + i));
        const int J = dim / 9;
        if (has_diff) {
          for (int jn = lane; jn < J; jn += 32) {
            float R[9], Q[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) { R[k] = __ldg(xc + jn * 9 + k); Q[k] = __ldg(xp + jn * 9 + k); }
            // Rrel = Q^T R  (utils.py:172), entries [i][j] = sum_k Q[k][i] R[k][j]
            float E[9];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
              for (int j = 0; j < 3; ++j)
                E[i * 3 + j] = Q[0 * 3 + i] * R[0 * 3 + j] + Q[1 * 3 + i] * R[1 * 3 + j] + Q[2 * 3 + i] * R[2 * 3 + j];
            float tr = E[0] + E[4] + E[8];
            tr = fminf(fmaxf(tr, -1.f + 1e-6f), 3.f - 1e-6f);
            const float c = (tr - 1.f) / 2.f;
            const float theta = acosf(c);
            // 2 sin(theta) with theta in [0, pi]: sin = sqrt((1-c)(1+c)); 1-c is exact in fp32 near c = 1, so this is
            // at least as accurate as sinf(acosf(c)) and costs one sqrt instead of a libm sine
            const float den = fmaxf(2.f * sqrtf((1.f - c) * (1.f + c)), 1e-6f);
            const float k = theta / den;
            const float wv[3] = {k * (E[7] - E[5]), k * (E[2] - E[6]), k * (E[3] - E[1])};
#pragma unroll
            for (int q = 0; q < 3; ++q) put1<O32, O16>(out, out16, dofs + jn * 3 + q, do16 + jn * 3 + q, nz(wv[q], dofs + jn * 3 + q));
          }
        }
      } else if (kind == TAG_KIND_PLAIN) {
        for (int i = lane; i < dim; i += 32) {
          const float x = __ldg(xc + i);
          put1<O32, O16>(out, out16, ro + i, ro16 + i, nz(x, ro + i));
          if (has_diff) put1<O32, O16>(out, out16, dofs + i, do16 + i, nz(x - __ldg(xp + i), dofs + i));
        }
      } else {  // TAG_KIND_PROCRUSTES: K = dim/2 <= 64 points, two per lane
        for (int i = lane; i < dim; i += 32) put1<O32, O16>(out, out16, ro + i, ro16 + i, nz(__ldg(xc + i), ro + i));
        if (has_diff) {
          const int K = dim / 2;
          const int k0 = lane, k1 = lane + 32;
          const bool a0 = k0 < K, a1 = k1 < K;
          // centre + Frobenius-normalise one frame's points (utils.py:192-196)
          auto load_norm = [&](const float* x, float& x0, float& y0, float& x1, float& y1) {
            x0 = a0 ? __ldg(x + 2 * k0) : 0.f; y0 = a0 ? __ldg(x + 2 * k0 + 1) : 0.f;
            x1 = a1 ? __ldg(x + 2 * k1) : 0.f; y1 = a1 ? __ldg(x + 2 * k1 + 1) : 0.f;
            const float mx = warp_sum(x0 + x1) / (float)K, my = warp_sum(y0 + y1) / (float)K;
            x0 = a0 ? x0 - mx : 0.f; y0 = a0 ? y0 - my : 0.f;
            x1 = a1 ? x1 - mx : 0.f; y1 = a1 ? y1 - my : 0.f;
            const float isc = 1.0f / fmaxf(sqrtf(warp_sum(x0 * x0 + y0 * y0 + x1 * x1 + y1 * y1)), 1e-6f);
            x0 *= isc; y0 *= isc; x1 *= isc; y1 *= isc;
          };
          float x0, y0, x1, y1, px0, py0, px1, py1;
          load_norm(xc, x0, y0, x1, y1);
          load_norm(xp, px0, py0, px1, py1);
          float d00 = 0.f, d01 = 0.f, d10 = 0.f, d11 = 0.f;
          if (t > 0) {
            // H = X^T Y (utils.py:207), X = previous frame, Y = current frame
            const float h00 = warp_sum(px0 * x0 + px1 * x1), h01 = warp_sum(px0 * y0 + px1 * y1);
            const float h10 = warp_sum(py0 * x0 + py1 * x1), h11 = warp_sum(py0 * y0 + py1 * y1);
            if (h00 * h11 - h01 * h10 < 0.f && lane == 0 && p.flags != nullptr) atomicAdd(p.flags, 1);
            // R = [[c, s], [-s, c]] with angle atan2(h10 - h01, h00 + h11): cos and sin are just the normalised pair
            const float ry = h10 - h01, rx = h00 + h11;
            const float rr = ry * ry + rx * rx;
            const float ir = rr > 0.f ? rsqrtf(rr) : 0.f;
            const float cs = rr > 0.f ? rx * ir : 1.f, sn = ry * ir;
            // X @ R with R = [[c, s], [-s, c]]  (== Vh @ U.T for det(H) > 0)
            d00 = x0 - (px0 * cs - py0 * sn); d01 = y0 - (px0 * sn + py0 * cs);
            d10 = x1 - (px1 * cs - py1 * sn); d11 = y1 - (px1 * sn + py1 * cs);
          }
          if (a0) {
            put1<O32, O16>(out, out16, dofs + 2 * k0, do16 + 2 * k0, nz(d00, dofs + 2 * k0));
            put1<O32, O16>(out, out16, dofs + 2 * k0 + 1, do16 + 2 * k0 + 1, nz(d01, dofs + 2 * k0 + 1));
          }
          if (a1) {
            put1<O32, O16>(out, out16, dofs + 2 * k1, do16 + 2 * k1, nz(d10, dofs + 2 * k1));
            put1<O32, O16>(out, out16, dofs + 2 * k1 + 1, do16 + 2 * k1 + 1, nz(d11, dofs + 2 * k1 + 1));
          }
        }
      }
    }
  }
}


// ------------------------------------------------------------------------------------------------------------------
// Shared-memory staged variant for the tensor-core path (fp16 padded output only).
// One CTA = one window x kS = 8 consecutive frames, one warp per frame:
//   load    the kS+1 source rows of every modality land in shared memory first — the wide cosine rows by
//           cp.async.bulk (one 16-byte-aligned row per copy, completion on an mbarrier), the small modalities by
//           coalesced loads issued up front by all 256 threads — so a CTA has ~49 KB of reads in flight at once
//           instead of a chain of dependent per-modality round trips;
//   compute each warp builds its frame's [raw || diff] row (z-scored, fp16, pad columns zero) in shared memory;
//   store   the kS output rows of a block are contiguous in feats16: ONE cp.async.bulk shared->global per CTA.
// kS = 4 frames per CTA: 5 input rows (27 KB) + 4 output rows (23 KB) -> 4 CTAs (32 warps) per SM, so one CTA's load
// phase overlaps the others' compute/store phases (kS = 8 left only 2 CTAs per SM and measured slower).
constexpr int kStagedFrames = 4;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int kS>
__global__ void __launch_bounds__(256) k_feature_fuse_staged(const FuseParams p, int in_floats_per_row, int smem_in_bytes, int dbg) {
  extern __shared__ __align__(128) unsigned char sm_raw[];
  float* s_in = reinterpret_cast<float*>(sm_raw);                                   // [kS+1][in_floats_per_row]
  __half* s_out = reinterpret_cast<__half*>(sm_raw + smem_in_bytes);               // [kS][D16]
  float* s_inv = reinterpret_cast<float*>(sm_raw + smem_in_bytes + kS * p.D16 * 2); // [M][kS+1]
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_inv + TAG_MAX_MODALITIES * (kS + 1) + 2);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int blocks_per_win = (p.T + kS - 1) / kS;
  const int64_t w = blockIdx.x / blocks_per_win;
  const int t0 = (int)(blockIdx.x - w * blocks_per_win) * kS;
  const int nf = min(kS, p.T - t0);
  const int vid = p.win_video[w];
  const int start = p.win_start[w];
  const int64_t f0 = p.frame_offset[vid];
  const int L = (int)(p.frame_offset[vid + 1] - f0);
  auto row_of = [&](int t) -> int64_t { return f0 + src_frame(start, t < 0 ? 0 : t, L); };
  const uint32_t bar = smem_addr(s_bar);

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // zero the output tile (pad columns stay zero)
  for (int i = tid; i < kS * p.D16 / 8; i += 256) reinterpret_cast<uint4*>(s_out)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();

  // ---- load phase
  int col = 0;                                   // float offset of modality m inside a staged row
  uint32_t tx = 0;
#pragma unroll 1
  for (int m = 0; m < p.M; ++m) {
    const int dim = p.raw_dim[m];
    if (p.kind[m] == TAG_KIND_COSINE) tx += (uint32_t)((nf + 1) * dim * 4);
    col += (dim + 3) & ~3;
  }
  if (dbg & 4) tx = 0;                           // experiment: no loads
  if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(tx) : "memory");
  __syncthreads();
  col = 0;
#pragma unroll 1
  for (int m = 0; m < p.M; ++m) {
    const int dim = p.raw_dim[m];
    const float* src = p.src[m];
    if (dbg & 4) { col += (dim + 3) & ~3; continue; }
    if (p.kind[m] == TAG_KIND_COSINE) {
      if (tid <= nf) {                           // one bulk copy per row (rows may repeat when the window is padded)
        const float* g = src + row_of(t0 + tid - 1) * dim;
        const uint32_t d = smem_addr(s_in + tid * in_floats_per_row + col);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(d), "l"(g), "r"((uint32_t)(dim * 4)), "r"(bar) : "memory");
      }
    } else {
      // warp r stages rows r and r+8 (only warp 0 has a second row). Explicit register batches: all loads of both rows
      // are issued before the first store — the compiler's own unrolling left a one-load-one-store remainder loop
      // (dim < 256) that paid a DRAM round trip per element, 28 % of the kernel's stall samples.
      const int r1 = warp, r2 = warp + 8;
      const float* g1 = src + row_of(t0 + r1 - 1) * dim;
      const float* g2 = src + row_of(t0 + (r2 <= nf ? r2 : r1) - 1) * dim;
      for (int i0 = 0; i0 < dim; i0 += 256) {
        float v[8], u[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int i = i0 + k * 32 + lane;
          v[k] = (r1 <= nf && i < dim) ? __ldg(g1 + i) : 0.f;
          u[k] = (r2 <= nf && i < dim) ? __ldg(g2 + i) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int i = i0 + k * 32 + lane;
          if (r1 <= nf && i < dim) s_in[r1 * in_floats_per_row + col + i] = v[k];
          if (r2 <= nf && i < dim) s_in[r2 * in_floats_per_row + col + i] = u[k];
        }
      }
    }
    col += (dim + 3) & ~3;
  }
  {  // wait for the bulk copies (bounded: a protocol bug must not hang the GPU)
    uint32_t ok = 0;
    const long long c0 = clock64();
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(bar), "r"(0) : "memory");
      if (!ok && clock64() - c0 > 4000000000LL) __trap();
    }
  }
  __syncthreads();

  // ---- row norms of the cosine modalities: one row per warp
  col = 0;
#pragma unroll 1
  for (int m = 0; m < ((dbg & 1) ? 0 : p.M); ++m) {
    const int dim = p.raw_dim[m];
    if (p.kind[m] == TAG_KIND_COSINE) {
      for (int r = warp; r <= nf; r += 8) {                               // one row per warp (two for warp 0 when kS = 8)
        {
          const float* x = s_in + r * in_floats_per_row + col;
          float ss = 0.f;
          for (int i = 2 * lane; i < dim; i += 64) {
            const float2 a = *reinterpret_cast<const float2*>(x + i);
            ss = fmaf(a.x, a.x, ss); ss = fmaf(a.y, a.y, ss);
          }
          ss = warp_sum(ss);
          if (lane == 0) s_inv[m * (kS + 1) + r] = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
        }
      }
    }
    col += (dim + 3) & ~3;
  }
  __syncthreads();

  // ---- compute phase, item-parallel over the whole CTA (no warp-serial chains): a thread owns a column (pair) for
  // all nf frames, so its two z-score table entries are loaded once and stay in registers; rotation joints and plain
  // differences are spread as (frame, item) pairs; only the keypoint alignment is a warp-per-frame reduction.
  const Norm nz{p.mean, p.stdv};
  col = 0;
#pragma unroll 1
  for (int m = 0; m < ((dbg & 1) ? 0 : p.M); ++m) {
    const int dim = p.raw_dim[m];
    const float* xin = s_in + col;                       // row r of this modality: xin + r * in_floats_per_row
    const int ro = p.raw_off[m], dofs = p.diff_off[m];
    const int ro16 = p.raw_off16[m], do16 = p.diff_off16[m];
    const int kind = p.kind[m];
    const bool has_diff = p.diff_dim[m] > 0;
    col += (dim + 3) & ~3;

    if (kind == TAG_KIND_COSINE) {
      const float* invm = s_inv + m * (kS + 1);
      for (int i = 2 * tid; i < dim; i += 512) {
        float2 sr = make_float2(1.f, 1.f), hr = make_float2(0.f, 0.f), sd = sr, hd = hr;
        if (nz.scale != nullptr) {
          sr = *reinterpret_cast<const float2*>(nz.scale + ro + i);    // ro, dofs even on this path (checked on the host)
          hr = *reinterpret_cast<const float2*>(nz.shift + ro + i);
          if (has_diff) {
            sd = *reinterpret_cast<const float2*>(nz.scale + dofs + i);
            hd = *reinterpret_cast<const float2*>(nz.shift + dofs + i);
          }
        }
        float2 prev = *reinterpret_cast<const float2*>(xin + i);
        prev.x *= invm[0]; prev.y *= invm[0];
#pragma unroll
        for (int f = 0; f < kS; ++f) {
          if (f < nf) {
            const float2 a = *reinterpret_cast<const float2*>(xin + (f + 1) * in_floats_per_row + i);
            __half* o16 = s_out + f * p.D16;
            *reinterpret_cast<__half2*>(o16 + ro16 + i) = __floats2half2_rn(fmaf(a.x, sr.x, hr.x), fmaf(a.y, sr.y, hr.y));
            if (has_diff) {
              const float2 cur = make_float2(a.x * invm[f + 1], a.y * invm[f + 1]);
              *reinterpret_cast<__half2*>(o16 + do16 + i) =
                  __floats2half2_rn(fmaf(cur.x - prev.x, sd.x, hd.x), fmaf(cur.y - prev.y, sd.y, hd.y));
              prev = cur;
            }
          }
        }
      }
      continue;
    }
    // raw columns of the small modalities: one column per thread, all frames
    for (int i = tid; i < dim; i += 256) {
      const float sc = nz.scale ? __ldg(nz.scale + ro + i) : 1.f, sh = nz.scale ? __ldg(nz.shift + ro + i) : 0.f;
#pragma unroll
      for (int f = 0; f < kS; ++f)
        if (f < nf) s_out[f * p.D16 + ro16 + i] = __float2half_rn(fmaf(xin[(f + 1) * in_floats_per_row + i], sc, sh));
    }
    if (!has_diff) continue;
    if (kind == TAG_KIND_ROTMAT) {
      const int J = dim / 9;
      for (int jn = lane; warp < nf && jn < J; jn += 32) {     // warp = frame, lane = joint
        const int f = warp;
        const float* xc = xin + (f + 1) * in_floats_per_row + jn * 9;
        const float* xp = xin + f * in_floats_per_row + jn * 9;
        float R[9], Q[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) { R[k] = xc[k]; Q[k] = xp[k]; }
        // Rrel = Q^T R  (utils.py:172), entries [i][j] = sum_k Q[k][i] R[k][j]
        float E[9];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j)
            E[i * 3 + j] = Q[0 * 3 + i] * R[0 * 3 + j] + Q[1 * 3 + i] * R[1 * 3 + j] + Q[2 * 3 + i] * R[2 * 3 + j];
        float tr = E[0] + E[4] + E[8];
        tr = fminf(fmaxf(tr, -1.f + 1e-6f), 3.f - 1e-6f);
        const float c = (tr - 1.f) / 2.f;
        const float theta = acosf(c);
        const float den = fmaxf(2.f * sqrtf((1.f - c) * (1.f + c)), 1e-6f);
        const float k = theta / den;
        const float wv[3] = {k * (E[7] - E[5]), k * (E[2] - E[6]), k * (E[3] - E[1])};
#pragma unroll
        for (int q = 0; q < 3; ++q) s_out[f * p.D16 + do16 + jn * 3 + q] = __float2half_rn(nz(wv[q], dofs + jn * 3 + q));
      }
    } else if (kind == TAG_KIND_PLAIN) {
      for (int i = lane; warp < nf && i < dim; i += 32) {      // warp = frame, lane = column
        const int f = warp;
        const float d = xin[(f + 1) * in_floats_per_row + i] - xin[f * in_floats_per_row + i];
        s_out[f * p.D16 + do16 + i] = __float2half_rn(nz(d, dofs + i));
      }
    } else if (warp < nf) {  // TAG_KIND_PROCRUSTES: warp f aligns frame t0+f-1 -> t0+f
      const int f = warp, t = t0 + f;
      const float* xc = xin + (f + 1) * in_floats_per_row;
      const float* xp = xin + f * in_floats_per_row;
      __half* o16 = s_out + f * p.D16;
      const int K = dim / 2;
      const int k0 = lane, k1 = lane + 32;
      const bool a0 = k0 < K, a1 = k1 < K;
      // centre + Frobenius-normalise one frame's points (utils.py:192-196)
      auto load_norm = [&](const float* x, float& x0, float& y0, float& x1, float& y1) {
        x0 = a0 ? x[2 * k0] : 0.f; y0 = a0 ? x[2 * k0 + 1] : 0.f;
        x1 = a1 ? x[2 * k1] : 0.f; y1 = a1 ? x[2 * k1 + 1] : 0.f;
        const float mx = warp_sum(x0 + x1) / (float)K, my = warp_sum(y0 + y1) / (float)K;
        x0 = a0 ? x0 - mx : 0.f; y0 = a0 ? y0 - my : 0.f;
        x1 = a1 ? x1 - mx : 0.f; y1 = a1 ? y1 - my : 0.f;
        const float isc = 1.0f / fmaxf(sqrtf(warp_sum(x0 * x0 + y0 * y0 + x1 * x1 + y1 * y1)), 1e-6f);
        x0 *= isc; y0 *= isc; x1 *= isc; y1 *= isc;
      };
      float x0, y0, x1, y1, px0, py0, px1, py1;
      load_norm(xc, x0, y0, x1, y1);
      load_norm(xp, px0, py0, px1, py1);
      float d00 = 0.f, d01 = 0.f, d10 = 0.f, d11 = 0.f;
      if (t > 0) {
        // H = X^T Y (utils.py:207), X = previous frame, Y = current frame
        const float h00 = warp_sum(px0 * x0 + px1 * x1), h01 = warp_sum(px0 * y0 + px1 * y1);
        const float h10 = warp_sum(py0 * x0 + py1 * x1), h11 = warp_sum(py0 * y0 + py1 * y1);
        if (h00 * h11 - h01 * h10 < 0.f && lane == 0 && p.flags != nullptr) atomicAdd(p.flags, 1);
        // R = [[c, s], [-s, c]] with angle atan2(h10 - h01, h00 + h11): cos and sin are just the normalised pair
        const float ry = h10 - h01, rx = h00 + h11;
        const float rr = ry * ry + rx * rx;
        const float ir = rr > 0.f ? rsqrtf(rr) : 0.f;
        const float cs = rr > 0.f ? rx * ir : 1.f, sn = ry * ir;
        d00 = x0 - (px0 * cs - py0 * sn); d01 = y0 - (px0 * sn + py0 * cs);
        d10 = x1 - (px1 * cs - py1 * sn); d11 = y1 - (px1 * sn + py1 * cs);
      }
      if (a0) {
        o16[do16 + 2 * k0] = __float2half_rn(nz(d00, dofs + 2 * k0));
        o16[do16 + 2 * k0 + 1] = __float2half_rn(nz(d01, dofs + 2 * k0 + 1));
      }
      if (a1) {
        o16[do16 + 2 * k1] = __float2half_rn(nz(d10, dofs + 2 * k1));
        o16[do16 + 2 * k1 + 1] = __float2half_rn(nz(d11, dofs + 2 * k1 + 1));
      }
    }
  }
  // ---- store phase: the block's rows are contiguous in feats16
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (tid == 0 && !(dbg & 2)) {
    __half* g = p.feats16 + ((int64_t)w * p.T + t0) * p.D16;
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(g), "r"(smem_addr(s_out)), "r"((uint32_t)(nf * p.D16 * 2)) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}

}  // namespace

cudaError_t launch_zscore_table(const float* mean, const float* stdv, float* scale, float* shift, int D, cudaStream_t s) {
  k_zscore_table<<<(D + 255) / 256, 256, 0, s>>>(mean, stdv, scale, shift, D);
  return cudaGetLastError();
}

cudaError_t launch_feature_fuse(const FuseParams& p, cudaStream_t s) {
  if (p.n_windows <= 0) return cudaSuccess;
  if (p.feats == nullptr && p.feats16 != nullptr) {
    // staged (bulk-copy) kernel for the tensor-core path, when the rows are 16-byte friendly and the tile fits in smem
    bool ok = true;
    int in_floats = 0;
    for (int m = 0; m < p.M; ++m) {
      if (p.kind[m] == TAG_KIND_COSINE && ((p.raw_dim[m] % 4) != 0 || (p.raw_off[m] & 1) || (p.diff_off[m] & 1))) ok = false;
      if (reinterpret_cast<uintptr_t>(p.src[m]) & 15) ok = false;
      in_floats += (p.raw_dim[m] + 3) & ~3;
    }
    constexpr int kS = kStagedFrames;
    const int smem_in = ((kS + 1) * in_floats * 4 + 127) & ~127;
    const int smem_total = smem_in + kS * p.D16 * 2 + (TAG_MAX_MODALITIES * (kS + 1) + 2) * 4 + 16;
    if (ok && smem_total <= 227 * 1024 && (reinterpret_cast<uintptr_t>(p.feats16) & 15) == 0) {
      static int configured = 0;
      if (configured < smem_total) {
        cudaError_t e = cudaFuncSetAttribute(k_feature_fuse_staged<kS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_total);
        if (e != cudaSuccess) return e;
        configured = smem_total;
      }
      const int64_t blocks = p.n_windows * ((p.T + kS - 1) / kS);
      static int dbg = -1;
      if (dbg < 0) { const char* e = getenv("TAG_K1_DEBUG"); dbg = e ? atoi(e) : 0; }   // bottleneck experiments only
      k_feature_fuse_staged<kS><<<(unsigned)blocks, 256, smem_total, s>>>(p, in_floats, smem_in, dbg);
      return cudaGetLastError();
    }
  }
  bool any_cos = false, any_small = false;
  for (int m = 0; m < p.M; ++m) { if (p.kind[m] == TAG_KIND_COSINE) any_cos = true; else any_small = true; }
  auto grid_for = [&](int kf) {
    const int64_t warps = p.n_windows * ((p.T + kf - 1) / kf);
    return (unsigned)((warps + kThreads / 32 - 1) / (kThreads / 32));
  };
  constexpr int KC = 8, KS = 2;
  if (p.feats == nullptr && p.feats16 == nullptr) return cudaErrorInvalidValue;
  if (any_cos) {
    if (p.feats != nullptr && p.feats16 != nullptr) k_feature_fuse<true, true, true, KC><<<grid_for(KC), kThreads, 0, s>>>(p);
    else if (p.feats != nullptr) k_feature_fuse<true, false, true, KC><<<grid_for(KC), kThreads, 0, s>>>(p);
    else k_feature_fuse<false, true, true, KC><<<grid_for(KC), kThreads, 0, s>>>(p);
  }
  if (any_small) {
    if (p.feats != nullptr && p.feats16 != nullptr) k_feature_fuse<true, true, false, KS><<<grid_for(KS), kThreads, 0, s>>>(p);
    else if (p.feats != nullptr) k_feature_fuse<true, false, false, KS><<<grid_for(KS), kThreads, 0, s>>>(p);
    else k_feature_fuse<false, true, false, KS><<<grid_for(KS), kThreads, 0, s>>>(p);
  }
  return cudaGetLastError();
}
