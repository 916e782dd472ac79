// K1 feature fuse — one pass from packed per-frame SMPL / keypoint / appearance arrays to the
// encoder input [raw blocks || diff blocks], z-scored.
//
// Replaces (reference, /root/reference):
//   utils.py:366-381  WindowDataset._slice_or_pad     (frame gather with nearest-repeat padding)
//   utils.py:396-404  raw flatten
//   utils.py:142-147  _vit_delta        (cosine: L2-normalise, first difference, row 0 = 0)
//   utils.py:165-174  _rotmat_delta  +  :130-140 _log_so3
//   utils.py:161-163  _betas_delta
//   utils.py:177-217  _procrustes_kp_delta (closed form of `Vh @ U.T` + det fix-up in both regimes: transposed polar
//                     rotation for det(H) >= 0, polar-reflection angle for det(H) < 0; the latter frames are also
//                     counted in flags[0] — SURVEY.md §8a A6)
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
    if (mean == nullptr) { scale[i] = 1.f; shift[i] = 0.f; return; }     // stats=None path: x*1+0 == x exactly
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
            const bool mirror = h00 * h11 - h01 * h10 < 0.f;
            if (mirror && lane == 0 && p.flags != nullptr) atomicAdd(p.flags, 1);
            // R = [[c, s], [-s, c]]: angle atan2(h10 - h01, h00 + h11) for det(H) >= 0 (transposed polar rotation), and
            // atan2(h10 + h01, h00 - h11) for det(H) < 0 (angle of the polar REFLECTION of H: what `Vh @ U.T` plus the
            // det fix-up of utils.py:209-212 gives with LAPACK's always-improper U); cos and sin are the normalised pair
            const float ry = mirror ? h10 + h01 : h10 - h01, rx = mirror ? h00 - h11 : h00 + h11;
            const float rr = ry * ry + rx * rx;
            const float ir = rr > 0.f ? rsqrtf(rr) : 0.f;
            const float cs = rr > 0.f ? rx * ir : 1.f, sn = ry * ir;
            // X @ R with R = [[c, s], [-s, c]]
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
// One CTA = one window x kS = 4 consecutive frames ("slots" 0..kS hold source frames t0-1 .. t0+kS-1):
//   load    every modality's slots land in shared memory first. Modalities whose rows are 16-byte friendly (vit, clip,
//           dino, kp2d) take one cp.async.bulk per slot, completing on an mbarrier; the others (9 / 207 / 10 floats
//           per row) are copied as ONE flat coalesced range when the slots are consecutive source frames (always,
//           except in nearest-repeat padded windows), compactly at a pitch of `dim` floats;
//   phase A warps 0..3 align the keypoints of frame `warp` (shuffle reductions); warps 4..7 take the cosine row norms,
//           the rotation-log joints (one (frame, joint) pair per lane, all rotation modalities in one item list), the
//           z-scored raw columns and plain differences of the small modalities (32-column chunks from a host-built
//           chunk list) — all independent, so no warp waits for another;
//   phase B all 256 threads stream the cosine modalities: 4 columns per thread, the four z-score table float4s loaded
//           once, the previous frame's normalised values carried in registers across the kS frames;
//   store   the kS output rows of a block are contiguous in feats16: ONE cp.async.bulk shared->global per CTA.
// ~51 KB of shared memory -> 4 CTAs (32 warps) per SM, so one CTA's load phase overlaps the others' compute / store.
// The first staged version spent 3,445 warp instructions per frame (ncu: issue slots 65 % busy, DRAM 15 %): 64-bit
// index divisions, eight predicated load/store slots per small row whatever its width, per-element table loads and
// 8-warp loops over 9-column modalities. The host-built plan below removes the per-element bookkeeping.
constexpr int kStagedFrames = 4;
constexpr int kMaxRawChunks = 48, kMaxPlainChunks = 16, kMaxRotItems = 64;

struct StagedPlan {
  int bpw;                                   // blocks (of kS frames) per window
  int wide_pitch;                            // floats between slots of the bulk-copied modalities
  int out_off, inv_off, bar_off;             // byte offsets in dynamic shared memory (staging area at 0)
  int base_off[TAG_MAX_MODALITIES];          // float offset of slot 0 of modality m
  int pitch[TAG_MAX_MODALITIES];             // floats between its slots (wide_pitch, or dim when compact)
  unsigned char bulk[TAG_MAX_MODALITIES];    // 1: one cp.async.bulk per slot
  int n_raw, n_plain, n_rot, n_proc, n_cos;
  unsigned char raw_mod[kMaxRawChunks];   short raw_col[kMaxRawChunks];      // z-scored raw copies, 64 columns per chunk
  unsigned char plain_mod[kMaxPlainChunks]; short plain_col[kMaxPlainChunks];
  unsigned char rot_mod[kMaxRotItems];    unsigned char rot_joint[kMaxRotItems];
  unsigned char proc_mod[TAG_MAX_MODALITIES];
  unsigned char cos_mod[TAG_MAX_MODALITIES];
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int kS, bool FULL>
__global__ void __launch_bounds__(256) k_feature_fuse_staged(const FuseParams p, const StagedPlan pl, int dbg) {
  extern __shared__ __align__(128) unsigned char sm_raw[];
  float* s_in = reinterpret_cast<float*>(sm_raw);
  __half* s_out = reinterpret_cast<__half*>(sm_raw + pl.out_off);                  // [kS][D16]
  float* s_inv = reinterpret_cast<float*>(sm_raw + pl.inv_off);                    // [n_cos][kS+1]
  int* s_cofs = reinterpret_cast<int*>(s_inv + TAG_MAX_MODALITIES * (kS + 1));     // [M] float offset of slot 0 (this CTA)
  const uint32_t bar = smem_addr(sm_raw + pl.bar_off);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned w = blockIdx.x / (unsigned)pl.bpw;
  const int t0 = (int)(blockIdx.x - w * (unsigned)pl.bpw) * kS;
  const int nf = FULL ? kS : min(kS, p.T - t0);      // FULL: T is a multiple of kS
  const int vid = p.win_video[w];
  const int start = p.win_start[w];
  const int64_t n_rows = *p.total_frames;
  const int64_t f0 = p.frame_offset[vid];
  const int L = (int)(p.frame_offset[vid + 1] - f0);
  const int k0 = t0 == 0 ? 1 : 0;                    // slot 0 (= the frame before the window) duplicates slot 1
  // slots k0..nf are consecutive source frames unless the window is padded (utils.py:371-381); `flat`: the compact
  // modalities of this block can be fetched as ONE 16-byte-aligned bulk range each (a few floats of the neighbouring
  // rows ride along on both sides, so the very last rows of the arrays take the scalar path instead)
  const bool consec = start >= 0 && start + t0 + nf - 1 <= L - 1;
  const bool flat = consec && f0 + start + t0 + nf + 4 <= n_rows;
  auto row_of_slot = [&](int k) -> int64_t { const int t = t0 - 1 + k; return f0 + src_frame(start, t < 0 ? 0 : t, L); };

  // ---- load phase: warp 0 alone owns the mbarrier (init, expect_tx, wait), plans (lane m = modality m) and issues every
  // bulk copy; the other warps go straight to zeroing the output tile, which overlaps the copies' flight time
  if (warp == 0) {
    if (lane == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t my_tx = 0, cnt = 0;
    int64_t gi = 0;
    bool my_flat = false;
    if (lane < p.M) {
      const int dim = p.raw_dim[lane];
      if (pl.bulk[lane]) {
        my_tx = (uint32_t)((nf + 1) * dim * 4);
        s_cofs[lane] = pl.base_off[lane];
      } else {
        gi = (f0 + start + t0 - 1 + k0) * (int64_t)dim;          // first float of slot k0
        const int sh = (int)(gi & 3);
        cnt = (uint32_t)((sh + (nf + 1 - k0) * dim + 3) & ~3);
        my_flat = flat;
        if (my_flat) { my_tx = cnt * 4u; s_cofs[lane] = pl.base_off[lane] + sh - k0 * dim; gi -= sh; }
        else s_cofs[lane] = pl.base_off[lane];
      }
    }
    if (dbg & 4) { my_tx = 0; my_flat = false; }
    const uint32_t tx = __reduce_add_sync(FULL_MASK, my_tx);
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(tx) : "memory");
    __syncwarp();
    if (my_flat) {
      const uint32_t d = smem_addr(s_in + pl.base_off[lane]);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(d), "l"(p.src[lane] + gi), "r"(cnt * 4u), "r"(bar) : "memory");
    }
    if (lane <= nf && !(dbg & 4)) {                   // lane k issues the per-row bulk copies of slot k
      const int64_t r = row_of_slot(lane);
      for (int m = 0; m < p.M; ++m) {
        if (!pl.bulk[m]) continue;
        const int dim = p.raw_dim[m];
        const float* g = p.src[m] + r * dim;
        const uint32_t d = smem_addr(s_in + pl.base_off[m] + lane * pl.wide_pitch);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(d), "l"(g), "r"((uint32_t)(dim * 4)), "r"(bar) : "memory");
      }
    }
  }
  {  // zero the output tile (pad columns stay zero); ordered before the phase A/B writes by the barrier below
    uint4* z = reinterpret_cast<uint4*>(s_out);
    const int n16 = kS * p.D16 / 8;
#pragma unroll 2
    for (int i = tid; i < n16; i += 256) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (!flat && !(dbg & 4)) {                          // padded windows / last rows of the arrays: slot by slot, scalar
#pragma unroll 1
    for (int m = 0; m < p.M; ++m) {
      if (pl.bulk[m]) continue;
      const int dim = p.raw_dim[m];
      float* dst = s_in + pl.base_off[m];
      for (int k = k0; k <= nf; ++k) {
        const float* g = p.src[m] + row_of_slot(k) * dim;
        for (int i = tid; i < dim; i += 256) dst[k * dim + i] = __ldg(g + i);
      }
    }
  }
  if (warp == 0) {  // one warp waits for the bulk copies (bounded: a protocol bug must not hang the GPU); the others park at the barrier
    uint32_t ok = 0;
    const long long c0 = clock64();
    while (!ok) {
#ifndef TAG_MBAR_NO_HINT
      // suspend-time hint: the warp is parked until the copies land instead of polling (see tc_common.cuh)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(bar), "r"(0), "r"(200000u) : "memory");
#else
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(bar), "r"(0) : "memory");
#endif
      if (!ok && clock64() - c0 > 4000000000LL) __trap();
    }
  }
  __syncthreads();

  const Norm nz{p.mean, p.stdv};
  if (!(dbg & 1)) {
    // ================= phase A: independent per-warp jobs =================
    const bool has_proc = pl.n_proc > 0;
    if (has_proc && warp < 4) {
      // ---- keypoint Procrustes delta: warp f aligns frame t0+f-1 -> t0+f (utils.py:177-217)
      const int f = warp;
      if (FULL || f < nf) {
        for (int q = 0; q < pl.n_proc; ++q) {
          const int m = pl.proc_mod[q];
          const int dim = p.raw_dim[m];
          const float* xc = s_in + s_cofs[m] + (f + 1) * pl.pitch[m];
          const float* xp = s_in + s_cofs[m] + ((f == 0 && k0) ? 1 : f) * pl.pitch[m];
          __half* o16 = s_out + f * p.D16 + p.diff_off16[m];
          const int dofs = p.diff_off[m];
          const int K = dim / 2;
          const int k1 = lane + 32;
          const bool a0 = lane < K, a1 = k1 < K;
          const float invK = 1.0f / (float)K;
          // centre + Frobenius-normalise one frame's points (utils.py:192-196)
          auto load_norm = [&](const float* x, float& x0, float& y0, float& x1, float& y1) {
            const float2 u0 = a0 ? *reinterpret_cast<const float2*>(x + 2 * lane) : make_float2(0.f, 0.f);
            const float2 u1 = a1 ? *reinterpret_cast<const float2*>(x + 2 * k1) : make_float2(0.f, 0.f);
            float sx = u0.x + u1.x, sy = u0.y + u1.y;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { sx += __shfl_xor_sync(FULL_MASK, sx, o); sy += __shfl_xor_sync(FULL_MASK, sy, o); }
            const float mx = sx * invK, my = sy * invK;
            x0 = a0 ? u0.x - mx : 0.f; y0 = a0 ? u0.y - my : 0.f;
            x1 = a1 ? u1.x - mx : 0.f; y1 = a1 ? u1.y - my : 0.f;
            const float isc = 1.0f / fmaxf(sqrtf(warp_sum(x0 * x0 + y0 * y0 + x1 * x1 + y1 * y1)), 1e-6f);
            x0 *= isc; y0 *= isc; x1 *= isc; y1 *= isc;
          };
          float x0, y0, x1, y1, px0, py0, px1, py1;
          load_norm(xc, x0, y0, x1, y1);
          load_norm(xp, px0, py0, px1, py1);
          float d00 = 0.f, d01 = 0.f, d10 = 0.f, d11 = 0.f;
          if (t0 + f > 0) {
            // H = X^T Y (utils.py:207), X = previous frame, Y = current frame
            float h00 = px0 * x0 + px1 * x1, h01 = px0 * y0 + px1 * y1, h10 = py0 * x0 + py1 * x1, h11 = py0 * y0 + py1 * y1;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              h00 += __shfl_xor_sync(FULL_MASK, h00, o); h01 += __shfl_xor_sync(FULL_MASK, h01, o);
              h10 += __shfl_xor_sync(FULL_MASK, h10, o); h11 += __shfl_xor_sync(FULL_MASK, h11, o);
            }
            const bool mirror = h00 * h11 - h01 * h10 < 0.f;
            if (mirror && lane == 0 && p.flags != nullptr) atomicAdd(p.flags, 1);
            // R = [[c, s], [-s, c]]: angle atan2(h10 - h01, h00 + h11) for det(H) >= 0 (transposed polar rotation), and
            // atan2(h10 + h01, h00 - h11) for det(H) < 0 (angle of the polar REFLECTION of H: what `Vh @ U.T` plus the
            // det fix-up of utils.py:209-212 gives with LAPACK's always-improper U); cos and sin are the normalised pair
            const float ry = mirror ? h10 + h01 : h10 - h01, rx = mirror ? h00 - h11 : h00 + h11;
            const float rr = ry * ry + rx * rx;
            const float ir = rr > 0.f ? rsqrtf(rr) : 0.f;
            const float cs = rr > 0.f ? rx * ir : 1.f, sn = ry * ir;
            d00 = x0 - (px0 * cs - py0 * sn); d01 = y0 - (px0 * sn + py0 * cs);
            d10 = x1 - (px1 * cs - py1 * sn); d11 = y1 - (px1 * sn + py1 * cs);
          }
          if (a0) *reinterpret_cast<__half2*>(o16 + 2 * lane) = __floats2half2_rn(nz(d00, dofs + 2 * lane), nz(d01, dofs + 2 * lane + 1));
          if (a1) *reinterpret_cast<__half2*>(o16 + 2 * k1) = __floats2half2_rn(nz(d10, dofs + 2 * k1), nz(d11, dofs + 2 * k1 + 1));
        }
      }
    } else {
      const int aw = has_proc ? warp - 4 : warp;      // auxiliary warp index
      const int naw = has_proc ? 4 : 8;
      // ---- cosine row norms: one (modality, slot) per job
      for (int j = aw; j < pl.n_cos * (nf + 1); j += naw) {
        const int q = j / (nf + 1), k = j - q * (nf + 1);
        const int m = pl.cos_mod[q];
        const int dim = p.raw_dim[m];
        const float4* x = reinterpret_cast<const float4*>(s_in + s_cofs[m] + k * pl.pitch[m]);
        float ss = 0.f;
        for (int i = lane; i < dim / 4; i += 32) {
          const float4 a = x[i];
          ss = fmaf(a.x, a.x, ss); ss = fmaf(a.y, a.y, ss); ss = fmaf(a.z, a.z, ss); ss = fmaf(a.w, a.w, ss);
        }
        ss = warp_sum(ss);
        if (lane == 0) s_inv[q * (kS + 1) + k] = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
      }
      // ---- rotation log-maps: one (frame, joint) pair per lane (utils.py:165-174, :130-140)
      for (int e = aw * 32 + lane; e < nf * pl.n_rot; e += naw * 32) {
        const int f = e / pl.n_rot, it = e - f * pl.n_rot;
        const int m = pl.rot_mod[it], jn = pl.rot_joint[it];
        const float* xc = s_in + s_cofs[m] + (f + 1) * pl.pitch[m] + jn * 9;
        const float* xp = s_in + s_cofs[m] + ((f == 0 && k0) ? 1 : f) * pl.pitch[m] + jn * 9;
        float R[9], Q[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) { R[k] = xc[k]; Q[k] = xp[k]; }
        // Rrel = Q^T R  (utils.py:172), entries [i][j] = sum_k Q[k][i] R[k][j]; only the trace and the skew part are used
        const float tr0 = Q[0] * R[0] + Q[3] * R[3] + Q[6] * R[6];
        const float tr1 = Q[1] * R[1] + Q[4] * R[4] + Q[7] * R[7];
        const float tr2 = Q[2] * R[2] + Q[5] * R[5] + Q[8] * R[8];
        const float e21 = Q[2] * R[1] + Q[5] * R[4] + Q[8] * R[7], e12 = Q[1] * R[2] + Q[4] * R[5] + Q[7] * R[8];
        const float e02 = Q[0] * R[2] + Q[3] * R[5] + Q[6] * R[8], e20 = Q[2] * R[0] + Q[5] * R[3] + Q[8] * R[6];
        const float e10 = Q[1] * R[0] + Q[4] * R[3] + Q[7] * R[6], e01 = Q[0] * R[1] + Q[3] * R[4] + Q[6] * R[7];
        float tr = tr0 + tr1 + tr2;
        tr = fminf(fmaxf(tr, -1.f + 1e-6f), 3.f - 1e-6f);
        const float c = (tr - 1.f) / 2.f;
        const float theta = acosf(c);
        const float den = fmaxf(2.f * sqrtf((1.f - c) * (1.f + c)), 1e-6f);
        const float kk = theta / den;
        const int dofs = p.diff_off[m] + jn * 3;
        __half* o = s_out + f * p.D16 + p.diff_off16[m] + jn * 3;
        o[0] = __float2half_rn(nz(kk * (e21 - e12), dofs));
        o[1] = __float2half_rn(nz(kk * (e02 - e20), dofs + 1));
        o[2] = __float2half_rn(nz(kk * (e10 - e01), dofs + 2));
      }
    }
    __syncthreads();                                   // row norms visible

    // ================= phase B: all warps =================
    // ---- z-scored raw columns of the non-cosine modalities: 64 columns per job (two per lane, one half2 store), all frames
    for (int j = warp; j < pl.n_raw; j += 8) {
      const int m = pl.raw_mod[j];
      const int c = pl.raw_col[j] + 2 * lane;
      const int dim = p.raw_dim[m];
      if (c < dim) {
        const bool two = c + 1 < dim;
        const int ro = p.raw_off[m] + c;
        const float sc0 = __ldg(nz.scale + ro), sh0 = __ldg(nz.shift + ro);
        const float sc1 = two ? __ldg(nz.scale + ro + 1) : 0.f, sh1 = two ? __ldg(nz.shift + ro + 1) : 0.f;
        const int pitch = pl.pitch[m];
        const float* x = s_in + s_cofs[m] + pitch + c;
        __half* o = s_out + p.raw_off16[m] + c;                 // even column of a 64-aligned block: 4-byte aligned
#pragma unroll
        for (int f = 0; f < kS; ++f) {
          if (FULL || f < nf) {
            const float x0 = x[f * pitch], x1 = two ? x[f * pitch + 1] : 0.f;   // the odd tail column is a zero pad column
            *reinterpret_cast<__half2*>(o + f * p.D16) = __floats2half2_rn(fmaf(x0, sc0, sh0), fmaf(x1, sc1, sh1));
          }
        }
      }
    }
    // ---- plain first differences (utils.py:161-163)
    for (int j = warp; j < pl.n_plain; j += 8) {
      const int m = pl.plain_mod[j];
      const int c = pl.plain_col[j] + lane;
      if (c < p.raw_dim[m]) {
        const int dofs = p.diff_off[m] + c;
        const float sc = __ldg(nz.scale + dofs), sh = __ldg(nz.shift + dofs);
        const float* x = s_in + s_cofs[m] + c;
        __half* o = s_out + p.diff_off16[m] + c;
        float prev = x[k0 * pl.pitch[m]];
#pragma unroll
        for (int f = 0; f < kS; ++f) {
          if (FULL || f < nf) {
            const float cur = x[(f + 1) * pl.pitch[m]];
            o[f * p.D16] = __float2half_rn(fmaf(cur - prev, sc, sh));
            prev = cur;
          }
        }
      }
    }
    // ---- cosine modalities, 4 columns per thread
#pragma unroll 1
    for (int q = 0; q < pl.n_cos; ++q) {
      const int m = pl.cos_mod[q];
      const int dim = p.raw_dim[m];
      const float* xin = s_in + s_cofs[m];
      const int pitch = pl.pitch[m];
      const int ro = p.raw_off[m], dofs = p.diff_off[m];
      const float* invm = s_inv + q * (kS + 1);
      float inv[kS + 1];
#pragma unroll
      for (int k = 0; k <= kS; ++k) inv[k] = invm[(FULL || k <= nf) ? k : 0];
      for (int i = 4 * tid; i < dim; i += 1024) {
        // ro, dofs even on this path (checked on the host): 8-byte table loads
        auto ld4 = [](const float* t) {
          const float2 a = __ldg(reinterpret_cast<const float2*>(t)), b = __ldg(reinterpret_cast<const float2*>(t + 2));
          return make_float4(a.x, a.y, b.x, b.y);
        };
        const float4 sr = ld4(nz.scale + ro + i), hr = ld4(nz.shift + ro + i);
        const float4 sd = ld4(nz.scale + dofs + i), hd = ld4(nz.shift + dofs + i);
        float4 prev = *reinterpret_cast<const float4*>(xin + k0 * pitch + i);
        const float inv0 = k0 ? inv[1] : inv[0];
        prev.x *= inv0; prev.y *= inv0; prev.z *= inv0; prev.w *= inv0;
        __half* oraw = s_out + p.raw_off16[m] + i;
        __half* odif = s_out + p.diff_off16[m] + i;
#pragma unroll
        for (int f = 0; f < kS; ++f) {
          if (FULL || f < nf) {
            const float4 a = *reinterpret_cast<const float4*>(xin + (f + 1) * pitch + i);
            const __half2 r0 = __floats2half2_rn(fmaf(a.x, sr.x, hr.x), fmaf(a.y, sr.y, hr.y));
            const __half2 r1 = __floats2half2_rn(fmaf(a.z, sr.z, hr.z), fmaf(a.w, sr.w, hr.w));
            *reinterpret_cast<uint2*>(oraw + f * p.D16) = make_uint2(*reinterpret_cast<const uint32_t*>(&r0), *reinterpret_cast<const uint32_t*>(&r1));
            {
              const float4 cur = make_float4(a.x * inv[f + 1], a.y * inv[f + 1], a.z * inv[f + 1], a.w * inv[f + 1]);
              const __half2 d0 = __floats2half2_rn(fmaf(cur.x - prev.x, sd.x, hd.x), fmaf(cur.y - prev.y, sd.y, hd.y));
              const __half2 d1 = __floats2half2_rn(fmaf(cur.z - prev.z, sd.z, hd.z), fmaf(cur.w - prev.w, sd.w, hd.w));
              *reinterpret_cast<uint2*>(odif + f * p.D16) = make_uint2(*reinterpret_cast<const uint32_t*>(&d0), *reinterpret_cast<const uint32_t*>(&d1));
              prev = cur;
            }
          }
        }
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
    // staged (bulk-copy) kernel for the tensor-core path, when the layout is 16-byte friendly and the tile fits in smem
    constexpr int kS = kStagedFrames;
    StagedPlan pl{};
    bool ok = p.mean != nullptr && p.stdv != nullptr && p.total_frames != nullptr && (reinterpret_cast<uintptr_t>(p.feats16) & 15) == 0 && (p.D16 % 8) == 0 && p.n_windows * ((p.T + kS - 1) / kS) < (1ll << 31);
    int wide = 0, compact = 0;
    for (int m = 0; m < p.M && ok; ++m) {
      const int dim = p.raw_dim[m];
      const bool aligned = (dim % 4) == 0 && (reinterpret_cast<uintptr_t>(p.src[m]) & 15) == 0;
      pl.bulk[m] = aligned ? 1 : 0;
      if (reinterpret_cast<uintptr_t>(p.src[m]) & 15) ok = false;          // bulk copies start at 16-byte boundaries of the array
      if ((p.raw_off16[m] & 1) || (p.diff_off16[m] & 1)) ok = false;       // half2 stores
      switch (p.kind[m]) {
        case TAG_KIND_COSINE:
          if (!aligned || p.diff_dim[m] != dim || (p.raw_off[m] & 1) || (p.diff_off[m] & 1) || (p.raw_off16[m] & 3) || (p.diff_off16[m] & 3)) ok = false;
          pl.cos_mod[pl.n_cos++] = (unsigned char)m;
          break;
        case TAG_KIND_ROTMAT:
          for (int j = 0; j < dim / 9; ++j) {
            if (pl.n_rot >= kMaxRotItems || j > 255) { ok = false; break; }
            pl.rot_mod[pl.n_rot] = (unsigned char)m; pl.rot_joint[pl.n_rot++] = (unsigned char)j;
          }
          break;
        case TAG_KIND_PLAIN:
          for (int c = 0; c < dim; c += 32) {
            if (pl.n_plain >= kMaxPlainChunks) { ok = false; break; }
            pl.plain_mod[pl.n_plain] = (unsigned char)m; pl.plain_col[pl.n_plain++] = (short)c;
          }
          break;
        default:  // TAG_KIND_PROCRUSTES: 2-D points, two per lane, half2 stores
          if (dim > 128 || (dim & 1) || (p.diff_off16[m] & 1) || !aligned) ok = false;
          pl.proc_mod[pl.n_proc++] = (unsigned char)m;
          break;
      }
      if (p.kind[m] != TAG_KIND_COSINE)
        for (int c = 0; c < dim; c += 64) {
          if (pl.n_raw >= kMaxRawChunks || dim > 32000) { ok = false; break; }
          pl.raw_mod[pl.n_raw] = (unsigned char)m; pl.raw_col[pl.n_raw++] = (short)c;
        }
      if (aligned) wide += dim; else compact += ((kS + 1) * dim + 7) & ~3;
    }
    if (ok) {
      pl.wide_pitch = wide;
      int wo = 0, co = (kS + 1) * wide;
      for (int m = 0; m < p.M; ++m) {
        if (pl.bulk[m]) { pl.base_off[m] = wo; pl.pitch[m] = wide; wo += p.raw_dim[m]; }
        else { pl.base_off[m] = co; pl.pitch[m] = p.raw_dim[m]; co += ((kS + 1) * p.raw_dim[m] + 7) & ~3; }   // + alignment slack of the flat copy
      }
      pl.bpw = (p.T + kS - 1) / kS;
      pl.out_off = (co * 4 + 127) & ~127;
      pl.inv_off = pl.out_off + kS * p.D16 * 2;
      pl.bar_off = (pl.inv_off + TAG_MAX_MODALITIES * (kS + 1) * 4 + TAG_MAX_MODALITIES * 4 + 15) & ~15;
      const int smem_total = pl.bar_off + 16;
      if (smem_total <= 227 * 1024) {
        static int configured_dev[64];                    // per device: function attributes are device state
        int dev = 0;
        cudaGetDevice(&dev);
        int& configured = configured_dev[dev & 63];
        if (configured < smem_total) {
          cudaError_t e = cudaFuncSetAttribute(k_feature_fuse_staged<kS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_total);
          if (e == cudaSuccess) e = cudaFuncSetAttribute(k_feature_fuse_staged<kS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_total);
          if (e != cudaSuccess) return e;
          configured = smem_total;
        }
        const int64_t blocks = p.n_windows * pl.bpw;
#ifdef TAG_EXPERIMENTS
        static int dbg = -1;
        if (dbg < 0) { const char* e = getenv("TAG_K1_DEBUG"); dbg = e ? atoi(e) : 0; }   // bottleneck experiments (tools/ build only)
#else
        const int dbg = 0;
#endif
        if (p.T % kS == 0) k_feature_fuse_staged<kS, true><<<(unsigned)blocks, 256, smem_total, s>>>(p, pl, dbg);
        else k_feature_fuse_staged<kS, false><<<(unsigned)blocks, 256, smem_total, s>>>(p, pl, dbg);
        return cudaGetLastError();
      }
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
