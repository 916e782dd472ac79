// Fused TemporalConvBlock (reference model.py:22-41) for the tensor-core encoder: ONE persistent cta_group::2 tcgen05 kernel per block
//
//   y1 = GELU(conv1(h))                            (dilated conv, 5 taps, 256 -> 256 channels, no bias)
//   h  = GroupNorm_1(GELU(conv2(y1) + h))          (same shape; one group over the whole (T x 256) window)
//
// The two-kernel path (gemm_tc.cu: conv1 + GELU, then conv2 + residual + GELU + GroupNorm) writes y1 to HBM (fp16) and reads it back
// as the activation operand of conv2: 410 MB of the 936 MB a block moves per 400,000 rows, and — measured with the load / store
// switches of the experiments build (profiles/r2_conv_loadskip_probe.log) — 5.6 % of conv1's time (its stores) plus 14-19 % of
// conv2's (its activation loads). A tile owns whole windows (T <= 128), so conv2 of a tile needs y1 of that tile only: here the
// conv1 epilogue writes y1 (fp16, the SAME rounding as the two-kernel path: results are bit-identical) straight into shared memory
// in the halo-tile layout that conv2's MMAs read — (t, window)-ordered rows, 128-byte swizzled 64-channel sub-tiles, zero rows
// before frame 0 and after frame T-1 (the conv padding) that are written once per kernel.
//
// Per 256-row pair tile:
//   warp 0   TMA producer: 4 x (h chunk with its time halo + 5 W1 tap tiles), then 4 x 5 W2 tap tiles, in consumption order
//   warp 1   MMA issuer (leader CTA): conv1 into TMEM columns 0-255; conv2 into columns 256-511, chunk by chunk as the epilogue
//            hands over the 64-channel sub-tiles of y1
//   warps 2-17  epilogue: (1) conv1 accumulator -> GELU -> fp16 -> y1 sub-tiles (every warp takes a 16-channel slice of EVERY
//            sub-tile, so sub-tile 0 is complete after a quarter of this phase and conv2 starts then); (2) conv2 accumulator ->
//            + residual -> GELU -> GroupNorm -> fp16 -> TMA stores, exactly as MODE 1 of gemm_tc.cu. Phase (2) of tile i runs
//            under conv1 of tile i + 1.
// Shared memory: six 128-row tiles (the 4 sub-tiles of y1, 2 activation stages) between shared zero halos of 2 dil NW rows | weight ring
// (7 / 6 / 5 / 4 stages for dilation 1 / 2 / 4 / 8 at T = 32) | barriers, GroupNorm exchange, gamma / beta. The epilogue's staging tiles alias the interior rows of the y1 sub-tiles (dead while phase (2) runs).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace {

using namespace tcx;

constexpr int BM = 128, BN = 256, BK = 64, UK = 16, KC = 4, TAPS = 5;
constexpr int EPI_WARPS = 16, THREADS = 64 + EPI_WARPS * 32, CW = 16;
constexpr int B_BYTES = (BN / 2) * BK * 2;                       // 16 KiB: this CTA's half of a weight tap tile
constexpr int MAX_B = 8, N_A = 2;
constexpr int BAR_BYTES = 512;
constexpr int GN_RED_BYTES = EPI_WARPS * 32 * 8;                 // per-lane (sum, sumsq) exchange. One buffer: between reading it for tile i and
                                                                 // writing it for tile i + 1 every warp passes the bar.sync of phase (1)
constexpr int PAR_BYTES = 2 * BN * 4;                            // gamma | beta (the reference's convs have no bias, model.py:25-29)
constexpr int STG_TILE = 32 * 64;                                // per-warp staging tile (32 rows x 64 B, SWIZZLE_64B)
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
constexpr int TMEM_COLS = 512;
// barrier slots
constexpr int B_FULLB = 0, B_EMPTYB = MAX_B, B_FULLA = 2 * MAX_B, B_EMPTYA = 2 * MAX_B + N_A, B_ACC1F = 2 * MAX_B + 2 * N_A,
              B_ACC1E = B_ACC1F + 1, B_ACC2F = B_ACC1F + 2, B_ACC2E = B_ACC1F + 3, B_YREADY = B_ACC1F + 4, N_BARS = B_YREADY + KC;
static_assert(8 * N_BARS + 8 <= BAR_BYTES, "barrier area");

struct TbParams {
  int64_t M, m_tiles;
  int T, lw, lt, nw;            // frames per window, log2(windows per tile), log2(frames per window), windows per tile
  int dil, tap_rows, halo_rows; // rows of one tap shift (dil * nw), zero rows on each side (2 * dil * nw)
  int y_sub_bytes;              // distance of the y1 sub-tiles: (halo + 128) rows when consecutive sub-tiles share a zero halo, else (2 halo + 128)
  int a_ring_off;               // first activation stage (from the start of y1); the stages are a_stage_bytes apart
  int a_dst_off, a_ct;          // where the TMA box lands inside a stage / its first frame: (halo rows, frame 0) in the compact layout — the
                                // zero halos are never written — or (0, -2 dil) with TMA's out-of-bounds zero fill supplying them
  int tiles_bytes;              // y1 + activation stages (zeroed once at kernel start)
  int a_stage_bytes, a_box_bytes, stg_off, b_stages;
  const float* gn_gamma; const float* gn_beta;
  const __half* res16; int ldr;
};

__global__ void __launch_bounds__(THREADS, 1)
k_tcn_block(const __grid_constant__ CUtensorMap map_h, const __grid_constant__ CUtensorMap map_w1,
            const __grid_constant__ CUtensorMap map_w2, const __grid_constant__ CUtensorMap map_out, const TbParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t y1 = smem_base;
  const uint32_t a_ring = y1 + (uint32_t)p.a_ring_off;
  const uint32_t b_ring = y1 + (uint32_t)p.tiles_bytes;
  const uint32_t bar_base = b_ring + (uint32_t)(p.b_stages * B_BYTES);
  auto bar = [&](int i) { return bar_base + 8u * (uint32_t)i; };
  const uint32_t tmem_slot = bar_base + 8u * N_BARS;
  const uint32_t red_base = bar_base + BAR_BYTES;
  float* s_par = reinterpret_cast<float*>(smem_raw + (bar_base - smem_u32(smem_raw)) + BAR_BYTES + GN_RED_BYTES);   // gamma | beta
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < BN; i += THREADS) {
    s_par[i] = __ldg(p.gn_gamma + i); s_par[BN + i] = __ldg(p.gn_beta + i);
  }
  // y1: everything zero once; the epilogue only ever writes the 128 interior rows of a sub-tile (and its staging tiles live there)
  for (int i = threadIdx.x; i < p.tiles_bytes / 16; i += THREADS) sts128(y1 + (uint32_t)i * 16u, make_uint4(0u, 0u, 0u, 0u));
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    for (int s = 0; s < MAX_B; ++s) { mbar_init(bar(B_FULLB + s), 1); mbar_init(bar(B_EMPTYB + s), 1); }
    for (int s = 0; s < N_A; ++s) { mbar_init(bar(B_FULLA + s), 1); mbar_init(bar(B_EMPTYA + s), 1); }
    mbar_init(bar(B_ACC1F), 1); mbar_init(bar(B_ACC2F), 1);
    mbar_init(bar(B_ACC1E), EPI_WARPS * 2); mbar_init(bar(B_ACC2E), EPI_WARPS * 2);
    for (int c = 0; c < KC; ++c) mbar_init(bar(B_YREADY + c), EPI_WARPS * 2);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int64_t total_tiles = (p.m_tiles + 1) / 2;
  const int64_t tile0 = (int64_t)(blockIdx.x >> 1), tile_step = (int64_t)(gridDim.x >> 1);

  if (warp == 0) {
    if (lane == 0) {
      // ================= TMA producer =================
      int sb = 0; uint32_t pb = 0;
      int sa = 0; uint32_t pa = 0;
      auto load_w = [&](const CUtensorMap* map, int kc, int j) {
        mbar_wait(bar(B_EMPTYB + sb), pb ^ 1u);
        if (leader) mbar_arrive_expect_tx(bar(B_FULLB + sb), 2u * B_BYTES);
        tma_load_2d_pair(b_ring + (uint32_t)(sb * B_BYTES), map, bar(B_FULLB + sb), (j * KC + kc) * BK, (int)rank * (BN / 2));
        if (++sb == p.b_stages) { sb = 0; pb ^= 1u; }
      };
      auto load_h = [&](int64_t tile, int kc) {               // 64-channel chunk kc of the tile's rows, with its time halo
        const int64_t m_tile = tile * 2 + rank;
        mbar_wait(bar(B_EMPTYA + sa), pa ^ 1u);
        if (leader) mbar_arrive_expect_tx(bar(B_FULLA + sa), 2u * (uint32_t)p.a_box_bytes);
        tma_load_3d_pair(a_ring + (uint32_t)(sa * p.a_stage_bytes + p.a_dst_off), &map_h, bar(B_FULLA + sa), kc * BK, (int)(m_tile * p.nw), p.a_ct);
        if (++sa == N_A) { sa = 0; pa ^= 1u; }
      };
      bool prefetched = false;                                // chunks 0 and 1 of `tile` were issued before the previous tile's W2 tiles
      for (int64_t tile = tile0; tile < total_tiles; tile += tile_step) {
        for (int kc = 0; kc < KC; ++kc) {                     // conv1: h chunk, then its five W1 tap tiles
          if (!(prefetched && kc < N_A)) load_h(tile, kc);
          for (int j = 0; j < TAPS; ++j) load_w(&map_w1, kc, j);
        }
        // the activation stages free up as conv1's last chunks retire, long before the W2 tiles below are consumed: fetch the next
        // tile's first chunks now so that its conv1 does not start with an exposed load
        prefetched = tile + tile_step < total_tiles;
        if (prefetched) for (int kc = 0; kc < N_A; ++kc) load_h(tile + tile_step, kc);
        for (int kc = 0; kc < KC; ++kc)                       // conv2: W2 tap tiles only (the activation operand is y1, on chip)
          for (int j = 0; j < TAPS; ++j) load_w(&map_w2, kc, j);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // ================= MMA issuer =================
      int sb = 0; uint32_t pb = 0;
      int sa = 0; uint32_t pa = 0;
      const uint32_t acc1 = tmem_base, acc2 = tmem_base + (uint32_t)BN;
      auto taps_of_chunk = [&](uint32_t a_chunk, uint32_t d_tmem, bool first_chunk) {
        for (int j = 0; j < TAPS; ++j) {
          mbar_wait(bar(B_FULLB + sb), pb);
          tc_fence_after();
          const uint64_t adesc = make_smem_desc(a_chunk + (uint32_t)(j * p.tap_rows) * 128u);     // tap j = the same tile, j*dil frames later
          const uint64_t bdesc = make_smem_desc(b_ring + (uint32_t)(sb * B_BYTES));
#pragma unroll
          for (int k = 0; k < BK / UK; ++k)
            umma_f16_pair(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), IDESC, (!first_chunk || (j | k) != 0) ? 1u : 0u);
          umma_commit_pair(bar(B_EMPTYB + sb));
          if (++sb == p.b_stages) { sb = 0; pb ^= 1u; }
        }
      };
      int64_t it = 0;
      for (int64_t tile = tile0; tile < total_tiles; tile += tile_step, ++it) {
        const uint32_t par = (uint32_t)(it & 1);
        // ---- conv1 -> acc1 (drained by phase (1) of the previous tile's epilogue long ago)
        mbar_wait(bar(B_ACC1E), par ^ 1u);
        tc_fence_after();
        for (int kc = 0; kc < KC; ++kc) {
          mbar_wait(bar(B_FULLA + sa), pa);
          tc_fence_after();
          taps_of_chunk(a_ring + (uint32_t)(sa * p.a_stage_bytes), acc1, kc == 0);
          umma_commit_pair(bar(B_EMPTYA + sa));
          if (++sa == N_A) { sa = 0; pa ^= 1u; }
        }
        umma_commit_pair(bar(B_ACC1F));
        // ---- conv2 -> acc2, one 64-channel sub-tile of y1 at a time
        mbar_wait(bar(B_ACC2E), par ^ 1u);                    // phase (2) of the previous tile has drained acc2
        tc_fence_after();
        for (int kc = 0; kc < KC; ++kc) {
          mbar_wait(bar(B_YREADY + kc), par);
          tc_fence_after();
          taps_of_chunk(y1 + (uint32_t)(kc * p.y_sub_bytes), acc2, kc == 0);
        }
        umma_commit_pair(bar(B_ACC2F));
      }
    }
  } else {
    // ================= epilogue warps =================
    const int q = warp & 3;                      // TMEM lane quarter
    const int part = (warp - 2) >> 2;            // phase (1): 16-channel slice of every sub-tile; phase (2): 64 output columns
    const int w16 = warp - 2;
    const uint32_t stg = y1 + (uint32_t)((w16 >> 2) * p.y_sub_bytes) + (uint32_t)p.stg_off + (uint32_t)((w16 & 3) * STG_TILE);
    const int row_abs = p.halo_rows + q * 32 + lane;           // this lane's row of a y1 sub-tile
    const uint32_t y_row = y1 + (uint32_t)row_abs * 128u;
    int64_t it = 0;
    for (int64_t tile = tile0; tile < total_tiles; tile += tile_step, ++it) {
      const int64_t m_tile = tile * 2 + rank;
      const uint32_t par = (uint32_t)(it & 1);
      const RowMap rm{m_tile * BM, q * 32, p.lw, p.lt};
      // ---------------- phase (1): y1 = GELU(acc1) -> fp16 -> shared memory
      mbar_wait(bar(B_ACC1F), par);
      tc_fence_after();
      // the staging tiles of the previous tile's phase (2) live inside y1: every warp's TMA stores must have read theirs
      if (lane == 0) bulk_wait_read<0>();
      asm volatile("bar.sync %0, %1;" ::"r"(2), "r"(EPI_WARPS * 32) : "memory");
#pragma unroll
      for (int kc = 0; kc < KC; ++kc) {
        uint32_t raw[CW];
        tmem_ld16_issue(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(kc * BK + part * CW), raw);
        tmem_ld16_wait(raw);
        uint4 o[2];
        __half2* hh = reinterpret_cast<__half2*>(o);
#pragma unroll
        for (int i = 0; i < CW / 2; ++i) {
          float v0, v1;
          upk2(gelu_fast2(pk2(__uint_as_float(raw[2 * i]), __uint_as_float(raw[2 * i + 1]))), v0, v1);
          hh[i] = __floats2half2_rn(v0, v1);
        }
        const uint32_t dst = y_row + (uint32_t)(kc * p.y_sub_bytes);
        sts128(dst + (uint32_t)((((part * 2) ^ (row_abs & 7)) & 7) << 4), o[0]);
        sts128(dst + (uint32_t)((((part * 2 + 1) ^ (row_abs & 7)) & 7) << 4), o[1]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to tcgen05.mma
        if (kc == KC - 1) tc_fence_before();                             // last TMEM read of acc1
        __syncwarp();
        if (lane == 0) {
          mbar_arrive_leader(bar(B_YREADY + kc));
          if (kc == KC - 1) mbar_arrive_leader(bar(B_ACC1E));
        }
      }
      // ---------------- phase (2): GroupNorm(GELU(acc2 + h)) -> fp16 -> TMA stores (as MODE 1 of gemm_tc.cu, halo tiles)
      {
        const int n_base = part * 64;
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(BN + n_base);
        uint32_t stash[32];
        f32x2 s1p = pk2(0.f), s2p = pk2(0.f);
        uint4 rres[4];
        unit_load(reinterpret_cast<const char*>(p.res16), (int64_t)p.ldr * 2, rm, p.M, (int64_t)n_base * 2, lane, rres);
        mbar_wait(bar(B_ACC2F), par);                          // conv2 done: acc2 complete, y1 (and the staging tiles in it) dead
        tc_fence_after();
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          unit_to_smem(stg, lane, rres);
          __syncwarp();
          if (u == 0) unit_load(reinterpret_cast<const char*>(p.res16), (int64_t)p.ldr * 2, rm, p.M, (int64_t)(n_base + 32) * 2, lane, rres);
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const int c = u * 2 + cc;
            uint32_t raw[CW];
            tmem_ld16_issue(t_row + (uint32_t)(c * CW), raw);
            tmem_ld16_wait(raw);
            f32x2 w[CW / 2];
#pragma unroll
            for (int i = 0; i < CW / 2; ++i) w[i] = pk2(__uint_as_float(raw[2 * i]), __uint_as_float(raw[2 * i + 1]));
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const uint4 r = lds128(stg_addr(stg, lane, cc * 2 + i));
              const __half2* hr = reinterpret_cast<const __half2*>(&r);
#pragma unroll
              for (int e = 0; e < 4; ++e) { const float2 f = __half22float2(hr[e]); w[i * 4 + e] = add2(w[i * 4 + e], pk2(f.x, f.y)); }
            }
#pragma unroll
            for (int i = 0; i < CW / 2; ++i) {
              w[i] = gelu_fast2(w[i]);
              s1p = add2(s1p, w[i]);
              s2p = fma2(w[i], w[i], s2p);
              float v0, v1;
              upk2(w[i], v0, v1);
              const __half2 h2 = __floats2half2_rn(v0, v1);
              stash[c * (CW / 2) + i] = *reinterpret_cast<const uint32_t*>(&h2);
            }
          }
          __syncwarp();                                         // everyone has read its residual rows of this unit
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(bar(B_ACC2E));
        float s1, s2;
        { float a, b; upk2(s1p, a, b); s1 = a + b; upk2(s2p, a, b); s2 = a + b; }
        // window statistics: lane l of EVERY epilogue warp holds rows of window l & (NW-1)
        const uint32_t red = red_base;
        for (int o = 16; o >= p.nw; o >>= 1) {
          s1 += __shfl_xor_sync(FULL_MASK, s1, o);
          s2 += __shfl_xor_sync(FULL_MASK, s2, o);
        }
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(red + (uint32_t)((w16 * 32 + lane) * 8)), "f"(s1), "f"(s2) : "memory");
        asm volatile("bar.sync %0, %1;" ::"r"(1), "r"(EPI_WARPS * 32) : "memory");
        float S1 = 0.f, S2 = 0.f;
#pragma unroll
        for (int e = 0; e < EPI_WARPS; ++e) {
          float a, b;
          asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a), "=f"(b) : "r"(red + (uint32_t)((e * 32 + lane) * 8)) : "memory");
          S1 += a; S2 += b;
        }
        const float inv_n = 1.0f / ((float)p.T * (float)BN);
        const float mean = S1 * inv_n;
        const float var = fmaxf(S2 * inv_n - mean * mean, 0.f);
        const float rstd = 1.0f / sqrtf(var + 1e-5f);
        const float nmr = -mean * rstd;
        const f32x2 rstd2 = pk2(rstd), nmr2 = pk2(nmr);
        // output coordinates of this warp's 32 rows: (column, window, frame) box of the 3-D tensor map
        int ow = (int)(m_tile * p.nw), ot = (q * 32) >> p.lw;
        if (p.nw >= 32) ow += (q * 32) & (p.nw - 1);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int col = n_base + u * 32 + i * 8;
            uint4 o;
            __half2* hh = reinterpret_cast<__half2*>(&o);
            const float4 g0 = *reinterpret_cast<const float4*>(s_par + col);
            const float4 g1 = *reinterpret_cast<const float4*>(s_par + col + 4);
            const float4 b0 = *reinterpret_cast<const float4*>(s_par + BN + col);
            const float4 b1 = *reinterpret_cast<const float4*>(s_par + BN + col + 4);
            const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 z = __half22float2(*reinterpret_cast<const __half2*>(&stash[u * 16 + i * 4 + e]));
              const f32x2 o2 = fma2(fma2(pk2(z.x, z.y), rstd2, nmr2), pk2(gg[2 * e], gg[2 * e + 1]), pk2(bb[2 * e], bb[2 * e + 1]));
              float oa, ob;
              upk2(o2, oa, ob);
              hh[e] = __floats2half2_rn(oa, ob);
            }
            if (i == 0 && u == 1) { if (lane == 0) bulk_wait_read<0>(); __syncwarp(); }   // unit 0's store has read the tile
            sts128(stg_addr(stg, lane, i), o);
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                         ::"l"(reinterpret_cast<uint64_t>(&map_out)), "r"(stg), "r"(n_base + u * 32), "r"(ow), "r"(ot) : "memory");
            bulk_commit();
          }
        }
      }
    }
    if (lane == 0) bulk_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct Plan { int nw, lw, lt, halo_rows, y_sub, a_ring_off, a_stage, a_dst_off, a_ct, a_box_rows, tiles_bytes, stg_off, b_stages, smem; bool ok; };

Plan make_plan(int64_t M, int T, int dil) {
  Plan pl{};
  pl.ok = false;
  if (T < 1 || T > BM || (T & (T - 1)) != 0 || dil < 1 || M <= BM || M % T != 0) return pl;
  pl.nw = BM / T;
  if (pl.nw > 32) return pl;
  while ((1 << pl.lw) < pl.nw) ++pl.lw;
  while ((1 << pl.lt) < T) ++pl.lt;
  pl.halo_rows = 2 * dil * pl.nw;
  if (T + 4 * dil > 256) return pl;                              // TMA box limit of the activation chunk
  if (pl.halo_rows % 8 == 0) {
    // compact layout: six 128-row tiles (four y1 sub-tiles, two activation stages) separated by ONE zero halo each — the rows behind
    // a tile double as the rows in front of the next; every tile stays on a 1024-byte boundary (the 128-byte swizzle pattern is a
    // function of the address). The activation box covers frames 0 .. T-1 only and lands behind the stage's leading halo.
    pl.y_sub = (pl.halo_rows + BM) * 128;
    pl.a_stage = pl.y_sub;
    pl.a_ring_off = KC * pl.y_sub;
    pl.a_dst_off = pl.halo_rows * 128; pl.a_ct = 0; pl.a_box_rows = BM;
    pl.tiles_bytes = (KC + N_A) * pl.y_sub + pl.halo_rows * 128;
  } else {
    // halo rows not a multiple of 8 (T = 128, dilation 1 .. 3): every tile carries both of its halos; the activation box starts at
    // frame -2 dil and TMA's out-of-bounds zero fill writes the halo rows
    pl.y_sub = ((BM + 2 * pl.halo_rows) * 128 + 1023) & ~1023;
    pl.a_stage = pl.y_sub;
    pl.a_ring_off = KC * pl.y_sub;
    pl.a_dst_off = 0; pl.a_ct = -2 * dil; pl.a_box_rows = BM + 2 * pl.halo_rows;
    pl.tiles_bytes = (KC + N_A) * pl.y_sub;
  }
  pl.tiles_bytes = (pl.tiles_bytes + 1023) & ~1023;
  pl.stg_off = (pl.halo_rows * 128 + 511) & ~511;                // staging tiles start inside the interior rows, 512-byte aligned
  if (pl.stg_off + 4 * STG_TILE > (pl.halo_rows + BM) * 128) return pl;
  const int fixed = 1024 + BAR_BYTES + GN_RED_BYTES + PAR_BYTES;
  const int left = 232448 - fixed - pl.tiles_bytes;
  pl.b_stages = left / B_BYTES;
  if (pl.b_stages > MAX_B) pl.b_stages = MAX_B;
  if (pl.b_stages < 4) return pl;                                // with 3 weight stages the kernel was 7 % SLOWER than the two GEMM launches (profiles/r2_tcn_block_micro.log)
  pl.smem = fixed + pl.tiles_bytes + pl.b_stages * B_BYTES;
  pl.ok = true;
  return pl;
}

}  // namespace

bool tcn_block_supported(int64_t M, int T, int dil) { return make_plan(M, T, dil).ok; }

// host-only view of the shared-memory plan (tag_debug_tcn_block_plan: lets the CPU test suite pin the budget arithmetic)
bool tcn_block_plan(int64_t M, int T, int dil, int* weight_stages, int* smem_bytes, int* tile_bytes) {
  const Plan pl = make_plan(M, T, dil);
  if (weight_stages) *weight_stages = pl.ok ? pl.b_stages : 0;
  if (smem_bytes) *smem_bytes = pl.ok ? pl.smem : 0;
  if (tile_bytes) *tile_bytes = pl.ok ? pl.tiles_bytes : 0;
  return pl.ok;
}

cudaError_t launch_tcn_block(void* encode_fn, int num_sms, const TcnBlock& t, cudaStream_t s, char* err, int errlen) {
  if (t.M <= 0) return cudaSuccess;
  auto bad = [&](const char* msg) {
    snprintf(err, errlen, "tcn_block: %s (M=%lld T=%d dil=%d)", msg, (long long)t.M, t.T, t.dil);
    return cudaErrorInvalidValue;
  };
  const Plan pl = make_plan(t.M, t.T, t.dil);
  if (!pl.ok) return bad("unsupported shape (T a power of two <= 128, more than 128 rows, halo tile + weight ring within shared memory)");
  if (!t.h16 || !t.W1_16 || !t.W2_16 || !t.gn_gamma || !t.gn_beta) return bad("NULL argument");
  if ((reinterpret_cast<uintptr_t>(t.h16) | reinterpret_cast<uintptr_t>(t.W1_16) | reinterpret_cast<uintptr_t>(t.W2_16)) & 15)
    return bad("pointers must be 16-byte aligned");
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(k_tcn_block, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) { snprintf(err, errlen, "cudaFuncSetAttribute(k_tcn_block) failed: %s", cudaGetErrorString(e)); return e; }
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  EncodeTiledFn encode = reinterpret_cast<EncodeTiledFn>(encode_fn);
  CUtensorMap m_h, m_w1, m_w2, m_out;
  cuuint32_t es3[3] = {1, 1, 1}, es2[2] = {1, 1};
  const int64_t W = t.M / t.T;
  {
    // activations as (channel, window, frame): the box (64, NW, T + 4 dil) starts at frame -2 dil; out-of-bounds frames are zero fill
    cuuint64_t gdim[3] = {(cuuint64_t)BN, (cuuint64_t)W, (cuuint64_t)t.T};
    cuuint64_t gstr[2] = {(cuuint64_t)t.T * BN * 2, (cuuint64_t)BN * 2};
    cuuint32_t box[3] = {BK, (cuuint32_t)pl.nw, (cuuint32_t)(pl.a_box_rows / pl.nw)};
    CUresult r = encode(&m_h, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, t.h16, gdim, gstr, box, es3, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { snprintf(err, errlen, "cuTensorMapEncodeTiled(h) failed with CUresult %d", (int)r); return cudaErrorInvalidValue; }
    // output: this warp's 32 (frame, window)-ordered rows x 32 columns = one SWIZZLE_64B box
    const int bw = pl.nw < 32 ? pl.nw : 32;
    cuuint32_t obox[3] = {32, (cuuint32_t)bw, (cuuint32_t)(32 / bw)};
    r = encode(&m_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, t.h16, gdim, gstr, obox, es3, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { snprintf(err, errlen, "cuTensorMapEncodeTiled(out) failed with CUresult %d", (int)r); return cudaErrorInvalidValue; }
  }
  auto wmap = [&](CUtensorMap* m, const __half* Wp) {
    cuuint64_t gdim[2] = {(cuuint64_t)(TAPS * BN), (cuuint64_t)BN}, gstr[1] = {(cuuint64_t)(TAPS * BN) * 2};
    cuuint32_t box[2] = {BK, BN / 2};
    return encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(Wp), gdim, gstr, box, es2, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  CUresult r = wmap(&m_w1, t.W1_16);
  if (r == CUDA_SUCCESS) r = wmap(&m_w2, t.W2_16);
  if (r != CUDA_SUCCESS) { snprintf(err, errlen, "cuTensorMapEncodeTiled(weights) failed with CUresult %d", (int)r); return cudaErrorInvalidValue; }

  TbParams p{};
  p.M = t.M; p.m_tiles = (t.M + BM - 1) / BM;
  p.T = t.T; p.lw = pl.nw > 1 ? pl.lw : 0; p.lt = pl.lt; p.nw = pl.nw;
  p.dil = t.dil; p.tap_rows = t.dil * pl.nw; p.halo_rows = pl.halo_rows;
  p.y_sub_bytes = pl.y_sub; p.a_ring_off = pl.a_ring_off; p.a_stage_bytes = pl.a_stage; p.a_dst_off = pl.a_dst_off; p.a_ct = pl.a_ct;
  p.a_box_bytes = pl.a_box_rows * 128; p.tiles_bytes = pl.tiles_bytes; p.stg_off = pl.stg_off; p.b_stages = pl.b_stages;
  p.gn_gamma = t.gn_gamma; p.gn_beta = t.gn_beta; p.res16 = t.h16; p.ldr = BN;
  const int64_t total = (p.m_tiles + 1) / 2;
  const int64_t clusters = total < num_sms / 2 ? total : num_sms / 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(2 * clusters));
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = (size_t)pl.smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, k_tcn_block, m_h, m_w1, m_w2, m_out, p);
}
