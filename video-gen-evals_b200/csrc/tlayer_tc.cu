// Fused tail of one post-norm transformer layer on sm_100a (reference model.py:143-146, nn.TransformerEncoderLayer with
// norm_first = False, ReLU, eps 1e-5):
//
//     x1 = LayerNorm1(x + attn @ Wo^T + bo)            (out-proj + residual + norm1)
//     x  = LayerNorm2(x1 + relu(x1 @ W1^T + b1) @ W2^T + b2)     (FFN + residual + norm2)
//
// as ONE persistent tcgen05 kernel per layer instead of three GEMM launches. Per 256-row tile (CTA pair, cta_group::2):
//
//   TMEM   cols   0..255  ACC_O : x1 (fp32, parked by the LayerNorm1 epilogue with tcgen05.st); FFN2 accumulates ON TOP of it
//          cols 256..511  ACC_H : the out-proj accumulator of the NEXT tile (issued while LayerNorm2 of this tile still reads
//                                 ACC_O), then the FFN1 accumulator of one 256-column hidden chunk at a time
//   SMEM   XA   64 KiB    x1 as the fp16 A operand of FFN1 (written by the LayerNorm1 epilogue in the 128B-swizzled K-major layout)
//          HB   2 x 32 KiB the two 128-column halves of relu(h) of a chunk as the fp16 A operand of FFN2
//          RING 10 x 8 KiB TMA granules: attention tile + Wo (out-proj), W1 row blocks, W2 column blocks
//
// The [rows x 1024] FFN hidden state, x1 and its fp16 copy never touch HBM: per layer pass the kernel reads attn (fp16) and x
// (fp32) and writes x (fp32) and its fp16 copy — 1.27 GB for 412,500 tokens instead of 4.1 GB for the three launches it replaces.
// Hidden chunks are software pipelined in halves: while the epilogue warps turn one 128-column half of chunk c around
// (TMEM -> +b1 -> ReLU -> fp16 -> SMEM), the tensor pipe runs FFN2 of the half before it (see the issue order in the MMA role).
//
// Roles (576 threads): warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer (leader CTA), warps 2..17 epilogue
// (4 per TMEM lane quarter x 4 column groups). Barrier protocol in the comments of each role; every wait is bounded.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace {

using namespace tcx;

constexpr int BM = 128, BK = 64, UK = 16, DM = 256;
constexpr int SUB = BM * BK * 2;                 // 16 KiB: one [128 x 64] fp16 sub-tile, SWIZZLE_128B K-major
constexpr int UNIT = SUB / 2;                    // ring granule: 8 KiB (a W1 block); attention / Wo / W2 blocks take two
constexpr int NS = 10;                           // ring granules (80 KiB). Two-granule loads always start on an even granule
                                                 // (single-granule loads come in fours), so they never straddle the wrap
constexpr int HCOLS = 128;                       // hidden columns per hand-over (half a chunk)
constexpr int WCOLS = 256;                       // hidden columns per FFN1 chunk (one N = 256 accumulator)
constexpr int EPI_WARPS = 16;
constexpr int THREADS = 64 + EPI_WARPS * 32;
constexpr int CW = 16;
constexpr int OFF_XA = 0, OFF_HB = 4 * SUB, OFF_RING = 8 * SUB, OFF_BAR = OFF_RING + NS * UNIT;
constexpr int BAR_BYTES = 512;
constexpr int RED_BYTES = EPI_WARPS * 32 * 8;              // (sum, sumsq) exchange of LN1 and LN2 (one buffer: every use is fenced
                                                           // from the next by a barrier that needs all epilogue warps)
constexpr int MAX_FFN = 1024;
constexpr int PAR_FLOATS = 6 * DM + MAX_FFN;               // bo, b2, ln1 gamma/beta, ln2 gamma/beta, b1
constexpr int SMEM_BYTES = OFF_BAR + BAR_BYTES + RED_BYTES + PAR_FLOATS * 4 + 1024;
constexpr int STG_WARP_BYTES = 2 * 32 * 64;            // two staging tiles per epilogue warp (TMA stores alternate between them)
static_assert(EPI_WARPS * STG_WARP_BYTES <= 4 * SUB, "the epilogue staging tiles alias the h buffers");
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
constexpr int TMEM_COLS = 512;
constexpr uint32_t IDESC_N256 = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);   // f16 x f16 -> f32, M = 256 (pair)

// barrier slots (8 B each)
constexpr int B_RFULL = 0, B_REMPTY = NS, B_OFULL = 2 * NS, B_X1 = 2 * NS + 1, B_HFULL = 2 * NS + 2, B_HREADY = 2 * NS + 4,
              B_HBFREE = 2 * NS + 6, B_O2FULL = 2 * NS + 8, B_COUNT = 2 * NS + 10;

struct TlParams {
  int64_t M;              // token rows
  int64_t m_tiles;
  int n_chunks;           // ffn_dim / 256
  const float* x_in;      // [M,256] fp32 residual stream (read)
  float* x_out;           // [M,256] fp32 (written; may alias x_in: a tile reads its rows before it writes them)
  __half* x16_out;        // [M,256] fp16 copy of x_out
  const float* bo; const float* b1; const float* b2;
  const float* ln1_g; const float* ln1_b; const float* ln2_g; const float* ln2_b;
};

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

#ifdef TAG_EXPERIMENTS
// timeline probe (tools/tl_trace.py): CTA 0's MMA thread / first epilogue warp / producer append (tag, clock64) pairs
__device__ long long* g_tl_trace = nullptr;
#define TL_TRACE(role, tag)                                                                       \
  do {                                                                                            \
    if (g_tl_trace != nullptr && blockIdx.x == 0 && tr_n < 2000) {                                \
      g_tl_trace[(role) * 4096 + 2 * tr_n] = (tag); g_tl_trace[(role) * 4096 + 2 * tr_n + 1] = clock64(); ++tr_n; \
    }                                                                                             \
  } while (0)
#else
#define TL_TRACE(role, tag) do {} while (0)
#endif

__global__ void __launch_bounds__(THREADS, 1)
k_tlayer_tail(const __grid_constant__ CUtensorMap map_x32, const __grid_constant__ CUtensorMap map_x16,
              const __grid_constant__ CUtensorMap map_att, const __grid_constant__ CUtensorMap map_wo,
              const __grid_constant__ CUtensorMap map_w1, const __grid_constant__ CUtensorMap map_w2, const TlParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t xa = base + OFF_XA, hb = base + OFF_HB, ring = base + OFF_RING, bars = base + OFF_BAR;
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };
  const uint32_t tmem_slot = bars + 8u * B_COUNT;
  const uint32_t red0 = bars + BAR_BYTES;
  float* s_par = reinterpret_cast<float*>(smem_raw + (bars - smem_u32(smem_raw)) + BAR_BYTES + RED_BYTES);
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < DM; i += THREADS) {
    s_par[i] = __ldg(p.bo + i); s_par[DM + i] = __ldg(p.b2 + i);
    s_par[2 * DM + i] = __ldg(p.ln1_g + i); s_par[3 * DM + i] = __ldg(p.ln1_b + i);
    s_par[4 * DM + i] = __ldg(p.ln2_g + i); s_par[5 * DM + i] = __ldg(p.ln2_b + i);
  }
  for (int i = threadIdx.x; i < p.n_chunks * WCOLS; i += THREADS) s_par[6 * DM + i] = __ldg(p.b1 + i);   // a global load per chunk sat on the
                                                                                                       // chunk turn-around path (ncu: 55 % of it)
  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(bar(B_RFULL + s), 1); mbar_init(bar(B_REMPTY + s), 1); }
    mbar_init(bar(B_OFULL), 1);
    mbar_init(bar(B_X1), EPI_WARPS * 2);
    mbar_init(bar(B_HFULL), 1);
    for (int b = 0; b < 2; ++b) { mbar_init(bar(B_HREADY + b), EPI_WARPS * 2); mbar_init(bar(B_HBFREE + b), 1); }
    mbar_init(bar(B_O2FULL), 1);
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
      // ================= TMA producer: the loads of a tile in exactly the order the MMA issuer consumes them =================
      int slot = 0; uint32_t phase = 0;
      int tr_n = 0; (void)tr_n;
      auto acquire = [&](int n_units) {                    // wait for n granules, arm the load's full barrier (leader: bytes of both CTAs)
        for (int u = 0; u < n_units; ++u) mbar_wait(bar(B_REMPTY + slot + u), phase ^ 1u);
        if (leader) mbar_arrive_expect_tx(bar(B_RFULL + slot), 2u * (uint32_t)(n_units * UNIT));
        return ring + (uint32_t)(slot * UNIT);
      };
      auto advance = [&](int n_units) { slot += n_units; if (slot >= NS) { slot = 0; phase ^= 1u; } };
      for (int64_t tile = tile0; tile < total_tiles; tile += tile_step) {
        const int m_tile = (int)(tile * 2 + rank);
        for (int kb = 0; kb < DM / BK; ++kb) {             // out-proj: attention rows, then this CTA's half of Wo
          uint32_t d = acquire(2);
          tma_load_3d_pair(d, &map_att, bar(B_RFULL + slot), kb * BK, m_tile * BM, 0);
          TL_TRACE(2, 100 + kb);
          advance(2);
          d = acquire(2);
          tma_load_2d_pair(d, &map_wo, bar(B_RFULL + slot), kb * BK, (int)rank * (DM / 2));
          advance(2);
        }
        auto load_f1 = [&](int c) {                        // W1 rows of wide hidden chunk c (256 hidden columns): 128 per CTA
          for (int kb = 0; kb < DM / BK; ++kb) {
            const uint32_t d = acquire(2);
            tma_load_2d_pair(d, &map_w1, bar(B_RFULL + slot), kb * BK, c * WCOLS + (int)rank * (WCOLS / 2));
            TL_TRACE(2, 200 + c * 4 + kb);
            advance(2);
          }
        };
        auto load_f2 = [&](int c, int half) {              // W2 columns of one 128-column half of chunk c: 128 output rows per CTA
          for (int kb = 0; kb < HCOLS / BK; ++kb) {
            const uint32_t d = acquire(2);
            tma_load_2d_pair(d, &map_w2, bar(B_RFULL + slot), c * WCOLS + half * HCOLS + kb * BK, (int)rank * (DM / 2));
            TL_TRACE(2, 300 + (c * 2 + half) * 2 + kb);
            advance(2);
          }
        };
        load_f1(0);
        for (int c = 0; c < p.n_chunks; ++c) {
          if (c > 0) load_f2(c - 1, 1);
          load_f2(c, 0);
          if (c + 1 < p.n_chunks) load_f1(c + 1);
        }
        load_f2(p.n_chunks - 1, 1);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // ================= MMA issuer (leader CTA) =================
      int slot = 0; uint32_t phase = 0;
      int tr_n = 0; (void)tr_n;
      uint32_t full_par = 0;                                // a load's full barrier is the one of its FIRST granule: a granule that was the
                                                            // second half of a two-granule load skipped a use, so every full barrier keeps
                                                            // its own parity (the empty barriers are used by every granule on every lap)
      auto wait_full = [&]() {                              // the load's TMA bytes (of both CTAs) have landed
        mbar_wait(bar(B_RFULL + slot), (full_par >> slot) & 1u);
        full_par ^= 1u << slot;
        tc_fence_after();
        return ring + (uint32_t)(slot * UNIT);
      };
      auto next = [&](int n_units) { const int cur = slot; slot += n_units; if (slot >= NS) { slot = 0; phase ^= 1u; } return cur; };
      auto release = [&](int first, int n_units) {          // granules are free (in both CTAs) once the MMAs issued so far retire
        for (int u = 0; u < n_units; ++u) umma_commit_pair(bar(B_REMPTY + first + u));
      };
      const uint32_t acc_o = tmem_base, acc_h = tmem_base + 256u;
      int64_t it = 0;
      for (int64_t tile = tile0; tile < total_tiles; tile += tile_step, ++it) {
        // ---- out-proj: ACC_H (all 256 columns) = attn @ Wo^T. ACC_H is free: every hidden chunk of the previous tile has been
        // read (h_ready waits above); ACC_O may still be read by LayerNorm2 of the previous tile — nothing touches it here
        for (int kb = 0; kb < DM / BK; ++kb) {
          const uint32_t sa = wait_full();
          const int slot_a = next(2);
          const uint32_t sb = wait_full();
          const int slot_b = next(2);
          const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sb);
#pragma unroll
          for (int k = 0; k < BK / UK; ++k)
            umma_f16_pair(acc_h, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), IDESC_N256, (kb | k) != 0 ? 1u : 0u);
          release(slot_a, 2);
          release(slot_b, 2);
        }
        umma_commit_pair(bar(B_OFULL));
        TL_TRACE(0, 1);
        // ---- x1 (fp16, SMEM) and its fp32 copy parked in ACC_O are ready; every epilogue warp has left LayerNorm2 of the
        // previous tile (ACC_O is ours) and LayerNorm1 of this one (ACC_H is free for the hidden chunks)
        mbar_wait(bar(B_X1), (uint32_t)(it & 1));
        tc_fence_after();
        TL_TRACE(0, 2);
        auto ffn1 = [&](int c) {                           // ACC_H (256 columns) = x1 @ W1[wide chunk c]^T
          for (int kb = 0; kb < DM / BK; ++kb) {
            const uint32_t sb = wait_full();
            const int sl = next(2);
            const uint64_t adesc = make_smem_desc(xa + (uint32_t)(kb * SUB)), bdesc = make_smem_desc(sb);
#pragma unroll
            for (int k = 0; k < BK / UK; ++k)
              umma_f16_pair(acc_h, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), IDESC_N256, (kb | k) != 0 ? 1u : 0u);
            release(sl, 2);
          }
          umma_commit_pair(bar(B_HFULL));
          TL_TRACE(0, 200 + c);
        };
        auto ffn2 = [&](int c, int half) {                 // ACC_O += relu(h)[half of chunk c] @ W2[:, those 128 columns]^T (on top of x1)
          for (int kb = 0; kb < HCOLS / BK; ++kb) {
            const uint32_t sb = wait_full();
            const int sl = next(2);
            const uint64_t adesc = make_smem_desc(hb + (uint32_t)((half * 2 + kb) * SUB)), bdesc = make_smem_desc(sb);
#pragma unroll
            for (int k = 0; k < BK / UK; ++k)
              umma_f16_pair(acc_o, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), IDESC_N256, 1u);
            release(sl, 2);
          }
          umma_commit_pair(bar(B_HBFREE + half));
          TL_TRACE(0, 400 + c * 2 + half);
        };
        // Wide chunks (256 hidden columns, N = 256 MMAs: the x1 operand is read from shared memory once per 256 output columns, not
        // once per 128) with ONE FFN1 accumulator, handed over in two 128-column halves. Issue order
        //   F1(c) | F2(c-1, h1) | F2(c, h0) | F1(c+1) | F2(c, h1) | ...
        // so that the tensor pipe has F2(c-1, h1) to run while the epilogue turns the first half of chunk c around, and F2(c, h0)
        // while it turns the second.
        ffn1(0);
        for (int c = 0; c < p.n_chunks; ++c) {
          const uint32_t par = (uint32_t)((it * p.n_chunks + c) & 1);
          if (c > 0) ffn2(c - 1, 1);
          mbar_wait(bar(B_HREADY + 0), par);               // relu(h) first half of chunk c is in HB[0]
          tc_fence_after();
          TL_TRACE(0, 300 + c * 2);
          ffn2(c, 0);
          mbar_wait(bar(B_HREADY + 1), par);               // second half in HB[1]; ACC_H has been read completely
          tc_fence_after();
          TL_TRACE(0, 301 + c * 2);
          if (c + 1 < p.n_chunks) ffn1(c + 1);
        }
        ffn2(p.n_chunks - 1, 1);
        umma_commit_pair(bar(B_O2FULL));
        TL_TRACE(0, 9);
      }
    }
  } else {
    // ================= epilogue warps =================
    const int q = warp & 3;                      // TMEM lane quarter
    const int part = (warp - 2) >> 2;            // column group: 64 of the 256 output columns, 32 of a 128-column hidden chunk
    constexpr int NCH = 64 / CW;
    const uint32_t stg = hb + (uint32_t)((warp - 2) * STG_WARP_BYTES);     // aliases HB: only used while no h chunk is live
    const int row = q * 32 + lane;               // this lane's row of the CTA's 128-row tile
    const uint32_t t_o = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(part * 64);
    const uint32_t t_h = tmem_base + ((uint32_t)(q * 32) << 16) + 256u + (uint32_t)(part * 32);
    const uint32_t t_p = tmem_base + ((uint32_t)(q * 32) << 16) + 256u + (uint32_t)(part * 64);      // out-proj accumulator (in ACC_H)
    const int nl = part * 64;
    const uint32_t xa_row = xa + (uint32_t)(part * SUB) + (uint32_t)(row * 128);
    int64_t it = 0;
    int tr_n = (warp == 2 && lane == 0) ? 0 : 1000000; (void)tr_n;
    for (int64_t tile = tile0; tile < total_tiles; tile += tile_step, ++it) {
      const int64_t m_tile = tile * 2 + rank;
      const RowMap rm{m_tile * BM, q * 32, 0, 0};
      // ---------------- LayerNorm1: x1 = LN(acc + bo + x), acc in ACC_H -> parked in ACC_O (fp32; this warp's own lanes and
      // columns, which it has finished reading in LayerNorm2 of the previous tile) and written to XA (fp16 A operand)
      {
        float s1 = 0.f, s2 = 0.f;
        uint4 rres[NCH][4];                                  // the whole 32 x 64 fp32 residual block of this warp, in flight before the wait
#pragma unroll
        for (int u = 0; u < NCH; ++u)
          unit_load(reinterpret_cast<const char*>(p.x_in), (int64_t)DM * 4, rm, p.M, (int64_t)(nl + u * CW) * 4, lane, rres[u]);
        TL_TRACE(1, 10);
        mbar_wait(bar(B_OFULL), (uint32_t)(it & 1));
        tc_fence_after();
        TL_TRACE(1, 1);
        if (lane == 0) bulk_wait_read<0>();                  // LayerNorm2 of the previous tile: its TMA stores have read the staging tiles
        __syncwarp();
#pragma unroll
        for (int u = 0; u < NCH; ++u) {
          unit_to_smem(stg, lane, rres[u]);
          __syncwarp();
          uint32_t raw[CW];
          tmem_ld16_issue(t_p + (uint32_t)(u * CW), raw);
          tmem_ld16_wait(raw);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint4 xr = lds128(stg_addr(stg, lane, i));
            const float4 bb = *reinterpret_cast<const float4*>(s_par + nl + u * CW + i * 4);
            const float v0 = __uint_as_float(raw[i * 4]) + bb.x + __uint_as_float(xr.x);
            const float v1 = __uint_as_float(raw[i * 4 + 1]) + bb.y + __uint_as_float(xr.y);
            const float v2 = __uint_as_float(raw[i * 4 + 2]) + bb.z + __uint_as_float(xr.z);
            const float v3 = __uint_as_float(raw[i * 4 + 3]) + bb.w + __uint_as_float(xr.w);
            s1 += (v0 + v1) + (v2 + v3);
            s2 = fmaf(v0, v0, s2); s2 = fmaf(v1, v1, s2); s2 = fmaf(v2, v2, s2); s2 = fmaf(v3, v3, s2);
            raw[i * 4] = __float_as_uint(v0); raw[i * 4 + 1] = __float_as_uint(v1);
            raw[i * 4 + 2] = __float_as_uint(v2); raw[i * 4 + 3] = __float_as_uint(v3);
          }
          tmem_st16(t_p + (uint32_t)(u * CW), raw);        // pre-norm sum parked in place (ACC_H)
          __syncwarp();
        }
        tmem_st_wait();
        TL_TRACE(1, 11);
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(red0 + (uint32_t)(((warp - 2) * 32 + lane) * 8)), "f"(s1), "f"(s2) : "memory");
        asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "r"(4 * 32) : "memory");
        float S1 = 0.f, S2 = 0.f;
#pragma unroll
        for (int pp = 0; pp < EPI_WARPS / 4; ++pp) {
          const int e = pp * 4 + ((q - 2) & 3);
          float a, b;
          asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a), "=f"(b) : "r"(red0 + (uint32_t)((e * 32 + lane) * 8)) : "memory");
          S1 += a; S2 += b;
        }
        const float mean = S1 * (1.0f / DM);
        const float var = fmaxf(S2 * (1.0f / DM) - mean * mean, 0.f);
        const float rstd = 1.0f / sqrtf(var + 1e-5f);
        const float nmr = -mean * rstd;
        TL_TRACE(1, 12);
#pragma unroll
        for (int u = 0; u < NCH; ++u) {
          uint32_t raw[CW];
          tmem_ld16_issue(t_p + (uint32_t)(u * CW), raw);
          tmem_ld16_wait(raw);
          float y[CW];
#pragma unroll
          for (int i = 0; i < CW; i += 4) {
            const float4 g = *reinterpret_cast<const float4*>(s_par + 2 * DM + nl + u * CW + i);
            const float4 b = *reinterpret_cast<const float4*>(s_par + 3 * DM + nl + u * CW + i);
            y[i] = fmaf(fmaf(__uint_as_float(raw[i]), rstd, nmr), g.x, b.x);
            y[i + 1] = fmaf(fmaf(__uint_as_float(raw[i + 1]), rstd, nmr), g.y, b.y);
            y[i + 2] = fmaf(fmaf(__uint_as_float(raw[i + 2]), rstd, nmr), g.z, b.z);
            y[i + 3] = fmaf(fmaf(__uint_as_float(raw[i + 3]), rstd, nmr), g.w, b.w);
          }
#pragma unroll
          for (int i = 0; i < CW; ++i) raw[i] = __float_as_uint(y[i]);
          tmem_st16(t_o + (uint32_t)(u * CW), raw);         // x1 (fp32) stays in the accumulator: FFN2 adds onto it
#pragma unroll
          for (int c = 0; c < 2; ++c) {                      // x1 (fp16) -> XA sub-tile `part`, 16-byte chunks u*2 + c
            uint4 o;
            __half2* hh = reinterpret_cast<__half2*>(&o);
#pragma unroll
            for (int e = 0; e < 4; ++e) hh[e] = __floats2half2_rn(y[c * 8 + 2 * e], y[c * 8 + 2 * e + 1]);
            sts128(xa_row + (uint32_t)((((u * 2 + c) ^ (row & 7)) & 7) << 4), o);
          }
        }
        tmem_st_wait();
        fence_async_smem();                                  // generic-proxy SMEM writes -> visible to tcgen05.mma (async proxy)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(bar(B_X1));
        TL_TRACE(1, 2);
      }
      // ---------------- hidden chunks: relu(ACC_H + b1) -> fp16 -> HB (A operand of FFN2)
      if (tile + tile_step < total_tiles) {                  // the residual rows LayerNorm1 of the NEXT tile reads: pull them into L2 now
        const int64_t nr = ((tile + tile_step) * 2 + rank) * BM + row;
        if (nr < p.M) {
          const char* a = reinterpret_cast<const char*>(p.x_in + nr * DM + nl);
          asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(a + 128));
        }
      }
      for (int c = 0; c < p.n_chunks; ++c) {
        const int64_t seq = it * p.n_chunks + c;             // completions of the chunk barriers before this chunk
        mbar_wait(bar(B_HFULL), (uint32_t)(seq & 1));
        tc_fence_after();
        TL_TRACE(1, 200 + c);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t r0[CW], r1[CW];
          tmem_ld16_issue(t_h + (uint32_t)(half * HCOLS), r0);
          tmem_ld16_issue(t_h + (uint32_t)(half * HCOLS + CW), r1);
          tmem_ld16_wait(r0);
          tmem_ld16_wait(r1);
          const float* b1 = s_par + 6 * DM + c * WCOLS + half * HCOLS + part * 32;
          uint4 o[4];
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            const uint32_t* src = cc < 2 ? r0 : r1;
            const int off = (cc & 1) * 8;
            const float4 ba = *reinterpret_cast<const float4*>(b1 + cc * 8);
            const float4 bb = *reinterpret_cast<const float4*>(b1 + cc * 8 + 4);
            __half2* hh = reinterpret_cast<__half2*>(&o[cc]);
            hh[0] = __floats2half2_rn(fmaxf(__uint_as_float(src[off]) + ba.x, 0.f), fmaxf(__uint_as_float(src[off + 1]) + ba.y, 0.f));
            hh[1] = __floats2half2_rn(fmaxf(__uint_as_float(src[off + 2]) + ba.z, 0.f), fmaxf(__uint_as_float(src[off + 3]) + ba.w, 0.f));
            hh[2] = __floats2half2_rn(fmaxf(__uint_as_float(src[off + 4]) + bb.x, 0.f), fmaxf(__uint_as_float(src[off + 5]) + bb.y, 0.f));
            hh[3] = __floats2half2_rn(fmaxf(__uint_as_float(src[off + 6]) + bb.z, 0.f), fmaxf(__uint_as_float(src[off + 7]) + bb.w, 0.f));
          }
          if (seq > 0) mbar_wait(bar(B_HBFREE + half), (uint32_t)((seq - 1) & 1));   // FFN2 of the previous chunk's half has finished reading HB[half]
          const uint32_t hrow = hb + (uint32_t)((half * 2 + (part >> 1)) * SUB) + (uint32_t)(row * 128);
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) sts128(hrow + (uint32_t)(((((part & 1) * 4 + cc) ^ (row & 7)) & 7) << 4), o[cc]);
          fence_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(bar(B_HREADY + half));
          TL_TRACE(1, 300 + c * 2 + half);
        }
      }
      // ---------------- LayerNorm2: x = LN(acc + b2) (acc already holds x1 + FFN) -> fp32 stream + fp16 copy
      {
        float s1 = 0.f, s2 = 0.f;
        mbar_wait(bar(B_O2FULL), (uint32_t)(it & 1));
        tc_fence_after();
        TL_TRACE(1, 9);
#pragma unroll
        for (int u = 0; u < NCH; ++u) {
          uint32_t raw[CW];
          tmem_ld16_issue(t_o + (uint32_t)(u * CW), raw);
          tmem_ld16_wait(raw);
#pragma unroll
          for (int i = 0; i < CW; i += 4) {
            const float4 bb = *reinterpret_cast<const float4*>(s_par + DM + nl + u * CW + i);
            const float v0 = __uint_as_float(raw[i]) + bb.x, v1 = __uint_as_float(raw[i + 1]) + bb.y;
            const float v2 = __uint_as_float(raw[i + 2]) + bb.z, v3 = __uint_as_float(raw[i + 3]) + bb.w;
            s1 += (v0 + v1) + (v2 + v3);
            s2 = fmaf(v0, v0, s2); s2 = fmaf(v1, v1, s2); s2 = fmaf(v2, v2, s2); s2 = fmaf(v3, v3, s2);
            raw[i] = __float_as_uint(v0); raw[i + 1] = __float_as_uint(v1); raw[i + 2] = __float_as_uint(v2); raw[i + 3] = __float_as_uint(v3);
          }
          tmem_st16(t_o + (uint32_t)(u * CW), raw);
        }
        tmem_st_wait();
        TL_TRACE(1, 91);
        const uint32_t red1 = red0;
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(red1 + (uint32_t)(((warp - 2) * 32 + lane) * 8)), "f"(s1), "f"(s2) : "memory");
        asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "r"(4 * 32) : "memory");
        float S1 = 0.f, S2 = 0.f;
#pragma unroll
        for (int pp = 0; pp < EPI_WARPS / 4; ++pp) {
          const int e = pp * 4 + ((q - 2) & 3);
          float a, b;
          asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a), "=f"(b) : "r"(red1 + (uint32_t)((e * 32 + lane) * 8)) : "memory");
          S1 += a; S2 += b;
        }
        const float mean = S1 * (1.0f / DM);
        const float var = fmaxf(S2 * (1.0f / DM) - mean * mean, 0.f);
        const float rstd = 1.0f / sqrtf(var + 1e-5f);
        const float nmr = -mean * rstd;
        TL_TRACE(1, 92);
        uint32_t h16[CW];
#pragma unroll
        for (int u = 0; u < NCH; ++u) {
          uint32_t raw[CW];
          tmem_ld16_issue(t_o + (uint32_t)(u * CW), raw);
          tmem_ld16_wait(raw);
          if (u == NCH - 1) tc_fence_before();                // last TMEM read of the tile (ordered before this warp's next barrier arrive)
          float y[CW];
#pragma unroll
          for (int i = 0; i < CW; i += 4) {
            const float4 g = *reinterpret_cast<const float4*>(s_par + 4 * DM + nl + u * CW + i);
            const float4 b = *reinterpret_cast<const float4*>(s_par + 5 * DM + nl + u * CW + i);
            y[i] = fmaf(fmaf(__uint_as_float(raw[i]), rstd, nmr), g.x, b.x);
            y[i + 1] = fmaf(fmaf(__uint_as_float(raw[i + 1]), rstd, nmr), g.y, b.y);
            y[i + 2] = fmaf(fmaf(__uint_as_float(raw[i + 2]), rstd, nmr), g.z, b.z);
            y[i + 3] = fmaf(fmaf(__uint_as_float(raw[i + 3]), rstd, nmr), g.w, b.w);
          }
          // fp32 stream: 16 columns x 32 rows = one staging tile, handed to the TMA engine by one lane. The six stores of a tile
          // (u0, u1, fp16, u2, u3, fp16) alternate between the warp's two staging tiles: store n fills tile n & 1 once store n - 2
          // has been read, i.e. with at most ONE store still pending
          const int n32 = u + (u >= 2 ? 1 : 0);
          const uint32_t st32 = stg + (uint32_t)((n32 & 1) * 2048);
          if (n32 >= 1) { if (lane == 0) bulk_wait_read<1>(); __syncwarp(); }
#pragma unroll
          for (int i = 0; i < 4; ++i)
            sts128(stg_addr(st32, lane, i), make_uint4(__float_as_uint(y[i * 4]), __float_as_uint(y[i * 4 + 1]),
                                                       __float_as_uint(y[i * 4 + 2]), __float_as_uint(y[i * 4 + 3])));
#pragma unroll
          for (int i = 0; i < CW / 2; ++i) {
            const __half2 t = __floats2half2_rn(y[2 * i], y[2 * i + 1]);
            h16[(u & 1) * (CW / 2) + i] = *reinterpret_cast<const uint32_t*>(&t);
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) { tma_store_2d(&map_x32, st32, nl + u * CW, (int)(m_tile * BM) + q * 32); bulk_commit(); }
          if (u & 1) {                                         // fp16 copy of the last two units: 32 columns x 32 rows, into the OTHER tile
            const uint32_t st16 = stg + (uint32_t)(((n32 + 1) & 1) * 2048);
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) sts128(stg_addr(st16, lane, i), make_uint4(h16[i * 4], h16[i * 4 + 1], h16[i * 4 + 2], h16[i * 4 + 3]));
            fence_async_smem();
            __syncwarp();
            if (lane == 0) { tma_store_2d(&map_x16, st16, nl + (u - 1) * CW, (int)(m_tile * BM) + q * 32); bulk_commit(); }
          }
          TL_TRACE(1, 93 + u);
        }
      }
    }
  }

  if (warp >= 2 && lane == 0) bulk_wait<0>();                 // this warp's TMA stores have completed
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

}  // namespace

bool tlayer_tail_supported(int64_t M, int ffn_dim) { return M > 128 && ffn_dim >= 256 && ffn_dim % 256 == 0 && ffn_dim <= MAX_FFN; }

cudaError_t launch_tlayer_tail(void* encode_fn, int num_sms, const TlayerTail& t, cudaStream_t s, char* err, int errlen) {
  if (t.M <= 0) return cudaSuccess;
  auto bad = [&](const char* msg) {
    snprintf(err, errlen, "tlayer_tail: %s (M=%lld ffn=%d)", msg, (long long)t.M, t.ffn_dim);
    return cudaErrorInvalidValue;
  };
  if (!tlayer_tail_supported(t.M, t.ffn_dim)) return bad("needs more than 128 rows and ffn_dim a multiple of 256, at most 1024");
  if (!t.att16 || !t.x32 || !t.x16 || !t.Wo16 || !t.W1_16 || !t.W2_16 || !t.bo || !t.b1 || !t.b2 || !t.ln1_g || !t.ln1_b || !t.ln2_g || !t.ln2_b)
    return bad("NULL argument");
  if ((reinterpret_cast<uintptr_t>(t.att16) | reinterpret_cast<uintptr_t>(t.x32) | reinterpret_cast<uintptr_t>(t.x16) |
       reinterpret_cast<uintptr_t>(t.Wo16) | reinterpret_cast<uintptr_t>(t.W1_16) | reinterpret_cast<uintptr_t>(t.W2_16) |
       reinterpret_cast<uintptr_t>(t.b1)) & 15)
    return bad("pointers must be 16-byte aligned");
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(k_tlayer_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) { snprintf(err, errlen, "cudaFuncSetAttribute(k_tlayer_tail, smem=%d) failed: %s", SMEM_BYTES, cudaGetErrorString(e)); return e; }
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  EncodeTiledFn encode = reinterpret_cast<EncodeTiledFn>(encode_fn);
  CUtensorMap m_att, m_wo, m_w1, m_w2, m_x32, m_x16;
  cuuint32_t es3[3] = {1, 1, 1}, es2[2] = {1, 1};
  {
    // outputs: the epilogue's staging tiles (32 rows x 64 bytes, SWIZZLE_64B) leave through TMA stores; rows past M are clipped
    cuuint64_t gdim[2] = {(cuuint64_t)DM, (cuuint64_t)t.M}, s32[1] = {(cuuint64_t)DM * 4}, s16[1] = {(cuuint64_t)DM * 2};
    cuuint32_t b32[2] = {16, 32}, b16[2] = {32, 32};
    CUresult r = encode(&m_x32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, t.x32, gdim, s32, b32, es2, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_SUCCESS)
      r = encode(&m_x16, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, t.x16, gdim, s16, b16, es2, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { snprintf(err, errlen, "cuTensorMapEncodeTiled(outputs) failed with CUresult %d", (int)r); return cudaErrorInvalidValue; }
  }
  {
    cuuint64_t gdim[3] = {(cuuint64_t)DM, (cuuint64_t)t.M, 1}, gstr[2] = {(cuuint64_t)DM * 2, (cuuint64_t)t.M * DM * 2};
    cuuint32_t box[3] = {BK, BM, 1};
    CUresult r = encode(&m_att, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(t.att16), gdim, gstr, box, es3,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { snprintf(err, errlen, "cuTensorMapEncodeTiled(attn) failed with CUresult %d", (int)r); return cudaErrorInvalidValue; }
  }
  auto wmap = [&](CUtensorMap* m, const __half* W, int K, int N, int box_rows) {
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)N}, gstr[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {BK, (cuuint32_t)box_rows};
    return encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(W), gdim, gstr, box, es2, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  CUresult r = wmap(&m_wo, t.Wo16, DM, DM, DM / 2);
  if (r == CUDA_SUCCESS) r = wmap(&m_w1, t.W1_16, DM, t.ffn_dim, WCOLS / 2);
  if (r == CUDA_SUCCESS) r = wmap(&m_w2, t.W2_16, t.ffn_dim, DM, DM / 2);
  if (r != CUDA_SUCCESS) { snprintf(err, errlen, "cuTensorMapEncodeTiled(weights) failed with CUresult %d", (int)r); return cudaErrorInvalidValue; }

  TlParams p{};
  p.M = t.M; p.m_tiles = (t.M + BM - 1) / BM; p.n_chunks = t.ffn_dim / WCOLS;
  p.x_in = t.x32; p.x_out = t.x32; p.x16_out = t.x16;
  p.bo = t.bo; p.b1 = t.b1; p.b2 = t.b2; p.ln1_g = t.ln1_g; p.ln1_b = t.ln1_b; p.ln2_g = t.ln2_g; p.ln2_b = t.ln2_b;
  const int64_t total = (p.m_tiles + 1) / 2;
  const int64_t clusters = total < num_sms / 2 ? total : num_sms / 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(2 * clusters));
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, k_tlayer_tail, m_x32, m_x16, m_att, m_wo, m_w1, m_w2, p);
}

#ifdef TAG_EXPERIMENTS
extern "C" int tag_exp_set_tlayer_trace(void* dev_buf) {        // experiments build only (tools/tl_trace.py); 3 roles x 4096 int64
  long long* p = reinterpret_cast<long long*>(dev_buf);
  return (int)cudaMemcpyToSymbol(g_tl_trace, &p, sizeof(p));
}
#endif
