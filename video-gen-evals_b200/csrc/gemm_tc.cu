// Tensor-core GEMM for the encoder's dense layers on sm_100a: TMA -> shared memory (128B swizzle) ->
// tcgen05.mma (fp16 operands, fp32 accumulators in TMEM) -> tcgen05.ld epilogue with fused
// bias / residual / activation.
//
//   C[r, n] = act( sum_{j<taps} sum_{k<K} A[r + (j - taps/2)*dil, k] * W[n, j*K + k] + bias[n] + res[r, n] )
//
// Replaces the reference's nn.Conv1d(k=5, dilation d, zero padding 2d) / 1x1 conv / nn.Linear calls
// (model.py:25-30, :46, :50, :84-97, :145). The dilated conv is NOT lowered through im2col: activations
// are viewed as a 3-D tensor (channel, t, window) and the five taps are five TMA loads whose t
// coordinate is shifted by (j-2)*dil — TMA's out-of-bounds zero fill IS the conv's zero padding — all
// accumulating into one TMEM tile.
//
// Halo mode (the default conv path since round 2): five shifted loads would re-read every activation row five times from
// L2 (~1 GB of the ~2 GB L2->SM traffic of a conv launch). In halo mode the producer loads each 64-channel chunk ONCE
// as a box (channel, window, t) = (64, NW, T' + 4*dil) starting at t = -2*dil: rows land in shared memory ordered
// (t, window) — row = (t + 2*dil) * NW + w — with TMA's zero fill supplying the conv padding rows, and tap j is the
// same tile read through a descriptor whose start address is advanced by j*dil*NW rows. The accumulator rows are then
// (t, window)-ordered too; the epilogue maps them back to [window][t] rows when it addresses global memory.
// Measured on B200 (same box, A/B, round 2: profiles/r2_halo_microbench.log): bit-identical results, 5x less activation
// traffic from L2, conv1 + GELU 225-231 us vs 237-262 us per 400,000-row launch (1.14-1.17 vs 1.00-1.10 PFLOP/s; the library's
// im2col GEMM of the same shape, cuBLAS fp16 [400000 x 1280] x [1280 x 256], runs at 1.09 PFLOP/s), conv2 + GroupNorm 261-269 vs
// 277-303 us, 2.5-3 % on the whole scoring step. (Round 1 measured it 2-3 % slower on that day's boxes and kept it off.)
//
// CTA = 576 threads (18 warps), one CTA per SM, persistent over output tiles; by default two CTAs form a cluster and one
// cta_group::2 MMA (256x256 tile per pair, 128 rows per CTA):
//   warp 0       TMA producer (one elected lane): ring of {A 128x64, B 128x64 (pair) / 256x64 (single)} fp16 tiles
//   warp 1       TMEM allocator + MMA issuer (one elected lane of the leader CTA): 4 x tcgen05.mma M256 N256 K16 per stage
//   warps 2..17  epilogue (4 per TMEM lane quarter x 4 column groups): tcgen05.ld 32 lanes x 16 columns at a time, math in
//                registers, every global access staged through a swizzled per-warp shared-memory tile; 2 TMEM accumulator
//                buffers (2 x 256 columns) so the epilogue of tile i overlaps the MMAs of tile i+1
// MODE 1 fuses GroupNorm over whole (T x 256) windows, MODE 2 LayerNorm over the 256 output columns (see k_gemm_tc).
// Mode 3 of the producer gathers the rows of each window from a per-frame table (GemmTC::g_*, frame-table mode).
// Descriptor encodings follow the PTX ISA tcgen05 matrix/instruction descriptor tables (cross-checked
// against CUTLASS cute/arch/mma_sm100_desc.hpp).
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

constexpr int BLOCK_M = 128, BLOCK_N = 256, BLOCK_K = 64, UMMA_K = 16;
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;       // 16 KiB
// 1-CTA tiles: the CTA stages all 256 weight rows (32 KiB) -> 4 stages of 48 KiB.
// CTA pairs (cta_group::2): each CTA stages its own 128 activation rows and HALF of the weight rows (16 KiB), the
// pair's tcgen05.mma reads both halves -> 6 stages of 32 KiB and 1/3 less L2->SM traffic per FLOP.
template <bool PAIR> struct Cfg {
  static constexpr int STAGES = PAIR ? 5 : 3;
  static constexpr int B_BYTES = (PAIR ? BLOCK_N / 2 : BLOCK_N) * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)((PAIR ? 2 * BLOCK_M : BLOCK_M) >> 4) << 24);
};
// barrier slots (8 B each): full_b[8], empty_b[8], full_a[4], empty_a[4], tmem_full[2], tmem_empty[2]. Without halo
// mode a stage holds {A, B} together and only the *_b barriers are used.
constexpr int MAX_STAGES = 8;
constexpr int MAX_A_STAGES = 4;               // halo mode: activation chunk stages
constexpr int MAX_BARS = 2 * MAX_STAGES + 2 * MAX_A_STAGES + 4;
constexpr int BAR_BYTES = 512;                // barriers, TMEM base word, GroupNorm pair-exchange slots
constexpr int EPI_WARPS = 16;                 // 4 per TMEM lane quarter, 64 accumulator columns each
constexpr int EPI_COLS = BLOCK_N / (EPI_WARPS / 4);   // 64
constexpr int CW = 16;                        // columns per tcgen05.ld chunk
constexpr int THREADS = 64 + EPI_WARPS * 32;
constexpr int TMEM_COLS = 512;
constexpr int GN_RED_BYTES = 2 * EPI_WARPS * 32 * 8;     // double-buffered per-lane (sum, sumsq) exchange
constexpr int RING_BYTES = 5 * (A_BYTES + (BLOCK_N / 2) * BLOCK_K * 2);      // 5 x 32 KiB (pair) >= 3 x 48 KiB (single)
constexpr int STG_WARP_BYTES = 32 * 64;            // per-epilogue-warp staging tile: 32 rows x 64 B, XOR-swizzled
constexpr int STG_BYTES = EPI_WARPS * STG_WARP_BYTES;
constexpr int GN_AFFINE_BYTES = 2 * BLOCK_N * 4;           // gamma[256] || beta[256]
constexpr int SMEM_BYTES = RING_BYTES + 1024 /*align slack*/ + BAR_BYTES + GN_RED_BYTES + GN_AFFINE_BYTES + STG_BYTES;

// tcgen05 instruction descriptor (Cfg::IDESC), kind::f16: D=f32 (bits 4-5 = 1), A=B=f16 (0), K-major A and B,
// N>>3 at bits 17-22, M>>4 at bits 24-28 (M = 256 for a CTA pair).

struct TcParams {
  int64_t M;
  int N;
  int kb_per_tap;       // K / 64
  int gn_xchg;          // fused GroupNorm with T == 256: a window spans both CTAs of the pair, statistics cross over DSMEM
  int kseg;             // plain GEMM over two activation tensors, C = [A | A2] W^T (K2 == K): number of extra K segments (0 / 1),
  int seg_flip;         //   reached through the third TMA coordinate; seg_flip: A2 lies below A in memory
  int taps, dil;
  int mode;             // 0 plain rows, 1 conv with T <= 128 (tile = 128/T windows), 2 conv with T % 128 == 0,
                        // 3 plain GEMM whose rows are gathered window by window from a per-frame table
  int T, wpt, tpw;      // frames per window, windows per tile (mode 1), tiles per window (mode 2)
  int64_t m_tiles;
  int n_tiles;
  const float* bias;
  const __half* res16; int ldr;
  const float* res32;
  __half* C16; int ldc;
  float* C32;
  int act;
  const float* gn_gamma;   // GN instantiation only: GroupNorm(1 group) affine, fused after the activation
  const float* gn_beta;
  const float* ln_gamma;   // MODE 2 only: LayerNorm over the 256 output columns (N == 256), fused after bias + residual;
  const float* ln_beta;    //   writes the fp32 stream (C32, may alias res32) and its fp16 copy (C16)
  // MODE 3 only (TCL forward, losses.py:14-34): the accumulator is the similarity tile S = Z Z^T; the epilogue keeps five masked
  // row sums per (row, 64-column slice) instead of storing S
  const int32_t* tcl_y;    // [M] class of every embedding (rows and columns index the same batch)
  float* tcl_part;         // [M][N / 64][5]: sum_pos exp(S/t), sum_pos exp(-S), sum_neg exp(S/t), sum_pos S/t, #pos
  float tcl_inv_temp;
  int tcl_valid;           // columns >= tcl_valid are padding (N rounded up to 256)
  int c_tma;               // fp16 output through TMA stores of the staging tiles: 0 off, 1 rows linear (2-D map), 2 (t, window)-ordered
                           //   halo tiles (3-D map: column, window, frame)
  int dbg;                 // bottleneck experiments (TAG_TC_DEBUG): 1 no epilogue stores, 2 no weight loads, 4 no activation loads
  // halo mode (conv): A chunk loaded once per 64 channels with its time halo, taps = shifted descriptor views
  int halo;                // 0 off, 1 on, 2 on with the descriptor base-offset field set for unaligned tap shifts (experiment)
  int a_stage_bytes;       // bytes of one staged A chunk (multiple of 1024)
  int a_stages;            // staged A chunks in flight (2..4)
  int a_box_bytes;         // bytes the TMA box delivers (zero-filled rows included)
  int b_stages;            // weight-tile stages behind the two A stages
  int g_L, g_wpv, g_stride;   // mode 3: frames per clip, windows per clip, window stride (frames)
  const float* row0_vec;      // fp16-out path: rows with (row % T) == 0 take this [N] vector
  int tap_rows;            // dil * NW: shared-memory rows between consecutive taps
  int lw, lt;              // log2(windows per tile), log2(frames per window): tile row r <-> window r & (NW-1), frame r >> lw
};

// v = act(acc + bias + res) for CW consecutive columns of one row; the residual comes from the staging tile
// (RES = 0 none, 16: two 16-byte chunks of fp16, 32: four 16-byte chunks of fp32)
template <int RES, int ACT>
__device__ __forceinline__ void epi_values(const TcParams& p, const uint32_t (&raw)[CW], uint32_t stg, int lane, int chunk0,
                                           float (&v)[CW], int n) {
  // per-element arithmetic on packed fp32 pairs (FFMA2 / FADD2: one issue slot per two elements)
  f32x2 w[CW / 2];
#pragma unroll
  for (int i = 0; i < CW / 2; ++i) w[i] = pk2(__uint_as_float(raw[2 * i]), __uint_as_float(raw[2 * i + 1]));
  // no branch between the TMEM load and the stores: register pairs that live across a branch cost a copy per register
  // (the host passes a zero vector when the GEMM has no bias; the activation is a template parameter of the kernel)
#pragma unroll
  for (int i = 0; i < CW; i += 4) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n + i));
    w[i / 2] = add2(w[i / 2], pk2(b.x, b.y));
    w[i / 2 + 1] = add2(w[i / 2 + 1], pk2(b.z, b.w));
  }
  if constexpr (RES == 16) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const uint4 u = lds128(stg_addr(stg, lane, chunk0 + i));
      const __half2* hh = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) { const float2 f = __half22float2(hh[e]); w[i * 4 + e] = add2(w[i * 4 + e], pk2(f.x, f.y)); }
    }
  }
  if constexpr (RES == 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint4 u = lds128(stg_addr(stg, lane, i));
      w[i * 2] = add2(w[i * 2], pk2(__uint_as_float(u.x), __uint_as_float(u.y)));
      w[i * 2 + 1] = add2(w[i * 2 + 1], pk2(__uint_as_float(u.z), __uint_as_float(u.w)));
    }
  }
  if constexpr (ACT == 1) {
#pragma unroll
    for (int i = 0; i < CW / 2; ++i) w[i] = gelu_fast2(w[i]);
  }
#pragma unroll
  for (int i = 0; i < CW / 2; ++i) upk2(w[i], v[2 * i], v[2 * i + 1]);
  if constexpr (ACT == 2) {
#pragma unroll
    for (int i = 0; i < CW; ++i) v[i] = fmaxf(v[i], 0.f);
  }
}

// GN = true: the epilogue additionally applies GroupNorm(1 group, 256 channels) over each whole (T x 256) window
// of the tile (reference model.py:32, :40) — conv2 + residual + GELU + GroupNorm in one kernel. Needs N == 256 and
// T dividing 128 so that a tile owns whole windows; statistics are exchanged between the 2*max(1,T/32) epilogue
// warps that share a window through shared memory and a named barrier.
//
// PAIR = true: the kernel is launched in clusters of two CTAs that form one cta_group::2 MMA (M = 256: each CTA owns
// 128 rows of the tile and their accumulators in its own TMEM). Each CTA's producer loads its 128 activation rows and
// its half of the weight rows; all TMA bytes are counted on the leader CTA's full barrier; the leader's elected thread
// issues the MMAs and its commits are multicast to the empty / accumulator-full barriers of both CTAs; the peer's
// epilogue warps release accumulators on the leader's barrier.
// MODE: 0 plain epilogue, 1 fused GroupNorm (GN), 2 fused LayerNorm (LN): out-proj / FFN2 of the transformer layer
// (post-norm, reference model.py:145): y = LN(acc + bias + x) * gamma + beta, x the fp32 token stream. The pre-norm sum
// is parked in the accumulator's own TMEM columns (tcgen05.st) between the statistics pass and the normalise pass, so
// neither registers nor shared memory have to hold the 128 x 256 fp32 tile.
// ---- fp16 output through TMA stores (round 2). The per-warp staging tile (32 rows x 64 B, 16-byte chunks XOR-ed with
// (row >> 1) & 3) IS the SWIZZLE_64B layout of a [32 rows x 32 fp16] TMA box, so one elected lane can hand the whole tile to the
// TMA engine instead of every lane re-reading it (LDS) and issuing predicated 16-byte global stores with 64-bit row arithmetic:
// fewer instructions on a power-capped board, no LSU work, and the M edge is clipped by the tensor map. For halo tiles the
// 32 rows of a warp are (frame, window)-ordered, which is a 3-D box (32 columns, NW windows, 32 / NW frames).
__device__ __forceinline__ void stg_reuse_wait(int lane) {     // the previous store of this warp has finished READING the staging tile
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  __syncwarp();
}
__device__ __forceinline__ void stg_store_drain(int lane) {    // all stores of this warp have completed (before the CTA exits)
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  __syncwarp();
}
__device__ __forceinline__ void tma_store_unit(const CUtensorMap* map, const TcParams& p, uint32_t stg, int64_t m_tile, int q, int col,
                                               int lane) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // this lane's st.shared -> visible to the async proxy
  __syncwarp();
  if (lane == 0) {
    if (p.c_tma == 1) {
      const int row = (int)(m_tile * BLOCK_M) + q * 32;
      asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                   ::"l"(reinterpret_cast<uint64_t>(map)), "r"(stg), "r"(col), "r"(row) : "memory");
    } else {
      const int nw = 1 << p.lw;
      int w, t;
      if (p.mode == 1) { w = (int)(m_tile * p.wpt); t = 0; }
      else { w = (int)(m_tile / p.tpw); t = (int)(m_tile - (int64_t)w * p.tpw) * BLOCK_M; }
      if (nw >= 32) w += (q * 32) & (nw - 1);
      t += (q * 32) >> p.lw;
      asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                   ::"l"(reinterpret_cast<uint64_t>(map)), "r"(stg), "r"(col), "r"(w), "r"(t) : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
}

// Every epilogue warp waits on the accumulator-full mbarrier itself (parked by the suspend-time hint, tc_common.cuh). The
// alternative — ONE warp watches the mbarrier and the other 15 park at a hardware named barrier, where a waiting warp issues
// nothing — was measured (TAG_EPI_LEADER_POLL, experiments build; same-box A/B in profiles/r2_mbar_hint_ab.log): conv GEMMs
// +1 %, but the epilogue-bound small-K GEMMs -2 % (the per-tile rendezvous couples the 16 warps), nothing on the step.
constexpr int kEpiStartBarrier = 9;
__device__ __forceinline__ void epi_wait_accumulator(uint32_t bar, uint32_t parity, int warp) {
#ifdef TAG_EPI_LEADER_POLL
  if (warp == 2) mbar_wait(bar, parity);
  asm volatile("bar.sync %0, %1;" ::"r"(kEpiStartBarrier), "r"(EPI_WARPS * 32) : "memory");
#else
  (void)warp;
  mbar_wait(bar, parity);
#endif
  tc_fence_after();
}

template <int MODE, bool PAIR, int ACT>
__global__ void __launch_bounds__(THREADS, 1)
k_gemm_tc(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
          const __grid_constant__ CUtensorMap map_c, const TcParams p) {
  constexpr int STAGES = Cfg<PAIR>::STAGES;
  constexpr int STAGE_BYTES = Cfg<PAIR>::STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;           // SWIZZLE_128B needs 1024 B alignment
  const uint32_t bar_base = smem_base + RING_BYTES;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  constexpr bool GN = MODE == 1;
  constexpr bool LN = MODE == 2;
  constexpr bool TCL = MODE == 3;
  // barrier slots (8 B each, layout at MAX_BARS); then the TMEM base word
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };
  auto fulla_bar = [&](int s) { return bar_base + 8u * (2 * MAX_STAGES + s); };
  auto emptya_bar = [&](int s) { return bar_base + 8u * (2 * MAX_STAGES + MAX_A_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * MAX_STAGES + 2 * MAX_A_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * MAX_STAGES + 2 * MAX_A_STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * MAX_BARS;
  // GroupNorm pair exchange (T == 256): two mbarriers and two (sum, sumsq) slots, alternating by tile parity
  auto xbar = [&](int i) { return bar_base + 8u * (MAX_BARS + 2) + 8u * i; };
  auto xslot = [&](int i) { return bar_base + 8u * (MAX_BARS + 4) + 8u * i; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* s_gb = reinterpret_cast<float*>(smem_raw + (bar_base - smem_u32(smem_raw)) + BAR_BYTES + GN_RED_BYTES);
  if constexpr (GN) {
    for (int i = threadIdx.x; i < BLOCK_N; i += THREADS) { s_gb[i] = __ldg(p.gn_gamma + i); s_gb[BLOCK_N + i] = __ldg(p.gn_beta + i); }
  }
  if constexpr (LN) {
    for (int i = threadIdx.x; i < BLOCK_N; i += THREADS) { s_gb[i] = __ldg(p.ln_gamma + i); s_gb[BLOCK_N + i] = __ldg(p.ln_beta + i); }
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < MAX_A_STAGES; ++s) { mbar_init(fulla_bar(s), 1); mbar_init(emptya_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), EPI_WARPS * (PAIR ? 2 : 1)); }
    for (int a = 0; a < 2; ++a) mbar_init(xbar(a), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // whole warp: allocate all 512 TMEM columns (this kernel runs one CTA per SM)
    if constexpr (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();     // both CTAs' barriers are initialised before any remote arrive / TMA signal
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int n_kb = (p.taps + p.kseg) * p.kb_per_tap;
  // work items: (m-tile, n-tile) per CTA, or (pair of m-tiles, n-tile) per cluster
  const int64_t total_tiles = (PAIR ? (p.m_tiles + 1) / 2 : p.m_tiles) * p.n_tiles;
  const int64_t tile0 = PAIR ? (int64_t)(blockIdx.x >> 1) : (int64_t)blockIdx.x;
  const int64_t tile_step = PAIR ? (int64_t)(gridDim.x >> 1) : (int64_t)gridDim.x;
  auto m_tile_of = [&](int64_t tile) { const int64_t mt = tile / p.n_tiles; return PAIR ? mt * 2 + rank : mt; };

  if (warp == 0) {
    if (lane == 0) {
      // ================= TMA producer =================
      int stage = 0; uint32_t phase = 0;
      int sa_i = 0; uint32_t pa = 0;
      const bool ld_a = !(p.dbg & 4), ld_b = !(p.dbg & 2);
      if (p.halo) {
        const int pad = (p.taps / 2) * p.dil;
        const uint32_t b_ring = smem_base + (uint32_t)(p.a_stages * p.a_stage_bytes);
        for (int64_t tile = tile0; tile < total_tiles; tile += tile_step) {
          const int64_t m_tile = m_tile_of(tile);
          const int n_tile = (int)(tile % p.n_tiles);
          int cw, ct;                                           // window / frame coordinate of the box
          if (p.mode == 1) { cw = (int)(m_tile * p.wpt); ct = -pad; }
          else { cw = (int)(m_tile / p.tpw); ct = (int)(m_tile - (int64_t)cw * p.tpw) * BLOCK_M - pad; }
          for (int kc = 0; kc < p.kb_per_tap; ++kc) {
            mbar_wait(emptya_bar(sa_i), pa ^ 1u);
            const uint32_t da = smem_base + (uint32_t)(sa_i * p.a_stage_bytes);
            if constexpr (PAIR) {
              if (leader) { if (ld_a) mbar_arrive_expect_tx(fulla_bar(sa_i), 2u * (uint32_t)p.a_box_bytes); else mbar_arrive(fulla_bar(sa_i)); }
              if (ld_a) tma_load_3d_pair(da, &map_a, fulla_bar(sa_i), kc * BLOCK_K, cw, ct);
            } else {
              if (ld_a) { mbar_arrive_expect_tx(fulla_bar(sa_i), (uint32_t)p.a_box_bytes); tma_load_3d(da, &map_a, fulla_bar(sa_i), kc * BLOCK_K, cw, ct); }
              else mbar_arrive(fulla_bar(sa_i));
            }
            if (++sa_i == p.a_stages) { sa_i = 0; pa ^= 1u; }
            for (int j = 0; j < p.taps; ++j) {
              mbar_wait(empty_bar(stage), phase ^ 1u);
              const uint32_t db = b_ring + (uint32_t)(stage * Cfg<PAIR>::B_BYTES);
              const int kcoord = (j * p.kb_per_tap + kc) * BLOCK_K;
              if constexpr (PAIR) {
                if (leader) { if (ld_b) mbar_arrive_expect_tx(full_bar(stage), 2u * Cfg<PAIR>::B_BYTES); else mbar_arrive(full_bar(stage)); }
                if (ld_b) tma_load_2d_pair(db, &map_b, full_bar(stage), kcoord, n_tile * BLOCK_N + (int)rank * (BLOCK_N / 2));
              } else {
                if (ld_b) { mbar_arrive_expect_tx(full_bar(stage), Cfg<PAIR>::B_BYTES); tma_load_2d(db, &map_b, full_bar(stage), kcoord, n_tile * BLOCK_N); }
                else mbar_arrive(full_bar(stage));
              }
              if (++stage == p.b_stages) { stage = 0; phase ^= 1u; }
            }
          }
        }
      } else
      for (int64_t tile = tile0; tile < total_tiles; tile += tile_step) {
        const int64_t m_tile = m_tile_of(tile);
        const int n_tile = (int)(tile % p.n_tiles);
        int c1_base, c2;
        int grow[8];                                            // mode 3: table row of each window of the tile
        if (p.mode == 3) {
          c1_base = 0; c2 = 0;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int64_t w = m_tile * p.wpt + i;
            const int64_t v = w / p.g_wpv;
            grow[i] = (int)(v * p.g_L + (w - v * p.g_wpv) * p.g_stride);
          }
        } else
        if (p.mode == 0) { c1_base = (int)(m_tile * BLOCK_M); c2 = 0; }
        else if (p.mode == 1) { c1_base = 0; c2 = (int)(m_tile * p.wpt); }
        else { c2 = (int)(m_tile / p.tpw); c1_base = (int)(m_tile - (int64_t)c2 * p.tpw) * BLOCK_M; }
        for (int kb = 0; kb < n_kb; ++kb) {
          const int j = kb / p.kb_per_tap;
          const int kc = kb - j * p.kb_per_tap;
          const int shift = (p.taps > 1) ? (j - p.taps / 2) * p.dil : 0;
          const int c2k = c2 + (p.kseg ? (p.seg_flip ? 1 - j : j) : 0);      // second K segment = second slice of the 3-D map
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          const uint32_t tx = (ld_a ? A_BYTES : 0) + (ld_b ? Cfg<PAIR>::B_BYTES : 0);
          if constexpr (PAIR) {
            if (leader) { if (tx) mbar_arrive_expect_tx(full_bar(stage), 2 * tx); else mbar_arrive(full_bar(stage)); }   // bytes of both CTAs
            if (ld_a && p.mode == 3) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (i < p.wpt) tma_load_3d_pair(sa + (uint32_t)(i * p.T) * 128u, &map_a, full_bar(stage), kc * BLOCK_K, grow[i], 0);
            } else
            if (ld_a) tma_load_3d_pair(sa, &map_a, full_bar(stage), kc * BLOCK_K, c1_base + shift, c2k);
            if (ld_b) tma_load_2d_pair(sa + A_BYTES, &map_b, full_bar(stage), kb * BLOCK_K, n_tile * BLOCK_N + (int)rank * (BLOCK_N / 2));
          } else {
            if (tx) mbar_arrive_expect_tx(full_bar(stage), tx); else mbar_arrive(full_bar(stage));
            if (ld_a && p.mode == 3) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (i < p.wpt) tma_load_3d(sa + (uint32_t)(i * p.T) * 128u, &map_a, full_bar(stage), kc * BLOCK_K, grow[i], 0);
            } else
            if (ld_a) tma_load_3d(sa, &map_a, full_bar(stage), kc * BLOCK_K, c1_base + shift, c2k);
            if (ld_b) tma_load_2d(sa + A_BYTES, &map_b, full_bar(stage), kb * BLOCK_K, n_tile * BLOCK_N);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // ================= MMA issuer (leader CTA of a pair) =================
      int stage = 0; uint32_t phase = 0;
      int sa_i = 0; uint32_t pa = 0;
      int64_t it = 0;
      for (int64_t tile = tile0; tile < total_tiles; tile += tile_step, ++it) {
        const int acc = (int)(it & 1);
        const uint32_t acc_phase = (uint32_t)((it >> 1) & 1);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);           // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
        if (p.halo) {
          const uint32_t b_ring = smem_base + (uint32_t)(p.a_stages * p.a_stage_bytes);
          for (int kc = 0; kc < p.kb_per_tap; ++kc) {
            mbar_wait(fulla_bar(sa_i), pa);                     // this 64-channel chunk (with its time halo) has landed
            tc_fence_after();
            const uint32_t a_chunk = smem_base + (uint32_t)(sa_i * p.a_stage_bytes);
            for (int j = 0; j < p.taps; ++j) {
              mbar_wait(full_bar(stage), phase);
              tc_fence_after();
              const uint32_t a_tap = a_chunk + (uint32_t)(j * p.tap_rows) * 128u;      // tap j = the same tile, j*dil frames later
              const uint64_t adesc = p.halo == 2 ? make_smem_desc_off(a_tap) : make_smem_desc(a_tap);
              const uint64_t bdesc = make_smem_desc(b_ring + (uint32_t)(stage * Cfg<PAIR>::B_BYTES));
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                if constexpr (PAIR) umma_f16_pair(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), Cfg<PAIR>::IDESC, (kc | j | k) != 0 ? 1u : 0u);
                else umma_f16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), Cfg<PAIR>::IDESC, (kc | j | k) != 0 ? 1u : 0u);
              }
              if constexpr (PAIR) umma_commit_pair(empty_bar(stage)); else umma_commit(empty_bar(stage));
              if (++stage == p.b_stages) { stage = 0; phase ^= 1u; }
            }
            if constexpr (PAIR) umma_commit_pair(emptya_bar(sa_i)); else umma_commit(emptya_bar(sa_i));
            if (++sa_i == p.a_stages) { sa_i = 0; pa ^= 1u; }
          }
        } else
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);                  // TMA bytes have landed
          tc_fence_after();
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          const uint64_t adesc = make_smem_desc(sa);
          const uint64_t bdesc = make_smem_desc(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 16 elements (32 B) along K inside the 128 B swizzle row: +2 in the (addr >> 4) field
            if constexpr (PAIR) umma_f16_pair(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), Cfg<PAIR>::IDESC, (kb | k) != 0 ? 1u : 0u);
            else umma_f16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), Cfg<PAIR>::IDESC, (kb | k) != 0 ? 1u : 0u);
          }
          // frees the smem stage (in both CTAs of a pair) when these MMAs retire
          if constexpr (PAIR) umma_commit_pair(empty_bar(stage)); else umma_commit(empty_bar(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        // accumulator complete -> epilogue warps (of both CTAs)
        if constexpr (PAIR) umma_commit_pair(tfull_bar(acc)); else umma_commit(tfull_bar(acc));
      }
    }
  } else {
    // ================= epilogue warps =================
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int part = (warp - 2) >> 2;          // which 64 accumulator columns
    constexpr int NCH = EPI_COLS / CW;         // 4 chunks of 16 columns
    const uint32_t stg = bar_base + (uint32_t)BAR_BYTES + GN_RED_BYTES + GN_AFFINE_BYTES + (uint32_t)((warp - 2) * STG_WARP_BYTES);
    const bool out32 = p.C32 != nullptr;
    int64_t it = 0;
    for (int64_t tile = tile0; tile < total_tiles; tile += tile_step, ++it) {
      const int64_t m_tile = m_tile_of(tile);
      const int n_tile = (int)(tile % p.n_tiles);
      const int acc = (int)(it & 1);
      const uint32_t acc_phase = (uint32_t)((it >> 1) & 1);
      const RowMap rm{m_tile * BLOCK_M, q * 32, p.lw, p.lt};   // this warp's 32 tile rows -> global rows
      const int n_base = n_tile * BLOCK_N + part * EPI_COLS;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + part * EPI_COLS);
      if constexpr (GN) {
        // ---- pass 1: z = GELU(acc + res), per-row partial sums, z stashed as fp16 pairs in registers
        uint32_t stash[EPI_COLS / 2];
        f32x2 s1p = pk2(0.f), s2p = pk2(0.f);                   // even / odd columns
        uint4 rres[4];
        unit_load(reinterpret_cast<const char*>(p.res16), (int64_t)p.ldr * 2, rm, p.M, (int64_t)n_base * 2, lane, rres);
        epi_wait_accumulator(tfull_bar(acc), acc_phase, warp);
        if (p.c_tma) stg_reuse_wait(lane);                      // the previous tile's last TMA store has read the staging tile
#pragma unroll
        for (int u = 0; u < 2; ++u) {                           // units of 32 fp16 columns
          unit_to_smem(stg, lane, rres);
          __syncwarp();
          if (u == 0) unit_load(reinterpret_cast<const char*>(p.res16), (int64_t)p.ldr * 2, rm, p.M, (int64_t)(n_base + 32) * 2, lane, rres);
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const int c = u * 2 + cc;
            uint32_t raw[CW];
            tmem_ld16_issue(t_row + (uint32_t)(c * CW), raw);
            tmem_ld16_wait(raw);
            float v[CW];
            epi_values<16, 1>(p, raw, stg, lane, cc * 2, v, n_base + c * CW);
            // rows past M need no masking: M is a whole number of windows and a tile owns whole windows, so such rows only feed the
            // statistics of windows that are never stored (their activations are TMA zero fill, their residual reads return 0)
#pragma unroll
            for (int i = 0; i < CW / 2; ++i) {
              const f32x2 vv = pk2(v[2 * i], v[2 * i + 1]);
              s1p = add2(s1p, vv);
              s2p = fma2(vv, vv, s2p);
            }
#pragma unroll
            for (int i = 0; i < CW / 2; ++i) {
              const __half2 h2 = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
              stash[c * (CW / 2) + i] = *reinterpret_cast<const uint32_t*>(&h2);
            }
          }
          __syncwarp();                                         // everyone has read its residual rows of this unit
        }
        // the accumulator has been read completely: hand the TMEM buffer back to the MMA warp now
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (PAIR) mbar_arrive_leader(tempty_bar(acc)); else mbar_arrive(tempty_bar(acc)); }
        float s1, s2;
        { float a, b; upk2(s1p, a, b); s1 = a + b; upk2(s2p, a, b); s2 = a + b; }
        // ---- window statistics: rows of a window = min(T,32) lanes x max(1,T/32) quarters x 4 column parts
        const uint32_t red = bar_base + (uint32_t)BAR_BYTES + (uint32_t)((it & 1) * EPI_WARPS * 32 * 8);
        float S1 = 0.f, S2 = 0.f;
        if (p.halo) {
          // (t, window)-ordered tile: lane l of EVERY epilogue warp holds rows of window l & (NW-1)
          for (int o = 16; o >= (1 << p.lw); o >>= 1) {
            s1 += __shfl_xor_sync(FULL_MASK, s1, o);
            s2 += __shfl_xor_sync(FULL_MASK, s2, o);
          }
          asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(red + (uint32_t)(((warp - 2) * 32 + lane) * 8)), "f"(s1), "f"(s2) : "memory");
          asm volatile("bar.sync %0, %1;" ::"r"(1), "r"(EPI_WARPS * 32) : "memory");
#pragma unroll
          for (int e = 0; e < EPI_WARPS; ++e) {
            float a, b;
            asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a), "=f"(b) : "r"(red + (uint32_t)((e * 32 + lane) * 8)) : "memory");
            S1 += a; S2 += b;
          }
        } else {
          const int width = p.T < 32 ? p.T : 32;
          for (int o = width >> 1; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(FULL_MASK, s1, o);
            s2 += __shfl_xor_sync(FULL_MASK, s2, o);
          }
          const int Tl = p.T < BLOCK_M ? p.T : BLOCK_M;       // rows of the window inside this CTA's tile
          const int G = Tl > 32 ? Tl / 32 : 1;
          asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(red + (uint32_t)(((warp - 2) * 32 + lane) * 8)), "f"(s1), "f"(s2) : "memory");
          asm volatile("bar.sync %0, %1;" ::"r"(1 + q / G), "r"(4 * G * 32) : "memory");
          for (int qq = 0; qq < G; ++qq) {
#pragma unroll
            for (int pp = 0; pp < EPI_WARPS / 4; ++pp) {
              const int e = pp * 4 + ((((q / G) * G + qq) - 2) & 3);
              float a, b;
              asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a), "=f"(b) : "r"(red + (uint32_t)((e * 32 + lane) * 8)) : "memory");
              S1 += a; S2 += b;
            }
          }
        }
        if constexpr (PAIR) {
          if (p.gn_xchg) {
            // the window's other 128 rows live in the peer CTA: swap the per-CTA sums through the peer's shared memory.
            // Barrier and slot alternate by tile parity: the peer can be at most one tile ahead (it needs this CTA's sums
            // of tile i+1 to finish tile i+1), so an (i & 1) slot is never overwritten before it has been read.
            const int xi = (int)(it & 1);
            if (warp == 2 && lane == 0) {
              st_peer_v2(map_to_peer(xslot(xi), rank ^ 1u), S1, S2);
              mbar_arrive_peer_release(map_to_peer(xbar(xi), rank ^ 1u));
            }
            mbar_wait_cluster(xbar(xi), (uint32_t)((it >> 1) & 1));
            float a, b;
            asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a), "=f"(b) : "r"(xslot(xi)) : "memory");
            S1 += a; S2 += b;
          }
        }
        const float inv_n = 1.0f / ((float)p.T * (float)BLOCK_N);
        const float mean = S1 * inv_n;
        const float var = fmaxf(S2 * inv_n - mean * mean, 0.f);
        const float rstd = 1.0f / sqrtf(var + 1e-5f);
        const float nmr = -mean * rstd;
        const f32x2 rstd2 = pk2(rstd), nmr2 = pk2(nmr);
        // ---- pass 2: normalise the stash, per-channel affine (from shared memory), stage, store coalesced
        const int nl = part * EPI_COLS;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {                         // 4 chunks of 8 columns
            const int col = nl + u * 32 + i * 8;
            uint4 o;
            __half2* hh = reinterpret_cast<__half2*>(&o);
            const float4 g0 = *reinterpret_cast<const float4*>(s_gb + col);
            const float4 g1 = *reinterpret_cast<const float4*>(s_gb + col + 4);
            const float4 b0 = *reinterpret_cast<const float4*>(s_gb + BLOCK_N + col);
            const float4 b1 = *reinterpret_cast<const float4*>(s_gb + BLOCK_N + col + 4);
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
            if (i == 0 && u == 1 && p.c_tma) stg_reuse_wait(lane);   // unit 0's store has read the tile
            sts128(stg_addr(stg, lane, i), o);
          }
          if (p.c_tma) {
            tma_store_unit(&map_c, p, stg, m_tile, q, n_tile * BLOCK_N + nl + u * 32, lane);
          } else {
            __syncwarp();
            unit_store(reinterpret_cast<char*>(p.C16), (int64_t)p.ldc * 2, rm, p.M, (int64_t)(nl + u * 32) * 2, lane, stg);
            __syncwarp();
          }
        }
      } else if constexpr (LN) {
        // ---- pass 1: v = acc + bias + x (fp32 residual, staged coalesced), row sums, v parked back into TMEM
        float s1 = 0.f, s2 = 0.f;
        uint4 rres[4];
        unit_load(reinterpret_cast<const char*>(p.res32), (int64_t)p.N * 4, rm, p.M, (int64_t)n_base * 4, lane, rres);
        epi_wait_accumulator(tfull_bar(acc), acc_phase, warp);
#pragma unroll
        for (int u = 0; u < NCH; ++u) {
          unit_to_smem(stg, lane, rres);
          __syncwarp();
          if (u + 1 < NCH) unit_load(reinterpret_cast<const char*>(p.res32), (int64_t)p.N * 4, rm, p.M, (int64_t)(n_base + (u + 1) * CW) * 4, lane, rres);
          uint32_t raw[CW];
          tmem_ld16_issue(t_row + (uint32_t)(u * CW), raw);
          tmem_ld16_wait(raw);
          float v[CW];
          epi_values<32, 0>(p, raw, stg, lane, 0, v, n_base + u * CW);
#pragma unroll
          for (int i = 0; i < CW; ++i) { s1 += v[i]; s2 = fmaf(v[i], v[i], s2); raw[i] = __float_as_uint(v[i]); }
          tmem_st16(t_row + (uint32_t)(u * CW), raw);
          __syncwarp();                                         // residual unit consumed before the next one is staged
        }
        tmem_st_wait();
        // ---- row statistics across the 4 column parts of this lane quarter
        const uint32_t red = bar_base + (uint32_t)BAR_BYTES + (uint32_t)((it & 1) * EPI_WARPS * 32 * 8);
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(red + (uint32_t)(((warp - 2) * 32 + lane) * 8)), "f"(s1), "f"(s2) : "memory");
        asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "r"(4 * 32) : "memory");
        float S1 = 0.f, S2 = 0.f;
#pragma unroll
        for (int pp = 0; pp < EPI_WARPS / 4; ++pp) {
          const int e = pp * 4 + ((q - 2) & 3);
          float a, b;
          asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a), "=f"(b) : "r"(red + (uint32_t)((e * 32 + lane) * 8)) : "memory");
          S1 += a; S2 += b;
        }
        const float mean = S1 * (1.0f / BLOCK_N);
        const float var = fmaxf(S2 * (1.0f / BLOCK_N) - mean * mean, 0.f);
        const float rstd = 1.0f / sqrtf(var + 1e-5f);
        const float nmr = -mean * rstd;
        // ---- pass 2: normalise, affine; fp32 stream (16 columns per unit) and fp16 copy (32 columns per two units)
        const int nl = part * EPI_COLS;
        uint32_t h16[CW];                                       // fp16 pairs of two consecutive units
#pragma unroll
        for (int u = 0; u < NCH; ++u) {
          uint32_t raw[CW];
          tmem_ld16_issue(t_row + (uint32_t)(u * CW), raw);
          tmem_ld16_wait(raw);
          if (u == NCH - 1) {                                   // last TMEM read: release the accumulator
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (PAIR) mbar_arrive_leader(tempty_bar(acc)); else mbar_arrive(tempty_bar(acc)); }
          }
          float y[CW];
#pragma unroll
          for (int i = 0; i < CW; i += 4) {
            const float4 g = *reinterpret_cast<const float4*>(s_gb + nl + u * CW + i);
            const float4 b = *reinterpret_cast<const float4*>(s_gb + BLOCK_N + nl + u * CW + i);
            y[i] = fmaf(fmaf(__uint_as_float(raw[i]), rstd, nmr), g.x, b.x);
            y[i + 1] = fmaf(fmaf(__uint_as_float(raw[i + 1]), rstd, nmr), g.y, b.y);
            y[i + 2] = fmaf(fmaf(__uint_as_float(raw[i + 2]), rstd, nmr), g.z, b.z);
            y[i + 3] = fmaf(fmaf(__uint_as_float(raw[i + 3]), rstd, nmr), g.w, b.w);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i)
            sts128(stg_addr(stg, lane, i), make_uint4(__float_as_uint(y[i * 4]), __float_as_uint(y[i * 4 + 1]),
                                                      __float_as_uint(y[i * 4 + 2]), __float_as_uint(y[i * 4 + 3])));
#pragma unroll
          for (int i = 0; i < CW / 2; ++i) {
            const __half2 t = __floats2half2_rn(y[2 * i], y[2 * i + 1]);
            h16[(u & 1) * (CW / 2) + i] = *reinterpret_cast<const uint32_t*>(&t);
          }
          __syncwarp();
          unit_store(reinterpret_cast<char*>(p.C32), (int64_t)p.N * 4, rm, p.M, (int64_t)(n_base + u * CW) * 4, lane, stg);
          __syncwarp();
          if (u & 1) {                                          // two units done: 32 fp16 columns = one 64-byte unit
#pragma unroll
            for (int i = 0; i < 4; ++i) sts128(stg_addr(stg, lane, i), make_uint4(h16[i * 4], h16[i * 4 + 1], h16[i * 4 + 2], h16[i * 4 + 3]));
            __syncwarp();
            unit_store(reinterpret_cast<char*>(p.C16), (int64_t)p.ldc * 2, rm, p.M, (int64_t)(n_base + (u - 1) * CW) * 2, lane, stg);
            __syncwarp();
          }
        }
      } else if constexpr (TCL) {
        // ---- TCL forward: masked row sums of the similarity tile (lane = anchor row i, columns = the other embeddings j)
        const int64_t i = tile_row(rm.tile_base, rm.rt0 + lane, rm.lw, rm.lt);
        constexpr int kPadLabel = (int)0x80000000;                        // INT_MIN never is a class
        const int yi = i < p.M ? __ldg(p.tcl_y + i) : kPadLabel;
        // classes of this warp's 64 columns: two per lane, broadcast by shuffle below
        const int j0 = n_base + lane, j1 = n_base + 32 + lane;
        const int ya = j0 < p.tcl_valid ? __ldg(p.tcl_y + j0) : kPadLabel;
        const int yb = j1 < p.tcl_valid ? __ldg(p.tcl_y + j1) : kPadLabel;
        float e_pos = 0.f, en_pos = 0.f, e_neg = 0.f, s_pos = 0.f, n_pos = 0.f;
        epi_wait_accumulator(tfull_bar(acc), acc_phase, warp);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          uint32_t raw[CW];
          tmem_ld16_issue(t_row + (uint32_t)(c * CW), raw);
          tmem_ld16_wait(raw);
          if (c == NCH - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (PAIR) mbar_arrive_leader(tempty_bar(acc)); else mbar_arrive(tempty_bar(acc)); }
          }
#pragma unroll
          for (int e = 0; e < CW; ++e) {
            const int col = c * CW + e;                                   // 0..63 inside the warp's slice
            const int yj = __shfl_sync(FULL_MASK, col < 32 ? ya : yb, col & 31);
            const float sv = __uint_as_float(raw[e]);
            const float st = sv * p.tcl_inv_temp;
            const float ex = __expf(st);
            if (yj == yi) {
              if ((int64_t)(n_base + col) != i) { e_pos += ex; en_pos += __expf(-sv); s_pos += st; n_pos += 1.f; }
            } else if (yj != kPadLabel) {
              e_neg += ex;
            }
          }
        }
        if (i < p.M) {
          float* o = p.tcl_part + ((size_t)i * (size_t)(p.N / EPI_COLS) + (size_t)(n_base / EPI_COLS)) * 5;
          o[0] = e_pos; o[1] = en_pos; o[2] = e_neg; o[3] = s_pos; o[4] = n_pos;
        }
      } else if (!out32) {
        // ---- fp16 output (optional fp16 residual): 2 units of 32 columns
        const bool has_res = p.res16 != nullptr;
        const bool row_is_t0 = p.row0_vec != nullptr && ((tile_row(rm.tile_base, rm.rt0 + lane, rm.lw, rm.lt) & (int64_t)(p.T - 1)) == 0);
        uint4 rres[4];
        if (has_res) unit_load(reinterpret_cast<const char*>(p.res16), (int64_t)p.ldr * 2, rm, p.M, (int64_t)n_base * 2, lane, rres);
        epi_wait_accumulator(tfull_bar(acc), acc_phase, warp);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (has_res) {
            if (p.c_tma) stg_reuse_wait(lane);
            unit_to_smem(stg, lane, rres);
            __syncwarp();
            if (u == 0) unit_load(reinterpret_cast<const char*>(p.res16), (int64_t)p.ldr * 2, rm, p.M, (int64_t)(n_base + 32) * 2, lane, rres);
          }
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const int c = u * 2 + cc;
            uint32_t raw[CW];
            tmem_ld16_issue(t_row + (uint32_t)(c * CW), raw);
            tmem_ld16_wait(raw);
            float v[CW];
            if (has_res) epi_values<16, ACT>(p, raw, stg, lane, cc * 2, v, n_base + c * CW);
            else epi_values<0, ACT>(p, raw, stg, lane, cc * 2, v, n_base + c * CW);
            if (row_is_t0) {                                   // first frame of a window: the fixed zero-motion row
#pragma unroll
              for (int i = 0; i < CW; i += 4) {
                const float4 z = __ldg(reinterpret_cast<const float4*>(p.row0_vec + n_base + c * CW + i));
                v[i] = z.x; v[i + 1] = z.y; v[i + 2] = z.z; v[i + 3] = z.w;
              }
            }
            if (cc == 0 && !has_res && p.c_tma) stg_reuse_wait(lane);   // the previous unit's TMA store has read the staging tile
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              uint4 o;
              __half2* hh = reinterpret_cast<__half2*>(&o);
#pragma unroll
              for (int e = 0; e < 4; ++e) hh[e] = __floats2half2_rn(v[i * 8 + 2 * e], v[i * 8 + 2 * e + 1]);
              sts128(stg_addr(stg, lane, cc * 2 + i), o);       // over this lane's own (already consumed) residual chunk
            }
          }
          if (u == 1) {                                         // last TMEM read done: release the accumulator early
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (PAIR) mbar_arrive_leader(tempty_bar(acc)); else mbar_arrive(tempty_bar(acc)); }
          }
          if (p.c_tma) {
            if (!(p.dbg & 1)) tma_store_unit(&map_c, p, stg, m_tile, q, n_base + u * 32, lane);
          } else {
            __syncwarp();
            if (!(p.dbg & 1)) unit_store(reinterpret_cast<char*>(p.C16), (int64_t)p.ldc * 2, rm, p.M, (int64_t)(n_base + u * 32) * 2, lane, stg);
            __syncwarp();
          }
        }
      } else {
        // ---- fp32 output (optional fp32 residual, ld = N): 4 units of 16 columns
        const bool has_res = p.res32 != nullptr;
        uint4 rres[4];
        if (has_res) unit_load(reinterpret_cast<const char*>(p.res32), (int64_t)p.N * 4, rm, p.M, (int64_t)n_base * 4, lane, rres);
        epi_wait_accumulator(tfull_bar(acc), acc_phase, warp);
#pragma unroll
        for (int u = 0; u < NCH; ++u) {
          if (has_res) {
            unit_to_smem(stg, lane, rres);
            __syncwarp();
            if (u + 1 < NCH) unit_load(reinterpret_cast<const char*>(p.res32), (int64_t)p.N * 4, rm, p.M, (int64_t)(n_base + (u + 1) * CW) * 4, lane, rres);
          }
          uint32_t raw[CW];
          tmem_ld16_issue(t_row + (uint32_t)(u * CW), raw);
          tmem_ld16_wait(raw);
          float v[CW];
          if (has_res) epi_values<32, ACT>(p, raw, stg, lane, 0, v, n_base + u * CW);
          else epi_values<0, ACT>(p, raw, stg, lane, 0, v, n_base + u * CW);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            sts128(stg_addr(stg, lane, i), make_uint4(__float_as_uint(v[i * 4]), __float_as_uint(v[i * 4 + 1]),
                                                      __float_as_uint(v[i * 4 + 2]), __float_as_uint(v[i * 4 + 3])));
          if (u == NCH - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (PAIR) mbar_arrive_leader(tempty_bar(acc)); else mbar_arrive(tempty_bar(acc)); }
          }
          __syncwarp();
          if (!(p.dbg & 1)) unit_store(reinterpret_cast<char*>(p.C32), (int64_t)p.N * 4, rm, p.M, (int64_t)(n_base + u * CW) * 4, lane, stg);
          __syncwarp();
        }
      }
    }
  }

  if (warp >= 2 && p.c_tma) stg_store_drain(lane);
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();     // no CTA exits (or frees TMEM) while its partner may still signal / read it
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

// the epilogue activation is a template parameter of the plain kernel (MODE 0); GroupNorm always follows a GELU
typedef void (*GemmKernel)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const TcParams);
GemmKernel pick_kernel(bool pair, int mode, int act) {
  if (pair) {
    if (mode == 3) return k_gemm_tc<3, true, 0>;
    if (mode == 2) return k_gemm_tc<2, true, 0>;
    if (mode == 1) return k_gemm_tc<1, true, 1>;
    return act == 1 ? k_gemm_tc<0, true, 1> : act == 2 ? k_gemm_tc<0, true, 2> : k_gemm_tc<0, true, 0>;
  }
  if (mode == 3) return k_gemm_tc<3, false, 0>;
  if (mode == 2) return k_gemm_tc<2, false, 0>;
  if (mode == 1) return k_gemm_tc<1, false, 1>;
  return act == 1 ? k_gemm_tc<0, false, 1> : act == 2 ? k_gemm_tc<0, false, 2> : k_gemm_tc<0, false, 0>;
}
constexpr int kZeroBiasFloats = 16384;

struct TcContext {
  float* zero_bias = nullptr;   // bias of a GEMM that has none (the epilogue adds its bias unconditionally)
  EncodeTiledFn encode = nullptr;
  int num_sms = 148;
  bool pair = true;       // CTA pairs (cta_group::2); TAG_TC_PAIR=0 selects the 1-CTA kernel (A/B testing)
  int dbg = 0;            // TAG_TC_DEBUG bottleneck experiments (results are wrong when set)
  bool tma_store = true;  // fp16 outputs leave through TMA stores of the staging tiles (TAG_TC_TMA_STORE=0 in the experiments build: A/B)
  int b_stage_cap = 0;    // experiments build: TAG_TC_BSTAGES caps the weight stages of halo mode (latency-sensitivity probe)
  int halo_a_stages = 2;  // activation chunks in flight in halo mode (2..4; 3 and 4 measured no faster, and slower at dilation 8
                          // where they leave only 4 weight stages — profiles/r2_halo_microbench.log)
  int halo = 2;           // conv activation loads: 2 (default) = halo tiles for every dilation — one load per 64-channel chunk, the five
                          // taps are descriptor views shifted by j*dil*NW rows (a start row that is not a multiple of 8 is read
                          // correctly by tcgen05.mma on B200 with base-offset 0; held bit-identical to mode 0 by the tests);
                          // 0 = five shifted TMA loads per chunk, 1 = halo only when the tap shift is 8-row aligned,
                          // 3 = 2 with the descriptor base-offset field set (WRONG on B200; probe). Only the experiments
                          // build can change it (TAG_TC_HALO).
};

TcContext* tc_context_create(int device, char* err, int errlen) {
  TcContext* c = new TcContext();
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
    snprintf(err, errlen, "cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: %s", cudaGetErrorString(e));
    delete c;
    return nullptr;
  }
  c->encode = reinterpret_cast<EncodeTiledFn>(fn);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->num_sms = prop.multiProcessorCount;
  e = cudaSuccess;
  for (int pair = 0; pair < 2 && e == cudaSuccess; ++pair)
    for (int mode = 0; mode < 4 && e == cudaSuccess; ++mode)
      for (int act = 0; act < (mode == 0 ? 3 : 1) && e == cudaSuccess; ++act)
        e = cudaFuncSetAttribute(reinterpret_cast<const void*>(pick_kernel(pair != 0, mode, mode == 1 ? 1 : act)),
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e == cudaSuccess) {
    e = cudaMalloc(reinterpret_cast<void**>(&c->zero_bias), (size_t)kZeroBiasFloats * sizeof(float));
    if (e == cudaSuccess) e = cudaMemset(c->zero_bias, 0, (size_t)kZeroBiasFloats * sizeof(float));
  }
#ifdef TAG_EXPERIMENTS   // tools/ build only (build.py --experiments -> libtag_b200_exp.so): the product library reads no environment
  const char* env = getenv("TAG_TC_PAIR");
  if (env != nullptr) c->pair = env[0] != '0';
  env = getenv("TAG_TC_DEBUG");
  if (env != nullptr) c->dbg = atoi(env);
  env = getenv("TAG_TC_HALO");
  if (env != nullptr) c->halo = atoi(env);
  env = getenv("TAG_TC_TMA_STORE");
  if (env != nullptr) c->tma_store = env[0] != '0';
  env = getenv("TAG_TC_BSTAGES");
  if (env != nullptr) c->b_stage_cap = atoi(env);
  env = getenv("TAG_TC_HALO_ASTAGES");
  if (env != nullptr) { c->halo_a_stages = atoi(env); if (c->halo_a_stages < 2) c->halo_a_stages = 2; if (c->halo_a_stages > MAX_A_STAGES) c->halo_a_stages = MAX_A_STAGES; }
#endif
  if (e != cudaSuccess) {
    snprintf(err, errlen, "cudaFuncSetAttribute(k_gemm_tc, smem=%d) failed: %s", SMEM_BYTES, cudaGetErrorString(e));
    delete c;
    return nullptr;
  }
  return c;
}

void tc_context_destroy(TcContext* c) {
  if (c != nullptr && c->zero_bias != nullptr) cudaFree(c->zero_bias);
  delete c;
}
void* tc_encode_fn(const TcContext* c) { return c ? reinterpret_cast<void*>(c->encode) : nullptr; }
int tc_num_sms(const TcContext* c) { return c ? c->num_sms : 148; }
bool tc_pair_enabled(const TcContext* c) { return c != nullptr && c->pair; }

cudaError_t launch_gemm_tc(TcContext* ctx, const GemmTC& g, cudaStream_t s, char* err, int errlen) {
  if (g.M <= 0) return cudaSuccess;
  auto bad = [&](const char* msg) {
    snprintf(err, errlen, "gemm_tc: %s (M=%lld N=%d K=%d taps=%d T=%d lda=%d)", msg, (long long)g.M, g.N, g.K, g.taps, g.T, g.lda);
    return cudaErrorInvalidValue;
  };
  if (ctx == nullptr) return bad("no tensor-core context");
  if (g.N % BLOCK_N) return bad("N must be a multiple of 256");
  if (g.K % BLOCK_K || g.K <= 0) return bad("K must be a positive multiple of 64");
  if (g.lda % 8 || (reinterpret_cast<uintptr_t>(g.A) & 15)) return bad("A must be 16-byte aligned with lda % 8 == 0");
  if (reinterpret_cast<uintptr_t>(g.W) & 15) return bad("W must be 16-byte aligned");
  if (g.A2 != nullptr && (g.taps != 1 || g.g_L > 0 || g.K2 != g.K || g.lda2 != g.lda || g.A2 == g.A ||
                          ((reinterpret_cast<uintptr_t>(g.A2) - reinterpret_cast<uintptr_t>(g.A)) & 15)))
    return bad("a second K segment needs a plain GEMM and a second activation tensor of the same shape and row pitch");
  const bool ln = g.ln_gamma != nullptr;
  const bool tcl = g.tcl_part != nullptr;
  if (tcl) {
    if (g.tcl_y == nullptr || g.taps != 1 || g.C16 != nullptr || g.C32 != nullptr || g.res16 != nullptr || g.res32 != nullptr ||
        g.bias != nullptr || g.gn_gamma != nullptr || ln || g.A2 != nullptr || g.g_L > 0 || g.tcl_valid < 1 || g.tcl_valid > g.N)
      return bad("the TCL epilogue needs a plain GEMM without outputs, bias or residual");
  } else
  if (ln) {
    if (g.ln_beta == nullptr || g.taps != 1 || g.N != BLOCK_N || g.C16 == nullptr || g.C32 == nullptr || g.res32 == nullptr ||
        g.res16 != nullptr || g.gn_gamma != nullptr || g.act != 0)
      return bad("fused LayerNorm needs a plain GEMM with N == 256, an fp32 residual and both outputs");
  } else {
    if ((g.C16 == nullptr) == (g.C32 == nullptr)) return bad("exactly one of C16 / C32 must be given");
    if (g.C32 != nullptr && g.res16 != nullptr) return bad("fp32 output takes an fp32 residual");
    if (g.C16 != nullptr && g.res32 != nullptr) return bad("fp16 output takes an fp16 residual");
  }
  if (g.res16 != nullptr && g.res32 != nullptr) return bad("one residual at most");
  if (g.res32 && (reinterpret_cast<uintptr_t>(g.res32) & 15)) return bad("res32 alignment");
  if (g.C32 && (reinterpret_cast<uintptr_t>(g.C32) & 15)) return bad("C32 alignment");
  if (g.C16 && ((g.ldc % 8) || (reinterpret_cast<uintptr_t>(g.C16) & 15))) return bad("C16 alignment");
  if (g.res16 && ((g.ldr % 8) || (reinterpret_cast<uintptr_t>(g.res16) & 15))) return bad("res16 alignment");

  TcParams p{};
  p.M = g.M; p.N = g.N; p.kb_per_tap = g.K / BLOCK_K; p.taps = g.taps; p.dil = g.dil; p.T = g.T;
  if (g.bias == nullptr && g.N > kZeroBiasFloats) return bad("a GEMM without bias needs N <= 16384");
  if (g.act < 0 || g.act > 2 || (g.gn_gamma != nullptr && g.act != 1)) return bad("act must be 0 (none), 1 (GELU) or 2 (ReLU); the fused GroupNorm follows a GELU");
  p.bias = g.bias != nullptr ? g.bias : ctx->zero_bias; p.res16 = g.res16; p.ldr = g.ldr; p.res32 = g.res32; p.C16 = g.C16; p.ldc = g.ldc; p.C32 = g.C32; p.act = g.act;
  p.gn_gamma = g.gn_gamma; p.gn_beta = g.gn_beta;
  p.ln_gamma = g.ln_gamma; p.ln_beta = g.ln_beta;
  p.tcl_y = g.tcl_y; p.tcl_part = g.tcl_part; p.tcl_inv_temp = g.tcl_inv_temp; p.tcl_valid = g.tcl_valid;
  p.dbg = ctx->dbg;
  const bool gn = g.gn_gamma != nullptr;
  if (gn) {
    const bool pair_ok = ctx->pair && (g.M + BLOCK_M - 1) / BLOCK_M >= 2;
    if (g.gn_beta == nullptr || g.taps <= 1 || g.N != BLOCK_N || g.C16 == nullptr || g.C32 != nullptr ||
        (g.T & (g.T - 1)) != 0 || (g.T > BLOCK_M && !(g.T == 2 * BLOCK_M && pair_ok)))
      return bad("fused GroupNorm needs a conv with N == 256, fp16 output and T a power of two <= 128 (256 with CTA pairs)");
    if (g.T == 2 * BLOCK_M) p.gn_xchg = 1;
  }
  p.n_tiles = g.N / BLOCK_N;
  p.m_tiles = (g.M + BLOCK_M - 1) / BLOCK_M;

  cuuint64_t gdim[3], gstr[2];
  cuuint32_t box[3], estr[3] = {1, 1, 1};
  const bool pair = ctx->pair && p.m_tiles >= 2;
  const int b_bytes = (pair ? BLOCK_N / 2 : BLOCK_N) * BLOCK_K * 2;
  if (g.taps > 1 && ctx->halo != 0 && (g.taps & 1) && !p.gn_xchg) {
    // halo tiles: (t, window)-ordered rows, one load per 64-channel chunk (see the header comment)
    if (g.T >= 1 && g.M % g.T == 0 && ((g.T <= BLOCK_M && BLOCK_M % g.T == 0) || g.T % BLOCK_M == 0)) {
      const int tp = g.T <= BLOCK_M ? g.T : BLOCK_M;              // frames of a tile
      const int nw = BLOCK_M / tp;                                // windows of a tile
      const int pad = (g.taps / 2) * g.dil;
      const int rows = nw * (tp + 2 * pad);
      const int a_stage = (rows * 128 + 1023) & ~1023;
      int a_stages = ctx->halo_a_stages;
      while (a_stages > 2 && (RING_BYTES - a_stages * a_stage) / b_bytes < 3) --a_stages;
      const int b_stages = (RING_BYTES - a_stages * a_stage) / b_bytes;
      const bool aligned = (g.dil * nw) % 8 == 0;
      if (nw <= 32 && tp + 2 * pad <= 256 && g.dil >= 1 && b_stages >= 2 && (aligned || ctx->halo >= 2)) {
        int lw = 0, lt = 0;
        while ((1 << lw) < nw) ++lw;
        while ((1 << lt) < tp) ++lt;
        p.halo = (!aligned && ctx->halo == 3) ? 2 : 1;
        p.a_stage_bytes = a_stage; p.a_box_bytes = rows * 128; p.a_stages = a_stages;
        p.b_stages = b_stages < MAX_STAGES ? b_stages : MAX_STAGES;
        if (ctx->b_stage_cap > 0 && p.b_stages > ctx->b_stage_cap) p.b_stages = ctx->b_stage_cap;
        p.tap_rows = g.dil * nw;
        p.lw = nw > 1 ? lw : 0; p.lt = lt;
      }
    }
  }
  if (p.halo) {
    const int64_t W = g.M / g.T;
    const int tp = g.T <= BLOCK_M ? g.T : BLOCK_M;
    if (g.T <= BLOCK_M) { p.mode = 1; p.wpt = BLOCK_M / g.T; p.tpw = 1; }
    else { p.mode = 2; p.wpt = 1; p.tpw = g.T / BLOCK_M; }
    const int pad = (g.taps / 2) * g.dil;
    box[0] = BLOCK_K; box[1] = (cuuint32_t)p.wpt; box[2] = (cuuint32_t)(tp + 2 * pad);
    gdim[0] = (cuuint64_t)g.K; gdim[1] = (cuuint64_t)W; gdim[2] = (cuuint64_t)g.T;
    gstr[0] = (cuuint64_t)g.T * g.lda * 2; gstr[1] = (cuuint64_t)g.lda * 2;
  } else if (g.taps > 1) {
    if (g.T < 1 || g.M % g.T) return bad("conv rows must be whole windows");
    const int64_t W = g.M / g.T;
    if (g.T <= BLOCK_M) {
      if (BLOCK_M % g.T) return bad("tensor-core conv needs T dividing 128 or a multiple of 128");
      p.mode = 1; p.wpt = BLOCK_M / g.T; p.tpw = 1;
      box[0] = BLOCK_K; box[1] = (cuuint32_t)g.T; box[2] = (cuuint32_t)p.wpt;
    } else {
      if (g.T % BLOCK_M) return bad("tensor-core conv needs T dividing 128 or a multiple of 128");
      p.mode = 2; p.wpt = 1; p.tpw = g.T / BLOCK_M;
      box[0] = BLOCK_K; box[1] = BLOCK_M; box[2] = 1;
    }
    gdim[0] = (cuuint64_t)g.K; gdim[1] = (cuuint64_t)g.T; gdim[2] = (cuuint64_t)W;
    gstr[0] = (cuuint64_t)g.lda * 2; gstr[1] = (cuuint64_t)g.T * g.lda * 2;
  } else if (g.g_L > 0) {
    if (g.T < 16 || g.T > BLOCK_M || (g.T & (g.T - 1)) || g.M % g.T || g.g_wpv < 1 || g.g_stride < 1 || g.g_rows < g.T ||
        g.g_rows >= (1ll << 31) || g.gn_gamma != nullptr || ln)
      return bad("window gather needs a plain GEMM with T a power of two in 16..128");
    p.mode = 3; p.wpt = BLOCK_M / g.T; p.tpw = 1;
    p.g_L = g.g_L; p.g_wpv = g.g_wpv; p.g_stride = g.g_stride;
    box[0] = BLOCK_K; box[1] = (cuuint32_t)g.T; box[2] = 1;
    gdim[0] = (cuuint64_t)g.K; gdim[1] = (cuuint64_t)g.g_rows; gdim[2] = 1;
    gstr[0] = (cuuint64_t)g.lda * 2; gstr[1] = (cuuint64_t)g.g_rows * g.lda * 2;
  } else {
    p.mode = 0; p.wpt = 1; p.tpw = 1;
    box[0] = BLOCK_K; box[1] = BLOCK_M; box[2] = 1;
    gdim[0] = (cuuint64_t)g.K; gdim[1] = (cuuint64_t)g.M; gdim[2] = 1;
    gstr[0] = (cuuint64_t)g.lda * 2; gstr[1] = (cuuint64_t)g.M * g.lda * 2;
    if (g.A2 != nullptr) {
      // the two activation tensors become the two slices of one 3-D map (slice stride = their distance in memory)
      const uintptr_t a = reinterpret_cast<uintptr_t>(g.A), a2 = reinterpret_cast<uintptr_t>(g.A2);
      const uint64_t dist = a2 > a ? a2 - a : a - a2;
      if (dist < (uint64_t)g.M * g.lda * 2 || dist >= (1ull << 40)) return bad("the two K-segment tensors overlap or are too far apart");
      p.kseg = 1; p.seg_flip = a2 < a ? 1 : 0;
      gdim[2] = 2; gstr[1] = dist;
    }
  }
  if (g.row0_vec != nullptr) {
    if (g.C16 == nullptr || g.T < 1 || (g.T & (g.T - 1)) || g.gn_gamma != nullptr || ln || (reinterpret_cast<uintptr_t>(g.row0_vec) & 15))
      return bad("row0_vec needs an fp16-output plain epilogue and T a power of two");
    p.row0_vec = g.row0_vec;
  }
  CUtensorMap map_a, map_b;
  const __half* a_base = (g.A2 != nullptr && p.seg_flip) ? g.A2 : g.A;
  CUresult r = ctx->encode(&map_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(a_base), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { snprintf(err, errlen, "cuTensorMapEncodeTiled(A) failed with CUresult %d", (int)r); return cudaErrorInvalidValue; }
  const int64_t ktot = (int64_t)(g.taps + p.kseg) * g.K;
  cuuint64_t bdim[2] = {(cuuint64_t)ktot, (cuuint64_t)g.N};
  cuuint64_t bstr[1] = {(cuuint64_t)ktot * 2};
  cuuint32_t bbox[2] = {BLOCK_K, (cuuint32_t)(pair ? BLOCK_N / 2 : BLOCK_N)};
  cuuint32_t bes[2] = {1, 1};
  r = ctx->encode(&map_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(g.W), bdim, bstr, bbox, bes,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { snprintf(err, errlen, "cuTensorMapEncodeTiled(W) failed with CUresult %d", (int)r); return cudaErrorInvalidValue; }

  CUtensorMap map_c;
  memset(&map_c, 0, sizeof(map_c));
  if (ctx->tma_store && g.C16 != nullptr && !ln && !tcl && g.ldc % 8 == 0 && reinterpret_cast<uintptr_t>(g.C16) % 16 == 0) {
    CUresult rc3;
    if (p.halo) {
      const int nw = 1 << p.lw, bw = nw < 32 ? nw : 32;
      cuuint64_t cdim[3] = {(cuuint64_t)g.N, (cuuint64_t)(g.M / g.T), (cuuint64_t)g.T};
      cuuint64_t cstr[2] = {(cuuint64_t)g.T * g.ldc * 2, (cuuint64_t)g.ldc * 2};
      cuuint32_t cbox[3] = {32, (cuuint32_t)bw, (cuuint32_t)(32 / bw)};
      rc3 = ctx->encode(&map_c, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, g.C16, cdim, cstr, cbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      p.c_tma = 2;
    } else {
      cuuint64_t cdim[2] = {(cuuint64_t)g.N, (cuuint64_t)g.M};
      cuuint64_t cstr[1] = {(cuuint64_t)g.ldc * 2};
      cuuint32_t cbox[2] = {32, 32};
      rc3 = ctx->encode(&map_c, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, g.C16, cdim, cstr, cbox, bes, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      p.c_tma = 1;
    }
    if (rc3 != CUDA_SUCCESS) { snprintf(err, errlen, "cuTensorMapEncodeTiled(C) failed with CUresult %d", (int)rc3); return cudaErrorInvalidValue; }
  }

  const int kmode = tcl ? 3 : gn ? 1 : ln ? 2 : 0;
  if (pair) {
    const int64_t total = ((p.m_tiles + 1) / 2) * p.n_tiles;
    const int64_t clusters = total < ctx->num_sms / 2 ? total : ctx->num_sms / 2;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(2 * clusters));
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, pick_kernel(true, kmode, p.act), map_a, map_b, map_c, p);
  }
  const int64_t total = p.m_tiles * p.n_tiles;
  const int grid = (int)(total < ctx->num_sms ? total : ctx->num_sms);
  pick_kernel(false, kmode, p.act)<<<grid, THREADS, SMEM_BYTES, s>>>(map_a, map_b, map_c, p);
  return cudaGetLastError();
}
