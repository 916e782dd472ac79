// Device-side building blocks shared by the tcgen05 kernels (gemm_tc.cu, tlayer_tc.cu): mbarrier / TMA / tcgen05 PTX wrappers,
// shared-memory matrix descriptors, TMEM loads and stores, and the per-warp swizzled staging tile that turns the row-per-lane
// TMEM view into coalesced global accesses. sm_100a only.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#include "common.cuh"

namespace tcx {

constexpr uint32_t kPeerMask = 0xFEFFFFFFu;        // clears the CTA-rank bit of a shared::cluster address -> the pair's leader CTA

// global row of tile row rt (identity unless the tile is (t, window)-ordered)
__device__ __forceinline__ int64_t tile_row(int64_t tile_base, int rt, int lw, int lt) {
  return tile_base + (int64_t)(((rt & ((1 << lw) - 1)) << lt) | (rt >> lw));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// bounded wait: a protocol bug must surface as a launch failure, never as a hung GPU
// The try_wait carries a suspend-time hint: the hardware parks the thread until the phase completes (or the hint expires) instead
// of returning after its short default time-out. Without it the 16 epilogue warps of a GEMM CTA, which wait for an accumulator
// 20-45 % of the time, spent that time executing this loop — ncu: half of the conv1 kernel's warp instructions were
// try_wait / branch / clock / compare — taking issue slots and power from the warps that had work.
#ifndef TAG_MBAR_NO_HINT
#define TAG_MBAR_HINT_OPERAND ", %3"
#else
#define TAG_MBAR_HINT_OPERAND
#endif
constexpr uint32_t kMbarSuspendNs = 200000u;

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2" TAG_MBAR_HINT_OPERAND ";\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(kMbarSuspendNs)
        : "memory");
    if (ok) break;
    if (clock64() - t0 > 4000000000LL) {
      printf("gemm_tc: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

// CTA-pair variants (executed by both CTAs; the mbarrier is the LEADER CTA's)
constexpr uint64_t kEvictNormal = 0x1000000000000000ull;
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar & kPeerMask), "r"(c0), "r"(c1), "r"(c2), "l"(kEvictNormal)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar & kPeerMask), "r"(c0), "r"(c1), "l"(kEvictNormal)
      : "memory");
}
// ---- staging tile -> global through the TMA engine (one elected lane; the tile is a SWIZZLE_64B box of 32 rows x 64 bytes, which is
// exactly the stg_addr layout below). Every lane fences its own st.shared to the async proxy before the warp elects the issuer.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this smem offset in BOTH CTAs of the pair once the MMAs issued so far retire
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerMask) : "memory");
}
// ---- CTA-pair exchange of GroupNorm statistics (T == 256: one window = the pair's two 128-row tiles)
__device__ __forceinline__ uint32_t map_to_peer(uint32_t smem_addr_cta, uint32_t peer_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr_cta), "r"(peer_rank));
  return r;
}
__device__ __forceinline__ void st_peer_v2(uint32_t addr_cluster, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr_cluster), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void mbar_arrive_peer_release(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2" TAG_MBAR_HINT_OPERAND ";\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(kMbarSuspendNs)
        : "memory");
    if (ok) break;
    if (clock64() - t0 > 4000000000LL) {
      printf("gemm_tc: GroupNorm exchange wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

// shared-memory matrix descriptor: K-major operand, 128-byte swizzle, rows 128 B apart, 8-row groups
// 1024 B apart (SBO), descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// same with the matrix base-offset field (bits 49-51) = position of the start row inside its 8-row swizzle atom
__device__ __forceinline__ uint64_t make_smem_desc_off(uint32_t saddr) {
  return make_smem_desc(saddr) | ((uint64_t)((saddr >> 7) & 7u) << 49);
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
// tcgen05.wait::ld with the loaded registers as in/out operands, so no consumer can be scheduled above the wait
__device__ __forceinline__ void tmem_ld16_wait(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :: "memory");
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- epilogue staging ------------------------------------------------------------------------------------------
// TMEM hands a lane one ROW (32 lanes = 32 rows); written straight to global memory that is 16 B per lane at a
// row-stride apart: 32 cache lines per warp instruction, and the LSU (one line per cycle) — not HBM — bounds the
// kernel (micro-benchmark, profiles/r1_tc_microbench_bottleneck.log: the K = 256 GEMMs ran 2x faster with the stores
// removed). So every epilogue global access goes through a per-warp shared-memory tile of 32 rows x 64 B: the
// "row" view (lane = row, 16-byte chunk c) is what the TMEM math reads/writes, the "coalesced" view (instruction j:
// row 8j + lane/4, chunk lane%4) is what global memory sees — 8 rows x 64 contiguous bytes per instruction.
// 16-byte chunks are XOR-swizzled by ((row >> 1) & 3): both views are bank-conflict free.
__device__ __forceinline__ uint32_t stg_addr(uint32_t base, int row, int chunk) {
  return base + (uint32_t)(row * 64) + (uint32_t)(((chunk ^ (row >> 1)) & 3) << 4);
}
__device__ __forceinline__ void sts128(uint32_t a, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
// coalesced global -> registers of one 32-row x 64-byte unit (rows row0.., byte offset `off` in each row of pitch
// `pitch` bytes); rows >= M read as zero
struct RowMap { int64_t tile_base; int rt0; int lw, lt; };      // rows rt0.. of the tile starting at global row tile_base
__device__ __forceinline__ void unit_load(const char* base, int64_t pitch, const RowMap& rm, int64_t M, int64_t off, int lane, uint4 (&r)[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int64_t row = tile_row(rm.tile_base, rm.rt0 + 8 * j + (lane >> 2), rm.lw, rm.lt);
    r[j] = (base != nullptr && row < M) ? __ldg(reinterpret_cast<const uint4*>(base + row * pitch + off + (lane & 3) * 16))
                                        : make_uint4(0u, 0u, 0u, 0u);
  }
}
__device__ __forceinline__ void unit_to_smem(uint32_t stg, int lane, const uint4 (&r)[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) sts128(stg_addr(stg, 8 * j + (lane >> 2), lane & 3), r[j]);
}
// staging tile -> global, coalesced
__device__ __forceinline__ void unit_store(char* base, int64_t pitch, const RowMap& rm, int64_t M, int64_t off, int lane, uint32_t stg) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int64_t row = tile_row(rm.tile_base, rm.rt0 + 8 * j + (lane >> 2), rm.lw, rm.lt);
    const uint4 v = lds128(stg_addr(stg, 8 * j + (lane >> 2), lane & 3));
    if (row < M) *reinterpret_cast<uint4*>(base + row * pitch + off + (lane & 3) * 16) = v;
  }
}

}  // namespace tcx
