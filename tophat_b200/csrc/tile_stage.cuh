// tile_stage.cuh -- warp-granular staging of a batch tile in shared memory with the bulk-copy engine (TMA, 1-D form).
//
// The batch arrays of both stages are SoA and ordered by read, so the records of 32 consecutive reads -- headers, segment
// counts, read planes, segment hits, CIGAR side records -- are five CONTIGUOUS byte ranges.  One lane arms an mbarrier with
// the byte total and issues one cp.async.bulk per range; the copy engine streams them into the warp's slice of shared
// memory while the SM's other warps compute, and the per-read logic then runs on shared-memory latency instead of a chain
// of dependent global loads (header -> counts -> hits -> CIGARs), which is what bounded the round-1 kernels
// (profiles/r1z_summary.md: long-scoreboard stalls, 8-20 of 32 lanes active, DRAM 10-40 %).
//
// Tiles are per WARP, not per CTA: no __syncthreads anywhere, a warp that is done fetches its next tile on its own.
// SASS: UBLKCP.S.G (the copy), SYNCS.ARRIVE.TRANS64 (expect_tx), SYNCS.PHASECHK.TRANS64.TRYWAIT (the wait).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace thb {

#ifndef THB_EMU
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");       // visible to the async proxy before the first copy
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory"); }
// global -> shared bulk copy; dst, src and bytes are multiples of 16
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
  uint32_t ok = 0;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
#else
// host emulation (tests/emu): the copy is a memcpy by the issuing lane; the __syncwarp() every caller places between issue
// and wait orders it before the other lanes' reads
__device__ __forceinline__ void mbar_init(uint64_t*, unsigned) {}
__device__ __forceinline__ void mbar_expect_tx(uint64_t*, uint32_t) {}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t*) { memcpy(dst, src, bytes); }
__device__ __forceinline__ void mbar_wait(uint64_t*, uint32_t) {}
#endif

// One contiguous byte range of a tile
struct TilePiece { const void* src; uint32_t bytes; uint32_t off; };

__device__ __forceinline__ bool piece_bulk_ok(const TilePiece& p)
{ return (p.bytes & 15u) == 0 && (reinterpret_cast<uintptr_t>(p.src) & 15u) == 0; }

// Stages NP pieces into `sm` (the calling warp's slice).  Full tiles of 16-byte-aligned arrays go through the copy engine;
// a tile with an odd piece (the last, partial tile of a batch) is copied by the lanes.  All 32 lanes must call.
template <int NP>
__device__ __forceinline__ void stage_tile(unsigned char* sm, const TilePiece (&pc)[NP], uint64_t* bar, uint32_t& parity, unsigned lane)
{
  bool bulk = true; uint32_t total = 0;
  #pragma unroll
  for (int i = 0; i < NP; ++i) { bulk = bulk && (pc[i].bytes == 0 || piece_bulk_ok(pc[i])); total += pc[i].bytes; }
  if (bulk) {
    if (lane == 0 && total) {
      mbar_expect_tx(bar, total);
      #pragma unroll
      for (int i = 0; i < NP; ++i) if (pc[i].bytes) bulk_g2s(sm + pc[i].off, pc[i].src, pc[i].bytes, bar);
    }
    __syncwarp();
    if (total) { mbar_wait(bar, parity); parity ^= 1u; }
  } else {
    #pragma unroll
    for (int i = 0; i < NP; ++i) {                       // every record type is a multiple of 2 bytes
      const uint16_t* s = reinterpret_cast<const uint16_t*>(pc[i].src); uint16_t* d = reinterpret_cast<uint16_t*>(sm + pc[i].off);
      for (uint32_t k = lane; k < pc[i].bytes / 2; k += 32) d[k] = s[k];
    }
    __syncwarp();
  }
}

}  // namespace thb
