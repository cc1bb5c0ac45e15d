// join_tile_kernel.cuh -- long_spanning_reads' per-read work as ONE pass over the batch: a warp stages the records of 32
// consecutive reads in shared memory through the bulk-copy engine (tile_stage.cuh), enumerates every read's segment-hit
// chains there (join_segments_for_read 2612-2667 + dfs_seg_hits 2222-2610, long_spanning_reads.cpp), and merges on the
// spot the chains that need no closure search -- every pair of neighbouring hits abuts exactly, the bulk of all chains:
// merge_chain (805-2038) is then one streaming pass (pair checks 930-949, per-hit finalisation 1888-1945), followed by
// check_editdist_consistency (bwt_map.cpp:2349-2465) and valid_hit (2045-2099).  Chains with a gap or an overlap between
// two hits are queued for chain_merge_kernel (join_kernel.cuh), which owns the junction / deletion / insertion closures.
//
// Replaces round 1's chain_enum_kernel + chain_merge_simple_kernel + chain_merge_abut_kernel: those three re-read the
// bundle header, the read planes and the hit records of every chain through a task queue (0.13 GB written and re-read per
// 10 M pairs) and spent their time on dependent global loads; here each input byte crosses HBM once.
#pragma once
#include "segjuncs_kernel.cuh"
#include "join_kernel.cuh"
#include "tile_stage.cuh"

namespace thb {

constexpr int JT_WARPS = 4;            // warps per CTA, each with its own tile
constexpr int JT_BYTES = 8192;         // shared memory per warp tile (typical need at 2x101 bp: 6.7 KB)

// counters: o.counters[0] chains enumerated, [1] closures (chain_merge_kernel), [2] records emitted,
//           tile_counters[0] chains merged here with single-match hits only, [1] other abutting chains merged here
template <int MINB>
__global__ void __launch_bounds__(JT_WARPS * 32, MINB)
join_tile_kernel(RefView ref, JoinParams P, JoinBatchView bv, ChainQueue q, JoinOut o, unsigned long long* tile_counters)
{
  __shared__ __align__(128) unsigned char tile_smem[JT_WARPS][JT_BYTES];
  __shared__ __align__(8) uint64_t tile_bar[JT_WARPS];
  const unsigned lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
  unsigned char* sm = tile_smem[wib]; uint64_t* bar = &tile_bar[wib];
  if (lane == 0) mbar_init(bar, 1);
  __syncwarp();
  uint32_t parity = 0;
  unsigned n_leaves = 0, n_emit = 0, n_simple = 0, n_abut = 0;
  const uint32_t n_tiles = (bv.n_bundles + 31u) / 32u;
  const uint32_t nsg = bv.n_segs, rw = bv.read_words;
  for (uint32_t tile = blockIdx.x * JT_WARPS + wib; tile < n_tiles; tile += gridDim.x * JT_WARPS) {
    const uint32_t b0 = tile * 32u, nb = min(32u, bv.n_bundles - b0);
    // ---- extents of the tile's hit / CIGAR ranges: the headers of its first read and of the read after its last
    uint32_t hx = 0, ex = 0;
    if (lane < 2) {
      const uint32_t bi = b0 + (lane ? nb : 0u);
      if (bi < bv.n_bundles) { const uint4 h = __ldg(reinterpret_cast<const uint4*>(bv.bundles + bi)); hx = h.y; ex = h.w; }
      else { hx = bv.hit_end; ex = bv.ops_end; }
    }
    const uint32_t h0 = __shfl_sync(0xffffffffu, hx, 0), h1 = __shfl_sync(0xffffffffu, hx, 1);
    const uint32_t e0 = __shfl_sync(0xffffffffu, ex, 0), e1 = __shfl_sync(0xffffffffu, ex, 1);
    // ---- layout: headers | segment counts | read planes | hits | CIGAR side records; what does not fit stays in HBM
    TilePiece pc[5];
    uint32_t off = 0;
    pc[0].src = bv.bundles + b0; pc[0].bytes = nb * 16u; pc[0].off = off; off += 32u * 16u;
    pc[1].src = bv.seg_count + (size_t)b0 * nsg; pc[1].bytes = nb * nsg * 2u; pc[1].off = off; off += (32u * nsg * 2u + 15u) & ~15u;
    pc[2].src = bv.reads + (size_t)b0 * 3u * rw; pc[2].bytes = nb * 3u * rw * 8u; pc[2].off = off; off += 32u * 3u * rw * 8u;
    const bool sane = h1 >= h0 && e1 >= e0;
    const uint32_t hbytes = sane ? (h1 - h0) * 16u : 0u, ebytes = sane ? (e1 - e0) * (uint32_t)sizeof(thb_jops) : 0u;
    const bool st_hits = sane && off + hbytes <= (uint32_t)JT_BYTES;
    pc[3].src = bv.hits + h0; pc[3].bytes = st_hits ? hbytes : 0u; pc[3].off = off; if (st_hits) off += hbytes;
    const bool st_ops = st_hits && off + ebytes <= (uint32_t)JT_BYTES;
    pc[4].src = bv.ops_ext + e0; pc[4].bytes = st_ops ? ebytes : 0u; pc[4].off = off;
    stage_tile<5>(sm, pc, bar, parity, lane);
    // virtual array bases: absolute hit / CIGAR indices keep working whether the records sit in shared memory or in HBM
    const thb_jhit* H = st_hits ? reinterpret_cast<const thb_jhit*>(sm + pc[3].off) - h0 : bv.hits;
    const thb_jops* E = st_ops ? reinterpret_cast<const thb_jops*>(sm + pc[4].off) - e0 : bv.ops_ext;
    const uint4* s_hdr = reinterpret_cast<const uint4*>(sm);
    const uint16_t* s_cnt = reinterpret_cast<const uint16_t*>(sm + pc[1].off);
    const uint64_t* s_rd = reinterpret_cast<const uint64_t*>(sm + pc[2].off);

    // ---- enumeration (thread = read): dfs_seg_hits' order and its budget of 10,000 complete chains per first-segment hit
    const bool act = lane < nb;
    const uint32_t bi = b0 + lane;
    ParkedChains park; park.n = 0;
    uint32_t offs[JMAXSEGS]; int n = 0, read_len = 0; uint32_t ops_begin = 0;
    bool fallback = false;             // this read's records are not (all) inside the staged ranges: leave it to the queue path
    if (act) {
      const uint4 hdr = s_hdr[lane];
      read_len = (int)(hdr.z & 0xffffu); n = (int)((hdr.z >> 16) & 0xffu); ops_begin = hdr.w;
      if (n < 1 || n > JMAXSEGS || n > (int)nsg) n = 0;
      int cnt[JMAXSEGS]; uint32_t a = hdr.y;
      for (int s = 0; s < n; ++s) { cnt[s] = (int)s_cnt[lane * nsg + s]; offs[s] = a; a += (uint32_t)cnt[s]; }
      bool skip = n == 0;
      if (P.bowtie2) for (int s = 0; s < n; ++s) if (cnt[s] > P.max_seg_multihits) skip = true;        // 2624-2632
      if (st_hits && (hdr.y < h0 || a > h1)) fallback = true;      // bundles whose hit ranges are not laid out back to back
      const thb_jhit* HH = fallback ? bv.hits : H;
      if (!skip) {
        int it[JMAXSEGS]; uint16_t sel[JMAXSEGS]; LiteHit top[JMAXSEGS]; bool simp[JMAXSEGS], abut[JMAXSEGS];
        auto load_lite_g = [&](uint32_t idx) -> LiteHit {
          const uint4 v = *reinterpret_cast<const uint4*>(HH + idx);
          LiteHit l; l.ref = v.x; l.left = (int)v.y; l.right = (int)v.z; l.anti = (v.w & THB_HIT_ANTISENSE) != 0; l.one_m = (v.w & THB_JHIT_ONE_MATCH) != 0;
          return l;
        };
        auto leaf = [&]() {
          ++n_leaves;
          const bool simple = n > 1 && simp[n - 1];
          const bool abutting = !simple && n > 1 && abut[n - 1];
          const int kind = simple ? 1 : (abutting ? 2 : 0);
          if (park.n < ENUM_PARK) {
            for (int s = 0; s < n; ++s) park.sel[park.n][s] = sel[s];
            park.kind[park.n] = (uint8_t)kind; ++park.n;
            return;
          }
          // more chains than the parking area holds (a multi-mapped read): the general kernel merges any chain
          const unsigned long long slot = agg_slot(q.count);
          write_chain(q, 0, slot, bi, n, offs, sel);
        };
        for (int i0 = 0; i0 < cnt[0]; ++i0) {
          sel[0] = (uint16_t)i0;
          int num_try = 10000;                                           // 2647
          if (n == 1) { --num_try; leaf(); continue; }
          int lvl = 1; it[1] = 0;
          top[0] = load_lite_g(offs[0] + (uint32_t)i0); simp[0] = top[0].one_m; abut[0] = true;
          while (lvl >= 1) {
            if (it[lvl] >= cnt[lvl]) { --lvl; if (lvl >= 1) ++it[lvl]; continue; }
            const LiteHit cand = load_lite_g(offs[lvl] + (uint32_t)it[lvl]);
            int dist;
            if (!chain_compatible(P, top[lvl - 1], cand, dist)) { ++it[lvl]; continue; }
            sel[lvl] = (uint16_t)it[lvl]; top[lvl] = cand; simp[lvl] = simp[lvl - 1] && cand.one_m && dist == 0; abut[lvl] = abut[lvl - 1] && dist == 0;
            if (lvl == n - 1) { --num_try; leaf(); if (num_try <= 0) break; ++it[lvl]; }
            else { ++lvl; it[lvl] = 0; }
          }
        }
      }
    }
    __syncwarp();
    // ---- chains that need a closure search (and single-segment reads): one queue reservation for the whole warp
    {
      unsigned c = 0;
      for (int k = 0; k < park.n; ++k) c += park.kind[k] == 0 ? 1u : 0u;
      unsigned incl = c;
      #pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, incl, d); if ((int)lane >= d) incl += v; }
      const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
      unsigned long long base = 0;
      if (total) {
        if (lane == 0) base = atomicAdd(q.count, (unsigned long long)total);
        base = __shfl_sync(0xffffffffu, base, 0);
      }
      unsigned long long slot = base + (unsigned long long)(incl - c);
      for (int k = 0; k < park.n; ++k) if (park.kind[k] == 0) write_chain(q, 0, slot++, bi, n, offs, park.sel[k]);
    }
    __syncwarp();
    // ---- abutting chains, merged here.  Round k: every lane takes its k-th parked chain (most reads have exactly one)
    int kmax = park.n;
    #pragma unroll
    for (int d = 16; d > 0; d >>= 1) kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, d));
    for (int k = 0; k < kmax; ++k) {
      bool ok = act && k < park.n && park.kind[k] != 0;
      uint32_t ref0 = 0; int left0 = 0, nLC = 0, num_mm = 0, num_smm = 0; bool anti = false, saw_as = false, saw_s = false;
      uint32_t LC[JMAXOPS];
      if (ok) {
        if (park.kind[k] == 1) ++n_simple; else ++n_abut;
        const thb_jhit* HH = fallback ? bv.hits : H; const thb_jops* EE = (fallback || !st_ops) ? bv.ops_ext : E;
        const uint16_t* sel = park.sel[k];
        anti = ((*(reinterpret_cast<const uint32_t*>(HH + offs[0] + sel[0]) + 3)) & THB_HIT_ANTISENSE) != 0;       // chain orientation (2117-2121)
        bool prev_spliced = false, prev_asplice = false, prev_last_match = false; uint32_t prev_ref = 0;
        for (int e = 0; e < n && ok; ++e) {
          const int sg = anti ? n - 1 - e : e;
          const uint4 a = *reinterpret_cast<const uint4*>(HH + offs[sg] + sel[sg]);
          const uint32_t fl = a.w & 0xfu; const bool asplice = (fl & THB_JHIT_ANTISENSE_SPLICE) != 0;
          int nops = 1; uint32_t ops[THB_JHIT_MAX_OPS]; ops[0] = mkop(OP_MATCH, (uint32_t)(a.z - a.y));
          if (!(fl & THB_JHIT_ONE_MATCH)) {
            nops = (int)((a.w >> 4) & 0xfu); if (nops > THB_JHIT_MAX_OPS) nops = THB_JHIT_MAX_OPS;
            const uint4* p = reinterpret_cast<const uint4*>(EE + ops_begin + ((a.w >> 8) & 0xffu));
            const uint4 b = p[0]; ops[0] = b.x; ops[1] = b.y; ops[2] = b.z; ops[3] = b.w;
            if (nops > 4) { const uint4 c = p[1]; ops[4] = c.x; ops[5] = c.y; ops[6] = c.z; ops[7] = c.w; }
            if (nops > 8) ops[8] = *reinterpret_cast<const uint32_t*>(p + 2);
          }
          if (nops < 1) { ok = false; break; }
          bool spliced = false;
          for (int x = 0; x < nops; ++x) spliced = spliced || opc(ops[x]) == OP_REF_SKIP;
          if (e == 0) { ref0 = a.x; left0 = (int)a.y; }
          else {
            if (!(prev_last_match || opc(ops[0]) == OP_MATCH)) { ok = false; break; }               // 930-934
            if (prev_spliced && spliced && prev_asplice != asplice) { ok = false; break; }            // 942-949
            if (a.x != prev_ref) { ok = false; break; }
          }
          // finalise this hit (1888-1945)
          num_mm += (int)((a.w >> 16) & 0xffu); num_smm += (int)(a.w >> 24);
          if (spliced) { if (asplice) { if (saw_s) { ok = false; break; } saw_as = true; } else { if (saw_as) { ok = false; break; } saw_s = true; } }
          int x0 = 0;
          if (nLC > 0 && opc(LC[nLC - 1]) == opc(ops[0])) { LC[nLC - 1] = mkop(opc(LC[nLC - 1]), opl(LC[nLC - 1]) + opl(ops[0])); x0 = 1; }
          for (; x0 < nops; ++x0) if (!cig_push(LC, nLC, ops[x0])) { ok = false; break; }
          prev_spliced = spliced; prev_asplice = asplice; prev_last_match = opc(ops[nops - 1]) == OP_MATCH; prev_ref = a.x;
        }
        if (ok && nLC == 0) ok = false;
      }
      __syncwarp();
      if (ok) {
        // the read, oriented like the chain; new_read_len == old_read_length (2023) holds: fusing equal neighbours keeps lengths
        uint64_t R[12];
        { const uint64_t* rd = s_rd + (size_t)lane * 3u * rw;
          #pragma unroll
          for (int pl = 0; pl < 3; ++pl)
            #pragma unroll
            for (int w = 0; w < 4; ++w) R[pl * 4 + w] = w < (int)rw ? rd[pl * rw + w] : 0ull; }
        if (anti) { uint64_t F[12];
          #pragma unroll
          for (int x = 0; x < 12; ++x) F[x] = R[x];
          revcomp_read(F, read_len, R); }
        ok = editdist_consistent(ref, ref0, left0, LC, nLC, R, 4, (uint8_t)num_mm) && valid_cigar(P, LC, nLC);
      }
      const unsigned em = __ballot_sync(0xffffffffu, ok);
      if (em) {
        unsigned long long slot0 = 0;
        if (lane == (unsigned)(__ffs((int)em) - 1)) slot0 = atomicAdd(o.count, (unsigned long long)__popc(em));
        slot0 = __shfl_sync(0xffffffffu, slot0, __ffs((int)em) - 1);
        if (ok) {
          const unsigned long long slot = slot0 + (unsigned long long)__popc(em & ((1u << lane) - 1u));
          if (slot >= o.cap) atomicOr(o.overflow, 1u);
          else {
            const uint32_t mism = (uint32_t)num_mm & 0xffu, ed = ((uint32_t)num_mm + (uint32_t)cig_gap_length(LC, nLC)) & 0xffu;
            const uint32_t flags = (anti ? (uint32_t)THB_HIT_ANTISENSE : 0u) | (saw_as ? (uint32_t)THB_JHIT_ANTISENSE_SPLICE : 0u);
            uint4* dst = reinterpret_cast<uint4*>(o.rec + slot);
            dst[0] = make_uint4(bi + bv.bundle_base, ref0, (uint32_t)left0, (uint32_t)nLC | (flags << 8) | (mism << 16) | (ed << 24));
            // whole 32-byte sectors only (a partly written sector costs a DRAM read to fill it): ops beyond n_ops are zero
            auto op_at = [&](int x) -> uint32_t { return x < nLC ? LC[x] : 0u; };
            dst[1] = make_uint4((uint32_t)num_smm & 0xffu, op_at(0), op_at(1), op_at(2));
            for (int qd = 2; 4 * qd - 5 < nLC; qd += 2) {
              dst[qd] = make_uint4(op_at(4 * qd - 5), op_at(4 * qd - 4), op_at(4 * qd - 3), op_at(4 * qd - 2));
              dst[qd + 1] = make_uint4(op_at(4 * qd - 1), op_at(4 * qd), op_at(4 * qd + 1), op_at(4 * qd + 2));
            }
          }
          ++n_emit;
        }
      }
      __syncwarp();
    }
    __syncwarp();             // every lane is done with the tile before the next one overwrites it
  }
  for (int k = 16; k > 0; k >>= 1) {
    n_leaves += __shfl_xor_sync(0xffffffffu, n_leaves, k); n_emit += __shfl_xor_sync(0xffffffffu, n_emit, k);
    n_simple += __shfl_xor_sync(0xffffffffu, n_simple, k); n_abut += __shfl_xor_sync(0xffffffffu, n_abut, k);
  }
  if (lane == 0) {
    if (n_leaves) atomicAdd(o.counters + 0, (unsigned long long)n_leaves);
    if (n_emit) atomicAdd(o.counters + 2, (unsigned long long)n_emit);
    if (n_simple) atomicAdd(tile_counters + 0, (unsigned long long)n_simple);
    if (n_abut) atomicAdd(tile_counters + 1, (unsigned long long)n_abut);
  }
}

}  // namespace thb
