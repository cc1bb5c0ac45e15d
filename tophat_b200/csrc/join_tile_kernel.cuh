// join_tile_kernel.cuh -- long_spanning_reads' per-read work as ONE pass over the batch, for reads of up to four segments
// (2x101 bp at the default --segment-length 25; longer layouts take the queue kernels of join_kernel.cuh).
//
// A warp stages the records of 32 consecutive reads in shared memory through the bulk-copy engine (tile_stage.cuh):
// headers, segment counts, read planes, segment hits, CIGAR side records -- five contiguous ranges, one cp.async.bulk each.
// Then, thread = read:
//   1. pair matrices: every hit of segment s against every hit of segment s+1 (dfs_seg_hits' pair test, long_spanning_reads.cpp
//      2250-2558 with --fusion-search off) -> one 64-bit "compatible" and one "abuts exactly" bit matrix per segment pair
//      (8 x 8 hits; a read with more than 8 hits in a segment takes the generic walk);
//   2. chain enumeration = path enumeration over the bit matrices, in dfs_seg_hits' order (2222-2610);
//   3. chains whose neighbouring hits all abut need no closure search: merge_chain (805-2038) is then one streaming pass over
//      the hits (pair checks 930-949, per-hit finalisation 1888-1945) + check_editdist_consistency (bwt_map.cpp:2349-2465) +
//      valid_hit (2045-2099), done here with the read planes and the trailing CIGAR op in registers;
//   4. the other chains (a gap or an overlap somewhere) are queued for chain_merge_kernel, which owns the closures.
//
// Why this shape: round 1's chain_enum / chain_merge_abut kernels were instruction-issue bound, not memory bound -- 870 and
// 1,500 thread-instructions per read at 11 and 19 active lanes of 32 (profiles/r1z_summary.md) -- because the walk kept its
// state in dynamically indexed local arrays and re-fetched header, read and hits per chain through a task queue.
#pragma once
#include "segjuncs_kernel.cuh"
#include "join_kernel.cuh"
#include "tile_stage.cuh"

namespace thb {

constexpr int JT_WARPS = 4;            // warps per CTA, each with its own tile
constexpr int JT_BYTES = 8192;         // shared memory per warp tile (typical need at 2x101 bp: 6.7 KB)
constexpr int JT_SEGS = 4;             // segments per read handled here
constexpr int JT_HITS = 8;             // hits per segment the bit matrices hold

// 64 bases starting at base `start` of a plane held in RW registers
template <int RW>
__device__ __forceinline__ uint64_t reg_slice(const uint64_t (&w)[RW], int start, int n)
{
  const int i = start >> 6, sh = start & 63;
  uint64_t lo = 0, hi = 0;
  #pragma unroll
  for (int k = 0; k < RW; ++k) { if (k == i) lo = w[k]; if (k == i + 1) hi = w[k]; }
  return shr128(lo, hi, sh) & maskn(n);
}

// reverse complement of a read held as RW-word planes (non-ACGT stays N, reads.cpp:191-207)
template <int RW>
__device__ __forceinline__ void reg_revcomp(uint64_t (&p0)[RW], uint64_t (&p1)[RW], uint64_t (&pn)[RW], int n)
{
  const int sh = 64 * RW - n, ws = sh >> 6, bs = sh & 63;
  auto rev = [&](uint64_t (&p)[RW]) {
    uint64_t t[RW + 1];
    #pragma unroll
    for (int w = 0; w < RW; ++w) t[w] = __brevll(p[RW - 1 - w]);
    t[RW] = 0;
    #pragma unroll
    for (int w = 0; w < RW; ++w) {
      uint64_t lo = 0, hi = 0;
      #pragma unroll
      for (int k = 0; k <= RW; ++k) { if (k == w + ws) lo = t[k]; if (k == w + ws + 1) hi = t[k]; }
      p[w] = bs ? ((lo >> bs) | (hi << (64 - bs))) : lo;
    }
  };
  rev(p0); rev(p1); rev(pn);
  #pragma unroll
  for (int w = 0; w < RW; ++w) {
    const int rem = n - 64 * w; const uint64_t valid = rem >= 64 ? ~0ull : (rem > 0 ? ((1ull << rem) - 1ull) : 0ull);
    const uint64_t keep = valid & ~pn[w];
    p0[w] = ~p0[w] & keep; p1[w] = ~p1[w] & keep;
  }
}

// BowtieHit::check_editdist_consistency (bwt_map.cpp:2349-2465) for a single-contig hit, read planes in registers.
// Written as a flat loop over <= 64-base chunks of the match operations: the t-th chunk of every lane is compared in the same
// iteration whatever the shape of the lane's CIGAR (101M and 51M 900N 50M both have two), so the expensive part -- reference
// fetch, read slices, popcounts -- runs converged.  All 32 lanes call; `alive` says whether the lane has a hit to check.
template <int RW>
__device__ __forceinline__ bool editdist_consistent_reg(const RefView& ref, uint32_t ref_id, int left, const uint32_t* ops, int n,
                                                        const uint64_t (&r0)[RW], const uint64_t (&r1)[RW], const uint64_t (&rn)[RW], unsigned mismatches, bool alive)
{
  int64_t len = 0; uint64_t cs = 0;
  if (alive) {
    if (!(ref_id >= 1 && ref_id <= ref.n_contigs)) alive = false;
    else { len = (int64_t)__ldg(ref.contig_len + ref_id - 1); cs = __ldg(ref.contig_start + ref_id - 1); if (len <= 0) alive = false; }
  }
  int64_t pos_ref = left; int pos_seq = 0, i = 0, rem = 0; unsigned mm = 0, nmm = 0;
  #pragma unroll 1
  for (;;) {
    while (alive && rem == 0 && i < n) {                                 // next match operation
      const uint32_t op = ops[i++]; const int c = opc(op), l = (int)opl(op);
      if (c == OP_MATCH) { if (pos_ref < 0 || pos_ref + l > len) alive = false; else rem = l; }   // the reference would read outside the contig
      else if (c == OP_INS) pos_seq += l;
      else if (c == OP_DEL || c == OP_REF_SKIP) pos_ref += l;
    }
    const bool work = alive && rem > 0;
    if (!__any_sync(0xffffffffu, work)) break;
    if (work) {
      const int m = min(64, rem);
      const P3 g = ref_fetch3(ref, cs + (uint64_t)pos_ref, m);
      const uint64_t q0 = reg_slice<RW>(r0, pos_seq, m), q1 = reg_slice<RW>(r1, pos_seq, m), qn = reg_slice<RW>(rn, pos_seq, m);
      mm += (unsigned)__popcll((g.p0 ^ q0) | (g.p1 ^ q1) | (g.pn ^ qn));
      nmm += (unsigned)__popcll(g.pn & qn);
      pos_ref += m; pos_seq += m; rem -= m;
    }
  }
  return alive && (mm == mismatches || mm + nmm == mismatches);
}

// the generic walk for a read with more than JT_HITS hits in a segment: every chain goes to the closure kernel's queue
// (it merges any chain; such reads are rare)
__device__ __noinline__ void enum_big_read(const JoinParams& P, const thb_jhit* HH, const ChainQueue& q, uint32_t bi, int n, const int* cnt, const uint32_t* offs,
                                           unsigned& n_leaves)
{
  int it[JT_SEGS]; uint16_t sel[JT_SEGS]; LiteHit top[JT_SEGS];
  auto lite = [&](uint32_t idx) -> LiteHit {
    const uint4 v = *reinterpret_cast<const uint4*>(HH + idx);
    LiteHit l; l.ref = v.x; l.left = (int)v.y; l.right = (int)v.z; l.anti = (v.w & THB_HIT_ANTISENSE) != 0; l.one_m = (v.w & THB_JHIT_ONE_MATCH) != 0;
    return l;
  };
  auto leaf = [&]() { ++n_leaves; write_chain(q, 0, agg_slot(q.count), bi, n, offs, sel); };
  for (int i0 = 0; i0 < cnt[0]; ++i0) {
    sel[0] = (uint16_t)i0;
    int num_try = 10000;                                           // 2647
    if (n == 1) { leaf(); continue; }
    int lvl = 1; it[1] = 0;
    top[0] = lite(offs[0] + (uint32_t)i0);
    while (lvl >= 1) {
      if (it[lvl] >= cnt[lvl]) { --lvl; if (lvl >= 1) ++it[lvl]; continue; }
      const LiteHit cand = lite(offs[lvl] + (uint32_t)it[lvl]);
      int dist;
      if (!chain_compatible(P, top[lvl - 1], cand, dist)) { ++it[lvl]; continue; }
      sel[lvl] = (uint16_t)it[lvl]; top[lvl] = cand;
      if (lvl == n - 1) { --num_try; leaf(); if (num_try <= 0) break; ++it[lvl]; }
      else { ++lvl; it[lvl] = 0; }
    }
  }
}

// counters: o.counters[0] chains enumerated, [1] closures (chain_merge_kernel), [2] records emitted,
//           tile_counters[0] chains merged here with single-match hits only, [1] other abutting chains merged here
template <int RW, int MINB>
__global__ void __launch_bounds__(JT_WARPS * 32, MINB)
join_tile_kernel(RefView ref, JoinParams P, JoinBatchView bv, ChainQueue q, JoinOut o, unsigned long long* tile_counters)
{
  __shared__ __align__(128) unsigned char tile_smem[JT_WARPS][JT_BYTES];
  __shared__ __align__(8) uint64_t tile_bar[JT_WARPS];
  __shared__ uint32_t tile_park[JT_WARPS][ENUM_PARK][32];       // parked chains: hit index per segment (8 bits each); kind in a lane register
  const unsigned lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
  unsigned char* sm = tile_smem[wib]; uint64_t* bar = &tile_bar[wib];
  if (lane == 0) mbar_init(bar, 1);
  __syncwarp();
  uint32_t parity = 0;
  unsigned n_leaves = 0, n_emit = 0, n_simple = 0, n_abut = 0;
  const uint32_t n_tiles = (bv.n_bundles + 31u) / 32u;
  const uint32_t nsg = bv.n_segs;
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
    pc[2].src = bv.reads + (size_t)b0 * 3u * RW; pc[2].bytes = nb * 3u * RW * 8u; pc[2].off = off; off += 32u * 3u * RW * 8u;
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

    // ---- 1 + 2: pair matrices and chain enumeration (thread = read)
    const bool act = lane < nb;
    const uint32_t bi = b0 + lane;
    int n = 0, read_len = 0; uint32_t ops_begin = 0;
    uint32_t offs[JT_SEGS] = {0u, 0u, 0u, 0u}; int cnt[JT_SEGS] = {0, 0, 0, 0};
    int n_park = 0; uint32_t park_kind = 0;                    // 2 bits per parked chain: 0 needs a closure, 1 single-match hits, 2 abutting
    if (act) {
      const uint4 hdr = s_hdr[lane];
      read_len = (int)(hdr.z & 0xffffu); n = (int)((hdr.z >> 16) & 0xffu); ops_begin = hdr.w;
      if (n < 1 || n > JT_SEGS || n > (int)nsg) n = 0;
      uint32_t a = hdr.y; bool big = false, skip = n == 0;
      #pragma unroll
      for (int s = 0; s < JT_SEGS; ++s) if (s < n) {
        cnt[s] = (int)s_cnt[lane * nsg + s]; offs[s] = a; a += (uint32_t)cnt[s];
        big = big || cnt[s] > JT_HITS;
        if (P.bowtie2 && cnt[s] > P.max_seg_multihits) skip = true;                                  // 2624-2632
      }
      const bool outside = st_hits && (hdr.y < h0 || a > h1);      // a read whose hit range is not inside the staged range
      const thb_jhit* HH = outside ? bv.hits : H;
      if (skip) n = 0;
      else if (big) { enum_big_read(P, HH, q, bi, n, cnt, offs, n_leaves); n = 0; }
      else {
        // pair matrices: bit 8 * i + j of M[s] <=> hit i of segment s and hit j of segment s+1 can be neighbours in a chain.
        // One flat loop over all (segment pair, i, j): the lanes' t-th pair tests run together.
        uint64_t M[JT_SEGS - 1] = {0ull, 0ull, 0ull}, A[JT_SEGS - 1] = {0ull, 0ull, 0ull}; uint32_t onem = 0;
        { const int tot = cnt[0] + cnt[1] + cnt[2] + cnt[3];            // hits of a read are contiguous: hit t = HH[offs[0] + t]
          #pragma unroll 1
          for (int t = 0; t < tot; ++t) if ((*(reinterpret_cast<const uint32_t*>(HH + offs[0] + t) + 3)) & THB_JHIT_ONE_MATCH) onem |= 1u << t; }
        { int ps = 0, pi = 0, pj = 0;
          #pragma unroll 1
          while (ps < n - 1) {
            const int ca = ps == 0 ? cnt[0] : (ps == 1 ? cnt[1] : cnt[2]), cb = ps == 0 ? cnt[1] : (ps == 1 ? cnt[2] : cnt[3]);
            const uint32_t oa = ps == 0 ? offs[0] : (ps == 1 ? offs[1] : offs[2]), ob = ps == 0 ? offs[1] : (ps == 1 ? offs[2] : offs[3]);
            const uint4 v = *reinterpret_cast<const uint4*>(HH + oa + pi), u = *reinterpret_cast<const uint4*>(HH + ob + pj);
            const bool anti = (v.w & THB_HIT_ANTISENSE) != 0;
            const int dist = anti ? (int)v.y - (int)u.z : (int)u.y - (int)v.z;                   // 2355-2379
            const bool okp = u.x == v.x && ((u.w & THB_HIT_ANTISENSE) != 0) == anti                // else: would need a fusion (2402)
                             && dist <= P.max_report_intron && dist >= -P.max_ins;                 // 2554-2556
            const uint64_t bit = okp ? (1ull << (8 * pi + pj)) : 0ull, abit = dist == 0 ? bit : 0ull;
            if (ps == 0) { M[0] |= bit; A[0] |= abit; } else if (ps == 1) { M[1] |= bit; A[1] |= abit; } else { M[2] |= bit; A[2] |= abit; }
            if (++pj == cb) { pj = 0; if (++pi == ca) { pi = 0; ++ps; } }
          } }
        // paths through the matrices = dfs_seg_hits' chains, in its order.  At most 8^3 chains per first-segment hit: the
        // budget of 10,000 (2647) cannot run out here.
        auto leaf = [&](uint32_t sel, bool all_abut) {
          ++n_leaves;
          bool all_one = true;
          #pragma unroll
          for (int s = 0; s < JT_SEGS; ++s) if (s < n) all_one = all_one && ((onem >> ((offs[s] - offs[0]) + ((sel >> (8 * s)) & 0xffu))) & 1u);
          const uint32_t kind = (n > 1 && all_abut) ? (all_one ? 1u : 2u) : 0u;
          if (n_park < ENUM_PARK) { tile_park[wib][n_park][lane] = sel; park_kind |= kind << (2 * n_park); ++n_park; return; }
          // more chains than the parking area holds (a multi-mapped read): the closure kernel merges any chain
          uint16_t s16[JT_SEGS];
          #pragma unroll
          for (int s = 0; s < JT_SEGS; ++s) s16[s] = (uint16_t)((sel >> (8 * s)) & 0xffu);
          write_chain(q, 0, agg_slot(q.count), bi, n, offs, s16);
        };
        for (int i0 = 0; i0 < cnt[0]; ++i0) {
          if (n == 1) { leaf((uint32_t)i0, false); continue; }
          uint32_t m1 = (uint32_t)(M[0] >> (8 * i0)) & 0xffu;
          while (m1) {
            const int i1 = __ffs((int)m1) - 1; m1 &= m1 - 1;
            const bool a1 = (A[0] >> (8 * i0 + i1)) & 1ull;
            const uint32_t s1 = (uint32_t)i0 | ((uint32_t)i1 << 8);
            if (n == 2) { leaf(s1, a1); continue; }
            uint32_t m2 = (uint32_t)(M[1] >> (8 * i1)) & 0xffu;
            while (m2) {
              const int i2 = __ffs((int)m2) - 1; m2 &= m2 - 1;
              const bool a2 = a1 && ((A[1] >> (8 * i1 + i2)) & 1ull);
              const uint32_t s2 = s1 | ((uint32_t)i2 << 16);
              if (n == 3) { leaf(s2, a2); continue; }
              uint32_t m3 = (uint32_t)(M[2] >> (8 * i2)) & 0xffu;
              while (m3) {
                const int i3 = __ffs((int)m3) - 1; m3 &= m3 - 1;
                leaf(s2 | ((uint32_t)i3 << 24), a2 && ((A[2] >> (8 * i2 + i3)) & 1ull));
              }
            }
          }
        }
      }
    }
    __syncwarp();
    // ---- 4: chains that need a closure search (and single-segment reads): one queue reservation for the whole warp
    {
      unsigned c = 0;
      #pragma unroll
      for (int k = 0; k < ENUM_PARK; ++k) c += (k < n_park && ((park_kind >> (2 * k)) & 3u) == 0u) ? 1u : 0u;
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
      #pragma unroll
      for (int k = 0; k < ENUM_PARK; ++k)
        if (k < n_park && ((park_kind >> (2 * k)) & 3u) == 0u) {
          const uint32_t sel = tile_park[wib][k][lane];
          uint16_t s16[JT_SEGS];
          #pragma unroll
          for (int s = 0; s < JT_SEGS; ++s) s16[s] = (uint16_t)((sel >> (8 * s)) & 0xffu);
          write_chain(q, 0, slot++, bi, n, offs, s16);
        }
    }
    // ---- 3: abutting chains, merged here.  Round k: every lane takes its k-th parked chain (most reads have exactly one)
    int kmax = n_park;
    #pragma unroll
    for (int d = 16; d > 0; d >>= 1) kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, d));
    const bool outside = act && n > 0 && st_hits && (offs[0] < h0 || offs[n - 1] + (uint32_t)cnt[n - 1] > h1);
    const thb_jhit* HH = outside ? bv.hits : H; const thb_jops* EE = (outside || !st_ops) ? bv.ops_ext : E;
    for (int k = 0; k < kmax; ++k) {
      const uint32_t kind = (k < n_park) ? ((park_kind >> (2 * k)) & 3u) : 0u;
      bool ok = kind != 0u;
      uint32_t ref0 = 0, last = 0; int left0 = 0, nfl = 0, num_mm = 0, num_smm = 0; bool anti = false, saw_as = false, saw_s = false;
      uint32_t LC[JMAXOPS];                                     // flushed ops; only chains with a multi-op hit touch it
      if (ok) {
        if (kind == 1u) ++n_simple; else ++n_abut;
        const uint32_t sel = tile_park[wib][k][lane];
        anti = ((*(reinterpret_cast<const uint32_t*>(HH + offs[0] + (sel & 0xffu)) + 3)) & THB_HIT_ANTISENSE) != 0;       // chain orientation (2117-2121)
        bool prev_spliced = false, prev_asplice = false, prev_last_match = false; uint32_t prev_ref = 0;
        #pragma unroll 1
        for (int e = 0; e < n; ++e) {
          const int sg = anti ? n - 1 - e : e;
          const uint32_t osg = sg == 0 ? offs[0] : (sg == 1 ? offs[1] : (sg == 2 ? offs[2] : offs[3]));
          const uint4 a = *reinterpret_cast<const uint4*>(HH + osg + ((sel >> (8 * sg)) & 0xffu));
          const uint32_t fl = a.w & 0xfu; const bool asplice = (fl & THB_JHIT_ANTISENSE_SPLICE) != 0, one = (fl & THB_JHIT_ONE_MATCH) != 0;
          num_mm += (int)((a.w >> 16) & 0xffu); num_smm += (int)(a.w >> 24);
          if (e == 0) { ref0 = a.x; left0 = (int)a.y; } else if (a.x != prev_ref) ok = false;
          // every hit goes through the same op loop: a single-match hit is the one-op case (ops synthesised from the 16-byte record)
          int nops = one ? 1 : (int)((a.w >> 4) & 0xfu); if (nops > THB_JHIT_MAX_OPS) nops = THB_JHIT_MAX_OPS;
          if (nops < 1) { ok = false; nops = 0; }
          const uint32_t* po = reinterpret_cast<const uint32_t*>(EE + ops_begin + ((a.w >> 8) & 0xffu));
          bool spliced = false; uint32_t op = 0;
          #pragma unroll 1
          for (int x = 0; x < nops; ++x) {
            op = one ? mkop(OP_MATCH, (uint32_t)(a.z - a.y)) : po[x];
            spliced = spliced || opc(op) == OP_REF_SKIP;
            if (x == 0) {
              if (e > 0 && !(prev_last_match || opc(op) == OP_MATCH)) ok = false;                  // 930-934
              if (last != 0u && opc(last) == opc(op)) { last += opl(op) << 4; continue; }          // equal neighbours fuse (1926-1936)
            }
            if (last != 0u) { if (nfl < JMAXOPS) LC[nfl++] = last; else { ok = false; atomicOr(o.overflow, 2u); } }
            last = op;
          }
          if (e > 0 && prev_spliced && spliced && prev_asplice != asplice) ok = false;             // 942-949
          if (spliced) { if (asplice) { if (saw_s) ok = false; saw_as = true; } else { if (saw_as) ok = false; saw_s = true; } }   // 1888-1945
          prev_spliced = spliced; prev_asplice = asplice; prev_last_match = opc(op) == OP_MATCH; prev_ref = a.x;
        }
        if (last == 0u) ok = false;
      }
      // the chain's CIGAR: LC[0 .. nfl) followed by `last`
      int nLC = nfl + 1;
      const uint32_t* cig = &last;
      if (ok && nfl > 0) { if (nfl < JMAXOPS) { LC[nfl] = last; cig = LC; } else { ok = false; atomicOr(o.overflow, 2u); } }
      if (ok && nfl == 0) nLC = 1;
      __syncwarp();
      {
        // the read, oriented like the chain; new_read_len == old_read_length (2023) holds: fusing equal neighbours keeps lengths
        uint64_t r0[RW], r1[RW], rn[RW];
        { const uint64_t* rd = s_rd + (size_t)lane * 3u * RW;
          #pragma unroll
          for (int w = 0; w < RW; ++w) { r0[w] = rd[w]; r1[w] = rd[RW + w]; rn[w] = rd[2 * RW + w]; } }
        if (ok && anti) reg_revcomp<RW>(r0, r1, rn, read_len);
        ok = editdist_consistent_reg<RW>(ref, ref0, left0, cig, nLC, r0, r1, rn, (unsigned)(num_mm & 0xff), ok);
        if (ok) ok = valid_cigar(P, cig, nLC);
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
            const uint32_t mism = (uint32_t)num_mm & 0xffu, ed = ((uint32_t)num_mm + (uint32_t)cig_gap_length(cig, nLC)) & 0xffu;
            const uint32_t flags = (anti ? (uint32_t)THB_HIT_ANTISENSE : 0u) | (saw_as ? (uint32_t)THB_JHIT_ANTISENSE_SPLICE : 0u);
            uint4* dst = reinterpret_cast<uint4*>(o.rec + slot);
            dst[0] = make_uint4(bi + bv.bundle_base, ref0, (uint32_t)left0, (uint32_t)nLC | (flags << 8) | (mism << 16) | (ed << 24));
            // whole 32-byte sectors only (a partly written sector costs a DRAM read to fill it): ops beyond n_ops are zero
            auto op_at = [&](int x) -> uint32_t { return x < nLC ? cig[x] : 0u; };
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
