// scan_tile_kernel.cuh -- the per-read bookkeeping of segment_juncs as ONE pass over the batch: a warp stages the records
// of 32 consecutive bundles (headers, segment counts, read planes, segment hits, partner hits) in shared memory through the
// bulk-copy engine (tile_stage.cuh) and runs, on shared-memory latency,
//   phase A (thread = bundle): segment bookkeeping, multihit guard, "must the mate anchor this read" (find_gaps 3361-3392),
//                              mate-flank rescue tasks, the hit -> (bundle, segment) map of the tile;
//   phase B (thread = hit)   : the hit against the hits of the next one / two segments of its read -- indel pair checks
//                              (find_insertions_and_deletions 2856-2938) and window construction (find_gaps 3508-3617),
//                              task emission for window_scan_kernel / indel_kernel.
// Replaces round 1's bundle_kernel + hit_kernel, which read the batch twice and exchanged a 4-byte-per-hit owner map and a
// per-bundle state word through HBM (bundle_kernel: 2.07x its algorithmic bytes in DRAM traffic, profiles/r1z_summary.md).
// Line numbers: segment_juncs.cpp of the reference.
#pragma once
#include "segjuncs_kernel.cuh"
#include "tile_stage.cuh"

namespace thb {

constexpr int ST_WARPS = 4;
template <int NSMAX> struct ScanTile { static constexpr int BYTES = NSMAX <= 4 ? 8192 : 11264; };

// The hit ranges of consecutive bundles must lie back to back (bundle i+1 starts where bundle i ends): that is how both
// hosts lay a batch out and what makes a tile's hits one contiguous range.  err bit 16: violated (host -> THB_EINVAL).
template <int NSMAX, int MINB>
__global__ void __launch_bounds__(ST_WARPS * 32, MINB)
scan_tile_kernel(RefView ref, SegParams P, BatchView bv, Queues q, uint32_t* __restrict__ bstate, uint32_t* __restrict__ owner_g, SegOutputs out)
{
  constexpr int TB = ScanTile<NSMAX>::BYTES;
  __shared__ __align__(128) unsigned char tile_smem[ST_WARPS][TB];
  __shared__ __align__(8) uint64_t tile_bar[ST_WARPS];
  __shared__ uint32_t tile_state[ST_WARPS][32];
  const unsigned lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
  unsigned char* sm = tile_smem[wib]; uint64_t* bar = &tile_bar[wib]; uint32_t* s_state = tile_state[wib];
  if (lane == 0) mbar_init(bar, 1);
  __syncwarp();
  uint32_t parity = 0;
  unsigned n_windows = 0, n_indel = 0;
  const uint32_t n_tiles = (bv.n_bundles + 31u) / 32u;
  const uint32_t nsg = bv.n_segs, rw = bv.read_words;
  const int look_bp = 8;
  for (uint32_t tile = blockIdx.x * ST_WARPS + wib; tile < n_tiles; tile += gridDim.x * ST_WARPS) {
    const uint32_t b0 = tile * 32u, nb = min(32u, bv.n_bundles - b0);
    // ---- extents of the tile's hit / partner-hit ranges
    uint64_t hx = 0, px = 0;
    if (lane < 2) {
      const uint32_t bi = b0 + (lane ? nb : 0u);
      if (bi < bv.n_bundles) { const uint4 h = __ldg(reinterpret_cast<const uint4*>(bv.bundles + bi)); hx = h.y; px = h.z; }
      else { hx = bv.hit_end; px = bv.partner_end; }
    }
    const uint64_t h0 = __shfl_sync(0xffffffffu, hx, 0), h1 = __shfl_sync(0xffffffffu, hx, 1);
    const uint64_t p0 = __shfl_sync(0xffffffffu, px, 0), p1 = __shfl_sync(0xffffffffu, px, 1);
    // ---- layout: headers | segment counts | read planes | partner hits | hits | owner map
    TilePiece pc[5];
    uint32_t off = 0;
    pc[0].src = bv.bundles + b0; pc[0].bytes = nb * 16u; pc[0].off = off; off += 32u * 16u;
    pc[1].src = bv.seg_count + (size_t)b0 * nsg; pc[1].bytes = nb * nsg * 2u; pc[1].off = off; off += (32u * nsg * 2u + 15u) & ~15u;
    pc[2].src = bv.reads + (size_t)b0 * 3u * rw; pc[2].bytes = nb * 3u * rw * 8u; pc[2].off = off; off += 32u * 3u * rw * 8u;
    const bool hsane = h1 >= h0 && h1 - h0 < (1ull << 24), psane = p1 >= p0 && p1 - p0 < (1ull << 24);
    const uint32_t nh = hsane ? (uint32_t)(h1 - h0) : 0u, np = psane ? (uint32_t)(p1 - p0) : 0u;
    const bool st_par = psane && off + np * 16u <= (uint32_t)TB;
    pc[3].src = bv.partner + p0; pc[3].bytes = st_par ? np * 16u : 0u; pc[3].off = off; if (st_par) off += np * 16u;
    const bool st_hits = hsane && off + nh * 20u <= (uint32_t)TB;                // 16-byte record + 4-byte owner word per hit
    pc[4].src = bv.hits + h0; pc[4].bytes = st_hits ? nh * 16u : 0u; pc[4].off = off;
    const uint32_t off_owner = off + nh * 16u;
    stage_tile<5>(sm, pc, bar, parity, lane);
    // virtual array bases: absolute indices keep working whether the records sit in shared memory or in HBM
    const thb_hit* H = st_hits ? reinterpret_cast<const thb_hit*>(sm + pc[4].off) - h0 : bv.hits;
    const thb_hit* PH = st_par ? reinterpret_cast<const thb_hit*>(sm + pc[3].off) - p0 : bv.partner;
    uint32_t* OW = st_hits ? reinterpret_cast<uint32_t*>(sm + off_owner) - h0 : owner_g - bv.hit_base;
    const uint4* s_hdr = reinterpret_cast<const uint4*>(sm);
    const uint16_t* s_cnt = reinterpret_cast<const uint16_t*>(sm + pc[1].off);
    const uint64_t* s_rd = reinterpret_cast<const uint64_t*>(sm + pc[2].off);

    // ================= phase A: thread = bundle =================
    const bool act = lane < nb;
    const uint32_t bi = b0 + lane;
    BundleView B; B.flags = 0; B.nsegs = 0; B.n_partner = 0; B.read_len = 0; B.partner_index = 0; B.partner = PH;
    #pragma unroll
    for (int s = 0; s < NSMAX; ++s) { B.seg_n[s] = 0; B.seg_ptr[s] = H; }
    uint32_t my_total = 0, my_begin = 0;
    if (act) {
      const uint4 hdr = s_hdr[lane];
      B.nsegs = (int)nsg; my_begin = hdr.y;
      uint32_t o = hdr.y;
      #pragma unroll
      for (int s = 0; s < NSMAX; ++s) {
        int c = 0; if (s < B.nsegs) c = (int)s_cnt[lane * nsg + s];
        B.seg_n[s] = c; B.seg_ptr[s] = H + o; o += (uint32_t)c;
      }
      my_total = o - hdr.y;
      B.partner_index = hdr.z; B.n_partner = (int)(hdr.w & 0xffffu);
      B.read_len = (int)((hdr.w >> 16) & 0xffu); B.flags = hdr.w >> 24;
      B.rw = (int)rw; B.rd = s_rd + (size_t)lane * 3u * rw;
      // a partner group outside the tile's range (groups shared between bundles) is read where it lies
      const bool pin = B.partner_index >= p0 && B.partner_index + (uint64_t)B.n_partner <= p1;
      B.partner = ((st_par && pin) ? PH : bv.partner) + B.partner_index;
    }
    // back-to-back check of the hit ranges
    {
      uint32_t incl = my_total;
      #pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d); if ((int)lane >= d) incl += v; }
      const bool bad = act && (uint64_t)my_begin != h0 + (uint64_t)(incl - my_total);
      const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
      const bool bad_end = !hsane || h0 + (uint64_t)total != h1;
      if (__any_sync(0xffffffffu, bad) || bad_end) { if (lane == 0) atomicOr(out.err, 16u); __syncwarp(); continue; }
    }
    // owner map: which bundle of the tile / which segment a hit belongs to
    if (act) {
      #pragma unroll
      for (int s = 0; s < NSMAX; ++s)
        for (int k = 0; k < B.seg_n[s]; ++k) OW[(B.seg_ptr[s] - H) + k] = (lane << 4) | (uint32_t)s;
    }
    // find_insertions_and_deletions: segment pairs (i, i+1), i < nsegs-2, up to the first empty segment (2856-2871)
    int indel_pairs = 0;
    if (act && (B.flags & THB_BUNDLE_INDELS) && B.nsegs > 1) {
      bool stop = false;
      #pragma unroll
      for (int i = 0; i + 2 < NSMAX; ++i)
        if (!stop && i + 2 < B.nsegs) {
          if (B.seg_n[i] == 0 || B.seg_n[i + 1] == 0 || i * P.seglen > B.read_len) stop = true; else indel_pairs = i + 1;
        }
    }
    const bool gaps = act && (B.flags & THB_BUNDLE_GAPS) && B.nsegs > 0;
    int last = B.nsegs - 1;                                       // find_gaps 3304-3313
    #pragma unroll
    for (int s = NSMAX - 1; s > 0; --s) if (s == last && B.seg_n[s] == 0) --last;
    // last == 0: the host never schedules find_gaps for a read with only segment-0 hits (3981)
    bool check_partner = gaps && last > 0;                        // 3361-3390
    if (check_partner) {
      for (int i = 0; i < B.seg_n[0] && check_partner; ++i) {
        const Hit l = load_hit(B.seg_ptr[0] + i);
        for (int j = 0; j < B.seg_n[last]; ++j) {
          const Hit r = load_hit(B.seg_ptr[last] + j);
          if (l.ref_id == r.ref_id && l.anti == r.anti) {
            const int dist = l.anti ? l.left - r.right : r.left - l.right;
            if (dist >= P.min_intron && dist < P.max_intron) { check_partner = false; break; }
          }
        }
      }
    }
    __syncwarp();
    const bool rescued = gaps && last > 0 && check_partner && B.n_partner > 0;    // 3392: the mate re-anchors the read
    // find_fusions re-anchors too, under its own pair rule (3123-3142); whether it gets that far is decided in
    // fusion_enum_kernel, here every partner hit it could ask for is queued (a superset; the search is a pure function)
    const bool fus = act && P.fusion_search && (B.flags & THB_BUNDLE_FUSIONS) && B.n_partner > 0 && B.seg_n[0] > 0;
    uint64_t need = 0; int n_need = 0;
    auto partner_needed = [&](int r) -> bool {
      const Hit rightHit = load_hit(B.partner + r);
      if (rescue_geom(ref, P, rightHit, B.read_len).status != RG_COMPUTE) return false;
      const int minus_dist = -P.max_ins * 2;
      bool any = false;
      for (int l = 0; l < B.seg_n[0] && !any; ++l) {
        const Hit leftHit = load_hit(B.seg_ptr[0] + l);
        const bool opposite = leftHit.ref_id == rightHit.ref_id && leftHit.anti != rightHit.anti;
        if (rescued) any = opposite;                                                 // 3412
        if (fus && !any) {
          const int dist = leftHit.anti ? leftHit.left - rightHit.right : rightHit.left - leftHit.right;
          any = !(opposite && dist > minus_dist && dist <= P.fusion_min_dist);       // 3132-3142
        }
      }
      return any;
    };
    if (rescued || fus)
      for (int r = 0; r < B.n_partner; ++r)
        if (partner_needed(r)) {
          if (r < 64) { need |= 1ull << r; ++n_need; }
          else {                                                                     // rare: reserve from here
            const unsigned long long slot = agg_slot(q.counts + 2);
            q.rescue[slot] = make_uint2(bi, (unsigned)(B.partner_index + (uint64_t)r - bv.partner_base));
          }
        }
    __syncwarp();
    {
      // warp-wide reservation: rescued-bundle list and rescue task queue
      const unsigned mres = __ballot_sync(0xffffffffu, rescued);
      unsigned long long rb0 = 0;
      if (mres) { if (lane == (unsigned)(__ffs((int)mres) - 1)) rb0 = atomicAdd(q.counts + 3, (unsigned long long)__popc(mres)); }
      unsigned incl = (unsigned)n_need;
      #pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, incl, d); if ((int)lane >= d) incl += v; }
      const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
      unsigned long long t0 = 0;
      if (total && lane == 0) t0 = atomicAdd(q.counts + 2, (unsigned long long)total);
      if (mres) rb0 = __shfl_sync(0xffffffffu, rb0, __ffs((int)mres) - 1);
      if (total) t0 = __shfl_sync(0xffffffffu, t0, 0);
      if (rescued) q.rbundle[rb0 + (unsigned long long)__popc(mres & ((1u << lane) - 1u))] = bi;
      unsigned long long slot = t0 + (unsigned long long)(incl - (unsigned)n_need);
      while (need) {
        const int r = __ffsll((long long)need) - 1; need &= need - 1;
        q.rescue[slot++] = make_uint2(bi, (unsigned)(B.partner_index + (uint64_t)r - bv.partner_base));
      }
    }
    {
      bool do_windows = gaps && last > 0 && !rescued;
      if (do_windows && P.bowtie2) {                              // 3499-3506
        #pragma unroll
        for (int s = 0; s < NSMAX; ++s) if (s <= last && B.seg_n[s] > P.max_multihits) do_windows = false;
      }
      // last(4) | indel_pairs(4) << 4 | do_windows << 8 | right_mate << 9 | re-anchored by the mate << 10
      const uint32_t st = (uint32_t)(last & 15) | ((uint32_t)indel_pairs << 4) | ((uint32_t)do_windows << 8) |
                          (((B.flags & THB_BUNDLE_RIGHT_MATE) ? 1u : 0u) << 9) | ((rescued ? 1u : 0u) << 10);
      s_state[lane] = act ? st : 0u;
      if (act) bstate[bi] = st;                                   // rescued_windows_kernel / fusion_enum_kernel read it
    }
    __syncwarp();

    // ================= phase B: thread = hit =================
    for (uint32_t hb = 0; hb < nh; hb += 32u) {
      const uint64_t hi = h0 + hb + lane;
      bool want_indel = false, want_win = false, right_mate = false;
      uint32_t bl = 0, o1 = 0, o2 = 0, hrel = 0; int s = 0, n1 = 0, n2 = 0, read_len = 0, lastb = 0;
      Hit bh; bh.ref_id = 0; bh.left = bh.right = 0; bh.read_len = bh.edit = bh.anti = bh.end = 0;
      if (hb + lane < nh) {
        const uint32_t ow = OW[hi];
        bl = ow >> 4; s = (int)(ow & 15u);
        const uint32_t st = s_state[bl];
        lastb = (int)(st & 15u); const int ip = (int)((st >> 4) & 15u);
        right_mate = (st >> 9) & 1u;
        want_indel = s < ip;
        want_win = ((st >> 8) & 1u) && s < lastb;                 // hits of the last segment never open a window (3513)
        if (want_indel || want_win) {
          const uint4 hdr = s_hdr[bl];
          read_len = (int)((hdr.w >> 16) & 0xffu);
          uint32_t o = hdr.y;
          #pragma unroll
          for (int k = 0; k < NSMAX; ++k) {
            const int c = k < (int)nsg ? (int)s_cnt[bl * nsg + k] : 0;
            if (k == s + 1) { o1 = o; n1 = c; }
            if (k == s + 2) { o2 = o; n2 = c; }
            o += (uint32_t)c;
          }
          bh = load_hit(H + hi);
          hrel = (uint32_t)(hi - hdr.y);                          // position of bh among the hits of its bundle
        }
      }
      const uint64_t* rd = s_rd + (size_t)bl * 3u * rw; const int rwi = (int)rw;
      __syncwarp();
      // ---- find_insertions_and_deletions 2856-2938, left hit = bh
      if (want_indel) {
        const int start = s * P.seglen;
        int plen = 2 * P.seglen; if (start + plen > read_len) plen = read_len - start;
        bool have_slice = false; P3 full, rc;
        for (int ri = 0; ri < n1; ++ri) {
          const Hit rh0 = load_hit(H + o1 + ri);
          if (bh.ref_id != rh0.ref_id) continue;                    // 2901
          if (bh.anti != rh0.anti) continue;                        // 2904
          const bool sw = bh.anti != 0;                             // 2914-2920
          const Hit& L = sw ? rh0 : bh; const Hit& R = sw ? bh : rh0;
          const int disc = (R.right - L.left) - plen;               // 2922-2923
          const bool is_del = disc > 0 && disc <= P.max_del, is_ins = disc < 0 && disc >= -P.max_ins;
          if (!is_del && !is_ins) continue;
          if (!ref_has_seq(ref, L.ref_id)) continue;
          if (L.left < 0) continue;                                 // 2574 / 2491
          const int64_t len = (int64_t)__ldg(ref.contig_len + L.ref_id - 1);
          if (is_del) {
            if (R.right < plen) continue;                           // 2578
            if ((int64_t)L.left + plen > len || (int64_t)R.right > len) continue;   // reference reads past its buffer
          } else {
            if ((int64_t)R.right > len || plen + disc <= 0) continue;
          }
          if (!have_slice) { full = read_slice(rd, rwi, start, plen); rc = revcomp(full, plen); have_slice = true; }   // 2882-2884
          const int adj = ((int)L.read_len + (int)R.read_len >= plen) ? -1 : 0;     // 2527-2529 / 2616-2618
          const uint64_t cs = __ldg(ref.contig_start + L.ref_id - 1);
          // insertion priority: bundle position, then the reference's enumeration order (segment pair, left hit, right hit)
          const uint64_t order = ((bv.order_base + b0 + bl) << ORDER_SHIFT) | ((uint64_t)min(hrel, 4095u) << 12) | (uint64_t)min(ri, 4095);
          push_indel(q, cs + (uint64_t)L.left, plen, is_del ? 0 : 1, is_del ? disc : -disc, (int)L.edit + (int)R.edit + adj, order, sw ? rc : full);
          ++n_indel;
        }
      }
      __syncwarp();
      // ---- adjacent / distant partners in the next two segments (3521-3570)
      bool found = false; int ndrs = 0, nrrs = 0;
      if (want_win) {
        for (int r = 0; r < n1; ++r) {
          const Hit rh = load_hit(H + o1 + r);
          if (bh.anti != rh.anti || bh.ref_id != rh.ref_id) continue;
          if ((bh.anti && rh.right == bh.left) || (!bh.anti && bh.right == rh.left)) { found = true; break; }
          const int dist = bh.anti ? bh.left - rh.right : rh.left - bh.right;
          if (dist >= P.min_intron && dist < P.max_intron) ++ndrs;
        }
        if (!found && s < lastb - 1) {
          for (int r = 0; r < n2; ++r) {
            const Hit rrh = load_hit(H + o2 + r);
            if (bh.anti != rrh.anti || bh.ref_id != rrh.ref_id) continue;
            const int dist = bh.anti ? bh.left - rrh.right : rrh.left - bh.right;
            if (dist >= P.min_intron + P.seglen && dist < P.max_intron + P.seglen) ++nrrs;
          }
        }
      }
      const bool use_rr = nrrs > 0;                                 // 3577
      const int start = (s + 1) * P.seglen - look_bp;               // 3582/3584
      const bool emit = want_win && !found && (ndrs > 0 || nrrs > 0) && start <= read_len && start >= 0;   // 3572
      __syncwarp();
      // ---- window tasks (3572-3617)
      if (emit) {
        int L = use_rr ? P.seglen + 2 * look_bp : 2 * look_bp;
        if (start + L > read_len) L = read_len - start;
        P3 sup = read_slice(rd, rwi, start, L);
        if (bh.anti) sup = revcomp(sup, L);                         // 3599
        const uint32_t od = use_rr ? o2 : o1; const int nd = use_rr ? n2 : n1;
        const int lo = use_rr ? P.min_intron + P.seglen : P.min_intron;
        const int hi2 = use_rr ? P.max_intron + P.seglen : P.max_intron;
        for (int r = 0; r < nd; ++r) {
          const Hit d = load_hit(H + od + r);
          if (bh.anti != d.anti || bh.ref_id != d.ref_id) continue;
          const int dist = bh.anti ? bh.left - d.right : d.left - bh.right;
          if (!(dist >= lo && dist < hi2)) continue;
          int64_t wl, wr;
          if (!bh.anti) { wl = (int64_t)bh.right - look_bp; if (wl < 0) wl = 0; wr = (int64_t)d.left + look_bp; }   // 3587-3593
          else          { wl = (int64_t)d.right - look_bp; wr = (int64_t)bh.left + look_bp; }                       // 3594-3605
          emit_window(ref, P, q, bh.ref_id, bh.anti != 0, right_mate, wl, wr, sup, L, n_windows);                    // 3618-3649
        }
      }
      __syncwarp();
    }
    __syncwarp();             // every lane is done with the tile before the next one overwrites it
  }
  warp_add(out.counters + 0, n_windows);
  warp_add(out.counters + 1, n_indel);
}

}  // namespace thb
