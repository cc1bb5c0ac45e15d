// fusion_kernel.cuh -- segment_juncs --fusion-search: find_fusions (segment_juncs.cpp:2976-3291) and detect_fusion
// (2629-2805) as two sm_100a kernels.
//
//   K6 fusion_enum_kernel    thread = bundle   the candidate pairs of find_fusions: segment-0 hits x (hits of the last
//                                               mapped segment ++ hits the mate re-anchors in its flank), filtered by the
//                                               reference's distance / strand / edit-distance rules  -> FusionTask queue
//   K7 fusion_detect_kernel  thread = pair     detect_fusion: simpleSplitAlignment (2390-2456) of the whole read against
//                                               the two loci, every optimal break position -> FusRec append buffer
//
// The mate-flank rescue (3123-3211) is the same map_read_to_contig search find_gaps runs, so its tasks ride in
// rescue_kernel's queue (bundle_kernel emits them) and K6 only reads rescue_out.  A bundle flagged
// THB_BUNDLE_FUSIONS_LAST is seen the way find_gaps left it (4005-4033): segments 1.. cleared and the rescued hits in
// the last segment (3395-3398, 3461, 3471), which K6 re-derives from rescue_out instead of storing.
// Records are reduced by key (count, minimum edit distance: fusions.h:87-101) when the batch sequence finishes.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "../../include/tophat_b200.h"
#include "bitplanes.cuh"
#include "segjuncs_kernel.cuh"
#include "join_kernel.cuh"

namespace thb {

enum { FUSION_FF = 7, FUSION_FR = 8, FUSION_RF = 9, FUSION_RR = 10 };      // CigarOpCode values, bwt_map.h:36-55

struct FusionParams { int fusion_min_dist, fusion_anchor, n_ignore; const uint32_t* ignore; };

// one detect_fusion call: left / right hit as passed (after the swap of 3273-3282)
struct __align__(16) FusionTask { uint32_t bundle, lref; int32_t lleft, lright; uint32_t rref; int32_t rleft, rright; uint32_t meta; };
//   meta = dir(4) | use_rc(1) << 4 | total_edit(16) << 8
struct __align__(16) FusRec { uint32_t r1, r2, left, right, dir, edit, pad0, pad1; };

struct FusionQueues {
  FusionTask* tasks; unsigned long long cap; unsigned long long* count;       // count: tasks emitted by K6
  FusRec* rec; unsigned long long rec_cap; unsigned long long* rec_count;
  unsigned int* err;                     // bit2: task queue overflow, bit3: record buffer overflow
  unsigned long long* counters;          // [2] rescue calls, [7] split alignments run
};

struct RightHit { uint32_t ref_id; int32_t left, right; uint32_t edit, anti; };

__device__ __forceinline__ bool fusion_ignored(const FusionParams& F, uint32_t ref_id)
{ for (int i = 0; i < F.n_ignore; ++i) if (__ldg(F.ignore + i) == ref_id) return true; return false; }

// K6
template <int NSMAX>
__device__ void fusion_enum_body(const RefView& ref, const SegParams& P, const FusionParams& F, const BatchView& bv, const Queues& q,
                                 const FusionQueues& fq, const uint32_t* __restrict__ bstate, uint32_t bi, unsigned& n_rescue)
{
  BundleView B; load_bundle<NSMAX>(bv, bi, B);
  if (!(B.flags & THB_BUNDLE_FUSIONS) || B.nsegs <= 0) return;
  int last = B.nsegs - 1;                                           // 2987-2994
  #pragma unroll
  for (int s = NSMAX - 1; s > 0; --s) if (s == last && B.seg_n[s] == 0) --last;
  const int n0 = B.seg_n[0];
  const bool mutated = (B.flags & THB_BUNDLE_FUSIONS_LAST) && ((__ldg(bstate + bi) >> 10) & 1u);
  const int minus_dist = -P.max_ins * 2;                            // 3119

  // the hits find_gaps left in the last segment of a re-anchored read (3406-3472)
  auto for_each_gaps_rescued = [&](auto&& f) {
    for (int l = 0; l < n0; ++l) {
      const Hit leftHit = load_hit(B.seg_ptr[0] + l);
      for (int r = 0; r < B.n_partner; ++r) {
        const Hit rightHit = load_hit(B.partner + r);
        if (leftHit.ref_id != rightHit.ref_id || leftHit.anti == rightHit.anti) continue;     // 3412
        const RescueGeom g = rescue_geom(ref, P, rightHit, B.read_len);
        if (g.status == RG_BREAK) break;
        if (g.status != RG_COMPUTE) continue;
        const int2 o = q.rescue_out[B.partner_index + (uint64_t)r - bv.partner_base];
        if (o.x != INT_MIN) { RightHit h; h.ref_id = rightHit.ref_id; h.left = o.x; h.right = o.x + g.crl; h.edit = 0; h.anti = 0; f(h); }
        if (o.y != INT_MIN) { RightHit h; h.ref_id = rightHit.ref_id; h.left = o.y; h.right = o.y + g.crl; h.edit = 0; h.anti = 1; f(h); }
      }
    }
  };
  // right_segment_hits before the rescue of find_fusions itself (3074-3079)
  auto for_each_base = [&](auto&& f) {
    if (mutated) for_each_gaps_rescued(f);
    else if (last != 0)
      for (int j = 0; j < B.seg_n[last]; ++j) {
        const Hit h0 = load_hit(B.seg_ptr[last] + j);
        RightHit h; h.ref_id = h0.ref_id; h.left = h0.left; h.right = h0.right; h.edit = h0.edit; h.anti = h0.anti; f(h);
      }
  };
  bool base_empty = true;
  if (mutated) for_each_gaps_rescued([&](const RightHit&) { base_empty = false; });
  else base_empty = last == 0;
  if (mutated && base_empty) last = 0;                              // every later segment is empty now
  if (n0 == 0) return;                                              // 3034 / empty pair loops
  if (last == 0 && (load_hit(B.seg_ptr[0]).end != 0)) return;       // 3034-3036

  bool check_partner = true;                                        // 3089-3117
  if (last != 0) {
    for (int i = 0; i < n0 && check_partner; ++i) {
      const Hit l = load_hit(B.seg_ptr[0] + i);
      for_each_base([&](const RightHit& r) {
        if (!check_partner) return;
        if (l.ref_id == r.ref_id && l.anti == r.anti) {
          const int dist = l.anti ? l.left - r.right : r.left - l.right;
          if (dist > -P.max_ins && dist <= F.fusion_min_dist) check_partner = false;
        }
      });
    }
  }
  const bool do_rescue = check_partner && B.n_partner > 0;          // 3121
  // hits the mate re-anchors for find_fusions (3123-3211); `count` = the reference's rescue calls that reach the search
  auto for_each_fusion_rescued = [&](bool count, auto&& f) {
    if (!do_rescue) return;
    for (int l = 0; l < n0; ++l) {
      const Hit leftHit = load_hit(B.seg_ptr[0] + l);
      for (int r = 0; r < B.n_partner; ++r) {
        const Hit rightHit = load_hit(B.partner + r);
        if (leftHit.ref_id == rightHit.ref_id && leftHit.anti != rightHit.anti) {
          const int dist = leftHit.anti ? leftHit.left - rightHit.right : rightHit.left - leftHit.right;
          if (dist > minus_dist && dist <= F.fusion_min_dist) continue;                       // 3140
        }
        const RescueGeom g = rescue_geom(ref, P, rightHit, B.read_len);
        if (g.status == RG_BREAK) break;
        if (g.status != RG_COMPUTE) continue;
        if (count) ++n_rescue;
        const int2 o = q.rescue_out[B.partner_index + (uint64_t)r - bv.partner_base];
        if (o.x != INT_MIN) { RightHit h; h.ref_id = rightHit.ref_id; h.left = o.x; h.right = o.x + g.crl; h.edit = 0; h.anti = 0; f(h); }
        if (o.y != INT_MIN) { RightHit h; h.ref_id = rightHit.ref_id; h.left = o.y; h.right = o.y + g.crl; h.edit = 0; h.anti = 1; f(h); }
      }
    }
  };
  for_each_fusion_rescued(true, [&](const RightHit&) {});           // the rescue loop runs once, before the pair loop

  for (int li = 0; li < n0; ++li) {                                 // 3221-3290
    const Hit lh = load_hit(B.seg_ptr[0] + li);
    auto pair = [&](const RightHit& rh) {
      if (F.n_ignore && (fusion_ignored(F, lh.ref_id) || fusion_ignored(F, rh.ref_id))) return;   // 3228-3230
      if (P.bowtie2 && (int)lh.edit + (int)rh.edit > (P.segmm << 1)) return;                      // 3232-3236
      if (lh.ref_id == rh.ref_id && lh.anti == rh.anti) {                                         // 3255-3268
        const int dist = lh.anti ? lh.left - rh.right : rh.left - lh.right;
        if (dist > minus_dist && dist <= F.fusion_min_dist) return;
      }
      uint32_t dir = FUSION_FF; bool use_rc = false, swap = false;
      if (lh.anti == rh.anti) { if (lh.anti) { swap = true; use_rc = true; } }                    // 3273-3282
      else if (!lh.anti && rh.anti) dir = FUSION_FR;
      else dir = FUSION_RF;
      const unsigned long long slot = atomicAdd(fq.count, 1ull);
      if (slot >= fq.cap) { atomicOr(fq.err, 4u); return; }
      FusionTask t; t.bundle = bi;
      if (!swap) { t.lref = lh.ref_id; t.lleft = lh.left; t.lright = lh.right; t.rref = rh.ref_id; t.rleft = rh.left; t.rright = rh.right; }
      else       { t.lref = rh.ref_id; t.lleft = rh.left; t.lright = rh.right; t.rref = lh.ref_id; t.rleft = lh.left; t.rright = lh.right; }
      t.meta = dir | ((use_rc ? 1u : 0u) << 4) | ((lh.edit + rh.edit) << 8);
      fq.tasks[slot] = t;
    };
    for_each_base(pair);
    for_each_fusion_rescued(false, pair);
  }
}

template <int NSMAX>
__global__ void __launch_bounds__(128)
fusion_enum_kernel(RefView ref, SegParams P, FusionParams F, BatchView bv, Queues q, FusionQueues fq, const uint32_t* __restrict__ bstate)
{
  unsigned n_rescue = 0;
  const unsigned lane = threadIdx.x & 31u;
  for (uint32_t base = blockIdx.x * blockDim.x + threadIdx.x - lane; base < bv.n_bundles; base += gridDim.x * blockDim.x) {
    if (base + lane < bv.n_bundles) fusion_enum_body<NSMAX>(ref, P, F, bv, q, fq, bstate, base + lane, n_rescue);
    __syncwarp();
  }
  warp_add(fq.counters + 2, n_rescue);
}

// n bases of the reference starting at global coordinate g, as stride-4 planes (code planes + N plane)
__device__ __forceinline__ void ref_fetch_long(const RefView& ref, uint64_t g, int n, uint64_t* G)
{
  #pragma unroll
  for (int w = 0; w < 4; ++w) {
    const int m = min(64, n - 64 * w);
    if (m > 0) { const P3 x = ref_fetch3(ref, g + 64ull * w, m); G[w] = x.p0; G[4 + w] = x.p1; G[8 + w] = x.pn; }
    else { G[w] = 0; G[4 + w] = 0; G[8 + w] = 0; }
  }
}

// K7: detect_fusion (2629-2805)
__device__ void fusion_detect_body(const RefView& ref, const FusionParams& F, const BatchView& bv, const FusionQueues& fq,
                                   unsigned long long ti, unsigned& n_split)
{
  const FusionTask t = fq.tasks[ti];
  const uint32_t dir = t.meta & 15u; const bool use_rc = (t.meta >> 4) & 1u; const uint32_t total_edit = t.meta >> 8;
  if (!ref_has_seq(ref, t.lref) || !ref_has_seq(ref, t.rref)) return;
  const uint4 hdr = __ldg(reinterpret_cast<const uint4*>(bv.bundles + t.bundle));
  const int n = (int)((hdr.w >> 16) & 0xffu);
  const int64_t llen = (int64_t)__ldg(ref.contig_len + t.lref - 1), rlen = (int64_t)__ldg(ref.contig_len + t.rref - 1);
  int64_t lg0, rg0;
  const bool lrev = !(dir == FUSION_FF || dir == FUSION_FR), rrev = !(dir == FUSION_FF || dir == FUSION_RF);
  if (!lrev) { if ((int64_t)t.lleft + n > llen || t.lleft < 0) return; lg0 = t.lleft; }                     // 2644-2650
  else       { if (t.lright < n || (int64_t)t.lright > llen) return; lg0 = (int64_t)t.lright - n; }        // 2651-2658
  if (!rrev) { if (t.rright < n || (int64_t)t.rright > rlen) return; rg0 = (int64_t)t.rright - n; }        // 2660-2666
  else       { if ((int64_t)t.rleft + n > rlen || t.rleft < 0) return; rg0 = t.rleft; }                    // 2667-2674
  if (n < 2) return;
  uint64_t R[12], LG[12], RG[12], T[12];
  { const uint64_t* rd = bv.reads + (size_t)t.bundle * 3 * bv.read_words; const int rw = (int)bv.read_words;
    #pragma unroll
    for (int pl = 0; pl < 3; ++pl)
      #pragma unroll
      for (int w = 0; w < 4; ++w) T[pl * 4 + w] = w < rw ? __ldg(rd + pl * rw + w) : 0ull; }
  if (use_rc) revcomp_read(T, n, R);
  else {
    #pragma unroll
    for (int k = 0; k < 12; ++k) R[k] = T[k]; }
  ref_fetch_long(ref, __ldg(ref.contig_start + t.lref - 1) + (uint64_t)lg0, n, T);
  if (lrev) revcomp_read(T, n, LG);
  else {
    #pragma unroll
    for (int k = 0; k < 12; ++k) LG[k] = T[k]; }
  ref_fetch_long(ref, __ldg(ref.contig_start + t.rref - 1) + (uint64_t)rg0, n, T);
  if (rrev) revcomp_read(T, n, RG);
  else {
    #pragma unroll
    for (int k = 0; k < 12; ++k) RG[k] = T[k]; }
  // mismatch indicators (2407-2434): different, or an N on either side
  uint64_t MA[4], MB[4]; int b = 0;
  #pragma unroll
  for (int w = 0; w < 4; ++w) {
    const int rem = n - 64 * w; const uint64_t valid = rem >= 64 ? ~0ull : (rem > 0 ? ((1ull << rem) - 1ull) : 0ull);
    MA[w] = ((LG[w] ^ R[w]) | (LG[4 + w] ^ R[4 + w]) | LG[8 + w] | R[8 + w]) & valid;
    MB[w] = ((RG[w] ^ R[w]) | (RG[4 + w] ^ R[4 + w]) | RG[8 + w] | R[8 + w]) & valid;
    b += __popcll(MB[w]);
  }
  ++n_split;
  // e(p) = afterErrors[p-1] + beforeErrors[p], p in [1, n)  (2443-2454): minimum, first and last position reaching it
  int a = 0, best = n + 1, first = 0, lastp = 0;
  for (int p = 1; p < n; ++p) {
    const int w = (p - 1) >> 6, j = (p - 1) & 63;
    a += (int)((MA[w] >> j) & 1ull); b -= (int)((MB[w] >> j) & 1ull);
    const int e = a + b;
    if (e < best) { best = e; first = p; lastp = p; } else if (e == best) lastp = p;
  }
  if (best > (int)total_edit || best > 2) return;                   // 2692-2697
  if (first < F.fusion_anchor || n - lastp < F.fusion_anchor) return;   // 2699-2708: any position too close to an end
  a = 0; b = 0;
  #pragma unroll
  for (int w = 0; w < 4; ++w) b += __popcll(MB[w]);
  for (int p = 1; p <= lastp; ++p) {                                // 2710-2804
    const int w = (p - 1) >> 6, j = (p - 1) & 63;
    a += (int)((MA[w] >> j) & 1ull); b -= (int)((MB[w] >> j) & 1ull);
    if (a + b != best) continue;
    uint32_t left = !lrev ? (uint32_t)(t.lleft + p - 1) : (uint32_t)(t.lright - p);
    uint32_t right = !rrev ? (uint32_t)(t.rright - (n - p)) : (uint32_t)(t.rleft + (n - p) - 1);
    uint32_t r1 = t.lref, r2 = t.rref, tdir = dir;
    if (r2 < r1 || (r1 == r2 && left > right)) {                    // 2772-2785
      const uint32_t x = r1; r1 = r2; r2 = x; const uint32_t y = left; left = right; right = y;
      if (dir == FUSION_FF) tdir = FUSION_RR;
    }
    const unsigned long long slot = atomicAdd(fq.rec_count, 1ull);
    if (slot >= fq.rec_cap) { atomicOr(fq.err, 8u); continue; }
    uint4* dst = reinterpret_cast<uint4*>(fq.rec + slot);
    dst[0] = make_uint4(r1, r2, left, right); dst[1] = make_uint4(tdir, total_edit, 0u, 0u);
  }
}

__global__ void __launch_bounds__(128)
fusion_detect_kernel(RefView ref, FusionParams F, BatchView bv, FusionQueues fq)
{
  unsigned n_split = 0;
  unsigned long long n = *fq.count; if (n > fq.cap) n = fq.cap;
  const unsigned lane = threadIdx.x & 31u;
  for (unsigned long long base = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x - lane; base < n; base += (unsigned long long)gridDim.x * blockDim.x) {
    if (base + lane < n) fusion_detect_body(ref, F, bv, fq, base + lane, n_split);
    __syncwarp();
  }
  warp_add(fq.counters + 7, n_split);
}

}  // namespace thb
