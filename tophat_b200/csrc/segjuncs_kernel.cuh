// segjuncs_kernel.cuh -- the segment_juncs per-read arithmetic as one sm_100a kernel.
//
// One thread owns one thb_bundle (= one read's `hits_for_read` vector) and runs, in the order the
// host encoded in bundle.flags, the device forms of
//   find_insertions_and_deletions  segment_juncs.cpp:2807-2942  (+2470-2541, 2554-2627, 2390-2456)
//   find_gaps                      segment_juncs.cpp:3293-3650  (+2946-2973 map_read_to_contig)
//   juncs_from_ref_segs<RecordSegmentJuncs> x {GT-AG, GC-AG, AT-AC}   2051-2377, 1669-1696
// All sequence comparisons are bit-plane XOR/popcount (bitplanes.cuh); the reference's whole-window
// copy (2157-2159, 85% of its CPU time) is replaced by two <=64-base fetches at the window ends,
// which are the only bases the algorithm ever reads.  Results go into device-resident sets: a
// 64-bit open-addressing hash set per record kind (std::set semantics, order restored by a final
// sort) and an append buffer for insertions (first-inserted-wins needs the processing order).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "../../include/tophat_b200.h"
#include "bitplanes.cuh"

namespace thb {

// ---------------------------------------------------------------------------------------------
constexpr uint64_t HS_EMPTY = ~0ull;
constexpr int      MAX_SEGS = 16;
constexpr int      RES_MAX  = 48;       // rescued mate-anchor hits kept per read (guard is 40)

// key layout of junction / deletion records: [ gl1 : 39 | len : 24 | antisense : 1 ]
//   gl1 = global coordinate of left+1 (first base inside the gap), len = right - left
constexpr int KEY_LEN_BITS = 24;
__host__ __device__ __forceinline__ uint64_t make_key(uint64_t gl1, uint32_t len, uint32_t anti)
{ return (gl1 << (KEY_LEN_BITS + 1)) | ((uint64_t)len << 1) | anti; }

struct HashSet {
  uint64_t* slots; uint64_t mask; unsigned int* overflow;
};

__device__ __forceinline__ uint64_t mix64(uint64_t x)
{ x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x; }

__device__ __forceinline__ void hs_insert(const HashSet& hs, uint64_t key)
{
  uint64_t h = mix64(key) & hs.mask;
  for (int probe = 0; probe < 256; ++probe) {
    uint64_t cur = __ldcg(hs.slots + h);
    if (cur == key) return;
    if (cur == HS_EMPTY) {
      cur = atomicCAS((unsigned long long*)(hs.slots + h), (unsigned long long)HS_EMPTY, (unsigned long long)key);
      if (cur == HS_EMPTY || cur == key) return;
    }
    h = (h + 1) & hs.mask;
  }
  atomicExch(hs.overflow, 1u);
}

struct InsRec { uint64_t key; uint64_t order; uint64_t seq; uint64_t pad; };   // key = gl1<<8 | len

struct SegParams {
  int seglen, segmm, min_intron, max_intron, max_ins, max_del, max_multihits;
  int inner_mean, inner_sd, bowtie2, library_type;
};

struct BatchView {
  const thb_bundle* bundles; const uint16_t* seg_count; const uint64_t* reads;
  const thb_hit* hits; const thb_hit* partner;
  uint32_t n_bundles, n_segs, read_words; uint64_t order_base;
};

struct SegOutputs {
  HashSet juncs, dels;
  InsRec* ins; unsigned long long* ins_count; unsigned long long ins_cap;
  unsigned long long* counters;   // [0] windows [1] indel tasks [2] rescue tasks [3] junction emits [4] hits read
  unsigned int* err;              // bit0: insertion buffer overflow, bit1: rescue overflow (bowtie1)
};

struct Hit { uint32_t ref_id; int32_t left, right; uint32_t read_len, edit, anti, end; };

__device__ __forceinline__ Hit load_hit(const thb_hit* p)
{
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
  Hit h; h.ref_id = v.x; h.left = (int32_t)v.y; h.right = (int32_t)v.z;
  h.read_len = v.w & 0xff; h.edit = (v.w >> 8) & 0xff;
  h.anti = (v.w >> 16) & THB_HIT_ANTISENSE ? 1u : 0u; h.end = (v.w >> 16) & THB_HIT_END ? 1u : 0u;
  return h;
}

struct Stats { unsigned windows, indel_tasks, rescue_tasks, emits; };

__device__ __forceinline__ bool ref_has_seq(const RefView& r, uint32_t id)
{ return id >= 1 && id <= r.n_contigs && __ldg(r.contig_len + id - 1) > 0; }

// ---------------------------------------------------------------------------------------------
// simpleSplitAlignment (2390-2456) on mismatch bit masks: MA bit j = leftRef[j] mismatches,
// MB bit j = rightRef[j] mismatches.  Returns the FIRST argmin over p in [1, n) of
// before[p] + after[p-1] (callers at 2516 / 2604 use element [0]); *min_err gets the minimum.
__device__ __forceinline__ int split_first_argmin(uint64_t MA, uint64_t MB, int n, int* min_err)
{
  int best = -1, bestv = n + 1;
  int a = 0;                                   // after[p-1]  = popc(MA & mask(p))
  int b = __popcll(MB);                        // before[p]   = popc(MB >> p)
  for (int p = 1; p < n; ++p) {
    a += (int)((MA >> (p - 1)) & 1ull);
    b -= (int)((MB >> (p - 1)) & 1ull);
    const int e = a + b;
    if (e < bestv) { bestv = e; best = p; }
  }
  *min_err = bestv;
  return best;
}

// detect_small_deletion 2554-2627.  `rd` = the (possibly reverse-complemented) 2*seglen read slice.
__device__ __forceinline__ void detect_small_deletion(const RefView& ref, const SegOutputs& out, const P3& rd, int P,
                                                      const Hit& L, const Hit& R, Stats& st)
{
  if (!ref_has_seq(ref, L.ref_id)) return;
  if (L.left < 0) return;                                         // 2574
  if (R.right < P) return;                                        // 2578
  const int64_t len = (int64_t)__ldg(ref.contig_len + L.ref_id - 1);
  if ((int64_t)L.left + P > len || (int64_t)R.right > len) return; // reference reads past its buffer here
  const int disc = (R.right - L.left) - P;                        // 2581
  const uint64_t cs = __ldg(ref.contig_start + L.ref_id - 1);
  const P3 lg = ref_fetch3(ref, cs + (uint64_t)L.left, P);        // 2582
  const P3 rg = ref_fetch3(ref, cs + (uint64_t)(R.right - P), P); // 2583
  const uint64_t m = maskn(P);
  const uint64_t MA = ((lg.p0 ^ rd.p0) | (lg.p1 ^ rd.p1) | lg.pn | rd.pn) & m;   // 2430
  const uint64_t MB = ((rg.p0 ^ rd.p0) | (rg.p1 ^ rd.p1) | rg.pn | rd.pn) & m;   // 2417
  int min_err; const int p = split_first_argmin(MA, MB, P, &min_err);
  st.indel_tasks++;
  if (p < 0) return;
  const int adj = ((int)L.read_len + (int)R.read_len >= P) ? -1 : 0;            // 2616-2618
  if (min_err <= (int)L.edit + (int)R.edit + adj) {                             // 2619
    // Deletion(ref, L.left + p - 1, L.left + p + disc): gl1 = left+1, len = disc + 1
    hs_insert(out.dels, make_key(cs + (uint64_t)(L.left + p), (uint32_t)(disc + 1), 0u));
  }
}

// detect_small_insertion 2470-2541
__device__ __forceinline__ void detect_small_insertion(const RefView& ref, const SegOutputs& out, const P3& rd, int P,
                                                       const Hit& L, const Hit& R, uint64_t order, unsigned& n_ins,
                                                       Stats& st)
{
  if (!ref_has_seq(ref, L.ref_id)) return;
  if (L.left < 0) return;                                         // 2491
  const int d = P - (R.right - L.left);                           // 2498, 1..max_ins
  const int G = P - d;
  const int64_t len = (int64_t)__ldg(ref.contig_len + L.ref_id - 1);
  if ((int64_t)R.right > len || G <= 0) return;
  const uint64_t cs = __ldg(ref.contig_start + L.ref_id - 1);
  const P2 gen = ref_fetch2(ref, cs + (uint64_t)L.left, G);       // 2499: DnaString, N -> A
  const uint64_t m = maskn(G);
  // left_read = rd[0, G), right_read = rd[P-G, P)  (2506-2507)
  const uint64_t MA = ((gen.p0 ^ rd.p0) | (gen.p1 ^ rd.p1) | rd.pn) & m;
  const uint64_t MB = ((gen.p0 ^ (rd.p0 >> d)) | (gen.p1 ^ (rd.p1 >> d)) | (rd.pn >> d)) & m;
  int min_err; const int p = split_first_argmin(MA, MB, G, &min_err);
  st.indel_tasks++;
  if (p < 0) return;                                              // 2513
  const int adj = ((int)L.read_len + (int)R.read_len >= P) ? -1 : 0;            // 2527-2529
  if (min_err <= (int)L.edit + (int)R.edit + adj && p + d <= G) {               // 2530-2531
    uint64_t seq = 0;                                             // left_read[p, p+d), 3 bits per base
    for (int k = 0; k < d; ++k) {
      const int j = p + k;
      const uint64_t c = ((rd.pn >> j) & 1ull) ? 4ull : (((rd.p0 >> j) & 1ull) | (((rd.p1 >> j) & 1ull) << 1));
      seq |= c << (3 * k);
    }
    const unsigned long long slot = atomicAdd(out.ins_count, 1ull);
    if (slot < out.ins_cap) {
      InsRec r; r.key = ((cs + (uint64_t)(L.left + p)) << 8) | (uint64_t)d;     // Insertion(ref, L.left+p-1, seq)
      r.order = (order << 12) | (uint64_t)(n_ins & 0xfffu); r.seq = seq; r.pad = 0;
      out.ins[slot] = r;
    } else atomicOr(out.err, 1u);
    n_ins++;
  }
}

// ---------------------------------------------------------------------------------------------
// map_read_to_contig 2946-2973: leftmost position in [0, clen - rl) with the fewest (<3) mismatches.
// Dna5 characters: N == N is a match there (plain char compare), anything else vs N mismatches.
__device__ __forceinline__ int map_read_to_contig(const RefView& ref, uint64_t g, int clen, const P3& q, int rl)
{
  int pos = -1, best = 3;
  const int npos = clen - rl;                                     // i < contig_len - read_len
  const uint64_t qm = maskn(rl);
  for (int c0 = 0; c0 < npos; c0 += 64 - rl + 1) {
    const int span = min(64, clen - c0);
    const P3 w = ref_fetch3(ref, g + (uint64_t)c0, span);
    const int lim = min(npos - c0, 64 - rl + 1);
    for (int j = 0; j < lim; ++j) {
      const uint64_t mm = (((w.p0 >> j) ^ q.p0) | ((w.p1 >> j) ^ q.p1) | ((w.pn >> j) ^ q.pn)) & qm;
      const int t = __popcll(mm);
      if (t < best) { best = t; pos = c0 + j; }
    }
    if (best == 0) break;
  }
  return pos;
}

// The mate-flank rescue of find_gaps (3406-3472).  Returns false where the reference `break`s.
__device__ __forceinline__ bool rescue_in_flank(const RefView& ref, const SegParams& P, const uint64_t* rd, int rw,
                                                int read_len, const Hit& rightHit, Hit* res, int& nres,
                                                const SegOutputs& out, Stats& st)
{
  if (!ref_has_seq(ref, rightHit.ref_id)) return true;
  const int part = P.inner_sd > P.inner_mean ? P.inner_sd - P.inner_mean : 0;   // 3425
  const int flank = P.inner_mean + P.inner_sd;                                  // 3426
  int64_t left;
  if (rightHit.anti) { if (flank <= rightHit.left) left = (int64_t)rightHit.left - flank; else return false; }
  else               { if (part <= rightHit.right) left = (int64_t)rightHit.right - part; else return false; }
  const int clen = flank + part;
  const int64_t len = (int64_t)__ldg(ref.contig_len + rightHit.ref_id - 1);
  if (clen <= 0 || left < 0 || left + clen > len) return true;    // past the contig end: undefined in the reference
  int crl = P.seglen - P.segmm - 3; if (crl > 15) crl = 15;       // 3451
  if (crl <= 0 || crl > read_len) return true;
  const uint64_t cs = __ldg(ref.contig_start + rightHit.ref_id - 1);
  const P3 fwd = read_slice(rd, rw, read_len - crl, crl);         // 3452
  const P3 rev = revcomp(fwd, crl);                               // 3453: rcRead[0, crl)
  st.rescue_tasks++;
  const int fpos = map_read_to_contig(ref, cs + (uint64_t)left, clen, fwd, crl);
  const int rpos = map_read_to_contig(ref, cs + (uint64_t)left, clen, rev, crl);
  #pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int pos = k ? rpos : fpos;
    if (pos < 0) continue;
    if (nres < RES_MAX) {
      Hit h; h.ref_id = rightHit.ref_id; h.left = (int32_t)(left + pos); h.right = h.left + crl;
      h.read_len = (uint32_t)crl; h.edit = 0; h.anti = (uint32_t)k; h.end = 1;
      res[nres] = h;
    } else if (!P.bowtie2) atomicOr(out.err, 2u);
    nres++;
  }
  return true;
}

// ---------------------------------------------------------------------------------------------
// One POINT_DIR_BOTH window, all three motifs (2097-2289 + 1686-1691).
//   sup      : support read, L bases (already reverse-complemented for antisense hits)
//   wleft/wright : RefSeg.left / RefSeg.right
__device__ __forceinline__ void scan_window(const RefView& ref, const SegParams& P, const SegOutputs& out,
                                            uint32_t ref_id, bool antisense, bool right_mate,
                                            int64_t wleft, int64_t wright, const P3& sup, int L, Stats& st)
{
  if (!ref_has_seq(ref, ref_id)) return;                          // 2104-2107
  const int64_t len = (int64_t)__ldg(ref.contig_len + ref_id - 1);
  if (wleft < 0 || wright >= len - 1) return;                     // 2154
  const int64_t S = wright - wleft;
  if (S < L + 2 || L < 2) return;
  bool skip_fwd = false, skip_rev = false;                        // 2109-2138
  if (P.library_type == 2) { if (!right_mate) { if (antisense) skip_rev = true; else skip_fwd = true; }
                             else             { if (antisense) skip_fwd = true; else skip_rev = true; } }
  if (P.library_type == 3) { if (!right_mate) { if (antisense) skip_fwd = true; else skip_rev = true; }
                             else             { if (antisense) skip_rev = true; else skip_fwd = true; } }
  st.windows++;
  const uint64_t cs = __ldg(ref.contig_start + ref_id - 1);
  const P2 WL = ref_fetch2(ref, cs + (uint64_t)wleft, L + 2);               // window[0, L+2)
  const P2 WR = ref_fetch2(ref, cs + (uint64_t)(wright - (L + 2)), L + 2);  // window[S-L-2, S)
  const uint64_t mL = maskn(L);
  // left_mismatches (2193-2205): bit j = window[j] != support[j]
  const uint64_t ML = ((WL.p0 ^ sup.p0) | (WL.p1 ^ sup.p1) | sup.pn) & mL;
  // right_mismatches (2210-2219): window[j + S - L] = WR[j + 2]
  const uint64_t MR = (((WR.p0 >> 2) ^ sup.p0) | ((WR.p1 >> 2) ^ sup.p1) | sup.pn) & mL;
  // `to`: index of the third left mismatch among j <= L-2, else L-2 (2177, 2200-2204)
  int to = L - 2;
  { uint64_t m = ML & maskn(L - 1); m &= m - 1; m &= m - 1; if (m) to = __ffsll((long long)m) - 1; }
  // k: index of the third mismatch met scanning right-to-left; right_mismatches[] stays 0 below it
  // (the zero-initialised array + break at 2217-2218: SURVEY.md quirk Q0)
  int k = -1;
  { uint64_t m = MR; if (m) m &= ~(1ull << (63 - __clzll((long long)m))); if (m) m &= ~(1ull << (63 - __clzll((long long)m)));
    if (m) k = 63 - __clzll((long long)m); }
  const uint64_t lA = eq_letter(WL, 0), lC = eq_letter(WL, 1), lG = eq_letter(WL, 2), lT = eq_letter(WL, 3);
  const uint64_t rA = eq_letter(WR, 0), rC = eq_letter(WR, 1), rG = eq_letter(WR, 2), rT = eq_letter(WR, 3);
  const uint64_t range = maskn(to + 1);
  // forward strand: donor at window[i], acceptor at window[pos] = WR[i]
  //   GT-AG, GC-AG, AT-AC
  uint64_t cand_f = ((lG & (lT >> 1)) & (rA & (rG >> 1))) | ((lG & (lC >> 1)) & (rA & (rG >> 1))) |
                    ((lA & (lT >> 1)) & (rA & (rC >> 1)));
  // reverse strand: rc(acceptor) at window[i], rc(donor) at WR[i]:  CT-AC, CT-GC, GT-AT
  uint64_t cand_r = ((lC & (lT >> 1)) & (rA & (rC >> 1))) | ((lC & (lT >> 1)) & (rG & (rC >> 1))) |
                    ((lG & (lT >> 1)) & (rA & (rT >> 1)));
  cand_f = skip_fwd ? 0ull : (cand_f & range);
  cand_r = skip_rev ? 0ull : (cand_r & range);
  uint64_t cand = cand_f | cand_r;
  while (cand) {
    const int i = __ffsll((long long)cand) - 1;
    cand &= cand - 1;
    const int lm = __popcll(ML & maskn(i));                        // left_mismatches[i-1]
    const int rm = (i > k) ? __popcll(MR >> i) : (i == k ? 3 : 0);  // right_mismatches[i]
    if (lm + rm <= 2) {                                            // 2265
      // Junction(ref, wleft + i - 1, wleft + pos + 2): gl1 = wleft + i, len = S - L + 1
      const uint64_t gl1 = cs + (uint64_t)(wleft + i);
      const uint32_t jl = (uint32_t)(S - L + 1);
      if ((cand_f >> i) & 1ull) { hs_insert(out.juncs, make_key(gl1, jl, 0u)); st.emits++; }
      if ((cand_r >> i) & 1ull) { hs_insert(out.juncs, make_key(gl1, jl, 1u)); st.emits++; }
    }
  }
}

// ---------------------------------------------------------------------------------------------
struct BundleView {
  const thb_hit* seg_ptr[MAX_SEGS]; int seg_n[MAX_SEGS]; int nsegs;
  const thb_hit* partner; int n_partner;
  const uint64_t* rd; int rw; int read_len;
};

// find_insertions_and_deletions 2807-2942
__device__ __forceinline__ void find_indels(const RefView& ref, const SegParams& P, const SegOutputs& out,
                                            const BundleView& B, uint64_t order, Stats& st)
{
  if (B.nsegs <= 1) return;                                       // 2812-2818
  unsigned n_ins = 0;
  for (int i = 0; i + 2 < B.nsegs; ++i) {                         // 2856
    if (B.seg_n[i] == 0 || B.seg_n[i + 1] == 0) return;           // 2870-2871
    const int start = i * P.seglen;
    if (start > B.read_len) return;
    int plen = 2 * P.seglen; if (start + plen > B.read_len) plen = B.read_len - start;
    const P3 full = read_slice(B.rd, B.rw, start, plen);          // 2882
    const P3 rc = revcomp(full, plen);                            // 2883-2884
    for (int li = 0; li < B.seg_n[i]; ++li) {
      const Hit lh0 = load_hit(B.seg_ptr[i] + li);
      for (int ri = 0; ri < B.seg_n[i + 1]; ++ri) {
        const Hit rh0 = load_hit(B.seg_ptr[i + 1] + ri);
        if (lh0.ref_id != rh0.ref_id) continue;                   // 2901
        if (lh0.anti != rh0.anti) continue;                       // 2904
        const bool sw = lh0.anti != 0;                            // 2914-2920
        const Hit& L = sw ? rh0 : lh0; const Hit& R = sw ? lh0 : rh0;
        const P3& mod = sw ? rc : full;
        const int disc = (R.right - L.left) - plen;               // 2922-2923
        if (disc > 0 && disc <= P.max_del) detect_small_deletion(ref, out, mod, plen, L, R, st);
        if (disc < 0 && disc >= -P.max_ins) detect_small_insertion(ref, out, mod, plen, L, R, order, n_ins, st);
      }
    }
  }
}

// find_gaps 3293-3650
__device__ __forceinline__ void find_gaps(const RefView& ref, const SegParams& P, const SegOutputs& out,
                                          const BundleView& B, bool right_mate, Stats& st)
{
  if (B.nsegs <= 0) return;
  int last = B.nsegs - 1;
  while (last > 0 && B.seg_n[last] == 0) --last;                  // 3304-3311
  const int n = last + 1;                                         // 3313
  if (last == 0) return;   // host never schedules find_gaps for a read with only segment-0 hits (3981)
  const bool has_partner = B.n_partner > 0;                       // 3322-3344 (host lookup)
  bool check_partner = true;                                      // 3361-3390
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
  Hit res[RES_MAX]; int nres = 0;
  const bool rescued = check_partner && has_partner;              // 3392
  if (rescued) {
    // segments 1..last are cleared (3395-3398); rescued hits become segment `last` (3461, 3471)
    for (int l = 0; l < B.seg_n[0]; ++l) {
      const Hit leftHit = load_hit(B.seg_ptr[0] + l);
      for (int r = 0; r < B.n_partner; ++r) {
        const Hit rightHit = load_hit(B.partner + r);
        if (leftHit.ref_id != rightHit.ref_id || leftHit.anti == rightHit.anti) continue;   // 3412
        // 3421 can never be true
        if (!rescue_in_flank(ref, P, B.rd, B.rw, B.read_len, rightHit, res, nres, out, st)) break;
      }
    }
  }
  if (nres > RES_MAX && !P.bowtie2) nres = RES_MAX;               // flagged through out.err bit 1
  auto count = [&](int s) -> int { return rescued ? (s == 0 ? B.seg_n[0] : (s == last ? nres : 0)) : B.seg_n[s]; };
  if (P.bowtie2) for (int s = 0; s < n; ++s) if (count(s) > P.max_multihits) return;   // 3499-3506
  auto get = [&](int s, int h) -> Hit { return (rescued && s == last) ? res[h] : load_hit(B.seg_ptr[s] + h); };
  const int look_bp = 8;
  for (int s = 0; s < n - 1; ++s) {          // hits of the last segment never open a window (3513)
    const int ns = count(s);
    for (int h = 0; h < ns; ++h) {
      const Hit bh = get(s, h);
      bool found = false; int ndrs = 0, nrrs = 0;
      { const int nr = count(s + 1);                              // 3521-3548
        for (int r = 0; r < nr; ++r) {
          const Hit rh = get(s + 1, r);
          if (bh.anti != rh.anti || bh.ref_id != rh.ref_id) continue;
          if ((bh.anti && rh.right == bh.left) || (!bh.anti && bh.right == rh.left)) { found = true; break; }
          const int dist = bh.anti ? bh.left - rh.right : rh.left - bh.right;
          if (dist >= P.min_intron && dist < P.max_intron) ++ndrs;
        } }
      if (found) continue;
      if (s < n - 2) {                                            // 3550-3570
        const int nr = count(s + 2);
        for (int r = 0; r < nr; ++r) {
          const Hit rrh = get(s + 2, r);
          if (bh.anti != rrh.anti || bh.ref_id != rrh.ref_id) continue;
          const int dist = bh.anti ? bh.left - rrh.right : rrh.left - bh.right;
          if (dist >= P.min_intron + P.seglen && dist < P.max_intron + P.seglen) ++nrrs;
        }
      }
      if (ndrs == 0 && nrrs == 0) continue;                       // 3572
      const bool use_rr = nrrs > 0;                               // 3577
      const int start = (s + 1) * P.seglen - look_bp;             // 3582/3584
      if (start > B.read_len || start < 0) continue;
      int L = use_rr ? P.seglen + 2 * look_bp : 2 * look_bp;
      if (start + L > B.read_len) L = B.read_len - start;
      P3 sup = read_slice(B.rd, B.rw, start, L);
      if (bh.anti) sup = revcomp(sup, L);                         // 3599
      const int sd = use_rr ? s + 2 : s + 1;
      const int nd = count(sd);
      const int lo = use_rr ? P.min_intron + P.seglen : P.min_intron;
      const int hi = use_rr ? P.max_intron + P.seglen : P.max_intron;
      for (int r = 0; r < nd; ++r) {
        const Hit d = get(sd, r);
        if (bh.anti != d.anti || bh.ref_id != d.ref_id) continue;
        const int dist = bh.anti ? bh.left - d.right : d.left - bh.right;
        if (!(dist >= lo && dist < hi)) continue;
        int64_t wl, wr;
        if (!bh.anti) { wl = (int64_t)bh.right - look_bp; if (wl < 0) wl = 0; wr = (int64_t)d.left + look_bp; }   // 3587-3593
        else          { wl = (int64_t)d.right - look_bp; wr = (int64_t)bh.left + look_bp; }                       // 3594-3605
        scan_window(ref, P, out, bh.ref_id, bh.anti != 0, right_mate, wl, wr, sup, L, st);                        // 3618-3649
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
segjuncs_kernel(RefView ref, SegParams P, BatchView bv, SegOutputs out)
{
  Stats st = {0u, 0u, 0u, 0u};
  for (uint32_t bi = blockIdx.x * blockDim.x + threadIdx.x; bi < bv.n_bundles; bi += gridDim.x * blockDim.x) {
    const uint4 hdr = __ldg(reinterpret_cast<const uint4*>(bv.bundles + bi));
    const uint32_t hit_begin = hdr.y, partner_begin = hdr.z;
    const uint32_t n_partner = hdr.w & 0xffffu, read_len = (hdr.w >> 16) & 0xffu, flags = hdr.w >> 24;
    BundleView B;
    B.nsegs = (int)bv.n_segs;
    uint32_t off = hit_begin;
    for (int s = 0; s < B.nsegs; ++s) {
      const int c = (int)__ldg(bv.seg_count + (size_t)bi * bv.n_segs + s);
      B.seg_ptr[s] = bv.hits + off; B.seg_n[s] = c; off += (uint32_t)c;
    }
    B.partner = bv.partner + partner_begin; B.n_partner = (int)n_partner;
    B.rw = (int)bv.read_words; B.rd = bv.reads + (size_t)bi * 3 * bv.read_words; B.read_len = (int)read_len;
    if (flags & THB_BUNDLE_INDELS) find_indels(ref, P, out, B, bv.order_base + bi, st);
    if (flags & THB_BUNDLE_GAPS) find_gaps(ref, P, out, B, (flags & THB_BUNDLE_RIGHT_MATE) != 0, st);
  }
  // block-level reduction of the task counters -> 4 atomics per warp
  unsigned v[4] = { st.windows, st.indel_tasks, st.rescue_tasks, st.emits };
  #pragma unroll
  for (int k = 0; k < 4; ++k) {
    unsigned x = v[k];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0 && x) atomicAdd(out.counters + k, (unsigned long long)x);
  }
}

// ---------------------------------------------------------------------------------------------
// set maintenance kernels

__global__ void hs_clear_kernel(uint64_t* slots, uint64_t n)
{
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) slots[i] = HS_EMPTY;
}
__global__ void hs_rehash_kernel(const uint64_t* old_slots, uint64_t n_old, HashSet dst)
{
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_old; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t k = old_slots[i];
    if (k != HS_EMPTY) hs_insert(dst, k);
  }
}
// insert an explicit key list (used after the multi-GPU all-gather)
__global__ void hs_insert_list_kernel(const uint64_t* keys, uint64_t n, HashSet dst)
{
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    if (keys[i] != HS_EMPTY) hs_insert(dst, keys[i]);
}
// compaction: warp-aggregated append of the occupied slots
__global__ void hs_compact_kernel(const uint64_t* slots, uint64_t n, uint64_t* out, unsigned long long* out_count)
{
  for (uint64_t base = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) & ~31ull; base < n; base += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t i = base + (threadIdx.x & 31);
    const uint64_t k = i < n ? slots[i] : HS_EMPTY;
    const unsigned m = __ballot_sync(0xffffffffu, k != HS_EMPTY);
    if (!m) continue;
    unsigned long long b = 0;
    if ((threadIdx.x & 31) == 0) b = atomicAdd(out_count, (unsigned long long)__popc(m));
    b = __shfl_sync(0xffffffffu, b, 0);
    if (k != HS_EMPTY) out[b + __popc(m & ((1u << (threadIdx.x & 31)) - 1u))] = k;
  }
}
// sorted keys -> thb_junction records (binary search of the contig table)
__global__ void decode_keys_kernel(const uint64_t* keys, uint64_t n, RefView ref, thb_junction* out)
{
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t k = keys[i];
    const uint64_t gl1 = k >> (KEY_LEN_BITS + 1);
    const uint32_t len = (uint32_t)((k >> 1) & ((1u << KEY_LEN_BITS) - 1u));
    int lo = 0, hi = (int)ref.n_contigs - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (ref.contig_start[mid] <= gl1) lo = mid; else hi = mid - 1; }
    thb_junction j; j.ref_id = (uint32_t)lo + 1u;
    j.left = (uint32_t)(gl1 - ref.contig_start[lo]) - 1u; j.right = j.left + len; j.antisense = (uint32_t)(k & 1ull);
    out[i] = j;
  }
}

}  // namespace thb
