// segjuncs_kernel.cuh -- the segment_juncs per-read arithmetic as five sm_100a kernels.
//
// The reference runs, per read, find_insertions_and_deletions (segment_juncs.cpp:2807-2942), find_gaps
// (3293-3650) and, inside it, juncs_from_ref_segs<RecordSegmentJuncs> x {GT-AG, GC-AG, AT-AC} (2051-2377).
// A thread-per-read kernel doing all of that diverges badly (ncu, profiles/r1a: 4 of 32 lanes active), so the
// work is split by *task type*; every kernel then runs one uniform code path over a dense task array:
//
//   K1a bundle_kernel          thread = bundle   segment bookkeeping, "must the mate anchor this read" (3361-3392),
//                                                 rescue tasks, hit -> bundle owner map
//   K1b hit_kernel             thread = hit      this hit against the next one / two segments: indel + window tasks
//   K2 rescue_kernel           thread = (bundle, partner hit)   map_read_to_contig fwd + rc (2946-2973), bit-sliced
//   K3 rescued_windows_kernel  thread = rescued bundle          multihit guard + window tasks from the rescued hits
//   K4 window_scan_kernel      thread = window                  the 3-motif donor/acceptor scan (2097-2289, 1686-1691)
//   K5 indel_kernel            thread = hit pair                simpleSplitAlignment (2390-2456) + accept rules
//
// All sequence comparisons are bit-plane XOR/popcount (bitplanes.cuh); the reference's whole-window copy
// (2157-2159, 85% of its CPU time) is replaced by two <=64-base fetches at the window ends, the only bases the
// algorithm reads.  Results go into device-resident sets: a 64-bit open-addressing hash set per record kind
// (std::set semantics, order restored by a final sort) and an append buffer for insertions (first-inserted-wins
// needs the processing order, carried as an explicit priority).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "../../include/tophat_b200.h"
#include "bitplanes.cuh"

namespace thb {

// ---------------------------------------------------------------------------------------------
constexpr uint64_t HS_EMPTY = ~0ull;
constexpr int HS_MAX_PROBES = 24;
constexpr int      MAX_SEGS = 16;
constexpr int      RES_MAX  = 64;       // rescued mate-anchor hits kept per read (bowtie2 guard is 40)

// key layout of junction / deletion records: [ gl1 : 39 | len : 24 | antisense : 1 ]
//   gl1 = global coordinate of left+1 (first base inside the gap), len = right - left
constexpr int KEY_LEN_BITS = 24;
__host__ __device__ __forceinline__ uint64_t make_key(uint64_t gl1, uint32_t len, uint32_t anti)
{ return (gl1 << (KEY_LEN_BITS + 1)) | ((uint64_t)len << 1) | anti; }

struct HashSet {
  uint64_t* slots; uint64_t mask; unsigned int* overflow;
};

// slot hash: two rounds of a 32-bit multiplicative mix over the folded key (64-bit multiplies cost 4 IMADs each)
__device__ __forceinline__ uint64_t mix64(uint64_t x)
{
  uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
  uint32_t h = (lo ^ (hi * 0x9E3779B1u)) * 0x85EBCA6Bu;
  h ^= h >> 15; h = (h ^ hi) * 0xC2B2AE35u; h ^= h >> 16;
  return ((uint64_t)(h * 0x27D4EB2Fu) << 32) | h;
}

__device__ __forceinline__ void hs_insert(const HashSet& hs, uint64_t key)
{
  uint64_t h = mix64(key) & hs.mask;
  // a probe run this long means the table is too full to be fast (linear probing: ~0.5 load): report overflow, the host doubles the
  // table and repeats the scan of the batch -- the larger table stays for the following batches and passes
  for (int probe = 0; probe < HS_MAX_PROBES; ++probe) {
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
constexpr int ORDER_SHIFT = 24;         // order = (global bundle position) << 24 | left hit index in bundle << 12 | right hit index

struct SegParams {
  int seglen, segmm, min_intron, max_intron, max_ins, max_del, max_multihits;
  int inner_mean, inner_sd, bowtie2, library_type;
  int fusion_search, fusion_min_dist;
};

struct BatchView {
  const thb_bundle* bundles; const uint16_t* seg_count; const uint64_t* reads;
  const thb_hit* hits; const thb_hit* partner;     // indexed by the absolute hit_begin / partner_begin of a bundle
  uint32_t n_bundles, n_segs, read_words; uint64_t order_base;
  uint64_t partner_base;                           // absolute index of the first partner hit present in this launch
  uint64_t hit_base;                               // absolute index of the first hit present in this launch
  uint64_t hit_end, partner_end;                   // absolute indices one past the last hit / partner hit of this launch
};

// ---- task records -----------------------------------------------------------------------------
// WindowTask: w0 = gl(39) << 25 | S(24) << 1 | antisense, gl = global coordinate of window[0], S = window length;
//             p0/p1 = support read code planes; pn = N plane (low 48 bits) | L << 56 | skip_fwd << 62 | skip_rev << 63
struct __align__(16) WindowTask { uint64_t w0, p0, p1, pn; };
// IndelTask: meta = gL(39) << 25 | P(7) << 18 | kind(1) << 17 | d(10) << 7 | (thr + 1)(7)
//            gL = global coordinate of leftHit.left, P = read slice length, kind 0 = deletion (d = discrepancy),
//            1 = insertion (d = inserted length), thr = accepted error count (clamped)
struct IndelTask { uint64_t meta, order, p0, p1, pn; };

struct Queues {
  WindowTask* win; unsigned long long cap_win;
  IndelTask* indel; unsigned long long cap_indel;
  uint2* rescue;                 // (bundle index in launch, absolute partner hit index); cap = partner hits in launch
  int2* rescue_out;              // [partner hits in launch] lefts of the fwd / rc rescue hit, INT_MIN = none
  uint32_t* rbundle;             // rescued bundles; cap = bundles in launch
  unsigned long long* counts;    // [0] windows [1] indel tasks [2] rescue tasks [3] rescued bundles
  unsigned int* overflow;        // bit0 window queue, bit1 indel queue
};

struct SegOutputs {
  HashSet juncs, dels;
  InsRec* ins; unsigned long long* ins_count; unsigned long long ins_cap;
  unsigned long long* counters;   // [0] windows [1] indel tasks [2] rescue tasks [3] junction emits
  unsigned int* err;              // bit0: insertion buffer overflow, bit1: rescue overflow (bowtie1)
};

struct Hit { uint32_t ref_id; int32_t left, right; uint32_t read_len, edit, anti, end; };

__device__ __forceinline__ Hit load_hit(const thb_hit* p)
{
  const uint4 v = *reinterpret_cast<const uint4*>(p);      // generic: hits may be staged in shared memory
  Hit h; h.ref_id = v.x; h.left = (int32_t)v.y; h.right = (int32_t)v.z;
  h.read_len = v.w & 0xff; h.edit = (v.w >> 8) & 0xff;
  h.anti = (v.w >> 16) & THB_HIT_ANTISENSE ? 1u : 0u; h.end = (v.w >> 16) & THB_HIT_END ? 1u : 0u;
  return h;
}

__device__ __forceinline__ bool ref_has_seq(const RefView& r, uint32_t id)
{ return id >= 1 && id <= r.n_contigs && __ldg(r.contig_len + id - 1) > 0; }

// warp-aggregated counter bump (all lanes of the warp must call)
__device__ __forceinline__ void warp_add(unsigned long long* ctr, unsigned v)
{
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0 && v) atomicAdd(ctr, (unsigned long long)v);
}

// Queue slot for the calling lane: the lanes that are converged here share ONE atomicAdd (same-address atomics
// were 14-38% of the stall samples in profiles/r1b).
__device__ __forceinline__ unsigned long long agg_slot(unsigned long long* ctr)
{
  const unsigned m = __activemask();
  const unsigned lane = threadIdx.x & 31u;
  const int leader = __ffs((int)m) - 1;
  unsigned long long base = 0;
  if ((int)lane == leader) base = atomicAdd(ctr, (unsigned long long)__popc(m));
  base = __shfl_sync(m, base, leader);
  return base + (unsigned long long)__popc(m & ((1u << lane) - 1u));
}

// ---------------------------------------------------------------------------------------------
// bundle access
struct BundleView {
  const thb_hit* seg_ptr[MAX_SEGS]; int seg_n[MAX_SEGS]; int nsegs;
  const thb_hit* partner; int n_partner; uint64_t partner_index;   // absolute index of partner[0]
  const uint64_t* rd; int rw; int read_len; uint32_t flags;
};

// `hits_base` / `partner_base` are virtual array bases: global arrays, or a shared-memory stage shifted so that the
// absolute indices of the bundle still apply.
template <int NSMAX>
__device__ __forceinline__ void bundle_header(const BatchView& bv, uint32_t bi, BundleView& B, uint32_t& hit_begin, uint32_t& hit_end)
{
  const uint4 hdr = __ldg(reinterpret_cast<const uint4*>(bv.bundles + bi));
  B.nsegs = (int)bv.n_segs;
  uint32_t off = hdr.y;
  #pragma unroll
  for (int s = 0; s < NSMAX; ++s) {
    int c = 0;
    if (s < B.nsegs) c = (int)__ldg(bv.seg_count + (size_t)bi * bv.n_segs + s);
    B.seg_n[s] = c; off += (uint32_t)c;
  }
  hit_begin = hdr.y; hit_end = off;
  B.partner_index = hdr.z; B.n_partner = (int)(hdr.w & 0xffffu);
  B.read_len = (int)((hdr.w >> 16) & 0xffu); B.flags = hdr.w >> 24;
  B.rw = (int)bv.read_words; B.rd = bv.reads + (size_t)bi * 3 * bv.read_words;
}
template <int NSMAX>
__device__ __forceinline__ void bundle_pointers(BundleView& B, const thb_hit* hits_base, uint32_t hit_begin, const thb_hit* partner_base)
{
  uint32_t off = hit_begin;
  #pragma unroll
  for (int s = 0; s < NSMAX; ++s) { B.seg_ptr[s] = hits_base + off; off += (uint32_t)B.seg_n[s]; }
  B.partner = partner_base + B.partner_index;
}
template <int NSMAX>
__device__ __forceinline__ void load_bundle(const BatchView& bv, uint32_t bi, BundleView& B)
{
  uint32_t hb, he; bundle_header<NSMAX>(bv, bi, B, hb, he);
  bundle_pointers<NSMAX>(B, bv.hits, hb, bv.partner);
}

// ---------------------------------------------------------------------------------------------
// task emission

// One POINT_DIR_BOTH window (find_gaps 3572-3605 -> juncs_from_ref_segs 2097-2177): pre-checks + queue push.
__device__ __forceinline__ void emit_window(const RefView& ref, const SegParams& P, const Queues& q, uint32_t ref_id,
                                            bool antisense, bool right_mate, int64_t wleft, int64_t wright,
                                            const P3& sup, int L, unsigned& n_windows)
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
  n_windows++;
  const unsigned long long slot = agg_slot(q.counts + 0);
  if (slot >= q.cap_win) { atomicOr(q.overflow, 1u); return; }
  const uint64_t cs = __ldg(ref.contig_start + ref_id - 1);
  WindowTask t;
  t.w0 = ((cs + (uint64_t)wleft) << 25) | ((uint64_t)S << 1) | (antisense ? 1ull : 0ull);
  t.p0 = sup.p0; t.p1 = sup.p1;
  t.pn = sup.pn | ((uint64_t)L << 56) | ((uint64_t)skip_fwd << 62) | ((uint64_t)skip_rev << 63);
  *reinterpret_cast<uint4*>(&q.win[slot]) = make_uint4((unsigned)t.w0, (unsigned)(t.w0 >> 32), (unsigned)t.p0, (unsigned)(t.p0 >> 32));
  *(reinterpret_cast<uint4*>(&q.win[slot]) + 1) = make_uint4((unsigned)t.p1, (unsigned)(t.p1 >> 32), (unsigned)t.pn, (unsigned)(t.pn >> 32));
}

__device__ __forceinline__ void push_indel(const Queues& q, uint64_t gL, int P, int kind, int d, int thr, uint64_t order,
                                           const P3& rd)
{
  const unsigned long long slot = agg_slot(q.counts + 1);
  if (slot >= q.cap_indel) { atomicOr(q.overflow, 2u); return; }
  int t1 = thr + 1; if (t1 < 0) t1 = 0; if (t1 > 127) t1 = 127;
  IndelTask t;
  t.meta = (gL << 25) | ((uint64_t)P << 18) | ((uint64_t)kind << 17) | ((uint64_t)d << 7) | (uint64_t)t1;
  t.order = order; t.p0 = rd.p0; t.p1 = rd.p1; t.pn = rd.pn;
  q.indel[slot] = t;
}

// ---------------------------------------------------------------------------------------------
// K1 is split in two so that every thread runs one short, uniform piece of the per-read logic:
//   K1a bundle_kernel  thread = bundle : segment bookkeeping, multihit guard, "does the mate have to anchor this read"
//                                        (3361-3392), rescue task emission, per-hit owner map
//   K1b hit_kernel     thread = hit    : this hit against the hits of the next one / two segments -- indel tasks
//                                        (2856-2938) and window tasks (3508-3617)
// A thread-per-bundle enumeration walks three nested data-dependent loops; lanes drift apart and are effectively
// serialised (profiles/r1b: 8 of 32 lanes active, 1.0-1.2 ms per 5.3 M bundles).

// per-bundle state written by K1a: last(4) | indel_pairs(4) << 4 | do_windows << 8 | right_mate << 9 | re-anchored by the mate << 10
struct HitOwner { uint32_t v; };     // bundle index in launch << 4 | segment

// Geometry of the mate-flank rescue for one partner hit (find_gaps 3421-3451).
enum { RG_NOSEQ = 0, RG_BREAK = 1, RG_SKIP = 2, RG_COMPUTE = 3 };
struct RescueGeom { int status; int64_t left; int clen, crl; };
__device__ __forceinline__ RescueGeom rescue_geom(const RefView& ref, const SegParams& P, const Hit& rightHit, int read_len)
{
  RescueGeom g; g.status = RG_NOSEQ; g.left = 0; g.clen = 0; g.crl = 0;
  if (!ref_has_seq(ref, rightHit.ref_id)) return g;
  const int part = P.inner_sd > P.inner_mean ? P.inner_sd - P.inner_mean : 0;   // 3425
  const int flank = P.inner_mean + P.inner_sd;                                  // 3426
  g.status = RG_BREAK;
  if (rightHit.anti) { if (flank <= rightHit.left) g.left = (int64_t)rightHit.left - flank; else return g; }
  else               { if (part <= rightHit.right) g.left = (int64_t)rightHit.right - part; else return g; }
  g.status = RG_SKIP;
  g.clen = flank + part;
  const int64_t len = (int64_t)__ldg(ref.contig_len + rightHit.ref_id - 1);
  if (g.clen <= 0 || g.left < 0 || g.left + g.clen > len) return g;   // past the contig end: undefined in the reference
  int crl = P.seglen - P.segmm - 3; if (crl > 15) crl = 15;           // 3451
  g.crl = crl;
  if (crl <= 0 || crl > read_len) return g;
  g.status = RG_COMPUTE;
  return g;
}


template <int NSMAX>
__global__ void __launch_bounds__(256)
bundle_kernel(RefView ref, SegParams P, BatchView bv, Queues q, uint32_t* __restrict__ bstate, uint32_t* __restrict__ owner)
{
  const unsigned lane = threadIdx.x & 31u;
  for (uint32_t base = blockIdx.x * blockDim.x + threadIdx.x - lane; base < bv.n_bundles; base += gridDim.x * blockDim.x) {
    const uint32_t bi = base + lane;
    const bool act = bi < bv.n_bundles;
    BundleView B; B.flags = 0; B.nsegs = 0; B.n_partner = 0;
    if (act) load_bundle<NSMAX>(bv, bi, B);
    __syncwarp();
    // owner map: which bundle / segment a hit belongs to
    if (act) {
      #pragma unroll
      for (int s = 0; s < NSMAX; ++s)
        for (int k = 0; k < B.seg_n[s]; ++k) owner[(B.seg_ptr[s] - bv.hits) + k - bv.hit_base] = (bi << 4) | (uint32_t)s;
    }
    __syncwarp();
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
    // last == 0: host never schedules find_gaps for a read with only segment-0 hits (3981)
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
    // Which partner hits need the flank search: decided per lane first (bit r of `need` for the first 64 partner hits), so
    // that queue space is reserved ONCE per warp at a converged point -- a returning atomic inside this data-dependent loop
    // stalled the diverged lanes for an L2 round trip per task (43% of this kernel's stall samples in profiles/r1u).
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
    __syncwarp();
    if (act) {
      bool do_windows = gaps && last > 0 && !rescued;
      if (do_windows && P.bowtie2) {                              // 3499-3506
        #pragma unroll
        for (int s = 0; s < NSMAX; ++s) if (s <= last && B.seg_n[s] > P.max_multihits) do_windows = false;
      }
      bstate[bi] = (uint32_t)last | ((uint32_t)indel_pairs << 4) | ((uint32_t)do_windows << 8) |
                   (((B.flags & THB_BUNDLE_RIGHT_MATE) ? 1u : 0u) << 9) | ((rescued ? 1u : 0u) << 10);
    }
    __syncwarp();
  }
}

// K1b: one hit `bh` of segment s against segments s+1 / s+2 of its read.  The body is cut into phases separated by
// __syncwarp(): lanes that skip a phase wait at the barrier instead of running ahead, so the expensive phases (read
// slice + task emission) execute once per warp with every qualifying lane converged (profiles/r1e: 1.7 lanes before).
template <int NSMAX>
__global__ void __launch_bounds__(256, 4)
hit_kernel(RefView ref, SegParams P, BatchView bv, Queues q, const uint32_t* __restrict__ bstate, const uint32_t* __restrict__ owner,
           uint64_t n_hits, SegOutputs out)
{
  unsigned n_windows = 0, n_indel = 0;
  const unsigned lane = threadIdx.x & 31u;
  const int look_bp = 8;
  for (uint64_t base = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x - lane; base < n_hits; base += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t hi = base + lane;
    // ---- phase 0: who am I
    bool want_indel = false, want_win = false, right_mate = false;
    uint32_t bi = 0, o1 = 0, o2 = 0, hrel = 0; int s = 0, n1 = 0, n2 = 0, read_len = 0, last = 0;
    Hit bh; bh.ref_id = 0; bh.left = bh.right = 0; bh.read_len = bh.edit = bh.anti = bh.end = 0;
    if (hi < n_hits) {
      const uint32_t ow = __ldg(owner + hi);
      bi = ow >> 4; s = (int)(ow & 15u);
      const uint32_t st = __ldg(bstate + bi);
      last = (int)(st & 15u); const int indel_pairs = (int)((st >> 4) & 15u);
      right_mate = (st >> 9) & 1u;
      want_indel = s < indel_pairs;
      want_win = ((st >> 8) & 1u) && s < last;                    // hits of the last segment never open a window (3513)
      if (want_indel || want_win) {
        const uint4 hdr = __ldg(reinterpret_cast<const uint4*>(bv.bundles + bi));
        read_len = (int)((hdr.w >> 16) & 0xffu);
        const int nsegs = (int)bv.n_segs;
        uint32_t off = hdr.y;
        #pragma unroll
        for (int k = 0; k < NSMAX; ++k) {
          const int c = k < nsegs ? (int)__ldg(bv.seg_count + (size_t)bi * bv.n_segs + k) : 0;
          if (k == s + 1) { o1 = off; n1 = c; }
          if (k == s + 2) { o2 = off; n2 = c; }
          off += (uint32_t)c;
        }
        bh = load_hit(bv.hits + bv.hit_base + hi);
        hrel = (uint32_t)(bv.hit_base + hi - hdr.y);             // position of bh among the hits of its bundle
      }
    }
    const uint64_t* rd = bv.reads + (size_t)bi * 3 * bv.read_words; const int rw = (int)bv.read_words;
    __syncwarp();
    // ---- phase 1: find_insertions_and_deletions 2856-2938, left hit = bh
    if (want_indel) {
      const int start = s * P.seglen;
      int plen = 2 * P.seglen; if (start + plen > read_len) plen = read_len - start;
      bool have_slice = false; P3 full, rc;
      for (int ri = 0; ri < n1; ++ri) {
        const Hit rh0 = load_hit(bv.hits + o1 + ri);
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
        if (!have_slice) { full = read_slice(rd, rw, start, plen); rc = revcomp(full, plen); have_slice = true; }   // 2882-2884
        const int adj = ((int)L.read_len + (int)R.read_len >= plen) ? -1 : 0;     // 2527-2529 / 2616-2618
        const uint64_t cs = __ldg(ref.contig_start + L.ref_id - 1);
        // insertion priority: bundle position, then the reference's enumeration order (segment pair, left hit, right hit)
        const uint64_t order = ((bv.order_base + bi) << ORDER_SHIFT) | ((uint64_t)min(hrel, 4095u) << 12) | (uint64_t)min(ri, 4095);
        push_indel(q, cs + (uint64_t)L.left, plen, is_del ? 0 : 1, is_del ? disc : -disc, (int)L.edit + (int)R.edit + adj, order, sw ? rc : full);
        ++n_indel;
      }
    }
    __syncwarp();
    // ---- phase 2: adjacent / distant partners in the next two segments (3521-3570)
    bool found = false; int ndrs = 0, nrrs = 0;
    if (want_win) {
      for (int r = 0; r < n1; ++r) {
        const Hit rh = load_hit(bv.hits + o1 + r);
        if (bh.anti != rh.anti || bh.ref_id != rh.ref_id) continue;
        if ((bh.anti && rh.right == bh.left) || (!bh.anti && bh.right == rh.left)) { found = true; break; }
        const int dist = bh.anti ? bh.left - rh.right : rh.left - bh.right;
        if (dist >= P.min_intron && dist < P.max_intron) ++ndrs;
      }
      if (!found && s < last - 1) {
        for (int r = 0; r < n2; ++r) {
          const Hit rrh = load_hit(bv.hits + o2 + r);
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
    // ---- phase 3: window tasks (3572-3617)
    if (emit) {
      int L = use_rr ? P.seglen + 2 * look_bp : 2 * look_bp;
      if (start + L > read_len) L = read_len - start;
      P3 sup = read_slice(rd, rw, start, L);
      if (bh.anti) sup = revcomp(sup, L);                         // 3599
      const uint32_t od = use_rr ? o2 : o1; const int nd = use_rr ? n2 : n1;
      const int lo = use_rr ? P.min_intron + P.seglen : P.min_intron;
      const int hi2 = use_rr ? P.max_intron + P.seglen : P.max_intron;
      for (int r = 0; r < nd; ++r) {
        const Hit d = load_hit(bv.hits + od + r);
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
  warp_add(out.counters + 0, n_windows);
  warp_add(out.counters + 1, n_indel);
}

// ---------------------------------------------------------------------------------------------
// K2: map_read_to_contig (2946-2973) for the read's last `crl` bases, forward and reverse-complemented, against
// the mate's flank.  Bit-sliced: for 64 candidate positions at once, mismatch indicators of read base k are
// (window >> k) XOR broadcast(read[k]), accumulated in a saturating 2-bit counter per position, so a position's
// count is exact for 0..2 and "3+" otherwise -- all the reference needs (it starts from mismatch = 3 and only
// accepts strict improvements, i.e. the leftmost position of the minimum below 3).
// Dna5 characters: N == N is a match there (plain char compare), anything else vs N mismatches.
struct SatCount { uint64_t c0, c1, sat; };
__device__ __forceinline__ void sat_add(SatCount& c, uint64_t d)
{
  const uint64_t carry = c.c0 & d;
  c.sat |= c.c1 & carry; c.c1 ^= carry; c.c0 ^= d;
}
__device__ __forceinline__ void best_of(const SatCount& c, uint64_t valid, int base, int& best, int& pos)
{
  const uint64_t live = valid & ~c.sat;
  const uint64_t z = live & ~c.c0 & ~c.c1, o = live & c.c0 & ~c.c1, t = live & ~c.c0 & c.c1;
  if (best > 0 && z) { best = 0; pos = base + __ffsll((long long)z) - 1; }
  else if (best > 1 && o) { best = 1; pos = base + __ffsll((long long)o) - 1; }
  else if (best > 2 && t) { best = 2; pos = base + __ffsll((long long)t) - 1; }
}

__global__ void __launch_bounds__(256)
rescue_kernel(RefView ref, SegParams P, BatchView bv, Queues q)
{
  const unsigned long long n = q.counts[2];
  const unsigned lane = threadIdx.x & 31u;
  for (unsigned long long base = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x - lane; base < n; base += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long ti = base + lane;
    if (ti < n) {
    const uint2 task = q.rescue[ti];
    const uint4 hdr = __ldg(reinterpret_cast<const uint4*>(bv.bundles + task.x));
    const int read_len = (int)((hdr.w >> 16) & 0xffu);
    const Hit rightHit = load_hit(bv.partner + bv.partner_base + task.y);
    const RescueGeom g = rescue_geom(ref, P, rightHit, read_len);       // RG_COMPUTE by construction
    const uint64_t* rd = bv.reads + (size_t)task.x * 3 * bv.read_words;
    const int rw = (int)bv.read_words, crl = g.crl;
    const P3 fwd = read_slice(rd, rw, read_len - crl, crl);         // 3452
    const P3 rev = revcomp(fwd, crl);                               // 3453: rcRead[0, crl)
    const uint64_t gbase = __ldg(ref.contig_start + rightHit.ref_id - 1) + (uint64_t)g.left;
    const int npos = g.clen - crl;                                  // i < contig_len - read_len
    int fbest = 3, fpos = -1, rbest = 3, rpos = -1;
    for (int c0 = 0; c0 < npos; c0 += 64) {
      const P3 lo = ref_fetch3(ref, gbase + (uint64_t)c0, 64);
      const P3 hi = ref_fetch3(ref, gbase + (uint64_t)c0 + 64, 16);
      SatCount cf = {0, 0, 0}, cr = {0, 0, 0};
      #pragma unroll 5
      for (int k = 0; k < crl; ++k) {
        const uint64_t w0 = shr128(lo.p0, hi.p0, k), w1 = shr128(lo.p1, hi.p1, k), wn = shr128(lo.pn, hi.pn, k);
        const uint64_t f0 = 0ull - ((fwd.p0 >> k) & 1ull), f1 = 0ull - ((fwd.p1 >> k) & 1ull), fn = 0ull - ((fwd.pn >> k) & 1ull);
        const uint64_t r0 = 0ull - ((rev.p0 >> k) & 1ull), r1 = 0ull - ((rev.p1 >> k) & 1ull), rn = 0ull - ((rev.pn >> k) & 1ull);
        sat_add(cf, (w0 ^ f0) | (w1 ^ f1) | (wn ^ fn));
        sat_add(cr, (w0 ^ r0) | (w1 ^ r1) | (wn ^ rn));
      }
      const uint64_t valid = maskn(min(64, npos - c0));
      best_of(cf, valid, c0, fbest, fpos);
      best_of(cr, valid, c0, rbest, rpos);
      if (fbest == 0 && rbest == 0) break;
    }
    q.rescue_out[task.y] = make_int2(fpos >= 0 ? (int)(g.left + fpos) : INT_MIN, rpos >= 0 ? (int)(g.left + rpos) : INT_MIN);
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// K3: find_gaps for a re-anchored read (3395-3472 bookkeeping, 3499-3617 on segment 0 x rescued hits)
struct ResHit { uint32_t ref_id; int32_t left; uint32_t anti; };

template <int NSMAX>
__device__ __forceinline__ void rescued_windows_body(const RefView& ref, const SegParams& P, const BatchView& bv, const Queues& q,
                                                     const SegOutputs& out, unsigned long long ti, unsigned& n_windows, unsigned& n_rescue)
{
  const uint32_t bi = q.rbundle[ti];
  BundleView B; load_bundle<NSMAX>(bv, bi, B);
  int last = B.nsegs - 1;
  #pragma unroll
  for (int s = NSMAX - 1; s > 0; --s) if (s == last && B.seg_n[s] == 0) --last;
  // segments 1..last are cleared (3395-3398); rescued hits become segment `last` (3461, 3471)
  ResHit res[RES_MAX]; int nres = 0; int crl = 0;
  for (int l = 0; l < B.seg_n[0]; ++l) {
    const Hit leftHit = load_hit(B.seg_ptr[0] + l);
    for (int r = 0; r < B.n_partner; ++r) {
      const Hit rightHit = load_hit(B.partner + r);
      if (leftHit.ref_id != rightHit.ref_id || leftHit.anti == rightHit.anti) continue;   // 3412
      // 3421 can never be true
      const RescueGeom g = rescue_geom(ref, P, rightHit, B.read_len);
      if (g.status == RG_BREAK) break;
      if (g.status != RG_COMPUTE) continue;
      n_rescue++; crl = g.crl;
      const int2 o = q.rescue_out[B.partner_index + (uint64_t)r - bv.partner_base];
      if (o.x != INT_MIN) { if (nres < RES_MAX) { res[nres].ref_id = rightHit.ref_id; res[nres].left = o.x; res[nres].anti = 0; } nres++; }
      if (o.y != INT_MIN) { if (nres < RES_MAX) { res[nres].ref_id = rightHit.ref_id; res[nres].left = o.y; res[nres].anti = 1; } nres++; }
    }
  }
  if (P.bowtie2 && (B.seg_n[0] > P.max_multihits || nres > P.max_multihits)) return;   // 3499-3506
  if (nres > RES_MAX) { atomicOr(out.err, 2u); nres = RES_MAX; }   // reported by the host as THB_EUNSUPPORTED
  if (last > 2 || nres == 0) return;      // the rescued hits sit in segment `last`; only segment 0 has other hits
  const bool use_rr = last == 2;            // s+1 is empty, the s+2 list is used (3550-3577)
  const bool right_mate = (B.flags & THB_BUNDLE_RIGHT_MATE) != 0;
  const int look_bp = 8;
  const int lo = use_rr ? P.min_intron + P.seglen : P.min_intron;
  const int hi = use_rr ? P.max_intron + P.seglen : P.max_intron;
  for (int h = 0; h < B.seg_n[0]; ++h) {
    const Hit bh = load_hit(B.seg_ptr[0] + h);
    bool found = false; int nd = 0;
    for (int r = 0; r < nres; ++r) {
      const ResHit rh = res[r]; const int rright = rh.left + crl;
      if (bh.anti != rh.anti || bh.ref_id != rh.ref_id) continue;
      if (!use_rr && ((bh.anti && rright == bh.left) || (!bh.anti && bh.right == rh.left))) { found = true; break; }
      const int dist = bh.anti ? bh.left - rright : rh.left - bh.right;
      if (dist >= lo && dist < hi) ++nd;
    }
    if (found || nd == 0) continue;
    const int start = P.seglen - look_bp;                       // 3582/3584 with s = 0
    if (start > B.read_len || start < 0) continue;
    int L = use_rr ? P.seglen + 2 * look_bp : 2 * look_bp;
    if (start + L > B.read_len) L = B.read_len - start;
    P3 sup = read_slice(B.rd, B.rw, start, L);
    if (bh.anti) sup = revcomp(sup, L);
    for (int r = 0; r < nres; ++r) {
      const ResHit d = res[r]; const int dright = d.left + crl;
      if (bh.anti != d.anti || bh.ref_id != d.ref_id) continue;
      const int dist = bh.anti ? bh.left - dright : d.left - bh.right;
      if (!(dist >= lo && dist < hi)) continue;
      int64_t wl, wr;
      if (!bh.anti) { wl = (int64_t)bh.right - look_bp; if (wl < 0) wl = 0; wr = (int64_t)d.left + look_bp; }
      else          { wl = (int64_t)dright - look_bp; wr = (int64_t)bh.left + look_bp; }
      emit_window(ref, P, q, bh.ref_id, bh.anti != 0, right_mate, wl, wr, sup, L, n_windows);
    }
  }
}

template <int NSMAX>
__global__ void __launch_bounds__(128)
rescued_windows_kernel(RefView ref, SegParams P, BatchView bv, Queues q, SegOutputs out)
{
  unsigned n_windows = 0, n_rescue = 0;
  const unsigned long long n = q.counts[3];
  const unsigned lane = threadIdx.x & 31u;
  for (unsigned long long base = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x - lane; base < n; base += (unsigned long long)gridDim.x * blockDim.x) {
    if (base + lane < n) rescued_windows_body<NSMAX>(ref, P, bv, q, out, base + lane, n_windows, n_rescue);
    __syncwarp();
  }
  warp_add(out.counters + 0, n_windows);
  warp_add(out.counters + 2, n_rescue);
}

// ---------------------------------------------------------------------------------------------
// K4: one POINT_DIR_BOTH window, all three motifs (2097-2289 + 1686-1691).
__global__ void __launch_bounds__(256)
window_scan_kernel(RefView ref, Queues q, SegOutputs out)
{
  unsigned emits = 0;
  unsigned long long n = q.counts[0]; if (n > q.cap_win) n = q.cap_win;
  const unsigned lane = threadIdx.x & 31u;
  for (unsigned long long base = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x - lane; base < n; base += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long ti = base + lane;
    uint64_t pass_f = 0, pass_r = 0, gl = 0; uint32_t jl = 0;
    if (ti < n) {
    const uint4 a = __ldcs(reinterpret_cast<const uint4*>(q.win + ti)), b = __ldcs(reinterpret_cast<const uint4*>(q.win + ti) + 1);
    const uint64_t w0 = ((uint64_t)a.y << 32) | a.x, pnw = ((uint64_t)b.w << 32) | b.z;
    P3 sup; sup.p0 = ((uint64_t)a.w << 32) | a.z; sup.p1 = ((uint64_t)b.y << 32) | b.x; sup.pn = pnw & 0x0000ffffffffffffull;
    const int L = (int)((pnw >> 56) & 63); const bool skip_fwd = (pnw >> 62) & 1, skip_rev = (pnw >> 63) & 1;
    gl = w0 >> 25; const int64_t S = (int64_t)((w0 >> 1) & 0xffffffull);
    const P2 WL = ref_fetch2(ref, gl, L + 2);                       // window[0, L+2)
    const P2 WR = ref_fetch2(ref, gl + (uint64_t)(S - (L + 2)), L + 2);  // window[S-L-2, S)
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
    // forward strand: donor at window[i], acceptor at window[pos] = WR[i]:  GT-AG, GC-AG, AT-AC
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
      if (lm + rm <= 2) { pass_f |= cand_f & (1ull << i); pass_r |= cand_r & (1ull << i); }   // 2265
    }
    jl = (uint32_t)(S - L + 1);
    }
    __syncwarp();
    // Junction(ref, wleft + i - 1, wleft + pos + 2): gl1 = wleft + i, len = S - L + 1
    emits += (unsigned)(__popcll(pass_f) + __popcll(pass_r));
    while (pass_f | pass_r) {
      const bool rev = pass_f == 0;
      const uint64_t m = rev ? pass_r : pass_f;
      const int i = __ffsll((long long)m) - 1;
      if (rev) pass_r = m & (m - 1); else pass_f = m & (m - 1);
      hs_insert(out.juncs, make_key(gl + (uint64_t)i, jl, rev ? 1u : 0u));
    }
    __syncwarp();
  }
  warp_add(out.counters + 3, emits);
}

// ---------------------------------------------------------------------------------------------
// K5: simpleSplitAlignment (2390-2456) on mismatch bit masks: MA bit j = leftRef[j] mismatches,
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

__device__ __forceinline__ void indel_body(const RefView& ref, const Queues& q, const SegOutputs& out, unsigned long long ti)
{
  const IndelTask t = q.indel[ti];
  const uint64_t gL = t.meta >> 25; const int P = (int)((t.meta >> 18) & 127); const int kind = (int)((t.meta >> 17) & 1);
  const int d = (int)((t.meta >> 7) & 1023); const int thr = (int)(t.meta & 127) - 1;
  P3 rd; rd.p0 = t.p0; rd.p1 = t.p1; rd.pn = t.pn;
  if (kind == 0) {                                                 // detect_small_deletion 2554-2627
    const P3 lg = ref_fetch3(ref, gL, P);                          // 2582
    const P3 rg = ref_fetch3(ref, gL + (uint64_t)d, P);            // 2583: genomic [right - P, right)
    const uint64_t m = maskn(P);
    const uint64_t MA = ((lg.p0 ^ rd.p0) | (lg.p1 ^ rd.p1) | lg.pn | rd.pn) & m;   // 2430
    const uint64_t MB = ((rg.p0 ^ rd.p0) | (rg.p1 ^ rd.p1) | rg.pn | rd.pn) & m;   // 2417
    int min_err; const int p = split_first_argmin(MA, MB, P, &min_err);
    if (p < 0) return;
    if (min_err <= thr)                                            // 2619
      // Deletion(ref, L.left + p - 1, L.left + p + disc): gl1 = left+1, len = disc + 1
      hs_insert(out.dels, make_key(gL + (uint64_t)p, (uint32_t)(d + 1), 0u));
  } else {                                                         // detect_small_insertion 2470-2541
    const int G = P - d;
    const P2 gen = ref_fetch2(ref, gL, G);                         // 2499: DnaString, N -> A
    const uint64_t m = maskn(G);
    // left_read = rd[0, G), right_read = rd[P-G, P)  (2506-2507)
    const uint64_t MA = ((gen.p0 ^ rd.p0) | (gen.p1 ^ rd.p1) | rd.pn) & m;
    const uint64_t MB = ((gen.p0 ^ (rd.p0 >> d)) | (gen.p1 ^ (rd.p1 >> d)) | (rd.pn >> d)) & m;
    int min_err; const int p = split_first_argmin(MA, MB, G, &min_err);
    if (p < 0) return;                                           // 2513
    if (min_err <= thr && p + d <= G) {                            // 2530-2531
      uint64_t seq = 0;                                            // left_read[p, p+d), 3 bits per base
      for (int k = 0; k < d; ++k) {
        const int j = p + k;
        const uint64_t c = ((rd.pn >> j) & 1ull) ? 4ull : (((rd.p0 >> j) & 1ull) | (((rd.p1 >> j) & 1ull) << 1));
        seq |= c << (3 * k);
      }
      const unsigned long long slot = atomicAdd(out.ins_count, 1ull);
      if (slot < out.ins_cap) {
        InsRec r; r.key = ((gL + (uint64_t)p) << 8) | (uint64_t)d;   // Insertion(ref, L.left+p-1, seq)
        r.order = t.order; r.seq = seq; r.pad = 0;
        out.ins[slot] = r;
      } else atomicOr(out.err, 1u);
    }
  }
}

__global__ void __launch_bounds__(256)
indel_kernel(RefView ref, Queues q, SegOutputs out)
{
  unsigned long long n = q.counts[1]; if (n > q.cap_indel) n = q.cap_indel;
  const unsigned lane = threadIdx.x & 31u;
  for (unsigned long long base = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x - lane; base < n; base += (unsigned long long)gridDim.x * blockDim.x) {
    if (base + lane < n) indel_body(ref, q, out, base + lane);
    __syncwarp();
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

// deletion keys as Junction(ref, left, right, antisense = false) keys (long_spanning_reads.cpp:2916-2944)
__global__ void clear_bit0_kernel(const uint64_t* in, uint64_t n, uint64_t* out)
{
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) out[i] = in[i] & ~1ull;
}

}  // namespace thb
