// join_kernel.cuh -- long_spanning_reads' per-read arithmetic as two sm_100a kernels:
//   chain_enum_kernel (thread = read) enumerates segment-hit chains, chain_merge_kernel (thread = chain) merges them.
//
// Replaces, for --fusion-search off and base-space reads, the reference's
//   join_segments_for_read   long_spanning_reads.cpp:2612-2667   (multihit guard, one DFS per first-segment hit)
//   dfs_seg_hits             2222-2610   (chain enumeration, 10,000-leaf budget per first-segment hit)
//   merge_segment_chain      2101-2220   (chain orientation, valid_hit 2045-2099)
//   merge_chain              805-2038    (gap closure by junction / deletion / insertion look-up with the +-4 bp
//                                         boundary adjustment, CIGAR stitching)
//   BowtieHit::check_editdist_consistency   bwt_map.cpp:2349-2465
// The junction / insertion sets are sorted device arrays searched with the std::set bounds the reference uses.
// Sorting + de-duplicating a read's joined hits, the read-level filters and the SAM fields stay on the host.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "../../include/tophat_b200.h"
#include "bitplanes.cuh"

namespace thb {

constexpr int JMAXOPS = THB_JOINED_MAX_OPS;
constexpr int JMAXSEGS = THB_MAX_SEGS;
// CigarOpCode values (bwt_map.h:36-55); ops are packed as length << 4 | opcode
enum { OP_MATCH = 1, OP_mATCH = 2, OP_INS = 3, OP_iNS = 4, OP_DEL = 5, OP_dEL = 6, OP_REF_SKIP = 11, OP_rEF_SKIP = 12,
       OP_SOFT_CLIP = 13, OP_HARD_CLIP = 14, OP_PAD = 15 };
__host__ __device__ __forceinline__ uint32_t mkop(int code, uint32_t len) { return (len << 4) | (uint32_t)code; }
__host__ __device__ __forceinline__ int opc(uint32_t o) { return (int)(o & 15u); }
__host__ __device__ __forceinline__ uint32_t opl(uint32_t o) { return o >> 4; }

struct JoinParams {
  int max_ins, max_del, min_report_intron, max_report_intron, fusion_min_dist, max_seg_multihits, bowtie2, seglen;
};

struct JoinSets {
  const thb_junction* juncs; uint32_t n_juncs;        // Junction order (junctions.h:39-57)
  const uint32_t* jidx; uint64_t n_buckets; int shift; // jidx[b] = first junction whose global left lies in bucket >= b; bucket = 2^shift bases
                                                      // (NULL: fall back to binary search)
  const thb_insertion* ins; uint32_t n_ins;           // (refid, left, length) order (insertions.h:52-67)
  const uint32_t* iidx; uint64_t n_ibuckets; int ishift;   // the same kind of bucket index over the insertion array (NULL: binary search)
};

struct JoinBatchView {
  const thb_join_bundle* bundles; const uint16_t* seg_count; const uint64_t* reads; const thb_jhit* hits;
  const thb_jops* ops_ext;            // CIGARs of the hits without THB_JHIT_ONE_MATCH; indexed bundle.ops_begin + hit.ops_index
  uint32_t n_bundles, n_segs, read_words;
  uint32_t bundle_base;               // added to the bundle index reported in thb_joined (chunked submission)
  uint32_t hit_end;                   // index one past the last hit of the last bundle of this view
  uint32_t ops_end;                   // index one past the last thb_jops record of the last bundle of this view
};

struct JoinOut {
  thb_joined* rec; unsigned long long cap; unsigned long long* count; unsigned int* overflow;
  unsigned long long* counters;      // [0] chains merged (leaves) [1] closures looked up [2] joined hits emitted
};

// working hit of merge_chain
struct WHit {
  uint32_t ref; int left; int n; uint32_t ops[JMAXOPS];
  bool anti, asplice; uint8_t mism, smm; int seq_pos, seq_len;
  int right, rlen; bool spliced;      // cached: right(), read_len(), is_spliced() of the ops
};

__device__ __forceinline__ int cig_right(int left, const uint32_t* ops, int n)
{ int r = left; for (int i = 0; i < n; ++i) { const int c = opc(ops[i]); if (c == OP_MATCH || c == OP_REF_SKIP || c == OP_DEL) r += (int)opl(ops[i]); } return r; }
__device__ __forceinline__ int cig_read_len(const uint32_t* ops, int n)
{ int r = 0; for (int i = 0; i < n; ++i) { const int c = opc(ops[i]); if (c == OP_MATCH || c == OP_INS || c == OP_SOFT_CLIP) r += (int)opl(ops[i]); } return r; }
__device__ __forceinline__ bool cig_spliced(const uint32_t* ops, int n)
{ for (int i = 0; i < n; ++i) if (opc(ops[i]) == OP_REF_SKIP) return true; return false; }
__device__ __forceinline__ int cig_gap_length(const uint32_t* ops, int n)
{ int r = 0; for (int i = 0; i < n; ++i) { const int c = opc(ops[i]); if (c == OP_INS || c == OP_DEL) r += (int)opl(ops[i]); } return r; }

// Dna5 code of global base g: 0..3, 4 = N
__device__ __forceinline__ int ref_code5(const RefView& r, uint64_t g)
{
  const uint64_t b = g >> 6; const int j = (int)(g & 63);
  if ((__ldg(r.nmask + b) >> j) & 1ull) return 4;
  const ulonglong2 w = __ldg(r.planes + b);
  return (int)((w.x >> j) & 1ull) | ((int)((w.y >> j) & 1ull) << 1);
}
// code of base i of the oriented read R (planes with stride 4 words): 0..3, 4 = N
__device__ __forceinline__ int read_code5(const uint64_t* R, int i)
{
  const int w = i >> 6, j = i & 63;
  if ((R[8 + w] >> j) & 1ull) return 4;
  return (int)((R[w] >> j) & 1ull) | ((int)((R[4 + w] >> j) & 1ull) << 1);
}

// std::set<Junction>::lower_bound / upper_bound on the sorted array
__device__ __forceinline__ bool junc_less(const thb_junction& a, uint32_t ref, uint32_t left, uint32_t right, uint32_t anti)
{
  if (a.ref_id != ref) return a.ref_id < ref;
  if (a.left != left) return a.left < left;
  if (a.right != right) return a.right < right;
  return a.antisense < anti;
}
__device__ __forceinline__ bool junc_greater(const thb_junction& a, uint32_t ref, uint32_t left, uint32_t right, uint32_t anti)
{
  if (a.ref_id != ref) return a.ref_id > ref;
  if (a.left != left) return a.left > left;
  if (a.right != right) return a.right > right;
  return a.antisense > anti;
}
__device__ __forceinline__ uint32_t junc_lower_bound(const JoinSets& S, uint32_t ref, uint32_t left, uint32_t right, uint32_t anti)
{ uint32_t lo = 0, hi = S.n_juncs; while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (junc_less(S.juncs[mid], ref, left, right, anti)) lo = mid + 1; else hi = mid; } return lo; }
__device__ __forceinline__ uint32_t junc_upper_bound(const JoinSets& S, uint32_t ref, uint32_t left, uint32_t right, uint32_t anti)
{ uint32_t lo = 0, hi = S.n_juncs; while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (junc_greater(S.juncs[mid], ref, left, right, anti)) hi = mid; else lo = mid + 1; } return lo; }
// The same bounds through the 64-base bucket index: one table load + a scan of the (few) junctions in the bucket instead of
// two 17-step dependent binary searches (SURVEY.md 8d counts the look-up as "bucket offset + bucket").
__device__ __forceinline__ uint32_t junc_bound_idx(const JoinSets& S, uint64_t cs, int64_t clen, uint32_t ref, uint32_t left, uint32_t right,
                                                   uint32_t anti, bool upper)
{
  if (S.jidx == nullptr || (int64_t)left > clen + 32) return upper ? junc_upper_bound(S, ref, left, right, anti) : junc_lower_bound(S, ref, left, right, anti);
  const uint64_t b = (cs + (uint64_t)left) >> S.shift;
  if (b >= S.n_buckets) return upper ? junc_upper_bound(S, ref, left, right, anti) : junc_lower_bound(S, ref, left, right, anti);
  // every junction of an earlier bucket is smaller, every junction of a later one greater: the bound lies in [jidx[b], jidx[b+1]]
  // (binary search inside the bucket: indel-dense regions put tens of records into one bucket, profiles/README.md)
  uint32_t lo = __ldg(S.jidx + b), hi = __ldg(S.jidx + b + 1);
  if (upper) { while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (junc_greater(S.juncs[mid], ref, left, right, anti)) hi = mid; else lo = mid + 1; } }
  else       { while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (junc_less(S.juncs[mid], ref, left, right, anti)) lo = mid + 1; else hi = mid; } }
  return lo;
}

// std::set<Insertion>::upper_bound(Insertion(ref, left, <string of length len>))
__device__ __forceinline__ uint32_t ins_upper_bound_bs(const JoinSets& S, uint32_t ref, uint32_t left, uint32_t len)
{
  uint32_t lo = 0, hi = S.n_ins;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1; const thb_insertion& a = S.ins[mid];
    bool greater;
    if (a.ref_id != ref) greater = a.ref_id > ref; else if (a.left != left) greater = a.left > left; else greater = a.len > len;
    if (greater) hi = mid; else lo = mid + 1;
  }
  return lo;
}

// head word 3 of a thb_jhit: flags (low nibble) | n_ops << 4 | ops_index << 8 | mismatches << 16 | splice_mms << 24
// the same bound through the bucket index: one table load + a scan of the few insertions in the bucket
__device__ __forceinline__ uint32_t ins_upper_bound(const JoinSets& S, uint64_t cs, int64_t clen, uint32_t ref, uint32_t left, uint32_t len)
{
  if (S.iidx == nullptr || (int64_t)left > clen + 32) return ins_upper_bound_bs(S, ref, left, len);
  const uint64_t b = (cs + (uint64_t)left) >> S.ishift;
  if (b >= S.n_ibuckets) return ins_upper_bound_bs(S, ref, left, len);
  uint32_t lo = __ldg(S.iidx + b), hi = __ldg(S.iidx + b + 1);
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1; const thb_insertion& a = S.ins[mid];
    bool greater;
    if (a.ref_id != ref) greater = a.ref_id > ref; else if (a.left != left) greater = a.left > left; else greater = a.len > len;
    if (greater) hi = mid; else lo = mid + 1;
  }
  return lo;
}

__device__ __forceinline__ void load_whit(WHit& w, const JoinBatchView& bv, uint32_t hit, uint32_t ops_begin, int seq_pos, int seq_len)
{
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(bv.hits + hit));
  w.ref = a.x; w.left = (int)a.y; w.right = (int)a.z;
  const uint32_t fl = a.w & 0xfu; w.mism = (uint8_t)((a.w >> 16) & 0xffu); w.smm = (uint8_t)(a.w >> 24);
  w.anti = (fl & THB_HIT_ANTISENSE) != 0; w.asplice = (fl & THB_JHIT_ANTISENSE_SPLICE) != 0;
  w.seq_pos = seq_pos; w.seq_len = seq_len;
  if (fl & THB_JHIT_ONE_MATCH) {                                  // the common case: everything is in the 16-byte record
    w.n = 1; w.ops[0] = mkop(OP_MATCH, (uint32_t)(a.z - a.y)); w.rlen = (int)(a.z - a.y); w.spliced = false;
  } else {
    w.n = (int)((a.w >> 4) & 0xfu); if (w.n > THB_JHIT_MAX_OPS) w.n = THB_JHIT_MAX_OPS;
    const uint4* p = reinterpret_cast<const uint4*>(bv.ops_ext + ops_begin + ((a.w >> 8) & 0xffu));
    const uint4 b = __ldg(p); w.ops[0] = b.x; w.ops[1] = b.y; w.ops[2] = b.z; w.ops[3] = b.w;
    if (w.n > 4) { const uint4 c = __ldg(p + 1); w.ops[4] = c.x; w.ops[5] = c.y; w.ops[6] = c.z; w.ops[7] = c.w; }
    if (w.n > 8) { w.ops[8] = __ldg(reinterpret_cast<const uint32_t*>(p + 2)); }
    w.rlen = cig_read_len(w.ops, w.n); w.spliced = cig_spliced(w.ops, w.n);
  }
}

// the few fields of a segment hit the DFS needs: one 16-byte load
struct LiteHit { uint32_t ref; int left, right; bool anti; bool one_m; };   // one_m: the CIGAR is a single match op
__device__ __forceinline__ LiteHit load_lite(const thb_jhit* h)
{
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(h));
  LiteHit l; l.ref = a.x; l.left = (int)a.y; l.right = (int)a.z; l.anti = (a.w & THB_HIT_ANTISENSE) != 0; l.one_m = (a.w & THB_JHIT_ONE_MATCH) != 0;
  return l;
}

// dfs_seg_hits' pair test with --fusion-search off (2250-2558): same contig, same strand, and the genomic gap between
// the (strand-ordered) pair within [-max_insertion_length, max_report_intron_length].
__device__ __forceinline__ bool chain_compatible(const JoinParams& P, const LiteHit& prev, const LiteHit& curr, int& dist)
{
  dist = 0x7fffffff;
  if (prev.ref != curr.ref || prev.anti != curr.anti) return false;     // would need a fusion (dir != FUSION_NOTHING, 2402)
  if (prev.anti) dist = prev.left - curr.right;                         // pair swapped for antisense hits (2355-2361)
  else {
    dist = curr.left - prev.right;                                      // 2366-2379
  }
  return dist <= P.max_report_intron && dist >= -P.max_ins;             // 2554-2556
}

// valid_hit 2045-2099
__device__ __forceinline__ bool valid_cigar(const JoinParams& P, const uint32_t* ops, int n)
{
  for (int i = 1; i < n; ++i) {
    const int c = opc(ops[i]), p = opc(ops[i - 1]);
    if (c != OP_MATCH && p != OP_MATCH) return false;
    if (c == OP_INS && (int)opl(ops[i]) > P.max_ins) return false;
    if (c == OP_DEL && (int)opl(ops[i]) > P.max_del) return false;
    if (c == OP_REF_SKIP && (int64_t)opl(ops[i]) < (int64_t)P.min_report_intron) return false;
  }
  return opc(ops[0]) == OP_MATCH && opc(ops[n - 1]) == OP_MATCH;
}

// BowtieHit::check_editdist_consistency (bwt_map.cpp:2349-2465) for a single-contig hit whose sequence is R
__device__ __forceinline__ bool editdist_consistent(const RefView& ref, uint32_t ref_id, int left, const uint32_t* ops, int n,
                                                    const uint64_t* R, int read_words4, uint8_t mismatches)
{
  if (!(ref_id >= 1 && ref_id <= ref.n_contigs)) return false;
  const int64_t len = (int64_t)__ldg(ref.contig_len + ref_id - 1);
  if (len <= 0) return false;
  const uint64_t cs = __ldg(ref.contig_start + ref_id - 1);
  int64_t pos_ref = left; int pos_seq = 0; unsigned mm = 0, nmm = 0;
  for (int i = 0; i < n; ++i) {
    const int c = opc(ops[i]); const int l = (int)opl(ops[i]);
    if (c == OP_MATCH) {
      if (pos_ref < 0 || pos_ref + l > len) return false;               // the reference would read outside the contig
      for (int o = 0; o < l; o += 64) {
        const int m = min(64, l - o);
        const P3 g = ref_fetch3(ref, cs + (uint64_t)(pos_ref + o), m);
        P3 r; r.p0 = plane_slice(R, 4, pos_seq + o, m); r.p1 = plane_slice(R + 4, 4, pos_seq + o, m); r.pn = plane_slice(R + 8, 4, pos_seq + o, m);
        mm += (unsigned)__popcll((g.p0 ^ r.p0) | (g.p1 ^ r.p1) | (g.pn ^ r.pn));
        nmm += (unsigned)__popcll(g.pn & r.pn);
      }
      pos_ref += l; pos_seq += l;
    } else if (c == OP_INS) pos_seq += l;
    else if (c == OP_DEL || c == OP_REF_SKIP) pos_ref += l;
  }
  (void)read_words4;
  return mm == (unsigned)mismatches || mm + nmm == (unsigned)mismatches;
}

__device__ __forceinline__ bool cig_push(uint32_t* ops, int& n, uint32_t op)
{ if (n >= JMAXOPS) return false; ops[n++] = op; return true; }

// ---- merge_chain (805-2038), fusion_dir == FUSION_NOTHING, base space -- the pieces ----------------------------------
// The closure searches are separate functions so that chain_merge_kernel can run them as warp-converged phases.

// junction / deletion closure (1311-1591) between prev (ending at pright, last op length prml) and curr (starting at
// cleft, first op length clml).  Returns false where the reference finds no closure.
struct JuncClosure { int dtl; uint32_t glen; bool anti; int new_diff; };
__device__ __forceinline__ int junction_closure(const RefView& ref, const JoinSets& S, const uint64_t* R, uint64_t cs, int64_t clen,
                                                uint32_t ref_id, int pright, int cleft, int prml, int clml, int prev_seq_end, int curr_seq_pos,
                                                JuncClosure& out)
{
  // return: 1 found, 0 not found, -1 chain invalid (reference would read outside the contig)
  const uint32_t lbnd = (uint32_t)(pright - 4), rbnd = (uint32_t)(cleft + 4);
  uint32_t it = junc_bound_idx(S, cs, clen, ref_id, lbnd, rbnd - 8u, 1u, true);
  const uint32_t ub = junc_bound_idx(S, cs, clen, ref_id, lbnd + 8u, rbnd, 0u, false);
  int new_diff = 0xff; bool found = false;
  for (; it != ub && it < S.n_juncs; ++it) {
    const thb_junction J = S.juncs[it];
    const int dtl = (int)(J.left - (uint32_t)pright + 1u), dtr = (int)(J.right - (uint32_t)cleft);
    if (!(abs(dtl) <= 4 && abs(dtr) <= 4 && dtl == dtr)) continue;
    if (dtl > clml || -dtl > prml) continue;                  // enough matched bases on either side (1343-1348)
    // mismatches of the <= 4 shifted read bases against their new and old reference positions: Dna5 character compares
    // (N equals N), done on bit planes -- two reference fetches instead of two loads per base
    int new_mm = 0, old_mm = 0;
    auto slice_mm = [&](int read_pos, uint64_t g, int len) -> int {
      const P3 r = ref_fetch3(ref, g, len);
      const uint64_t q0 = plane_slice(R, 4, read_pos, len), q1 = plane_slice(R + 4, 4, read_pos, len), qn = plane_slice(R + 8, 4, read_pos, len);
      return __popcll((r.p0 ^ q0) | (r.p1 ^ q1) | (r.pn ^ qn));
    };
    if (dtl > 0) {
      if ((int64_t)pright + dtl > clen || (int64_t)cleft + dtl > clen) return -1;
      new_mm = slice_mm(curr_seq_pos, cs + (uint64_t)pright, dtl);
      old_mm = slice_mm(curr_seq_pos, cs + (uint64_t)cleft, dtl);
    } else if (dtl < 0) {
      const int ad = -dtl;
      if ((int64_t)J.right + ad > clen) return -1;
      new_mm = slice_mm(prev_seq_end - ad, cs + (uint64_t)J.right, ad);
      old_mm = slice_mm(prev_seq_end - ad, cs + (uint64_t)J.left + 1u, ad);
    }
    const int temp = new_mm - old_mm;
    if (temp >= new_diff || new_mm >= 2) continue;            // first strictly better candidate in Junction order (1497-1512)
    new_diff = temp; out.dtl = dtl; out.glen = J.right - J.left - 1u; out.anti = J.antisense != 0; found = true;
  }
  out.new_diff = new_diff;
  return found ? 1 : 0;
}

// insertion closure (1010-1306).  Returns 1 found (itpr / len / mismatch filled), 0 none, -1 chain invalid
struct InsClosure { int itpr; uint32_t len; int mismatch; };
__device__ __forceinline__ int insertion_closure(const RefView& ref, const JoinSets& S, const JoinParams& P, const uint64_t* R, uint64_t cs,
                                                 int64_t clen, uint32_t ref_id, int pright, int cleft, int prml, int clml, int prev_seq_end,
                                                 int curr_seq_pos, InsClosure& out)
{
  const uint32_t lbnd = (uint32_t)(pright - 4), rbnd = (uint32_t)(cleft + 4);
  uint32_t it = ins_upper_bound(S, cs, clen, ref_id, lbnd, 0u);
  const uint32_t ub = ins_upper_bound(S, cs, clen, ref_id, rbnd, (uint32_t)P.max_ins);
  bool found = false;
  for (; it != ub && it < S.n_ins; ++it) {
    const thb_insertion& I = S.ins[it];
    if ((int)I.len != pright - cleft) continue;
    const int itpr = pright - (int)I.left - 1;               // insert_to_prev_right
    const int clti = (int)I.left - cleft + 1;                // curr_left_to_insert
    if (itpr > prml || clti > clml) continue;
    int this_ref_mm = 0, ins_mm = 0; const int ilen = (int)I.len;
    auto ins_code = [&](int k) -> int { const char ch = I.seq[k]; return ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : 4; };
    if (itpr > 0) {
      // referenceSequence = ref[I.left + 1, prev.right); old/new segment = last itpr bases of prev's sequence
      const int64_t g0 = (int64_t)I.left + 1; const int s0 = prev_seq_end - itpr;
      if (g0 < 0 || g0 + itpr > clen) return -1;
      for (int ri = 0; ri < itpr; ++ri) {
        const int rc = ref_code5(ref, cs + (uint64_t)(g0 + ri)), sc = read_code5(R, s0 + ri);
        if (rc == 4 || rc != sc) ++this_ref_mm;
        if (ri < ilen) { const int ic = ins_code(ri); if (ic == 4 || ic != sc) { ++ins_mm; break; } }
        else { const int rc2 = ref_code5(ref, cs + (uint64_t)(g0 + ri - ilen)); if (rc2 == 4 || rc2 != sc) --this_ref_mm; }
      }
    }
    if (clti > 0) {
      // referenceSequence = ref[curr.left, I.left + 1); old/new segment = first clti bases of curr's sequence
      const int64_t g0 = cleft; const int s0 = curr_seq_pos;
      if (g0 < 0 || g0 + clti > clen) return -1;
      for (int ri = 0; ri < clti; ++ri) {
        const int sp = clti - ri - 1, ip = ilen - ri - 1;
        const int rc = ref_code5(ref, cs + (uint64_t)(g0 + sp)), sc = read_code5(R, s0 + sp);
        if (rc == 4 || rc != sc) ++this_ref_mm;
        if (ri < ilen) { const int ic = ins_code(ip); if (ic == 4 || ic != sc) { ++ins_mm; break; } }
        else { const int rc2 = ref_code5(ref, cs + (uint64_t)(g0 + sp + ilen)); if (rc2 == 4 || rc2 != sc) --this_ref_mm; }
      }
    }
    if (found) return -1;                                     // a second same-length candidate rejects the chain (1246-1250)
    if (ins_mm == 0) { out.mismatch = -this_ref_mm; out.itpr = itpr; out.len = I.len; found = true; }
  }
  return found ? 1 : 0;
}

__device__ __forceinline__ void emit_joined(const JoinOut& o, const thb_joined& j)
{
  const unsigned long long slot = atomicAdd(o.count, 1ull);
  if (slot >= o.cap) { atomicOr(o.overflow, 1u); return; }
  o.rec[slot] = j;
}

// Reverse complement of a whole read held as stride-4 planes (non-ACGT stays N, reads.cpp:191-207): reverse the 256-bit
// planes (bit reversal of every word, words in reverse order), shift right by 256 - n, complement the code bits.
__device__ __forceinline__ void revcomp_read(const uint64_t* F, int n, uint64_t* R)
{
  const int sh = 256 - n, ws = sh >> 6, bs = sh & 63;
  #pragma unroll
  for (int pl = 0; pl < 3; ++pl) {
    uint64_t t[5];
    #pragma unroll
    for (int w = 0; w < 4; ++w) t[w] = __brevll(F[pl * 4 + 3 - w]);
    t[4] = 0;
    #pragma unroll
    for (int w = 0; w < 4; ++w) {
      uint64_t lo = 0, hi = 0;
      #pragma unroll
      for (int k = 0; k < 5; ++k) { if (k == w + ws) lo = t[k]; if (k == w + ws + 1) hi = t[k]; }
      R[pl * 4 + w] = bs ? ((lo >> bs) | (hi << (64 - bs))) : lo;
    }
  }
  #pragma unroll
  for (int w = 0; w < 4; ++w) {
    const int rem = n - 64 * w; const uint64_t valid = rem >= 64 ? ~0ull : (rem > 0 ? ((1ull << rem) - 1ull) : 0ull);
    const uint64_t keep = valid & ~R[8 + w];
    R[w] = ~R[w] & keep; R[4 + w] = ~R[4 + w] & keep;
  }
}

// jidx[b] = number of junctions whose global left coordinate lies before bucket b (= index of the first one at or after it);
// a bucket is 2^shift bases, sized by the host so that there are a few buckets per junction
template <class Rec>
__global__ void junction_index_kernel(const Rec* juncs, uint32_t n, const uint64_t* contig_start, uint32_t* jidx, uint64_t n_buckets, int shift)
{
  for (uint64_t b = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; b <= n_buckets; b += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t g0 = b << shift;
    uint32_t lo = 0, hi = n;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; const Rec j = juncs[mid];
      if (contig_start[j.ref_id - 1] + (uint64_t)j.left < g0) lo = mid + 1; else hi = mid; }
    jidx[b] = lo;
  }
}

struct ChainQueue {
  uint32_t* tasks; unsigned long long cap; uint32_t stride;     // task = [bundle, hit index of segment 0 .. n-1] (absolute indices)
  unsigned long long* count; unsigned int* overflow;
  // chains of single-match hits that abut exactly (no closure needed): merged by chain_merge_simple_kernel
  uint32_t* simple_tasks; unsigned long long* simple_count;
  // chains whose hits all abut exactly but carry multi-op CIGARs (spliced hits): no closure search needed
  uint32_t* abut_tasks; unsigned long long* abut_count;
};

// K-J1: join_segments_for_read (2612-2667) + dfs_seg_hits (2222-2610): enumerate the compatible segment-hit chains of one
// read in the reference's DFS order, with its budget of 10,000 complete chains per first-segment hit.
// (Measured alternatives, both slower on B200 than this per-thread walk: the walk as a warp lock-step state machine, and the
// hits of the warp's 32 reads staged in shared memory first -- profiles/README.md.)
// Complete chains are NOT appended to the queues from inside the walk: a returning atomic there costs an L2 round trip per
// chain with the lanes diverged (a third of all stall samples in profiles/r1w).  Each lane parks its first ENUM_PARK chains
// in local memory; after the walk the warp is converged again and reserves queue space with one atomic per queue for all
// 32 reads.  Only a read with more chains than that falls back to reserving from inside the walk.
constexpr int ENUM_PARK = 4;
struct ParkedChains { uint16_t sel[ENUM_PARK][JMAXSEGS]; uint8_t kind[ENUM_PARK]; int n; };    // kind: 0 general, 1 simple, 2 abutting

__device__ __forceinline__ void write_chain(const ChainQueue& q, int kind, unsigned long long slot, uint32_t bi, int n, const uint32_t* off, const uint16_t* sel)
{
  if (slot >= q.cap) { atomicOr(q.overflow, 1u); return; }
  uint32_t* t = (kind == 1 ? q.simple_tasks : (kind == 2 ? q.abut_tasks : q.tasks)) + slot * q.stride;
  t[0] = bi;
  for (int s = 0; s < n; ++s) t[1 + s] = off[s] + (uint32_t)sel[s];
}

__device__ void enum_read(const JoinParams& P, const JoinBatchView& bv, const ChainQueue& q, uint32_t bi, unsigned& n_leaves,
                          ParkedChains& park, int& n_out, uint32_t* off)
{
  const uint4 hdr = __ldg(reinterpret_cast<const uint4*>(bv.bundles + bi));
  const int n = (int)((hdr.z >> 16) & 0xffu);
  if (n < 1 || n > JMAXSEGS) return;
  n_out = n;
  int cnt[JMAXSEGS];
  { uint32_t a = hdr.y;
    for (int s = 0; s < n; ++s) { cnt[s] = (int)__ldg(bv.seg_count + (size_t)bi * bv.n_segs + s); off[s] = a; a += (uint32_t)cnt[s]; } }
  if (P.bowtie2) for (int s = 0; s < n; ++s) if (cnt[s] > P.max_seg_multihits) return;      // 2624-2632
  int it[JMAXSEGS]; uint16_t sel[JMAXSEGS]; LiteHit top[JMAXSEGS]; bool simp[JMAXSEGS], abut[JMAXSEGS];
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
    // overflow of the parking area: reserve from here (lanes that are converged at this point share the atomic)
    unsigned long long* ctr = simple ? q.simple_count : (abutting ? q.abut_count : q.count);
    unsigned long long slot;
    { const unsigned m = __activemask(); const unsigned ms = __ballot_sync(m, simple); const unsigned ma = __ballot_sync(m, abutting);
      const unsigned grp = simple ? ms : (abutting ? ma : (m & ~ms & ~ma));
      const unsigned ln = threadIdx.x & 31u; const int leader = __ffs((int)grp) - 1;
      unsigned long long b0 = 0; if ((int)ln == leader) b0 = atomicAdd(ctr, (unsigned long long)__popc(grp));
      b0 = __shfl_sync(grp, b0, leader); slot = b0 + (unsigned long long)__popc(grp & ((1u << ln) - 1u)); }
    write_chain(q, kind, slot, bi, n, off, sel);
  };
  for (int i0 = 0; i0 < cnt[0]; ++i0) {
    sel[0] = (uint16_t)i0;
    int num_try = 10000;                                           // 2647
    if (n == 1) { --num_try; leaf(); continue; }
    int lvl = 1; it[1] = 0;
    top[0] = load_lite(bv.hits + off[0] + i0); simp[0] = top[0].one_m; abut[0] = true;
    while (lvl >= 1) {
      if (it[lvl] >= cnt[lvl]) { --lvl; if (lvl >= 1) ++it[lvl]; continue; }
      const LiteHit cand = load_lite(bv.hits + off[lvl] + it[lvl]);
      int dist;
      if (!chain_compatible(P, top[lvl - 1], cand, dist)) { ++it[lvl]; continue; }
      sel[lvl] = (uint16_t)it[lvl]; top[lvl] = cand; simp[lvl] = simp[lvl - 1] && cand.one_m && dist == 0; abut[lvl] = abut[lvl - 1] && dist == 0;
      if (lvl == n - 1) { --num_try; leaf(); if (num_try <= 0) break; ++it[lvl]; }
      else { ++lvl; it[lvl] = 0; }
    }
  }
}

__global__ void __launch_bounds__(256)
chain_enum_kernel(JoinParams P, JoinBatchView bv, ChainQueue q, unsigned long long* counters)
{
  unsigned n_leaves = 0;
  const unsigned lane = threadIdx.x & 31u;
  for (uint32_t base = blockIdx.x * blockDim.x + threadIdx.x - lane; base < bv.n_bundles; base += gridDim.x * blockDim.x) {
    ParkedChains park; park.n = 0; int n = 0; uint32_t off[JMAXSEGS];
    const uint32_t bi = base + lane;
    if (bi < bv.n_bundles) enum_read(P, bv, q, bi, n_leaves, park, n, off);
    __syncwarp();
    // converged: one reservation per queue for the parked chains of all 32 reads
    unsigned c[3] = {0u, 0u, 0u};
    for (int k = 0; k < park.n; ++k) c[park.kind[k]]++;
    unsigned long long slot[3];
    #pragma unroll
    for (int kd = 0; kd < 3; ++kd) {
      unsigned incl = c[kd];
      #pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, incl, d); if ((int)lane >= d) incl += v; }
      const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
      unsigned long long b0 = 0;
      if (total) {
        if (lane == 0) b0 = atomicAdd(kd == 1 ? q.simple_count : (kd == 2 ? q.abut_count : q.count), (unsigned long long)total);
        b0 = __shfl_sync(0xffffffffu, b0, 0);
      }
      slot[kd] = b0 + (unsigned long long)(incl - c[kd]);
    }
    for (int k = 0; k < park.n; ++k) {
      const int kd = park.kind[k];
      const unsigned long long sl = kd == 1 ? slot[1]++ : (kd == 2 ? slot[2]++ : slot[0]++);
      write_chain(q, kd, sl, bi, n, off, park.sel[k]);
    }
    __syncwarp();
  }
  for (int k = 16; k > 0; k >>= 1) n_leaves += __shfl_xor_sync(0xffffffffu, n_leaves, k);
  if (lane == 0 && n_leaves) atomicAdd(counters + 0, (unsigned long long)n_leaves);
}

// K-J2a: chains of single-match hits that abut exactly.  merge_chain reduces to: no closure anywhere (dist == 0, 1592),
// CIGAR = one match op over the whole read (1926-1936), mismatches = sum of the segments' (1900), then the consistency
// re-read (2023-2035); valid_hit holds trivially.  One uniform straight-line path, so these chains get their own queue.
__global__ void __launch_bounds__(256)
chain_merge_simple_kernel(RefView ref, JoinParams P, JoinBatchView bv, ChainQueue q, JoinOut o)
{
  unsigned n_emit = 0;
  unsigned long long nq = *q.simple_count; if (nq > q.cap) nq = q.cap;
  const unsigned lane = threadIdx.x & 31u;
  for (unsigned long long base = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x - lane; base < nq; base += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long ti = base + lane;
    bool ok = ti < nq;
    uint32_t bi = 0, ref0 = 0, total = 0; int left0 = 0; unsigned mism = 0, smm = 0; bool anti = false;
    if (ok) {
      const uint32_t* __restrict__ t = q.simple_tasks + ti * q.stride;
      bi = t[0];
      const uint4 hdr = __ldg(reinterpret_cast<const uint4*>(bv.bundles + bi));
      const int read_len = (int)(hdr.z & 0xffffu); const int n = (int)((hdr.z >> 16) & 0xffu);
      uint64_t R[12];
      { const uint64_t* rd = bv.reads + (size_t)bi * 3 * bv.read_words; const int rw = (int)bv.read_words;
        #pragma unroll
        for (int pl = 0; pl < 3; ++pl)
          #pragma unroll
          for (int w = 0; w < 4; ++w) R[pl * 4 + w] = w < rw ? __ldg(rd + pl * rw + w) : 0ull; }
      int minleft = 0x7fffffff;
      for (int s = 0; s < n; ++s) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(bv.hits + t[1 + s]));
        if (s == 0) { ref0 = a.x; anti = (a.w & THB_HIT_ANTISENSE) != 0; }
        minleft = min(minleft, (int)a.y);                     // chain[0] is the leftmost hit in either orientation
        mism += (a.w >> 16) & 0xffu; smm += a.w >> 24; total += a.z - a.y;
      }
      left0 = minleft;
      if (anti) { uint64_t F[12];
        #pragma unroll
        for (int k = 0; k < 12; ++k) F[k] = R[k];
        revcomp_read(F, read_len, R); }
      // new_read_len == old_read_length (2023) holds by construction
      const uint32_t op = mkop(OP_MATCH, total);
      ok = editdist_consistent(ref, ref0, left0, &op, 1, R, 4, (uint8_t)mism);
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
          uint4* dst = reinterpret_cast<uint4*>(o.rec + slot);
          const uint32_t m8 = mism & 0xffu;
          dst[0] = make_uint4(bi + bv.bundle_base, ref0, (uint32_t)left0, 1u | ((anti ? (uint32_t)THB_HIT_ANTISENSE : 0u) << 8) | (m8 << 16) | (m8 << 24));
          dst[1] = make_uint4(smm & 0xffu, mkop(OP_MATCH, total), 0u, 0u);
        }
        ++n_emit;
      }
    }
    __syncwarp();
  }
  for (int k = 16; k > 0; k >>= 1) n_emit += __shfl_xor_sync(0xffffffffu, n_emit, k);
  if (lane == 0 && n_emit) atomicAdd(o.counters + 2, (unsigned long long)n_emit);
}

// K-J2b: chains whose hits all abut exactly (gap 0 between neighbours) but carry multi-op CIGARs -- reads crossing a
// junction through a hit against the junction index.  merge_chain never searches a closure for them: every hit is
// "finalised" in turn (1888-1945: mismatch sums, splice-strand agreement, CIGAR appended with equal neighbouring ops fused),
// after the pair checks of 930-949.  That is one streaming pass over the chain's hit records, so these chains (the bulk of
// the spliced reads) get a kernel without the closure machinery of K-J2: no working copies of the hits, no phases.
template <bool DEFER, int MINB>
__global__ void __launch_bounds__(128, MINB)
chain_merge_abut_kernel(RefView ref, JoinParams P, JoinBatchView bv, ChainQueue q, JoinOut o)
{
  unsigned n_emit = 0;
  unsigned long long nq = *q.abut_count; if (nq > q.cap) nq = q.cap;
  const unsigned lane = threadIdx.x & 31u;
  // The record of a chain is written one iteration late: the atomic that reserves the warp's output slots is issued at the
  // end of iteration i and its result is first needed at the end of iteration i+1, so its L2 round trip (a quarter of the
  // stall samples of this kernel in profiles/r1w) overlaps the next chain's loads.  CIGAR buffers ping-pong between the two.
  uint32_t LCB[2][JMAXOPS]; int par = 0;
  unsigned p_em = 0; unsigned long long p_slot0 = 0; bool p_ok = false; uint4 p_hdr = make_uint4(0u, 0u, 0u, 0u); uint32_t p_smm = 0; int p_n = 0;
  auto flush = [&]() {
    if (!p_em) return;
    const unsigned long long slot0 = __shfl_sync(0xffffffffu, p_slot0, __ffs((int)p_em) - 1);
    if (p_ok) {
      const unsigned long long slot = slot0 + (unsigned long long)__popc(p_em & ((1u << lane) - 1u));
      if (slot >= o.cap) atomicOr(o.overflow, 1u);
      else {
        const uint32_t* PL = LCB[par ^ 1];
        uint4* dst = reinterpret_cast<uint4*>(o.rec + slot);
        dst[0] = p_hdr;
        auto op_at = [&](int k) -> uint32_t { return k < p_n ? PL[k] : 0u; };
        dst[1] = make_uint4(p_smm, op_at(0), op_at(1), op_at(2));
        for (int qd = 2; 4 * qd - 5 < p_n; qd += 2) {
          dst[qd] = make_uint4(op_at(4 * qd - 5), op_at(4 * qd - 4), op_at(4 * qd - 3), op_at(4 * qd - 2));
          dst[qd + 1] = make_uint4(op_at(4 * qd - 1), op_at(4 * qd), op_at(4 * qd + 1), op_at(4 * qd + 2));
        }
      }
      ++n_emit;
    }
    p_em = 0;
  };
  for (unsigned long long base = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x - lane; base < nq; base += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long ti = base + lane;
    bool ok = ti < nq;
    uint32_t bi = 0, ref0 = 0; int left0 = 0, nLC = 0, num_mm = 0, num_smm = 0; bool anti = false, saw_as = false, saw_s = false;
    uint32_t* LC = LCB[par];
    if (ok) {
      const uint32_t* __restrict__ t = q.abut_tasks + ti * q.stride;
      bi = t[0];
      const uint4 hdr = __ldg(reinterpret_cast<const uint4*>(bv.bundles + bi));
      const int read_len = (int)(hdr.z & 0xffffu); const int n = (int)((hdr.z >> 16) & 0xffu);
      uint64_t R[12];
      { const uint64_t* rd = bv.reads + (size_t)bi * 3 * bv.read_words; const int rw = (int)bv.read_words;
        #pragma unroll
        for (int pl = 0; pl < 3; ++pl)
          #pragma unroll
          for (int w = 0; w < 4; ++w) R[pl * 4 + w] = w < rw ? __ldg(rd + pl * rw + w) : 0ull; }
      anti = (__ldg(reinterpret_cast<const uint32_t*>(bv.hits + t[1]) + 3) & THB_HIT_ANTISENSE) != 0;   // chain orientation (2117-2121)
      bool prev_spliced = false, prev_asplice = false, prev_last_match = false; uint32_t prev_ref = 0;
      for (int e = 0; e < n && ok; ++e) {
        const int sg = anti ? n - 1 - e : e;
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(bv.hits + t[1 + sg]));
        const uint32_t fl = a.w & 0xfu; const bool asplice = (fl & THB_JHIT_ANTISENSE_SPLICE) != 0;
        int nops = 1; uint32_t ops[THB_JHIT_MAX_OPS]; ops[0] = mkop(OP_MATCH, (uint32_t)(a.z - a.y));
        if (!(fl & THB_JHIT_ONE_MATCH)) {
          nops = (int)((a.w >> 4) & 0xfu); if (nops > THB_JHIT_MAX_OPS) nops = THB_JHIT_MAX_OPS;
          const uint4* p = reinterpret_cast<const uint4*>(bv.ops_ext + hdr.w + ((a.w >> 8) & 0xffu));
          const uint4 b = __ldg(p); ops[0] = b.x; ops[1] = b.y; ops[2] = b.z; ops[3] = b.w;
          if (nops > 4) { const uint4 c = __ldg(p + 1); ops[4] = c.x; ops[5] = c.y; ops[6] = c.z; ops[7] = c.w; }
          if (nops > 8) ops[8] = __ldg(reinterpret_cast<const uint32_t*>(p + 2));
        }
        if (nops < 1) { ok = false; break; }
        bool spliced = false;
        for (int k = 0; k < nops; ++k) spliced = spliced || opc(ops[k]) == OP_REF_SKIP;
        if (e == 0) { ref0 = a.x; left0 = (int)a.y; }
        else {
          if (!(prev_last_match || opc(ops[0]) == OP_MATCH)) { ok = false; break; }               // 930-934
          if (prev_spliced && spliced && prev_asplice != asplice) { ok = false; break; }            // 942-949
          if (a.x != prev_ref) { ok = false; break; }
        }
        // finalise this hit (1888-1945)
        num_mm += (int)((a.w >> 16) & 0xffu); num_smm += (int)(a.w >> 24);
        if (spliced) { if (asplice) { if (saw_s) { ok = false; break; } saw_as = true; } else { if (saw_as) { ok = false; break; } saw_s = true; } }
        int b0 = 0;
        if (nLC > 0 && opc(LC[nLC - 1]) == opc(ops[0])) { LC[nLC - 1] = mkop(opc(LC[nLC - 1]), opl(LC[nLC - 1]) + opl(ops[0])); b0 = 1; }
        for (; b0 < nops; ++b0) if (!cig_push(LC, nLC, ops[b0])) { ok = false; break; }
        prev_spliced = spliced; prev_asplice = asplice; prev_last_match = opc(ops[nops - 1]) == OP_MATCH; prev_ref = a.x;
      }
      if (ok && nLC == 0) ok = false;
      if (ok) {
        if (anti) { uint64_t F[12];
          #pragma unroll
          for (int k = 0; k < 12; ++k) F[k] = R[k];
          revcomp_read(F, read_len, R); }
        // new_read_len == old_read_length (2023): fusing equal neighbours preserves the lengths
        ok = editdist_consistent(ref, ref0, left0, LC, nLC, R, 4, (uint8_t)num_mm) && valid_cigar(P, LC, nLC);
      }
    }
    const unsigned em = __ballot_sync(0xffffffffu, ok);
    flush();                                                     // the previous iteration's records, CIGARs in LCB[par ^ 1]
    p_em = em; p_ok = ok; p_n = nLC; p_smm = (uint32_t)num_smm & 0xffu;
    if (ok) {
      const uint32_t mism = (uint32_t)num_mm & 0xffu, ed = ((uint32_t)num_mm + (uint32_t)cig_gap_length(LC, nLC)) & 0xffu;
      const uint32_t flags = (anti ? (uint32_t)THB_HIT_ANTISENSE : 0u) | (saw_as ? (uint32_t)THB_JHIT_ANTISENSE_SPLICE : 0u);
      p_hdr = make_uint4(bi + bv.bundle_base, ref0, (uint32_t)left0, (uint32_t)nLC | (flags << 8) | (mism << 16) | (ed << 24));
    }
    if (em && lane == (unsigned)(__ffs((int)em) - 1)) p_slot0 = atomicAdd(o.count, (unsigned long long)__popc(em));
    par ^= 1;
    if (!DEFER) flush();
    __syncwarp();
  }
  flush();
  for (int k = 16; k > 0; k >>= 1) n_emit += __shfl_xor_sync(0xffffffffu, n_emit, k);
  if (lane == 0 && n_emit) atomicAdd(o.counters + 2, (unsigned long long)n_emit);
}

// K-J2: merge_segment_chain (2101-2220) + merge_chain (805-2038) + valid_hit (2045-2099), one chain per thread.
// Every lane walks the same phase sequence with an `alive` predicate and the warp re-converges at each __syncwarp():
// the closure searches (equal-length binary searches) and the consistency re-read then run with all the lanes that
// need them side by side (profiles/r1h: 4 of 32 lanes active in the straightforward per-thread form).
template <int MINB>
__global__ void __launch_bounds__(128, MINB)
chain_merge_kernel(RefView ref, JoinParams P, JoinSets S, JoinBatchView bv, ChainQueue q, JoinOut o)
{
  unsigned n_closures = 0, n_emit = 0;
  unsigned long long nq = *q.count; if (nq > q.cap) nq = q.cap;
  const uint32_t* __restrict__ qtasks = q.tasks;
  const unsigned lane = threadIdx.x & 31u;
  for (unsigned long long base = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x - lane; base < nq; base += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long ti = base + lane;
    // ---- phase 0: task, read, orientation
    bool alive = ti < nq;
    const uint32_t* __restrict__ t = qtasks + (alive ? ti : 0) * q.stride;
    uint32_t bi = 0, ops_begin = 0; int read_len = 0, n = 0; bool anti = false;
    uint64_t R[12];
    #pragma unroll
    for (int k = 0; k < 12; ++k) R[k] = 0;
    if (alive) {
      bi = t[0];
      const uint4 hdr = __ldg(reinterpret_cast<const uint4*>(bv.bundles + bi));
      read_len = (int)(hdr.z & 0xffffu); n = (int)((hdr.z >> 16) & 0xffu); ops_begin = hdr.w;
      const uint64_t* rd = bv.reads + (size_t)bi * 3 * bv.read_words; const int rw = (int)bv.read_words;
      #pragma unroll
      for (int pl = 0; pl < 3; ++pl)
        #pragma unroll
        for (int w = 0; w < 4; ++w) R[pl * 4 + w] = w < rw ? __ldg(rd + pl * rw + w) : 0ull;
      anti = (__ldg(reinterpret_cast<const uint32_t*>(bv.hits + t[1]) + 3) & THB_HIT_ANTISENSE) != 0;     // chain orientation (2117-2121)
    }
    __syncwarp();
    if (alive && anti && n > 1) { uint64_t F[12];
      #pragma unroll
      for (int k = 0; k < 12; ++k) F[k] = R[k];
      revcomp_read(F, read_len, R); }
    __syncwarp();
    int nmax = n;
    #pragma unroll
    for (int k = 16; k > 0; k >>= 1) nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, k));
    auto chain_hit = [&](int e) -> uint32_t { const int s = anti ? n - 1 - e : e; return t[1 + s]; };
    auto chain_len = [&](int e) -> int { const int s = anti ? n - 1 - e : e; return (s == n - 1) ? read_len - s * P.seglen : P.seglen; };

    thb_joined j; j.bundle = bi; j.reserved8[0] = j.reserved8[1] = j.reserved8[2] = 0; j.n_ops = 0;
    // ---- single-segment reads: merge_segment_chain 2196-2213
    const bool single = alive && n == 1;
    if (single) {
      WHit w; load_whit(w, bv, t[1], ops_begin, 0, read_len);
      alive = w.n > 0 && valid_cigar(P, w.ops, w.n);
      if (alive) { j.ref_id = w.ref; j.left = w.left; j.n_ops = (uint8_t)w.n; j.flags = (uint8_t)((w.anti ? THB_HIT_ANTISENSE : 0) | (w.asplice ? THB_JHIT_ANTISENSE_SPLICE : 0));
                   j.mismatches = w.mism; j.edit_dist = (uint8_t)(w.mism + cig_gap_length(w.ops, w.n)); j.splice_mms = w.smm;
                   for (int k = 0; k < w.n; ++k) j.ops[k] = w.ops[k]; }
    }
    bool multi = alive && !single;
    __syncwarp();
    // the first pass of merge_chain (843-897: two gaps that only a fusion could explain -> give up) is folded into the stitch
    // loop below: its `gap` is the `dist` of phase 1 (a merged block keeps the right end of its last hit)
    int num_fusions = 0;
    uint64_t cs = 0; int64_t clen = 0;
    if (multi) {
      const uint32_t r0 = __ldg(&bv.hits[chain_hit(0)].ref_id);
      if (!(r0 >= 1 && r0 <= ref.n_contigs)) multi = false;
      else { cs = __ldg(ref.contig_start + r0 - 1); clen = (int64_t)__ldg(ref.contig_len + r0 - 1); }
    }
    // accumulators of the final pass (1888-1945), filled as blocks are finalised
    uint32_t LC[JMAXOPS]; int nLC = 0; int num_mm = 0, num_smm = 0; bool saw_as = false, saw_s = false;
    bool cig_ovf = false;              // a CIGAR outgrew JMAXOPS: reported, never silently dropped (the reference has no such limit)
    int old_read_length = 0;
    auto finalize = [&](const WHit& h) -> bool {
      num_mm += h.mism; num_smm += h.smm;
      if (h.spliced) {
        if (h.asplice) { if (saw_s) return false; saw_as = true; } else { if (saw_as) return false; saw_s = true; }
      }
      int b = 0;
      if (nLC > 0 && opc(LC[nLC - 1]) == opc(h.ops[0])) { LC[nLC - 1] = mkop(opc(LC[nLC - 1]), opl(LC[nLC - 1]) + opl(h.ops[0])); b = 1; }
      for (; b < h.n; ++b) if (!cig_push(LC, nLC, h.ops[b])) { cig_ovf = true; return false; }
      return true;
    };
    WHit prev; prev.n = 0; prev.ref = 0; prev.left = 0; prev.anti = prev.asplice = false; prev.mism = prev.smm = 0; prev.seq_pos = prev.seq_len = 0;
    prev.right = 0; prev.rlen = 0; prev.spliced = false;
    int left0 = 0; uint32_t ref0 = 0; bool antisense = false;
    if (multi) {
      load_whit(prev, bv, chain_hit(0), ops_begin, 0, chain_len(0));
      left0 = prev.left; ref0 = prev.ref; antisense = prev.anti;
      old_read_length += prev.rlen;
    }
    __syncwarp();
    // ---- stitch loop (899-1880)
    for (int e = 1; e < nmax; ++e) {
      const bool on = multi && e < n;
      WHit curr; curr.n = 0;
      int kind = 0;                 // 0 contiguous, 1 insertion closure, 2 junction closure
      bool found = false, antisense_closure = false; int mismatch = 0, prml = 0, clml = 0, pright = 0;
      // phase 1: load + classify
      if (on) {
        load_whit(curr, bv, chain_hit(e), ops_begin, prev.seq_pos + prev.seq_len, chain_len(e));
        old_read_length += curr.rlen;
        antisense = prev.anti;
        const bool ps = prev.spliced, csp = curr.spliced;
        if (!(opc(prev.ops[prev.n - 1]) == OP_MATCH || opc(curr.ops[0]) == OP_MATCH)) multi = false;          // 930-934
        else if (ps && csp && prev.asplice != curr.asplice) multi = false;                                      // 942-949
        else if (curr.ref != prev.ref) multi = false;
        else {
          antisense_closure = ps ? prev.asplice : curr.asplice;
          prml = (int)opl(prev.ops[prev.n - 1]); clml = (int)opl(curr.ops[0]);
          pright = prev.right;
          const int dist = curr.left - pright;
          const bool same_strand = prev.anti == curr.anti;
          { const int hi = min(P.max_report_intron, P.fusion_min_dist);
            if (dist < -P.max_ins || (dist > P.max_del && (dist < P.min_report_intron || dist > hi))) ++num_fusions;
            if (num_fusions >= 2) multi = false; }
          if (dist < 0 && dist >= -P.max_ins && same_strand) kind = 1;
          else if (dist > 0 && dist <= P.max_report_intron && same_strand) kind = 2;
          else if (!(dist == 0 && same_strand)) multi = false;     // only a fusion could close this gap (1592-1819)
        }
      }
      __syncwarp();
      // phase 2: junction / deletion closure
      JuncClosure jc; jc.dtl = 0; jc.glen = 0; jc.anti = false; jc.new_diff = 0;
      if (on && multi && kind == 2) {
        n_closures++;
        const int rc = junction_closure(ref, S, R, cs, clen, prev.ref, pright, curr.left, prml, clml, prev.seq_pos + prev.seq_len, curr.seq_pos, jc);
        if (rc <= 0) multi = false; else { found = true; mismatch = jc.new_diff; }
      }
      __syncwarp();
      // phase 3: insertion closure
      InsClosure ic; ic.itpr = 0; ic.len = 0; ic.mismatch = 0;
      if (on && multi && kind == 1) {
        n_closures++;
        const int rc = insertion_closure(ref, S, P, R, cs, clen, prev.ref, pright, curr.left, prml, clml, prev.seq_pos + prev.seq_len, curr.seq_pos, ic);
        if (rc <= 0) multi = false; else { found = true; mismatch = ic.mismatch; }
      }
      __syncwarp();
      // phase 4: stitch (1262-1290, 1520-1585, 1822-1880)
      if (on && multi) {
        if (found) {
          uint32_t NC[JMAXOPS]; int nNC = 0; bool okc = true;
          for (int k = 0; k < prev.n; ++k) NC[nNC++] = prev.ops[k];
          if (kind == 1) {
            // lengths are uint32 in the reference, the arithmetic wraps the same way
            { const uint32_t bl = opl(NC[nNC - 1]) - (uint32_t)ic.itpr; if ((bl & 0x0fffffffu) == 0) --nNC; else NC[nNC - 1] = mkop(opc(NC[nNC - 1]), bl & 0x0fffffffu); }
            okc = cig_push(NC, nNC, mkop(OP_INS, ic.len));
            const uint32_t fl = (opl(curr.ops[0]) + (uint32_t)(ic.itpr - (int)ic.len)) & 0x0fffffffu;
            for (int k = fl > 0 ? 0 : 1; k < curr.n && okc; ++k) okc = cig_push(NC, nNC, k == 0 ? mkop(opc(curr.ops[0]), fl) : curr.ops[k]);
          } else {
            { const int nlb = prml + jc.dtl; if (nlb > 0) NC[nNC - 1] = mkop(opc(NC[nNC - 1]), (uint32_t)nlb); else --nNC; }
            if (jc.glen <= (uint32_t)P.max_del) okc = cig_push(NC, nNC, mkop(OP_DEL, jc.glen));
            else { okc = cig_push(NC, nNC, mkop(OP_REF_SKIP, jc.glen)); antisense_closure = jc.anti; }
            const int nrf = clml - jc.dtl;
            for (int k = nrf > 0 ? 0 : 1; k < curr.n && okc; ++k) okc = cig_push(NC, nNC, k == 0 ? mkop(opc(curr.ops[0]), (uint32_t)nrf) : curr.ops[k]);
          }
          if (!okc) cig_ovf = true;
          if (!okc || nNC == 0) multi = false;
          else {
            // merged_hit (1822-1838); _mismatches / _edit_dist are unsigned chars in the reference
            const int mm = (int)prev.mism + (int)curr.mism + mismatch;
            prev.n = nNC; for (int k = 0; k < nNC; ++k) prev.ops[k] = NC[k];
            prev.asplice = antisense_closure; prev.anti = antisense; prev.mism = (uint8_t)mm; prev.smm = (uint8_t)(prev.smm + curr.smm);
            prev.seq_len += curr.seq_len;
            prev.right = curr.right; prev.rlen += curr.rlen;
            prev.spliced = prev.spliced || curr.spliced || (kind == 2 && jc.glen > (uint32_t)P.max_del);
          }
        } else {
          if (!finalize(prev)) multi = false;
          else {
            prev.ref = curr.ref; prev.left = curr.left; prev.n = curr.n; prev.anti = curr.anti; prev.asplice = curr.asplice;
            prev.mism = curr.mism; prev.smm = curr.smm; prev.seq_pos = curr.seq_pos; prev.seq_len = curr.seq_len;
            prev.right = curr.right; prev.rlen = curr.rlen; prev.spliced = curr.spliced;
            for (int k = 0; k < curr.n; ++k) prev.ops[k] = curr.ops[k];
          }
        }
      }
      __syncwarp();
    }
    // ---- new_hit (1947-1957) + final checks (2023-2035) + valid_hit
    if (multi) {
      if (!finalize(prev) || nLC == 0) multi = false;
      else if (cig_read_len(LC, nLC) != old_read_length) multi = false;
    }
    __syncwarp();
    if (multi) {
      const uint8_t mism = (uint8_t)num_mm;
      if (!editdist_consistent(ref, ref0, left0, LC, nLC, R, 4, mism) || !valid_cigar(P, LC, nLC)) multi = false;
      else {
        j.ref_id = ref0; j.left = left0; j.n_ops = (uint8_t)nLC;
        j.flags = (uint8_t)((antisense ? THB_HIT_ANTISENSE : 0) | (saw_as ? THB_JHIT_ANTISENSE_SPLICE : 0));
        j.mismatches = mism; j.edit_dist = (uint8_t)(num_mm + cig_gap_length(LC, nLC)); j.splice_mms = (uint8_t)num_smm;
      }
    }
    if (cig_ovf) atomicOr(o.overflow, 2u);
    __syncwarp();
    // ---- emit (warp-aggregated slot allocation)
    const bool out_ok = (single && alive) || multi;
    const unsigned em = __ballot_sync(0xffffffffu, out_ok);
    if (em) {
      unsigned long long slot0 = 0;
      if (lane == (unsigned)(__ffs((int)em) - 1)) slot0 = atomicAdd(o.count, (unsigned long long)__popc(em));
      slot0 = __shfl_sync(0xffffffffu, slot0, __ffs((int)em) - 1);
      if (out_ok) {
        const unsigned long long slot = slot0 + (unsigned long long)__popc(em & ((1u << lane) - 1u));
        if (slot >= o.cap) atomicOr(o.overflow, 1u);
        else {
          uint4* dst = reinterpret_cast<uint4*>(o.rec + slot);
          const uint32_t hdr3 = (uint32_t)j.n_ops | ((uint32_t)j.flags << 8) | ((uint32_t)j.mismatches << 16) | ((uint32_t)j.edit_dist << 24);
          dst[0] = make_uint4(j.bundle + bv.bundle_base, j.ref_id, (uint32_t)j.left, hdr3);
          // whole 32-byte sectors only (a partly written sector costs a DRAM read to fill it): ops beyond n_ops are zero
          const int nop = j.n_ops;
          const uint32_t* src = single ? j.ops : LC;
          auto op_at = [&](int k) -> uint32_t { return k < nop ? src[k] : 0u; };
          dst[1] = make_uint4((uint32_t)j.splice_mms, op_at(0), op_at(1), op_at(2));
          for (int q = 2; 4 * q - 5 < nop; q += 2) {
            dst[q] = make_uint4(op_at(4 * q - 5), op_at(4 * q - 4), op_at(4 * q - 3), op_at(4 * q - 2));
            dst[q + 1] = make_uint4(op_at(4 * q - 1), op_at(4 * q), op_at(4 * q + 1), op_at(4 * q + 2));
          }
        }
        ++n_emit;
      }
    }
    __syncwarp();
  }
  for (int k = 16; k > 0; k >>= 1) { n_closures += __shfl_xor_sync(0xffffffffu, n_closures, k); n_emit += __shfl_xor_sync(0xffffffffu, n_emit, k); }
  if (lane == 0) { if (n_closures) atomicAdd(o.counters + 1, (unsigned long long)n_closures); if (n_emit) atomicAdd(o.counters + 2, (unsigned long long)n_emit); }
}

}  // namespace thb
