// fusion_join_kernel.cuh -- long_spanning_reads with --fusion-search: the segment-chain join where a chain may cross ONE
// fusion point (two loci on different contigs, strands or far apart), thread = read.
//
// Replaces, for --fusion-search on, the reference's (src/long_spanning_reads.cpp, src/bwt_map.h)
//   join_segments_for_read 2612-2667, dfs_seg_hits 2222-2610 (pair rules with fusion directions, hit reversal),
//   merge_segment_chain 2101-2220 (chain orientation), merge_chain 805-2038 (insertion / junction / deletion closures in
//   forward and reversed orientation, the fusion closure against the fusion set), valid_hit 2045-2099,
//   BowtieHit::reverse bwt_map.h:331-442, right / is_forwarding_* 213-320, check_editdist_consistency bwt_map.cpp:2349-2465,
//   fusions_from_spliced_hit fusions.cpp:441-496.
// With --fusion-search every read takes this kernel -- the pair rules themselves change (any two hits may be neighbours if a
// fusion could explain them) -- so it is written for exactness, not speed: configs[4] is a correctness configuration of a few
// hundred reads.  A hit is a small record in local memory (CIGAR of <= 27 ops, the read pieces its sequence consists of), the
// DFS an explicit stack.  Sequence comparisons are per base on the bit planes.
#pragma once
#include "join_kernel.cuh"

#ifdef THB_EMU
#define FJ_DBG(...) do { if (getenv("THB_FJ_DEBUG")) { fprintf(stderr, "[fj] " __VA_ARGS__); fputc('\n', stderr); } } while (0)
#else
#define FJ_DBG(...) do { } while (0)
#endif

namespace thb {

enum { FUS_NOTHING = 0, FUS_FF = 7, FUS_FR = 8, FUS_RF = 9, FUS_RR = 10 };     // CigarOpCode values, bwt_map.h:36-55
__device__ __forceinline__ bool is_fus(int c) { return c >= FUS_FF && c <= FUS_RR; }

struct FusionKey { uint32_t r1, r2, left, right, dir; };                       // Fusion::operator< order (fusions.h:40-70)
struct FJoinSets { JoinSets base; const FusionKey* fus; uint32_t n_fus; };

constexpr int FJ_PIECES = THB_MAX_SEGS;
// A BowtieHit of the join.  seq() is not stored: it is the concatenation of `np` pieces of the read, piece k = bases
// [start, start + len) of the read, reverse-complemented when rc is set (a segment's BAM sequence is the read's segment,
// reverse-complemented for an antisense alignment; BowtieHit::reverse complements it again; merges concatenate).
struct FHit {
  uint32_t ref, ref2; int left; int n; uint32_t ops[JMAXOPS];
  bool anti, asplice; uint8_t mism, smm;
  int np; uint16_t pstart[FJ_PIECES]; uint16_t plen[FJ_PIECES]; uint8_t prc[FJ_PIECES];
};

__device__ __forceinline__ int fh_fusion_opcode(const FHit& h) { for (int i = 0; i < h.n; ++i) if (is_fus(opc(h.ops[i]))) return opc(h.ops[i]); return FUS_NOTHING; }
__device__ __forceinline__ int fh_right(const FHit& h)                          // bwt_map.h:213-243
{
  int r = h.left;
  for (int i = 0; i < h.n; ++i) { const int c = opc(h.ops[i]); const int l = (int)opl(h.ops[i]);
    if (c == OP_MATCH || c == OP_REF_SKIP || c == OP_DEL) r += l; else if (c == OP_mATCH || c == OP_rEF_SKIP || c == OP_dEL) r -= l; else if (is_fus(c)) r = l; }
  return r;
}
__device__ __forceinline__ bool fh_spliced(const FHit& h) { for (int i = 0; i < h.n; ++i) { const int c = opc(h.ops[i]); if (c == OP_REF_SKIP || c == OP_rEF_SKIP) return true; } return false; }
__device__ __forceinline__ bool fwd_code(int c) { return c == OP_MATCH || c == OP_REF_SKIP || c == OP_INS || c == OP_DEL; }
__device__ __forceinline__ bool bwd_code(int c) { return c == OP_mATCH || c == OP_rEF_SKIP || c == OP_iNS || c == OP_dEL; }
__device__ __forceinline__ bool fh_fwd_left(const FHit& h)                      // is_forwarding_left 266-284
{ for (int i = 0; i < h.n; ++i) { const int c = opc(h.ops[i]); if (fwd_code(c)) return true; if (bwd_code(c)) return false; if (is_fus(c)) break; } return true; }
__device__ __forceinline__ bool fh_fwd_right(const FHit& h)                     // is_forwarding_right 290-308
{ for (int i = h.n - 1; i >= 0; --i) { const int c = opc(h.ops[i]); if (fwd_code(c)) return true; if (bwd_code(c)) return false; if (is_fus(c)) break; } return true; }
__device__ __forceinline__ bool fh_anti2(const FHit& h)                         // antisense_align2 314-325
{ const int f = fh_fusion_opcode(h); return (f == FUS_NOTHING || f == FUS_FF || f == FUS_RR) ? h.anti : !h.anti; }
__device__ __forceinline__ int fh_read_len(const FHit& h)
{ int r = 0; for (int i = 0; i < h.n; ++i) { const int c = opc(h.ops[i]); if (c == OP_MATCH || c == OP_mATCH || c == OP_INS || c == OP_iNS || c == OP_SOFT_CLIP) r += (int)opl(h.ops[i]); } return r; }
__device__ __forceinline__ int fh_gap_length(const FHit& h)
{ int r = 0; for (int i = 0; i < h.n; ++i) { const int c = opc(h.ops[i]); if (c == OP_INS || c == OP_iNS || c == OP_DEL || c == OP_dEL) r += (int)opl(h.ops[i]); } return r; }
__device__ __forceinline__ int fh_seq_len(const FHit& h) { int r = 0; for (int k = 0; k < h.np; ++k) r += h.plen[k]; return r; }

__device__ __forceinline__ int flip_case(int c)
{
  switch (c) { case OP_MATCH: return OP_mATCH; case OP_mATCH: return OP_MATCH; case OP_INS: return OP_iNS; case OP_iNS: return OP_INS;
               case OP_DEL: return OP_dEL; case OP_dEL: return OP_DEL; case OP_REF_SKIP: return OP_rEF_SKIP; case OP_rEF_SKIP: return OP_REF_SKIP; default: return c; }
}

// BowtieHit::reverse (bwt_map.h:331-442)
__device__ void fh_reverse(const FHit& h, FHit& o)
{
  o.ref = h.ref2; o.ref2 = h.ref;
  uint32_t right = (uint32_t)h.left, fusion_pos = (uint32_t)h.left;
  for (int i = 0; i < h.n; ++i) { const int c = opc(h.ops[i]); const uint32_t l = opl(h.ops[i]);
    if (c == OP_MATCH || c == OP_REF_SKIP || c == OP_DEL) right += l; else if (c == OP_mATCH || c == OP_rEF_SKIP || c == OP_dEL) right -= l;
    else if (is_fus(c)) { fusion_pos = right; right = l; } }
  const bool fl = fh_fwd_left(h);
  if (fl) fusion_pos -= 1; else fusion_pos += 1;
  const int f = fh_fusion_opcode(h);
  if (f == FUS_NOTHING || f == FUS_FF || f == FUS_RR) o.left = fl ? (int)(right - 1) : (int)(right + 1);
  else o.left = f == FUS_FR ? (int)(right + 1) : (int)(right - 1);
  o.n = h.n;
  for (int i = 0; i < h.n; ++i) { const uint32_t op = h.ops[h.n - 1 - i]; const int c = opc(op);
    o.ops[i] = is_fus(c) ? mkop(c, fusion_pos & 0x0fffffffu) : mkop(flip_case(c), opl(op)); }
  o.anti = (f == FUS_FR || f == FUS_RF) ? !h.anti : h.anti;
  o.asplice = h.asplice; o.mism = h.mism; o.smm = h.smm;
  o.np = h.np;                                                    // reverse complement of the sequence
  for (int k = 0; k < h.np; ++k) { o.pstart[k] = h.pstart[h.np - 1 - k]; o.plen[k] = h.plen[h.np - 1 - k]; o.prc[k] = h.prc[h.np - 1 - k] ^ 1; }
}

// Dna5 code (0..3, 4 = N) of base i of the read R (stride-4 planes) / of the hit's sequence
__device__ __forceinline__ int fj_read_code(const uint64_t* R, int i) { return read_code5(R, i); }
__device__ __forceinline__ int comp5(int c) { return c > 3 ? 4 : 3 - c; }
__device__ int fh_seq_code(const FHit& h, const uint64_t* R, int i)
{
  for (int k = 0; k < h.np; ++k) {
    if (i < (int)h.plen[k]) return h.prc[k] ? comp5(fj_read_code(R, (int)h.pstart[k] + (int)h.plen[k] - 1 - i)) : fj_read_code(R, (int)h.pstart[k] + i);
    i -= (int)h.plen[k];
  }
  return 4;
}
// is the hit's sequence the read itself (0), its reverse complement (1), or neither (2)?
__device__ int fh_seq_orientation(const FHit& h, int read_len)
{
  bool fw = true, rc = true; int pos = 0;
  for (int k = 0; k < h.np; ++k) {
    if (h.plen[k] == 0) continue;
    if (!(h.prc[k] == 0 && (int)h.pstart[k] == pos)) fw = false;
    if (!(h.prc[k] == 1 && (int)h.pstart[k] + (int)h.plen[k] == read_len - pos)) rc = false;
    pos += (int)h.plen[k];
  }
  if (pos != read_len) return 2;
  return fw ? 0 : (rc ? 1 : 2);
}

struct FRef { const RefView* ref; uint64_t cs; int64_t len; };
__device__ __forceinline__ bool fref_of(const RefView& ref, uint32_t id, FRef& o)
{
  if (!(id >= 1 && id <= ref.n_contigs)) return false;
  o.ref = &ref; o.len = (int64_t)__ldg(ref.contig_len + id - 1); o.cs = __ldg(ref.contig_start + id - 1);
  return o.len > 0;
}
// base at contig position p; -1 where the reference would read outside the contig
__device__ __forceinline__ int fref_code(const FRef& r, int64_t p) { if (p < 0 || p >= r.len) return -1; return ref_code5(*r.ref, r.cs + (uint64_t)p); }

// BowtieHit::check_editdist_consistency (bwt_map.cpp:2349-2465), fusion-aware
__device__ bool fh_editdist_consistent(const RefView& ref, const FHit& h, const uint64_t* R)
{
  FRef r1, r2;
  if (!fref_of(ref, h.ref, r1) || !fref_of(ref, h.ref2, r2)) return false;
  const FRef* cur = &r1;
  int64_t pos_ref = h.left; int pos_seq = 0; unsigned mm = 0, nmm = 0; bool saw = false;
  const int slen = fh_seq_len(h);
  for (int i = 0; i < h.n; ++i) {
    const int c = opc(h.ops[i]); const int l = (int)opl(h.ops[i]);
    if (c == OP_MATCH || c == OP_mATCH) {
      for (int j = 0; j < l; ++j) {
        const int64_t p = c == OP_MATCH ? pos_ref + j : pos_ref - j;
        int g = fref_code(*cur, p); if (g < 0) return false;
        if (c == OP_mATCH) g = comp5(g);
        if (pos_seq >= slen) return false;
        const int s = fh_seq_code(h, R, pos_seq);
        if (s != g) ++mm; else if (s == 4) ++nmm;
        ++pos_seq;
      }
      pos_ref += c == OP_MATCH ? l : -l;
    } else if (c == OP_INS || c == OP_iNS) pos_seq += l;
    else if (c == OP_DEL || c == OP_REF_SKIP) pos_ref += l;
    else if (c == OP_dEL || c == OP_rEF_SKIP) pos_ref -= l;
    else if (is_fus(c)) { if (saw) return false; cur = &r2; pos_ref = l; saw = true; }
  }
  return mm == (unsigned)h.mism || mm + nmm == (unsigned)h.mism;
}

// valid_hit (2045-2099)
__device__ bool fh_valid(const JoinParams& P, const FHit& h)
{
  if (h.n < 1) return false;
  for (int i = 1; i < h.n; ++i) {
    const int c = opc(h.ops[i]), p = opc(h.ops[i - 1]); const int64_t l = (int64_t)opl(h.ops[i]);
    if (!(c == OP_MATCH || c == OP_mATCH) && !(p == OP_MATCH || p == OP_mATCH)) return false;
    if ((c == OP_INS || c == OP_iNS) && l > P.max_ins) return false;
    if ((c == OP_DEL || c == OP_dEL) && l > P.max_del) return false;
    if ((c == OP_REF_SKIP || c == OP_rEF_SKIP) && l < (int64_t)P.min_report_intron) return false;
  }
  const int f = opc(h.ops[0]), b = opc(h.ops[h.n - 1]);
  return (f == OP_MATCH || f == OP_mATCH) && (b == OP_MATCH || b == OP_mATCH);
}

// do_reverse of merge_segment_chain 2196-2210 / merge_chain 1984-1999 (fusions_from_spliced_hit with auto_sort = false)
__device__ bool fh_do_reverse(const FHit& h)
{
  if (h.ref != h.ref2) return h.ref > h.ref2;
  uint32_t pos = (uint32_t)h.left;
  for (int i = 0; i < h.n; ++i) { const int c = opc(h.ops[i]); const uint32_t l = opl(h.ops[i]);
    if (c == OP_REF_SKIP || c == OP_MATCH || c == OP_DEL) pos += l; else if (c == OP_rEF_SKIP || c == OP_mATCH || c == OP_dEL) pos -= l;
    else if (is_fus(c)) { pos = (c == FUS_RF || c == FUS_RR) ? pos + 1 : pos - 1; return pos > l; } }
  return false;
}

// segment hit -> FHit (the hit's sequence = its segment of the read, reverse-complemented for an antisense alignment)
__device__ void fh_load(FHit& w, const JoinBatchView& bv, uint32_t hit, uint32_t ops_begin, int seg_pos, int seg_len, int read_len)
{
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(bv.hits + hit));
  w.ref = a.x; w.ref2 = a.x; w.left = (int)a.y;
  const uint32_t fl = a.w & 0xfu; w.mism = (uint8_t)((a.w >> 16) & 0xffu); w.smm = (uint8_t)(a.w >> 24);
  w.anti = (fl & THB_HIT_ANTISENSE) != 0; w.asplice = (fl & THB_JHIT_ANTISENSE_SPLICE) != 0;
  bool seq_flipped = false;
  if (fl & THB_JHIT_ONE_MATCH) { w.n = 1; w.ops[0] = mkop(OP_MATCH, (uint32_t)(a.z - a.y)); }
  else {
    w.n = (int)((a.w >> 4) & 0xfu); if (w.n > THB_JHIT_MAX_OPS) w.n = THB_JHIT_MAX_OPS;
    const uint32_t* p = reinterpret_cast<const uint32_t*>(bv.ops_ext + ops_begin + ((a.w >> 8) & 0xffu));
    for (int k = 0; k < w.n; ++k) w.ops[k] = __ldg(p + k);
    if (fh_fusion_opcode(w) != FUS_NOTHING) w.ref2 = __ldg(p + 11);            // thb_jops.ops[11]: second contig of a fusion hit
    seq_flipped = __ldg(p + 10) != 0u;                                         // thb_jops.ops[10]: THB_JHIT_SEQ_FLIPPED
  }
  // an antisense alignment covers the mirrored part of the read: segment s of the read was reverse-complemented by the mapper
  w.np = 1; w.pstart[0] = (uint16_t)seg_pos; w.plen[0] = (uint16_t)seg_len; w.prc[0] = (w.anti != seq_flipped) ? 1 : 0;
  (void)read_len;
}

__device__ __forceinline__ bool fh_push(FHit& h, uint32_t op) { if (h.n >= JMAXOPS) return false; h.ops[h.n++] = op; return true; }

// std::set<Fusion>::upper_bound / lower_bound over the sorted array (dir = FUSION_FF in both probe keys, 1629-1630)
__device__ bool fus_less(const FusionKey& a, uint32_t r1, uint32_t r2, uint32_t l, uint32_t r, uint32_t d)
{ if (a.r1 != r1) return a.r1 < r1; if (a.r2 != r2) return a.r2 < r2; if (a.left != l) return a.left < l; if (a.right != r) return a.right < r; return a.dir < d; }
__device__ bool fus_greater(const FusionKey& a, uint32_t r1, uint32_t r2, uint32_t l, uint32_t r, uint32_t d)
{ if (a.r1 != r1) return a.r1 > r1; if (a.r2 != r2) return a.r2 > r2; if (a.left != l) return a.left > l; if (a.right != r) return a.right > r; return a.dir > d; }

// ---- merge_chain (805-2038).  Returns false where the reference returns an empty BowtieHit. ------------------------------
__device__ bool fj_merge_chain(const RefView& ref, const JoinParams& P, const FJoinSets& S, const uint64_t* R, int read_len, FHit* chain, int nchain,
                               int fusion_dir, FHit& out, unsigned& n_closures, bool& cig_ovf)
{
  bool antisense = chain[0].anti;
  const int left = chain[0].left;
  int old_read_length = 0;
  for (int i = 0; i < nchain; ++i) old_read_length += fh_read_len(chain[i]);
  // the chain's sequence = concatenation of the hits' sequences (824-829)
  FHit seqh; seqh.np = 0;
  for (int i = 0; i < nchain; ++i) for (int k = 0; k < chain[i].np; ++k) { if (seqh.np >= FJ_PIECES) return false; seqh.pstart[seqh.np] = chain[i].pstart[k]; seqh.plen[seqh.np] = chain[i].plen[k]; seqh.prc[seqh.np] = chain[i].prc[k]; ++seqh.np; }

  // first pass (843-897): at most one fusion along the chain
  {
    int num_fusions = fh_fusion_opcode(chain[0]) == FUS_NOTHING ? 0 : 1; bool fusion_passed = false;
    for (int k = 1; k < nchain; ++k) {
      const FHit& p = chain[k - 1]; const FHit& c = chain[k];
      if (p.ref != p.ref2 || p.ref2 != c.ref) fusion_passed = true;
      if (p.ref2 != c.ref) ++num_fusions;
      if (fh_fusion_opcode(c) != FUS_NOTHING) ++num_fusions;
      if (p.ref2 == c.ref) {
        const bool reversed = (fusion_dir == FUS_FR && fusion_passed) || (fusion_dir == FUS_RF && !fusion_passed);
        const int gap = reversed ? fh_right(p) - c.left : c.left - fh_right(p);
        const int hi = min(P.max_report_intron, P.fusion_min_dist);
        if (gap < -P.max_ins || (gap > P.max_del && (gap < P.min_report_intron || gap > hi))) { fusion_passed = true; ++num_fusions; }
      }
      if (num_fusions >= 2) { FJ_DBG("merge: first pass num_fusions %d at k=%d", num_fusions, k); return false; }
    }
  }
  // stitch loop (899-1880): `prev` is the block being grown, finalised blocks go to fin[]
  FHit fin[JMAXSEGS]; int nfin = 0;
  FHit prev = chain[0];
  bool fusion_passed = false;
  for (int k = 1; k < nchain; ++k) {
    const FHit& curr = chain[k];
    const int curr_seg_index = k;
    antisense = prev.anti;
    if (fh_fusion_opcode(prev) != FUS_NOTHING || prev.ref2 != curr.ref) fusion_passed = true;
    { const int pb = opc(prev.ops[prev.n - 1]), cf = opc(curr.ops[0]);
      if (!(pb == OP_MATCH || cf == OP_MATCH || pb == OP_mATCH || cf == OP_mATCH)) return false; }              // 930-934
    const bool ps = fh_spliced(prev), csp = fh_spliced(curr);
    if (ps && csp && prev.asplice != curr.asplice) return false;                                                // 942-949
    bool found = false; bool antisense_closure = ps ? prev.asplice : curr.asplice;
    FHit nc; nc.n = 0; int new_left = -1; int mismatch = 0;
    const int prml = (int)opl(prev.ops[prev.n - 1]), clml = (int)opl(curr.ops[0]);
    const int pright = fh_right(prev);
    bool check_fusion = prev.ref2 != curr.ref;
    if (prev.ref2 == curr.ref) {
      const bool reversed = (fusion_dir == FUS_FR && fusion_passed) || (fusion_dir == FUS_RF && !fusion_passed);
      FRef rs; if (!fref_of(ref, prev.ref2, rs)) return false;
      const uint32_t reference_id = prev.ref2;
      const int left_boundary = reversed ? curr.left - 4 : pright - 4, right_boundary = reversed ? pright + 4 : curr.left + 4;
      const int dist_btw_two = reversed ? pright - curr.left : curr.left - pright;
      const bool strands_ok = fh_anti2(prev) == curr.anti;
      if (dist_btw_two < 0 && dist_btw_two >= -P.max_ins && strands_ok) {
        // ---- insertion closure (1010-1306)
        ++n_closures;
        uint32_t it = ins_upper_bound(S.base, rs.cs, rs.len, reference_id, (uint32_t)left_boundary, 0u);
        const uint32_t ub = ins_upper_bound(S.base, rs.cs, rs.len, reference_id, (uint32_t)right_boundary, (uint32_t)P.max_ins);
        for (; it != ub && it < S.base.n_ins; ++it) {
          const thb_insertion& I = S.base.ins[it];
          if ((int)I.len != (reversed ? curr.left - pright : pright - curr.left)) continue;
          int itpr, clti;
          if (reversed) { itpr = (int)I.left - pright; clti = curr.left - (int)I.left; }
          else { itpr = pright - (int)I.left - 1; clti = (int)I.left - curr.left + 1; }
          if (itpr > prml || clti > clml) continue;
          int this_ref_mm = 0, ins_mm = 0; const int ilen = (int)I.len;
          auto ins_code = [&](int x) -> int {                      // insertionSequence (reverse-complemented when reversed, 1069-1073)
            const char ch = reversed ? I.seq[ilen - 1 - x] : I.seq[x];
            const int c = ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : 4;
            return reversed ? comp5(c) : c; };
          if (itpr > 0) {
            // referenceSequence / oldSegmentSequence (1083-1113)
            const int plen_seq = fh_seq_len(prev);
            for (int ri = 0; ri < itpr; ++ri) {
              int rc_, rc2 = 4, sc;
              if (reversed) {
                // rc(ref[pright + 1, I.left + 1)) ; read[curr_seg_index * seglen - itpr + ri]
                const int64_t hi_ = (int64_t)I.left;            // last base of the infix
                rc_ = fref_code(rs, hi_ - ri); if (rc_ < 0) return false; rc_ = comp5(rc_);
                if (ri >= ilen) { rc2 = fref_code(rs, hi_ - (ri - ilen)); if (rc2 < 0) return false; rc2 = comp5(rc2); }
                const int rp = curr_seg_index * P.seglen - itpr + ri; sc = (rp >= 0 && rp < read_len) ? fj_read_code(R, rp) : 4;
              } else {
                const int64_t g0 = (int64_t)I.left + 1;
                rc_ = fref_code(rs, g0 + ri); if (rc_ < 0) return false;
                if (ri >= ilen) { rc2 = fref_code(rs, g0 + ri - ilen); if (rc2 < 0) return false; }
                sc = fh_seq_code(prev, R, plen_seq - itpr + ri);
              }
              if (rc_ == 4 || rc_ != sc) ++this_ref_mm;
              if (ri < ilen) { const int ic = ins_code(ri); if (ic == 4 || ic != sc) { ++ins_mm; break; } }
              else if (rc2 == 4 || rc2 != sc) --this_ref_mm;
            }
          }
          if (clti > 0) {
            for (int ri = 0; ri < clti; ++ri) {
              const int sp = clti - ri - 1, ip = ilen - ri - 1;
              int rc_, rc2 = 4, sc;
              if (reversed) {
                // rc(ref[I.left + 1, curr.left + 1)); read[curr_seg_index * seglen + sp]
                const int64_t hi_ = (int64_t)curr.left;
                rc_ = fref_code(rs, hi_ - sp); if (rc_ < 0) return false; rc_ = comp5(rc_);
                if (ri >= ilen) { rc2 = fref_code(rs, hi_ - (sp + ilen)); if (rc2 < 0) return false; rc2 = comp5(rc2); }
                const int rp = curr_seg_index * P.seglen + sp; sc = (rp >= 0 && rp < read_len) ? fj_read_code(R, rp) : 4;
              } else {
                const int64_t g0 = curr.left;
                rc_ = fref_code(rs, g0 + sp); if (rc_ < 0) return false;
                if (ri >= ilen) { rc2 = fref_code(rs, g0 + sp + ilen); if (rc2 < 0) return false; }
                sc = fh_seq_code(curr, R, sp);
              }
              if (rc_ == 4 || rc_ != sc) ++this_ref_mm;
              if (ri < ilen) { const int ic = ins_code(ip); if (ic == 4 || ic != sc) { ++ins_mm; break; } }
              else if (rc2 == 4 || rc2 != sc) --this_ref_mm;
            }
          }
          if (found) return false;                                  // a second same-length candidate rejects the chain (1246-1250)
          if (ins_mm == 0) {
            mismatch = -this_ref_mm; found = true; new_left = prev.left;
            nc = prev;
            { const uint32_t bl = (opl(nc.ops[nc.n - 1]) - (uint32_t)itpr) & 0x0fffffffu; if ((int)(opl(nc.ops[nc.n - 1])) - itpr <= 0) --nc.n; else nc.ops[nc.n - 1] = mkop(opc(nc.ops[nc.n - 1]), bl); }
            if (!fh_push(nc, mkop(reversed ? OP_iNS : OP_INS, I.len))) { cig_ovf = true; return false; }
            const int fl = (int)opl(curr.ops[0]) + (itpr - (int)I.len);
            for (int x = fl > 0 ? 0 : 1; x < curr.n; ++x) if (!fh_push(nc, x == 0 ? mkop(opc(curr.ops[0]), (uint32_t)fl) : curr.ops[x])) { cig_ovf = true; return false; }
          }
        }
        if (!found) return false;
      } else if (dist_btw_two > 0 && dist_btw_two <= P.max_report_intron && strands_ok) {
        // ---- junction / deletion closure (1311-1591)
        ++n_closures;
        uint32_t it = junc_bound_idx(S.base, rs.cs, rs.len, reference_id, (uint32_t)left_boundary, (uint32_t)(right_boundary - 8), 1u, true);
        const uint32_t ub = junc_bound_idx(S.base, rs.cs, rs.len, reference_id, (uint32_t)(left_boundary + 8), (uint32_t)right_boundary, 0u, false);
        int new_diff = 0xff;
        for (; it != ub && it < S.base.n_juncs; ++it) {
          const thb_junction J = S.base.juncs[it];
          int dtl, dtr;
          if (reversed) { dtl = (int)J.left - curr.left; dtr = (int)J.right - pright - 1; }
          else { dtl = (int)J.left - pright + 1; dtr = (int)J.right - curr.left; }
          if (!(abs(dtl) <= 4 && abs(dtr) <= 4 && dtl == dtr)) continue;
          if ((reversed && (dtl > prml || -dtl > clml)) || (!reversed && (dtl > clml || -dtl > prml))) continue;
          int new_mm = 0, old_mm = 0;
          if (dtl > 0) {
            for (int i = 0; i < dtl; ++i) {
              int ncmp, ocmp, cs_;
              if (reversed) {
                // new = rc(ref[curr.left + 1, J.left + 1)), old = rc(ref[pright + 1, J.right)); read[curr_seg_index * seglen - dtl + i]
                ncmp = fref_code(rs, (int64_t)J.left - i); ocmp = fref_code(rs, (int64_t)J.right - 1 - i);
                if (ncmp < 0 || ocmp < 0) return false; ncmp = comp5(ncmp); ocmp = comp5(ocmp);
                const int rp = curr_seg_index * P.seglen - dtl + i; cs_ = (rp >= 0 && rp < read_len) ? fj_read_code(R, rp) : 4;
              } else {
                ncmp = fref_code(rs, (int64_t)pright + i); ocmp = fref_code(rs, (int64_t)curr.left + i);
                if (ncmp < 0 || ocmp < 0) return false;
                cs_ = fh_seq_code(curr, R, i);
              }
              if (cs_ != ncmp) ++new_mm;
              if (cs_ != ocmp) ++old_mm;
            }
          } else if (dtl < 0) {
            const int ad = -dtl;
            const int plen_seq = fh_seq_len(prev);
            for (int i = 0; i < ad; ++i) {
              int ncmp, ocmp, ps_;
              if (reversed) {
                // new = rc(ref[J.right, pright + 1)), old = rc(ref[J.left + 1, curr.left + 1)); prev_hit_seq = read[curr_seg_index * seglen, + ad)
                ncmp = fref_code(rs, (int64_t)pright - i); ocmp = fref_code(rs, (int64_t)curr.left - i);
                if (ncmp < 0 || ocmp < 0) return false; ncmp = comp5(ncmp); ocmp = comp5(ocmp);
                const int rp = curr_seg_index * P.seglen + i; ps_ = (rp >= 0 && rp < read_len) ? fj_read_code(R, rp) : 4;     // prev_seq[len - (ad - i)], len == ad
              } else {
                ncmp = fref_code(rs, (int64_t)J.right + i); ocmp = fref_code(rs, (int64_t)J.left + 1 + i);
                if (ncmp < 0 || ocmp < 0) return false;
                ps_ = fh_seq_code(prev, R, plen_seq - (ad - i));
              }
              if (ps_ != ncmp) ++new_mm;
              if (ps_ != ocmp) ++old_mm;
            }
          }
          const int temp = new_mm - old_mm;
          if (temp >= new_diff || new_mm >= 2) continue;
          new_diff = temp; new_left = prev.left;
          nc = prev;
          { const int nlb = reversed ? prml - dtl : prml + dtl; if (nlb > 0) nc.ops[nc.n - 1] = mkop(opc(nc.ops[nc.n - 1]), (uint32_t)nlb); else --nc.n; }
          const uint32_t glen = J.right - J.left - 1u;
          if (glen <= (uint32_t)P.max_del) { if (!fh_push(nc, mkop(reversed ? OP_dEL : OP_DEL, glen))) { cig_ovf = true; return false; } antisense_closure = ps ? prev.asplice : curr.asplice; }
          else { if (!fh_push(nc, mkop(reversed ? OP_rEF_SKIP : OP_REF_SKIP, glen))) { cig_ovf = true; return false; } antisense_closure = J.antisense != 0; }
          const int nrf = reversed ? clml + dtr : clml - dtr;
          for (int x = nrf > 0 ? 0 : 1; x < curr.n; ++x) if (!fh_push(nc, x == 0 ? mkop(opc(curr.ops[0]), (uint32_t)nrf) : curr.ops[x])) { cig_ovf = true; return false; }
          mismatch = new_diff; found = true;
        }
        if (!found) return false;
      } else if (!(dist_btw_two == 0 && strands_ok)) check_fusion = true;
    }
    if (check_fusion) {
      // ---- fusion closure (1596-1819)
      ++n_closures;
      uint32_t id1 = prev.ref2, id2 = curr.ref; uint32_t fl_ = (uint32_t)(pright - 4), fr_ = (uint32_t)(curr.left - 4);
      bool reversed = false;
      if (fusion_dir != FUS_FF && (id2 < id1 || (id1 == id2 && fl_ > fr_))) { reversed = true; uint32_t t = id1; id1 = id2; id2 = t; t = fl_; fl_ = fr_; fr_ = t; }
      uint32_t lo = 0, hi = S.n_fus;            // upper_bound(Fusion(id1, id2, fl_, fr_)) with dir = FUSION_FF
      while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (fus_greater(S.fus[mid], id1, id2, fl_, fr_, (uint32_t)FUS_FF)) hi = mid; else lo = mid + 1; }
      uint32_t it = lo; lo = 0; hi = S.n_fus;   // lower_bound(Fusion(id1, id2, fl_ + 8, fr_ + 8))
      while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (fus_less(S.fus[mid], id1, id2, fl_ + 8u, fr_ + 8u, (uint32_t)FUS_FF)) lo = mid + 1; else hi = mid; }
      const uint32_t ub = lo;
      FRef ra, rb; if (!fref_of(ref, prev.ref2, ra) || !fref_of(ref, curr.ref, rb)) return false;
      int new_diff = 0xff;
      const int plen_seq = fh_seq_len(prev);
      for (; it != ub && it < S.n_fus; ++it) {
        const FusionKey F = S.fus[it];
        const int lb_left = reversed ? (int)F.right : (int)F.left, lb_right = reversed ? (int)F.left : (int)F.right;
        const int dtl = fusion_dir == FUS_RF ? pright - lb_left + 1 : lb_left - pright + 1;
        const int dtr = fusion_dir == FUS_FR ? curr.left - lb_right : lb_right - curr.left;
        if (!(abs(dtl) <= 4 && abs(dtr) <= 4 && dtl == dtr)) continue;
        if (dtl > clml || -dtl > prml) continue;
        int new_mm = 0, old_mm = 0;
        if (dtl > 0) {
          for (int i = 0; i < dtl; ++i) {
            int ncmp, ocmp;
            if (fusion_dir == FUS_RF) { ncmp = fref_code(ra, (int64_t)pright - i); if (ncmp >= 0) ncmp = comp5(ncmp); }      // rc(ref[lb_left, pright + 1))
            else ncmp = fref_code(ra, (int64_t)pright + i);                                                               // ref[pright, lb_left + 1)
            if (fusion_dir == FUS_FR) { ocmp = fref_code(rb, (int64_t)curr.left - i); if (ocmp >= 0) ocmp = comp5(ocmp); }    // rc(ref2[lb_right + 1, curr.left + 1))
            else ocmp = fref_code(rb, (int64_t)curr.left + i);                                                            // ref2[curr.left, lb_right)
            if (ncmp < 0 || ocmp < 0) return false;
            int cs_;
            if (fusion_dir == FUS_FF || fusion_dir == FUS_RR) cs_ = fh_seq_code(curr, R, i);
            else { const int rp = curr_seg_index * P.seglen + i; cs_ = (rp >= 0 && rp < read_len) ? fj_read_code(R, rp) : 4; }
            if (cs_ != ncmp) ++new_mm;
            if (cs_ != ocmp) ++old_mm;
          }
        } else if (dtl < 0) {
          const int ad = -dtl;
          for (int i = 0; i < ad; ++i) {
            int ncmp, ocmp;
            if (fusion_dir == FUS_FR) { ncmp = fref_code(rb, (int64_t)lb_right - i); if (ncmp >= 0) ncmp = comp5(ncmp); }    // rc(ref2[curr.left + 1, lb_right + 1))
            else ncmp = fref_code(rb, (int64_t)lb_right + i);                                                             // ref2[lb_right, curr.left)
            if (fusion_dir == FUS_RF) { ocmp = fref_code(ra, (int64_t)lb_left - 1 - i); if (ocmp >= 0) ocmp = comp5(ocmp); }  // rc(ref[pright + 1, lb_left))
            else ocmp = fref_code(ra, (int64_t)lb_left + 1 + i);                                                          // ref[lb_left + 1, pright)
            if (ncmp < 0 || ocmp < 0) return false;
            int ps_;
            if (fusion_dir == FUS_FF || fusion_dir == FUS_RR) ps_ = fh_seq_code(prev, R, plen_seq - (ad - i));
            else { const int seg0 = (curr_seg_index - 1) * P.seglen; int sl = P.seglen; if (seg0 + sl > read_len) sl = read_len - seg0;
                   const int rp = seg0 + sl - (ad - i); ps_ = (rp >= 0 && rp < read_len) ? fj_read_code(R, rp) : 4; }
            if (ps_ != ncmp) ++new_mm;
            if (ps_ != ocmp) ++old_mm;
          }
        }
        const int temp = new_mm - old_mm;
        if (temp >= new_diff || new_mm >= 2) continue;
        new_diff = temp; new_left = prev.left;
        nc = prev;
        { const int nlb = prml + dtl; if (nlb > 0) nc.ops[nc.n - 1] = mkop(opc(nc.ops[nc.n - 1]), (uint32_t)nlb); else --nc.n; }
        if (!fh_push(nc, mkop(fusion_dir, (uint32_t)lb_right & 0x0fffffffu))) { cig_ovf = true; return false; }
        antisense_closure = ps ? prev.asplice : curr.asplice;
        const int nrf = clml - dtr;
        for (int x = nrf > 0 ? 0 : 1; x < curr.n; ++x) if (!fh_push(nc, x == 0 ? mkop(opc(curr.ops[0]), (uint32_t)nrf) : curr.ops[x])) { cig_ovf = true; return false; }
        mismatch = new_diff; found = true;
      }
      if (!found) { FJ_DBG("merge: no fusion closure k=%d dir=%d pright=%d cleft=%d reversed=%d range [%u,%u) n_fus=%u ids %u %u l %u r %u", k, fusion_dir, pright, curr.left, (int)reversed, it, ub, S.n_fus, id1, id2, fl_, fr_); return false; }
    }
    if (found) {
      // merged_hit (1822-1838)
      const int mm = (int)prev.mism + (int)curr.mism + mismatch;
      FHit m = nc; m.ref = prev.ref; m.ref2 = curr.ref2; m.left = new_left; m.anti = antisense; m.asplice = antisense_closure;
      m.mism = (uint8_t)mm; m.smm = (uint8_t)(prev.smm + curr.smm);
      m.np = 0;
      for (int x = 0; x < prev.np; ++x) { m.pstart[m.np] = prev.pstart[x]; m.plen[m.np] = prev.plen[x]; m.prc[m.np] = prev.prc[x]; ++m.np; }
      for (int x = 0; x < curr.np; ++x) { if (m.np >= FJ_PIECES) return false; m.pstart[m.np] = curr.pstart[x]; m.plen[m.np] = curr.plen[x]; m.prc[m.np] = curr.prc[x]; ++m.np; }
      prev = m;
    } else { fin[nfin++] = prev; prev = curr; }
  }
  fin[nfin++] = prev;
  // final pass (1888-1945)
  bool saw_as = false, saw_s = false; int num_mm = 0, num_smm = 0;
  out.n = 0;
  for (int s = 0; s < nfin; ++s) {
    const FHit& h = fin[s];
    num_mm += h.mism; num_smm += h.smm;
    if (fh_spliced(h)) { if (h.asplice) { if (saw_s) return false; saw_as = true; } else { if (saw_as) return false; saw_s = true; } }
    int b = 0;
    if (out.n > 0 && opc(out.ops[out.n - 1]) == opc(h.ops[0])) { out.ops[out.n - 1] = mkop(opc(h.ops[0]), opl(out.ops[out.n - 1]) + opl(h.ops[0])); b = 1; }
    for (; b < h.n; ++b) if (!fh_push(out, h.ops[b])) { cig_ovf = true; return false; }
  }
  out.ref = fin[0].ref; out.ref2 = fin[nfin - 1].ref2; out.left = left; out.anti = antisense; out.asplice = saw_as;
  out.mism = (uint8_t)num_mm; out.smm = (uint8_t)num_smm;
  if (fusion_dir == FUS_NOTHING || fusion_dir == FUS_FF || fusion_dir == FUS_RR) { out.np = seqh.np; for (int k = 0; k < seqh.np; ++k) { out.pstart[k] = seqh.pstart[k]; out.plen[k] = seqh.plen[k]; out.prc[k] = seqh.prc[k]; } }
  else { out.np = 1; out.pstart[0] = 0; out.plen[0] = (uint16_t)read_len; out.prc[0] = 0; }                       // new_hit.seq(read_seq), 1976-1979
  if (fh_do_reverse(out)) { FHit t; fh_reverse(out, t); out = t; }
  if (fusion_dir != FUS_NOTHING) out.anti = fh_seq_orientation(out, read_len) != 0;                              // 2007-2013: seq != read_seq
  if (fh_read_len(out) != old_read_length || !fh_editdist_consistent(ref, out, R)) { FJ_DBG("merge: final check failed len %d vs %d, n=%d left=%d ref %u-%u mism %d ops0 %u/%u", fh_read_len(out), old_read_length, out.n, out.left, out.ref, out.ref2, (int)out.mism, opl(out.ops[0]), (unsigned)opc(out.ops[0])); return false; }               // 2023-2035
  return true;
}

// merge_segment_chain (2101-2220): orient the chain, merge it, valid_hit.  `hits` = the DFS stack.
__device__ bool fj_merge_segment_chain(const RefView& ref, const JoinParams& P, const FJoinSets& S, const uint64_t* R, int read_len, const FHit* hits, int n,
                                       int fusion_dir, FHit& out, unsigned& n_closures, bool& cig_ovf)
{
  if (n > 1) {
    FHit chain[JMAXSEGS];
    if (fusion_dir == FUS_NOTHING || fusion_dir == FUS_FF || fusion_dir == FUS_RR) {
      if (hits[0].anti) for (int i = 0; i < n; ++i) chain[i] = hits[n - 1 - i]; else for (int i = 0; i < n; ++i) chain[i] = hits[i];
    } else {
      bool saw = false;
      for (int i = 0; i < n; ++i) {
        bool pushed = false;
        if (!saw && i > 0) {
          if (hits[i - 1].ref != hits[i].ref) saw = true;
          else if (hits[i - 1].anti != hits[i].anti) saw = true;
          else { const int dist = hits[i].anti ? hits[i - 1].left - fh_right(hits[i]) : hits[i].left - fh_right(hits[i - 1]);
                 if (dist >= P.max_report_intron || dist < -P.max_ins) saw = true; }
        }
        const int fo = fh_fusion_opcode(hits[i]);
        if (fo == FUS_NOTHING && ((fusion_dir == FUS_FR && saw) || (fusion_dir == FUS_RF && !saw)) && hits[i].left < fh_right(hits[i])) { fh_reverse(hits[i], chain[i]); pushed = true; }
        if (i > 0 && fo != FUS_NOTHING && hits[i].ref != hits[i - 1].ref) {
          if (pushed) return false;          // the reference would push the hit twice; such a chain never merges (lengths differ)
          fh_reverse(hits[i], chain[i]); pushed = true;
        }
        if (!saw && fo != FUS_NOTHING) saw = true;
        if (!pushed) chain[i] = hits[i];
      }
    }
    if (!fj_merge_chain(ref, P, S, R, read_len, chain, n, fusion_dir, out, n_closures, cig_ovf)) return false;
  } else {
    out = hits[0];
    if (fh_do_reverse(out)) { FHit t; fh_reverse(out, t); out = t; }
  }
  return fh_valid(P, out);
}

struct FJoinOut { thb_joined* rec; unsigned long long cap; unsigned long long* count; unsigned int* overflow; unsigned long long* counters; };

// Output record: thb_joined with, for fusion alignments, ops[THB_JOINED_MAX_OPS - 1] = ref_id2 and the flag bits
// THB_JOINED_FUSION (the alignment has a second contig) / THB_JOINED_SEQ_RC (the hit's own sequence is the reverse complement of
// the read: bowtie_sam_extra reads it, bwt_map.cpp:2467-2648).
__device__ void fj_emit(const FJoinOut& o, const JoinBatchView& bv, uint32_t bi, const FHit& h, int read_len, unsigned& n_emit)
{
  const int so = fh_seq_orientation(h, read_len);
  if (so == 2 || h.n > JMAXOPS - 1) { atomicOr(o.overflow, so == 2 ? 4u : 2u); return; }
  const unsigned long long slot = atomicAdd(o.count, 1ull);
  ++n_emit;
  if (slot >= o.cap) { atomicOr(o.overflow, 1u); return; }
  thb_joined j; memset(&j, 0, sizeof j);
  j.bundle = bi + bv.bundle_base; j.ref_id = h.ref; j.left = h.left; j.n_ops = (uint8_t)h.n;
  const bool fused = fh_fusion_opcode(h) != FUS_NOTHING;
  j.flags = (uint8_t)((h.anti ? THB_HIT_ANTISENSE : 0) | (h.asplice ? THB_JHIT_ANTISENSE_SPLICE : 0) | (fused ? THB_JOINED_FUSION : 0) | (so == 1 ? THB_JOINED_SEQ_RC : 0));
  j.mismatches = h.mism; j.edit_dist = (uint8_t)(h.mism + fh_gap_length(h)); j.splice_mms = h.smm;
  for (int k = 0; k < h.n; ++k) j.ops[k] = h.ops[k];
  j.ops[THB_JOINED_MAX_OPS - 1] = h.ref2;
  o.rec[slot] = j;
}

// dfs_seg_hits' pair rules (2244-2558) for one candidate: prev = copy of the stack's top, curr = copy of the candidate.
// Returns whether the pair is accepted; *dir_out = the fusion direction to pass down.
__device__ bool fj_pair(const JoinParams& P, FHit& bh_prev, FHit& bh, int fusion_dir, int& dir_out)
{
  FHit* prevHit = &bh_prev; FHit* currHit = &bh;
  const int pf = fh_fusion_opcode(*prevHit), cf = fh_fusion_opcode(*currHit);
  const bool prev_fused = pf != FUS_NOTHING, curr_fused = cf != FUS_NOTHING;
  const int num_fusions = (prev_fused ? 1 : 0) + (curr_fused ? 1 : 0);
  int dir = prev_fused ? pf : cf;
  if (num_fusions >= 2) return false;
  if (fusion_dir != FUS_NOTHING && curr_fused) return false;
  if (fusion_dir == FUS_FF || fusion_dir == FUS_RR)
    if ((currHit->anti && currHit->ref != prevHit->ref) || (!currHit->anti && currHit->ref != prevHit->ref2)) return false;
  if ((fusion_dir == FUS_FR || fusion_dir == FUS_RF) && prevHit->ref2 != currHit->ref) return false;
  if ((fusion_dir == FUS_FR && !currHit->anti) || (fusion_dir == FUS_RF && currHit->anti)) return false;
  FHit t;
  if (curr_fused && dir == FUS_RR) { fh_reverse(*currHit, t); *currHit = t; }
  if (fusion_dir == FUS_FR || fusion_dir == FUS_RF || (curr_fused && currHit->ref == currHit->ref2 && (dir == FUS_FR || dir == FUS_RF))) {
    if (curr_fused) { if ((dir == FUS_FR && currHit->anti) || (dir == FUS_RF && !currHit->anti)) { fh_reverse(*currHit, t); *currHit = t; } }
    else if (fusion_dir == FUS_FR && currHit->anti) { fh_reverse(*currHit, t); *currHit = t; }
  }
  else if ((num_fusions == 0 && prevHit->anti && currHit->anti && prevHit->ref == currHit->ref &&
            (prevHit->left <= fh_right(*currHit) + P.max_report_intron && prevHit->left + P.max_ins >= fh_right(*currHit))) ||
           (num_fusions == 1 && (dir == FUS_FF || dir == FUS_RR) && ((!prev_fused && prevHit->anti) || (!curr_fused && currHit->anti)))) {
    FHit* tmp = prevHit; prevHit = currHit; currHit = tmp;             // 2349-2361
  }
  else if (num_fusions == 0) {
    if (prevHit->ref2 == currHit->ref && prevHit->anti == currHit->anti) {
      const int dist = prevHit->anti ? prevHit->left - fh_right(*currHit) : currHit->left - fh_right(*prevHit);
      if (dist > P.max_report_intron || dist < -P.max_ins)
        dir = ((prevHit->anti && prevHit->left > currHit->left) || (!prevHit->anti && prevHit->left < currHit->left)) ? FUS_FF : FUS_RR;
    } else {
      if (prevHit->anti == currHit->anti) dir = ((prevHit->anti && prevHit->ref > currHit->ref) || (!prevHit->anti && prevHit->ref < currHit->ref)) ? FUS_FF : FUS_RR;
      else if (!prevHit->anti) dir = FUS_FR;
      else dir = FUS_RF;
      if (dir == FUS_FR) { fh_reverse(*currHit, t); *currHit = t; }
      else if (dir == FUS_RF) { fh_reverse(*prevHit, t); *prevHit = t; }
    }
  }
  if (num_fusions == 1 && dir != FUS_FF && dir != FUS_RR) {                // 2439-2508: orient the fused segment
    bool prev_rep = false, curr_rep = false;
    if (prev_fused) {
      if ((dir == FUS_FR && !currHit->anti) || (dir == FUS_RF && currHit->anti)) return false;
      if (prevHit->ref2 != currHit->ref) prev_rep = true;
      else if ((dir == FUS_FR && prevHit->anti) || (dir == FUS_RF && !prevHit->anti)) prev_rep = true;
    }
    if (curr_fused) {
      if ((dir == FUS_FR && prevHit->anti) || (dir == FUS_RF && !prevHit->anti)) return false;
      if (currHit->ref != prevHit->ref2) curr_rep = true;
    }
    if (prev_rep) { fh_reverse(*prevHit, t); *prevHit = t; }
    if (curr_rep) { fh_reverse(*currHit, t); *currHit = t; }
    prev_rep = curr_rep = false;
    if (prev_fused) { if (fh_fwd_right(*prevHit) != fh_fwd_left(*currHit)) curr_rep = true; }
    else { if (fh_fwd_right(*prevHit) != fh_fwd_left(*currHit)) prev_rep = true; }
    if (prev_rep) { fh_reverse(*prevHit, t); *prevHit = t; }
    if (curr_rep) { fh_reverse(*currHit, t); *currHit = t; }
  }
  const bool same_contig = prevHit->ref2 == currHit->ref;
  if (!same_contig && num_fusions > 0) return false;
  if (same_contig && num_fusions >= 1 && fh_anti2(*prevHit) != currHit->anti) return false;
  int dist = 0;
  if (same_contig) {
    int bh_l, back_right;
    if ((fusion_dir == FUS_FR || fusion_dir == FUS_RF || dir == FUS_FR || dir == FUS_RF) && fh_anti2(*prevHit)) { bh_l = fh_right(*prevHit) + 1; back_right = currHit->left + 1; }
    else { bh_l = currHit->left; back_right = fh_right(*prevHit); }
    dist = bh_l - back_right;
  }
  dir_out = dir == FUS_NOTHING ? fusion_dir : dir;
  return !same_contig ||
         (same_contig && num_fusions == 0 && dir != FUS_NOTHING && fusion_dir == FUS_NOTHING) ||
         (same_contig && dist <= P.max_report_intron && dist >= -P.max_ins && fh_fwd_right(*prevHit) == fh_fwd_left(*currHit));
}

// K-FJ: join_segments_for_read with --fusion-search, thread = read
__global__ void __launch_bounds__(64)
fusion_join_kernel(RefView ref, JoinParams P, FJoinSets S, JoinBatchView bv, FJoinOut o)
{
  unsigned n_leaves = 0, n_closures = 0, n_emit = 0;
  for (uint32_t bi = blockIdx.x * blockDim.x + threadIdx.x; bi < bv.n_bundles; bi += gridDim.x * blockDim.x) {
    const uint4 hdr = __ldg(reinterpret_cast<const uint4*>(bv.bundles + bi));
    const int read_len = (int)(hdr.z & 0xffffu); const int n = (int)((hdr.z >> 16) & 0xffu); const uint32_t ops_begin = hdr.w;
    if (n < 1 || n > JMAXSEGS) continue;
    uint32_t off[JMAXSEGS]; int cnt[JMAXSEGS];
    { uint32_t a = hdr.y; for (int s = 0; s < n; ++s) { cnt[s] = (int)__ldg(bv.seg_count + (size_t)bi * bv.n_segs + s); off[s] = a; a += (uint32_t)cnt[s]; } }
    bool skip = false;
    if (P.bowtie2) for (int s = 0; s < n; ++s) if (cnt[s] > P.max_seg_multihits) skip = true;          // 2624-2632
    if (skip) continue;
    uint64_t R[12];
    { const uint64_t* rd = bv.reads + (size_t)bi * 3 * bv.read_words; const int rw = (int)bv.read_words;
      for (int pl = 0; pl < 3; ++pl) for (int w = 0; w < 4; ++w) R[pl * 4 + w] = w < rw ? __ldg(rd + pl * rw + w) : 0ull; }
    auto seg_pos = [&](int s) { return s * P.seglen; };
    auto seg_len = [&](int s) { return s == n - 1 ? read_len - s * P.seglen : P.seglen; };
    FHit stack[JMAXSEGS], saved[JMAXSEGS]; int it[JMAXSEGS], fdir[JMAXSEGS];
    bool cig_ovf = false;
    for (int i0 = 0; i0 < cnt[0]; ++i0) {
      { FHit h0; fh_load(h0, bv, off[0] + (uint32_t)i0, ops_begin, seg_pos(0), seg_len(0), read_len);
        if (fh_fusion_opcode(h0) == FUS_RR) fh_reverse(h0, stack[0]); else stack[0] = h0; }                 // 2645-2648
      int num_try = 10000;
      auto leaf = [&](int fusion_dir) {
        --num_try; ++n_leaves;
        FHit outh;
        const bool okm = fj_merge_segment_chain(ref, P, S, R, read_len, stack, n, fusion_dir, outh, n_closures, cig_ovf);
        FJ_DBG("read %u leaf dir %d -> %s", __ldg(&bv.bundles[bi].read_id), fusion_dir, okm ? "merged" : "rejected");
        if (okm) fj_emit(o, bv, bi, outh, read_len, n_emit);
      };
      if (n == 1) { leaf(FUS_NOTHING); continue; }
      int L = 1; it[1] = 0; fdir[1] = FUS_NOTHING;
      while (L >= 1) {
        if (it[L] >= cnt[L]) {                                 // this level is exhausted: back to the parent
          --L;
          if (L >= 1) { stack[L - 1] = saved[L]; if (num_try <= 0) break; ++it[L]; }
          continue;
        }
        FHit bh, bh_prev = stack[L - 1];
        fh_load(bh, bv, off[L] + (uint32_t)it[L], ops_begin, seg_pos(L), seg_len(L), read_len);
        int ndir = FUS_NOTHING;
        if (!fj_pair(P, bh_prev, bh, fdir[L], ndir)) { FJ_DBG("read %u L=%d it=%d pair rejected (fdir %d)", __ldg(&bv.bundles[bi].read_id), L, it[L], fdir[L]); ++it[L]; continue; }
        FJ_DBG("read %u L=%d it=%d pair accepted ndir %d", __ldg(&bv.bundles[bi].read_id), L, it[L], ndir);
        saved[L] = stack[L - 1]; stack[L - 1] = bh_prev; stack[L] = bh;      // 2571-2577
        if (L == n - 1) {
          leaf(ndir);
          stack[L - 1] = saved[L];
          if (num_try <= 0) break;
          ++it[L];
        } else { fdir[L + 1] = ndir; ++L; it[L] = 0; }
      }
    }
    if (cig_ovf) atomicOr(o.overflow, 2u);
  }
  if (n_leaves) atomicAdd(o.counters + 0, (unsigned long long)n_leaves);
  if (n_closures) atomicAdd(o.counters + 1, (unsigned long long)n_closures);
  if (n_emit) atomicAdd(o.counters + 2, (unsigned long long)n_emit);
}

}  // namespace thb
