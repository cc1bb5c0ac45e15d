// flank_kernel.cuh -- the junction index between the two hot stages, as device passes over the resident reference image.
//
// Replaces, for the segments-vs-junction-contigs search (tophat.py:2546-2600, 3686-3741):
//   juncs_db      print_splice / print_insertion / print_fusion, src/juncs_db.cpp:72-229   -> flank_build_kernel
//   bowtie-build  (third party, suffix array + BWT of segment_juncs.fa)                    -> flank_seed_kernel + one radix sort
//   bowtie -v <segment_mismatches> -k <max_seg_multihits> -m <max_seg_multihits>           -> flank_match_kernel (+ filter / decode)
//
// A contig is at most 128 bases (two flanks of <= max_seg_len bases, an inserted sequence between them for `ins` contigs), held
// as three bit planes of CW words.  The search is seed-and-verify on the q-gram lemma: the first `smin` bases of a segment are cut
// into v+2 pieces; a placement with <= v mismatches leaves at least two pieces exact, so every placement is found from the
// (piece a, piece b) seed of its two FIRST exact pieces -- and is reported from that pair only, which makes the result a set without a
// de-duplication pass.  The index is a direct-address table per piece pair: bucket -> range of (contig, offset) entries.
// Everything is exact: seeds only propose, the verify step counts mismatches over the whole segment.
#pragma once
#include "../../include/tophat_b200.h"
#include "bitplanes.cuh"

namespace thb {

constexpr int FLANK_MAX_PIECES = 5;        // v <= 3
constexpr int FLANK_POS_BITS = 7;          // contig offsets < 128

struct FlankDesc {           // one contig: where its bases come from (32 bytes)
  uint64_t a_start;          // global coordinate of the left part's first base in the image
  uint64_t b_start;          // .. of the right part
  uint64_t ins_code;         // 2 bits per inserted base, base 0 in the low bits
  uint8_t  a_len, b_len, ins_len, flags;     // flags: 1 = left part reverse-complemented, 2 = right part
  uint32_t pad;
};

template <int CW> struct FlankSeq { uint64_t p0[CW], p1[CW], pn[CW]; uint64_t len; };      // 32 bytes (CW=1) / 56 -> padded below
template <> struct FlankSeq<2> { uint64_t p0[2], p1[2], pn[2]; uint64_t len; uint64_t pad; };   // 64 bytes

struct FlankIndexParams {
  int npieces, piece_len, npairs, smin;      // pieces cover bases [0, npieces*piece_len) of a segment
  int bbits;                                 // bucket bits per pair
  int hashed;                                // seed code wider than bbits: bucket = mix(code) >> (64 - bbits)
  int max_mm, max_hits, ref_n_mismatch;
  int fp_bases;                              // bases of every NON-seed piece kept in an index entry (2 bits each, <= 32 bits in all)
};

constexpr int FLANK_INLINE = 7;              // entries held in the 64-byte bucket record itself
struct FlankBucket { uint32_t count, overflow; uint64_t e[FLANK_INLINE]; };      // overflow: index of entry #7 in the sorted entry array

// The bases of the v pieces that are NOT the seed pair (a, b), fp_bases of each: piece i of them at bits [2*i*fb, 2*i*fb + fb) (plane 0)
// and the next fb bits (plane 1).  An index entry carries this fingerprint of its contig window, so a proposed placement whose
// non-seed pieces alone already differ from the read in more than v bases is rejected without touching the contig's planes.
__device__ __forceinline__ uint32_t flank_fingerprint(const FlankIndexParams& ip, uint64_t p0, uint64_t p1, int a, int b)
{
  const int fb = ip.fp_bases, pl = ip.piece_len; const uint64_t m = maskn(fb);
  uint32_t fp = 0; int k = 0;
  for (int i = 0; i < ip.npieces; ++i) {
    if (i == a || i == b) continue;
    fp |= (uint32_t)(((p0 >> (i * pl)) & m) | (((p1 >> (i * pl)) & m) << fb)) << (2 * k * fb);
    ++k;
  }
  return fp;
}
// mismatching bases between two fingerprints (+ the read's N bits of the same bases, gathered the same way into plane-0 positions)
__device__ __forceinline__ int flank_fp_mismatches(const FlankIndexParams& ip, uint32_t x, uint32_t nbits)
{
  const int fb = ip.fp_bases; const uint32_t m = (uint32_t)maskn(fb);
  int n = 0;
  for (int k = 0; k < ip.npieces - 2; ++k) n += __popc((((x >> (2 * k * fb)) | (x >> (2 * k * fb + fb))) & m) | ((nbits >> (2 * k * fb)) & m));
  return n;
}

__device__ __forceinline__ uint64_t flank_mix(uint64_t x)
{
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull; x ^= x >> 27; x *= 0x94d049bb133111ebull; x ^= x >> 31; return x;
}
// pair index of pieces (a < b) in lexicographic order
__device__ __forceinline__ int flank_pair_index(int a, int b, int np) { return a * np - a * (a + 1) / 2 + (b - a - 1); }

__device__ __forceinline__ uint32_t flank_bucket(const FlankIndexParams& ip, uint64_t p0, uint64_t p1, int a, int b)
{
  const int pl = ip.piece_len; const uint64_t m = maskn(pl);
  const uint64_t code = ((p0 >> (a * pl)) & m) | (((p1 >> (a * pl)) & m) << pl) | (((p0 >> (b * pl)) & m) << (2 * pl)) | (((p1 >> (b * pl)) & m) << (3 * pl));
  return ip.hashed ? (uint32_t)(flank_mix(code) >> (64 - ip.bbits)) : (uint32_t)code;
}

// ---- 128-bit plane helpers ---------------------------------------------------------------------------------------------
template <int CW> __device__ __forceinline__ void flank_put(uint64_t* w, int pos, uint64_t v)
{
  const int i = pos >> 6, sh = pos & 63;
  if (CW == 1) { w[0] |= v << sh; return; }
  w[i] |= v << sh;
  if (sh && i + 1 < CW) w[i + 1] |= v >> (64 - sh);
}
template <int CW> __device__ __forceinline__ uint64_t flank_get(const uint64_t* w, int pos, int n)
{
  if (CW == 1) return (w[0] >> pos) & maskn(n);
  const int i = pos >> 6, sh = pos & 63;
  const uint64_t lo = w[i], hi = (i + 1 < CW) ? w[i + 1] : 0ull;
  return shr128(lo, hi, sh) & maskn(n);
}

// ---- contig sequences from the image (juncs_db.cpp:72-229) ----------------------------------------------------------------
template <int CW>
__global__ void flank_build_kernel(RefView ref, const FlankDesc* __restrict__ desc, FlankSeq<CW>* __restrict__ seq, uint32_t n)
{
  for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
    const FlankDesc d = desc[c];
    FlankSeq<CW> s;
    for (int i = 0; i < CW; ++i) s.p0[i] = s.p1[i] = s.pn[i] = 0ull;
    P3 a = ref_fetch3(ref, d.a_start, d.a_len);
    if (d.flags & 1) a = revcomp(a, d.a_len);
    flank_put<CW>(s.p0, 0, a.p0); flank_put<CW>(s.p1, 0, a.p1); flank_put<CW>(s.pn, 0, a.pn);
    int pos = d.a_len;
    if (d.ins_len) {
      uint64_t i0 = 0, i1 = 0;
      for (int k = 0; k < d.ins_len; ++k) { const uint64_t cd = (d.ins_code >> (2 * k)) & 3ull; i0 |= (cd & 1ull) << k; i1 |= (cd >> 1) << k; }
      flank_put<CW>(s.p0, pos, i0); flank_put<CW>(s.p1, pos, i1);
      pos += d.ins_len;
    }
    P3 b = ref_fetch3(ref, d.b_start, d.b_len);
    if (d.flags & 2) b = revcomp(b, d.b_len);
    flank_put<CW>(s.p0, pos, b.p0); flank_put<CW>(s.p1, pos, b.p1); flank_put<CW>(s.pn, pos, b.pn);
    s.len = (uint64_t)(pos + d.b_len);
    seq[c] = s;
  }
}

// ---- seed entries: one per (contig, offset, piece pair) ---------------------------------------------------------------------
// entry_base[c] = number of offsets of the contigs before c; an offset o is indexed when a segment of smin bases fits at it.
// Entries whose seed pieces touch an 'N' of the contig get the key `npairs << bbits` (sorted to the end, never looked up).
template <int CW>
__global__ void flank_seed_kernel(const FlankSeq<CW>* __restrict__ seq, const uint64_t* __restrict__ entry_base, uint32_t n_contigs,
                                  FlankIndexParams ip, uint32_t* __restrict__ keys, uint64_t* __restrict__ vals)
{
  for (uint32_t c = blockIdx.x; c < n_contigs; c += gridDim.x) {
    const int len = (int)seq[c].len;
    const int noff = len - ip.smin + 1;
    if (noff <= 0) continue;
    const uint64_t base = entry_base[c] * (uint64_t)ip.npairs;
    for (int t = threadIdx.x; t < noff * ip.npairs; t += blockDim.x) {
      const int o = t / ip.npairs, pr = t - o * ip.npairs;
      int a = 0, rem = pr;
      while (rem >= ip.npieces - 1 - a) { rem -= ip.npieces - 1 - a; ++a; }
      const int b = a + 1 + rem;
      const int span = ip.npieces * ip.piece_len;
      const uint64_t w0 = flank_get<CW>(seq[c].p0, o, span), w1 = flank_get<CW>(seq[c].p1, o, span), wn = flank_get<CW>(seq[c].pn, o, span);
      const uint64_t pm = maskn(ip.piece_len);
      const bool has_n = (((wn >> (a * ip.piece_len)) | (wn >> (b * ip.piece_len))) & pm) != 0ull;
      keys[base + t] = has_n ? ((uint32_t)ip.npairs << ip.bbits) : (((uint32_t)pr << ip.bbits) | flank_bucket(ip, w0, w1, a, b));
      vals[base + t] = ((uint64_t)flank_fingerprint(ip, w0, w1, a, b) << 32) | (uint64_t)((c << FLANK_POS_BITS) | (uint32_t)o);
    }
  }
}

// start[k] = index of the first sorted entry with key >= k, for k in [0, n_keys]  (n_keys = npairs << bbits)
__global__ void flank_bucket_kernel(const uint32_t* __restrict__ keys, uint64_t n, uint32_t n_keys, uint32_t* __restrict__ start)
{
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t hi = (i < n) ? min(keys[i], n_keys) : n_keys;
    const int64_t lo = (i == 0) ? -1 : (int64_t)min(keys[i - 1], n_keys);
    for (int64_t k = lo + 1; k <= (int64_t)hi; ++k) start[k] = (uint32_t)i;
  }
}

// bucket records: the first FLANK_INLINE entries of a key inline, so that a lookup is ONE 64-byte gather (count + entries with their
// fingerprints) instead of three dependent ones (range, entries, contig planes)
__global__ void flank_fill_kernel(const uint32_t* __restrict__ start, const uint64_t* __restrict__ vals, uint32_t n_keys, FlankBucket* __restrict__ buckets)
{
  for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_keys; k += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t lo = start[k], hi = start[k + 1];
    FlankBucket b;
    b.count = hi - lo; b.overflow = lo + FLANK_INLINE;
    for (int i = 0; i < FLANK_INLINE; ++i) b.e[i] = (lo + i < hi) ? vals[lo + i] : 0ull;
    buckets[k] = b;
  }
}

// ---- the search ----------------------------------------------------------------------------------------------------------
struct FlankBatchView {
  const uint64_t* reads; uint32_t n_reads, read_words, n_segs;
  uint16_t seg_bounds[THB_MAX_SEGS + 1];
};
struct FlankOut {
  uint64_t* keys; uint32_t* mm;            // append buffer: sortable placement key, mismatches
  unsigned long long* count; uint64_t cap; // *count may pass cap: the batch is then repeated with a larger buffer
  uint32_t* per_seg;                       // [n_reads * n_segs] placements found so far (the -m rule)
  unsigned long long* n_verified;
};
// read(27) | seg(4) | contig(25) | pos(7) | antisense(1)
__device__ __forceinline__ uint64_t flank_hit_key(uint32_t read, int seg, uint32_t contig, int pos, int anti)
{
  return ((uint64_t)read << 37) | ((uint64_t)seg << 33) | ((uint64_t)contig << 8) | ((uint64_t)pos << 1) | (uint64_t)anti;
}

// lane = (read, segment, strand, piece pair) for the seed lookup: one 64-byte bucket record.  The proposed placements of the warp's
// 32 lookups are then checked as ONE flattened list, 32 per round (buckets differ in size by orders of magnitude: a lane that walked
// its own bucket left the other 31 idle -- 7 of 32 lanes active in the first version, profiles/r2m_ncu_table.md): fingerprint first,
// the contig's planes only for the survivors.
struct FlankTask { uint64_t p0, p1, pn; uint32_t read; uint32_t info; uint32_t fp, fpn; uint32_t overflow, pad; };
// info: seg | anti << 4 | a << 5 | b << 8 | s << 11

template <int CW>
__global__ void __launch_bounds__(256)
flank_match_kernel(const FlankSeq<CW>* __restrict__ seq, const FlankBucket* __restrict__ buckets, const uint64_t* __restrict__ vals,
                   FlankIndexParams ip, FlankBatchView bv, FlankOut o)
{
  __shared__ FlankTask tasks[8][32];
  __shared__ uint64_t inl[8][32][FLANK_INLINE];
  const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  const uint64_t total = (uint64_t)bv.n_reads * bv.n_segs * 2u * (uint32_t)ip.npairs;
  const int pl = ip.piece_len; const uint64_t pm = maskn(pl);
  unsigned long long verified = 0;
  // task t = ((read * n_segs + seg) * 2 + anti) * npairs + pair.  64-bit divisions cost ~100 instructions each: the warp keeps
  // (read, remainder) of its tile's first task incrementally and the lanes finish with 32-bit arithmetic on small numbers.
  const uint32_t per_read = bv.n_segs * 2u * (uint32_t)ip.npairs;
  const uint64_t stride = (uint64_t)gridDim.x * 256u;
  const uint32_t stride_reads = (uint32_t)(stride / per_read), stride_rem = (uint32_t)(stride % per_read);
  uint64_t tile = ((uint64_t)blockIdx.x * 8u + w) * 32u;
  uint32_t read0 = (uint32_t)(tile / per_read), rem0 = (uint32_t)(tile % per_read);
  for (; tile < total; tile += stride, read0 += stride_reads, rem0 += stride_rem) {
    if (rem0 >= per_read) { rem0 -= per_read; ++read0; }
    const uint64_t t = tile + lane;
    uint32_t cnt = 0;
    if (t < total) {
      const uint32_t off = rem0 + lane, dr = off / per_read, x = off - dr * per_read;
      const uint32_t read = read0 + dr;
      const uint32_t y = x / (uint32_t)ip.npairs;                 // (seg * 2 + anti)
      const int pr = (int)(x - y * (uint32_t)ip.npairs), anti = (int)(y & 1u), seg = (int)(y >> 1);
      const int s0 = bv.seg_bounds[seg], s = bv.seg_bounds[seg + 1] - s0;
      P3 q = read_slice(bv.reads + (uint64_t)read * 3u * bv.read_words, (int)bv.read_words, s0, s);
      if (anti) q = revcomp(q, s);
      int a = 0, rem = pr;
      while (rem >= ip.npieces - 1 - a) { rem -= ip.npieces - 1 - a; ++a; }
      const int b = a + 1 + rem;
      FlankTask tk; tk.p0 = q.p0; tk.p1 = q.p1; tk.pn = q.pn; tk.read = read; tk.overflow = 0; tk.pad = 0;
      tk.info = (uint32_t)seg | ((uint32_t)anti << 4) | ((uint32_t)a << 5) | ((uint32_t)b << 8) | ((uint32_t)s << 11);
      tk.fp = flank_fingerprint(ip, q.p0, q.p1, a, b); tk.fpn = flank_fingerprint(ip, q.pn, 0ull, a, b);
      if ((((q.pn >> (a * pl)) | (q.pn >> (b * pl))) & pm) == 0ull) {              // an N in a seed piece of the read never matches
        const uint32_t key = ((uint32_t)pr << ip.bbits) | flank_bucket(ip, q.p0, q.p1, a, b);
        const ulonglong2* bp = reinterpret_cast<const ulonglong2*>(buckets + key);
        const ulonglong2 h0 = __ldg(bp), h1 = __ldg(bp + 1), h2 = __ldg(bp + 2), h3 = __ldg(bp + 3);
        cnt = (uint32_t)h0.x; tk.overflow = (uint32_t)(h0.x >> 32);
        uint64_t* dst = inl[w][lane];
        dst[0] = h0.y; dst[1] = h1.x; dst[2] = h1.y; dst[3] = h2.x; dst[4] = h2.y; dst[5] = h3.x; dst[6] = h3.y;
      }
      tasks[w][lane] = tk;
    }
    uint32_t incl = cnt;
    for (int d = 1; d < 32; d <<= 1) { const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d); if ((int)lane >= d) incl += up; }
    const uint32_t excl = incl - cnt;
    const uint32_t all = __shfl_sync(0xffffffffu, incl, 31);
    __syncwarp();
    for (uint32_t j0 = 0; j0 < all; j0 += 32u) {
      const uint32_t j = j0 + lane;
      int owner = 0;                                                               // the last lane whose list starts at or before j
      for (int step = 16; step; step >>= 1) {
        const int cand = owner + step;
        const uint32_t e = __shfl_sync(0xffffffffu, excl, cand & 31);
        if (cand < 32 && e <= j) owner = cand;
      }
      const uint32_t oex = __shfl_sync(0xffffffffu, excl, owner), ocnt = __shfl_sync(0xffffffffu, cnt, owner);
      if (j >= all) continue;
      const FlankTask& tk = tasks[w][owner];
      const int seg = (int)(tk.info & 15u), anti = (int)((tk.info >> 4) & 1u), a = (int)((tk.info >> 5) & 7u), b = (int)((tk.info >> 8) & 7u), s = (int)(tk.info >> 11);
      uint32_t* pcnt = o.per_seg + (uint64_t)tk.read * bv.n_segs + seg;
      if (ocnt > 64u && *(volatile uint32_t*)pcnt > (uint32_t)ip.max_hits) continue;     // a big bucket of a segment already suppressed by -m
      const uint32_t idx = j - oex;
      const uint64_t ent = idx < (uint32_t)FLANK_INLINE ? inl[w][owner][idx] : __ldg(vals + tk.overflow + (idx - FLANK_INLINE));
      if (flank_fp_mismatches(ip, (uint32_t)(ent >> 32) ^ tk.fp, tk.fpn) > ip.max_mm) continue;
      const uint32_t v = (uint32_t)ent;
      const uint32_t c = v >> FLANK_POS_BITS; const int pos = (int)(v & ((1u << FLANK_POS_BITS) - 1u));
      const FlankSeq<CW>& cs = seq[c];
      if (pos + s > (int)cs.len) continue;
      ++verified;
      const uint64_t w0 = flank_get<CW>(cs.p0, pos, s), w1 = flank_get<CW>(cs.p1, pos, s), wn = flank_get<CW>(cs.pn, pos, s);
      if (wn != 0ull && !ip.ref_n_mismatch) continue;                              // bowtie 1: no placement over an ambiguous reference base
      const uint64_t mism = ((w0 ^ tk.p0) | (w1 ^ tk.p1) | tk.pn | wn) & maskn(s);
      const int nm = __popcll(mism);
      if (nm > ip.max_mm) continue;
      int first = -1, second = -1;                                                 // reported from the pair of its two first exact pieces only
      for (int i = 0; i < ip.npieces; ++i)
        if (((mism >> (i * pl)) & pm) == 0ull) { if (first < 0) first = i; else if (second < 0) second = i; }
      if (first != a || second != b) continue;
      atomicAdd(pcnt, 1u);
      const unsigned long long slot = atomicAdd(o.count, 1ull);
      if (slot < o.cap) { o.keys[slot] = flank_hit_key(tk.read, seg, c, pos, anti); o.mm[slot] = (uint32_t)nm; }
    }
    __syncwarp();
  }
  if (verified) atomicAdd(o.n_verified, verified);
}

// the -m rule: placements of a segment with more than max_hits of them are dropped (the append order is arbitrary, the sort follows)
__global__ void flank_filter_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ mm, uint64_t n, const uint32_t* __restrict__ per_seg,
                                    uint32_t n_segs, uint32_t max_hits, uint64_t* __restrict__ out_keys, uint32_t* __restrict__ out_mm,
                                    unsigned long long* out_count)
{
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t k = keys[i];
    const uint32_t read = (uint32_t)(k >> 37); const uint32_t seg = (uint32_t)((k >> 33) & 15u);
    if (per_seg[(uint64_t)read * n_segs + seg] > max_hits) continue;
    const unsigned long long slot = atomicAdd(out_count, 1ull);
    out_keys[slot] = k; out_mm[slot] = mm[i];
  }
}

struct FlankHitRec { uint32_t read, contig; uint8_t seg, pos, antisense, mismatches; };     // = thb_flank_hit

__global__ void flank_decode_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ mm, uint64_t n, FlankHitRec* __restrict__ out)
{
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t k = keys[i];
    FlankHitRec h;
    h.read = (uint32_t)(k >> 37); h.seg = (uint8_t)((k >> 33) & 15u); h.contig = (uint32_t)((k >> 8) & 0x1ffffffu);
    h.pos = (uint8_t)((k >> 1) & 127u); h.antisense = (uint8_t)(k & 1u); h.mismatches = (uint8_t)mm[i];
    out[i] = h;
  }
}

// ---- placements back on the genome: SplicedBAMHitFactory::get_hit_from_buf + spliceCigar for an un-gapped hit ------------------
// (bwt_map.cpp:1469-1770, 678-883).  A placement is `s` matched bases at offset pos of a contig; on the genome it is
//   junction / deletion   left = left_start + pos, x = left + 1 .. : xM <gap>N|D (s-x)M         (x bases on the left flank)
//   insertion             xM <len>I (s-x-len)M, mismatches on inserted bases are not counted
//   fusion                xM <right>F (s-x)M with the lower-case match on the side that is read leftwards, strand flipped for rf / rr
// and is dropped (n_ops = 0) where the reference drops it: the hit does not reach over the event, or starts at / behind it.
// splice_mms = mismatches within min_anchor_len of the gap (1675-1689); the arithmetic follows the contig NAME, as the reference's does.
struct FlankContigDev { uint32_t kind, ref_id, ref_id2, left_start, left, right, right_end, aux, length; char ins_seq[20]; };   // = thb_flank_contig
struct FlankJHit { uint32_t ref_id; int32_t left; uint8_t n_ops, flags, mismatches, splice_mms; uint32_t ops[THB_JHIT_MAX_OPS]; };   // = thb_jhit_full

template <int CW>
__global__ void flank_splice_kernel(const FlankSeq<CW>* __restrict__ seq, const FlankContigDev* __restrict__ cdesc, const uint64_t* __restrict__ keys,
                                    uint64_t n, FlankBatchView bv, int min_anchor_len, int ref_n_mismatch, FlankJHit* __restrict__ out)
{
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t k = keys[i];
    const uint32_t read = (uint32_t)(k >> 37); const int seg = (int)((k >> 33) & 15u); const uint32_t c = (uint32_t)((k >> 8) & 0x1ffffffu);
    const int pos = (int)((k >> 1) & 127u); const int anti = (int)(k & 1u);
    const int s0 = bv.seg_bounds[seg], s = bv.seg_bounds[seg + 1] - s0;
    P3 q = read_slice(bv.reads + (uint64_t)read * 3u * bv.read_words, (int)bv.read_words, s0, s);
    if (anti) q = revcomp(q, s);
    const FlankSeq<CW>& cs = seq[c];
    const uint64_t w0 = flank_get<CW>(cs.p0, pos, s), w1 = flank_get<CW>(cs.p1, pos, s), wn = flank_get<CW>(cs.pn, pos, s);
    const uint64_t mism = ((w0 ^ q.p0) | (w1 ^ q.p1) | q.pn | (ref_n_mismatch ? wn : 0ull)) & maskn(s);      // bit o: contig offset pos + o
    const int nm = __popcll(mism);
    const FlankContigDev T = cdesc[c];
    FlankJHit h;
    h.ref_id = T.ref_id; h.left = 0; h.n_ops = 0; h.flags = 0; h.mismatches = 0; h.splice_mms = 0;
    for (int j = 0; j < THB_JHIT_MAX_OPS; ++j) h.ops[j] = 0u;
    const int end = (seg == (int)bv.n_segs - 1) ? THB_HIT_END : 0;
    if (T.kind == THB_FLANK_INS) {
      const int left = (int)T.left_start + pos;
      const int at = (int)T.left + 1 - left, len = (int)T.aux, ev_end = at + len;
      if (left <= (int)T.left && s > ev_end) {
        const int smm = __popcll(mism & (maskn(ev_end) & ~maskn(at)));
        h.left = left; h.n_ops = 3;
        h.ops[0] = ((uint32_t)at << 4) | 1u; h.ops[1] = ((uint32_t)len << 4) | 3u; h.ops[2] = ((uint32_t)(s - ev_end) << 4) | 1u;
        h.flags = (uint8_t)((anti ? THB_HIT_ANTISENSE : 0) | end);
        h.mismatches = (uint8_t)(nm - smm); h.splice_mms = 0;                      // create_hit(.., splice_mms = 0) for insertions (1652-1664)
      }
    } else {
      const bool fusion = T.kind == THB_FLANK_FUS;
      const bool leftwards = fusion && (T.aux == 9u || T.aux == 10u);
      const int left = leftwards ? (int)T.left_start - pos : (int)T.left_start + pos;
      const int lsp = leftwards ? (int)T.left - 1 : (int)T.left + 1;
      const bool reach = leftwards ? left > lsp : left < lsp;
      int at = lsp - left; if (at < 0) at = -at;
      const int gap = fusion ? (int)T.right : (int)T.right - (int)T.left - 1;
      if (reach && at < s && gap > 0) {
        const int lo = at - min_anchor_len + 1 > 0 ? at - min_anchor_len + 1 : 0, hi = at + min_anchor_len < s ? at + min_anchor_len : s;
        const int smm = hi > lo ? __popcll(mism & (maskn(hi) & ~maskn(lo))) : 0;
        const uint32_t code = fusion ? T.aux : (T.kind == THB_FLANK_DEL ? 5u : 11u);
        const uint32_t before = (fusion && (T.aux == 9u || T.aux == 10u)) ? 2u : 1u, after = (fusion && (T.aux == 8u || T.aux == 10u)) ? 2u : 1u;
        h.left = left; h.n_ops = 3;
        h.ops[0] = ((uint32_t)at << 4) | before; h.ops[1] = ((uint32_t)gap << 4) | code; h.ops[2] = ((uint32_t)(s - at) << 4) | after;
        if (fusion) h.ops[THB_JHIT_MAX_OPS - 1] = T.ref_id2;
        const int anti_out = leftwards ? !anti : anti;
        h.flags = (uint8_t)((anti_out ? THB_HIT_ANTISENSE : 0) | end | ((T.kind == THB_FLANK_JUNC && T.aux) ? THB_JHIT_ANTISENSE_SPLICE : 0) |
                            (leftwards ? THB_JHIT_SEQ_FLIPPED : 0));
        h.mismatches = (uint8_t)nm; h.splice_mms = (uint8_t)smm;
      }
    }
    out[i] = h;
  }
}

}  // namespace thb
