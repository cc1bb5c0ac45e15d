/*
 * segjuncs_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's
 * segment_juncs per-read arithmetic (TopHat 2.1.2, /root/reference/src/segment_juncs.cpp).
 *
 * Nothing under tophat_b200/ may call this file.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline leg of bench.py load it, and only as the checker for the CUDA path.
 *
 * Parity pin: this restatement is checked (tests/test_oracle_vs_reference.py) against the
 * reference's own segment_juncs binary built from /root/reference/src by oracle/Makefile.ref, on
 * identical BAM/FASTA inputs, and against the committed outputs of that binary under
 * tests/golden/.  The reference ships no golden vectors for this boundary.
 *
 * The code deliberately follows the reference statement by statement on unpacked ASCII strings
 * (no bit tricks) so that it is an independent check of the bit-parallel CUDA kernels.  Every
 * function cites the reference lines it restates.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include "../include/tophat_b200.h"

/* ------------------------------------------------------------------------------------------- */
/* result containers                                                                            */

typedef struct { uint32_t ref_id, left, right, antisense; } orc_junc;
typedef struct { uint32_t ref_id, left, len; char seq[20]; uint64_t order; } orc_ins;
typedef struct { uint32_t ref1, ref2, left, right, dir, count, edit_dist; } orc_fus;

typedef struct {
  orc_junc* juncs; size_t n_juncs, c_juncs;
  orc_junc* dels;  size_t n_dels,  c_dels;
  orc_ins*  ins;   size_t n_ins,   c_ins;
  orc_fus*  fus;   size_t n_fus,   c_fus;
  uint64_t  order;          /* running insertion order (processing order = first wins)        */
  /* task counters (used by bench.py to compute algorithmic bytes) */
  uint64_t n_windows, n_indel_tasks, n_rescue_tasks, n_fusion_tasks, n_juncs_emitted;
} orc_results;

/* CigarOpCode values used for fusion directions (bwt_map.h:36-55) */
enum { FUSION_FF = 7, FUSION_FR = 8, FUSION_RF = 9, FUSION_RR = 10 };

#define PUSH(arr, n, c, T, v) do { if ((n) == (c)) { (c) = (c) ? 2 * (c) : 1024; \
    (arr) = (T*)realloc((arr), (c) * sizeof(T)); } (arr)[(n)++] = (v); } while (0)

/* ------------------------------------------------------------------------------------------- */
/* unpacking of the packed inputs into the ASCII the reference works on                        */

typedef struct {
  const thb_ref_image* img;
} orc_ref;

/* Dna5 character of global base g: what RefSequenceTable::Sequence holds (bwt_map.h:582). */
static char ref_char5(const orc_ref* r, uint64_t g)
{
  uint64_t b = g >> 6; unsigned j = (unsigned)(g & 63);
  if ((r->img->nmask[b] >> j) & 1) return 'N';
  unsigned c = (unsigned)((r->img->planes[2 * b] >> j) & 1) | ((unsigned)((r->img->planes[2 * b + 1] >> j) & 1) << 1);
  return "ACGT"[c];
}

static int ref_has_seq(const orc_ref* r, uint32_t ref_id)
{
  return ref_id >= 1 && ref_id <= r->img->n_contigs && r->img->contig_len[ref_id - 1] > 0;
}
static int64_t ref_len(const orc_ref* r, uint32_t ref_id) { return (int64_t)r->img->contig_len[ref_id - 1]; }

/* seqan::infix(*ref_str, b, e) -> Dna5String; out must hold e-b+1 chars.  Returns 0 if the range
 * leaves the contig (the reference reads past the end of its buffer there: undefined). */
static int ref_infix5(const orc_ref* r, uint32_t ref_id, int64_t b, int64_t e, char* out)
{
  if (b < 0 || e > ref_len(r, ref_id) || e < b) return 0;
  uint64_t g0 = r->img->contig_start[ref_id - 1];
  for (int64_t i = b; i < e; ++i) out[i - b] = ref_char5(r, g0 + (uint64_t)i);
  out[e - b] = 0;
  return 1;
}
/* Dna5 -> Dna conversion: N becomes A (value & 3; alphabet_residue.h:873-876) */
static void to_dna4(char* s) { for (; *s; ++s) if (*s == 'N') *s = 'A'; }

static void unpack_read(const thb_segjuncs_batch* b, uint32_t idx, uint32_t len, char* out)
{
  const uint64_t* p = b->reads + (size_t)idx * 3 * b->read_words;
  for (uint32_t i = 0; i < len; ++i) {
    uint32_t w = i >> 6, j = i & 63;
    if ((p[2 * b->read_words + w] >> j) & 1) { out[i] = 'N'; continue; }
    unsigned c = (unsigned)((p[w] >> j) & 1) | ((unsigned)((p[b->read_words + w] >> j) & 1) << 1);
    out[i] = "ACGT"[c];
  }
  out[len] = 0;
}

/* reads.cpp:191-207 reverse_complement: non-ACGT -> N.  (seqan::reverseComplement on the
 * String<char> copies used at segment_juncs.cpp:2884, 3084, 3403 agrees on the ACGTN alphabet.) */
static void revcomp(const char* in, int n, char* out)
{
  for (int i = 0; i < n; ++i) {
    char c = in[n - 1 - i], o;
    switch (c) { case 'A': o = 'T'; break; case 'T': o = 'A'; break;
                 case 'C': o = 'G'; break; case 'G': o = 'C'; break; default: o = 'N'; }
    out[i] = o;
  }
  out[n] = 0;
}

/* ------------------------------------------------------------------------------------------- */
/* hit vectors (vector<HitsForRead>)                                                            */

#define MAX_SEGS 16
typedef struct { thb_hit* h; int n, c; } hitvec;
static void hv_push(hitvec* v, thb_hit x)
{ if (v->n == v->c) { v->c = v->c ? 2 * v->c : 8; v->h = (thb_hit*)realloc(v->h, v->c * sizeof(thb_hit)); } v->h[v->n++] = x; }
static void hv_free(hitvec* v) { free(v->h); v->h = NULL; v->n = v->c = 0; }
static void hv_copy(hitvec* d, const hitvec* s) { d->n = 0; for (int i = 0; i < s->n; ++i) hv_push(d, s->h[i]); }

#define ANTI(h) (((h).flags & THB_HIT_ANTISENSE) != 0)
#define ISEND(h) (((h).flags & THB_HIT_END) != 0)

/* ------------------------------------------------------------------------------------------- */
/* simpleSplitAlignment, segment_juncs.cpp:2390-2456                                            */
/* positions[] receives every argmin in ascending order; returns their count.                   */
static int simple_split_alignment(const char* shorter, int n, const char* left_ref, const char* right_ref,
                                  int* positions, int* mismatch_count)
{
  int* before = (int*)malloc(sizeof(int) * (n > 0 ? n : 1));
  int* after = (int*)malloc(sizeof(int) * (n > 0 ? n : 1));
  for (int idx = n - 1; idx >= 0; --idx) {                       /* 2407-2421 */
    int prev = (idx < n - 1) ? before[idx + 1] : 0;
    int mm = (right_ref[idx] == 'N' || shorter[idx] == 'N' || right_ref[idx] != shorter[idx]) ? 1 : 0;
    before[idx] = prev + mm;
  }
  for (int idx = 0; idx < n; ++idx) {                            /* 2424-2434 */
    int prev = idx > 0 ? after[idx - 1] : 0;
    int mm = (left_ref[idx] == 'N' || shorter[idx] == 'N' || left_ref[idx] != shorter[idx]) ? 1 : 0;
    after[idx] = prev + mm;
  }
  *mismatch_count = n + 1;                                       /* 2436 */
  int np = 0;
  for (int p = 1; p < n; ++p) {                                  /* 2443-2454 */
    int e = before[p] + after[p - 1];
    if (e < *mismatch_count) { *mismatch_count = e; np = 0; positions[np++] = p; }
    else if (e == *mismatch_count) positions[np++] = p;
  }
  free(before); free(after);
  return np;
}

/* ------------------------------------------------------------------------------------------- */
/* detect_small_insertion, segment_juncs.cpp:2470-2541 (base space; color is out of scope)      */
static void detect_small_insertion(const thb_params* P, const orc_ref* rt, const char* read, int read_length,
                                   const thb_hit* L, const thb_hit* R, orc_results* out)
{
  (void)P;
  if (!ref_has_seq(rt, L->ref_id)) return;                       /* 2476-2478 */
  if (L->left < 0) return;                                       /* 2491 */
  int discrepancy = read_length - (R->right - L->left);         /* 2498 */
  int glen = R->right - L->left;
  if (glen < 0 || glen > 4096) return;
  char genomic[4097];
  if (!ref_infix5(rt, L->ref_id, L->left, R->right, genomic)) return; /* 2499 (DnaString: N->A) */
  to_dna4(genomic);
  /* 2506-2507: left/right read slices of the genomic length */
  if (glen > read_length) return;
  const char* left_read = read;
  const char* right_read = read + (read_length - glen);
  int positions[4097]; int min_errors = -1;
  int np = simple_split_alignment(genomic, glen, left_read, right_read, positions, &min_errors); /* 2511 */
  out->n_indel_tasks++;
  if (np <= 0) return;                                           /* 2513 */
  int best = positions[0];                                       /* 2516 */
  int adjustment = 0;
  if ((int)L->read_len + (int)R->read_len >= read_length) adjustment = -1;        /* 2527-2529 */
  if (min_errors <= ((int)L->edit_dist + (int)R->edit_dist + adjustment) &&
      best + discrepancy <= glen) {                              /* 2530-2531 */
    orc_ins ins; memset(&ins, 0, sizeof ins);
    ins.ref_id = L->ref_id; ins.left = (uint32_t)(L->left + best - 1);            /* 2535-2537 */
    ins.len = (uint32_t)discrepancy;
    if (discrepancy > (int)sizeof(ins.seq) - 1) return;
    memcpy(ins.seq, left_read + best, (size_t)discrepancy);      /* 2532 */
    ins.order = out->order++;
    PUSH(out->ins, out->n_ins, out->c_ins, orc_ins, ins);
  }
}

/* detect_small_deletion, segment_juncs.cpp:2554-2627 */
static void detect_small_deletion(const thb_params* P, const orc_ref* rt, const char* read, int read_length,
                                  const thb_hit* L, const thb_hit* R, orc_results* out)
{
  (void)P;
  if (!ref_has_seq(rt, L->ref_id)) return;
  if (L->left < 0) return;                                       /* 2574 */
  if (R->right < read_length) return;                            /* 2578 */
  int discrepancy = (R->right - L->left) - read_length;         /* 2581 */
  char lg[4097], rg[4097];
  if (read_length > 4096) return;
  if (!ref_infix5(rt, L->ref_id, L->left, L->left + read_length, lg)) return;     /* 2582 (Dna5: N kept) */
  if (!ref_infix5(rt, L->ref_id, R->right - read_length, R->right, rg)) return;   /* 2583 */
  int positions[4097]; int min_errors = -1;
  int np = simple_split_alignment(read, read_length, lg, rg, positions, &min_errors); /* 2601 */
  out->n_indel_tasks++;
  if (np <= 0) return;
  int best = positions[0];                                       /* 2604 */
  int adjustment = 0;
  if ((int)L->read_len + (int)R->read_len >= read_length) adjustment = -1;        /* 2616-2618 */
  if (min_errors <= ((int)L->edit_dist + (int)R->edit_dist + adjustment)) {      /* 2619 */
    orc_junc d = { L->ref_id, (uint32_t)(L->left + best - 1), (uint32_t)(L->left + best + discrepancy), 0 }; /* 2620-2623 */
    PUSH(out->dels, out->n_dels, out->c_dels, orc_junc, d);
  }
}

/* find_insertions_and_deletions, segment_juncs.cpp:2807-2942 */
static void find_insertions_and_deletions(const thb_params* P, const orc_ref* rt, const char* read_seq, int read_len,
                                          hitvec* H, int nsegs, orc_results* out)
{
  if (nsegs <= 0) return;                                        /* 2812 */
  if (nsegs - 1 == 0) return;                                    /* 2815-2818 */
  const int seglen = P->segment_length;
  for (int i = 0; i + 2 < nsegs; ++i) {                          /* 2856: i < size-2 */
    hitvec* ls = &H[i]; hitvec* rs = &H[i + 1];
    if (ls->n == 0 || rs->n == 0) return;                        /* 2870-2871 */
    /* 2882-2884: fullRead = read.seq.substr(i*seglen, 2*seglen); rcRead = revcomp */
    int start = i * seglen;
    if (start > read_len) return;                                /* substr would throw */
    int plen = 2 * seglen; if (start + plen > read_len) plen = read_len - start;
    char full[512], rc[512];
    memcpy(full, read_seq + start, (size_t)plen); full[plen] = 0;
    revcomp(full, plen, rc);
    for (int li = 0; li < ls->n; ++li)
      for (int ri = 0; ri < rs->n; ++ri) {
        const thb_hit* lh = &ls->h[li]; const thb_hit* rh = &rs->h[ri];
        if (lh->ref_id != rh->ref_id) continue;                  /* 2901 */
        if (ANTI(*lh) != ANTI(*rh)) continue;                    /* 2904 */
        const char* mod = full;
        if (ANTI(*lh)) { const thb_hit* t = lh; lh = rh; rh = t; mod = rc; }      /* 2914-2920 */
        int apparent = rh->right - lh->left;                     /* 2922 */
        int disc = apparent - plen;                              /* 2923 */
        if (disc > 0 && disc <= P->max_deletion_length)          /* 2924 */
          detect_small_deletion(P, rt, mod, plen, lh, rh, out);
        if (disc < 0 && disc >= -P->max_insertion_length)        /* 2932 */
          detect_small_insertion(P, rt, mod, plen, lh, rh, out);
      }
  }
}

/* map_read_to_contig, segment_juncs.cpp:2946-2973 */
static int map_read_to_contig(const char* contig, int contig_len, const char* read, int read_len)
{
  int pos = -1, mismatch = 3;
  for (int i = 0; i < contig_len - read_len; ++i) {
    int t = 0;
    for (int j = 0; j < read_len; ++j) {
      if (contig[i + j] != read[j]) ++t;
      if (t >= mismatch) break;
    }
    if (t < mismatch) { pos = i; mismatch = t; }
  }
  return pos;
}

/* The mate-flank rescue shared by find_gaps (3406-3491) and find_fusions (3123-3211).
 * Returns 0 when the reference `break`s out of the partner loop. */
static int rescue_in_flank(const thb_params* P, const orc_ref* rt, const char* read, const char* rcread, int read_length,
                           const thb_hit* rightHit, hitvec* dest, orc_results* out)
{
  if (!ref_has_seq(rt, rightHit->ref_id)) return 1;
  const int part_seq_len = P->inner_dist_std_dev > P->inner_dist_mean ? P->inner_dist_std_dev - P->inner_dist_mean : 0; /* 3425 */
  const int flanking_seq_len = P->inner_dist_mean + P->inner_dist_std_dev;                                              /* 3426 */
  int64_t left = 0;
  if (ANTI(*rightHit)) {                                         /* 3430-3439 */
    if (flanking_seq_len <= rightHit->left) left = rightHit->left - flanking_seq_len; else return 0;
  } else {                                                       /* 3440-3449 */
    if (part_seq_len <= rightHit->right) left = rightHit->right - part_seq_len; else return 0;
  }
  int clen = flanking_seq_len + part_seq_len;
  if (clen <= 0 || clen > 8192) return 1;
  char contig[8193];
  if (!ref_infix5(rt, rightHit->ref_id, left, left + clen, contig)) return 1;    /* past the contig end: undefined in the reference */
  int check_read_len = P->segment_length - P->segment_mismatches - 3; if (check_read_len > 15) check_read_len = 15; /* 3451 */
  if (check_read_len <= 0 || check_read_len > read_length) return 1;
  const char* fwd_read = read + (read_length - check_read_len);  /* 3452 */
  const char* rev_read = rcread;                                 /* 3453 */
  out->n_rescue_tasks++;
  int fwd_pos = map_read_to_contig(contig, clen, fwd_read, check_read_len);      /* 3455 */
  if (fwd_pos >= 0) {                                            /* 3456-3462 */
    thb_hit h; memset(&h, 0, sizeof h);
    h.ref_id = rightHit->ref_id; h.left = (int32_t)(left + fwd_pos); h.right = h.left + check_read_len;
    h.read_len = (uint8_t)check_read_len; h.edit_dist = 0; h.flags = THB_HIT_END;
    hv_push(dest, h);
  }
  int rev_pos = map_read_to_contig(contig, clen, rev_read, check_read_len);      /* 3464 */
  if (rev_pos >= 0) {                                            /* 3466-3472 */
    thb_hit h; memset(&h, 0, sizeof h);
    h.ref_id = rightHit->ref_id; h.left = (int32_t)(left + rev_pos); h.right = h.left + check_read_len;
    h.read_len = (uint8_t)check_read_len; h.edit_dist = 0; h.flags = THB_HIT_END | THB_HIT_ANTISENSE;
    hv_push(dest, h);
  }
  return 1;
}

/* ------------------------------------------------------------------------------------------- */
/* juncs_from_ref_segs<RecordSegmentJuncs> for one POINT_DIR_BOTH window and one motif,
 * segment_juncs.cpp:2097-2289 + RecordSegmentJuncs::record 1669-1696.                          */
typedef struct { uint32_t ref_id; int antisense; int right_mate; int64_t left, right; char support[160]; int slen; } refseg;

static void scan_window(const thb_params* P, const orc_ref* rt, const refseg* seg,
                        const char* donor, const char* acceptor, orc_results* out)
{
  if (!ref_has_seq(rt, seg->ref_id)) return;                     /* 2104-2107 */
  int skip_fwd = 0, skip_rev = 0;                                /* 2109-2138 */
  if (P->library_type == 2) {            /* FR_FIRSTSTRAND */
    if (!seg->right_mate) { if (seg->antisense) skip_rev = 1; else skip_fwd = 1; }
    else                  { if (seg->antisense) skip_fwd = 1; else skip_rev = 1; }
  }
  if (P->library_type == 3) {            /* FR_SECONDSTRAND */
    if (!seg->right_mate) { if (seg->antisense) skip_fwd = 1; else skip_rev = 1; }
    else                  { if (seg->antisense) skip_rev = 1; else skip_fwd = 1; }
  }
  if (seg->left < 0 || seg->right >= ref_len(rt, seg->ref_id) - 1) return;       /* 2154 */
  int64_t seg_len = seg->right - seg->left;                      /* 2172 */
  int read_len = seg->slen;
  if (seg_len < read_len + 2 || read_len < 2) return;            /* cannot happen for find_gaps windows (dist >= 50) */
  /* rev_donor / rev_acceptor dinucleotides (2071-2074) */
  char rev_donor[3], rev_acceptor[3];
  revcomp(donor, 2, rev_donor); revcomp(acceptor, 2, rev_acceptor);
  /* only the two window ends are ever read: [0, read_len+1) and [seg_len-read_len-2, seg_len) */
  char lw[200], rw[200];
  int nl = read_len + 2;
  ref_infix5(rt, seg->ref_id, seg->left, seg->left + nl, lw); to_dna4(lw);      /* 2157: DnaString (N->A) */
  ref_infix5(rt, seg->ref_id, seg->right - nl, seg->right, rw); to_dna4(rw);
  #define W(idx) (((idx) < nl) ? lw[(idx)] : rw[(idx) - (seg_len - nl)])
  int to = read_len - 2;                                         /* 2177 */
  uint8_t left_mm[160] = {0}, right_mm[160] = {0};               /* 2180-2181 */
  { int n = 0;                                                   /* 2193-2205 */
    for (int i = 0; i < read_len - 1; ++i) {
      if (W(i) != seg->support[i]) ++n;
      left_mm[i] = (uint8_t)n;
      if (n > 2) { to = i; break; }
    } }
  { int n = 0;                                                   /* 2210-2219 */
    for (int i = read_len - 1; i >= 0; --i) {
      if (W(i + (seg_len - read_len)) != seg->support[i]) ++n;
      right_mm[i] = (uint8_t)n;
      if (n > 2) break;
    } }
  out->n_windows++;
  for (int i = 0; i <= to; ++i) {                                /* 2242 */
    char curr[2] = { W(i), W(i + 1) };
    int is_donor = (curr[0] == donor[0] && curr[1] == donor[1]);
    int is_racc = (curr[0] == rev_acceptor[0] && curr[1] == rev_acceptor[1]);
    if ((!skip_fwd && is_donor) || (!skip_rev && is_racc)) {     /* 2247 */
      const char* partner = is_donor ? acceptor : rev_donor;     /* 2249-2253 */
      int lm = i > 0 ? left_mm[i - 1] : 0;                       /* 2255-2257 */
      if (lm + right_mm[i] <= 2) {                               /* 2265 */
        int64_t pos = seg_len - (read_len - i) - 2;              /* 2267 */
        if (partner[0] == W(pos) && partner[1] == W(pos + 1)) {  /* 2268 */
          /* 2270-2279 + RecordSegmentJuncs::record 1686-1691 */
          orc_junc j = { seg->ref_id, (uint32_t)(seg->left + i - 1), (uint32_t)(seg->left + pos + 2), is_donor ? 0u : 1u };
          PUSH(out->juncs, out->n_juncs, out->c_juncs, orc_junc, j);
          out->n_juncs_emitted++;
        }
      }
    }
  }
  #undef W
}

/* find_gaps, segment_juncs.cpp:3293-3650.  H is mutated exactly as the reference mutates
 * hits_for_read (resize at 3313, clear at 3395-3398, rescue push at 3461/3471). *nsegs is updated. */
static void find_gaps(const thb_params* P, const orc_ref* rt, const char* read_seq, int read_len,
                      hitvec* H, int* nsegs, const thb_hit* partner, int n_partner, int right_mate, orc_results* out)
{
  if (*nsegs <= 0) return;                                       /* 3301 */
  int last = *nsegs - 1;
  while (last > 0) { if (H[last].n) break; --last; }             /* 3304-3311 */
  for (int s = last + 1; s < *nsegs; ++s) H[s].n = 0;
  *nsegs = last + 1;                                             /* 3313 */
  if (last == 0 && (H[0].n == 0 || ISEND(H[0].h[0]))) return;    /* 3316-3318 */
  int has_partner = n_partner > 0;                               /* 3322-3344 (host lookup) */
  hitvec right_copy = {0, 0, 0};                                 /* 3354-3356 */
  if (last != 0) hv_copy(&right_copy, &H[last]);
  int check_partner = 1;
  if (last != 0) {                                               /* 3362-3390 */
    for (int i = 0; i < H[0].n && check_partner; ++i)
      for (int j = 0; j < right_copy.n; ++j) {
        const thb_hit* l = &H[0].h[i]; const thb_hit* r = &right_copy.h[j];
        if (l->ref_id == r->ref_id && ANTI(*l) == ANTI(*r)) {
          int dist = ANTI(*l) ? l->left - r->right : r->left - l->right;
          if (dist >= P->min_segment_intron_length && dist < P->max_segment_intron_length) { check_partner = 0; break; }
        }
      }
  }
  hv_free(&right_copy);
  if (check_partner && has_partner) {                            /* 3392 */
    for (int i = 1; i < *nsegs; ++i) H[i].n = 0;                 /* 3395-3398 */
    char rc[512]; revcomp(read_seq, read_len, rc);               /* 3400-3403 */
    int nleft = H[0].n;    /* last >= 1 on every call path, so H[0] is not the push target */
    for (int l = 0; l < (last == 0 ? H[0].n : nleft); ++l) {     /* 3406 */
      for (int r = 0; r < n_partner; ++r) {                      /* 3409 */
        const thb_hit* leftHit = &H[0].h[l]; const thb_hit* rightHit = &partner[r];
        if (leftHit->ref_id != rightHit->ref_id || ANTI(*leftHit) == ANTI(*rightHit)) continue; /* 3412 */
        /* 3421: `dist < min && dist >= max` is never true */
        if (!rescue_in_flank(P, rt, read_seq, rc, read_len, rightHit, &H[last], out)) break;   /* 3424-3472 */
      }
    }
  }
  if (P->bowtie2)                                                /* 3499-3506 */
    for (int s = 0; s < *nsegs; ++s) if (H[s].n > P->max_seg_multihits) return;

  const int seglen = P->segment_length;
  refseg* wins = NULL; size_t nw = 0, cw = 0;
  for (int s = 0; s < *nsegs; ++s) {                             /* 3508 */
    for (int h = 0; h < H[s].n; ++h) {
      int found = (s == *nsegs - 1);                             /* 3513 */
      const thb_hit* bh = &H[s].h[h];
      const thb_hit* drs[4096]; int ndrs = 0; const thb_hit* rrs[4096]; int nrrs = 0;
      if (s < *nsegs - 1) {                                      /* 3521-3548 */
        for (int r = 0; r < H[s + 1].n; ++r) {
          const thb_hit* rh = &H[s + 1].h[r];
          if (ANTI(*bh) != ANTI(*rh) || bh->ref_id != rh->ref_id) continue;
          if ((ANTI(*bh) && rh->right == bh->left) || (!ANTI(*bh) && bh->right == rh->left)) { found = 1; break; }
          int dist = ANTI(*bh) ? bh->left - rh->right : rh->left - bh->right;
          if (dist >= P->min_segment_intron_length && dist < P->max_segment_intron_length && ndrs < 4096) drs[ndrs++] = rh;
        }
      }
      if (!found && s < *nsegs - 2) {                            /* 3550-3570 */
        for (int r = 0; r < H[s + 2].n; ++r) {
          const thb_hit* rrh = &H[s + 2].h[r];
          if (ANTI(*bh) != ANTI(*rrh) || bh->ref_id != rrh->ref_id) continue;
          int dist = ANTI(*bh) ? bh->left - rrh->right : rrh->left - bh->right;
          if (dist >= P->min_segment_intron_length + seglen && dist < P->max_segment_intron_length + seglen && nrrs < 4096) rrs[nrrs++] = rrh;
        }
      }
      if (!found && (ndrs > 0 || nrrs > 0)) {                    /* 3572 */
        const int look_bp = 8;
        const thb_hit** d = nrrs > 0 ? rrs : drs; int nd = nrrs > 0 ? nrrs : ndrs;   /* 3577 */
        for (int r = 0; r < nd; ++r) {
          refseg w; memset(&w, 0, sizeof w);
          int start = (s + 1) * seglen - look_bp;                /* 3582/3584 */
          int want = nrrs <= 0 ? look_bp * 2 : seglen + look_bp * 2;
          if (start > read_len || start < 0) continue;           /* substr would throw */
          int sl = want; if (start + sl > read_len) sl = read_len - start;
          w.ref_id = bh->ref_id; w.antisense = ANTI(*bh); w.right_mate = right_mate; w.slen = sl;
          if (!ANTI(*bh)) {                                      /* 3587-3593 */
            memcpy(w.support, read_seq + start, (size_t)sl); w.support[sl] = 0;
            w.left = bh->right - look_bp; if (w.left < 0) w.left = 0;
            w.right = d[r]->left + look_bp;
          } else {                                               /* 3594-3605 */
            char tmp[160]; memcpy(tmp, read_seq + start, (size_t)sl); tmp[sl] = 0;
            revcomp(tmp, sl, w.support);
            w.left = d[r]->right - look_bp;
            w.right = bh->left + look_bp;
          }
          PUSH(wins, nw, cw, refseg, w);
        }
      }
    }
  }
  /* 3618-3649: three motif passes over the same window list */
  static const char* motifs[3][2] = { {"GT", "AG"}, {"GC", "AG"}, {"AT", "AC"} };
  for (int m = 0; m < 3; ++m)
    for (size_t r = 0; r < nw; ++r) scan_window(P, rt, &wins[r], motifs[m][0], motifs[m][1], out);
  out->n_windows -= 2 * nw;      /* count each window once, not once per motif */
  free(wins);
}

/* ------------------------------------------------------------------------------------------- */
/* detect_fusion, segment_juncs.cpp:2629-2805 */
static void add_fusion(orc_results* out, uint32_t r1, uint32_t r2, uint32_t left, uint32_t right, uint32_t dir, uint32_t ed)
{
  for (size_t i = 0; i < out->n_fus; ++i) {                      /* 2788-2803 (map find) */
    orc_fus* f = &out->fus[i];
    if (f->ref1 == r1 && f->ref2 == r2 && f->left == left && f->right == right && f->dir == dir) {
      f->count += 1; if (ed < f->edit_dist) f->edit_dist = ed; return;
    }
  }
  orc_fus f = { r1, r2, left, right, dir, 1, ed };
  PUSH(out->fus, out->n_fus, out->c_fus, orc_fus, f);
}

static void detect_fusion(const thb_params* P, const orc_ref* rt, const char* read, int read_length,
                          const thb_hit* L, const thb_hit* R, uint32_t dir, orc_results* out)
{
  if (!ref_has_seq(rt, L->ref_id) || !ref_has_seq(rt, R->ref_id)) return;
  if (read_length > 400) return;
  char lg[512], rg[512], tmp[512];
  if (dir == FUSION_FF || dir == FUSION_FR) {                    /* 2644-2650 */
    if (L->left + read_length > ref_len(rt, L->ref_id)) return;
    if (!ref_infix5(rt, L->ref_id, L->left, L->left + read_length, lg)) return;
  } else {                                                       /* 2651-2658 */
    if (L->right < read_length) return;
    if (!ref_infix5(rt, L->ref_id, L->right - read_length, L->right, tmp)) return;
    revcomp(tmp, read_length, lg);
  }
  if (dir == FUSION_FF || dir == FUSION_RF) {                    /* 2660-2666 */
    if (R->right < read_length) return;
    if (!ref_infix5(rt, R->ref_id, R->right - read_length, R->right, rg)) return;
  } else {                                                       /* 2667-2674 */
    if (R->left + read_length > ref_len(rt, R->ref_id)) return;
    if (!ref_infix5(rt, R->ref_id, R->left, R->left + read_length, tmp)) return;
    revcomp(tmp, read_length, rg);
  }
  int positions[512]; int min_errors = -1;
  int np = simple_split_alignment(read, read_length, lg, rg, positions, &min_errors);  /* 2686 */
  out->n_fusion_tasks++;
  uint32_t total_edit_dist = (uint32_t)L->edit_dist + (uint32_t)R->edit_dist;          /* 2692 */
  if (min_errors > (int)total_edit_dist) return;                 /* 2693 */
  if (min_errors > 2) return;                                    /* 2696 */
  for (int i = 0; i < np; ++i) {                                 /* 2699-2708 */
    int left = positions[i];
    if (left < P->fusion_anchor_length) return;
    int right = left;
    if (read_length - right < P->fusion_anchor_length) return;
  }
  for (int i = 0; i < np; ++i) {                                 /* 2710-2804 */
    int bl = positions[i], br = bl;
    uint32_t left, right;
    if (dir == FUSION_FF || dir == FUSION_FR) left = (uint32_t)(L->left + bl - 1); else left = (uint32_t)(L->right - bl);
    if (dir == FUSION_FF || dir == FUSION_RF) right = (uint32_t)(R->right - (read_length - br));
    else right = (uint32_t)(R->left + (read_length - br) - 1);
    uint32_t r1 = L->ref_id, r2 = R->ref_id, tdir = dir;
    if (r2 < r1 || (r1 == r2 && left > right)) {                 /* 2772-2785 */
      uint32_t t = r1; r1 = r2; r2 = t; t = left; left = right; right = t;
      if (dir == FUSION_FF) tdir = FUSION_RR;
    }
    add_fusion(out, r1, r2, left, right, tdir, total_edit_dist);
  }
}

/* find_fusions, segment_juncs.cpp:2976-3291.  --fusion-ignore-chromosomes is not modelled. */
static void find_fusions(const thb_params* P, const orc_ref* rt, const char* read_seq, int read_len,
                         hitvec* H, int nsegs, const thb_hit* partner, int n_partner, orc_results* out)
{
  if (nsegs <= 0) return;
  int last = nsegs - 1;
  while (last > 0) { if (H[last].n) break; --last; }             /* 2987-2994 */
  if (last == 0 && (H[0].n == 0 || ISEND(H[0].h[0]))) return;    /* 3034-3036 */
  int has_partner = n_partner > 0;
  hitvec right = {0, 0, 0};                                      /* 3074-3079 */
  if (last != 0) hv_copy(&right, &H[last]);
  char rc[512]; revcomp(read_seq, read_len, rc);
  int check_partner = 1;
  if (last != 0) {                                               /* 3089-3117 */
    for (int i = 0; i < H[0].n && check_partner; ++i)
      for (int j = 0; j < right.n; ++j) {
        const thb_hit* l = &H[0].h[i]; const thb_hit* r = &right.h[j];
        if (l->ref_id == r->ref_id && ANTI(*l) == ANTI(*r)) {
          int dist = ANTI(*l) ? l->left - r->right : r->left - l->right;
          if (dist > -P->max_insertion_length && dist <= P->fusion_min_dist) { check_partner = 0; break; }
        }
      }
  }
  const int minus_dist = -P->max_insertion_length * 2;          /* 3119 */
  if (check_partner && has_partner) {                            /* 3121-3212 */
    for (int l = 0; l < H[0].n; ++l)
      for (int r = 0; r < n_partner; ++r) {
        const thb_hit* leftHit = &H[0].h[l]; const thb_hit* rightHit = &partner[r];
        if (leftHit->ref_id == rightHit->ref_id && ANTI(*leftHit) != ANTI(*rightHit)) {
          int dist = ANTI(*leftHit) ? leftHit->left - rightHit->right : rightHit->left - leftHit->right;
          if (dist > minus_dist && dist <= P->fusion_min_dist) continue;          /* 3140 */
        }
        if (!rescue_in_flank(P, rt, read_seq, rc, read_len, rightHit, &right, out)) break;
      }
  }
  for (int li = 0; li < H[0].n; ++li)                            /* 3221-3290 */
    for (int ri = 0; ri < right.n; ++ri) {
      const thb_hit* lh = &H[0].h[li]; const thb_hit* rh = &right.h[ri];
      if (P->bowtie2 && (int)lh->edit_dist + (int)rh->edit_dist > (P->segment_mismatches << 1)) continue; /* 3232-3236 */
      if (lh->ref_id == rh->ref_id && ANTI(*lh) == ANTI(*rh)) {  /* 3255-3268 */
        int dist = ANTI(*lh) ? lh->left - rh->right : rh->left - lh->right;
        if (dist > minus_dist && dist <= P->fusion_min_dist) continue;
      }
      uint32_t dir = FUSION_FF; const char* mod = read_seq;
      if (ANTI(*lh) == ANTI(*rh)) { if (ANTI(*lh)) { const thb_hit* t = lh; lh = rh; rh = t; mod = rc; } }  /* 3273-3282 */
      else if (!ANTI(*lh) && ANTI(*rh)) dir = FUSION_FR;         /* 3283 */
      else dir = FUSION_RF;
      detect_fusion(P, rt, mod, read_len, lh, rh, dir, out);
    }
  hv_free(&right);
}

/* ------------------------------------------------------------------------------------------- */
/* set semantics of the result containers                                                       */

static int cmp_junc(const void* a, const void* b)
{ /* Junction::operator<, junctions.h:39-57 */
  const orc_junc* x = (const orc_junc*)a; const orc_junc* y = (const orc_junc*)b;
  if (x->ref_id != y->ref_id) return x->ref_id < y->ref_id ? -1 : 1;
  if (x->left != y->left) return x->left < y->left ? -1 : 1;
  if (x->right != y->right) return x->right < y->right ? -1 : 1;
  if (x->antisense != y->antisense) return x->antisense < y->antisense ? -1 : 1;
  return 0;
}
static int cmp_ins(const void* a, const void* b)
{ /* Insertion::operator< (ref, left, LENGTH of sequence), insertions.h:52-67; then insertion order */
  const orc_ins* x = (const orc_ins*)a; const orc_ins* y = (const orc_ins*)b;
  if (x->ref_id != y->ref_id) return x->ref_id < y->ref_id ? -1 : 1;
  if (x->left != y->left) return x->left < y->left ? -1 : 1;
  if (x->len != y->len) return x->len < y->len ? -1 : 1;
  if (x->order != y->order) return x->order < y->order ? -1 : 1;
  return 0;
}
static int cmp_fus(const void* a, const void* b)
{ /* Fusion::operator<, fusions.h:40-70 */
  const orc_fus* x = (const orc_fus*)a; const orc_fus* y = (const orc_fus*)b;
  if (x->ref1 != y->ref1) return x->ref1 < y->ref1 ? -1 : 1;
  if (x->ref2 != y->ref2) return x->ref2 < y->ref2 ? -1 : 1;
  if (x->left != y->left) return x->left < y->left ? -1 : 1;
  if (x->right != y->right) return x->right < y->right ? -1 : 1;
  if (x->dir != y->dir) return x->dir < y->dir ? -1 : 1;
  return 0;
}
static size_t uniq_junc(orc_junc* a, size_t n)
{
  if (!n) return 0;
  qsort(a, n, sizeof *a, cmp_junc);
  size_t k = 1;
  for (size_t i = 1; i < n; ++i) if (cmp_junc(&a[i], &a[k - 1]) != 0) a[k++] = a[i];
  return k;
}

/* ------------------------------------------------------------------------------------------- */
/* public entry points (loaded through ctypes by tests/ and bench.py only)                      */

orc_results* orc_results_new(void) { return (orc_results*)calloc(1, sizeof(orc_results)); }
void orc_results_free(orc_results* r) { if (!r) return; free(r->juncs); free(r->dels); free(r->ins); free(r->fus); free(r); }

/* Processes every bundle of one batch in order (SegmentSearchWorker loop, 4624-4652, with the
 * per-bundle call order chosen by look_for_hit_group / process_next_hit_group, 4005-4033 and
 * 4092-4117, which the host encodes in thb_bundle.flags). */
int orc_segjuncs_batch(const thb_params* P, const thb_ref_image* img, const thb_segjuncs_batch* b, orc_results* out)
{
  orc_ref rt = { img };
  if (b->n_segs > MAX_SEGS) return THB_EUNSUPPORTED;
  hitvec H[MAX_SEGS]; memset(H, 0, sizeof H);
  char read[512];
  for (uint32_t bi = 0; bi < b->n_bundles; ++bi) {
    const thb_bundle* bu = &b->bundles[bi];
    unpack_read(b, bi, bu->read_len, read);
    /* insertion priority = global processing position of the bundle (first inserted wins,
     * insertions.h:52-67); order_base lets a sharded run reproduce the single-process order */
    out->order = (b->order_base + bi) << 12;
    uint32_t off = bu->hit_begin;
    for (uint32_t s = 0; s < b->n_segs; ++s) {
      H[s].n = 0;
      uint32_t c = b->seg_count[(size_t)bi * b->n_segs + s];
      for (uint32_t k = 0; k < c; ++k) hv_push(&H[s], b->hits[off + k]);
      off += c;
    }
    const thb_hit* partner = b->partner_hits + bu->partner_begin;
    int nsegs = (int)b->n_segs;
    int right_mate = (bu->flags & THB_BUNDLE_RIGHT_MATE) != 0;
    if (bu->flags & THB_BUNDLE_INDELS)
      find_insertions_and_deletions(P, &rt, read, bu->read_len, H, nsegs, out);
    if ((bu->flags & THB_BUNDLE_FUSIONS) && !(bu->flags & THB_BUNDLE_FUSIONS_LAST))
      find_fusions(P, &rt, read, bu->read_len, H, nsegs, partner, bu->n_partner, out);
    if (bu->flags & THB_BUNDLE_GAPS)
      find_gaps(P, &rt, read, bu->read_len, H, &nsegs, partner, bu->n_partner, right_mate, out);
    if ((bu->flags & THB_BUNDLE_FUSIONS) && (bu->flags & THB_BUNDLE_FUSIONS_LAST))
      find_fusions(P, &rt, read, bu->read_len, H, nsegs, partner, bu->n_partner, out);
    /* keep the working sets bounded like std::set would: dedup when they grow large */
    if (out->n_juncs > (1u << 22)) out->n_juncs = uniq_junc(out->juncs, out->n_juncs);
    if (out->n_dels > (1u << 22)) out->n_dels = uniq_junc(out->dels, out->n_dels);
  }
  for (uint32_t s = 0; s < b->n_segs; ++s) hv_free(&H[s]);
  return THB_OK;
}

/* Final std::set views: sorted unique junctions (capped at max_seg_juncs = 10,000,000 by erasing
 * the largest, 58/1692-1693), deletions, insertions with first-inserted-wins, fusions by key. */
int orc_segjuncs_finish(orc_results* r)
{
  r->n_juncs = uniq_junc(r->juncs, r->n_juncs);
  if (r->n_juncs > 10000000u) r->n_juncs = 10000000u;
  r->n_dels = uniq_junc(r->dels, r->n_dels);
  if (r->n_ins) {
    qsort(r->ins, r->n_ins, sizeof *r->ins, cmp_ins);
    size_t k = 1;
    for (size_t i = 1; i < r->n_ins; ++i) {
      const orc_ins* p = &r->ins[k - 1]; const orc_ins* q = &r->ins[i];
      if (p->ref_id == q->ref_id && p->left == q->left && p->len == q->len) continue;
      r->ins[k++] = *q;
    }
    r->n_ins = k;
  }
  if (r->n_fus) qsort(r->fus, r->n_fus, sizeof *r->fus, cmp_fus);
  return THB_OK;
}

size_t orc_n_juncs(const orc_results* r) { return r->n_juncs; }
size_t orc_n_dels(const orc_results* r) { return r->n_dels; }
size_t orc_n_ins(const orc_results* r) { return r->n_ins; }
size_t orc_n_fus(const orc_results* r) { return r->n_fus; }
void orc_get_juncs(const orc_results* r, thb_junction* out) { for (size_t i = 0; i < r->n_juncs; ++i) { out[i].ref_id = r->juncs[i].ref_id; out[i].left = r->juncs[i].left; out[i].right = r->juncs[i].right; out[i].antisense = r->juncs[i].antisense; } }
void orc_get_dels(const orc_results* r, thb_junction* out) { for (size_t i = 0; i < r->n_dels; ++i) { out[i].ref_id = r->dels[i].ref_id; out[i].left = r->dels[i].left; out[i].right = r->dels[i].right; out[i].antisense = 0; } }
void orc_get_ins(const orc_results* r, thb_insertion* out) { for (size_t i = 0; i < r->n_ins; ++i) { memset(&out[i], 0, sizeof out[i]); out[i].ref_id = r->ins[i].ref_id; out[i].left = r->ins[i].left; out[i].len = r->ins[i].len; memcpy(out[i].seq, r->ins[i].seq, sizeof out[i].seq); } }
void orc_get_fus(const orc_results* r, thb_fusion* out) { for (size_t i = 0; i < r->n_fus; ++i) { out[i].ref_id1 = r->fus[i].ref1; out[i].ref_id2 = r->fus[i].ref2; out[i].left = r->fus[i].left; out[i].right = r->fus[i].right; out[i].dir = r->fus[i].dir; out[i].count = r->fus[i].count; out[i].edit_dist = r->fus[i].edit_dist; out[i].reserved = 0; } }
void orc_get_ins_order(const orc_results* r, uint64_t* out) { for (size_t i = 0; i < r->n_ins; ++i) out[i] = r->ins[i].order; }
void orc_get_counters(const orc_results* r, uint64_t* out5)
{ out5[0] = r->n_windows; out5[1] = r->n_indel_tasks; out5[2] = r->n_rescue_tasks; out5[3] = r->n_fusion_tasks; out5[4] = r->n_juncs_emitted; }
