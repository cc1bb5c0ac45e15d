/*
 * tophat_b200.h -- C ABI of libtophat_b200.so
 *
 * The drop-in boundary of the B200-native splice-junction hot path.  The reference (TopHat 2.1.2)
 * exposes this path only as two executables (segment_juncs, long_spanning_reads) driven by
 * tophat.py; our replacements of those executables are thin C++ hosts that do argv/BAM/FASTA/text
 * I/O and call the entry points below for every per-read computation.  Each entry point names the
 * reference function(s) it replaces (paths relative to the reference tree, src/...).
 *
 * Conventions
 *   - plain C, POD structs, pointers + sizes only; no C++/torch types cross the boundary;
 *   - every function returns 0 on success and a negative THB_E* code on failure; the message is
 *     available from thb_last_error(); the library never calls exit();
 *   - the caller owns every host buffer it passes in; the context owns all device memory;
 *   - one context per process per device; a context is not re-entrant;
 *   - there is NO CPU fallback: without a usable sm_100 device thb_create() fails.
 *
 * Coordinates are the reference's: 0-based, `left` inclusive, `right` as BowtieHit::right()
 * (bwt_map.h:213-243, one past the last aligned reference base).  Reference ids are 1-based in
 * SAM-header order (bwt_map.h:608-674); id 0 is "none".
 */
#ifndef TOPHAT_B200_H
#define TOPHAT_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define THB_OK              0
#define THB_ENODEVICE      -1   /* no CUDA device / wrong architecture            */
#define THB_ECUDA          -2   /* a CUDA runtime call failed                     */
#define THB_EINVAL         -3   /* malformed argument / batch                     */
#define THB_ENOMEM         -4   /* host or device allocation failed               */
#define THB_EUNSUPPORTED   -5   /* parameter combination outside the GPU path     */
#define THB_ESTATE         -6   /* call sequence error (e.g. no reference loaded) */
#define THB_ENCCL          -7   /* NCCL failure                                   */

typedef struct thb_ctx thb_ctx;

/* Segments per read both stages accept (reads of up to 12 x --segment-length bases; the reference has no fixed limit, tophat.py
 * splits 2x101 bp reads into 4).  One limit for batches and executables alike.                                                  */
#define THB_MAX_SEGS 12

/* ---- option globals consumed on the hot path (common.cpp:79-180, parsed at common.cpp:459-721) */
typedef struct thb_params {
  int32_t segment_length;             /* --segment-length            (common.cpp:121) 25      */
  int32_t segment_mismatches;         /* --segment-mismatches        (common.cpp:122) 2       */
  int32_t min_segment_intron_length;  /* --min-segment-intron        (common.cpp:115) 50      */
  int32_t max_segment_intron_length;  /* --max-segment-intron        (common.cpp:116) 500000  */
  int32_t max_insertion_length;       /* --max-insertion-length      (common.cpp:98)  3       */
  int32_t max_deletion_length;        /* --max-deletion-length       (common.cpp:99)  3       */
  int32_t max_seg_multihits;          /* --max-seg-multihits         (common.cpp:135) 40      */
  int32_t inner_dist_mean;            /* --inner-dist-mean           (common.cpp:101) 200     */
  int32_t inner_dist_std_dev;         /* --inner-dist-std-dev        (common.cpp:102) 20      */
  int32_t bowtie2;                    /* 0 with --bowtie1            (common.cpp:79)  1       */
  int32_t library_type;               /* eLIBRARY_TYPE (common.h): 0 none, 1 fr-unstranded,
                                         2 fr-firststrand, 3 fr-secondstrand, 4.. ff-*        */
  int32_t fusion_search;              /* --fusion-search             (common.cpp:171) 0       */
  int32_t fusion_anchor_length;       /* --fusion-anchor-length      (common.cpp:172) 20      */
  int32_t fusion_min_dist;            /* --fusion-min-dist           (common.cpp:173) 10000000*/
  /* long_spanning_reads only */
  int32_t max_report_intron_length;   /* --max-report-intron         (common.cpp:107) 500000  */
  int32_t min_report_intron_length;   /* --min-report-intron         (common.cpp:106) 50      */
  int32_t min_anchor_len;             /* --min-anchor                (common.cpp:105) 8       */
  int32_t read_mismatches;            /* --read-mismatches           (common.cpp:123) 2       */
  int32_t read_gap_length;            /* --read-gap-length           (common.cpp:124) 2       */
  int32_t read_edit_dist;             /* --read-edit-dist            (common.cpp:125) 2       */
  int32_t bowtie2_max_penalty;        /* (common.cpp:87) 6 */
  int32_t bowtie2_min_penalty;        /* (common.cpp:88) 2 */
  int32_t bowtie2_penalty_for_N;      /* (common.cpp:89) 1 */
  int32_t bowtie2_read_gap_open;      /* (common.cpp:90) 5 */
  int32_t bowtie2_read_gap_cont;      /* (common.cpp:91) 3 */
  int32_t bowtie2_ref_gap_open;       /* (common.cpp:92) 5 */
  int32_t bowtie2_ref_gap_cont;       /* (common.cpp:93) 3 */
  int32_t reserved[5];
} thb_params;

/* Fills *p with the reference binaries' defaults (common.cpp:79-180). */
void thb_params_default(thb_params* p);

/* ---- device context ---------------------------------------------------------------------------*/

/* Creates a context on CUDA device `device` (must be compute capability 10.x). */
int  thb_create(int device, thb_ctx** out);
void thb_destroy(thb_ctx* ctx);
const char* thb_last_error(const thb_ctx* ctx);   /* ctx may be NULL: last creation error */
const char* thb_version(void);

/* ---- reference genome image ---------------------------------------------------------------------
 * Replaces RefSequenceTable + get_seqs (bwt_map.h:579-788; segment_juncs.cpp:64-88): the genome
 * as bit planes instead of seqan::String<Dna5,Packed<>>.
 *
 * Global base coordinate g = contig_start[id-1] + pos.  contig_start[] are multiples of 64 and
 * leave >= 64 bases of zero padding after every contig.  For every 64-base block b:
 *   planes[2*b+0] bit j = low  bit of the 2-bit code of base 64*b+j   (A=0 C=1 G=2 T=3)
 *   planes[2*b+1] bit j = high bit
 *   nmask[b]      bit j = 1 if the FASTA byte was not one of ACGTUacgtu (SeqAn Dna5 'N',
 *                         alphabet_residue_tabs.h:107-140); such bases carry code 0 in `planes`,
 *                         which is exactly the Dna5->Dna conversion (`& 3`) the junction scan and
 *                         insertion search apply (segment_juncs.cpp:2157, 2499).
 */
typedef struct thb_ref_image {
  uint32_t        n_contigs;
  const uint64_t* contig_start;   /* [n_contigs] global coordinate of base 0 of each contig    */
  const uint32_t* contig_len;     /* [n_contigs] 0 = id known from the SAM header, no sequence */
  uint64_t        n_blocks;       /* number of 64-base blocks                                  */
  const uint64_t* planes;         /* [2*n_blocks]                                              */
  const uint64_t* nmask;          /* [n_blocks]                                                */
} thb_ref_image;

int thb_ref_upload(thb_ctx* ctx, const thb_ref_image* img);

/* Host helper: packs `len` FASTA bytes of one contig (no newlines) into planes/nmask at global
 * coordinate `gstart` (multiple of 64).  Buffers must be zero-initialised by the caller.       */
void thb_pack_bases(const char* seq, uint64_t len, uint64_t gstart, uint64_t* planes, uint64_t* nmask);

/* ---- segment_juncs batch ------------------------------------------------------------------------
 * One thb_bundle = one call of find_insertions_and_deletions / find_fusions / find_gaps for one
 * read, i.e. the `hits_for_read` vector that look_for_hit_group / process_next_hit_group
 * (segment_juncs.cpp:3823-4123) assemble, plus the partner hit group find_gaps looks up
 * (segment_juncs.cpp:3322-3344).  Bundles are stored in the reference's processing order.
 */
typedef struct thb_hit {          /* the BowtieHit fields the path reads (bwt_map.h:78-536)      */
  uint32_t ref_id;                /* ref_id()                                                    */
  int32_t  left;                  /* left()                                                      */
  int32_t  right;                 /* right()                                                     */
  uint8_t  read_len;              /* read_len()                                                  */
  uint8_t  edit_dist;             /* edit_dist()                                                 */
  uint8_t  flags;                 /* THB_HIT_*                                                   */
  uint8_t  reserved;
} thb_hit;                        /* 16 bytes */

#define THB_HIT_ANTISENSE  0x01   /* antisense_align() */
#define THB_HIT_END        0x02   /* end()             */

typedef struct thb_bundle {
  uint32_t read_id;               /* insert_id                                                   */
  uint32_t hit_begin;             /* first hit of segment 0 in hits[]; segments follow in order  */
  uint32_t partner_begin;         /* first partner hit in partner_hits[]                         */
  uint16_t n_partner;             /* partner group size; 0 <=> !has_partner                      */
  uint8_t  read_len;              /* bases in the read                                           */
  uint8_t  flags;                 /* THB_BUNDLE_*                                                */
} thb_bundle;                     /* 16 bytes */

#define THB_BUNDLE_INDELS        0x01  /* run find_insertions_and_deletions (2807-2942)          */
#define THB_BUNDLE_GAPS          0x02  /* run find_gaps (3293-3650)                              */
#define THB_BUNDLE_FUSIONS       0x04  /* run find_fusions (2976-3291)                           */
#define THB_BUNDLE_FUSIONS_LAST  0x08  /* find_fusions sees the bundle as mutated by find_gaps
                                          (different-group branch, 4005-4033)                    */
#define THB_BUNDLE_RIGHT_MATE    0x10  /* eREAD read_side == READ_RIGHT (segments.h:12-17)       */

typedef struct thb_segjuncs_batch {
  uint32_t          n_bundles;
  uint32_t          n_segs;        /* number of segment files (hits_for_read.size())             */
  uint32_t          read_words;    /* 64-bit words per bit plane of a read = ceil(maxlen/64)     */
  uint32_t          reserved;
  const thb_bundle* bundles;       /* [n_bundles]                                                */
  const uint16_t*   seg_count;     /* [n_bundles*n_segs] hits per segment                        */
  const uint64_t*   reads;         /* [n_bundles*3*read_words]: plane0 | plane1 | planeN words;
                                      bit j of word w = base 64*w+j; planeN set for every read
                                      byte outside ACGT (such bases have code 0)                 */
  uint64_t          n_hits;
  const thb_hit*    hits;          /* [n_hits]                                                   */
  uint64_t          n_partner_hits;
  const thb_hit*    partner_hits;  /* [n_partner_hits]                                           */
  uint64_t          order_base;    /* added to the bundle index to form the insertion
                                      first-wins priority (insertions.h:52-67): lower wins       */
} thb_segjuncs_batch;

/* Host helper: packs an ASCII read into the three planes (read_words words each). */
void thb_pack_read(const char* seq, uint32_t len, uint32_t read_words, uint64_t* out3planes);

/* Result records.  Same value types and orderings as junctions.h:27-80, deletions.h:26,
 * insertions.h:31-74, fusions.h:24-116.                                                         */
typedef struct thb_junction { uint32_t ref_id; uint32_t left; uint32_t right; uint32_t antisense; } thb_junction;
typedef struct thb_insertion { uint32_t ref_id; uint32_t left; uint32_t len; char seq[20]; } thb_insertion;
typedef struct thb_fusion { uint32_t ref_id1; uint32_t ref_id2; uint32_t left; uint32_t right;
                            uint32_t dir; uint32_t count; uint32_t edit_dist; uint32_t reserved; } thb_fusion;

typedef struct thb_segjuncs_results {
  uint64_t            n_junctions;  const thb_junction*  junctions;   /* Junction order          */
  uint64_t            n_deletions;  const thb_junction*  deletions;   /* as Deletion(ref,l,r)    */
  uint64_t            n_insertions; const thb_insertion* insertions;  /* (ref,left,len) order    */
  uint64_t            n_fusions;    const thb_fusion*    fusions;     /* Fusion order            */
} thb_segjuncs_results;

/* Clears the accumulated junction / deletion / insertion / fusion sets of the context. */
int thb_segjuncs_begin(thb_ctx* ctx, const thb_params* params);

/* --fusion-ignore-chromosomes (common.cpp:174, segment_juncs.cpp:3213-3230): contig ids (1-based, as in thb_hit.ref_id)
 * whose hits find_fusions skips.  Call after thb_segjuncs_begin; cleared by the next thb_segjuncs_begin.           */
int thb_segjuncs_fusion_ignore(thb_ctx* ctx, const uint32_t* ref_ids, uint32_t n);

/* Processes one batch whose arrays live in HOST memory (copied to the device inside the call).
 * Replaces, for every bundle, find_insertions_and_deletions -> detect_small_{deletion,insertion}
 * -> simpleSplitAlignment (2807-2942, 2554-2627, 2470-2541, 2390-2456), find_gaps (3293-3650)
 * with map_read_to_contig (2946-2973) and juncs_from_ref_segs<RecordSegmentJuncs> for GT-AG,
 * GC-AG, AT-AC (2051-2377, 1669-1696), and find_fusions/detect_fusion (2976-3291, 2629-2805). */
int thb_segjuncs_submit(thb_ctx* ctx, const thb_segjuncs_batch* host_batch);

/* Same, for a batch whose arrays already live in DEVICE memory of ctx's device (every array 16-byte aligned:
 * the kernels use 128-bit loads; cudaMalloc / torch allocations are).                                        */
int thb_segjuncs_submit_device(thb_ctx* ctx, const thb_segjuncs_batch* device_batch);

/* Sorts and de-duplicates the accumulated sets (the std::set semantics of 4907-4922, insertion
 * first-wins, the max_seg_juncs cap of 58/1692-1693) and returns host pointers owned by ctx,
 * valid until the next thb_segjuncs_begin / thb_destroy.                                       */
int thb_segjuncs_finish(thb_ctx* ctx, thb_segjuncs_results* out);
/* thb_segjuncs_finish in two halves, for a caller that goes on with the device-resident sets (thb_join_begin_resident, thb_flank_begin
 * on the host arrays later): _finish_resident builds the sets on the device, fills the counts and the pointers of `out` and queues
 * the download of the arrays on a copy stream WITHOUT waiting for it; the arrays are valid once thb_segjuncs_fetch has returned. */
int thb_segjuncs_finish_resident(thb_ctx* ctx, thb_segjuncs_results* out);
int thb_segjuncs_fetch(thb_ctx* ctx);

/* Multi-GPU exchange (replaces the per-thread set union at 4911-4922): all-gathers the
 * de-duplicated junction/deletion/insertion/fusion sets of every rank over NCCL and merges them,
 * after which thb_segjuncs_finish returns the union on every rank.  `nccl_unique_id` is the 128
 * byte ncclUniqueId obtained on rank 0 via thb_nccl_unique_id and distributed by the caller.   */
int thb_nccl_unique_id(void* out128);
int thb_comm_init(thb_ctx* ctx, const void* nccl_unique_id128, int rank, int world);
int thb_segjuncs_allgather(thb_ctx* ctx);

/* ---- long_spanning_reads: segment-chain join -----------------------------------------------------------
 * One thb_join_bundle = one call of join_segments_for_read (long_spanning_reads.cpp:2612-2667): the read plus,
 * per segment, the hit group JoinSegmentsWorker assembled for it (contiguous hits from the segment BAM followed
 * by the hits against juncs_db contigs converted to genomic coordinates, 2706-2765 / 87-163).  The host keeps
 * the worker's stream rules (every segment must have a hit, the last one must carry the end flag, 2768-2785);
 * the kernel runs the chain DFS (2222-2610), merge_chain (805-2038) and the edit-distance consistency check
 * (bwt_map.cpp:2349-2465) and returns every valid merged alignment.  Sorting / de-duplicating a read's
 * alignments (2805-2807), the read-level filters (2810-2813) and the SAM fields are the caller's.              */
#define THB_JHIT_MAX_OPS    9
#define THB_JOINED_MAX_OPS  27
#define THB_JHIT_ANTISENSE_SPLICE 0x04    /* antisense_splice() (with THB_HIT_ANTISENSE / THB_HIT_END)        */
/* CIGAR ops are packed as length << 4 | CigarOpCode (bwt_map.h:36-55: MATCH 1, INS 3, DEL 5, REF_SKIP 11,
 * SOFT_CLIP 13, PAD 15; lower-case fusion-side codes never occur without --fusion-search).                    */
#define THB_JHIT_ONE_MATCH  0x08   /* the CIGAR is a single MATCH of right-left bases: no thb_jops record        */
/* --fusion-search: CigarOpCodes 2 / 4 / 6 / 12 (mATCH, iNS, dEL, rEF_SKIP: the lower-case ops of a part that is read leftwards) and
 * 7 .. 10 (FUSION_FF / FR / RF / RR, length = position on the second contig) occur (bwt_map.h:36-55).  A segment hit against a
 * `fus` contig of the junction index carries its second contig in thb_jhit_full.ops[8] (so it has at most 8 CIGAR ops), which
 * thb_join_pack_hits moves to thb_jops.ops[11].  thb_joined.flags of a merged alignment with a fusion has THB_JOINED_FUSION and
 * its second contig in ops[THB_JOINED_MAX_OPS - 1] (such an alignment has at most 26 ops); THB_JOINED_SEQ_RC says that the
 * alignment's own sequence (BowtieHit::seq(), what bowtie_sam_extra reads) is the reverse complement of the read.             */
#define THB_JHIT_SEQ_FLIPPED 0x10  /* thb_jhit_full.flags only: the hit's BAM sequence is oriented opposite to what THB_HIT_ANTISENSE
                                      says (a hit on an rf / rr fusion contig, whose strand the reference flips, bwt_map.cpp:1740-1741);
                                      thb_join_pack_hits moves it to thb_jops.ops[10]                                              */
#define THB_JOINED_FUSION   0x10
#define THB_JOINED_SEQ_RC   0x20
/* Wire form of one segment's BowtieHit (BAMHitFactory / SplicedBAMHitFactory).  Nearly every hit is an un-gapped match,
 * so the record the kernels stream is 16 bytes: position, right() and the small fields; the CIGAR of the other hits
 * (against the junction index: M N M, indels ...) sits in a side array, one thb_jops per such hit, in hit order.       */
typedef struct thb_jhit {
  uint32_t ref_id;
  int32_t  left;
  int32_t  right;                  /* right() = left + bases of MATCH / DEL / REF_SKIP ops                       */
  uint8_t  flags_nops;             /* THB_HIT_ANTISENSE | THB_HIT_END | THB_JHIT_ANTISENSE_SPLICE | THB_JHIT_ONE_MATCH
                                      in the low nibble, number of CIGAR ops (1..9) in the high nibble           */
  uint8_t  ops_index;              /* without THB_JHIT_ONE_MATCH: ops_ext[bundle.ops_begin + ops_index]          */
  uint8_t  mismatches, splice_mms;
} thb_jhit;                        /* 16 bytes */
typedef struct thb_jops { uint32_t ops[12]; } thb_jops;    /* 48 bytes; ops[0 .. n_ops) valid, the rest zero     */

/* Unpacked form (one record per hit, CIGAR inline) and the host helper that converts the hits of ONE read -- all its
 * segments, in batch order -- to the wire form.  Returns the number of thb_jops records written (<= n), or
 * THB_EUNSUPPORTED when the read has more than 256 hits with a multi-op CIGAR.                                    */
typedef struct thb_jhit_full {
  uint32_t ref_id;
  int32_t  left;
  uint8_t  n_ops, flags, mismatches, splice_mms;
  uint32_t ops[THB_JHIT_MAX_OPS];
} thb_jhit_full;                   /* 48 bytes */
int thb_join_pack_hits(const thb_jhit_full* hits, uint32_t n, thb_jhit* heads, thb_jops* ops_ext);

typedef struct thb_join_bundle {
  uint32_t read_id;
  uint32_t hit_begin;              /* first hit of segment 0; segments follow in order                         */
  uint16_t read_len;
  uint8_t  n_segs;                 /* segments of THIS read (all non-empty)                                    */
  uint8_t  reserved;
  uint32_t ops_begin;              /* first thb_jops record of this read's hits in ops_ext[]                   */
} thb_join_bundle;                 /* 16 bytes */

typedef struct thb_join_batch {
  uint32_t               n_bundles;
  uint32_t               n_segs;       /* row stride of seg_count (maximum segments per read)                  */
  uint32_t               read_words;
  uint32_t               reserved;
  const thb_join_bundle* bundles;
  const uint16_t*        seg_count;    /* [n_bundles * n_segs]                                                  */
  const uint64_t*        reads;        /* [n_bundles * 3 * read_words], as in thb_segjuncs_batch                */
  uint64_t               n_hits;
  const thb_jhit*        hits;
  uint64_t               n_ops_ext;
  const thb_jops*        ops_ext;      /* [n_ops_ext] CIGARs of the hits without THB_JHIT_ONE_MATCH, in hit order     */
} thb_join_batch;

typedef struct thb_joined {        /* the BowtieHit merge_chain returns                                         */
  uint32_t bundle;                 /* index of the read in the submitted batch                                  */
  uint32_t ref_id;
  int32_t  left;
  uint8_t  n_ops, flags, mismatches, edit_dist;
  uint8_t  splice_mms, reserved8[3];
  uint32_t ops[THB_JOINED_MAX_OPS];
} thb_joined;                      /* 128 bytes */

/* Uploads the junction set (junction files + deletions as Junction(ref, left - 1, right), 2895-2944) and the
 * insertion set (2952-2980).  Arrays must be sorted and unique in the reference's set orders.                */
int thb_join_begin(thb_ctx* ctx, const thb_params* params, const thb_junction* juncs, uint64_t n_juncs,
                   const thb_insertion* insertions, uint64_t n_insertions);
/* Same with the sets of the segment_juncs pass this context has just finished (thb_segjuncs_finish or _finish_resident), taken where they
 * lie in device memory: junctions and deletions merged into one Junction-ordered set (deletions with antisense = false), insertions. */
int thb_join_begin_resident(thb_ctx* ctx, const thb_params* params);
/* --fusion-search (params->fusion_search != 0 in thb_join_begin): the fusion set of long_spanning_reads.cpp:2996-3040, sorted and
 * unique in Fusion order (refid1, refid2, left, right, dir; fusions.h:40-70); count / edit_dist of the records are ignored.
 * Call after thb_join_begin; thb_join_begin resets the set to empty.                                                          */
int thb_join_set_fusions(thb_ctx* ctx, const thb_fusion* fusions, uint64_t n_fusions);
/* Joins one batch (host arrays); *out / *n_out receive the merged alignments of this batch, grouped by nothing
 * in particular (use .bundle), owned by ctx and valid until the next join call.                              */
int thb_join_submit(thb_ctx* ctx, const thb_join_batch* host_batch, const thb_joined** out, uint64_t* n_out);
/* Same for a batch whose arrays already live in device memory; the merged alignments stay on the device until
 * thb_join_fetch copies them to the host (same lifetime as above).                                          */
int thb_join_submit_device(thb_ctx* ctx, const thb_join_batch* device_batch, uint64_t* n_out);
int thb_join_fetch(thb_ctx* ctx, const thb_joined** out, uint64_t* n_out);
/* Since round 2 the enumeration and the closure-free merges are one kernel (join_tile_kernel): its time is enum_ms,
 * merge_simple_ms / merge_abutting_ms stay 0, merge_general_ms is the closure kernel.
 * h2d_ms: wall time of thb_join_submit's pipeline (upload, kernels and download of consecutive chunks overlap);
 * d2h_ms: thb_join_fetch only.                                                                                 */
typedef struct thb_join_timing { float h2d_ms, kernel_ms, d2h_ms; float enum_ms, merge_ms;   /* kernel_ms = enum + merge */
                                 uint32_t launches;
                                 uint64_t n_chains, n_closures, n_joined, algorithmic_bytes;
                                 uint64_t n_simple_chains, n_abutting_chains;   /* chains merged without a closure search */
                                 float merge_simple_ms, merge_abutting_ms, merge_general_ms;   /* the three merge kernels  */
                                 float begin_ms;   /* thb_join_begin: set upload, validation, bucket index build (CUDA events) */
                               } thb_join_timing;
int thb_join_last_timing(thb_ctx* ctx, thb_join_timing* out);

/* Kernel timing of the last submit (CUDA events on the context's stream), milliseconds.        */
typedef struct thb_timing {
  float h2d_ms;
  float scan_kernel_ms;         /* sum of the five scan-phase kernels below                              */
  float finish_ms; float total_ms;
  float bundle_ms, hit_ms, rescue_ms, rescued_windows_ms, window_scan_ms, indel_ms;   /* per kernel, CUDA events */
  float fusion_enum_ms, fusion_detect_ms;                                              /* --fusion-search only    */
  uint64_t n_windows; uint64_t n_indel_tasks; uint64_t n_rescue_tasks; uint64_t n_juncs_emitted;
  uint64_t n_fusion_tasks;      /* detect_fusion calls that reached simpleSplitAlignment                 */
  uint64_t algorithmic_bytes;
  uint32_t kernel_launches;   /* launches of the scan kernel                                       */
  uint32_t total_launches;    /* every kernel of this library since thb_segjuncs_begin (valid after finish) */
} thb_timing;
int thb_last_timing(thb_ctx* ctx, thb_timing* out);

/* ------------------------------------------------------------------------------------------------------------------
 * Junction-flank matcher: the step BETWEEN the two stages (tophat.py:2546-2600 build_juncs_index, 3686-3741 map2juncs).
 * Replaces  juncs_db <min_anchor> <max_seg_len> juncs insertions deletions fusions ref.fa   (src/juncs_db.cpp:72-229, 481-528)
 *           bowtie-build segment_juncs.fa
 *           bowtie -v <segment_mismatches> -k <max_seg_multihits> -m <max_seg_multihits> <junction index> <segments>
 * for one resident reference: thb_flank_begin builds the contigs and their seed index on the device, thb_flank_submit reports, for a
 * batch of reads cut into segments, every un-gapped placement of a segment (or of its reverse complement) on a contig with at
 * most max_mismatches mismatches -- none for a segment that has more than max_multihits placements (bowtie's -m).  Placements come
 * back sorted by (read, segment, contig, pos, antisense).  An 'N' of the read is a mismatch; a placement over an 'N' of the contig is
 * invalid (bowtie 1) or, with ref_n_is_mismatch, a mismatch (bowtie 2's treatment).                                              */
#define THB_FLANK_JUNC 0u   /* >ref|left_start|left-right|right_end|GTAG|fwd  (or |rev when aux != 0)                           */
#define THB_FLANK_DEL  1u   /* >ref|left_start|left-right|right_end|del|fwd                                                      */
#define THB_FLANK_INS  2u   /* >ref|left_start|left-SEQ|right_end|ins|fwd                                                        */
#define THB_FLANK_FUS  3u   /* >ref1-ref2|left_start|left-right|right_end|fus|{ff,fr,rf,rr}   (aux = dir, CigarOpCode 7..10)     */
typedef struct thb_flank_contig {        /* what juncs_db encodes in the FASTA name of one contig; contigs are in its output order  */
  uint32_t kind;                         /* THB_FLANK_*                                                                         */
  uint32_t ref_id, ref_id2;              /* 1-based; ref_id2 differs from ref_id only for fusions                                */
  uint32_t left_start;                   /* name field 2                                                                        */
  uint32_t left, right;                  /* junction / deletion / fusion coordinates; insertion: left, right unused              */
  uint32_t right_end;                    /* name field 4; 0xffffffff where juncs_db prints (size_t)-1 (fusion, right flank at 0) */
  uint32_t aux;                          /* junction: antisense; fusion: dir; insertion: inserted length                         */
  uint32_t length;                       /* bases of the contig                                                                 */
  char     ins_seq[20];                  /* insertion: NUL-terminated inserted bases                                            */
} thb_flank_contig;                      /* 56 bytes */
typedef struct thb_flank_params {
  int32_t max_mismatches;                /* bowtie -v: --segment-mismatches (0..3)                                              */
  int32_t max_multihits;                 /* bowtie -k / -m: --max-seg-multihits                                                 */
  int32_t min_seg_len, max_seg_len;      /* min_seg_len: shortest segment that will be submitted (the seeds cover its bases),
                                            4*(max_mismatches+2) <= min_seg_len; max_seg_len: juncs_db's <read_length>, the flank
                                            length on either side of an event (tophat.py:3483-3492), <= 56.  Segments of up to 64
                                            bases may be submitted; one longer than a contig has no placement on it.             */
  int32_t min_anchor;                    /* juncs_db's <min_anchor>: tophat.py passes 3                                          */
  int32_t ref_n_is_mismatch;             /* 0: bowtie 1 (placement over a contig N invalid); 1: the N counts as a mismatch       */
} thb_flank_params;
typedef struct thb_flank_batch {
  uint32_t        n_reads;               /* < 2^27                                                                              */
  uint32_t        read_words;            /* 64-bit words per bit plane of a read                                                */
  uint32_t        n_segs;                /* segments per read (<= THB_MAX_SEGS)                                                  */
  uint32_t        reserved;
  const uint64_t* reads;                 /* [n_reads*3*read_words] as in thb_segjuncs_batch                                      */
  uint16_t        seg_bounds[THB_MAX_SEGS + 1];   /* segment k = read bases [seg_bounds[k], seg_bounds[k+1])                     */
} thb_flank_batch;
typedef struct thb_flank_hit { uint32_t read; uint32_t contig; uint8_t seg; uint8_t pos; uint8_t antisense; uint8_t mismatches; } thb_flank_hit; /* 12 bytes */
typedef struct thb_flank_timing { float index_ms;                 /* thb_flank_begin: contigs + seed entries + sort + bucket table  */
                                  float h2d_ms, match_ms, post_ms, d2h_ms;   /* last submit: upload, search kernel, -m filter + sort + decode, download */
                                  uint64_t n_contigs, n_index_entries, n_verified, n_hits, algorithmic_bytes;
                                  uint32_t launches; uint32_t reserved; } thb_flank_timing;
/* The four sets as thb_segjuncs_finish returns them (sorted, unique, in their set orders); any may be empty.  Needs thb_ref_upload.       */
int thb_flank_begin(thb_ctx* ctx, const thb_flank_params* params,
                    const thb_junction* junctions, uint64_t n_junctions, const thb_junction* deletions, uint64_t n_deletions,
                    const thb_insertion* insertions, uint64_t n_insertions, const thb_fusion* fusions, uint64_t n_fusions);
/* The contigs of the index, in juncs_db's output order (owned by ctx, valid until the next thb_flank_begin). */
int thb_flank_contigs(thb_ctx* ctx, const thb_flank_contig** contigs, uint64_t* n_contigs);
/* Searches one batch (host arrays); *hits is owned by ctx and valid until the next thb_flank call. */
int thb_flank_submit(thb_ctx* ctx, const thb_flank_batch* host_batch, const thb_flank_hit** hits, uint64_t* n_hits);
/* Same for reads that already live in device memory (e.g. the `reads` array of a batch submitted to the other stages). */
int thb_flank_submit_device(thb_ctx* ctx, const thb_flank_batch* device_batch, const thb_flank_hit** hits, uint64_t* n_hits);
/* The placements of the LAST submit as the BowtieHits SplicedBAMHitFactory makes of them (bwt_map.cpp:1469-1770: genomic left, the
 * CIGAR with the event spliced in, mismatches, splice_mms within min_anchor_len of a gap, strand flags, second contig of a fusion in
 * ops[8]) -- record i belongs to hit i of that submit; n_ops == 0 where the reference discards the hit (it does not reach over the
 * event).  After thb_flank_submit_device the batch's device reads must still be alive.  Owned by ctx until the next thb_flank call. */
int thb_flank_spliced_hits(thb_ctx* ctx, int min_anchor_len, const thb_jhit_full** jhits, uint64_t* n);
int thb_flank_last_timing(thb_ctx* ctx, thb_flank_timing* out);

/* Page-locked host memory for batch arrays (full-speed, asynchronous host->device copies).
 * Returns NULL on failure. */
void* thb_alloc_pinned(size_t bytes);
void  thb_free_pinned(void* p);

/* Raw access for the measurement harness: CUDA stream of the context (cudaStream_t as void*). */
void* thb_stream(thb_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* TOPHAT_B200_H */
