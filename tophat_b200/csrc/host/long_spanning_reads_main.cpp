// long_spanning_reads -- drop-in replacement of the reference stage binary (src/long_spanning_reads.cpp main/driver,
// 2870-3329): same argv, same inputs, BAM output with the same records.  The host merges the id-sorted segment
// streams the way JoinSegmentsWorker does (2706-2785, look_right_for_hit_group 87-163), packs per-read bundles, runs
// the chain join on the GPU through libtophat_b200.so, and then does what is left of the worker loop: sort + unique
// of a read's alignments (2805-2807), the read-level filters (2810-2813), bowtie_sam_extra (bwt_map.cpp:2467-2648) and
// print_bamhit (1888-2093).  There is no CPU fallback for the join itself.
#include <algorithm>
#include <chrono>
#include <thread>
#include <atomic>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <map>
#include <set>
#include <string>
#include <vector>
#include "tophat_b200.h"
#include "thb_options.hpp"
#include "thb_input.hpp"
#include "thb_join_input.hpp"
#include "thb_bamwrite.hpp"

using namespace thbhost;

static void print_usage()
{
  fprintf(stderr, "Usage:   long_spanning_reads <reads.fa/.fq> <possible_juncs1,...,possible_juncsN> <possible_insertions1,...,possible_insertionsN> "
                  "<possible_deletions1,...,possible_deletionsN> <seg1.bwtout,...,segN.bwtout> [spliced_seg1.bwtout,...,spliced_segN.bwtout]\n");
}
[[noreturn]] static void die(const char* fmt, const std::string& a = "") { fprintf(stderr, fmt, a.c_str()); fputc('\n', stderr); exit(1); }

// CigarOpCode (bwt_map.h:36-55) -> BAM op code / SAM letter
static int bam_op(int code) { switch (code) { case 1: return 0; case 3: return 1; case 5: return 2; case 11: return 3; case 13: return 4; case 14: return 5; case 15: return 6; default: return 0; } }

struct Joined {
  uint32_t ref_id, ref_id2; int32_t left; bool anti, asplice, fused, seq_rc; uint8_t mism, edit, smm; std::vector<uint32_t> ops;   // ops: len << 4 | CigarOpCode
};
// BowtieHit::operator< (bwt_map.h:180-207) within one read
static bool joined_less(const Joined& a, const Joined& b)
{
  if (a.ref_id != b.ref_id) return a.ref_id < b.ref_id;
  if (a.ref_id2 != b.ref_id2) return a.ref_id2 < b.ref_id2;
  if (a.left != b.left) return a.left < b.left;
  if (a.anti != b.anti) return a.anti < b.anti;
  if (a.mism != b.mism) return a.mism < b.mism;
  if (a.edit != b.edit) return a.edit < b.edit;
  if (a.ops != b.ops) {
    if (a.ops.size() != b.ops.size()) return a.ops.size() < b.ops.size();
    for (size_t i = 0; i < a.ops.size(); ++i) if (a.ops[i] != b.ops[i]) {
      const uint32_t ac = a.ops[i] & 15, bc = b.ops[i] & 15;
      return ac < bc || (ac == bc && (a.ops[i] >> 4) < (b.ops[i] >> 4));
    }
  }
  return false;
}
// BowtieHit::operator== (167-178)
static bool joined_equal(const Joined& a, const Joined& b)
{ return a.ref_id == b.ref_id && a.ref_id2 == b.ref_id2 && a.anti == b.anti && a.left == b.left && a.asplice == b.asplice && a.edit == b.edit && a.ops == b.ops; }

// CIGAR text of print_bamhit (bwt_map.cpp:1944-2003): lower-case letters for the ops of a part read leftwards, <pos + 1>F for a fusion
static std::string cigar_text(const std::vector<uint32_t>& ops)
{
  std::string t;
  for (uint32_t op : ops) {
    const int c = (int)(op & 15); const uint32_t l = op >> 4;
    switch (c) {
      case 1: t += std::to_string(l) + "M"; break; case 2: t += std::to_string(l) + "m"; break;
      case 3: t += std::to_string(l) + "I"; break; case 4: t += std::to_string(l) + "i"; break;
      case 5: t += std::to_string(l) + "D"; break; case 6: t += std::to_string(l) + "d"; break;
      case 11: t += std::to_string(l) + "N"; break; case 12: t += std::to_string(l) + "n"; break;
      case 7: case 8: case 9: case 10: t += std::to_string(l + 1) + "F"; break;
      default: break;
    }
  }
  return t;
}

// extract_partial_hits (bwt_map.cpp:2148-2346): the two single-contig records of a fusion alignment.  `seq` / `qual` are the
// (already oriented) sequence of the whole alignment; BAM CIGARs come out as len << 4 | BAM op.
struct FusionParts { std::vector<uint32_t> cig1, cig2; std::string seq1, seq2, qual1, qual2; int left1 = -1, left2 = -1; };
static void reverse_complement(std::string& s);
static FusionParts split_fusion(const Joined& J, const std::string& seq, const std::string& qual)
{
  FusionParts P; int right = J.left, fusion_left = -1, fusion_right = -1, dir = 0; size_t fidx = 0, left_len = 0;
  for (size_t c = 0; c < J.ops.size(); ++c) {
    const int code = (int)(J.ops[c] & 15); const int l = (int)(J.ops[c] >> 4);
    if (code == 1 || code == 11 || code == 5) right += l; else if (code == 2 || code == 12 || code == 6) right -= l;
    else if (code >= 7 && code <= 10) { dir = code; fidx = c; fusion_left = (code == 7 || code == 8) ? right - 1 : right + 1; fusion_right = right = l; }
    if (!dir && (code == 1 || code == 2 || code == 3 || code == 4)) left_len += (size_t)l;
  }
  auto bam_of = [](int code) -> uint32_t { switch (code) { case 1: case 2: return 0; case 3: case 4: return 1; case 5: case 6: return 2; case 11: case 12: return 3; default: return 0; } };
  if (dir == 7 || dir == 8) for (size_t c = 0; c < fidx; ++c) P.cig1.push_back(((J.ops[c] >> 4) << 4) | bam_of((int)(J.ops[c] & 15)));
  else if (dir == 9 || dir == 10) for (size_t c = fidx; c-- > 0;) P.cig1.push_back(((J.ops[c] >> 4) << 4) | bam_of((int)(J.ops[c] & 15)));
  if (dir == 7 || dir == 9) for (size_t c = fidx + 1; c < J.ops.size(); ++c) P.cig2.push_back(((J.ops[c] >> 4) << 4) | bam_of((int)(J.ops[c] & 15)));
  else if (dir == 8 || dir == 10) for (size_t c = J.ops.size() - 1; c > fidx; --c) P.cig2.push_back(((J.ops[c] >> 4) << 4) | bam_of((int)(J.ops[c] & 15)));
  if (left_len > seq.size()) left_len = seq.size();
  P.seq1 = seq.substr(0, left_len); P.qual1 = qual.substr(0, std::min(left_len, qual.size()));
  if (dir == 9 || dir == 10) { reverse_complement(P.seq1); std::reverse(P.qual1.begin(), P.qual1.end()); }
  P.seq2 = seq.substr(left_len); P.qual2 = left_len < qual.size() ? qual.substr(left_len) : std::string();
  if (dir == 8 || dir == 10) { reverse_complement(P.seq2); std::reverse(P.qual2.begin(), P.qual2.end()); }
  P.left1 = (dir == 7 || dir == 8) ? J.left : fusion_left;
  P.left2 = (dir == 7 || dir == 9) ? fusion_right : right + 1;
  return P;
}

struct GenomeView {
  const Genome* g;
  char base(uint32_t ref_id, int64_t pos) const {      // Dna5 character
    const uint64_t gl = g->contig_start[ref_id - 1] + (uint64_t)pos; const uint64_t b = gl >> 6; const unsigned j = (unsigned)(gl & 63);
    if ((g->nmask[b] >> j) & 1) return 'N';
    return "ACGT"[((g->planes[2 * b] >> j) & 1) | (((g->planes[2 * b + 1] >> j) & 1) << 1)];
  }
};

static void reverse_complement(std::string& s)
{
  std::reverse(s.begin(), s.end());
  for (char& c : s) switch (c) { case 'A': c = 'T'; break; case 'C': c = 'G'; break; case 'G': c = 'C'; break; case 'T': c = 'A'; break; default: c = 'N'; }
}

int main(int argc, char** argv)
{
  fprintf(stderr, "long_spanning_reads v%s (%s)\n", "2.1.2", "tophat_b200");
  fprintf(stderr, "--------------------------------------------\n");
  Options o;
  if (parse_options(argc, argv, o, print_usage)) return 1;
  const std::vector<std::string>& a = o.positional;
  if (a.size() < 8) { print_usage(); return 1; }
  const std::string ref_fname = a[0], reads_fname = a[1];
  const std::vector<std::string> juncs_files = split_list(a[2]), ins_files = split_list(a[3]), del_files = split_list(a[4]);
  const std::string bam_out = a[6];
  const std::vector<std::string> seg_files = split_list(a[7]);
  std::vector<std::string> spl_files; if (a.size() >= 9) spl_files = split_list(a[8]);
  if (seg_files.empty()) { fprintf(stderr, "No hits to process, exiting\n"); return 0; }          // 2883-2887
  if (o.color) die("Error: colorspace reads are outside the GPU path");
  if (seg_files.size() > THB_MAX_SEGS) die("Error: more than %s segments per read are not supported by the GPU path", std::to_string(THB_MAX_SEGS));

  thb_ctx* ctx = nullptr;
  const char* dev_env = getenv("TOPHAT_GPU_DEVICE");
  // the CUDA context (0.3 - 1.5 s per process) comes up on its own thread while this one loads the genome
  int create_rc = THB_OK; std::thread ctx_thread([&] { create_rc = thb_create(dev_env ? atoi(dev_env) : 0, &ctx); });
  auto join_ctx = [&]() { if (ctx_thread.joinable()) ctx_thread.join(); if (create_rc != THB_OK) die("Error: %s", thb_last_error(nullptr)); };
  auto t0 = std::chrono::steady_clock::now();
  RefTable rt; std::string err;
  if (!o.sam_header.empty() && !rt.load_sam_header(o.sam_header, &err)) die("%s", err);
  fprintf(stderr, "Loading reference sequences...\n");
  Genome g;
  if (!load_fasta(ref_fname, rt, g, false, std::max(8, o.num_threads), &err)) { join_ctx(); die("Error: %s", err); }
  join_ctx();
  fprintf(stderr, "        reference sequences loaded.\n");

  // junctions + deletions -> std::set<Junction>; insertions -> std::set<Insertion> (2895-2980)
  struct JL { bool operator()(const thb_junction& x, const thb_junction& y) const {
    if (x.ref_id != y.ref_id) return x.ref_id < y.ref_id; if (x.left != y.left) return x.left < y.left;
    if (x.right != y.right) return x.right < y.right; return x.antisense < y.antisense; } };
  struct IL { bool operator()(const thb_insertion& x, const thb_insertion& y) const {
    if (x.ref_id != y.ref_id) return x.ref_id < y.ref_id; if (x.left != y.left) return x.left < y.left; return x.len < y.len; } };
  std::set<thb_junction, JL> jset; std::set<thb_insertion, IL> iset;
  std::set<thb_junction, JL> jonly, donly;      // the junction / deletion files on their own: what juncs_db reads (TOPHAT_GPU_FLANK_SEARCH)
  fprintf(stderr, "Loading junctions...");
  for (const std::string& f : juncs_files) {
    FILE* fp = fopen(f.c_str(), "r"); if (!fp) die("Error: cannot open %s for reading", f);
    char buf[2048];
    while (fgets(buf, sizeof buf, fp)) { char name[256]; int l, r; char orient;
      if (sscanf(buf, "%255s %d %d %c", name, &l, &r, &orient) != 4) continue;
      thb_junction j; j.ref_id = rt.get_id(name); j.left = (uint32_t)l; j.right = (uint32_t)r; j.antisense = orient == '-'; jset.insert(j); jonly.insert(j); }
    fclose(fp);
  }
  fprintf(stderr, "done\nLoading deletions...");
  for (const std::string& f : del_files) {
    FILE* fp = fopen(f.c_str(), "r"); if (!fp) continue;
    char buf[2048];
    while (fgets(buf, sizeof buf, fp)) { char* nl = strrchr(buf, '\n'); if (nl) *nl = 0;
      char* t1 = strchr(buf, '\t'); if (!t1) die("Error: malformed deletion coordinate record"); *t1++ = 0;
      char* t2 = strchr(t1, '\t'); if (!t2) die("Error: malformed deletion coordinate record"); *t2++ = 0;
      char* t3 = strchr(t2, '\t'); if (t3) *t3 = 0;
      thb_junction j; j.ref_id = rt.get_id(buf); j.left = (uint32_t)atoi(t1) - 1u; j.right = (uint32_t)atoi(t2); j.antisense = 0; jset.insert(j); donly.insert(j); }
    fclose(fp);
  }
  fprintf(stderr, "done\nLoading insertions...");
  for (const std::string& f : ins_files) {
    FILE* fp = fopen(f.c_str(), "r"); if (!fp) continue;
    char buf[2048];
    while (fgets(buf, sizeof buf, fp)) { char* nl = strrchr(buf, '\n'); if (nl) *nl = 0;
      char* t1 = strchr(buf, '\t'); if (!t1) die("Error: malformed insertion coordinate record"); *t1++ = 0;
      char* t2 = strchr(t1, '\t'); if (!t2) die("Error: malformed insertion coordinate record"); *t2++ = 0;
      char* t3 = strchr(t2, '\t'); if (!t3) die("Error: malformed insertion coordinate record"); *t3++ = 0;
      char* t4 = strchr(t3, '\t'); if (t4) *t4 = 0;
      thb_insertion in; memset(&in, 0, sizeof in); in.ref_id = rt.get_id(buf); in.left = (uint32_t)atoi(t1); in.len = (uint32_t)strlen(t3);
      if (in.len > 19) die("Error: insertion longer than 19 bases is not supported by the GPU path");
      memcpy(in.seq, t3, in.len); iset.insert(in); }
    fclose(fp);
  }
  fprintf(stderr, "done\n");
  // fusions (2996-3040)
  struct FL { bool operator()(const thb_fusion& x, const thb_fusion& y) const {
    if (x.ref_id1 != y.ref_id1) return x.ref_id1 < y.ref_id1; if (x.ref_id2 != y.ref_id2) return x.ref_id2 < y.ref_id2;
    if (x.left != y.left) return x.left < y.left; if (x.right != y.right) return x.right < y.right; return x.dir < y.dir; } };
  std::set<thb_fusion, FL> fset;
  if (o.p.fusion_search) {
    fprintf(stderr, "Loading fusions...");
    for (const std::string& f : split_list(a[5])) {
      FILE* fp = fopen(f.c_str(), "r"); if (!fp) continue;
      char buf[2048];
      while (fgets(buf, sizeof buf, fp)) { char* nl = strrchr(buf, '\n'); if (nl) *nl = 0;
        char* t1 = strchr(buf, '\t'); if (!t1) die("Error: malformed insertion coordinate record"); *t1++ = 0;
        char* t2 = strchr(t1, '\t'); if (!t2) die("Error: malformed insertion coordinate record"); *t2++ = 0;
        char* t3 = strchr(t2, '\t'); if (!t3) die("Error: malformed insertion coordinate record"); *t3++ = 0;
        char* t4 = strchr(t3, '\t'); if (!t4) die("Error: malformed insertion coordinate record"); *t4++ = 0;
        char* t5 = strchr(t4, '\t'); if (t5) *t5 = 0;
        thb_fusion fu; memset(&fu, 0, sizeof fu); fu.ref_id1 = rt.get_id(buf); fu.left = (uint32_t)atoi(t1); fu.ref_id2 = rt.get_id(t2); fu.right = (uint32_t)atoi(t3);
        fu.dir = strcmp(t4, "fr") == 0 ? 8u : strcmp(t4, "rf") == 0 ? 9u : strcmp(t4, "rr") == 0 ? 10u : 7u;
        fset.insert(fu); }
      fclose(fp);
    }
    fprintf(stderr, "done\n");
  }
  // contigs named only by the junction files must exist in the genome image before it is uploaded
  if (g.contig_len.size() < rt.size()) {
    Genome g2 = g; const size_t old = g.contig_len.size(); uint64_t gpos = (g.n_blocks ? (g.n_blocks - 1) * 64 : 0);
    for (size_t i = old; i < rt.size(); ++i) { g2.contig_start.push_back(gpos); g2.contig_len.push_back(0); gpos += 64; }
    g2.n_blocks = gpos / 64 + 1; g2.planes.resize(2 * g2.n_blocks, 0); g2.nmask.resize(g2.n_blocks, 0); g = g2;
  }
  std::vector<thb_junction> jv(jset.begin(), jset.end()); std::vector<thb_insertion> iv(iset.begin(), iset.end());

  std::mutex rtm;
  { thb_ref_image img = g.image(); if (thb_ref_upload(ctx, &img) != THB_OK) die("Error: thb_ref_upload: %s", thb_last_error(ctx)); }
  if (thb_join_begin(ctx, &o.p, jv.data(), jv.size(), iv.data(), iv.size()) != THB_OK) die("Error: thb_join_begin: %s", thb_last_error(ctx));
  if (o.p.fusion_search) { std::vector<thb_fusion> fv(fset.begin(), fset.end());
    if (thb_join_set_fusions(ctx, fv.data(), fv.size()) != THB_OK) die("Error: thb_join_set_fusions: %s", thb_last_error(ctx)); }
  // TOPHAT_GPU_FLANK_SEARCH=1: the junction index is searched in this process (thb_flank_*) instead of reading the segment hits that
  // juncs_db + bowtie-build + bowtie + fix_map_ordering would have left in <spliced_segK.bwtout> (tophat.py:2546-2600, 3686-3741);
  // a spliced argument, if given, is ignored.  The flank length is tophat.py's max_seg_len (3483-3492) for the longest read.
  const bool flank_search = getenv("TOPHAT_GPU_FLANK_SEARCH") != nullptr && atoi(getenv("TOPHAT_GPU_FLANK_SEARCH")) != 0;
  const int seglen = o.p.segment_length;
  auto segment_bounds = [seglen](int L, std::vector<uint16_t>& b) {       // split_reads, tophat.py:2975-2990
    b.clear(); int n = L / seglen;
    for (int i = 0; i <= n; ++i) b.push_back((uint16_t)(seglen * i));
    if (L % seglen >= std::min(seglen - 2, 20)) { b.push_back((uint16_t)L); ++n; } else b.back() = (uint16_t)L;
    if (n <= 1) { b.clear(); b.push_back(0); b.push_back((uint16_t)L); }
  };
  if (flank_search) {
    if (!spl_files.empty()) { fprintf(stderr, "TOPHAT_GPU_FLANK_SEARCH: ignoring the spliced segment files\n"); spl_files.clear(); }
    int max_len = 0, min_seg = 1 << 30;
    { FullReadStream all(reads_fname);                       // lengths only: the flank length follows the longest read
      std::vector<uint16_t> b; std::vector<char> seen(256, 0);
      while (const FullRead* r = all.next()) {
        const int L = (int)r->seq.size(); if (L > 255) die("Error: reads longer than 255 bases are not supported");
        if (L < seglen || seen[(size_t)L]) continue; seen[(size_t)L] = 1; max_len = std::max(max_len, L);
        segment_bounds(L, b); for (size_t k = 0; k + 1 < b.size(); ++k) min_seg = std::min(min_seg, (int)b[k + 1] - (int)b[k]); }
      if (!all.ok()) die("Error: %s", all.error()); }
    if (max_len == 0) { max_len = seglen; min_seg = seglen; }
    int max_seg_len = seglen;
    { int n = max_len / seglen; if (!(max_len % seglen >= std::min(seglen - 2, 20)) && n > 1) max_seg_len += max_len % seglen; }
    if (const char* e = getenv("TOPHAT_GPU_FLANK_LENGTH")) if (atoi(e) >= seglen) max_seg_len = atoi(e);      // juncs_db's <read_length> given explicitly
    thb_flank_params fp; memset(&fp, 0, sizeof fp);
    fp.max_mismatches = std::min(o.p.segment_mismatches, 3); fp.max_multihits = o.p.max_seg_multihits; fp.min_seg_len = std::min(min_seg, max_seg_len);
    fp.max_seg_len = max_seg_len; fp.min_anchor = 3; fp.ref_n_is_mismatch = o.p.bowtie2 ? 1 : 0;
    std::vector<thb_junction> jo(jonly.begin(), jonly.end()), dn(donly.begin(), donly.end());
    std::vector<thb_fusion> fv(fset.begin(), fset.end());
    if (thb_flank_begin(ctx, &fp, jo.data(), jo.size(), dn.data(), dn.size(), iv.data(), iv.size(), fv.data(), fv.size()) != THB_OK)
      die("Error: thb_flank_begin: %s", thb_last_error(ctx));
    thb_flank_timing ft; thb_flank_last_timing(ctx, &ft);
    fprintf(stderr, "Junction index: %llu contigs, %llu seed entries, built in %.1f ms\n", (unsigned long long)ft.n_contigs, (unsigned long long)ft.n_index_entries, ft.index_ms);
  }
  auto t1 = std::chrono::steady_clock::now();

  // Read-id ranges, one output BAM per range -- the reference's own -p N layout (<out minus .bam><i>.bam, 3056-3064, which
  // tophat.py:3775-3779 picks up): contiguous id ranges split at entries of the reads file's .index, every stream positioned by
  // its own index and filtered by id.  The library is entered by one range at a time (the kernels are a small part of a range).
  std::vector<uint32_t> splits;
  if (o.num_threads > 1) { BamIndex ix; if (ix.load(reads_fname) || ix.load(seg_files.back())) splits = split_ids(ix, std::min(o.num_threads, 64)); }
  const size_t n_ranges = splits.size() + 1;
  // tophat.py prefers <out>.bam over the per-range files when it exists and is not empty (3772-3779): one left over from an
  // earlier run must not shadow this run's output
  if (n_ranges > 1 && bam_out.compare(0, 5, "/dev/") != 0) { remove(bam_out.c_str()); remove((bam_out + ".index").c_str()); }
  std::mutex ctx_mutex, stat_mutex;
  uint64_t n_reads = 0, n_out = 0;
  double submit_s = 0, post_s = 0;           // TOPHAT_GPU_STATS: time inside thb_join_submit / in sorting, SAM fields and BAM output (summed over ranges)
  const GenomeView gv{&g};
  const size_t nseg = seg_files.size();
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };

  auto run_range = [&](size_t ri) {
  const uint32_t begin_id = ri ? splits[ri - 1] : 0u, end_id = ri < splits.size() ? splits[ri] : 0xffffffffu;
  std::string out_path = bam_out, err;
  if (n_ranges > 1) out_path = bam_out.substr(0, bam_out.size() >= 4 ? bam_out.size() - 4 : bam_out.size()) + std::to_string(ri) + ".bam";
  const int T = n_ranges > 1 ? 1 : std::max(1, std::min(o.num_threads, 64));       // record-building threads of one flush
  std::vector<std::unique_ptr<JoinHitStream>> contig, spliced;
  for (auto& f : seg_files) contig.emplace_back(new JoinHitStream(f, rt, rtm, false, o.p.max_report_intron_length, o.p.min_anchor_len, range_for(f, begin_id, end_id)));
  for (auto& f : spl_files) spliced.emplace_back(new JoinHitStream(f, rt, rtm, true, o.p.max_report_intron_length, o.p.min_anchor_len, range_for(f, begin_id, end_id)));
  FullReadStream reads(reads_fname, range_for(reads_fname, begin_id, end_id));
  uint64_t r_reads = 0, r_out = 0; double r_submit = 0, r_post = 0;
  // in-process junction index search: the placements of this range's reads, per segment, in id order
  std::vector<std::vector<JHitRec>> mem_spl(flank_search ? nseg : 0); std::vector<size_t> mem_pos(mem_spl.size(), 0);
  if (flank_search) {
    FullReadStream pre(reads_fname, range_for(reads_fname, begin_id, end_id));
    std::map<int, std::pair<std::vector<uint32_t>, std::vector<uint64_t>>> by_len;        // read length -> ids, bit planes (read_words = 4)
    auto search = [&](int L, std::vector<uint32_t>& ids, std::vector<uint64_t>& planes) {
      if (ids.empty()) return;
      std::vector<uint16_t> b; segment_bounds(L, b);
      thb_flank_batch fb; memset(&fb, 0, sizeof fb);
      fb.n_reads = (uint32_t)ids.size(); fb.read_words = 4; fb.n_segs = (uint32_t)b.size() - 1; fb.reads = planes.data();
      if (fb.n_segs > nseg) die("Error: a read of %s bases has more segments than segment files were given", std::to_string(L));
      for (size_t k = 0; k < b.size(); ++k) fb.seg_bounds[k] = b[k];
      std::lock_guard<std::mutex> cl(ctx_mutex);
      const thb_flank_hit* fh = nullptr; const thb_jhit_full* jh = nullptr; uint64_t n = 0, n2 = 0;
      if (thb_flank_submit(ctx, &fb, &fh, &n) != THB_OK) die("Error: thb_flank_submit: %s", thb_last_error(ctx));
      if (thb_flank_spliced_hits(ctx, o.p.min_anchor_len, &jh, &n2) != THB_OK || n2 != n) die("Error: thb_flank_spliced_hits: %s", thb_last_error(ctx));
      for (uint64_t i = 0; i < n; ++i) if (jh[i].n_ops) { JHitRec r; r.id = ids[fh[i].read]; r.h = jh[i]; mem_spl[fh[i].seg].push_back(r); }
      ids.clear(); planes.clear();
    };
    while (const FullRead* r = pre.next()) {
      const int L = (int)r->seq.size();
      if (L < seglen || L > 255) continue;
      auto& slot = by_len[L];
      ReadRec rr; pack_read_ascii(r->seq.data(), (uint32_t)L, rr);
      slot.first.push_back(r->id); slot.second.insert(slot.second.end(), rr.planes, rr.planes + 12);
      if (slot.first.size() >= (1u << 20)) search(L, slot.first, slot.second);
    }
    if (!pre.ok()) die("Error: %s", pre.error());
    for (auto& kv : by_len) search(kv.first, kv.second.first, kv.second.second);
    for (auto& v : mem_spl) std::stable_sort(v.begin(), v.end(), [](const JHitRec& x, const JHitRec& y) { return x.id < y.id; });
  }
  const size_t n_spl = flank_search ? nseg : spliced.size();
  auto spl_next_id = [&](size_t s) -> uint32_t {
    if (!flank_search) return spliced[s]->next_group_id();
    return mem_pos[s] < mem_spl[s].size() ? mem_spl[s][mem_pos[s]].id : 0u; };
  auto spl_next_group = [&](size_t s, std::vector<thb_jhit_full>& out) {
    if (!flank_search) { spliced[s]->next_group(out); return; }
    const uint32_t id = spl_next_id(s); if (!id) return;
    while (mem_pos[s] < mem_spl[s].size() && mem_spl[s][mem_pos[s]].id == id) out.push_back(mem_spl[s][mem_pos[s]++].h); };
  auto spl_skip = [&](size_t s) {
    if (!flank_search) { spliced[s]->skip_group(); return; }
    const uint32_t id = spl_next_id(s); while (mem_pos[s] < mem_spl[s].size() && mem_spl[s][mem_pos[s]].id == id) ++mem_pos[s]; };
  BamWriter bw;
  if (!bw.open(out_path, o.sam_header, out_path + ".index", &err)) die("Error: %s", err);
  std::vector<int> ref2tid;

  struct Pending { uint32_t id; FullRead read; };
  std::vector<thb_join_bundle> bundles; std::vector<uint16_t> segc; std::vector<uint64_t> rplanes; std::vector<thb_jhit> hits; std::vector<thb_jops> ops_ext; std::vector<Pending> pend;
  std::vector<thb_jhit_full> read_hits; std::vector<thb_joined> joined_copy;
  auto flush = [&]() {
    if (bundles.empty()) return;
    const auto f0 = now();
    // reads of one batch share read_words = 4 (reads up to 255 bases)
    thb_join_batch jb; memset(&jb, 0, sizeof jb);
    jb.n_bundles = (uint32_t)bundles.size(); jb.n_segs = (uint32_t)nseg; jb.read_words = 4; jb.bundles = bundles.data(); jb.seg_count = segc.data();
    jb.reads = rplanes.data(); jb.n_hits = hits.size(); jb.hits = hits.data(); jb.n_ops_ext = ops_ext.size(); jb.ops_ext = ops_ext.data();
    // the result buffer belongs to the context and lives until the next join call: copied out before the next range enters
    uint64_t no = 0;
    { std::lock_guard<std::mutex> cl(ctx_mutex);
      const thb_joined* res = nullptr;
      if (thb_join_submit(ctx, &jb, &res, &no) != THB_OK) die("Error: thb_join_submit: %s", thb_last_error(ctx));
      joined_copy.assign(res, res + no); }
    const thb_joined* out = joined_copy.data();
    const auto f1 = now(); r_submit += secs(f0, f1);
    std::vector<std::vector<Joined>> per(bundles.size());
    for (uint64_t i = 0; i < no; ++i) {
      const thb_joined& j = out[i]; Joined J; J.ref_id = j.ref_id; J.left = j.left; J.anti = (j.flags & THB_HIT_ANTISENSE) != 0;
      J.fused = (j.flags & THB_JOINED_FUSION) != 0; J.ref_id2 = o.p.fusion_search ? j.ops[THB_JOINED_MAX_OPS - 1] : j.ref_id;
      J.seq_rc = o.p.fusion_search ? (j.flags & THB_JOINED_SEQ_RC) != 0 : J.anti;
      J.asplice = (j.flags & THB_JHIT_ANTISENSE_SPLICE) != 0; J.mism = j.mismatches; J.edit = j.edit_dist; J.smm = j.splice_mms;
      J.ops.assign(j.ops, j.ops + j.n_ops); per[j.bundle].push_back(std::move(J));
    }
    // the records of the batch are built on -p threads (each a contiguous slice of the reads, so the output order is kept),
    // then handed to the writer, which deflates the BGZF blocks in parallel too
    { std::lock_guard<std::mutex> l(rtm);
      ref2tid.assign(rt.size() + 1, -1);
      for (uint32_t id = 1; id <= rt.size(); ++id) ref2tid[id] = bw.target_id(rt.name(id)); }
    std::vector<BamWriter::RecordPart> parts((size_t)T);
    auto build_slice = [&](int t) {
    BamWriter::RecordPart& part = parts[(size_t)t];
    std::vector<uint8_t> aux; std::vector<uint32_t> bcig; std::string MD;
    const size_t nbu = bundles.size();
    for (size_t b = nbu * (size_t)t / (size_t)T; b < nbu * (size_t)(t + 1) / (size_t)T; ++b) {
      std::vector<Joined>& v = per[b];
      std::sort(v.begin(), v.end(), joined_less);
      v.erase(std::unique(v.begin(), v.end(), joined_equal), v.end());
      const FullRead& rd = pend[b].read;
      for (const Joined& J : v) {
        int gap = 0, rlen = 0; bool has_splice = false;
        for (uint32_t op : J.ops) { const int c = (int)(op & 15), l = (int)(op >> 4);
          if (c == 3 || c == 4 || c == 5 || c == 6) gap += l; if (c == 1 || c == 2 || c == 3 || c == 4 || c == 13) rlen += l; if (c == 11 || c == 12) has_splice = true; }
        if ((int)J.mism > o.p.read_mismatches || gap > o.p.read_gap_length || (int)J.edit > o.p.read_edit_dist) continue;   // 2810-2813
        // the hit's own sequence / qualities (merge_chain 1959-1979, BowtieHit::reverse): the read or its reverse complement
        std::string hseq = rd.seq, hqual = rd.qual;
        if (J.seq_rc) { reverse_complement(hseq); std::reverse(hqual.begin(), hqual.end()); }
        // bowtie_sam_extra (bwt_map.cpp:2467-2648); lower-case ops walk leftwards over the reverse-complemented reference, a
        // fusion op continues on the second contig
        aux.clear(); MD.clear();
        auto has_seq = [&](uint32_t id) { return id >= 1 && id <= g.contig_len.size() && g.contig_len[id - 1] > 0; };
        auto comp = [](char c) { return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : 'N'; };
        if (has_seq(J.ref_id) && has_seq(J.ref_id2)) {
          long pos_ref = J.left; size_t pos_seq = 0; int pos_mm = 0, mm = 0, gap_opens = 0, gap_conts = 0, AS = 0; uint32_t cur_ref = J.ref_id; bool saw_fusion = false, ok_extra = true;
          for (uint32_t op : J.ops) {
            const int c = (int)(op & 15), l = (int)(op >> 4);
            if (c == 1 || c == 2) {
              for (int k = 0; k < l; ++k, ++pos_seq) {
                const char rn = c == 1 ? gv.base(cur_ref, pos_ref + k) : comp(gv.base(cur_ref, pos_ref - k));
                const char sc = pos_seq < hseq.size() ? hseq[pos_seq] : 'N';
                if (sc != rn) {
                  ++mm;
                  if (pos_seq < hqual.size()) {
                    if (sc == 'N' || rn == 'N') AS -= o.p.bowtie2_penalty_for_N;
                    else { const float pen = o.p.bowtie2_min_penalty + (o.p.bowtie2_max_penalty - o.p.bowtie2_min_penalty) * std::min((int)(hqual[pos_seq] - '!'), 40) / 40.0f; AS -= (int)pen; }
                  }
                  MD += std::to_string(pos_mm); MD.push_back(rn); pos_mm = 0;
                } else { if (rn == 'N') AS -= o.p.bowtie2_penalty_for_N; ++pos_mm; }
              }
              pos_ref += c == 1 ? l : -l;
            } else if (c == 3 || c == 4) { pos_seq += (size_t)l; AS -= o.p.bowtie2_read_gap_open; AS -= o.p.bowtie2_read_gap_cont * l; gap_opens += 1; gap_conts += l; }
            else if (c == 5 || c == 6) {
              AS -= o.p.bowtie2_ref_gap_open; AS -= o.p.bowtie2_ref_gap_cont * l; gap_opens += 1; gap_conts += l;
              MD += std::to_string(pos_mm); MD.push_back('^');
              for (int k = 0; k < l; ++k) MD.push_back(c == 5 ? gv.base(cur_ref, pos_ref + k) : comp(gv.base(cur_ref, pos_ref - k)));
              pos_ref += c == 5 ? l : -l; pos_mm = 0;
            } else if (c == 11) pos_ref += l;
            else if (c == 12) pos_ref -= l;
            else if (c >= 7 && c <= 10) { if (saw_fusion) { ok_extra = false; break; } cur_ref = J.ref_id2; pos_ref = l; saw_fusion = true; }
          }
          if (ok_extra) {
            MD += std::to_string(pos_mm);
            BamWriter::aux_int(aux, "AS", AS); BamWriter::aux_int(aux, "XM", mm); BamWriter::aux_int(aux, "XO", gap_opens);
            BamWriter::aux_int(aux, "XG", gap_conts); BamWriter::aux_str(aux, "MD", MD);
          }
        }
        // print_bamhit (bwt_map.cpp:1888-2093): read sequence cut to the hit's read length, rc for antisense hits
        std::string seq = rd.seq, quals = rd.qual; seq.resize((size_t)rlen, '\0'); quals.resize((size_t)rlen, '\0');
        if (J.anti) { reverse_complement(seq); std::reverse(quals.begin(), quals.end()); }
        BamWriter::aux_int(aux, "NM", (int)J.mism + gap);
        if (has_splice) BamWriter::aux_char(aux, "XS", J.asplice ? '-' : '+');
        const int tid = J.ref_id < ref2tid.size() ? ref2tid[J.ref_id] : -1;
        if (!J.fused) {
          bcig.clear(); for (uint32_t op : J.ops) bcig.push_back(((op >> 4) << 4) | (uint32_t)bam_op((int)(op & 15)));
          const size_t before = part.bytes.size();
          BamWriter::encode(part.bytes, rd.name, J.anti ? 0x10 : 0, tid, J.left, 255, bcig, seq, quals, aux);
          part.sizes.push_back((uint32_t)(part.bytes.size() - before)); part.ids.push_back(pend[b].id);
          continue;
        }
        // a fusion alignment is written as two records, one per contig, each carrying the whole alignment in XF (2047-2083)
        const FusionParts FP = split_fusion(J, seq, quals);
        const int tid2 = J.ref_id2 < ref2tid.size() ? ref2tid[J.ref_id2] : -1;
        std::string n1, n2; { std::lock_guard<std::mutex> l(rtm); n1 = rt.name(J.ref_id); n2 = rt.name(J.ref_id2); }
        const std::string xf_tail = " " + n1 + "-" + n2 + " " + std::to_string(J.left + 1) + " " + cigar_text(J.ops) + " " + seq + " " + quals;
        for (int partno = 1; partno <= 2; ++partno) {
          std::vector<uint8_t> aux2 = aux; BamWriter::aux_str(aux2, "XF", std::to_string(partno) + xf_tail);
          const size_t before = part.bytes.size();
          if (partno == 1) BamWriter::encode(part.bytes, rd.name, J.anti ? 0x10 : 0, tid, FP.left1, 255, FP.cig1, FP.seq1, FP.qual1, aux2);
          else BamWriter::encode(part.bytes, rd.name, J.anti ? 0x10 : 0, tid2, FP.left2, 255, FP.cig2, FP.seq2, FP.qual2, aux2);
          part.sizes.push_back((uint32_t)(part.bytes.size() - before)); part.ids.push_back(pend[b].id);
        }
        continue;
      }
    }
    };
    if (T == 1) build_slice(0);
    else { std::vector<std::thread> th; for (int t = 1; t < T; ++t) th.emplace_back(build_slice, t); build_slice(0); for (auto& x : th) x.join(); }
    for (const auto& part : parts) r_out += part.sizes.size();
    bw.append_records(parts, T);
    bundles.clear(); segc.clear(); rplanes.clear(); hits.clear(); ops_ext.clear(); pend.clear();
    r_post += secs(f1, now());
  };

  // JoinSegmentsWorker::operator() (2671-2845)
  const size_t BATCH = n_ranges > 1 ? (1u << 18) : (1u << 20);
  std::vector<std::vector<thb_jhit_full>> seg_hits(nseg);
  for (;;) {
    const uint32_t cid = contig[0]->next_group_id();
    const uint32_t sid = n_spl == 0 ? 0 : spl_next_id(0);
    if (!cid && !sid) break;
    const uint32_t id = (cid && (!sid || cid <= sid)) ? cid : sid;
    for (auto& v : seg_hits) v.clear();
    if (cid == id) contig[0]->next_group(seg_hits[0]);
    if (sid == id) spl_next_group(0, seg_hits[0]);
    // look_right_for_hit_group (87-163): stop at the first segment without contiguous or spliced hits
    for (size_t s = 1; s < nseg; ++s) {
      uint32_t gidc; while ((gidc = contig[s]->next_group_id()) != 0 && gidc < id) contig[s]->skip_group();
      if (gidc == id) contig[s]->next_group(seg_hits[s]);
      if (s < n_spl) { uint32_t gids; while ((gids = spl_next_id(s)) != 0 && gids < id) spl_skip(s);
                       if (gids == id) spl_next_group(s, seg_hits[s]); }
      if (seg_hits[s].empty()) break;
    }
    int last_non_empty = (int)nseg - 1;
    while (last_non_empty >= 0 && seg_hits[last_non_empty].empty()) --last_non_empty;
    if (last_non_empty < 0) continue;
    if (!(seg_hits[last_non_empty][0].flags & THB_HIT_END)) continue;                                    // 2777-2785
    const FullRead* rd = reads.get(id);
    if (!rd) { if (!reads.ok()) die("Error: %s", reads.error()); fprintf(stderr, "Error: could not get read # %d from stream\n", (int)id); exit(1); }
    if (rd->seq.size() > 255) die("Error: reads longer than 255 bases are not supported");
    thb_join_bundle bu; memset(&bu, 0, sizeof bu);
    bu.read_id = id; bu.hit_begin = (uint32_t)hits.size(); bu.read_len = (uint16_t)rd->seq.size(); bu.n_segs = (uint8_t)(last_non_empty + 1);
    bu.ops_begin = (uint32_t)ops_ext.size();
    read_hits.clear();
    for (size_t s = 0; s < nseg; ++s) {
      const size_t c = (int)s <= last_non_empty ? seg_hits[s].size() : 0;
      if (c > 65535) die("Error: more than 65535 hits for one segment of one read");
      segc.push_back((uint16_t)c);
      if ((int)s <= last_non_empty) read_hits.insert(read_hits.end(), seg_hits[s].begin(), seg_hits[s].end());
    }
    {                                   // wire form: 16-byte records + the CIGARs of the multi-op hits
      const size_t h0 = hits.size(), e0 = ops_ext.size();
      hits.resize(h0 + read_hits.size()); ops_ext.resize(e0 + read_hits.size());
      const int ne = thb_join_pack_hits(read_hits.data(), (uint32_t)read_hits.size(), hits.data() + h0, ops_ext.data() + e0);
      if (ne < 0) die("Error: a read has more than 256 segment hits with a gapped / spliced CIGAR (outside the GPU path)");
      ops_ext.resize(e0 + (size_t)ne);
    }
    ReadRec rr; pack_read_ascii(rd->seq.data(), (uint32_t)rd->seq.size(), rr);
    rplanes.insert(rplanes.end(), rr.planes, rr.planes + 12);
    bundles.push_back(bu); pend.push_back(Pending{id, *rd}); ++r_reads;
    if (bundles.size() >= BATCH) flush();
  }
  flush();
  for (auto& h : contig) if (!h->ok()) die("Error: %s", h->error());
  for (auto& h : spliced) if (!h->ok()) die("Error: %s", h->error());
  if (!reads.ok()) die("Error: %s", reads.error());
  { uint64_t dropped = 0; for (auto& h : contig) dropped += h->dropped_long_cigars(); for (auto& h : spliced) dropped += h->dropped_long_cigars();
    if (dropped) die("Error: %s segment hits carry more than 9 CIGAR operations (outside the GPU path; the output would lack their alignments)", std::to_string(dropped)); }
  if (!bw.close(&err)) die("Error: %s", err);
  { std::lock_guard<std::mutex> sl(stat_mutex); n_reads += r_reads; n_out += r_out; submit_s += r_submit; post_s += r_post; }
  };   // run_range

  {
    std::atomic<size_t> next(0);
    auto worker = [&]() { for (;;) { const size_t i = next.fetch_add(1); if (i >= n_ranges) break; run_range(i); } };
    const int nthr = (int)std::max<size_t>(1, std::min<size_t>((size_t)std::max(1, o.num_threads), n_ranges));
    std::vector<std::thread> pool;
    for (int k = 1; k < nthr; ++k) pool.emplace_back(worker);
    worker();
    for (auto& th : pool) th.join();
  }
  auto t2 = std::chrono::steady_clock::now();
  if (getenv("TOPHAT_GPU_STATS")) {
    thb_join_timing tm; thb_join_last_timing(ctx, &tm);
    auto sec = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    fprintf(stderr, "{\"gpu_stats\": {\"reads\": %llu, \"records\": %llu, \"ref_load_s\": %.3f, \"join_s\": %.3f, \"submit_s\": %.3f, \"post_s\": %.3f, "
                    "\"kernel_ms\": %.3f, \"chains\": %llu, \"closures\": %llu}}\n",
            (unsigned long long)n_reads, (unsigned long long)n_out, sec(t0, t1), sec(t1, t2), submit_s, post_s, tm.kernel_ms, (unsigned long long)tm.n_chains,
            (unsigned long long)tm.n_closures);
  }
  thb_destroy(ctx);
  return 0;
}
