#include "thb_join_input.hpp"
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <zlib.h>

namespace thbhost {

namespace {

// CigarOpCode values of the reference (bwt_map.h:36-55)
enum { C_MATCH = 1, C_mATCH = 2, C_INS = 3, C_iNS = 4, C_DEL = 5, C_dEL = 6, C_FUSION_FF = 7, C_FUSION_FR = 8, C_FUSION_RF = 9, C_FUSION_RR = 10,
       C_REF_SKIP = 11, C_rEF_SKIP = 12, C_SOFT_CLIP = 13, C_HARD_CLIP = 14, C_PAD = 15 };
struct Op { int code; int len; };

inline uint32_t pack_op(const Op& o) { return ((uint32_t)o.len << 4) | (uint32_t)o.code; }

// ---- placing a junction-index hit back on the genome ------------------------------------------------------------------
// A segment that bowtie aligned to a juncs_db contig (the two flanks of a junction / deletion / fusion glued together, or
// a reference stretch with an insertion's bases spliced in) carries a CIGAR against that contig.  On the genome the same
// alignment has one more operation -- the event -- at reference offset `at` from the hit's start:
//   gap events (REF_SKIP, DEL, FUSION_*): zero-width on the contig; the operation covering `at` is cut in two around it, or,
//       when `at` falls on the boundary in front of an operation, the gap goes in front of that operation;
//   INS: the contig bases [at, at + len) are the inserted bases; what the hit's operations cover of them becomes one INS
//       operation (shortened when the hit starts inside the insertion).
// Behaviour contract = the reference's spliceCigar (bwt_map.cpp:678-883), including two quirks that change records:
//   * its appender lengthens an equal trailing operation AND still appends the new one (bwt_map.cpp:672-677);
//   * a hit that does not reach the event from the left (event at or before the hit's first base) or that ends exactly at it
//     never gains the two extra operations and is rejected by the final size test (874).
// Fusion events: the operations on the far side of the fusion point are the lower-case codes when that side is read
// leftwards (FR / RR: after the point, RF / RR: before it; bwt_map.h:36-55).
struct EventOnContig { int code; int at; int len; };

inline bool is_fusion_code(int c) { return c >= C_FUSION_FF && c <= C_FUSION_RR; }
inline int lower_case(int c) { return c == C_MATCH ? C_mATCH : c == C_INS ? C_iNS : c == C_DEL ? C_dEL : c == C_REF_SKIP ? C_rEF_SKIP : c; }

struct SplicedCigarBuilder {
  std::vector<Op>& out;
  void put(int code, int len)
  {
    if (len <= 0) return;
    if (!out.empty() && out.back().code == code) out.back().len += len;      // sic: and the operation is appended as well
    out.push_back(Op{code, len});
  }
};

// Returns false where the reference discards the hit.  `spl_mismatches` receives the MD mismatches the reference attributes to
// the event: within min_anchor_len of a gap, or on inserted bases.
bool splice_cigar(std::vector<Op>& spl, const std::vector<Op>& cigar, const std::vector<bool>& mism, int left, int spl_start,
                  int spl_len, int spl_code, int& spl_mismatches, int min_anchor_len)
{
  EventOnContig ev{spl_code, spl_start - left, spl_len};
  const bool fusion = is_fusion_code(spl_code), insertion = spl_code == C_INS;
  if (fusion && ev.at < 0) ev.at = -ev.at;
  const int ev_end = insertion ? ev.at + ev.len : ev.at;          // first contig offset behind the event
  if (ev_end <= 0) return false;                                  // the hit lies wholly behind the event: never spliced (871-874)
  const bool lower_before = spl_code == C_FUSION_RF || spl_code == C_FUSION_RR;
  const bool lower_after = spl_code == C_FUSION_FR || spl_code == C_FUSION_RR;
  SplicedCigarBuilder b{spl};
  int ref_at = 0, read_at = 0; bool behind_event = false;
  for (const Op& op : cigar) {
    const int a = ref_at, r0 = read_at;
    const bool on_ref = op.code == C_MATCH || op.code == C_DEL || op.code == C_REF_SKIP || op.code == C_PAD;
    const bool on_read = op.code == C_MATCH || op.code == C_INS || op.code == C_SOFT_CLIP;
    if (on_ref) ref_at += op.len;
    if (on_read) read_at += op.len;
    const int z = ref_at;                                         // the operation covers contig offsets [a, z)
    if (op.code == C_MATCH)
      for (int r = r0; r < read_at; ++r) {
        if (r < 0 || (size_t)r >= mism.size() || !mism[r]) continue;
        const int o = a + (r - r0);
        if (insertion ? (o >= ev.at && o < ev_end) : std::abs(ev.at - o) < min_anchor_len) ++spl_mismatches;
      }
    if (a >= ev_end || z <= ev.at) {                              // untouched by the event
      if (a == ev_end && !insertion && op.code != C_INS) { behind_event = true; b.put(ev.code, ev.len); }
      const bool lower = behind_event ? lower_after : lower_before;
      b.put(lower ? lower_case(op.code) : op.code, op.len);
      continue;
    }
    behind_event = true;
    if (insertion) {
      b.put(op.code, ev.at - a);
      b.put(C_INS, ev.at < 0 ? ev.len + ev.at : ev.len);          // a hit starting inside the insertion sees only its tail
      b.put(op.code, z - ev_end);
    } else {
      b.put(lower_before ? lower_case(op.code) : op.code, ev.at - a);
      b.put(ev.code, ev.len);
      b.put(lower_after ? lower_case(op.code) : op.code, z - ev.at);
    }
  }
  if (spl.size() < cigar.size() + 2) return false;
  const int f = spl.front().code, l = spl.back().code;
  return (f == C_MATCH || f == C_mATCH) && (l == C_MATCH || l == C_mATCH);
}

// one juncs_db contig: where its event lies on the genome
struct SplTarget { bool ok = false, ins = false; uint32_t ref_id = 0, ref_id2 = 0; int left_edge = 0, splice_left = 0, splice_right = 0; std::string inserted;
                   int opcode = C_REF_SKIP; bool rev = false; };

void split(const std::string& s, char sep, std::vector<std::string>& out, bool strict)
{
  out.clear(); std::string cur;
  for (char c : s) { if (c == sep) { if (strict || !cur.empty()) out.push_back(cur); cur.clear(); } else cur.push_back(c); }
  if (strict || !cur.empty()) out.push_back(cur);
}

}  // namespace

JoinHitStream::JoinHitStream(const std::string& path, RefTable& rt, std::mutex& rt_mutex, bool spliced, int max_report_intron, int min_anchor_len,
                             StreamRange range)
  : path_(path), rt_(rt), rt_mutex_(rt_mutex), spliced_(spliced), max_report_intron_(max_report_intron), min_anchor_len_(min_anchor_len), range_(range), q_(4)
{
  th_ = std::thread([this] { produce(); });
}
JoinHitStream::~JoinHitStream() { q_.stop(); if (th_.joinable()) th_.join(); }

void JoinHitStream::produce()
{
  BamReader br;
  if (!br.open_shared(path_, range_.voffset)) { err_ = br.error(); q_.finish(); return; }
  const auto& tnames = br.header().target_name;
  std::vector<uint32_t> tid2ref(tnames.size(), 0);
  if (!spliced_) { std::lock_guard<std::mutex> l(rt_mutex_); for (size_t i = 0; i < tid2ref.size(); ++i) tid2ref[i] = rt_.get_id(tnames[i]); }
  // spliced streams: contig names "ref|left_start|L-R|right_end|type|strand" are parsed once per file and process (the table is
  // shared by the read-id ranges of the file: a junction index has a contig per junction)
  std::shared_ptr<const std::vector<SplTarget>> spl_tab;
  if (spliced_) {
    static std::mutex cm; static std::map<std::string, std::shared_ptr<const std::vector<SplTarget>>> cache;
    std::lock_guard<std::mutex> cl(cm);                            // one builder at a time; later ranges find the table ready
    auto& slot = cache[path_];
    if (!slot) {
      auto tab = std::make_shared<std::vector<SplTarget>>(tnames.size());
      std::vector<std::string> toks, st;
      for (size_t i = 0; i < tnames.size(); ++i) {
        split(tnames[i], '|', toks, true);
        const int extra = (int)toks.size() - 6;
        if (extra < 0) continue;                                    // malformed splice record -> every hit on it is skipped
        std::string contig = toks[0];
        for (int t = 1; t <= extra; ++t) { contig += "|"; contig += toks[t]; }
        split(toks[extra + 2], '-', st, false);
        if (st.size() != 2) continue;
        const std::string& jtype = toks[extra + 4]; const std::string& jstrand = toks[extra + 5];
        SplTarget T; T.left_edge = atoi(toks[extra + 1].c_str()); T.splice_left = atoi(st[0].c_str());
        if (jtype == "ins") { T.ins = true; T.inserted = st[1]; T.rev = jstrand == "rev"; }
        else {
          if (!(jstrand == "rev" || jstrand == "fwd" || jstrand == "ff" || jstrand == "fr" || jstrand == "rf" || jstrand == "rr")) continue;
          T.splice_right = atoi(st[1].c_str()); T.rev = jstrand == "rev";
          if (jtype == "fus") {
            // a fusion contig "ref1-ref2|...|L-R|...|fus|ff": the part behind the fusion point lies on ref2 (1690-1704, 1729-1742)
            T.opcode = jstrand == "ff" ? C_FUSION_FF : jstrand == "fr" ? C_FUSION_FR : jstrand == "rf" ? C_FUSION_RF : C_FUSION_RR;
            std::vector<std::string> two; split(contig, '-', two, false);
            if (two.size() != 2) continue;
            { std::lock_guard<std::mutex> l(rt_mutex_); T.ref_id = rt_.get_id(two[0]); T.ref_id2 = rt_.get_id(two[1]); }
            T.ok = true; (*tab)[i] = T;
            continue;
          }
          T.opcode = jtype == "del" ? C_DEL : C_REF_SKIP;
        }
        { std::lock_guard<std::mutex> l(rt_mutex_); T.ref_id = rt_.get_id(contig); T.ref_id2 = T.ref_id; }
        T.ok = true; (*tab)[i] = T;
      }
      slot = tab;
    }
    spl_tab = slot;
  }
  static const std::vector<SplTarget> no_targets;
  const std::vector<SplTarget>& spl = spl_tab ? *spl_tab : no_targets;
  uint32_t star_id = 0;
  const size_t CH = 1 << 15;
  std::vector<JHitRec> chunk; chunk.reserve(CH);
  std::vector<Op> cig, scig; std::vector<bool> mism;
  BamRecord r;
  while (br.next(r)) {
    bool end = true;
    const char* pipe = strrchr(r.qname, '|');
    if (pipe && strchr(pipe + 1, ':')) end = last_segment_from_suffix(pipe + 1);
    JHitRec hr; memset(&hr, 0, sizeof hr);
    hr.id = (uint32_t)atoi(r.qname);
    if (hr.id < range_.begin_id) continue;        // id range of this stream (files are id-sorted)
    if (hr.id >= range_.end_id) break;
    if (r.tid < 0) {
      if (!star_id) { std::lock_guard<std::mutex> l(rt_mutex_); star_id = rt_.get_id("*"); }
      hr.h.ref_id = star_id; hr.h.n_ops = 1; hr.h.ops[0] = pack_op(Op{C_MATCH, 0}); hr.h.flags = end ? THB_HIT_END : 0;
      chunk.push_back(hr);
      if (chunk.size() >= CH) { q_.push(std::move(chunk)); chunk = std::vector<JHitRec>(); chunk.reserve(CH); }
      continue;
    }
    cig.clear(); bool bad = false, spliced_aln = false;
    for (int i = 0; i < r.n_cigar && !bad; ++i) {
      const uint32_t c = r.cigar_at(i); const int len = (int)(c >> 4); const int op = (int)(c & 15);
      if (len <= 0) { bad = true; break; }
      switch (op) {
        case 0: cig.push_back(Op{C_MATCH, len}); break;
        case 1: cig.push_back(Op{C_INS, len}); break;
        case 2: cig.push_back(Op{C_DEL, len}); break;
        case 3: if (len > max_report_intron_) bad = true; spliced_aln = true; cig.push_back(Op{C_REF_SKIP, len}); break;
        case 4: cig.push_back(Op{C_SOFT_CLIP, len}); break;
        case 5: break;                                            // hard clips are dropped
        case 6: cig.push_back(Op{C_PAD, len}); break;
        default: bad = true;
      }
    }
    if (bad) continue;
    if (r.mtid >= 0 && r.mtid != r.tid) continue;
    const bool anti = (r.flag & 0x10) != 0;
    if (!spliced_) {
      if (r.aux_str("XF")) continue;                              // fusion-pass records (not produced without --fusion-search)
      int64_t nm = 0; r.aux_int("NM", &nm);
      unsigned char m = (unsigned char)nm;
      for (const Op& o : cig) if (o.code == C_INS || o.code == C_DEL) m -= (unsigned char)o.len;
      bool asplice = false;
      if (spliced_aln) { const char xs = r.aux_char("XS"); asplice = xs == '-'; }
      if (cig.size() > THB_JHIT_MAX_OPS) { ++dropped_; continue; }
      hr.h.ref_id = (size_t)r.tid < tid2ref.size() ? tid2ref[r.tid] : 0; hr.h.left = r.pos;
      hr.h.n_ops = (uint8_t)cig.size(); for (size_t k = 0; k < cig.size(); ++k) hr.h.ops[k] = pack_op(cig[k]);
      hr.h.flags = (uint8_t)((anti ? THB_HIT_ANTISENSE : 0) | (end ? THB_HIT_END : 0) | (asplice ? THB_JHIT_ANTISENSE_SPLICE : 0));
      hr.h.mismatches = m; hr.h.splice_mms = 0;
    } else {
      if ((size_t)r.tid >= spl.size() || !spl[r.tid].ok) continue;
      const SplTarget& T = spl[r.tid];
      // getBAMmismatches (bwt_map.cpp:410-475): mismatch positions from MD
      mism.assign((size_t)std::max(0, r.l_seq), false); int num_mm = 0;
      if (const char* p = r.aux_str("MD")) {
        int bi = 0;
        while (*p) {
          if (isdigit((unsigned char)*p)) { const int v = atoi(p); do { ++p; } while (isdigit((unsigned char)*p)); bi += v; }
          while (isalpha((unsigned char)*p)) { ++p; ++num_mm; if (bi >= 0 && (size_t)bi < mism.size()) mism[bi] = true; ++bi; }
          if (*p == '^') { ++p; while (isalpha((unsigned char)*p)) { ++p; ++bi; } }
          if (*p && !isdigit((unsigned char)*p) && !isalpha((unsigned char)*p) && *p != '^') ++p;
        }
      }
      scig.clear(); int spl_mm = 0; int left; bool anti_out = anti;
      if (T.ins) {
        left = T.left_edge + r.pos;
        if (left > T.splice_left) continue;
        if (!splice_cigar(scig, cig, mism, left, T.splice_left + 1, (int)T.inserted.size(), C_INS, spl_mm, min_anchor_len_)) continue;
        if (spl_mm < 0) continue;
        num_mm -= spl_mm; spl_mm = 0;                             // create_hit(..., splice_mms = 0) for insertions (1652-1664)
      } else {
        // "del", "intron" or "fusion" (1668-1755)
        const bool fusion = is_fusion_code(T.opcode), leftwards = T.opcode == C_FUSION_RF || T.opcode == C_FUSION_RR;
        left = leftwards ? T.left_edge - r.pos : T.left_edge + r.pos;
        int left_splice_pos = T.splice_left; int gap_len;
        if (fusion) gap_len = T.splice_right; else gap_len = T.splice_right - T.splice_left - 1;
        if (leftwards) { left_splice_pos -= 1; if (left <= left_splice_pos) continue; }
        else { left_splice_pos += 1; if (left >= left_splice_pos) continue; }
        if (!splice_cigar(scig, cig, mism, left, left_splice_pos, gap_len, T.opcode, spl_mm, min_anchor_len_)) continue;
        if (spl_mm < 0) continue;
        if (leftwards) anti_out = !anti;                              // 1740-1741
      }
      const bool fused = is_fusion_code(T.opcode) && !T.ins;
      if (scig.size() > (size_t)(fused ? THB_JHIT_MAX_OPS - 1 : THB_JHIT_MAX_OPS)) { ++dropped_; continue; }
      hr.h.ref_id = T.ref_id; hr.h.left = left;
      hr.h.n_ops = (uint8_t)scig.size(); for (size_t k = 0; k < scig.size(); ++k) hr.h.ops[k] = pack_op(scig[k]);
      if (fused) hr.h.ops[THB_JHIT_MAX_OPS - 1] = T.ref_id2;          // second contig (include/tophat_b200.h)
      hr.h.flags = (uint8_t)((anti_out ? THB_HIT_ANTISENSE : 0) | (end ? THB_HIT_END : 0) | (T.rev ? THB_JHIT_ANTISENSE_SPLICE : 0) |
                             (anti_out != anti ? THB_JHIT_SEQ_FLIPPED : 0));
      hr.h.mismatches = (uint8_t)num_mm; hr.h.splice_mms = (uint8_t)spl_mm;
    }
    chunk.push_back(hr);
    if (chunk.size() >= CH) { q_.push(std::move(chunk)); chunk = std::vector<JHitRec>(); chunk.reserve(CH); }
  }
  if (!br.error().empty()) err_ = br.error();
  if (!chunk.empty()) q_.push(std::move(chunk));
  q_.finish();
}

bool JoinHitStream::ensure()
{
  while (pos_ >= cur_.size()) {
    if (end_) return false;
    cur_.clear(); pos_ = 0;
    if (!q_.pop(cur_)) { end_ = true; return false; }
  }
  return true;
}
uint32_t JoinHitStream::next_group_id() { while (ensure()) { if (cur_[pos_].id != 0) return cur_[pos_].id; ++pos_; } return 0; }
void JoinHitStream::next_group(std::vector<thb_jhit_full>& out)
{ const uint32_t id = next_group_id(); if (!id) return; while (ensure() && cur_[pos_].id == id) { out.push_back(cur_[pos_].h); ++pos_; } }
void JoinHitStream::skip_group()
{ const uint32_t id = next_group_id(); if (!id) return; while (ensure() && cur_[pos_].id == id) ++pos_; }

// ---- FullReadStream ---------------------------------------------------------------------------------------
FullReadStream::FullReadStream(const std::string& path, StreamRange range) : path_(path), range_(range), q_(4)
{
  // a reads file that is not BAM is FASTQ / FASTA text, possibly gzip-ed (ReadStream::init, reads.cpp:528-560; next_fastx_read 94-190)
  const bool bam = path.size() >= 4 && path.compare(path.size() - 4, 4, ".bam") == 0;
  th_ = std::thread([this, bam] { if (bam) produce(); else produce_fastx(); });
}
FullReadStream::~FullReadStream() { q_.stop(); if (th_.joinable()) th_.join(); }

void FullReadStream::produce()
{
  BamReader br;
  if (!br.open_shared(path_, range_.voffset)) { err_ = br.error(); q_.finish(); return; }
  static const char* nt16 = "=ACMGRSVTWYHKDBN";
  const size_t CH = 1 << 14;
  std::vector<FullRead> chunk; chunk.reserve(CH);
  BamRecord r;
  while (br.next(r)) {
    if (r.flag & 0x200) continue;
    const uint32_t rid = (uint32_t)atol(r.qname);
    if (rid < range_.begin_id) continue;
    if (rid >= range_.end_id) break;
    FullRead fr; fr.id = rid; fr.name = r.qname;
    fr.seq.resize((size_t)r.l_seq); fr.qual.resize((size_t)r.l_seq);
    for (int k = 0; 2 * k < r.l_seq; ++k) { const uint8_t b = r.seq[k]; fr.seq[2 * k] = nt16[b >> 4]; if (2 * k + 1 < r.l_seq) fr.seq[2 * k + 1] = nt16[b & 15]; }
    for (int i = 0; i < r.l_seq; ++i) fr.qual[i] = (char)(r.qual[i] + 33);
    chunk.push_back(std::move(fr));
    if (chunk.size() >= CH) { q_.push(std::move(chunk)); chunk = std::vector<FullRead>(); chunk.reserve(CH); }
  }
  if (!br.error().empty()) err_ = br.error();
  if (!chunk.empty()) q_.push(std::move(chunk));
  q_.finish();
}
void FullReadStream::produce_fastx()
{
  gzFile f = gzopen(path_.c_str(), "rb");
  if (!f) { err_ = "cannot open " + path_ + " for reading"; q_.finish(); return; }
  gzbuffer(f, 1 << 20);
  const size_t CH = 1 << 14;
  std::vector<FullRead> chunk; chunk.reserve(CH);
  std::vector<char> line(1 << 16);
  auto getl = [&](std::string& s) -> bool { s.clear(); if (!gzgets(f, line.data(), (int)line.size())) return false; s = line.data();
    while (!s.empty() && (s.back() == '\n' || s.back() == '\r')) s.pop_back(); return true; };
  std::string l, seq, plus, qual;
  while (getl(l)) {
    if (l.empty()) continue;
    const bool fq = l[0] == '@';
    if (!fq && l[0] != '>') continue;
    if (!getl(seq)) break;
    if (fq) { getl(plus); getl(qual); }
    FullRead fr; fr.name = l.substr(1);
    const size_t sp = fr.name.find_first_of(" \t");          // the name ends at the first blank (reads.cpp:114-118)
    if (sp != std::string::npos) fr.name.resize(sp);
    fr.id = (uint32_t)atol(fr.name.c_str());
    if (fr.id < range_.begin_id) continue;                    // text read files have no index: a range scans from the start
    if (fr.id >= range_.end_id) break;
    for (char& c : seq) if (c == '.') c = 'N';                // reads.cpp:125
    fr.seq = seq; fr.qual = fq ? qual : std::string();
    chunk.push_back(std::move(fr));
    if (chunk.size() >= CH) { q_.push(std::move(chunk)); chunk = std::vector<FullRead>(); chunk.reserve(CH); }
  }
  gzclose(f);
  if (!chunk.empty()) q_.push(std::move(chunk));
  q_.finish();
}
bool FullReadStream::ensure()
{
  while (pos_ >= cur_.size()) {
    if (end_) return false;
    cur_.clear(); pos_ = 0;
    if (!q_.pop(cur_)) { end_ = true; return false; }
  }
  return true;
}
const FullRead* FullReadStream::next() { if (!ensure()) return nullptr; return &cur_[pos_++]; }
const FullRead* FullReadStream::get(uint32_t id)
{
  while (ensure()) { const FullRead& r = cur_[pos_]; if (r.id == id) return &r; if (r.id > id) return nullptr; ++pos_; }
  return nullptr;
}

}  // namespace thbhost
