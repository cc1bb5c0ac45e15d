#include "thb_join_input.hpp"
#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace thbhost {

namespace {

// CigarOpCode values of the reference (bwt_map.h:36-55)
enum { C_MATCH = 1, C_INS = 3, C_DEL = 5, C_REF_SKIP = 11, C_SOFT_CLIP = 13, C_HARD_CLIP = 14, C_PAD = 15 };
struct Op { int code; int len; };

inline uint32_t pack_op(const Op& o) { return ((uint32_t)o.len << 4) | (uint32_t)o.code; }

// cigar_add (bwt_map.cpp:670-676): extends an equal trailing op AND still appends the op (sic)
void cigar_add(std::vector<Op>& c, const Op& op)
{
  if (op.len <= 0) return;
  if (!c.empty() && c.back().code == op.code) c.back().len += op.len;
  c.push_back(op);
}

// spliceCigar (bwt_map.cpp:678-883) for INS / DEL / REF_SKIP events
bool splice_cigar(std::vector<Op>& spl, const std::vector<Op>& cigar, const std::vector<bool>& mism, int& left, int spl_start,
                  int spl_len, int spl_code, int& spl_mismatches, int min_anchor_len)
{
  const int spl_ofs = spl_start - left;
  int spl_ofs_end = spl_ofs;
  const Op gapop{spl_code, spl_len};
  if (spl_code == C_INS) spl_ofs_end += spl_len;
  int ref_ofs = 0, read_ofs = 0; bool xfound = false;
  if (spl_ofs_end > 0) {
    for (size_t c = 0; c < cigar.size(); ++c) {
      const int prev_read_ofs = read_ofs, cur_op_ofs = ref_ofs, cur_opcode = cigar[c].code, cur_oplen = cigar[c].len;
      switch (cur_opcode) {
        case C_MATCH:
          ref_ofs += cur_oplen; read_ofs += cur_oplen;
          if (spl_code == C_REF_SKIP || spl_code == C_DEL) {
            for (int o = cur_op_ofs; o < ref_ofs; ++o) { const int rofs = prev_read_ofs + (o - cur_op_ofs);
              if (std::abs(spl_ofs - o) < min_anchor_len && rofs >= 0 && (size_t)rofs < mism.size() && mism[rofs]) spl_mismatches++; }
          } else if (spl_code == C_INS) {
            for (int o = cur_op_ofs; o < ref_ofs; ++o) { const int rofs = prev_read_ofs + (o - cur_op_ofs);
              if (o >= spl_ofs && o < spl_ofs_end && rofs >= 0 && (size_t)rofs < mism.size() && mism[rofs]) spl_mismatches++; }
          }
          break;
        case C_DEL: case C_REF_SKIP: case C_PAD: ref_ofs += cur_oplen; break;
        case C_SOFT_CLIP: case C_INS: read_ofs += cur_oplen; break;
      }
      if (cur_op_ofs >= spl_ofs_end || ref_ofs <= spl_ofs) {
        if (cur_op_ofs == spl_ofs_end && spl_code != C_INS && cur_opcode != C_INS) { xfound = true; cigar_add(spl, gapop); }
        cigar_add(spl, cigar[c]);
      } else {
        xfound = true;
        if (spl_code == C_INS) {
          Op op = cigar[c]; op.len = spl_ofs - cur_op_ofs;
          if (spl_ofs > cur_op_ofs) cigar_add(spl, op);
          if (spl_ofs < 0) { Op t = gapop; t.len += spl_ofs; if (t.len > 0) cigar_add(spl, t); }
          else cigar_add(spl, gapop);
          op.len = ref_ofs - spl_ofs_end;
          if (ref_ofs > spl_ofs_end) cigar_add(spl, op);
        } else {
          Op op = cigar[c]; op.len = spl_ofs - cur_op_ofs; cigar_add(spl, op);
          cigar_add(spl, gapop);
          op.len = ref_ofs - spl_ofs; cigar_add(spl, op);
        }
      }
    }
  }
  (void)xfound;
  if (spl_ofs_end <= 0) { if (spl_code == C_INS) left -= spl_len; else left += spl_len; spl = cigar; }
  if (spl.size() < cigar.size() + 2) return false;
  if (spl.front().code != C_MATCH) return false;
  if (spl.back().code != C_MATCH) return false;
  return true;
}

void split(const std::string& s, char sep, std::vector<std::string>& out, bool strict)
{
  out.clear(); std::string cur;
  for (char c : s) { if (c == sep) { if (strict || !cur.empty()) out.push_back(cur); cur.clear(); } else cur.push_back(c); }
  if (strict || !cur.empty()) out.push_back(cur);
}

}  // namespace

JoinHitStream::JoinHitStream(const std::string& path, RefTable& rt, std::mutex& rt_mutex, bool spliced, int max_report_intron, int min_anchor_len)
  : path_(path), rt_(rt), rt_mutex_(rt_mutex), spliced_(spliced), max_report_intron_(max_report_intron), min_anchor_len_(min_anchor_len), q_(4)
{
  th_ = std::thread([this] { produce(); });
}
JoinHitStream::~JoinHitStream() { q_.stop(); if (th_.joinable()) th_.join(); }

void JoinHitStream::produce()
{
  BamReader br;
  if (!br.open(path_)) { err_ = br.error(); q_.finish(); return; }
  const auto& tnames = br.header().target_name;
  std::vector<uint32_t> tid2ref(tnames.size(), 0);
  if (!spliced_) { std::lock_guard<std::mutex> l(rt_mutex_); for (size_t i = 0; i < tid2ref.size(); ++i) tid2ref[i] = rt_.get_id(tnames[i]); }
  // spliced streams: contig names "ref|left_start|L-R|right_end|type|strand" are parsed once per target
  struct SplTarget { bool ok = false, ins = false; uint32_t ref_id = 0; int left_edge = 0, splice_left = 0, splice_right = 0; std::string inserted;
                     int opcode = C_REF_SKIP; bool rev = false; };
  std::vector<SplTarget> spl(spliced_ ? tnames.size() : 0);
  if (spliced_) {
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
        if (jtype == "fus") continue;                             // fusion contigs only exist with --fusion-search (unsupported)
        if (!(jstrand == "rev" || jstrand == "fwd" || jstrand == "ff" || jstrand == "fr" || jstrand == "rf" || jstrand == "rr")) continue;
        T.opcode = jtype == "del" ? C_DEL : C_REF_SKIP; T.splice_right = atoi(st[1].c_str()); T.rev = jstrand == "rev";
      }
      { std::lock_guard<std::mutex> l(rt_mutex_); T.ref_id = rt_.get_id(contig); }
      T.ok = true; spl[i] = T;
    }
  }
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
      scig.clear(); int spl_mm = 0; int left;
      if (T.ins) {
        left = T.left_edge + r.pos;
        if (left > T.splice_left) continue;
        if (!splice_cigar(scig, cig, mism, left, T.splice_left + 1, (int)T.inserted.size(), C_INS, spl_mm, min_anchor_len_)) continue;
        if (spl_mm < 0) continue;
        num_mm -= spl_mm; spl_mm = 0;                             // create_hit(..., splice_mms = 0) for insertions (1652-1664)
      } else {
        left = T.left_edge + r.pos;
        const int left_splice_pos = T.splice_left + 1;
        const int gap_len = T.splice_right - T.splice_left - 1;
        if (left >= left_splice_pos) continue;
        if (!splice_cigar(scig, cig, mism, left, left_splice_pos, gap_len, T.opcode, spl_mm, min_anchor_len_)) continue;
        if (spl_mm < 0) continue;
      }
      if (scig.size() > THB_JHIT_MAX_OPS) { ++dropped_; continue; }
      hr.h.ref_id = T.ref_id; hr.h.left = left;
      hr.h.n_ops = (uint8_t)scig.size(); for (size_t k = 0; k < scig.size(); ++k) hr.h.ops[k] = pack_op(scig[k]);
      hr.h.flags = (uint8_t)((anti ? THB_HIT_ANTISENSE : 0) | (end ? THB_HIT_END : 0) | (T.rev ? THB_JHIT_ANTISENSE_SPLICE : 0));
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
FullReadStream::FullReadStream(const std::string& path) : path_(path), q_(4) { th_ = std::thread([this] { produce(); }); }
FullReadStream::~FullReadStream() { q_.stop(); if (th_.joinable()) th_.join(); }

void FullReadStream::produce()
{
  BamReader br;
  if (!br.open(path_)) { err_ = br.error(); q_.finish(); return; }
  static const char* nt16 = "=ACMGRSVTWYHKDBN";
  const size_t CH = 1 << 14;
  std::vector<FullRead> chunk; chunk.reserve(CH);
  BamRecord r;
  while (br.next(r)) {
    if (r.flag & 0x200) continue;
    FullRead fr; fr.id = (uint32_t)atol(r.qname); fr.name = r.qname;
    fr.seq.resize((size_t)r.l_seq); fr.qual.resize((size_t)r.l_seq);
    for (int i = 0; i < r.l_seq; ++i) { fr.seq[i] = nt16[(r.seq[i >> 1] >> ((~i & 1) << 2)) & 15]; fr.qual[i] = (char)(r.qual[i] + 33); }
    chunk.push_back(std::move(fr));
    if (chunk.size() >= CH) { q_.push(std::move(chunk)); chunk = std::vector<FullRead>(); chunk.reserve(CH); }
  }
  if (!br.error().empty()) err_ = br.error();
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
const FullRead* FullReadStream::get(uint32_t id)
{
  while (ensure()) { const FullRead& r = cur_[pos_]; if (r.id == id) return &r; if (r.id > id) return nullptr; ++pos_; }
  return nullptr;
}

}  // namespace thbhost
