#include "thb_input.hpp"
#include <sys/stat.h>
#include <unistd.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <atomic>
#include <zlib.h>
#include <time.h>

namespace thbhost {

static std::atomic<long long> g_producer_cpu_us(0);
double thread_cpu_seconds() { timespec ts; clock_gettime(CLOCK_THREAD_CPUTIME_ID, &ts); return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec; }
void add_producer_cpu() { g_producer_cpu_us += (long long)(thread_cpu_seconds() * 1e6); }
double producer_cpu_seconds() { return 1e-6 * (double)g_producer_cpu_us.load(); }

// ---- RefTable -----------------------------------------------------------------------------------------
uint32_t RefTable::get_id(const std::string& name)
{
  auto it = by_name_.find(name);
  if (it != by_name_.end()) return it->second;
  names_.push_back(name);
  const uint32_t id = (uint32_t)names_.size();
  by_name_.emplace(name, id);
  return id;
}
uint32_t RefTable::find(const std::string& name) const { auto it = by_name_.find(name); return it == by_name_.end() ? 0 : it->second; }

bool RefTable::load_sam_header(const std::string& path, std::string* err)
{
  std::ifstream in(path.c_str());
  if (!in.good()) { *err = "Failed to open SAM header file " + path; return false; }
  std::string line;
  while (std::getline(in, line)) {
    if (line.compare(0, 3, "@SQ") != 0) continue;
    size_t p = 0;
    while ((p = line.find('\t', p)) != std::string::npos) {
      ++p;
      if (line.compare(p, 3, "SN:") == 0) {
        size_t e = line.find_first_of("\t\r\n", p);
        get_id(line.substr(p + 3, e == std::string::npos ? std::string::npos : e - p - 3));
        break;
      }
    }
  }
  return true;
}

// ---- Genome ---------------------------------------------------------------------------------------------
thb_ref_image Genome::image() const
{
  thb_ref_image img; img.n_contigs = (uint32_t)contig_len.size(); img.contig_start = contig_start.data();
  img.contig_len = contig_len.data(); img.n_blocks = n_blocks; img.planes = planes.data(); img.nmask = nmask.data();
  return img;
}

// ---- packed genome image cached next to the FASTA ---------------------------------------------------------------------
// Every stage process loads the same FASTA (three processes per sample; the reference spends ~12 ms per Mbp on it, 37 s for an
// hg38-sized genome, each time).  The first process to parse a FASTA of >= 16 MB leaves <fasta>.thb_img: the contigs in file
// order (name, length) + the bit planes; later processes check size / mtime of the FASTA, replay the name -> id assignments and
// read the planes.  TOPHAT_GPU_NO_IMAGE_CACHE=1 turns it off; an unwritable directory just means no cache.
namespace {
struct ImgHeader { char magic[8]; uint64_t fasta_size, fasta_mtime_ns, n_blocks; uint32_t n_fasta_contigs, reserved; };
const char IMG_MAGIC[8] = {'T', 'H', 'B', 'I', 'M', 'G', '2', 0};

bool fasta_stamp(const std::string& path, uint64_t* size, uint64_t* mtime_ns)
{
  struct stat st; if (stat(path.c_str(), &st) != 0) return false;
  *size = (uint64_t)st.st_size; *mtime_ns = (uint64_t)st.st_mtim.tv_sec * 1000000000ull + (uint64_t)st.st_mtim.tv_nsec; return true;
}

void layout_genome(const RefTable& rt, const std::map<uint32_t, uint64_t>& len_of, Genome& g)
{
  const uint32_t nc = rt.size();
  g.contig_start.assign(nc, 0); g.contig_len.assign(nc, 0);
  uint64_t gpos = 0;
  for (uint32_t id = 1; id <= nc; ++id) {
    auto it = len_of.find(id);
    const uint64_t len = it == len_of.end() ? 0 : it->second;
    g.contig_start[id - 1] = gpos; g.contig_len[id - 1] = (uint32_t)len;
    gpos += ((len + 63) / 64 + 1) * 64;
  }
  g.n_blocks = gpos / 64 + 1;
}

bool load_image_cache(const std::string& path, RefTable& rt, Genome& g, bool log)
{
  if (getenv("TOPHAT_GPU_NO_IMAGE_CACHE")) return false;
  uint64_t fsz = 0, fmt = 0; if (!fasta_stamp(path, &fsz, &fmt)) return false;
  FILE* f = fopen((path + ".thb_img").c_str(), "rb"); if (!f) return false;
  bool ok = false; ImgHeader h;
  std::vector<std::pair<std::string, uint64_t>> contigs;
  if (fread(&h, sizeof h, 1, f) == 1 && memcmp(h.magic, IMG_MAGIC, 8) == 0 && h.fasta_size == fsz && h.fasta_mtime_ns == fmt && h.n_fasta_contigs < (1u << 26)) {
    ok = true;
    for (uint32_t i = 0; ok && i < h.n_fasta_contigs; ++i) {
      uint32_t nl = 0; uint64_t len = 0;
      if (fread(&nl, 4, 1, f) != 1 || nl > 65535) { ok = false; break; }
      std::string nm(nl, '\0');
      if ((nl && fread(&nm[0], 1, nl, f) != nl) || fread(&len, 8, 1, f) != 1) { ok = false; break; }
      contigs.emplace_back(nm, len);
    }
  }
  if (ok) {
    // the same get_id sequence the parse performs (a repeated name keeps the later record, as std::map::operator[] + swap does)
    RefTable trial = rt; std::map<uint32_t, uint64_t> len_of;
    for (auto& c : contigs) len_of[trial.get_id(c.first)] = c.second;
    Genome t; layout_genome(trial, len_of, t);
    if (t.n_blocks != h.n_blocks) ok = false;
    else {
      t.planes.resize(2 * t.n_blocks); t.nmask.resize(t.n_blocks);
      ok = fread(t.planes.data(), 8, 2 * t.n_blocks, f) == 2 * t.n_blocks && fread(t.nmask.data(), 8, t.n_blocks, f) == t.n_blocks;
      if (ok) {
        rt = trial; g = std::move(t);
        if (log) for (auto& c : contigs) { if (!c.first.empty()) fprintf(stderr, "\tLoading %s...", c.first.c_str()); if (c.second) fprintf(stderr, " done (%ld bases).\n", (long)c.second); }
      }
    }
  }
  fclose(f);
  return ok;
}

void save_image_cache(const std::string& path, const std::vector<std::pair<std::string, uint64_t>>& contigs, const Genome& g)
{
  if (getenv("TOPHAT_GPU_NO_IMAGE_CACHE")) return;
  ImgHeader h; memset(&h, 0, sizeof h); memcpy(h.magic, IMG_MAGIC, 8);
  if (!fasta_stamp(path, &h.fasta_size, &h.fasta_mtime_ns) || h.fasta_size < (16u << 20)) return;
  h.n_blocks = g.n_blocks; h.n_fasta_contigs = (uint32_t)contigs.size();
  const std::string tmp = path + ".thb_img.tmp" + std::to_string((long)getpid());
  FILE* f = fopen(tmp.c_str(), "wb"); if (!f) return;
  bool ok = fwrite(&h, sizeof h, 1, f) == 1;
  for (auto& c : contigs) { const uint32_t nl = (uint32_t)c.first.size(); ok = ok && fwrite(&nl, 4, 1, f) == 1 && (!nl || fwrite(c.first.data(), 1, nl, f) == nl) && fwrite(&c.second, 8, 1, f) == 1; }
  ok = ok && fwrite(g.planes.data(), 8, g.planes.size(), f) == g.planes.size() && fwrite(g.nmask.data(), 8, g.nmask.size(), f) == g.nmask.size();
  ok = (fclose(f) == 0) && ok;
  if (!ok || rename(tmp.c_str(), (path + ".thb_img").c_str()) != 0) remove(tmp.c_str());
}
}  // namespace

bool load_fasta(const std::string& path, RefTable& rt, Genome& g, bool log, int threads, std::string* err)
{
  if (load_image_cache(path, rt, g, log)) return true;
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) { *err = "cannot open " + path + " for reading"; return false; }
  std::vector<std::pair<std::string, uint64_t>> fasta_order;
  std::map<uint32_t, std::string> seqs;          // id -> bases (no line breaks)
  std::vector<char> buf(1 << 22);
  std::string name, seq, header; bool in_header = false, have_record = false, at_line_start = true;
  auto flush = [&]() {
    if (!have_record) return;
    if (log && !name.empty()) fprintf(stderr, "\tLoading %s...", name.c_str());
    if (log && !seq.empty()) fprintf(stderr, " done (%ld bases).\n", (long)seq.size());
    const uint32_t id = rt.get_id(name);
    fasta_order.emplace_back(name, (uint64_t)seq.size());
    seqs[id].swap(seq); seq.clear();
  };
  size_t n;
  while ((n = fread(buf.data(), 1, buf.size(), f)) > 0) {
    size_t i = 0;
    while (i < n) {
      if (in_header) {
        const char* nl = (const char*)memchr(buf.data() + i, '\n', n - i);
        const size_t e = nl ? (size_t)(nl - buf.data()) : n;
        header.append(buf.data() + i, e - i);
        i = e;
        if (nl) { in_header = false; at_line_start = true; ++i;
          size_t cut = header.find_first_of(" \t\r"); name = cut == std::string::npos ? header : header.substr(0, cut); header.clear(); }
        continue;
      }
      if (at_line_start && buf[i] == '>') { flush(); have_record = true; in_header = true; ++i; continue; }
      if (!have_record) { have_record = true; name.clear(); }     // sequence before any header: empty name
      const char* nl = (const char*)memchr(buf.data() + i, '\n', n - i);
      const size_t e = nl ? (size_t)(nl - buf.data()) : n;
      size_t e2 = e; if (e2 > i && buf[e2 - 1] == '\r') --e2;
      // strip any embedded CR (rare); everything else is a base
      for (size_t k = i; k < e2; ++k) if (buf[k] == '\r') { seq.append(buf.data() + i, k - i); i = k + 1; }
      if (e2 > i) seq.append(buf.data() + i, e2 - i);
      i = e; at_line_start = false;
      if (nl) { at_line_start = true; ++i; }
    }
  }
  fclose(f);
  if (in_header) { size_t cut = header.find_first_of(" \t\r"); name = cut == std::string::npos ? header : header.substr(0, cut); }
  flush();
  { std::map<uint32_t, uint64_t> len_of;
    for (auto& kv : seqs) { if (kv.second.size() > 0xffffffffull) { *err = "contig longer than 2^32 bases"; return false; } len_of[kv.first] = kv.second.size(); }
    layout_genome(rt, len_of, g); }
  g.planes.assign(2 * g.n_blocks, 0); g.nmask.assign(g.n_blocks, 0);
  // pack 64-aligned slices in parallel (slices never share a plane word)
  struct Job { uint32_t id; uint64_t off, len; };
  std::vector<Job> jobs; const uint64_t SL = 1ull << 22;
  for (auto& kv : seqs) for (uint64_t o = 0; o < kv.second.size(); o += SL) jobs.push_back({kv.first, o, std::min<uint64_t>(SL, kv.second.size() - o)});
  std::atomic<size_t> nexti(0);
  auto work = [&]() { for (;;) { size_t j = nexti.fetch_add(1); if (j >= jobs.size()) break; const Job& jb = jobs[j];
      thb_pack_bases(seqs[jb.id].data() + jb.off, jb.len, g.contig_start[jb.id - 1] + jb.off, g.planes.data(), g.nmask.data()); } };
  std::vector<std::thread> th; const int nt = std::max(1, std::min<int>(threads, (int)jobs.size()));
  for (int t = 1; t < nt; ++t) th.emplace_back(work);
  work(); for (auto& t : th) t.join();
  save_image_cache(path, fasta_order, g);
  return true;
}

// ---- HitStream ------------------------------------------------------------------------------------------
bool BamIndex::load(const std::string& bam_path)
{
  entries.clear();
  FILE* f = fopen((bam_path + ".index").c_str(), "r");
  if (!f) return false;
  unsigned long long id = 0; long long off = 0;
  while (fscanf(f, "%llu %lld", &id, &off) == 2) if (off > 0) entries.emplace_back((uint32_t)id, (uint64_t)off);
  fclose(f);
  return !entries.empty();
}
uint64_t BamIndex::offset_for(uint32_t begin_id) const
{
  // entries are in increasing id order; upper_bound - 1 = last entry with id <= begin_id
  size_t lo = 0, hi = entries.size();
  while (lo < hi) { const size_t mid = (lo + hi) / 2; if (entries[mid].first <= begin_id) lo = mid + 1; else hi = mid; }
  return lo ? entries[lo - 1].second : 0;
}
std::vector<uint32_t> split_ids(const BamIndex& idx, int parts)
{
  std::vector<uint32_t> out;
  const size_t n = idx.entries.size();
  if (parts < 2 || n < 2) return out;
  const size_t np = std::min<size_t>((size_t)parts, n);
  for (size_t k = 1; k < np; ++k) {
    const uint32_t id = idx.entries[n * k / np].first;
    if (id > 0 && (out.empty() || id > out.back())) out.push_back(id);
  }
  return out;
}
StreamRange range_for(const std::string& bam_path, uint32_t begin_id, uint32_t end_id)
{
  StreamRange r; r.begin_id = begin_id; r.end_id = end_id;
  if (begin_id > 0) { BamIndex ix; if (ix.load(bam_path)) r.voffset = ix.offset_for(begin_id); }
  return r;
}

HitStream::HitStream(const std::string& path, RefTable& rt, std::mutex& rt_mutex, int max_report_intron, StreamRange range)
  : path_(path), rt_(rt), rt_mutex_(rt_mutex), max_report_intron_(max_report_intron), range_(range), q_(4)
{
  th_ = std::thread([this] { produce(); });
}
HitStream::~HitStream() { q_.stop(); if (th_.joinable()) th_.join(); }

void HitStream::produce()
{
  BamReader br;
  if (!br.open_shared(path_, range_.voffset)) { err_ = br.error(); q_.finish(); return; }
  std::vector<uint32_t> tid2ref(br.header().target_name.size(), 0);
  { std::lock_guard<std::mutex> l(rt_mutex_);
    for (size_t i = 0; i < tid2ref.size(); ++i) tid2ref[i] = rt_.get_id(br.header().target_name[i]); }
  uint32_t star_id = 0;
  const size_t CH = 1 << 16;
  std::vector<HitRec> chunk; chunk.reserve(CH);
  BamRecord r;
  while (br.next(r)) {
    // qname "<id>|<offset>:<seg>:<nsegs>" (bwt_map.cpp:1126-1143)
    bool end = true;
    const char* pipe = strrchr(r.qname, '|');
    if (pipe && strchr(pipe + 1, ':')) {
      end = last_segment_from_suffix(pipe + 1);
    }
    HitRec hr; memset(&hr, 0, sizeof hr);
    hr.id = (uint32_t)atoi(r.qname);          // atoi stops at '|' (ReadTable::get_id, bwt_map.h:546-552)
    if (hr.id < range_.begin_id) continue;    // id range of this stream (files are id-sorted: fix_map_ordering)
    if (hr.id >= range_.end_id) break;
    if (r.tid < 0) {                          // unmapped record -> hit on "*" (1145-1156)
      if (!star_id) { std::lock_guard<std::mutex> l(rt_mutex_); star_id = rt_.get_id("*"); }
      hr.h.ref_id = star_id; hr.h.flags = end ? THB_HIT_END : 0;
      chunk.push_back(hr);
    } else {
      if (r.aux_str("XF")) continue;          // fusion-pass records never reach this stage's inputs
      int64_t nm = 0; r.aux_int("NM", &nm);
      unsigned char mism = (unsigned char)nm; // the reference keeps it in an unsigned char (1182-1188)
      int64_t right = r.pos; unsigned rlen = 0, gaps = 0; bool bad = false;
      for (int i = 0; i < r.n_cigar && !bad; ++i) {
        const uint32_t c = r.cigar_at(i); const int len = (int)(c >> 4); const int op = (int)(c & 15);
        if (len <= 0) { bad = true; break; }
        switch (op) {
          case 0: right += len; rlen += len; break;                          // M
          case 1: rlen += len; gaps += len; mism -= (unsigned char)len; break; // I
          case 2: right += len; gaps += len; mism -= (unsigned char)len; break; // D
          case 3: if (len > max_report_intron_) bad = true; right += len; break; // N
          case 4: rlen += len; break;                                        // S
          case 5: case 6: break;                                             // H, P
          default: bad = true;                                               // invalid CIGAR operation
        }
      }
      if (bad) continue;
      if (r.mtid >= 0 && r.mtid != r.tid) continue;                          // 1409-1416
      hr.h.ref_id = (size_t)r.tid < tid2ref.size() ? tid2ref[r.tid] : 0;
      hr.h.left = r.pos; hr.h.right = (int32_t)right; hr.h.read_len = (uint8_t)std::min(rlen, 255u);
      hr.h.edit_dist = (uint8_t)(mism + gaps);
      hr.h.flags = (uint8_t)(((r.flag & 0x10) ? THB_HIT_ANTISENSE : 0) | (end ? THB_HIT_END : 0));
      chunk.push_back(hr);
    }
    if (chunk.size() >= CH) { q_.push(std::move(chunk)); chunk = std::vector<HitRec>(); chunk.reserve(CH); }
  }
  if (!br.error().empty()) err_ = br.error();
  if (!chunk.empty()) q_.push(std::move(chunk));
  add_producer_cpu();
  q_.finish();
}

bool HitStream::ensure()
{
  while (pos_ >= cur_.size()) {
    if (end_) return false;
    cur_.clear(); pos_ = 0;
    if (!q_.pop(cur_)) { end_ = true; return false; }
  }
  return true;
}

uint32_t HitStream::next_group_id()
{
  // records with id 0 terminate grouping in the reference (insert_id 0 == "no hit"); skip them
  while (ensure()) { if (cur_[pos_].id != 0) return cur_[pos_].id; ++pos_; }
  return 0;
}
void HitStream::next_group(std::vector<thb_hit>& out)
{
  const uint32_t id = next_group_id();
  if (!id) return;
  while (ensure() && cur_[pos_].id == id) { out.push_back(cur_[pos_].h); ++pos_; ++n_records_; }
}
void HitStream::skip_group()
{
  const uint32_t id = next_group_id();
  if (!id) return;
  while (ensure() && cur_[pos_].id == id) { ++pos_; ++n_records_; }
}

// ---- ReadStream -----------------------------------------------------------------------------------------
void pack_read_ascii(const char* s, uint32_t len, ReadRec& r)
{
  memset(r.planes, 0, sizeof r.planes);
  r.len = len;
  // bit0 / bit1 = 2-bit code, bit2 = "not ACGT" (code bits 0 there); branch-free: this runs once per read in the consumer loop
  static const struct Table { uint8_t v[256]; Table() { for (int i = 0; i < 256; ++i) v[i] = 4; v[(int)'A'] = 0; v[(int)'C'] = 1; v[(int)'G'] = 2; v[(int)'T'] = 3; } } T;
  for (uint32_t i = 0; i < len && i < 256; ++i) {
    const uint64_t c = T.v[(unsigned char)s[i]];
    const uint32_t w = i >> 6, j = i & 63;
    r.planes[w] |= (c & 1ull) << j;
    r.planes[4 + w] |= ((c >> 1) & 1ull) << j;
    r.planes[8 + w] |= (c >> 2) << j;
  }
}

ReadStream::ReadStream(const std::string& path, StreamRange range) : path_(path), range_(range), q_(4)
{
  const bool bam = path.size() >= 4 && path.compare(path.size() - 4, 4, ".bam") == 0;
  th_ = std::thread([this, bam] { if (bam) produce_bam(); else produce_fastx(); });
}
ReadStream::~ReadStream() { q_.stop(); if (th_.joinable()) th_.join(); }

namespace {
struct Nib2 { uint8_t p0[256], p1[256], pn[256]; };
Nib2 make_nib2(const int8_t (&code)[16])
{
  Nib2 t;
  for (int b = 0; b < 256; ++b) {
    uint8_t a0 = 0, a1 = 0, an = 0;
    for (int h = 0; h < 2; ++h) { const int c = code[h == 0 ? (b >> 4) : (b & 15)];
      if (c < 0) an |= (uint8_t)(1u << h); else { if (c & 1) a0 |= (uint8_t)(1u << h); if (c & 2) a1 |= (uint8_t)(1u << h); } }
    t.p0[b] = a0; t.p1[b] = a1; t.pn[b] = an;
  }
  return t;
}
}  // namespace

void ReadStream::produce_bam()
{
  BamReader br;
  if (!br.open_shared(path_, range_.voffset)) { err_ = br.error(); q_.finish(); return; }
  // 4-bit BAM base codes "=ACMGRSVTWYHKDBN": A=1 C=2 G=4 T=8, everything else is treated as N
  static const int8_t code[16] = {-1, 0, 1, -1, 2, -1, -1, -1, 3, -1, -1, -1, -1, -1, -1, -1};
  const size_t CH = 1 << 15;
  std::vector<ReadRec> chunk; chunk.reserve(CH);
  BamRecord r;
  while (br.next(r)) {
    if (r.flag & 0x200) continue;                    // BAM_FQCFAIL reads are skipped (reads.cpp:552)
    ReadRec rr; memset(&rr, 0, sizeof rr);
    rr.id = (uint32_t)atol(r.qname);
    if (rr.id < range_.begin_id) continue;
    if (rr.id >= range_.end_id) break;
    if (r.l_seq > 255) { err_ = path_ + ": read longer than 255 bases"; break; }
    rr.len = (uint32_t)r.l_seq;
    // a BAM byte carries two bases (high nibble first): table -> two bits per plane, 32 bytes per plane word
    { static const Nib2 tab = make_nib2(code);
      const int nbytes = (r.l_seq + 1) >> 1;
      for (int k = 0; k < nbytes; ++k) {
        const uint8_t b = r.seq[k]; const int w = k >> 5, sh = (k & 31) << 1;
        rr.planes[w] |= (uint64_t)tab.p0[b] << sh; rr.planes[4 + w] |= (uint64_t)tab.p1[b] << sh; rr.planes[8 + w] |= (uint64_t)tab.pn[b] << sh;
      }
      if (r.l_seq & 1) {                                  // the unused low nibble of the last byte is not a base
        const int i = r.l_seq, w = i >> 6; const uint64_t keep = ~(1ull << (i & 63));
        rr.planes[w] &= keep; rr.planes[4 + w] &= keep; rr.planes[8 + w] &= keep;
      } }
    chunk.push_back(rr);
    if (chunk.size() >= CH) { q_.push(std::move(chunk)); chunk = std::vector<ReadRec>(); chunk.reserve(CH); }
  }
  if (!br.error().empty() && err_.empty()) err_ = br.error();
  if (!chunk.empty()) q_.push(std::move(chunk));
  add_producer_cpu();
  q_.finish();
}

void ReadStream::produce_fastx()
{
  gzFile f = gzopen(path_.c_str(), "rb");
  if (!f) { err_ = "cannot open " + path_ + " for reading"; q_.finish(); return; }
  gzbuffer(f, 1 << 20);
  const size_t CH = 1 << 15;
  std::vector<ReadRec> chunk; chunk.reserve(CH);
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
    ReadRec rr; rr.id = (uint32_t)atol(l.c_str() + 1);
    if (rr.id < range_.begin_id) continue;             // text read files have no index: a range scans from the start
    if (rr.id >= range_.end_id) break;
    if (seq.size() > 255) { err_ = path_ + ": read longer than 255 bases"; break; }
    for (char& c : seq) c = (char)toupper((unsigned char)c);
    pack_read_ascii(seq.data(), (uint32_t)seq.size(), rr);
    chunk.push_back(rr);
    if (chunk.size() >= CH) { q_.push(std::move(chunk)); chunk = std::vector<ReadRec>(); chunk.reserve(CH); }
  }
  gzclose(f);
  if (!chunk.empty()) q_.push(std::move(chunk));
  q_.finish();
}

bool ReadStream::ensure()
{
  while (pos_ >= cur_.size()) {
    if (end_) return false;
    cur_.clear(); pos_ = 0;
    if (!q_.pop(cur_)) { end_ = true; return false; }
  }
  return true;
}

const ReadRec* ReadStream::get(uint32_t id)
{
  while (ensure()) {
    const ReadRec& r = cur_[pos_];
    if (r.id == id) return &r;             // not consumed: the same read may be requested again
    if (r.id > id) return nullptr;
    ++pos_;
  }
  return nullptr;
}

}  // namespace thbhost
