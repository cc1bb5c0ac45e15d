#include "thb_bam.hpp"
#include <zlib.h>
#include <cstring>
#include <thread>
#include <atomic>
#include <map>
#include <memory>
#include <mutex>

namespace thbhost {

namespace {

inline uint16_t rd16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }
inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }

struct BlockRef { size_t raw_off, raw_len, out_off, out_len; };

bool inflate_block(const uint8_t* src, size_t n, uint8_t* dst, size_t dst_len)
{
  z_stream zs; memset(&zs, 0, sizeof zs);
  if (inflateInit2(&zs, -15) != Z_OK) return false;
  zs.next_in = const_cast<Bytef*>(src); zs.avail_in = (uInt)n;
  zs.next_out = dst; zs.avail_out = (uInt)dst_len;
  int rc = inflate(&zs, Z_FINISH);
  bool ok = (rc == Z_STREAM_END) && zs.total_out == dst_len;
  inflateEnd(&zs);
  return ok;
}

// Returns a pointer to the type byte of `tag`, or NULL; a field whose value does not lie wholly inside the record (truncated /
// corrupt input, unterminated string) ends the search.
const uint8_t* aux_find(const uint8_t* p, int n, const char tag[2])
{
  const uint8_t* e = p + n;
  while (p + 3 <= e) {
    const bool hit = p[0] == (uint8_t)tag[0] && p[1] == (uint8_t)tag[1];
    const uint8_t type = p[2];
    const uint8_t* v = p + 3;
    switch (type) {
      case 'A': case 'c': case 'C': v += 1; break;
      case 's': case 'S': v += 2; break;
      case 'i': case 'I': case 'f': v += 4; break;
      case 'd': v += 8; break;
      case 'Z': case 'H': while (v < e && *v) ++v; if (v >= e) return nullptr; ++v; break;
      case 'B': { if (v + 5 > e) return nullptr; const uint8_t st = v[0]; const uint32_t cnt = rd32(v + 1);
                  const int sz = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4;
                  if ((uint64_t)sz * cnt > (uint64_t)(e - v)) return nullptr; v += 5 + (size_t)sz * cnt; break; }
      default: return nullptr;
    }
    if (v > e) return nullptr;
    if (hit) return p + 2;
    p = v;
  }
  return nullptr;
}

}  // namespace

bool BamRecord::aux_int(const char tag[2], int64_t* out) const
{
  const uint8_t* t = aux_find(aux, l_aux, tag);
  if (!t) return false;
  const uint8_t* v = t + 1;
  switch (*t) {
    case 'c': *out = (int8_t)v[0]; return true;
    case 'C': *out = v[0]; return true;
    case 's': { int16_t x; memcpy(&x, v, 2); *out = x; return true; }
    case 'S': *out = rd16(v); return true;
    case 'i': { int32_t x; memcpy(&x, v, 4); *out = x; return true; }
    case 'I': *out = rd32(v); return true;
    default: *out = 0; return true;      // bam_aux2i returns 0 for non-integer types
  }
}
const char* BamRecord::aux_str(const char tag[2]) const
{
  const uint8_t* t = aux_find(aux, l_aux, tag);
  return (t && (*t == 'Z' || *t == 'H')) ? (const char*)(t + 1) : nullptr;
}
char BamRecord::aux_char(const char tag[2]) const
{
  const uint8_t* t = aux_find(aux, l_aux, tag);
  return (t && *t == 'A') ? (char)t[1] : 0;
}

BamReader::~BamReader() { close(); }
void BamReader::close() { if (f_) fclose(f_); f_ = nullptr; }

// Reads the next window of BGZF blocks (up to ~8 MB of raw data), inflates them in parallel and appends the
// result behind the unread tail of buf_.
bool BamReader::refill(size_t need)
{
  while (end_ - pos_ < need) {
    if (eof_) return false;
    if (pos_ > 0) { memmove(buf_.data(), buf_.data() + pos_, end_ - pos_); end_ -= pos_; pos_ = 0; }
    std::vector<BlockRef> blocks; size_t raw_used = 0, out_total = 0;
    // compressed bytes per refill: starts small (a reader of a short id range must not inflate megabytes it will never
    // deliver) and doubles up to 8 MB while the reader keeps going
    const size_t RAW_WINDOW = window_; if (window_ < (8u << 20)) window_ *= 2;
    raw_.resize(RAW_WINDOW + 65536);
    while (raw_used < RAW_WINDOW) {
      uint8_t h[18];
      size_t got = fread(h, 1, 18, f_);
      if (got == 0) { eof_ = true; break; }
      if (got != 18 || h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) { err_ = path_ + ": not a BGZF block"; eof_ = true; return false; }
      const int xlen = rd16(h + 10);
      if (xlen < 6) { err_ = path_ + ": BGZF block without BC field"; eof_ = true; return false; }
      // the BC subfield carries the block size; it is normally the only subfield (xlen == 6)
      int bsize = -1;
      if (h[12] == 'B' && h[13] == 'C' && rd16(h + 14) == 2) bsize = rd16(h + 16);
      if (xlen > 6) {
        std::vector<uint8_t> extra((size_t)xlen); memcpy(extra.data(), h + 12, 6);
        if (fread(extra.data() + 6, 1, (size_t)xlen - 6, f_) != (size_t)(xlen - 6)) { err_ = path_ + ": truncated BGZF header"; eof_ = true; return false; }
        for (int o = 0; bsize < 0 && o + 4 <= xlen;) { const int sl = rd16(extra.data() + o + 2);
          if (extra[o] == 'B' && extra[o + 1] == 'C' && sl == 2 && o + 6 <= xlen) bsize = rd16(extra.data() + o + 4); o += 4 + sl; }
      }
      if (bsize < 0 || bsize + 1 < 12 + xlen + 8) { err_ = path_ + ": bad BGZF block header"; eof_ = true; return false; }
      const size_t remain = (size_t)bsize + 1 - 12 - (size_t)xlen;     // deflate data + crc32 + isize
      if (remain < 8) { err_ = path_ + ": bad BGZF block size"; eof_ = true; return false; }
      if (raw_used + remain > raw_.size()) raw_.resize(raw_used + remain + 65536);
      if (fread(raw_.data() + raw_used, 1, remain, f_) != remain) { err_ = path_ + ": truncated BGZF block"; eof_ = true; return false; }
      const uint32_t isize = rd32(raw_.data() + raw_used + remain - 4);
      if (isize) blocks.push_back({raw_used, remain - 8, out_total, isize});
      raw_used += remain; out_total += isize;
    }
    if (blocks.empty()) { if (eof_) return end_ - pos_ >= need; continue; }
    if (buf_.size() < end_ + out_total) buf_.resize(end_ + out_total);
    uint8_t* dst = buf_.data() + end_;
    std::atomic<size_t> nexti(0); std::atomic<bool> ok(true);
    auto work = [&]() { for (;;) { size_t i = nexti.fetch_add(1); if (i >= blocks.size()) break;
        if (!inflate_block(raw_.data() + blocks[i].raw_off, blocks[i].raw_len, dst + blocks[i].out_off, blocks[i].out_len)) ok = false; } };
    const int nt = (int)std::min<size_t>((size_t)threads_, blocks.size());
    if (nt <= 1) work();
    else { std::vector<std::thread> th; for (int t = 1; t < nt; ++t) th.emplace_back(work); work(); for (auto& t : th) t.join(); }
    if (!ok) { err_ = path_ + ": inflate failed (corrupt BGZF block)"; eof_ = true; return false; }
    end_ += out_total;
  }
  return true;
}

bool BamReader::read_exact(void* dst, size_t n)
{
  if (!refill(n)) return false;
  memcpy(dst, buf_.data() + pos_, n); pos_ += n;
  return true;
}

bool BamReader::open(const std::string& path, int inflate_threads)
{
  close(); path_ = path; err_.clear(); eof_ = false; pos_ = end_ = 0; threads_ = inflate_threads < 1 ? 1 : inflate_threads;
  f_ = fopen(path.c_str(), "rb");
  if (!f_) { err_ = "cannot open " + path + " for reading"; return false; }
  setvbuf(f_, nullptr, _IOFBF, 1 << 20);
  char magic[4]; int32_t l_text = 0, n_ref = 0;
  if (!read_exact(magic, 4) || memcmp(magic, "BAM\1", 4) != 0) { if (err_.empty()) err_ = path + ": not a BAM file"; return false; }
  if (!read_exact(&l_text, 4) || l_text < 0) { err_ = path + ": bad BAM header"; return false; }
  hdr_.text.resize((size_t)l_text);
  if (l_text && !read_exact(&hdr_.text[0], (size_t)l_text)) { err_ = path + ": truncated BAM header"; return false; }
  if (!read_exact(&n_ref, 4) || n_ref < 0) { err_ = path + ": bad BAM header"; return false; }
  hdr_.target_name.clear(); hdr_.target_len.clear();
  for (int i = 0; i < n_ref; ++i) {
    int32_t l_name = 0; uint32_t l_ref = 0;
    if (!read_exact(&l_name, 4) || l_name <= 0) { err_ = path + ": bad BAM reference entry"; return false; }
    std::string name((size_t)l_name, '\0');
    if (!read_exact(&name[0], (size_t)l_name) || !read_exact(&l_ref, 4)) { err_ = path + ": truncated BAM header"; return false; }
    name.resize(strlen(name.c_str()));
    hdr_.target_name.push_back(name); hdr_.target_len.push_back(l_ref);
  }
  return true;
}

bool BamReader::open_shared(const std::string& path, uint64_t voff, int inflate_threads)
{
  static std::mutex m; static std::map<std::string, std::shared_ptr<const BamHeader>> cache;
  std::shared_ptr<const BamHeader> h;
  { std::lock_guard<std::mutex> l(m); auto it = cache.find(path); if (it != cache.end()) h = it->second; }
  if (!h || voff == 0) {
    if (!open(path, inflate_threads)) return false;                 // parses the header (and leaves the reader behind it)
    if (!h) { std::lock_guard<std::mutex> l(m); auto& slot = cache[path]; if (!slot) slot = std::make_shared<BamHeader>(hdr_); h = slot; }
    shared_hdr_ = h; hdr_ = BamHeader();
    return voff == 0 ? true : seek(voff);
  }
  close(); path_ = path; err_.clear(); eof_ = false; pos_ = end_ = 0; threads_ = inflate_threads < 1 ? 1 : inflate_threads;
  f_ = fopen(path.c_str(), "rb");
  if (!f_) { err_ = "cannot open " + path + " for reading"; return false; }
  setvbuf(f_, nullptr, _IOFBF, 1 << 20);
  shared_hdr_ = h;
  return seek(voff);
}

bool BamReader::seek(uint64_t voff)
{
  if (!f_) return false;
  if (fseeko(f_, (off_t)(voff >> 16), SEEK_SET) != 0) { err_ = path_ + ": cannot seek"; return false; }
  pos_ = end_ = 0; eof_ = false;
  const size_t within = (size_t)(voff & 0xffffu);
  if (within) { if (!refill(within)) { if (err_.empty()) err_ = path_ + ": bad virtual offset"; return false; } pos_ += within; }
  return true;
}

bool BamReader::next(BamRecord& r)
{
  if (!refill(4)) return false;
  const int32_t bs = (int32_t)rd32(buf_.data() + pos_);
  if (bs < 32) { err_ = path_ + ": corrupt BAM record"; return false; }
  if (!refill(4 + (size_t)bs)) { if (err_.empty()) err_ = path_ + ": truncated BAM record"; return false; }
  const uint8_t* p = buf_.data() + pos_ + 4;
  r.tid = (int32_t)rd32(p); r.pos = (int32_t)rd32(p + 4);
  r.l_qname = p[8]; r.mapq = p[9]; r.n_cigar = rd16(p + 12); r.flag = rd16(p + 14);
  r.l_seq = (int32_t)rd32(p + 16); r.mtid = (int32_t)rd32(p + 20); r.mpos = (int32_t)rd32(p + 24); r.tlen = (int32_t)rd32(p + 28);
  // the variable-length fields must fit the record before any pointer into them is formed
  if (r.l_seq < 0 || r.l_qname < 1 ||
      32ull + (uint64_t)r.l_qname + 4ull * (uint64_t)r.n_cigar + ((uint64_t)r.l_seq + 1) / 2 + (uint64_t)r.l_seq > (uint64_t)bs ||
      p[32 + r.l_qname - 1] != 0) { err_ = path_ + ": corrupt BAM record"; return false; }
  const uint8_t* q = p + 32;
  r.qname = (const char*)q; q += r.l_qname;
  r.cigar = (const uint32_t*)q; q += 4 * (size_t)r.n_cigar;
  r.seq = q; q += (size_t)(r.l_seq + 1) / 2;
  r.qual = q; q += (size_t)r.l_seq;
  const uint8_t* e = p + bs;
  if (q > e) { err_ = path_ + ": corrupt BAM record"; return false; }
  r.aux = q; r.l_aux = (int)(e - q);
  pos_ += 4 + (size_t)bs;
  return true;
}

}  // namespace thbhost
