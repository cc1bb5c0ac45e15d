#include "thb_bamwrite.hpp"
#include <zlib.h>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <thread>
#include <algorithm>

namespace thbhost {

namespace {
const size_t BLOCK_DATA = 0xff00;        // uncompressed payload per BGZF block (samtools uses 64 KiB - slack)
int reg2bin(int beg, int end)
{
  --end;
  if (beg >> 14 == end >> 14) return ((1 << 15) - 1) / 7 + (beg >> 14);
  if (beg >> 17 == end >> 17) return ((1 << 12) - 1) / 7 + (beg >> 17);
  if (beg >> 20 == end >> 20) return ((1 << 9) - 1) / 7 + (beg >> 20);
  if (beg >> 23 == end >> 23) return ((1 << 6) - 1) / 7 + (beg >> 23);
  if (beg >> 26 == end >> 26) return ((1 << 3) - 1) / 7 + (beg >> 26);
  return 0;
}
}  // namespace

// Deflates one BGZF block payload (<= BLOCK_DATA bytes) into `out` (header + data + crc + isize); returns the block's size or 0.
// BSIZE is 16 bits: a payload that deflate cannot shrink below 64 KiB - 26 is stored instead (level 0: payload + 5 bytes always fits).
// Compression level: the stage's BAM is an intermediate file that tophat_reports reads once; level 1 deflates the 4-bit packed
// bases and the quality strings three to four times faster than zlib's default 6 (which the reference uses through samtools'
// bgzf) for about a tenth more bytes.  TOPHAT_GPU_BAM_LEVEL=<0..9> overrides.
static int bgzf_level()
{
  static const int v = [] { const char* e = getenv("TOPHAT_GPU_BAM_LEVEL"); const int l = e ? atoi(e) : 1; return l < 0 ? 0 : (l > 9 ? 9 : l); }();
  return v;
}
static size_t bgzf_compress(const uint8_t* src, size_t n, uint8_t* out, size_t out_cap)
{
  size_t clen = 0;
  for (int level : {bgzf_level(), 0}) {
    z_stream zs; memset(&zs, 0, sizeof zs);
    deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
    zs.next_in = const_cast<Bytef*>(src); zs.avail_in = (uInt)n; zs.next_out = out + 18; zs.avail_out = (uInt)(out_cap - 18 - 8);
    const int rc = deflate(&zs, Z_FINISH); clen = zs.total_out; deflateEnd(&zs);
    if (rc == Z_STREAM_END && clen + 26 <= 65536) break;
    if (level == 0) return 0;
  }
  const uint8_t hdr[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
  memcpy(out, hdr, 16);
  const uint16_t bsize = (uint16_t)(clen + 18 + 8 - 1); memcpy(out + 16, &bsize, 2);
  const uint32_t crc = (uint32_t)crc32(crc32(0L, nullptr, 0), src, (uInt)n), isize = (uint32_t)n;
  memcpy(out + 18 + clen, &crc, 4); memcpy(out + 18 + clen + 4, &isize, 4);
  return clen + 26;
}

BamWriter::~BamWriter() { std::string e; close(&e); }

bool BamWriter::open(const std::string& path, const std::string& header_sam_path, const std::string& index_path, std::string* err)
{
  // written under a temporary name, renamed by close(): a killed run leaves no plausible partial BAM for tophat.py --resume
  path_ = path; index_path_ = index_path; tmp_ = path.compare(0, 5, "/dev/") != 0;
  f_ = fopen((tmp_ ? path + ".thb_tmp" : path).c_str(), "wb");
  if (!f_) { *err = "cannot open " + path + " for writing"; return false; }
  if (!index_path.empty()) fidx_ = fopen((tmp_ ? index_path + ".thb_tmp" : index_path).c_str(), "w");
  std::string text; std::vector<uint32_t> tlen;
  { std::ifstream in(header_sam_path.c_str()); if (!in.good()) { *err = "Failed to open SAM header file " + header_sam_path; return false; }
    std::stringstream ss; ss << in.rdbuf(); text = ss.str(); }
  { std::istringstream ls(text); std::string line;
    while (std::getline(ls, line)) {
      if (line.compare(0, 3, "@SQ") != 0) continue;
      std::string sn; uint32_t ln = 0; size_t p = 0;
      while ((p = line.find('\t', p)) != std::string::npos) { ++p; size_t e = line.find_first_of("\t\r\n", p);
        const std::string fld = line.substr(p, e == std::string::npos ? std::string::npos : e - p);
        if (fld.compare(0, 3, "SN:") == 0) sn = fld.substr(3); else if (fld.compare(0, 3, "LN:") == 0) ln = (uint32_t)strtoul(fld.c_str() + 3, nullptr, 10); }
      tnames_.push_back(sn); tlen.push_back(ln);
    } }
  put("BAM\1", 4);
  const int32_t l_text = (int32_t)text.size(); put(&l_text, 4); put(text.data(), text.size());
  const int32_t n_ref = (int32_t)tnames_.size(); put(&n_ref, 4);
  for (size_t i = 0; i < tnames_.size(); ++i) { const int32_t l = (int32_t)tnames_[i].size() + 1; put(&l, 4); put(tnames_[i].c_str(), (size_t)l); put(&tlen[i], 4); }
  flush_block();
  return true;
}

int BamWriter::target_id(const std::string& name) const
{ for (size_t i = 0; i < tnames_.size(); ++i) if (tnames_[i] == name) return (int)i; return -1; }

void BamWriter::put(const void* p, size_t n)
{
  const uint8_t* s = (const uint8_t*)p;
  while (n) { const size_t room = BLOCK_DATA - blk_.size(); const size_t k = n < room ? n : room; blk_.insert(blk_.end(), s, s + k); s += k; n -= k;
    if (blk_.size() >= BLOCK_DATA) flush_block(); }
}

void BamWriter::flush_block()
{
  if (blk_.empty() || !f_) return;
  uint8_t out[70000];
  const size_t total = bgzf_compress(blk_.data(), blk_.size(), out, sizeof out);
  if (!total) { fail_ = true; return; }
  if (fwrite(out, 1, total, f_) != total) fail_ = true;
  file_off_ += total; blk_.clear();
}

void BamWriter::encode(std::vector<uint8_t>& rec, const std::string& qname, int flag, int tid, int pos0, int mapq, const std::vector<uint32_t>& cigar,
                       const std::string& seq, const std::string& qual, const std::vector<uint8_t>& aux)
{
  int end = pos0; for (uint32_t c : cigar) { const int op = (int)(c & 15); if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) end += (int)(c >> 4); }
  if (end == pos0) end = pos0 + 1;
  const int32_t l_seq = (int32_t)seq.size();
  const int32_t block_size = 32 + (int32_t)qname.size() + 1 + 4 * (int32_t)cigar.size() + (l_seq + 1) / 2 + l_seq + (int32_t)aux.size();
  auto p32 = [&](int32_t v) { const uint8_t* b = (const uint8_t*)&v; rec.insert(rec.end(), b, b + 4); };
  p32(block_size); p32(tid); p32(pos0);
  const uint32_t bin_mq_nl = ((uint32_t)reg2bin(pos0, end) << 16) | ((uint32_t)mapq << 8) | (uint32_t)(qname.size() + 1);
  p32((int32_t)bin_mq_nl); p32((int32_t)(((uint32_t)flag << 16) | (uint32_t)cigar.size()));
  p32(l_seq); p32(-1); p32(-1); p32(0);
  rec.insert(rec.end(), qname.begin(), qname.end()); rec.push_back(0);
  for (uint32_t c : cigar) p32((int32_t)c);
  auto nt = [](char c) -> uint8_t { switch (c) { case '=': return 0; case 'A': case 'a': return 1; case 'C': case 'c': return 2; case 'M': return 3; case 'G': case 'g': return 4;
    case 'R': return 5; case 'S': return 6; case 'V': return 7; case 'T': case 't': return 8; case 'W': return 9; case 'Y': return 10; case 'H': return 11; case 'K': return 12;
    case 'D': return 13; case 'B': return 14; default: return 15; } };
  for (int i = 0; i < l_seq; i += 2) rec.push_back((uint8_t)((nt(seq[i]) << 4) | (i + 1 < l_seq ? nt(seq[i + 1]) : 0)));
  for (int i = 0; i < l_seq; ++i) rec.push_back((uint8_t)(i < (int)qual.size() ? qual[i] - 33 : 0xff));
  rec.insert(rec.end(), aux.begin(), aux.end());
}

void BamWriter::write(const std::string& qname, uint32_t read_id, int flag, int tid, int pos0, int mapq, const std::vector<uint32_t>& cigar,
                      const std::string& seq, const std::string& qual, const std::vector<uint8_t>& aux)
{
  // side index: an entry every >= 1000 records, at a read-id boundary (common.h:577-611)
  int64_t pre_pos = 0; bool write_index = false;
  if (fidx_ && read_id) {
    if (idxcount_ >= 1000 && (long)read_id != idx_last_id_) { pre_pos = tell(); write_index = true; }
    idx_last_id_ = (long)read_id; idxcount_++;
  }
  std::vector<uint8_t> rec;
  encode(rec, qname, flag, tid, pos0, mapq, cigar, seq, qual, aux);
  // samtools' bgzf_write flushes the current block first when a record does not fit in it
  if (blk_.size() + rec.size() > BLOCK_DATA && !blk_.empty() && rec.size() <= BLOCK_DATA) flush_block();
  put(rec.data(), rec.size());
  wcount_++;
  if (write_index) { fprintf(fidx_, "%ld\t%ld\n", (long)read_id, (long)pre_pos); idxcount_ = 0; }
}

void BamWriter::append_records(const std::vector<RecordPart>& parts, int threads)
{
  if (!f_) return;
  // 1. block formation, exactly as a sequence of write() calls would do it: `payload` = the bytes of the blocks completed by
  //    this batch (the first one starts with what blk_ already holds); the tail stays in blk_ for the next batch / close()
  std::vector<uint8_t> payload(blk_.begin(), blk_.end());
  std::vector<size_t> block_start(1, 0);                   // offsets into payload where completed blocks begin
  struct IdxEntry { long id; size_t block; size_t in_block; };
  std::vector<IdxEntry> idx;
  size_t cur = blk_.size();                                // bytes in the block being filled
  auto close_block = [&]() { block_start.push_back(payload.size()); cur = 0; };
  for (const RecordPart& p : parts) {
    size_t off = 0;
    for (size_t r = 0; r < p.sizes.size(); ++r) {
      const size_t n = p.sizes[r]; const uint32_t read_id = p.ids[r];
      bool write_index = false; IdxEntry ie{0, 0, 0};
      if (fidx_ && read_id) {          // position taken before a possible flush, like write() does with tell()
        if (idxcount_ >= 1000 && (long)read_id != idx_last_id_) { ie = IdxEntry{(long)read_id, block_start.size() - 1, cur}; write_index = true; }
        idx_last_id_ = (long)read_id; idxcount_++;
      }
      if (cur + n > BLOCK_DATA && cur > 0 && n <= BLOCK_DATA) close_block();
      size_t left = n; const uint8_t* src = p.bytes.data() + off;
      while (left) { const size_t room = BLOCK_DATA - cur; const size_t k = left < room ? left : room;
        payload.insert(payload.end(), src, src + k); src += k; left -= k; cur += k; if (cur >= BLOCK_DATA) close_block(); }
      off += n; wcount_++;
      if (write_index) { idx.push_back(ie); idxcount_ = 0; }
    }
  }
  const size_t n_blocks = block_start.size() - 1;          // completed blocks; payload[block_start.back() ..) is the new tail
  // 2. parallel deflate of the completed blocks
  std::vector<std::vector<uint8_t>> comp(n_blocks); std::vector<size_t> clen(n_blocks, 0);
  auto work = [&](size_t t, size_t nt) {
    for (size_t b = t; b < n_blocks; b += nt) {
      comp[b].resize(70000);
      clen[b] = bgzf_compress(payload.data() + block_start[b], block_start[b + 1] - block_start[b], comp[b].data(), comp[b].size());
    }
  };
  const size_t nt = (size_t)std::max(1, std::min<int>(threads, (int)n_blocks));
  if (nt <= 1) work(0, 1);
  else { std::vector<std::thread> th; for (size_t t = 1; t < nt; ++t) th.emplace_back(work, t, nt); work(0, nt); for (auto& x : th) x.join(); }
  // 3. ordered output + side index (virtual offset = file offset of the block << 16 | offset inside it)
  std::vector<uint64_t> block_file_off(n_blocks + 1, file_off_);
  for (size_t b = 0; b < n_blocks; ++b) {
    if (clen[b] == 0 || fwrite(comp[b].data(), 1, clen[b], f_) != clen[b]) fail_ = true;
    file_off_ += clen[b]; block_file_off[b + 1] = file_off_;
  }
  if (fidx_) for (const IdxEntry& e : idx) fprintf(fidx_, "%ld\t%ld\n", e.id, (long)((block_file_off[e.block] << 16) | (uint64_t)e.in_block));
  blk_.assign(payload.begin() + (long)block_start.back(), payload.end());
}

bool BamWriter::close(std::string* err)
{
  if (!f_) return !fail_;
  flush_block();
  static const uint8_t eof_block[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (fwrite(eof_block, 1, 28, f_) != 28) fail_ = true;
  if (fclose(f_) != 0) fail_ = true;
  f_ = nullptr;
  const bool had_index = fidx_ != nullptr;
  if (fidx_) { if (fclose(fidx_) != 0) fail_ = true; fidx_ = nullptr; }
  if (!fail_ && tmp_) {
    if (had_index && rename((index_path_ + ".thb_tmp").c_str(), index_path_.c_str()) != 0) fail_ = true;
    if (rename((path_ + ".thb_tmp").c_str(), path_.c_str()) != 0) fail_ = true;
  }
  if (fail_ && err) *err = "error writing BAM output";
  return !fail_;
}

void BamWriter::aux_int(std::vector<uint8_t>& a, const char tag[2], long long x)
{
  a.push_back((uint8_t)tag[0]); a.push_back((uint8_t)tag[1]);
  auto putn = [&](char t, const void* p, int n) { a.push_back((uint8_t)t); const uint8_t* b = (const uint8_t*)p; a.insert(a.end(), b, b + n); };
  if (x < 0) {
    if (x >= -127) { const int8_t v = (int8_t)x; putn('c', &v, 1); }
    else if (x >= -32767) { const int16_t v = (int16_t)x; putn('s', &v, 2); }
    else { const int32_t v = (int32_t)x; putn('i', &v, 4); }
  } else {
    if (x <= 255) { const uint8_t v = (uint8_t)x; putn('C', &v, 1); }
    else if (x <= 65535) { const uint16_t v = (uint16_t)x; putn('S', &v, 2); }
    else { const uint32_t v = (uint32_t)x; putn('I', &v, 4); }
  }
}
void BamWriter::aux_char(std::vector<uint8_t>& a, const char tag[2], char c)
{ a.push_back((uint8_t)tag[0]); a.push_back((uint8_t)tag[1]); a.push_back('A'); a.push_back((uint8_t)c); }
void BamWriter::aux_str(std::vector<uint8_t>& a, const char tag[2], const std::string& s)
{ a.push_back((uint8_t)tag[0]); a.push_back((uint8_t)tag[1]); a.push_back('Z'); a.insert(a.end(), s.begin(), s.end()); a.push_back(0); }

}  // namespace thbhost
