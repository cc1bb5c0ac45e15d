// thb_bam.hpp -- minimal BGZF/BAM input (and later output) for the stage hosts.
//
// Replaces the vendored samtools-0.1.18 readers the reference goes through (samopen/samread,
// bwt_map.cpp:245-267; reads.cpp:542-560).  BGZF blocks are independent deflate members, so a reader
// inflates a window of blocks with a few worker threads and hands out records from the joined buffer.
#pragma once
#include <cstdint>
#include <cstdio>
#include <memory>
#include <string>
#include <vector>

namespace thbhost {

struct BamHeader {
  std::string text;
  std::vector<std::string> target_name;
  std::vector<uint32_t> target_len;
};

// View of one alignment record inside the reader's buffer (valid until the next call of next()).
struct BamRecord {
  int32_t tid, pos, mtid, mpos, tlen, l_seq;
  uint16_t flag, n_cigar;
  uint8_t mapq;
  const char* qname; int l_qname;          // NUL-terminated
  const uint32_t* cigar;                   // may be unaligned: use cigar_at()
  const uint8_t* seq; const uint8_t* qual; const uint8_t* aux; int l_aux;
  uint32_t cigar_at(int i) const { uint32_t v; __builtin_memcpy(&v, (const uint8_t*)cigar + 4 * i, 4); return v; }
  // integer-valued aux tag (types c C s S i I); returns false if absent
  bool aux_int(const char tag[2], int64_t* out) const;
  const char* aux_str(const char tag[2]) const;   // Z tag or NULL
  char aux_char(const char tag[2]) const;          // A tag or 0
};

class BamReader {
 public:
  BamReader() = default;
  ~BamReader();
  BamReader(const BamReader&) = delete;
  BamReader& operator=(const BamReader&) = delete;
  // Opens and parses the header.  Returns false (message in error()) on failure.
  bool open(const std::string& path, int inflate_threads = 2);
  // Same, for one of several readers of the same file (read-id ranges on different threads): the header -- 100 MB of text for a
  // junction index with a million contigs -- is parsed once per process and shared; a reader that starts at a virtual offset
  // never touches it.
  bool open_shared(const std::string& path, uint64_t virtual_offset, int inflate_threads = 1);
  const BamHeader& header() const { return shared_hdr_ ? *shared_hdr_ : hdr_; }
  bool next(BamRecord& r);                 // false at end of file or on error (check error())
  // Continues at a BGZF virtual offset (compressed block start << 16 | offset inside the inflated block), e.g. one taken from a
  // <file>.index side file (common.h:577-611).  Call after open().
  bool seek(uint64_t virtual_offset);
  const std::string& error() const { return err_; }
  void close();

 private:
  bool refill(size_t need);                // makes >= need bytes available at buf_[pos_..]
  bool read_exact(void* dst, size_t n);
  FILE* f_ = nullptr;
  std::string path_, err_;
  BamHeader hdr_; std::shared_ptr<const BamHeader> shared_hdr_;
  std::vector<uint8_t> buf_; size_t pos_ = 0, end_ = 0;
  std::vector<uint8_t> raw_;
  bool eof_ = false; int threads_ = 2; size_t window_ = 128u << 10;
};

}  // namespace thbhost
