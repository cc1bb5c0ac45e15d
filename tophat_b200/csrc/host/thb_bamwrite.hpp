// thb_bamwrite.hpp -- BAM output of long_spanning_reads: what GBamWriter / GBamRecord (common.h:486-611,
// common.cpp:974-1200) produce through samtools, written directly as BGZF blocks, plus the "<id>\t<offset>" side
// index the reference writes every >= 1000 records for its thread partitioning (common.h:577-611).
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace thbhost {

class BamWriter {
 public:
  ~BamWriter();
  // header_sam: text of the --sam-header file (becomes the BAM text header; @SQ lines define the targets)
  bool open(const std::string& path, const std::string& header_sam_path, const std::string& index_path, std::string* err);
  int target_id(const std::string& name) const;
  // one alignment; cigar ops as (len << 4 | BAM op code 0..8); aux = already encoded tag bytes
  void write(const std::string& qname, uint32_t read_id, int flag, int tid, int pos0, int mapq, const std::vector<uint32_t>& cigar,
             const std::string& seq, const std::string& qual, const std::vector<uint8_t>& aux);
  // The same in two steps, for batches: `encode` serialises one alignment record (thread-safe, no writer state) behind
  // `out`; `append_records` takes the records of a batch in output order -- any number of byte buffers, each with the sizes
  // and read ids of its records -- forms the BGZF blocks (same boundaries as record-by-record `write` calls), deflates them on
  // `threads` threads and writes blocks and side-index entries in order.
  static void encode(std::vector<uint8_t>& out, const std::string& qname, int flag, int tid, int pos0, int mapq, const std::vector<uint32_t>& cigar,
                     const std::string& seq, const std::string& qual, const std::vector<uint8_t>& aux);
  struct RecordPart { std::vector<uint8_t> bytes; std::vector<uint32_t> sizes, ids; };
  void append_records(const std::vector<RecordPart>& parts, int threads);
  bool close(std::string* err);
  uint64_t written() const { return wcount_; }
  // aux encoders (GBamRecord::add_aux typing rules, common.cpp:1111-1200)
  static void aux_int(std::vector<uint8_t>& a, const char tag[2], long long v);
  static void aux_char(std::vector<uint8_t>& a, const char tag[2], char c);
  static void aux_str(std::vector<uint8_t>& a, const char tag[2], const std::string& s);
 private:
  void put(const void* p, size_t n);
  void flush_block();
  int64_t tell() const { return (int64_t)((file_off_ << 16) | (uint64_t)blk_.size()); }
  FILE* f_ = nullptr; FILE* fidx_ = nullptr;
  std::vector<uint8_t> blk_; uint64_t file_off_ = 0; bool fail_ = false;
  std::vector<std::string> tnames_;
  std::string path_, index_path_; bool tmp_ = false;
  uint64_t wcount_ = 0; int idxcount_ = 0; long idx_last_id_ = 0;
};

}  // namespace thbhost
