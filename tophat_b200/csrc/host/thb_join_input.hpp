// thb_join_input.hpp -- hit streams of long_spanning_reads: BAMHitFactory (contiguous segment hits) and
// SplicedBAMHitFactory (hits against juncs_db contigs mapped back to genomic coordinates), each producing the
// packed thb_jhit_full record of the C ABI, plus the full read records (name, bases, qualities) the BAM writer needs.
#pragma once
#include "thb_input.hpp"

namespace thbhost {

struct JHitRec { uint32_t id; thb_jhit_full h; };

class JoinHitStream {
 public:
  // spliced = SplicedBAMHitFactory semantics (bwt_map.cpp:1469-1770), else BAMHitFactory (1101-1452)
  JoinHitStream(const std::string& path, RefTable& rt, std::mutex& rt_mutex, bool spliced, int max_report_intron, int min_anchor_len,
                StreamRange range = StreamRange());
  ~JoinHitStream();
  bool ok() const { return err_.empty(); }
  const std::string& error() const { return err_; }
  uint32_t next_group_id();
  void next_group(std::vector<thb_jhit_full>& out);
  void skip_group();
  uint64_t dropped_long_cigars() const { return dropped_; }
 private:
  void produce();
  bool ensure();
  std::string path_, err_;
  RefTable& rt_; std::mutex& rt_mutex_; bool spliced_; int max_report_intron_, min_anchor_len_; StreamRange range_;
  ChunkQueue<JHitRec> q_; std::thread th_;
  std::vector<JHitRec> cur_; size_t pos_ = 0; bool end_ = false; uint64_t dropped_ = 0;
};

struct FullRead { uint32_t id; std::string name, seq, qual; };

class FullReadStream {
 public:
  explicit FullReadStream(const std::string& path, StreamRange range = StreamRange());
  ~FullReadStream();
  bool ok() const { return err_.empty(); }
  const std::string& error() const { return err_; }
  const FullRead* get(uint32_t id);        // increasing ids; NULL if absent
  const FullRead* next();                  // sequential access: the next read of the stream, NULL at its end
 private:
  void produce(); void produce_fastx();
  bool ensure();
  std::string path_, err_; StreamRange range_;
  ChunkQueue<FullRead> q_; std::thread th_;
  std::vector<FullRead> cur_; size_t pos_ = 0; bool end_ = false;
};

}  // namespace thbhost
