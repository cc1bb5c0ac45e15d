// thb_input.hpp -- host-side ingestion for the stage binaries: reference table + genome image, hit
// streams, read stream.  Each input file is decoded by its own producer thread into compact records
// (no BowtieHit / Read objects), which the bundle builder merges by read id.
#pragma once
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "tophat_b200.h"
#include "thb_bam.hpp"

namespace thbhost {

// "<offset>:<seg>:<nsegs>" behind the last '|' of a segment read name (bwt_map.cpp:1126-1143): is this the read's last segment?
// Same answer as sscanf(s, "%u:%u:%u") followed by seg + 1 == nsegs (fields that do not parse stay 0).
inline bool last_segment_from_suffix(const char* s)
{
  unsigned long sn = 0, ns = 0; char* e = nullptr;
  strtoul(s, &e, 10);
  if (e != s && *e == ':') {
    const char* p2 = e + 1; const unsigned long v2 = strtoul(p2, &e, 10);
    if (e != p2) { sn = v2; if (*e == ':') { const char* p3 = e + 1; const unsigned long v3 = strtoul(p3, &e, 10); if (e != p3) ns = v3; } }
  }
  return sn + 1 == ns;
}

// CPU seconds spent by the decoder threads (TOPHAT_GPU_STATS)
void add_producer_cpu();            // called by a decoder thread when it finishes: adds its own CPU time
double producer_cpu_seconds();
double thread_cpu_seconds();        // CPU time of the calling thread

// RefSequenceTable (bwt_map.h:579-788): ids are 1-based in first-seen order, SAM header first.
class RefTable {
 public:
  uint32_t get_id(const std::string& name);            // never returns 0; appends unknown names
  uint32_t find(const std::string& name) const;        // 0 if unknown
  const std::string& name(uint32_t id) const { return names_[id - 1]; }
  uint32_t size() const { return (uint32_t)names_.size(); }
  bool load_sam_header(const std::string& path, std::string* err);
 private:
  std::vector<std::string> names_;
  std::map<std::string, uint32_t> by_name_;
};

// Genome as bit planes in the layout thb_ref_upload expects.
struct Genome {
  std::vector<uint64_t> contig_start; std::vector<uint32_t> contig_len;
  std::vector<uint64_t> planes, nmask; uint64_t n_blocks = 0;
  thb_ref_image image() const;
};
// get_seqs (segment_juncs.cpp:64-88): names cut at the first space/tab/CR; every byte of a sequence line
// except CR/LF is a base; contigs missing from the FASTA keep length 0.  `log` mirrors the reference's
// "Loading <name>... done (<n> bases)." lines on stderr.
bool load_fasta(const std::string& path, RefTable& rt, Genome& g, bool log, int threads, std::string* err);

// ---- bounded chunk queue between a producer thread and the consumer ---------------------------------
template <typename T>
class ChunkQueue {
 public:
  explicit ChunkQueue(size_t depth = 4) : depth_(depth) {}
  void push(std::vector<T>&& c) { std::unique_lock<std::mutex> l(m_); cv_.wait(l, [&] { return q_.size() < depth_ || stop_; }); if (stop_) return; q_.push_back(std::move(c)); cv_.notify_all(); }
  void finish() { std::lock_guard<std::mutex> l(m_); done_ = true; cv_.notify_all(); }
  void stop() { std::lock_guard<std::mutex> l(m_); stop_ = true; cv_.notify_all(); }
  bool pop(std::vector<T>& c) { std::unique_lock<std::mutex> l(m_); cv_.wait(l, [&] { return !q_.empty() || done_ || stop_; });
    if (q_.empty()) return false; c = std::move(q_.front()); q_.pop_front(); cv_.notify_all(); return true; }
 private:
  std::mutex m_; std::condition_variable cv_; std::deque<std::vector<T>> q_; size_t depth_; bool done_ = false, stop_ = false;
};

// A contiguous read-id range of an id-sorted BAM stream: where to start reading (BGZF virtual offset, 0 = behind the header) and
// the ids to deliver, [begin_id, end_id).  The reference partitions its threads the same way (utils.cpp:22-129), but starts every
// stream at the index entry of the LAST file's split id and so can lose hit groups at the range boundaries; here every stream is
// positioned by its own index and filtered by id, so the union over the ranges is exactly the whole file.
struct StreamRange { uint64_t voffset = 0; uint32_t begin_id = 0; uint32_t end_id = 0xffffffffu; };

// <bam>.index side file: "read_id <TAB> virtual offset" of the first record of that read, an entry every >= 1000 records.
struct BamIndex {
  std::vector<std::pair<uint32_t, uint64_t>> entries;
  bool load(const std::string& bam_path);                 // false if the side file is missing / empty
  uint64_t offset_for(uint32_t begin_id) const;            // offset of the last entry with id <= begin_id, else 0
};
// Split ids for `parts` ranges with about equal numbers of index entries; fewer parts when the index is short.  Returns the
// begin ids of parts 1.. (part 0 begins at id 0).
std::vector<uint32_t> split_ids(const BamIndex& idx, int parts);
StreamRange range_for(const std::string& bam_path, uint32_t begin_id, uint32_t end_id);

struct HitRec { uint32_t id; thb_hit h; };

// HitStream + BAMHitFactory::get_hit_from_buf (bwt_map.h:1040-1227; bwt_map.cpp:1101-1452) reduced to the
// fields the hot path reads.  Groups = maximal runs of records with the same numeric qname prefix.
class HitStream {
 public:
  HitStream(const std::string& path, RefTable& rt, std::mutex& rt_mutex, int max_report_intron, StreamRange range = StreamRange());
  ~HitStream();
  bool ok() const { return err_.empty(); }
  const std::string& error() const { return err_; }
  uint32_t next_group_id();                         // 0 at end of stream (HitStream::next_group_id)
  void next_group(std::vector<thb_hit>& out);       // appends the hits of the next group and consumes it
  void skip_group();
  uint64_t records() const { return n_records_; }
 private:
  void produce();
  bool ensure();
  std::string path_, err_;
  RefTable& rt_; std::mutex& rt_mutex_; int max_report_intron_; StreamRange range_;
  ChunkQueue<HitRec> q_; std::thread th_;
  std::vector<HitRec> cur_; size_t pos_ = 0; bool end_ = false; uint64_t n_records_ = 0;
};

struct ReadRec { uint32_t id; uint32_t len; uint64_t planes[12]; };   // plane0[4] | plane1[4] | planeN[4]

// ReadStream::getRead (reads.cpp:571-630) for BAM and FASTA/FASTQ read files with numeric names.
class ReadStream {
 public:
  explicit ReadStream(const std::string& path, StreamRange range = StreamRange());
  ~ReadStream();
  bool ok() const { return err_.empty(); }
  const std::string& error() const { return err_; }
  // reads must be requested in increasing id order; returns NULL if the id is not in the file
  const ReadRec* get(uint32_t id);
 private:
  void produce_bam(); void produce_fastx();
  bool ensure();
  std::string path_, err_; StreamRange range_;
  ChunkQueue<ReadRec> q_; std::thread th_;
  std::vector<ReadRec> cur_; size_t pos_ = 0; bool end_ = false;
};

void pack_read_ascii(const char* s, uint32_t len, ReadRec& r);

}  // namespace thbhost
