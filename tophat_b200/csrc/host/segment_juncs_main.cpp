// segment_juncs -- drop-in replacement of the reference stage binary (src/segment_juncs.cpp main/driver,
// 5186-5364 / 4704-5184): same argv, same input files, same four output files, same stderr banner and exit
// codes.  The host only does I/O: it merges the id-sorted segment BAM streams into per-read bundles
// (look_for_hit_group / process_next_hit_group rule table, 3823-4123), looks up the mate's hit group
// (find_gaps 3322-3344) and the read sequence (ReadStream::getRead), packs them into the C-ABI batch and calls
// libtophat_b200.so; every per-read computation runs on the GPU.  There is no CPU fallback.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include <chrono>
#include <algorithm>
#include <mutex>
#include <thread>
#include <atomic>
#include "tophat_b200.h"
#include "thb_options.hpp"
#include "thb_input.hpp"

using namespace thbhost;

static void print_usage()
{
  fprintf(stderr, "Usage:   segment_juncs <ref.fa> <segment.juncs> <segment.insertions> <segment.deletions> <segment.fusions> "
                  "<left_reads.fq> <left_reads.bwtout> <left_seg1.bwtout,...,segN.bwtout> "
                  "[right_reads.fq right_reads.bwtout right_seg1.bwtout,...,right_segN.bwtout]\n");
}

[[noreturn]] static void die(const char* fmt, const std::string& a = "")
{
  fprintf(stderr, fmt, a.c_str()); fputc('\n', stderr); exit(1);
}

struct Batch {
  std::vector<thb_bundle> bundles; std::vector<uint16_t> seg_count; std::vector<uint64_t> reads4;   // stride 12 words
  std::vector<thb_hit> hits, partner; uint32_t max_len = 0;
  void clear() { bundles.clear(); seg_count.clear(); reads4.clear(); hits.clear(); partner.clear(); max_len = 0; }
};

struct Stats { uint64_t bundles = 0, hits = 0; double submit_s = 0, sides_s = 0; };
static Stats* g_stats = nullptr;

static std::mutex g_submit_mutex;      // the two mate sides are assembled on two threads; the library is entered by one at a time

static void submit(thb_ctx* ctx, Batch& b, uint32_t nseg, uint64_t& order_base, std::vector<uint64_t>& packed)
{
  if (b.bundles.empty()) return;
  const uint32_t rw = std::max<uint32_t>(1, (b.max_len + 63) / 64);
  const size_t n = b.bundles.size();
  packed.resize(n * 3 * rw);
  for (size_t i = 0; i < n; ++i)
    for (int pl = 0; pl < 3; ++pl)
      memcpy(&packed[(i * 3 + pl) * rw], &b.reads4[i * 12 + pl * 4], rw * sizeof(uint64_t));
  thb_segjuncs_batch sb; memset(&sb, 0, sizeof sb);
  sb.n_bundles = (uint32_t)n; sb.n_segs = nseg; sb.read_words = rw; sb.bundles = b.bundles.data(); sb.seg_count = b.seg_count.data();
  sb.reads = packed.data(); sb.n_hits = b.hits.size(); sb.hits = b.hits.data(); sb.n_partner_hits = b.partner.size();
  sb.partner_hits = b.partner.data(); sb.order_base = order_base;
  {
    std::lock_guard<std::mutex> l(g_submit_mutex);
    const auto s0 = std::chrono::steady_clock::now();
    if (thb_segjuncs_submit(ctx, &sb) != THB_OK) die("Error: thb_segjuncs_submit: %s", thb_last_error(ctx));
    if (g_stats) g_stats->submit_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - s0).count();
  }
  order_base += n;
  b.clear();
}

// One mate side: SegmentSearchWorker::operator() (4565-4663) at -p1, i.e. every hit group in increasing id order.
static void process_side(thb_ctx* ctx, const Options& o, RefTable& rt, std::mutex& rtm, const std::string& reads_fname,
                         const std::vector<std::string>& segs, const std::string& partner_map, const std::string& partner_seg,
                         bool right_mate, uint64_t& order_base, Stats& st, uint32_t begin_id, uint32_t end_id, size_t batch_bundles)
{
  // the read ids [begin_id, end_id) of this side (SegmentSearchWorker's begin_id / end_id, 4565-4663): every stream is positioned
  // by its own .index side file and filtered by id
  const uint32_t nseg = (uint32_t)segs.size();
  if (nseg > THB_MAX_SEGS) die("Error: more than %s segments per read are not supported by the GPU path", std::to_string(THB_MAX_SEGS));
  std::vector<std::unique_ptr<HitStream>> hs;
  for (auto& f : segs) hs.emplace_back(new HitStream(f, rt, rtm, o.p.max_report_intron_length, range_for(f, begin_id, end_id)));
  std::unique_ptr<HitStream> pm, ps;
  if (!partner_map.empty()) pm.reset(new HitStream(partner_map, rt, rtm, o.p.max_report_intron_length, range_for(partner_map, begin_id, end_id)));
  if (!partner_seg.empty()) ps.reset(new HitStream(partner_seg, rt, rtm, o.p.max_report_intron_length, range_for(partner_seg, begin_id, end_id)));
  ReadStream rs(reads_fname, range_for(reads_fname, begin_id, end_id));
  const bool fusion = o.p.fusion_search != 0;
  Batch b; std::vector<uint64_t> packed; std::vector<thb_hit> tmp;
  const size_t BATCH = batch_bundles;
  for (;;) {
    uint32_t id = 0;
    for (auto& h : hs) { const uint32_t g = h->next_group_id(); if (g && (!id || g < id)) id = g; }
    if (!id) break;
    // t = highest segment with hits decides what runs for the read (4092-4117, 4005-4033, 3981)
    int t = -1;
    for (uint32_t s = 0; s < nseg; ++s) if (hs[s]->next_group_id() == id) t = (int)s;
    if (t == 0 && !fusion) { hs[0]->skip_group(); continue; }
    thb_bundle bu; memset(&bu, 0, sizeof bu);
    bu.read_id = id; bu.hit_begin = (uint32_t)b.hits.size(); bu.partner_begin = (uint32_t)b.partner.size();
    for (uint32_t s = 0; s < nseg; ++s) {
      const size_t before = b.hits.size();
      if (hs[s]->next_group_id() == id) hs[s]->next_group(b.hits);
      const size_t c = b.hits.size() - before;
      if (c > 65535) die("Error: more than 65535 hits for one segment of one read");
      b.seg_count.push_back((uint16_t)c);
    }
    uint8_t flags = right_mate ? THB_BUNDLE_RIGHT_MATE : 0;
    if (t > 0) flags |= THB_BUNDLE_INDELS | THB_BUNDLE_GAPS;
    if (fusion) { flags |= THB_BUNDLE_FUSIONS; if (t > 0 && t < (int)nseg - 1) flags |= THB_BUNDLE_FUSIONS_LAST; }
    bu.flags = flags;
    // partner group: the mate's full-read hits, else the mate's last-segment hits (3322-3344; find_fusions 3041-3064)
    if (t > 0 || fusion) {
      bool has = false;
      if (pm) { uint32_t g; while ((g = pm->next_group_id()) != 0 && g < id) pm->skip_group();
                if (g == id) { pm->next_group(b.partner); has = true; } }
      if (!has && ps) { uint32_t g; while ((g = ps->next_group_id()) != 0 && g < id) ps->skip_group();
                        if (g == id) { ps->next_group(b.partner); has = true; } }
      const size_t np = b.partner.size() - bu.partner_begin;
      if (np > 65535) die("Error: more than 65535 partner hits for one read");
      bu.n_partner = (uint16_t)np;
    }
    const ReadRec* rr = rs.get(id);
    if (!rr) {
      if (!rs.ok()) die("Error: %s", rs.error());
      fprintf(stderr, "Error: could not get read# %d from stream!\n", (int)id); exit(1);     // 3347-3351
    }
    bu.read_len = (uint8_t)rr->len;
    b.max_len = std::max(b.max_len, rr->len);
    b.reads4.insert(b.reads4.end(), rr->planes, rr->planes + 12);
    b.bundles.push_back(bu);
    st.bundles++;
    if (b.bundles.size() >= BATCH) { st.hits += b.hits.size(); submit(ctx, b, nseg, order_base, packed); }
  }
  st.hits += b.hits.size();
  submit(ctx, b, nseg, order_base, packed);
  for (auto& h : hs) if (!h->ok()) die("Error: %s", h->error());
  if (pm && !pm->ok()) die("Error: %s", pm->error());
  if (ps && !ps->ok()) die("Error: %s", ps->error());
  if (!rs.ok()) die("Error: %s", rs.error());
}

int main(int argc, char** argv)
{
  fprintf(stderr, "segment_juncs v%s (%s)\n", "2.1.2", "tophat_b200");
  fprintf(stderr, "---------------------------\n");
  Options o;
  if (parse_options(argc, argv, o, print_usage)) return 1;
  const std::vector<std::string>& a = o.positional;
  if (a.size() < 8 || a.size() == 9 || a.size() == 10) { print_usage(); return 1; }
  const std::string ref_fname = a[0], juncs_fname = a[1], ins_fname = a[2], del_fname = a[3], fus_fname = a[4];
  const std::string left_reads = a[5], left_map = a[6];
  const std::vector<std::string> left_segs = split_list(a[7]);
  std::string right_reads, right_map; std::vector<std::string> right_segs;
  if (a.size() >= 11) { right_reads = a[8]; right_map = a[9]; right_segs = split_list(a[10]); }

  { FILE* f = fopen(ref_fname.c_str(), "r"); if (!f) die("Error: cannot open %s for reading", ref_fname); fclose(f); }
  // The four outputs are written under temporary names and renamed when complete: tophat.py --resume skips this stage when
  // segment.juncs exists (tophat.py:3087-3088), so a killed run must never leave a plausible partial file behind.
  auto tmp_of = [](const std::string& n) { return n.compare(0, 5, "/dev/") == 0 ? n : n + ".thb_tmp"; };
  FILE* juncs_out = fopen(tmp_of(juncs_fname).c_str(), "w"); if (!juncs_out) die("Error: cannot open %s for writing", juncs_fname);
  FILE* ins_out = fopen(tmp_of(ins_fname).c_str(), "w"); if (!ins_out) die("Error: cannot open %s for writing", ins_fname);
  FILE* del_out = fopen(tmp_of(del_fname).c_str(), "w"); if (!del_out) die("Error: cannot open %s for writing", del_fname);
  FILE* fus_out = fopen(tmp_of(fus_fname).c_str(), "w"); if (!fus_out) die("Error: cannot open %s for writing", fus_fname);

  auto publish_outputs = [&]() {
    for (const std::string* n : { &juncs_fname, &ins_fname, &del_fname, &fus_fname })
      if (tmp_of(*n) != *n && rename(tmp_of(*n).c_str(), n->c_str()) != 0) die("Error: cannot move the finished output into place: %s", *n);
  };
  if (left_segs.empty()) {                                                                       // 4724-4728: the (empty) files exist
    fclose(juncs_out); fclose(ins_out); fclose(del_out); fclose(fus_out); publish_outputs();
    fprintf(stderr, "No hits to process, exiting\n"); return 0;
  }
  if (!o.no_coverage_search || !o.no_microexon_search || o.butterfly_search)
    die("Error: coverage / microexon / butterfly search are outside the GPU path (tophat.py passes --no-coverage-search "
        "--no-microexon-search for reads of >= 3 segments)");

  const char* dev_env = getenv("TOPHAT_GPU_DEVICE");
  thb_ctx* ctx = nullptr;
  // the CUDA context (0.3 - 1.5 s per process) comes up on its own thread while this one loads the genome
  int create_rc = THB_OK; std::thread ctx_thread([&] { create_rc = thb_create(dev_env ? atoi(dev_env) : 0, &ctx); });
  auto join_ctx = [&]() { if (ctx_thread.joinable()) ctx_thread.join(); if (create_rc != THB_OK) die("Error: %s", thb_last_error(nullptr)); };

  auto t0 = std::chrono::steady_clock::now();
  RefTable rt; std::string err;
  if (!o.sam_header.empty() && !rt.load_sam_header(o.sam_header, &err)) die("%s", err);
  fprintf(stderr, "Loading reference sequences...\n");
  Genome g;
  if (!load_fasta(ref_fname, rt, g, true, std::max(8, o.num_threads), &err)) { join_ctx(); die("Error: %s", err); }
  join_ctx();
  { thb_ref_image img = g.image(); if (thb_ref_upload(ctx, &img) != THB_OK) die("Error: thb_ref_upload: %s", thb_last_error(ctx)); }
  auto t1 = std::chrono::steady_clock::now();
  if (thb_segjuncs_begin(ctx, &o.p) != THB_OK) die("Error: %s", thb_last_error(ctx));
  if (o.p.fusion_search && !o.fusion_ignore_chromosomes.empty()) {                               // 3213-3219
    std::vector<uint32_t> ids;
    for (auto& nm : o.fusion_ignore_chromosomes) ids.push_back(rt.get_id(nm));
    if (thb_segjuncs_fusion_ignore(ctx, ids.data(), (uint32_t)ids.size()) != THB_OK) die("Error: %s", thb_last_error(ctx));
  }

  // Work list = (mate side) x (read-id range), like the reference's threads (4756-4826 left mates, 4835-4905 right mates): the sides
  // are independent passes over disjoint files, the ranges are contiguous id ranges split at entries of the reads file's .index.
  // -p N threads take the tasks; the library is entered by one task at a time (its kernels are a small part of a task).
  // Insertions keep the single-threaded first-wins order: a bundle's priority is its position in the -p1 processing order
  // (left mates by id, then right mates by id) -- order_base = side << 38 | range << 31.
  std::mutex rtm; Stats st; g_stats = &st;
  std::vector<uint32_t> splits;
  if (o.num_threads > 1) {
    BamIndex ix;
    if (ix.load(left_reads) || ix.load(left_segs.back())) splits = split_ids(ix, std::min(o.num_threads, 64));
  }
  struct Task { bool right; uint32_t begin_id, end_id; uint64_t order_base; Stats st; };
  std::vector<Task> tasks;
  for (int side = 0; side < 2; ++side) {
    if ((side == 0 ? left_segs.size() : right_segs.size()) <= 1) continue;
    for (size_t r = 0; r <= splits.size(); ++r)
      tasks.push_back(Task{side == 1, r ? splits[r - 1] : 0u, r < splits.size() ? splits[r] : 0xffffffffu, ((uint64_t)side << 38) | ((uint64_t)r << 31), Stats()});
  }
  fprintf(stderr, ">> Performing segment-search:\n");
  if (left_segs.size() > 1) fprintf(stderr, "Loading left segment hits... done.\n");
  if (right_segs.size() > 1) fprintf(stderr, "Loading right segment hits...done.\n");
  {
    const size_t batch = tasks.size() > 2 ? (1u << 19) : (1u << 21);
    std::atomic<size_t> next(0);
    auto run = [&]() {
      for (;;) {
        const size_t i = next.fetch_add(1); if (i >= tasks.size()) break;
        Task& t = tasks[i];
        if (!t.right) process_side(ctx, o, rt, rtm, left_reads, left_segs, right_map, right_segs.empty() ? std::string() : right_segs.back(), false, t.order_base, t.st, t.begin_id, t.end_id, batch);
        else process_side(ctx, o, rt, rtm, right_reads, right_segs, left_map, left_segs.back(), true, t.order_base, t.st, t.begin_id, t.end_id, batch);
      }
    };
    const int nthr = (int)std::max<size_t>(1, std::min<size_t>((size_t)std::max(1, o.num_threads), tasks.size()));
    std::vector<std::thread> pool;
    for (int k = 1; k < nthr; ++k) pool.emplace_back(run);
    run();
    for (auto& th : pool) th.join();
    for (const Task& t : tasks) { st.bundles += t.st.bundles; st.hits += t.st.hits; }
  }
  st.sides_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
  thb_segjuncs_results r;
  if (thb_segjuncs_finish(ctx, &r) != THB_OK) die("Error: thb_segjuncs_finish: %s", thb_last_error(ctx));
  auto t2 = std::chrono::steady_clock::now();

  fprintf(stderr, "\tfound %ld potential split-segment junctions\n", (long)r.n_junctions);
  fprintf(stderr, "\tfound %ld potential small deletions\n", (long)r.n_deletions);
  fprintf(stderr, "\tfound %ld potential small insertions\n", (long)r.n_insertions);
  for (uint64_t i = 0; i < r.n_junctions; ++i)                                                    // 5035-5053
    fprintf(juncs_out, "%s\t%d\t%d\t%c\n", rt.name(r.junctions[i].ref_id).c_str(), (int)r.junctions[i].left, (int)r.junctions[i].right,
            r.junctions[i].antisense ? '-' : '+');
  fclose(juncs_out);
  fprintf(stderr, "Reported %d total potential splices\n", (int)r.n_junctions);
  fprintf(stderr, "Reporting %lu potential deletions...\n", (unsigned long)r.n_deletions);
  for (uint64_t i = 0; i < r.n_deletions; ++i)                                                    // 5065-5075
    fprintf(del_out, "%s\t%d\t%d\n", rt.name(r.deletions[i].ref_id).c_str(), (int)r.deletions[i].left + 1, (int)r.deletions[i].right);
  fclose(del_out);
  fprintf(stderr, "Reporting %lu potential insertions...\n", (unsigned long)r.n_insertions);
  for (uint64_t i = 0; i < r.n_insertions; ++i)                                                   // 5083-5091
    fprintf(ins_out, "%s\t%d\t%d\t%s\n", rt.name(r.insertions[i].ref_id).c_str(), (int)r.insertions[i].left, (int)r.insertions[i].left,
            r.insertions[i].seq);
  fclose(ins_out);
  {
    // segment.fusions (5096-5180): neighbouring fusions (same contigs, left ends < 10 apart, same direction and the same
    // offset on both sides) compete -- the better supported one stays, ties go to the one whose ends coincide with
    // splice-junction coordinates
    struct Coord { uint32_t ref; int c; bool operator<(const Coord& r) const { return ref < r.ref || (ref == r.ref && c < r.c); } };
    std::vector<Coord> coords;
    if (o.p.fusion_search)
      for (uint64_t i = 0; i < r.n_junctions; ++i) { coords.push_back({r.junctions[i].ref_id, (int)r.junctions[i].left}); coords.push_back({r.junctions[i].ref_id, (int)r.junctions[i].right}); }
    std::sort(coords.begin(), coords.end());
    const uint64_t nf = r.n_fusions;
    std::vector<char> skip(nf, 0), lc(nf, 0), rc(nf, 0);
    for (uint64_t i = 0; i < nf; ++i) {
      lc[i] = std::binary_search(coords.begin(), coords.end(), Coord{r.fusions[i].ref_id1, (int)r.fusions[i].left});
      rc[i] = std::binary_search(coords.begin(), coords.end(), Coord{r.fusions[i].ref_id2, (int)r.fusions[i].right});
    }
    for (uint64_t i = 0; i < nf; ++i) {
      const thb_fusion& f = r.fusions[i];
      for (uint64_t k = i + 1; k < nf; ++k) {
        const thb_fusion& g2 = r.fusions[k];
        const int left_diff = abs((int)f.left - (int)g2.left);
        if (!(f.ref_id1 == g2.ref_id1 && f.ref_id2 == g2.ref_id2 && left_diff < 10)) break;
        if (f.dir == g2.dir && left_diff == abs((int)f.right - (int)g2.right)) {
          if (g2.count > f.count) skip[i] = 1;
          else if (g2.count == f.count) { if ((int)lc[i] + (int)rc[i] < (int)lc[k] + (int)rc[k]) skip[i] = 1; else skip[k] = 1; }
          else skip[k] = 1;
        }
      }
      if (skip[i] && !o.fusion_do_not_resolve_conflicts) continue;
      const char* dir = f.dir == 8 ? "fr" : f.dir == 9 ? "rf" : f.dir == 10 ? "rr" : "ff";
      fprintf(fus_out, "%s\t%d\t%s\t%d\t%s\n", rt.name(f.ref_id1).c_str(), (int)f.left, rt.name(f.ref_id2).c_str(), (int)f.right, dir);
    }
  }
  fclose(fus_out);
  publish_outputs();
  fprintf(stderr, "Reporting potential fusions...\n");
  if (getenv("TOPHAT_GPU_STATS")) {
    thb_timing tm; thb_last_timing(ctx, &tm);
    auto sec = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    fprintf(stderr, "{\"gpu_stats\": {\"bundles\": %llu, \"hits\": %llu, \"ref_load_s\": %.3f, \"search_s\": %.3f, \"scan_kernel_ms\": %.3f, "
                    "\"h2d_ms\": %.3f, \"submit_s\": %.3f, \"begin_and_sides_s\": %.3f, \"main_thread_cpu_s\": %.3f, \"decoder_threads_cpu_s\": %.3f, \"windows\": %llu, \"indel_tasks\": %llu, \"rescue_tasks\": %llu}}\n",
            (unsigned long long)st.bundles, (unsigned long long)st.hits, sec(t0, t1), sec(t1, t2), tm.scan_kernel_ms, tm.h2d_ms, st.submit_s, st.sides_s, thread_cpu_seconds(), producer_cpu_seconds(),
            (unsigned long long)tm.n_windows, (unsigned long long)tm.n_indel_tasks, (unsigned long long)tm.n_rescue_tasks);
  }
  thb_destroy(ctx);
  return 0;
}
