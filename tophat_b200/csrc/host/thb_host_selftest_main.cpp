// thb_host_selftest -- exercises the host-side I/O layer without a GPU (used by the CPU test-suite):
//   bam   <in.bam> <out.bam> <header.sam>   BamReader -> BamWriter round trip (every field, aux bytes verbatim)
//   fasta <ref.fa> <header.sam> <out.bin>   RefTable + load_fasta -> dump of the genome image
//   hits  <seg.bam> <header.sam>            HitStream records as text: id ref_id left right read_len edit flags
//   jhits <seg.bam> <header.sam> <spliced 0|1>   JoinHitStream records as text
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "tophat_b200.h"
#include "thb_bam.hpp"
#include "thb_bamwrite.hpp"
#include "thb_input.hpp"
#include "thb_join_input.hpp"

using namespace thbhost;

int main(int argc, char** argv)
{
  if (argc < 2) return 2;
  const std::string mode = argv[1];
  std::string err;
  if (mode == "bam" && argc == 5) {
    BamReader br; if (!br.open(argv[2])) { fprintf(stderr, "%s\n", br.error().c_str()); return 1; }
    BamWriter bw; if (!bw.open(argv[3], argv[4], std::string(argv[3]) + ".index", &err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    static const char* nt16 = "=ACMGRSVTWYHKDBN";
    BamRecord r; std::vector<uint32_t> cig; std::vector<uint8_t> aux; std::string seq, qual;
    while (br.next(r)) {
      cig.clear(); for (int i = 0; i < r.n_cigar; ++i) cig.push_back(r.cigar_at(i));
      seq.resize((size_t)r.l_seq); qual.resize((size_t)r.l_seq);
      for (int i = 0; i < r.l_seq; ++i) { seq[i] = nt16[(r.seq[i >> 1] >> ((~i & 1) << 2)) & 15]; qual[i] = (char)(r.qual[i] + 33); }
      aux.assign(r.aux, r.aux + r.l_aux);
      bw.write(r.qname, (uint32_t)atol(r.qname), r.flag, r.tid, r.pos, r.mapq, cig, seq, qual, aux);
    }
    if (!br.error().empty()) { fprintf(stderr, "%s\n", br.error().c_str()); return 1; }
    if (!bw.close(&err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    printf("%llu\n", (unsigned long long)bw.written());
    return 0;
  }
  if (mode == "fasta" && argc == 5) {
    RefTable rt; if (!rt.load_sam_header(argv[3], &err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    Genome g; if (!load_fasta(argv[2], rt, g, false, 4, &err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    FILE* f = fopen(argv[4], "wb"); if (!f) return 1;
    const uint64_t nc = g.contig_len.size();
    fwrite(&nc, 8, 1, f); fwrite(&g.n_blocks, 8, 1, f);
    fwrite(g.contig_start.data(), 8, nc, f); fwrite(g.contig_len.data(), 4, nc, f);
    fwrite(g.planes.data(), 8, g.planes.size(), f); fwrite(g.nmask.data(), 8, g.nmask.size(), f);
    fclose(f);
    for (uint32_t i = 1; i <= rt.size(); ++i) printf("%s\n", rt.name(i).c_str());
    return 0;
  }
  if (mode == "hits" && argc == 4) {
    RefTable rt; if (!rt.load_sam_header(argv[3], &err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    std::mutex m; HitStream hs(argv[2], rt, m, 500000);
    std::vector<thb_hit> v;
    for (uint32_t id; (id = hs.next_group_id()) != 0;) {
      v.clear(); hs.next_group(v);
      for (const thb_hit& h : v) printf("%u %u %d %d %u %u %u\n", id, h.ref_id, h.left, h.right, h.read_len, h.edit_dist, h.flags);
    }
    if (!hs.ok()) { fprintf(stderr, "%s\n", hs.error().c_str()); return 1; }
    return 0;
  }
  if (mode == "jhits" && argc == 5) {
    RefTable rt; if (!rt.load_sam_header(argv[3], &err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    std::mutex m; JoinHitStream hs(argv[2], rt, m, atoi(argv[4]) != 0, 500000, 8);
    std::vector<thb_jhit_full> v;
    for (uint32_t id; (id = hs.next_group_id()) != 0;) {
      v.clear(); hs.next_group(v);
      for (const thb_jhit_full& h : v) {
        printf("%u %s %d %u %u %u", id, rt.name(h.ref_id).c_str(), h.left, h.flags, h.mismatches, h.splice_mms);
        for (int k = 0; k < h.n_ops; ++k) printf(" %u:%u", h.ops[k] >> 4, h.ops[k] & 15);
        printf("\n");
      }
    }
    if (!hs.ok()) { fprintf(stderr, "%s\n", hs.error().c_str()); return 1; }
    return 0;
  }
  fprintf(stderr, "usage: thb_host_selftest bam|fasta|hits|jhits ...\n");
  return 2;
}
