#include "thb_options.hpp"
#include <getopt.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace thbhost {

namespace {

enum Kind { K_FLAG, K_INT, K_STR, K_IGNORE_ARG, K_IGNORE_FLAG, K_LIBTYPE, K_STRLIST };

struct Spec {
  const char* name; char short_opt; Kind kind; long lower;
  int32_t thb_params::* pfield; int Options::* ifield; bool Options::* bfield; std::string Options::* sfield;
};

#define PI(f) &thb_params::f, nullptr, nullptr, nullptr
#define OI(f) nullptr, &Options::f, nullptr, nullptr
#define OB(f) nullptr, nullptr, &Options::f, nullptr
#define OS(f) nullptr, nullptr, nullptr, &Options::f
#define NONE nullptr, nullptr, nullptr, nullptr

// One row per option of the reference table (name, short alias, kind, lower bound, destination).
const Spec SPECS[] = {
  {"fasta", 0, K_IGNORE_FLAG, 0, NONE}, {"fastq", 0, K_IGNORE_FLAG, 0, NONE},
  {"min-anchor", 0, K_INT, 3, PI(min_anchor_len)}, {"sam-header", 0, K_STR, 0, OS(sam_header)},
  {"rg-id", 0, K_IGNORE_ARG, 0, NONE}, {"splice-mismatches", 0, K_IGNORE_ARG, 0, NONE},
  {"verbose", 0, K_FLAG, 0, OB(verbose)},
  {"inner-dist-mean", 0, K_INT, -1024, PI(inner_dist_mean)}, {"inner-dist-std-dev", 0, K_INT, 0, PI(inner_dist_std_dev)},
  {"output-dir", 0, K_STR, 0, OS(output_dir)}, {"gene-filter", 0, K_IGNORE_ARG, 0, NONE},
  {"gtf-annotations", 0, K_IGNORE_ARG, 0, NONE}, {"max-multihits", 0, K_INT, 1, OI(max_multihits)},
  {"suppress-hits", 0, K_IGNORE_FLAG, 0, NONE}, {"max-seg-multihits", 0, K_INT, 1, PI(max_seg_multihits)},
  {"no-closure-search", 0, K_FLAG, 0, OB(no_closure_search)}, {"no-coverage-search", 0, K_FLAG, 0, OB(no_coverage_search)},
  {"no-microexon-search", 0, K_FLAG, 0, OB(no_microexon_search)},
  {"segment-length", 0, K_INT, 4, PI(segment_length)}, {"segment-mismatches", 0, K_INT, 0, PI(segment_mismatches)},
  {"read-mismatches", 'N', K_INT, 0, PI(read_mismatches)}, {"read-gap-length", 0, K_INT, 0, PI(read_gap_length)},
  {"read-edit-dist", 0, K_INT, 0, PI(read_edit_dist)}, {"read-realign-edit-dist", 0, K_IGNORE_ARG, 0, NONE},
  {"min-closure-exon", 0, K_IGNORE_ARG, 0, NONE}, {"min-closure-intron", 0, K_IGNORE_ARG, 0, NONE},
  {"max-closure-intron", 0, K_IGNORE_ARG, 0, NONE}, {"min-coverage-intron", 0, K_IGNORE_ARG, 0, NONE},
  {"max-coverage-intron", 0, K_IGNORE_ARG, 0, NONE},
  {"min-segment-intron", 0, K_INT, 1, PI(min_segment_intron_length)}, {"max-segment-intron", 0, K_INT, 1, PI(max_segment_intron_length)},
  {"min-report-intron", 0, K_INT, 1, PI(min_report_intron_length)}, {"max-report-intron", 0, K_INT, 1, PI(max_report_intron_length)},
  {"min-isoform-fraction", 0, K_IGNORE_ARG, 0, NONE}, {"ium-reads", 0, K_STR, 0, OS(ium_reads)},
  {"butterfly-search", 0, K_FLAG, 0, OB(butterfly_search)}, {"solexa-quals", 0, K_IGNORE_FLAG, 0, NONE},
  {"phred64-quals", 0, K_IGNORE_FLAG, 0, NONE}, {"quals", 'Q', K_IGNORE_FLAG, 0, NONE}, {"integer-quals", 0, K_IGNORE_FLAG, 0, NONE},
  {"color", 'C', K_FLAG, 0, OB(color)}, {"library-type", 0, K_LIBTYPE, 0, NONE},
  {"max-deletion-length", 0, K_INT, 0, PI(max_deletion_length)}, {"max-insertion-length", 0, K_INT, 0, PI(max_insertion_length)},
  {"num-threads", 'p', K_INT, 1, OI(num_threads)}, {"zpacker", 'z', K_STR, 0, OS(zpacker)},
  {"samtools", 0, K_IGNORE_ARG, 0, NONE}, {"aux-outfile", 0, K_IGNORE_ARG, 0, NONE}, {"outfile", 'w', K_IGNORE_ARG, 0, NONE},
  {"index-outfile", 0, K_IGNORE_ARG, 0, NONE}, {"gtf-juncs", 0, K_IGNORE_ARG, 0, NONE}, {"flt-reads", 0, K_IGNORE_ARG, 0, NONE},
  {"flt-hits", 0, K_IGNORE_ARG, 0, NONE}, {"flt-side", 0, K_IGNORE_ARG, 0, NONE},
  {"report-secondary-alignments", 0, K_IGNORE_FLAG, 0, NONE}, {"report-discordant-pair-alignments", 0, K_IGNORE_FLAG, 0, NONE},
  {"report-mixed-alignments", 0, K_IGNORE_FLAG, 0, NONE},
  {"fusion-search", 0, K_INT /* flag -> 1 */, 0, PI(fusion_search)},
  {"fusion-anchor-length", 0, K_INT, 10, PI(fusion_anchor_length)}, {"fusion-min-dist", 0, K_INT, 0, PI(fusion_min_dist)},
  {"fusion-read-mismatches", 0, K_INT, 0, OI(fusion_read_mismatches)}, {"fusion-multireads", 0, K_INT, 1, OI(fusion_multireads)},
  {"fusion-multipairs", 0, K_INT, 1, OI(fusion_multipairs)}, {"fusion-ignore-chromosomes", 0, K_STRLIST, 0, NONE},
  {"fusion-do-not-resolve-conflicts", 0, K_FLAG, 0, OB(fusion_do_not_resolve_conflicts)},
  {"bowtie1", 0, K_IGNORE_FLAG /* handled below */, 0, NONE},
  {"bowtie2-min-score", 'W', K_IGNORE_ARG, 0, NONE},
  {"bowtie2-max-penalty", 0, K_INT, 0, PI(bowtie2_max_penalty)}, {"bowtie2-min-penalty", 0, K_INT, 0, PI(bowtie2_min_penalty)},
  {"bowtie2-penalty-for-N", 0, K_INT, 0, PI(bowtie2_penalty_for_N)}, {"bowtie2-read-gap-open", 0, K_INT, 0, PI(bowtie2_read_gap_open)},
  {"bowtie2-read-gap-cont", 0, K_INT, 0, PI(bowtie2_read_gap_cont)}, {"bowtie2-ref-gap-open", 0, K_INT, 0, PI(bowtie2_ref_gap_open)},
  {"bowtie2-ref-gap-cont", 0, K_INT, 0, PI(bowtie2_ref_gap_cont)},
};
const int NSPEC = (int)(sizeof(SPECS) / sizeof(SPECS[0]));

int libtype_from_name(const char* s)
{
  const char* names[] = {"", "fr-unstranded", "fr-firststrand", "fr-secondstrand", "ff-unstranded", "ff-firststrand", "ff-secondstrand"};
  for (int i = 1; i < 7; ++i) if (strcmp(s, names[i]) == 0) return i;
  return 0;      // unknown names leave LIBRARY_TYPE_NONE, like the reference's if/else chain
}

}  // namespace

std::vector<std::string> split_list(const std::string& s, char sep)
{
  std::vector<std::string> out; std::string cur;
  for (char c : s) { if (c == sep) { if (!cur.empty()) out.push_back(cur); cur.clear(); } else cur.push_back(c); }
  if (!cur.empty()) out.push_back(cur);
  return out;
}

int parse_options(int argc, char** argv, Options& o, void (*usage)())
{
  thb_params_default(&o.p);
  std::vector<option> lo; std::string shorts;
  for (int i = 0; i < NSPEC; ++i) {
    const Spec& s = SPECS[i];
    bool is_flag = s.kind == K_FLAG || s.kind == K_IGNORE_FLAG || !strcmp(s.name, "fusion-search");
    option op; op.name = s.name; op.has_arg = is_flag ? no_argument : required_argument; op.flag = nullptr; op.val = 1000 + i;
    lo.push_back(op);
    if (s.short_opt) { shorts.push_back(s.short_opt); if (!is_flag) shorts.push_back(':'); }
  }
  option end; memset(&end, 0, sizeof end); lo.push_back(end);
  optind = 1;
  int c, idx = 0;
  while ((c = getopt_long(argc, argv, shorts.c_str(), lo.data(), &idx)) != -1) {
    int si = -1;
    if (c >= 1000) si = c - 1000;
    else for (int i = 0; i < NSPEC; ++i) if (SPECS[i].short_opt && SPECS[i].short_opt == c) { si = i; break; }
    if (si < 0) { usage(); return 1; }
    const Spec& s = SPECS[si];
    if (!strcmp(s.name, "fusion-search")) { o.p.fusion_search = 1; continue; }
    if (!strcmp(s.name, "bowtie1")) { o.p.bowtie2 = 0; continue; }
    switch (s.kind) {
      case K_FLAG: o.*(s.bfield) = true; break;
      case K_INT: {
        char* endp = nullptr; long v = strtol(optarg, &endp, 10);
        if (v < s.lower) { fprintf(stderr, "--%s arg must be at least %ld\n", s.name, s.lower); usage(); return 1; }
        if (s.pfield) o.p.*(s.pfield) = (int32_t)v; else o.*(s.ifield) = (int)v;
        break; }
      case K_STR: o.*(s.sfield) = optarg; break;
      case K_LIBTYPE: o.library_type_name = optarg; o.p.library_type = libtype_from_name(optarg); break;
      case K_STRLIST: o.fusion_ignore_chromosomes = split_list(optarg); break;
      case K_IGNORE_ARG: case K_IGNORE_FLAG: break;
    }
  }
  for (int i = optind; i < argc; ++i) o.positional.push_back(argv[i]);
  return 0;
}

}  // namespace thbhost
