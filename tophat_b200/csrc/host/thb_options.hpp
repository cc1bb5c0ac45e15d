// thb_options.hpp -- the option surface shared by the TopHat stage binaries.
//
// tophat.py passes the same ~30 flags (TopHatParams.cmd, tophat.py:824-900) to every stage program, so a
// drop-in replacement has to accept the whole table of common.cpp:265-422 even where it ignores the value.
// Defaults are the reference binaries' (common.cpp:79-180); lower bounds follow parse_options
// (common.cpp:459-721): a violated bound or an unknown option prints the usage text and fails.
#pragma once
#include <string>
#include <vector>
#include "tophat_b200.h"

namespace thbhost {

struct Options {
  thb_params p;                       // everything the kernels consume
  std::string sam_header, ium_reads, output_dir = "./", zpacker, library_type_name;
  int num_threads = 1;
  bool no_closure_search = false, no_coverage_search = false, no_microexon_search = false, butterfly_search = false;
  bool color = false, fusion_do_not_resolve_conflicts = false, verbose = false;
  int max_multihits = 20, fusion_read_mismatches = 2, fusion_multireads = 2, fusion_multipairs = 2;
  std::vector<std::string> fusion_ignore_chromosomes;
  std::vector<std::string> positional;
};

// Returns 0 on success; on error prints the message (and the usage via `usage`) to stderr and returns 1.
int parse_options(int argc, char** argv, Options& o, void (*usage)());

std::vector<std::string> split_list(const std::string& s, char sep = ',');

}  // namespace thbhost
