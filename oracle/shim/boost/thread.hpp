// Test-infrastructure shim: the reference hot path uses boost::thread only as a plain
// joinable thread (segment_juncs.cpp:4776-4825, long_spanning_reads.cpp:3125-3140).
#pragma once
#include <thread>
namespace boost { using thread = std::thread; }
