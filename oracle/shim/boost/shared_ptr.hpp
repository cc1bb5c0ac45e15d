// Test-infrastructure shim: boost::shared_ptr / make_shared -> std (reads.h:16,147; reads.cpp:534).
#pragma once
#include <memory>
namespace boost { using std::shared_ptr; using std::make_shared; }
