// HOST EMULATION SHIM of the one CUB entry point the library uses (see ../../cuda_runtime.h).
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
namespace cub {
struct DeviceRadixSort {
  template <class K>
  static cudaError_t SortKeys(void* tmp, size_t& bytes, const K* in, K* out, int n, int = 0, int = 64, cudaStream_t = 0)
  {
    if (!tmp) { bytes = 16; return cudaSuccess; }
    std::copy(in, in + n, out); std::sort(out, out + n);
    return cudaSuccess;
  }
  template <class K, class V>
  static cudaError_t SortPairs(void* tmp, size_t& bytes, const K* kin, K* kout, const V* vin, V* vout, int n, int = 0, int = 64, cudaStream_t = 0)
  {
    if (!tmp) { bytes = 16; return cudaSuccess; }
    std::vector<int> idx(n); for (int i = 0; i < n; ++i) idx[i] = i;
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return kin[a] < kin[b]; });
    for (int i = 0; i < n; ++i) { kout[i] = kin[idx[i]]; vout[i] = vin[idx[i]]; }
    return cudaSuccess;
  }
};
}
