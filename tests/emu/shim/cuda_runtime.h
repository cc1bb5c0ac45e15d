// cuda_runtime.h -- HOST EMULATION SHIM (development harness, not a product path).
//
// tests/emu compiles tophat_b200/csrc/thb_api.cu + the kernel headers with g++ against this header so that
// the kernels' LOGIC can be debugged in a container without a GPU.  Every warp runs as 32 cooperative fibers
// (ucontext) on one OS thread; full-mask warp collectives (__syncwarp, __shfl*_sync, __ballot_sync) are
// barriers between the fibers, and a lane that reaches a different collective than its siblings aborts the
// run -- which is how warp-divergence bugs show up here instead of as a hang on the GPU box.
// Nothing under tophat_b200/ knows about this build; only tests/test_emu_*.py load it, explicitly.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>
#include <stdio.h>
#include <limits.h>
#include <ucontext.h>
#include <algorithm>
#include <functional>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static      /* blocks and warps run one after the other */
#define THB_EMU 1

struct uint2 { unsigned x, y; };
struct int2 { int x, y; };
struct uint3 { unsigned x, y, z; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
struct __attribute__((aligned(16))) ulonglong2 { unsigned long long x, y; };
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
static inline uint2 make_uint2(unsigned x, unsigned y) { uint2 r = { x, y }; return r; }
static inline int2 make_int2(int x, int y) { int2 r = { x, y }; return r; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }

// ---- runtime API ------------------------------------------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0, cudaDevAttrMultiProcessorCount = 16 };
struct cudaDeviceProp { int major, minor; char name[64]; };
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { memset(p, 0, sizeof *p); p->major = 10; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 1; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (void*)1; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (void*)1; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (void*)1; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
// device memory = host heap, filled with a pattern so that reads of never-written memory stand out
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n ? n : 1); if (!*p) return cudaErrorMemoryAllocation; memset(*p, 0xCD, n); return cudaSuccess; }
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return cudaSuccess; }

// ---- device intrinsics --------------------------------------------------------------------------
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
template <class T> static inline T __ldcs(const T* p) { return *p; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
static inline unsigned long long __brevll(unsigned long long x)
{ unsigned long long r = 0; for (int i = 0; i < 64; ++i) { r = (r << 1) | (x & 1ull); x >>= 1; } return r; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <class T> static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <class T> static inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> static inline T atomicCAS(T* p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }

// ---- the warp executor ------------------------------------------------------------------------
namespace emu {
enum { K_NONE = 0, K_SYNC, K_SHFL, K_BALLOT };
struct Warp {
  ucontext_t main; ucontext_t lane[32]; bool done[32]; int phase[32]; int kind[32];
  unsigned long long buf[2][32]; int cur; const std::function<void()>* fn;
};
inline Warp g_warp;
inline char* g_stacks = nullptr;
constexpr size_t STACK = 512 * 1024;
}
inline uint3 threadIdx, blockIdx;
inline dim3 blockDim, gridDim;
namespace emu {
inline void die(const char* what) { fprintf(stderr, "[thb emu] %s (block %u thread %u)\n", what, blockIdx.x, threadIdx.x); abort(); }
inline void trampoline() { Warp& w = g_warp; const int l = w.cur; (*w.fn)(); w.done[l] = true; w.kind[l] = K_NONE; }
// parks the calling lane at a collective of `kind` carrying v; returns the buffer index holding every lane's value
inline int exchange(int kind, unsigned long long v)
{
  Warp& w = g_warp; const int l = w.cur;
  const int p = w.phase[l] & 1;
  w.buf[p][l] = v; w.phase[l]++; w.kind[l] = kind;
  swapcontext(&w.lane[l], &w.main);
  threadIdx.x = (threadIdx.x & ~31u) | (unsigned)l;
  return p;
}
inline void run_warp(const std::function<void()>& fn, unsigned warp_in_block)
{
  Warp& w = g_warp;
  if (!g_stacks) g_stacks = (char*)malloc(STACK * 32);
  w.fn = &fn;
  for (int l = 0; l < 32; ++l) {
    w.done[l] = false; w.phase[l] = 0; w.kind[l] = K_NONE;
    getcontext(&w.lane[l]);
    w.lane[l].uc_stack.ss_sp = g_stacks + STACK * l; w.lane[l].uc_stack.ss_size = STACK; w.lane[l].uc_link = &w.main;
    makecontext(&w.lane[l], (void (*)())trampoline, 0);
  }
  for (;;) {
    bool any = false;
    for (int l = 0; l < 32; ++l) {
      if (w.done[l]) continue;
      any = true; w.cur = l; threadIdx.x = warp_in_block * 32 + (unsigned)l;
      swapcontext(&w.main, &w.lane[l]);
    }
    if (!any) break;
    // every live lane is now parked at a collective: they must all be at the same one
    int ph = -1, kd = -1; bool live = false, dead = false;
    for (int l = 0; l < 32; ++l) {
      if (w.done[l]) { dead = true; continue; }
      live = true;
      if (ph < 0) { ph = w.phase[l]; kd = w.kind[l]; }
      else if (ph != w.phase[l] || kd != w.kind[l]) die("warp divergence: lanes parked at different collectives");
    }
    if (live && dead) die("a lane exited while others wait at a full-mask collective");
  }
}
inline void launch(dim3 g, dim3 b, const std::function<void()>& fn)
{
  if (b.x % 32) die("block size must be a multiple of 32 in the emulator");
  gridDim = g; blockDim = b;
  for (unsigned bx = 0; bx < g.x; ++bx) {
    blockIdx.x = bx; blockIdx.y = blockIdx.z = 0;
    for (unsigned wi = 0; wi < b.x / 32; ++wi) run_warp(fn, wi);
  }
}
}  // namespace emu

static inline unsigned __activemask() { return 1u << (threadIdx.x & 31u); }   // aggregation degenerates to one lane
static inline void __syncwarp(unsigned mask = 0xffffffffu)
{ if (mask != 0xffffffffu) emu::die("partial-mask __syncwarp"); emu::exchange(emu::K_SYNC, 0); }
template <class T> static inline T __shfl_sync(unsigned mask, T v, int src)
{
  static_assert(sizeof(T) <= 8, "shfl type");
  const unsigned lane = threadIdx.x & 31u;
  if (mask == (1u << lane)) return v;                                         // single-lane group
  if (mask != 0xffffffffu) emu::die("partial-mask __shfl_sync");
  unsigned long long raw = 0; memcpy(&raw, &v, sizeof(T));
  const int p = emu::exchange(emu::K_SHFL, raw);
  T o; memcpy(&o, &emu::g_warp.buf[p][src & 31], sizeof(T)); return o;
}
template <class T> static inline T __shfl_xor_sync(unsigned mask, T v, int x)
{ return __shfl_sync(mask, v, (int)((threadIdx.x & 31u) ^ (unsigned)x)); }
static inline unsigned __ballot_sync(unsigned mask, int pred)
{
  const unsigned lane = threadIdx.x & 31u;
  if (mask == (1u << lane)) return pred ? mask : 0u;
  if (mask != 0xffffffffu) emu::die("partial-mask __ballot_sync");
  const int p = emu::exchange(emu::K_BALLOT, pred ? 1ull : 0ull);
  unsigned r = 0; for (int l = 0; l < 32; ++l) if (emu::g_warp.buf[p][l]) r |= 1u << l;
  return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0u; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == 0xffffffffu; }
template <class T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta)
{ const unsigned lane = threadIdx.x & 31u; const T o = __shfl_sync(mask, v, (int)(lane >= delta ? lane - delta : lane)); return o; }
template <class T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta)
{ const unsigned lane = threadIdx.x & 31u; const T o = __shfl_sync(mask, v, (int)(lane + delta < 32u ? lane + delta : lane)); return o; }
