// bitplanes.cuh -- bit-plane sequence primitives shared by every kernel of the hot path.
//
// A sequence is held as three bit planes (bit j of a plane word = base j): plane0/plane1 carry the
// 2-bit code (A=0 C=1 G=2 T=3) and planeN marks 'N' (code bits are 0 under an N).  Mismatch
// counting between a read slice and a reference window then is XOR + OR + popcount on 64-bit words
// (the per-base loops of segment_juncs.cpp:2191-2220, 2406-2434, 2953-2970 collapse to a handful of
// integer instructions), and a reverse complement is a bit reversal + NOT.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace thb {

struct P2 { uint64_t p0, p1; };          // up to 64 bases, code planes only
struct P3 { uint64_t p0, p1, pn; };      // up to 64 bases, code planes + N plane

__device__ __forceinline__ uint64_t maskn(int n) { return n >= 64 ? ~0ull : ((1ull << n) - 1ull); }

// 64 bits starting at bit `sh` of the 128-bit value hi:lo
__device__ __forceinline__ uint64_t shr128(uint64_t lo, uint64_t hi, int sh)
{
  return sh ? ((lo >> sh) | (hi << (64 - sh))) : lo;
}

// ---- reference image ---------------------------------------------------------------------------
struct RefView {
  const ulonglong2* planes;      // [n_blocks] {plane0, plane1} of 64 bases
  const uint64_t*   nmask;       // [n_blocks]
  const uint64_t*   contig_start;
  const uint32_t*   contig_len;
  uint32_t          n_contigs;
};

// Code planes of bases [g, g+n), n <= 64.  An 'N' reads as 'A' (code 0): the Dna5 -> Dna
// conversion of segment_juncs.cpp:2157 / 2499.
__device__ __forceinline__ P2 ref_fetch2(const RefView& r, uint64_t g, int n)
{
  const uint64_t b = g >> 6; const int sh = (int)(g & 63);
  const ulonglong2 lo = __ldg(r.planes + b);
  P2 o;
  if (sh + n > 64) {
    const ulonglong2 hi = __ldg(r.planes + b + 1);
    o.p0 = shr128(lo.x, hi.x, sh); o.p1 = shr128(lo.y, hi.y, sh);
  } else { o.p0 = lo.x >> sh; o.p1 = lo.y >> sh; }
  const uint64_t m = maskn(n);
  o.p0 &= m; o.p1 &= m;
  return o;
}
// Same with the N plane (Dna5 view: segment_juncs.cpp:2582-2583, 3428-3445, 2649-2672).
__device__ __forceinline__ P3 ref_fetch3(const RefView& r, uint64_t g, int n)
{
  const uint64_t b = g >> 6; const int sh = (int)(g & 63);
  const ulonglong2 lo = __ldg(r.planes + b);
  const uint64_t nlo = __ldg(r.nmask + b);
  P3 o;
  if (sh + n > 64) {
    const ulonglong2 hi = __ldg(r.planes + b + 1);
    const uint64_t nhi = __ldg(r.nmask + b + 1);
    o.p0 = shr128(lo.x, hi.x, sh); o.p1 = shr128(lo.y, hi.y, sh); o.pn = shr128(nlo, nhi, sh);
  } else { o.p0 = lo.x >> sh; o.p1 = lo.y >> sh; o.pn = nlo >> sh; }
  const uint64_t m = maskn(n);
  o.p0 &= m; o.p1 &= m; o.pn &= m;
  return o;
}

// ---- reads ------------------------------------------------------------------------------------
// A read is 3*rw words: plane0[rw] | plane1[rw] | planeN[rw].  Bases [start, start+n), n <= 64.
__device__ __forceinline__ uint64_t plane_slice(const uint64_t* __restrict__ w, int rw, int start, int n)
{
  const int i = start >> 6, sh = start & 63;
  const uint64_t lo = w[i];
  const uint64_t hi = (sh + n > 64 && i + 1 < rw) ? w[i + 1] : 0ull;
  return shr128(lo, hi, sh) & maskn(n);
}
__device__ __forceinline__ P3 read_slice(const uint64_t* __restrict__ rd, int rw, int start, int n)
{
  P3 o;
  o.p0 = plane_slice(rd, rw, start, n);
  o.p1 = plane_slice(rd + rw, rw, start, n);
  o.pn = plane_slice(rd + 2 * rw, rw, start, n);
  return o;
}

// Reverse complement of an n-base slice; non-ACGT stays N (reads.cpp:191-207).
__device__ __forceinline__ P3 revcomp(const P3& a, int n)
{
  const int sh = 64 - n;
  P3 o;
  o.pn = __brevll(a.pn) >> sh;
  const uint64_t keep = maskn(n) & ~o.pn;
  o.p0 = (~(__brevll(a.p0) >> sh)) & keep;
  o.p1 = (~(__brevll(a.p1) >> sh)) & keep;
  return o;
}

// bit j set where base j of X equals letter c (0..3)
__device__ __forceinline__ uint64_t eq_letter(const P2& x, int c)
{
  return ((c & 1) ? x.p0 : ~x.p0) & ((c & 2) ? x.p1 : ~x.p1);
}

// position of the k-th (0-based) lowest set bit; caller guarantees it exists
__device__ __forceinline__ int nth_lowest(uint64_t m, int k)
{
  for (int i = 0; i < k; ++i) m &= m - 1;
  return __ffsll((long long)m) - 1;
}

}  // namespace thb
