// thb_api.cu -- implementation of the C ABI declared in include/tophat_b200.h.
//
// Owns the device context: reference image, result sets, staging buffers, streams.  The host
// batch path (thb_segjuncs_submit) is a two-stream, double-buffered pipeline: chunk c+1 is copied
// host->device on the copy stream while the scan kernel of chunk c runs on the compute stream.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdarg>
#include <string>
#include <vector>
#include <algorithm>
#include <chrono>
#include <dlfcn.h>
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include "../../include/tophat_b200.h"
#include "segjuncs_kernel.cuh"
#include "scan_tile_kernel.cuh"
#include "join_kernel.cuh"
#include "join_tile_kernel.cuh"
#include "fusion_join_kernel.cuh"
#include "fusion_kernel.cuh"
#include "flank_kernel.cuh"

using namespace thb;

namespace {

char g_create_error[512] = "";

struct DevBuf {
  void* p = nullptr; size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = n + n / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// grow-only page-locked host array for result records (device->host copies at full PCIe speed, no zero fill)
template <class T> struct PinnedVec {
  T* p = nullptr; size_t n = 0, cap = 0;
  cudaError_t resize(size_t m) {
    if (m > cap) {
      if (p) cudaFreeHost(p);
      p = nullptr; cap = 0;
      const size_t want = m + m / 4 + 1024;
      cudaError_t e = cudaHostAlloc((void**)&p, want * sizeof(T), cudaHostAllocDefault);
      if (e != cudaSuccess) { n = 0; return e; }
      cap = want;
    }
    n = m; return cudaSuccess;
  }
  void clear() { n = 0; }
  size_t size() const { return n; }
  T* data() { return p; }
  void release() { if (p) cudaFreeHost(p); p = nullptr; n = cap = 0; }
};

struct Staging { DevBuf bundles, seg_count, reads, hits, partner; cudaEvent_t copied = nullptr, consumed = nullptr; bool used = false; };
// join pipeline stage: input staging + the chunk's result buffer (its device->host copy overlaps the next chunk's kernels)
struct JStage { DevBuf bundles, seg_count, reads, hits, ops, out; uint64_t cap_out = 0; cudaEvent_t copied = nullptr, out_free = nullptr; bool out_busy = false; };

// dynamically bound NCCL (the library is only needed for the multi-GPU exchange)
struct NcclUid { char b[128]; };          // ncclUniqueId is passed BY VALUE to ncclCommInitRank
struct Nccl {
  typedef NcclUid Uid;
  void* h = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, NcclUid, int) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

// junction-flank matcher (thb_flank_*)
struct FlankState {
  DevBuf desc, seq, base, keys, vals, keys2, vals2, start, buckets, reads, hkeys, hmm, hkeys2, hmm2, per_seg, out, scalars, cdesc, jout;
  std::vector<thb_flank_contig> contigs; PinnedVec<thb_flank_hit> hits; PinnedVec<thb_jhit_full> jhits;
  FlankBatchView last_bv{}; uint64_t last_n = 0; bool have_last = false;
  FlankIndexParams ip{}; int cw = 1; bool begun = false; uint64_t n_contigs = 0, n_entries = 0, cap_hits = 0;
  int min_seg_len = 0, max_seg_len = 0; thb_flank_timing timing{};
  void release() { for (DevBuf* b : { &desc, &seq, &base, &keys, &vals, &keys2, &vals2, &start, &buckets, &reads, &hkeys, &hmm, &hkeys2, &hmm2, &per_seg, &out, &scalars, &cdesc, &jout }) b->release(); hits.release(); jhits.release(); }
};

}  // namespace

struct thb_ctx {
  int device = 0;
  cudaStream_t compute = nullptr, copy = nullptr, d2h = nullptr;
  cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_c = nullptr, ev_d = nullptr;
  std::string err;
  // reference
  DevBuf d_planes, d_nmask, d_cstart, d_clen;
  std::vector<uint64_t> h_cstart; std::vector<uint32_t> h_clen;
  RefView ref{}; bool have_ref = false;
  // parameters
  thb_params params{}; SegParams sp{}; bool begun = false;
  // result sets
  DevBuf d_juncs, d_dels; uint64_t cap_juncs = 0, cap_dels = 0;
  DevBuf d_ins; uint64_t cap_ins = 0;
  // --fusion-search: candidate pair queue, record append buffer, ignored contigs
  DevBuf d_fus, q_fus, d_fus_ignore; uint64_t cap_fus = 0, cap_fustask = 0; FusionParams fp{};
  unsigned long long* d_fus_count = nullptr; unsigned long long* d_fustask_count = nullptr; unsigned long long h_fus_count = 0;
  std::vector<FusRec> h_fusrec;
  DevBuf d_scalars;                 // [0..7] counters (u64), then ins_count(u64), then flags (u32 x4)
  unsigned long long* d_counters = nullptr; unsigned long long* d_ins_count = nullptr;
  unsigned int* d_ovf_juncs = nullptr; unsigned int* d_ovf_dels = nullptr; unsigned int* d_err = nullptr;
  DevBuf d_keys, d_keys_sorted, d_cub_tmp, d_decoded, d_count;
  // device-resident results of the last finish (thb_join_begin_resident, asynchronous download of thb_segjuncs_finish_resident)
  DevBuf d_decoded_dels, d_skeys_j, d_skeys_d; uint64_t res_n[3] = {0, 0, 0}; bool have_resident = false, fetch_pending = false;
  cudaEvent_t ev_sets = nullptr;
  // task queues of the scan phase
  DevBuf ag_send, ag_recv;
  const uint64_t* ag_keys[2] = {nullptr, nullptr}; uint64_t ag_n[2] = {0, 0};     // after the all-gather: every rank's junction / deletion keys (HS_EMPTY padded)
  DevBuf q_win, q_indel, q_rescue, q_rescue_out, q_rbundle, q_bstate, q_owner;
  uint64_t cap_win = 0, cap_indel = 0;
  unsigned long long* d_qcounts = nullptr; unsigned int* d_qovf = nullptr;
  cudaEvent_t kev[9] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool kev_pending = false;
  int sms = 148;
  Staging stage[2];
  // host results
  PinnedVec<thb_junction> h_juncs, h_dels; PinnedVec<thb_insertion> h_ins; std::vector<thb_fusion> h_fus;
  DevBuf d_ins_a, d_ins_b, d_ins_c, d_ins_d, d_ins_out;     // scratch of the insertion reduction
  // accounting
  thb_timing timing{}; uint64_t n_bundles_total = 0, n_hits_total = 0, n_partner_total = 0;
  uint64_t n_ins_out = 0, n_del_out = 0;
  unsigned long long h_ins_count = 0;   // host mirror of *d_ins_count after the last completed launch
  uint32_t own_launches = 0;        // every kernel of this library launched since thb_segjuncs_begin
  // long_spanning_reads join
  DevBuf j_idx; uint64_t j_nbuckets = 0; bool j_use_idx = false; int j_shift = 6;
  DevBuf j_iidx; uint64_t j_nibuckets = 0; bool j_use_iidx = false; int j_ishift = 6;
  DevBuf j_juncs, j_ins, j_bundles, j_segc, j_reads, j_hits, j_out, j_chain; uint64_t j_cap_chain = 0, j_cap_out = 0, j_n_juncs = 0, j_n_ins = 0;
  DevBuf j_fus; uint64_t j_n_fus = 0;              // --fusion-search: the fusion set of the join (FusionKey records)
  JoinParams jp{}; bool join_begun = false; thb_join_timing jtiming{}; unsigned long long j_last_n = 0;
  JStage jstage[2];
  thb_joined* h_joined = nullptr; uint64_t h_joined_cap = 0;      // page-locked result buffer, grow-only
  FlankState fl;
  // nccl
  Nccl nccl; void* comm = nullptr; int rank = 0, world = 1;
};

namespace {

int fail(thb_ctx* c, int code, const char* fmt, ...)
{
  char buf[512]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  if (c) c->err = buf; else snprintf(g_create_error, sizeof g_create_error, "%s", buf);
  return code;
}
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(ctx, THB_ECUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

// THB_TINY_CAPS=1 (test knob): every growable structure starts tiny, so that the overflow -> grow -> repeat paths run on small inputs
bool tiny_caps() { static const bool v = getenv("THB_TINY_CAPS") != nullptr; return v; }
// Measurement knobs.  THB_SCAN_LEGACY=1: round 1's bundle_kernel + hit_kernel instead of the TMA-staged scan tile kernel.
// THB_JOIN_TILE=1: the TMA-staged join tile kernel instead of the queue kernels (chain_enum + merge_simple + merge_abut) -- measured
// on B200 it is instruction-issue bound at 13 of 32 lanes and 11 % slower than the queue kernels (profiles/r2c_*), so it is not the default.
bool join_legacy() { static const bool v = getenv("THB_JOIN_TILE") == nullptr || getenv("THB_JOIN_LEGACY") != nullptr; return v; }
bool scan_legacy() { static const bool v = getenv("THB_SCAN_LEGACY") != nullptr; return v; }

HashSet make_set(DevBuf& b, uint64_t cap, unsigned int* ovf) { HashSet h; h.slots = (uint64_t*)b.p; h.mask = cap - 1; h.overflow = ovf; return h; }

int grid_for(uint64_t n, int block) {
  int dev = 0, sms = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  uint64_t g = (n + block - 1) / block; uint64_t cap = (uint64_t)sms * 16;
  if (g > cap) g = cap; if (g < 1) g = 1; return (int)g;
}

int alloc_set(thb_ctx* ctx, DevBuf& b, uint64_t cap)
{
  CU(b.reserve(cap * sizeof(uint64_t)));
  hs_clear_kernel<<<grid_for(cap, 256), 256, 0, ctx->compute>>>((uint64_t*)b.p, cap);
  CU(cudaGetLastError()); ctx->own_launches++;
  return THB_OK;
}

// doubles a hash set, re-inserting its content
int grow_set(thb_ctx* ctx, DevBuf& b, uint64_t& cap, unsigned int* ovf)
{
  DevBuf nb; uint64_t ncap = cap * 2;
  CU(nb.reserve(ncap * sizeof(uint64_t)));
  hs_clear_kernel<<<grid_for(ncap, 256), 256, 0, ctx->compute>>>((uint64_t*)nb.p, ncap); ctx->own_launches += 2;
  CU(cudaMemsetAsync(ovf, 0, sizeof(unsigned int), ctx->compute));
  HashSet dst; dst.slots = (uint64_t*)nb.p; dst.mask = ncap - 1; dst.overflow = ovf;
  hs_rehash_kernel<<<grid_for(cap, 256), 256, 0, ctx->compute>>>((const uint64_t*)b.p, cap, dst);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(ctx->compute));
  b.release(); b = nb; cap = ncap;
  return THB_OK;
}

SegOutputs outputs(thb_ctx* ctx)
{
  SegOutputs o;
  o.juncs = make_set(ctx->d_juncs, ctx->cap_juncs, ctx->d_ovf_juncs);
  o.dels = make_set(ctx->d_dels, ctx->cap_dels, ctx->d_ovf_dels);
  o.ins = (InsRec*)ctx->d_ins.p; o.ins_count = ctx->d_ins_count; o.ins_cap = ctx->cap_ins;
  o.counters = ctx->d_counters; o.err = ctx->d_err;
  return o;
}

struct Flags { unsigned int ovf_juncs, ovf_dels, err, qovf; };

int read_state(thb_ctx* ctx, Flags* f, unsigned long long* ins_count, unsigned long long* fus_counts2)
{
  CU(cudaMemcpyAsync(f, ctx->d_ovf_juncs, sizeof(Flags), cudaMemcpyDeviceToHost, ctx->compute));
  CU(cudaMemcpyAsync(ins_count, ctx->d_ins_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->compute));
  CU(cudaMemcpyAsync(fus_counts2, ctx->d_fus_count, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->compute));
  CU(cudaStreamSynchronize(ctx->compute));
  return THB_OK;
}

// The scan phase of one bundle range: K1 enumerate -> K2 rescue -> K3 rescued windows -> K4 window scan -> K5 indel.
// `n_partner` = partner hits present in this launch (upper bound of the rescue queue).
template <int NSMAX>
void launch_phase(thb_ctx* ctx, const BatchView& bv, const Queues& q, const SegOutputs& o, uint64_t n_hits)
{
  const int g_b = grid_for(bv.n_bundles, 256), g_h = grid_for(n_hits, 256), g_t = ctx->sms * 8;
  uint32_t* bstate = (uint32_t*)ctx->q_bstate.p; uint32_t* owner = (uint32_t*)ctx->q_owner.p;
  cudaEventRecord(ctx->kev[0], ctx->compute);
  if (scan_legacy()) {
    bundle_kernel<NSMAX><<<g_b, 256, 0, ctx->compute>>>(ctx->ref, ctx->sp, bv, q, bstate, owner);
    cudaEventRecord(ctx->kev[1], ctx->compute);
    if (n_hits) hit_kernel<NSMAX><<<g_h, 256, 0, ctx->compute>>>(ctx->ref, ctx->sp, bv, q, bstate, owner, n_hits, o);
  } else {
    // one pass over the batch: TMA-staged tiles of 32 bundles, bundle bookkeeping + per-hit task construction (scan_tile_kernel.cuh)
    const uint32_t n_tiles = (bv.n_bundles + 31u) / 32u;
    const int grid = (int)std::min<uint64_t>((n_tiles + ST_WARPS - 1) / ST_WARPS, (uint64_t)ctx->sms * 64);
    // residency chosen by measurement on B200 (hg38-sized workload, 6.25 M pairs: 1.17 ms at 6 CTAs / 80 registers vs 1.40 ms at 4 / 115;
    // THB_SCAN_MINB=4 selects the other build): the kernel is issue-bound, more resident warps fill more issue slots
    static const int minb = [] { const char* e = getenv("THB_SCAN_MINB"); return e ? atoi(e) : 6; }();
    if (minb >= 6) scan_tile_kernel<NSMAX, 6><<<grid, ST_WARPS * 32, 0, ctx->compute>>>(ctx->ref, ctx->sp, bv, q, bstate, owner, o);
    else scan_tile_kernel<NSMAX, 4><<<grid, ST_WARPS * 32, 0, ctx->compute>>>(ctx->ref, ctx->sp, bv, q, bstate, owner, o);
    cudaEventRecord(ctx->kev[1], ctx->compute);
  }
  cudaEventRecord(ctx->kev[2], ctx->compute);
  rescue_kernel<<<g_t, 256, 0, ctx->compute>>>(ctx->ref, ctx->sp, bv, q);
  cudaEventRecord(ctx->kev[3], ctx->compute);
  rescued_windows_kernel<NSMAX><<<g_t, 128, 0, ctx->compute>>>(ctx->ref, ctx->sp, bv, q, o);
  cudaEventRecord(ctx->kev[4], ctx->compute);
  window_scan_kernel<<<g_t, 256, 0, ctx->compute>>>(ctx->ref, q, o);
  cudaEventRecord(ctx->kev[5], ctx->compute);
  indel_kernel<<<g_t, 256, 0, ctx->compute>>>(ctx->ref, q, o);
  cudaEventRecord(ctx->kev[6], ctx->compute);
  if (ctx->sp.fusion_search) {
    FusionQueues fq; fq.tasks = (FusionTask*)ctx->q_fus.p; fq.cap = ctx->cap_fustask; fq.count = ctx->d_fustask_count;
    fq.rec = (FusRec*)ctx->d_fus.p; fq.rec_cap = ctx->cap_fus; fq.rec_count = ctx->d_fus_count; fq.err = ctx->d_err; fq.counters = ctx->d_counters;
    fusion_enum_kernel<NSMAX><<<grid_for(bv.n_bundles, 128), 128, 0, ctx->compute>>>(ctx->ref, ctx->sp, ctx->fp, bv, q, fq, bstate);
    cudaEventRecord(ctx->kev[7], ctx->compute);
    fusion_detect_kernel<<<g_t, 128, 0, ctx->compute>>>(ctx->ref, ctx->fp, bv, fq);
    cudaEventRecord(ctx->kev[8], ctx->compute);
  }
}

int launch_scan(thb_ctx* ctx, const BatchView& bv, uint64_t n_partner, uint64_t n_hits)
{
  if (bv.n_bundles == 0) return THB_OK;
  if (tiny_caps()) { ctx->cap_win = std::max<uint64_t>(ctx->cap_win, 64); ctx->cap_indel = std::max<uint64_t>(ctx->cap_indel, 64); }
  else {
    ctx->cap_win = std::max<uint64_t>(ctx->cap_win, std::max<uint64_t>(2ull * bv.n_bundles, 1u << 16));
    ctx->cap_indel = std::max<uint64_t>(ctx->cap_indel, std::max<uint64_t>(bv.n_bundles / 2, 1u << 16));
  }
  CU(ctx->q_win.reserve(ctx->cap_win * sizeof(WindowTask))); CU(ctx->q_indel.reserve(ctx->cap_indel * sizeof(IndelTask)));
  CU(ctx->q_rescue.reserve((n_partner + 1) * sizeof(uint2))); CU(ctx->q_rescue_out.reserve((n_partner + 1) * sizeof(int2)));
  CU(ctx->q_rbundle.reserve((size_t)bv.n_bundles * sizeof(uint32_t)));
  CU(ctx->q_bstate.reserve((size_t)bv.n_bundles * sizeof(uint32_t))); CU(ctx->q_owner.reserve((size_t)(n_hits + 1) * sizeof(uint32_t)));
  Queues q;
  q.win = (WindowTask*)ctx->q_win.p; q.cap_win = ctx->cap_win; q.indel = (IndelTask*)ctx->q_indel.p; q.cap_indel = ctx->cap_indel;
  q.rescue = (uint2*)ctx->q_rescue.p; q.rescue_out = (int2*)ctx->q_rescue_out.p; q.rbundle = (uint32_t*)ctx->q_rbundle.p;
  q.counts = ctx->d_qcounts; q.overflow = ctx->d_qovf;
  CU(cudaMemsetAsync(ctx->d_qcounts, 0, 4 * sizeof(unsigned long long), ctx->compute));
  if (ctx->sp.fusion_search) {
    ctx->cap_fustask = std::max<uint64_t>(ctx->cap_fustask, tiny_caps() ? 64 : std::max<uint64_t>(bv.n_bundles, 1u << 16));
    CU(ctx->q_fus.reserve(ctx->cap_fustask * sizeof(FusionTask)));
    CU(cudaMemsetAsync(ctx->d_fustask_count, 0, sizeof(unsigned long long), ctx->compute));
  }
  const SegOutputs o = outputs(ctx);
  if (bv.n_segs <= 4) launch_phase<4>(ctx, bv, q, o, n_hits);
  else if (bv.n_segs <= 8) launch_phase<8>(ctx, bv, q, o, n_hits);
  else launch_phase<THB_MAX_SEGS>(ctx, bv, q, o, n_hits);
  CU(cudaGetLastError());
  ctx->kev_pending = true;
  ctx->timing.kernel_launches++; ctx->own_launches += (ctx->sp.fusion_search ? 8 : 6) - (scan_legacy() ? 0 : 1);
  return THB_OK;
}

// adds the per-kernel CUDA-event times of the last launch_scan (the stream must be idle)
void collect_kernel_times(thb_ctx* ctx, float* total)
{
  *total = 0.f;
  if (!ctx->kev_pending) return;
  ctx->kev_pending = false;
  float* dst[8] = { &ctx->timing.bundle_ms, &ctx->timing.hit_ms, &ctx->timing.rescue_ms, &ctx->timing.rescued_windows_ms, &ctx->timing.window_scan_ms, &ctx->timing.indel_ms,
                    &ctx->timing.fusion_enum_ms, &ctx->timing.fusion_detect_ms };
  for (int k = 0; k < (ctx->sp.fusion_search ? 8 : 6); ++k) { float ms = 0.f; if (cudaEventElapsedTime(&ms, ctx->kev[k], ctx->kev[k + 1]) == cudaSuccess) { *dst[k] += ms; *total += ms; } }
}

int validate_batch(thb_ctx* ctx, const thb_segjuncs_batch* b)
{
  if (!ctx->have_ref) return fail(ctx, THB_ESTATE, "no reference image uploaded");
  if (!ctx->begun) return fail(ctx, THB_ESTATE, "thb_segjuncs_begin not called");
  if (ctx->ag_keys[0]) return fail(ctx, THB_ESTATE, "batch submitted after thb_segjuncs_allgather (the exchange closes the pass)");
  if (b->n_segs < 1 || b->n_segs > (uint32_t)THB_MAX_SEGS) return fail(ctx, THB_EUNSUPPORTED, "n_segs %u outside [1,%d]", b->n_segs, THB_MAX_SEGS);
  if (b->n_bundles >= (1u << 28)) return fail(ctx, THB_EUNSUPPORTED, "more than 2^28 bundles in one batch");
  if (b->read_words < 1 || b->read_words > 4) return fail(ctx, THB_EUNSUPPORTED, "read_words %u outside [1,4] (reads up to 255 bp)", b->read_words);
  if (b->n_bundles && (!b->bundles || !b->seg_count || !b->reads)) return fail(ctx, THB_EINVAL, "null batch array");
  if (b->n_hits && !b->hits) return fail(ctx, THB_EINVAL, "null hits array");
  if (b->n_partner_hits && !b->partner_hits) return fail(ctx, THB_EINVAL, "null partner_hits array");
  return THB_OK;
}

// After a scan: grow any structure that overflowed and report whether the scan must be repeated
// (set inserts are idempotent; the insertion buffer is rolled back to `ins_before`).
int check_and_grow(thb_ctx* ctx, const unsigned long long* ins_before_p, bool* redo)
{
  Flags f; unsigned long long ins_now, fus_now[2]; *redo = false;
  int rc = read_state(ctx, &f, &ins_now, fus_now); if (rc) return rc;
  const unsigned long long ins_before = *ins_before_p;      // valid only after the sync above
  const unsigned long long fus_before = ctx->h_fus_count;   // fusion records of the launches completed before this one
  if ((f.err & 4u) || fus_now[1] > ctx->cap_fustask) { ctx->cap_fustask = std::max<uint64_t>(ctx->cap_fustask * 2, fus_now[1] + 1024); *redo = true; }
  if ((f.err & 8u) || fus_now[0] > ctx->cap_fus) {
    uint64_t ncap = std::max<uint64_t>(ctx->cap_fus * 2, fus_now[0] + 1024);
    DevBuf nb; CU(nb.reserve(ncap * sizeof(FusRec)));
    CU(cudaMemcpyAsync(nb.p, ctx->d_fus.p, (size_t)std::min<uint64_t>(fus_before, ctx->cap_fus) * sizeof(FusRec), cudaMemcpyDeviceToDevice, ctx->compute));
    CU(cudaStreamSynchronize(ctx->compute));
    ctx->d_fus.release(); ctx->d_fus = nb; ctx->cap_fus = ncap; *redo = true;
  }
  if (f.err & 16u) return fail(ctx, THB_EINVAL, "hit ranges of consecutive bundles are not laid out back to back (hit_begin[i+1] != hit_begin[i] + hits of bundle i)");
  if (f.err & 2u) return fail(ctx, THB_EUNSUPPORTED, "more than %d rescued mate-anchor hits for one read (raise RES_MAX)", RES_MAX);
  if (f.qovf & 1u) { ctx->cap_win *= 2; *redo = true; }
  if (f.qovf & 2u) { ctx->cap_indel *= 2; *redo = true; }
  if (f.ovf_juncs) { rc = grow_set(ctx, ctx->d_juncs, ctx->cap_juncs, ctx->d_ovf_juncs); if (rc) return rc; *redo = true; }
  if (f.ovf_dels) { rc = grow_set(ctx, ctx->d_dels, ctx->cap_dels, ctx->d_ovf_dels); if (rc) return rc; *redo = true; }
  if ((f.err & 1u) || ins_now > ctx->cap_ins) {
    uint64_t ncap = std::max<uint64_t>(ctx->cap_ins * 2, ins_now + 1024);
    DevBuf nb; CU(nb.reserve(ncap * sizeof(InsRec)));
    CU(cudaMemcpyAsync(nb.p, ctx->d_ins.p, (size_t)std::min<uint64_t>(ins_before, ctx->cap_ins) * sizeof(InsRec), cudaMemcpyDeviceToDevice, ctx->compute));
    CU(cudaStreamSynchronize(ctx->compute));
    ctx->d_ins.release(); ctx->d_ins = nb; ctx->cap_ins = ncap; *redo = true;
  }
  if (!*redo) { ctx->h_ins_count = ins_now; ctx->h_fus_count = fus_now[0]; }
  if (*redo) {
    CU(cudaMemcpyAsync(ctx->d_ins_count, &ins_before, sizeof ins_before, cudaMemcpyHostToDevice, ctx->compute));
    CU(cudaMemcpyAsync(ctx->d_fus_count, &fus_before, sizeof fus_before, cudaMemcpyHostToDevice, ctx->compute));
    CU(cudaMemsetAsync(ctx->d_err, 0, 2 * sizeof(unsigned int), ctx->compute));   // err + queue overflow flags
    CU(cudaStreamSynchronize(ctx->compute));
  }
  return THB_OK;
}

uint64_t algorithmic_bytes(const thb_ctx* ctx, const unsigned long long* cnt)
{
  // SURVEY.md section 8(d), B_segjuncs, evaluated on the actual task counts of the run.
  const uint64_t per_read = 16 + 40;
  return ctx->n_bundles_total * per_read + 16ull * (ctx->n_hits_total + ctx->n_partner_total) +
         cnt[0] * (32ull + 64ull) + 16ull * cnt[3] +               // windows + junction records
         cnt[1] * (32ull + 64ull) + 16ull * (ctx->n_ins_out + ctx->n_del_out) +
         cnt[2] * (32ull + 128ull);
}

// thb_join_begin: order / uniqueness of the uploaded sets and monotonicity of the junctions' global coordinates
__global__ void join_sets_check_kernel(const thb_junction* juncs, uint64_t n_juncs, const thb_insertion* ins, uint64_t n_ins, RefView ref, uint64_t n_buckets,
                                       int shift, uint64_t n_ibuckets, int ishift, unsigned int* flags)
{
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_juncs || i < n_ins; i += (uint64_t)gridDim.x * blockDim.x) {
    if (i < n_juncs) {
      const thb_junction b = juncs[i];
      bool mono = b.ref_id >= 1 && b.ref_id <= ref.n_contigs && (uint64_t)b.left <= (uint64_t)ref.contig_len[b.ref_id >= 1 && b.ref_id <= ref.n_contigs ? b.ref_id - 1 : 0] + 63;
      uint64_t g = 0;
      if (mono) { g = ref.contig_start[b.ref_id - 1] + b.left; if ((g >> shift) >= n_buckets) mono = false; }
      if (i) {
        const thb_junction a = juncs[i - 1];
        const bool lt = a.ref_id != b.ref_id ? a.ref_id < b.ref_id : a.left != b.left ? a.left < b.left : a.right != b.right ? a.right < b.right : a.antisense < b.antisense;
        if (!lt) { atomicOr(flags, 1u); flags[1] = (unsigned int)i; }
        if (mono && a.ref_id >= 1 && a.ref_id <= ref.n_contigs && ref.contig_start[a.ref_id - 1] + a.left > g) mono = false;
      }
      if (!mono) atomicOr(flags, 4u);
    }
    if (i < n_ins) {
      const thb_insertion b = ins[i];
      bool mono = b.ref_id >= 1 && b.ref_id <= ref.n_contigs && (uint64_t)b.left <= (uint64_t)ref.contig_len[b.ref_id >= 1 && b.ref_id <= ref.n_contigs ? b.ref_id - 1 : 0] + 63;
      uint64_t g = 0;
      if (mono) { g = ref.contig_start[b.ref_id - 1] + b.left; if ((g >> ishift) >= n_ibuckets) mono = false; }
      if (i) {
        const thb_insertion a = ins[i - 1];
        const bool lt = a.ref_id != b.ref_id ? a.ref_id < b.ref_id : a.left != b.left ? a.left < b.left : a.len < b.len;
        if (!lt) { atomicOr(flags, 2u); flags[1] = (unsigned int)i; }
        if (mono && a.ref_id >= 1 && a.ref_id <= ref.n_contigs && ref.contig_start[a.ref_id - 1] + a.left > g) mono = false;
      }
      if (!mono) atomicOr(flags, 8u);
    }
  }
}

// insertion reduction (thb_segjuncs_finish): field extraction for the two stable sorts, winner selection, decoding
__global__ void ins_field_kernel(const InsRec* rec, const uint64_t* idx, uint64_t n, int want_order, uint64_t* key_out, uint64_t* val_out)
{
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = idx ? idx[i] : i;
    key_out[i] = want_order ? rec[r].order : rec[r].key; val_out[i] = r;
  }
}
__global__ void ins_winner_kernel(const InsRec* rec, const uint64_t* keys_sorted, const uint64_t* idx_sorted, uint64_t n, uint64_t* wkey, uint64_t* wseq,
                                  unsigned long long* count)
{
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    if (i && keys_sorted[i - 1] == keys_sorted[i]) continue;       // not the first-inserted record of its key
    const unsigned long long slot = atomicAdd(count, 1ull);
    wkey[slot] = keys_sorted[i]; wseq[slot] = rec[idx_sorted[i]].seq;
  }
}
__global__ void ins_decode_kernel(const uint64_t* keys, const uint64_t* seqs, uint64_t n, RefView ref, thb_insertion* out)
{
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t gl1 = keys[i] >> 8; const uint32_t len = (uint32_t)(keys[i] & 0xffu); const uint64_t sq = seqs[i];
    int lo = 0, hi = (int)ref.n_contigs - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (ref.contig_start[mid] <= gl1) lo = mid; else hi = mid - 1; }
    thb_insertion o; o.ref_id = (uint32_t)lo + 1u; o.left = (uint32_t)(gl1 - ref.contig_start[lo]) - 1u; o.len = len;
    for (int k = 0; k < 20; ++k) o.seq[k] = 0;
    for (uint32_t k = 0; k < len && k < 19; ++k) { const unsigned c = (unsigned)((sq >> (3 * k)) & 7u); o.seq[k] = c == 0 ? 'A' : c == 1 ? 'C' : c == 2 ? 'G' : c == 3 ? 'T' : 'N'; }
    out[i] = o;
  }
}

// appends the valid records of a gathered (padded) insertion / fusion buffer; padding = all-ones first word
template <class Rec>
__global__ void rec_append_kernel(const Rec* src, uint64_t n, Rec* dst, unsigned long long* count, unsigned long long cap, unsigned int* err, unsigned int err_bit)
{
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const Rec r = src[i];
    if (*reinterpret_cast<const unsigned long long*>(&r) == ~0ull) continue;
    const unsigned long long slot = atomicAdd(count, 1ull);
    if (slot < cap) dst[slot] = r; else atomicOr(err, err_bit);
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
extern "C" {

const char* thb_version(void) { return "tophat_b200 0.1 (sm_100a)"; }

void thb_params_default(thb_params* p)
{
  memset(p, 0, sizeof *p);
  p->segment_length = 25; p->segment_mismatches = 2; p->min_segment_intron_length = 50;
  p->max_segment_intron_length = 500000; p->max_insertion_length = 3; p->max_deletion_length = 3;
  p->max_seg_multihits = 40; p->inner_dist_mean = 200; p->inner_dist_std_dev = 20; p->bowtie2 = 1;
  p->library_type = 0; p->fusion_search = 0; p->fusion_anchor_length = 20; p->fusion_min_dist = 10000000;
  p->max_report_intron_length = 500000; p->min_report_intron_length = 50; p->min_anchor_len = 8;
  p->read_mismatches = 2; p->read_gap_length = 2; p->read_edit_dist = 2;
  p->bowtie2_max_penalty = 6; p->bowtie2_min_penalty = 2; p->bowtie2_penalty_for_N = 1;
  p->bowtie2_read_gap_open = 5; p->bowtie2_read_gap_cont = 3; p->bowtie2_ref_gap_open = 5; p->bowtie2_ref_gap_cont = 3;
}

const char* thb_last_error(const thb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error; }

int thb_create(int device, thb_ctx** out)
{
  thb_ctx* ctx = nullptr;
  if (!out) return fail(nullptr, THB_EINVAL, "null out pointer");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) return fail(nullptr, THB_ENODEVICE, "no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(nullptr, THB_ENODEVICE, "device %d out of range (have %d)", device, n);
  cudaDeviceProp prop; e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return fail(nullptr, THB_ECUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10) return fail(nullptr, THB_ENODEVICE, "device %d is sm_%d%d; kernels are built for sm_100a only", device, prop.major, prop.minor);
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail(nullptr, THB_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  ctx = new thb_ctx(); ctx->device = device;
  auto bail = [&](const char* what, cudaError_t ee) { int rc = fail(nullptr, THB_ECUDA, "%s: %s", what, cudaGetErrorString(ee)); delete ctx; return rc; };
  if ((e = cudaStreamCreateWithFlags(&ctx->compute, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
  if ((e = cudaStreamCreateWithFlags(&ctx->copy, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
  if ((e = cudaStreamCreateWithFlags(&ctx->d2h, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
  for (auto& s : ctx->jstage) { cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming); cudaEventCreateWithFlags(&s.out_free, cudaEventDisableTiming); }
  cudaEventCreate(&ctx->ev_a); cudaEventCreate(&ctx->ev_b); cudaEventCreate(&ctx->ev_c); cudaEventCreate(&ctx->ev_d);
  for (auto& s : ctx->stage) { cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming); cudaEventCreateWithFlags(&s.consumed, cudaEventDisableTiming); }
  if ((e = ctx->d_scalars.reserve(256)) != cudaSuccess) return bail("cudaMalloc", e);
  ctx->d_counters = (unsigned long long*)ctx->d_scalars.p;
  ctx->d_ins_count = ctx->d_counters + 8;
  ctx->d_ovf_juncs = (unsigned int*)(ctx->d_counters + 9);
  ctx->d_ovf_dels = ctx->d_ovf_juncs + 1; ctx->d_err = ctx->d_ovf_juncs + 2; ctx->d_qovf = ctx->d_ovf_juncs + 3;
  ctx->d_qcounts = ctx->d_counters + 12;
  ctx->d_fus_count = ctx->d_counters + 16; ctx->d_fustask_count = ctx->d_counters + 17;
  for (auto& e : ctx->kev) cudaEventCreate(&e);
  cudaEventCreateWithFlags(&ctx->ev_sets, cudaEventDisableTiming);
  cudaDeviceGetAttribute(&ctx->sms, cudaDevAttrMultiProcessorCount, device);
  cudaMemset(ctx->d_scalars.p, 0, 256);
  thb_params_default(&ctx->params);
  *out = ctx;
  return THB_OK;
}

void thb_destroy(thb_ctx* ctx)
{
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  if (ctx->comm && ctx->nccl.CommDestroy) ctx->nccl.CommDestroy(ctx->comm);
  for (DevBuf* b : { &ctx->d_planes, &ctx->d_nmask, &ctx->d_cstart, &ctx->d_clen, &ctx->d_juncs, &ctx->d_dels, &ctx->d_ins,
                     &ctx->d_scalars, &ctx->d_keys, &ctx->d_keys_sorted, &ctx->d_cub_tmp, &ctx->d_decoded, &ctx->d_count,
                     &ctx->q_win, &ctx->q_indel, &ctx->q_rescue, &ctx->q_rescue_out, &ctx->q_rbundle, &ctx->q_bstate, &ctx->q_owner, &ctx->ag_send, &ctx->ag_recv,
                     &ctx->d_fus, &ctx->q_fus, &ctx->d_fus_ignore, &ctx->j_juncs, &ctx->j_ins, &ctx->j_bundles, &ctx->j_segc, &ctx->j_reads, &ctx->j_hits, &ctx->j_out, &ctx->j_chain, &ctx->j_idx, &ctx->j_iidx, &ctx->j_fus }) b->release();
  for (auto& e : ctx->kev) if (e) cudaEventDestroy(e);
  if (ctx->ev_sets) cudaEventDestroy(ctx->ev_sets);
  for (DevBuf* b : { &ctx->d_decoded_dels, &ctx->d_skeys_j, &ctx->d_skeys_d }) b->release();
  for (auto& s : ctx->stage) { for (DevBuf* b : { &s.bundles, &s.seg_count, &s.reads, &s.hits, &s.partner }) b->release(); cudaEventDestroy(s.copied); cudaEventDestroy(s.consumed); }
  cudaEventDestroy(ctx->ev_a); cudaEventDestroy(ctx->ev_b); cudaEventDestroy(ctx->ev_c); cudaEventDestroy(ctx->ev_d);
  for (auto& s : ctx->jstage) { for (DevBuf* b : { &s.bundles, &s.seg_count, &s.reads, &s.hits, &s.ops, &s.out }) b->release(); cudaEventDestroy(s.copied); cudaEventDestroy(s.out_free); }
  if (ctx->h_joined) cudaFreeHost(ctx->h_joined);
  ctx->h_juncs.release(); ctx->h_dels.release(); ctx->h_ins.release(); ctx->fl.release();
  for (DevBuf* b : { &ctx->d_ins_a, &ctx->d_ins_b, &ctx->d_ins_c, &ctx->d_ins_d, &ctx->d_ins_out }) b->release();
  cudaStreamDestroy(ctx->compute); cudaStreamDestroy(ctx->copy); cudaStreamDestroy(ctx->d2h);
  delete ctx;
}

void* thb_stream(thb_ctx* ctx) { return ctx ? (void*)ctx->compute : nullptr; }

void* thb_alloc_pinned(size_t bytes)
{
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
  return p;
}
void thb_free_pinned(void* p) { if (p) cudaFreeHost(p); }

void thb_pack_bases(const char* seq, uint64_t len, uint64_t gstart, uint64_t* planes, uint64_t* nmask)
{
  for (uint64_t i = 0; i < len; ++i) {
    const uint64_t g = gstart + i, b = g >> 6; const unsigned j = (unsigned)(g & 63);
    unsigned c; bool isn = false;
    switch (seq[i]) { case 'A': case 'a': c = 0; break; case 'C': case 'c': c = 1; break; case 'G': case 'g': c = 2; break;
                      case 'T': case 't': case 'U': case 'u': c = 3; break; default: c = 0; isn = true; }
    if (c & 1) planes[2 * b] |= 1ull << j;
    if (c & 2) planes[2 * b + 1] |= 1ull << j;
    if (isn) nmask[b] |= 1ull << j;
  }
}

void thb_pack_read(const char* seq, uint32_t len, uint32_t read_words, uint64_t* out)
{
  memset(out, 0, sizeof(uint64_t) * 3 * read_words);
  for (uint32_t i = 0; i < len && i < read_words * 64; ++i) {
    const uint32_t w = i >> 6, j = i & 63; unsigned c; bool isn = false;
    switch (seq[i]) { case 'A': c = 0; break; case 'C': c = 1; break; case 'G': c = 2; break; case 'T': c = 3; break; default: c = 0; isn = true; }
    if (c & 1) out[w] |= 1ull << j;
    if (c & 2) out[read_words + w] |= 1ull << j;
    if (isn) out[2 * read_words + w] |= 1ull << j;
  }
}

int thb_join_pack_hits(const thb_jhit_full* hits, uint32_t n, thb_jhit* heads, thb_jops* ops_ext)
{
  if ((n && (!hits || !heads)) ) return THB_EINVAL;
  uint32_t n_ext = 0;
  for (uint32_t i = 0; i < n; ++i) {
    const thb_jhit_full& f = hits[i]; thb_jhit& h = heads[i];
    uint32_t nops = f.n_ops; if (nops > THB_JHIT_MAX_OPS) return THB_EINVAL;
    int64_t right = f.left;
    for (uint32_t k = 0; k < nops; ++k) {                         // BowtieHit::right(), bwt_map.h:213-243
      const uint32_t c = f.ops[k] & 15u; const int64_t l = (int64_t)(f.ops[k] >> 4);
      if (c == 1 || c == 5 || c == 11) right += l; else if (c == 2 || c == 6 || c == 12) right -= l; else if (c >= 7 && c <= 10) right = l;
    }
    const bool one_match = nops == 1 && (f.ops[0] & 15u) == 1u;
    h.ref_id = f.ref_id; h.left = f.left; h.right = (int32_t)right;
    h.flags_nops = (uint8_t)((f.flags & 0x7u) | (one_match ? THB_JHIT_ONE_MATCH : 0) | (nops << 4));
    h.mismatches = f.mismatches; h.splice_mms = f.splice_mms; h.ops_index = 0;
    if (!one_match) {
      if (n_ext > 255 || !ops_ext) return THB_EUNSUPPORTED;
      h.ops_index = (uint8_t)n_ext;
      thb_jops& o = ops_ext[n_ext++]; memset(&o, 0, sizeof o);
      bool fused = false;
      for (uint32_t k = 0; k < nops; ++k) { o.ops[k] = f.ops[k]; const uint32_t c = f.ops[k] & 15u; fused = fused || (c >= 7u && c <= 10u); }
      if (fused) { if (nops > 8) return THB_EINVAL; o.ops[11] = f.ops[8]; }          // second contig of a fusion hit
      if (f.flags & THB_JHIT_SEQ_FLIPPED) o.ops[10] = 1u;
    }
  }
  return (int)n_ext;
}

int thb_ref_upload(thb_ctx* ctx, const thb_ref_image* img)
{
  if (!ctx || !img) return THB_EINVAL;
  CU(cudaSetDevice(ctx->device));
  if (img->n_contigs == 0 || !img->contig_start || !img->contig_len || !img->planes || !img->nmask)
    return fail(ctx, THB_EINVAL, "incomplete reference image");
  for (uint32_t i = 0; i < img->n_contigs; ++i) {
    if (img->contig_start[i] & 63) return fail(ctx, THB_EINVAL, "contig_start[%u] is not a multiple of 64", i);
    const uint64_t end = img->contig_start[i] + img->contig_len[i] + 64;
    if (end > img->n_blocks * 64) return fail(ctx, THB_EINVAL, "contig %u overruns the image (needs 64 bases of padding)", i);
    if (i && img->contig_start[i] < img->contig_start[i - 1] + img->contig_len[i - 1] + 64) return fail(ctx, THB_EINVAL, "contigs %u/%u overlap or lack padding", i - 1, i);
  }
  if (img->n_blocks * 64 >= (1ull << 39)) return fail(ctx, THB_EUNSUPPORTED, "reference larger than 2^39 bases");
  CU(ctx->d_planes.reserve((img->n_blocks + 2) * 16)); CU(ctx->d_nmask.reserve((img->n_blocks + 2) * 8));
  CU(ctx->d_cstart.reserve(img->n_contigs * 8)); CU(ctx->d_clen.reserve(img->n_contigs * 4));
  CU(cudaMemsetAsync(ctx->d_planes.p, 0, (img->n_blocks + 2) * 16, ctx->compute));
  CU(cudaMemsetAsync(ctx->d_nmask.p, 0, (img->n_blocks + 2) * 8, ctx->compute));
  CU(cudaMemcpyAsync(ctx->d_planes.p, img->planes, img->n_blocks * 16, cudaMemcpyHostToDevice, ctx->compute));
  CU(cudaMemcpyAsync(ctx->d_nmask.p, img->nmask, img->n_blocks * 8, cudaMemcpyHostToDevice, ctx->compute));
  CU(cudaMemcpyAsync(ctx->d_cstart.p, img->contig_start, img->n_contigs * 8, cudaMemcpyHostToDevice, ctx->compute));
  CU(cudaMemcpyAsync(ctx->d_clen.p, img->contig_len, img->n_contigs * 4, cudaMemcpyHostToDevice, ctx->compute));
  CU(cudaStreamSynchronize(ctx->compute));
  ctx->h_cstart.assign(img->contig_start, img->contig_start + img->n_contigs);
  ctx->h_clen.assign(img->contig_len, img->contig_len + img->n_contigs);
  ctx->ref.planes = (const ulonglong2*)ctx->d_planes.p; ctx->ref.nmask = (const uint64_t*)ctx->d_nmask.p;
  ctx->ref.contig_start = (const uint64_t*)ctx->d_cstart.p; ctx->ref.contig_len = (const uint32_t*)ctx->d_clen.p;
  ctx->ref.n_contigs = img->n_contigs; ctx->have_ref = true;
  return THB_OK;
}

int thb_segjuncs_begin(thb_ctx* ctx, const thb_params* p)
{
  if (!ctx || !p) return THB_EINVAL;
  CU(cudaSetDevice(ctx->device));
  if (p->segment_length < 4 || p->segment_length > 32)
    return fail(ctx, THB_EUNSUPPORTED, "--segment-length %d outside the GPU path's [4,32] (two segments must fit one 64-bit plane word)", p->segment_length);
  if (p->max_segment_intron_length + p->segment_length + 64 >= (1 << KEY_LEN_BITS))
    return fail(ctx, THB_EUNSUPPORTED, "--max-segment-intron %d too large for the 24-bit span field", p->max_segment_intron_length);
  if (p->max_insertion_length > 19 || p->max_deletion_length > 1000)      // thb_insertion.seq holds 19 bases + NUL
    return fail(ctx, THB_EUNSUPPORTED, "--max-insertion-length > 19 / --max-deletion-length > 1000 not supported");
  if (p->inner_dist_mean + p->inner_dist_std_dev + std::max(0, p->inner_dist_std_dev - p->inner_dist_mean) > 100000)
    return fail(ctx, THB_EUNSUPPORTED, "mate flank longer than 100000 bases");
  ctx->params = *p;
  SegParams& s = ctx->sp;
  s.seglen = p->segment_length; s.segmm = p->segment_mismatches; s.min_intron = p->min_segment_intron_length;
  s.max_intron = p->max_segment_intron_length; s.max_ins = p->max_insertion_length; s.max_del = p->max_deletion_length;
  s.max_multihits = p->max_seg_multihits; s.inner_mean = p->inner_dist_mean; s.inner_sd = p->inner_dist_std_dev;
  s.bowtie2 = p->bowtie2; s.library_type = p->library_type;
  s.fusion_search = p->fusion_search ? 1 : 0; s.fusion_min_dist = p->fusion_min_dist;
  ctx->fp.fusion_min_dist = p->fusion_min_dist; ctx->fp.fusion_anchor = p->fusion_anchor_length; ctx->fp.n_ignore = 0; ctx->fp.ignore = nullptr;
  if (s.fusion_search) { if (ctx->cap_fus == 0) ctx->cap_fus = tiny_caps() ? 64 : 1ull << 16; CU(ctx->d_fus.reserve(ctx->cap_fus * sizeof(FusRec))); }
  if (ctx->cap_juncs == 0) ctx->cap_juncs = tiny_caps() ? 64 : 1ull << 21;
  if (ctx->cap_dels == 0) ctx->cap_dels = tiny_caps() ? 64 : 1ull << 18;
  if (ctx->cap_ins == 0) ctx->cap_ins = tiny_caps() ? 64 : 1ull << 18;
  int rc;
  if ((rc = alloc_set(ctx, ctx->d_juncs, ctx->cap_juncs))) return rc;
  if ((rc = alloc_set(ctx, ctx->d_dels, ctx->cap_dels))) return rc;
  CU(ctx->d_ins.reserve(ctx->cap_ins * sizeof(InsRec)));
  CU(cudaMemsetAsync(ctx->d_scalars.p, 0, 256, ctx->compute));
  CU(cudaStreamSynchronize(ctx->compute));
  ctx->h_juncs.clear(); ctx->h_dels.clear(); ctx->h_ins.clear(); ctx->h_fus.clear();
  memset(&ctx->timing, 0, sizeof ctx->timing);
  ctx->own_launches = 2;            // the two hs_clear launches above
  ctx->h_ins_count = 0; ctx->h_fus_count = 0;
  ctx->n_bundles_total = ctx->n_hits_total = ctx->n_partner_total = 0; ctx->n_ins_out = ctx->n_del_out = 0;
  ctx->ag_keys[0] = ctx->ag_keys[1] = nullptr; ctx->ag_n[0] = ctx->ag_n[1] = 0;
  ctx->begun = true;
  return THB_OK;
}

int thb_segjuncs_fusion_ignore(thb_ctx* ctx, const uint32_t* ref_ids, uint32_t n)
{
  if (!ctx || (n && !ref_ids)) return THB_EINVAL;
  CU(cudaSetDevice(ctx->device));
  if (!ctx->begun) return fail(ctx, THB_ESTATE, "thb_segjuncs_begin not called");
  if (n > 4096) return fail(ctx, THB_EUNSUPPORTED, "more than 4096 ignored contigs");
  CU(ctx->d_fus_ignore.reserve((size_t)(n + 1) * 4));
  if (n) CU(cudaMemcpyAsync(ctx->d_fus_ignore.p, ref_ids, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->compute));
  CU(cudaStreamSynchronize(ctx->compute));
  ctx->fp.n_ignore = (int)n; ctx->fp.ignore = (const uint32_t*)ctx->d_fus_ignore.p;
  return THB_OK;
}

int thb_segjuncs_submit_device(thb_ctx* ctx, const thb_segjuncs_batch* b)
{
  if (!ctx || !b) return THB_EINVAL;
  CU(cudaSetDevice(ctx->device));
  int rc = validate_batch(ctx, b); if (rc) return rc;
  BatchView bv; bv.bundles = b->bundles; bv.seg_count = b->seg_count; bv.reads = b->reads; bv.hits = b->hits;
  bv.partner = b->partner_hits; bv.n_bundles = b->n_bundles; bv.n_segs = b->n_segs; bv.read_words = b->read_words;
  bv.order_base = b->order_base;
  bv.partner_base = 0; bv.hit_base = 0; bv.hit_end = b->n_hits; bv.partner_end = b->n_partner_hits;
  unsigned long long ins_before = ctx->h_ins_count;
  float ms_total = 0.f; bool done = false;
  for (int attempt = 0; attempt < 24 && !done; ++attempt) {
    unsigned long long cnt0[8];
    CU(cudaMemcpyAsync(cnt0, ctx->d_counters, sizeof cnt0, cudaMemcpyDeviceToHost, ctx->compute));
    rc = launch_scan(ctx, bv, b->n_partner_hits, b->n_hits); if (rc) return rc;
    bool redo = false;
    rc = check_and_grow(ctx, &ins_before, &redo); if (rc) return rc;
    if (!redo) { collect_kernel_times(ctx, &ms_total); done = true; break; }
    ctx->kev_pending = false;
    // roll the task counters back so that a repeated scan is not double counted
    CU(cudaMemcpyAsync(ctx->d_counters, cnt0, sizeof cnt0, cudaMemcpyHostToDevice, ctx->compute));
    CU(cudaStreamSynchronize(ctx->compute));
  }
  if (!done) return fail(ctx, THB_ENOMEM, "result structures still overflow after 24 growth attempts");
  ctx->timing.scan_kernel_ms += ms_total; ctx->timing.total_ms += ms_total;
  ctx->n_bundles_total += b->n_bundles; ctx->n_hits_total += b->n_hits; ctx->n_partner_total += b->n_partner_hits;
  return THB_OK;
}

int thb_segjuncs_submit(thb_ctx* ctx, const thb_segjuncs_batch* b)
{
  if (!ctx || !b) return THB_EINVAL;
  CU(cudaSetDevice(ctx->device));
  int rc = validate_batch(ctx, b); if (rc) return rc;
  if (b->n_bundles == 0) return THB_OK;
  const uint32_t CH = 1u << 20;                         // bundles per pipeline chunk
  const size_t rdw = (size_t)3 * b->read_words;
  CU(cudaEventRecord(ctx->ev_c, ctx->compute));
  float kernel_ms = 0.f;
  const uint32_t nchunks = (b->n_bundles + CH - 1) / CH;
  struct Range { uint32_t b0, nb; uint64_t h0, h1, p0, p1; };
  auto range_of = [&](uint32_t c, Range* r) -> bool {
    r->b0 = c * CH; const uint32_t b1 = std::min<uint32_t>(b->n_bundles, r->b0 + CH); r->nb = b1 - r->b0;
    r->h0 = b->bundles[r->b0].hit_begin; r->h1 = (b1 < b->n_bundles) ? b->bundles[b1].hit_begin : b->n_hits;
    r->p0 = b->bundles[r->b0].partner_begin; r->p1 = (b1 < b->n_bundles) ? b->bundles[b1].partner_begin : b->n_partner_hits;
    return !(r->h1 < r->h0 || r->h1 > b->n_hits || r->p1 < r->p0 || r->p1 > b->n_partner_hits);
  };
  // host -> device copy of chunk c into staging buffer c&1 on the copy stream
  auto enqueue_copy = [&](uint32_t c) -> int {
    Staging& s = ctx->stage[c & 1]; Range r;
    if (!range_of(c, &r)) return fail(ctx, THB_EINVAL, "hit_begin / partner_begin not monotonic near bundle %u", c * CH);
    if (s.used) CU(cudaStreamWaitEvent(ctx->copy, s.consumed, 0));
    CU(s.bundles.reserve((size_t)r.nb * sizeof(thb_bundle))); CU(s.seg_count.reserve((size_t)r.nb * b->n_segs * 2));
    CU(s.reads.reserve((size_t)r.nb * rdw * 8)); CU(s.hits.reserve((size_t)(r.h1 - r.h0 + 1) * sizeof(thb_hit)));
    CU(s.partner.reserve((size_t)(r.p1 - r.p0 + 1) * sizeof(thb_hit)));
    CU(cudaMemcpyAsync(s.bundles.p, b->bundles + r.b0, (size_t)r.nb * sizeof(thb_bundle), cudaMemcpyHostToDevice, ctx->copy));
    CU(cudaMemcpyAsync(s.seg_count.p, b->seg_count + (size_t)r.b0 * b->n_segs, (size_t)r.nb * b->n_segs * 2, cudaMemcpyHostToDevice, ctx->copy));
    CU(cudaMemcpyAsync(s.reads.p, b->reads + (size_t)r.b0 * rdw, (size_t)r.nb * rdw * 8, cudaMemcpyHostToDevice, ctx->copy));
    if (r.h1 > r.h0) CU(cudaMemcpyAsync(s.hits.p, b->hits + r.h0, (size_t)(r.h1 - r.h0) * sizeof(thb_hit), cudaMemcpyHostToDevice, ctx->copy));
    if (r.p1 > r.p0) CU(cudaMemcpyAsync(s.partner.p, b->partner_hits + r.p0, (size_t)(r.p1 - r.p0) * sizeof(thb_hit), cudaMemcpyHostToDevice, ctx->copy));
    CU(cudaEventRecord(s.copied, ctx->copy));
    return THB_OK;
  };
  rc = enqueue_copy(0); if (rc) return rc;
  for (uint32_t c = 0; c < nchunks; ++c) {
    Staging& s = ctx->stage[c & 1]; Range r; range_of(c, &r);
    CU(cudaStreamWaitEvent(ctx->compute, s.copied, 0));
    BatchView bv; bv.bundles = (const thb_bundle*)s.bundles.p; bv.seg_count = (const uint16_t*)s.seg_count.p;
    bv.reads = (const uint64_t*)s.reads.p; bv.hits = (const thb_hit*)s.hits.p - r.h0; bv.partner = (const thb_hit*)s.partner.p - r.p0;
    bv.n_bundles = r.nb; bv.n_segs = b->n_segs; bv.read_words = b->read_words; bv.order_base = b->order_base + r.b0;
    bv.partner_base = r.p0; bv.hit_base = r.h0; bv.hit_end = r.h1; bv.partner_end = r.p1;
    unsigned long long ins_before = ctx->h_ins_count;
    bool done = false;
    for (int attempt = 0; attempt < 24; ++attempt) {
      unsigned long long cnt0[8];
      CU(cudaMemcpyAsync(cnt0, ctx->d_counters, sizeof cnt0, cudaMemcpyDeviceToHost, ctx->compute));
      rc = launch_scan(ctx, bv, r.p1 - r.p0, r.h1 - r.h0); if (rc) return rc;
      // the next chunk's copy (other staging buffer, whose kernel already completed) overlaps this scan
      if (attempt == 0 && c + 1 < nchunks) { rc = enqueue_copy(c + 1); if (rc) return rc; }
      bool redo = false;
      rc = check_and_grow(ctx, &ins_before, &redo); if (rc) return rc;      // synchronises the compute stream
      if (!redo) { float ms = 0.f; collect_kernel_times(ctx, &ms); kernel_ms += ms; done = true; break; }
      ctx->kev_pending = false;
      CU(cudaMemcpyAsync(ctx->d_counters, cnt0, sizeof cnt0, cudaMemcpyHostToDevice, ctx->compute));
      CU(cudaStreamSynchronize(ctx->compute));
    }
    if (!done) return fail(ctx, THB_ENOMEM, "result structures still overflow after 24 growth attempts");
    CU(cudaEventRecord(s.consumed, ctx->compute)); s.used = true;
  }
  CU(cudaEventRecord(ctx->ev_d, ctx->compute));
  CU(cudaStreamSynchronize(ctx->compute));
  float tot = 0.f; CU(cudaEventElapsedTime(&tot, ctx->ev_c, ctx->ev_d));
  ctx->timing.scan_kernel_ms += kernel_ms; ctx->timing.total_ms += tot; ctx->timing.h2d_ms += std::max(0.f, tot - kernel_ms);
  ctx->n_bundles_total += b->n_bundles; ctx->n_hits_total += b->n_hits; ctx->n_partner_total += b->n_partner_hits;
  return THB_OK;
}

// THB_TRACE=1: host wall clock of the phases of thb_segjuncs_finish on stderr (development aid)
struct Trace {
  bool on = getenv("THB_TRACE") != nullptr; std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now(); std::string line;
  void mark(const char* what) { if (!on) return; const auto n = std::chrono::steady_clock::now(); char b[96];
    snprintf(b, sizeof b, " %s %.3f", what, std::chrono::duration<double, std::milli>(n - t).count()); line += b; t = n; }
  ~Trace() { if (on && !line.empty()) fprintf(stderr, "[thb trace ms]%s\n", line.c_str()); }
};

// decoded: device buffer that keeps the records (resident sets); skeys: keeps the sorted keys; async: the download is queued on the
// d2h stream behind the decode and NOT waited for (thb_segjuncs_fetch does)
static int finish_set(thb_ctx* ctx, DevBuf& set, uint64_t cap, PinnedVec<thb_junction>& out, uint64_t limit, Trace* tr,
                      const uint64_t* gathered, uint64_t n_gathered, DevBuf& decoded, DevBuf& skeys, bool async)
{
  CU(ctx->d_count.reserve(64));
  unsigned long long n = 0;
  const uint64_t* sorted = nullptr;
  size_t tmp = 0;
  if (gathered) {
    // after thb_segjuncs_allgather: the union is the sorted, de-duplicated concatenation of every rank's keys -- one radix sort and
    // one unique pass instead of world x keys random inserts into a hash set sized for all of them
    out.clear();
    if (n_gathered == 0) return THB_OK;
    CU(ctx->d_keys.reserve(n_gathered * 8)); CU(ctx->d_keys_sorted.reserve(n_gathered * 8));
    cub::DeviceRadixSort::SortKeys(nullptr, tmp, gathered, (uint64_t*)ctx->d_keys.p, (int)n_gathered, 0, 64, ctx->compute);
    CU(ctx->d_cub_tmp.reserve(tmp + 16));
    CU(cub::DeviceRadixSort::SortKeys(ctx->d_cub_tmp.p, tmp, gathered, (uint64_t*)ctx->d_keys.p, (int)n_gathered, 0, 64, ctx->compute));
    tmp = 0;
    cub::DeviceSelect::Unique(nullptr, tmp, (const uint64_t*)ctx->d_keys.p, (uint64_t*)ctx->d_keys_sorted.p, (unsigned long long*)ctx->d_count.p, (int)n_gathered, ctx->compute);
    CU(ctx->d_cub_tmp.reserve(tmp + 16));
    CU(cub::DeviceSelect::Unique(ctx->d_cub_tmp.p, tmp, (const uint64_t*)ctx->d_keys.p, (uint64_t*)ctx->d_keys_sorted.p, (unsigned long long*)ctx->d_count.p, (int)n_gathered, ctx->compute));
    uint64_t last = 0;
    CU(cudaMemcpyAsync(&n, ctx->d_count.p, 8, cudaMemcpyDeviceToHost, ctx->compute));
    CU(cudaStreamSynchronize(ctx->compute));
    if (n) { CU(cudaMemcpy(&last, (const uint64_t*)ctx->d_keys_sorted.p + (n - 1), 8, cudaMemcpyDeviceToHost)); if (last == HS_EMPTY) --n; }    // the padding
    ctx->own_launches += 2;
    if (tr) tr->mark("sort+unique(gathered)");
    if (n == 0) return THB_OK;
    sorted = (const uint64_t*)ctx->d_keys_sorted.p;
  } else {
    CU(ctx->d_keys.reserve(cap * 8)); CU(ctx->d_keys_sorted.reserve(cap * 8));
    CU(cudaMemsetAsync(ctx->d_count.p, 0, 8, ctx->compute));
    hs_compact_kernel<<<grid_for(cap, 256), 256, 0, ctx->compute>>>((const uint64_t*)set.p, cap, (uint64_t*)ctx->d_keys.p, (unsigned long long*)ctx->d_count.p);
    CU(cudaGetLastError()); ctx->own_launches++;
    CU(cudaMemcpyAsync(&n, ctx->d_count.p, 8, cudaMemcpyDeviceToHost, ctx->compute));
    CU(cudaStreamSynchronize(ctx->compute));
    if (tr) tr->mark("compact");
    out.clear();
    if (n == 0) return THB_OK;
    cub::DeviceRadixSort::SortKeys(nullptr, tmp, (const uint64_t*)ctx->d_keys.p, (uint64_t*)ctx->d_keys_sorted.p, (int)n, 0, 64, ctx->compute);
    CU(ctx->d_cub_tmp.reserve(tmp + 16));
    CU(cub::DeviceRadixSort::SortKeys(ctx->d_cub_tmp.p, tmp, (const uint64_t*)ctx->d_keys.p, (uint64_t*)ctx->d_keys_sorted.p, (int)n, 0, 64, ctx->compute));
    sorted = (const uint64_t*)ctx->d_keys_sorted.p;
  }
  if (n > limit) n = limit;       // std::set capped at max_seg_juncs by erasing the largest (1692-1693)
  CU(decoded.reserve(n * sizeof(thb_junction))); CU(skeys.reserve(n * 8));
  CU(cudaMemcpyAsync(skeys.p, sorted, n * 8, cudaMemcpyDeviceToDevice, ctx->compute));
  decode_keys_kernel<<<grid_for(n, 256), 256, 0, ctx->compute>>>(sorted, n, ctx->ref, (thb_junction*)decoded.p);
  CU(cudaGetLastError()); ctx->own_launches++;
  if (tr) tr->mark("sort+decode(enqueue)");
  CU(out.resize(n));
  if (tr) tr->mark("host_resize");
  if (async) {
    CU(cudaEventRecord(ctx->ev_sets, ctx->compute));
    CU(cudaStreamWaitEvent(ctx->d2h, ctx->ev_sets, 0));
    CU(cudaMemcpyAsync(out.data(), decoded.p, n * sizeof(thb_junction), cudaMemcpyDeviceToHost, ctx->d2h));
    ctx->fetch_pending = true;
  } else {
    CU(cudaMemcpyAsync(out.data(), decoded.p, n * sizeof(thb_junction), cudaMemcpyDeviceToHost, ctx->compute));
    CU(cudaStreamSynchronize(ctx->compute));
  }
  if (tr) tr->mark("d2h+sync");
  return THB_OK;
}

static int finish_impl(thb_ctx* ctx, thb_segjuncs_results* out, bool async)
{
  if (!ctx || !out) return THB_EINVAL;
  CU(cudaSetDevice(ctx->device));
  if (!ctx->begun) return fail(ctx, THB_ESTATE, "thb_segjuncs_begin not called");
  if (ctx->fetch_pending) { CU(cudaStreamSynchronize(ctx->d2h)); ctx->fetch_pending = false; }
  ctx->have_resident = false;
  Trace tr;
  CU(cudaEventRecord(ctx->ev_a, ctx->compute));
  int rc;
  if ((rc = finish_set(ctx, ctx->d_juncs, ctx->cap_juncs, ctx->h_juncs, 10000000ull, &tr, ctx->ag_keys[0], ctx->ag_n[0], ctx->d_decoded, ctx->d_skeys_j, async))) return rc;
  if ((rc = finish_set(ctx, ctx->d_dels, ctx->cap_dels, ctx->h_dels, ~0ull, &tr, ctx->ag_keys[1], ctx->ag_n[1], ctx->d_decoded_dels, ctx->d_skeys_d, async))) return rc;
  // insertions: first inserted wins among equal (ref, left, length) -- insertions.h:52-67
  unsigned long long nins = 0;
  CU(cudaMemcpyAsync(&nins, ctx->d_ins_count, 8, cudaMemcpyDeviceToHost, ctx->compute));
  CU(cudaStreamSynchronize(ctx->compute));
  // Reduced on the device: stable radix sort by processing order, then by (position, length) key -- the first record of every
  // key run is the one std::set would have kept -- winners compacted, sorted by key, decoded.  (The first version copied every
  // raw record to the host and std::sort-ed them: 60 ms per 5 M indel-heavy pairs.)
  ctx->h_ins.clear();
  if (nins) {
    const int n = (int)nins;
    CU(ctx->d_ins_a.reserve(nins * 8)); CU(ctx->d_ins_b.reserve(nins * 8)); CU(ctx->d_ins_c.reserve(nins * 8)); CU(ctx->d_ins_d.reserve(nins * 8));
    uint64_t* ka = (uint64_t*)ctx->d_ins_a.p; uint64_t* kb = (uint64_t*)ctx->d_ins_b.p; uint64_t* va = (uint64_t*)ctx->d_ins_c.p; uint64_t* vb = (uint64_t*)ctx->d_ins_d.p;
    const InsRec* rec = (const InsRec*)ctx->d_ins.p;
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, ka, kb, va, vb, n, 0, 64, ctx->compute);
    CU(ctx->d_cub_tmp.reserve(tmp + 16));
    ins_field_kernel<<<grid_for(nins, 256), 256, 0, ctx->compute>>>(rec, nullptr, nins, 1, ka, va);                 // ka = order, va = index
    CU(cub::DeviceRadixSort::SortPairs(ctx->d_cub_tmp.p, tmp, ka, kb, va, vb, n, 0, 64, ctx->compute));              // vb = indices by order
    ins_field_kernel<<<grid_for(nins, 256), 256, 0, ctx->compute>>>(rec, vb, nins, 0, ka, va);                      // ka = key of vb[i], va = vb[i]
    CU(cub::DeviceRadixSort::SortPairs(ctx->d_cub_tmp.p, tmp, ka, kb, va, vb, n, 0, 64, ctx->compute));              // kb = keys sorted, vb = indices (ties by order)
    CU(ctx->d_count.reserve(64)); CU(cudaMemsetAsync(ctx->d_count.p, 0, 8, ctx->compute));
    ins_winner_kernel<<<grid_for(nins, 256), 256, 0, ctx->compute>>>(rec, kb, vb, nins, ka, va, (unsigned long long*)ctx->d_count.p);   // ka / va = winners' key / seq
    unsigned long long nw = 0;
    CU(cudaMemcpyAsync(&nw, ctx->d_count.p, 8, cudaMemcpyDeviceToHost, ctx->compute));
    CU(cudaStreamSynchronize(ctx->compute));
    CU(cub::DeviceRadixSort::SortPairs(ctx->d_cub_tmp.p, tmp, ka, kb, va, vb, (int)nw, 0, 64, ctx->compute));        // winners in key order
    CU(ctx->d_ins_out.reserve(nw * sizeof(thb_insertion)));
    ins_decode_kernel<<<grid_for(nw, 256), 256, 0, ctx->compute>>>(kb, vb, nw, ctx->ref, (thb_insertion*)ctx->d_ins_out.p);
    CU(cudaGetLastError()); ctx->own_launches += 4;
    CU(ctx->h_ins.resize(nw));
    if (async) {
      CU(cudaEventRecord(ctx->ev_sets, ctx->compute));
      CU(cudaStreamWaitEvent(ctx->d2h, ctx->ev_sets, 0));
      CU(cudaMemcpyAsync(ctx->h_ins.data(), ctx->d_ins_out.p, nw * sizeof(thb_insertion), cudaMemcpyDeviceToHost, ctx->d2h));
      ctx->fetch_pending = true;
    } else {
      CU(cudaMemcpyAsync(ctx->h_ins.data(), ctx->d_ins_out.p, nw * sizeof(thb_insertion), cudaMemcpyDeviceToHost, ctx->compute));
      CU(cudaStreamSynchronize(ctx->compute));
    }
  }
  // fusions: reduce the appended records by key -- count, minimum edit distance (2787-2803, fusions.h:87-101)
  ctx->h_fus.clear();
  if (ctx->sp.fusion_search) {
    unsigned long long nfus = 0;
    CU(cudaMemcpyAsync(&nfus, ctx->d_fus_count, 8, cudaMemcpyDeviceToHost, ctx->compute));
    CU(cudaStreamSynchronize(ctx->compute));
    ctx->h_fusrec.resize(nfus);
    if (nfus) { CU(cudaMemcpyAsync(ctx->h_fusrec.data(), ctx->d_fus.p, nfus * sizeof(FusRec), cudaMemcpyDeviceToHost, ctx->compute)); CU(cudaStreamSynchronize(ctx->compute)); }
    auto key_lt = [](const FusRec& a, const FusRec& b) {          // Fusion::operator<, fusions.h:40-70
      if (a.r1 != b.r1) return a.r1 < b.r1; if (a.r2 != b.r2) return a.r2 < b.r2; if (a.left != b.left) return a.left < b.left;
      if (a.right != b.right) return a.right < b.right; return a.dir < b.dir; };
    std::sort(ctx->h_fusrec.begin(), ctx->h_fusrec.end(), key_lt);
    for (size_t i = 0; i < ctx->h_fusrec.size(); ++i) {
      const FusRec& r = ctx->h_fusrec[i];
      const uint32_t cnt1 = r.pad0 ? r.pad0 : 1u;                  // pad0: count carried by an already reduced record (all-gather)
      if (!ctx->h_fus.empty()) {
        thb_fusion& b = ctx->h_fus.back();
        if (b.ref_id1 == r.r1 && b.ref_id2 == r.r2 && b.left == r.left && b.right == r.right && b.dir == r.dir) { b.count += cnt1; b.edit_dist = std::min(b.edit_dist, r.edit); continue; }
      }
      thb_fusion o; o.ref_id1 = r.r1; o.ref_id2 = r.r2; o.left = r.left; o.right = r.right; o.dir = r.dir; o.count = cnt1; o.edit_dist = r.edit; o.reserved = 0;
      ctx->h_fus.push_back(o);
    }
  }
  tr.mark("ins+fus");
  CU(cudaEventRecord(ctx->ev_b, ctx->compute));
  CU(cudaStreamSynchronize(ctx->compute));
  float ms = 0.f; CU(cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b));
  ctx->timing.finish_ms = ms;
  tr.mark("events");
  unsigned long long cnt[8];
  CU(cudaMemcpy(cnt, ctx->d_counters, sizeof cnt, cudaMemcpyDeviceToHost));
  ctx->n_ins_out = nins; ctx->n_del_out = ctx->h_dels.size();
  ctx->timing.n_windows = cnt[0]; ctx->timing.n_indel_tasks = cnt[1]; ctx->timing.n_rescue_tasks = cnt[2]; ctx->timing.n_juncs_emitted = cnt[3];
  ctx->timing.n_fusion_tasks = cnt[7];
  ctx->timing.algorithmic_bytes = algorithmic_bytes(ctx, cnt);
  ctx->timing.total_launches = ctx->own_launches;
  out->n_junctions = ctx->h_juncs.size(); out->junctions = ctx->h_juncs.data();
  out->n_deletions = ctx->h_dels.size(); out->deletions = ctx->h_dels.data();
  out->n_insertions = ctx->h_ins.size(); out->insertions = ctx->h_ins.data();
  out->n_fusions = ctx->h_fus.size(); out->fusions = ctx->h_fus.data();
  ctx->res_n[0] = ctx->h_juncs.size(); ctx->res_n[1] = ctx->h_dels.size(); ctx->res_n[2] = ctx->h_ins.size(); ctx->have_resident = true;
  return THB_OK;
}

int thb_segjuncs_finish(thb_ctx* ctx, thb_segjuncs_results* out) { return finish_impl(ctx, out, false); }
int thb_segjuncs_finish_resident(thb_ctx* ctx, thb_segjuncs_results* out) { return finish_impl(ctx, out, true); }
int thb_segjuncs_fetch(thb_ctx* ctx)
{
  if (!ctx) return THB_EINVAL;
  CU(cudaSetDevice(ctx->device));
  if (ctx->fetch_pending) { CU(cudaStreamSynchronize(ctx->d2h)); ctx->fetch_pending = false; }
  return THB_OK;
}

int thb_last_timing(thb_ctx* ctx, thb_timing* out)
{
  if (!ctx || !out) return THB_EINVAL;
  *out = ctx->timing;
  return THB_OK;
}


// ---- long_spanning_reads join -------------------------------------------------------------------------
static int join_begin_checks(thb_ctx* ctx, const thb_params* p)
{
  if (!ctx->have_ref) return fail(ctx, THB_ESTATE, "no reference image uploaded");
  if (p->segment_length < 4 || p->segment_length > 255) return fail(ctx, THB_EUNSUPPORTED, "--segment-length %d outside [4,255]", p->segment_length);
  if (p->max_insertion_length > 19) return fail(ctx, THB_EUNSUPPORTED, "--max-insertion-length > 19 not supported");
  return THB_OK;
}
static int join_index_sets(thb_ctx* ctx, const thb_params* p, uint64_t n_juncs, uint64_t n_ins);     // validation + bucket index + parameters (ev_a recorded by the caller)

int thb_join_begin(thb_ctx* ctx, const thb_params* p, const thb_junction* juncs, uint64_t n_juncs, const thb_insertion* ins, uint64_t n_ins)
{
  if (!ctx || !p) return THB_EINVAL;
  CU(cudaSetDevice(ctx->device));
  int rc = join_begin_checks(ctx, p); if (rc) return rc;
  if ((n_juncs && !juncs) || (n_ins && !ins) || n_juncs >= (1ull << 31) || n_ins >= (1ull << 31)) return fail(ctx, THB_EINVAL, "bad junction / insertion set");
  CU(ctx->j_juncs.reserve((n_juncs + 1) * sizeof(thb_junction))); CU(ctx->j_ins.reserve((n_ins + 1) * sizeof(thb_insertion)));
  CU(cudaEventRecord(ctx->ev_a, ctx->compute));
  if (n_juncs) CU(cudaMemcpyAsync(ctx->j_juncs.p, juncs, n_juncs * sizeof(thb_junction), cudaMemcpyHostToDevice, ctx->compute));
  if (n_ins) CU(cudaMemcpyAsync(ctx->j_ins.p, ins, n_ins * sizeof(thb_insertion), cudaMemcpyHostToDevice, ctx->compute));
  return join_index_sets(ctx, p, n_juncs, n_ins);
}

// The sets of the finished segment_juncs pass, still on the device, become the join's sets: junctions and deletions (as
// Junction(ref, left, right, antisense = false)) merged in Junction order without duplicates, as long_spanning_reads builds its set from
// the two files (2895-2944); the insertions as they are.  No host round trip.
int thb_join_begin_resident(thb_ctx* ctx, const thb_params* p)
{
  if (!ctx || !p) return THB_EINVAL;
  CU(cudaSetDevice(ctx->device));
  int rc = join_begin_checks(ctx, p); if (rc) return rc;
  if (!ctx->have_resident) return fail(ctx, THB_ESTATE, "thb_join_begin_resident without a finished segment_juncs pass in this context");
  const uint64_t nj = ctx->res_n[0], nd = ctx->res_n[1], ni = ctx->res_n[2];
  uint64_t n_juncs = nj;
  CU(ctx->j_juncs.reserve((nj + nd + 1) * sizeof(thb_junction))); CU(ctx->j_ins.reserve((ni + 1) * sizeof(thb_insertion)));
  CU(cudaEventRecord(ctx->ev_a, ctx->compute));
  if (nd == 0) {
    if (nj) CU(cudaMemcpyAsync(ctx->j_juncs.p, ctx->d_decoded.p, nj * sizeof(thb_junction), cudaMemcpyDeviceToDevice, ctx->compute));
  } else {
    CU(ctx->d_keys.reserve((nj + nd) * 8)); CU(ctx->d_keys_sorted.reserve((nj + nd) * 8)); CU(ctx->d_count.reserve(64));
    uint64_t* cat = (uint64_t*)ctx->d_keys.p;
    if (nj) CU(cudaMemcpyAsync(cat, ctx->d_skeys_j.p, nj * 8, cudaMemcpyDeviceToDevice, ctx->compute));
    clear_bit0_kernel<<<grid_for(nd, 256), 256, 0, ctx->compute>>>((const uint64_t*)ctx->d_skeys_d.p, nd, cat + nj);
    size_t tmp = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, tmp, (const uint64_t*)cat, (uint64_t*)ctx->d_keys_sorted.p, (int)(nj + nd), 0, 64, ctx->compute);
    CU(ctx->d_cub_tmp.reserve(tmp + 16));
    CU(cub::DeviceRadixSort::SortKeys(ctx->d_cub_tmp.p, tmp, (const uint64_t*)cat, (uint64_t*)ctx->d_keys_sorted.p, (int)(nj + nd), 0, 64, ctx->compute));
    tmp = 0;
    cub::DeviceSelect::Unique(nullptr, tmp, (const uint64_t*)ctx->d_keys_sorted.p, cat, (unsigned long long*)ctx->d_count.p, (int)(nj + nd), ctx->compute);
    CU(ctx->d_cub_tmp.reserve(tmp + 16));
    CU(cub::DeviceSelect::Unique(ctx->d_cub_tmp.p, tmp, (const uint64_t*)ctx->d_keys_sorted.p, cat, (unsigned long long*)ctx->d_count.p, (int)(nj + nd), ctx->compute));
    unsigned long long nu = 0;
    CU(cudaMemcpyAsync(&nu, ctx->d_count.p, 8, cudaMemcpyDeviceToHost, ctx->compute));
    CU(cudaStreamSynchronize(ctx->compute));
    n_juncs = nu;
    decode_keys_kernel<<<grid_for(nu, 256), 256, 0, ctx->compute>>>((const uint64_t*)cat, nu, ctx->ref, (thb_junction*)ctx->j_juncs.p);
    CU(cudaGetLastError());
  }
  if (ni) CU(cudaMemcpyAsync(ctx->j_ins.p, ctx->d_ins_out.p, ni * sizeof(thb_insertion), cudaMemcpyDeviceToDevice, ctx->compute));
  return join_index_sets(ctx, p, n_juncs, ni);
}

static int join_index_sets(thb_ctx* ctx, const thb_params* p, uint64_t n_juncs, uint64_t n_ins)
{
  // The sets are validated where they now live (a pass over 1.4 M junctions of an hg38-sized run on the host cost more than
  // the join kernels): strictly increasing in Junction / Insertion order; and, for the bucket index, monotone global lefts
  // (every junction inside its contig's slot of the image -- otherwise the kernels binary-search).
  // Bucket size: a power of two >= 64 bases giving about four buckets per junction -- on an hg38-sized image 64-base buckets
  // would be 48 M entries, rebuilt per thb_join_begin.
  const uint64_t blocks = ctx->d_nmask.cap / 8;     // 64-base blocks allocated for the image (>= n_blocks)
  int shift = 6;
  while (shift < 20 && ((blocks << 6) >> shift) > 4 * std::max<uint64_t>(n_juncs, 1u << 16)) ++shift;
  const uint64_t nb = ((blocks << 6) >> shift) + 1;
  ctx->j_shift = shift;
  int ishift = 6;
  while (ishift < 20 && ((blocks << 6) >> ishift) > 4 * std::max<uint64_t>(n_ins, 1u << 16)) ++ishift;
  const uint64_t nib = ((blocks << 6) >> ishift) + 1;
  ctx->j_ishift = ishift;
  CU(ctx->d_count.reserve(64));
  CU(cudaMemsetAsync(ctx->d_count.p, 0, 16, ctx->compute));
  unsigned int* vflags = (unsigned int*)ctx->d_count.p;            // [0] bit0 junctions unsorted, bit1 insertions unsorted, bit2 not monotone; [1] first bad index
  if (n_juncs || n_ins)
    join_sets_check_kernel<<<grid_for(std::max(n_juncs, n_ins), 256), 256, 0, ctx->compute>>>((const thb_junction*)ctx->j_juncs.p, n_juncs, (const thb_insertion*)ctx->j_ins.p, n_ins,
                                                                                              ctx->ref, nb, shift, nib, ishift, vflags);
  unsigned int hflags[2] = {0, 0};
  CU(cudaMemcpyAsync(hflags, vflags, 8, cudaMemcpyDeviceToHost, ctx->compute));
  CU(cudaStreamSynchronize(ctx->compute));
  if (hflags[0] & 1u) return fail(ctx, THB_EINVAL, "junction set not sorted / unique near %u", hflags[1]);
  if (hflags[0] & 2u) return fail(ctx, THB_EINVAL, "insertion set not sorted / unique near %u", hflags[1]);
  {
    const bool mono = n_juncs > 0 && !(hflags[0] & 4u);
    ctx->j_use_idx = mono;
    if (mono) {
      CU(ctx->j_idx.reserve((nb + 1) * 4));
      junction_index_kernel<thb_junction><<<grid_for(nb + 1, 256), 256, 0, ctx->compute>>>((const thb_junction*)ctx->j_juncs.p, (uint32_t)n_juncs, ctx->ref.contig_start,
                                                                                         (uint32_t*)ctx->j_idx.p, nb, shift);
      CU(cudaGetLastError());
      ctx->j_nbuckets = nb;
    }
    const bool imono = n_ins > 0 && !(hflags[0] & 8u);
    ctx->j_use_iidx = imono;
    if (imono) {
      CU(ctx->j_iidx.reserve((nib + 1) * 4));
      junction_index_kernel<thb_insertion><<<grid_for(nib + 1, 256), 256, 0, ctx->compute>>>((const thb_insertion*)ctx->j_ins.p, (uint32_t)n_ins, ctx->ref.contig_start,
                                                                                           (uint32_t*)ctx->j_iidx.p, nib, ishift);
      CU(cudaGetLastError());
      ctx->j_nibuckets = nib;
    }
  }
  CU(cudaEventRecord(ctx->ev_b, ctx->compute));
  CU(cudaStreamSynchronize(ctx->compute));
  float begin_ms = 0.f; cudaEventElapsedTime(&begin_ms, ctx->ev_a, ctx->ev_b);
  ctx->j_n_juncs = n_juncs; ctx->j_n_ins = n_ins;
  JoinParams& j = ctx->jp;
  j.max_ins = p->max_insertion_length; j.max_del = p->max_deletion_length; j.min_report_intron = p->min_report_intron_length;
  j.max_report_intron = p->max_report_intron_length; j.fusion_min_dist = p->fusion_min_dist; j.max_seg_multihits = p->max_seg_multihits;
  j.bowtie2 = p->bowtie2; j.seglen = p->segment_length;
  ctx->params = *p; ctx->join_begun = true; ctx->j_n_fus = 0;
  memset(&ctx->jtiming, 0, sizeof ctx->jtiming);
  ctx->jtiming.begin_ms = begin_ms;
  return THB_OK;
}

static int join_validate(thb_ctx* ctx, const thb_join_batch* b)
{
  if (!ctx->join_begun) return fail(ctx, THB_ESTATE, "thb_join_begin not called");
  if (b->n_segs < 1 || b->n_segs > (uint32_t)JMAXSEGS) return fail(ctx, THB_EUNSUPPORTED, "n_segs %u outside [1,%d]", b->n_segs, JMAXSEGS);
  if (b->read_words < 1 || b->read_words > 4) return fail(ctx, THB_EUNSUPPORTED, "read_words %u outside [1,4]", b->read_words);
  if (!b->bundles || !b->seg_count || !b->reads || (b->n_hits && !b->hits) || (b->n_ops_ext && !b->ops_ext)) return fail(ctx, THB_EINVAL, "null batch array");
  if (b->n_hits >= (1ull << 32)) return fail(ctx, THB_EUNSUPPORTED, "more than 2^32 segment hits in one batch (hit_begin is 32 bits)");
  return THB_OK;
}

// launches the chain join over device-resident arrays; results stay in ctx->j_out, *n receives their number
static int join_run(thb_ctx* ctx, const JoinBatchView& bv, uint64_t n_hits, uint64_t n_ext, unsigned long long* n_res, DevBuf& out_buf, uint64_t& cap_out)
{
  cap_out = std::max<uint64_t>(cap_out, tiny_caps() ? 64 : std::max<uint64_t>(2ull * bv.n_bundles, 1u << 16));
  ctx->j_cap_chain = std::max<uint64_t>(ctx->j_cap_chain, tiny_caps() ? 64 : std::max<uint64_t>(2ull * bv.n_bundles, 1u << 16));
  JoinSets S; S.juncs = (const thb_junction*)ctx->j_juncs.p; S.n_juncs = (uint32_t)ctx->j_n_juncs; S.ins = (const thb_insertion*)ctx->j_ins.p; S.n_ins = (uint32_t)ctx->j_n_ins;
  S.jidx = ctx->j_use_idx ? (const uint32_t*)ctx->j_idx.p : nullptr; S.n_buckets = ctx->j_nbuckets; S.shift = ctx->j_shift;
  S.iidx = ctx->j_use_iidx ? (const uint32_t*)ctx->j_iidx.p : nullptr; S.n_ibuckets = ctx->j_nibuckets; S.ishift = ctx->j_ishift;
  unsigned long long n = 0; unsigned long long cnt[3] = {0, 0, 0}; float kms = 0.f; unsigned long long qn_simple = 0, qn_abut = 0;
  const uint32_t stride = bv.n_segs + 1;
  for (int attempt = 0; attempt < 24; ++attempt) {
    CU(out_buf.reserve(cap_out * sizeof(thb_joined)));
    // reads of up to four segments (2x101 bp at --segment-length 25) take the tile kernel; other layouts the queue kernels
    const bool legacy = join_legacy() || bv.n_segs > (uint32_t)JT_SEGS || !(bv.read_words == 2 || bv.read_words == 4);
    CU(ctx->j_chain.reserve((legacy ? 3 : 1) * ctx->j_cap_chain * stride * sizeof(uint32_t)));     // legacy: general + simple + abutting queues
    CU(cudaMemsetAsync(ctx->d_qcounts, 0, 4 * sizeof(unsigned long long), ctx->compute));
    CU(cudaMemsetAsync(ctx->d_qovf, 0, sizeof(unsigned int), ctx->compute));
    CU(cudaMemsetAsync(ctx->d_counters + 4, 0, 3 * sizeof(unsigned long long), ctx->compute));
    JoinOut o; o.rec = (thb_joined*)out_buf.p; o.cap = cap_out; o.count = ctx->d_qcounts; o.overflow = ctx->d_qovf; o.counters = ctx->d_counters + 4;
    ChainQueue q; q.tasks = (uint32_t*)ctx->j_chain.p; q.cap = ctx->j_cap_chain; q.stride = stride; q.count = ctx->d_qcounts + 1; q.overflow = ctx->d_qovf;
    q.simple_tasks = q.tasks + ctx->j_cap_chain * stride; q.simple_count = ctx->d_qcounts + 2;
    q.abut_tasks = q.tasks + 2 * ctx->j_cap_chain * stride; q.abut_count = ctx->d_qcounts + 3;
    CU(cudaMemsetAsync(ctx->d_counters + 18, 0, 2 * sizeof(unsigned long long), ctx->compute));
    CU(cudaEventRecord(ctx->kev[0], ctx->compute));
    if (ctx->params.fusion_search) {
      // --fusion-search: every read takes the fusion-aware walk + merge (fusion_join_kernel.cuh); no queues
      FJoinSets FS; FS.base = S; FS.fus = (const FusionKey*)ctx->j_fus.p; FS.n_fus = (uint32_t)ctx->j_n_fus;
      FJoinOut fo; fo.rec = o.rec; fo.cap = o.cap; fo.count = o.count; fo.overflow = o.overflow; fo.counters = o.counters;
      fusion_join_kernel<<<grid_for(bv.n_bundles, 64), 64, 0, ctx->compute>>>(ctx->ref, ctx->jp, FS, bv, fo);
      CU(cudaEventRecord(ctx->kev[1], ctx->compute));
      CU(cudaEventRecord(ctx->kev[3], ctx->compute)); CU(cudaEventRecord(ctx->kev[4], ctx->compute));
    } else if (legacy) {
      chain_enum_kernel<<<grid_for(bv.n_bundles, 256), 256, 0, ctx->compute>>>(ctx->jp, bv, q, ctx->d_counters + 4);
      CU(cudaEventRecord(ctx->kev[1], ctx->compute));
      chain_merge_simple_kernel<<<ctx->sms * 8, 256, 0, ctx->compute>>>(ctx->ref, ctx->jp, bv, q, o);
      CU(cudaEventRecord(ctx->kev[3], ctx->compute));
      chain_merge_abut_kernel<false, 8><<<ctx->sms * 16, 128, 0, ctx->compute>>>(ctx->ref, ctx->jp, bv, q, o);
      CU(cudaEventRecord(ctx->kev[4], ctx->compute));
    } else {
      // one pass over the batch: TMA-staged tiles of 32 reads, chain enumeration + the closure-free merges (join_tile_kernel.cuh)
      const uint32_t n_tiles = (bv.n_bundles + 31u) / 32u;
      const int grid = (int)std::min<uint64_t>((n_tiles + JT_WARPS - 1) / JT_WARPS, (uint64_t)ctx->sms * 64);
      if (bv.read_words == 2) join_tile_kernel<2, 6><<<grid, JT_WARPS * 32, 0, ctx->compute>>>(ctx->ref, ctx->jp, bv, q, o, ctx->d_counters + 18);
      else join_tile_kernel<4, 6><<<grid, JT_WARPS * 32, 0, ctx->compute>>>(ctx->ref, ctx->jp, bv, q, o, ctx->d_counters + 18);
      CU(cudaEventRecord(ctx->kev[1], ctx->compute));
      CU(cudaEventRecord(ctx->kev[3], ctx->compute)); CU(cudaEventRecord(ctx->kev[4], ctx->compute));
    }
    // launch bounds chosen by measurement on B200 (profiles/README.md): 80 registers for the closure kernel -- it is
    // latency-bound, so residency is traded against spills
    if (!ctx->params.fusion_search) chain_merge_kernel<6><<<ctx->sms * 16, 128, 0, ctx->compute>>>(ctx->ref, ctx->jp, S, bv, q, o);
    CU(cudaGetLastError());
    CU(cudaEventRecord(ctx->kev[2], ctx->compute));
    ctx->jtiming.launches += legacy ? 4 : 2;
    unsigned int ovf = 0; unsigned long long qn[4] = {0, 0, 0, 0}, tc[2] = {0, 0};
    CU(cudaMemcpyAsync(qn, ctx->d_qcounts, 32, cudaMemcpyDeviceToHost, ctx->compute));
    CU(cudaMemcpyAsync(tc, ctx->d_counters + 18, 16, cudaMemcpyDeviceToHost, ctx->compute));
    CU(cudaMemcpyAsync(&ovf, ctx->d_qovf, 4, cudaMemcpyDeviceToHost, ctx->compute));
    CU(cudaMemcpyAsync(cnt, ctx->d_counters + 4, sizeof cnt, cudaMemcpyDeviceToHost, ctx->compute));
    CU(cudaStreamSynchronize(ctx->compute));
    if (ovf & 2u) return fail(ctx, THB_EUNSUPPORTED, "a merged alignment needs more than %d CIGAR operations (outside the GPU path)", JMAXOPS);
    if (ovf & 4u) return fail(ctx, THB_EUNSUPPORTED, "a fusion alignment whose sequence is neither the read nor its reverse complement (outside the GPU path)");
    n = qn[0]; qn_simple = legacy ? qn[2] : tc[0]; qn_abut = legacy ? qn[3] : tc[1];
    if (!ovf && n <= cap_out && qn[1] <= ctx->j_cap_chain && qn[2] <= ctx->j_cap_chain && qn[3] <= ctx->j_cap_chain) {
      float a = 0.f, b2 = 0.f; CU(cudaEventElapsedTime(&a, ctx->kev[0], ctx->kev[1])); CU(cudaEventElapsedTime(&b2, ctx->kev[1], ctx->kev[2]));
      kms = a + b2; ctx->jtiming.enum_ms += a; ctx->jtiming.merge_ms += b2;
      float m1 = 0.f, m2 = 0.f, m3 = 0.f;
      cudaEventElapsedTime(&m1, ctx->kev[1], ctx->kev[3]); cudaEventElapsedTime(&m2, ctx->kev[3], ctx->kev[4]); cudaEventElapsedTime(&m3, ctx->kev[4], ctx->kev[2]);
      ctx->jtiming.merge_simple_ms += m1; ctx->jtiming.merge_abutting_ms += m2; ctx->jtiming.merge_general_ms += m3; break;
    }
    cap_out = std::max<uint64_t>(cap_out, 2 * n + 1024);
    ctx->j_cap_chain = std::max<uint64_t>(ctx->j_cap_chain, 2 * std::max(qn[1], std::max(qn[2], qn[3])) + 1024);
    if (attempt == 23) return fail(ctx, THB_ENOMEM, "joined-hit buffer still overflows");
  }
  thb_join_timing& t = ctx->jtiming;
  t.kernel_ms += kms; t.n_chains += cnt[0]; t.n_closures += cnt[1]; t.n_joined += cnt[2];
  t.n_simple_chains += qn_simple; t.n_abutting_chains += qn_abut;
  // SURVEY.md 8(d) B_join on this batch's actual counts: 16 (header) + 40 (read) per read, 16 per segment hit + 32 per hit
  // with a multi-op CIGAR, per closure 64 (set lookup) + 64 (reference), 96 for the consistency re-read per merged chain,
  // 128 per output record
  t.algorithmic_bytes += 56ull * bv.n_bundles + 16ull * n_hits + 32ull * n_ext + 128ull * cnt[1] + 96ull * cnt[0] + 128ull * n;
  ctx->j_last_n = n;
  *n_res = n;
  return THB_OK;
}

int thb_join_set_fusions(thb_ctx* ctx, const thb_fusion* fus, uint64_t n)
{
  if (!ctx || (n && !fus)) return THB_EINVAL;
  CU(cudaSetDevice(ctx->device));
  if (!ctx->join_begun) return fail(ctx, THB_ESTATE, "thb_join_begin not called");
  if (n >= (1ull << 31)) return fail(ctx, THB_EINVAL, "bad fusion set");
  std::vector<FusionKey> keys((size_t)n);
  for (uint64_t i = 0; i < n; ++i) {
    keys[i] = FusionKey{fus[i].ref_id1, fus[i].ref_id2, fus[i].left, fus[i].right, fus[i].dir};
    if (i) { const FusionKey& a = keys[i - 1]; const FusionKey& b = keys[i];
      const bool lt = a.r1 != b.r1 ? a.r1 < b.r1 : a.r2 != b.r2 ? a.r2 < b.r2 : a.left != b.left ? a.left < b.left : a.right != b.right ? a.right < b.right : a.dir < b.dir;
      if (!lt) return fail(ctx, THB_EINVAL, "fusion set not sorted / unique near %llu", (unsigned long long)i); }
  }
  CU(ctx->j_fus.reserve((n + 1) * sizeof(FusionKey)));
  if (n) CU(cudaMemcpyAsync(ctx->j_fus.p, keys.data(), n * sizeof(FusionKey), cudaMemcpyHostToDevice, ctx->compute));
  CU(cudaStreamSynchronize(ctx->compute));
  ctx->j_n_fus = n;
  return THB_OK;
}

// page-locked host buffer for joined records: at least `need` records, the first `keep` preserved
static int join_host_reserve(thb_ctx* ctx, uint64_t need, uint64_t keep)
{
  if (need <= ctx->h_joined_cap) return THB_OK;
  const uint64_t ncap = std::max<uint64_t>(need + need / 4, 1u << 16);
  thb_joined* nb = nullptr;
  if (cudaHostAlloc((void**)&nb, ncap * sizeof(thb_joined), cudaHostAllocDefault) != cudaSuccess) return fail(ctx, THB_ENOMEM, "cannot page-lock %llu joined records", (unsigned long long)ncap);
  if (keep) { CU(cudaStreamSynchronize(ctx->d2h)); memcpy(nb, ctx->h_joined, keep * sizeof(thb_joined)); }
  if (ctx->h_joined) cudaFreeHost(ctx->h_joined);
  ctx->h_joined = nb; ctx->h_joined_cap = ncap;
  return THB_OK;
}

int thb_join_fetch(thb_ctx* ctx, const thb_joined** out, uint64_t* n_out)
{
  if (!ctx || !out || !n_out) return THB_EINVAL;
  CU(cudaSetDevice(ctx->device));
  const unsigned long long n = ctx->j_last_n;
  int rc = join_host_reserve(ctx, n, 0); if (rc) return rc;
  CU(cudaEventRecord(ctx->ev_c, ctx->compute));
  if (n) CU(cudaMemcpyAsync(ctx->h_joined, ctx->j_out.p, n * sizeof(thb_joined), cudaMemcpyDeviceToHost, ctx->compute));
  CU(cudaEventRecord(ctx->ev_d, ctx->compute));
  CU(cudaStreamSynchronize(ctx->compute));
  float d2h = 0.f; cudaEventElapsedTime(&d2h, ctx->ev_c, ctx->ev_d); ctx->jtiming.d2h_ms += d2h;
  *out = ctx->h_joined; *n_out = n;
  return THB_OK;
}

int thb_join_submit_device(thb_ctx* ctx, const thb_join_batch* b, uint64_t* n_out)
{
  if (!ctx || !b || !n_out) return THB_EINVAL;
  CU(cudaSetDevice(ctx->device));
  *n_out = 0; ctx->j_last_n = 0;
  int rc = join_validate(ctx, b); if (rc) return rc;
  if (b->n_bundles == 0) return THB_OK;
  JoinBatchView bv; bv.bundles = b->bundles; bv.seg_count = b->seg_count; bv.reads = b->reads; bv.hits = b->hits; bv.ops_ext = b->ops_ext;
  bv.n_bundles = b->n_bundles; bv.n_segs = b->n_segs; bv.read_words = b->read_words; bv.bundle_base = 0; bv.hit_end = (uint32_t)b->n_hits; bv.ops_end = (uint32_t)b->n_ops_ext;
  unsigned long long n = 0;
  rc = join_run(ctx, bv, b->n_hits, b->n_ops_ext, &n, ctx->j_out, ctx->j_cap_out); if (rc) return rc;
  *n_out = n;
  return THB_OK;
}

// Host batch: a three-stream pipeline over chunks of reads.  While the kernels of chunk c run on the compute stream, the
// copy stream uploads chunk c+1 and the d2h stream downloads the records of chunk c-1 into the page-locked result buffer.
int thb_join_submit(thb_ctx* ctx, const thb_join_batch* b, const thb_joined** out, uint64_t* n_out)
{
  if (!ctx || !b || !out || !n_out) return THB_EINVAL;
  CU(cudaSetDevice(ctx->device));
  *out = nullptr; *n_out = 0; ctx->j_last_n = 0;
  int rc = join_validate(ctx, b); if (rc) return rc;
  if (b->n_bundles == 0) return THB_OK;
  const size_t rdw = (size_t)3 * b->read_words;
  uint32_t CH = 1u << 19;                                // reads per pipeline chunk (tuning knob: THB_JOIN_CHUNK_READS)
  if (const char* e = getenv("THB_JOIN_CHUNK_READS")) { const long v = atol(e); if (v >= 32 && v <= (1l << 26)) CH = (uint32_t)v; }
  const uint32_t nchunks = (b->n_bundles + CH - 1) / CH;
  struct Range { uint32_t b0, nb; uint64_t h0, h1, e0, e1; };
  auto range_of = [&](uint32_t c, Range* r) -> bool {
    r->b0 = c * CH; const uint32_t b1 = std::min<uint32_t>(b->n_bundles, r->b0 + CH); r->nb = b1 - r->b0;
    r->h0 = b->bundles[r->b0].hit_begin; r->h1 = (b1 < b->n_bundles) ? b->bundles[b1].hit_begin : b->n_hits;
    r->e0 = b->bundles[r->b0].ops_begin; r->e1 = (b1 < b->n_bundles) ? b->bundles[b1].ops_begin : b->n_ops_ext;
    return !(r->h1 < r->h0 || r->h1 > b->n_hits || r->e1 < r->e0 || r->e1 > b->n_ops_ext);
  };
  auto enqueue_copy = [&](uint32_t c) -> int {
    JStage& s = ctx->jstage[c & 1]; Range r;
    if (!range_of(c, &r)) return fail(ctx, THB_EINVAL, "hit_begin / ops_begin not monotonic near bundle %u", c * CH);
    CU(s.bundles.reserve((size_t)r.nb * sizeof(thb_join_bundle))); CU(s.seg_count.reserve((size_t)r.nb * b->n_segs * 2));
    CU(s.reads.reserve((size_t)r.nb * rdw * 8)); CU(s.hits.reserve((size_t)(r.h1 - r.h0 + 1) * sizeof(thb_jhit)));
    CU(s.ops.reserve((size_t)(r.e1 - r.e0 + 1) * sizeof(thb_jops)));
    if (r.e1 > r.e0) CU(cudaMemcpyAsync(s.ops.p, b->ops_ext + r.e0, (size_t)(r.e1 - r.e0) * sizeof(thb_jops), cudaMemcpyHostToDevice, ctx->copy));
    CU(cudaMemcpyAsync(s.bundles.p, b->bundles + r.b0, (size_t)r.nb * sizeof(thb_join_bundle), cudaMemcpyHostToDevice, ctx->copy));
    CU(cudaMemcpyAsync(s.seg_count.p, b->seg_count + (size_t)r.b0 * b->n_segs, (size_t)r.nb * b->n_segs * 2, cudaMemcpyHostToDevice, ctx->copy));
    CU(cudaMemcpyAsync(s.reads.p, b->reads + (size_t)r.b0 * rdw, (size_t)r.nb * rdw * 8, cudaMemcpyHostToDevice, ctx->copy));
    if (r.h1 > r.h0) CU(cudaMemcpyAsync(s.hits.p, b->hits + r.h0, (size_t)(r.h1 - r.h0) * sizeof(thb_jhit), cudaMemcpyHostToDevice, ctx->copy));
    CU(cudaEventRecord(s.copied, ctx->copy));
    return THB_OK;
  };
  rc = join_host_reserve(ctx, (uint64_t)b->n_bundles + b->n_bundles / 2, 0); if (rc) return rc;
  CU(cudaEventRecord(ctx->ev_a, ctx->compute));
  rc = enqueue_copy(0); if (rc) return rc;
  uint64_t total = 0;
  for (uint32_t c = 0; c < nchunks; ++c) {
    JStage& s = ctx->jstage[c & 1]; Range r; range_of(c, &r);
    CU(cudaStreamWaitEvent(ctx->compute, s.copied, 0));
    // staging buffer (c+1)&1 was last read by the kernels of chunk c-1, which have completed (join_run synchronises)
    if (c + 1 < nchunks) { rc = enqueue_copy(c + 1); if (rc) return rc; }
    if (s.out_busy) CU(cudaStreamWaitEvent(ctx->compute, s.out_free, 0));      // records of chunk c-2 have left this buffer
    JoinBatchView bv; bv.bundles = (const thb_join_bundle*)s.bundles.p; bv.seg_count = (const uint16_t*)s.seg_count.p; bv.reads = (const uint64_t*)s.reads.p;
    bv.hits = (const thb_jhit*)s.hits.p - r.h0; bv.ops_ext = (const thb_jops*)s.ops.p - r.e0; bv.n_bundles = r.nb; bv.n_segs = b->n_segs; bv.read_words = b->read_words; bv.bundle_base = r.b0; bv.hit_end = (uint32_t)r.h1; bv.ops_end = (uint32_t)r.e1;
    unsigned long long n = 0;
    rc = join_run(ctx, bv, r.h1 - r.h0, r.e1 - r.e0, &n, s.out, s.cap_out); if (rc) return rc;       // returns with the compute stream idle
    rc = join_host_reserve(ctx, total + n, total); if (rc) return rc;
    if (n) CU(cudaMemcpyAsync(ctx->h_joined + total, s.out.p, n * sizeof(thb_joined), cudaMemcpyDeviceToHost, ctx->d2h));
    CU(cudaEventRecord(s.out_free, ctx->d2h)); s.out_busy = true;
    total += n;
  }
  CU(cudaStreamSynchronize(ctx->d2h));
  CU(cudaEventRecord(ctx->ev_b, ctx->compute));
  CU(cudaStreamSynchronize(ctx->compute));
  float tot_ms = 0.f; cudaEventElapsedTime(&tot_ms, ctx->ev_a, ctx->ev_b);
  ctx->jtiming.h2d_ms += tot_ms;          // wall time of the pipelined submit (copies, kernels and downloads overlap)
  ctx->j_last_n = 0;                      // nothing left on the device for thb_join_fetch
  *out = ctx->h_joined; *n_out = total;
  return THB_OK;
}

int thb_join_last_timing(thb_ctx* ctx, thb_join_timing* out)
{
  if (!ctx || !out) return THB_EINVAL;
  *out = ctx->jtiming;
  return THB_OK;
}

// ---- junction-flank matcher -------------------------------------------------------------------------------------------------
extern "C++" {
namespace {

// the flank arithmetic of print_splice / print_insertion / print_fusion (juncs_db.cpp:72-229); false = juncs_db prints nothing
struct FlankGeom { uint64_t ls, le, rs, re; bool rc_a, rc_b; };

bool flank_splice(uint64_t n, uint64_t left, uint64_t right, int half, FlankGeom& g)
{
  if (!(left <= n && right <= n)) return false;
  g.ls = (int64_t)left - half + 1 >= 0 ? left - half + 1 : 0; g.le = g.ls + half;
  g.rs = right; g.re = g.rs + half < n ? g.rs + half : n;
  g.rc_a = g.rc_b = false;
  return g.ls < g.le && g.le <= n && g.rs < g.re && g.re <= n;
}
bool flank_insertion(uint64_t n, uint64_t left, int half, FlankGeom& g)
{
  if (!(left <= n) || half <= 0) return false;
  g.ls = (int64_t)left - half + 1 >= 0 ? left - half + 1 : 0; g.le = g.ls + half;
  g.rs = g.le; g.re = g.rs + half < n ? g.rs + half : n;
  g.rc_a = g.rc_b = false;
  return g.ls < g.le && g.le <= n && g.rs < g.re && g.re <= n;
}
bool flank_fusion(uint64_t n1, uint64_t n2, uint64_t left, uint64_t right, uint32_t dir, int half, FlankGeom& g)
{
  if (!(left < n1 && right < n2) || half <= 0) return false;
  if (dir == 7u || dir == 8u) { g.ls = left + 1 >= (uint64_t)half ? left - half + 1 : 0; g.le = g.ls + half; }
  else { g.ls = left; g.le = g.ls + half < n1 ? g.ls + half : n1; }
  if (dir == 7u || dir == 9u) { g.rs = right; g.re = g.rs + half < n2 ? g.rs + half : n2; }
  else { g.re = right + 1; g.rs = g.re >= (uint64_t)half ? g.re - half : 0; }
  g.rc_a = dir == 9u || dir == 10u; g.rc_b = dir == 8u || dir == 10u;
  return g.ls < g.le && g.le <= n1 && g.rs < g.re && g.re <= n2;
}

template <int CW>
int flank_build_index(thb_ctx* ctx, uint64_t n_contigs, uint64_t n_entries, int sort_bits, uint32_t n_keys)
{
  FlankState& f = ctx->fl;
  flank_build_kernel<CW><<<grid_for(n_contigs, 128), 128, 0, ctx->compute>>>(ctx->ref, (const FlankDesc*)f.desc.p, (FlankSeq<CW>*)f.seq.p, (uint32_t)n_contigs);
  CU(cudaGetLastError());
  flank_seed_kernel<CW><<<(unsigned)std::min<uint64_t>(n_contigs, (uint64_t)ctx->sms * 64), 64, 0, ctx->compute>>>(
      (const FlankSeq<CW>*)f.seq.p, (const uint64_t*)f.base.p, (uint32_t)n_contigs, f.ip, (uint32_t*)f.keys.p, (uint64_t*)f.vals.p);
  CU(cudaGetLastError());
  size_t tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp, (const uint32_t*)f.keys.p, (uint32_t*)f.keys2.p, (const uint64_t*)f.vals.p, (uint64_t*)f.vals2.p, (int)n_entries, 0, sort_bits, ctx->compute);
  CU(ctx->d_cub_tmp.reserve(tmp + 16));
  CU(cub::DeviceRadixSort::SortPairs(ctx->d_cub_tmp.p, tmp, (const uint32_t*)f.keys.p, (uint32_t*)f.keys2.p, (const uint64_t*)f.vals.p, (uint64_t*)f.vals2.p, (int)n_entries, 0, sort_bits, ctx->compute));
  flank_bucket_kernel<<<grid_for(n_entries + 1, 256), 256, 0, ctx->compute>>>((const uint32_t*)f.keys2.p, n_entries, n_keys, (uint32_t*)f.start.p);
  CU(cudaGetLastError());
  flank_fill_kernel<<<grid_for(n_keys, 256), 256, 0, ctx->compute>>>((const uint32_t*)f.start.p, (const uint64_t*)f.vals2.p, n_keys, (FlankBucket*)f.buckets.p);
  CU(cudaGetLastError());
  f.timing.launches = 5;
  return THB_OK;
}

}  // namespace
}  // extern "C++"

int thb_flank_begin(thb_ctx* ctx, const thb_flank_params* P,
                    const thb_junction* junctions, uint64_t n_junctions, const thb_junction* deletions, uint64_t n_deletions,
                    const thb_insertion* insertions, uint64_t n_insertions, const thb_fusion* fusions, uint64_t n_fusions)
{
  if (!ctx || !P) return THB_EINVAL;
  if (!ctx->have_ref) return fail(ctx, THB_ESTATE, "thb_flank_begin before thb_ref_upload");
  if ((n_junctions && !junctions) || (n_deletions && !deletions) || (n_insertions && !insertions) || (n_fusions && !fusions))
    return fail(ctx, THB_EINVAL, "thb_flank_begin: null set with a non-zero count");
  if (P->max_mismatches < 0 || P->max_mismatches > 3) return fail(ctx, THB_EUNSUPPORTED, "thb_flank_begin: max_mismatches %d outside 0..3 (tophat.py:2286 clamps --segment-mismatches to 3)", P->max_mismatches);
  if (P->max_multihits < 1) return fail(ctx, THB_EINVAL, "thb_flank_begin: max_multihits %d", P->max_multihits);
  const int npieces = P->max_mismatches + 2;
  if (P->min_seg_len > P->max_seg_len || P->min_seg_len < 4 * npieces || P->max_seg_len > 56)
    return fail(ctx, THB_EUNSUPPORTED, "thb_flank_begin: segment lengths %d..%d outside %d..56", P->min_seg_len, P->max_seg_len, 4 * npieces);
  if (P->min_anchor < 0 || P->min_anchor >= P->max_seg_len) return fail(ctx, THB_EINVAL, "thb_flank_begin: min_anchor %d", P->min_anchor);
  CU(cudaSetDevice(ctx->device));
  FlankState& f = ctx->fl;
  f.begun = false; f.have_last = false; f.contigs.clear(); f.timing = thb_flank_timing{};
  f.min_seg_len = P->min_seg_len; f.max_seg_len = P->max_seg_len;
  const int h = P->max_seg_len;
  std::vector<FlankDesc> desc; std::vector<uint64_t> base;
  desc.reserve(n_junctions + n_deletions + n_insertions + n_fusions); base.reserve(desc.capacity() + 1);
  const uint32_t nref = (uint32_t)ctx->h_clen.size();
  uint64_t n_off = 0; int max_len = 0;
  auto ref_len = [&](uint32_t id) -> uint64_t { return (id >= 1 && id <= nref) ? ctx->h_clen[id - 1] : 0; };    // 0: no sequence (rt.get_seq == NULL)
  auto push = [&](const FlankGeom& g, uint32_t r1, uint32_t r2, uint64_t ins_code, int ins_len, thb_flank_contig c) {
    FlankDesc d{};
    d.a_start = ctx->h_cstart[r1 - 1] + g.ls; d.b_start = ctx->h_cstart[r2 - 1] + g.rs;
    d.a_len = (uint8_t)(g.le - g.ls); d.b_len = (uint8_t)(g.re - g.rs); d.ins_len = (uint8_t)ins_len; d.ins_code = ins_code;
    d.flags = (uint8_t)((g.rc_a ? 1 : 0) | (g.rc_b ? 2 : 0));
    const int len = d.a_len + d.ins_len + d.b_len;
    c.length = (uint32_t)len; c.ref_id = r1; c.ref_id2 = r2;
    base.push_back(n_off);
    if (len >= P->min_seg_len) n_off += (uint64_t)(len - P->min_seg_len + 1);
    if (len > max_len) max_len = len;
    desc.push_back(d); f.contigs.push_back(c);
  };
  for (int pass = 0; pass < 2; ++pass) {          // junctions, then deletions read back as junctions (juncs_db.cpp:481-503)
    const thb_junction* set = pass ? deletions : junctions; const uint64_t n = pass ? n_deletions : n_junctions;
    for (uint64_t i = 0; i < n; ++i) {
      const thb_junction& j = set[i]; const uint64_t rl = ref_len(j.ref_id); FlankGeom g;
      if (!rl || !flank_splice(rl, j.left, j.right, h, g)) continue;
      thb_flank_contig c{}; c.kind = pass ? THB_FLANK_DEL : THB_FLANK_JUNC; c.left_start = (uint32_t)g.ls; c.left = j.left; c.right = j.right;
      c.right_end = (uint32_t)g.re; c.aux = pass ? 0u : (j.antisense ? 1u : 0u);
      push(g, j.ref_id, j.ref_id, 0, 0, c);
    }
  }
  for (uint64_t i = 0; i < n_insertions; ++i) {
    const thb_insertion& in = insertions[i]; const uint64_t rl = ref_len(in.ref_id); FlankGeom g;
    if (!rl || in.len == 0 || in.len > 19) continue;
    uint64_t code = 0; bool amb = false;
    for (uint32_t k = 0; k < in.len; ++k) {
      uint64_t cd;
      switch (in.seq[k]) { case 'A': cd = 0; break; case 'C': cd = 1; break; case 'G': cd = 2; break; case 'T': cd = 3; break; default: cd = 0; amb = true; }
      code |= cd << (2 * k);
    }
    if (amb || !flank_insertion(rl, in.left, h - P->min_anchor, g)) continue;       // juncs_db.cpp:418-430: no ambiguity in an insertion
    thb_flank_contig c{}; c.kind = THB_FLANK_INS; c.left_start = (uint32_t)g.ls; c.left = in.left; c.right = 0; c.right_end = (uint32_t)g.re; c.aux = in.len;
    memcpy(c.ins_seq, in.seq, in.len); c.ins_seq[in.len] = 0;
    push(g, in.ref_id, in.ref_id, code, (int)in.len, c);
  }
  for (uint64_t i = 0; i < n_fusions; ++i) {
    const thb_fusion& fu = fusions[i]; const uint64_t r1 = ref_len(fu.ref_id1), r2 = ref_len(fu.ref_id2); FlankGeom g;
    if (!r1 || !r2 || fu.dir < 7u || fu.dir > 10u || !flank_fusion(r1, r2, fu.left, fu.right, fu.dir, h - P->min_anchor, g)) continue;
    thb_flank_contig c{}; c.kind = THB_FLANK_FUS; c.left = fu.left; c.right = fu.right; c.aux = fu.dir;
    c.left_start = (uint32_t)(g.rc_a ? g.le - 1 : g.ls); c.right_end = (uint32_t)(g.rc_b ? g.rs - 1 : g.re);      // juncs_db.cpp:205-215
    push(g, fu.ref_id1, fu.ref_id2, 0, 0, c);
  }
  base.push_back(n_off);
  const uint64_t nc = desc.size();
  if (nc >= (1ull << 25)) return fail(ctx, THB_EUNSUPPORTED, "thb_flank_begin: %llu contigs (limit 2^25)", (unsigned long long)nc);
  FlankIndexParams ip{};
  ip.npieces = npieces; ip.piece_len = P->min_seg_len / npieces; if (ip.piece_len > 16) ip.piece_len = 16;
  ip.npairs = npieces * (npieces - 1) / 2; ip.smin = P->min_seg_len;
  ip.max_mm = P->max_mismatches; ip.max_hits = P->max_multihits; ip.ref_n_mismatch = P->ref_n_is_mismatch ? 1 : 0;
  // ~4 entries per 64-byte bucket record on average (7 fit inline); 2^23 buckets per pair at most (0.5 GB of records per pair)
  static const int bb_max = getenv("THB_FLANK_BBITS") ? atoi(getenv("THB_FLANK_BBITS")) : 23;      // measurement knob
  int bb = 8; while (bb < bb_max && (4ull << bb) < n_off) ++bb;
  if (4 * ip.piece_len <= bb) { ip.bbits = 4 * ip.piece_len; ip.hashed = 0; } else { ip.bbits = bb; ip.hashed = 1; }
  ip.fp_bases = P->max_mismatches ? std::min(ip.piece_len, 16 / P->max_mismatches) : 0;
  f.ip = ip; f.cw = max_len > 64 ? 2 : 1; f.n_contigs = nc;
  const uint64_t n_entries = n_off * (uint64_t)ip.npairs;
  if (n_entries >= 0x7fffffffull) return fail(ctx, THB_EUNSUPPORTED, "thb_flank_begin: %llu index entries (limit 2^31)", (unsigned long long)n_entries);
  f.n_entries = n_entries;
  f.timing.n_contigs = nc; f.timing.n_index_entries = n_entries;
  CU(f.scalars.reserve(64));
  if (nc == 0 || n_entries == 0) { f.begun = true; return THB_OK; }
  const uint32_t n_keys = (uint32_t)ip.npairs << ip.bbits;
  int sort_bits = ip.bbits; while ((1u << (sort_bits - ip.bbits)) <= (uint32_t)ip.npairs) ++sort_bits;
  CU(f.desc.reserve(nc * sizeof(FlankDesc))); CU(f.base.reserve((nc + 1) * 8));
  CU(f.seq.reserve(nc * (f.cw == 2 ? sizeof(FlankSeq<2>) : sizeof(FlankSeq<1>))));
  CU(f.keys.reserve(n_entries * 4)); CU(f.vals.reserve(n_entries * 8)); CU(f.keys2.reserve(n_entries * 4)); CU(f.vals2.reserve(n_entries * 8));
  CU(f.start.reserve(((uint64_t)n_keys + 2) * 4)); CU(f.buckets.reserve((uint64_t)n_keys * sizeof(FlankBucket)));
  CU(cudaEventRecord(ctx->ev_a, ctx->compute));
  CU(cudaMemcpyAsync(f.desc.p, desc.data(), nc * sizeof(FlankDesc), cudaMemcpyHostToDevice, ctx->compute));
  CU(cudaMemcpyAsync(f.base.p, base.data(), (nc + 1) * 8, cudaMemcpyHostToDevice, ctx->compute));
  CU(f.cdesc.reserve(nc * sizeof(thb_flank_contig)));
  CU(cudaMemcpyAsync(f.cdesc.p, f.contigs.data(), nc * sizeof(thb_flank_contig), cudaMemcpyHostToDevice, ctx->compute));
  int rc = f.cw == 2 ? flank_build_index<2>(ctx, nc, n_entries, sort_bits, n_keys) : flank_build_index<1>(ctx, nc, n_entries, sort_bits, n_keys);
  if (rc != THB_OK) return rc;
  CU(cudaEventRecord(ctx->ev_b, ctx->compute));
  CU(cudaStreamSynchronize(ctx->compute));      // desc / base are host vectors of this call
  CU(cudaEventElapsedTime(&f.timing.index_ms, ctx->ev_a, ctx->ev_b));
  f.keys.release(); f.vals.release(); f.keys2.release(); f.desc.release(); f.base.release(); f.start.release();      // the search needs seq, the bucket records and the sorted entries (overflow)
  f.begun = true;
  return THB_OK;
}

int thb_flank_contigs(thb_ctx* ctx, const thb_flank_contig** contigs, uint64_t* n_contigs)
{
  if (!ctx || !contigs || !n_contigs) return THB_EINVAL;
  if (!ctx->fl.begun) return fail(ctx, THB_ESTATE, "thb_flank_contigs before thb_flank_begin");
  *contigs = ctx->fl.contigs.data(); *n_contigs = ctx->fl.contigs.size();
  return THB_OK;
}

static int flank_submit(thb_ctx* ctx, const thb_flank_batch* b, bool on_device, const thb_flank_hit** hits, uint64_t* n_hits)
{
  if (!ctx || !b || !hits || !n_hits) return THB_EINVAL;
  FlankState& f = ctx->fl;
  if (!f.begun) return fail(ctx, THB_ESTATE, "thb_flank_submit before thb_flank_begin");
  *hits = nullptr; *n_hits = 0;
  if (b->n_segs < 1 || b->n_segs > THB_MAX_SEGS || b->read_words < 1 || b->read_words > 4) return fail(ctx, THB_EINVAL, "thb_flank_submit: n_segs %u / read_words %u", b->n_segs, b->read_words);
  if (b->n_reads >= (1u << 27)) return fail(ctx, THB_EUNSUPPORTED, "thb_flank_submit: %u reads in one batch (limit 2^27)", b->n_reads);
  if (b->n_reads && !b->reads) return fail(ctx, THB_EINVAL, "thb_flank_submit: null reads");
  for (uint32_t k = 0; k < b->n_segs; ++k) {
    const int s = (int)b->seg_bounds[k + 1] - (int)b->seg_bounds[k];
    if (s < f.min_seg_len || s > 64 || b->seg_bounds[k + 1] > 64u * b->read_words)
      return fail(ctx, THB_EINVAL, "thb_flank_submit: segment %u has %d bases (index seeds need >= %d, a segment is at most 64) or ends past the read", k, s, f.min_seg_len);
  }
  CU(cudaSetDevice(ctx->device));
  const uint32_t launches0 = f.timing.launches;
  f.timing.h2d_ms = f.timing.match_ms = f.timing.post_ms = f.timing.d2h_ms = 0; f.timing.n_verified = f.timing.n_hits = 0;
  f.hits.clear(); f.have_last = false;
  if (b->n_reads == 0 || f.n_contigs == 0 || f.n_entries == 0) { *hits = f.hits.data(); f.have_last = true; f.last_n = 0; return THB_OK; }
  const size_t rbytes = (size_t)b->n_reads * 3 * b->read_words * 8;
  CU(cudaEventRecord(ctx->ev_a, ctx->compute));
  const uint64_t* d_reads = b->reads;
  if (!on_device) {
    CU(f.reads.reserve(rbytes));
    CU(cudaMemcpyAsync(f.reads.p, b->reads, rbytes, cudaMemcpyHostToDevice, ctx->compute));
    d_reads = (const uint64_t*)f.reads.p;
  }
  CU(cudaEventRecord(ctx->ev_b, ctx->compute));
  const uint64_t nsegs_total = (uint64_t)b->n_reads * b->n_segs;
  CU(f.per_seg.reserve(nsegs_total * 4));
  if (f.cap_hits == 0) f.cap_hits = tiny_caps() ? 64 : std::max<uint64_t>(1u << 16, nsegs_total / 2);
  FlankBatchView bv{}; bv.reads = d_reads; bv.n_reads = b->n_reads; bv.read_words = b->read_words; bv.n_segs = b->n_segs;
  for (uint32_t k = 0; k <= b->n_segs; ++k) bv.seg_bounds[k] = b->seg_bounds[k];
  unsigned long long* sc = (unsigned long long*)f.scalars.p;      // [0] appended, [1] kept, [2] verified
  unsigned long long counts[3] = {0, 0, 0};
  const uint64_t threads = nsegs_total * 2u * (uint64_t)f.ip.npairs;
  for (;;) {
    CU(f.hkeys.reserve(f.cap_hits * 8)); CU(f.hmm.reserve(f.cap_hits * 4));
    CU(cudaMemsetAsync(f.per_seg.p, 0, nsegs_total * 4, ctx->compute));
    CU(cudaMemsetAsync(sc, 0, 24, ctx->compute));
    FlankOut o{}; o.keys = (uint64_t*)f.hkeys.p; o.mm = (uint32_t*)f.hmm.p; o.count = sc; o.cap = f.cap_hits; o.per_seg = (uint32_t*)f.per_seg.p; o.n_verified = sc + 2;
    const int mgrid = tiny_caps() ? 3 : grid_for(threads, 256);      // THB_TINY_CAPS: many grid-stride rounds per warp on small inputs
    if (f.cw == 2) flank_match_kernel<2><<<mgrid, 256, 0, ctx->compute>>>((const FlankSeq<2>*)f.seq.p, (const FlankBucket*)f.buckets.p, (const uint64_t*)f.vals2.p, f.ip, bv, o);
    else           flank_match_kernel<1><<<mgrid, 256, 0, ctx->compute>>>((const FlankSeq<1>*)f.seq.p, (const FlankBucket*)f.buckets.p, (const uint64_t*)f.vals2.p, f.ip, bv, o);
    CU(cudaGetLastError()); f.timing.launches++;
    CU(cudaMemcpyAsync(counts, sc, 24, cudaMemcpyDeviceToHost, ctx->compute));
    CU(cudaStreamSynchronize(ctx->compute));
    if (counts[0] <= f.cap_hits) break;
    while (f.cap_hits < counts[0]) f.cap_hits *= 2;               // the append buffer was too small: repeat the batch with a larger one
  }
  CU(cudaEventRecord(ctx->ev_c, ctx->compute));
  const uint64_t n_app = counts[0];
  uint64_t n_keep = 0;
  if (n_app) {
    CU(f.hkeys2.reserve(n_app * 8)); CU(f.hmm2.reserve(n_app * 4));
    flank_filter_kernel<<<grid_for(n_app, 256), 256, 0, ctx->compute>>>((const uint64_t*)f.hkeys.p, (const uint32_t*)f.hmm.p, n_app, (const uint32_t*)f.per_seg.p,
                                                                       b->n_segs, (uint32_t)f.ip.max_hits, (uint64_t*)f.hkeys2.p, (uint32_t*)f.hmm2.p, sc + 1);
    CU(cudaGetLastError()); f.timing.launches++;
    CU(cudaMemcpyAsync(counts + 1, sc + 1, 8, cudaMemcpyDeviceToHost, ctx->compute));
    CU(cudaStreamSynchronize(ctx->compute));
    n_keep = counts[1];
  }
  if (n_keep) {
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, (const uint64_t*)f.hkeys2.p, (uint64_t*)f.hkeys.p, (const uint32_t*)f.hmm2.p, (uint32_t*)f.hmm.p, (int)n_keep, 0, 64, ctx->compute);
    CU(ctx->d_cub_tmp.reserve(tmp + 16));
    CU(cub::DeviceRadixSort::SortPairs(ctx->d_cub_tmp.p, tmp, (const uint64_t*)f.hkeys2.p, (uint64_t*)f.hkeys.p, (const uint32_t*)f.hmm2.p, (uint32_t*)f.hmm.p, (int)n_keep, 0, 64, ctx->compute));
    CU(f.out.reserve(n_keep * sizeof(thb_flank_hit)));
    flank_decode_kernel<<<grid_for(n_keep, 256), 256, 0, ctx->compute>>>((const uint64_t*)f.hkeys.p, (const uint32_t*)f.hmm.p, n_keep, (FlankHitRec*)f.out.p);
    CU(cudaGetLastError()); f.timing.launches += 2;
  }
  CU(cudaEventRecord(ctx->ev_d, ctx->compute));
  CU(f.hits.resize(n_keep));
  if (n_keep) CU(cudaMemcpyAsync(f.hits.data(), f.out.p, n_keep * sizeof(thb_flank_hit), cudaMemcpyDeviceToHost, ctx->compute));
  CU(cudaEventRecord(ctx->kev[0], ctx->compute));
  CU(cudaStreamSynchronize(ctx->compute));
  CU(cudaEventElapsedTime(&f.timing.h2d_ms, ctx->ev_a, ctx->ev_b));
  CU(cudaEventElapsedTime(&f.timing.match_ms, ctx->ev_b, ctx->ev_c));
  CU(cudaEventElapsedTime(&f.timing.post_ms, ctx->ev_c, ctx->ev_d));
  CU(cudaEventElapsedTime(&f.timing.d2h_ms, ctx->ev_d, ctx->kev[0]));
  f.timing.n_verified = counts[2]; f.timing.n_hits = n_keep;
  // per segment and strand: the read words once; per proposed placement: its index entry and the contig's planes; per placement kept: the record
  f.timing.algorithmic_bytes = rbytes + threads * 8 + counts[2] * (4 + (f.cw == 2 ? 48 : 24)) + n_keep * sizeof(thb_flank_hit);
  (void)launches0;
  f.last_bv = bv; f.last_n = n_keep; f.have_last = true;
  *hits = f.hits.data(); *n_hits = n_keep;
  return THB_OK;
}

int thb_flank_submit(thb_ctx* ctx, const thb_flank_batch* b, const thb_flank_hit** hits, uint64_t* n_hits) { return flank_submit(ctx, b, false, hits, n_hits); }
int thb_flank_submit_device(thb_ctx* ctx, const thb_flank_batch* b, const thb_flank_hit** hits, uint64_t* n_hits) { return flank_submit(ctx, b, true, hits, n_hits); }

int thb_flank_spliced_hits(thb_ctx* ctx, int min_anchor_len, const thb_jhit_full** jhits, uint64_t* n)
{
  if (!ctx || !jhits || !n) return THB_EINVAL;
  FlankState& f = ctx->fl;
  if (!f.begun || !f.have_last) return fail(ctx, THB_ESTATE, "thb_flank_spliced_hits without a preceding thb_flank_submit");
  if (min_anchor_len < 0) return fail(ctx, THB_EINVAL, "thb_flank_spliced_hits: min_anchor_len %d", min_anchor_len);
  static_assert(sizeof(FlankJHit) == sizeof(thb_jhit_full) && sizeof(FlankContigDev) == sizeof(thb_flank_contig) && sizeof(FlankHitRec) == sizeof(thb_flank_hit), "wire records");
  CU(cudaSetDevice(ctx->device));
  *jhits = nullptr; *n = f.last_n;
  CU(f.jhits.resize(f.last_n));
  if (f.last_n == 0) { *jhits = f.jhits.data(); return THB_OK; }
  CU(f.jout.reserve(f.last_n * sizeof(thb_jhit_full)));
  if (f.cw == 2) flank_splice_kernel<2><<<grid_for(f.last_n, 128), 128, 0, ctx->compute>>>((const FlankSeq<2>*)f.seq.p, (const FlankContigDev*)f.cdesc.p, (const uint64_t*)f.hkeys.p, f.last_n, f.last_bv, min_anchor_len, f.ip.ref_n_mismatch, (FlankJHit*)f.jout.p);
  else           flank_splice_kernel<1><<<grid_for(f.last_n, 128), 128, 0, ctx->compute>>>((const FlankSeq<1>*)f.seq.p, (const FlankContigDev*)f.cdesc.p, (const uint64_t*)f.hkeys.p, f.last_n, f.last_bv, min_anchor_len, f.ip.ref_n_mismatch, (FlankJHit*)f.jout.p);
  CU(cudaGetLastError()); f.timing.launches++;
  CU(cudaMemcpyAsync(f.jhits.data(), f.jout.p, f.last_n * sizeof(thb_jhit_full), cudaMemcpyDeviceToHost, ctx->compute));
  CU(cudaStreamSynchronize(ctx->compute));
  *jhits = f.jhits.data();
  return THB_OK;
}

int thb_flank_last_timing(thb_ctx* ctx, thb_flank_timing* out)
{
  if (!ctx || !out) return THB_EINVAL;
  *out = ctx->fl.timing;
  return THB_OK;
}

// ---- multi-GPU exchange -------------------------------------------------------------------------
static int nccl_bind(thb_ctx* ctx, Nccl& n)
{
  if (n.h) return THB_OK;
  const char* names[] = { "libnccl.so.2", "libnccl.so" };
  for (const char* nm : names) { n.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (n.h) break; }
  if (!n.h) return fail(ctx, THB_ENCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
  n.GetUniqueId = (int (*)(void*))dlsym(n.h, "ncclGetUniqueId");
  *(void**)(&n.CommInitRank) = dlsym(n.h, "ncclCommInitRank");
  *(void**)(&n.AllGather) = dlsym(n.h, "ncclAllGather");
  n.CommDestroy = (int (*)(void*))dlsym(n.h, "ncclCommDestroy");
  n.GroupStart = (int (*)())dlsym(n.h, "ncclGroupStart");
  n.GroupEnd = (int (*)())dlsym(n.h, "ncclGroupEnd");
  n.GetErrorString = (const char* (*)(int))dlsym(n.h, "ncclGetErrorString");
  if (!n.GetUniqueId || !n.CommInitRank || !n.AllGather) return fail(ctx, THB_ENCCL, "libnccl lacks required symbols");
  return THB_OK;
}

int thb_nccl_unique_id(void* out128)
{
  static Nccl n; int rc = nccl_bind(nullptr, n); if (rc) return rc;
  return n.GetUniqueId(out128) == 0 ? THB_OK : THB_ENCCL;
}

int thb_comm_init(thb_ctx* ctx, const void* uid, int rank, int world)
{
  if (!ctx || !uid) return THB_EINVAL;
  CU(cudaSetDevice(ctx->device));
  int rc = nccl_bind(ctx, ctx->nccl); if (rc) return rc;
  Nccl::Uid u; memcpy(u.b, uid, 128);
  int e = ctx->nccl.CommInitRank(&ctx->comm, world, u, rank);
  if (e != 0) return fail(ctx, THB_ENCCL, "ncclCommInitRank: %s", ctx->nccl.GetErrorString ? ctx->nccl.GetErrorString(e) : "?");
  ctx->rank = rank; ctx->world = world;
  return THB_OK;
}

// Union of the per-rank sets (replaces the per-thread set union of segment_juncs.cpp:4911-4922).  One all-gather of
// the three set sizes, one host read of them, then one grouped all-gather of the padded payloads and device-side
// inserts; staging buffers persist in the context.
int thb_segjuncs_allgather(thb_ctx* ctx)
{
  if (!ctx) return THB_EINVAL;
  CU(cudaSetDevice(ctx->device));
  if (!ctx->comm) return fail(ctx, THB_ESTATE, "thb_comm_init not called");
  const int W = ctx->world;
  const uint64_t capj = ctx->cap_juncs, capd = ctx->cap_dels;
  CU(ctx->d_keys.reserve((capj + capd) * 8)); CU(ctx->d_count.reserve(8 * (size_t)(4 + 4 * W)));
  unsigned long long* d_cnt = (unsigned long long*)ctx->d_count.p;
  uint64_t* keys_j = (uint64_t*)ctx->d_keys.p; uint64_t* keys_d = keys_j + capj;
  CU(cudaMemsetAsync(d_cnt, 0, 32, ctx->compute));
  hs_compact_kernel<<<grid_for(capj, 256), 256, 0, ctx->compute>>>((const uint64_t*)ctx->d_juncs.p, capj, keys_j, d_cnt + 0);
  hs_compact_kernel<<<grid_for(capd, 256), 256, 0, ctx->compute>>>((const uint64_t*)ctx->d_dels.p, capd, keys_d, d_cnt + 1);
  CU(cudaGetLastError()); ctx->own_launches += 2;
  CU(cudaMemcpyAsync(d_cnt + 2, ctx->d_ins_count, 8, cudaMemcpyDeviceToDevice, ctx->compute));
  if (ctx->sp.fusion_search) CU(cudaMemcpyAsync(d_cnt + 3, ctx->d_fus_count, 8, cudaMemcpyDeviceToDevice, ctx->compute));
  // ncclUint64 = 5
  if (ctx->nccl.AllGather(d_cnt, d_cnt + 4, 4, 5, ctx->comm, ctx->compute) != 0) return fail(ctx, THB_ENCCL, "ncclAllGather(counts)");
  std::vector<unsigned long long> counts(4 + 4 * (size_t)W);
  CU(cudaMemcpyAsync(counts.data(), d_cnt, 8 * counts.size(), cudaMemcpyDeviceToHost, ctx->compute));
  CU(cudaStreamSynchronize(ctx->compute));
  unsigned long long mx[4] = {0, 0, 0, 0}, tot[4] = {0, 0, 0, 0};
  for (int r = 0; r < W; ++r) for (int k = 0; k < 4; ++k) { const unsigned long long c = counts[4 + 4 * r + k]; mx[k] = std::max(mx[k], c); tot[k] += c; }
  // key sections are padded to an even number of 8-byte keys so that the record sections behind them stay 16-byte aligned for
  // every world size (FusRec is loaded with 128-bit accesses); padding keys are HS_EMPTY and ignored on insertion
  mx[0] = (mx[0] + 1) & ~1ull; mx[1] = (mx[1] + 1) & ~1ull;
  if (counts[2] > ctx->cap_ins) return fail(ctx, THB_ESTATE, "insertion buffer overflow before the all-gather");
  if (counts[3] > ctx->cap_fus) return fail(ctx, THB_ESTATE, "fusion buffer overflow before the all-gather");
  // staging: [juncs send | dels send | ins send | fusion send] and the W-fold receive areas
  const size_t sj = mx[0] * 8, sd = mx[1] * 8, si = mx[2] * sizeof(InsRec), sf = mx[3] * sizeof(FusRec);
  CU(ctx->ag_send.reserve(sj + sd + si + sf + 64)); CU(ctx->ag_recv.reserve((sj + sd + si + sf) * (size_t)W + 64));
  uint8_t* send = (uint8_t*)ctx->ag_send.p; uint8_t* recv = (uint8_t*)ctx->ag_recv.p;
  if (sj + sd + si + sf) CU(cudaMemsetAsync(send, 0xff, sj + sd + si + sf, ctx->compute));           // pad with HS_EMPTY
  if (counts[0]) CU(cudaMemcpyAsync(send, keys_j, counts[0] * 8, cudaMemcpyDeviceToDevice, ctx->compute));
  if (counts[1]) CU(cudaMemcpyAsync(send + sj, keys_d, counts[1] * 8, cudaMemcpyDeviceToDevice, ctx->compute));
  if (counts[2]) CU(cudaMemcpyAsync(send + sj + sd, ctx->d_ins.p, counts[2] * sizeof(InsRec), cudaMemcpyDeviceToDevice, ctx->compute));
  if (counts[3]) CU(cudaMemcpyAsync(send + sj + sd + si, ctx->d_fus.p, counts[3] * sizeof(FusRec), cudaMemcpyDeviceToDevice, ctx->compute));
  uint8_t* rj = recv; uint8_t* rd = recv + sj * W; uint8_t* ri = rd + sd * W; uint8_t* rf = ri + si * W;
  if (ctx->nccl.GroupStart) ctx->nccl.GroupStart();
  int e = 0;
  if (mx[0]) e |= ctx->nccl.AllGather(send, rj, mx[0], 5, ctx->comm, ctx->compute);
  if (mx[1]) e |= ctx->nccl.AllGather(send + sj, rd, mx[1], 5, ctx->comm, ctx->compute);
  if (mx[2]) e |= ctx->nccl.AllGather(send + sj + sd, ri, mx[2] * 4, 5, ctx->comm, ctx->compute);
  if (mx[3]) e |= ctx->nccl.AllGather(send + sj + sd + si, rf, mx[3] * 4, 5, ctx->comm, ctx->compute);
  if (ctx->nccl.GroupEnd) e |= ctx->nccl.GroupEnd();
  if (e != 0) return fail(ctx, THB_ENCCL, "ncclAllGather(payload)");
  // the key sets are NOT re-inserted into the hash sets: thb_segjuncs_finish sorts and de-duplicates the gathered lists (which
  // contain this rank's own keys as well)
  ctx->ag_keys[0] = (const uint64_t*)rj; ctx->ag_n[0] = mx[0] * W; ctx->ag_keys[1] = (const uint64_t*)rd; ctx->ag_n[1] = mx[1] * W;
  if (!mx[0]) ctx->ag_keys[0] = (const uint64_t*)recv;       // empty everywhere: still "gathered" (n = 0)
  if (!mx[1]) ctx->ag_keys[1] = (const uint64_t*)recv;
  if (mx[2]) {
    if (tot[2] > ctx->cap_ins) {                                   // the gathered records replace the local buffer
      ctx->d_ins.release(); ctx->cap_ins = tot[2] + 1024; CU(ctx->d_ins.reserve(ctx->cap_ins * sizeof(InsRec)));
    }
    CU(cudaMemsetAsync(ctx->d_ins_count, 0, 8, ctx->compute));
    rec_append_kernel<InsRec><<<grid_for(mx[2] * W, 256), 256, 0, ctx->compute>>>((const InsRec*)ri, mx[2] * W, (InsRec*)ctx->d_ins.p, ctx->d_ins_count, ctx->cap_ins, ctx->d_err, 1u);
    ctx->own_launches++;
    ctx->h_ins_count = tot[2];
  }
  if (mx[3]) {                                                     // fusion records: every rank ends up with the union
    if (tot[3] > ctx->cap_fus) { ctx->d_fus.release(); ctx->cap_fus = tot[3] + 1024; CU(ctx->d_fus.reserve(ctx->cap_fus * sizeof(FusRec))); }
    CU(cudaMemsetAsync(ctx->d_fus_count, 0, 8, ctx->compute));
    rec_append_kernel<FusRec><<<grid_for(mx[3] * W, 256), 256, 0, ctx->compute>>>((const FusRec*)rf, mx[3] * W, (FusRec*)ctx->d_fus.p, ctx->d_fus_count, ctx->cap_fus, ctx->d_err, 8u);
    ctx->own_launches++;
    ctx->h_fus_count = tot[3];
  }
  CU(cudaGetLastError());
  return THB_OK;
}

}  // extern "C"
