/* Stand-in for libnccl.so.2 (emulation harness only): lets thb_segjuncs_allgather's packing, padding and union logic run in a
 * container without GPUs.  World size 1: an all-gather is a copy.  World size W > 1: every "peer" contributes a copy of this rank's
 * send buffer (W replicas) -- enough to exercise the receive layout, the padding keys and the de-duplication of the union. */
#include <string.h>
#include <stddef.h>
#include <stdint.h>
static size_t tsize(int t) { return (t == 0 || t == 1) ? 1 : (t == 2 || t == 3 || t == 7) ? 4 : (t == 6 ? 2 : 8); }
int ncclGetUniqueId(void* id) { memset(id, 7, 128); return 0; }
typedef struct { char b[128]; } uid_t_;
int ncclCommInitRank(void** comm, int world, uid_t_ id, int rank) { (void)id; (void)rank; if (world < 1 || world > 8) return 5; *comm = (void*)(intptr_t)world; return 0; }
int ncclAllGather(const void* s, void* r, size_t count, int dtype, void* comm, void* stream)
{
  (void)stream;
  const int world = (int)(intptr_t)comm; const size_t n = count * tsize(dtype);
  for (int k = world - 1; k >= 0; --k) memmove((char*)r + (size_t)k * n, s, n);      /* s may alias r's first slot */
  return 0;
}
int ncclCommDestroy(void* c) { (void)c; return 0; }
int ncclGroupStart(void) { return 0; }
int ncclGroupEnd(void) { return 0; }
const char* ncclGetErrorString(int e) { (void)e; return "fake nccl (replicating stand-in)"; }
