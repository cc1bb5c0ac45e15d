/* Single-rank stand-in for libnccl.so.2 (emulation harness only): lets thb_segjuncs_allgather's packing, padding
 * and re-insertion logic run in a container without GPUs.  World size 1: an all-gather is a copy. */
#include <string.h>
#include <stddef.h>
static size_t tsize(int t) { return (t == 0 || t == 1) ? 1 : (t == 2 || t == 3 || t == 7) ? 4 : (t == 6 ? 2 : 8); }
int ncclGetUniqueId(void* id) { memset(id, 7, 128); return 0; }
typedef struct { char b[128]; } uid_t_;
int ncclCommInitRank(void** comm, int world, uid_t_ id, int rank) { (void)id; (void)rank; if (world != 1) return 5; *comm = (void*)1; return 0; }
int ncclAllGather(const void* s, void* r, size_t count, int dtype, void* comm, void* stream)
{ (void)comm; (void)stream; memmove(r, s, count * tsize(dtype)); return 0; }
int ncclCommDestroy(void* c) { (void)c; return 0; }
int ncclGroupStart(void) { return 0; }
int ncclGroupEnd(void) { return 0; }
const char* ncclGetErrorString(int e) { (void)e; return "fake nccl (world size 1 only)"; }
