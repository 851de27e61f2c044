// NCCL bound at run time; see dist.h.
#include "host/dist.h"

#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace ries {
namespace {

/* the few pieces of NCCL's C ABI used here (stable across NCCL 2.x) */
typedef void* nccl_comm_t;
typedef struct { char internal[128]; } nccl_unique_id;
enum { NCCL_SUCCESS = 0, NCCL_INT8 = 0 };

struct Api {
  void* lib = 0;
  int (*GetUniqueId)(nccl_unique_id*) = 0;
  int (*CommInitRank)(nccl_comm_t*, int, nccl_unique_id, int) = 0;
  int (*CommDestroy)(nccl_comm_t) = 0;
  int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = 0;
  int (*Send)(const void*, size_t, int, int, nccl_comm_t, cudaStream_t) = 0;
  int (*Recv)(void*, size_t, int, int, nccl_comm_t, cudaStream_t) = 0;
  int (*GroupStart)() = 0;
  int (*GroupEnd)() = 0;
  const char* (*GetErrorString)(int) = 0;
};
Api g_api;
DistState g_dist;

bool load_api() {
  if (g_api.lib) return true;
  /* prefer the copy already in the process (torch's), so that two NCCLs never coexist */
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    fprintf(stderr, "ries_b200: cannot load libnccl.so.2: %s\n", dlerror());
    return false;
  }
#define BIND(field, name)                                               \
  *(void**)(&g_api.field) = dlsym(h, name);                             \
  if (!g_api.field) { fprintf(stderr, "ries_b200: %s missing in libnccl\n", name); return false; }
  BIND(GetUniqueId, "ncclGetUniqueId")
  BIND(CommInitRank, "ncclCommInitRank")
  BIND(CommDestroy, "ncclCommDestroy")
  BIND(AllGather, "ncclAllGather")
  BIND(Send, "ncclSend")
  BIND(Recv, "ncclRecv")
  BIND(GroupStart, "ncclGroupStart")
  BIND(GroupEnd, "ncclGroupEnd")
  BIND(GetErrorString, "ncclGetErrorString")
#undef BIND
  g_api.lib = h;
  return true;
}

void check(int r, const char* what) {
  if (r != NCCL_SUCCESS) {
    fprintf(stderr, "ries_b200: NCCL error in %s: %s\n", what, g_api.GetErrorString ? g_api.GetErrorString(r) : "?");
    abort();
  }
}

}  // namespace

const DistState& dist_state() {
  static bool init = false;
  if (!init) {
    init = true;
    const char* f = getenv("RIES_FAKE_SHARD");
    if (f && atoi(f) > 1) g_dist.fake_world = atoi(f);
  }
  return g_dist;
}

int dist_unique_id(unsigned char* id128) {
  if (!load_api()) return -1;
  nccl_unique_id id;
  int r = g_api.GetUniqueId(&id);
  if (r != NCCL_SUCCESS) return r;
  memcpy(id128, id.internal, 128);
  return 0;
}

int dist_init(int rank, int world, const unsigned char* id128) {
  if (world <= 1) return 0;
  if (!load_api()) return -1;
  if (g_dist.comm) dist_finalize();
  nccl_unique_id id;
  memcpy(id.internal, id128, 128);
  nccl_comm_t comm = 0;
  int r = g_api.CommInitRank(&comm, world, id, rank);
  if (r != NCCL_SUCCESS) {
    fprintf(stderr, "ries_b200: ncclCommInitRank failed: %s\n", g_api.GetErrorString(r));
    return r;
  }
  g_dist.comm = comm;
  g_dist.rank = rank;
  g_dist.world = world;
  return 0;
}

void dist_finalize() {
  if (g_dist.comm) g_api.CommDestroy((nccl_comm_t)g_dist.comm);
  g_dist.comm = 0;
  g_dist.rank = 0;
  g_dist.world = 1;
}

void dist_allgather(const void* send, void* recv, size_t bytes_per_rank, cudaStream_t st) {
  check(g_api.AllGather(send, recv, bytes_per_rank, NCCL_INT8, (nccl_comm_t)g_dist.comm, st), "ncclAllGather");
}

void dist_allgather3(const void* s0, void* r0, size_t b0, const void* s1, void* r1, size_t b1, const void* s2, void* r2,
                     size_t b2, cudaStream_t st) {
  nccl_comm_t c = (nccl_comm_t)g_dist.comm;
  check(g_api.GroupStart(), "ncclGroupStart");
  check(g_api.AllGather(s0, r0, b0, NCCL_INT8, c, st), "ncclAllGather");
  check(g_api.AllGather(s1, r1, b1, NCCL_INT8, c, st), "ncclAllGather");
  check(g_api.AllGather(s2, r2, b2, NCCL_INT8, c, st), "ncclAllGather");
  check(g_api.GroupEnd(), "ncclGroupEnd");
}

void dist_allgather4(const void* s0, void* r0, size_t b0, const void* s1, void* r1, size_t b1, const void* s2, void* r2,
                     size_t b2, const void* s3, void* r3, size_t b3, cudaStream_t st) {
  nccl_comm_t c = (nccl_comm_t)g_dist.comm;
  check(g_api.GroupStart(), "ncclGroupStart");
  check(g_api.AllGather(s0, r0, b0, NCCL_INT8, c, st), "ncclAllGather");
  check(g_api.AllGather(s1, r1, b1, NCCL_INT8, c, st), "ncclAllGather");
  check(g_api.AllGather(s2, r2, b2, NCCL_INT8, c, st), "ncclAllGather");
  check(g_api.AllGather(s3, r3, b3, NCCL_INT8, c, st), "ncclAllGather");
  check(g_api.GroupEnd(), "ncclGroupEnd");
}

}  // namespace ries
