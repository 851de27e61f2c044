// One search sharded over the GPUs of one box: the NCCL plumbing.
//
// north_star row (e): "the LHS index space shards across the GPUs of one box, the RHS table is generated
// redundantly or all-gathered over NVLink with NCCL".  Here both sides shard the same way: the forms of a
// large pass are dealt round-robin to the ranks for enumerate+evaluate (K1) and the hash pre-dedupe, what
// survives is all-gathered in fixed-size slots, and everything downstream runs replicated (device/shard.cuh).
//
// NCCL is bound at run time (dlopen of the libnccl.so.2 already in the process, e.g. torch's), so the
// single-GPU library and the `ries` binary have no NCCL dependency.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace ries {

#define RIES_MAX_RANKS 64 /* the engine's pinned per-rank control blocks (engine.cu h_allctl_) */

struct DistState {
  int rank = 0, world = 1;
  void* comm = 0;     /* ncclComm_t */
  int fake_world = 0; /* test hook RIES_FAKE_SHARD=W: W virtual ranks run one after the other on one GPU */
};

const DistState& dist_state();
/* rank 0: a fresh 128-byte ncclUniqueId to hand to the other ranks (over any side channel) */
int dist_unique_id(unsigned char* id128);
/* every rank, after cudaSetDevice: joins the communicator.  0 on success. */
int dist_init(int rank, int world, const unsigned char* id128);
void dist_finalize();

/* stream-ordered collectives over raw bytes */
void dist_allgather(const void* send, void* recv, size_t bytes_per_rank, cudaStream_t st);
/* three all-gathers as one NCCL group (headers, rows, form indices of a sharded pass) */
void dist_allgather4(const void* s0, void* r0, size_t b0, const void* s1, void* r1, size_t b1, const void* s2, void* r2,
                     size_t b2, const void* s3, void* r3, size_t b3, cudaStream_t st);
void dist_allgather3(const void* s0, void* r0, size_t b0, const void* s1, void* r1, size_t b1, const void* s2, void* r2,
                     size_t b2, cudaStream_t st);

}  // namespace ries
