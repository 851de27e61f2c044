// Pass begin / end kernels.
// Part of the single translation unit engine.cu (included inside namespace ries).
#pragma once

/* ---------------------------------------------------------------- pass begin / end ------ */

__device__ __forceinline__ void pass_begin_body(const DevPtrs* D, const PassArgs* A, PassCtl* C, TableState* T) {
  const u32 nforms = (u32)A->nforms;
  /* one search over several GPUs: every rank starts from all the forms; the unit dealt to the ranks is (form, first
     symbol), at the first expression position (k1_free.cuh) */
  const u32 nroots = nforms;
  for (u32 j = threadIdx.x; j < nroots; j += blockDim.x) {
    Node r;
    r.val = 0; r.der = 0; r.sym.lo = 0; r.sym.hi = 0;
    r.below = RIES_NO_NODE; r.form = (uint16_t)j; r.cplx = 0; r.tag = 0; r.len = 0; r.sp = 0;
    r.pad0 = 0; r.pad1 = 0;
    D->nodes[j] = r;
  }
  u32* sc = (u32*)D->sortctl;
  for (u32 i = threadIdx.x; i < sizeof(SortCtl) / 4; i += blockDim.x) sc[i] = 0;
  if (threadIdx.x == 0) {
    C->lvl_begin = 0; C->lvl_end = nroots; C->node_count = nroots; C->done_count = 0;
    C->tile_ticket = 0; C->blocks_done = 0; C->n_events = 0; C->n_uniq = 0; C->overflow = 0;
    C->chunk_ticket = 0; C->nt_before = T->nt; C->dead = 0; C->aux_ticket = 0; C->n_kept = 0; C->n_keys = 0; C->gen = 0; C->n_tail = 0; C->max_nodes = 0; C->max_done = 0;
    C->stamp[1] = C->stamp[2] = C->stamp[3] = C->stamp[4] = 0;
    u32 e = (T->epoch + 16) & 0x3ffffff0u; /* 16 tags per pass: 8 radix digits + classify */
    if (e == 0) e = 16;
    T->epoch = e;
  }
}

__global__ void k_pass_begin(const DevPtrs* D, const PassArgs* A, PassCtl* C, TableState* T) {
  stage_stamp(C, 0);
  pass_begin_body(D, A, C, T);
}

__device__ __forceinline__ void pass_end_body(const DevPtrs* D, const PassArgs* A, PassCtl* C, TableState* T) {
  if (threadIdx.x == 0 && !C->overflow && C->n_uniq) {
    T->cur ^= 1u;
    T->nt += C->n_uniq;
  }
  /* the first few candidates ride along with the counters (C is the head of a PassOut) */
  u32 ne = C->n_events;
  if (ne > A->cap_events) ne = A->cap_events;
  if (ne > HEAD_EVENTS) ne = HEAD_EVENTS;
  const u32* src = (const u32*)D->events;
  u32* dst = (u32*)((PassOut*)C)->head;
  for (u32 i = threadIdx.x; i < ne * (sizeof(MatchEvent) / 4); i += blockDim.x) dst[i] = src[i];
}

/* The very end of a pass (single block, after pass_end_body and the last stamp): the control block and the head
   candidates go straight into the host's mapped buffer, then the pass number -- the host polls for it instead of
   paying for a copy operation and a stream synchronisation per pass. */
__device__ __forceinline__ void post_to_host(const DevPtrs* D, const PassArgs* A, PassCtl* C) {
  __syncthreads();
  PassOut* h = D->host_out;
  u32 ne = C->n_events;
  if (ne > A->cap_events) ne = A->cap_events;
  if (ne > HEAD_EVENTS) ne = HEAD_EVENTS;
  const u32 words = (u32)(offsetof(PassOut, head) / 4) + ne * (u32)(sizeof(MatchEvent) / 4);
  const volatile u32* src = (const volatile u32*)C;
  volatile u32* dst = (volatile u32*)h;
  for (u32 i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    *(volatile u32*)&h->seq = A->seq;
    __threadfence_system();
  }
}

__global__ void k_pass_end(const DevPtrs* D, const PassArgs* A, PassCtl* C, TableState* T) {
  pass_end_body(D, A, C, T);
  __syncthreads();
  stage_stamp(C, 4);
  post_to_host(D, A, C);
}

/* Consumer side of a shared RHS pass (engine.cu SharedRhs): instead of K1, the pass starts from the batch's
   enumeration.  Control block reset as in k_pass_begin; the dead-end count is the shared pass's, corrected for this
   target's overflow bound by the records that apply to it. */
__global__ void k_shared_front(const DevPtrs* D, const PassArgs* A, PassCtl* C, TableState* T, const SharedView* Vp) {
  const SharedView V = *Vp;
  __shared__ long long s_sum[32];
  stage_stamp(C, 0);
  pass_begin_body(D, A, C, T);
  long long acc = 0;
  for (u32 i = threadIdx.x; i < V.n_tail; i += blockDim.x) {
    const TailRec r = V.tail[i];
    if (r.lo < V.p_t && V.p_t <= r.hi) acc += r.val;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long tot = 0;
    for (u32 w = 0; w < (blockDim.x + 31) / 32; w++) tot += s_sum[w];
    C->dead = (u64)((long long)V.dead_base + tot);
    C->node_count = V.node_count;
    C->lvl_begin = C->lvl_end = 0;
  }
}
