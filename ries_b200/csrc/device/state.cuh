// Device-resident state of a pass: control blocks, pointer table, per-pass arguments.
// Part of the single translation unit engine.cu (included inside namespace ries).
#pragma once

/* ---------------------------------------------------------------- device state ---------- */

#define OVF_NODES 1u
#define OVF_DONE 2u
#define OVF_EVENTS 4u
#define OVF_SMALL 8u /* a pass predicted small was not: rerun it through the general kernels */

struct SortCtl {
  u32 ghist[8][256];
  u32 ticket[8];
};

struct __attribute__((aligned(8))) Staged { /* a child between phase 3 and phase 4 of K1 (L2-resident) */
  double val, der;
  u32 below;
  unsigned char tag, sp, pad0, pad1;
};

/* Per-pass scalars, written by the host (one small H2D copy) before the pass graph is launched. */
struct PassArgs {
  int using_x, cplx, skip_exact, nforms;
  double best_match0;
  u32 cap_nodes, cap_done, cap_events;
  u32 seq; /* number of this pass launch: the last kernel posts it to the host together with the results */
  /* one search over several GPUs: this rank enumerates the forms f with f % shard_world == shard_rank, and
     the regroup kernels read the all-gathered complete expressions instead of this rank's own */
  int shard_rank, shard_world, gathered, pad2;
};

/* Written by the kernels of a pass; copied back to pinned host memory at its end. */
struct PassCtl {
  u32 lvl_begin, lvl_end; /* frontier = nodes[lvl_begin, lvl_end) */
  u32 node_count;         /* live prefixes so far (may exceed capacity: then overflow is set) */
  u32 done_count;         /* complete expressions so far */
  u32 tile_ticket, blocks_done;
  u32 n_events, n_uniq;
  u32 overflow;
  u32 chunk_ticket;       /* k_classify_compact */
  u32 nt_before;          /* table rows before this pass */
  u32 pad;
  u64 dead;               /* dead-ends of the pass */
  u32 aux_ticket;         /* k_local_filter / k_compact_kept (sharded passes) */
  u32 n_kept;             /* this rank's complete expressions that survive its local pre-dedupe */
  /* device clock (%globaltimer, ns) at the start of each stage and at the end of the pass: [0] K1, [1] regroup +
     sort, [2] classify/match/merge, [3] refine, [4] end.  Written by thread 0 of block 0 of the stage's first
     kernel (the kernels of a pass run back to back on one stream), read by the host from the copied-back
     block: per-stage device time without a CUDA event per stage on the host's critical path. */
  u64 stamp[5];
};

__device__ __forceinline__ u64 dev_now_ns() {
  u64 t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void stage_stamp(const PassCtl* C, int i) {
  if (blockIdx.x == 0 && threadIdx.x == 0) ((PassCtl*)C)->stamp[i] = dev_now_ns();
}

#define HEAD_EVENTS 16 /* candidates copied back together with the counters (one copy, one sync per pass) */
struct PassOut {
  PassCtl ctl;
  MatchEvent head[HEAD_EVENTS];
  u32 seq; /* host copy only: PassArgs::seq of the pass whose results are complete in this block */
  u32 pad[3];
};

/* What the host sends per pass, in one copy: the scalars, then the forms. */
#define PASSIN_FORMS_OFFSET 256
#define PASSIN_MAX_FORMS 65536

/* Survives passes. */
struct TableState {
  u32 cur; /* which table buffer is live */
  u32 nt;  /* rows in it */
  u32 epoch; /* look-back tag, bumped once per pass; each kernel adds its own offset */
  u32 pad;
};

/* Every buffer the kernels touch.  Lives in device memory; the host rewrites it when a buffer grows,
   so kernel arguments (and the captured graphs) never change. */
struct DevPtrs {
  const RiesParams* P;
  const TestTables* tt;
  FormDesc* forms;
  Node* nodes;
  Row* done;
  uint16_t* done_form;
  Row* gdone;           /* all ranks' complete expressions, rank-major (sharded passes only) */
  uint16_t* gdone_form;
  Row* surv;
  uint16_t* surv_form; /* form index of surv[i] (sharded passes: needed again after the exchange) */
  unsigned char* keep; /* sharded passes: surv[i] takes part in the exchange */
  Row* ksend;          /* ... compacted rows / forms this rank sends */
  uint16_t* ksend_form;
  u64* keys[2];
  u32* idx[2];
  u32* form_first;
  u32* form_cnt;
  u32* form_base;
  u32* uniq_seq;
  u32* uniq_gap;
  Row* trow[2];
  double* tval[2];
  MatchEvent* events;
  u64* tiles;     /* K1 look-back state, one word per tile */
  Staged* staged; /* K1 scratch, 2 x EXP_MAX_ITEMS per block */
  u64* rs_state;  /* radix look-back state, 256 words per chunk */
  u64* cc_state;  /* classify look-back state, one word per chunk */
  SortCtl* sortctl;
  u64* opcnt;
  PassOut* host_out; /* mapped pinned host memory: the last kernel of a pass writes its results there itself */
};
