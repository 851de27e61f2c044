// Device-resident state of a pass: control blocks, pointer table, per-pass arguments.
// Part of the single translation unit engine.cu (included inside namespace ries).
#pragma once

/* ---------------------------------------------------------------- device state ---------- */

#define OVF_NODES 1u
#define OVF_DONE 2u
#define OVF_EVENTS 4u
#define OVF_SMALL 8u /* a pass predicted small was not: rerun it through the general kernels */
#define OVF_KEPT 16u /* sharded passes: a rank's send slot was too small */
#define OVF_TAIL 32u /* shared RHS pass: the list of bound-dependent records was too small */

/* One bound-dependent correction of a shared RHS pass's dead-end count: it applies to a target whose overflow
   bound P satisfies lo < P <= hi (see k1_free.cuh and DESIGN.md section 7). */
struct TailRec {
  double lo, hi;
  int val, pad;
};

/* A shared RHS pass as its consumers see it (engine.cu SharedRhs; passed to the kernels by value). */
struct SharedView {
  const Row* rows;         /* val, M (in `der`), symbols; slots may be holes */
  const uint16_t* forms;
  const TailRec* tail;
  u32 n_slots, n_tail, node_count, pad;
  u64 dead_base;           /* dead-ends under the largest bound of the batch */
  double p_t;              /* this target's bound */
};

/* Header of a rank's send slot in the exchange of a sharded pass (shard.cuh). */
struct __attribute__((aligned(16))) SlotHdr {
  u32 n_kept, overflow, node_count, done_slots;
  u64 dead, gen;
  u32 pad[8];
};

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
     k_keys reads the all-gathered send slots instead of this rank's own expressions */
  int shard_rank, shard_world, gathered;
  int shard_bound; /* rows per rank in the exchange of a sharded pass (send slot, stride of the gathered array) */
  /* sharded pass, second half split by VALUE RANGE (shard.cuh): this rank sorts, dedupes and matches only the
     gathered rows whose value lies in its range; rows per rank / candidates per rank in the second exchange */
  int range_split, range_bound, range_ebound;
  /* K1: a position that cannot fill the grid with full tiles is spread over all the blocks, one small tile each
     (k1_free.cuh): the latency of a lone search; off for the engines of a batch, which share the GPU and want
     full batches instead */
  int k1_spread;
  /* RHS pass enumerated once for a whole batch of targets (engine.cu, SharedRhs): `der` of an RHS prefix carries
     M = max |value| along the prefix, and what depends on a target's overflow bound is recorded (TailRec) for
     every prefix / child with M >= p_min, the smallest bound in the batch */
  int track_m;
  u32 cap_tail;
  double p_min;
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
  u32 n_keys;             /* complete expressions that survive the pre-dedupe and go through the sort */
  u64 dead;               /* dead-ends of the pass */
  u32 aux_ticket;         /* spare ticket */
  u32 n_kept;             /* this rank's complete expressions that survive its local pre-dedupe */
  u64 gen;                /* complete expressions generated (g_ne, ries.c:9190): done_count counts SLOTS, holes included */
  u32 max_nodes, max_done; /* sharded passes: the largest rank's node and expression slots (arena sizing) */
  u32 n_tail;              /* TailRec records written (may exceed the capacity: then OVF_TAIL is set) */
  u32 bar_gen;             /* k_expand_all: grid-barrier generation (monotonic, survives passes) */
  u64 klo, khi;            /* sharded pass split by value range: this rank keeps the keys in [klo, khi) */
  u32 max_new, max_ev;     /* ... the largest range's new rows and candidates (slot sizes of the next pass) */
  /* device clock (%globaltimer, ns) at the start of each stage and at the end of the pass: [0] K1, [1] keys +
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

/* What the host sends per pass, in one copy: the scalars, (the view of a shared RHS pass,) then the forms. */
#define PASSIN_VIEW_OFFSET 128
#define PASSIN_FORMS_OFFSET 256
#define PASSIN_MAX_FORMS 65536
static_assert(sizeof(PassArgs) <= PASSIN_VIEW_OFFSET, "PassArgs outgrew its slot in the pass block");
static_assert(sizeof(SharedView) <= PASSIN_FORMS_OFFSET - PASSIN_VIEW_OFFSET, "SharedView outgrew its slot in the pass block");

/* Survives passes. */
struct TableState {
  u32 cur; /* which table buffer is live */
  u32 nt;  /* rows in it */
  u32 epoch; /* look-back tag, bumped once per pass; each kernel adds its own offset */
  u32 pad;
};

/* INSERTION ORDER WITHOUT ORDERED OUTPUT.  The reference inserts expressions one at a time, and what it prints
   depends on the order: forms in pre-order, within a form the lexicographic order of symbol TRIAL positions
   (ge_2's loop, ries.c:9388).  Trial order within a class is the order of the symbol table (setup.cc fills
   trial[] by ascending table index), so for two expressions of one pass
        a was inserted before b   <=>   (form_a, code_a[0], code_a[1], ...) < (form_b, code_b[0], ...)
   lexicographically, the codes being the 6-bit table indices stored in Sym128.  K1 therefore does not have to
   EMIT expressions in insertion order (which took a look-back chain over all tiles in flight and a second phase
   per tile): every consumer that needs "who came first" compares these keys. */
struct SeqKey {
  u64 lo, hi; /* Sym128 of the expression as enumerated */
  u32 form, pad;
};

__device__ __forceinline__ bool sym_before(u64 alo, u64 ahi, u64 blo, u64 bhi) {
  /* position 0 sits in the lowest 6 bits: the first differing position is the lowest differing bit's group */
  const u64 xl = alo ^ blo;
  if (xl) {
    const int g = (__ffsll((long long)xl) - 1) / 6;
    if (g < 10) return ((alo >> (6 * g)) & 63ull) < ((blo >> (6 * g)) & 63ull);
    /* position 10 straddles the two words (bits 60..65) */
    return (((alo >> 60) | (ahi << 4)) & 63ull) < (((blo >> 60) | (bhi << 4)) & 63ull);
  }
  const u64 xh = ahi ^ bhi;
  if (!xh) return false;
  const int t = __ffsll((long long)xh) - 1; /* bit t of hi is bit 64 + t of the whole */
  if (t < 2) return (((alo >> 60) | (ahi << 4)) & 63ull) < (((blo >> 60) | (bhi << 4)) & 63ull); /* equal low words: decided by the two high bits of position 10 */
  const int g = (t - 2) / 6; /* position 11 + g */
  return ((ahi >> (2 + 6 * g)) & 63ull) < ((bhi >> (2 + 6 * g)) & 63ull);
}
__device__ __forceinline__ bool seq_before(u32 fa, const Sym128& a, u32 fb, const Sym128& b) {
  if (fa != fb) return fa < fb;
  return sym_before(a.lo, a.hi, b.lo, b.hi);
}
__device__ __forceinline__ bool seq_before(const SeqKey& a, const SeqKey& b) {
  if (a.form != b.form) return a.form < b.form;
  return sym_before(a.lo, a.hi, b.lo, b.hi);
}

/* The buffer pointers come out of the device-resident pointer table, so the compiler cannot see their address
   space and would emit generic loads and stores; every one of them is global memory. */
#ifdef RIES_NO_ASSUME_GLOBAL
#define ASSUME_GLOBAL(p) ((void)0)
#else
#define ASSUME_GLOBAL(p) __builtin_assume(__isGlobal(p))
#endif

/* Bulk asynchronous copy global -> shared (the TMA engine's 1-D path, SASS UBLKCP) completing on an mbarrier: one
   thread issues it, nobody's registers or issue slots are spent on the transfer, and the block does something
   else until it needs the data.  Source, destination and size must be multiples of 16 bytes. */
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, u32 bytes, unsigned long long* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, u32 parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

#define FORM_HOLE 0xffffu
#define TREF_LHS 0x80000000u /* the side of a table row rides in its reference: der != 0 <=> LHS (ries.c:8496) */ /* done_form[] / Node::form of a slot that was allocated but not filled */

/* Hash set of table keys, one per side: 4-slot buckets of 32 bytes (one DRAM sector per probe).  It only has to be
   SOUND: a key it reports is in the table on that side.  A key that found its bucket full is simply not
   remembered (the sort and the table search after it are the arbiter), so there is no overflow handling. */
__device__ __forceinline__ u64 hset_bucket(u64 key, u32 log2b) { return (key * 0x9E3779B97F4A7C15ull) >> (64 - log2b); }
__device__ __forceinline__ bool hset_contains(const u64* __restrict__ slots, u32 log2b, u64 key) {
  if (!key) return false;
  const ulonglong2* b = (const ulonglong2*)(slots + (hset_bucket(key, log2b) << 2));
  const ulonglong2 a = b[0], c = b[1];
  return a.x == key || a.y == key || c.x == key || c.y == key;
}
__device__ __forceinline__ void hset_insert(u64* slots, u32 log2b, u64 key) {
  if (!key) return;
  unsigned long long* b = (unsigned long long*)(slots + (hset_bucket(key, log2b) << 2));
  /* one look at the bucket (a single 32-byte sector), then one compare-and-swap on the first free slot: a CAS per
     slot in turn costs a DRAM round trip each, and this is the hot spot of the merge kernel */
  const ulonglong2 a = __ldcg((const ulonglong2*)b), c = __ldcg((const ulonglong2*)b + 1);
  if (a.x == key || a.y == key || c.x == key || c.y == key) return;
  int s = !a.x ? 0 : !a.y ? 1 : !c.x ? 2 : !c.y ? 3 : 4;
  for (; s < 4; s++) {
    const unsigned long long old = atomicCAS(b + s, 0ull, (unsigned long long)key);
    if (old == 0ull || old == key) return;
  }
}

/* Every buffer the kernels touch.  Lives in device memory; the host rewrites it when a buffer grows,
   so kernel arguments (and the captured graphs) never change. */
struct DevPtrs {
  const RiesParams* P;
  const TestTables* tt;
  FormDesc* forms;
  Node* nodes;
  Row* done;
  uint16_t* done_form;
  Row* gdone;           /* all ranks' send slots, rank r at [r * shard_bound ...) (sharded passes only) */
  uint16_t* gdone_form;
  Row* surv;           /* tiny passes: the pass's expressions in insertion order (SP_CAP rows) */
  Row* ksend;          /* sharded passes: this rank's send slot (rows, form indices, header) */
  uint16_t* ksend_form;
  SlotHdr* shdr;
  SlotHdr* ghdr;
  SlotHdr* ghdr2;        /* second exchange of a range-split pass: headers, new rows (index, gap), candidates */
  u32* g2_idx;
  u32* g2_gap;
  MatchEvent* g2_events;
  TailRec* tail;  /* shared RHS passes being built: bound-dependent records */ /* ... and the gathered headers, one per rank */
  u64* keys[2];
  u32* idx[2];
  u32* uniq_seq;  /* new rows in value order: index of the row in done[] (tiny passes: insertion number) */
  u32* uniq_gap;
  SeqKey* uniq_sk; /* ... and its insertion-order key (general passes) */
  u64* hset[2];   /* per side (0 RHS, 1 LHS): hash set of the keys resident in the table, 4-slot buckets */
  u32 hset_log2[2]; /* log2 of the number of buckets */
  Row* heap;      /* the table's rows, append-only: a row never moves once inserted */
  u32* tref[2];   /* per table position (sorted by value, ping-pong with tval): TREF_LHS | index of its row in heap[] */
  double* tval[2];
  MatchEvent* events;
  Staged* staged; /* K1 of the tiny passes: scratch, 2 x EXP_MAX_ITEMS */
  u64* rs_state;  /* radix look-back state, 256 words per chunk */
  u64* cc_state;  /* classify look-back state, one word per chunk */
  SortCtl* sortctl;
  u64* opcnt;
  PassOut* host_out; /* mapped pinned host memory: the last kernel of a pass writes its results there itself */
};
