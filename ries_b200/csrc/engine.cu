// The sm_100a device engine: every hot stage of the RIES search as hand-written CUDA kernels.
//
//   K1  k_expand            enumerate + evaluate one expression position of every live prefix
//                           [ge_2 + exec, ries.c:8908-9475, 5308-5950]
//   K2  k_reorder/k_rs_onesweep       stable LSD radix sort of the pass's survivors on the fp64 value
//       k_classify_compact  duplicate pruning inside the pass and against the resident table
//       k_match_merge_new, k_merge_old   merge into the resident sorted table   [bt_insert, ries.c:8622-8724]
//   K3  k_match_merge_new   gap-local neighbour test with the derivative-scaled window
//                                                           [check_sides/check_match, 8488, 8243]
//   K4  k_refine            cv_simplify + Newton on the compacted candidates        [7297, 7139]
//
// A whole complexity pass is DEVICE-DRIVEN: every kernel has a fixed grid and reads its problem size
// (frontier bounds, number of survivors, number of new rows, table size) from a control block in HBM
// that the previous kernel wrote, and every buffer pointer lives in a device-resident pointer table.
// Kernel arguments are therefore constant for the life of the engine, the per-pass launch sequence
// is captured once into a CUDA graph per expression length, and the host synchronises ONCE per pass
// (to read counters and candidates for the sequential replay that decides the next pass).
//
// Data layout in HBM (all device-resident for the whole search):
//   nodes[]   48 B AoS, one per live prefix of the running pass (cactus stack, see params.h)
//   table     rows 32 B AoS sorted by value + a dense 8 B key array for searches, ping-pong
//   surv[]    32 B rows, the pass's complete expressions in insertion (DFS) order
// No tensor cores: nothing here is a dense contraction.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <vector>

#include "core/expand.h"
#include "core/refine.h"
#include "host/dist.h"
#include "host/ries_host.h"

namespace ries {

#define CUDA_CHECK(x)                                                                        \
  do {                                                                                       \
    cudaError_t e_ = (x);                                                                    \
    if (e_ != cudaSuccess) {                                                                 \
      fprintf(stderr, "ries_b200: CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e_), __FILE__, \
              __LINE__, cudaGetErrorString(e_));                                             \
      abort();                                                                               \
    }                                                                                        \
  } while (0)

typedef unsigned long long u64;
typedef unsigned int u32;

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
  u32 cap_nodes, cap_done, cap_events, pad;
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
};

#define HEAD_EVENTS 16 /* candidates copied back together with the counters (one copy, one sync per pass) */
struct PassOut {
  PassCtl ctl;
  MatchEvent head[HEAD_EVENTS];
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
};

/* ---------------------------------------------------------------- pass begin / end ------ */

__global__ void k_pass_begin(const DevPtrs* D, const PassArgs* A, PassCtl* C, TableState* T) {
  const u32 nforms = (u32)A->nforms;
  const u32 world = (u32)A->shard_world, rank = (u32)A->shard_rank;
  const u32 nroots = (nforms > rank) ? (nforms - rank + world - 1) / world : 0; /* forms dealt round-robin */
  for (u32 j = threadIdx.x; j < nroots; j += blockDim.x) {
    Node r;
    r.val = 0; r.der = 0; r.sym.lo = 0; r.sym.hi = 0;
    r.below = RIES_NO_NODE; r.form = (uint16_t)(rank + j * world); r.cplx = 0; r.tag = 0; r.len = 0; r.sp = 0;
    r.pad0 = 0; r.pad1 = 0;
    D->nodes[j] = r;
  }
  for (u32 i = threadIdx.x; i < nforms; i += blockDim.x) {
    D->form_first[i] = 0;
    D->form_cnt[i] = 0;
  }
  u32* sc = (u32*)D->sortctl;
  for (u32 i = threadIdx.x; i < sizeof(SortCtl) / 4; i += blockDim.x) sc[i] = 0;
  if (threadIdx.x == 0) {
    C->lvl_begin = 0; C->lvl_end = nroots; C->node_count = nroots; C->done_count = 0;
    C->tile_ticket = 0; C->blocks_done = 0; C->n_events = 0; C->n_uniq = 0; C->overflow = 0;
    C->chunk_ticket = 0; C->nt_before = T->nt; C->dead = 0; C->aux_ticket = 0; C->n_kept = 0;
    u32 e = (T->epoch + 16) & 0x3ffffff0u; /* 16 tags per pass: 8 radix digits + classify */
    if (e == 0) e = 16;
    T->epoch = e;
  }
}

__global__ void k_pass_end(const DevPtrs* D, const PassArgs* A, PassCtl* C, TableState* T) {
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

/* ---------------------------------------------------------------- K1: enumerate+evaluate - */

/* K1 works on tiles of 256 consecutive live prefixes (one block per tile at a time):
     phase 1  thread = prefix: the integer symbol-append tests for all candidate symbols -> bit mask
     phase 2  the (prefix, symbol) pairs that passed are listed in shared memory GROUPED BY SYMBOL SLOT
     phase 3  thread = listed pair: exec + value prunes; warps are (almost) operator-uniform
     phase 4  ordered append: children land in (prefix, symbol) order = the reference's DFS order
   The last block to leave advances the frontier for the next level. */
#ifndef EXP_THREADS
#define EXP_THREADS 256
#endif
#define EXP_PARENTS EXP_THREADS
#define EXP_MAX_ITEMS (EXP_PARENTS * RIES_SLOTS)
#ifndef EXP_MIN_BLOCKS
#define EXP_MIN_BLOCKS 3
#endif
#define IT_LIVE 0x4000u
#define IT_DONE 0x8000u

#define TS_AGG (1ull << 62)
#define TS_INC (2ull << 62)
#define TS_STATUS(x) ((x) >> 62)
#define TS_LIVE(x) ((u32)(((x) >> 31) & 0x7fffffffull))
#define TS_DONE(x) ((u32)((x)&0x7fffffffull))
#define TS_PACK(st, live, done) ((st) | ((u64)(live) << 31) | (u64)(done))

/* Decoupled look-back over tile aggregates (two counters packed in one word), in two steps so that a
   block can publish a tile's aggregate, go on to the next tile, and only then resolve the first one's
   exclusive prefix (by then its predecessors have long published). */
__device__ __forceinline__ void tile_publish(volatile u64* state, u32 tile, u32 live_agg, u32 done_agg) {
  __threadfence();
  state[tile] = TS_PACK(tile == 0 ? TS_INC : TS_AGG, live_agg, done_agg);
}

/* Called by warp 0. */
__device__ __forceinline__ void tile_lookback(volatile u64* state, u32 tile, u32 live_agg, u32 done_agg,
                                              u32& live_excl, u32& done_excl) {
  const u32 lane = threadIdx.x & 31;
  live_excl = 0; done_excl = 0;
  if (tile == 0) return;
  u32 lsum = 0, dsum = 0;
  int idx = (int)tile - 1;
  while (true) {
    int my = idx - (int)lane;
    u64 w;
    if (my >= 0) {
      do { w = state[my]; } while (TS_STATUS(w) == 0);
    } else {
      w = TS_PACK(TS_INC, 0, 0);
    }
    u32 inc_mask = __ballot_sync(0xffffffffu, TS_STATUS(w) == 2);
    /* lanes closer than the first inclusive one contribute their aggregates; that one its prefix */
    u32 first_inc = inc_mask ? (u32)__ffs(inc_mask) - 1 : 32;
    u32 lv = (lane <= first_inc) ? TS_LIVE(w) : 0;
    u32 dv = (lane <= first_inc) ? TS_DONE(w) : 0;
    for (int o = 16; o > 0; o >>= 1) {
      lv += __shfl_xor_sync(0xffffffffu, lv, o);
      dv += __shfl_xor_sync(0xffffffffu, dv, o);
    }
    lsum += lv; dsum += dv;
    if (inc_mask) break;
    idx -= 32;
  }
  if (lane == 0) {
    __threadfence();
    state[tile] = TS_PACK(TS_INC, lsum + live_agg, dsum + done_agg);
  }
  live_excl = lsum; done_excl = dsum;
}

/* Shared-memory state of one tile between its front half (phases 1-3) and its back half (phase 4). */
struct TileBuf {
  Node par[EXP_PARENTS];
  unsigned short item[EXP_MAX_ITEMS];
  u32 live[EXP_PARENTS], done[EXP_PARENTS];   /* per prefix: bit per slot */
  u32 lbase[EXP_PARENTS], dbase[EXP_PARENTS]; /* per prefix: exclusive child offsets inside the tile */
  unsigned char cls[EXP_PARENTS], complete[EXP_PARENTS];
  u32 total, tile, live_agg, done_agg;
};

#define EXP_DYN_SMEM (2 * sizeof(TileBuf))

__global__ void __launch_bounds__(EXP_THREADS, EXP_MIN_BLOCKS)
k_expand(const DevPtrs* __restrict__ Dp, const PassArgs* __restrict__ Ap, PassCtl* ctl) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  TileBuf* s_buf = (TileBuf*)s_dyn; /* [2] */
  __shared__ u32 s_opcnt[RIES_MAX_SYMS];
  __shared__ RiesParams P;
  __shared__ TestTables TT;
  __shared__ u32 s_wcnt[EXP_THREADS / 32][RIES_SLOTS];
  __shared__ u32 s_slot_base[RIES_SLOTS];
  __shared__ u32 s_wl[EXP_THREADS / 32], s_wd[EXP_THREADS / 32];
  __shared__ u32 s_live_base, s_done_base, s_dead, s_tile, s_last, s_next;

  const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u32 lvl_begin = ctl->lvl_begin, lvl_end = ctl->lvl_end;
  const u32 n = lvl_end - lvl_begin;
  const u32 ntiles = (n + EXP_PARENTS - 1) / EXP_PARENTS;
  const DevPtrs D = *Dp;
  const int using_x = Ap->using_x, cplx = Ap->cplx;
  const u32 cap_nodes = Ap->cap_nodes, cap_done = Ap->cap_done;
  if (ntiles != 0 && !(ctl->overflow & (OVF_NODES | OVF_DONE))) {
    {
      const u32* src = (const u32*)D.P;
      u32* dst = (u32*)&P;
      for (u32 i = tid; i < sizeof(RiesParams) / 4; i += EXP_THREADS) dst[i] = src[i];
      const u32* tsrc = (const u32*)D.tt;
      u32* tdst = (u32*)&TT;
      for (u32 i = tid; i < sizeof(TestTables) / 4; i += EXP_THREADS) tdst[i] = tsrc[i];
      if (tid == 0) s_dead = 0;
      if (tid < RIES_MAX_SYMS) s_opcnt[tid] = 0;
    }
    __syncthreads();
    const u32 done_base0 = ctl->done_count; /* read-only during the tile loop (updated by the last block) */
    const u32 lt = (1u << lane) - 1;
    Node* nodes = D.nodes;
    const FormDesc* forms = D.forms;
    volatile u64* tile_state = (volatile u64*)D.tiles;
    u32 my_dead = 0;
    int cur = 0;
    bool have_prev = false;

    /* Tiles are handed out by ticket, so a tile's predecessors (which its look-back waits on) are always
       held by blocks that are already running: safe even when not every block of the grid is resident
       (several engines sharing the GPU on different streams).  The loop is software-pipelined: the front
       half of tile t+1 runs between the publication of tile t's aggregate and t's look-back. */
    while (true) {
      if (tid == 0) { s_tile = atomicAdd(&ctl->tile_ticket, 1u); s_next = 0; }
      __syncthreads();
      const u32 tile = s_tile;
      const bool have_cur = tile < ntiles;
      if (have_cur) {
        TileBuf& B = s_buf[cur];
        Staged* staged = D.staged + ((size_t)blockIdx.x * 2 + cur) * EXP_MAX_ITEMS;
        const u32 p0 = lvl_begin + tile * EXP_PARENTS;
        /* ---- phase 1: integer tests, one thread per prefix ---- */
        u32 mask = 0;
        {
          const u32 p = p0 + tid;
          if (p < lvl_end) {
            const Node par = nodes[p];
            B.par[tid] = par;
            const FormDesc& F = forms[par.form];
            const int cls = F.cls[par.len];
            const int ns = P.cls_n[cls];
            const int atts = prefix_attrs(P, F, par, using_x);
            mask = expand_mask(P, TT, F, par, cls, atts, using_x, cplx, cplx);
            my_dead += (u32)ns - (u32)__popc(mask);
            B.cls[tid] = (unsigned char)cls;
            B.complete[tid] = (unsigned char)(par.len + 1 >= F.flen);
          }
          B.live[tid] = 0;
          B.done[tid] = 0;
        }
        /* ---- phase 2: list the surviving pairs grouped by slot ---- */
#pragma unroll
        for (int k = 0; k < RIES_SLOTS; k++) {
          u32 b = __ballot_sync(0xffffffffu, (mask >> k) & 1u);
          if (lane == 0) s_wcnt[warp][k] = __popc(b);
        }
        __syncthreads();
        if (tid < RIES_SLOTS) { /* per slot: exclusive prefix over warps, and the slot's total */
          u32 run = 0;
          for (int w = 0; w < EXP_THREADS / 32; w++) { u32 c = s_wcnt[w][tid]; s_wcnt[w][tid] = run; run += c; }
          s_slot_base[tid] = run;
        }
        __syncthreads();
        if (tid == 0) {
          u32 run = 0;
          for (int k = 0; k < RIES_SLOTS; k++) { u32 c = s_slot_base[k]; s_slot_base[k] = run; run += c; }
          B.total = run;
          B.tile = tile;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < RIES_SLOTS; k++) {
          u32 b = __ballot_sync(0xffffffffu, (mask >> k) & 1u);
          if ((mask >> k) & 1u) B.item[s_slot_base[k] + s_wcnt[warp][k] + __popc(b & lt)] = (unsigned short)((tid << 4) | k);
        }
        __syncthreads();
        const u32 T = B.total;
        /* ---- phase 3: exec + value prunes; warps pull 32 listed pairs at a time (a warp's batch is
           almost always one operator, and heavy operators spread over the warps) ---- */
        while (true) {
          u32 b0 = 0;
          if (lane == 0) b0 = atomicAdd(&s_next, 32u);
          b0 = __shfl_sync(0xffffffffu, b0, 0);
          if (b0 >= T) break;
          const u32 i = b0 + lane;
          if (i < T) {
            const u32 it = B.item[i];
            const u32 pl = it >> 4, k = it & 15u;
            const int cls = B.cls[pl];
            const int g = P.trial[cls][k];
            Node child;
            atomicAdd(&s_opcnt[g], 1u); /* exec() calls per symbol: the algorithmic work of this kernel */
            int oc = expand_exec(P, cls, g, B.complete[pl] != 0, nodes, B.par[pl], p0 + pl, using_x, child);
            if (oc == OUT_DEAD) my_dead++;
            else if (oc == OUT_LIVE || oc == OUT_DONE) {
              Staged st;
              st.val = child.val; st.der = child.der; st.below = child.below; st.tag = child.tag; st.sp = child.sp;
              st.pad0 = 0; st.pad1 = 0;
              staged[i] = st;
              if (oc == OUT_LIVE) { atomicOr(&B.live[pl], 1u << k); B.item[i] = (unsigned short)(it | IT_LIVE); }
              else { atomicOr(&B.done[pl], 1u << k); B.item[i] = (unsigned short)(it | IT_DONE); }
            }
          }
        }
        __syncthreads();
        /* child offsets inside the tile, and the tile's aggregate -> published at once */
        {
          u32 nl = __popc(B.live[tid]), nd = __popc(B.done[tid]);
          u32 il = nl, id = nd;
          for (int o = 1; o < 32; o <<= 1) {
            u32 a = __shfl_up_sync(0xffffffffu, il, o), b = __shfl_up_sync(0xffffffffu, id, o);
            if (lane >= (u32)o) { il += a; id += b; }
          }
          if (lane == 31) { s_wl[warp] = il; s_wd[warp] = id; }
          __syncthreads();
          u32 wl = 0, wd = 0, la = 0, da = 0;
          for (u32 w = 0; w < EXP_THREADS / 32; w++) {
            const u32 a = s_wl[w], b = s_wd[w];
            if (w < warp) { wl += a; wd += b; }
            la += a; da += b;
          }
          B.lbase[tid] = wl + il - nl;
          B.dbase[tid] = wd + id - nd;
          if (tid == 0) {
            B.live_agg = la; B.done_agg = da;
            tile_publish(tile_state, tile, la, da);
          }
        }
      }
      /* ---- phase 4 of the PREVIOUS tile: ordered append ---- */
      if (have_prev) {
        TileBuf& B = s_buf[cur ^ 1];
        const Staged* staged = D.staged + ((size_t)blockIdx.x * 2 + (cur ^ 1)) * EXP_MAX_ITEMS;
        if (warp == 0) {
          u32 le, de;
          tile_lookback(tile_state, B.tile, B.live_agg, B.done_agg, le, de);
          if (lane == 0) { s_live_base = le; s_done_base = de; }
        }
        __syncthreads();
        const u32 T = B.total;
        const u32 p_tile = B.tile;
        (void)p_tile;
        for (u32 i = tid; i < T; i += EXP_THREADS) {
          const u32 it = B.item[i];
          if (!(it & (IT_LIVE | IT_DONE))) continue;
          const u32 pl = (it >> 4) & 1023u, k = it & 15u;
          const Node& par = B.par[pl];
          const int g = P.trial[B.cls[pl]][k];
          const Staged st = staged[i];
          Node child;
          child.val = st.val; child.der = st.der; child.below = st.below; child.tag = st.tag; child.sp = st.sp;
          expand_finish(P, par, g, child);
          const u32 below_k = (1u << k) - 1;
          if (it & IT_LIVE) {
            u32 pos = lvl_end + s_live_base + B.lbase[pl] + __popc(B.live[pl] & below_k);
            if (pos < cap_nodes) nodes[pos] = child;
          } else {
            u32 pos = done_base0 + s_done_base + B.dbase[pl] + __popc(B.done[pl] & below_k);
            if (pos < cap_done) {
              Row r; r.val = child.val; r.der = child.der; r.sym = child.sym;
              D.done[pos] = r;
              D.done_form[pos] = child.form;
            }
          }
        }
      }
      __syncthreads(); /* s_tile, s_live_base and the buffers are reused by the next round */
      if (!have_cur) break;
      have_prev = true;
      cur ^= 1;
    }
    /* block-level dead-end count -> one atomic per block */
    for (int o = 16; o > 0; o >>= 1) my_dead += __shfl_xor_sync(0xffffffffu, my_dead, o);
    if (lane == 0 && my_dead) atomicAdd(&s_dead, my_dead);
    __syncthreads();
    if (tid == 0 && s_dead) atomicAdd(&ctl->dead, (u64)s_dead);
    if (tid < RIES_MAX_SYMS && s_opcnt[tid]) atomicAdd(&D.opcnt[(using_x ? RIES_MAX_SYMS : 0) + tid], (u64)s_opcnt[tid]);
  }

  /* ---- level finish by the last block to get here (replaces a separate launch) ---- */
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    s_last = (atomicAdd(&ctl->blocks_done, 1u) == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  volatile u64* tile_state = (volatile u64*)D.tiles;
  const u64 last = ntiles ? tile_state[ntiles - 1] : 0;
  __syncthreads();
  for (u32 i = tid; i < ntiles; i += EXP_THREADS) tile_state[i] = 0;
  if (tid == 0) {
    u32 ovf = ctl->overflow;
    if (!(ovf & (OVF_NODES | OVF_DONE))) {
      u32 live = TS_LIVE(last), dn = TS_DONE(last);
      u64 new_nodes = (u64)lvl_end + live;
      u64 new_done = (u64)ctl->done_count + dn;
      if (new_nodes > cap_nodes) ovf |= OVF_NODES;
      if (new_done > cap_done) ovf |= OVF_DONE;
      ctl->overflow = ovf;
      ctl->lvl_begin = lvl_end;
      /* on overflow stop expanding: the host grows the arenas and reruns the pass */
      ctl->lvl_end = (ovf & (OVF_NODES | OVF_DONE)) ? lvl_end : (u32)new_nodes;
      ctl->node_count = (new_nodes > 0xffffffffull) ? 0xffffffffu : (u32)new_nodes;
      ctl->done_count = (new_done > 0xffffffffull) ? 0xffffffffu : (u32)new_done;
    }
    ctl->tile_ticket = 0;
    ctl->blocks_done = 0;
  }
}

/* fp64 FMA peak of this device, measured (denominator for K1's fp64 figure): 8 independent chains/thread */
__global__ void k_dfma_peak(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

/* ------------------------------------------------ survivors into insertion (DFS) order --- */
/* Within one level the completed expressions are ordered (form, lexicographic) and every form
   completes at exactly one level, so a form's expressions are contiguous in done[]. */
__device__ __forceinline__ bool pass_ok(const PassCtl* C) { return (C->overflow & (OVF_NODES | OVF_DONE)) == 0; }

__global__ void k_form_bounds(const DevPtrs* Dp, const PassArgs* A, const PassCtl* C) {
  if (!pass_ok(C)) return;
  const u32 n = C->done_count;
  const uint16_t* done_form = A->gathered ? Dp->gdone_form : Dp->done_form;
  u32* form_first = Dp->form_first;
  u32* form_cnt = Dp->form_cnt;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    uint16_t f = done_form[i];
    if (i == 0 || done_form[i - 1] != f) form_first[f] = i;
    if (i + 1 == n || done_form[i + 1] != f) form_cnt[f] = i + 1; /* end; fixed up in k_form_scan */
  }
}

__global__ void k_form_scan(const DevPtrs* Dp, const PassArgs* A, const PassCtl* C) {
  /* single block: the number of forms is at most a few thousand */
  __shared__ u32 s_carry;
  __shared__ u32 s_tmp[1024];
  if (!pass_ok(C) || C->done_count == 0) return;
  const u32 nforms = (u32)A->nforms;
  const u32* form_first = Dp->form_first;
  u32* form_cnt = Dp->form_cnt;
  u32* form_base = Dp->form_base;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (u32 base = 0; base < nforms; base += 1024) {
    u32 f = base + threadIdx.x;
    u32 c = 0;
    if (f < nforms) {
      u32 e = form_cnt[f];
      c = e ? (e - form_first[f]) : 0;
    }
    s_tmp[threadIdx.x] = c;
    __syncthreads();
    for (u32 o = 1; o < 1024; o <<= 1) {
      u32 v = (threadIdx.x >= o) ? s_tmp[threadIdx.x - o] : 0;
      __syncthreads();
      s_tmp[threadIdx.x] += v;
      __syncthreads();
    }
    if (f < nforms) form_base[f] = s_carry + s_tmp[threadIdx.x] - c;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry += s_tmp[1023];
    __syncthreads();
  }
}

__device__ __forceinline__ u64 key_of(double v) { /* order-preserving map double -> u64 */
  u64 b = (u64)__double_as_longlong(v);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

/* Survivors into insertion order, their sort keys, and (fused, so the keys are not read again) the
   histograms of all eight key digits for the radix sort below. */
__global__ void __launch_bounds__(256)
k_reorder(const DevPtrs* Dp, const PassArgs* A, const PassCtl* C) {
  __shared__ u32 s_h[8][256];
  if (!pass_ok(C)) return;
  const u32 n = C->done_count;
  if (blockIdx.x * blockDim.x >= n) return;
  const DevPtrs D = *Dp;
  const Row* __restrict__ done = A->gathered ? D.gdone : D.done;
  const uint16_t* __restrict__ done_form = A->gathered ? D.gdone_form : D.done_form;
  for (int d = 0; d < 8; d++) s_h[d][threadIdx.x] = 0;
  __syncthreads();
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    uint16_t f = done_form[i];
    u32 pos = D.form_base[f] + (i - D.form_first[f]);
    Row r = done[i];
    D.surv[pos] = r;
    D.surv_form[pos] = f;
    const u64 k = key_of(r.val);
    D.keys[0][pos] = k;
    D.idx[0][pos] = pos;
#pragma unroll
    for (int d = 0; d < 8; d++) atomicAdd(&s_h[d][(u32)(k >> (8 * d)) & 255u], 1u);
  }
  __syncthreads();
  SortCtl* ctl = D.sortctl;
  for (int d = 0; d < 8; d++) {
    u32 c = s_h[d][threadIdx.x];
    if (c) atomicAdd(&ctl->ghist[d][threadIdx.x], c);
  }
}

/* ---------------------------------------------------------------- K2: radix sort --------- */
/* Stable LSD radix sort of (u64 key, u32 payload), 8 digits of 8 bits, "onesweep" style: the histograms
   of all eight digits come from k_reorder (which writes the keys); each digit pass is then a single
   kernel in which every 4096-key chunk gets its per-digit output offsets by decoupled look-back over the
   chunks before it (chunks are taken in ticket order, so equal keys keep their input order = insertion
   order). */
#define RS_THREADS 256
#define RS_WARPS (RS_THREADS / 32)
#define RS_ITEMS 16
#define RS_CHUNK (RS_THREADS * RS_ITEMS)
#define RS_FLAG_AGG 1ull
#define RS_FLAG_INC 2ull
#define RS_DYN_SMEM (RS_CHUNK * (sizeof(u64) + sizeof(u32)))

/* One digit pass.  Per 4096-key chunk:
     rank     each warp ranks its own 512 keys (warp-striped, 16 rounds of match_any against a private
              histogram): no block barrier inside
     offsets  per bin: prefix over the warps, exclusive scan over the bins, look-back over earlier chunks
     exchange keys go to shared memory in their chunk-local sorted position, and are written out from
              there so that neighbouring threads write neighbouring addresses of the same bin
   The order inside a bin is (chunk, warp, round, lane) = input order, so the sort is stable. */
__global__ void __launch_bounds__(RS_THREADS, 2)
k_rs_onesweep(const DevPtrs* Dp, const PassCtl* C, const TableState* T, int pass) {
  extern __shared__ __align__(16) unsigned char s_rs_dyn[];
  u64* s_keys = (u64*)s_rs_dyn;
  u32* s_pays = (u32*)(s_keys + RS_CHUNK);
  __shared__ u32 s_whist[RS_WARPS][256];
  __shared__ u32 s_binex[256]; /* chunk-local exclusive offset of a bin */
  __shared__ u32 s_goff[256];  /* global offset of a bin's first key of this chunk, minus s_binex */
  __shared__ u32 s_wtot[RS_WARPS];
  __shared__ u32 s_chunk;
  if (!pass_ok(C)) return;
  const u32 n = C->done_count;
  const u32 nchunks = (n + RS_CHUNK - 1) / RS_CHUNK;
  if (blockIdx.x >= nchunks) return;
  const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u32 lt = (1u << lane) - 1;
  const int shift = pass * 8;
  const int src = pass & 1;
  const u64* __restrict__ keys_in = Dp->keys[src];
  const u32* __restrict__ idx_in = Dp->idx[src];
  u64* keys_out = Dp->keys[src ^ 1];
  u32* idx_out = Dp->idx[src ^ 1];
  SortCtl* ctl = Dp->sortctl;
  volatile u64* state = (volatile u64*)Dp->rs_state;
  const u64 epoch = (u64)(T->epoch + (u32)pass);
  /* exclusive scan of the digit's global histogram: bin `tid` starts at gbase */
  u32 gbase;
  {
    const u32 mine = ctl->ghist[pass][tid];
    u32 inc = mine;
    for (int o = 1; o < 32; o <<= 1) {
      u32 v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= (u32)o) inc += v;
    }
    if (lane == 31) s_wtot[warp] = inc;
    __syncthreads();
    u32 wbase = 0;
    for (u32 w = 0; w < warp; w++) wbase += s_wtot[w];
    gbase = wbase + inc - mine;
    __syncthreads();
  }

  while (true) {
    if (tid == 0) s_chunk = atomicAdd(&ctl->ticket[pass], 1u);
    for (u32 j = tid; j < RS_WARPS * 256; j += RS_THREADS) (&s_whist[0][0])[j] = 0;
    __syncthreads();
    const u32 chunk = s_chunk;
    if (chunk >= nchunks) break;
    const u32 base = chunk * RS_CHUNK;
    const u32 m = min(RS_CHUNK, n - base); /* keys in this chunk */
    u64 key[RS_ITEMS];
    u32 pay[RS_ITEMS];
    u32 rk[RS_ITEMS]; /* rank of the key among its warp's keys of the same digit */
    const u32 wbase_i = warp * (RS_ITEMS * 32) + lane;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
      const u32 li = wbase_i + r * 32;
      const bool valid = li < m;
      key[r] = valid ? keys_in[base + li] : 0;
      pay[r] = valid ? idx_in[base + li] : 0;
    }
    u32* wh = s_whist[warp];
    /* all the match_any's first (independent of each other), then the short dependent chain through
       the warp's histogram: every peer reads the running count, the leader bumps it */
    u32 peers[RS_ITEMS];
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
      const bool valid = (wbase_i + r * 32) < m;
      const u32 d = (u32)(key[r] >> shift) & 255u;
      peers[r] = __match_any_sync(0xffffffffu, valid ? d : (0x100u | lane));
    }
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
      const bool valid = (wbase_i + r * 32) < m;
      const u32 d = (u32)(key[r] >> shift) & 255u;
      const u32 prev = valid ? wh[d] : 0;
      __syncwarp();
      if (valid && (peers[r] & lt) == 0) wh[d] = prev + __popc(peers[r]);
      __syncwarp();
      rk[r] = prev + __popc(peers[r] & lt);
    }
    __syncthreads();
    /* bin `tid`: exclusive prefix over the warps; its count in the chunk */
    u32 cnt = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; w++) { u32 c = s_whist[w][tid]; s_whist[w][tid] = cnt; cnt += c; }
    /* publish the chunk's count of this bin right away; the look-back comes after the exchange, when
       the chunks before this one have had time to publish theirs */
    volatile u64* st = state + (size_t)chunk * 256 + tid;
    const u64 tag = epoch << 34;
    *st = tag | ((chunk == 0 ? RS_FLAG_INC : RS_FLAG_AGG) << 32) | cnt;
    /* exclusive scan over the bins */
    u32 inc = cnt;
    for (int o = 1; o < 32; o <<= 1) {
      u32 v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= (u32)o) inc += v;
    }
    if (lane == 31) s_wtot[warp] = inc;
    __syncthreads();
    u32 binex;
    {
      u32 wb = 0;
      for (u32 w = 0; w < warp; w++) wb += s_wtot[w];
      binex = wb + inc - cnt;
      s_binex[tid] = binex;
    }
    __syncthreads();
    /* exchange through shared memory */
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
      if ((wbase_i + r * 32) < m) {
        const u32 d = (u32)(key[r] >> shift) & 255u;
        const u32 lp = s_binex[d] + wh[d] + rk[r];
        s_keys[lp] = key[r];
        s_pays[lp] = pay[r];
      }
    }
    /* decoupled look-back for bin `tid` over the chunks before this one */
    {
      u32 ex = 0;
      if (chunk != 0) {
        int c = (int)chunk - 1;
        while (c >= 0) {
          u64 w;
          do { w = state[(size_t)c * 256 + tid]; } while ((w >> 34) != epoch || ((w >> 32) & 3ull) == 0);
          ex += (u32)w;
          if (((w >> 32) & 3ull) == RS_FLAG_INC) break;
          c--;
        }
        *st = tag | (RS_FLAG_INC << 32) | (u64)(ex + cnt);
      }
      s_goff[tid] = gbase + ex - binex;
    }
    __syncthreads();
#pragma unroll 4
    for (u32 j = tid; j < m; j += RS_THREADS) {
      const u64 k = s_keys[j];
      const u32 dest = s_goff[(u32)(k >> shift) & 255u] + j;
      keys_out[dest] = k;
      idx_out[dest] = s_pays[j];
    }
  }
}

/* ---------------------------------------------------------------- dedupe / merge / match - */

__device__ __forceinline__ u32 lower_bound_val(const double* __restrict__ a, u32 n, double v) {
  u32 lo = 0, hi = n;
  while (lo < hi) {
    u32 mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ void push_event(MatchEvent* events, u32 cap, PassCtl* ctl, u32 seq, u32 kind,
                                           const Row& L, const Row& R) {
  u32 e = atomicAdd(&ctl->n_events, 1u);
  if (e < cap) {
    MatchEvent ev;
    ev.seq = seq; ev.kind = kind; ev.lhs = L; ev.rhs = R;
    ev.lhs_s.lo = ev.lhs_s.hi = ev.rhs_s.lo = ev.rhs_s.hi = 0;
    ev.root = 0; ev.diff_dx = 0; ev.newton_err = 0; ev.pad = 0;
    events[e] = ev;
  } else {
    atomicOr(&ctl->overflow, OVF_EVENTS);
  }
}

/* Inverse of key_of. */
__device__ __forceinline__ double val_of_key(u64 k) {
  u64 b = (k & 0x8000000000000000ull) ? (k ^ 0x8000000000000000ull) : ~k;
  return __longlong_as_double((long long)b);
}

/* Lower bound by a whole warp: 32 probes per step split the range in 33, so a 3e7-row table takes five
   dependent loads instead of twenty-five. */
__device__ __forceinline__ u32 warp_lower_bound(const double* __restrict__ a, u32 n, double v) {
  const u32 lane = threadIdx.x & 31;
  u32 lo = 0, hi = n; /* the answer is in [lo, hi] */
  while (hi - lo > 32) {
    const u32 step = (hi - lo + 32) / 33;
    const u32 q = lo + (lane + 1) * step - 1;
    const bool lt = (q < hi) && (a[q] < v);
    const u32 c = __popc(__ballot_sync(0xffffffffu, lt));
    const u32 nhi = lo + (c + 1) * step - 1;
    lo += c * step;
    if (c < 32 && nhi < hi) hi = nhi; /* probe c exists, is in range and is >= v */
  }
  const bool lt = (lo + lane < hi) && (a[lo + lane] < v);
  return lo + __popc(__ballot_sync(0xffffffffu, lt));
}

/* For every survivor in sorted order: is it a new distinct value?  New rows are compacted (ticketed
   chunks + look-back keep them in value order) together with their gap = lower bound in the old
   table.  Duplicates of an opposite-side table row become EXACT events (check_exact_match's side
   test, ries.c:8550-8593).
   The survivors of a chunk are consecutive in value order, so their lower bounds all fall between the
   lower bounds of the chunk's first and last value: two warp searches find that window, the window of
   table keys is staged in shared memory with coalesced loads, and the per-survivor searches run there.
   Values come straight from the sorted keys (the key map is invertible): no gather from surv[]. */
#define CC_THREADS 256
#define CC_ITEMS 8
#define CC_CHUNK (CC_THREADS * CC_ITEMS)
#define CC_WIN 4096
__global__ void __launch_bounds__(CC_THREADS)
k_classify_compact(const DevPtrs* Dp, const PassArgs* A, PassCtl* C, const TableState* T) {
  __shared__ double s_win[CC_WIN + 1];
  __shared__ u32 s_cnt[CC_ITEMS * (CC_THREADS / 32)];
  __shared__ u32 s_chunk, s_base, s_lo, s_hi;
  if (!pass_ok(C)) return;
  const u32 n = C->done_count;
  const u32 nchunks = (n + CC_CHUNK - 1) / CC_CHUNK;
  if (blockIdx.x >= nchunks) return;
  const DevPtrs D = *Dp;
  const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u32 cur = T->cur, nt = T->nt;
  const double* __restrict__ tval = D.tval[cur];
  const Row* __restrict__ trow = D.trow[cur];
  const u64* __restrict__ keys = D.keys[0]; /* 8 digit passes: the sorted order ends in buffer 0 */
  const u32* __restrict__ idx = D.idx[0];
  const Row* __restrict__ surv = D.surv;
  const int using_x = A->using_x, skip_exact = A->skip_exact;
  volatile u64* state = (volatile u64*)D.cc_state;
  const u64 epoch = (u64)(T->epoch + 8u);
  while (true) {
    if (tid == 0) s_chunk = atomicAdd(&C->chunk_ticket, 1u);
    __syncthreads();
    const u32 chunk = s_chunk;
    if (chunk >= nchunks) break;
    const u32 base = chunk * CC_CHUNK;
    const u32 last = min(base + CC_CHUNK, n) - 1;
    /* window of the table that can hold this chunk's lower bounds */
    if (warp == 0) {
      u32 r = warp_lower_bound(tval, nt, val_of_key(keys[base]));
      if (lane == 0) s_lo = r;
    } else if (warp == 1) {
      u32 r = warp_lower_bound(tval, nt, val_of_key(keys[last]));
      if (lane == 0) s_hi = r;
    }
    __syncthreads();
    const u32 lo = s_lo, hi = s_hi;        /* every lower bound of the chunk is in [lo, hi] */
    const u32 wn = hi - lo;                 /* keys that take part in the searches */
    const u32 wload = min(hi + 1, nt) - lo; /* + the one at `hi`, for the equality test */
    const bool in_smem = wn <= CC_WIN;
    if (in_smem)
      for (u32 j = tid; j < wload; j += CC_THREADS) s_win[j] = tval[lo + j];
    __syncthreads();
    const double* win = in_smem ? s_win : tval + lo;
    u32 pos[CC_ITEMS], sq[CC_ITEMS];
    u32 flags = 0;
#pragma unroll
    for (int r = 0; r < CC_ITEMS; r++) {
      const u32 i = base + r * CC_THREADS + tid;
      bool fl = false;
      pos[r] = 0; sq[r] = 0;
      if (i < n) {
        const u64 k = keys[i];
        const double v = val_of_key(k);
        const u32 s = idx[i];
        sq[r] = s;
        const bool dup_prev = (i > 0) && (keys[i - 1] == k);
        u32 a = 0, b = wn;
        while (a < b) {
          u32 mid = (a + b) >> 1;
          if (win[mid] < v) a = mid + 1; else b = mid;
        }
        const u32 p = lo + a;
        pos[r] = p;
        const bool in_table = (p < nt) && (win[a] == v);
        if (in_table) {
          if (!skip_exact) {
            Row t = trow[p];
            bool tree_is_lhs = (t.der != 0.0);
            if (using_x ? !tree_is_lhs : tree_is_lhs) {
              Row f = surv[s];
              if (using_x) push_event(D.events, A->cap_events, C, s, EV_EXACT, f, t);
              else push_event(D.events, A->cap_events, C, s, EV_EXACT, t, f);
            }
          }
        } else if (!dup_prev) {
          fl = true;
        }
      }
      const u32 b = __ballot_sync(0xffffffffu, fl);
      if (lane == 0) s_cnt[r * (CC_THREADS / 32) + warp] = __popc(b);
      if (fl) flags |= 1u << r;
    }
    __syncthreads();
    /* exclusive scan of the 64 (row, warp) counts in row-major order = survivor order */
    u32 total = 0;
    if (warp == 0) {
      const u32 c0 = s_cnt[2 * lane], c1 = s_cnt[2 * lane + 1];
      u32 inc = c0 + c1;
      for (int o = 1; o < 32; o <<= 1) {
        u32 v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (u32)o) inc += v;
      }
      s_cnt[2 * lane] = inc - c0 - c1;
      s_cnt[2 * lane + 1] = inc - c1;
      total = __shfl_sync(0xffffffffu, inc, 31);
      /* look-back (one word per chunk: tag | flag | count) */
      if (lane == 0) {
        const u64 tag = epoch << 34;
        u32 ex = 0;
        if (chunk == 0) {
          state[0] = tag | (RS_FLAG_INC << 32) | total;
        } else {
          state[chunk] = tag | (RS_FLAG_AGG << 32) | total;
          int c = (int)chunk - 1;
          while (c >= 0) {
            u64 w;
            do { w = state[c]; } while ((w >> 34) != epoch || ((w >> 32) & 3ull) == 0);
            ex += (u32)w;
            if (((w >> 32) & 3ull) == RS_FLAG_INC) break;
            c--;
          }
          state[chunk] = tag | (RS_FLAG_INC << 32) | (u64)(ex + total);
        }
        s_base = ex;
        if (chunk == nchunks - 1) C->n_uniq = ex + total;
      }
    }
    __syncthreads();
    const u32 cbase = s_base;
#pragma unroll
    for (int r = 0; r < CC_ITEMS; r++) {
      const bool fl = (flags >> r) & 1u;
      const u32 b = __ballot_sync(0xffffffffu, fl);
      if (fl) {
        const u32 u = cbase + s_cnt[r * (CC_THREADS / 32) + warp] + __popc(b & ((1u << lane) - 1));
        D.uniq_seq[u] = sq[r];
        D.uniq_gap[u] = pos[r];
      }
    }
    __syncthreads();
  }
}

/* K3 + merge of the new rows.  A new row's predecessor AT ITS INSERTION TIME is the old-table row below
   its gap unless a same-pass row with a smaller sequence number already sits between them (that one is
   same-side, so no test happens: check_sides, ries.c:8496-8538).  Likewise for the successor.  The row is
   read once: it goes to its place in the merged table and into the neighbour tests. */
__global__ void k_match_merge_new(const DevPtrs* Dp, const PassArgs* A, PassCtl* C, const TableState* T) {
  if (!pass_ok(C)) return;
  const u32 nu = C->n_uniq;
  const DevPtrs D = *Dp;
  const u32 nt = T->nt;
  const Row* __restrict__ trow = D.trow[T->cur];
  Row* mrow = D.trow[T->cur ^ 1];
  double* mval = D.tval[T->cur ^ 1];
  const u32* __restrict__ uniq_seq = D.uniq_seq;
  const u32* __restrict__ uniq_gap = D.uniq_gap;
  const int using_x = A->using_x;
  const double best_match0 = A->best_match0;
  for (u32 u = blockIdx.x * blockDim.x + threadIdx.x; u < nu; u += gridDim.x * blockDim.x) {
    const u32 s = uniq_seq[u], g = uniq_gap[u];
    const Row f = D.surv[s];
    mrow[u + g] = f;
    mval[u + g] = f.val;
    bool test_prev = g > 0, test_next = g < nt;
    if (test_prev) {
      for (u32 j = u; j-- > 0 && uniq_gap[j] == g;)
        if (uniq_seq[j] < s) { test_prev = false; break; }
    }
    if (test_next) {
      for (u32 j = u + 1; j < nu && uniq_gap[j] == g; j++)
        if (uniq_seq[j] < s) { test_next = false; break; }
    }
    for (int side = 0; side < 2; side++) {
      if (side == 0 ? !test_prev : !test_next) continue;
      const Row o = trow[side == 0 ? g - 1 : g];
      bool other_is_lhs = (o.der != 0.0);
      if (using_x ? other_is_lhs : !other_is_lhs) continue;
      const Row& L = using_x ? f : o;
      const Row& R = using_x ? o : f;
      double total = L.der - R.der;
      double score = fabs((R.val - L.val) / total);
      if (score > best_match0) continue; /* NaN passes, as in the reference's test (8270) */
      push_event(D.events, A->cap_events, C, s, side == 0 ? EV_PREV : EV_NEXT, L, R);
    }
  }
}

/* ------------------------------------------------ small passes: one block does it all ---- */
/* A pass with at most SP_CAP complete expressions (most passes of a search: sizes grow ~1.3x per
   complexity step from one expression up) is launch-latency bound in the general kernels, so its
   regroup + sort + classify + match + merge-new chain runs in ONE block here, with the keys resident in
   shared memory.  Same algorithm, same order of results.  Insertion order is recovered by a stable
   sort on the form index (every form completes at exactly one level, levels are appended in order). */
#define SP_THREADS 1024
#define SP_WARPS (SP_THREADS / 32)
#define SP_ITEMS 4
#define SP_CAP (SP_THREADS * SP_ITEMS)
#define SP_DYN_SMEM (SP_CAP * (sizeof(u64) + sizeof(u32)) + SP_WARPS * 256 * sizeof(u32))

/* one stable 8-bit digit pass over m keys held in shared memory (in place) */
__device__ __forceinline__ void small_sort_pass(u64* s_keys, u32* s_pays, u32* s_whist, u32* s_binex,
                                                u32* s_wtot, u32 m, int shift) {
  const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u32 lt = (1u << lane) - 1;
  for (u32 j = tid; j < SP_WARPS * 256; j += SP_THREADS) s_whist[j] = 0;
  u64 key[SP_ITEMS];
  u32 pay[SP_ITEMS], rk[SP_ITEMS], peers[SP_ITEMS];
  const u32 li0 = warp * (SP_ITEMS * 32) + lane;
#pragma unroll
  for (int r = 0; r < SP_ITEMS; r++) {
    const u32 li = li0 + r * 32;
    key[r] = li < m ? s_keys[li] : 0;
    pay[r] = li < m ? s_pays[li] : 0;
  }
  __syncthreads();
  u32* wh = s_whist + warp * 256;
#pragma unroll
  for (int r = 0; r < SP_ITEMS; r++) {
    const bool valid = (li0 + r * 32) < m;
    const u32 d = (u32)(key[r] >> shift) & 255u;
    peers[r] = __match_any_sync(0xffffffffu, valid ? d : (0x100u | lane));
  }
#pragma unroll
  for (int r = 0; r < SP_ITEMS; r++) {
    const bool valid = (li0 + r * 32) < m;
    const u32 d = (u32)(key[r] >> shift) & 255u;
    const u32 prev = valid ? wh[d] : 0;
    __syncwarp();
    if (valid && (peers[r] & lt) == 0) wh[d] = prev + __popc(peers[r]);
    __syncwarp();
    rk[r] = prev + __popc(peers[r] & lt);
  }
  __syncthreads();
  u32 cnt = 0, inc = 0;
  if (tid < 256) {
    for (int w = 0; w < SP_WARPS; w++) { u32 c = s_whist[w * 256 + tid]; s_whist[w * 256 + tid] = cnt; cnt += c; }
    inc = cnt;
    for (int o = 1; o < 32; o <<= 1) {
      u32 v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= (u32)o) inc += v;
    }
    if (lane == 31) s_wtot[warp] = inc;
  }
  __syncthreads();
  if (tid < 256) {
    u32 wb = 0;
    for (u32 w = 0; w < warp; w++) wb += s_wtot[w];
    s_binex[tid] = wb + inc - cnt;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < SP_ITEMS; r++) {
    if ((li0 + r * 32) < m) {
      const u32 d = (u32)(key[r] >> shift) & 255u;
      const u32 lp = s_binex[d] + wh[d] + rk[r];
      s_keys[lp] = key[r];
      s_pays[lp] = pay[r];
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(SP_THREADS, 1)
k_small_post(const DevPtrs* Dp, const PassArgs* A, PassCtl* C, const TableState* T) {
  extern __shared__ __align__(16) unsigned char s_sp_dyn[];
  u64* s_keys = (u64*)s_sp_dyn;
  u32* s_pays = (u32*)(s_keys + SP_CAP);
  u32* s_whist = s_pays + SP_CAP;      /* SP_WARPS x 256 while sorting ... */
  u32* s_useq = s_whist;               /* ... then the new rows: sequence number and gap */
  u32* s_ugap = s_whist + SP_CAP;
  __shared__ u32 s_binex[256];
  __shared__ u32 s_wtot[8];
  __shared__ u32 s_cnt[SP_ITEMS * SP_WARPS];
  __shared__ u32 s_nu;
  if (!pass_ok(C)) return;
  const u32 n = C->done_count;
  if (n == 0) return;
  if (n > SP_CAP) { /* the host's size prediction was wrong: it reruns the pass through the general kernels */
    if (threadIdx.x == 0) atomicOr(&C->overflow, OVF_SMALL);
    return;
  }
  const DevPtrs D = *Dp;
  const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  /* insertion order: stable sort of the done[] indices by form */
  const Row* __restrict__ done = A->gathered ? D.gdone : D.done;
  const uint16_t* __restrict__ done_form = A->gathered ? D.gdone_form : D.done_form;
  for (u32 i = tid; i < n; i += SP_THREADS) { s_keys[i] = done_form[i]; s_pays[i] = i; }
  __syncthreads();
  small_sort_pass(s_keys, s_pays, s_whist, s_binex, s_wtot, n, 0);
  if (A->nforms > 256) small_sort_pass(s_keys, s_pays, s_whist, s_binex, s_wtot, n, 8);
  for (u32 j = tid; j < n; j += SP_THREADS) {
    const Row r = done[s_pays[j]];
    D.surv[j] = r;
    s_keys[j] = key_of(r.val);
    s_pays[j] = j;
  }
  __syncthreads();
  /* stable sort by value */
  for (int pass = 0; pass < 8; pass++) small_sort_pass(s_keys, s_pays, s_whist, s_binex, s_wtot, n, pass * 8);
  /* classify: new distinct values, EXACT events */
  const u32 cur = T->cur, nt = T->nt;
  const double* __restrict__ tval = D.tval[cur];
  const Row* __restrict__ trow = D.trow[cur];
  const int using_x = A->using_x, skip_exact = A->skip_exact;
  u32 pos[SP_ITEMS], sq[SP_ITEMS], flags = 0;
  {
    /* the thread's SP_ITEMS lower bounds advance in lock step: SP_ITEMS independent loads per round */
    double v[SP_ITEMS];
    u32 lo[SP_ITEMS], hi[SP_ITEMS];
#pragma unroll
    for (int r = 0; r < SP_ITEMS; r++) {
      const u32 i = r * SP_THREADS + tid;
      v[r] = (i < n) ? val_of_key(s_keys[i]) : 0.0;
      lo[r] = 0;
      hi[r] = (i < n) ? nt : 0;
    }
    for (u32 span = nt; span != 0; span >>= 1) {
      double t[SP_ITEMS];
#pragma unroll
      for (int r = 0; r < SP_ITEMS; r++) t[r] = (lo[r] < hi[r]) ? tval[(lo[r] + hi[r]) >> 1] : 0.0;
#pragma unroll
      for (int r = 0; r < SP_ITEMS; r++) {
        if (lo[r] < hi[r]) {
          const u32 mid = (lo[r] + hi[r]) >> 1;
          if (t[r] < v[r]) lo[r] = mid + 1; else hi[r] = mid;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < SP_ITEMS; r++) pos[r] = lo[r];
  }
#pragma unroll
  for (int r = 0; r < SP_ITEMS; r++) {
    const u32 i = r * SP_THREADS + tid;
    bool fl = false;
    sq[r] = 0;
    if (i < n) {
      const u64 k = s_keys[i];
      const double v = val_of_key(k);
      const u32 sidx = s_pays[i];
      sq[r] = sidx;
      const bool dup_prev = (i > 0) && (s_keys[i - 1] == k);
      const u32 p = pos[r];
      const bool in_table = (p < nt) && (tval[p] == v);
      if (in_table) {
        if (!skip_exact) {
          Row t = trow[p];
          bool tree_is_lhs = (t.der != 0.0);
          if (using_x ? !tree_is_lhs : tree_is_lhs) {
            Row f = D.surv[sidx];
            if (using_x) push_event(D.events, A->cap_events, C, sidx, EV_EXACT, f, t);
            else push_event(D.events, A->cap_events, C, sidx, EV_EXACT, t, f);
          }
        }
      } else if (!dup_prev) {
        fl = true;
      }
    }
    const u32 b = __ballot_sync(0xffffffffu, fl);
    if (lane == 0) s_cnt[r * SP_WARPS + warp] = __popc(b);
    if (fl) flags |= 1u << r;
  }
  __syncthreads(); /* also: every read of s_keys/s_pays/s_whist is done, s_useq/s_ugap may be written */
  if (warp == 0) { /* exclusive scan of the SP_ITEMS*SP_WARPS = 128 counts, 4 per lane */
    u32 c[4], t = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) { c[q] = s_cnt[4 * lane + q]; t += c[q]; }
    u32 inc = t;
    for (int o = 1; o < 32; o <<= 1) {
      u32 v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= (u32)o) inc += v;
    }
    u32 run = inc - t;
#pragma unroll
    for (int q = 0; q < 4; q++) { s_cnt[4 * lane + q] = run; run += c[q]; }
    if (lane == 31) { s_nu = inc; C->n_uniq = inc; }
  }
  __syncthreads();
  const u32 nu = s_nu;
#pragma unroll
  for (int r = 0; r < SP_ITEMS; r++) {
    const bool fl = (flags >> r) & 1u;
    const u32 b = __ballot_sync(0xffffffffu, fl);
    if (fl) {
      const u32 u = s_cnt[r * SP_WARPS + warp] + __popc(b & ((1u << lane) - 1));
      s_useq[u] = sq[r]; s_ugap[u] = pos[r];
      D.uniq_seq[u] = sq[r]; D.uniq_gap[u] = pos[r];
    }
  }
  __syncthreads();
  /* match (K3) and merge of the new rows */
  const double best_match0 = A->best_match0;
  Row* mrow = D.trow[cur ^ 1];
  double* mval = D.tval[cur ^ 1];
  for (u32 u = tid; u < nu; u += SP_THREADS) {
    const u32 sidx = s_useq[u], g = s_ugap[u];
    const Row f = D.surv[sidx];
    mrow[u + g] = f;
    mval[u + g] = f.val;
    bool test_prev = g > 0, test_next = g < nt;
    if (test_prev) {
      for (u32 j = u; j-- > 0 && s_ugap[j] == g;)
        if (s_useq[j] < sidx) { test_prev = false; break; }
    }
    if (test_next) {
      for (u32 j = u + 1; j < nu && s_ugap[j] == g; j++)
        if (s_useq[j] < sidx) { test_next = false; break; }
    }
    for (int side = 0; side < 2; side++) {
      if (side == 0 ? !test_prev : !test_next) continue;
      const Row o = trow[side == 0 ? g - 1 : g];
      bool other_is_lhs = (o.der != 0.0);
      if (using_x ? other_is_lhs : !other_is_lhs) continue;
      const Row& L = using_x ? f : o;
      const Row& R = using_x ? o : f;
      double total = L.der - R.der;
      double score = fabs((R.val - L.val) / total);
      if (score > best_match0) continue; /* NaN passes, as in the reference's test (8270) */
      push_event(D.events, A->cap_events, C, sidx, side == 0 ? EV_PREV : EV_NEXT, L, R);
    }
  }
}

/* ------------------------------------------------ sharded passes: local pre-dedupe -------- */
/* One search over several GPUs: before the ranks exchange their complete expressions, each drops the ones
   that cannot matter, using its replica of the table and its own survivors in sorted order:
     value already in the table on the SAME side      -> dropped (check_exact_match discards it, 8550-8593)
     value already in the table on the opposite side  -> kept, every occurrence (each is an EXACT event)
     new value                                        -> only the first occurrence in insertion order is kept
   The global first-inserted winner of a value is the minimum over the ranks' first occurrences, so the
   replicated dedupe that follows sees exactly the rows it needs, in the same relative order. */
__global__ void __launch_bounds__(CC_THREADS)
k_local_filter(const DevPtrs* Dp, const PassArgs* A, PassCtl* C, const TableState* T) {
  __shared__ double s_win[CC_WIN + 1];
  __shared__ u32 s_lo, s_hi;
  if (!pass_ok(C)) return;
  const u32 n = C->done_count;
  const u32 nchunks = (n + CC_CHUNK - 1) / CC_CHUNK;
  const DevPtrs D = *Dp;
  const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u32 cur = T->cur, nt = T->nt;
  const double* __restrict__ tval = D.tval[cur];
  const Row* __restrict__ trow = D.trow[cur];
  const u64* __restrict__ keys = D.keys[0];
  const u32* __restrict__ idx = D.idx[0];
  const int using_x = A->using_x;
  for (u32 chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
    const u32 base = chunk * CC_CHUNK;
    const u32 last = min(base + CC_CHUNK, n) - 1;
    if (warp == 0) {
      u32 r = warp_lower_bound(tval, nt, val_of_key(keys[base]));
      if (lane == 0) s_lo = r;
    } else if (warp == 1) {
      u32 r = warp_lower_bound(tval, nt, val_of_key(keys[last]));
      if (lane == 0) s_hi = r;
    }
    __syncthreads();
    const u32 lo = s_lo, hi = s_hi;
    const u32 wn = hi - lo;
    const u32 wload = min(hi + 1, nt) - lo;
    const bool in_smem = wn <= CC_WIN;
    if (in_smem)
      for (u32 j = tid; j < wload; j += CC_THREADS) s_win[j] = tval[lo + j];
    __syncthreads();
    const double* win = in_smem ? s_win : tval + lo;
#pragma unroll
    for (int r = 0; r < CC_ITEMS; r++) {
      const u32 i = base + r * CC_THREADS + tid;
      if (i < n) {
        const u64 k = keys[i];
        const double v = val_of_key(k);
        const bool dup_prev = (i > 0) && (keys[i - 1] == k);
        u32 a = 0, b = wn;
        while (a < b) {
          u32 mid = (a + b) >> 1;
          if (win[mid] < v) a = mid + 1; else b = mid;
        }
        const u32 p = lo + a;
        bool kp;
        if ((p < nt) && (win[a] == v)) {
          const bool tree_is_lhs = (trow[p].der != 0.0);
          kp = using_x ? !tree_is_lhs : tree_is_lhs;
        } else {
          kp = !dup_prev;
        }
        D.keep[idx[i]] = kp ? 1 : 0;
      }
    }
    __syncthreads();
  }
}

/* Kept rows, compacted in insertion order (chunks by ticket, look-back over the chunk counts). */
__global__ void __launch_bounds__(256)
k_compact_kept(const DevPtrs* Dp, PassCtl* C, const TableState* T) {
  __shared__ u32 s_warp[8];
  __shared__ u32 s_chunk, s_base;
  if (!pass_ok(C)) return;
  const u32 n = C->done_count;
  const u32 nchunks = (n + 255) / 256;
  if (blockIdx.x >= nchunks) return;
  const DevPtrs D = *Dp;
  const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  volatile u64* state = (volatile u64*)D.cc_state;
  const u64 epoch = (u64)(T->epoch + 9u);
  while (true) {
    if (tid == 0) s_chunk = atomicAdd(&C->aux_ticket, 1u);
    __syncthreads();
    const u32 chunk = s_chunk;
    if (chunk >= nchunks) break;
    const u32 i = chunk * 256 + tid;
    const bool fl = (i < n) && D.keep[i];
    const u32 b = __ballot_sync(0xffffffffu, fl);
    if (lane == 0) s_warp[warp] = __popc(b);
    __syncthreads();
    u32 woff = 0, total = 0;
    for (u32 w = 0; w < 8; w++) { u32 c = s_warp[w]; if (w < warp) woff += c; total += c; }
    if (tid == 0) {
      const u64 tag = epoch << 34;
      u32 ex = 0;
      if (chunk == 0) {
        state[0] = tag | (RS_FLAG_INC << 32) | total;
      } else {
        state[chunk] = tag | (RS_FLAG_AGG << 32) | total;
        int c = (int)chunk - 1;
        while (c >= 0) {
          u64 w;
          do { w = state[c]; } while ((w >> 34) != epoch || ((w >> 32) & 3ull) == 0);
          ex += (u32)w;
          if (((w >> 32) & 3ull) == RS_FLAG_INC) break;
          c--;
        }
        state[chunk] = tag | (RS_FLAG_INC << 32) | (u64)(ex + total);
      }
      s_base = ex;
      if (chunk == nchunks - 1) C->n_kept = ex + total;
    }
    __syncthreads();
    if (fl) {
      const u32 u = s_base + woff + __popc(b & ((1u << lane) - 1));
      D.ksend[u] = D.surv[i];
      D.ksend_form[u] = D.surv_form[i];
    }
    __syncthreads();
  }
}

/* Between the local regroup+sort of a sharded pass and the replicated one: fresh look-back tags, zeroed
   form tables, histograms and tickets. */
__global__ void k_stage1_reset(const DevPtrs* D, const PassArgs* A, PassCtl* C, TableState* T) {
  const u32 nforms = (u32)A->nforms;
  for (u32 i = threadIdx.x; i < nforms; i += blockDim.x) { D->form_first[i] = 0; D->form_cnt[i] = 0; }
  u32* sc = (u32*)D->sortctl;
  for (u32 i = threadIdx.x; i < sizeof(SortCtl) / 4; i += blockDim.x) sc[i] = 0;
  if (threadIdx.x == 0) {
    C->chunk_ticket = 0; C->aux_ticket = 0;
    u32 e = (T->epoch + 16) & 0x3ffffff0u;
    if (e == 0) e = 16;
    T->epoch = e;
  }
}

/* Upper bound by a whole warp over a sorted u32 array: #{u : a[u] <= key}. */
__device__ __forceinline__ u32 warp_upper_bound_u32(const u32* __restrict__ a, u32 n, u32 key) {
  const u32 lane = threadIdx.x & 31;
  u32 lo = 0, hi = n;
  while (hi - lo > 32) {
    const u32 step = (hi - lo + 32) / 33;
    const u32 q = lo + (lane + 1) * step - 1;
    const bool le = (q < hi) && (a[q] <= key);
    const u32 c = __popc(__ballot_sync(0xffffffffu, le));
    const u32 nhi = lo + (c + 1) * step - 1;
    lo += c * step;
    if (c < 32 && nhi < hi) hi = nhi;
  }
  const bool le = (lo + lane < hi) && (a[lo + lane] <= key);
  return lo + __popc(__ballot_sync(0xffffffffu, le));
}

/* Old rows move up by the number of new rows whose gap is at or below them.  A chunk of consecutive
   old rows only needs the slice of the (sorted) gap list between the counts of its first and last row:
   found by two warp searches, staged in shared memory, searched there. */
#define MO_THREADS 256
#define MO_ITEMS 8
#define MO_CHUNK (MO_THREADS * MO_ITEMS)
#define MO_WIN 4096
__global__ void __launch_bounds__(MO_THREADS)
k_merge_old(const DevPtrs* Dp, const PassCtl* C, const TableState* T) {
  __shared__ u32 s_win[MO_WIN];
  __shared__ u32 s_lo, s_hi;
  if (C->overflow) return;
  const u32 nu = C->n_uniq;
  if (nu == 0) return; /* nothing inserted: the live buffer stays as it is */
  const u32 nt = T->nt;
  const DevPtrs D = *Dp;
  const Row* __restrict__ trow = D.trow[T->cur];
  Row* mrow = D.trow[T->cur ^ 1];
  double* mval = D.tval[T->cur ^ 1];
  const u32* __restrict__ uniq_gap = D.uniq_gap;
  const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  (void)lane;
  const u32 nchunks = (nt + MO_CHUNK - 1) / MO_CHUNK;
  for (u32 chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
    const u32 base = chunk * MO_CHUNK;
    const u32 last = min(base + MO_CHUNK, nt) - 1;
    if (warp == 0) {
      u32 r = warp_upper_bound_u32(uniq_gap, nu, base);
      if (lane == 0) s_lo = r;
    } else if (warp == 1) {
      u32 r = warp_upper_bound_u32(uniq_gap, nu, last);
      if (lane == 0) s_hi = r;
    }
    __syncthreads();
    const u32 lo = s_lo, wn = s_hi - s_lo;
    const bool in_smem = wn <= MO_WIN;
    if (in_smem)
      for (u32 j = tid; j < wn; j += MO_THREADS) s_win[j] = uniq_gap[lo + j];
    __syncthreads();
    const u32* win = in_smem ? s_win : uniq_gap + lo;
#pragma unroll
    for (int r = 0; r < MO_ITEMS; r++) {
      const u32 j = base + r * MO_THREADS + tid;
      if (j < nt) {
        u32 a = 0, b = wn;
        if (wn) {
          while (a < b) {
            u32 mid = (a + b) >> 1;
            if (win[mid] <= j) a = mid + 1; else b = mid;
          }
        }
        const Row row = trow[j];
        const u32 dst = j + lo + a;
        mrow[dst] = row;
        mval[dst] = row.val;
      }
    }
    __syncthreads();
  }
}

/* K4 */
__global__ void k_refine(const DevPtrs* Dp, const PassArgs* A, const PassCtl* C) {
  __shared__ RiesParams P;
  if (C->overflow) return;
  u32 n = C->n_events;
  if (n == 0) return;
  if (n > A->cap_events) n = A->cap_events;
  {
    const u32* src = (const u32*)Dp->P;
    u32* dst = (u32*)&P;
    for (u32 i = threadIdx.x; i < sizeof(RiesParams) / 4; i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  MatchEvent* events = Dp->events;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (events[i].kind == EV_EXACT) continue;
    MatchEvent ev = events[i];
    refine_event(P, ev);
    events[i] = ev;
  }
}

/* ---------------------------------------------------------------- host-side engine ------- */

template <typename T>
struct DevBuf {
  T* p = 0;
  size_t cap = 0;
  /* returns true when the pointer changed */
  bool ensure(size_t n, cudaStream_t st, bool keep = false, bool zero = false) {
    if (n <= cap) return false;
    size_t ncap = std::max(n, cap + cap / 2);
    T* q = 0;
    if (!(keep && p && cap) && p) { /* nothing to carry over: give the old block back first, so the pool can
                                       reuse its physical memory for the bigger one */
      CUDA_CHECK(cudaFreeAsync(p, st));
      p = 0;
    }
    CUDA_CHECK(cudaMallocAsync(&q, ncap * sizeof(T), st));
    if (zero) CUDA_CHECK(cudaMemsetAsync(q, 0, ncap * sizeof(T), st));
    if (keep && p && cap) CUDA_CHECK(cudaMemcpyAsync(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, st));
    if (p) CUDA_CHECK(cudaFreeAsync(p, st));
    p = q;
    cap = ncap;
    return true;
  }
  void release(cudaStream_t st) { if (p) cudaFreeAsync(p, st); p = 0; cap = 0; }
};


class DeviceEngine : public Engine {
 public:
  DeviceEngine() {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
      fprintf(stderr, "ries_b200: no CUDA device available (%s). This build has no CPU path.\n",
              cudaGetErrorString(e));
      throw ExitRequest{3};
    }
    const char* dev_env = getenv("RIES_DEVICE");
    int dev = dev_env ? atoi(dev_env) : 0;
    CUDA_CHECK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
    num_sms_ = prop.multiProcessorCount;
    {
      /* keep freed blocks in the stream-ordered pool instead of returning them to the OS */
      cudaMemPool_t pool;
      if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
      }
    }
    CUDA_CHECK(cudaStreamCreateWithFlags(&st_, cudaStreamNonBlocking));
    int occ = 1;
    CUDA_CHECK(cudaFuncSetAttribute(k_expand, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EXP_DYN_SMEM));
    CUDA_CHECK(cudaFuncSetAttribute(k_rs_onesweep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RS_DYN_SMEM));
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_expand, EXP_THREADS, EXP_DYN_SMEM));
    if (occ < 1) occ = 1;
    expand_grid_ = (u32)(num_sms_ * occ);
    wide_grid_ = (u32)(num_sms_ * 8);
    use_graphs_ = getenv("RIES_NO_GRAPHS") == 0;
    use_small_post_ = getenv("RIES_NO_SMALL_POST") == 0; /* test hook: every pass through the general kernels */
    force_small_post_ = getenv("RIES_FORCE_SMALL_POST") != 0;
    pass_log_ = getenv("RIES_PASS_LOG") != 0; /* one line per pass on stderr: sizes and device time per stage */
    CUDA_CHECK(cudaFuncSetAttribute(k_small_post, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SP_DYN_SMEM));
    CUDA_CHECK(cudaMalloc(&d_P_, sizeof(RiesParams)));
    CUDA_CHECK(cudaMalloc(&d_tt_, sizeof(TestTables)));
    CUDA_CHECK(cudaMalloc(&d_out_, sizeof(PassOut)));
    CUDA_CHECK(cudaMemset(d_out_, 0, sizeof(PassOut)));
    d_ctl_ = &d_out_->ctl;
    const size_t passin_bytes = PASSIN_FORMS_OFFSET + (size_t)PASSIN_MAX_FORMS * sizeof(FormDesc);
    CUDA_CHECK(cudaMalloc(&d_passin_, passin_bytes));
    d_args_ = (PassArgs*)d_passin_;
    d_forms_ = (FormDesc*)(d_passin_ + PASSIN_FORMS_OFFSET);
    CUDA_CHECK(cudaMalloc(&d_tab_, sizeof(TableState)));
    CUDA_CHECK(cudaMalloc(&d_ptrs_, sizeof(DevPtrs)));
    CUDA_CHECK(cudaMalloc(&d_sortctl_, sizeof(SortCtl)));
    CUDA_CHECK(cudaMalloc(&d_opcnt_, 2 * RIES_MAX_SYMS * sizeof(u64)));
    CUDA_CHECK(cudaMallocHost(&h_out_, sizeof(PassOut)));
    h_ctl_ = &h_out_->ctl;
    h_head_ = h_out_->head;
    CUDA_CHECK(cudaMallocHost(&h_passin_, passin_bytes));
    h_args_ = (PassArgs*)h_passin_;
    h_forms_ = (FormDesc*)(h_passin_ + PASSIN_FORMS_OFFSET);
    CUDA_CHECK(cudaMallocHost(&h_ptrs_, sizeof(DevPtrs)));
    CUDA_CHECK(cudaMallocHost(&h_allctl_, 64 * sizeof(PassCtl)));
    if (const char* e = getenv("RIES_SHARD_MIN")) shard_min_ = (u32)atol(e);
    memset(h_ptrs_, 0, sizeof(DevPtrs));
    for (int i = 0; i < 5; i++) CUDA_CHECK(cudaEventCreate(&ev_[i]));
    for (int i = 0; i < 2; i++) CUDA_CHECK(cudaEventCreate(&ev_g_[i]));
    const char* arena_env = getenv("RIES_ARENA_CAP"); /* test hook: start tiny to exercise overflow-and-rerun */
    const size_t arena0 = arena_env ? (size_t)atoi(arena_env) : (size_t)(1u << 20);
    const char* cap_env = getenv("RIES_EVENTS_CAP"); /* test hook: start tiny to exercise the regrow path */
    staged_.ensure((size_t)expand_grid_ * 2 * EXP_MAX_ITEMS, st_);
    events_.ensure(std::max<size_t>(cap_env ? (size_t)atoi(cap_env) : (size_t)(1u << 14), 1), st_);
    for (int i = 0; i < 3; i++) form_[i].ensure(4096, st_);
    grow_arena(arena0, arena0);
    ensure_table(1u << 16);
    ptrs_dirty_ = true;
  }
  ~DeviceEngine() override {
    cudaStreamSynchronize(st_);
    for (auto& kv : graphs_) cudaGraphExecDestroy(kv.second);
    nodes_.release(st_); done_.release(st_); done_form_.release(st_); surv_.release(st_);
    for (int b = 0; b < 2; b++) { keys_[b].release(st_); idx_[b].release(st_); trow_[b].release(st_); tval_[b].release(st_); }
    for (int i = 0; i < 3; i++) form_[i].release(st_);
    uniq_seq_.release(st_); uniq_gap_.release(st_); events_.release(st_); tiles_.release(st_); staged_.release(st_);
    rs_state_.release(st_); cc_state_.release(st_); gdone_.release(st_); gdone_form_.release(st_); allctl_.release(st_);
    surv_form_.release(st_); keep_.release(st_); ksend_.release(st_); ksend_form_.release(st_);
    cudaStreamSynchronize(st_);
    cudaFreeHost(h_allctl_);
    cudaFree(d_P_); cudaFree(d_tt_); cudaFree(d_out_); cudaFree(d_passin_); cudaFree(d_tab_); cudaFree(d_ptrs_); cudaFree(d_sortctl_);
    cudaFree(d_opcnt_);
    cudaFreeHost(h_out_); cudaFreeHost(h_passin_); cudaFreeHost(h_ptrs_);
    if (h_events_) cudaFreeHost(h_events_);
    for (int i = 0; i < 5; i++) cudaEventDestroy(ev_[i]);
    for (int i = 0; i < 2; i++) cudaEventDestroy(ev_g_[i]);
    cudaStreamDestroy(st_);
  }
  const char* name() const override { return "sm_100a"; }

  void begin(const RiesParams& P) override {
    if (!test_tables_cover(P)) {
      fprintf(stderr, "ries_b200: a symbol weight is outside the test tables' range\n");
      abort();
    }
    build_test_tables(P, h_tt_);
    CUDA_CHECK(cudaMemcpyAsync(d_P_, &P, sizeof(P), cudaMemcpyHostToDevice, st_));
    CUDA_CHECK(cudaMemcpyAsync(d_tt_, &h_tt_, sizeof(TestTables), cudaMemcpyHostToDevice, st_));
    CUDA_CHECK(cudaMemsetAsync(d_opcnt_, 0, 2 * RIES_MAX_SYMS * sizeof(u64), st_));
    TableState t0;
    t0.cur = 0; t0.nt = 0; t0.epoch = epoch_host_; t0.pad = 0;
    CUDA_CHECK(cudaMemcpyAsync(d_tab_, &t0, sizeof(t0), cudaMemcpyHostToDevice, st_));
    CUDA_CHECK(cudaStreamSynchronize(st_));
    memcpy(sym_chars_, P.sym_char, sizeof(sym_chars_));
    nt_ = 0;
    cur_ = 0;
    last_nu_ = 0;
    prev_nodes_[0] = prev_nodes_[1] = prev_done_[0] = prev_done_[1] = 0;
    kernel_launches = 0;
    h2d_bytes = sizeof(P);
    d2h_bytes = 0;
    for (int i = 0; i < 4; i++) { stage_ms[i] = 0; stage_launches[i] = 0; }
    sort_keys = merge_rows = expand_nodes = 0;
    sharded_passes = gather_bytes = 0;
    gather_ms = 0;
  }

  long inserted_before(uint32_t seq, bool inclusive) override {
    if (!last_nu_) return 0;
    std::vector<u32> h(last_nu_);
    CUDA_CHECK(cudaMemcpyAsync(h.data(), uniq_seq_.p, last_nu_ * sizeof(u32), cudaMemcpyDeviceToHost, st_));
    CUDA_CHECK(cudaStreamSynchronize(st_));
    long c = 0;
    for (u32 s : h) c += inclusive ? (s <= seq) : (s < seq);
    return c;
  }

  void run_pass(int using_x, int cplx, const std::vector<FormDesc>& forms, double best_match0,
                bool skip_exact, PassOutput& out) override {
    last_nu_ = 0;
    const u32 nforms = (u32)forms.size();
    if (nforms == 0) return;
    if (nforms > PASSIN_MAX_FORMS - 1) {
      fprintf(stderr, "ries_b200: %u forms in one pass exceeds the 16-bit form index\n", nforms);
      abort();
    }
    int max_len = 0;
    for (const FormDesc& f : forms) max_len = std::max(max_len, (int)f.flen);
    for (int i = 0; i < 3; i++)
      if (form_[i].ensure(nforms, st_)) ptrs_dirty_ = true;
    if (nodes_.cap < (size_t)nforms + 16) grow_arena(nforms + 16, done_.cap);
    memcpy(h_forms_, forms.data(), nforms * sizeof(FormDesc)); /* pinned; the last pass's copy has completed */
    h2d_bytes += nforms * sizeof(FormDesc) + sizeof(PassArgs);

    bool no_small_post = !use_small_post_;
    unsigned long long gen_k1 = 0; /* sharded passes: complete expressions of all ranks' K1, before the pre-dedupe */
    bool was_sharded = false;
    {
      /* sizes grow ~1.3-1.6x per complexity step: keep twice the previous same-side pass in hand, so that
         the overflow-and-rerun path stays the exception (it costs a whole pass) */
      const int side = using_x ? 1 : 0;
      const size_t want_nodes = (size_t)prev_nodes_[side] * 2, want_done = (size_t)prev_done_[side] * 2;
      if (!getenv("RIES_ARENA_CAP") && (want_nodes > nodes_.cap || want_done > done_.cap))
        grow_arena(std::max(want_nodes, nodes_.cap), std::max(want_done, done_.cap));
    }
    while (true) {
      /* the merged table needs room for every survivor the arena can hold */
      ensure_table((size_t)nt_ + done_.cap);
      if (ptrs_dirty_) upload_ptrs();
      h_args_->using_x = using_x; h_args_->cplx = cplx; h_args_->skip_exact = skip_exact ? 1 : 0;
      h_args_->nforms = (int)nforms; h_args_->best_match0 = best_match0;
      h_args_->cap_nodes = (u32)std::min<size_t>(nodes_.cap, 0xffffffffu);
      h_args_->cap_done = (u32)std::min<size_t>(done_.cap, 0xffffffffu);
      h_args_->cap_events = (u32)events_.cap; h_args_->pad = 0;
      /* size class from the previous pass of the same side (sizes grow ~1.3-1.6x per complexity step) */
      int small = (prev_nodes_[using_x ? 1 : 0] <= 12000 && prev_done_[using_x ? 1 : 0] <= 3000) ? 1 : 0;
      if (small && !no_small_post && prev_done_[using_x ? 1 : 0] <= SP_CAP / 3 && nt_ <= 200000u) small = 2;
      if (force_small_post_ && !no_small_post) small = 2; /* test hook: mispredict, to exercise the rerun */
      /* one search over several GPUs: only passes big enough to pay for the exchange are sharded */
      const DistState& ds = dist_state();
      const int shard_world = ds.fake_world > 1 ? ds.fake_world : ds.world;
      const bool shard = shard_world > 1 && prev_done_[using_x ? 1 : 0] >= shard_min_;
      h_args_->shard_rank = (shard && ds.fake_world <= 1) ? ds.rank : 0;
      h_args_->shard_world = shard ? shard_world : 1;
      h_args_->gathered = 0; /* a sharded pass switches to the gathered arrays after the exchange */
      h_args_->pad2 = 0;
      if (shard && small == 2) small = 1; /* the single-block kernel has no local/global split */
      CUDA_CHECK(cudaMemcpyAsync(d_passin_, h_passin_, PASSIN_FORMS_OFFSET + nforms * sizeof(FormDesc),
                                 cudaMemcpyHostToDevice, st_));
      if (!shard) {
        launch_stages(0, 3, max_len, small);
        CUDA_CHECK(cudaMemcpyAsync(h_out_, d_out_, sizeof(PassOut), cudaMemcpyDeviceToHost, st_));
        CUDA_CHECK(cudaStreamSynchronize(st_));
        CUDA_CHECK(cudaGetLastError());
        d2h_bytes += sizeof(PassOut);
        collect_stage_times(0, 3);
      } else {
        size_t want_nodes = nodes_.cap, want_done = done_.cap;
        const u32 fovf = sharded_front(max_len, small, &want_nodes, &want_done, &gen_k1);
        if (fovf) {
          grow_arena(std::max(want_nodes, nodes_.cap * ((fovf & OVF_NODES) ? 2 : 1)),
                     std::max(want_done, done_.cap * ((fovf & OVF_DONE) ? 2 : 1)));
          continue;
        }
        sharded_passes++;
        was_sharded = true;
        launch_stages(1, 3, max_len, small);
        CUDA_CHECK(cudaMemcpyAsync(h_out_, d_out_, sizeof(PassOut), cudaMemcpyDeviceToHost, st_));
        CUDA_CHECK(cudaStreamSynchronize(st_));
        CUDA_CHECK(cudaGetLastError());
        d2h_bytes += sizeof(PassOut);
        collect_stage_times(1, 3);
        if (gather_pending_) {
          float ms = 0;
          if (cudaEventElapsedTime(&ms, ev_g_[0], ev_g_[1]) == cudaSuccess) gather_ms += ms;
          (void)cudaGetLastError();
          gather_pending_ = false;
        }
      }
      const u32 ovf = h_ctl_->overflow;
      if (!ovf) break;
      /* grow what overflowed and rerun the pass; the live table buffer has not been touched */
      size_t want_nodes = nodes_.cap, want_done = done_.cap;
      if (ovf & OVF_NODES) want_nodes = std::max<size_t>(nodes_.cap * 2, (size_t)h_ctl_->node_count * 2);
      if (ovf & OVF_DONE) want_done = std::max<size_t>(done_.cap * 2, (size_t)h_ctl_->done_count * 2);
      if (ovf & (OVF_NODES | OVF_DONE)) grow_arena(want_nodes, want_done);
      if (ovf & OVF_EVENTS) {
        if (events_.ensure((size_t)h_ctl_->n_events + 1024, st_)) ptrs_dirty_ = true;
      }
      if (ovf & OVF_SMALL) no_small_post = true;
    }
    const u32 n = h_ctl_->done_count, nu = h_ctl_->n_uniq, ne = h_ctl_->n_events;
    if (pass_log_) {
      float ms[4] = {0, 0, 0, 0};
      for (int i = 0; i < 4; i++) (void)cudaEventElapsedTime(&ms[i], ev_[i], ev_[i + 1]);
      (void)cudaGetLastError();
      fprintf(stderr, "pass %s cplx %3d forms %5u len %2d nodes %9u done %9u new %8u table %9u events %5u | K1 %7.3f sort %7.3f merge %7.3f refine %6.3f ms\n",
              using_x ? "LHS" : "RHS", cplx, nforms, max_len, h_ctl_->node_count, n, nu, nt_, ne, ms[0], ms[1], ms[2], ms[3]);
    }
    const unsigned long long gen = was_sharded ? gen_k1 : (unsigned long long)n;
    prev_nodes_[using_x ? 1 : 0] = h_ctl_->node_count;
    prev_done_[using_x ? 1 : 0] = (u32)std::min<unsigned long long>(gen, 0xffffffffull);
    expand_nodes += h_ctl_->node_count;
    sort_keys += n;
    out.generated = (double)gen;
    out.dead_ends = (double)h_ctl_->dead;
    out.inserted = (long)nu;
    last_nu_ = nu;
    if (nu) {
      merge_rows += (unsigned long long)nt_ + nu;
      cur_ ^= 1;
      nt_ += nu;
    }
    if (ne) {
      if (ne <= HEAD_EVENTS) {
        out.events.assign(h_head_, h_head_ + ne);
      } else {
        if (ne > h_events_cap_) {
          if (h_events_) cudaFreeHost(h_events_);
          h_events_cap_ = std::max<size_t>(ne, h_events_cap_ * 2);
          CUDA_CHECK(cudaMallocHost(&h_events_, h_events_cap_ * sizeof(MatchEvent)));
        }
        CUDA_CHECK(cudaMemcpyAsync(h_events_, events_.p, ne * sizeof(MatchEvent), cudaMemcpyDeviceToHost, st_));
        CUDA_CHECK(cudaStreamSynchronize(st_));
        d2h_bytes += (unsigned long long)ne * sizeof(MatchEvent);
        out.events.assign(h_events_, h_events_ + ne);
      }
      std::sort(out.events.begin(), out.events.end(), [](const MatchEvent& a, const MatchEvent& b) {
        return a.seq != b.seq ? a.seq < b.seq : a.kind < b.kind;
      });
    }
  }

  unsigned long long kernel_launches = 0;
  unsigned long long h2d_bytes = 0, d2h_bytes = 0;
  /* device time per stage, CUDA events on the launching stream (ms): 0 K1 enumerate+evaluate,
     1 order+radix sort, 2 classify/compact/match/merge, 3 refine */
  double stage_ms[4] = {0, 0, 0, 0};
  unsigned long long stage_launches[4] = {0, 0, 0, 0};
  unsigned long long sort_keys = 0;      /* survivors that went through the radix sort */
  unsigned long long merge_rows = 0;     /* rows written by the merges (old + new) */
  unsigned long long expand_nodes = 0;   /* live prefixes expanded (= exec'd prefixes) */
  unsigned long long sharded_passes = 0; /* passes whose K1 ran sharded over the ranks */
  unsigned long long gather_bytes = 0;   /* bytes this rank received/kept in the all-gathers */
  double gather_ms = 0;                  /* device time of the all-gathers (CUDA events) */
  u32 table_rows() const { return nt_; }
  unsigned char sym_chars_[RIES_MAX_SYMS];
  void read_opcounts(unsigned long long* out /* [2][RIES_MAX_SYMS] */) {
    CUDA_CHECK(cudaMemcpyAsync(out, d_opcnt_, 2 * RIES_MAX_SYMS * sizeof(u64), cudaMemcpyDeviceToHost, st_));
    CUDA_CHECK(cudaStreamSynchronize(st_));
  }
  double dfma_peak_tflops() {
    const int blocks = num_sms_ * 8, threads = 512, iters = 1 << 14;
    double* d = 0;
    CUDA_CHECK(cudaMalloc(&d, (size_t)blocks * threads * sizeof(double)));
    cudaEvent_t a, b;
    CUDA_CHECK(cudaEventCreate(&a)); CUDA_CHECK(cudaEventCreate(&b));
    double best = 0;
    for (int rep = 0; rep < 5; rep++) {
      CUDA_CHECK(cudaEventRecord(a, st_));
      k_dfma_peak<<<blocks, threads, 0, st_>>>(d, iters);
      CUDA_CHECK(cudaEventRecord(b, st_));
      CUDA_CHECK(cudaEventSynchronize(b));
      float ms = 0;
      CUDA_CHECK(cudaEventElapsedTime(&ms, a, b));
      double tf = 2.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
      if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(d);
    return best;
  }

 private:
  /* The fixed launch sequence of one pass, in four stages.  Each stage is issued directly or as a CUDA
     graph captured once (stage 0 once per expression length); the events between stages are recorded
     outside the graphs so that per-stage device time can be read on this stream. */
  void enqueue_stage(int stage, int max_len, int small) {
    /* every kernel loops over its work (ticket / grid-stride), so the grid only sets the parallelism:
       the "small" variant is for passes predicted to be tiny, where launching a thousand blocks that
       exit at once costs more than the work itself.  A wrong prediction is slow, never wrong. */
    const u32 g_exp = small ? 24u : expand_grid_;
    const u32 g_wide = small ? 12u : wide_grid_;
    const u32 g_sort = small ? 12u : (u32)num_sms_ * 2;
    switch (stage) {
      case 0:
        k_pass_begin<<<1, 1024, 0, st_>>>(d_ptrs_, d_args_, d_ctl_, d_tab_);
        for (int lvl = 0; lvl < max_len; lvl++) k_expand<<<g_exp, EXP_THREADS, EXP_DYN_SMEM, st_>>>(d_ptrs_, d_args_, d_ctl_);
        break;
      case 1:
        if (small == 2) {
          k_small_post<<<1, SP_THREADS, SP_DYN_SMEM, st_>>>(d_ptrs_, d_args_, d_ctl_, d_tab_);
          break;
        }
        k_form_bounds<<<g_wide, 256, 0, st_>>>(d_ptrs_, d_args_, d_ctl_);
        k_form_scan<<<1, 1024, 0, st_>>>(d_ptrs_, d_args_, d_ctl_);
        k_reorder<<<g_wide, 256, 0, st_>>>(d_ptrs_, d_args_, d_ctl_);
        for (int pass = 0; pass < 8; pass++)
          k_rs_onesweep<<<g_sort, RS_THREADS, RS_DYN_SMEM, st_>>>(d_ptrs_, d_ctl_, d_tab_, pass);
        break;
      case 2:
        if (small == 2) {
          k_merge_old<<<48u, MO_THREADS, 0, st_>>>(d_ptrs_, d_ctl_, d_tab_);
          break;
        }
        k_classify_compact<<<g_wide, CC_THREADS, 0, st_>>>(d_ptrs_, d_args_, d_ctl_, d_tab_);
        k_match_merge_new<<<g_wide, 256, 0, st_>>>(d_ptrs_, d_args_, d_ctl_, d_tab_);
        k_merge_old<<<small ? 48u : wide_grid_, MO_THREADS, 0, st_>>>(d_ptrs_, d_ctl_, d_tab_);
        break;
      default:
        k_refine<<<small ? 8 : 64, 64, 0, st_>>>(d_ptrs_, d_args_, d_ctl_);
        k_pass_end<<<1, 128, 0, st_>>>(d_ptrs_, d_args_, d_ctl_, d_tab_);
        break;
    }
  }

  /* mirrors what k_pass_begin / k_stage1_reset do to TableState::epoch */
  void bump_epoch() {
    epoch_host_ = (epoch_host_ + 16) & 0x3ffffff0u;
    if (epoch_host_ == 0) {
      /* 2^26 bumps later the tags wrap: clear the look-back state so old tags cannot alias new ones */
      epoch_host_ = 16;
      CUDA_CHECK(cudaMemsetAsync(rs_state_.p, 0, rs_state_.cap * sizeof(u64), st_));
      CUDA_CHECK(cudaMemsetAsync(cc_state_.p, 0, cc_state_.cap * sizeof(u64), st_));
    }
  }

  /* Stages [first, last] of a pass on the stream; ev_[k] is recorded before stage k (and after the last). */
  void launch_stages(int first, int last, int max_len, int small) {
    if (first == 0) bump_epoch();
    CUDA_CHECK(cudaEventRecord(ev_[first], st_));
    for (int stage = first; stage <= last; stage++) {
      if (!use_graphs_) {
        enqueue_stage(stage, max_len, small);
      } else {
        const int key = ((stage == 0) ? max_len : -stage) * 4 + ((stage == 0 && small == 2) ? 1 : small) + 1000;
        auto it = graphs_.find(key);
        if (it == graphs_.end()) {
          cudaGraph_t g;
          CUDA_CHECK(cudaStreamBeginCapture(st_, cudaStreamCaptureModeThreadLocal));
          enqueue_stage(stage, max_len, small);
          CUDA_CHECK(cudaStreamEndCapture(st_, &g));
          cudaGraphExec_t ge;
          CUDA_CHECK(cudaGraphInstantiate(&ge, g, 0));
          CUDA_CHECK(cudaGraphDestroy(g));
          it = graphs_.emplace(key, ge).first;
        }
        CUDA_CHECK(cudaGraphLaunch(it->second, st_));
      }
      CUDA_CHECK(cudaEventRecord(ev_[stage + 1], st_));
      const unsigned long long k = (stage == 0) ? (unsigned long long)max_len + 1
                                 : (stage == 1) ? ((small == 2) ? 1 : 11)
                                 : (stage == 2) ? ((small == 2) ? 1 : 3) : 2;
      kernel_launches += k;
      stage_launches[stage] += k;
    }
  }

  void collect_stage_times(int first, int last) {
    for (int i = first; i <= last; i++) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, ev_[i], ev_[i + 1]) == cudaSuccess) stage_ms[i] += ms;
    }
    (void)cudaGetLastError();
  }

  /* Front half of a pass sharded over `world` ranks (or virtual ranks): K1 on this rank's forms, the local
     regroup + sort + pre-dedupe (k_local_filter), then the all-gather of what is left into
     gdone/gdone_form in rank-major order and the totals into the device control block.  Returns the
     overflow flags (any rank's); on overflow nothing was gathered and *want_* hold the sizes to grow to.
     Every rank computes the same totals, so every rank takes the same decisions afterwards. */
  u32 sharded_front(int max_len, int small, size_t* want_nodes, size_t* want_done, unsigned long long* gen_k1) {
    const DistState& ds = dist_state();
    const bool fake = ds.fake_world > 1;
    const int world = fake ? ds.fake_world : ds.world;
    std::vector<PassCtl> all((size_t)world);
    bool ch = false;
    ch |= keep_.ensure(done_.cap, st_);
    ch |= ksend_.ensure(done_.cap, st_);
    ch |= ksend_form_.ensure(done_.cap, st_);
    if (ch) { ptrs_dirty_ = true; upload_ptrs(); }
    const u32 g_wide = small ? 12u : wide_grid_;
    /* K1 -> regroup + sort of this rank's own expressions -> drop what cannot matter -> compact */
    auto local_front = [&]() {
      launch_stages(0, 1, max_len, small);
      k_local_filter<<<g_wide, CC_THREADS, 0, st_>>>(d_ptrs_, d_args_, d_ctl_, d_tab_);
      k_compact_kept<<<g_wide, 256, 0, st_>>>(d_ptrs_, d_ctl_, d_tab_);
      bump_epoch();
      k_stage1_reset<<<1, 1024, 0, st_>>>(d_ptrs_, d_args_, d_ctl_, d_tab_);
      kernel_launches += 3;
      stage_launches[1] += 3;
    };
    if (fake) {
      /* test hook: the virtual ranks run one after the other on this GPU; the "all-gather" is a copy */
      size_t off = 0;
      u32 ovf = 0;
      for (int r = 0; r < world; r++) {
        h_args_->shard_rank = r;
        CUDA_CHECK(cudaMemcpyAsync(d_passin_, h_passin_, PASSIN_FORMS_OFFSET, cudaMemcpyHostToDevice, st_));
        local_front();
        CUDA_CHECK(cudaMemcpyAsync(h_ctl_, d_ctl_, sizeof(PassCtl), cudaMemcpyDeviceToHost, st_));
        CUDA_CHECK(cudaStreamSynchronize(st_));
        collect_stage_times(0, 1);
        all[r] = *h_ctl_;
        ovf |= h_ctl_->overflow;
        if (h_ctl_->overflow) break;
        const size_t c = h_ctl_->n_kept;
        if (off + c > done_.cap) { ovf |= OVF_DONE; *want_done = std::max(*want_done, (off + c) * 2); break; }
        if (gdone_.ensure(off + c, st_, true)) ptrs_dirty_ = true;
        if (gdone_form_.ensure(off + c, st_, true)) ptrs_dirty_ = true;
        if (c) {
          CUDA_CHECK(cudaMemcpyAsync(gdone_.p + off, ksend_.p, c * sizeof(Row), cudaMemcpyDeviceToDevice, st_));
          CUDA_CHECK(cudaMemcpyAsync(gdone_form_.p + off, ksend_form_.p, c * sizeof(uint16_t), cudaMemcpyDeviceToDevice, st_));
        }
        off += c;
      }
      if (ovf) {
        for (int r = 0; r < world; r++) {
          if (all[r].overflow & OVF_NODES) *want_nodes = std::max(*want_nodes, (size_t)all[r].node_count * 2);
          if (all[r].overflow & OVF_DONE) *want_done = std::max(*want_done, (size_t)all[r].done_count * 2);
        }
        return ovf;
      }
    } else {
      local_front();
      allctl_.ensure((size_t)world, st_); /* not in the pointer table */
      dist_allgather(d_ctl_, allctl_.p, sizeof(PassCtl), st_);
      CUDA_CHECK(cudaMemcpyAsync(h_allctl_, allctl_.p, (size_t)world * sizeof(PassCtl), cudaMemcpyDeviceToHost, st_));
      CUDA_CHECK(cudaStreamSynchronize(st_));
      collect_stage_times(0, 1);
      d2h_bytes += (size_t)world * sizeof(PassCtl);
      u32 ovf = 0;
      size_t total = 0;
      for (int r = 0; r < world; r++) {
        all[r] = h_allctl_[r];
        ovf |= all[r].overflow;
        total += all[r].n_kept;
        if (all[r].overflow & OVF_NODES) *want_nodes = std::max(*want_nodes, (size_t)all[r].node_count * 2);
        if (all[r].overflow & OVF_DONE) *want_done = std::max(*want_done, (size_t)all[r].done_count * 2);
      }
      if (!ovf && total > done_.cap) { ovf |= OVF_DONE; *want_done = std::max(*want_done, total * 2); }
      if (ovf) return ovf;
      if (gdone_.ensure(total, st_)) ptrs_dirty_ = true;
      if (gdone_form_.ensure(total, st_)) ptrs_dirty_ = true;
      std::vector<size_t> bytes((size_t)world), offs((size_t)world);
      size_t off = 0;
      for (int r = 0; r < world; r++) { bytes[r] = (size_t)all[r].n_kept * sizeof(Row); offs[r] = off * sizeof(Row); off += all[r].n_kept; }
      CUDA_CHECK(cudaEventRecord(ev_g_[0], st_));
      dist_allgatherv(ksend_.p, gdone_.p, bytes.data(), offs.data(), st_);
      off = 0;
      for (int r = 0; r < world; r++) { bytes[r] = (size_t)all[r].n_kept * sizeof(uint16_t); offs[r] = off * sizeof(uint16_t); off += all[r].n_kept; }
      dist_allgatherv(ksend_form_.p, gdone_form_.p, bytes.data(), offs.data(), st_);
      CUDA_CHECK(cudaEventRecord(ev_g_[1], st_));
      gather_pending_ = true;
      gather_bytes += total * (sizeof(Row) + sizeof(uint16_t));
    }
    /* totals into the device control block, as if one K1 had produced exactly the kept expressions; from
       here on the regroup kernels read the gathered arrays */
    PassCtl t = all[0];
    unsigned long long nodes = 0, done = 0, kept = 0, dead = 0;
    for (int r = 0; r < world; r++) { nodes += all[r].node_count; done += all[r].done_count; kept += all[r].n_kept; dead += all[r].dead; }
    t.node_count = (u32)std::min<unsigned long long>(nodes, 0xffffffffull);
    t.done_count = (u32)kept;
    t.dead = dead;
    t.lvl_begin = t.lvl_end = 0; t.tile_ticket = 0; t.blocks_done = 0; t.chunk_ticket = 0; t.aux_ticket = 0;
    *gen_k1 = done;
    *h_ctl_ = t;
    CUDA_CHECK(cudaMemcpyAsync(d_ctl_, h_ctl_, sizeof(PassCtl), cudaMemcpyHostToDevice, st_));
    h_args_->gathered = 1;
    CUDA_CHECK(cudaMemcpyAsync(d_passin_, h_passin_, PASSIN_FORMS_OFFSET, cudaMemcpyHostToDevice, st_));
    if (ptrs_dirty_) upload_ptrs();
    return 0;
  }

  /* nodes / done arenas and everything sized by them */
  void grow_arena(size_t want_nodes, size_t want_done) {
    bool ch = false;
    ch |= nodes_.ensure(want_nodes, st_);
    ch |= done_.ensure(want_done, st_);
    ch |= done_form_.ensure(done_.cap, st_);
    ch |= surv_.ensure(done_.cap, st_);
    ch |= surv_form_.ensure(done_.cap, st_);
    for (int b = 0; b < 2; b++) { ch |= keys_[b].ensure(done_.cap, st_); ch |= idx_[b].ensure(done_.cap, st_); }
    ch |= uniq_seq_.ensure(done_.cap, st_);
    ch |= uniq_gap_.ensure(done_.cap, st_);
    /* look-back state must start as zeros (stale bits must not look like a live epoch / status) */
    ch |= tiles_.ensure(nodes_.cap / EXP_PARENTS + 2, st_, false, true);
    ch |= rs_state_.ensure((done_.cap / RS_CHUNK + 2) * 256, st_, false, true);
    ch |= cc_state_.ensure(done_.cap / 256 + 2, st_, false, true); /* one word per chunk; the smallest chunk is 256 (k_compact_kept) */
    if (ch) ptrs_dirty_ = true;
  }
  void ensure_table(size_t rows) {
    /* the live buffer keeps its contents; the other one is the merge destination */
    if (trow_[cur_].ensure(rows, st_, true)) ptrs_dirty_ = true;
    if (tval_[cur_].ensure(rows, st_, true)) ptrs_dirty_ = true;
    if (trow_[cur_ ^ 1].ensure(rows, st_)) ptrs_dirty_ = true;
    if (tval_[cur_ ^ 1].ensure(rows, st_)) ptrs_dirty_ = true;
  }
  void upload_ptrs() {
    CUDA_CHECK(cudaStreamSynchronize(st_)); /* h_ptrs_ may still be in flight from the last upload */
    DevPtrs& h = *h_ptrs_;
    h.P = d_P_; h.tt = d_tt_; h.forms = d_forms_; h.nodes = nodes_.p; h.done = done_.p; h.done_form = done_form_.p; h.gdone = gdone_.p; h.gdone_form = gdone_form_.p; h.surv = surv_.p;
    h.surv_form = surv_form_.p; h.keep = keep_.p; h.ksend = ksend_.p; h.ksend_form = ksend_form_.p;
    for (int b = 0; b < 2; b++) { h.keys[b] = keys_[b].p; h.idx[b] = idx_[b].p; h.trow[b] = trow_[b].p; h.tval[b] = tval_[b].p; }
    h.form_first = form_[0].p; h.form_cnt = form_[1].p; h.form_base = form_[2].p;
    h.uniq_seq = uniq_seq_.p; h.uniq_gap = uniq_gap_.p; h.events = events_.p; h.tiles = tiles_.p; h.staged = staged_.p;
    h.rs_state = rs_state_.p; h.cc_state = cc_state_.p; h.sortctl = d_sortctl_; h.opcnt = d_opcnt_;
    CUDA_CHECK(cudaMemcpyAsync(d_ptrs_, h_ptrs_, sizeof(DevPtrs), cudaMemcpyHostToDevice, st_));
    ptrs_dirty_ = false;
  }

  cudaStream_t st_;
  int num_sms_ = 0;
  u32 expand_grid_ = 0, wide_grid_ = 0;
  bool use_graphs_ = true, use_small_post_ = true, force_small_post_ = false, pass_log_ = false;
  std::map<int, cudaGraphExec_t> graphs_;
  RiesParams* d_P_ = 0;
  TestTables* d_tt_ = 0;
  TestTables h_tt_;
  PassOut* d_out_ = 0;
  PassOut* h_out_ = 0;
  PassCtl* d_ctl_ = 0; /* = &d_out_->ctl */
  PassCtl* h_ctl_ = 0;
  unsigned char* d_passin_ = 0; /* PassArgs, then the forms */
  unsigned char* h_passin_ = 0;
  PassArgs* d_args_ = 0;
  PassArgs* h_args_ = 0;
  FormDesc* d_forms_ = 0;
  TableState* d_tab_ = 0;
  DevPtrs* d_ptrs_ = 0;
  DevPtrs* h_ptrs_ = 0;
  bool ptrs_dirty_ = true;
  SortCtl* d_sortctl_ = 0;
  u64* d_opcnt_ = 0;
  cudaEvent_t ev_[5];
  cudaEvent_t ev_g_[2];
  bool gather_pending_ = false;
  FormDesc* h_forms_ = 0;
  DevBuf<Node> nodes_;
  DevBuf<Row> done_, surv_, gdone_, ksend_;
  DevBuf<uint16_t> done_form_, gdone_form_, surv_form_, ksend_form_;
  DevBuf<unsigned char> keep_;
  DevBuf<PassCtl> allctl_;
  PassCtl* h_allctl_ = 0; /* pinned, MAX_RANKS entries */
  u32 shard_min_ = 100000; /* complete expressions in the previous same-side pass from which a pass is sharded */
  DevBuf<u64> tiles_, rs_state_, cc_state_;
  DevBuf<Staged> staged_;
  DevBuf<u32> form_[3];
  DevBuf<u64> keys_[2];
  DevBuf<u32> idx_[2];
  DevBuf<u32> uniq_seq_, uniq_gap_;
  DevBuf<Row> trow_[2];
  DevBuf<double> tval_[2];
  DevBuf<MatchEvent> events_;
  MatchEvent* h_head_ = 0;
  MatchEvent* h_events_ = 0;
  size_t h_events_cap_ = 0;
  u32 epoch_host_ = 16; /* mirrors TableState::epoch: tags stay unique across searches of this engine */
  u32 nt_ = 0;
  int cur_ = 0;
  u32 last_nu_ = 0;
  u32 prev_nodes_[2] = {0, 0}, prev_done_[2] = {0, 0};
};

Engine* make_device_engine() { return new DeviceEngine(); }

void device_engine_report(Engine* e, DeviceReport* r) {
  DeviceEngine* d = static_cast<DeviceEngine*>(e);
  r->kernel_launches = d->kernel_launches;
  r->h2d_bytes = d->h2d_bytes;
  r->d2h_bytes = d->d2h_bytes;
  for (int i = 0; i < 4; i++) { r->stage_ms[i] = d->stage_ms[i]; r->stage_launches[i] = d->stage_launches[i]; }
  r->sort_keys = d->sort_keys;
  r->merge_rows = d->merge_rows;
  r->expand_nodes = d->expand_nodes;
  r->sharded_passes = d->sharded_passes;
  r->gather_bytes = d->gather_bytes;
  r->gather_ms = d->gather_ms;
  r->table_rows = d->table_rows();
  d->read_opcounts(&r->op_count[0][0]);
  memcpy(r->op_char, d->sym_chars_, sizeof(r->op_char));
}

double device_engine_dfma_peak(Engine* e) { return static_cast<DeviceEngine*>(e)->dfma_peak_tflops(); }

}  // namespace ries
