// The sm_100a device engine: every hot stage of the RIES search as hand-written CUDA kernels.
//
//   K1  k_expand        enumerate + evaluate, one thread per (live prefix, candidate symbol)
//                       [ge_2 + exec, ries.c:8908-9475, 5308-5950]
//   K2  k_rs_*          stable LSD radix sort of the pass's survivors on the fp64 value
//       k_classify      duplicate pruning inside the pass and against the resident table
//       k_merge_*       merge into the resident sorted table            [bt_insert, ries.c:8622-8724]
//   K3  k_match         gap-local neighbour test with the derivative-scaled window
//                                                           [check_sides/check_match, 8488, 8243]
//   K4  k_refine        cv_simplify + Newton on the compacted candidates       [7297, 7139]
//
// Data layout in HBM (all device-resident for the whole search):
//   nodes[]   48 B AoS, one per live prefix of the running pass (cactus stack, see params.h)
//   table_*   the value table: rows 32 B AoS sorted by value + a dense 8 B key array for searches
//   surv[]    32 B rows, the pass's complete expressions in insertion (DFS) order
// No tensor cores: nothing here is a dense contraction.  K1 is fp64-pipe bound, K2/K3 HBM bound.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "core/expand.h"
#include "core/refine.h"
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

struct PassCtl {          /* lives in device memory, mirrored to pinned host memory per pass */
  u32 lvl_begin, lvl_end; /* frontier = nodes[lvl_begin, lvl_end) */
  u32 node_count;         /* nodes appended so far (may exceed capacity: then overflow is set) */
  u32 done_count;         /* complete expressions so far */
  u32 tile_ticket;
  u32 n_events;
  u32 n_uniq;
  u32 overflow;           /* bit0 nodes, bit1 done, bit2 events */
  u64 dead;               /* dead-ends of the pass */
  u64 pad;
};

/* ---------------------------------------------------------------- K1: enumerate+evaluate - */

/* K1 works on tiles of 256 consecutive live prefixes (one block per tile at a time):
     phase 1  thread = prefix: the integer symbol-append tests for all candidate symbols -> bit mask
     phase 2  the (prefix, symbol) pairs that passed are listed in shared memory GROUPED BY SYMBOL SLOT
     phase 3  thread = listed pair: exec + value prunes; warps are (almost) operator-uniform
     phase 4  ordered append: children land in (prefix, symbol) order = the reference's DFS order */
#define EXP_THREADS 256
#define EXP_PARENTS 256
#define EXP_MAX_ITEMS (EXP_PARENTS * RIES_SLOTS)
#define IT_LIVE 0x1000u
#define IT_DONE 0x2000u

struct __attribute__((aligned(8))) Staged { /* a child between phase 3 and phase 4 (L2-resident scratch) */
  double val, der;
  u32 below;
  unsigned char tag, sp, pad0, pad1;
};

#define TS_AGG (1ull << 62)
#define TS_INC (2ull << 62)
#define TS_STATUS(x) ((x) >> 62)
#define TS_LIVE(x) ((u32)(((x) >> 31) & 0x7fffffffull))
#define TS_DONE(x) ((u32)((x)&0x7fffffffull))
#define TS_PACK(st, live, done) ((st) | ((u64)(live) << 31) | (u64)(done))

__global__ void k_init_roots(Node* nodes, u32 nforms, PassCtl* ctl) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nforms) {
    Node r;
    r.val = 0; r.der = 0; r.sym.lo = 0; r.sym.hi = 0;
    r.below = RIES_NO_NODE; r.form = (uint16_t)i; r.cplx = 0; r.tag = 0; r.len = 0; r.sp = 0;
    r.pad0 = 0; r.pad1 = 0;
    nodes[i] = r;
  }
  if (i == 0) {
    ctl->lvl_begin = 0; ctl->lvl_end = nforms; ctl->node_count = nforms; ctl->done_count = 0;
    ctl->tile_ticket = 0; ctl->n_events = 0; ctl->n_uniq = 0; ctl->overflow = 0; ctl->dead = 0;
  }
}

/* Decoupled look-back over tile aggregates (two counters packed in one word). Called by warp 0. */
__device__ __forceinline__ void tile_lookback(volatile u64* state, u32 tile, u32 live_agg, u32 done_agg,
                                              u32& live_excl, u32& done_excl) {
  const u32 lane = threadIdx.x & 31;
  if (tile == 0) {
    if (lane == 0) {
      state[0] = TS_PACK(TS_INC, live_agg, done_agg);
    }
    live_excl = 0; done_excl = 0;
    return;
  }
  if (lane == 0) {
    state[tile] = TS_PACK(TS_AGG, live_agg, done_agg);
  }
  __threadfence();
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

#ifndef EXP_MIN_BLOCKS
#define EXP_MIN_BLOCKS 3
#endif
__global__ void __launch_bounds__(EXP_THREADS, EXP_MIN_BLOCKS)
k_expand(const RiesParams* __restrict__ gP, const FormDesc* __restrict__ forms, Node* nodes,
         Row* done, uint16_t* done_form, u32 cap_nodes, u32 cap_done, PassCtl* ctl,
         volatile u64* tile_state, Staged* staged_all, u64* opcnt, int using_x, int cplx) {
  __shared__ u32 s_opcnt[RIES_MAX_SYMS];
  __shared__ RiesParams P;
  __shared__ Node s_par[EXP_PARENTS];
  __shared__ unsigned short s_item[EXP_MAX_ITEMS];
  __shared__ u32 s_live[EXP_PARENTS], s_done[EXP_PARENTS];   /* per prefix: bit per slot */
  __shared__ u32 s_lbase[EXP_PARENTS], s_dbase[EXP_PARENTS]; /* per prefix: exclusive child offsets */
  __shared__ unsigned char s_cls[EXP_PARENTS], s_complete[EXP_PARENTS];
  __shared__ u32 s_wcnt[EXP_THREADS / 32][RIES_SLOTS];
  __shared__ u32 s_slot_base[RIES_SLOTS];
  __shared__ u32 s_wl[EXP_THREADS / 32], s_wd[EXP_THREADS / 32];
  __shared__ u32 s_total, s_live_base, s_done_base, s_dead;
  {
    const u32* src = (const u32*)gP;
    u32* dst = (u32*)&P;
    for (u32 i = threadIdx.x; i < sizeof(RiesParams) / 4; i += EXP_THREADS) dst[i] = src[i];
    if (threadIdx.x == 0) s_dead = 0;
    if (threadIdx.x < RIES_MAX_SYMS) s_opcnt[threadIdx.x] = 0;
  }
  __syncthreads();
  const u32 lvl_begin = ctl->lvl_begin, lvl_end = ctl->lvl_end;
  const u32 done_base0 = ctl->done_count; /* read-only during this launch (updated by k_level_finish) */
  const u32 n = lvl_end - lvl_begin;
  const u32 ntiles = (n + EXP_PARENTS - 1) / EXP_PARENTS;
  const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u32 lt = (1u << lane) - 1;
  Staged* staged = staged_all + (size_t)blockIdx.x * EXP_MAX_ITEMS;
  u32 my_dead = 0;

  /* Tiles are handed out by ticket, so a tile's predecessors (which its look-back waits on) are always
     held by blocks that are already running: safe even when not every block of the grid is resident
     (several engines sharing the GPU on different streams). */
  __shared__ u32 s_tile;
  while (true) {
    if (tid == 0) s_tile = atomicAdd(&ctl->tile_ticket, 1u);
    __syncthreads();
    const u32 tile = s_tile;
    if (tile >= ntiles) break;
    const u32 p0 = lvl_begin + tile * EXP_PARENTS;
    /* ---- phase 1: integer tests, one thread per prefix ---- */
    u32 mask = 0;
    {
      const u32 p = p0 + tid;
      if (p < lvl_end) {
        const Node par = nodes[p];
        s_par[tid] = par;
        const FormDesc& F = forms[par.form];
        const int cls = F.cls[par.len];
        const int ns = P.cls_n[cls];
        const int atts = prefix_attrs(P, F, par, using_x);
        for (int k = 0; k < ns; k++) {
          if (expand_tests(P, F, par, k, atts, using_x, cplx, cplx)) mask |= 1u << k;
          else my_dead++;
        }
        s_cls[tid] = (unsigned char)cls;
        s_complete[tid] = (unsigned char)(par.len + 1 >= F.flen);
      }
      s_live[tid] = 0;
      s_done[tid] = 0;
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
      s_total = run;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < RIES_SLOTS; k++) {
      u32 b = __ballot_sync(0xffffffffu, (mask >> k) & 1u);
      if ((mask >> k) & 1u) s_item[s_slot_base[k] + s_wcnt[warp][k] + __popc(b & lt)] = (unsigned short)((tid << 4) | k);
    }
    __syncthreads();
    const u32 T = s_total;
    /* ---- phase 3: exec + value prunes, one thread per listed pair ---- */
    for (u32 i = tid; i < T; i += EXP_THREADS) {
      const u32 it = s_item[i];
      const u32 pl = it >> 4, k = it & 15u;
      const int cls = s_cls[pl];
      const int g = P.trial[cls][k];
      Node child;
      atomicAdd(&s_opcnt[g], 1u); /* exec() calls per symbol: the algorithmic work of this kernel */
      int oc = expand_exec(P, cls, g, s_complete[pl] != 0, nodes, s_par[pl], p0 + pl, using_x, child);
      if (oc == OUT_DEAD) my_dead++;
      else if (oc == OUT_LIVE || oc == OUT_DONE) {
        Staged st;
        st.val = child.val; st.der = child.der; st.below = child.below; st.tag = child.tag; st.sp = child.sp;
        st.pad0 = 0; st.pad1 = 0;
        staged[i] = st;
        if (oc == OUT_LIVE) { atomicOr(&s_live[pl], 1u << k); s_item[i] = (unsigned short)(it | IT_LIVE); }
        else { atomicOr(&s_done[pl], 1u << k); s_item[i] = (unsigned short)(it | IT_DONE); }
      }
    }
    __syncthreads();
    /* ---- phase 4: ordered append ---- */
    {
      u32 nl = __popc(s_live[tid]), nd = __popc(s_done[tid]);
      u32 il = nl, id = nd;
      for (int o = 1; o < 32; o <<= 1) {
        u32 a = __shfl_up_sync(0xffffffffu, il, o), b = __shfl_up_sync(0xffffffffu, id, o);
        if (lane >= (u32)o) { il += a; id += b; }
      }
      if (lane == 31) { s_wl[warp] = il; s_wd[warp] = id; }
      __syncthreads();
      u32 wl = 0, wd = 0;
      for (u32 w = 0; w < warp; w++) { wl += s_wl[w]; wd += s_wd[w]; }
      s_lbase[tid] = wl + il - nl;
      s_dbase[tid] = wd + id - nd;
      if (warp == 0) {
        u32 la = 0, da = 0;
        for (int w = 0; w < EXP_THREADS / 32; w++) { la += s_wl[w]; da += s_wd[w]; }
        u32 le, de;
        tile_lookback(tile_state, tile, la, da, le, de);
        if (lane == 0) { s_live_base = le; s_done_base = de; }
      }
      __syncthreads();
    }
    for (u32 i = tid; i < T; i += EXP_THREADS) {
      const u32 it = s_item[i];
      if (!(it & (IT_LIVE | IT_DONE))) continue;
      const u32 pl = (it >> 4) & 255u, k = it & 15u;
      const Node& par = s_par[pl];
      const int g = P.trial[s_cls[pl]][k];
      const Staged st = staged[i];
      Node child;
      child.val = st.val; child.der = st.der; child.below = st.below; child.tag = st.tag; child.sp = st.sp;
      expand_finish(P, par, g, child);
      const u32 below_k = (1u << k) - 1;
      if (it & IT_LIVE) {
        u32 pos = lvl_end + s_live_base + s_lbase[pl] + __popc(s_live[pl] & below_k);
        if (pos < cap_nodes) nodes[pos] = child;
      } else {
        u32 pos = done_base0 + s_done_base + s_dbase[pl] + __popc(s_done[pl] & below_k);
        if (pos < cap_done) {
          Row r; r.val = child.val; r.der = child.der; r.sym = child.sym;
          done[pos] = r;
          done_form[pos] = child.form;
        }
      }
    }
    __syncthreads(); /* shared tile state is reused by the next tile */
  }
  /* block-level dead-end count -> one atomic per block */
  for (int o = 16; o > 0; o >>= 1) my_dead += __shfl_xor_sync(0xffffffffu, my_dead, o);
  if (lane == 0 && my_dead) atomicAdd(&s_dead, my_dead);
  __syncthreads();
  if (tid == 0 && s_dead) atomicAdd(&ctl->dead, (u64)s_dead);
  if (tid < RIES_MAX_SYMS && s_opcnt[tid]) atomicAdd(&opcnt[(using_x ? RIES_MAX_SYMS : 0) + tid], (u64)s_opcnt[tid]);
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

/* After a level: advance the frontier, fold the level's totals, clear the tile states it used. */
__global__ void k_level_finish(PassCtl* ctl, u64* tile_state, u32 cap_nodes, u32 cap_done) {
  __shared__ u32 s_ntiles;
  __shared__ u64 s_last;
  if (threadIdx.x == 0) {
    u32 n = ctl->lvl_end - ctl->lvl_begin;
    u32 ntiles = (n + EXP_PARENTS - 1) / EXP_PARENTS;
    s_ntiles = ntiles;
    s_last = ntiles ? tile_state[ntiles - 1] : 0;
  }
  __syncthreads();
  const u32 ntiles = s_ntiles;
  for (u32 i = threadIdx.x; i < ntiles; i += blockDim.x) tile_state[i] = 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    u32 live = TS_LIVE(s_last), dn = TS_DONE(s_last);
    u32 old_end = ctl->lvl_end;
    u64 new_nodes = (u64)old_end + live;
    u64 new_done = (u64)ctl->done_count + dn;
    u32 ovf = ctl->overflow;
    if (new_nodes > cap_nodes) { ovf |= 1; }
    if (new_done > cap_done) { ovf |= 2; }
    ctl->overflow = ovf;
    ctl->tile_ticket = 0;
    ctl->lvl_begin = old_end;
    /* on overflow stop expanding: the host grows the arenas and reruns the pass */
    ctl->lvl_end = (ovf & 3) ? old_end : (u32)new_nodes;
    ctl->node_count = (new_nodes > 0xffffffffull) ? 0xffffffffu : (u32)new_nodes;
    ctl->done_count = (new_done > 0xffffffffull) ? 0xffffffffu : (u32)new_done;
  }
}

/* ------------------------------------------------ survivors into insertion (DFS) order --- */
/* Within one level the completed expressions are ordered (form, lexicographic) and every form
   completes at exactly one level, so a form's expressions are contiguous in done[]. */
__global__ void k_form_bounds(const uint16_t* done_form, u32 n, u32* form_first, u32* form_cnt) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint16_t f = done_form[i];
  if (i == 0 || done_form[i - 1] != f) form_first[f] = i;
  if (i + 1 == n || done_form[i + 1] != f) form_cnt[f] = i + 1; /* end; fixed up in k_form_scan */
}

__global__ void k_form_scan(u32 nforms, const u32* form_first, u32* form_cnt, u32* form_base) {
  /* single block: nforms is at most a few thousand */
  __shared__ u32 s_carry;
  __shared__ u32 s_tmp[1024];
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (u32 base = 0; base < nforms; base += 1024) {
    u32 f = base + threadIdx.x;
    u32 c = 0;
    if (f < nforms) {
      u32 e = form_cnt[f];
      c = e ? (e - form_first[f]) : 0;
      form_cnt[f] = c;
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

__global__ void k_reorder(const Row* done, const uint16_t* done_form, u32 n, const u32* form_first,
                          const u32* form_base, Row* surv, u64* keys, u32* idx) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint16_t f = done_form[i];
  u32 pos = form_base[f] + (i - form_first[f]);
  Row r = done[i];
  surv[pos] = r;
  keys[pos] = key_of(r.val);
  idx[pos] = pos;
}

/* ---------------------------------------------------------------- generic exclusive scan - */
#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_CHUNK (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ u32 block_excl_scan(u32 v, u32* s_warp, u32& total) {
  const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  u32 inc = v;
  for (int o = 1; o < 32; o <<= 1) {
    u32 t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= (u32)o) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  u32 woff = 0, tot = 0;
  for (u32 w = 0; w < SCAN_THREADS / 32; w++) {
    u32 t = s_warp[w];
    if (w < warp) woff += t;
    tot += t;
  }
  __syncthreads();
  total = tot;
  return woff + inc - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const u32* in, u32 n, u32* sums) {
  __shared__ u32 s_warp[SCAN_THREADS / 32];
  u32 base = blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_ITEMS;
  u32 v = 0;
  for (int k = 0; k < SCAN_ITEMS; k++)
    if (base + k < n) v += in[base + k];
  u32 total;
  block_excl_scan(v, s_warp, total);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_apply(const u32* in, u32* out, u32 n, const u32* sums_excl, u32* total_out) {
  __shared__ u32 s_warp[SCAN_THREADS / 32];
  u32 base = blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_ITEMS;
  u32 x[SCAN_ITEMS];
  u32 v = 0;
  for (int k = 0; k < SCAN_ITEMS; k++) {
    x[k] = (base + k < n) ? in[base + k] : 0;
    v += x[k];
  }
  u32 total;
  u32 off = block_excl_scan(v, s_warp, total) + (sums_excl ? sums_excl[blockIdx.x] : 0);
  for (int k = 0; k < SCAN_ITEMS; k++) {
    if (base + k < n) out[base + k] = off;
    off += x[k];
  }
  if (total_out && blockIdx.x == gridDim.x - 1 && threadIdx.x == SCAN_THREADS - 1) *total_out = off;
}

/* ---------------------------------------------------------------- K2: radix sort --------- */
/* Stable LSD radix sort of (u64 key, u32 payload), 8 digits of 8 bits, "onesweep" style: ONE pass over
   the keys builds all eight digit histograms; each digit pass is then a single kernel in which every
   2048-key chunk gets its per-digit output offsets by decoupled look-back over the chunks before it
   (chunks are taken in ticket order, so equal keys keep their input order = insertion order). */
#define RS_THREADS 256
#define RS_ITEMS 8
#define RS_CHUNK (RS_THREADS * RS_ITEMS)
#define RS_FLAG_AGG 1ull
#define RS_FLAG_INC 2ull

struct SortCtl {
  u32 ghist[8][256];
  u32 ticket[8];
};

__global__ void __launch_bounds__(RS_THREADS)
k_rs_hist_all(const u64* __restrict__ keys, u32 n, SortCtl* ctl) {
  __shared__ u32 s_h[8][256];
  for (int d = 0; d < 8; d++) s_h[d][threadIdx.x] = 0;
  __syncthreads();
  for (u32 base = blockIdx.x * RS_CHUNK; base < n; base += gridDim.x * RS_CHUNK) {
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
      u32 i = base + r * RS_THREADS + threadIdx.x;
      if (i < n) {
        u64 k = keys[i];
#pragma unroll
        for (int d = 0; d < 8; d++) atomicAdd(&s_h[d][(u32)(k >> (8 * d)) & 255u], 1u);
      }
    }
  }
  __syncthreads();
  for (int d = 0; d < 8; d++) {
    u32 c = s_h[d][threadIdx.x];
    if (c) atomicAdd(&ctl->ghist[d][threadIdx.x], c);
  }
}

__global__ void __launch_bounds__(RS_THREADS)
k_rs_onesweep(const u64* __restrict__ keys_in, const u32* __restrict__ idx_in, u64* keys_out, u32* idx_out,
              u32 n, int pass, SortCtl* ctl, volatile u64* state, u32 epoch) {
  __shared__ u32 s_off[256];
  __shared__ u32 s_scan[256];
  __shared__ u32 s_wcnt[RS_THREADS / 32][256];
  __shared__ u32 s_chunk;
  const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int shift = pass * 8;
  if (tid == 0) s_chunk = atomicAdd(&ctl->ticket[pass], 1u);
  /* exclusive scan of the digit's global histogram (256 bins, Hillis-Steele in shared memory) */
  u32 mine = ctl->ghist[pass][tid];
  s_scan[tid] = mine;
  for (int w = 0; w < RS_THREADS / 32; w++) s_wcnt[w][tid] = 0;
  s_off[tid] = 0;
  __syncthreads();
  for (u32 o = 1; o < 256; o <<= 1) {
    u32 v = (tid >= o) ? s_scan[tid - o] : 0;
    __syncthreads();
    s_scan[tid] += v;
    __syncthreads();
  }
  const u32 gbase = s_scan[tid] - mine;
  const u32 chunk = s_chunk;
  const u32 base = chunk * RS_CHUNK;

  /* local histogram of this chunk */
  u64 key[RS_ITEMS];
  u32 pay[RS_ITEMS];
#pragma unroll
  for (int r = 0; r < RS_ITEMS; r++) {
    u32 i = base + r * RS_THREADS + tid;
    bool valid = i < n;
    key[r] = valid ? keys_in[i] : 0;
    pay[r] = valid ? idx_in[i] : 0;
    if (valid) atomicAdd(&s_off[(u32)(key[r] >> shift) & 255u], 1u);
  }
  __syncthreads();
  const u32 local = s_off[tid];
  /* decoupled look-back for bin `tid` over the chunks before this one */
  {
    volatile u64* st = state + (size_t)chunk * 256 + tid;
    const u64 tag = (u64)epoch << 34;
    u32 ex = 0;
    if (chunk == 0) {
      *st = tag | (RS_FLAG_INC << 32) | local;
    } else {
      *st = tag | (RS_FLAG_AGG << 32) | local;
      int c = (int)chunk - 1;
      while (c >= 0) {
        u64 w;
        do { w = state[(size_t)c * 256 + tid]; } while ((w >> 34) != epoch || ((w >> 32) & 3ull) == 0);
        ex += (u32)w;
        if (((w >> 32) & 3ull) == RS_FLAG_INC) break;
        c--;
      }
      *st = tag | (RS_FLAG_INC << 32) | (u64)(ex + local);
    }
    __syncthreads();
    s_off[tid] = gbase + ex;
  }
  __syncthreads();
  /* stable scatter: rows of 256 keys in order; inside a row, rank by (warp, lane) */
#pragma unroll
  for (int r = 0; r < RS_ITEMS; r++) {
    u32 i = base + r * RS_THREADS + tid;
    bool valid = i < n;
    u32 d = (u32)(key[r] >> shift) & 255u;
    u32 vmask = __ballot_sync(0xffffffffu, valid);
    u32 peers = 0, rank = 0;
    bool leader = false;
    if (valid) {
      peers = __match_any_sync(vmask, d);
      rank = __popc(peers & ((1u << lane) - 1));
      leader = (rank == 0);
      if (leader) s_wcnt[warp][d] = __popc(peers);
    }
    __syncthreads();
    u32 dest = 0;
    if (valid) {
      dest = s_off[d] + rank;
      for (u32 w = 0; w < warp; w++) dest += s_wcnt[w][d];
    }
    __syncthreads();
    if (valid && leader) {
      atomicAdd(&s_off[d], (u32)__popc(peers));
      s_wcnt[warp][d] = 0;
    }
    if (valid) { keys_out[dest] = key[r]; idx_out[dest] = pay[r]; }
    __syncthreads();
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
    atomicOr(&ctl->overflow, 4u);
  }
}

/* For sorted position i: is it a new distinct value?  Emits EXACT events (check_exact_match's
   side test, ries.c:8550-8593) for duplicates of an opposite-side table row. */
__global__ void k_classify(const Row* __restrict__ surv, const u32* __restrict__ idx, u32 n,
                           const double* __restrict__ tval, const Row* __restrict__ trow, u32 nt,
                           u32* flag, u32* gap, int using_x, int skip_exact, MatchEvent* events,
                           u32 cap_events, PassCtl* ctl) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u32 s = idx[i];
  double v = surv[s].val;
  bool dup_prev = (i > 0) && (surv[idx[i - 1]].val == v);
  u32 pos = lower_bound_val(tval, nt, v);
  bool in_table = (pos < nt) && (tval[pos] == v);
  u32 fl = 0;
  if (in_table) {
    if (!skip_exact) {
      Row t = trow[pos];
      bool tree_is_lhs = (t.der != 0.0);
      if (using_x ? !tree_is_lhs : tree_is_lhs) {
        Row f = surv[s];
        if (using_x) push_event(events, cap_events, ctl, s, EV_EXACT, f, t);
        else push_event(events, cap_events, ctl, s, EV_EXACT, t, f);
      }
    }
  } else if (!dup_prev) {
    fl = 1;
  }
  flag[i] = fl;
  gap[i] = pos;
}

__global__ void k_compact_uniq(const u32* __restrict__ idx, const u32* __restrict__ flag,
                               const u32* __restrict__ flag_scan, const u32* __restrict__ gap, u32 n,
                               u32* uniq_seq, u32* uniq_gap) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (flag[i]) {
    u32 u = flag_scan[i];
    uniq_seq[u] = idx[i];
    uniq_gap[u] = gap[i];
  }
}

/* K3.  A new row's predecessor AT ITS INSERTION TIME is the old-table row below its gap unless a
   same-pass row with a smaller sequence number already sits between them (that one is same-side, so
   no test happens: check_sides, ries.c:8496-8538).  Likewise for the successor. */
__global__ void k_match(const Row* __restrict__ surv, const u32* __restrict__ uniq_seq,
                        const u32* __restrict__ uniq_gap, u32 nu, const Row* __restrict__ trow, u32 nt,
                        int using_x, double best_match0, MatchEvent* events, u32 cap_events, PassCtl* ctl) {
  u32 u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= nu) return;
  const u32 s = uniq_seq[u], g = uniq_gap[u];
  bool test_prev = g > 0, test_next = g < nt;
  if (test_prev) {
    for (u32 j = u; j-- > 0 && uniq_gap[j] == g;)
      if (uniq_seq[j] < s) { test_prev = false; break; }
  }
  if (test_next) {
    for (u32 j = u + 1; j < nu && uniq_gap[j] == g; j++)
      if (uniq_seq[j] < s) { test_next = false; break; }
  }
  if (!test_prev && !test_next) return;
  const Row f = surv[s];
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
    push_event(events, cap_events, ctl, s, side == 0 ? EV_PREV : EV_NEXT, L, R);
  }
}

__global__ void k_merge_new(const Row* __restrict__ surv, const u32* __restrict__ uniq_seq,
                            const u32* __restrict__ uniq_gap, u32 nu, Row* mrow, double* mval) {
  u32 u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= nu) return;
  Row r = surv[uniq_seq[u]];
  u32 pos = u + uniq_gap[u];
  mrow[pos] = r;
  mval[pos] = r.val;
}

__global__ void k_merge_old(const Row* __restrict__ trow, u32 nt, const u32* __restrict__ uniq_gap, u32 nu,
                            Row* mrow, double* mval) {
  u32 j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nt) return;
  /* new rows below table row j = #{u : gap[u] <= j} */
  u32 lo = 0, hi = nu;
  while (lo < hi) {
    u32 mid = (lo + hi) >> 1;
    if (uniq_gap[mid] <= j) lo = mid + 1; else hi = mid;
  }
  Row r = trow[j];
  mrow[j + lo] = r;
  mval[j + lo] = r.val;
}

/* K4 */
__global__ void k_refine(const RiesParams* __restrict__ gP, MatchEvent* events, u32 n) {
  __shared__ RiesParams P;
  {
    const u32* src = (const u32*)gP;
    u32* dst = (u32*)&P;
    for (u32 i = threadIdx.x; i < sizeof(RiesParams) / 4; i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (events[i].kind == EV_EXACT) return;
  MatchEvent ev = events[i];
  refine_event(P, ev);
  events[i] = ev;
}

/* ---------------------------------------------------------------- host-side engine ------- */

template <typename T>
struct DevBuf {
  T* p = 0;
  size_t cap = 0;
  void ensure(size_t n, bool keep = false) {
    if (n <= cap) return;
    size_t ncap = std::max(n, cap + cap / 2);
    T* q = 0;
    CUDA_CHECK(cudaMalloc(&q, ncap * sizeof(T)));
    if (keep && p && cap) CUDA_CHECK(cudaMemcpy(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice));
    if (p) CUDA_CHECK(cudaFree(p));
    p = q;
    cap = ncap;
  }
  ~DevBuf() { if (p) cudaFree(p); }
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
    CUDA_CHECK(cudaStreamCreateWithFlags(&st_, cudaStreamNonBlocking));
    int occ = 1;
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_expand, EXP_THREADS, 0));
    if (occ < 1) occ = 1;
    expand_grid_ = (u32)(num_sms_ * occ);
    staged_.ensure((size_t)expand_grid_ * EXP_MAX_ITEMS);
    CUDA_CHECK(cudaMalloc(&d_P_, sizeof(RiesParams)));
    CUDA_CHECK(cudaMalloc(&d_ctl_, sizeof(PassCtl)));
    CUDA_CHECK(cudaMalloc(&d_sortctl_, sizeof(SortCtl)));
    CUDA_CHECK(cudaMalloc(&d_opcnt_, 2 * RIES_MAX_SYMS * sizeof(u64)));
    CUDA_CHECK(cudaMallocHost(&h_ctl_, sizeof(PassCtl)));
    const char* arena_env = getenv("RIES_ARENA_CAP"); /* test hook: start tiny to exercise overflow-and-rerun */
    const size_t arena0 = arena_env ? (size_t)atoi(arena_env) : (size_t)(1u << 20);
    nodes_.ensure(arena0);
    done_.ensure(arena0);
    done_form_.ensure(arena0);
    events_.ensure(1u << 14);
    ensure_tiles();
  }
  ~DeviceEngine() override {
    cudaStreamSynchronize(st_);
    cudaFree(d_P_); cudaFree(d_ctl_); cudaFreeHost(h_ctl_);
    if (h_events_) cudaFreeHost(h_events_);
    cudaStreamDestroy(st_);
  }
  const char* name() const override { return "sm_100a"; }

  void begin(const RiesParams& P) override {
    CUDA_CHECK(cudaMemcpyAsync(d_P_, &P, sizeof(P), cudaMemcpyHostToDevice, st_));
    CUDA_CHECK(cudaMemsetAsync(d_opcnt_, 0, 2 * RIES_MAX_SYMS * sizeof(u64), st_));
    CUDA_CHECK(cudaStreamSynchronize(st_));
    memcpy(sym_chars_, P.sym_char, sizeof(sym_chars_));
    nt_ = 0;
    cur_ = 0;
    last_nu_ = 0;
    kernel_launches = 0;
    h2d_bytes = sizeof(P);
    d2h_bytes = 0;
    for (int i = 0; i < 4; i++) { stage_ms[i] = 0; stage_launches[i] = 0; }
    sort_keys = merge_rows = expand_nodes = 0;
  }

  long inserted_before(uint32_t seq, bool inclusive) override {
    if (!last_nu_) return 0;
    std::vector<u32> h(last_nu_);
    CUDA_CHECK(cudaMemcpy(h.data(), uniq_seq_.p, last_nu_ * sizeof(u32), cudaMemcpyDeviceToHost));
    long c = 0;
    for (u32 s : h) c += inclusive ? (s <= seq) : (s < seq);
    return c;
  }

  void run_pass(int using_x, int cplx, const std::vector<FormDesc>& forms, double best_match0,
                bool skip_exact, PassOutput& out) override {
    last_nu_ = 0;
    const u32 nforms = (u32)forms.size();
    if (nforms == 0) return;
    if (nforms > 65535) {
      fprintf(stderr, "ries_b200: %u forms in one pass exceeds the 16-bit form index\n", nforms);
      abort();
    }
    int max_len = 0;
    for (const FormDesc& f : forms) max_len = std::max(max_len, (int)f.flen);
    forms_.ensure(nforms);
    CUDA_CHECK(cudaMemcpyAsync(forms_.p, forms.data(), nforms * sizeof(FormDesc), cudaMemcpyHostToDevice, st_));
    form_first_.ensure(nforms); form_cnt_.ensure(nforms); form_base_.ensure(nforms);

    h2d_bytes += nforms * sizeof(FormDesc);
    /* ---- phase A: K1, rerun with larger arenas on overflow ---- */
    while (true) {
      nodes_.ensure(nforms + 16);
      cudaEvent_t ta = stage_begin();
      unsigned long long l0 = kernel_launches;
      k_init_roots<<<(nforms + 255) / 256, 256, 0, st_>>>(nodes_.p, nforms, d_ctl_);
      kernel_launches++;
      for (int lvl = 0; lvl < max_len; lvl++) {
        k_expand<<<expand_grid_, EXP_THREADS, 0, st_>>>(d_P_, forms_.p, nodes_.p, done_.p, done_form_.p,
                                                       (u32)nodes_.cap, (u32)done_.cap, d_ctl_,
                                                       (volatile u64*)tiles_.p, staged_.p, d_opcnt_, using_x, cplx);
        k_level_finish<<<1, 1024, 0, st_>>>(d_ctl_, tiles_.p, (u32)nodes_.cap, (u32)done_.cap);
        kernel_launches += 2;
      }
      stage_end(ta, 0, l0);
      CUDA_CHECK(cudaMemcpyAsync(h_ctl_, d_ctl_, sizeof(PassCtl), cudaMemcpyDeviceToHost, st_));
      CUDA_CHECK(cudaStreamSynchronize(st_));
      CUDA_CHECK(cudaGetLastError());
      d2h_bytes += sizeof(PassCtl);
      if (!(h_ctl_->overflow & 3)) break;
      /* grow (the counters say how much the level that overflowed needed; later levels may need
         more, so grow generously) and retry; the table has not been touched yet */
      if (h_ctl_->overflow & 1) nodes_.ensure(std::max<size_t>(nodes_.cap * 2, (size_t)h_ctl_->node_count * 2));
      if (h_ctl_->overflow & 2) {
        size_t want = std::max<size_t>(done_.cap * 2, (size_t)h_ctl_->done_count * 2);
        done_.ensure(want); done_form_.ensure(want);
      }
      ensure_tiles();
    }
    const u32 n = h_ctl_->done_count;
    expand_nodes += h_ctl_->node_count;
    out.generated = (double)n;
    out.dead_ends = (double)h_ctl_->dead;
    out.inserted = 0;
    if (n == 0) return;

    /* ---- phase B: order, sort, dedupe, match, merge, refine ---- */
    surv_.ensure(n);
    for (int b = 0; b < 2; b++) { keys_[b].ensure(n); idx_[b].ensure(n); }
    flag_.ensure(n); flag_scan_.ensure(n); gap_.ensure(n); uniq_seq_.ensure(n); uniq_gap_.ensure(n);
    const u32 nb256 = (n + 255) / 256;
    cudaEvent_t tb = stage_begin();
    unsigned long long l1 = kernel_launches;
    sort_keys += n;
    CUDA_CHECK(cudaMemsetAsync(form_cnt_.p, 0, nforms * sizeof(u32), st_));
    CUDA_CHECK(cudaMemsetAsync(form_first_.p, 0, nforms * sizeof(u32), st_));
    k_form_bounds<<<nb256, 256, 0, st_>>>(done_form_.p, n, form_first_.p, form_cnt_.p);
    k_form_scan<<<1, 1024, 0, st_>>>(nforms, form_first_.p, form_cnt_.p, form_base_.p);
    k_reorder<<<nb256, 256, 0, st_>>>(done_.p, done_form_.p, n, form_first_.p, form_base_.p, surv_.p,
                                     keys_[0].p, idx_[0].p);
    kernel_launches += 3;

    /* K2: one histogram pass + 8 stable onesweep digit passes */
    const u32 rs_blocks = (n + RS_CHUNK - 1) / RS_CHUNK;
    rs_state_.ensure((size_t)rs_blocks * 256);
    if (rs_state_.cap != rs_state_zeroed_) { /* fresh allocation: stale bits must not look like an epoch */
      CUDA_CHECK(cudaMemsetAsync(rs_state_.p, 0, rs_state_.cap * sizeof(u64), st_));
      rs_state_zeroed_ = rs_state_.cap;
    }
    CUDA_CHECK(cudaMemsetAsync(d_sortctl_, 0, sizeof(SortCtl), st_));
    k_rs_hist_all<<<std::min<u32>(rs_blocks, (u32)num_sms_ * 4), RS_THREADS, 0, st_>>>(keys_[0].p, n, d_sortctl_);
    kernel_launches++;
    int src = 0;
    for (int pass = 0; pass < 8; pass++) {
      rs_epoch_ = (rs_epoch_ + 1) & 0x3fffffffu;
      if (rs_epoch_ == 0) rs_epoch_ = 1;
      k_rs_onesweep<<<rs_blocks, RS_THREADS, 0, st_>>>(keys_[src].p, idx_[src].p, keys_[src ^ 1].p, idx_[src ^ 1].p, n,
                                                       pass, d_sortctl_, (volatile u64*)rs_state_.p, rs_epoch_);
      kernel_launches++;
      src ^= 1;
    }
    const u32* sidx = idx_[src].p;
    stage_end(tb, 1, l1);

    events_cap_check();
    cudaEvent_t tc = stage_begin();
    unsigned long long l2 = kernel_launches;
    k_classify<<<nb256, 256, 0, st_>>>(surv_.p, sidx, n, tval_[cur_].p, trow_[cur_].p, nt_, flag_.p, gap_.p,
                                      using_x, skip_exact ? 1 : 0, events_.p, (u32)events_.cap, d_ctl_);
    kernel_launches++;
    exclusive_scan(flag_.p, flag_scan_.p, n, &d_ctl_->n_uniq);
    k_compact_uniq<<<nb256, 256, 0, st_>>>(sidx, flag_.p, flag_scan_.p, gap_.p, n, uniq_seq_.p, uniq_gap_.p);
    kernel_launches++;
    CUDA_CHECK(cudaMemcpyAsync(h_ctl_, d_ctl_, sizeof(PassCtl), cudaMemcpyDeviceToHost, st_));
    CUDA_CHECK(cudaStreamSynchronize(st_));
    d2h_bytes += sizeof(PassCtl);
    const u32 nu = h_ctl_->n_uniq;
    out.inserted = (long)nu;
    last_nu_ = nu;

    if (nu) {
      const u32 nbu = (nu + 255) / 256;
      k_match<<<nbu, 256, 0, st_>>>(surv_.p, uniq_seq_.p, uniq_gap_.p, nu, trow_[cur_].p, nt_, using_x,
                                   best_match0, events_.p, (u32)events_.cap, d_ctl_);
      trow_[cur_ ^ 1].ensure((size_t)nt_ + nu);
      tval_[cur_ ^ 1].ensure((size_t)nt_ + nu);
      k_merge_new<<<nbu, 256, 0, st_>>>(surv_.p, uniq_seq_.p, uniq_gap_.p, nu, trow_[cur_ ^ 1].p, tval_[cur_ ^ 1].p);
      kernel_launches += 2;
      if (nt_) {
        k_merge_old<<<(nt_ + 255) / 256, 256, 0, st_>>>(trow_[cur_].p, nt_, uniq_gap_.p, nu, trow_[cur_ ^ 1].p,
                                                       tval_[cur_ ^ 1].p);
        kernel_launches++;
      }
    }
    if (nu) merge_rows += (unsigned long long)nt_ + nu;
    stage_end(tc, 2, l2);
    /* events: count, (re)run on overflow is not needed -- cap is generous and checked */
    CUDA_CHECK(cudaMemcpyAsync(h_ctl_, d_ctl_, sizeof(PassCtl), cudaMemcpyDeviceToHost, st_));
    CUDA_CHECK(cudaStreamSynchronize(st_));
    d2h_bytes += sizeof(PassCtl);
    while (h_ctl_->overflow & 4) {
      /* more candidates than the buffer holds (huge --max-match-distance, or -i with many exact hits):
         grow it and re-emit; the old table is still intact (ping-pong) and flags/gaps are unchanged */
      events_.ensure((size_t)h_ctl_->n_events + 1024);
      CUDA_CHECK(cudaMemsetAsync(&d_ctl_->n_events, 0, sizeof(u32), st_));
      CUDA_CHECK(cudaMemsetAsync(&d_ctl_->overflow, 0, sizeof(u32), st_));
      k_classify<<<nb256, 256, 0, st_>>>(surv_.p, sidx, n, tval_[cur_].p, trow_[cur_].p, nt_, flag_.p, gap_.p,
                                        using_x, skip_exact ? 1 : 0, events_.p, (u32)events_.cap, d_ctl_);
      kernel_launches++;
      if (nu) {
        k_match<<<(nu + 255) / 256, 256, 0, st_>>>(surv_.p, uniq_seq_.p, uniq_gap_.p, nu, trow_[cur_].p, nt_, using_x,
                                                  best_match0, events_.p, (u32)events_.cap, d_ctl_);
        kernel_launches++;
      }
      CUDA_CHECK(cudaMemcpyAsync(h_ctl_, d_ctl_, sizeof(PassCtl), cudaMemcpyDeviceToHost, st_));
      CUDA_CHECK(cudaStreamSynchronize(st_));
      d2h_bytes += sizeof(PassCtl);
    }
    const u32 ne = h_ctl_->n_events;
    if (ne) {
      cudaEvent_t td = stage_begin();
      unsigned long long l3 = kernel_launches;
      k_refine<<<(ne + 63) / 64, 64, 0, st_>>>(d_P_, events_.p, ne);
      kernel_launches++;
      stage_end(td, 3, l3);
      d2h_bytes += (unsigned long long)ne * sizeof(MatchEvent);
      if (ne > h_events_cap_) {
        if (h_events_) cudaFreeHost(h_events_);
        h_events_cap_ = std::max<size_t>(ne, h_events_cap_ * 2);
        CUDA_CHECK(cudaMallocHost(&h_events_, h_events_cap_ * sizeof(MatchEvent)));
      }
      CUDA_CHECK(cudaMemcpyAsync(h_events_, events_.p, ne * sizeof(MatchEvent), cudaMemcpyDeviceToHost, st_));
      CUDA_CHECK(cudaStreamSynchronize(st_));
      out.events.assign(h_events_, h_events_ + ne);
      std::sort(out.events.begin(), out.events.end(), [](const MatchEvent& a, const MatchEvent& b) {
        return a.seq != b.seq ? a.seq < b.seq : a.kind < b.kind;
      });
    }
    if (nu) {
      cur_ ^= 1;
      nt_ += nu;
    }
    CUDA_CHECK(cudaGetLastError());
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
  u32 table_rows() const { return nt_; }
  unsigned char sym_chars_[RIES_MAX_SYMS];
  void read_opcounts(unsigned long long* out /* [2][RIES_MAX_SYMS] */) {
    CUDA_CHECK(cudaMemcpy(out, d_opcnt_, 2 * RIES_MAX_SYMS * sizeof(u64), cudaMemcpyDeviceToHost));
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
  void collect_stage_times() {
    for (size_t i = 0; i < pend_.size(); i++) {
      float ms = 0;
      cudaEventElapsedTime(&ms, pend_[i].a, pend_[i].b);
      stage_ms[pend_[i].stage] += ms;
      ev_pool_.push_back(pend_[i].a);
      ev_pool_.push_back(pend_[i].b);
    }
    pend_.clear();
  }

 private:
  struct Pend { cudaEvent_t a, b; int stage; };
  std::vector<Pend> pend_;
  std::vector<cudaEvent_t> ev_pool_;
  cudaEvent_t get_event() {
    if (ev_pool_.empty()) {
      cudaEvent_t e;
      CUDA_CHECK(cudaEventCreate(&e));
      return e;
    }
    cudaEvent_t e = ev_pool_.back();
    ev_pool_.pop_back();
    return e;
  }
  cudaEvent_t stage_begin() {
    cudaEvent_t a = get_event();
    CUDA_CHECK(cudaEventRecord(a, st_));
    return a;
  }
  void stage_end(cudaEvent_t a, int stage, unsigned long long launches_before) {
    cudaEvent_t b = get_event();
    CUDA_CHECK(cudaEventRecord(b, st_));
    pend_.push_back(Pend{a, b, stage});
    stage_launches[stage] += kernel_launches - launches_before;
  }
  void ensure_tiles() {
    size_t want = nodes_.cap / EXP_PARENTS + 2;
    if (want > tiles_.cap) {
      tiles_.ensure(want);
      CUDA_CHECK(cudaMemsetAsync(tiles_.p, 0, tiles_.cap * sizeof(u64), st_));
    }
  }
  void events_cap_check() {
    /* the window filter keeps the real number tiny; on overflow run_pass grows the buffer and re-emits */
    const char* cap_env = getenv("RIES_EVENTS_CAP"); /* test hook: start tiny to exercise the regrow path */
    events_.ensure(cap_env ? (size_t)atoi(cap_env) : (size_t)(1u << 16));
  }
  void exclusive_scan(const u32* in, u32* out, u32 n, u32* total_out) {
    u32 nb = (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
    if (nb <= 1) {
      k_scan_apply<<<1, SCAN_THREADS, 0, st_>>>(in, out, n, 0, total_out);
      kernel_launches++;
      return;
    }
    int lvl = scan_depth_++;
    if ((int)scan_tmp_.size() <= lvl) scan_tmp_.resize(lvl + 1);
    if (!scan_tmp_[lvl]) scan_tmp_[lvl] = new DevBuf<u32>();
    scan_tmp_[lvl]->ensure(nb);
    u32* sums = scan_tmp_[lvl]->p;
    k_scan_reduce<<<nb, SCAN_THREADS, 0, st_>>>(in, n, sums);
    kernel_launches++;
    exclusive_scan(sums, sums, nb, 0);
    k_scan_apply<<<nb, SCAN_THREADS, 0, st_>>>(in, out, n, sums, total_out);
    kernel_launches++;
    scan_depth_--;
  }

  cudaStream_t st_;
  int num_sms_ = 0;
  u32 expand_grid_ = 0;
  RiesParams* d_P_ = 0;
  PassCtl* d_ctl_ = 0;
  PassCtl* h_ctl_ = 0;
  SortCtl* d_sortctl_ = 0;
  u64* d_opcnt_ = 0;
  DevBuf<u64> rs_state_;
  size_t rs_state_zeroed_ = 0;
  u32 rs_epoch_ = 0;
  DevBuf<FormDesc> forms_;
  DevBuf<Node> nodes_;
  DevBuf<Row> done_, surv_;
  DevBuf<uint16_t> done_form_;
  DevBuf<u64> tiles_;
  DevBuf<Staged> staged_;
  DevBuf<u32> form_first_, form_cnt_, form_base_;
  DevBuf<u64> keys_[2];
  DevBuf<u32> idx_[2];
  DevBuf<u32> hist_, flag_, flag_scan_, gap_, uniq_seq_, uniq_gap_;
  DevBuf<Row> trow_[2];
  DevBuf<double> tval_[2];
  DevBuf<MatchEvent> events_;
  MatchEvent* h_events_ = 0;
  size_t h_events_cap_ = 0;
  std::vector<DevBuf<u32>*> scan_tmp_;
  int scan_depth_ = 0;
  u32 nt_ = 0;
  int cur_ = 0;
  u32 last_nu_ = 0;
};

Engine* make_device_engine() { return new DeviceEngine(); }

void device_engine_report(Engine* e, DeviceReport* r) {
  DeviceEngine* d = static_cast<DeviceEngine*>(e);
  d->collect_stage_times();
  r->kernel_launches = d->kernel_launches;
  r->h2d_bytes = d->h2d_bytes;
  r->d2h_bytes = d->d2h_bytes;
  for (int i = 0; i < 4; i++) { r->stage_ms[i] = d->stage_ms[i]; r->stage_launches[i] = d->stage_launches[i]; }
  r->sort_keys = d->sort_keys;
  r->merge_rows = d->merge_rows;
  r->expand_nodes = d->expand_nodes;
  r->table_rows = d->table_rows();
  d->read_opcounts(&r->op_count[0][0]);
  memcpy(r->op_char, d->sym_chars_, sizeof(r->op_char));
}

double device_engine_dfma_peak(Engine* e) { return static_cast<DeviceEngine*>(e)->dfma_peak_tflops(); }

}  // namespace ries
