// K1 of the TINY passes: enumerate + evaluate one expression position of every live prefix (ge_2 + exec) inside
// the single block of k_tiny_front (tiny_pass.cuh), children appended in the reference's insertion order.
// The general kernel is k1_free.cuh.
// Part of the single translation unit engine.cu (included inside namespace ries).
#pragma once

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
/* live prefixes per thread and tile: a tile of EXP_PPT x EXP_THREADS prefixes yields ~1.3 x that many
   (prefix, symbol) pairs, i.e. more 32-pair batches per warp in phase 3 (less waiting for the last batch)
   and fewer block barriers per prefix */
#ifndef EXP_PPT
#define EXP_PPT 1
#endif
#define EXP_PARENTS (EXP_PPT * EXP_THREADS)
#define EXP_VWARPS (EXP_PARENTS / 32)
#define EXP_MAX_ITEMS (EXP_PARENTS * RIES_SLOTS)
#ifndef EXP_MIN_BLOCKS
#define EXP_MIN_BLOCKS 3
#endif
#define IT_LIVE 0x4000u
#define IT_DONE 0x8000u

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

/* One expression position of the running pass.  TINY = called by the first EXP_THREADS threads of the
   1024-thread single-block kernel k_tiny_front (which loops over the positions itself): the block barrier is
   then a named barrier over those threads only. */
template <bool TINY>
__device__ __forceinline__ void exp_sync() {
  if (TINY) asm volatile("bar.sync 1, %0;" ::"n"(EXP_THREADS) : "memory");
  else __syncthreads();
}

/* TINY (single block, see tiny_pass.cuh): no other block exists, so the tile ticket, the published tile
   aggregates, the look-back and the last-block election -- each a dependent round trip to L2 -- are replaced
   by a loop counter and running sums in shared memory, the search constants are loaded once per pass
   (`first`), and the frontier bounds travel in shared memory (s_lvl) instead of the control block. */
template <bool TINY>
__device__ __forceinline__ void expand_level(const DevPtrs* __restrict__ Dp, const PassArgs* __restrict__ Ap, PassCtl* ctl,
                                             const bool first = true) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  TileBuf* s_buf = (TileBuf*)s_dyn; /* [2] */
  __shared__ u32 s_opcnt[RIES_MAX_SYMS];
  __shared__ RiesParams P;
  __shared__ TestTables TT;
  __shared__ u32 s_wcnt[EXP_VWARPS][RIES_SLOTS];
  __shared__ u32 s_slot_base[RIES_SLOTS];
  __shared__ u32 s_wl[EXP_VWARPS], s_wd[EXP_VWARPS];
  __shared__ u32 s_live_base, s_done_base, s_dead, s_tile, s_last, s_next;
  __shared__ u32 s_run_live, s_run_done; /* TINY: children appended by the tiles before the current one */
  __shared__ u32 s_lvl[4];               /* TINY: lvl_begin, lvl_end, done_count, overflow */

  const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (TINY && first) {
    if (tid == 0) { s_lvl[0] = ctl->lvl_begin; s_lvl[1] = ctl->lvl_end; s_lvl[2] = ctl->done_count; s_lvl[3] = ctl->overflow; }
    exp_sync<TINY>();
  }
  const DevPtrs D = *Dp;
  const u32 lvl_begin = TINY ? s_lvl[0] : ctl->lvl_begin, lvl_end = TINY ? s_lvl[1] : ctl->lvl_end;
  const u32 n = lvl_end - lvl_begin;
  const u32 ntiles = (n + EXP_PARENTS - 1) / EXP_PARENTS;
  const int using_x = Ap->using_x, cplx = Ap->cplx;
  const u32 cap_nodes = Ap->cap_nodes, cap_done = Ap->cap_done;
  if (ntiles != 0 && !((TINY ? s_lvl[3] : ctl->overflow) & (OVF_NODES | OVF_DONE))) {
    if (TINY && first) {
      const u32* src = (const u32*)D.P;
      u32* dst = (u32*)&P;
      for (u32 i = tid; i < sizeof(RiesParams) / 4; i += EXP_THREADS) dst[i] = src[i];
      const u32* tsrc = (const u32*)D.tt;
      u32* tdst = (u32*)&TT;
      for (u32 i = tid; i < sizeof(TestTables) / 4; i += EXP_THREADS) tdst[i] = tsrc[i];
    }
    if (tid == 0) { s_dead = 0; s_run_live = 0; s_run_done = 0; }
    if (tid < RIES_MAX_SYMS) s_opcnt[tid] = 0;
    exp_sync<TINY>();
    const u32 done_base0 = TINY ? s_lvl[2] : ctl->done_count; /* read-only during the tile loop (updated by the last block) */
    const u32 lt = (1u << lane) - 1;
    Node* nodes = D.nodes;
    const FormDesc* forms = D.forms;
    u32 my_dead = 0;
    int cur = 0;
    bool have_prev = false;
    u32 tiny_tile = 0;

    /* The one block takes the tiles in order and keeps running sums of the children appended so far; a tile's
       children are placed (phase 4) after the next tile's front half, as in the round-1 general kernel this
       code comes from. */
    while (true) {
      if (tid == 0) s_next = 0;
      exp_sync<TINY>();
      const u32 tile = TINY ? tiny_tile++ : s_tile;
      const bool have_cur = tile < ntiles;
      if (have_cur) {
        TileBuf& B = s_buf[cur];
        Staged* staged = D.staged + ((size_t)blockIdx.x * 2 + cur) * EXP_MAX_ITEMS;
        const u32 p0 = lvl_begin + tile * EXP_PARENTS;
        /* ---- phase 1: integer tests, one thread per prefix ---- */
        u32 mask[EXP_PPT];
#pragma unroll
        for (int j = 0; j < EXP_PPT; j++) {
          const u32 pl = (u32)j * EXP_THREADS + tid;
          const u32 p = p0 + pl;
          mask[j] = 0;
          if (p < lvl_end) {
            const Node par = nodes[p];
            B.par[pl] = par;
            const FormDesc& F = forms[par.form];
            const int cls = F.cls[par.len];
            const int ns = P.cls_n[cls];
            const int atts = prefix_attrs(P, F, par, using_x);
            mask[j] = expand_mask(P, TT, F, par, cls, atts, using_x, cplx, cplx);
            my_dead += (u32)ns - (u32)__popc(mask[j]);
            B.cls[pl] = (unsigned char)cls;
            B.complete[pl] = (unsigned char)(par.len + 1 >= F.flen);
          }
          B.live[pl] = 0;
          B.done[pl] = 0;
        }
        /* ---- phase 2: list the surviving pairs grouped by slot ---- */
#pragma unroll
        for (int j = 0; j < EXP_PPT; j++) {
#pragma unroll
          for (int k = 0; k < RIES_SLOTS; k++) {
            u32 b = __ballot_sync(0xffffffffu, (mask[j] >> k) & 1u);
            if (lane == 0) s_wcnt[j * (EXP_THREADS / 32) + warp][k] = __popc(b);
          }
        }
        exp_sync<TINY>();
        if (tid < RIES_SLOTS) { /* per slot: exclusive prefix over (virtual) warps, and the slot's total */
          u32 run = 0;
          for (int w = 0; w < EXP_VWARPS; w++) { u32 c = s_wcnt[w][tid]; s_wcnt[w][tid] = run; run += c; }
          s_slot_base[tid] = run;
        }
        exp_sync<TINY>();
        if (tid == 0) {
          u32 run = 0;
          for (int k = 0; k < RIES_SLOTS; k++) { u32 c = s_slot_base[k]; s_slot_base[k] = run; run += c; }
          B.total = run;
          B.tile = tile;
        }
        exp_sync<TINY>();
#pragma unroll
        for (int j = 0; j < EXP_PPT; j++) {
          const u32 pl = (u32)j * EXP_THREADS + tid;
#pragma unroll
          for (int k = 0; k < RIES_SLOTS; k++) {
            u32 b = __ballot_sync(0xffffffffu, (mask[j] >> k) & 1u);
            if ((mask[j] >> k) & 1u)
              B.item[s_slot_base[k] + s_wcnt[j * (EXP_THREADS / 32) + warp][k] + __popc(b & lt)] = (unsigned short)((pl << 4) | k);
          }
        }
        exp_sync<TINY>();
        const u32 T = B.total;
        /* ---- phase 3: exec + value prunes; warps pull 32 listed pairs at a time (a warp's batch is
           almost always one operator, and heavy operators spread over the warps) ---- */
        while (true) {
          u32 b0 = 0;
          if (lane == 0) b0 = atomicAdd(&s_next, 32u);
          b0 = __shfl_sync(0xffffffffu, b0, 0);
          if (b0 >= T) break;
#ifndef EXP_FORWARD_BATCHES
          /* batches from the end of the list first: the list is in slot order and the heavy operators of a class
             (^ v L A, the trigonometric functions) are its last slots, so they start first and the cheap
             constants fill the tail of the phase (K1 45.9 -> 44.8 ms at -l6) */
          const u32 i = (((T + 31u) >> 5) - 1u - (b0 >> 5)) * 32u + lane;
#else
          const u32 i = b0 + lane;
#endif
          if (i < T) {
            const u32 it = B.item[i];
            const u32 pl = (it >> 4) & 1023u, k = it & 15u;
            const int cls = B.cls[pl];
            const int g = P.trial[cls][k];
            Node child;
            atomicAdd(&s_opcnt[g], 1u); /* exec() calls per symbol: the algorithmic work of this kernel */
            int oc = expand_exec(P, cls, g, B.complete[pl] != 0, nodes, B.par[pl], p0 + pl, using_x, child);
            if (oc == OUT_DEAD) my_dead++;
            else if (oc == OUT_LIVE || oc == OUT_DONE) {
              Staged st;
              st.val = child.val; st.der = child.der; st.below = child.below; st.tag = child.tag; st.sp = child.sp;
              st.pad0 = child.pad0; st.pad1 = 0; /* pad0: canon_apply's code */
              staged[i] = st;
              if (oc == OUT_LIVE) { atomicOr(&B.live[pl], 1u << k); B.item[i] = (unsigned short)(it | IT_LIVE); }
              else { atomicOr(&B.done[pl], 1u << k); B.item[i] = (unsigned short)(it | IT_DONE); }
            }
          }
        }
        exp_sync<TINY>();
        /* child offsets inside the tile, and the tile's aggregate -> published at once */
        {
          u32 xl[EXP_PPT], xd[EXP_PPT]; /* exclusive offsets inside the (virtual) warp */
#pragma unroll
          for (int j = 0; j < EXP_PPT; j++) {
            const u32 pl = (u32)j * EXP_THREADS + tid;
            u32 nl = __popc(B.live[pl]), nd = __popc(B.done[pl]);
            u32 il = nl, id = nd;
            for (int o = 1; o < 32; o <<= 1) {
              u32 a = __shfl_up_sync(0xffffffffu, il, o), b = __shfl_up_sync(0xffffffffu, id, o);
              if (lane >= (u32)o) { il += a; id += b; }
            }
            if (lane == 31) { s_wl[j * (EXP_THREADS / 32) + warp] = il; s_wd[j * (EXP_THREADS / 32) + warp] = id; }
            xl[j] = il - nl; xd[j] = id - nd;
          }
          exp_sync<TINY>();
          u32 la = 0, da = 0;
          for (u32 w = 0; w < EXP_VWARPS; w++) { la += s_wl[w]; da += s_wd[w]; }
#pragma unroll
          for (int j = 0; j < EXP_PPT; j++) {
            const u32 vw = (u32)j * (EXP_THREADS / 32) + warp;
            const u32 pl = (u32)j * EXP_THREADS + tid;
            u32 wl = 0, wd = 0;
            for (u32 w = 0; w < vw; w++) { wl += s_wl[w]; wd += s_wd[w]; }
            B.lbase[pl] = wl + xl[j];
            B.dbase[pl] = wd + xd[j];
          }
          if (tid == 0) {
            B.live_agg = la; B.done_agg = da;
          }
        }
      }
      /* ---- phase 4 of the PREVIOUS tile: ordered append ---- */
      if (have_prev) {
        TileBuf& B = s_buf[cur ^ 1];
        const Staged* staged = D.staged + ((size_t)blockIdx.x * 2 + (cur ^ 1)) * EXP_MAX_ITEMS;
        if (tid == 0) {
          s_live_base = s_run_live; s_done_base = s_run_done;
          s_run_live += B.live_agg; s_run_done += B.done_agg;
        }
        exp_sync<TINY>();
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
          child.pad0 = st.pad0;
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
      exp_sync<TINY>(); /* s_tile, s_live_base and the buffers are reused by the next round */
      if (!have_cur) break;
      have_prev = true;
      cur ^= 1;
    }
    /* block-level dead-end count -> one atomic per block */
    for (int o = 16; o > 0; o >>= 1) my_dead += __shfl_xor_sync(0xffffffffu, my_dead, o);
    if (lane == 0 && my_dead) atomicAdd(&s_dead, my_dead);
    exp_sync<TINY>();
    if (tid == 0 && s_dead) atomicAdd(&ctl->dead, (u64)s_dead);
    if (tid < RIES_MAX_SYMS && s_opcnt[tid]) atomicAdd(&D.opcnt[(using_x ? RIES_MAX_SYMS : 0) + tid], (u64)s_opcnt[tid]);
  }

  if (TINY) {
    /* level finish, single block: same bookkeeping, no election, no tile states to clear */
    exp_sync<TINY>();
    if (tid == 0) {
      u32 ovf = s_lvl[3];
      if (!(ovf & (OVF_NODES | OVF_DONE))) {
        const u32 live = ntiles ? s_run_live : 0, dn = ntiles ? s_run_done : 0;
        u64 new_nodes = (u64)lvl_end + live;
        u64 new_done = (u64)s_lvl[2] + dn;
        if (new_nodes > cap_nodes) ovf |= OVF_NODES;
        if (new_done > cap_done) ovf |= OVF_DONE;
        s_lvl[3] = ovf;
        s_lvl[0] = lvl_end;
        s_lvl[1] = (ovf & (OVF_NODES | OVF_DONE)) ? lvl_end : (u32)new_nodes;
        s_lvl[2] = (new_done > 0xffffffffull) ? 0xffffffffu : (u32)new_done;
        if (ovf) atomicOr(&ctl->overflow, ovf);
        ctl->lvl_begin = s_lvl[0];
        ctl->lvl_end = s_lvl[1];
        ctl->node_count = (new_nodes > 0xffffffffull) ? 0xffffffffu : (u32)new_nodes;
        ctl->done_count = s_lvl[2];
      }
    }
    exp_sync<TINY>();
    return;
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
