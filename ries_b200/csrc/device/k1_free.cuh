// K1, general kernel: enumerate + evaluate one expression position of every live prefix (ge_2 + exec).
// Operator queues in shared memory, children written where they are produced -- no ordered append.
// Part of the single translation unit engine.cu (included inside namespace ries).
#pragma once

/* Round 1's K1 emitted children in the reference's insertion order (decoupled look-back over tile aggregates plus
   a placement phase per tile) and ran exec() on batches cut from ONE tile's (prefix, symbol) pairs.  ncu: 40 % of
   the stall samples at block barriers (look-back wait, slowest warp of the exec phase), 24 of 32 lanes active.
   A first order-free version of this kernel (profiles/r2_*) removed the look-back and showed what was under it:
   a tile of 256 prefixes yields ~340 pairs over ~20 operators, i.e. batches of two or three operators each
   (17 of 32 lanes active) hopping through 100 KB of exec() code (stall_no_instruction 44 %).  Hence:

     ORDER-FREE  insertion order is a pure function of (form, symbols) (state.cuh, SeqKey): nothing downstream
                 needs it materialised, so children go wherever there is room.
     QUEUES      every block keeps one queue PER (complete?, SYMBOL) in shared memory, across tiles.  A round
                 = phase P: thread = prefix of the next tile (taken by ticket): the integer symbol-append tests
                   for ALL candidate symbols at once from host-built bit-mask tables (expand_mask;
                   ries.c:9401-9445), every surviving pair appended to its symbol's queue;   -- block barrier
                 phase C: only FULL batches of 32 are taken off the queues (the rest waits for the next tile),
                   in queue order, so a batch is exactly one operator with 32 active lanes and neighbouring
                   warps run the same code: exec() + value prunes (ries.c:5308-5950, 8962-9076) without
                   divergence; the child -- a longer prefix or a complete expression -- is written straight to
                   slot (round base + index in the round); thread 0 reserves the round's room with one atomic
                   per output array while the first batches are already running.
                 The last round of a block drains what is left.  A pair that fails leaves a HOLE (form =
                 FORM_HOLE), skipped by whoever reads the slot next (~11 % of the slots at -l6).
     SPREAD      a position with fewer prefixes than blocks x XF_TILE (four launches in five) would leave most
                 blocks idle while a few run ~8 batches per warp one after the other: it is cut into one small
                 tile per block instead (block b = tile b, no ticket), whose single round drains the queues.
                 PassArgs::k1_spread; off for the engines of a batch, which share the GPU and want full batches.
   Two barriers per round, no look-back, no staging of children, no partial batches but the last. */
#ifndef XF_THREADS
#define XF_THREADS 448 /* 2 blocks of 14 warps per SM; measured 256 x 2 or 3: 47.5 ms, 384 x 2: 43.2, 448 x 2: 42.0, 512 x 1: 45.4 */
#endif
#ifndef XF_TILE
#define XF_TILE 416 /* prefixes per round: two blocks' queues (62 x (XF_TILE + 32) x 4 B at the default symbol set) fill an SM's shared memory */
#endif
#ifndef XF_MIN_BLOCKS
#define XF_MIN_BLOCKS 2
#endif
#ifndef XF_MIN_TILE
#define XF_MIN_TILE 8 /* smallest tile of a position spread over the grid; measured at -l6: fixed 416: K1 42.8 ms, >= 128: 42.2, >= 64: 41.6, >= 32: 40.9, >= 16: 40.3, >= 8: 40.1 */
#endif
#define XF_CAP (XF_TILE + 32) /* a queue holds at most 31 left-overs + one tile's pairs of that symbol */
#define XF_MAXQ (2 * RIES_MAX_SYMS)

__device__ __forceinline__ void push_tail(TailRec* tail, u32 cap, PassCtl* ctl, double lo, double hi, int val) {
  const u32 e = atomicAdd(&ctl->n_tail, 1u);
  if (e < cap) { TailRec r; r.lo = lo; r.hi = hi; r.val = val; r.pad = 0; tail[e] = r; }
  else atomicOr(&ctl->overflow, OVF_TAIL);
}

/* dynamic shared memory: nq queues of XF_CAP entries (u32: index of the prefix in the node arena) */
__host__ __device__ inline size_t xf_dyn_smem(int nq) { return (size_t)nq * XF_CAP * sizeof(u32); }

/* PERSIST = false: one launch per expression position (the launches of a pass are chained by programmatic
   dependent launch).  PERSIST = true (k_expand_all): ONE launch runs every position of the pass, the blocks meeting
   at a grid barrier between positions -- for the small and mid-size passes, where a launch per position costs
   more than the position (a pass of 1e5 prefixes: ~15 us per launch against ~4 us per barrier).  The grid must be
   co-resident (cooperative launch, at most occupancy x SMs blocks). */
template <bool PERSIST>
__device__ __forceinline__ void expand_body(const DevPtrs* __restrict__ Dp, const PassArgs* __restrict__ Ap, PassCtl* ctl) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  u32* const s_ent = (u32*)s_dyn; /* [nq][XF_CAP] rings */
  __shared__ RiesParams P;
  __shared__ TestTables TT;
  __shared__ u32 s_tail[XF_MAXQ];    /* entries ever appended to queue q */
  __shared__ u32 s_head[2][XF_MAXQ]; /* entries ever taken off, by round parity (read-only during a round) */
  __shared__ u32 s_cum[XF_MAXQ + 1]; /* this round: first batch of queue q in the round's batch list */
  __shared__ u32 s_opcnt[RIES_MAX_SYMS];
  __shared__ u32 s_tile, s_next[2], s_dead, s_gen, s_last, s_drain;
  __shared__ volatile u32 s_node_base, s_done_base;
  __shared__ volatile u32 s_alloc_tag;
  const u32 tid = threadIdx.x, lane = tid & 31;
  const DevPtrs D = *Dp;
  {
    /* Programmatic dependent launch: the launches of one pass follow each other with the attribute set (see
       enqueue_stage), so this kernel's blocks may start while the previous expression position is still
       draining.  What does not depend on it -- the search constants -- is fetched first; then the next position
       is allowed to start its own prologue, and only then do we wait for the previous one to be complete and
       visible.  (Both instructions are no-ops in a launch without the attribute.) */
    const u32* src = (const u32*)D.P;
    u32* dst = (u32*)&P;
    for (u32 i = tid; i < sizeof(RiesParams) / 4; i += XF_THREADS) dst[i] = src[i];
    const u32* tsrc = (const u32*)D.tt;
    u32* tdst = (u32*)&TT;
    for (u32 i = tid; i < sizeof(TestTables) / 4; i += XF_THREADS) tdst[i] = tsrc[i];
    if (tid == 0) { s_dead = 0; s_gen = 0; s_alloc_tag = 0; }
    if (tid < RIES_MAX_SYMS) s_opcnt[tid] = 0;
    if (tid < XF_MAXQ) { s_tail[tid] = 0; s_head[0][tid] = 0; s_head[1][tid] = 0; }
    if (!PERSIST) {
      asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
      asm volatile("griddepcontrol.wait;" ::: "memory");
    }
  }
  __syncthreads();
  const int using_x = Ap->using_x, cplx = Ap->cplx;
  const u32 cap_nodes = Ap->cap_nodes, cap_done = Ap->cap_done;
  /* shared RHS pass (SharedRhs, engine.cu): M rides in `der`; records for what depends on a target's bound */
  const bool track_m = Ap->track_m != 0;
  const double p_min = Ap->p_min;
  const u32 cap_tail = Ap->cap_tail;
  const bool spread = Ap->k1_spread != 0;
  const u32 shard_world = (u32)Ap->shard_world, shard_rank = (u32)Ap->shard_rank;
  /* queue of (complete, class, slot) = complete * nsq + qoff[class] + slot; symbol of queue q = sym_of_q */
  const u32 qoff1 = (u32)P.cls_n[0], qoff2 = qoff1 + (u32)P.cls_n[1], nsq = qoff2 + (u32)P.cls_n[2];
  const u32 nq = 2 * nsq;
  Node* nodes = D.nodes;
  const FormDesc* __restrict__ forms = D.forms;
  Row* const done_rows = D.done;
  uint16_t* const done_forms = D.done_form;
  ASSUME_GLOBAL(nodes); ASSUME_GLOBAL(forms); ASSUME_GLOBAL(done_rows); ASSUME_GLOBAL(done_forms); ASSUME_GLOBAL(ctl);
  u32 my_dead = 0, my_gen = 0;
  const u32 bar_gen0 = PERSIST ? *(volatile u32*)&ctl->bar_gen : 0u;
  for (u32 lvl = 0; lvl == 0 || (PERSIST && lvl <= RIES_MAX_ELEN); lvl++) {
  if (PERSIST && lvl) { /* the queues and the hand-off word start every position empty */
    if (tid == 0) s_alloc_tag = 0;
    if (tid < XF_MAXQ) { s_tail[tid] = 0; s_head[0][tid] = 0; s_head[1][tid] = 0; }
    __syncthreads();
  }
  const u32 lvl_begin = *(volatile u32*)&ctl->lvl_begin, lvl_end = *(volatile u32*)&ctl->lvl_end;
  const u32 n = lvl_end - lvl_begin;
  /* A position that cannot fill the grid with full tiles (most of the launches of a search) is spread over all the
     blocks instead, one smaller tile each: its time is the chain of batches a warp runs one after the other, not
     the throughput.  Block b then takes tile b -- no ticket -- and its only round drains the queues. */
  const bool one_round = spread && n < XF_TILE * gridDim.x;
  const u32 tile_sz = one_round ? min((u32)XF_TILE, max((u32)XF_MIN_TILE, ((n + gridDim.x - 1) / gridDim.x + XF_MIN_TILE - 1u) & ~(XF_MIN_TILE - 1u))) : (u32)XF_TILE;
  const u32 ntiles = (n + tile_sz - 1) / tile_sz;
  if (PERSIST && ntiles == 0) break; /* uniform over the grid: every block reads the same bounds after the barrier */
  if (ntiles != 0 && !(*(volatile u32*)&ctl->overflow & (OVF_NODES | OVF_DONE))) {
    bool draining = false;
    for (u32 par2 = 0; !draining; par2 ^= 1u) {
      /* the batch cursor and the queue heads alternate between two copies: there is no barrier at the end of a
         round, so a warp that is done may be here while another still pulls batches of the previous round; the
         copy written here was last read two rounds ago, which every warp has left (it passed the barriers) */
      if (tid == 0) { s_tile = one_round ? blockIdx.x : atomicAdd(&ctl->tile_ticket, 1u); s_next[par2] = 0; }
      __syncthreads(); /* also: every warp has left phase C of the previous round (queue entries may be overwritten) */
      const u32 tile = s_tile;
      const bool have_tile = tile < ntiles;
      /* ---- phase P ---- */
      if (have_tile) {
        const u32 p = lvl_begin + tile * tile_sz + tid;
        if (tid < tile_sz && p < lvl_end) {
          const Node par = nodes[p];
          if (par.form != FORM_HOLE) {
            const FormDesc& F = forms[par.form];
            const int cls = F.cls[par.len];
            const int ns = P.cls_n[cls];
            const int atts = prefix_attrs(P, F, par, using_x);
            u32 mask = expand_mask(P, TT, F, par, cls, atts, using_x, cplx, cplx);
            u32 rejected = (u32)ns - (u32)__popc(mask);
            if (shard_world > 1 && par.len == 0) {
              /* one search over several GPUs: rank r enumerates the (form, first symbol) pairs with
                 (13 form + slot) mod world == r -- finer than whole forms, whose sizes differ by orders of
                 magnitude; a rejected first symbol is a dead-end of the rank that owns the pair */
              u32 own = 0;
              for (int k = 0; k < ns; k++)
                if ((13u * (u32)par.form + (u32)k) % shard_world == shard_rank) own |= 1u << k;
              rejected = (u32)__popc(own & ~mask);
              mask &= own;
            }
            my_dead += rejected;
            if (track_m && par.der >= p_min && rejected) /* gone with the prefix, for a smaller bound */
              push_tail(D.tail, cap_tail, ctl, -1.0 / 0.0, par.der, -(int)rejected);
            const u32 q0 = ((par.len + 1 >= F.flen) ? nsq : 0u) + (cls == 0 ? 0u : cls == 1 ? qoff1 : qoff2);
            while (mask) {
              const u32 k = (u32)__ffs((int)mask) - 1;
              mask &= mask - 1;
              const u32 pos = atomicAdd(&s_tail[q0 + k], 1u);
              s_ent[(q0 + k) * XF_CAP + pos % XF_CAP] = p;
            }
          }
        }
      }
      /* the block's LAST round takes everything off the queues: when it got no tile, or when the tile it just got
         was the last one anybody will get (the ticket counter has run past the end) -- which for a small position,
         where a block sees one tile at most, saves the whole extra round */
#ifdef XF_NO_DRAIN_FUSION
      if (tid == 0) s_drain = have_tile ? 0u : 1u;
#else
      if (tid == 0) s_drain = (one_round || !have_tile || *(volatile u32*)&ctl->tile_ticket >= ntiles) ? 1u : 0u;
#endif
      __syncthreads();
      draining = s_drain != 0;
      /* ---- the round's batch list: queue q contributes its full batches (everything when draining) ---- */
      const u32* head = s_head[par2];
      u32 nb_total, nb_live; /* batches of the round; of which from the queues of incomplete prefixes */
      {
        /* each lane looks after queues lane, lane + 32, lane + 64; exclusive scan in queue order: all of j = 0
           first, then j = 1, then j = 2 (a group past the last queue is skipped: the branch is uniform) */
        u32 run = 0;
#pragma unroll
        for (int j = 0; j < 3; j++) {
          if ((u32)j * 32 > nq) break;
          const u32 q = (u32)j * 32 + lane;
          const u32 avail = (q < nq) ? s_tail[q] - head[q] : 0;
          const u32 c = draining ? (avail + 31) >> 5 : avail >> 5;
          u32 inc = c;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const u32 v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (u32)o) inc += v;
          }
          if (q <= nq) s_cum[q] = run + inc - c; /* every warp writes the same values; q == nq: the sentinel */
          run += __shfl_sync(0xffffffffu, inc, 31);
        }
        nb_total = run;
        if (lane == 0 && nq >= 96) s_cum[96] = run; /* nq == XF_MAXQ: the sentinel is not any lane's queue */
        __syncwarp();
        /* the queues of incomplete prefixes come first: their batches are the ones before queue nsq */
        nb_live = s_cum[nsq];
      }
      /* room for every pair of the round: a live batch b owns node slots [32 b, 32 b + 32) of the round's range,
         a done batch expression slots likewise; only the last batch of a queue can be partial (draining) */
      const u32 slots_live = nb_live * 32u, slots_done = (nb_total - nb_live) * 32u;
      if (tid == 0) {
        const u32 nbase = slots_live ? atomicAdd(&ctl->node_count, slots_live) : 0;
        const u32 dbase = slots_done ? atomicAdd(&ctl->done_count, slots_done) : 0;
        u32 ovf = 0;
        if ((u64)nbase + slots_live > cap_nodes) ovf |= OVF_NODES;
        if ((u64)dbase + slots_done > cap_done) ovf |= OVF_DONE;
        if (ovf) atomicOr(&ctl->overflow, ovf);
        s_node_base = nbase; s_done_base = dbase;
        __threadfence_block();
        s_alloc_tag = (par2 << 30) + tile + 1;
      }
      if (tid < XF_MAXQ) { /* next round's heads */
        const u32 q = tid;
        const u32 avail = (q < nq) ? s_tail[q] - head[q] : 0;
        s_head[par2 ^ 1][q] = head[q] + (draining ? avail : (avail & ~31u));
      }
      /* ---- phase C ---- */
      while (true) {
        u32 b = 0;
        if (lane == 0) b = atomicAdd(&s_next[par2], 1u);
        b = __shfl_sync(0xffffffffu, b, 0);
        if (b >= nb_total) break;
        /* from the END of the list: the queues are in slot order and the heavy operators of a class (^ v L A, the
           trigonometric functions) are its last slots, so they start first and the constants fill the tail */
        b = nb_total - 1u - b;
        /* queue of batch b: the last q with s_cum[q] <= b (s_cum is non-decreasing, s_cum[0] = 0): the lanes look at
           three queues each and count */
        u32 q = 0;
#pragma unroll
        for (int j = 0; j < 3; j++) {
          const u32 qq = (u32)j * 32 + lane;
          q += __popc(__ballot_sync(0xffffffffu, qq < nq && s_cum[qq] <= b));
        }
        q -= 1u;
        const u32 j0 = (b - s_cum[q]) * 32u + lane;      /* index in the queue, from its head */
        const u32 avail = s_tail[q] - head[q];
        const bool act = j0 < avail;
        const bool complete = q >= nsq;
        const u32 qs = complete ? q - nsq : q;
        const int cls = qs < qoff1 ? 0 : qs < qoff2 ? 1 : 2;
        const u32 k = qs - (cls == 0 ? 0u : cls == 1 ? qoff1 : qoff2);
        const int g = P.trial[cls][k];
        if (lane == 0) atomicAdd(&s_opcnt[g], min(32u, avail - (b - s_cum[q]) * 32u)); /* exec() calls per symbol */
        int oc = OUT_NONE;
        Node child;
        Node par;
        if (act) {
          const u32 p = s_ent[q * XF_CAP + (head[q] + j0) % XF_CAP];
          par = nodes[p];
          oc = expand_exec(P, cls, g, complete, nodes, par, p, using_x, child, track_m);
          if (oc == OUT_DEAD) my_dead++;
          if (track_m) {
            if (oc == OUT_DEAD) { if (par.der >= p_min) push_tail(D.tail, cap_tail, ctl, -1.0 / 0.0, par.der, -1); }
            else if ((oc == OUT_LIVE || oc == OUT_DONE) && child.der >= p_min) /* one dead-end where the bound cuts in */
              push_tail(D.tail, cap_tail, ctl, par.der, child.der, 1);
          }
        }
        while (s_alloc_tag != (par2 << 30) + tile + 1) {} /* set long ago, as a rule */
        if (!complete) {
          const u32 pos = s_node_base + b * 32u + lane;
          if (pos < cap_nodes) {
            if (oc == OUT_LIVE) {
              expand_finish(P, par, g, child);
              nodes[pos] = child;
            } else { /* hole (failed pair, or the tail of a partial batch): only the word that holds the form */
              ((uint4*)(nodes + pos))[2] = make_uint4(RIES_NO_NODE, FORM_HOLE, 0u, 0u);
            }
          }
        } else {
          const u32 pos = s_done_base + (b - nb_live) * 32u + lane;
          if (pos < cap_done) {
            if (oc == OUT_DONE) {
              expand_finish(P, par, g, child);
              Row r; r.val = child.val; r.der = child.der; r.sym = child.sym;
              done_rows[pos] = r;
              done_forms[pos] = child.form;
            } else {
              done_forms[pos] = (uint16_t)FORM_HOLE;
            }
          }
          if (oc == OUT_DONE) my_gen++;
        }
      }
    }
  }
  /* ---- level finish by the last block to get here (replaces a separate launch); PERSIST: the others wait for it ---- */
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    s_last = (atomicAdd(&ctl->blocks_done, 1u) == gridDim.x - 1) ? 1u : 0u;
    if (s_last) {
      __threadfence();
      const u32 ovf = *(volatile u32*)&ctl->overflow;
      const u32 alloc = *(volatile u32*)&ctl->node_count;
      ctl->lvl_begin = lvl_end;
      /* on overflow stop expanding: the host grows the arenas and reruns the pass */
      ctl->lvl_end = (ovf & (OVF_NODES | OVF_DONE)) ? lvl_end : alloc;
      ctl->tile_ticket = 0;
      ctl->blocks_done = 0;
      if (PERSIST) { __threadfence(); *(volatile u32*)&ctl->bar_gen = bar_gen0 + lvl + 1u; }
    } else if (PERSIST) {
      while (*(volatile u32*)&ctl->bar_gen != bar_gen0 + lvl + 1u) {}
      __threadfence();
    }
  }
  __syncthreads();
  } /* positions */
  /* block-level counters -> one atomic each per block */
  for (int o = 16; o > 0; o >>= 1) {
    my_dead += __shfl_xor_sync(0xffffffffu, my_dead, o);
    my_gen += __shfl_xor_sync(0xffffffffu, my_gen, o);
  }
  if (lane == 0) { if (my_dead) atomicAdd(&s_dead, my_dead); if (my_gen) atomicAdd(&s_gen, my_gen); }
  __syncthreads();
  if (tid == 0 && s_dead) atomicAdd(&ctl->dead, (u64)s_dead);
  if (tid == 0 && s_gen) atomicAdd(&ctl->gen, (u64)s_gen);
  if (tid < RIES_MAX_SYMS && s_opcnt[tid]) atomicAdd(&D.opcnt[(using_x ? RIES_MAX_SYMS : 0) + tid], (u64)s_opcnt[tid]);
}

__global__ void __launch_bounds__(XF_THREADS, XF_MIN_BLOCKS)
k_expand(const DevPtrs* __restrict__ Dp, const PassArgs* __restrict__ Ap, PassCtl* ctl) {
  expand_body<false>(Dp, Ap, ctl);
}

__global__ void __launch_bounds__(XF_THREADS, XF_MIN_BLOCKS)
k_expand_all(const DevPtrs* __restrict__ Dp, const PassArgs* __restrict__ Ap, PassCtl* ctl) {
  expand_body<true>(Dp, Ap, ctl);
}
