// One search over several GPUs: the sharded front of a pass and the exchange bookkeeping.
// Part of the single translation unit engine.cu (included inside namespace ries).
#pragma once

/* A sharded pass (passes big enough to pay for an exchange):

     every rank   K1 on its share of the forms (dealt round-robin; insertion order does not depend on who
                  enumerates what: it is a function of (form, symbols), state.cuh)
                  k_keys_send: what the table already holds on the same side is dropped by the hash probe
                  (70-84 % of a big pass); the rest -- row + form index -- is compacted into this rank's SEND
                  SLOT, a buffer of `bound` rows sized by the host from the previous pass of the side
                  k_send_hdr: the slot's header (count, this rank's counters and overflow flags)
     exchange     three fixed-size all-gathers (headers, rows, form indices): rank r's slot lands at
                  gdone[r * bound ...] on every rank.  Fixed sizes mean NO HOST ROUND TRIP between the counts
                  and the payload: the whole pass is enqueued at once and the host polls for its end, exactly
                  as on one GPU.  The padding costs ~30 % more bytes than the rows; a slot that turns out too
                  small sets an overflow flag that every rank sees in the gathered headers, and the pass is
                  rerun with a larger bound.
     every rank   k_gather_hdr: counters summed, flags or-ed, into the pass's control block;
                  then the ordinary pipeline on the gathered rows (k_keys in `gathered` mode reads the slots),
                  replicated: sort, dedupe, match, merge.  Every rank holds the same table and the same hash
                  sets, takes the same decisions and prints the same report. */
/* Survivors of the hash probe, compacted into the send slot (any order). */
__global__ void __launch_bounds__(KK_THREADS)
k_keys_send(const DevPtrs* Dp, const PassArgs* A, PassCtl* C) {
  __shared__ u32 s_cnt[KK_ITEMS * (KK_THREADS / 32)];
  __shared__ u32 s_base;
  stage_stamp(C, 1);
  if (!pass_ok(C)) return;
  const u32 n = C->done_count;
  if ((u64)blockIdx.x * KK_CHUNK >= n) return;
  const DevPtrs& D = *Dp;
  const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const Row* __restrict__ done = D.done;
  const uint16_t* __restrict__ done_form = D.done_form;
  const int side = A->using_x ? 1 : 0;
  const u64* __restrict__ hs = Dp->hset[side];
  const u32 hlog = Dp->hset_log2[side];
  const u32 bound = (u32)A->shard_bound;
  Row* ksend = D.ksend;
  uint16_t* ksend_form = D.ksend_form;
  for (u32 base = blockIdx.x * KK_CHUNK; base < n; base += gridDim.x * KK_CHUNK) {
    u32 flags = 0;
#pragma unroll
    for (int r = 0; r < KK_ITEMS; r++) {
      const u32 i = base + r * KK_THREADS + tid;
      bool keep = false;
      if (i < n && done_form[i] != FORM_HOLE) keep = !hset_contains(hs, hlog, key_of(done[i].val));
      const u32 b = __ballot_sync(0xffffffffu, keep);
      if (lane == 0) s_cnt[r * (KK_THREADS / 32) + warp] = __popc(b);
      if (keep) flags |= (1u << r) | (__popc(b & ((1u << lane) - 1)) << (8 + 6 * r));
    }
    __syncthreads();
    if (warp == 0) {
      const u32 c = s_cnt[lane];
      u32 inc = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u32 v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (u32)o) inc += v;
      }
      s_cnt[lane] = inc - c;
      if (lane == 31) s_base = inc ? atomicAdd(&C->n_kept, inc) : 0;
    }
    __syncthreads();
    const u32 cbase = s_base;
#pragma unroll
    for (int r = 0; r < KK_ITEMS; r++) {
      if ((flags >> r) & 1u) {
        const u32 pos = cbase + s_cnt[r * (KK_THREADS / 32) + warp] + ((flags >> (8 + 6 * r)) & 63u);
        const u32 i = base + r * KK_THREADS + tid;
        if (pos < bound) {
          ksend[pos] = done[i];
          ksend_form[pos] = done_form[i];
        }
      }
    }
    __syncthreads();
  }
}

__global__ void k_send_hdr(const DevPtrs* Dp, const PassArgs* A, PassCtl* C) {
  if (threadIdx.x != 0) return;
  SlotHdr h;
  h.overflow = C->overflow;
  h.n_kept = C->n_kept;
  if (h.n_kept > (u32)A->shard_bound) h.overflow |= OVF_KEPT;
  h.node_count = C->node_count; h.done_slots = C->done_count;
  h.dead = C->dead; h.gen = C->gen;
  for (int i = 0; i < 8; i++) h.pad[i] = 0;
  *Dp->shdr = h;
}

/* After the exchange: the ranks' counters into the control block, as if one K1 had produced the gathered rows;
   fresh look-back tags, histograms, tickets for the replicated half of the pass. */
__global__ void k_gather_hdr(const DevPtrs* D, const PassArgs* A, PassCtl* C, TableState* T) {
  u32* sc = (u32*)D->sortctl;
  for (u32 i = threadIdx.x; i < sizeof(SortCtl) / 4; i += blockDim.x) sc[i] = 0;
  if (threadIdx.x == 0) {
    const int world = A->shard_world;
    const SlotHdr* g = D->ghdr;
    u32 ovf = 0, max_nodes = 0, max_done = 0, max_kept = 0;
    u64 nodes = 0, dead = 0, gen = 0;
    for (int r = 0; r < world; r++) {
      ovf |= g[r].overflow;
      nodes += g[r].node_count; dead += g[r].dead; gen += g[r].gen;
      max_nodes = max(max_nodes, g[r].node_count); max_done = max(max_done, g[r].done_slots); max_kept = max(max_kept, g[r].n_kept);
    }
    C->overflow = ovf;
    C->node_count = (nodes > 0xffffffffull) ? 0xffffffffu : (u32)nodes;
    C->done_count = (u32)world * (u32)A->shard_bound; /* slots of the gathered array */
    C->dead = dead; C->gen = gen;
    C->max_nodes = max_nodes; C->max_done = max_done; C->n_kept = max_kept;
    C->n_keys = 0; C->chunk_ticket = 0; C->aux_ticket = 0; C->tile_ticket = 0; C->blocks_done = 0;
    C->n_events = 0; C->n_uniq = 0;
    C->klo = 0; C->khi = ~0ull;
    if (A->range_split) {
      /* this rank's value range: between the table's values at its quantiles, so that every gap of the table --
         and with it every neighbour test and every run of equal values -- belongs to exactly one rank */
      const u32 nt = T->nt;
      const double* tv = D->tval[T->cur];
      const u32 r = (u32)A->shard_rank, w = (u32)world;
      if (r > 0) C->klo = key_of(tv[(u64)nt * r / w]);
      if (r + 1 < w) C->khi = key_of(tv[(u64)nt * (r + 1) / w]);
    }
    u32 e = (T->epoch + 16) & 0x3ffffff0u;
    if (e == 0) e = 16;
    T->epoch = e;
  }
}

/* ---- second half of a sharded pass, split by value range -------------------------------------------------- */
/* The table stays replicated (its per-pass merge moves 12 bytes per row), but sorting, deduping and matching the
   gathered rows is split: rank r takes the rows whose value lies between the table's values at its quantiles
   (k_keys' range filter), sorts and classifies them, runs the neighbour tests of its new rows (k_match_merge_new,
   mode 1) -- and then the ranks exchange what they found, which is small: (index into the gathered rows, gap) per new
   row, and the candidates.  k_gather2 strings the ranges together (they are in value order by rank), and the merge
   (mode 2, k_merge_old), the refinement and the end of the pass run replicated on identical data as before. */
__global__ void k_send2_hdr(const DevPtrs* Dp, const PassArgs* A, PassCtl* C) {
  if (threadIdx.x != 0) return;
  SlotHdr h;
  h.overflow = C->overflow;
  h.n_kept = C->n_uniq;
  h.node_count = C->n_events;
  if (C->n_uniq > (u32)A->range_bound) h.overflow |= OVF_KEPT;
  if (C->n_events > (u32)A->range_ebound) h.overflow |= OVF_EVENTS;
  h.done_slots = 0; h.dead = 0; h.gen = 0;
  for (int i = 0; i < 8; i++) h.pad[i] = 0;
  *Dp->shdr = h;
}

__global__ void __launch_bounds__(256)
k_gather2(const DevPtrs* Dp, const PassArgs* A, PassCtl* C) {
  __shared__ u32 s_off[RIES_MAX_RANKS + 1], s_eoff[RIES_MAX_RANKS + 1];
  const DevPtrs& D = *Dp;
  const int world = A->shard_world;
  const SlotHdr* g = D.ghdr2;
  if (threadIdx.x == 0) {
    u32 off = 0, eoff = 0;
    for (int r = 0; r < world; r++) { s_off[r] = off; s_eoff[r] = eoff; off += g[r].n_kept; eoff += g[r].node_count; }
    s_off[world] = off; s_eoff[world] = eoff;
  }
  __syncthreads();
  u32 ovf = 0;
  for (int r = 0; r < world; r++) ovf |= g[r].overflow;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    C->overflow |= ovf;
    C->n_uniq = s_off[world];
    C->n_events = s_eoff[world];
    if (s_eoff[world] > A->cap_events) C->overflow |= OVF_EVENTS;
    u32 mk = 0, me = 0;
    for (int r = 0; r < world; r++) { mk = max(mk, g[r].n_kept); me = max(me, g[r].node_count); }
    C->max_new = mk;
    C->max_ev = me;
  }
  if (ovf) return;
  const u32 bound = (u32)A->range_bound, eb = (u32)A->range_ebound;
  const Row* __restrict__ gd = D.gdone;
  const uint16_t* __restrict__ gf = D.gdone_form;
  for (int r = 0; r < world; r++) {
    const u32 n = g[r].n_kept, o = s_off[r];
    for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
      const u32 gi = D.g2_idx[(size_t)r * bound + j];
      D.uniq_seq[o + j] = gi;
      D.uniq_gap[o + j] = D.g2_gap[(size_t)r * bound + j];
      SeqKey sk;
      const Sym128 sy = gd[gi].sym;
      sk.lo = sy.lo; sk.hi = sy.hi; sk.form = gf[gi]; sk.pad = 0;
      D.uniq_sk[o + j] = sk;
    }
    const u32 ne = g[r].node_count, eo = s_eoff[r];
    if (s_eoff[world] <= A->cap_events)
      for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j < ne; j += gridDim.x * blockDim.x)
        D.events[eo + j] = D.g2_events[(size_t)r * eb + j];
  }
}
