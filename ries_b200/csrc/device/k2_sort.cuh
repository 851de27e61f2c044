// Sort keys + pre-dedupe, and K2: onesweep LSD radix sort on the fp64 value.
// Part of the single translation unit engine.cu (included inside namespace ries).
#pragma once

/* ------------------------------------------------ sort keys + pre-dedupe ------------------- */
__device__ __forceinline__ bool pass_ok(const PassCtl* C) { return (C->overflow & (OVF_NODES | OVF_DONE | OVF_KEPT)) == 0; }

__device__ __forceinline__ u64 key_of(double v) { /* order-preserving map double -> u64 */
  u64 b = (u64)__double_as_longlong(v);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

/* Every filled slot of done[] gets its sort key -- unless the value is already in the table ON THE SAME SIDE:
   bt_insert finds the equal node and check_exact_match returns at once for a same-side pair (ries.c:8648-8661,
   8550-8593), so such an expression changes nothing but the "expressions" counter, which K1 has already
   counted.  At -l6 that is 70-84 % of what the big passes generate, found here with ONE 32-byte probe of the
   side's hash set instead of eight digit passes and a table search.  The survivors are compacted (block
   aggregate, one atomic per 1024 slots; their order is free: insertion order is recovered from (form, symbols)
   wherever it matters) together with the histograms of all eight key digits for the radix sort below. */
#define KK_THREADS 256
#define KK_ITEMS 4
#define KK_CHUNK (KK_THREADS * KK_ITEMS)
__global__ void __launch_bounds__(KK_THREADS)
k_keys(const DevPtrs* Dp, const PassArgs* A, PassCtl* C) {
  __shared__ u32 s_h[8][256];
  __shared__ u32 s_cnt[KK_ITEMS * (KK_THREADS / 32)];
  __shared__ u32 s_base;
  if (!A->gathered) stage_stamp(C, 1); /* a sharded pass: k_keys_send has stamped the end of K1 */
  if (!pass_ok(C)) return;
  const u32 n = C->done_count;
  if ((u64)blockIdx.x * KK_CHUNK >= n) return;
  const DevPtrs& D = *Dp;
  const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const Row* __restrict__ done = A->gathered ? D.gdone : D.done;
  const uint16_t* __restrict__ done_form = A->gathered ? D.gdone_form : D.done_form;
  const int side = A->using_x ? 1 : 0;
  const u64* __restrict__ hs = Dp->hset[side];
  const u32 hlog = Dp->hset_log2[side];
  /* gathered rows of a sharded pass: slot j of rank r is filled iff j < that rank's count */
  const bool gathered = A->gathered != 0;
  const u32 bound = gathered ? (u32)A->shard_bound : 1u;
  const SlotHdr* __restrict__ ghdr = D.ghdr;
  const bool ranged = gathered && A->range_split != 0;
  const u64 klo = C->klo, khi = C->khi;
  u64* keys = D.keys[0];
  u32* idx = D.idx[0];
  ASSUME_GLOBAL(done); ASSUME_GLOBAL(done_form); ASSUME_GLOBAL(hs); ASSUME_GLOBAL(keys); ASSUME_GLOBAL(idx); ASSUME_GLOBAL(ghdr);
  for (int d = 0; d < 8; d++) s_h[d][tid] = 0;
  __syncthreads();
  for (u32 base = blockIdx.x * KK_CHUNK; base < n; base += gridDim.x * KK_CHUNK) {
    u64 k[KK_ITEMS];
    u32 flags = 0;
#pragma unroll
    for (int r = 0; r < KK_ITEMS; r++) {
      const u32 i = base + r * KK_THREADS + tid;
      bool keep = false;
      k[r] = 0;
      if (i < n) {
        const bool filled = gathered ? (i % bound) < ghdr[i / bound].n_kept : done_form[i] != FORM_HOLE;
        if (filled) {
          k[r] = key_of(done[i].val);
          keep = (!ranged || (k[r] >= klo && k[r] < khi)) && !hset_contains(hs, hlog, k[r]);
        }
      }
      const u32 b = __ballot_sync(0xffffffffu, keep);
      if (lane == 0) s_cnt[r * (KK_THREADS / 32) + warp] = __popc(b);
      if (keep) flags |= (1u << r) | (__popc(b & ((1u << lane) - 1)) << (8 + 6 * r)); /* rank in the warp: 5 bits used of 6 */
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
      if (lane == 31) s_base = inc ? atomicAdd(&C->n_keys, inc) : 0;
    }
    __syncthreads();
    const u32 cbase = s_base;
#pragma unroll
    for (int r = 0; r < KK_ITEMS; r++) {
      if ((flags >> r) & 1u) {
        const u32 pos = cbase + s_cnt[r * (KK_THREADS / 32) + warp] + ((flags >> (8 + 6 * r)) & 63u);
        keys[pos] = k[r];
        idx[pos] = base + r * KK_THREADS + tid;
#pragma unroll
        for (int d = 0; d < 8; d++) atomicAdd(&s_h[d][(u32)(k[r] >> (8 * d)) & 255u], 1u);
      }
    }
    __syncthreads();
  }
  SortCtl* ctl = D.sortctl;
  for (int d = 0; d < 8; d++) {
    u32 c = s_h[d][tid];
    if (c) atomicAdd(&ctl->ghist[d][tid], c);
  }
}

/* k_keys for the consumer of a shared RHS pass: the rows come from the batch's enumeration; a row exists for this
   target iff M < its overflow bound (that many are "generated"); what the hash probe lets through is copied into
   this engine's done[] with `der` back to 0 (der == 0 <=> RHS everywhere downstream). */
__global__ void __launch_bounds__(KK_THREADS)
k_keys_shared(const DevPtrs* Dp, const PassArgs* A, PassCtl* C, const SharedView* Vp) {
  const SharedView V = *Vp;
  __shared__ u32 s_h[8][256];
  __shared__ u32 s_cnt[KK_ITEMS * (KK_THREADS / 32)];
  __shared__ u32 s_base, s_gen;
  stage_stamp(C, 1);
  const u32 n = V.n_slots;
  if ((u64)blockIdx.x * KK_CHUNK >= n) return;
  const DevPtrs& D = *Dp;
  const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u64* __restrict__ hs = Dp->hset[0];
  const u32 hlog = Dp->hset_log2[0];
  const u32 cap_done = A->cap_done;
  u64* keys = D.keys[0];
  u32* idx = D.idx[0];
  for (int d = 0; d < 8; d++) s_h[d][tid] = 0;
  if (tid == 0) s_gen = 0;
  __syncthreads();
  u32 my_gen = 0;
  for (u32 base = blockIdx.x * KK_CHUNK; base < n; base += gridDim.x * KK_CHUNK) {
    u64 k[KK_ITEMS];
    u32 flags = 0;
#pragma unroll
    for (int r = 0; r < KK_ITEMS; r++) {
      const u32 i = base + r * KK_THREADS + tid;
      bool keep = false;
      k[r] = 0;
      if (i < n && V.forms[i] != FORM_HOLE) {
        const Row row = V.rows[i];
        if (row.der < V.p_t) {
          my_gen++;
          k[r] = key_of(row.val);
          keep = !hset_contains(hs, hlog, k[r]);
        }
      }
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
      if (lane == 31) {
        const u32 b0 = inc ? atomicAdd(&C->n_keys, inc) : 0;
        if (inc && (u64)b0 + inc > cap_done) { atomicOr(&C->overflow, OVF_DONE); atomicMax(&C->done_count, b0 + inc); }
        s_base = b0;
      }
    }
    __syncthreads();
    const u32 cbase = s_base;
#pragma unroll
    for (int r = 0; r < KK_ITEMS; r++) {
      if ((flags >> r) & 1u) {
        const u32 pos = cbase + s_cnt[r * (KK_THREADS / 32) + warp] + ((flags >> (8 + 6 * r)) & 63u);
        if (pos < cap_done) {
          const u32 i = base + r * KK_THREADS + tid;
          Row row = V.rows[i];
          row.der = 0.0;
          D.done[pos] = row;
          D.done_form[pos] = V.forms[i];
          keys[pos] = k[r];
          idx[pos] = pos;
#pragma unroll
          for (int d = 0; d < 8; d++) atomicAdd(&s_h[d][(u32)(k[r] >> (8 * d)) & 255u], 1u);
        }
      }
    }
    __syncthreads();
  }
  for (int o = 16; o > 0; o >>= 1) my_gen += __shfl_xor_sync(0xffffffffu, my_gen, o);
  if (lane == 0 && my_gen) atomicAdd(&s_gen, my_gen);
  __syncthreads();
  if (tid == 0 && s_gen) atomicAdd((unsigned long long*)&C->gen, (unsigned long long)s_gen);
  SortCtl* ctl = D.sortctl;
  for (int d = 0; d < 8; d++) {
    u32 c = s_h[d][tid];
    if (c) atomicAdd(&ctl->ghist[d][tid], c);
  }
}

/* ---------------------------------------------------------------- K2: radix sort --------- */
/* Stable LSD radix sort of (u64 key, u32 payload), 8 digits of 8 bits, "onesweep" style: the histograms
   of all eight digits come from k_keys (which writes the keys); each digit pass is then a single kernel in
   which every 4096-key chunk gets its per-digit output offsets by decoupled look-back over the chunks before
   it (chunks are taken in ticket order, so the sort is stable -- which nothing relies on any more: equal
   values are ranked by their insertion-order keys in k_classify_compact). */
#define RS_THREADS 256
#define RS_WARPS (RS_THREADS / 32)
#ifndef RS_ITEMS
#define RS_ITEMS 16
#endif
#ifndef RS_MIN_BLOCKS
#define RS_MIN_BLOCKS 2
#endif
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
__global__ void __launch_bounds__(RS_THREADS, RS_MIN_BLOCKS)
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
  const u32 n = C->n_keys;
  const u32 nchunks = (n + RS_CHUNK - 1) / RS_CHUNK;
  if (blockIdx.x >= nchunks) return;
  const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u32 lt = (1u << lane) - 1;
#ifndef RS_NO_SINGLE
  if (nchunks == 1) {
    /* A pass of at most one chunk (most passes of a search are): this block sorts it through all eight digits in
       the launch of digit 0 -- the keys stay in registers and shared memory from the first load to the last
       store, there is no ticket, no look-back, no global histogram -- and the other seven launches find nothing
       to do.  Same ranking as below, so the order inside a bin is again the input order. */
    if (pass != 0) return;
    u64* keys_io = Dp->keys[0];
    u32* idx_io = Dp->idx[0];
    ASSUME_GLOBAL(keys_io); ASSUME_GLOBAL(idx_io);
    const u32 wbase_i = warp * (RS_ITEMS * 32) + lane;
    u64 key[RS_ITEMS];
    u32 pay[RS_ITEMS];
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
      const u32 li = wbase_i + r * 32;
      key[r] = li < n ? keys_io[li] : 0;
      pay[r] = li < n ? idx_io[li] : 0;
    }
    u32* wh = s_whist[warp];
#pragma unroll 1
    for (int shift = 0; shift < 64; shift += 8) {
      for (u32 j = tid; j < RS_WARPS * 256; j += RS_THREADS) (&s_whist[0][0])[j] = 0;
      __syncthreads(); /* also: the reload below is over before the exchange buffers are written again */
      u32 rk[RS_ITEMS];
#pragma unroll
      for (int r = 0; r < RS_ITEMS; r++) {
        const bool valid = (wbase_i + r * 32) < n;
        const u32 d = (u32)(key[r] >> shift) & 255u;
        const u32 peers = __match_any_sync(0xffffffffu, valid ? d : (0x100u | lane));
        const u32 prev = valid ? wh[d] : 0;
        __syncwarp();
        if (valid && (peers & lt) == 0) wh[d] = prev + __popc(peers);
        __syncwarp();
        rk[r] = prev + __popc(peers & lt);
      }
      __syncthreads();
      u32 cnt = 0;
#pragma unroll
      for (int w = 0; w < RS_WARPS; w++) { u32 c = s_whist[w][tid]; s_whist[w][tid] = cnt; cnt += c; }
      u32 inc = cnt;
      for (int o = 1; o < 32; o <<= 1) {
        u32 v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (u32)o) inc += v;
      }
      if (lane == 31) s_wtot[warp] = inc;
      __syncthreads();
      {
        u32 wb = 0;
        for (u32 w = 0; w < warp; w++) wb += s_wtot[w];
        s_binex[tid] = wb + inc - cnt;
      }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < RS_ITEMS; r++) {
        if ((wbase_i + r * 32) < n) {
          const u32 d = (u32)(key[r] >> shift) & 255u;
          const u32 lp = s_binex[d] + wh[d] + rk[r];
          s_keys[lp] = key[r];
          s_pays[lp] = pay[r];
        }
      }
      __syncthreads();
      if (shift < 56) {
#pragma unroll
        for (int r = 0; r < RS_ITEMS; r++) {
          const u32 li = wbase_i + r * 32;
          if (li < n) { key[r] = s_keys[li]; pay[r] = s_pays[li]; }
        }
      }
    }
    for (u32 j = tid; j < n; j += RS_THREADS) { keys_io[j] = s_keys[j]; idx_io[j] = s_pays[j]; }
    return;
  }
#endif
  const int shift = pass * 8;
  const int src = pass & 1;
  const u64* __restrict__ keys_in = Dp->keys[src];
  const u32* __restrict__ idx_in = Dp->idx[src];
  u64* keys_out = Dp->keys[src ^ 1];
  u32* idx_out = Dp->idx[src ^ 1];
  SortCtl* ctl = Dp->sortctl;
  volatile u64* state = (volatile u64*)Dp->rs_state;
  ASSUME_GLOBAL(keys_in); ASSUME_GLOBAL(idx_in); ASSUME_GLOBAL(keys_out); ASSUME_GLOBAL(idx_out); ASSUME_GLOBAL(ctl);
  ASSUME_GLOBAL((const u64*)state);
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
