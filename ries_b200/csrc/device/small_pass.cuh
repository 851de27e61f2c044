// Tiny passes: regroup + sort + classify + match + merge-new in one block (called by k_tiny_front).
// Part of the single translation unit engine.cu (included inside namespace ries).
#pragma once

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

/* m <= SP_THREADS keys: the final position of a key = the number of keys that sort before it (smaller, or
   equal and earlier: stable).  `tpk` threads share a key (a power of two, as many as the block has to
   spare, all in one warp) and each counts over every tpk-th key; the partial counts meet by shuffle.  m / tpk
   broadcast reads of shared memory per thread and two barriers instead of eight digit passes of five. */
__device__ __forceinline__ void small_rank_sort(u64* s_keys, u32* s_pays, u32 m) {
  const u32 tid = threadIdx.x;
  u32 tpk = 1;
  while (tpk < 32 && m * tpk * 2 <= SP_THREADS) tpk <<= 1;
  const u32 ki = tid / tpk, sub = tid % tpk;
  u64 k = 0;
  u32 p = 0, rank = 0;
  if (ki < m) {
    k = s_keys[ki];
    p = s_pays[ki];
    for (u32 j = sub; j < m; j += tpk) {
      const u64 o = s_keys[j];
      rank += (u32)((o < k) || (o == k && j < ki));
    }
  }
  for (u32 o = 1; o < tpk; o <<= 1) rank += __shfl_xor_sync(0xffffffffu, rank, o);
  __syncthreads();
  if (ki < m && sub == 0) { s_keys[rank] = k; s_pays[rank] = p; }
  __syncthreads();
}

__device__ __forceinline__ void small_post_body(const DevPtrs* Dp, const PassArgs* A, PassCtl* C, const TableState* T) {
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
  /* (no __restrict__: in k_tiny_front these were written earlier in the same kernel) */
  const Row* done = A->gathered ? D.gdone : D.done;
  const uint16_t* done_form = A->gathered ? D.gdone_form : D.done_form;
  for (u32 i = tid; i < n; i += SP_THREADS) { s_keys[i] = done_form[i]; s_pays[i] = i; }
  __syncthreads();
  const bool by_rank = n <= SP_THREADS;
  if (by_rank) {
    small_rank_sort(s_keys, s_pays, n);
  } else {
    small_sort_pass(s_keys, s_pays, s_whist, s_binex, s_wtot, n, 0);
    if (A->nforms > 256) small_sort_pass(s_keys, s_pays, s_whist, s_binex, s_wtot, n, 8);
  }
  for (u32 j = tid; j < n; j += SP_THREADS) {
    const Row r = done[s_pays[j]];
    D.surv[j] = r;
    s_keys[j] = key_of(r.val);
    s_pays[j] = j;
  }
  __syncthreads();
  /* stable sort by value */
  if (by_rank) {
    small_rank_sort(s_keys, s_pays, n);
  } else {
    for (int pass = 0; pass < 8; pass++) small_sort_pass(s_keys, s_pays, s_whist, s_binex, s_wtot, n, pass * 8);
  }
  /* classify: new distinct values, EXACT events */
  const u32 cur = T->cur, nt = T->nt;
  const double* __restrict__ tval = D.tval[cur];
  const u32* __restrict__ tref = D.tref[cur];
  Row* heap = D.heap;
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
          const u32 ref = tref[p];
          const bool tree_is_lhs = (ref & TREF_LHS) != 0;
          if (using_x ? !tree_is_lhs : tree_is_lhs) {
            const Row t = heap[ref & ~TREF_LHS];
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
  u32* mref = D.tref[cur ^ 1];
  double* mval = D.tval[cur ^ 1];
  for (u32 u = tid; u < nu; u += SP_THREADS) {
    const u32 sidx = s_useq[u], g = s_ugap[u];
    const Row f = D.surv[sidx];
    heap[nt + u] = f;
    mval[u + g] = f.val;
    mref[u + g] = (nt + u) | (using_x ? TREF_LHS : 0u);
    hset_insert(Dp->hset[using_x ? 1 : 0], Dp->hset_log2[using_x ? 1 : 0], key_of(f.val)); /* later passes' pre-dedupe */
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
      const u32 oref = tref[side == 0 ? g - 1 : g];
      const bool other_is_lhs = (oref & TREF_LHS) != 0;
      if (using_x ? other_is_lhs : !other_is_lhs) continue;
      const Row o = heap[oref & ~TREF_LHS];
      const Row& L = using_x ? f : o;
      const Row& R = using_x ? o : f;
      double total = L.der - R.der;
      double score = fabs((R.val - L.val) / total);
      if (score > best_match0) continue; /* NaN passes, as in the reference's test (8270) */
      push_event(D.events, A->cap_events, C, sidx, side == 0 ? EV_PREV : EV_NEXT, L, R);
    }
  }
}

