// Classify/compact, K3 neighbour test + merge of the new rows.
// Part of the single translation unit engine.cu (included inside namespace ries).
#pragma once

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

/* For every sorted key: is it a new distinct value?  Of a run of equal new values the row the tree would have
   kept is the one inserted first (bt_insert, ries.c:8648-8661): the first element of the run walks the run and
   picks the smallest insertion-order key (SeqKey; runs of NEW values are short -- what the table already holds
   on this side never got here).  New rows are compacted (ticketed chunks + look-back keep them in value
   order) together with their gap = lower bound in the old table.  Duplicates of an opposite-side table row
   become EXACT events (check_exact_match's side test, ries.c:8550-8593).
   The survivors of a chunk are consecutive in value order, so their lower bounds all fall between the
   lower bounds of the chunk's first and last value: two warp searches find that window, the window of
   table keys is staged in shared memory with coalesced loads, and the per-survivor searches run there.
   Values come straight from the sorted keys (the key map is invertible): no gather from surv[]. */
#define CC_THREADS 256
#define CC_ITEMS 8
#define CC_CHUNK (CC_THREADS * CC_ITEMS)
#define CC_WIN 3072

/* The first 64 bits of an insertion-order key (form, then the first eight symbol codes, first position most
   significant): decides "who came first" between two expressions of a pass unless they share form and the first
   eight symbols, in which case the full keys are compared. */
__device__ __forceinline__ u64 seq_prefix(u32 form, u64 sym_lo) {
  u64 o = (u64)form << 48;
#pragma unroll
  for (int i = 0; i < 8; i++) o |= ((sym_lo >> (6 * i)) & 63ull) << (6 * (7 - i));
  return o;
}
__global__ void __launch_bounds__(CC_THREADS)
k_classify_compact(const DevPtrs* Dp, const PassArgs* A, PassCtl* C, const TableState* T) {
  __shared__ __align__(16) double s_winraw[CC_WIN + 4];
  __shared__ __align__(8) unsigned long long s_mbar;
  __shared__ u64 s_ord[CC_CHUNK]; /* seq_prefix of the chunk's keys that have an equal neighbour */
  __shared__ u32 s_cnt[CC_ITEMS * (CC_THREADS / 32)];
  __shared__ u32 s_chunk, s_base, s_lo, s_hi;
  stage_stamp(C, 2);
  if (!pass_ok(C)) return;
  const u32 n = C->n_keys;
  const u32 nchunks = (n + CC_CHUNK - 1) / CC_CHUNK;
  if (blockIdx.x >= nchunks) return;
  const DevPtrs& D = *Dp;
  const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u32 cur = T->cur, nt = T->nt;
  const double* __restrict__ tval = D.tval[cur];
  const u32* __restrict__ tref = D.tref[cur];
  const Row* __restrict__ heap = D.heap;
  const u64* __restrict__ keys = D.keys[0]; /* 8 digit passes: the sorted order ends in buffer 0 */
  const u32* __restrict__ idx = D.idx[0];
  const Row* __restrict__ surv = A->gathered ? D.gdone : D.done;
  const uint16_t* __restrict__ surv_form = A->gathered ? D.gdone_form : D.done_form;
  const int using_x = A->using_x, skip_exact = A->skip_exact;
  volatile u64* state = (volatile u64*)D.cc_state;
  if (threadIdx.x == 0) mbar_init(&s_mbar, 1);
  u32 mbar_phase = 0;
  __syncthreads();
  u32* const uniq_seq_o = D.uniq_seq;
  u32* const uniq_gap_o = D.uniq_gap;
  SeqKey* const uniq_sk_o = D.uniq_sk;
  ASSUME_GLOBAL(tval); ASSUME_GLOBAL(tref); ASSUME_GLOBAL(heap); ASSUME_GLOBAL(keys); ASSUME_GLOBAL(idx); ASSUME_GLOBAL(surv); ASSUME_GLOBAL(surv_form);
  ASSUME_GLOBAL(uniq_seq_o); ASSUME_GLOBAL(uniq_gap_o); ASSUME_GLOBAL(uniq_sk_o); ASSUME_GLOBAL((const u64*)state);
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
    /* the window goes to shared memory -- every S-th key of it when it is longer than the buffer (after the
       pre-dedupe the keys of a chunk are spread over a wider stretch of the table than they used to be): the
       search runs over the samples there and finishes among the S keys between two samples in global memory */
    const u32 S = (wload + CC_WIN - 1) / CC_WIN > 1 ? (wload + CC_WIN - 1) / CC_WIN : 1;
    const u32 ns = (wload + S - 1) / S;
    /* a contiguous window is fetched by ONE bulk asynchronous copy (16-byte granules: from the even index at or
       below lo), which runs under the fetch of the order keys below; a sampled window is gathered by the threads */
    const u32 lo_al = lo & ~1u;
    const double* s_win = s_winraw + (S == 1 ? lo - lo_al : 0u);
    if (S == 1) {
      if (tid == 0) bulk_g2s(s_winraw, tval + lo_al, (((wload + (lo - lo_al)) + 1u) & ~1u) * (u32)sizeof(double), &s_mbar);
    } else {
      for (u32 j = tid; j < ns; j += CC_THREADS) s_winraw[j] = tval[lo + j * S];
    }
    /* every key with an equal neighbour fetches the head of its own insertion-order key -- all at once, instead of
       the first key of a run fetching them one after the other while it walks the run */
#pragma unroll
    for (int r = 0; r < CC_ITEMS; r++) {
      const u32 i = base + r * CC_THREADS + tid;
      if (i < n) {
        const u64 k = keys[i];
        if ((i > 0 && keys[i - 1] == k) || (i + 1 < n && keys[i + 1] == k)) {
          const u32 s = idx[i];
          s_ord[r * CC_THREADS + tid] = seq_prefix(surv_form[s], surv[s].sym.lo);
        }
      }
    }
    if (S == 1) { mbar_wait(&s_mbar, mbar_phase); mbar_phase ^= 1u; }
    __syncthreads();
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
        u32 a = 0, b = (S == 1) ? wn : ns;
        while (a < b) {
          u32 mid = (a + b) >> 1;
          if (s_win[mid] < v) a = mid + 1; else b = mid;
        }
        bool in_table;
        if (S == 1) {
          in_table = (lo + a < nt) && (s_win[a] == v);
        } else { /* sample a is the first one >= v: the lower bound lies in ((a-1) S, a S] */
          u32 l2 = a ? (a - 1) * S + 1 : 0, h2 = min(a * S, wn);
          while (l2 < h2) {
            u32 mid = (l2 + h2) >> 1;
            if (tval[lo + mid] < v) l2 = mid + 1; else h2 = mid;
          }
          a = l2;
          in_table = (lo + a < nt) && (tval[lo + a] == v);
        }
        const u32 p = lo + a;
        pos[r] = p;
        if (in_table) {
          if (!skip_exact) {
            const u32 ref = tref[p];
            const bool tree_is_lhs = (ref & TREF_LHS) != 0;
            if (using_x ? !tree_is_lhs : tree_is_lhs) {
              const Row t = heap[ref & ~TREF_LHS];
              Row f = surv[s];
              if (using_x) push_event(D.events, A->cap_events, C, surv_form[s], EV_EXACT, f, t);
              else push_event(D.events, A->cap_events, C, surv_form[s], EV_EXACT, t, f);
            }
          }
        } else if (!dup_prev) {
          fl = true;
          /* the run of this value: the expression inserted first wins */
          if (i + 1 < n && keys[i + 1] == k) {
            u32 w = s;
            u64 wo = s_ord[i - base];
            const u32 cend = base + CC_CHUNK;
            for (u32 j = i + 1; j < n && keys[j] == k; j++) {
              const u32 s2 = idx[j];
              const u64 o2 = (j < cend) ? s_ord[j - base] : seq_prefix(surv_form[s2], surv[s2].sym.lo);
              bool before = o2 < wo;
              if (o2 == wo) before = seq_before(surv_form[s2], surv[s2].sym, surv_form[w], surv[w].sym);
              if (before) { w = s2; wo = o2; }
            }
            sq[r] = w;
          }
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
        uniq_seq_o[u] = sq[r];
        uniq_gap_o[u] = pos[r];
        SeqKey sk;
        const Sym128 sy = surv[sq[r]].sym;
        sk.lo = sy.lo; sk.hi = sy.hi; sk.form = surv_form[sq[r]]; sk.pad = 0;
        uniq_sk_o[u] = sk;
      }
    }
    __syncthreads();
  }
}

/* K3 + merge of the new rows.  A new row's predecessor AT ITS INSERTION TIME is the old-table row below
   its gap unless a same-pass row inserted earlier already sits between them (that one is same-side, so no test
   happens: check_sides, ries.c:8496-8538).  Likewise for the successor.  "Inserted earlier" = smaller
   insertion-order key (uniq_sk).  The row is read once: it is appended to the heap of table rows (rows never
   move; the sorted order lives in the (value, reference) arrays), its value and reference go to their place in
   the merged order, its key into the side's hash set, and it into the neighbour tests -- which fetch a
   neighbour's row only when its reference says it is of the other side. */
/* mode 0: both halves (every pass but the range-split ones); mode 1: the neighbour tests only, on this rank's
   value range; mode 2: the merge only, on the new rows of all ranges after the second exchange (shard.cuh) */
__global__ void k_match_merge_new(const DevPtrs* Dp, const PassArgs* A, PassCtl* C, const TableState* T, int mode) {
  if (!pass_ok(C)) return;
  const u32 nu = C->n_uniq;
  const DevPtrs& D = *Dp;
  const u32 nt = T->nt;
  const u32* __restrict__ tref = D.tref[T->cur];
  u32* mref = D.tref[T->cur ^ 1];
  Row* heap = D.heap;
  double* mval = D.tval[T->cur ^ 1];
  const u32* __restrict__ uniq_seq = D.uniq_seq;
  const u32* __restrict__ uniq_gap = D.uniq_gap;
  const SeqKey* __restrict__ uniq_sk = D.uniq_sk;
  const Row* __restrict__ surv = A->gathered ? D.gdone : D.done;
  const int using_x = A->using_x;
  const double best_match0 = A->best_match0;
  u64* hs = Dp->hset[using_x ? 1 : 0];
  const u32 hlog = Dp->hset_log2[using_x ? 1 : 0];
  ASSUME_GLOBAL(tref); ASSUME_GLOBAL(mref); ASSUME_GLOBAL(heap); ASSUME_GLOBAL(mval); ASSUME_GLOBAL(uniq_seq); ASSUME_GLOBAL(uniq_gap);
  ASSUME_GLOBAL(uniq_sk); ASSUME_GLOBAL(surv); ASSUME_GLOBAL(hs);
  for (u32 u = blockIdx.x * blockDim.x + threadIdx.x; u < nu; u += gridDim.x * blockDim.x) {
    const u32 g = uniq_gap[u];
    const Row f = surv[uniq_seq[u]];
    if (mode != 1) {
      heap[nt + u] = f;
      mval[u + g] = f.val;
      mref[u + g] = (nt + u) | (using_x ? TREF_LHS : 0u);
      hset_insert(hs, hlog, key_of(f.val));
    }
    if (mode == 2) continue;
    /* the old-table neighbours' side first (one word each): a neighbour of this pass's own side is never tested, so
       the walk through the gap is not needed for it either -- in a one-sided search that is every gap but the ones
       next to `x`, where one thread would otherwise walk all the new rows of the gap */
    u32 oref_prev = 0, oref_next = 0;
    bool test_prev = g > 0, test_next = g < nt;
    if (test_prev) {
      oref_prev = tref[g - 1];
      const bool other_is_lhs = (oref_prev & TREF_LHS) != 0;
      if (using_x ? other_is_lhs : !other_is_lhs) test_prev = false;
    }
    if (test_next) {
      oref_next = tref[g];
      const bool other_is_lhs = (oref_next & TREF_LHS) != 0;
      if (using_x ? other_is_lhs : !other_is_lhs) test_next = false;
    }
    const bool nb_prev = u > 0 && uniq_gap[u - 1] == g, nb_next = u + 1 < nu && uniq_gap[u + 1] == g;
    if ((test_prev && nb_prev) || (test_next && nb_next)) {
      const SeqKey me = uniq_sk[u];
      if (test_prev && nb_prev) {
        for (u32 j = u; j-- > 0 && uniq_gap[j] == g;)
          if (seq_before(uniq_sk[j], me)) { test_prev = false; break; }
      }
      if (test_next && nb_next) {
        for (u32 j = u + 1; j < nu && uniq_gap[j] == g; j++)
          if (seq_before(uniq_sk[j], me)) { test_next = false; break; }
      }
    }
    for (int side = 0; side < 2; side++) {
      if (side == 0 ? !test_prev : !test_next) continue;
      const Row o = heap[(side == 0 ? oref_prev : oref_next) & ~TREF_LHS];
      const Row& L = using_x ? f : o;
      const Row& R = using_x ? o : f;
      double total = L.der - R.der;
      double score = fabs((R.val - L.val) / total);
      if (score > best_match0) continue; /* NaN passes, as in the reference's test (8270) */
      push_event(D.events, A->cap_events, C, uniq_sk[u].form, side == 0 ? EV_PREV : EV_NEXT, L, R);
    }
  }
}
