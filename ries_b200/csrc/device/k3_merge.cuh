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
  stage_stamp(C, 2);
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
