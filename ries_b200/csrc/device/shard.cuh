// Sharded passes: local pre-dedupe, compaction, stage-1 reset.
// Part of the single translation unit engine.cu (included inside namespace ries).
#pragma once

/* ------------------------------------------------ sharded passes: local pre-dedupe -------- */
/* One search over several GPUs: before the ranks exchange their complete expressions, each drops the ones
   that cannot matter, using its replica of the table and its own survivors in sorted order:
     value already in the table on the SAME side      -> dropped (check_exact_match discards it, 8550-8593);
                                                         k_keys' hash probe has removed nearly all of them already
     value already in the table on the opposite side  -> kept, every occurrence (each is an EXACT event)
     new value                                        -> only the occurrence inserted first is kept
   The global first-inserted winner of a value is the minimum over the ranks' winners, so the replicated
   dedupe that follows sees exactly the rows it needs.  keep[] was cleared by k_keys. */
__global__ void __launch_bounds__(CC_THREADS)
k_local_filter(const DevPtrs* Dp, const PassArgs* A, PassCtl* C, const TableState* T) {
  __shared__ double s_win[CC_WIN + 1];
  __shared__ u32 s_lo, s_hi;
  if (!pass_ok(C)) return;
  const u32 n = C->n_keys;
  const u32 nchunks = (n + CC_CHUNK - 1) / CC_CHUNK;
  const DevPtrs& D = *Dp;
  const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const u32 cur = T->cur, nt = T->nt;
  const double* __restrict__ tval = D.tval[cur];
  const Row* __restrict__ trow = D.trow[cur];
  const u64* __restrict__ keys = D.keys[0];
  const u32* __restrict__ idx = D.idx[0];
  const Row* __restrict__ done = D.done;
  const uint16_t* __restrict__ done_form = D.done_form;
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
        if ((p < nt) && (win[a] == v)) {
          const bool tree_is_lhs = (trow[p].der != 0.0);
          if (using_x ? !tree_is_lhs : tree_is_lhs) D.keep[idx[i]] = 1; /* every occurrence is an EXACT event */
        } else if (!dup_prev) {
          /* a new value: of this rank's run only the expression inserted first can be the global winner */
          u32 w = idx[i];
          if (i + 1 < n && keys[i + 1] == k) {
            Sym128 ws = done[w].sym;
            u32 wf = done_form[w];
            for (u32 j = i + 1; j < n && keys[j] == k; j++) {
              const u32 s2 = idx[j];
              const Sym128 s2s = done[s2].sym;
              const u32 s2f = done_form[s2];
              if (seq_before(s2f, s2s, wf, ws)) { w = s2; ws = s2s; wf = s2f; }
            }
          }
          D.keep[w] = 1;
        }
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
  const u32 n = C->done_count; /* slots of done[]; keep[] is 0 for holes and for everything the filters dropped */
  const u32 nchunks = (n + 255) / 256;
  if (blockIdx.x >= nchunks) return;
  const DevPtrs& D = *Dp;
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
      D.ksend[u] = D.done[i];
      D.ksend_form[u] = D.done_form[i];
    }
    __syncthreads();
  }
}

/* Between the local sort of a sharded pass and the replicated one: fresh look-back tags, zeroed histograms,
   counters and tickets. */
__global__ void k_stage1_reset(const DevPtrs* D, const PassArgs* A, PassCtl* C, TableState* T) {
  u32* sc = (u32*)D->sortctl;
  for (u32 i = threadIdx.x; i < sizeof(SortCtl) / 4; i += blockDim.x) sc[i] = 0;
  if (threadIdx.x == 0) {
    C->chunk_ticket = 0; C->aux_ticket = 0; C->n_keys = 0;
    u32 e = (T->epoch + 16) & 0x3ffffff0u;
    if (e == 0) e = 16;
    T->epoch = e;
  }
}
