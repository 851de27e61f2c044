// Regroup into insertion order + K2: stable onesweep LSD radix sort on the fp64 value.
// Part of the single translation unit engine.cu (included inside namespace ries).
#pragma once

/* ------------------------------------------------ survivors into insertion (DFS) order --- */
/* Within one level the completed expressions are ordered (form, lexicographic) and every form
   completes at exactly one level, so a form's expressions are contiguous in done[]. */
__device__ __forceinline__ bool pass_ok(const PassCtl* C) { return (C->overflow & (OVF_NODES | OVF_DONE)) == 0; }

__global__ void k_form_bounds(const DevPtrs* Dp, const PassArgs* A, const PassCtl* C) {
  stage_stamp(C, 1);
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
