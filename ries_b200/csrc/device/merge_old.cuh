// Merge of the old table rows (window of the gap list in shared memory).
// Part of the single translation unit engine.cu (included inside namespace ries).
#pragma once

/* Upper bound by a whole warp over a sorted u32 array: #{u : a[u] <= key}. */
__device__ __forceinline__ u32 warp_upper_bound_u32(const u32* __restrict__ a, u32 n, u32 key) {
  const u32 lane = threadIdx.x & 31;
  u32 lo = 0, hi = n;
  while (hi - lo > 32) {
    const u32 step = (hi - lo + 32) / 33;
    const u32 q = lo + (lane + 1) * step - 1;
    const bool le = (q < hi) && (a[q] <= key);
    const u32 c = __popc(__ballot_sync(0xffffffffu, le));
    const u32 nhi = lo + (c + 1) * step - 1;
    lo += c * step;
    if (c < 32 && nhi < hi) hi = nhi;
  }
  const bool le = (lo + lane < hi) && (a[lo + lane] <= key);
  return lo + __popc(__ballot_sync(0xffffffffu, le));
}

/* Old table positions -- (value, row reference) pairs, 12 bytes: the rows themselves stay where they are in the
   heap -- move up by the number of new rows whose gap is at or below them.  A chunk of consecutive
   old rows only needs the slice of the (sorted) gap list between the counts of its first and last row:
   found by two warp searches, staged in shared memory, searched there. */
#define MO_THREADS 256
#define MO_ITEMS 8
#define MO_CHUNK (MO_THREADS * MO_ITEMS)
#define MO_WIN 4096
__global__ void __launch_bounds__(MO_THREADS)
k_merge_old(const DevPtrs* Dp, const PassCtl* C, const TableState* T) {
  __shared__ __align__(16) u32 s_winraw[MO_WIN + 8];
  __shared__ __align__(8) unsigned long long s_mbar;
  __shared__ u32 s_lo, s_hi;
  if (C->overflow) return;
  const u32 nu = C->n_uniq;
  if (nu == 0) return; /* nothing inserted: the live buffer stays as it is */
  const u32 nt = T->nt;
  const DevPtrs& D = *Dp;
  const u32* __restrict__ tref = D.tref[T->cur];
  const double* __restrict__ tval = D.tval[T->cur];
  u32* mref = D.tref[T->cur ^ 1];
  double* mval = D.tval[T->cur ^ 1];
  const u32* __restrict__ uniq_gap = D.uniq_gap;
  ASSUME_GLOBAL(tref); ASSUME_GLOBAL(tval); ASSUME_GLOBAL(mref); ASSUME_GLOBAL(mval); ASSUME_GLOBAL(uniq_gap);
  const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  (void)lane;
  const u32 nchunks = (nt + MO_CHUNK - 1) / MO_CHUNK;
  if (tid == 0) mbar_init(&s_mbar, 1);
  u32 mbar_phase = 0;
  __syncthreads();
  for (u32 chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
    const u32 base = chunk * MO_CHUNK;
    const u32 last = min(base + MO_CHUNK, nt) - 1;
    if (warp == 0) {
      u32 r = warp_upper_bound_u32(uniq_gap, nu, base);
      if (lane == 0) s_lo = r;
    } else if (warp == 1) {
      u32 r = warp_upper_bound_u32(uniq_gap, nu, last);
      if (lane == 0) s_hi = r;
    }
    __syncthreads();
    const u32 lo = s_lo, wn = s_hi - s_lo;
    const bool in_smem = wn <= MO_WIN;
    /* the slice of the gap list arrives by one bulk asynchronous copy (16-byte granules: from the multiple of four
       at or below lo) while the threads fetch the chunk's (value, reference) pairs */
    const u32 lo_al = lo & ~3u;
    if (in_smem && wn && tid == 0) bulk_g2s(s_winraw, uniq_gap + lo_al, (((wn + (lo - lo_al)) + 3u) & ~3u) * 4u, &s_mbar);
    double tv[MO_ITEMS];
    u32 tr[MO_ITEMS];
#pragma unroll
    for (int r = 0; r < MO_ITEMS; r++) {
      const u32 j = base + r * MO_THREADS + tid;
      tv[r] = (j < nt) ? tval[j] : 0.0;
      tr[r] = (j < nt) ? tref[j] : 0u;
    }
    if (in_smem && wn) { mbar_wait(&s_mbar, mbar_phase); mbar_phase ^= 1u; }
    const u32* win = in_smem ? s_winraw + (lo - lo_al) : uniq_gap + lo;
#pragma unroll
    for (int r = 0; r < MO_ITEMS; r++) {
      const u32 j = base + r * MO_THREADS + tid;
      if (j < nt) {
        u32 a = 0, b = wn;
        if (wn) {
          while (a < b) {
            u32 mid = (a + b) >> 1;
            if (win[mid] <= j) a = mid + 1; else b = mid;
          }
        }
        const u32 dst = j + lo + a;
        mval[dst] = tv[r];
        mref[dst] = tr[r];
      }
    }
    __syncthreads();
  }
}

/* A side's hash set re-entered from the table (after the set has grown, or after a pass whose merge was undone). */
__global__ void k_hset_rebuild(const DevPtrs* Dp, const TableState* T, int side) {
  const DevPtrs& D = *Dp;
  const u32 nt = T->nt;
  const u32* __restrict__ tref = D.tref[T->cur];
  const double* __restrict__ tval = D.tval[T->cur];
  u64* hs = Dp->hset[side];
  const u32 hlog = Dp->hset_log2[side];
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < nt; i += gridDim.x * blockDim.x)
    if ((int)(tref[i] >> 31) == side) hset_insert(hs, hlog, key_of(tval[i]));
}
