// K4: cv_simplify + Newton on the candidates.
// Part of the single translation unit engine.cu (included inside namespace ries).
#pragma once

/* K4 */
__device__ __forceinline__ void refine_body(const DevPtrs* Dp, const PassArgs* A, const PassCtl* C) {
  __shared__ RiesParams P;
  u32 n = C->overflow ? 0 : C->n_events;
  if (n == 0) return; /* uniform over the block */
  if (n > A->cap_events) n = A->cap_events;
  {
    const u32* src = (const u32*)Dp->P;
    u32* dst = (u32*)&P;
    for (u32 i = threadIdx.x; i < sizeof(RiesParams) / 4; i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  MatchEvent* events = Dp->events;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (events[i].kind == EV_EXACT) continue;
    MatchEvent ev = events[i];
    refine_event(P, ev);
    events[i] = ev;
  }
}

__global__ void k_refine(const DevPtrs* Dp, const PassArgs* A, const PassCtl* C) {
  stage_stamp(C, 3);
  refine_body(Dp, A, C);
}
