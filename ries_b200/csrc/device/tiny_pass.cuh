// Tiny passes: the whole pass in three launches instead of sixteen.
// Part of the single translation unit engine.cu (included inside namespace ries).
#pragma once

/* Most passes of a search are tiny (sizes grow ~1.3x per complexity step from a single expression up: at the
   default level two thirds of the passes hold fewer than a thousand prefixes), and a tiny pass through the
   general sequence -- pass begin, one K1 launch per expression position, the single-block post kernel,
   merge, refine, pass end -- is nothing but launch latency.  For a pass predicted to be tiny the host
   launches
     k_tiny_front  <<<1, 1024>>>  pass begin; K1 for every expression position (the first EXP_THREADS threads
                                  loop over the positions, a named barrier between them); then regroup +
                                  sort + classify + match + merge-new (small_post_body) by the whole block
     k_merge_old   <<<48, ...>>>  unchanged
     k_refine_end  <<<1, 128>>>   refine, then pass end
   Same device functions, same results.  A frontier wider than the prediction allows sets OVF_SMALL and the host
   reruns the pass through the general kernels. */
#define TINY_THREADS SP_THREADS
#define TINY_DYN_SMEM ((EXP_DYN_SMEM > SP_DYN_SMEM) ? EXP_DYN_SMEM : SP_DYN_SMEM)

__global__ void __launch_bounds__(TINY_THREADS, 1)
k_tiny_front(const DevPtrs* Dp, const PassArgs* Ap, PassCtl* ctl, TableState* T, u32 max_frontier) {
  stage_stamp(ctl, 0);
  pass_begin_body(Dp, Ap, ctl, T);
  __threadfence();
  __syncthreads();
  if (threadIdx.x < EXP_THREADS) {
    /* the frontier bounds of the next position come back through the control block, written by thread 0
       before expand_level's closing barrier */
    volatile PassCtl* vc = ctl;
    for (int lvl = 0; lvl <= RIES_MAX_ELEN; lvl++) {
      const u32 width = vc->lvl_end - vc->lvl_begin;
      if (width == 0) break;
      if (width > max_frontier) { /* not tiny after all */
        if (threadIdx.x == 0) atomicOr(&ctl->overflow, OVF_SMALL);
        break;
      }
      expand_level<true>(Dp, Ap, ctl, lvl == 0);
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) ctl->gen = ctl->done_count; /* this K1 leaves no holes */
  stage_stamp(ctl, 1);
  if (ctl->overflow & OVF_SMALL) return; /* uniform */
  small_post_body(Dp, Ap, ctl, T);
  __syncthreads();
  stage_stamp(ctl, 2);
}

__global__ void __launch_bounds__(128)
k_refine_end(const DevPtrs* Dp, const PassArgs* A, PassCtl* C, TableState* T) {
  stage_stamp(C, 3);
  refine_body(Dp, A, C);
  __threadfence();
  __syncthreads();
  pass_end_body(Dp, A, C, T);
  __syncthreads();
  stage_stamp(C, 4);
  post_to_host(Dp, A, C);
}
