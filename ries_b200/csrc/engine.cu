// The sm_100a device engine: every hot stage of the RIES search as hand-written CUDA kernels.
//
//   K1  k_expand            enumerate + evaluate one expression position of every live prefix
//                           [ge_2 + exec, ries.c:8908-9475, 5308-5950]
//   K2  k_keys              sort keys; what the table already holds on this side is dropped by a hash probe
//       k_rs_onesweep       LSD radix sort of the rest on the fp64 value
//       k_classify_compact  duplicate pruning inside the pass and against the resident table
//       k_match_merge_new, k_merge_old   merge into the resident sorted table   [bt_insert, ries.c:8622-8724]
//   K3  k_match_merge_new   gap-local neighbour test with the derivative-scaled window
//                                                           [check_sides/check_match, 8488, 8243]
//   K4  k_refine            cv_simplify + Newton on the compacted candidates        [7297, 7139]
//
// A whole complexity pass is DEVICE-DRIVEN: every kernel has a fixed grid and reads its problem size
// (frontier bounds, number of survivors, number of new rows, table size) from a control block in HBM
// that the previous kernel wrote, and every buffer pointer lives in a device-resident pointer table.
// Kernel arguments are therefore constant for the life of the engine, the per-pass launch sequence
// is captured once into a CUDA graph per expression length, and the host synchronises ONCE per pass
// (to read counters and candidates for the sequential replay that decides the next pass).
//
// Files (one translation unit): device/state.cuh control blocks, pointer table, insertion-order keys, hash set;
// device/pass.cuh; device/k1_free.cuh (K1); device/k1_expand.cuh (K1 of the tiny passes); device/k2_sort.cuh (keys +
// onesweep); device/k3_merge.cuh (classify, match + merge-new); device/merge_old.cuh; device/small_pass.cuh +
// device/tiny_pass.cuh (single-block path); device/shard.cuh (one search over several GPUs); device/k4_refine.cuh;
// this file: buffers, graphs, the per-pass host loop.
//
// Data layout in HBM (all device-resident for the whole search):
//   nodes[]   48 B AoS, one per live prefix of the running pass (cactus stack, see params.h)
//   table     (value 8 B, row reference 4 B) sorted by value, ping-pong; the 32 B rows in an append-only heap
//   done[]    32 B rows + 2 B form index, the pass's complete expressions in NO particular order (state.cuh, SeqKey)
//   hset[2]   per side, a hash set of the keys resident in the table (pre-dedupe before the sort)
// No tensor cores: nothing here is a dense contraction.
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sched.h>
#include <time.h>

#include <algorithm>
#include <atomic>
#include <map>
#include <mutex>
#include <vector>

#include "core/expand.h"
#include "core/refine.h"
#include "host/dist.h"
#include "host/ries_host.h"

namespace ries {

/* A failed allocation is the caller's request being too deep for this GPU, not a fault: it unwinds as the
   reference's out-of-memory exit (status 1, ries.c:4879) and ries_main() drops the engine so that the next call
   starts from clean buffers.  Anything else is a genuine fault of the device or of this code. */
inline void cuda_check_fail(cudaError_t e, const char* file, int line) {
  if (e == cudaErrorMemoryAllocation) {
    (void)cudaGetLastError();
    fprintf(stderr, "ries_b200: out of device memory at %s:%d (search level too deep for this GPU)\n", file, line);
    throw ExitRequest{1};
  }
  fprintf(stderr, "ries_b200: CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e), file, line, cudaGetErrorString(e));
  abort();
}
#define CUDA_CHECK(x)                                          \
  do {                                                         \
    cudaError_t e_ = (x);                                      \
    if (e_ != cudaSuccess) cuda_check_fail(e_, __FILE__, __LINE__); \
  } while (0)

typedef unsigned long long u64;
typedef unsigned int u32;

/* The kernels, in pipeline order (one translation unit: they share the structs of device/state.cuh). */
#include "device/state.cuh"
#include "device/pass.cuh"
#include "device/k1_expand.cuh"
#include "device/k1_free.cuh"
#include "device/k2_sort.cuh"
#include "device/k3_merge.cuh"
#include "device/small_pass.cuh"
#include "device/shard.cuh"
#include "device/merge_old.cuh"
#include "device/k4_refine.cuh"
#include "device/tiny_pass.cuh"

/* ---------------------------------------------------------------- host-side engine ------- */

template <typename T>
struct DevBuf {
  T* p = 0;
  size_t cap = 0;
  /* returns true when the pointer changed */
  bool ensure(size_t n, cudaStream_t st, bool keep = false, bool zero = false) {
    if (n <= cap) return false;
    size_t ncap = std::max(n, cap + cap / 2);
    T* q = 0;
    if (!(keep && p && cap) && p) { /* nothing to carry over: give the old block back first, so the pool can
                                       reuse its physical memory for the bigger one */
      CUDA_CHECK(cudaFreeAsync(p, st));
      p = 0;
      cap = 0;
    }
    CUDA_CHECK(cudaMallocAsync(&q, ncap * sizeof(T), st));
    if (zero) CUDA_CHECK(cudaMemsetAsync(q, 0, ncap * sizeof(T), st));
    if (keep && p && cap) CUDA_CHECK(cudaMemcpyAsync(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, st));
    if (p) CUDA_CHECK(cudaFreeAsync(p, st));
    p = q;
    cap = ncap;
    return true;
  }
  void release(cudaStream_t st) { if (p) cudaFreeAsync(p, st); p = 0; cap = 0; }
};


/* ---------------------------------------------------------------- one RHS database per batch ---- */
/* BASELINE config 5: "a batch of targets sharing one RHS database".  Right-hand sides contain no x, and what the
   reference does with them depends on the target in exactly three ways: (1) the overflow bound
   p_ovr = |T| / k_prune_deriv cuts every prefix whose value reaches it (ries.c:12250, test at 9020) -- so the set
   of generated expressions and the dead-end counters shrink with |T|; (2) an RHS value equal to an LHS value
   already in the tree is reported, not inserted (8648-8661); (3) the scheduler interleaves the sides by their
   insert counts (12424-12460).  (2) and (3) happen AFTER enumeration, so what is shared here is the enumeration:
   the RHS pass of a complexity is run ONCE per batch, by whichever engine needs it first, under the batch's
   largest bound, with every prefix carrying M = the largest |value| along it.  A target with bound P then takes the
   rows with M < P (exactly the expressions its own run would have generated; their number is its "expressions"
   count), and its dead-end count is the shared one corrected by the recorded cases where P cuts in (TailRec).
   Everything after that -- pre-dedupe against ITS table, sort, cross-side exact matches, neighbour tests, merge --
   is the target's own pipeline, unchanged.  Reports stay byte-identical (tests: 64 targets against the reference). */
struct SharedPass {
  Row* rows = 0;
  uint16_t* forms = 0;
  TailRec* tail = 0;
  u32 n_slots = 0, n_tail = 0, node_count = 0;
  u64 dead_base = 0, gen = 0;
};
struct SharedRhs {
  std::mutex mu;
  std::map<int, SharedPass> passes;
  RiesParams sig;
  double p_max = 0, p_min = 0;
  int c_share = -1; /* RHS passes from this complexity on are shared (set by the first engine that gets there) */
  int dev = 0;
  std::atomic<unsigned long long> built{0}, consumed{0};
  ~SharedRhs() {
    cudaSetDevice(dev);
    for (auto& kv : passes) {
      cudaFreeAsync(kv.second.rows, cudaStreamPerThread); cudaFreeAsync(kv.second.forms, cudaStreamPerThread);
      cudaFreeAsync(kv.second.tail, cudaStreamPerThread);
    }
    cudaStreamSynchronize(cudaStreamPerThread);
  }
};

RiesParams shared_rhs_signature(const RiesParams& P) {
  RiesParams s = P;
  s.target = 0; s.p_ovr = 0; s.n_ovr = 0; s.targ_tags = 0;
  return s;
}
void* shared_rhs_create(const RiesParams& sig, double p_max, double p_min) {
  SharedRhs* r = new SharedRhs();
  r->sig = sig; r->p_max = p_max; r->p_min = p_min;
  const char* dev_env = getenv("RIES_DEVICE");
  r->dev = dev_env ? atoi(dev_env) : 0;
  return r;
}
void shared_rhs_destroy(void* ctx) { delete static_cast<SharedRhs*>(ctx); }
void shared_rhs_stats(void* ctx, unsigned long long* b, unsigned long long* c) {
  SharedRhs* r = static_cast<SharedRhs*>(ctx);
  *b = r->built.load(); *c = r->consumed.load();
}

class DeviceEngine : public Engine {
 public:
  DeviceEngine() {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
      fprintf(stderr, "ries_b200: no CUDA device available (%s). This build has no CPU path.\n",
              cudaGetErrorString(e));
      throw ExitRequest{3};
    }
    const char* dev_env = getenv("RIES_DEVICE");
    int dev = dev_env ? atoi(dev_env) : 0;
    CUDA_CHECK(cudaSetDevice(dev));
    dev_ = dev;
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
    num_sms_ = prop.multiProcessorCount;
    {
      /* keep freed blocks in the stream-ordered pool instead of returning them to the OS */
      cudaMemPool_t pool;
      if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
      }
    }
    CUDA_CHECK(cudaStreamCreateWithFlags(&st_, cudaStreamNonBlocking));
    int occ = 1;
    CUDA_CHECK(cudaFuncSetAttribute(k_rs_onesweep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RS_DYN_SMEM));
    const int xf_max_smem = (int)std::min<size_t>(xf_dyn_smem(XF_MAXQ), (size_t)prop.sharedMemPerBlockOptin - 8192);
    CUDA_CHECK(cudaFuncSetAttribute(k_expand, cudaFuncAttributeMaxDynamicSharedMemorySize, xf_max_smem));
    CUDA_CHECK(cudaFuncSetAttribute(k_expand, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CUDA_CHECK(cudaFuncSetAttribute(k_expand_all, cudaFuncAttributeMaxDynamicSharedMemorySize, xf_max_smem));
    CUDA_CHECK(cudaFuncSetAttribute(k_expand_all, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    /* measured (tools/ab_l6.py, -l6 and the default level): no faster than the chained launches -- a small
       position costs its dependent-load chain and two block barriers, not its launch -- so it stays an option */
    use_persist_ = getenv("RIES_PERSIST") != 0;
    if (const char* e = getenv("RIES_PERSIST_MAX")) persist_max_ = (u32)strtoul(e, 0, 10);
    (void)occ;
    wide_grid_ = (u32)(num_sms_ * 8);
    use_graphs_ = getenv("RIES_NO_GRAPHS") == 0;
    use_small_post_ = getenv("RIES_NO_SMALL_POST") == 0; /* test hook: every pass through the general kernels */
    force_small_post_ = getenv("RIES_FORCE_SMALL_POST") != 0;
    /* tiny passes (tiny_pass.cuh): RIES_NO_TINY switches them off, RIES_TINY_NODES / RIES_TINY_FRONTIER move the
       prediction threshold and the frontier width at which the kernel gives up (test hooks) */
    use_tiny_ = getenv("RIES_NO_TINY") == 0;
    /* Programmatic dependent launch between the K1 launches of a pass (the next position's prologue under the
       previous position's tail) was worth 0.4 ms of 44 at -l6 in round 1 -- and with this round's K1 it DEADLOCKS
       now and then when a pass overflows an arena and the remaining positions are empty, microsecond-long kernels
       (tools/hang_repro.sh: 10 of 14 runs of `ries 143 -i -l3` with RIES_ARENA_CAP=4096 hang with it, 0 of 14
       without; the device is then stuck between the last level's election and the first kernel of the next stage).
       Off unless asked for. */
    use_pdl_ = getenv("RIES_PDL") != 0 && getenv("RIES_NO_PDL") == 0;
    if (const char* e = getenv("RIES_TINY_NODES")) tiny_nodes_ = (u32)strtoul(e, 0, 10);
    if (const char* e = getenv("RIES_TINY_FRONTIER")) tiny_frontier_ = (u32)strtoul(e, 0, 10);
    CUDA_CHECK(cudaFuncSetAttribute(k_tiny_front, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TINY_DYN_SMEM));
    host_prof_ = getenv("RIES_HOST_PROF") != 0; /* host time per pass phase, printed at the next begin()/exit */
    pass_log_ = getenv("RIES_PASS_LOG") != 0; /* one line per pass on stderr: sizes and device time per stage */
    CUDA_CHECK(cudaMalloc(&d_P_, sizeof(RiesParams)));
    CUDA_CHECK(cudaMalloc(&d_tt_, sizeof(TestTables)));
    CUDA_CHECK(cudaMalloc(&d_out_, sizeof(PassOut)));
    CUDA_CHECK(cudaMemset(d_out_, 0, sizeof(PassOut)));
    d_ctl_ = &d_out_->ctl;
    const size_t passin_bytes = PASSIN_FORMS_OFFSET + (size_t)PASSIN_MAX_FORMS * sizeof(FormDesc);
    CUDA_CHECK(cudaMalloc(&d_passin_, passin_bytes));
    d_args_ = (PassArgs*)d_passin_;
    d_forms_ = (FormDesc*)(d_passin_ + PASSIN_FORMS_OFFSET);
    CUDA_CHECK(cudaMalloc(&d_tab_, sizeof(TableState)));
    CUDA_CHECK(cudaMalloc(&d_ptrs_, sizeof(DevPtrs)));
    CUDA_CHECK(cudaMalloc(&d_sortctl_, sizeof(SortCtl)));
    CUDA_CHECK(cudaMalloc(&d_opcnt_, 2 * RIES_MAX_SYMS * sizeof(u64)));
    CUDA_CHECK(cudaHostAlloc(&h_out_, sizeof(PassOut), cudaHostAllocMapped));
    memset(h_out_, 0, sizeof(PassOut));
    CUDA_CHECK(cudaHostGetDevicePointer((void**)&d_host_out_, h_out_, 0));
    poll_host_ = getenv("RIES_NO_HOST_POLL") == 0; /* test hook: copy + synchronise per pass instead */
    h_ctl_ = &h_out_->ctl;
    h_head_ = h_out_->head;
    CUDA_CHECK(cudaMallocHost(&h_passin_, passin_bytes));
    h_args_ = (PassArgs*)h_passin_;
    h_forms_ = (FormDesc*)(h_passin_ + PASSIN_FORMS_OFFSET);
    CUDA_CHECK(cudaMallocHost(&h_ptrs_, sizeof(DevPtrs)));
    CUDA_CHECK(cudaMallocHost(&h_args2_, sizeof(PassArgs)));
    CUDA_CHECK(cudaMalloc(&d_shdr_, sizeof(SlotHdr)));
    CUDA_CHECK(cudaMalloc(&d_ghdr_, RIES_MAX_RANKS * sizeof(SlotHdr)));
    CUDA_CHECK(cudaMalloc(&d_ghdr2_, RIES_MAX_RANKS * sizeof(SlotHdr)));
    CUDA_CHECK(cudaMemset(d_ghdr2_, 0, RIES_MAX_RANKS * sizeof(SlotHdr)));
    if (const char* e = getenv("RIES_WATCHDOG_SEC")) watchdog_s_ = atof(e);
    no_spread_ = getenv("RIES_NO_K1_SPREAD") != 0; /* A/B, test hook */
    if (const char* e = getenv("RIES_SMALL_GRID")) small_grid_ = (u32)std::max(1, atoi(e)); /* tuning hook */
    use_range_ = getenv("RIES_NO_RANGE_SPLIT") == 0; /* A/B, test hook: the whole second half replicated */
    if (const char* e = getenv("RIES_RANGE_MIN")) range_min_ = std::max(16u, (u32)strtoul(e, 0, 10)); /* test hook */
    CUDA_CHECK(cudaMemset(d_ghdr_, 0, RIES_MAX_RANKS * sizeof(SlotHdr)));
    if (const char* e = getenv("RIES_SHARD_MIN")) shard_min_ = (u32)atol(e);
    memset(h_ptrs_, 0, sizeof(DevPtrs));
    for (int i = 0; i < 5; i++) CUDA_CHECK(cudaEventCreate(&ev_[i]));
    for (int i = 0; i < 2; i++) CUDA_CHECK(cudaEventCreate(&ev_g_[i]));
    const char* arena_env = getenv("RIES_ARENA_CAP"); /* test hook: start tiny to exercise overflow-and-rerun */
    const size_t arena0 = arena_env ? (size_t)atoi(arena_env) : (size_t)(1u << 20);
    const char* cap_env = getenv("RIES_EVENTS_CAP"); /* test hook: start tiny to exercise the regrow path */
    staged_.ensure((size_t)2 * EXP_MAX_ITEMS, st_); /* K1 of the tiny passes: one block */
    surv_.ensure(SP_CAP, st_);
    events_.ensure(std::max<size_t>(cap_env ? (size_t)atoi(cap_env) : (size_t)(1u << 14), 1), st_);
    grow_arena(arena0, arena0);
    ensure_table(1u << 16);
    ptrs_dirty_ = true;
  }
  ~DeviceEngine() override {
    cudaSetDevice(dev_);
    cudaStreamSynchronize(st_);
    for (auto& kv : graphs_) cudaGraphExecDestroy(kv.second);
    nodes_.release(st_); done_.release(st_); done_form_.release(st_); surv_.release(st_); uniq_sk_.release(st_);
    hset_[0].release(st_); hset_[1].release(st_); tail_.release(st_);
    for (int b = 0; b < 2; b++) { keys_[b].release(st_); idx_[b].release(st_); tref_[b].release(st_); tval_[b].release(st_); }
    heap_.release(st_);
    uniq_seq_.release(st_); uniq_gap_.release(st_); events_.release(st_); staged_.release(st_);
    rs_state_.release(st_); cc_state_.release(st_); gdone_.release(st_); gdone_form_.release(st_);
    ksend_.release(st_); ksend_form_.release(st_);
    cudaStreamSynchronize(st_);
    cudaFreeHost(h_args2_); cudaFree(d_shdr_); cudaFree(d_ghdr_); cudaFree(d_ghdr2_);
    g2_idx_.release(st_); g2_gap_.release(st_); g2_events_.release(st_);
    cudaFree(d_P_); cudaFree(d_tt_); cudaFree(d_out_); cudaFree(d_passin_); cudaFree(d_tab_); cudaFree(d_ptrs_); cudaFree(d_sortctl_);
    cudaFree(d_opcnt_);
    cudaFreeHost(h_out_); cudaFreeHost(h_passin_); cudaFreeHost(h_ptrs_);
    if (h_events_) cudaFreeHost(h_events_);
    for (int i = 0; i < 5; i++) cudaEventDestroy(ev_[i]);
    for (int i = 0; i < 2; i++) cudaEventDestroy(ev_g_[i]);
    cudaStreamDestroy(st_);
  }
  const char* name() const override { return "sm_100a"; }
  void set_shared(void* ctx) override { shared_ = static_cast<SharedRhs*>(ctx); }
  void set_throughput_mode(bool on) override { throughput_mode_ = on && getenv("RIES_BATCH_FULL_GRIDS") == 0; }

  void begin(const RiesParams& P) override {
    /* the calling thread may be a new one (batch workers take pooled engines): its current device is 0 */
    CUDA_CHECK(cudaSetDevice(dev_));
    if (host_prof_ && hp_[5] > 0) {
      fprintf(stderr, "host per pass (us): prepare %.1f enqueue %.1f wait %.1f collect %.1f events %.1f over %.0f passes\n",
              hp_[0] / hp_[5], hp_[1] / hp_[5], hp_[2] / hp_[5], hp_[3] / hp_[5], hp_[4] / hp_[5], hp_[5]);
      for (double& v : hp_) v = 0;
    }
    if (!test_tables_cover(P)) {
      fprintf(stderr, "ries_b200: a symbol weight is outside the test tables' range\n");
      abort();
    }
    build_test_tables(P, h_tt_);
    CUDA_CHECK(cudaMemcpyAsync(d_P_, &P, sizeof(P), cudaMemcpyHostToDevice, st_));
    CUDA_CHECK(cudaMemcpyAsync(d_tt_, &h_tt_, sizeof(TestTables), cudaMemcpyHostToDevice, st_));
    CUDA_CHECK(cudaMemsetAsync(d_opcnt_, 0, 2 * RIES_MAX_SYMS * sizeof(u64), st_));
    TableState t0;
    t0.cur = 0; t0.nt = 0; t0.epoch = epoch_host_; t0.pad = 0;
    CUDA_CHECK(cudaMemcpyAsync(d_tab_, &t0, sizeof(t0), cudaMemcpyHostToDevice, st_));
    CUDA_CHECK(cudaStreamSynchronize(st_));
    memcpy(sym_chars_, P.sym_char, sizeof(sym_chars_));
    nt_ = 0;
    cur_ = 0;
    last_nu_ = 0;
    reset_hash();
    cur_P_ = P;
    /* this search may use the batch's RHS enumeration iff it differs from the batch's common parameters in the
       target alone and its overflow bound lies in the range the shared passes keep records for */
    shared_ok_ = false;
    if (shared_ && !P.canon_ops && dist_state().world <= 1 && dist_state().fake_world <= 1) {
      const RiesParams sg = shared_rhs_signature(P);
      shared_ok_ = memcmp(&sg, &shared_->sig, sizeof(RiesParams)) == 0 && P.p_ovr <= shared_->p_max && P.p_ovr >= shared_->p_min;
    }
    {
      /* K1's operator queues: one per (complete?, enabled symbol); their number sets the kernel's shared memory
         and with it the grid (blocks per SM) and the captured launch sequences */
      const int nq = 2 * (P.cls_n[0] + P.cls_n[1] + P.cls_n[2]);
      if (nq != k1_nq_) {
        for (auto& kv : graphs_) cudaGraphExecDestroy(kv.second);
        graphs_.clear();
        k1_nq_ = nq;
        int occ = 1;
        CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_expand_all, XF_THREADS, xf_dyn_smem(nq)));
        if (occ < 1) occ = 1;
        expand_grid_ = (u32)(num_sms_ * occ);
      }
    }
    prev_nodes_[0] = prev_nodes_[1] = prev_done_[0] = prev_done_[1] = 0;
    prev_gen_[0] = prev_gen_[1] = prev_keys_[0] = prev_keys_[1] = shard_kept_[0] = shard_kept_[1] = 0;
    range_new_[0] = range_new_[1] = 0;
    kernel_launches = 0;
    h2d_bytes = sizeof(P);
    d2h_bytes = 0;
    for (int i = 0; i < 4; i++) { stage_ms[i] = 0; stage_launches[i] = 0; }
    sort_keys = merge_rows = expand_nodes = 0;
    sharded_passes = gather_bytes = 0;
    tiny_passes = 0;
    gather_ms = 0;
    /* cross-check mode (bench.py): the stages of every pass bracketed by CUDA events on this stream, as before the
       device clock stamps; tiny passes go through the general kernels so that the stages stay separable */
    stage_events_ = getenv("RIES_STAGE_EVENTS") != 0;
  }

  /* host-side twin of seq_before (state.cuh): a was inserted before b */
  static bool sym_before_host(const Sym128& a, const Sym128& b) {
    for (int i = 0; i < RIES_MAX_ELEN + 1; i++) {
      const unsigned ca = sym_get(a, i), cb = sym_get(b, i);
      if (ca != cb) return ca < cb;
    }
    return false;
  }
  static bool seq_before_host(u32 fa, const Sym128& a, u32 fb, const Sym128& b) {
    return fa != fb ? fa < fb : sym_before_host(a, b);
  }

  long inserted_before(const MatchEvent& ev, bool inclusive) override {
    if (!last_nu_) return 0;
    long c = 0;
    if (last_tiny_) { /* tiny passes number their expressions in insertion order */
      std::vector<u32> h(last_nu_);
      CUDA_CHECK(cudaMemcpyAsync(h.data(), uniq_seq_.p, last_nu_ * sizeof(u32), cudaMemcpyDeviceToHost, st_));
      CUDA_CHECK(cudaStreamSynchronize(st_));
      for (u32 s : h) c += inclusive ? (s <= ev.seq) : (s < ev.seq);
      return c;
    }
    std::vector<SeqKey> h(last_nu_);
    CUDA_CHECK(cudaMemcpyAsync(h.data(), uniq_sk_.p, last_nu_ * sizeof(SeqKey), cudaMemcpyDeviceToHost, st_));
    CUDA_CHECK(cudaStreamSynchronize(st_));
    const Sym128& me = last_using_x_ ? ev.lhs.sym : ev.rhs.sym; /* the NEW expression of the event */
    for (const SeqKey& k : h) {
      Sym128 ks; ks.lo = k.lo; ks.hi = k.hi;
      const bool before = seq_before_host(k.form, ks, ev.seq, me);
      const bool same = k.form == ev.seq && ks.lo == me.lo && ks.hi == me.hi;
      c += (before || (inclusive && same)) ? 1 : 0;
    }
    return c;
  }

  void run_pass(int using_x, int cplx, const std::vector<FormDesc>& forms, double best_match0,
                bool skip_exact, PassOutput& out) override {
    last_nu_ = 0;
    const u32 nforms = (u32)forms.size();
    if (nforms == 0) return;
    double hp_t0 = host_prof_ ? now_us() : 0, hp_t1 = 0, hp_t2 = 0, hp_t3 = 0;
    if (nforms > PASSIN_MAX_FORMS - 1) {
      fprintf(stderr, "ries_b200: %u forms in one pass exceeds the 16-bit form index\n", nforms);
      abort();
    }
    int max_len = 0;
    for (const FormDesc& f : forms) max_len = std::max(max_len, (int)f.flen);
    if (nodes_.cap < (size_t)nforms + 16) grow_arena(nforms + 16, done_.cap);
    memcpy(h_forms_, forms.data(), nforms * sizeof(FormDesc)); /* pinned; the last pass's copy has completed */
    h2d_bytes += nforms * sizeof(FormDesc) + sizeof(PassArgs);

    const int side = using_x ? 1 : 0;
    bool no_tiny = !use_small_post_;
    unsigned long long gen_k1 = 0; /* sharded passes: complete expressions of all ranks' K1, before the pre-dedupe */
    bool was_sharded = false, was_tiny = false, was_shared = false, was_ranged = false;
    {
      /* sizes grow ~1.3-1.6x per complexity step: keep twice the previous same-side pass in hand, so that
         the overflow-and-rerun path stays the exception (it costs a whole pass) */
      const size_t want_nodes = (size_t)prev_nodes_[side] * 2, want_done = (size_t)prev_done_[side] * 2;
      if (!getenv("RIES_ARENA_CAP") && (want_nodes > nodes_.cap || want_done > done_.cap))
        grow_arena(std::max(want_nodes, nodes_.cap), std::max(want_done, done_.cap));
    }
    while (true) {
      /* the merged table needs room for every survivor the arena can hold */
      ensure_table((size_t)nt_ + done_.cap);
      ensure_hash(side);
      if (ptrs_dirty_) upload_ptrs();
      h_args_->using_x = using_x; h_args_->cplx = cplx; h_args_->skip_exact = skip_exact ? 1 : 0;
      h_args_->nforms = (int)nforms; h_args_->best_match0 = best_match0;
      h_args_->cap_nodes = (u32)std::min<size_t>(nodes_.cap, 0xffffffffu);
      h_args_->cap_done = (u32)std::min<size_t>(done_.cap, 0xffffffffu);
      h_args_->cap_events = (u32)events_.cap; h_args_->seq = ++pass_seq_;
      /* size class from the previous pass of the same side (sizes grow ~1.3-1.6x per complexity step) */
      int small = (prev_nodes_[side] <= 12000 && prev_done_[side] <= 3000) ? 1 : 0;
      /* throughput mode (the engines of a batch share the GPU): a mid-size pass on full grids keeps the whole
         machine busy scheduling blocks that find nothing to do, and K1's queues fill better when a block sees
         several tiles; a lone search wants the latency of full grids instead */
      if (!small && throughput_mode_ && nt_ <= 4000000u) {
        if (prev_nodes_[side] <= 150000) small = 2;
        else if (prev_nodes_[side] <= 1200000) small = 3;
      }
      /* one search over several GPUs: only passes big enough to pay for the exchange are sharded */
      const DistState& ds = dist_state();
      const int shard_world = ds.fake_world > 1 ? ds.fake_world : ds.world;
      const bool shard = shard_world > 1 && prev_gen_[side] >= shard_min_;
      h_args_->shard_rank = (shard && ds.fake_world <= 1) ? ds.rank : 0;
      h_args_->shard_world = shard ? shard_world : 1;
      h_args_->gathered = 0; /* a sharded pass switches to the gathered arrays after the exchange */
      const u32 bound = shard ? shard_bound(side, shard_world) : 0;
      h_args_->shard_bound = (int)bound;
      h_args_->track_m = 0; h_args_->cap_tail = 0; h_args_->p_min = 0;
      /* a sharded pass's second half is split by value range once the table has quantiles worth the name */
      const bool range = shard && use_range_ && nt_ >= range_min_;
      const u32 rbound = range ? range_bound(side, shard_world) : 0, rebound = range ? range_ebound_ : 0;
      h_args_->range_split = range ? 1 : 0;
      h_args_->range_bound = (int)rbound; h_args_->range_ebound = (int)rebound;
      h_args_->k1_spread = (throughput_mode_ || no_spread_) ? 0 : 1;
      /* tiny: the whole front of the pass in one single-block kernel (tiny_pass.cuh); force_small_post_ is the
         test hook that mispredicts, to exercise the rerun through the general kernels */
      const bool tiny = !shard && !no_tiny && !stage_events_ &&
                        (force_small_post_ || (use_tiny_ && small && prev_done_[side] <= SP_CAP / 3 && nt_ <= 200000u &&
                                               prev_nodes_[side] <= tiny_nodes_));
      CUDA_CHECK(cudaMemcpyAsync(d_passin_, h_passin_, PASSIN_FORMS_OFFSET + nforms * sizeof(FormDesc),
                                 cudaMemcpyHostToDevice, st_));
      was_tiny = tiny;
      /* small and mid-size passes: every expression position in ONE launch with grid barriers (k_expand_all); a
         lone search only -- the barrier needs the whole grid resident, which engines sharing the GPU cannot promise */
      persist_ = use_persist_ && !throughput_mode_ && !shard && prev_nodes_[side] <= persist_max_;
      const bool use_shared = !tiny && !shard && shared_wanted(using_x, cplx, side);
      was_shared = use_shared;
      if (use_shared) {
        const SharedPass sp = shared_get(cplx, forms, max_len);
        /* the build above may have used the pass's argument block: (re)send it */
        h_args_->using_x = 0; h_args_->cplx = cplx; h_args_->skip_exact = skip_exact ? 1 : 0;
        h_args_->nforms = (int)nforms; h_args_->best_match0 = best_match0;
        h_args_->cap_nodes = (u32)std::min<size_t>(nodes_.cap, 0xffffffffu);
        h_args_->cap_done = (u32)std::min<size_t>(done_.cap, 0xffffffffu);
        h_args_->cap_events = (u32)events_.cap; h_args_->seq = ++pass_seq_;
        h_args_->shard_rank = 0; h_args_->shard_world = 1; h_args_->gathered = 0; h_args_->shard_bound = 0;
        h_args_->track_m = 0; h_args_->cap_tail = 0; h_args_->p_min = 0;
        if (ptrs_dirty_) upload_ptrs();
        SharedView& V = *(SharedView*)(h_passin_ + PASSIN_VIEW_OFFSET);
        V.rows = sp.rows; V.forms = sp.forms; V.tail = sp.tail; V.n_slots = sp.n_slots; V.n_tail = sp.n_tail;
        V.node_count = sp.node_count; V.pad = 0; V.dead_base = sp.dead_base; V.p_t = cur_P_.p_ovr;
        CUDA_CHECK(cudaMemcpyAsync(d_passin_, h_passin_, PASSIN_FORMS_OFFSET + nforms * sizeof(FormDesc),
                                   cudaMemcpyHostToDevice, st_));
        if (host_prof_) hp_t1 = now_us();
        launch_pass(1, max_len, small); /* the shared front (control block, keys, sort), then stages 2 and 3 */
        if (host_prof_) hp_t2 = now_us();
        wait_pass();
        if (host_prof_) hp_t3 = now_us();
        collect_stamp_times();
        shared_->consumed++;
      } else if (tiny) {
        if (host_prof_) hp_t1 = now_us();
        launch_tiny();
        if (host_prof_) hp_t2 = now_us();
        wait_pass();
        if (host_prof_) hp_t3 = now_us();
        collect_stamp_times();
      } else if (!shard) {
        if (host_prof_) hp_t1 = now_us();
        if (stage_events_) launch_stages(0, 3, max_len, small, true);
        else launch_pass(0, max_len, small);
        if (host_prof_) hp_t2 = now_us();
        wait_pass();
        if (stage_events_) CUDA_CHECK(cudaStreamSynchronize(st_)); /* the last event must have completed */
        if (host_prof_) hp_t3 = now_us();
        if (stage_events_) collect_stage_times(0, 3); else collect_stamp_times();
      } else {
        /* one search over several GPUs: the front on this rank's forms, the fixed-size exchange, the replicated
           back half -- all enqueued at once, no host round trip in between (shard.cuh) */
        if (host_prof_) hp_t1 = now_us();
        sharded_front(max_len, small, bound, shard_world);
        if (range) range_back(max_len, small, rbound, rebound, shard_world);
        else { gather_hdr(); launch_stages(1, 3, max_len, small, false); }
        was_ranged = range;
        if (host_prof_) hp_t2 = now_us();
        wait_pass();
        if (host_prof_) hp_t3 = now_us();
        collect_stamp_times();
        if (gather_pending_) {
          CUDA_CHECK(cudaEventSynchronize(ev_g_[1]));
          float ms = 0;
          if (cudaEventElapsedTime(&ms, ev_g_[0], ev_g_[1]) == cudaSuccess) gather_ms += ms;
          (void)cudaGetLastError();
          gather_pending_ = false;
        }
        was_sharded = true;
        if (!h_ctl_->overflow) {
          sharded_passes++;
          gather_bytes += (unsigned long long)shard_world * bound * (sizeof(Row) + sizeof(uint16_t));
        }
      }
      const u32 ovf = h_ctl_->overflow;
      if (!ovf) break;
      /* grow what overflowed and rerun the pass; the live table buffer has not been touched */
      size_t want_nodes = nodes_.cap, want_done = done_.cap;
      /* a sharded pass reports the largest rank's sizes (every rank sees the same gathered headers) */
      if (ovf & OVF_NODES) want_nodes = std::max<size_t>(nodes_.cap * 2, (size_t)(was_sharded ? h_ctl_->max_nodes : h_ctl_->node_count) * 2);
      if (ovf & OVF_DONE) want_done = std::max<size_t>(done_.cap * 2, (size_t)(was_sharded ? h_ctl_->max_done : h_ctl_->done_count) * 2);
      if (ovf & (OVF_NODES | OVF_DONE)) grow_arena(want_nodes, want_done);
      if (ovf & OVF_KEPT) {
        shard_kept_[side] = std::max<u32>(h_ctl_->n_kept, shard_kept_[side] * 2);
        if (was_ranged) range_new_[side] = std::max<u32>(h_ctl_->max_new, range_new_[side] * 2 + 1024);
      }
      if ((ovf & OVF_EVENTS) && was_ranged) range_ebound_ = std::max<u32>(range_ebound_ * 2, h_ctl_->max_ev + 64);
      if (ovf & OVF_EVENTS) {
        if (events_.ensure((size_t)h_ctl_->n_events + 1024, st_)) ptrs_dirty_ = true;
        /* the merge kernels had already entered this pass's new keys into the hash set, but the table was not
           flipped: forget them, or the rerun's pre-dedupe would drop the very rows it has to insert */
        rebuild_hash(0);
        rebuild_hash(1);
      }
      if (ovf & OVF_SMALL) no_tiny = true;
    }
    const u32 nu = h_ctl_->n_uniq, ne = h_ctl_->n_events;
    const double hp_t4 = host_prof_ ? now_us() : 0;
    const unsigned long long gen = h_ctl_->gen;
    (void)gen_k1;
    if (pass_log_) {
      const float* ms = last_ms_;
      fprintf(stderr, "pass %s cplx %3d forms %5u len %2d nodes %9u done %9llu keys %9u new %8u table %9u events %5u | K1 %7.3f sort %7.3f merge %7.3f refine %6.3f ms\n",
              using_x ? (was_tiny ? "lhs" : "LHS") : (was_tiny ? "rhs" : "RHS"), cplx, nforms, max_len, h_ctl_->node_count, gen,
              was_tiny ? h_ctl_->done_count : h_ctl_->n_keys, nu, nt_, ne, ms[0], ms[1], ms[2], ms[3]);
    }
    /* sizes for the next pass of this side; a sharded pass sizes the arenas by its largest rank */
    if (was_shared) { /* the consumer of a shared pass only holds what survived its pre-dedupe */
      h_ctl_->done_count = std::max(h_ctl_->done_count, h_ctl_->n_keys);
    }
    prev_nodes_[side] = was_sharded ? h_ctl_->max_nodes : h_ctl_->node_count;
    prev_done_[side] = was_sharded ? h_ctl_->max_done : (u32)std::min<unsigned long long>(std::max<unsigned long long>(gen, h_ctl_->done_count), 0xffffffffull);
    prev_gen_[side] = (u32)std::min<unsigned long long>(gen, 0xffffffffull);
    prev_keys_[side] = h_ctl_->n_keys;
    shard_kept_[side] = was_sharded ? h_ctl_->n_kept : 0;
    range_new_[side] = was_ranged ? h_ctl_->max_new : 0;
    prev_new_[side] = nu;
    side_rows_[side] += nu;
    expand_nodes += h_ctl_->node_count;
    sort_keys += was_tiny ? h_ctl_->done_count : h_ctl_->n_keys;
    out.generated = (double)gen;
    out.dead_ends = (double)h_ctl_->dead;
    out.inserted = (long)nu;
    last_nu_ = nu;
    last_tiny_ = was_tiny;
    last_using_x_ = using_x != 0;
    if (nu) {
      merge_rows += (unsigned long long)nt_ + nu;
      cur_ ^= 1;
      nt_ += nu;
    }
    if (ne) {
      if (ne <= HEAD_EVENTS) {
        out.events.assign(h_head_, h_head_ + ne);
      } else {
        if (ne > h_events_cap_) {
          if (h_events_) cudaFreeHost(h_events_);
          h_events_cap_ = std::max<size_t>(ne, h_events_cap_ * 2);
          CUDA_CHECK(cudaMallocHost(&h_events_, h_events_cap_ * sizeof(MatchEvent)));
        }
        CUDA_CHECK(cudaMemcpyAsync(h_events_, events_.p, ne * sizeof(MatchEvent), cudaMemcpyDeviceToHost, st_));
        CUDA_CHECK(cudaStreamSynchronize(st_));
        d2h_bytes += (unsigned long long)ne * sizeof(MatchEvent);
        out.events.assign(h_events_, h_events_ + ne);
      }
      /* replay order = insertion order of the NEW expression, then exact < previous < next (check_sides) */
      if (was_tiny) {
        std::sort(out.events.begin(), out.events.end(), [](const MatchEvent& a, const MatchEvent& b) {
          return a.seq != b.seq ? a.seq < b.seq : a.kind < b.kind;
        });
      } else {
        const bool ux = using_x != 0;
        std::sort(out.events.begin(), out.events.end(), [ux](const MatchEvent& a, const MatchEvent& b) {
          const Sym128& sa = ux ? a.lhs.sym : a.rhs.sym;
          const Sym128& sb = ux ? b.lhs.sym : b.rhs.sym;
          if (a.seq != b.seq || sa.lo != sb.lo || sa.hi != sb.hi) return seq_before_host(a.seq, sa, b.seq, sb);
          return a.kind < b.kind;
        });
      }
    }
    if (host_prof_ && hp_t1 > 0) {
      hp_[0] += hp_t1 - hp_t0; hp_[1] += hp_t2 - hp_t1; hp_[2] += hp_t3 - hp_t2; hp_[3] += hp_t4 - hp_t3;
      hp_[4] += now_us() - hp_t4; hp_[5] += 1;
    }
  }

  unsigned long long kernel_launches = 0;
  unsigned long long h2d_bytes = 0, d2h_bytes = 0;
  /* device time per stage, CUDA events on the launching stream (ms): 0 K1 enumerate+evaluate,
     1 order+radix sort, 2 classify/compact/match/merge, 3 refine */
  double stage_ms[4] = {0, 0, 0, 0};
  unsigned long long stage_launches[4] = {0, 0, 0, 0};
  unsigned long long sort_keys = 0;      /* survivors that went through the radix sort */
  unsigned long long merge_rows = 0;     /* rows written by the merges (old + new) */
  unsigned long long expand_nodes = 0;   /* live prefixes expanded (= exec'd prefixes) */
  unsigned long long tiny_passes = 0;    /* passes that went through the three-launch tiny path */
  unsigned long long sharded_passes = 0; /* passes whose K1 ran sharded over the ranks */
  unsigned long long gather_bytes = 0;   /* bytes this rank received/kept in the all-gathers */
  double gather_ms = 0;                  /* device time of the all-gathers (CUDA events) */
  u32 table_rows() const { return nt_; }
  unsigned char sym_chars_[RIES_MAX_SYMS];
  void read_opcounts(unsigned long long* out /* [2][RIES_MAX_SYMS] */) {
    CUDA_CHECK(cudaMemcpyAsync(out, d_opcnt_, 2 * RIES_MAX_SYMS * sizeof(u64), cudaMemcpyDeviceToHost, st_));
    CUDA_CHECK(cudaStreamSynchronize(st_));
  }
  double dfma_peak_tflops() {
    const int blocks = num_sms_ * 8, threads = 512, iters = 1 << 14;
    double* d = 0;
    CUDA_CHECK(cudaMalloc(&d, (size_t)blocks * threads * sizeof(double)));
    cudaEvent_t a, b;
    CUDA_CHECK(cudaEventCreate(&a)); CUDA_CHECK(cudaEventCreate(&b));
    double best = 0;
    for (int rep = 0; rep < 5; rep++) {
      CUDA_CHECK(cudaEventRecord(a, st_));
      k_dfma_peak<<<blocks, threads, 0, st_>>>(d, iters);
      CUDA_CHECK(cudaEventRecord(b, st_));
      CUDA_CHECK(cudaEventSynchronize(b));
      float ms = 0;
      CUDA_CHECK(cudaEventElapsedTime(&ms, a, b));
      double tf = 2.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
      if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(d);
    return best;
  }

 private:
  /* The fixed launch sequence of one pass, in four stages.  Each stage is issued directly or as a CUDA
     graph captured once (stage 0 once per expression length); the events between stages are recorded
     outside the graphs so that per-stage device time can be read on this stream. */
  void enqueue_stage(int stage, int max_len, int small) {
    /* every kernel loops over its work (ticket / grid-stride), so the grid only sets the parallelism:
       the "small" variant is for passes predicted to be tiny, where launching a thousand blocks that
       exit at once costs more than the work itself.  A wrong prediction is slow, never wrong. */
    /* `small`: 0 full grids; 1 a pass predicted to be small; 2, 3 (throughput mode only, see run_pass): mid-size
       passes on a sixteenth / a quarter of the machine */
    const u32 div = small == 2 ? 16u : small == 3 ? 4u : 1u;
    const u32 g_exp = small == 1 ? small_grid_ : std::max(24u, expand_grid_ / div);
    const u32 g_wide = small == 1 ? 12u : std::max(24u, wide_grid_ / div);
    const u32 g_keys = small == 1 ? 12u : std::max(16u, (u32)num_sms_ * 4 / div);
    const u32 g_sort = small == 1 ? 12u : std::max(12u, (u32)num_sms_ * RS_MIN_BLOCKS / div);
    switch (stage) {
      case 0:
        k_pass_begin<<<1, 1024, 0, st_>>>(d_ptrs_, d_args_, d_ctl_, d_tab_);
        if (persist_) { /* every expression position in one cooperative launch (k_expand_all) */
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3(std::min(g_exp, expand_grid_)); cfg.blockDim = dim3(XF_THREADS);
          cfg.dynamicSmemBytes = xf_dyn_smem(k1_nq_); cfg.stream = st_;
          cudaLaunchAttribute at[1];
          at[0].id = cudaLaunchAttributeCooperative;
          at[0].val.cooperative = 1;
          cfg.attrs = at; cfg.numAttrs = 1;
          const DevPtrs* a0 = d_ptrs_; const PassArgs* a1 = d_args_; PassCtl* a2 = d_ctl_;
          CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_expand_all, a0, a1, a2));
          break;
        }
        for (int lvl = 0; lvl < max_len; lvl++) {
          /* positions after the first may overlap their prologue with the previous position's tail
             (programmatic dependent launch; the kernel waits on the dependency itself) */
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3(g_exp); cfg.blockDim = dim3(XF_THREADS); cfg.dynamicSmemBytes = xf_dyn_smem(k1_nq_); cfg.stream = st_;
          cudaLaunchAttribute at[1];
          at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
          at[0].val.programmaticStreamSerializationAllowed = 1;
          cfg.attrs = at; cfg.numAttrs = (use_pdl_ && lvl > 0) ? 1 : 0;
          const DevPtrs* a0 = d_ptrs_; const PassArgs* a1 = d_args_; PassCtl* a2 = d_ctl_;
          CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_expand, a0, a1, a2));
        }
        break;
      case 1:
        k_keys<<<g_keys, KK_THREADS, 0, st_>>>(d_ptrs_, d_args_, d_ctl_);
        for (int pass = 0; pass < 8; pass++)
          k_rs_onesweep<<<g_sort, RS_THREADS, RS_DYN_SMEM, st_>>>(d_ptrs_, d_ctl_, d_tab_, pass);
        break;
      case 2:
        k_classify_compact<<<g_wide, CC_THREADS, 0, st_>>>(d_ptrs_, d_args_, d_ctl_, d_tab_);
        k_match_merge_new<<<g_wide, 256, 0, st_>>>(d_ptrs_, d_args_, d_ctl_, d_tab_, 0);
        k_merge_old<<<small == 1 ? 48u : std::max(48u, wide_grid_ / div), MO_THREADS, 0, st_>>>(d_ptrs_, d_ctl_, d_tab_);
        break;
      case 3:
        k_refine<<<small ? 8 : 64, 64, 0, st_>>>(d_ptrs_, d_args_, d_ctl_);
        k_pass_end<<<1, 128, 0, st_>>>(d_ptrs_, d_args_, d_ctl_, d_tab_);
        break;
      default: { /* 4: a shared RHS pass's front (instead of stages 0 and 1): the view rides in the pass's argument block */
        const SharedView* vp = (const SharedView*)(d_passin_ + PASSIN_VIEW_OFFSET);
        k_shared_front<<<1, 256, 0, st_>>>(d_ptrs_, d_args_, d_ctl_, d_tab_, vp);
        k_keys_shared<<<g_keys, KK_THREADS, 0, st_>>>(d_ptrs_, d_args_, d_ctl_, vp);
        for (int pass = 0; pass < 8; pass++)
          k_rs_onesweep<<<g_sort, RS_THREADS, RS_DYN_SMEM, st_>>>(d_ptrs_, d_ctl_, d_tab_, pass);
        break;
      }
    }
  }

  /* mirrors what k_pass_begin / k_stage1_reset do to TableState::epoch */
  void bump_epoch() {
    epoch_host_ = (epoch_host_ + 16) & 0x3ffffff0u;
    if (epoch_host_ == 0) {
      /* 2^26 bumps later the tags wrap: clear the look-back state so old tags cannot alias new ones */
      epoch_host_ = 16;
      CUDA_CHECK(cudaMemsetAsync(rs_state_.p, 0, rs_state_.cap * sizeof(u64), st_));
      CUDA_CHECK(cudaMemsetAsync(cc_state_.p, 0, cc_state_.cap * sizeof(u64), st_));
    }
  }

  /* Stages [first, last] of a pass on the stream.  Per-stage device time comes from the device clock stamps
     the kernels leave in the control block (collect_stamp_times); only sharded passes, whose stages are
     interleaved with the exchange, bracket them with CUDA events (ev_[k] before stage k and after the last). */
  void count_stage(int stage, int max_len) {
    const unsigned long long k = (stage == 0) ? (persist_ ? 2ull : (unsigned long long)max_len + 1) : (stage == 1) ? 9 : (stage == 2) ? 3 : (stage == 3) ? 2 : 10;
    kernel_launches += k;
    if (stage == 4) { stage_launches[0] += 1; stage_launches[1] += 9; }
    else stage_launches[stage] += (stage == 0) ? k - 1 : k; /* stage 0: the K1 launches, not k_pass_begin */
  }

  /* A whole pass as ONE graph launch (per-stage device times come from the kernels' clock stamps, so nothing has to
     be recorded between the stages): kind 0 the general pass (stages 0-3), 1 the consumer of a shared RHS pass
     (stages 4, 2, 3), 2 a tiny pass.  With the per-pass argument copy that is two driver calls per pass -- what
     a batch of concurrent engines, which serialise on the driver, is bound by. */
  void launch_pass(int kind, int max_len, int small) {
    bump_epoch();
    static const int seq0[] = {0, 1, 2, 3}, seq1[] = {4, 2, 3};
    const int* seq = kind == 0 ? seq0 : seq1;
    const int nseq = kind == 0 ? 4 : kind == 1 ? 3 : 0;
    auto enqueue = [&]() {
      if (kind == 2) {
        k_tiny_front<<<1, TINY_THREADS, TINY_DYN_SMEM, st_>>>(d_ptrs_, d_args_, d_ctl_, d_tab_, tiny_frontier_);
        k_merge_old<<<48u, MO_THREADS, 0, st_>>>(d_ptrs_, d_ctl_, d_tab_);
        k_refine_end<<<1, 128, 0, st_>>>(d_ptrs_, d_args_, d_ctl_, d_tab_);
      } else {
        for (int i = 0; i < nseq; i++) enqueue_stage(seq[i], max_len, small);
      }
    };
    if (!use_graphs_) {
      enqueue();
    } else {
      const int key = 100000 + kind * 10000 + (kind == 0 ? (persist_ ? 30 : max_len) : 0) * 4 + small;
      auto it = graphs_.find(key);
      if (it == graphs_.end()) {
        cudaGraph_t g;
        CUDA_CHECK(cudaStreamBeginCapture(st_, cudaStreamCaptureModeThreadLocal));
        enqueue();
        CUDA_CHECK(cudaStreamEndCapture(st_, &g));
        cudaGraphExec_t ge;
        CUDA_CHECK(cudaGraphInstantiate(&ge, g, 0));
        CUDA_CHECK(cudaGraphDestroy(g));
        it = graphs_.emplace(key, ge).first;
      }
      CUDA_CHECK(cudaGraphLaunch(it->second, st_));
    }
    if (kind == 2) {
      kernel_launches += 3;
      stage_launches[0] += 1; stage_launches[2] += 1; stage_launches[3] += 1;
      tiny_passes++;
    } else {
      for (int i = 0; i < nseq; i++) count_stage(seq[i], max_len);
    }
  }

  void launch_stages(int first, int last, int max_len, int small, bool events = false) {
    if (first == 0 || first == 4) bump_epoch();
    if (events) CUDA_CHECK(cudaEventRecord(ev_[first], st_));
    for (int stage = first; stage <= last; stage++) {
      if (!use_graphs_) {
        enqueue_stage(stage, max_len, small);
      } else {
        const int key = ((stage == 0) ? (persist_ ? 30 : max_len) : -stage) * 4 + small + 1000;
        auto it = graphs_.find(key);
        if (it == graphs_.end()) {
          cudaGraph_t g;
          CUDA_CHECK(cudaStreamBeginCapture(st_, cudaStreamCaptureModeThreadLocal));
          enqueue_stage(stage, max_len, small);
          CUDA_CHECK(cudaStreamEndCapture(st_, &g));
          cudaGraphExec_t ge;
          CUDA_CHECK(cudaGraphInstantiate(&ge, g, 0));
          CUDA_CHECK(cudaGraphDestroy(g));
          it = graphs_.emplace(key, ge).first;
        }
        CUDA_CHECK(cudaGraphLaunch(it->second, st_));
      }
      if (events) CUDA_CHECK(cudaEventRecord(ev_[stage + 1], st_));
      count_stage(stage, max_len);
    }
  }

  /* Results of the pass just enqueued.  The last kernel of a pass writes them into the mapped host block and
     then posts the pass number; the host polls for it (no copy operation, no stream synchronisation), looking
     at the stream now and then so that a faulted kernel is reported instead of waited for. */
  void wait_pass() {
    if (poll_host_) {
      volatile u32* seq = &h_out_->seq;
      const u32 want = pass_seq_;
      double t_wait0 = 0;
      for (unsigned spins = 1;; spins++) {
        if (*seq == want) break;
        if ((spins & 0xffffffu) == 0) { /* watchdog: a pass that posts nothing for minutes is a hung kernel, not work */
          const double now = now_us();
          if (t_wait0 == 0) t_wait0 = now;
          else if (now - t_wait0 > watchdog_s_ * 1e6) {
            fprintf(stderr, "ries_b200: pass %u has not completed after %.0f s (device hung?); giving up\n", want, watchdog_s_);
            { /* what the device says about it, read on another stream (the engine's is stuck behind the kernel) */
              cudaStream_t s2;
              PassCtl c;
              if (cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking) == cudaSuccess &&
                  cudaMemcpyAsync(&c, d_ctl_, sizeof(c), cudaMemcpyDeviceToHost, s2) == cudaSuccess &&
                  cudaStreamSynchronize(s2) == cudaSuccess)
                fprintf(stderr, "  control block: lvl [%u, %u) nodes %u done %u ticket %u blocks_done %u keys %u uniq %u events %u overflow %u "
                                "chunk_ticket %u stamps %llu %llu %llu %llu %llu\n", c.lvl_begin, c.lvl_end, c.node_count, c.done_count,
                        c.tile_ticket, c.blocks_done, c.n_keys, c.n_uniq, c.n_events, c.overflow, c.chunk_ticket,
                        (unsigned long long)c.stamp[0], (unsigned long long)(c.stamp[1] ? c.stamp[1] - c.stamp[0] : 0),
                        (unsigned long long)(c.stamp[2] ? c.stamp[2] - c.stamp[0] : 0), (unsigned long long)(c.stamp[3] ? c.stamp[3] - c.stamp[0] : 0),
                        (unsigned long long)(c.stamp[4] ? c.stamp[4] - c.stamp[0] : 0));
            }
            abort();
          }
        }
        /* engines of a batch: more pollers than there may be free cores -- let whoever has work to enqueue run */
        if (throughput_mode_ && (spins & 0xffu) == 0) sched_yield();
        if ((spins & 0x3ffffu) == 0) { /* every few milliseconds: a driver call takes the lock the launching threads of a batch need */
          const cudaError_t q = cudaStreamQuery(st_);
          if (q == cudaSuccess) {
            if (*seq == want) break;
            /* the stream is idle and nothing was posted: fall back to the explicit copy */
            CUDA_CHECK(cudaMemcpyAsync(h_out_, d_out_, offsetof(PassOut, seq), cudaMemcpyDeviceToHost, st_));
            CUDA_CHECK(cudaStreamSynchronize(st_));
            break;
          }
          if (q != cudaErrorNotReady) CUDA_CHECK(q);
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
      }
      __sync_synchronize();
      u32 ne = h_ctl_->n_events;
      if (ne > HEAD_EVENTS) ne = HEAD_EVENTS;
      d2h_bytes += offsetof(PassOut, head) + (unsigned long long)ne * sizeof(MatchEvent) + 4;
    } else {
      CUDA_CHECK(cudaMemcpyAsync(h_out_, d_out_, offsetof(PassOut, seq), cudaMemcpyDeviceToHost, st_));
      CUDA_CHECK(cudaStreamSynchronize(st_));
      CUDA_CHECK(cudaGetLastError());
      d2h_bytes += offsetof(PassOut, seq);
    }
  }

  /* A pass predicted tiny: three direct launches (tiny_pass.cuh). */
  void launch_tiny() { launch_pass(2, 0, 0); }

  /* device time per stage of the pass just synchronised, from the stamps in the copied-back control block */
  void collect_stamp_times() {
    const u64* t = h_ctl_->stamp;
    for (int i = 0; i < 4; i++) {
      last_ms_[i] = 0;
      if (!t[i]) continue;
      int j = i + 1;
      while (j < 4 && !t[j]) j++; /* a stage that never started (overflow) leaves no stamp */
      if (t[j] >= t[i]) last_ms_[i] = (float)((double)(t[j] - t[i]) * 1e-6);
      stage_ms[i] += last_ms_[i];
    }
  }

  void collect_stage_times(int first, int last) {
    for (int i = first; i <= last; i++) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, ev_[i], ev_[i + 1]) == cudaSuccess) { stage_ms[i] += ms; last_ms_[i] = ms; }
    }
    (void)cudaGetLastError();
  }

  /* Is the coming RHS pass taken from the batch's shared enumeration?  Shared from the first complexity at which
     some engine of the batch finds a pass worth it (a tiny pass costs less than the bookkeeping). */
  bool shared_wanted(int using_x, int cplx, int side) {
    if (!shared_ok_ || using_x) return false;
    std::lock_guard<std::mutex> lock(shared_->mu);
    if (shared_->c_share < 0 && prev_gen_[side] >= 2048) shared_->c_share = cplx;
    return shared_->c_share >= 0 && cplx >= shared_->c_share;
  }

  /* The shared RHS pass of complexity `cplx`; the first engine to ask enumerates it (K1 with the batch's largest
     overflow bound, M carried in `der`, bound-dependent records) and keeps a copy for the whole batch. */
  SharedPass shared_get(int cplx, const std::vector<FormDesc>& forms, int max_len) {
    std::lock_guard<std::mutex> lock(shared_->mu);
    auto it = shared_->passes.find(cplx);
    if (it != shared_->passes.end()) return it->second;
    const u32 nforms = (u32)forms.size();
    /* sizes from the previous shared pass (RHS passes grow ~1.3x per complexity step) */
    size_t want_nodes = nodes_.cap, want_done = done_.cap, want_tail = std::max<size_t>(tail_.cap, 1u << 16);
    auto pv = shared_->passes.find(cplx - 1);
    if (pv != shared_->passes.end()) {
      want_nodes = std::max<size_t>(want_nodes, (size_t)pv->second.node_count * 2);
      want_done = std::max<size_t>(want_done, (size_t)pv->second.n_slots * 2);
      want_tail = std::max<size_t>(want_tail, (size_t)pv->second.n_tail * 2 + 4096);
    }
    RiesParams P2 = cur_P_;
    P2.p_ovr = shared_->p_max; P2.n_ovr = -shared_->p_max;
    CUDA_CHECK(cudaMemcpyAsync(d_P_, &P2, sizeof(P2), cudaMemcpyHostToDevice, st_));
    CUDA_CHECK(cudaStreamSynchronize(st_));
    PassCtl c;
    while (true) {
      if (want_nodes > nodes_.cap || want_done > done_.cap) grow_arena(std::max(want_nodes, nodes_.cap), std::max(want_done, done_.cap));
      if (tail_.ensure(want_tail, st_)) ptrs_dirty_ = true;
      if (ptrs_dirty_) upload_ptrs();
      h_args_->using_x = 0; h_args_->cplx = cplx; h_args_->skip_exact = 0; h_args_->nforms = (int)nforms; h_args_->best_match0 = 0;
      h_args_->cap_nodes = (u32)std::min<size_t>(nodes_.cap, 0xffffffffu);
      h_args_->cap_done = (u32)std::min<size_t>(done_.cap, 0xffffffffu);
      h_args_->cap_events = (u32)events_.cap; h_args_->seq = 0;
      h_args_->shard_rank = 0; h_args_->shard_world = 1; h_args_->gathered = 0; h_args_->shard_bound = 0;
      h_args_->track_m = 1; h_args_->cap_tail = (u32)tail_.cap; h_args_->p_min = shared_->p_min;
      CUDA_CHECK(cudaMemcpyAsync(d_passin_, h_passin_, PASSIN_FORMS_OFFSET + nforms * sizeof(FormDesc), cudaMemcpyHostToDevice, st_));
      launch_stages(0, 0, max_len, 0, false);
      CUDA_CHECK(cudaMemcpyAsync(&c, d_ctl_, sizeof(PassCtl), cudaMemcpyDeviceToHost, st_));
      CUDA_CHECK(cudaStreamSynchronize(st_));
      if (!c.overflow) break;
      if (c.overflow & OVF_NODES) want_nodes = std::max<size_t>(nodes_.cap * 2, (size_t)c.node_count * 2);
      if (c.overflow & OVF_DONE) want_done = std::max<size_t>(done_.cap * 2, (size_t)c.done_count * 2);
      if (c.overflow & OVF_TAIL) want_tail = std::max<size_t>(tail_.cap * 2, (size_t)c.n_tail * 2);
    }
    SharedPass sp;
    sp.n_slots = c.done_count; sp.n_tail = c.n_tail; sp.node_count = c.node_count; sp.dead_base = c.dead; sp.gen = c.gen;
    /* stream-ordered allocations: cudaMalloc / cudaFree would synchronise the whole device, i.e. every engine of the batch */
    CUDA_CHECK(cudaMallocAsync(&sp.rows, std::max<size_t>(sp.n_slots, 1) * sizeof(Row), st_));
    CUDA_CHECK(cudaMallocAsync(&sp.forms, std::max<size_t>(sp.n_slots, 1) * sizeof(uint16_t), st_));
    CUDA_CHECK(cudaMallocAsync(&sp.tail, std::max<size_t>(sp.n_tail, 1) * sizeof(TailRec), st_));
    if (sp.n_slots) {
      CUDA_CHECK(cudaMemcpyAsync(sp.rows, done_.p, (size_t)sp.n_slots * sizeof(Row), cudaMemcpyDeviceToDevice, st_));
      CUDA_CHECK(cudaMemcpyAsync(sp.forms, done_form_.p, (size_t)sp.n_slots * sizeof(uint16_t), cudaMemcpyDeviceToDevice, st_));
    }
    if (sp.n_tail) CUDA_CHECK(cudaMemcpyAsync(sp.tail, tail_.p, (size_t)sp.n_tail * sizeof(TailRec), cudaMemcpyDeviceToDevice, st_));
    CUDA_CHECK(cudaMemcpyAsync(d_P_, &cur_P_, sizeof(RiesParams), cudaMemcpyHostToDevice, st_)); /* this target's own bound again */
    CUDA_CHECK(cudaStreamSynchronize(st_)); /* other engines read the copy from their own streams */
    shared_->passes.emplace(cplx, sp);
    shared_->built++;
    return sp;
  }

  /* Rows per rank in the exchange of the coming sharded pass: from what the largest rank kept in the side's last
     sharded pass (same-side passes grow ~1.33x), or, for the first sharded pass of a side, from the keys of its
     last unsharded one.  Too small is not wrong: the pass is rerun (OVF_KEPT). */
  u32 shard_bound(int side, int world) {
    size_t est = shard_kept_[side] ? (size_t)shard_kept_[side] * 3 / 2
                                   : ((size_t)prev_keys_[side] * 2 / (size_t)world) * 3 / 2;
    est = std::max<size_t>(est, 4096);
    est = (est + 1023) & ~(size_t)1023;
    return (u32)std::min<size_t>(est, 0x7fffffffu / (size_t)world);
  }

  /* Front half of a pass sharded over `world` ranks (or virtual ranks): K1 on this rank's forms, the hash
     pre-dedupe into the send slot, the fixed-size all-gather of the slots, the totals into the control block
     (shard.cuh).  Nothing here waits for the device: the overflow flags of all ranks arrive in the gathered
     headers and end up in the control block the host reads at the end of the pass, so every rank takes the same
     decision. */
  void sharded_front(int max_len, int small, u32 bound, int world) {
    const DistState& ds = dist_state();
    const bool fake = ds.fake_world > 1;
    bool ch = false;
    ch |= ksend_.ensure(bound, st_);
    ch |= ksend_form_.ensure(bound, st_);
    ch |= gdone_.ensure((size_t)world * bound, st_);
    ch |= gdone_form_.ensure((size_t)world * bound, st_);
    /* the replicated half sorts up to world x bound keys */
    if ((size_t)world * bound > done_.cap) { grow_arena(nodes_.cap, (size_t)world * bound); ch = true; }
    ensure_table((size_t)nt_ + (size_t)world * bound);
    if (ch || ptrs_dirty_) { ptrs_dirty_ = true; upload_ptrs(); }
    const u32 g_keys = small ? 12u : (u32)num_sms_ * 4;
    auto local_front = [&]() {
      launch_stages(0, 0, max_len, small, false);
      k_keys_send<<<g_keys, KK_THREADS, 0, st_>>>(d_ptrs_, d_args_, d_ctl_);
      k_send_hdr<<<1, 32, 0, st_>>>(d_ptrs_, d_args_, d_ctl_);
      kernel_launches += 2;
      stage_launches[1] += 2;
    };
    if (fake) {
      /* test hook: the virtual ranks run one after the other on this GPU; the "all-gather" is three copies */
      for (int r = 0; r < world; r++) {
        CUDA_CHECK(cudaStreamSynchronize(st_)); /* the pinned arguments are about to change */
        h_args_->shard_rank = r;
        CUDA_CHECK(cudaMemcpyAsync(d_passin_, h_passin_, PASSIN_FORMS_OFFSET, cudaMemcpyHostToDevice, st_));
        local_front();
        CUDA_CHECK(cudaMemcpyAsync(gdone_.p + (size_t)r * bound, ksend_.p, (size_t)bound * sizeof(Row), cudaMemcpyDeviceToDevice, st_));
        CUDA_CHECK(cudaMemcpyAsync(gdone_form_.p + (size_t)r * bound, ksend_form_.p, (size_t)bound * sizeof(uint16_t), cudaMemcpyDeviceToDevice, st_));
        CUDA_CHECK(cudaMemcpyAsync(d_ghdr_ + r, d_shdr_, sizeof(SlotHdr), cudaMemcpyDeviceToDevice, st_));
      }
      CUDA_CHECK(cudaStreamSynchronize(st_));
    } else {
      local_front();
      CUDA_CHECK(cudaEventRecord(ev_g_[0], st_));
      dist_allgather3(d_shdr_, d_ghdr_, sizeof(SlotHdr), ksend_.p, gdone_.p, (size_t)bound * sizeof(Row),
                      ksend_form_.p, gdone_form_.p, (size_t)bound * sizeof(uint16_t), st_);
      CUDA_CHECK(cudaEventRecord(ev_g_[1], st_));
      gather_pending_ = true;
    }
    /* from here on the kernels read the gathered slots (second copy of the pass's scalars: the first may not
       have been consumed yet) */
    memcpy(h_args2_, h_args_, sizeof(PassArgs));
    h_args2_->gathered = 1;
    CUDA_CHECK(cudaMemcpyAsync(d_passin_, h_args2_, sizeof(PassArgs), cudaMemcpyHostToDevice, st_));
  }

  void gather_hdr() {
    bump_epoch();
    k_gather_hdr<<<1, 256, 0, st_>>>(d_ptrs_, d_args_, d_ctl_, d_tab_);
    kernel_launches += 1;
    stage_launches[1] += 1;
  }

  /* rows per rank in the second exchange of a range-split pass: from the largest range's new rows in the side's last
     such pass, else from the side's last pass as a whole */
  u32 range_bound(int side, int world) {
    size_t est = range_new_[side] ? (size_t)range_new_[side] * 3 / 2 : ((size_t)prev_new_[side] * 2 / (size_t)world) * 3 / 2;
    est = std::max<size_t>(est, 4096);
    return (u32)((est + 1023) & ~(size_t)1023);
  }

  /* Second half of a sharded pass split by value range (shard.cuh): keys + sort + classify + neighbour tests on
     this rank's range of the gathered rows, the second (small) exchange, then the replicated merge, refinement and
     end of the pass.  Enqueued at once like the rest. */
  void range_back(int max_len, int small, u32 rbound, u32 ebound, int world) {
    const DistState& ds = dist_state();
    const bool fake = ds.fake_world > 1;
    bool ch = false;
    ch |= g2_idx_.ensure((size_t)world * rbound, st_);
    ch |= g2_gap_.ensure((size_t)world * rbound, st_);
    ch |= g2_events_.ensure((size_t)world * ebound, st_);
    if (events_.cap < ebound) ch |= events_.ensure(ebound, st_);
    if (ch || ptrs_dirty_) { ptrs_dirty_ = true; upload_ptrs(); }
    const u32 g_wide = small ? 12u : wide_grid_;
    auto local_back = [&]() {
      gather_hdr();
      launch_stages(1, 1, max_len, small, false);
      k_classify_compact<<<g_wide, CC_THREADS, 0, st_>>>(d_ptrs_, d_args_, d_ctl_, d_tab_);
      k_match_merge_new<<<g_wide, 256, 0, st_>>>(d_ptrs_, d_args_, d_ctl_, d_tab_, 1);
      k_send2_hdr<<<1, 32, 0, st_>>>(d_ptrs_, d_args_, d_ctl_);
      kernel_launches += 3;
      stage_launches[2] += 3;
    };
    if (fake) {
      for (int r = 0; r < world; r++) {
        CUDA_CHECK(cudaStreamSynchronize(st_)); /* the pinned arguments are about to change */
        h_args2_->shard_rank = r;
        CUDA_CHECK(cudaMemcpyAsync(d_passin_, h_args2_, sizeof(PassArgs), cudaMemcpyHostToDevice, st_));
        local_back();
        CUDA_CHECK(cudaMemcpyAsync(g2_idx_.p + (size_t)r * rbound, uniq_seq_.p, (size_t)rbound * sizeof(u32), cudaMemcpyDeviceToDevice, st_));
        CUDA_CHECK(cudaMemcpyAsync(g2_gap_.p + (size_t)r * rbound, uniq_gap_.p, (size_t)rbound * sizeof(u32), cudaMemcpyDeviceToDevice, st_));
        CUDA_CHECK(cudaMemcpyAsync(g2_events_.p + (size_t)r * ebound, events_.p, (size_t)ebound * sizeof(MatchEvent), cudaMemcpyDeviceToDevice, st_));
        CUDA_CHECK(cudaMemcpyAsync(d_ghdr2_ + r, d_shdr_, sizeof(SlotHdr), cudaMemcpyDeviceToDevice, st_));
      }
      CUDA_CHECK(cudaStreamSynchronize(st_));
    } else {
      local_back();
      dist_allgather4(d_shdr_, d_ghdr2_, sizeof(SlotHdr), uniq_seq_.p, g2_idx_.p, (size_t)rbound * sizeof(u32),
                      uniq_gap_.p, g2_gap_.p, (size_t)rbound * sizeof(u32), events_.p, g2_events_.p,
                      (size_t)ebound * sizeof(MatchEvent), st_);
      gather_bytes += (unsigned long long)world * ((size_t)rbound * 8 + (size_t)ebound * sizeof(MatchEvent));
    }
    k_gather2<<<g_wide, 256, 0, st_>>>(d_ptrs_, d_args_, d_ctl_);
    k_match_merge_new<<<g_wide, 256, 0, st_>>>(d_ptrs_, d_args_, d_ctl_, d_tab_, 2);
    k_merge_old<<<small ? 48u : wide_grid_, MO_THREADS, 0, st_>>>(d_ptrs_, d_ctl_, d_tab_);
    kernel_launches += 3;
    stage_launches[2] += 3;
    launch_stages(3, 3, max_len, small, false);
  }

  /* The hash sets restart small with every search and grow with the side's row count: at most half full after
     the coming pass (estimated from the side's last pass; an underestimate only costs filtering), a quarter
     full right after growing.  Growing means: clear, then re-enter the side's keys from the table. */
  void reset_hash() {
    for (int sd = 0; sd < 2; sd++) {
      hset_log2_[sd] = 12;
      if (hset_[sd].ensure((size_t)4 << hset_log2_[sd], st_)) ptrs_dirty_ = true;
      CUDA_CHECK(cudaMemsetAsync(hset_[sd].p, 0, ((size_t)4 << hset_log2_[sd]) * sizeof(u64), st_));
      side_rows_[sd] = 0;
      prev_new_[sd] = 0;
    }
    ptrs_dirty_ = true;
  }
  void rebuild_hash(int sd) {
    if (ptrs_dirty_) upload_ptrs();
    CUDA_CHECK(cudaMemsetAsync(hset_[sd].p, 0, ((size_t)4 << hset_log2_[sd]) * sizeof(u64), st_));
    if (nt_) k_hset_rebuild<<<(u32)num_sms_ * 8, 256, 0, st_>>>(d_ptrs_, d_tab_, sd);
    kernel_launches += nt_ ? 1 : 0;
  }
  void ensure_hash(int sd) {
    const size_t expect = side_rows_[sd] + 3 * (size_t)prev_new_[sd] + 4096;
    if (expect * 2 <= ((size_t)4 << hset_log2_[sd])) return;
    u32 lg = hset_log2_[sd];
    while (expect * 4 > ((size_t)4 << lg) && lg < 30) lg++;
    hset_log2_[sd] = lg;
    hset_[sd].ensure((size_t)4 << lg, st_);
    ptrs_dirty_ = true;
    rebuild_hash(sd);
  }

  /* nodes / done arenas and everything sized by them */
  void grow_arena(size_t want_nodes, size_t want_done) {
    bool ch = false;
    ch |= nodes_.ensure(want_nodes, st_);
    ch |= done_.ensure(want_done, st_);
    ch |= done_form_.ensure(done_.cap, st_);
    for (int b = 0; b < 2; b++) { ch |= keys_[b].ensure(done_.cap, st_); ch |= idx_[b].ensure(done_.cap, st_); }
    ch |= uniq_seq_.ensure(done_.cap, st_);
    ch |= uniq_gap_.ensure(done_.cap + 16, st_); /* + slack: bulk copies read whole 16-byte granules */
    ch |= uniq_sk_.ensure(done_.cap, st_);
    /* look-back state must start as zeros (stale bits must not look like a live epoch / status) */
    ch |= rs_state_.ensure((done_.cap / RS_CHUNK + 2) * 256, st_, false, true);
    ch |= cc_state_.ensure(done_.cap / 256 + 2, st_, false, true); /* one word per chunk of k_classify_compact (2048 keys); sized for chunks of 256 */
    if (ch) ptrs_dirty_ = true;
  }
  void ensure_table(size_t rows) {
    /* the live buffer keeps its contents; the other one is the merge destination */
    if (heap_.ensure(rows, st_, true)) ptrs_dirty_ = true;
    if (tref_[cur_].ensure(rows + 16, st_, true)) ptrs_dirty_ = true; /* + slack: bulk copies read whole 16-byte granules */
    if (tval_[cur_].ensure(rows + 16, st_, true)) ptrs_dirty_ = true;
    if (tref_[cur_ ^ 1].ensure(rows + 16, st_)) ptrs_dirty_ = true;
    if (tval_[cur_ ^ 1].ensure(rows + 16, st_)) ptrs_dirty_ = true;
  }
  void upload_ptrs() {
    CUDA_CHECK(cudaStreamSynchronize(st_)); /* h_ptrs_ may still be in flight from the last upload */
    DevPtrs& h = *h_ptrs_;
    h.P = d_P_; h.tt = d_tt_; h.forms = d_forms_; h.nodes = nodes_.p; h.done = done_.p; h.done_form = done_form_.p; h.gdone = gdone_.p; h.gdone_form = gdone_form_.p; h.surv = surv_.p;
    h.ksend = ksend_.p; h.ksend_form = ksend_form_.p; h.shdr = d_shdr_; h.ghdr = d_ghdr_; h.tail = tail_.p;
    h.ghdr2 = d_ghdr2_; h.g2_idx = g2_idx_.p; h.g2_gap = g2_gap_.p; h.g2_events = g2_events_.p;
    for (int b = 0; b < 2; b++) { h.keys[b] = keys_[b].p; h.idx[b] = idx_[b].p; h.tref[b] = tref_[b].p; h.tval[b] = tval_[b].p; }
    h.heap = heap_.p;
    h.uniq_seq = uniq_seq_.p; h.uniq_gap = uniq_gap_.p; h.uniq_sk = uniq_sk_.p; h.events = events_.p; h.staged = staged_.p;
    for (int sd = 0; sd < 2; sd++) { h.hset[sd] = hset_[sd].p; h.hset_log2[sd] = hset_log2_[sd]; }
    h.rs_state = rs_state_.p; h.cc_state = cc_state_.p; h.sortctl = d_sortctl_; h.opcnt = d_opcnt_;
    h.host_out = d_host_out_;
    CUDA_CHECK(cudaMemcpyAsync(d_ptrs_, h_ptrs_, sizeof(DevPtrs), cudaMemcpyHostToDevice, st_));
    ptrs_dirty_ = false;
  }

  cudaStream_t st_;
  int num_sms_ = 0, dev_ = 0;
  u32 expand_grid_ = 0, wide_grid_ = 0;
  int k1_nq_ = 0;
  bool use_graphs_ = true, use_small_post_ = true, force_small_post_ = false, pass_log_ = false;
  bool use_tiny_ = true, host_prof_ = false, use_pdl_ = true, stage_events_ = false;
  u32 small_grid_ = 24; /* K1 blocks of a pass predicted to be small */
  bool no_spread_ = false;
  double hp_[6] = {0, 0, 0, 0, 0, 0}; /* prepare, enqueue, wait, collect, events, passes */
  static double now_us() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
  }
  u32 tiny_nodes_ = 1200, tiny_frontier_ = 2048;
  std::map<int, cudaGraphExec_t> graphs_;
  RiesParams* d_P_ = 0;
  TestTables* d_tt_ = 0;
  TestTables h_tt_;
  PassOut* d_out_ = 0;
  PassOut* h_out_ = 0;
  PassOut* d_host_out_ = 0; /* device address of h_out_ (mapped) */
  u32 pass_seq_ = 0;
  bool poll_host_ = true;
  PassCtl* d_ctl_ = 0; /* = &d_out_->ctl */
  PassCtl* h_ctl_ = 0;
  unsigned char* d_passin_ = 0; /* PassArgs, then the forms */
  unsigned char* h_passin_ = 0;
  PassArgs* d_args_ = 0;
  PassArgs* h_args_ = 0;
  FormDesc* d_forms_ = 0;
  TableState* d_tab_ = 0;
  DevPtrs* d_ptrs_ = 0;
  DevPtrs* h_ptrs_ = 0;
  bool ptrs_dirty_ = true;
  SortCtl* d_sortctl_ = 0;
  u64* d_opcnt_ = 0;
  cudaEvent_t ev_[5];
  float last_ms_[4] = {0, 0, 0, 0}; /* per-stage device time of the last pass (pass log) */
  cudaEvent_t ev_g_[2];
  bool gather_pending_ = false;
  FormDesc* h_forms_ = 0;
  DevBuf<Node> nodes_;
  DevBuf<Row> done_, surv_, gdone_, ksend_;
  DevBuf<uint16_t> done_form_, gdone_form_, ksend_form_;
  DevBuf<SeqKey> uniq_sk_;
  DevBuf<TailRec> tail_;
  SharedRhs* shared_ = 0; /* the batch's RHS database, if this engine works for a sharing batch */
  bool shared_ok_ = false, throughput_mode_ = false;
  bool use_persist_ = false, persist_ = false;
  u32 persist_max_ = 3000000;
  RiesParams cur_P_;
  /* per side: hash set of the table's keys (state.cuh); rows that side holds, new rows of its last pass */
  DevBuf<u64> hset_[2];
  u32 hset_log2_[2] = {0, 0};
  size_t side_rows_[2] = {0, 0};
  u32 prev_new_[2] = {0, 0};
  bool last_tiny_ = false, last_using_x_ = false;
  PassArgs* h_args2_ = 0; /* pinned: the scalars of a sharded pass's replicated half */
  SlotHdr* d_shdr_ = 0;   /* this rank's send-slot header; the gathered headers (RIES_MAX_RANKS) */
  SlotHdr* d_ghdr_ = 0;
  SlotHdr* d_ghdr2_ = 0;  /* second exchange of a range-split pass */
  DevBuf<u32> g2_idx_, g2_gap_;
  DevBuf<MatchEvent> g2_events_;
  double watchdog_s_ = 300.0; /* RIES_WATCHDOG_SEC */
  bool use_range_ = true;
  u32 range_min_ = 65536; /* table rows from which the second half of a sharded pass is split by value range */
  u32 range_new_[2] = {0, 0}, range_ebound_ = 256;
  u32 prev_gen_[2] = {0, 0}, prev_keys_[2] = {0, 0}, shard_kept_[2] = {0, 0};
  u32 shard_min_ = 100000; /* complete expressions in the previous same-side pass from which a pass is sharded */
  DevBuf<u64> rs_state_, cc_state_;
  DevBuf<Staged> staged_;
  DevBuf<u64> keys_[2];
  DevBuf<u32> idx_[2];
  DevBuf<u32> uniq_seq_, uniq_gap_;
  DevBuf<Row> heap_;
  DevBuf<u32> tref_[2];
  DevBuf<double> tval_[2];
  DevBuf<MatchEvent> events_;
  MatchEvent* h_head_ = 0;
  MatchEvent* h_events_ = 0;
  size_t h_events_cap_ = 0;
  u32 epoch_host_ = 16; /* mirrors TableState::epoch: tags stay unique across searches of this engine */
  u32 nt_ = 0;
  int cur_ = 0;
  u32 last_nu_ = 0;
  u32 prev_nodes_[2] = {0, 0}, prev_done_[2] = {0, 0};
};

Engine* make_device_engine() { return new DeviceEngine(); }

void device_engine_report(Engine* e, DeviceReport* r) {
  DeviceEngine* d = static_cast<DeviceEngine*>(e);
  r->kernel_launches = d->kernel_launches;
  r->h2d_bytes = d->h2d_bytes;
  r->d2h_bytes = d->d2h_bytes;
  for (int i = 0; i < 4; i++) { r->stage_ms[i] = d->stage_ms[i]; r->stage_launches[i] = d->stage_launches[i]; }
  r->sort_keys = d->sort_keys;
  r->merge_rows = d->merge_rows;
  r->expand_nodes = d->expand_nodes;
  r->sharded_passes = d->sharded_passes;
  r->gather_bytes = d->gather_bytes;
  r->gather_ms = d->gather_ms;
  r->table_rows = d->table_rows();
  d->read_opcounts(&r->op_count[0][0]);
  memcpy(r->op_char, d->sym_chars_, sizeof(r->op_char));
}

double device_engine_dfma_peak(Engine* e) { return static_cast<DeviceEngine*>(e)->dfma_peak_tflops(); }

}  // namespace ries
