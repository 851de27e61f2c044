// C ABI (include/ries_b200.h): main(argc, argv) of the reference (ries.c:12062) as a library call.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <exception>

#include <atomic>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ries_b200.h"
#include "host/dist.h"
#include "host/ries_host.h"

namespace ries {
Engine* make_device_engine();
void device_engine_report(Engine* e, DeviceReport* r);
double device_engine_dfma_peak(Engine* e);
}  // namespace ries

static ries_run_stats g_last;
/* One engine per process, reused by successive ries_main() calls (the web worker calls callMain
   repeatedly on one module instance, ries.worker.js:64-123): device buffers stay allocated. */
static ries::Engine* g_engine = 0;

extern "C" int ries_main(int argc, char** argv, ries_write_fn out, void* ctx) {
  ries::Sink sink{out, ctx};
  ries::Options opt;
  ries::RunStats st;
  int code = 0;
  memset(&g_last, 0, sizeof(g_last));
  try {
    if (argc > 1) {
      ries::parse_args(argc, argv, opt, sink);
    } else { /* ries.c:12089 */
      sink.printf("%s: Please specify a target number.\n\n", argc > 0 ? argv[0] : "ries");
      ries::brief_help(sink);
      throw ries::ExitRequest{1};
    }
    if (!g_engine) g_engine = ries::make_device_engine(); /* throws ExitRequest{3} without a GPU: no CPU path */
    code = ries::run_search(opt, *g_engine, sink, &st);
  } catch (const ries::ExitRequest& e) {
    code = e.code;
    if (code == 1 && g_engine && opt.got_target) { /* out of device memory: start the next call from clean buffers */
      delete g_engine;
      g_engine = 0;
    }
  } catch (const std::exception& e) { /* nothing may unwind across the C ABI */
    fprintf(stderr, "ries_b200: %s\n", e.what());
    code = 3;
    delete g_engine;
    g_engine = 0;
  } catch (...) {
    code = 3;
    delete g_engine;
    g_engine = 0;
  }
  if (g_engine) {
    ries::DeviceReport r;
    ries::device_engine_report(g_engine, &r);
    g_last.kernel_launches = r.kernel_launches;
    g_last.h2d_bytes = r.h2d_bytes;
    g_last.d2h_bytes = r.d2h_bytes;
    for (int i = 0; i < 4; i++) { g_last.stage_ms[i] = r.stage_ms[i]; g_last.stage_launches[i] = r.stage_launches[i]; }
    g_last.sort_keys = r.sort_keys;
    g_last.merge_rows = r.merge_rows;
    g_last.expand_nodes = r.expand_nodes;
    g_last.sharded_passes = r.sharded_passes;
    g_last.gather_bytes = r.gather_bytes;
    g_last.gather_ms = r.gather_ms;
    for (int i = 0; i < 48; i++) {
      g_last.op_char[i] = (char)r.op_char[i];
      g_last.op_count_rhs[i] = r.op_count[0][i];
      g_last.op_count_lhs[i] = r.op_count[1][i];
    }
  }
  g_last.lhs_expressions = st.lhs_gen; g_last.rhs_expressions = st.rhs_gen;
  g_last.lhs_dead_ends = st.lhs_prune; g_last.rhs_dead_ends = st.rhs_prune;
  g_last.lhs_distinct = st.lhs_insert; g_last.rhs_distinct = st.rhs_insert;
  g_last.lhs_max_complexity = st.lmax; g_last.rhs_max_complexity = st.rmax;
  g_last.passes = st.passes;
  g_last.exit_code = code;
  g_last.seconds = st.exec_seconds;
  g_last.engine_seconds = st.engine_seconds;
  g_last.forms_seconds = st.forms_seconds;
  return code;
}

extern "C" void ries_last_stats(ries_run_stats* out) {
  if (out) *out = g_last;
}

extern "C" int ries_dist_unique_id(unsigned char id[128]) { return ries::dist_unique_id(id); }

extern "C" int ries_dist_init(int rank, int world, const unsigned char id[128]) {
  if (world < 1 || world > RIES_MAX_RANKS || rank < 0 || rank >= world) return -3;
  /* the communicator binds to the current device: select it the way the engine will */
  const char* dev_env = getenv("RIES_DEVICE");
  if (cudaSetDevice(dev_env ? atoi(dev_env) : 0) != cudaSuccess) return -2;
  return ries::dist_init(rank, world, id);
}

extern "C" void ries_dist_finalize(void) { ries::dist_finalize(); }

extern "C" double ries_fp64_peak_tflops(void) {
  try {
    if (!g_engine) g_engine = ries::make_device_engine();
    return ries::device_engine_dfma_peak(g_engine);
  } catch (...) {
    return 0.0;
  }
}

extern "C" const char* ries_engine_name(void) {
  try {
    if (!g_engine) g_engine = ries::make_device_engine();
    return "sm_100a";
  } catch (...) {
    return "unavailable: no CUDA device";
  }
}

/* ---- batch of independent targets ---- */
namespace {
void append_sink(void* ctx, const char* buf, size_t n) { static_cast<std::string*>(ctx)->append(buf, n); }
}  // namespace

/* Engines of the batch workers, kept across calls (device buffers, pinned memory, CUDA graphs: building one costs
   more than a -l2 search).  A worker takes one for the whole call and hands it back at the end. */
static unsigned long long g_batch_shared_built = 0, g_batch_shared_used = 0;
extern "C" void ries_batch_shared_stats(unsigned long long* passes_built, unsigned long long* passes_taken) {
  if (passes_built) *passes_built = g_batch_shared_built;
  if (passes_taken) *passes_taken = g_batch_shared_used;
}
static std::mutex g_pool_mu;
static std::vector<ries::Engine*> g_pool;

static ries::Engine* pool_take() {
  {
    std::lock_guard<std::mutex> lock(g_pool_mu);
    if (!g_pool.empty()) {
      ries::Engine* e = g_pool.back();
      g_pool.pop_back();
      return e;
    }
  }
  try {
    return ries::make_device_engine();
  } catch (...) {
    return 0;
  }
}

static void pool_give(ries::Engine* e) {
  if (!e) return;
  std::lock_guard<std::mutex> lock(g_pool_mu);
  g_pool.push_back(e);
}

extern "C" int ries_batch(int n_targets, const char* const* targets, int n_flags, const char* const* flags,
                          int n_workers, ries_batch_done_fn done, void* ctx) {
  if (n_targets <= 0) return 0;
  /* the workers' engines would all issue collectives on the one communicator of a sharded search */
  if (ries::dist_state().world > 1) return -1;
  if (n_workers < 1) n_workers = 1;
  if (n_workers > n_targets) n_workers = n_targets;
  std::atomic<int> next(0), failures(0);
  std::mutex mu;
  /* One RHS database for the batch (BASELINE config 5): the targets whose search parameters differ in the target
     alone share one enumeration of every right-hand-side pass (engine.cu, SharedRhs).  Planned here from the
     command lines: the common parameters, and the range of overflow bounds p_ovr = |T| / k_prune_deriv the shared
     passes must serve.  Targets more than four decades below the largest one (one in ten thousand of a uniform
     batch) would make every pass keep records for that much of its tail, so they search on their own. */
  void* shared = 0;
  if (n_targets >= 2 && !getenv("RIES_BATCH_NO_SHARE")) {
    std::vector<::RiesParams> ps((size_t)n_targets);
    std::vector<char> ok((size_t)n_targets, 0);
    for (int i = 0; i < n_targets; i++) {
      std::vector<std::string> args;
      args.push_back("ries");
      for (int k = 0; k < n_flags; k++) args.push_back(flags[k]);
      args.push_back(targets[i]);
      std::vector<char*> argv;
      for (auto& a : args) argv.push_back(&a[0]);
      ok[i] = ries::search_params((int)argv.size(), argv.data(), &ps[i]) ? 1 : 0;
    }
    int first = -1;
    for (int i = 0; i < n_targets && first < 0; i++) if (ok[i]) first = i;
    if (first >= 0) {
      const ::RiesParams sig = ries::shared_rhs_signature(ps[first]);
      double p_max = 0;
      for (int i = 0; i < n_targets; i++) {
        if (!ok[i]) continue;
        const ::RiesParams si = ries::shared_rhs_signature(ps[i]);
        ok[i] = memcmp(&si, &sig, sizeof(sig)) == 0 && ps[i].p_ovr > 0 && ps[i].p_ovr < 1e300;
        if (ok[i] && ps[i].p_ovr > p_max) p_max = ps[i].p_ovr;
      }
      double p_min = p_max;
      int sharing = 0;
      for (int i = 0; i < n_targets; i++)
        if (ok[i] && ps[i].p_ovr >= p_max * 1e-4) { sharing++; if (ps[i].p_ovr < p_min) p_min = ps[i].p_ovr; }
      if (sharing >= 2) shared = ries::shared_rhs_create(sig, p_max, p_min);
    }
  }
  auto worker = [&]() {
    ries::Engine* eng = pool_take();
    if (eng) { eng->set_shared(shared); eng->set_throughput_mode(n_workers > 1); }
    while (true) {
      int i = next.fetch_add(1);
      if (i >= n_targets) break;
      std::string text;
      int code = 3;
      if (eng) {
        std::vector<std::string> args;
        args.push_back("ries");
        for (int k = 0; k < n_flags; k++) args.push_back(flags[k]);
        args.push_back(targets[i]);
        std::vector<char*> argv;
        for (auto& a : args) argv.push_back(&a[0]);
        ries::Sink sink{append_sink, &text};
        ries::Options opt;
        try {
          ries::parse_args((int)argv.size(), argv.data(), opt, sink);
          code = ries::run_search(opt, *eng, sink, 0);
        } catch (const ries::ExitRequest& e) {
          code = e.code;
          if (code == 1 && opt.got_target) { /* out of device memory: do not reuse half-grown buffers */
            delete eng;
            eng = pool_take();
            if (eng) { eng->set_shared(shared); eng->set_throughput_mode(n_workers > 1); }
          }
        } catch (...) {
          code = 3;
          delete eng;
          eng = pool_take();
          if (eng) { eng->set_shared(shared); eng->set_throughput_mode(n_workers > 1); }
        }
      }
      if (code != 0 && code != RIES_EXIT_NO_ERROR) failures++;
      std::lock_guard<std::mutex> lock(mu);
      if (done) done(ctx, i, code, text.data(), text.size());
    }
    if (eng) eng->set_shared(0);
    pool_give(eng);
  };
  std::vector<std::thread> pool;
  for (int w = 0; w < n_workers; w++) pool.emplace_back(worker);
  for (auto& t : pool) t.join();
  if (shared) {
    ries::shared_rhs_stats(shared, &g_batch_shared_built, &g_batch_shared_used);
    ries::shared_rhs_destroy(shared);
  }
  return failures.load();
}
