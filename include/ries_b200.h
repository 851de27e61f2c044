/* ries_b200 -- C ABI of the B200-native RIES search engine (libries_b200.so).
 *
 * This is the boundary a maintainer of thomasahle/ries binds instead of the emscripten module: the
 * web worker calls `moduleInstance.callMain([...args, target])` and collects what the program prints
 * (src/workers/ries.worker.js:100-113, 226-234); natively that is ries.c's
 * `int main(int nargs, char *argv[])` (public/ries.c:12062) writing to stdout.  Plain pointers and
 * sizes only; no C++ or torch types cross this line.  INTEGRATION.md shows the N-API / ctypes stubs.
 */
#ifndef RIES_B200_H
#define RIES_B200_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Receives program output in order (the replacement for Module.print / stdout). */
typedef void (*ries_write_fn)(void *ctx, const char *buf, size_t n);

/* Replaces main(argc, argv) of public/ries.c:12062 (wasm export `_main` / `callMain`).
 * argv[0] is the program name, the target number and options follow exactly as for `ries`.
 * Output that the reference prints to stdout goes to `out` (NULL => the process's stdout).
 * Returns the status the reference would have passed to exit(): 0 normal end, -9998 for every
 * "(Stopping now because ...)" early stop (ries.c:4541, seen by a shell as 242), -1 usage error,
 * 1 missing target.  3 = no CUDA device (this library has no CPU search path).
 * Not re-entrant (one search at a time per process, like the reference's global state,
 * ries.c:3201-3538); the device is chosen with the RIES_DEVICE environment variable. */
int ries_main(int argc, char **argv, ries_write_fn out, void *ctx);

/* Counters of the most recent ries_main() call (the stats block of print_end, ries.c:7599-7628,
 * as numbers) -- what bench.py divides by time. */
typedef struct ries_run_stats {
  double lhs_expressions, rhs_expressions; /* lhs_gen, rhs_gen */
  double lhs_dead_ends, rhs_dead_ends;     /* lhs_prune, rhs_prune */
  long long lhs_distinct, rhs_distinct;    /* lhs_insert, rhs_insert */
  int lhs_max_complexity, rhs_max_complexity;
  int passes;
  int exit_code;
  double seconds;                           /* wall time inside ries_main */
  double engine_seconds, forms_seconds;     /* of which: device passes (launch + wait), host form enumeration */
  unsigned long long kernel_launches;      /* CUDA kernels this call launched */
  unsigned long long h2d_bytes, d2h_bytes; /* bytes copied host->device / device->host */
  /* device time per stage from CUDA events on the engine's stream, milliseconds, and the launches in
     each: [0] K1 enumerate+evaluate  [1] order + radix sort  [2] dedupe/match/merge  [3] Newton refine */
  double stage_ms[4];
  unsigned long long stage_launches[4];
  unsigned long long sort_keys;            /* complete expressions that went through the radix sort */
  unsigned long long merge_rows;           /* table rows written by the per-pass merges */
  unsigned long long expand_nodes;         /* live prefixes evaluated by K1 */
  /* exec() calls per symbol (the reference's g_exec_calls, ries.c:8943, split by symbol): op_char[i] is
     the symbol, counted separately for RHS passes (value only) and LHS passes (value + d/dx) */
  char op_char[48];
  unsigned long long op_count_rhs[48], op_count_lhs[48];
  /* one search over several GPUs (ries_dist_init): passes whose K1 ran sharded, bytes all-gathered */
  unsigned long long sharded_passes, gather_bytes;
  double gather_ms;                         /* device time of those all-gathers */
} ries_run_stats;
void ries_last_stats(ries_run_stats *out);

/* Many independent searches (BASELINE config 5: a batch of targets, each searched with the same flags).
 * `flags` are the argv entries between the program name and the target, e.g. {"-l4", "-F3"}.  The searches
 * run on `n_workers` host threads, each with its own device engine and CUDA stream on the same GPU, so the
 * many small kernels of different searches overlap; results arrive through `done` (serialised, any order),
 * `index` being the position in `targets`.  Returns the number of searches that ended with an error status
 * other than 0 / -9998, or -1 (nothing done) while this process is part of a sharded search (ries_dist_init).
 * The workers' engines are kept for the next call.  The reference has no batch mode: this is
 * `for t in targets: ries <flags> t`. */
typedef void (*ries_batch_done_fn)(void *ctx, int index, int exit_code, const char *text, size_t n);
int ries_batch(int n_targets, const char *const *targets, int n_flags, const char *const *flags,
               int n_workers, ries_batch_done_fn done, void *ctx);

/* The targets of a batch whose search parameters differ in the target alone share ONE enumeration of every
 * right-hand-side pass (BASELINE config 5: "sharing one RHS database"): the pass is run once, by the first worker
 * that needs it, and every target filters it by its own overflow bound (public/ries.c:12250, 9020); everything
 * after enumeration stays per target, so each report is still byte-identical to `ries <flags> t`.  Counters of the
 * most recent ries_batch() call: RHS passes enumerated for the batch, RHS passes of targets served from them.
 * (RIES_BATCH_NO_SHARE=1 in the environment switches the sharing off.) */
void ries_batch_shared_stats(unsigned long long *passes_built, unsigned long long *passes_taken);

/* ONE search sharded over the GPUs of one box (north_star row (e); the reference is single-threaded, there is
 * nothing to cite): one process per GPU, each calls ries_main() with the same argv after joining an NCCL
 * communicator.  The forms of every large pass are dealt round-robin to the ranks for enumerate+evaluate,
 * the ranks' complete expressions are all-gathered over NVLink in rank-major order, and the rest of the pass
 * runs replicated, so every rank prints the identical report.  NCCL is bound at run time (libnccl.so.2).
 *   rank 0:      ries_dist_unique_id(id)  -> hand the 128 bytes to the other ranks (any side channel)
 *   every rank:  ries_dist_init(rank, world, id)   after selecting its GPU (RIES_DEVICE)
 *   ...          ries_main(...) as usual
 *   every rank:  ries_dist_finalize()
 * Return 0 on success. */
int ries_dist_unique_id(unsigned char id[128]);
int ries_dist_init(int rank, int world, const unsigned char id[128]);
void ries_dist_finalize(void);

/* Measured fp64 FMA throughput of the device (TFLOP/s, a short DFMA microbenchmark on the engine's
 * stream): the denominator bench.py uses for K1's fp64 figure.  0 without a GPU. */
double ries_fp64_peak_tflops(void);

/* "sm_100a" when the CUDA engine is usable in this process, else a diagnostic string. */
const char *ries_engine_name(void);

#ifdef __cplusplus
}
#endif
#endif
