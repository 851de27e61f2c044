// Host side of the B200 RIES engine: the thin C++ layer that mirrors the reference's
// main()/parse_args()/init2()/report_match()/print_end() boundary (ries.c:12062, 11117, 10314,
// 7887, 7574) and drives the device engine one complexity pass at a time.
#pragma once
#include <stdarg.h>
#include <stddef.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "core/params.h"

namespace ries {

/* Where stdout goes (the C-ABI's write callback; NULL => fwrite(stdout)). */
struct Sink {
  void (*fn)(void* ctx, const char* buf, size_t n);
  void* ctx;
  void write(const char* s, size_t n) const;
  void printf(const char* fmt, ...) const __attribute__((format(printf, 2, 3)));
};

/* Thrown wherever the reference calls exit()/print_end(nonzero). */
struct ExitRequest {
  int code;
};

#define RIES_EXIT_NO_ERROR (-9998) /* ries.c:4541 */

enum { OF_POSTFIX = 0, OF_CONDENSED = 1, OF_NORMAL = 2, OF_FORTH = 3 };

/* Per-symbol attributes the host needs beyond RiesParams (names for output). */
struct SymInfo {
  unsigned char ch;
  char seft;             /* 'a' 'b' 'c' */
  short wgt;
  short alwd;
  const char* name_forth; /* may be NULL */
  const char* name_infix;
};

/* Everything parse_args()/init1() decide (ries.c:10228-10310, 11117-11964). */
struct Options {
  std::string argv0;
  double levadj = 2.0;
  double tlevel = 4.0;
  bool got_target = false;
  double target = 1.0;
  double mag_ulp = 0.0;
  short sa_alwd[256];
  int restrict_subexpr = TYPE_NONE;
  int restrict_exponents = TYPE_NONE;
  int restrict_trig_args = TYPE_NONE;
  bool exact_exit = false;
  bool relative_x = true;
  bool wide_output = false;
  bool show_ss = false;
  bool S_option = false;
  bool NOS_options = false;
  bool allow_slow_message = true;
  bool no_cv_simplify = false;
  int canon_ops = 0; /* --canon-reduction (ries.c:11255): CANONVAL_* bits */
  bool solve_for_x = false; /* -s / --try-solve-for-x (ries.c:11756) */
  double derivative_margin = 0;       /* --derivative-margin (11284): 0 = automatic */
  double min_equ_val = -9.9e99, max_equ_val = 9.9e99; /* --min/--max-equate-value (11484, 11373) */
  double sincos_max_arg = 1.0;        /* --max-trig-cycles (11536) */
  double trig_scale = 0;              /* --trig-argument-scale (11731): k_sincos_arg_scale, <= 0 means pi */
  bool trig_scale_default = false;    /* g_trig_scale_default as parse_args leaves it */
  double sig_loss = 0.01;             /* --significance-loss-margin (11622) */
  bool match_all_digits = false;      /* --match-all-digits / --mad (11366) */
  bool explicit_multiply = false;     /* --explicit-multiply (11338) */
  std::string sym_name[256];          /* --symbol-names (11648): sa_name = name_forth = given text */
  bool sym_named[256];
  bool refinement = true;   /* --no-refinement (11577): the threshold never shrinks; every match inside it is printed once */
  bool one_sided = false;   /* --one-sided (11600): 'x' is the only LHS, RHS passes from then on (12425) */
  bool anagram = false;     /* --numeric-anagram (11586): digit budget through sa_alwd, implies one_sided */
  short preempt_weight[256]; /* --symbol-weights (11689): -1 = the built-in weight */
  int out_format = OF_NORMAL;
  double max_matches = 100;
  double k_max_match_dist = -0.01;
  double min_matchsize = 0;
  int debug_y = 0;
  Options();
};

/* State init2() + main()'s parameter set-up derive (ries.c:10314-10744, 12111-12317). */
struct Setup {
  RiesParams P;
  bool trig_pi_names = true;
  std::string name_store[256]; /* owns the --symbol-names strings the SymInfo pointers refer to */
  const char* any_name[256];   /* sa_name of every character code, symbol or not ('=', '(', '.', ...) */
  const char* any_forth[256];  /* name_forth of every character code (0 = none) */ /* sinpi/cospi/tanpi vs sin/cos/tan (ries.c:10419-10441) */
  SymInfo info[RIES_MAX_SYMS];
  int idx_of[256]; /* char -> table index, -1 if unknown */
  double k_min_best_match;
  double init_match_dist;
  double tgt_to_print;
  double k_precision_ulp;
  int k_usable_digits;
  int k_nominal_digits;
  bool x_lhs_only;
};

/* parse_args: throws ExitRequest on usage errors, printing what the reference prints. */
void parse_args(int argc, char** argv, Options& opt, const Sink& out);
/* init2 + the threshold set-up of main(); may print notices ("Replacing -i with -r ..."). */
void build_setup(Options& opt, Setup& S, const Sink& out);
/* main()'s parameter block between init2() and the header: -i demotion, thresholds (12111-12317). */
void finish_setup(Options& opt, Setup& S, const Sink& out);
void show_symset(const Options& opt, const Setup& S, const Sink& out);
void brief_help(const Sink& out);
/* `pi_names`: the trig symbols are called sinpi/cospi/tanpi (k_sincos_arg_scale == k_pi, ries.c:10419) */
const char* sym_defn(unsigned char ch, bool normal, bool pi_names = true);

/* gen_forms/gf_1 + ge_1's suffix tables for one pass, in the reference's visiting order. */
void gen_forms(const RiesParams& P, int e_minw, int e_maxw, std::vector<FormDesc>& forms,
               double* n_leaf_forms);

/* One pass as seen by the host. */
struct PassOutput {
  double generated = 0;  /* g_ne */
  double dead_ends = 0;  /* prune_count */
  long inserted = 0;     /* insert_count */
  std::vector<MatchEvent> events; /* sorted by (seq, kind) */
};

/* The device engine (K1..K4).  The product has exactly one implementation (engine.cu); the CPU
   emulation used by the tests lives in oracle/emul.cc and is never linked into the product. */
class Engine {
 public:
  virtual ~Engine() {}
  virtual void begin(const RiesParams& P) = 0;
  virtual void run_pass(int using_x, int cplx, const std::vector<FormDesc>& forms,
                        double best_match0, bool skip_exact, PassOutput& out) = 0;
  /* rows of the last pass inserted before the new expression of candidate `ev` (or up to and including it):
     early-exit accounting */
  virtual long inserted_before(const MatchEvent& ev, bool inclusive) = 0;
  virtual const char* name() const = 0;
  /* batch of targets sharing one enumeration of the right-hand sides (device engine only): the context made by
     shared_rhs_create(), or 0 */
  virtual void set_shared(void* ctx) { (void)ctx; }
  /* several engines share the GPU (a batch): size the grids of mid-size passes by the work, not for latency */
  virtual void set_throughput_mode(bool on) { (void)on; }
};

/* One RHS database for a batch of targets (BASELINE config 5).  `sig` is the RiesParams all sharing targets have in
   common (target-dependent fields cleared, see shared_rhs_signature), p_max / p_min the largest and smallest
   overflow bound p_ovr among them.  The context owns device memory: destroy it after the batch, from any thread. */
void* shared_rhs_create(const RiesParams& sig, double p_max, double p_min);
void shared_rhs_destroy(void* ctx);
void shared_rhs_stats(void* ctx, unsigned long long* passes_built, unsigned long long* consumer_passes);
RiesParams shared_rhs_signature(const RiesParams& P);
/* the search parameters a command line leads to, without searching (what ries_batch needs to plan the sharing);
   false when the arguments do not describe a search */
bool search_params(int argc, char** argv, RiesParams* out);

struct RunStats {
  double lhs_gen = 0, rhs_gen = 0, gen_total = 0;
  double lhs_prune = 0, rhs_prune = 0;
  long lhs_insert = 0, rhs_insert = 0;
  int lmax = 0, rmax = 0;
  int passes = 0;
  double exec_seconds = 0;
  double engine_seconds = 0; /* inside Engine::run_pass (launch + wait) */
  double forms_seconds = 0;  /* host form enumeration */
};

/* What the device engine measured during one search (CUDA events on its own stream). */
struct DeviceReport {
  unsigned long long kernel_launches, h2d_bytes, d2h_bytes;
  double stage_ms[4];
  unsigned long long stage_launches[4];
  unsigned long long sort_keys, merge_rows, expand_nodes, table_rows;
  unsigned long long sharded_passes, gather_bytes;
  double gather_ms;
  unsigned long long op_count[2][RIES_MAX_SYMS]; /* exec() calls per symbol: [0] RHS passes, [1] LHS passes */
  unsigned char op_char[RIES_MAX_SYMS];
};

/* main() after parse_args: the whole search.  Returns the process exit code. */
int run_search(Options& opt, Engine& eng, const Sink& out, RunStats* stats_out);

/* number formatting (ries_spfg ries.c:4069, pf_intfloat_wid 7542) */
std::string spfg(int prec, double x);
std::string pf_intfloat_wid(double x, int width);

}  // namespace ries
