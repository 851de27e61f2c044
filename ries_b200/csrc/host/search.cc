// The pass scheduler, the sequential replay of device-found candidates, and the report: the
// host-side restatement of main()'s outer loop (/root/reference/public/ries.c:12327-12580),
// check_match()'s threshold logic (8243-8369), check_exact_match() (8541-8618), report_match()
// (7887-8236), check_exit() (7805-7843) and print_end() (7574-7633).
//
// Division of labour: the device engine generates, evaluates, sorts, dedupes, merges and finds
// every (new expression, neighbour) pair that could beat the match threshold as it stood at the
// START of the pass, already refined by cv_simplify+Newton.  The threshold only ever shrinks, so
// replaying those candidates here in insertion order with the live threshold reproduces the
// reference's streaming decisions exactly.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#include <algorithm>
#include <future>
#include <map>
#include <memory>

#include "core/exec.h"
#include "host/ries_host.h"

namespace ries {

/* ---------------- number formatting ---------------- */

static double intpow10(int p) { /* ries_intpow(10.0, p) ries.c:3794 */
  double rv = 1.0, pow2 = 10.0;
  int recip = 0;
  if (p < 0) { p = -p; recip = 1; }
  int i = p;
  while (i > 0) {
    if (i & 1) rv = rv * pow2;
    i >>= 1;
    if (i) pow2 = pow2 * pow2;
  }
  return recip ? (1.0 / rv) : rv;
}

/* ries_to_digits ries.c:3839-3955 */
static bool to_digits(double x, char* s, int* expn, int* sign, int precision) {
  int dig1 = precision + 1;
  int sgn = 0;
  if (x < 0.0) { x = -x; sgn = 1; }
  *sign = sgn;
  double r = x;
  if (x == 0.0) {
    *expn = 0;
    for (int i = 0; i < precision; i++) s[i] = '0';
    s[precision] = 0;
    return true;
  }
  int e = (int)(floor(log10(fabs(x))));
  if (e < -300) {
    r = r * intpow10(300);
    r = r / intpow10(e + 300);
  } else if (e > 0) {
    r = r / intpow10(e);
  } else if (e < 0) {
    r = r * intpow10(-e);
  }
  if (r >= 10.0) { r = r / 10.0; e++; }
  else if (r < 1.0) { r = r * 10.0; e--; }
  if (r >= 10.0 || r < 1.0) return false;
  for (int i = 0; i < dig1; i++) {
    int d = (int)r;
    r = r - (double)d;
    r = r * 10.0;
    s[i] = (char)('0' + d);
  }
  for (int i = dig1 - 1; i > 0; i--) {
    if (s[i] < '0') { s[i - 1]--; s[i] = (char)(s[i] + 10); }
  }
  if (s[0] <= '0') return false;
  if (s[dig1 - 1] >= '5') {
    s[dig1 - 2]++;
    int i = dig1 - 2;
    while (i > 0 && s[i] > '9') {
      s[i] = (char)(s[i] - 10);
      i--;
      s[i] = (char)(s[i] + 1);
    }
  }
  if (s[0] > '9') {
    e++;
    for (int i = precision; i >= 2; i--) s[i] = s[i - 1];
    s[0] = '1';
    s[1] = '0';
  }
  s[precision] = 0;
  *expn = e;
  return true;
}

static void bltr0(char* s) { /* ries_bltr0 ries.c:4025 */
  if (!s || !*s) return;
  char* t = s;
  while (*t) t++;
  t--;
  while (t >= s && *t == '0') { *t = 0; t--; }
}

/* ries_spfg(tmp, 100, 0, prec, x) ries.c:4069-4246 with a generous buffer */
std::string spfg(int precision, double x) {
  if (x <= -HUGE_VAL) return std::string("-inf").substr(0, precision);
  if (x >= HUGE_VAL) return std::string("+inf").substr(0, precision);
  if (!((x > -HUGE_VAL) && (x < HUGE_VAL))) return std::string("NaN").substr(0, precision);
  if (precision > 35) precision = 35;
  else if (precision < 1) precision = 1;
  char s2[48];
  int exponent = 0, sign = 0;
  memset(s2, 0, sizeof(s2));
  if (!to_digits(x, s2, &exponent, &sign, precision)) return std::string();
  std::string o;
  if (sign) o.push_back('-');
  if ((exponent > 0) && (exponent < precision)) {
    int i = 0;
    o.push_back(s2[i++]);
    while (exponent) { o.push_back(s2[i++]); exponent--; }
    bltr0(s2);
    if (*s2 == 0) return o;
    o.push_back('.');
    for (; s2[i] && (i < precision); i++) o.push_back(s2[i]);
    return o;
  } else if (exponent == 0) {
    o.push_back(s2[0]);
    bltr0(s2 + 1);
    if (s2[1] == 0) return o;
    o.push_back('.');
    for (int i = 1; s2[i] && (i < precision); i++) o.push_back(s2[i]);
    return o;
  } else if ((exponent < 0) && (exponent > -5)) {
    o += "0.";
    exponent++;
    while (exponent < 0) { o.push_back('0'); exponent++; }
    bltr0(s2);
    if (*s2 == 0) return o;
    for (int i = 0; s2[i] && (i < precision); i++) o.push_back(s2[i]);
    return o;
  }
  char sexp[16];
  if (exponent >= 0) snprintf(sexp, sizeof(sexp), "+%d", exponent);
  else snprintf(sexp, sizeof(sexp), "%d", exponent);
  o.push_back(s2[0]);
  bltr0(s2 + 1);
  if (s2[1]) {
    o.push_back('.');
    for (int i = 1; (i < precision) && s2[i]; i++) o.push_back(s2[i]);
  }
  o.push_back('e');
  o += sexp;
  return o;
}

/* pf_intfloat_wid ries.c:7542-7567 */
std::string pf_intfloat_wid(double x, int width) {
  char fmt1[32], fmt2[32], buf[40];
  int w = (x < 0) ? (width - 1) : width;
  snprintf(fmt1, sizeof(fmt1), "%%.%dg", w);
  snprintf(fmt2, sizeof(fmt2), "%%.%de", w - 6);
  snprintf(buf, 40, fmt1, x);
  if (strchr(buf, 'e') || strchr(buf, 'E')) {
    snprintf(buf, 40, fmt2, x);
  } else if (strchr(buf, '.')) {
    int i = (int)strlen(buf) - 1;
    while ((i > 0) && (buf[i] == '0')) { buf[i] = 0; i--; }
    if ((i > 0) && (buf[i] == '.')) buf[i] = 0;
  }
  return std::string(buf);
}

/* ---------------- infix formatting (-F1 / -F2) ---------------- */
/* infix_preproc 6269, infix_1 5966-6240, infix_expand 6536: postfix -> one-character-per-symbol infix ->
   names.  Phantom symbols (2900-2903): reversed subtract / divide / power. */
namespace {
const char PS_REVSUB = 1, PS_REVDIV = 2, PS_REVPOW = 4;

struct InfixCtx {
  const Setup* S;
  bool explicit_multiply;
  bool used_atan1, used_atan2;
  char seft(unsigned char c) const {
    if (c == (unsigned char)PS_REVSUB || c == (unsigned char)PS_REVDIV || c == 3 || c == (unsigned char)PS_REVPOW) return 'c';
    int g = S->idx_of[c];
    return g >= 0 ? S->info[g].seft : 0;
  }
};

std::string infix_preproc(const InfixCtx& cx, const std::string& e) {
  std::string o;
  size_t n = e.size();
  for (size_t i = 0; i < n; i++) {
    char c = e[i];
    char c1 = (i + 1 < n) ? e[i + 1] : 0, c2 = (i + 2 < n) ? e[i + 2] : 0;
    if (c == '-' && c1 == 'n') { o.push_back(PS_REVSUB); i++; }
    else if (c == '/' && c1 == 'r') { o.push_back(PS_REVDIV); i++; }
    else if (c == ' ') { }
    else if (c == 'n' && c1 && cx.seft((unsigned char)c1) == 'a' && c2 == '+') { o.push_back(c1); o.push_back(PS_REVSUB); i += 2; }
    else if (c == 's') { o.push_back('2'); o.push_back('^'); }
    else if (c == 'r') { o.push_back('1'); o.push_back(PS_REVDIV); }
    else if (c == 'E') { o.push_back('e'); o.push_back(PS_REVPOW); }
    else o.push_back(c);
  }
  return o;
}

bool in_set(const char* set, unsigned char c) { return c != 0 && strchr(set, (char)c) != 0; }

/* Pulls one complete sub-expression off the END of `expr` (shortening it) and returns it as infix. */
bool infix_1(InfixCtx& cx, std::string& expr, std::string& term, unsigned char& t_op) {
  term.clear();
  if (expr.empty()) return false;
  unsigned char op = (unsigned char)expr.back();
  expr.pop_back();
  t_op = op;
  switch (cx.seft(op)) {
    case 'a':
      term.push_back((char)op);
      break;
    case 'b': {
      std::string a;
      unsigned char op_a;
      if (!infix_1(cx, expr, a, op_a)) return false;
      bool swap = (op == 's');
      bool paren_a = true;
      if (op == 'n' && cx.seft(op_a) != 'c') paren_a = false;
      if (!swap) term.push_back((char)op);
      if (paren_a) term.push_back('(');
      term += a;
      if (paren_a) term.push_back(')');
      if (swap) term.push_back((char)op);
      break;
    }
    case 'c': {
      std::string st_a, st_b;
      unsigned char op_a, op_b;
      if (!infix_1(cx, expr, st_b, op_b)) return false;
      if (!infix_1(cx, expr, st_a, op_a)) return false;
      bool swap = false;
      if (op == '+' && cx.seft(op_a) == 'a' && op_a != 'x') swap = true;
      if (op == '*') {
        bool a_has_x = st_a.find('x') != std::string::npos;
        bool b_has_x = st_b.find('x') != std::string::npos;
        if (!a_has_x && op_b == 'x') {
        } else if (cx.seft(op_a) != 'a' && cx.seft(op_b) == 'a') swap = true;
        else if (op_a == 'x' && !b_has_x) swap = true;
        if (cx.seft(op_a) == 'a' && op_b >= '0' && op_b <= '9') swap = true;
        else if (op_a == 'x' && (op_b == 'e' || op_b == 'f' || op_b == 'p')) swap = true;
      }
      if (op == 'v') swap = true;
      else if (op == 'L') swap = true;
      if (op == (unsigned char)PS_REVSUB) { swap = true; op = '-'; t_op = '-'; }
      else if (op == (unsigned char)PS_REVDIV) { swap = true; op = '/'; t_op = '/'; }
      else if (op == (unsigned char)PS_REVPOW) { swap = true; op = '^'; t_op = '^'; }
      const std::string* term_a = &st_a;
      const std::string* term_b = &st_b;
      if (swap) { term_a = &st_b; term_b = &st_a; unsigned char t = op_a; op_a = op_b; op_b = t; }
      bool paren_a = false, paren_b = false, atan2_show_denom = true;
      if (op == 'L') {
        paren_b = true;
        if (cx.seft(op_a) != 'a') paren_a = true;
      } else if (op == 'A') {
        if (op_b == '1') { atan2_show_denom = false; cx.used_atan1 = true; }
        else cx.used_atan2 = true;
      }
      if (cx.seft(op_a) == 'c') {
        paren_a = true;
        if (in_set("+*/^vA", op_a) && in_set("+-", op)) paren_a = false;
        else if (in_set("*^vA", op_a) && in_set("*/A", op)) paren_a = false;
      }
      if (cx.seft(op_b) == 'c') {
        paren_b = true;
        if (op == '+' && in_set("+-*/^vA", op_b)) paren_b = false;
        else if (op == '-' && in_set("*/^vA", op_b)) paren_b = false;
        else if (op == '*' && in_set("*/^vA", op_b)) paren_b = false;
        else if (in_set("/A", op) && in_set("^vA", op_b)) paren_b = false;
      }
      if (in_set("LA", op)) term.push_back((char)op);
      if (op == 'A') term.push_back('(');
      if (paren_a) term.push_back('(');
      term += *term_a;
      if (paren_a) term.push_back(')');
      char tb0 = term_b->empty() ? 0 : (*term_b)[0];
      if (op == '*') {
        if (cx.explicit_multiply) term.push_back((char)op);
        else if (op_a <= '9' && paren_b) { }
        else if (op_a <= '9' && tb0 <= '9') term.push_back((char)op);
        else if (op_b == 'n') term.push_back((char)op);
        else if (cx.seft(op_a) != 'a' || cx.seft(op_b) != 'a' || tb0 > '9') term.push_back('.');
        else term.push_back((char)op);
      } else if (op == 'A') {
        if (atan2_show_denom) term.push_back(',');
      } else if (op == 'L') {
      } else {
        term.push_back((char)op);
      }
      if (atan2_show_denom) {
        if (paren_b) term.push_back('(');
        term += *term_b;
        if (paren_b) term.push_back(')');
      }
      if (op == 'A') term.push_back(')');
      break;
    }
    default:
      return false;
  }
  return true;
}

std::string infix_expand(const InfixCtx& cx, const std::string& in) {
  std::string o;
  for (size_t i = 0; i < in.size(); i++) {
    unsigned char c = (unsigned char)in[i];
    switch (c) {
      case '(': case ',': case ')': case '=': o += cx.S->any_name[c]; break;
      case '.': o += " "; break;
      default: {
        int g = cx.S->idx_of[c];
        if (g >= 0 && cx.S->info[g].name_infix) o += cx.S->info[g].name_infix;
      }
    }
  }
  return o;
}
}  // namespace

/* ---------------- the search ---------------- */

namespace {

double now_sec() {
  struct timeval tv;
  gettimeofday(&tv, 0);
  return (double)tv.tv_sec + 1e-6 * (double)tv.tv_usec;
}

struct Search {
  Options& opt;
  Setup S;
  Engine& eng;
  const Sink& out;
  RunStats st;
  double t_start;

  /* live search state (ries.c:3461, 3231-3233) */
  double best_match = 0;
  bool got_exact = false;
  double num_matches = 0;
  int16_t lmin = 1, lmax = 1, rmin = 1, rmax = 1; /* s16 in the reference (ries.c:3218): a search that never
                                                    inserts anything ends by wrapping around, 12430/12445 */

  /* early-exit bookkeeping: which event of the running pass we are replaying */
  long resident_rows = 0;   /* rows inserted by completed passes */
  const MatchEvent* cur_ev = 0;

  Search(Options& o, Engine& e, const Sink& s) : opt(o), eng(e), out(s) { t_start = now_sec(); }
  ~Search() { if (ahead.valid()) ahead.wait(); }

  /* The forms of a complexity do not depend on the side, and the passes visit complexities 1,2,3,... on
     each side: every list is built once, and the next one is built on a helper thread while the device
     runs the current pass (gen_forms only reads the immutable parameter block). */
  struct FormList { std::vector<FormDesc> forms; double n_leaf = 0; };
  std::map<int, std::unique_ptr<FormList>> form_cache;
  std::future<std::unique_ptr<FormList>> ahead;
  int ahead_cplx = -1;
  std::unique_ptr<FormList> build_forms(int cplx) const {
    std::unique_ptr<FormList> f(new FormList());
    gen_forms(S.P, cplx, cplx, f->forms, &f->n_leaf);
    return f;
  }
  const std::vector<FormDesc>& forms_for(int cplx, double* n_leaf) {
    if (ahead.valid() && (ahead_cplx == cplx || form_cache.size() > 4096)) {
      form_cache[ahead_cplx] = ahead.get();
    }
    auto it = form_cache.find(cplx);
    if (it == form_cache.end()) it = form_cache.emplace(cplx, build_forms(cplx)).first;
    if (!ahead.valid() && !form_cache.count(cplx + 1)) {
      ahead_cplx = cplx + 1;
      const int c = cplx + 1;
      ahead = std::async(std::launch::async, [this, c]() { return build_forms(c); });
    }
    if (n_leaf) *n_leaf = it->second->n_leaf;
    return it->second->forms;
  }

  std::string sym_chars(const Sym128& s) const {
    std::string r;
    for (int i = 0; i < RIES_MAX_ELEN + 1; i++) {
      unsigned c = sym_get(s, i);
      if (!c) break;
      r.push_back((char)S.P.sym_char[c - 1]);
    }
    return r;
  }

  int complexity(const std::string& e) const { /* ries.c:6614 */
    int c = 0;
    for (size_t i = 0; i < e.size(); i++) {
      int g = S.idx_of[(unsigned char)e[i]];
      if (g >= 0) c += S.P.sym_wgt[g];
    }
    return c;
  }

  std::string postfix_formatter(const std::string& e) const { /* ries.c:6572 */
    std::string o;
    for (size_t i = 0; i < e.size(); i++) {
      int g = S.idx_of[(unsigned char)e[i]];
      if (g >= 0 && S.info[g].name_forth) o += S.info[g].name_forth;
      else o.push_back(e[i]);
      o.push_back(' ');
    }
    while (!o.empty() && o.back() == ' ') o.pop_back();
    return o;
  }

  /* --no-refinement bookkeeping (ries.c:10697-10738, 7848-7873, 7919-7942): the list of reported equations
     as one symbol string " lhs=rhs lhs=rhs ...", searched by substring; the reported deltas; and what both
     add to the reference's memory counters. */
  std::string nr_matches;
  size_t nr_alloc = 0;
  std::vector<double> nr_deltas;
  long nr_kib = 0;
  unsigned long nr_bytes = 0;
  void nr_init() {
    if (opt.refinement) return;
    nr_deltas.assign((size_t)opt.max_matches + 1, 0.0);
    size_t space = ((size_t)opt.max_matches) * (size_t)(RIES_MAX_ELEN + 1);
    nr_alloc = space;
    if (opt.solve_for_x) nr_alloc += space;
    if (opt.out_format == OF_CONDENSED || opt.out_format == OF_NORMAL) nr_alloc += space * 4;
    nr_matches = " ";
    nr_kib = (long)(nr_alloc >> 10);
    nr_bytes = 2;
  }
  bool unique_eqn(const std::string& lhs, const std::string& rhs, bool addit) { /* ries.c:7848 */
    const size_t TS_ALLOC_R = (RIES_MAX_ELEN + 1) * 4;
    std::string te = " " + lhs + "=" + rhs + " ";
    if (te.size() > TS_ALLOC_R - 1) te.resize(TS_ALLOC_R - 1); /* symstrncat truncates */
    if (nr_matches.find(te) != std::string::npos) return false;
    if (addit) {
      std::string add = te.substr(1);
      size_t room = (nr_alloc > nr_matches.size() + 1) ? nr_alloc - 1 - nr_matches.size() : 0;
      nr_matches += add.substr(0, std::min(room, add.size()));
      nr_bytes += (unsigned long)(add.size() + 1);
    }
    return true;
  }

  /* Memory the reference's allocator would report for `nodes` tree nodes: sizeof(expr) = 72 on
     LP64 (ries.c:2955), 64 KiB blocks with an 8-byte back link (my_alloc 4837-4900). */
  static void ref_memory(long nodes, long* kib, unsigned long* bytes) {
    const long per_block = (65536 - 8) / 72;
    long blocks = (nodes + per_block - 1) / per_block;
    *kib = blocks * 64;
    *bytes = (unsigned long)(blocks * 8 + nodes * 72);
  }

  long live_rows() {
    if (!cur_ev) return resident_rows;
    /* An EXACT event's expression is never inserted; a neighbour event fires right after its
       expression was inserted (bt_insert 8686-8721). */
    return resident_rows + eng.inserted_before(*cur_ev, cur_ev->kind != EV_EXACT);
  }

  void print_end(int exit_code) { /* ries.c:7574 */
    double tsec = now_sec() - t_start;
    if (got_exact && (opt.restrict_subexpr < TYPE_INT)) {
      out.printf("\n  NOTE: 'exact' match may result from floating-point roundoff error.\n");
    }
    if (opt.out_format <= OF_NORMAL) describe_symbols();
    long total_insert = st.lhs_insert + st.rhs_insert;
    if (total_insert) {
      out.printf("\n");
      out.printf("                         --LHS--      --RHS--      -Total-\n");
      out.printf("     max complexity: %11d  %11d  %11d\n", lmax, rmax, lmax + rmax);
      out.printf("          dead-ends: %11s", pf_intfloat_wid(st.lhs_prune, 11).c_str());
      out.printf("  %11s", pf_intfloat_wid(st.rhs_prune, 11).c_str());
      out.printf("  %11s", pf_intfloat_wid(st.lhs_prune + st.rhs_prune, 11).c_str());
      out.printf("  Time: %.3f\n", tsec);
      out.printf("        expressions: %11s", pf_intfloat_wid(st.lhs_gen, 11).c_str());
      out.printf("  %11s", pf_intfloat_wid(st.rhs_gen, 11).c_str());
      out.printf("  %11s\n", pf_intfloat_wid(st.gen_total, 11).c_str());
      out.printf("           distinct: %11ld  %11ld  %11ld", st.lhs_insert, st.rhs_insert, total_insert);
      long kib; unsigned long bytes;
      ref_memory(live_rows(), &kib, &bytes);
      kib += nr_kib; bytes += nr_bytes;
      if (kib > 1024L) out.printf("  Memory: %ldKiB\n", kib);
      else out.printf("  Memory: %ld B\n", bytes);
      out.printf("\n");
      double combos = ((double)st.lhs_insert) * ((double)st.rhs_insert);
      out.printf("        Total equations tested:    %20s", pf_intfloat_wid(combos, 20).c_str());
      if (combos > 9999) out.printf(" (%.4g)", combos);
      out.printf("\n");
    }
    if (exit_code) throw ExitRequest{exit_code};
  }

  InfixCtx icx{0, false, false, false};
  /* legend bookkeeping: defsym_used ries.c:7417, describe_symbols 7434-7516 */
  bool def_needed[256] = {false};
  bool def_given[256] = {false};
  bool used_trig = false;
  void defsym_used(const std::string& e) {
    for (size_t i = 0; i < e.size(); i++) {
      if (sym_defn((unsigned char)e[i], opt.out_format == OF_NORMAL, S.trig_pi_names)) def_needed[(unsigned char)e[i]] = true;
      if (e[i] == 'C' || e[i] == 'S' || e[i] == 'T') used_trig = true;
    }
  }
  void describe_symbols() {
    /* used_atan1/used_atan2 are only ever set by the infix formatter (6128-6133), so in postfix mode 'A' is
       never described; symbol 0 is the phantom "one-argument arctangent" (PHSYM_ATAN1, 3243) */
    if (def_needed[(unsigned char)'A']) {
      if (icx.used_atan1) def_needed[0] = true;
      if (!icx.used_atan2) def_needed[(unsigned char)'A'] = false;
    }
    bool normal = (opt.out_format == OF_NORMAL);
    auto defn = [&](int sym) -> const char* {
      if (sym == 0) return "atan2(s) = Angle of ray from origin through point (1,s)";
      return sym_defn((unsigned char)sym, normal, S.trig_pi_names);
    };
    const int kLineLeftInit = 79 - 2;
    int lineleft = kLineLeftInit;
    bool need_lf = true, going = true;
    while (going) {
      going = false;
      for (int sym = 0; sym < 256; sym++) {
        if (def_needed[sym] && !def_given[sym]) {
          going = true;
          if (need_lf) { out.printf("\n"); need_lf = false; }
        }
      }
      if (going) {
        const char* candidate = 0;
        int clen = 0, csym = 0;
        for (int sym = 0; sym < 256; sym++) {
          if (def_needed[sym] && !def_given[sym]) {
            int sl = (int)strlen(defn(sym));
            if ((sl > clen) && (sl < lineleft)) {
              csym = sym;
              candidate = defn(sym);
              clen = sl;
            }
          }
        }
        if (candidate) {
          out.printf("  %s", candidate);
          def_given[csym] = true;
          lineleft = lineleft - clen - 2;
        } else {
          out.printf("\n");
          lineleft = kLineLeftInit;
        }
      }
    }
    if (lineleft < kLineLeftInit) out.printf("\n");
    if (used_trig && !S.P.trig_scale_default) { /* 7501-7514 */
      out.printf("  For the trig functions, ");
      if (S.P.sincos_scale == 1.0) out.printf("2 pi");
      else out.printf("%g", 2.0 * S.P.k_pi / S.P.sincos_scale);
      out.printf(" units are a full circle.\n");
    }
  }

  void check_exit(bool is_exact) { /* ries.c:7805 */
    if (is_exact && (opt.restrict_subexpr == TYPE_INT) && opt.exact_exit) {
      out.printf("  (Stopping now because -ie option was given.)\n");
    } else if (is_exact && (opt.restrict_subexpr == TYPE_RAT) && opt.exact_exit) {
      out.printf("  (Stopping now because -re option was given.)\n");
    } else if (is_exact && (opt.restrict_subexpr == TYPE_CONS) && opt.exact_exit) {
      out.printf("  (Stopping now because -ce option was given.)\n");
    } else if (is_exact && (opt.k_max_match_dist == 0)) {
      out.printf("  (Stopping now because an exact match was found.)\n");
    } else if (is_exact && opt.exact_exit) {
      out.printf("  (Stopping now because --min-match-distance 0 or --max-match-distance 0\n"
                 "  was given.)\n");
    } else if (best_match < opt.min_matchsize) {
      double t = (S.k_min_best_match > opt.min_matchsize) ? S.k_min_best_match : opt.min_matchsize;
      out.printf("  (Stopping now because best match is within %7.3g of target value.)\n", t);
    } else if (num_matches >= opt.max_matches) {
      if (num_matches == 1) out.printf("  (Stopping now because 1 match was found.)\n");
      else out.printf("  (Stopping now because %ld matches were found.)\n", (long)num_matches);
    } else {
      return;
    }
    print_end(RIES_EXIT_NO_ERROR);
  }

  /* eval() ries.c:6777 on a symbol string of any length (only try_solve's sign test needs it): value at
     x = target, 0 on any error exactly like the reference's default `*val = k_0`. */
  double eval_string(const std::string& e) const {
    const RiesParams& P = S.P;
    Val stk[RIES_STK_MAX + 2];
    int sp = 0;
    bool has_x = e.find('x') != std::string::npos;
    for (size_t i = 0; i < e.size(); i++) {
      if ((int)i >= RIES_MAX_ELEN) return 0.0; /* ERR_EVAL_TOO_LONG */
      int g = S.idx_of[(unsigned char)e[i]];
      if (g < 0) return 0.0;
      Val r;
      int cls = P.sym_cls[g], err;
      if (cls == 0) {
        if (sp + 1 > RIES_STK_MAX) return 0.0;
        err = exec_push(P, (unsigned char)e[i], P.target, r);
        if (err) return 0.0;
        stk[sp++] = r;
      } else if (cls == 1) {
        if (sp < 1) return 0.0;
        err = exec_unary(P, (unsigned char)e[i], has_x, stk[sp - 1], r);
        if (err) return 0.0;
        stk[sp - 1] = r;
      } else {
        if (sp < 2) return 0.0;
        err = exec_binary(P, (unsigned char)e[i], has_x, stk[sp - 2], stk[sp - 1], r);
        if (err) return 0.0;
        sp--;
        stk[sp - 1] = r;
      }
    }
    return sp > 0 ? stk[sp - 1].v : 0.0;
  }

  /* expr_break ries.c:6688 over endstack 6630: root operator and its argument(s) */
  bool expr_break(const std::string& e, char* op, char* seft, std::string* a1, std::string* a2) const {
    int sp = 0, last_sp_1 = -1;
    a1->clear(); a2->clear();
    if (e.empty()) return false;
    for (size_t i = 0; i < e.size(); i++) {
      int g = S.idx_of[(unsigned char)e[i]];
      char sf = g >= 0 ? S.info[g].seft : 0;
      if (sf == 'a') sp++;
      else if (sf == 'c') sp--;
      if (sp <= 0) return false;
      if (sp == 1 && i + 1 < e.size()) last_sp_1 = (int)i;
    }
    if (sp > 1) return false;
    int g = S.idx_of[(unsigned char)e.back()];
    *op = e.back();
    *seft = g >= 0 ? S.info[g].seft : 0;
    if (*seft == 'b') *a1 = e.substr(0, e.size() - 1);
    else if (*seft == 'c') {
      *a1 = e.substr(0, (size_t)last_sp_1 + 1);
      *a2 = e.substr((size_t)last_sp_1 + 1, e.size() - 1 - ((size_t)last_sp_1 + 1));
    }
    return true;
  }

  /* try_solve ries.c:6886-7123: peel the LHS's root operator onto the RHS until x (or a function with no
     inverse here) is left.  Buffers are TS_ALLOC_L = 22 / TS_ALLOC_R = 88 symbols in the reference. */
  void try_solve(std::string& lhs, std::string& rhs) const {
    const size_t kR = 4 * RIES_EXPR_ALLOC - 1;
    auto cat = [&](std::string& to, const std::string& add) {
      for (size_t i = 0; i < add.size() && to.size() < kR; i++) to.push_back(add[i]);
    };
    bool going = true;
    while (going) {
      char op = 0, seft = 0;
      std::string p1, p2;
      if (!expr_break(lhs, &op, &seft, &p1, &p2)) break;
      switch (op) {
        case '+': cat(rhs, p2); cat(rhs, "-"); lhs = p1; break;
        case '-': cat(rhs, p2); cat(rhs, "+"); lhs = p1; break;
        case '*': cat(rhs, p2); cat(rhs, "/"); lhs = p1; break;
        case '/': cat(rhs, p2); cat(rhs, "*"); lhs = p1; break;
        case '^': cat(rhs, p2); cat(rhs, "v"); lhs = p1; break;
        case 'L': { std::string t = p2; cat(t, rhs); cat(t, "^"); rhs = t; lhs = p1; break; }
        case 'v': cat(rhs, p2); cat(rhs, "^"); lhs = p1; break;
        case 'E': cat(rhs, "l"); lhs = p1; break;
        case 'l': cat(rhs, "E"); lhs = p1; break;
        case 'n': cat(rhs, "n"); lhs = p1; break;
        case 'q': cat(rhs, "s"); lhs = p1; break;
        case 'r': cat(rhs, "r"); lhs = p1; break;
        case 's': {
          cat(rhs, "q"); lhs = p1;
          double l = eval_string(lhs), r = eval_string(rhs);
          if (l * r < 0) cat(rhs, "n");
          break;
        }
        default: going = false; break;
      }
    }
  }

  /* report_match ries.c:7887 */
  void report_match(const std::string& le_in, const std::string& re_in, double root, double delta) {
    if ((delta == 0.0) && got_exact) return;
    if (!opt.refinement) { /* prune on the delta, then on the equation itself (7919-7953) */
      size_t i = 0;
      for (; i < (size_t)num_matches; i++)
        if (nr_deltas[i] == delta) return;
      if (i < nr_deltas.size()) nr_deltas[i] = delta;
      nr_bytes += sizeof(double);
      if (!unique_eqn(le_in, re_in, true)) return;
    }
    std::string le = le_in, re = re_in;
    if (opt.solve_for_x) {
      try_solve(le, re);
      if (!opt.refinement && !unique_eqn(le, re, true)) return; /* 8013 */
    }
    std::string l_fmt, r_fmt;
    if (opt.out_format == OF_POSTFIX) {
      l_fmt = le; r_fmt = re;
    } else if (opt.out_format == OF_CONDENSED || opt.out_format == OF_NORMAL) {
      std::string cond[2];
      bool okk[2];
      for (int side = 0; side < 2; side++) {
        std::string pre = infix_preproc(icx, side ? re : le);
        unsigned char top;
        okk[side] = infix_1(icx, pre, cond[side], top);
        if (side == 0) { /* the LHS is finished (expanded) before the RHS is started (8033-8047) */
          if (!okk[0]) l_fmt.clear();
          else l_fmt = (opt.out_format == OF_NORMAL) ? infix_expand(icx, cond[0]) : cond[0];
        }
      }
      if (!opt.refinement && !unique_eqn(cond[0], cond[1], true)) return; /* 8062 */
      if (!okk[1]) r_fmt.clear();
      else r_fmt = (opt.out_format == OF_NORMAL) ? infix_expand(icx, cond[1]) : cond[1];
    } else {
      l_fmt = postfix_formatter(le); r_fmt = postfix_formatter(re);
    }
    int lf_len = (int)l_fmt.size(), rf_len = (int)r_fmt.size();
    int width = opt.relative_x ? 44 : 40;
    int equ_width = 5;
    int t_width = width + equ_width;
    int l_pad = (opt.one_sided || opt.anagram || opt.solve_for_x) ? width / 8 : width / 2; /* ries.c:8113 */
    int r_pad = width - l_pad;
    l_pad -= lf_len; if (l_pad < 0) l_pad = 0;
    r_pad -= rf_len; if (r_pad < 0) r_pad = 0;
    width = l_pad + lf_len + equ_width + rf_len + r_pad;
    while ((width > t_width) && (l_pad > 0)) { l_pad--; width--; }
    while ((width > t_width) && (r_pad > 0)) { r_pad--; width--; }
    int justify_overflow = width - t_width;
    if (justify_overflow < 0) justify_overflow = 0;

    std::string line(l_pad, ' ');
    line += " " + l_fmt + " " + S.any_name[(unsigned char)'='] + " " + r_fmt + " ";
    line.append(r_pad, ' ');

    std::string froot = spfg(S.k_usable_digits, root);
    const char* x_T_intro = "for x ";
    if (opt.wide_output) {
      char tmp[64];
      snprintf(tmp, sizeof(tmp), "x = %-20s  ", froot.c_str());
      line += tmp;
      x_T_intro = "";
    }
    char fval[32];
    fval[0] = 0;
    int posn = 0;
    double adelta = delta;
    if (delta == 0.0) {
      line += "('exact' match)";
      posn = 15;
      got_exact = true;
    } else if (!opt.relative_x) {
      line += "for x = ";
      snprintf(fval, sizeof(fval), "%s", froot.c_str());
      posn = 8 + (int)strlen(fval);
    } else if (delta < 0.0) {
      adelta = -delta;
      line += x_T_intro; line += "= T - ";
      snprintf(fval, 30, "%.6g", adelta);
      posn = 6 + (int)strlen(x_T_intro) + (int)strlen(fval);
    } else {
      line += x_T_intro; line += "= T + ";
      snprintf(fval, 30, "%.6g", adelta);
      posn = 6 + (int)strlen(x_T_intro) + (int)strlen(fval);
    }
    if (opt.relative_x) width = 12 + 6 + 5;
    else width = 8 + S.k_usable_digits + 6;
    line += fval;
    posn += justify_overflow;
    for (int i = posn; i < width; i++) line.push_back(' ');
    if (opt.wide_output) {
      char temp[64];
      double ratio = 0;
      if (adelta != 0) ratio = fabs(S.P.target / adelta);
      if (adelta == 0) snprintf(temp, 30, "%s", "(1 part in infinity)");
      else if (ratio < 100) {
        int r10 = (int)((ratio * 10) + 0.5);
        int r1 = r10 % 10; r10 = (r10 - r1) / 10;
        snprintf(temp, 30, "(1 part in %d.%d)", r10, r1);
      } else if (ratio < 100000000) snprintf(temp, 30, "(1 part in %d)", (int)(ratio + 0.5));
      else snprintf(temp, 30, "(1 part in %.3e)", ratio);
      char t2[64];
      snprintf(t2, sizeof(t2), "%-22s", temp);
      line += t2;
    }
    char tail[32];
    snprintf(tail, sizeof(tail), " {%d}\n", complexity(le) + complexity(re));
    line += tail;
    out.write(line.data(), line.size());

    num_matches++;
    defsym_used(le);
    defsym_used(re);
    check_exit(delta == 0.0);
  }

  /* check_exact_match ries.c:8541; `tree` is the row already in the table */
  void replay_exact(const MatchEvent& ev, int using_x) {
    const Row& tree = using_x ? ev.rhs : ev.lhs;
    const Row& fresh = using_x ? ev.lhs : ev.rhs;
    double total_deriv;
    if (!using_x) { /* new item is RHS, tree item must be an LHS */
      if (tree.der == 0.0) return;
      if (fabs(tree.der) < S.P.k_vanished_dx) return;
      total_deriv = tree.der;
    } else {
      if (tree.der != 0.0) return;
      if (fabs(fresh.der) < S.P.k_vanished_dx) return;
      total_deriv = 0 - fresh.der;
    }
    double x = tree.val, eps = x * 1.0e-14, xpe = x + eps;
    if (x == xpe) return;
    if (total_deriv == 0) return;
    double delta = eps / total_deriv;
    double score = fabs(delta);
    if ((best_match > 0) && (score > best_match)) return;
    if (score < opt.min_matchsize) return;
    if (!opt.refinement && !unique_eqn(sym_chars(ev.lhs.sym), sym_chars(ev.rhs.sym), false)) return; /* 8611 */
    report_match(sym_chars(ev.lhs.sym), sym_chars(ev.rhs.sym), S.P.target, 0.0);
  }

  /* check_match ries.c:8243 */
  void replay_neighbour(const MatchEvent& ev) {
    double total_deriv = ev.lhs.der - ev.rhs.der;
    if (total_deriv == 0.0) {
      fprintf(stderr, "check_match got dx = 0!\n");
      return;
    }
    double delta = (ev.rhs.val - ev.lhs.val) / total_deriv;
    double score = fabs(delta);
    if ((score > best_match) || (score < opt.min_matchsize)) return;
    if (!opt.refinement && !unique_eqn(sym_chars(ev.lhs.sym), sym_chars(ev.rhs.sym), false)) return; /* 8280 */
    if (ev.newton_err) return;
    delta = ev.root - S.P.target;
    score = fabs(delta);
    std::string le = sym_chars(ev.lhs_s), re = sym_chars(ev.rhs_s);
    if (score == 0.0) {
      report_match(le, re, ev.root, delta);
      return;
    } else if (score < best_match) {
      report_match(le, re, ev.root, delta);
      if (opt.refinement) best_match = (score * 0.999) - S.k_min_best_match; /* 8351 */
      check_exit(false);
    }
  }

  int run() {
    build_setup(opt, S, out);
    if (opt.show_ss) { /* ries.c:12100 */
      show_symset(opt, S, out);
      out.printf("%s: %s", opt.argv0.c_str(),
                 "Exiting now (to do an equation search, omit '-S' or include\n"
                 "  symbol names, for example: 'ries 3.14159 -S123456789+/')\n");
      return 0;
    }
    finish_setup(opt, S, out);
    icx.S = &S;
    icx.explicit_multiply = opt.explicit_multiply;
    best_match = S.init_match_dist;
    nr_init();

    /* header ries.c:12327-12337 */
    {
      char fmt[32], buf[128];
      out.printf("\n");
      out.printf("   Your target value: T = ");
      snprintf(fmt, sizeof(fmt), "%%-%ds", S.k_usable_digits + 6);
      snprintf(buf, sizeof(buf), fmt, spfg(S.k_usable_digits, S.tgt_to_print).c_str());
      out.printf("%s", buf);
      snprintf(fmt, sizeof(fmt), "  %%%ds", 44 - S.k_usable_digits);
      out.printf(fmt, "mrob.com/ries");
      out.printf("%s", "\n\n");
    }

    double lim_gentotal = 2000.0;
    if (opt.tlevel >= -2.0) lim_gentotal *= pow(4.0, opt.tlevel);
    double lim_prune = lim_gentotal * 15.0;
    if (opt.debug_y) out.printf("Will stop after generating searchmax=%g expressions\n", lim_gentotal);

    eng.begin(S.P);
    const bool one_sided = opt.one_sided || opt.anagram; /* ries.c:12319 */
    bool reported_exhaustion = false;
    PassOutput po;
    bool going = true;
    while (going) {
      int using_x;
      int cplx;
      /* --one-sided: 'x' is the sole LHS expression, RHS passes from then on (ries.c:12425) */
      if ((one_sided && st.lhs_insert >= 1) || (st.lhs_insert > st.rhs_insert)) {
        rmin = (int16_t)(rmin + 1); rmax = (int16_t)(rmax + 1); using_x = 0; cplx = rmax;
      }
      else { lmin = (int16_t)(lmin + 1); lmax = (int16_t)(lmax + 1); using_x = 1; cplx = lmax; }
      double genf = 0;
      double tf0 = now_sec();
      const std::vector<FormDesc>& forms = forms_for(cplx, &genf);
      po = PassOutput();
      cur_ev = 0;
      double tf1 = now_sec();
      st.forms_seconds += tf1 - tf0;
      /* The engine indexes forms with 16 bits.  A pass with more forms (extreme --symbol-weights) is fed in
         consecutive slices: any split of a pass that keeps insertion order is as good as the whole, since
         the rows of a pass are all one side and same-side neighbours are never tested (ries.c:8496). */
      size_t kMaxForms = 60000;
      if (const char* e = getenv("RIES_FORMS_SLICE")) kMaxForms = (size_t)std::max(1, atoi(e)); /* test hook */
      for (size_t f0 = 0; f0 < forms.size() || f0 == 0; f0 += kMaxForms) {
        PassOutput part;
        double te0 = now_sec();
        if (forms.size() <= kMaxForms) {
          eng.run_pass(using_x, cplx, forms, best_match, got_exact, part);
        } else {
          std::vector<FormDesc> slice(forms.begin() + f0, forms.begin() + std::min(forms.size(), f0 + kMaxForms));
          eng.run_pass(using_x, cplx, slice, best_match, got_exact, part);
        }
        st.engine_seconds += now_sec() - te0;
        /* replay the candidates in insertion order with the live threshold */
        for (size_t k = 0; k < part.events.size(); k++) {
          cur_ev = &part.events[k];
          if (cur_ev->kind == EV_EXACT) replay_exact(*cur_ev, using_x);
          else replay_neighbour(*cur_ev);
        }
        cur_ev = 0;
        resident_rows += part.inserted;
        po.generated += part.generated; po.dead_ends += part.dead_ends; po.inserted += part.inserted;
        if (forms.size() <= kMaxForms) break;
      }
      st.passes++;
      if (using_x) {
        if (opt.debug_y) {
          out.printf("finished LHS from {%d} to {%d}: %g forms expanded, %g expressions.\n", lmin, lmax,
                     genf, po.generated);
          out.printf("\n");
        }
        st.lhs_gen += po.generated; st.lhs_insert += po.inserted; st.lhs_prune += po.dead_ends;
      } else {
        if (opt.debug_y)
          out.printf("finished RHS from {%d} to {%d}: %g forms expanded, %g expressions.\n", rmin, rmax,
                     genf, po.generated);
        st.rhs_gen += po.generated; st.rhs_insert += po.inserted; st.rhs_prune += po.dead_ends;
      }
      st.gen_total += po.generated;

      if ((lmax > S.P.exhaust_cpx) && (one_sided || (rmax > S.P.exhaust_cpx))) { /* 12484 */
        if (num_matches == 0) {
          out.printf("exhaustion timeout (complexity LHS {%d}, RHS {%d}). Possible causes:\n"
                     "  MAX_ELEN %d too small\n", lmax, rmax, RIES_MAX_ELEN);
          if (opt.NOS_options) out.printf("  too many symbols excluded via -N, -O or -S\n");
          if (opt.anagram) out.printf("  --numeric-anagram too restrictive\n");
          if (opt.derivative_margin == 0) /* ries.c:12504 */
            out.printf("  target too big or small for k_vanished_dx (use --derivative-margin)\n");
          reported_exhaustion = true;
        }
        st.gen_total = lim_gentotal;
      }
      if (opt.debug_y) {
        out.printf("gen_total %g (lhs %ld rhs %ld) searchmax %g gnm %ld\n", st.gen_total, st.lhs_insert,
                   st.rhs_insert, lim_gentotal, (long)num_matches);
        /* bt_stats (ries.c:8399-8428, called at 12515): the node count is the table's row count.  The average
           node depth describes the shape of the reference's unbalanced tree; the table here is a sorted array
           and has no such shape -- a listed difference, except for the empty tree. */
        out.printf("Current tree stats:\n");
        out.printf("       Nodes: %ld\n", (long)resident_rows);
        if (resident_rows == 0) out.printf("  Avg. Depth: -nan\n"); /* 0/0 on x86-64 */
        else out.printf("  Avg. Depth: (not tracked: the value table is a sorted array)\n");
      }
      if (num_matches == 0) { /* 12522 */
        double tsec = now_sec() - t_start;
        if (opt.debug_y || ((tsec > 2.0) && opt.allow_slow_message)) {
          out.printf("  Still searching: %11s expr and %.3f sec so far...     \r",
                     pf_intfloat_wid(st.gen_total, 11).c_str(), tsec);
        }
      }
      if ((st.lhs_prune + st.rhs_prune) >= lim_prune) going = false;
      if (st.gen_total >= lim_gentotal) going = false;
    }

    if (opt.debug_y) /* ries.c:12550 */
      out.printf("Level %.g search complete (gen_total exceeded lim_gentotal).\n", opt.levadj);
    if (reported_exhaustion) {
    } else if (num_matches == 0) {
      out.printf("No solution was found (try using -l%d", (int)floor(opt.levadj + 1.0));
      if (S.init_match_dist < (0.001 * S.P.target)) out.printf(" or a larger --max-match-distance");
      out.printf(").  \n");
    } else if (opt.exact_exit) {
      out.printf("               No 'exact' solution was found (try using -l%d).\n",
                 (int)floor(opt.levadj + 1.0));
    } else {
      out.printf("                  (for more results, use the option '-l%d')\n",
                 (int)floor(opt.levadj + 1.0));
    }
    print_end(0);
    return 0;
  }
};

}  // namespace

static void null_sink_fn(void*, const char*, size_t) {}

bool search_params(int argc, char** argv, RiesParams* outp) {
  Sink sink{null_sink_fn, 0};
  try {
    Options opt;
    parse_args(argc, argv, opt, sink);
    if (opt.show_ss) return false;
    Setup S;
    build_setup(opt, S, sink);
    finish_setup(opt, S, sink);
    *outp = S.P;
    return true;
  } catch (const ExitRequest&) {
    return false;
  }
}

int run_search(Options& opt, Engine& eng, const Sink& out, RunStats* stats_out) {
  Search s(opt, eng, out);
  int code = 0;
  try {
    code = s.run();
  } catch (const ExitRequest& e) {
    code = e.code;
  }
  s.st.lmax = s.lmax;
  s.st.rmax = s.rmax;
  s.st.exec_seconds = now_sec() - s.t_start;
  if (stats_out) *stats_out = s.st;
  return code;
}

}  // namespace ries
