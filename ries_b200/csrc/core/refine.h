// Candidate refinement: cv_simplify (/root/reference/public/ries.c:7297-7415) followed by
// newton (7139-7286).  A pure function of (lhs, rhs, target): it does not depend on the live match
// threshold, so the device can run it for every candidate of a pass before the host replays them
// in insertion order.
#pragma once
#include "exec.h"
#include "params.h"

#define RIES_NEWTON_MAX_ITER 12 /* ries.c:7125 */

struct SymStr {
  unsigned char s[RIES_EXPR_ALLOC + 2];
  int n;
};

RIES_HD void symstr_from(const RiesParams& P, const Sym128& p, SymStr& o) {
  o.n = 0;
  for (int i = 0; i < RIES_MAX_ELEN + 1; i++) {
    unsigned c = sym_get(p, i);
    if (!c) break;
    o.s[o.n++] = P.sym_char[c - 1];
  }
  o.s[o.n] = 0;
}

RIES_HD unsigned sym_code_of(const RiesParams& P, unsigned char ch) {
  for (int g = 0; g < P.nsym; g++)
    if (P.sym_char[g] == ch) return (unsigned)(g + 1);
  return 0;
}

RIES_HD void symstr_to(const RiesParams& P, const SymStr& i, Sym128& o) {
  o.lo = 0; o.hi = 0;
  for (int k = 0; k < i.n && k <= RIES_MAX_ELEN; k++) sym_set(o, k, sym_code_of(P, i.s[k]));
}

/* symstrtrail ries.c:6363 */
RIES_HD bool ss_trail(const SymStr& big, const char* little, int l2) {
  if (big.n == 0 || l2 == 0 || big.n < l2) return false;
  for (int i = 0; i < l2; i++)
    if (big.s[big.n - l2 + i] != (unsigned char)little[i]) return false;
  return true;
}
/* symstrclip ries.c:6393 */
RIES_HD void ss_clip(SymStr& s, int len) {
  if (s.n >= len) { s.n -= len; s.s[s.n] = 0; }
}
/* symstrncat with TS_ALLOC_L ries.c:6502 */
RIES_HD void ss_cat(SymStr& s, const char* add, int l2) {
  for (int i = 0; i < l2 && (s.n + 1 < RIES_EXPR_ALLOC); i++) s.s[s.n++] = (unsigned char)add[i];
  s.s[s.n] = 0;
}

/* The rewrite loop of cv_simplify (ries.c:7312-7400). */
RIES_HD void cv_simplify_syms(SymStr& lhs, SymStr& rhs) {
  bool gg = true;
  while (gg) {
    gg = false;
    if (ss_trail(lhs, "2/", 2) && ss_trail(rhs, "2/", 2)) { ss_clip(lhs, 2); ss_clip(rhs, 2); gg = true; }
    if (ss_trail(lhs, "2*", 2) && ss_trail(rhs, "2*", 2)) { ss_clip(lhs, 2); ss_clip(rhs, 2); gg = true; }
    if (ss_trail(lhs, "r", 1) && ss_trail(rhs, "r", 1)) { ss_clip(lhs, 1); ss_clip(rhs, 1); gg = true; }
    if (ss_trail(lhs, "n", 1) && ss_trail(rhs, "n", 1)) { ss_clip(lhs, 1); ss_clip(rhs, 1); gg = true; }
    if (ss_trail(lhs, "nr", 2) && ss_trail(rhs, "n", 1)) {
      ss_clip(lhs, 2); ss_cat(lhs, "r", 1); ss_clip(rhs, 1); gg = true;
    }
    if (ss_trail(lhs, "n", 1) && ss_trail(rhs, "nr", 2)) {
      ss_clip(lhs, 1); ss_clip(rhs, 2); ss_cat(rhs, "r", 1); gg = true;
    }
    if (ss_trail(lhs, "rr", 2)) { ss_clip(lhs, 2); gg = true; }
    if (ss_trail(rhs, "rr", 2)) { ss_clip(rhs, 2); gg = true; }
    if (ss_trail(lhs, "r2/", 3) && ss_trail(rhs, "r", 1)) {
      ss_clip(lhs, 3); ss_cat(lhs, "2*", 2); ss_clip(rhs, 1); gg = true;
    }
    if (ss_trail(lhs, "r", 1) && ss_trail(rhs, "r2/", 3)) {
      ss_clip(lhs, 1); ss_clip(rhs, 3); ss_cat(rhs, "2*", 2); gg = true;
    }
  }
}

RIES_HD bool sym_has_x(const RiesParams& P, const Sym128& s) {
  for (int i = 0; i < RIES_MAX_ELEN + 1; i++) {
    unsigned c = sym_get(s, i);
    if (!c) break;
    if (P.sym_char[c - 1] == 'x') return true;
  }
  return false;
}

/* newton() ries.c:7139-7286.  Returns 0 / ERR_*; *root defaults to the target. */
RIES_HD int newton_solve(const RiesParams& P, const Sym128& lhs, const Sym128& rhs, double* root,
                         double* diff_dx) {
  Val R, L;
  double curr, prev, prv2;
  int err;
  *root = P.target;
  *diff_dx = 0;
  bool rhs_has_x = sym_has_x(P, rhs);
  R.v = 0; R.d = 0; R.tag = 0;
  err = eval_expr(P, rhs, P.target, R);
  if (err) return err;
  L.v = 0; L.d = 0; L.tag = 0;
  curr = P.target; prev = curr - 1.0; prv2 = prev - 1.0;
  for (int i = 0; i < RIES_NEWTON_MAX_ITER; i++) {
    prv2 = prev; prev = curr;
    err = eval_expr(P, lhs, curr, L);
    if (err) return err;
    if (L.d == 0.0) return ERR_NEWTON_ZERO_DX;
    if (rhs_has_x) {
      err = eval_expr(P, rhs, curr, R);
      if (err) return err;
    }
    curr = curr + ((R.v - L.v) / (L.d - R.d));
    if (curr == prev) {
      i = RIES_NEWTON_MAX_ITER + 1;
    } else if (curr == prv2) {
      double margin = (fabs(curr - prev) * 9007199254740992.0 / fabs(curr));
      if (margin <= 2.0) {
        curr = prev; i = RIES_NEWTON_MAX_ITER + 1;
      } else if ((i >= 6) && (margin <= 4.0)) {
        curr = prev; i = RIES_NEWTON_MAX_ITER + 1;
      }
    }
  }
  if (curr == prev) {
    *root = curr; *diff_dx = L.d - R.d;
    return 0;
  } else if (fabs(curr - prev) < fabs(P.target * P.k_newton_settled)) {
    *root = curr; *diff_dx = L.d - R.d;
    return 0;
  }
  return ERR_NEWTON_NO_CONVERGE;
}

/* check_match's second stage up to the Newton result (ries.c:8292-8314). */
RIES_HD void refine_event(const RiesParams& P, MatchEvent& ev) {
  SymStr l, r;
  symstr_from(P, ev.lhs.sym, l);
  symstr_from(P, ev.rhs.sym, r);
  if (!P.no_cv_simplify) cv_simplify_syms(l, r);
  symstr_to(P, l, ev.lhs_s);
  symstr_to(P, r, ev.rhs_s);
  ev.newton_err = newton_solve(P, ev.lhs_s, ev.rhs_s, &ev.root, &ev.diff_dx);
}
