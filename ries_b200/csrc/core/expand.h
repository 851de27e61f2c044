// One enumeration step: try candidate symbol `slot` after one live prefix.
//
// This is ge_2 (/root/reference/public/ries.c:8908-9475) unrolled into a data-parallel step: the
// reference's recursion "choose symbol (9388-9465) -> recurse -> exec (8946) -> value prunes
// (8962-9076) -> complete? (9085-9190)" becomes a pure function of (parent prefix, slot) so that one
// GPU thread can own one (prefix, symbol) pair.  The function reports exactly one outcome per pair, which
// is what makes the dead-end counter (prune_count, 8970..9137 and 9461) reproducible:
//   OUT_NONE  slot beyond the class's symbol list (no such iteration in the reference)
//   OUT_DEAD  one dead-end (symbol-loop rejection OR value prune)
//   OUT_ERR   exec() error: abandoned WITHOUT counting (8947-8953)
//   OUT_LIVE  a longer live prefix
//   OUT_DONE  a complete expression handed to the value table (counts as "generated", 9190)
#pragma once
#include "exec.h"
#include "params.h"

/* Attribute bits of a prefix (ries.c:9240-9379); independent of the candidate symbol. */
RIES_HD int prefix_attrs(const RiesParams& P, const FormDesc& F, const Node& par, int using_x) {
  int ip = par.len;
  int atts = using_x ? 0 : AM_RHS;
  if (ip > 2) {
    if (F.cls[ip - 1] == 0 && F.cls[ip - 2] == 2 && F.cls[ip - 3] == 0) {
      unsigned char s1 = P.sym_char[sym_get(par.sym, ip - 1) - 1];
      unsigned char s2 = P.sym_char[sym_get(par.sym, ip - 2) - 1];
      unsigned char s3 = P.sym_char[sym_get(par.sym, ip - 3) - 1];
      if (s1 == s3) {
        if (s2 == '*') atts |= AM_KxK;
        else if (s2 == '+') atts |= AM_KpK;
      }
    }
  }
  if (ip > 1) {
    if (F.cls[ip] == 2) {
      int a1i = F.arg1[ip];
      if (a1i >= 0) {
        unsigned char c = P.sym_char[sym_get(par.sym, a1i) - 1];
        if (c == 'e') atts |= AM_a1_e;
        else if (c == '1') atts |= AM_a1_1;
        else if (c == 'r') atts |= AM_a1_r;
      }
    }
    if (F.cls[ip - 1] == 0 && F.cls[ip - 2] == 0) {
      unsigned char s1 = P.sym_char[sym_get(par.sym, ip - 1) - 1];
      unsigned char s2 = P.sym_char[sym_get(par.sym, ip - 2) - 1];
      if (s1 == s2) atts |= AM_KK;
      if ((s1 < '6') && (s2 < '6')) atts |= AM_55;
      if (s2 == '1') atts |= AM_1K;
      if (s1 > s2) atts |= AM_jK;
    }
  }
  if (ip > 0) atts |= P.sym_amkey[sym_get(par.sym, ip - 1) - 1];
  return atts;
}

/* Symbol-append tests of ge_2's loop, in the reference's order (ries.c:9401-9445): complexity bounds,
   rule mask, LHS-starts-with-x, -O symbol count.  false = one dead-end. */
RIES_HD bool expand_tests(const RiesParams& P, const FormDesc& F, const Node& par, int slot, int atts,
                          int using_x, int e_minw, int e_maxw) {
  int ip = par.len;
  int g = P.trial[F.cls[ip]][slot];
  int t_cpx = par.cplx + P.sym_wgt[g];
  if (t_cpx + F.rminw[ip + 1] > e_maxw) return false;
  if (t_cpx + F.rmaxw[ip + 1] < e_minw) return false;
  if (atts & P.sym_mask[g]) return false;
  if (using_x && (ip == 0) && (P.sym_char[g] != 'x')) return false;
  if (P.sym_alwd[g] < RIES_MAX_ELEN) {
    int ct = 0;
    for (int i = 0; i < ip; i++) ct += (sym_get(par.sym, i) == (unsigned)(g + 1));
    if (ct >= P.sym_alwd[g]) return false;
  }
  return true;
}

/* The same tests for ALL candidate slots of a prefix at once, as bit masks (bit k = slot k survives).
   Per class, host-built tables list the slots by weight bound, by rule-mask bit, the 'x' slot and the
   slots with an -O limit; only the last needs a per-slot look at the prefix.  Must agree with
   expand_tests() slot by slot (the CPU oracle uses expand_tests; the parity tests compare). */
#define RIES_TT_W 104 /* weights 0..103: --symbol-weights caps a weight at 100 (ries.c:3197) */
struct TestTables {
  unsigned short wge[3][RIES_TT_W]; /* slots whose symbol weight is >= w */
  unsigned short wle[3][RIES_TT_W]; /* ... <= w */
  unsigned short attr[3][32]; /* slots whose rule mask has attribute bit b */
  unsigned short xslot[3];    /* slots holding 'x' */
  unsigned short limited[3];  /* slots whose symbol has an -O count limit */
};

inline void build_test_tables(const RiesParams& P, TestTables& T) {
  for (int c = 0; c < 3; c++) {
    for (int w = 0; w < RIES_TT_W; w++) { T.wge[c][w] = 0; T.wle[c][w] = 0; }
    for (int b = 0; b < 32; b++) T.attr[c][b] = 0;
    T.xslot[c] = 0;
    T.limited[c] = 0;
    for (int k = 0; k < P.cls_n[c]; k++) {
      const int g = P.trial[c][k];
      const int wt = P.sym_wgt[g];
      for (int w = 0; w < RIES_TT_W; w++) {
        if (wt >= w) T.wge[c][w] |= (unsigned short)(1u << k);
        if (wt <= w) T.wle[c][w] |= (unsigned short)(1u << k);
      }
      for (int b = 0; b < 32; b++)
        if ((unsigned)P.sym_mask[g] & (1u << b)) T.attr[c][b] |= (unsigned short)(1u << k);
      if (P.sym_char[g] == 'x') T.xslot[c] |= (unsigned short)(1u << k);
      if (P.sym_alwd[g] < RIES_MAX_ELEN) T.limited[c] |= (unsigned short)(1u << k);
    }
  }
}

/* true when every enabled symbol's weight fits the tables (always, short of --symbol-weights) */
inline bool test_tables_cover(const RiesParams& P) {
  for (int c = 0; c < 3; c++)
    for (int k = 0; k < P.cls_n[c]; k++) {
      const int wt = P.sym_wgt[P.trial[c][k]];
      if (wt < 0 || wt >= RIES_TT_W) return false;
    }
  return true;
}

#ifdef __CUDA_ARCH__
#define RIES_CTZ(x) (__ffs((int)(x)) - 1)
#else
#define RIES_CTZ(x) __builtin_ctz(x)
#endif

RIES_HD unsigned expand_mask(const RiesParams& P, const TestTables& T, const FormDesc& F, const Node& par,
                             int cls, int atts, int using_x, int e_minw, int e_maxw) {
  const int ip = par.len;
  const int whi = e_maxw - F.rminw[ip + 1] - par.cplx; /* weight <= whi  <=>  t_cpx + rminw <= e_maxw */
  const int wlo = e_minw - F.rmaxw[ip + 1] - par.cplx; /* weight >= wlo  <=>  t_cpx + rmaxw >= e_minw */
  if (whi < 0 || wlo >= RIES_TT_W) return 0;
  unsigned m = (unsigned)T.wge[cls][wlo < 0 ? 0 : wlo] & (unsigned)T.wle[cls][whi >= RIES_TT_W ? RIES_TT_W - 1 : whi];
  unsigned a = (unsigned)atts;
  while (a && m) {
    const int b = RIES_CTZ(a);
    a &= a - 1;
    m &= ~(unsigned)T.attr[cls][b];
  }
  if (using_x && ip == 0) m &= T.xslot[cls];
  unsigned lim = m & T.limited[cls];
  while (lim) {
    const int k = RIES_CTZ(lim);
    lim &= lim - 1;
    const int g = P.trial[cls][k];
    int ct = 0;
    for (int i = 0; i < ip; i++) ct += (sym_get(par.sym, i) == (unsigned)(g + 1));
    if (ct >= P.sym_alwd[g]) m &= ~(1u << k);
  }
  return m;
}

/* canonval (ries.c:8753-8892): after a complete expression has passed every prune, up to four more
   operators are appended so that its value moves towards [1,2): [n] when negative, [r] when |v| < 1,
   [2/] when |v| >= 2, [2*] when |v| < 1 -- each only if its symbols may still be used (sa_ct < sa_alwd,
   the counts being those of the expression before canonval touched it) and the longer expression fits
   MAX_ELEN.  `r` is updated in place; the returned code (bit per operator, same bits as canon_ops) tells
   expand_finish() which symbols to append.  An exec() error abandons the expression uncounted (9145-9151;
   the reference also leaves its metastack one undo short there -- no search reaches that case: 1/v of
   a value that has passed the prunes, v/2 and v*2 do not fail). */
enum { CANON_NEG = 1, CANON_RECIP = 2, CANON_DIV2 = 4, CANON_MUL2 = 8 };

RIES_HD bool canon_may_use(const RiesParams& P, int which, const Sym128& sym, int len) {
  const int g = P.canon_sym[which];
  if (g < 0) return false;
  const int alwd = P.sym_alwd[g];
  if (alwd >= RIES_MAX_ELEN) return true;
  int ct = 0;
  for (int i = 0; i < len; i++) ct += (sym_get(sym, i) == (unsigned)(g + 1));
  return ct < alwd;
}

RIES_HD int canon_apply(const RiesParams& P, bool do_dx, const Sym128& sym, int len, Val& r, int& code) {
  int ip = len;
  code = 0;
  Val t;
  if ((P.canon_ops & CANON_NEG) && canon_may_use(P, 0, sym, len) && (r.v < 0.0) && (ip < RIES_MAX_ELEN)) {
    if (exec_unary(P, 'n', do_dx, r, t)) return 1;
    r = t; ip += 1; code |= CANON_NEG;
  }
  if ((P.canon_ops & CANON_RECIP) && canon_may_use(P, 1, sym, len) && (r.v * r.v < 1.0) && (ip < RIES_MAX_ELEN)) {
    if (exec_unary(P, 'r', do_dx, r, t)) return 1;
    r = t; ip += 1; code |= CANON_RECIP;
  }
  if ((P.canon_ops & (CANON_DIV2 | CANON_MUL2)) && canon_may_use(P, 2, sym, len)) {
    Val two;
    if ((P.canon_ops & CANON_DIV2) && canon_may_use(P, 3, sym, len) && (r.v * r.v >= 4.0) && (ip + 1 < RIES_MAX_ELEN)) {
      if (exec_push(P, '2', P.target, two)) return 1;
      if (exec_binary(P, '/', do_dx, r, two, t)) return 1;
      r = t; ip += 2; code |= CANON_DIV2;
    }
    if ((P.canon_ops & CANON_MUL2) && canon_may_use(P, 4, sym, len) && (r.v * r.v < 1.0) && (ip + 1 < RIES_MAX_ELEN)) {
      if (exec_push(P, '2', P.target, two)) return 1;
      if (exec_binary(P, '*', do_dx, r, two, t)) return 1;
      r = t; ip += 2; code |= CANON_MUL2;
    }
  }
  return 0;
}

/* exec of the chosen symbol on the cactus stack (the child's ge_2 entry, 8942-8954), the value prunes
   (8962-9076) and the complete-expression tests (9085-9139).  `nodes` is the pass's node arena
   (parents of every level stay resident for `below` links); `par_idx` is the parent's own index.
   Fills only the evaluation part of `child` (val, der, tag, below, sp). */
/* `track_m` (RHS passes of a batch that shares its right-hand sides, engine.cu SharedRhs): an RHS prefix has no
   derivative, and its `der` field carries instead M = the largest |value| any exec() along the prefix has
   produced -- what decides, target by target, whether the prefix survives that target's overflow bound
   (|v| >= p_ovr, ries.c:9020).  The operands then enter exec() with derivative 0, as they would have. */
RIES_HD int expand_exec(const RiesParams& P, int cls, int g, bool complete, const Node* nodes,
                        const Node& par, uint32_t par_idx, int using_x, Node& child, bool track_m = false) {
  unsigned char sym = P.sym_char[g];
  Val r;
  int err;
  uint32_t below;
  int sp;
  if (cls == 0) {
    err = exec_push(P, sym, P.target, r);
    below = par_idx;
    sp = par.sp + 1;
  } else if (cls == 1) {
    Val a = {par.val, track_m ? 0.0 : par.der, par.tag};
    err = exec_unary(P, sym, using_x != 0, a, r);
    below = par.below;
    sp = par.sp;
  } else {
    const Node& lo = nodes[par.below];
    Val a = {lo.val, track_m ? 0.0 : lo.der, lo.tag};
    Val b = {par.val, track_m ? 0.0 : par.der, par.tag};
    err = exec_binary(P, sym, using_x != 0, a, b, r);
    below = lo.below;
    sp = par.sp - 1;
  }
  if (err) return OUT_ERR;

  /* value prunes (8962-9076); each is one dead-end */
  if (r.v == 0.0) return OUT_DEAD;
  if (P.restrict_subexpr == TYPE_INT) {
    if (r.tag != TYPE_INT) return OUT_DEAD;
  }
  if (r.tag < P.restrict_subexpr) return OUT_DEAD;
  if ((r.v >= P.p_ovr) || (r.v <= P.n_ovr)) return OUT_DEAD;
  if (using_x) {
    if ((r.d != 0.0) && (fabs(r.d) / (1.0 + fabs(r.v)) < P.k_vanished_dx)) return OUT_DEAD;
    if ((sp == 1) && (fabs(r.d) / (1.0 + fabs(r.v)) < P.k_vanished_dx)) return OUT_DEAD;
  }
  child.pad0 = 0;
  const double m_here = track_m ? fmax(par.der, fabs(r.v)) : 0.0; /* canonval's results are not overflow-tested */
  if (complete) {
    /* complete expression (9085-9139) */
    if (using_x && (fabs(r.d) < P.k_vanished_dx)) return OUT_DEAD;
    if (r.v < P.min_equ_val) return OUT_DEAD;
    if (r.v > P.max_equ_val) return OUT_DEAD;
    if (P.canon_ops) {
      /* the value the table sees is the canonical one (9144); pad0 carries which symbols to append */
      Sym128 full = par.sym;
      sym_set(full, par.len, (unsigned)(g + 1));
      int code;
      if (canon_apply(P, using_x != 0, full, par.len + 1, r, code)) return OUT_ERR;
      child.pad0 = (uint8_t)code;
    }
  }
  child.val = r.v;
  child.der = track_m ? m_here : r.d;
  child.tag = (uint8_t)r.tag;
  child.below = below;
  child.sp = (uint8_t)sp;
  return complete ? OUT_DONE : OUT_LIVE;
}

/* The bookkeeping part of a child: symbols, form, complexity, length.  child.pad0 is canon_apply's code
   (nonzero only for complete expressions under --canon-reduction): its symbols are appended here; the
   complexity stays that of the expression as enumerated (canonval leaves bpe->cplx alone). */
RIES_HD void expand_finish(const RiesParams& P, const Node& par, int g, Node& child) {
  child.sym = par.sym;
  sym_set(child.sym, par.len, (unsigned)(g + 1));
  child.form = par.form;
  child.cplx = (int16_t)(par.cplx + P.sym_wgt[g]);
  int len = par.len + 1;
  const int code = child.pad0;
  if (code) {
    if (code & CANON_NEG) sym_set(child.sym, len++, (unsigned)(P.canon_sym[0] + 1));
    if (code & CANON_RECIP) sym_set(child.sym, len++, (unsigned)(P.canon_sym[1] + 1));
    if (code & CANON_DIV2) {
      sym_set(child.sym, len++, (unsigned)(P.canon_sym[2] + 1));
      sym_set(child.sym, len++, (unsigned)(P.canon_sym[3] + 1));
    }
    if (code & CANON_MUL2) {
      sym_set(child.sym, len++, (unsigned)(P.canon_sym[2] + 1));
      sym_set(child.sym, len++, (unsigned)(P.canon_sym[4] + 1));
    }
  }
  child.len = (uint8_t)len;
  child.pad0 = 0;
  child.pad1 = 0;
}

/* One (prefix, slot) pair end to end (used by the CPU emulation; the kernel runs the three parts in
   separate phases so that threads can be regrouped by symbol in between). */
RIES_HD int expand_one(const RiesParams& P, const FormDesc& F, const Node* nodes, const Node& par,
                       uint32_t par_idx, int slot, int atts, int using_x, int e_minw, int e_maxw,
                       Node& child) {
  int ip = par.len;
  int cls = F.cls[ip];
  if (slot >= P.cls_n[cls]) return OUT_NONE;
  if (!expand_tests(P, F, par, slot, atts, using_x, e_minw, e_maxw)) return OUT_DEAD;
  int g = P.trial[cls][slot];
  int oc = expand_exec(P, cls, g, ip + 1 >= F.flen, nodes, par, par_idx, using_x, child);
  if (oc == OUT_LIVE || oc == OUT_DONE) expand_finish(P, par, g, child);
  return oc;
}
