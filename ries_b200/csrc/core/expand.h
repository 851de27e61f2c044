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

/* `nodes` is the pass's node arena (parents of every level stay resident for `below` links);
   `par_idx` is the parent's own index in it. */
RIES_HD int expand_one(const RiesParams& P, const FormDesc& F, const Node* nodes, const Node& par,
                       uint32_t par_idx, int slot, int atts, int using_x, int e_minw, int e_maxw,
                       Node& child) {
  int ip = par.len;
  int cls = F.cls[ip];
  if (slot >= P.cls_n[cls]) return OUT_NONE;
  int g = P.trial[cls][slot];
  unsigned char sym = P.sym_char[g];

  /* symbol-append loop tests, in the reference's order (9401-9445) */
  int t_cpx = par.cplx + P.sym_wgt[g];
  if (t_cpx + F.rminw[ip + 1] > e_maxw) return OUT_DEAD;
  if (t_cpx + F.rmaxw[ip + 1] < e_minw) return OUT_DEAD;
  if (atts & P.sym_mask[g]) return OUT_DEAD;
  if (using_x && (ip == 0) && (sym != 'x')) return OUT_DEAD;
  if (P.sym_alwd[g] < RIES_MAX_ELEN) {
    int ct = 0;
    for (int i = 0; i < ip; i++) ct += (sym_get(par.sym, i) == (unsigned)(g + 1));
    if (ct >= P.sym_alwd[g]) return OUT_DEAD;
  }

  /* exec the new symbol on the cactus stack (child's ge_2 entry, 8942-8954) */
  Val r;
  int err;
  uint32_t below;
  int sp;
  if (cls == 0) {
    err = exec_push(P, sym, P.target, r);
    below = par_idx;
    sp = par.sp + 1;
  } else if (cls == 1) {
    Val a = {par.val, par.der, par.tag};
    err = exec_unary(P, sym, using_x != 0, a, r);
    below = par.below;
    sp = par.sp;
  } else {
    const Node& lo = nodes[par.below];
    Val a = {lo.val, lo.der, lo.tag};
    Val b = {par.val, par.der, par.tag};
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

  child.val = r.v;
  child.der = r.d;
  child.sym = par.sym;
  sym_set(child.sym, ip, (unsigned)(g + 1));
  child.below = below;
  child.form = par.form;
  child.cplx = (int16_t)t_cpx;
  child.tag = (uint8_t)r.tag;
  child.len = (uint8_t)(ip + 1);
  child.sp = (uint8_t)sp;
  child.pad0 = 0;
  child.pad1 = 0;

  if (ip + 1 >= F.flen) {
    /* complete expression (9085-9139) */
    if (using_x && (fabs(r.d) < P.k_vanished_dx)) return OUT_DEAD;
    if (r.v < P.min_equ_val) return OUT_DEAD;
    if (r.v > P.max_equ_val) return OUT_DEAD;
    return OUT_DONE;
  }
  return OUT_LIVE;
}
