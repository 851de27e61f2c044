// One postfix symbol applied to (value, d/dx, class) operands -- the B200 restatement of
// exec() (/root/reference/public/ries.c:5308-5950) without the metastack: operands arrive in
// registers and the result is returned in a register; the caller owns the stack.
//
// Operation ORDER inside every expression below follows the reference's source text exactly
// (gcc -O2 on x86-64 does not contract or reassociate), and the build uses nvcc -fmad=false /
// gcc -ffp-contract=off, so value and derivative are bit-identical to the reference for identical
// libm results.
#pragma once
#include "params.h"
#include "rmath.h"

RIES_HD int tgmin(int a, int b) { return (a < b) ? a : b; } /* TGMIN ries.c:2940 */
RIES_HD bool dfinite(double d) {                              /* (d < inf) && (d > -inf), ries.c:5394 */
  return (d < HUGE_VAL) && (d > -HUGE_VAL);
}

/* trig result class, ries.c:5533-5537 */
RIES_HD int trig_tag(const RiesParams& P, int tga) {
  if (P.trig_scale_default) return (tga >= TYPE_RAT) ? TYPE_ALG : TYPE_NONE;
  return TYPE_NONE;
}

/* Class-a symbols: push a constant or x (ries.c:5335-5362). */
RIES_HD int exec_push(const RiesParams& P, unsigned char op, double xval, Val& out) {
  out.d = 0.0;
  switch (op) {
    case '1': out.v = 1.0; out.tag = TYPE_INT; return 0;
    case '2': out.v = 2.0; out.tag = TYPE_INT; return 0;
    case '3': out.v = 3.0; out.tag = TYPE_INT; return 0;
    case '4': out.v = 4.0; out.tag = TYPE_INT; return 0;
    case '5': out.v = 5.0; out.tag = TYPE_INT; return 0;
    case '6': out.v = 6.0; out.tag = TYPE_INT; return 0;
    case '7': out.v = 7.0; out.tag = TYPE_INT; return 0;
    case '8': out.v = 8.0; out.tag = TYPE_INT; return 0;
    case '9': out.v = 9.0; out.tag = TYPE_INT; return 0;
    case 'f': out.v = P.k_phi; out.tag = P.tg_phi; return 0;
    case 'e': out.v = P.k_e; out.tag = P.tg_e; return 0;
    case 'p': out.v = P.k_pi; out.tag = P.tg_pi; return 0;
    case 'x': out.v = xval; out.d = 1.0; out.tag = P.targ_tags; return 0;
    default: return ERR_EXEC_ILLEGAL_SYMBOL;
  }
}

/* Class-b symbols: one operand (ries.c:5365-5652). */
RIES_HD int exec_unary(const RiesParams& P, unsigned char op, bool do_dx, const Val& A, Val& out) {
  double a = A.v, da = A.d;
  int tga = A.tag;
  double rv, drv = 0.0;
  int trv = TYPE_NONE;
  switch (op) {
    case 'I':
      rv = a; drv = da; trv = tga;
      break;
    case 'n': /* 5374 */
      if (do_dx) drv = -da;
      rv = -a;
      trv = tga;
      break;
    case 'r': /* 5385 */
      if (a == 0.0) return ERR_EXEC_DIV_ZERO;
      if (do_dx) drv = (-da) / (a * a);
      if (!dfinite(drv)) return ERR_EXEC_ILLEGAL_DERIV;
      rv = 1.0 / a;
      trv = tgmin(tga, TYPE_RAT);
      break;
    case 's': /* 5403 */
      if (do_dx) drv = 2.0 * a * da;
      rv = a * a;
      trv = tga;
      break;
    case 'q': /* 5413 */
      if (a < 0.0) return ERR_EXEC_ROOT_NEG;
      rv = sqrt(a);
      if (do_dx) drv = da / (2.0 * rv);
      if ((tga == TYPE_INT) && (rv < 1.0e6) && (rv == floor(rv))) trv = TYPE_INT;
      else trv = tgmin(TYPE_CONS, tga);
      break;
    case 'l': /* 5436 */
      if (P.restrict_exponents) {
        if ((da != 0) || (tga < P.restrict_exponents)) return ERR_EXEC_ILLEGAL_EXPONENT;
      }
      if (a <= 0.0) return ERR_EXEC_LOG_NEG;
      if (fabs(a - 1.0) <= P.k_sig_loss) return ERR_EXEC_SIG_LOSS;
      if (do_dx) drv = da / a;
      if (!dfinite(drv)) return ERR_EXEC_ILLEGAL_DERIV;
      rv = rm_log(a);
      trv = tgmin(tga, TYPE_TCEL);
      break;
    case 'E': /* 5464 */
      if (P.restrict_exponents) {
        if ((da != 0) || (tga < P.restrict_exponents)) return ERR_EXEC_ILLEGAL_EXPONENT;
      }
      if (a > P.k_eXlim) return ERR_EXEC_OVERFLOW;
      if (a < P.k_sig_loss) return ERR_EXEC_SIG_LOSS;
      rv = rm_exp(a);
      if (do_dx) drv = rv * da;
      trv = tgmin(tga, TYPE_TCEL);
      break;
    case 'S': { /* 5488 */
      if (P.restrict_trig_args) {
        if ((da != 0) || (tga < P.restrict_trig_args)) return ERR_EXEC_TRIG_ARGTYPE;
      }
      a *= P.sincos_scale;
      if ((a > P.trig_max_arg) || (-a > P.trig_max_arg)) return ERR_EXEC_TRIG_RANGE;
      double ca;
      rm_sincos(a, &rv, &ca); /* gcc -O2 folds the reference's SIN(a)/COS(a) into one sincos() */
      if (fabs(rv) > P.k_sin_clip) return ERR_EXEC_TRIG_LOW_DX;
      if (fabs(ca) < P.k_sig_loss) return ERR_EXEC_SIG_LOSS;
      if (fabs(rv) < (fabs(a) * P.k_sig_loss)) return ERR_EXEC_SIG_LOSS;
      if (do_dx) {
        drv = ca * P.sincos_scale * da;
        if (fabs(drv) < (fabs(da) * P.k_sig_loss)) return ERR_EXEC_SIG_LOSS;
      }
      trv = trig_tag(P, tga);
      break;
    }
    case 'C': { /* 5541 */
      if (P.restrict_trig_args) {
        if ((da != 0) || (tga < P.restrict_trig_args)) return ERR_EXEC_TRIG_ARGTYPE;
      }
      a *= P.sincos_scale;
      if ((a > P.trig_max_arg) || (-a > P.trig_max_arg)) return ERR_EXEC_TRIG_RANGE;
      double sa;
      rm_sincos(a, &sa, &rv);
      if (fabs(rv) < (fabs(a) * P.k_sig_loss)) return ERR_EXEC_SIG_LOSS;
      if (fabs(sa) < P.k_sig_loss) return ERR_EXEC_SIG_LOSS;
      if (fabs(rv) > P.k_sin_clip) return ERR_EXEC_TRIG_LOW_DX;
      if (do_dx) drv = 0.0 - (sa * P.sincos_scale * da);
      trv = trig_tag(P, tga);
      break;
    }
    case 'T': /* 5582 */
      if (P.restrict_trig_args) {
        if ((da != 0) || (tga < P.restrict_trig_args)) return ERR_EXEC_TRIG_ARGTYPE;
      }
      a *= P.sincos_scale;
      if ((a > P.trig_max_arg) || (-a > P.trig_max_arg)) return ERR_EXEC_TRIG_RANGE;
      rv = rm_tan(a);
      if (fabs(rv) > (1 / P.k_sig_loss)) return ERR_EXEC_SIG_LOSS;
      if ((a != 0) && (fabs(rv / a) < P.k_sig_loss)) return ERR_EXEC_SIG_LOSS;
      if (do_dx) drv = P.sincos_scale * (1.0 + rv * rv) * da;
      trv = trig_tag(P, tga);
      break;
    default:
      return ERR_EXEC_ILLEGAL_SYMBOL;
  }
  out.v = rv; out.d = drv; out.tag = trv;
  return 0;
}

/* Class-c symbols: A is the deeper operand (first argument), B the top of stack (ries.c:5655-5909). */
RIES_HD int exec_binary(const RiesParams& P, unsigned char op, bool do_dx, const Val& A, const Val& B,
                        Val& out) {
  double a = A.v, da = A.d, b = B.v, db = B.d;
  int tga = A.tag, tgb = B.tag;
  double rv, drv = 0.0, lb;
  int trv = TYPE_NONE;
  switch (op) {
    case '-': /* 5655: shares '+' below */
      b = -b; db = -db;
      /* fallthrough */
    case '+': /* 5662 */
      rv = a + b;
      if ((fabs(a) < (fabs(b) * P.k_sig_loss)) || (fabs(b) < (fabs(a) * P.k_sig_loss))) {
        if (((rv - b) != a) || ((rv - a) != b)) return ERR_EXEC_SIG_LOSS;
      }
      if ((fabs(rv) < (fabs(a) * P.k_sig_loss)) || (fabs(rv) < (fabs(b) * P.k_sig_loss))) {
        if (((rv - b) != a) || ((rv - a) != b)) return ERR_EXEC_SIG_LOSS;
      }
      if (do_dx) drv = da + db;
      trv = tgmin(tga, tgb);
      break;
    case '*': /* 5701 */
      if (do_dx) drv = (b * da) + (a * db);
      rv = a * b;
      trv = tgmin(tga, tgb);
      break;
    case '/': /* 5714 */
      if (b == 0.0) return ERR_EXEC_DIV_ZERO;
      if (do_dx) drv = ((b * da) - (a * db)) / (b * b);
      if (!dfinite(drv)) return ERR_EXEC_ILLEGAL_DERIV;
      rv = a / b;
      if ((tga >= TYPE_RAT) && (tgb >= TYPE_RAT)) trv = (rv == floor(rv)) ? TYPE_INT : TYPE_RAT;
      else trv = tgmin(tga, tgb);
      break;
    case '^': /* 5739 */
      if (P.restrict_exponents) {
        if ((db != 0) || (tgb < P.restrict_exponents)) return ERR_EXEC_ILLEGAL_EXPONENT;
      }
      if (a <= 0.0) return ERR_EXEC_POW_NEG_BASE;
      if (fabs(b) < P.k_sig_loss) return ERR_EXEC_SIG_LOSS;
      rv = rm_pow(a, b);
      if (do_dx) {
        drv = rv * ((b * da / a) + (rm_log(a) * db));
        if (!dfinite(drv)) return ERR_EXEC_ILLEGAL_DERIV;
        if ((da || db) && (drv == 0.0)) return ERR_EXEC_ZERO_DERIV;
      }
      if (tgb == TYPE_INT) {
        if (b >= 0) trv = tga;
        else trv = tgmin(tga, TYPE_RAT);
      } else if (tgb == TYPE_RAT) {
        lb = fabs(b);
        if ((lb == 0.5) || (lb == 0.25) || (lb == 0.125)) trv = tgmin(tga, TYPE_CONS);
        else trv = tgmin(tga, TYPE_ALG);
      } else if ((tgb == TYPE_CONS) || (tgb == TYPE_ALG)) {
        trv = tgmin(tga, TYPE_TCEL);
      } else {
        trv = tgmin(tga, tgb);
      }
      break;
    case 'v': { /* 5793 */
      if (P.restrict_exponents) {
        if ((db != 0) || (tgb < P.restrict_exponents)) return ERR_EXEC_ILLEGAL_EXPONENT;
      }
      if (b == 0.0) return ERR_EXEC_OVERFLOW;
      if (fabs(b) > (1.0 / P.k_sig_loss)) return ERR_EXEC_SIG_LOSS;
      int f1 = 0;
      if (b == 3.0) {
        if (a < 0.0) { a = -a; f1 = 1; }
      } else if (a < 0.0) {
        return ERR_EXEC_ROOT_NEG;
      }
      rv = rm_pow(a, 1.0 / b);
      if (do_dx) {
        drv = rv * ((da / (a * b)) - (rm_log(a) * db / (b * b)));
        if ((da || db) && (drv == 0.0)) return ERR_EXEC_ZERO_DERIV;
      }
      if (f1) rv = -rv;
      if (!dfinite(drv)) return ERR_EXEC_ILLEGAL_DERIV;
      if (tgb >= TYPE_RAT) {
        lb = fabs(b);
        if ((lb == 2.0) || (lb == 4.0) || (lb == 8.0)) trv = tgmin(tga, TYPE_CONS);
        else trv = tgmin(tga, TYPE_ALG);
      } else if ((tgb == TYPE_CONS) || (tgb == TYPE_ALG)) {
        trv = tgmin(tga, TYPE_TCEL);
      } else {
        trv = tgmin(tga, tgb);
      }
      break;
    }
    case 'L': /* 5851 */
      if (P.restrict_exponents) {
        if ((db != 0) || (tgb < P.restrict_exponents)) return ERR_EXEC_ILLEGAL_EXPONENT;
      }
      if ((b <= 0.0) || (b == 1.0)) return ERR_EXEC_LOG_BAD_BASE;
      if (fabs(1.0 - b) < P.k_sig_loss) return ERR_EXEC_SIG_LOSS;
      if (P.restrict_exponents) {
        if ((da != 0) || (tga < P.restrict_exponents)) return ERR_EXEC_ILLEGAL_EXPONENT;
      }
      if (a <= 0.0) return ERR_EXEC_LOG_NEG;
      if (fabs(a - 1.0) <= P.k_sig_loss) return ERR_EXEC_SIG_LOSS;
      lb = rm_log(b);
      rv = rm_log(a) / lb;
      if (do_dx) {
        drv = (da / (a * lb)) - (rv * db / (b * lb));
        if (!dfinite(drv)) return ERR_EXEC_ILLEGAL_DERIV;
        if ((da || db) && (drv == 0.0)) return ERR_EXEC_ZERO_DERIV;
      }
      trv = tgmin(tga, tgmin(tgb, TYPE_TCEL));
      break;
    case 'A': /* 5895; arctan2() ries.c:5188 maps (0,0) to 0 */
      rv = ((a == 0) && (b == 0)) ? 0.0 : rm_atan2(a, b);
      if (do_dx) drv = (b * da - a * db) / (b * b + a * a);
      trv = TYPE_NONE;
      break;
    default:
      return ERR_EXEC_ILLEGAL_SYMBOL;
  }
  out.v = rv; out.d = drv; out.tag = trv;
  return 0;
}

/* Whole-expression evaluation on a local stack: eval() ries.c:6777-6862 (used by Newton only).
   `do_dx` = expression contains x (6787). Returns 0 or an ERR_ code. */
RIES_HD int eval_expr(const RiesParams& P, const Sym128& s, double xval, Val& result) {
  Val stk[RIES_STK_MAX + 1];
  int sp = 0;
  bool has_x = false;
  for (int i = 0; i < RIES_MAX_ELEN + 1; i++) {
    unsigned c = sym_get(s, i);
    if (!c) break;
    if (P.sym_char[c - 1] == 'x') has_x = true;
  }
  result.v = 0; result.d = 0; result.tag = 0;
  for (int i = 0;; i++) {
    unsigned c = (i <= RIES_MAX_ELEN) ? sym_get(s, i) : 0;
    if (!c) break;
    if (i >= RIES_MAX_ELEN) return ERR_EVAL_TOO_LONG;
    unsigned char op = P.sym_char[c - 1];
    int cls = P.sym_cls[c - 1];
    int err;
    Val r;
    if (cls == 0) {
      if (sp + 1 > RIES_STK_MAX) return ERR_EVAL_STACK_OVERFLOW;
      err = exec_push(P, op, xval, r);
      if (err) return err;
      stk[sp++] = r;
    } else if (cls == 1) {
      if (sp < 1) return ERR_EVAL_STACK_UNDERFLOW;
      err = exec_unary(P, op, has_x, stk[sp - 1], r);
      if (err) return err;
      stk[sp - 1] = r;
    } else {
      if (sp < 2) return ERR_EVAL_STACK_UNDERFLOW;
      err = exec_binary(P, op, has_x, stk[sp - 2], stk[sp - 1], r);
      if (err) return err;
      sp--;
      stk[sp - 1] = r;
    }
  }
  if (sp > 0) result = stk[sp - 1];
  return 0;
}
