// rmath implementation: glibc 2.39 (x86-64, FMA ifunc variants) restated for host and device.
//
// The reference links libm dynamically (exp log pow sincos tan atan2: `nm -D` of gcc -O2 ries.c --
// gcc folds every sin()/cos() pair of exec() into one sincos() call), glibc is not part of
// /root/reference, and the search dedupes on `==` of doubles; so the device must reproduce glibc's
// results bit for bit.  Each function below follows the published glibc algorithm named in its
// header comment.  Where the C source leaves the compiler free to contract a*b+c, the contraction
// glibc's own build chose for its *_fma variants is written out as an explicit fma(), so that
// nvcc -fmad=false and gcc -ffp-contract=off both produce exactly those operations.
// Pinned by tests/test_rmath.py: billions-scale differential runs against libm.so.6 are expected to
// report 0 mismatching bits on the argument ranges exec() can produce (DESIGN.md lists them).
//
// Domain notes (enforced by the callers in exec.h, ries.c:5308-5950):
//   exp: 0.01 <= x <= 690        log: x > 0            pow: x > 0, y finite
//   sincos, tan: |x| <= pi       atan2: both finite, not both zero
// Outside those ranges the functions still follow glibc except for the huge-argument trig
// reduction (__branred, |x| > 1.05e8 for sincos, > 25 for tan here), which returns NaN.
#pragma once

#if defined(__CUDA_ARCH__)
#define RM_ASU(x) ((uint64_t)__double_as_longlong(x))
#define RM_ASD(x) (__longlong_as_double((long long)(x)))
#else
#include <string.h>
RIES_HD uint64_t rm_asu_(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
RIES_HD double rm_asd_(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }
#define RM_ASU(x) rm_asu_(x)
#define RM_ASD(x) rm_asd_(x)
#endif

/* ---- tables (bit patterns; see tools/extract_glibc_tables.py) ---- */
static const uint64_t rm_tab_exp_h[256] = {
#include "rmath_tab_exp.inc"
};
static const uint64_t rm_tab_log_h[256] = {
#include "rmath_tab_log.inc"
};
static const uint64_t rm_tab_powlog_h[512] = {
#include "rmath_tab_powlog.inc"
};
static const uint64_t rm_tab_sincos_h[440] = {
#include "rmath_tab_sincos.inc"
};
static const uint64_t rm_tab_tan_h[744] = {
#include "rmath_tab_tan.inc"
};
static const uint64_t rm_tab_atan_h[1687] = {
#include "rmath_tab_atan.inc"
};
#if defined(__CUDACC__)
static __device__ const uint64_t rm_tab_exp_d[256] = {
#include "rmath_tab_exp.inc"
};
static __device__ const uint64_t rm_tab_log_d[256] = {
#include "rmath_tab_log.inc"
};
static __device__ const uint64_t rm_tab_powlog_d[512] = {
#include "rmath_tab_powlog.inc"
};
static __device__ const uint64_t rm_tab_sincos_d[440] = {
#include "rmath_tab_sincos.inc"
};
static __device__ const uint64_t rm_tab_tan_d[744] = {
#include "rmath_tab_tan.inc"
};
static __device__ const uint64_t rm_tab_atan_d[1687] = {
#include "rmath_tab_atan.inc"
};
#endif
#if defined(__CUDA_ARCH__)
#define RM_TAB(name) rm_tab_##name##_d
#else
#define RM_TAB(name) rm_tab_##name##_h
#endif

#define RM_NAN (RM_ASD(0x7ff8000000000000ull))

/* ------------------------------------------------------------------------------------------
   exp -- sysdeps/ieee754/dbl-64/e_exp.c (Szabolcs Nagy / ARM optimized routines), N = 128,
   degree-5 polynomial; constants from e_exp_data.c.
   ------------------------------------------------------------------------------------------ */
#define RM_EXP_INVLN2N 0x1.71547652b82fep+7
#define RM_EXP_SHIFT 0x1.8p52
#define RM_EXP_NEGLN2HIN (-0x1.62e42fefa0000p-8)
#define RM_EXP_NEGLN2LON (-0x1.cf79abc9e3b3ap-47)
#define RM_EXP_C2 0x1.ffffffffffdbdp-2
#define RM_EXP_C3 0x1.555555555543cp-3
#define RM_EXP_C4 0x1.55555cf172b91p-5
#define RM_EXP_C5 0x1.1111167a4d017p-7

/* specialcase() of e_exp.c: the scale 2^k over- or underflows */
RIES_HD double rm_exp_special(double tmp, uint64_t sbits, uint64_t ki) {
  double scale, y;
  if ((ki & 0x80000000ull) == 0) {
    sbits -= 1009ull << 52;
    scale = RM_ASD(sbits);
    y = 0x1p1009 * fma(scale, tmp, scale);
    return y;
  }
  sbits += 1022ull << 52;
  scale = RM_ASD(sbits);
  y = scale + scale * tmp;
  if (fabs(y) < 1.0) {
    double hi, lo, one = 1.0;
    if (y < 0.0) one = -1.0;
    lo = scale - y + scale * tmp;
    hi = one + y;
    lo = one - hi + y + lo;
    y = (hi + lo) - one;
    if (y == 0.0) y = RM_ASD(sbits & 0x8000000000000000ull);
  }
  y = 0x1p-1022 * y;
  return y;
}

/* exp_inline(x, xtail, sign_bias) shared by exp (xtail = 0, sign_bias = 0) and pow */
RIES_HD double rm_exp_core(double x, double xtail, int use_tail, uint64_t sign_bias) {
  uint32_t abstop = (uint32_t)(RM_ASU(x) >> 52) & 0x7ff;
  if (abstop - 0x3c9u >= 0x3fu) { /* |x| < 2^-54 or |x| >= 512 */
    if (abstop - 0x3c9u >= 0x80000000u) {
      /* tiny: exp(x) = 1 + x */
      double one = use_tail ? 1.0 + x : 1.0 + x;
      return sign_bias ? -one : one;
    }
    if (abstop >= 0x409u) { /* |x| >= 1024 */
      if (!use_tail && RM_ASU(x) == 0xfff0000000000000ull) return 0.0;
      if (abstop >= 0x7ffu) return 1.0 + x;
      if (RM_ASU(x) >> 63) return sign_bias ? -0.0 : 0.0; /* underflow */
      return sign_bias ? -HUGE_VAL : HUGE_VAL;             /* overflow */
    }
    abstop = 0; /* large but finite result possible: handled in specialcase */
  }
  double kd = fma(x, RM_EXP_INVLN2N, RM_EXP_SHIFT);
  uint64_t ki = RM_ASU(kd);
  kd -= RM_EXP_SHIFT;
  double r = fma(kd, RM_EXP_NEGLN2HIN, x);
  r = fma(kd, RM_EXP_NEGLN2LON, r);
  if (use_tail) r = xtail + r;
  uint64_t idx = 2 * (ki & 127);
  uint64_t top = (ki + sign_bias) << 45;
  double tail = RM_ASD(RM_TAB(exp)[idx]);
  uint64_t sbits = RM_TAB(exp)[idx + 1] + top;
  double r2 = r * r;
  double p23 = fma(r, RM_EXP_C3, RM_EXP_C2);
  double p45 = fma(r, RM_EXP_C5, RM_EXP_C4);
  double tmp = fma(p23, r2, r + tail);
  tmp = fma(p45, r2 * r2, tmp);
  if (abstop == 0) return rm_exp_special(tmp, sbits, ki);
  double scale = RM_ASD(sbits);
  return fma(tmp, scale, scale);
}

RIES_HD double rm_exp(double x) { return rm_exp_core(x, 0.0, 0, 0); }

/* ------------------------------------------------------------------------------------------
   log -- sysdeps/ieee754/dbl-64/e_log.c, 128-entry table, degree-5 (A) / degree-11 near 1 (B).
   ------------------------------------------------------------------------------------------ */
#define RM_LN2HI 0x1.62e42fefa3800p-1
#define RM_LN2LO 0x1.ef35793c76730p-45

RIES_HD double rm_log(double x) {
  const double A0 = -0x1.0000000000001p-1, A1 = 0x1.555555551305bp-2, A2 = -0x1.fffffffeb4590p-3,
               A3 = 0x1.999b324f10111p-3, A4 = -0x1.55575e506c89fp-3;
  const double B0 = -0x1p-1, B1 = 0x1.5555555555577p-2, B2 = -0x1.ffffffffffdcbp-3,
               B3 = 0x1.999999995dd0cp-3, B4 = -0x1.55555556745a7p-3, B5 = 0x1.24924a344de30p-3,
               B6 = -0x1.fffffa4423d65p-4, B7 = 0x1.c7184282ad6cap-4, B8 = -0x1.999eb43b068ffp-4,
               B9 = 0x1.78182f7afd085p-4, B10 = -0x1.5521375d145cdp-4;
  uint64_t ix = RM_ASU(x);
  uint32_t top = (uint32_t)(ix >> 48);
  /* 1 - 2^-4 <= x < 1 + 0x1.09p-4 */
  if (ix - 0x3fee000000000000ull < 0x3090000000000ull) {
    if (ix == 0x3ff0000000000000ull) return 0.0;
    double r = x - 1.0;
    double r2 = r * r;
    double r3 = r * r2;
    double p1 = fma(r2, B3, fma(r, B2, B1));
    double p2 = fma(r2, B6, fma(r, B5, B4));
    double p3 = fma(r3, B10, fma(r2, B9, fma(r, B8, B7)));
    double y = fma(fma(p3, r3, p2), r3, p1);
    /* split r: rhi = (r + r*2^27) - r*2^27 */
    double w = fma(r, 0x1p27, r);
    double rhi = fma(-0x1p27, r, w);
    double rlo = r - rhi;
    double rhi2 = rhi * rhi;
    double hi = fma(rhi2, B0, r);
    double lo = fma(rhi2, B0, r - hi);
    lo = fma(B0 * rlo, r + rhi, lo);
    y = fma(y, r3, lo);
    return hi + y;
  }
  if (top - 0x0010u >= 0x7ff0u - 0x0010u) {
    /* x < 0x1p-1022 or inf or nan */
    if (ix * 2 == 0) return -HUGE_VAL;
    if (ix == 0x7ff0000000000000ull) return x;
    if ((top & 0x8000u) || (top & 0x7ff0u) == 0x7ff0u) return RM_NAN;
    /* subnormal: normalize */
    ix = RM_ASU(x * 0x1p52);
    ix -= 52ull << 52;
  }
  uint64_t tmp = ix - 0x3fe6000000000000ull;
  int i = (int)((tmp >> 45) & 127);
  int k = (int)((int64_t)tmp >> 52);
  uint64_t iz = ix - (tmp & (0xfffull << 52));
  double invc = RM_ASD(RM_TAB(log)[2 * i]);
  double logc = RM_ASD(RM_TAB(log)[2 * i + 1]);
  double z = RM_ASD(iz);
  double kd = (double)k;
  double r = fma(z, invc, -1.0);
  double w = fma(kd, RM_LN2HI, logc);
  double hi = w + r;
  double lo = fma(kd, RM_LN2LO, (w - hi) + r);
  double r2 = r * r;
  double r3 = r * r2;
  double p = fma(fma(r, A4, A3), r2, fma(r, A2, A1));
  double y = fma(r3, p, fma(r2, A0, lo));
  return y + hi;
}

/* ------------------------------------------------------------------------------------------
   pow -- sysdeps/ieee754/dbl-64/e_pow.c: log_inline (128-entry table with logctail, degree-7
   polynomial) -> exp_inline above.
   ------------------------------------------------------------------------------------------ */
RIES_HD int rm_checkint(uint64_t iy) { /* 0 not int, 1 odd, 2 even */
  int e = (int)(iy >> 52) & 0x7ff;
  if (e < 0x3ff) return 0;
  if (e > 0x3ff + 52) return 2;
  if (iy & ((1ull << (0x3ff + 52 - e)) - 1)) return 0;
  if (iy & (1ull << (0x3ff + 52 - e))) return 1;
  return 2;
}

RIES_HD double rm_pow(double x, double y) {
  const double A0 = -0x1p-1, A1 = -0x1.5555555555560p-1, A2 = 0x1.0000000000006p-1,
               A3 = 0x1.999999959554ep-1, A4 = -0x1.555555529a47ap-1, A5 = -0x1.2495b9b4845e9p+0,
               A6 = 0x1.0002b8b263fc3p+0;
  uint64_t sign_bias = 0;
  uint64_t ix = RM_ASU(x), iy = RM_ASU(y);
  uint32_t topx = (uint32_t)(ix >> 52), topy = (uint32_t)(iy >> 52);
  if (topx - 0x001u >= 0x7ffu - 0x001u || (topy & 0x7ff) - 0x3beu >= 0x43eu - 0x3beu) {
    /* special cases: (x < 0x1p-1022 or inf or nan) or (|y| < 0x1p-65 or |y| >= 0x1p63 or nan) */
    if (2 * iy - 1 >= 2 * 0x7ff0000000000000ull - 1) { /* y zero, inf or nan */
      if (2 * iy == 0) return 1.0;
      if (ix == 0x3ff0000000000000ull) return 1.0;
      if (2 * ix > 2 * 0x7ff0000000000000ull || 2 * iy > 2 * 0x7ff0000000000000ull) return x + y;
      if (2 * ix == 2 * 0x3ff0000000000000ull) return 1.0;
      if ((2 * ix < 2 * 0x3ff0000000000000ull) == !(iy >> 63)) return 0.0;
      return y * y;
    }
    if (2 * ix - 1 >= 2 * 0x7ff0000000000000ull - 1) { /* x zero, inf or nan */
      double x2 = x * x;
      if ((ix >> 63) && rm_checkint(iy) == 1) x2 = -x2;
      return (iy >> 63) ? (1 / x2) : x2;
    }
    if (ix >> 63) { /* finite x < 0 */
      int yint = rm_checkint(iy);
      if (yint == 0) return RM_NAN;
      if (yint == 1) sign_bias = 0x800ull << 7; /* SIGN_BIAS = 0x800 << EXP_TABLE_BITS */
      ix &= 0x7fffffffffffffffull;
      topx &= 0x7ff;
    }
    if ((topy & 0x7ff) - 0x3beu >= 0x43eu - 0x3beu) {
      if (ix == 0x3ff0000000000000ull) return 1.0;
      if ((topy & 0x7ff) < 0x3beu) return (ix > 0x3ff0000000000000ull) ? 1.0 + y : 1.0 - y;
      return ((ix > 0x3ff0000000000000ull) == (topy < 0x800u)) ? HUGE_VAL : 0.0;
    }
    if (topx == 0) { /* subnormal x */
      ix = RM_ASU(RM_ASD(ix) * 0x1p52);
      ix &= 0x7fffffffffffffffull;
      ix -= 52ull << 52;
    }
  }
  /* log_inline */
  uint64_t tmp = ix - 0x3fe6955500000000ull;
  int i = (int)((tmp >> 45) & 127);
  int k = (int)((int64_t)tmp >> 52);
  uint64_t iz = ix - (tmp & (0xfffull << 52));
  double z = RM_ASD(iz);
  double kd = (double)k;
  double invc = RM_ASD(RM_TAB(powlog)[4 * i]);
  double logc = RM_ASD(RM_TAB(powlog)[4 * i + 2]);
  double logctail = RM_ASD(RM_TAB(powlog)[4 * i + 3]);
  double r = fma(z, invc, -1.0);
  double t1 = fma(kd, RM_LN2HI, logc);
  double t2 = t1 + r;
  double lo1 = fma(kd, RM_LN2LO, logctail);
  double lo2 = t1 - t2 + r;
  double ar = A0 * r;
  double ar2 = r * ar;
  double ar3 = r * ar2;
  double hi = t2 + ar2;
  double lo3 = fma(ar, r, -ar2);
  double lo4 = t2 - hi + ar2;
  double p = fma(ar2, fma(ar2, fma(r, A6, A5), fma(r, A4, A3)), fma(r, A2, A1));
  double lo = fma(ar3, p, lo1 + lo2 + lo3 + lo4);
  double lhi = hi + lo;
  double llo = hi - lhi + lo;
  /* exp_inline(y * log(x)) */
  double ehi = y * lhi;
  double elo = fma(y, llo, fma(y, lhi, -ehi));
  return rm_exp_core(ehi, elo, 1, sign_bias);
}

/* ------------------------------------------------------------------------------------------
   sincos -- sysdeps/ieee754/dbl-64/s_sincos.c over the helpers of s_sin.c (IBM Accurate
   Mathematical Library): Taylor for |x| < 0.126, else the 440-entry sin/cos table at spacing
   2^-7 with degree-5 corrections; Cody-Waite reduction by pi/2 in three pieces.
   ------------------------------------------------------------------------------------------ */
#define RM_BIG 0x1.8p45
#define RM_HP0 0x1.921fb54442d18p+0
#define RM_HP1 0x1.1a62633145c07p-54
#define RM_HPINV 0x1.45f306dc9c883p-1
#define RM_TOINT 0x1.8p52
#define RM_MP1 0x1.921fb58000000p+0
#define RM_MP2 (-0x1.dde973c000000p-27)
#define RM_PP3 (-0x1.cb3b398000000p-55)
#define RM_PP4 (-0x1.d747f23e32ed7p-83)

RIES_HD double rm_taylor_sin(double xx, double x, double dx) {
  const double s1 = -0x1.5555555555555p-3, s2 = 0x1.1111111110ecep-7, s3 = -0x1.a01a019db08b8p-13,
               s4 = 0x1.71de27b9a7ed9p-19, s5 = -0x1.addffc2fcdf59p-26;
  double poly = fma(fma(fma(fma(s5, xx, s4), xx, s3), xx, s2), xx, s1);
  double t = fma(fma(poly, x, -(0.5 * dx)), xx, dx);
  return x + t;
}

#define RM_SN3 (-0x1.5555555555515p-3)
#define RM_SN5 0x1.11110e829872fp-7
#define RM_CS2 0x1p-1
#define RM_CS4 (-0x1.5555555555535p-5)
#define RM_CS6 0x1.6c16bedd9e239p-10

RIES_HD double rm_do_sin(double x, double dx) {
  double xold = x;
  if (fabs(x) < 0x1.020c49ba5e354p-3) return rm_taylor_sin(x * x, x, dx); /* 0.126 */
  if (x <= 0) dx = -dx;
  double u = RM_BIG + fabs(x);
  x = fabs(x) - (u - RM_BIG);
  int k = (int)(uint32_t)RM_ASU(u) * 4;
  double xx = x * x;
  double s = x + fma(x * xx, fma(xx, RM_SN5, RM_SN3), dx);
  double c = fma(x, dx, xx * fma(fma(RM_CS6, xx, RM_CS4), xx, RM_CS2));
  double sn = RM_ASD(RM_TAB(sincos)[k]), ssn = RM_ASD(RM_TAB(sincos)[k + 1]);
  double cs = RM_ASD(RM_TAB(sincos)[k + 2]), ccs = RM_ASD(RM_TAB(sincos)[k + 3]);
  double cor = fma(cs, s, fma(-sn, c, fma(s, ccs, ssn)));
  return copysign(sn + cor, xold);
}

RIES_HD double rm_do_cos(double x, double dx) {
  if (x < 0) dx = -dx;
  double u = RM_BIG + fabs(x);
  x = fabs(x) - (u - RM_BIG) + dx;
  int k = (int)(uint32_t)RM_ASU(u) * 4;
  double xx = x * x;
  double s = fma(x * xx, fma(xx, RM_SN5, RM_SN3), x);
  double c = xx * fma(fma(RM_CS6, xx, RM_CS4), xx, RM_CS2);
  double sn = RM_ASD(RM_TAB(sincos)[k]), ssn = RM_ASD(RM_TAB(sincos)[k + 1]);
  double cs = RM_ASD(RM_TAB(sincos)[k + 2]), ccs = RM_ASD(RM_TAB(sincos)[k + 3]);
  double cor = fma(-sn, s, fma(-cs, c, fma(-s, ssn, ccs)));
  return cs + cor;
}

RIES_HD int rm_reduce_sincos(double x, double* a, double* da) {
  double t = fma(x, RM_HPINV, RM_TOINT);
  double xn = t - RM_TOINT;
  int n = (int)(uint32_t)RM_ASU(t) & 3;
  double y = fma(-xn, RM_MP2, fma(-xn, RM_MP1, x));
  double t2 = fma(-xn, RM_PP3, y);
  double db = fma(-xn, RM_PP3, y - t2);
  double b = fma(-xn, RM_PP4, t2);
  db += fma(-xn, RM_PP4, t2 - b);
  *a = b;
  *da = db;
  return n;
}

RIES_HD double rm_do_sincos(double a, double da, int n) {
  double r = (n & 1) ? rm_do_cos(a, da) : rm_do_sin(a, da);
  return (n & 2) ? -r : r;
}

RIES_HD void rm_sincos(double x, double* sinx, double* cosx) {
  uint32_t k = (uint32_t)(RM_ASU(x) >> 32) & 0x7fffffffu;
  if (k < 0x400368fdu) {
    if (k < 0x3e400000u) { /* |x| < 2^-27 */
      *sinx = x; *cosx = 1.0;
      return;
    }
    if (k < 0x3feb6000u) { /* |x| < 0.855469 */
      *sinx = (k < 0x3e500000u) ? x : rm_do_sin(x, 0);
      *cosx = rm_do_cos(x, 0);
      return;
    }
    /* |x| < 2.426265 */
    double y = RM_HP0 - fabs(x);
    double a = y + RM_HP1;
    double da = (y - a) + RM_HP1;
    *sinx = copysign(rm_do_cos(a, da), x);
    *cosx = rm_do_sin(a, da);
    return;
  }
  if (k < 0x419921FBu) { /* |x| < 105414350 */
    double a, da;
    int n = rm_reduce_sincos(x, &a, &da);
    *sinx = rm_do_sincos(a, da, n);
    *cosx = rm_do_sincos(a, da, n + 1);
    return;
  }
  *sinx = RM_NAN; *cosx = RM_NAN; /* __branred range / inf / nan: never reached from exec() */
}

/* ------------------------------------------------------------------------------------------
   tan -- sysdeps/ieee754/dbl-64/s_tan.c (IBM), |x| <= 1e8: odd polynomial to 0.0608, the
   186-entry {x, tan, cot} table to 0.787, reduction by pi/2 with tan / -cot above (one extra term of
   pi/2 to 25, two extra terms beyond: --max-trig-cycles lets exec() reach 100 pi).
   ------------------------------------------------------------------------------------------ */
RIES_HD double rm_tan_poly(double a2) { /* d3 + a2 (d5 + a2 (d7 + a2 (d9 + a2 d11))) */
  const double d3 = 0x1.5555555555555p-2, d5 = 0x1.11111111107c6p-3, d7 = 0x1.ba1ba1cdb8745p-5,
               d9 = 0x1.664ed49cfc666p-6, d11 = 0x1.2385a3cf2e4eap-7;
  return fma(fma(fma(fma(d11, a2, d9), a2, d7), a2, d5), a2, d3);
}

RIES_HD double rm_tan(double x) {
  const double g1 = 0x1.b096c00000000p-27, g2 = 0x1.f212d00000000p-5, g3 = 0x1.92f1a00000000p-1,
               g4 = 25.0, g5 = 1.0e8;
  const double e0 = 0x1.5555555554dbdp-2, e1 = 0x1.11112e0a6b45fp-3;
  const double mp3 = -0x1.cb3b399d747f2p-55;
  uint32_t ux = (uint32_t)(RM_ASU(x) >> 32);
  if ((ux & 0x7ff00000u) == 0x7ff00000u) return x - x;
  double w = (x < 0.0) ? -x : x;
  if (w <= g1) return x;
  if (w <= g2) {
    double x2 = x * x;
    return fma(x * x2, rm_tan_poly(x2), x);
  }
  if (w <= g3) {
    int i = (int)fma(256.0, w, -15.5);
    double z = w - RM_ASD(RM_TAB(tan)[4 * i]);
    double z2 = z * z;
    double s = (x < 0.0) ? -1.0 : 1.0;
    double pz = fma(z * z2, fma(z2, e1, e0), z);
    double fi = RM_ASD(RM_TAB(tan)[4 * i + 1]);
    double gi = RM_ASD(RM_TAB(tan)[4 * i + 2]);
    double t2 = pz * (gi + fi) / (gi - pz);
    return s * (fi + t2);
  }
  if (w <= g5) {
    double t = fma(x, RM_HPINV, RM_TOINT);
    double xn = t - RM_TOINT;
    int n = (int)(uint32_t)RM_ASU(t) & 1;
    double t1 = fma(-xn, RM_MP2, fma(-xn, RM_MP1, x));
    double a, da;
    if (w <= g4) { /* |x| <= 25: one more term of pi/2 */
      a = fma(-xn, mp3, t1);
      da = fma(-xn, mp3, t1 - a);
    } else { /* 25 < |x| <= 1e8: two more terms (pp3, pp4), then EADD(a, da) */
      double tt = fma(-xn, RM_PP3, t1);
      da = fma(-xn, RM_PP3, t1 - tt);
      a = fma(-xn, RM_PP4, tt);
      da = da + fma(-xn, RM_PP4, tt - a);
      double z = a + da;
      double zz = (fabs(a) > fabs(da)) ? ((a - z) + da) : ((da - z) + a);
      a = z;
      da = zz;
    }
    double ya, yya, sy;
    if (a < 0.0) { ya = -a; yya = -da; sy = -1.0; }
    else { ya = a; yya = da; sy = 1.0; }
    if (ya <= g2) {
      double a2 = a * a;
      double t2 = fma(a * a2, rm_tan_poly(a2), da);
      if (n) {
        /* -cot: EADD(a, t2, b, db); DIV2(1, 0, b, db, c, dc) */
        double b = a + t2;
        double db = (fabs(a) > fabs(t2)) ? ((a - b) + t2) : ((t2 - b) + a);
        double c = 1.0 / b;
        double u = c * b;
        double uu = fma(c, b, -u);
        double cc = fma(-db, c, ((1.0 - u) - uu) + 0.0) / b;
        double zz0 = c + cc;
        double zz1 = (c - zz0) + cc;
        return -(zz0 + zz1);
      }
      return a + t2;
    }
    int i = (int)fma(256.0, ya, -15.5);
    double z = (ya - RM_ASD(RM_TAB(tan)[4 * i])) + yya;
    double z2 = z * z;
    double pz = fma(z * z2, fma(z2, e1, e0), z);
    double fi = RM_ASD(RM_TAB(tan)[4 * i + 1]);
    double gi = RM_ASD(RM_TAB(tan)[4 * i + 2]);
    if (n) {
      double t2 = pz * (fi + gi) / (fi + pz);
      return (-sy) * (gi - t2);
    }
    double t2 = pz * (gi + fi) / (gi - pz);
    return sy * (fi + t2);
  }
  return RM_NAN; /* |x| > 1e8 (__branred range): never reached from exec(), |a| <= 100 pi */
}

/* ------------------------------------------------------------------------------------------
   atan2 -- sysdeps/ieee754/dbl-64/e_atan2.c (IBM): u = min/max with an exact correction du,
   degree-13 odd polynomial below 1/16, else the 241-entry cij table; four quadrant cases.
   ------------------------------------------------------------------------------------------ */
#define RM_HPI 0x1.921fb54442d18p+0
#define RM_HPI1 0x1.1a62633145c07p-54
#define RM_OPI 0x1.921fb54442d18p+1
#define RM_OPI1 0x1.1a62633145c07p-53

RIES_HD double rm_atan_poly(double v) { /* d3 + v (d5 + v (d7 + v (d9 + v (d11 + v d13)))) */
  const double d3 = -0x1.5555555555555p-2, d5 = 0x1.99999999997fdp-3, d7 = -0x1.24924923f7603p-3,
               d9 = 0x1.c71c6e5129a3bp-4, d11 = -0x1.7458022b13c25p-4, d13 = 0x1.375f08b31cbcep-4;
  return fma(fma(fma(fma(fma(d13, v, d11), v, d9), v, d7), v, d5), v, d3);
}

RIES_HD int rm_atan_index(double u) { return (int)(fma(u, 256.0, 0x1p52) - 0x1p52) - 16; }

RIES_HD double rm_atan_tabpoly(const uint64_t* c, double v) { /* c2 + v (c3 + v (c4 + v (c5 + v c6))) */
  return fma(fma(fma(fma(RM_ASD(c[6]), v, RM_ASD(c[5])), v, RM_ASD(c[4])), v, RM_ASD(c[3])), v, RM_ASD(c[2]));
}

RIES_HD double rm_atan2(double y, double x) {
  uint64_t bx = RM_ASU(x), by = RM_ASU(y);
  uint32_t ux = (uint32_t)(bx >> 32), dx = (uint32_t)bx, uy = (uint32_t)(by >> 32), dy = (uint32_t)by;
  if ((ux & 0x7ff00000u) == 0x7ff00000u && (((ux & 0x000fffffu) | dx) != 0)) return x + y;
  if ((uy & 0x7ff00000u) == 0x7ff00000u && (((uy & 0x000fffffu) | dy) != 0)) return y + y;
  if (uy == 0 && dy == 0) return (ux & 0x80000000u) ? RM_OPI : 0.0;
  if (uy == 0x80000000u && dy == 0) return (ux & 0x80000000u) ? -RM_OPI : -0.0;
  if (x == 0) return (uy & 0x80000000u) ? -RM_HPI : RM_HPI;
  if ((ux & 0x7fffffffu) == 0x7ff00000u && dx == 0) { /* x = +-inf */
    if ((uy & 0x7fffffffu) == 0x7ff00000u && dy == 0) {
      double q = (ux & 0x80000000u) ? 0x1.2d97c7f3321d2p+1 : 0x1.921fb54442d18p-1;
      return (uy & 0x80000000u) ? -q : q;
    }
    if (ux & 0x80000000u) return (uy & 0x80000000u) ? -RM_OPI : RM_OPI;
    return (uy & 0x80000000u) ? -0.0 : 0.0;
  }
  if ((uy & 0x7fffffffu) == 0x7ff00000u && dy == 0) return (uy & 0x80000000u) ? -RM_HPI : RM_HPI;

  double ax = (x < 0) ? -x : x;
  double ay = (y < 0) ? -y : y;
  int de = (int)(uy & 0x7ff00000u) - (int)(ux & 0x7ff00000u);
  if (de >= 59768832) return (y > 0) ? RM_HPI : -RM_HPI;
  if (de <= -59768832) {
    if (x > 0) return copysign(ay / ax, y);
    return (y > 0) ? RM_OPI : -RM_OPI;
  }
  if (ax < 0x1p-500 || ay < 0x1p-500) { ax *= 0x1p500; ay *= 0x1p500; }
  if (ax > 0x1p500 || ay > 0x1p500) { ax *= 0x1p-500; ay *= 0x1p-500; }

  double u, du;
  if (ay < ax) {
    u = ay / ax;
    double v = u * ax;
    double vv = fma(u, ax, -v);
    du = ((ay - v) - vv) / ax;
  } else {
    u = ax / ay;
    double v = u * ay;
    double vv = fma(u, ay, -v);
    du = ((ax - v) - vv) / ay;
  }
  double z;
  if (x > 0) {
    if (ay < ax) { /* (i) atan(ay/ax) */
      if (u < 0.0625) {
        double v = u * u;
        z = u + fma(u * v, rm_atan_poly(v), du);
        return copysign(z, y);
      }
      const uint64_t* c = &RM_TAB(atan)[7 * rm_atan_index(u)];
      double t3 = u - RM_ASD(c[0]);
      double v = t3 + du;
      double dv = (fabs(t3) > fabs(du)) ? ((t3 - v) + du) : ((du - v) + t3);
      double t2 = RM_ASD(c[2]);
      double q = fma(fma(fma(RM_ASD(c[6]), v, RM_ASD(c[5])), v, RM_ASD(c[4])), v, RM_ASD(c[3]));
      double zz = fma(v, t2, fma(dv, t2, (v * v) * q));
      z = zz + RM_ASD(c[1]);
      return copysign(z, y);
    }
    /* (ii) pi/2 - atan(ax/ay) */
    if (u < 0.0625) {
      double v = u * u;
      double zz = (u * v) * rm_atan_poly(v);
      double t2 = RM_HPI - u;
      double cor = (RM_HPI > fabs(u)) ? ((RM_HPI - t2) - u) : (RM_HPI - (u + t2));
      double t3 = ((cor + RM_HPI1) - du) - zz;
      return copysign(t3 + t2, y);
    }
    const uint64_t* c = &RM_TAB(atan)[7 * rm_atan_index(u)];
    double v = (u - RM_ASD(c[0])) + du;
    double zz = fma(-rm_atan_tabpoly(c, v), v, RM_HPI1);
    z = (RM_HPI - RM_ASD(c[1])) + zz;
    return copysign(z, y);
  }
  if (ax < ay) { /* (iii) pi/2 + atan(ax/ay) */
    if (u < 0.0625) {
      double v = u * u;
      double zz = (u * v) * rm_atan_poly(v);
      double t2 = u + RM_HPI;
      double cor = (RM_HPI > fabs(u)) ? ((RM_HPI - t2) + u) : ((u - t2) + RM_HPI);
      double t3 = ((cor + RM_HPI1) + du) + zz;
      return copysign(t3 + t2, y);
    }
    const uint64_t* c = &RM_TAB(atan)[7 * rm_atan_index(u)];
    double v = (u - RM_ASD(c[0])) + du;
    double zz = fma(rm_atan_tabpoly(c, v), v, RM_HPI1);
    z = (RM_HPI + RM_ASD(c[1])) + zz;
    return copysign(z, y);
  }
  /* (iv) pi - atan(ay/ax) */
  if (u < 0.0625) {
    double v = u * u;
    double zz = (u * v) * rm_atan_poly(v);
    double t2 = RM_OPI - u;
    double cor = (RM_OPI > fabs(u)) ? ((RM_OPI - t2) - u) : (RM_OPI - (u + t2));
    double t3 = ((cor + RM_OPI1) - du) - zz;
    return copysign(t3 + t2, y);
  }
  const uint64_t* c = &RM_TAB(atan)[7 * rm_atan_index(u)];
  double v = (u - RM_ASD(c[0])) + du;
  double zz = fma(-rm_atan_tabpoly(c, v), v, RM_OPI1);
  z = (RM_OPI - RM_ASD(c[1])) + zz;
  return copysign(z, y);
}
