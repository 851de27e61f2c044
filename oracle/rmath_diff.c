/* TEST INFRASTRUCTURE.  Differential test of rmath (host twin) against the platform libm -- the
   third-party dependency whose bits the reference's dedupe depends on.
   usage: rmath_diff <func> <n> [seed]   prints mismatch count and the first few offenders. */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

double rmh_exp(double), rmh_log(double), rmh_pow(double, double), rmh_tan(double), rmh_atan2(double, double);
void rmh_sincos(double, double *, double *);

static uint64_t s[2];
static uint64_t rnd(void) { /* xorshift128+ */
  uint64_t a = s[0], b = s[1];
  s[0] = b; a ^= a << 23; s[1] = a ^ b ^ (a >> 17) ^ (b >> 26);
  return s[1] + b;
}
static double unif(void) { return (double)(rnd() >> 11) * 0x1p-53; }
/* log-uniform magnitude in [2^lo, 2^hi) with random sign if sgn */
static double logu(double lo, double hi, int sgn) {
  double e = lo + (hi - lo) * unif();
  double v = exp2(e);
  return (sgn && (rnd() & 1)) ? -v : v;
}
static int same(double a, double b) {
  uint64_t x, y; memcpy(&x, &a, 8); memcpy(&y, &b, 8);
  return x == y || (a != a && b != b);
}

int main(int argc, char **argv) {
  const char *f = argc > 1 ? argv[1] : "exp";
  long n = argc > 2 ? atol(argv[2]) : 1000000, bad = 0;
  s[0] = 0x9E3779B97F4A7C15ull ^ (argc > 3 ? (uint64_t)atol(argv[3]) : 1); s[1] = 0xD1B54A32D192ED03ull;
  for (long i = 0; i < n; i++) {
    double x, y, a, b, a2 = 0, b2 = 0;
    int mode = (int)(rnd() % 4);
    if (!strcmp(f, "exp")) {
      x = mode == 0 ? 0.01 + 690 * unif() : mode == 1 ? logu(-7, 9.5, 0) : mode == 2 ? logu(-60, 10, 1) : -745 + 1500 * unif();
      a = exp(x); b = rmh_exp(x); y = 0;
    } else if (!strcmp(f, "log")) {
      x = mode == 0 ? logu(-40, 40, 0) : mode == 1 ? 0.9 + 0.2 * unif() : mode == 2 ? logu(-1074, 1023, 0) : 20 * unif();
      a = log(x); b = rmh_log(x); y = 0;
    } else if (!strcmp(f, "pow")) {
      x = mode == 0 ? logu(-34, 34, 0) : mode == 1 ? 10 * unif() : mode == 2 ? logu(-3, 3, 0) : (double)(1 + rnd() % 9);
      y = mode == 0 ? logu(-7, 7, 1) : mode == 1 ? 1.0 / (double)(1 + rnd() % 100) : mode == 2 ? logu(-7, 34, 1) : (double)((long)(rnd() % 41) - 20);
      if (rnd() % 16 == 0) { x = logu(-1074, -1000, 0); }
      a = pow(x, y); b = rmh_pow(x, y);
    } else if (!strcmp(f, "sincos")) {
      x = mode == 0 ? -M_PI + 2 * M_PI * unif() : mode == 1 ? logu(-30, 2, 1) : mode == 2 ? logu(-3, 26, 1) : M_PI * (double)((long)(rnd() % 2001) - 1000) / 1000.0;
      if (rnd() % 4 == 0) x = -400 + 800 * unif(); /* --max-trig-cycles: to 100 pi */
      sincos(x, &a, &a2); rmh_sincos(x, &b, &b2); y = 0;
    } else if (!strcmp(f, "tan")) {
      x = mode == 0 ? -M_PI + 2 * M_PI * unif() : mode == 1 ? logu(-30, 2, 1) : mode == 2 ? -25 + 50 * unif() : M_PI * (double)((long)(rnd() % 2001) - 1000) / 1000.0;
      if (rnd() % 4 == 0) x = (rnd() % 2) ? -400 + 800 * unif() : logu(4, 26, 1); /* --max-trig-cycles: to 100 pi, and up to 1e8 */
      a = tan(x); b = rmh_tan(x); y = 0;
    } else {
      y = mode == 0 ? logu(-34, 34, 1) : mode == 1 ? -10 + 20 * unif() : mode == 2 ? logu(-3, 3, 1) : (double)((long)(rnd() % 19) - 9);
      x = mode == 0 ? logu(-34, 34, 1) : mode == 1 ? -10 + 20 * unif() : mode == 2 ? logu(-3, 3, 1) : (double)((long)(rnd() % 19) - 9);
      if (rnd() % 8 == 0) x = logu(-200, 200, 1);
      if (x == 0 && y == 0) continue;
      a = atan2(y, x); b = rmh_atan2(y, x);
    }
    if (!same(a, b) || !same(a2, b2)) {
      if (bad < 8) printf("MISMATCH %s(%a, %a): libm %a %a  rmath %a %a\n", f, x, y, a, a2, b, b2);
      bad++;
    }
  }
  printf("%s: %ld / %ld mismatches\n", f, bad, n);
  return bad != 0;
}
