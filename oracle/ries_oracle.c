/* TEST INFRASTRUCTURE -- never built into, linked with or executed by the product (ries_b200/).
 *
 * ries_oracle: a plain-C, single-threaded CPU restatement of the hot path of the reference
 * (/root/reference/public/ries.c, "2024 May 20"), with the reference's own STREAMING semantics:
 * depth-first expression generation, one insert at a time into a binary search tree keyed on the
 * exact double value, immediate-neighbour match test at the moment of insertion, Newton refinement,
 * report.  It exists so that parity of the CUDA path can be checked on machines where
 * /root/reference is absent, and as an independent second derivation next to the batch-semantics
 * emulation (oracle/emul.cc).
 *
 * Pinned: tests/test_oracle_cpu.py compares its output byte for byte with tests/golden/*.txt, which
 * tools/make_golden.py produced by running the UNMODIFIED reference here (gcc -O2, glibc 2.39).
 * libm: stock glibc by default; `make ries_oracle_rm` interposes the repo's rmath instead.
 *
 * Each function cites the reference lines it restates.  Scope = SURVEY.md section 8: options
 * -l -S -O -N -E -i/-ie -r/-re -a/-ae -c/-ce -F0/-F3 -x -n --wide --max-match-distance
 * --no-slow-messages -Dy; output formats postfix and FORTH.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#define MAXLEN 21                 /* MAX_ELEN ries.c:2779 */
#define STKMAX ((MAXLEN + 1) / 2) /* MS_STK_MAX 3105 */
enum { T_NONE, T_TRAN, T_TCEL, T_ALG, T_CONS, T_RAT, T_INT }; /* 2927-2939 */
enum { KK = 1, A1 = 2, A2 = 4, An = 8, Ar = 0x10, A55 = 0x20, AjK = 0x40, ARHS = 0x80, Asq = 0x100,
       A1K = 0x200, Al = 0x400, AE = 0x800, Api = 0x1000, AKxK = 0x2000, AKpK = 0x4000, Aplus = 0x8000,
       Amul = 0x10000, Apow = 0x20000, Aa1e = 0x40000, Aa11 = 0x80000, Aa1r = 0x100000 }; /* 3279-3299 */
#define STOP_EXIT (-9998) /* EXIT_NO_ERROR 4541 */

/* ---------------- symbols (add_symbol calls 10375-10468) ---------------- */
typedef struct { char ch, seft; int w; const char *forth, *terse; } symdef;
static const symdef DEFS[] = {
  {'1','a',0,0,0},{'f','a',8,"phi","f = phi, the golden ratio, (1+sqrt(5))/2"},{'2','a',3,0,0},
  {'e','a',6,"e","e = base of natural logarithms, 2.71828..."},{'3','a',5,0,0},{'p','a',4,"pi","p = pi, 3.14159..."},
  {'4','a',6,0,0},{'5','a',7,0,0},{'6','a',8,0,0},{'7','a',8,0,0},{'8','a',9,0,0},{'9','a',9,0,0},{'x','a',5,0,0},
  {'n','b',-3,"neg","n = negative"},{'r','b',-3,"recip",0},{'s','b',-1,"dup*",0},{'q','b',-1,"sqrt","q = square root"},
  {'l','b',3,"ln","l = ln, natural logarithm or log base e"},{'E','b',3,"exp",0},
  {'S','b',3,"sinpi","S(x) = sinpi(x) = sin(pi * x)"},{'C','b',3,"cospi","C(x) = cospi(x) = cos(pi * x)"},
  {'T','b',6,"tanpi","T(x) = tanpi(x) = tan(pi * x)"},{'W','b',5,"W","W(x) = LambertW(x) = inverse(x=w*e^w)"},
  {'+','c',-6,"+",0},{'-','c',-5,"-",0},{'*','c',-6,"*",0},{'/','c',-5,"/",0},{'^','c',-4,"**",0},
  {'v','c',-3,"root","A v B = Ath root of B"},{'L','c',-1,"logN","A L B = logarithm to base A of B = ln(B) / ln(A)"},
  {'A','c',-1,"atan2","s A c = arctan2(s,c)"},
};
#define NDEFS ((int)(sizeof(DEFS) / sizeof(DEFS[0])))
static const char *INFIX_NAME[256];
static const char *DESC[256];

static int alwd[256], wgt[256], mask[256], amkey[256], cnt[256];
static char seft_of[256];
static const char *forth_of[256], *terse_of[256];
static unsigned char lst[3][20];
static int nlst[3], minw[3], maxw[3], s_minw, exhaust_cpx;

/* ---------------- options / thresholds ---------------- */
static const char *argv0;
static double levadj = 2.0, tlevel = 4.0, target = 1.0, exec_x;
static int got_target, restrict_sub, restrict_exp, restrict_trig, exact_exit, relative_x = 1, wide_out;
static int show_ss, S_opt, NOS_opt, slow_msg = 1, fmt = 2, debug_y, no_simplify;
static int one_sided, anagram_opt, preempt_w[256]; /* --one-sided 11600, --numeric-anagram 11586, --symbol-weights 11689 */
static double max_matches = 100, k_max_match_dist = -0.01, min_matchsize;
static double k_pi, k_e, k_phi, scale_trig, sig_loss = 0.01, vanished_dx = 1e-6, p_ovr, n_ovr;
static double best_match, k_min_best_match, init_match_dist, precision_ulp;
static int tg_phi = T_CONS, tg_e = T_TCEL, tg_pi = T_TCEL, targ_tag;
static int got_exact; static double num_matches;
static int lmin = 1, lmax = 1, rmin = 1, rmax = 1;
static double lhs_gen, rhs_gen, gen_total, lhs_prune, rhs_prune, g_ne, prune_count;
static long lhs_insert, rhs_insert, insert_count, nodes_total;
static double t0;
static int def_needed[256], def_given[256];

static double now(void) { struct timeval tv; gettimeofday(&tv, 0); return tv.tv_sec + 1e-6 * tv.tv_usec; }
static void print_end(int code);

static void help(void) { /* brief_help 3755 */
  printf("%s",
"Usage:\n  ries [options] target-value\n\n"
"Target value is required and may be any number. Options include:\n"
"  -l3   Search further (``level-3 search'')\n"
"  -x    Show matched values as ``x = value'' rather than ``x = T + epsilon''\n"
"  -s    (Sort of) solve by transforming to ``x = ...'' form\n"
"  -N+-/ Do not use symbols +, - or /\n"
"  -S    Show list of all available symbols\n");
  printf("%s",
"  -l    Restrict to Liouvillian solutions\n"
"  -a    Restrict to algebraic solutions\n"
"  -c    Restrict to ``constructible'' (straightedge and compass) solutions\n"
"  -r    Restrict to rational solutions (-re for exact match)\n"
"  -i    Restrict to integer solutions (-ie for exact match)\n\n"
"There are many more options; get the full manual at mrob.com/ries\n");
}

/* ---------------- number printing (ries_to_digits 3839, ries_spfg 4069) ---------------- */
static double p10(int p) { double r = 1, b = 10; int n = p < 0 ? -p : p; while (n) { if (n & 1) r *= b; n >>= 1; if (n) b *= b; } return p < 0 ? 1 / r : r; }
static void spfg(char *out, int prec, double x) {
  char d[48]; int e, i, sign = 0, n = prec + 1; char *o = out;
  memset(d, 0, sizeof d);
  if (x < 0) { x = -x; sign = 1; }
  if (x == 0) { for (i = 0; i < prec; i++) d[i] = '0'; e = 0; }
  else {
    double r = x; e = (int)floor(log10(x));
    if (e < -300) { r = r * p10(300); r = r / p10(e + 300); } else if (e > 0) r = r / p10(e); else if (e < 0) r = r * p10(-e);
    if (r >= 10) { r /= 10; e++; } else if (r < 1) { r *= 10; e--; }
    for (i = 0; i < n; i++) { int k = (int)r; r = (r - k) * 10; d[i] = (char)('0' + k); }
    for (i = n - 1; i > 0; i--) if (d[i] < '0') { d[i - 1]--; d[i] += 10; }
    if (d[n - 1] >= '5') { d[n - 2]++; i = n - 2; while (i > 0 && d[i] > '9') { d[i] -= 10; i--; d[i]++; } }
    if (d[0] > '9') { e++; for (i = prec; i >= 2; i--) d[i] = d[i - 1]; d[0] = '1'; d[1] = '0'; }
    d[prec] = 0;
  }
#define TRIM(p) do { char *t_ = (p) + strlen(p); while (t_ > (p) && t_[-1] == '0') *--t_ = 0; } while (0)
  if (sign) *o++ = '-';
  if (e > 0 && e < prec) {
    i = 0; *o++ = d[i++]; while (e--) *o++ = d[i++];
    TRIM(d); if (!*d) { *o = 0; return; }
    *o++ = '.'; for (; d[i] && i < prec; i++) *o++ = d[i]; *o = 0; return;
  } else if (e == 0) {
    *o++ = d[0]; TRIM(d + 1); if (!d[1]) { *o = 0; return; }
    *o++ = '.'; for (i = 1; d[i] && i < prec; i++) *o++ = d[i]; *o = 0; return;
  } else if (e < 0 && e > -5) {
    *o++ = '0'; *o++ = '.'; for (e++; e < 0; e++) *o++ = '0';
    TRIM(d); if (!*d) { *o = 0; return; }
    for (i = 0; d[i] && i < prec; i++) *o++ = d[i]; *o = 0; return;
  }
  *o++ = d[0]; TRIM(d + 1);
  if (d[1]) { *o++ = '.'; for (i = 1; i < prec && d[i]; i++) *o++ = d[i]; }
  sprintf(o, "e%s%d", e >= 0 ? "+" : "", e);
}
static const char *pfw(double x, int width) { /* pf_intfloat_wid 7542 */
  static char b[8][40]; static int k; char f1[16], f2[16]; char *s = b[k++ & 7];
  int w = x < 0 ? width - 1 : width, i;
  sprintf(f1, "%%.%dg", w); sprintf(f2, "%%.%de", w - 6);
  snprintf(s, 40, f1, x);
  if (strchr(s, 'e') || strchr(s, 'E')) snprintf(s, 40, f2, x);
  else if (strchr(s, '.')) { i = (int)strlen(s) - 1; while (i > 0 && s[i] == '0') s[i--] = 0; if (i > 0 && s[i] == '.') s[i] = 0; }
  return s;
}

/* ---------------- one symbol on the value stack: exec() 5308-5950 ---------------- */
typedef struct { double v, d; int t; } cell;
static int imin(int a, int b) { return a < b ? a : b; }
#define FIN(z) ((z) < HUGE_VAL && (z) > -HUGE_VAL)
static int trig_tag(int ta) { return ta >= T_RAT ? T_ALG : T_NONE; } /* default trig scale only */

/* Applies `op` to stack s (depth *sp). Returns nonzero on the reference's error returns. */
static int apply(unsigned char op, cell *s, int *sp, int dx) {
  cell r; double a, b, da, db, lb; int ta, tb, f1;
  r.d = 0; r.t = T_NONE;
  switch (seft_of[op]) {
  case 'a':
    r.t = T_INT;
    if (op >= '1' && op <= '9') r.v = op - '0';
    else if (op == 'f') { r.v = k_phi; r.t = tg_phi; } else if (op == 'e') { r.v = k_e; r.t = tg_e; }
    else if (op == 'p') { r.v = k_pi; r.t = tg_pi; } else if (op == 'x') { r.v = exec_x; r.d = 1; r.t = targ_tag; }
    else return -10;
    s[(*sp)++] = r; return 0;
  case 'b':
    a = s[*sp - 1].v; da = s[*sp - 1].d; ta = s[*sp - 1].t;
    switch (op) {
    case 'n': if (dx) r.d = -da; r.v = -a; r.t = ta; break;                                       /* 5374 */
    case 'r': if (a == 0) return -1; if (dx) r.d = (-da) / (a * a); if (!FIN(r.d)) return -22;      /* 5385 */
      r.v = 1.0 / a; r.t = imin(ta, T_RAT); break;
    case 's': if (dx) r.d = 2.0 * a * da; r.v = a * a; r.t = ta; break;                           /* 5403 */
    case 'q': if (a < 0) return -2; r.v = sqrt(a); if (dx) r.d = da / (2.0 * r.v);                 /* 5413 */
      r.t = (ta == T_INT && r.v < 1.0e6 && r.v == floor(r.v)) ? T_INT : imin(T_CONS, ta); break;
    case 'l': if (restrict_exp && (da != 0 || ta < restrict_exp)) return -20;                      /* 5436 */
      if (a <= 0) return -3; if (fabs(a - 1.0) <= sig_loss) return -7;
      if (dx) r.d = da / a; if (!FIN(r.d)) return -22; r.v = log(a); r.t = imin(ta, T_TCEL); break;
    case 'E': if (restrict_exp && (da != 0 || ta < restrict_exp)) return -20;                      /* 5464 */
      if (a > 690.0) return -4; if (a < sig_loss) return -7;
      r.v = exp(a); if (dx) r.d = r.v * da; r.t = imin(ta, T_TCEL); break;
    case 'S': if (restrict_trig && (da != 0 || ta < restrict_trig)) return -21;                    /* 5488 */
      a *= scale_trig; if (a > k_pi * 1.0 || -a > k_pi * 1.0) return -5;
      r.v = sin(a); if (fabs(r.v) > 0.99999) return -6; if (fabs(cos(a)) < sig_loss) return -7;
      if (fabs(r.v) < fabs(a) * sig_loss) return -7;
      if (dx) { r.d = cos(a) * scale_trig * da; if (fabs(r.d) < fabs(da) * sig_loss) return -7; }
      r.t = trig_tag(ta); break;
    case 'C': if (restrict_trig && (da != 0 || ta < restrict_trig)) return -21;                    /* 5541 */
      a *= scale_trig; if (a > k_pi * 1.0 || -a > k_pi * 1.0) return -5;
      r.v = cos(a); if (fabs(r.v) < fabs(a) * sig_loss) return -7; if (fabs(sin(a)) < sig_loss) return -7;
      if (fabs(r.v) > 0.99999) return -6;
      if (dx) r.d = 0.0 - (sin(a) * scale_trig * da);
      r.t = trig_tag(ta); break;
    case 'T': if (restrict_trig && (da != 0 || ta < restrict_trig)) return -21;                    /* 5582 */
      a *= scale_trig; if (a > k_pi * 1.0 || -a > k_pi * 1.0) return -5;
      r.v = tan(a); if (fabs(r.v) > 1 / sig_loss) return -7; if (a != 0 && fabs(r.v / a) < sig_loss) return -7;
      if (dx) r.d = scale_trig * (1.0 + r.v * r.v) * da;
      r.t = trig_tag(ta); break;
    default: return -10;
    }
    s[*sp - 1] = r; return 0;
  case 'c':
    b = s[*sp - 1].v; db = s[*sp - 1].d; tb = s[*sp - 1].t;
    a = s[*sp - 2].v; da = s[*sp - 2].d; ta = s[*sp - 2].t;
    switch (op) {
    case '-': b = -b; db = -db; /* fall through: shares the significance tests of '+' (5655) */
    case '+': r.v = a + b;                                                                         /* 5662 */
      if (fabs(a) < fabs(b) * sig_loss || fabs(b) < fabs(a) * sig_loss) if ((r.v - b) != a || (r.v - a) != b) return -7;
      if (fabs(r.v) < fabs(a) * sig_loss || fabs(r.v) < fabs(b) * sig_loss) if ((r.v - b) != a || (r.v - a) != b) return -7;
      if (dx) r.d = da + db; r.t = imin(ta, tb); break;
    case '*': if (dx) r.d = (b * da) + (a * db); r.v = a * b; r.t = imin(ta, tb); break;           /* 5701 */
    case '/': if (b == 0) return -1; if (dx) r.d = ((b * da) - (a * db)) / (b * b);                /* 5714 */
      if (!FIN(r.d)) return -22; r.v = a / b;
      r.t = (ta >= T_RAT && tb >= T_RAT) ? (r.v == floor(r.v) ? T_INT : T_RAT) : imin(ta, tb); break;
    case '^': if (restrict_exp && (db != 0 || tb < restrict_exp)) return -20;                      /* 5739 */
      if (a <= 0) return -8; if (fabs(b) < sig_loss) return -7;
      r.v = pow(a, b);
      if (dx) { r.d = r.v * ((b * da / a) + (log(a) * db)); if (!FIN(r.d)) return -22; if ((da || db) && r.d == 0) return -23; }
      if (tb == T_INT) r.t = b >= 0 ? ta : imin(ta, T_RAT);
      else if (tb == T_RAT) { lb = fabs(b); r.t = (lb == 0.5 || lb == 0.25 || lb == 0.125) ? imin(ta, T_CONS) : imin(ta, T_ALG); }
      else if (tb == T_CONS || tb == T_ALG) r.t = imin(ta, T_TCEL); else r.t = imin(ta, tb);
      break;
    case 'v': if (restrict_exp && (db != 0 || tb < restrict_exp)) return -20;                      /* 5793 */
      if (b == 0) return -4; if (fabs(b) > 1.0 / sig_loss) return -7;
      f1 = 0; if (b == 3.0) { if (a < 0) { a = -a; f1 = 1; } } else if (a < 0) return -2;
      r.v = pow(a, 1.0 / b);
      if (dx) { r.d = r.v * ((da / (a * b)) - (log(a) * db / (b * b))); if ((da || db) && r.d == 0) return -23; }
      if (f1) r.v = -r.v;
      if (!FIN(r.d)) return -22;
      if (tb >= T_RAT) { lb = fabs(b); r.t = (lb == 2.0 || lb == 4.0 || lb == 8.0) ? imin(ta, T_CONS) : imin(ta, T_ALG); }
      else if (tb == T_CONS || tb == T_ALG) r.t = imin(ta, T_TCEL); else r.t = imin(ta, tb);
      break;
    case 'L': if (restrict_exp && (db != 0 || tb < restrict_exp)) return -20;                      /* 5851 */
      if (b <= 0 || b == 1.0) return -9; if (fabs(1.0 - b) < sig_loss) return -7;
      if (restrict_exp && (da != 0 || ta < restrict_exp)) return -20;
      if (a <= 0) return -3; if (fabs(a - 1.0) <= sig_loss) return -7;
      lb = log(b); r.v = log(a) / lb;
      if (dx) { r.d = (da / (a * lb)) - (r.v * db / (b * lb)); if (!FIN(r.d)) return -22; if ((da || db) && r.d == 0) return -23; }
      r.t = imin(ta, imin(tb, T_TCEL)); break;
    case 'A': r.v = (a == 0 && b == 0) ? 0 : atan2(a, b);                                           /* 5895, 5188 */
      if (dx) r.d = (b * da - a * db) / (b * b + a * a);
      r.t = T_NONE; break;
    default: return -10;
    }
    (*sp)--; s[*sp - 1] = r; return 0;
  }
  return -10;
}

/* eval() 6777: whole expression at the current exec_x */
static int eval(const char *e, double *v, double *d) {
  cell s[STKMAX + 2]; int sp = 0, i, err, hasx = strchr(e, 'x') != 0;
  *v = 0; *d = 0;
  for (i = 0; e[i]; i++) {
    unsigned char c = (unsigned char)e[i];
    if (i >= MAXLEN) return -13;
    if (seft_of[c] == 'a' && sp + 1 > STKMAX) return -16;
    if ((seft_of[c] == 'b' && sp < 1) || (seft_of[c] == 'c' && sp < 2)) return -17;
    if (!seft_of[c]) return -15;
    if ((err = apply(c, s, &sp, hasx))) return err;
  }
  if (sp > 0) { *v = s[sp - 1].v; *d = s[sp - 1].d; }
  return 0;
}

/* ---------------- the value database: bt_insert 8622 ---------------- */
typedef struct node { double val, der; struct node *l, *r; char sym[MAXLEN + 2]; } node;
static node *root;
static node *pool; static long pool_left;
static node *new_node(void) { if (!pool_left) { pool = malloc(sizeof(node) * 65536); pool_left = 65536; } pool_left--; nodes_total++; return pool++; }

/* bt_stats 8399-8428 (called at 12515 under -Dy): node count and average node depth (root = 1), summed in an
   int like the reference.  Iterative walk: the tree is unbalanced and can be deep. */
static void bt_stats(void) {
  long n = 0; int tdepth = 0;
  static node **stk; static int *dep; static long cap;
  long sp = 0;
  if (!cap) { cap = 1 << 16; stk = malloc(cap * sizeof *stk); dep = malloc(cap * sizeof *dep); }
  if (root) { stk[0] = root; dep[0] = 1; sp = 1; }
  while (sp) {
    node *it = stk[--sp]; int d = dep[sp];
    n++; tdepth += d;
    if (sp + 2 > cap) { cap *= 2; stk = realloc(stk, cap * sizeof *stk); dep = realloc(dep, cap * sizeof *dep); }
    if (it->l) { stk[sp] = it->l; dep[sp++] = d + 1; }
    if (it->r) { stk[sp] = it->r; dep[sp++] = d + 1; }
  }
  printf("Current tree stats:\n");
  printf("       Nodes: %ld\n", n);
  printf("  Avg. Depth: %f\n", ((double)tdepth) / ((double)n));
}

static int complexity(const char *e) { int c = 0; for (; *e; e++) c += wgt[(unsigned char)*e]; return c; }

static void defsym_used(const char *e) { for (; *e; e++) if (terse_of[(unsigned char)*e]) def_needed[(unsigned char)*e] = 1; }

static void check_exit(int is_exact) { /* 7805 */
  if (is_exact && restrict_sub == T_INT && exact_exit) printf("  (Stopping now because -ie option was given.)\n");
  else if (is_exact && restrict_sub == T_RAT && exact_exit) printf("  (Stopping now because -re option was given.)\n");
  else if (is_exact && restrict_sub == T_CONS && exact_exit) printf("  (Stopping now because -ce option was given.)\n");
  else if (is_exact && k_max_match_dist == 0) printf("  (Stopping now because an exact match was found.)\n");
  else if (is_exact && exact_exit) printf("  (Stopping now because --min-match-distance 0 or --max-match-distance 0\n  was given.)\n");
  else if (best_match < min_matchsize)
    printf("  (Stopping now because best match is within %7.3g of target value.)\n", k_min_best_match > min_matchsize ? k_min_best_match : min_matchsize);
  else if (num_matches >= max_matches) {
    if (num_matches == 1) printf("  (Stopping now because 1 match was found.)\n");
    else printf("  (Stopping now because %ld matches were found.)\n", (long)num_matches);
  } else return;
  print_end(STOP_EXIT);
}

static void fmt_expr(const char *e, char *out) { /* postfix 6561 / postfix_formatter 6572 */
  if (fmt == 0) { strcpy(out, e); return; }
  out[0] = 0;
  for (; *e; e++) { const char *f = forth_of[(unsigned char)*e]; if (f) strcat(out, f); else strncat(out, e, 1); strcat(out, " "); }
  while (*out && out[strlen(out) - 1] == ' ') out[strlen(out) - 1] = 0;
}

static void report(const char *le, const char *re, double rootv, double delta) { /* report_match 7887 */
  char lf[512], rf[512], fval[64], froot[64]; int ll, rl, width, lp, rp, tw, over, posn = 0, i; const char *intro = "for x ";
  if (delta == 0 && got_exact) return;
  fmt_expr(le, lf); fmt_expr(re, rf); ll = (int)strlen(lf); rl = (int)strlen(rf);
  width = relative_x ? 44 : 40; tw = width + 5; lp = one_sided ? width / 8 : width / 2; rp = width - lp; /* 8113 */
  lp -= ll; if (lp < 0) lp = 0; rp -= rl; if (rp < 0) rp = 0;
  width = lp + ll + 5 + rl + rp;
  while (width > tw && lp > 0) { lp--; width--; }
  while (width > tw && rp > 0) { rp--; width--; }
  over = width - tw; if (over < 0) over = 0;
  for (i = 0; i < lp; i++) putchar(' ');
  printf(" %s = %s ", lf, rf);
  for (i = 0; i < rp; i++) putchar(' ');
  spfg(froot, 15, rootv);
  if (wide_out) { printf("x = %-20s  ", froot); intro = ""; }
  fval[0] = 0;
  if (delta == 0) { printf("('exact' match)"); posn = 15; got_exact = 1; }
  else if (!relative_x) { printf("for x = "); strcpy(fval, froot); posn = 8 + (int)strlen(fval); }
  else if (delta < 0) { delta = -delta; printf("%s= T - ", intro); snprintf(fval, 30, "%.6g", delta); posn = 6 + (int)strlen(intro) + (int)strlen(fval); }
  else { printf("%s= T + ", intro); snprintf(fval, 30, "%.6g", delta); posn = 6 + (int)strlen(intro) + (int)strlen(fval); }
  width = relative_x ? 23 : 8 + 15 + 6;
  printf("%s", fval);
  for (i = posn + over; i < width; i++) putchar(' ');
  if (wide_out) {
    char tmp[64]; double ratio = delta != 0 ? fabs(target / delta) : 0;
    if (delta == 0) strcpy(tmp, "(1 part in infinity)");
    else if (ratio < 100) { int r10 = (int)(ratio * 10 + 0.5), r1 = r10 % 10; snprintf(tmp, 30, "(1 part in %d.%d)", (r10 - r1) / 10, r1); }
    else if (ratio < 100000000) snprintf(tmp, 30, "(1 part in %d)", (int)(ratio + 0.5));
    else snprintf(tmp, 30, "(1 part in %.3e)", ratio);
    printf("%-22s", tmp);
  }
  printf(" {%d}\n", complexity(le) + complexity(re));
  num_matches++;
  defsym_used(le); defsym_used(re);
  check_exit(delta == 0);
}

static int ends(const char *s, const char *t) { size_t a = strlen(s), b = strlen(t); return a && b && a >= b && !strcmp(s + a - b, t); }
static void clip(char *s, size_t n) { size_t a = strlen(s); if (a >= n) s[a - n] = 0; }

/* cv_simplify 7297 + newton 7139 */
static int refine(char *l, char *r, double *rootv) {
  int go = !no_simplify, i, err, rx; double rv, rd, lv, ld, cur, prev, prv2;
  while (go) {
    go = 0;
    if (ends(l, "2/") && ends(r, "2/")) { clip(l, 2); clip(r, 2); go = 1; }
    if (ends(l, "2*") && ends(r, "2*")) { clip(l, 2); clip(r, 2); go = 1; }
    if (ends(l, "r") && ends(r, "r")) { clip(l, 1); clip(r, 1); go = 1; }
    if (ends(l, "n") && ends(r, "n")) { clip(l, 1); clip(r, 1); go = 1; }
    if (ends(l, "nr") && ends(r, "n")) { clip(l, 2); strcat(l, "r"); clip(r, 1); go = 1; }
    if (ends(l, "n") && ends(r, "nr")) { clip(l, 1); clip(r, 2); strcat(r, "r"); go = 1; }
    if (ends(l, "rr")) { clip(l, 2); go = 1; }
    if (ends(r, "rr")) { clip(r, 2); go = 1; }
    if (ends(l, "r2/") && ends(r, "r")) { clip(l, 3); strcat(l, "2*"); clip(r, 1); go = 1; }
    if (ends(l, "r") && ends(r, "r2/")) { clip(l, 1); clip(r, 3); strcat(r, "2*"); go = 1; }
  }
  *rootv = target; rx = strchr(r, 'x') != 0;
  if ((err = eval(r, &rv, &rd))) return err;
  cur = target; prev = cur - 1.0; prv2 = prev - 1.0; lv = ld = 0;
  for (i = 0; i < 12; i++) {
    prv2 = prev; prev = cur; exec_x = cur;
    if ((err = eval(l, &lv, &ld))) { exec_x = target; return err; }
    if (ld == 0) { exec_x = target; return -11; }
    if (rx && (err = eval(r, &rv, &rd))) { exec_x = target; return err; }
    cur = cur + ((rv - lv) / (ld - rd));
    if (cur == prev) i = 13;
    else if (cur == prv2) {
      double m = fabs(cur - prev) * 9007199254740992.0 / fabs(cur);
      if (m <= 2.0) { cur = prev; i = 13; } else if (i >= 6 && m <= 4.0) { cur = prev; i = 13; }
    }
  }
  exec_x = target;
  if (cur == prev || fabs(cur - prev) < fabs(target * 1.0e-15)) { *rootv = cur; return 0; }
  return -12;
}

static void check_match(const node *lhs, const node *rhs) { /* 8243 */
  double td = lhs->der - rhs->der, score, rootv, delta; char l[MAXLEN + 4], r[MAXLEN + 4];
  if (td == 0) { fprintf(stderr, "check_match got dx = 0!\n"); return; }
  score = fabs((rhs->val - lhs->val) / td);
  if (score > best_match || score < min_matchsize) return;
  strcpy(l, lhs->sym); strcpy(r, rhs->sym);
  if (refine(l, r, &rootv)) return;
  delta = rootv - target; score = fabs(delta);
  if (score == 0) report(l, r, rootv, delta);
  else if (score < best_match) { report(l, r, rootv, delta); best_match = score * 0.999 - k_min_best_match; check_exit(0); }
}

static void exact_match(const node *it, double new_dx, const char *sym) { /* check_exact_match 8541 */
  const char *l, *r; double td, x, eps, score;
  if (new_dx == 0) { if (it->der == 0) return; if (fabs(it->der) < vanished_dx) return; l = it->sym; r = sym; td = it->der; }
  else { if (it->der != 0) return; if (fabs(new_dx) < vanished_dx) return; l = sym; r = it->sym; td = 0 - new_dx; }
  x = it->val; eps = x * 1.0e-14;
  if (x == x + eps) return;
  if (td == 0) return;
  score = fabs(eps / td);
  if (best_match > 0 && score > best_match) return;
  if (score < min_matchsize) return;
  report(l, r, target, 0);
}

/* Returns 1 when a node was added. Predecessor/successor of a new leaf = the last ancestors where the
   descent turned right/left (what bt_prev/bt_next 8431-8480 find through parent pointers). */
static int insert(double x, double dx, const char *sym) {
  node *it = root, *pred = 0, *succ = 0, *n;
  while (it) {
    if (it->val == x) { exact_match(it, dx, sym); return 0; }
    if (x < it->val) { succ = it; if (!it->l) break; it = it->l; }
    else { pred = it; if (!it->r) break; it = it->r; }
  }
  n = new_node(); n->val = x; n->der = dx; n->l = n->r = 0; strcpy(n->sym, sym);
  if (!it) root = n; else if (x < it->val) it->l = n; else it->r = n;
  insert_count++;
  if (it) { /* check_sides 8488 (the very first node has no neighbours) */
    if (dx == 0) {
      if (pred && pred->der != 0) check_match(pred, n);
      if (succ && succ->der != 0) check_match(succ, n);
    } else {
      if (pred && pred->der == 0) check_match(n, pred);
      if (succ && succ->der == 0) check_match(n, succ);
    }
  }
  return 1;
}

/* ---------------- expression generation: ge_2 8908, ge_1 9481, gf_1 9599 ---------------- */
typedef struct { int len; char cls[MAXLEN + 3]; int stk[MAXLEN + 3], arg1[MAXLEN + 3], rmin[MAXLEN + 4], rmax[MAXLEN + 4]; } form;
static char cur[MAXLEN + 3];

static double gen(const form *F, int ip, int cplx, cell *s, int sp, int e_min, int e_max, int lhs) {
  double n = 0; int atts, k, c;
  if (ip > 0) {
    cell save[2]; int osp = sp, err;
    if (sp > 0) save[0] = s[sp - 1];
    if (sp > 1) save[1] = s[sp - 2];
    err = apply((unsigned char)cur[ip - 1], s, &sp, lhs);
#define UNDO() do { if (seft_of[(unsigned char)cur[ip - 1]] != 'a') { s[osp - 1] = save[0]; if (osp > 1) s[osp - 2] = save[1]; } } while (0)
    if (err) { UNDO(); return 0; }
    {
      double v = s[sp - 1].v, d = s[sp - 1].d; int t = s[sp - 1].t, dead = 0;
      if (v == 0) dead = 1;                                                      /* 8962 */
      else if (restrict_sub == T_INT && t != T_INT) dead = 1;                    /* 8975 */
      else if (t < restrict_sub) dead = 1;                                       /* 8995 */
      else if (v >= p_ovr || v <= n_ovr) dead = 1;                               /* 9020 */
      else if (lhs && d != 0 && fabs(d) / (1.0 + fabs(v)) < vanished_dx) dead = 1; /* 9042 */
      else if (lhs && sp == 1 && fabs(d) / (1.0 + fabs(v)) < vanished_dx) dead = 1; /* 9061 */
      if (!dead && ip >= F->len) {                                               /* 9085 */
        if (lhs && fabs(d) < vanished_dx) dead = 1;                              /* 9109 */
        else if (v < -9.9e99 || v > 9.9e99) dead = 1;
        else { cur[ip] = 0; insert(v, d, cur); UNDO(); return 1; }
      }
      if (dead) { prune_count += 1; UNDO(); return 0; }
    }
    /* attributes of the prefix 9240-9379 */
    atts = lhs ? 0 : ARHS;
    if (ip > 2 && F->cls[ip - 1] == 'a' && F->cls[ip - 2] == 'c' && F->cls[ip - 3] == 'a' && cur[ip - 1] == cur[ip - 3]) {
      if (cur[ip - 2] == '*') atts |= AKxK; else if (cur[ip - 2] == '+') atts |= AKpK;
    }
    if (ip > 1) {
      if (F->cls[ip] == 'c' && F->arg1[ip] >= 0) {
        char a1 = cur[F->arg1[ip]];
        if (a1 == 'e') atts |= Aa1e; else if (a1 == '1') atts |= Aa11; else if (a1 == 'r') atts |= Aa1r;
      }
      if (F->cls[ip - 1] == 'a' && F->cls[ip - 2] == 'a') {
        if (cur[ip - 1] == cur[ip - 2]) atts |= KK;
        if (cur[ip - 1] < '6' && cur[ip - 2] < '6') atts |= A55;
        if (cur[ip - 2] == '1') atts |= A1K;
        if ((unsigned char)cur[ip - 1] > (unsigned char)cur[ip - 2]) atts |= AjK;
      }
    }
    atts |= amkey[(unsigned char)cur[ip - 1]];
    c = F->cls[ip] - 'a';
    for (k = 0; k < nlst[c]; k++) {                                              /* 9390-9465 */
      unsigned char sy = lst[c][k]; int tc = cplx + wgt[sy];
      if (tc + F->rmin[ip + 1] > e_max || tc + F->rmax[ip + 1] < e_min || (atts & mask[sy]) || cnt[sy] >= alwd[sy]) { prune_count += 1; continue; }
      cur[ip] = (char)sy; cnt[sy]++;
      n += gen(F, ip + 1, tc, s, sp, e_min, e_max, lhs);
      cnt[sy]--;
    }
    UNDO();
    return n;
  }
  /* ip == 0: first symbol is always class a; LHS must start with x (9433) */
  for (k = 0; k < nlst[0]; k++) {
    unsigned char sy = lst[0][k]; int tc = wgt[sy];
    if (tc + F->rmin[1] > e_max || tc + F->rmax[1] < e_min || ((lhs ? 0 : ARHS) & mask[sy]) || (lhs && sy != 'x') || cnt[sy] >= alwd[sy]) { prune_count += 1; continue; }
    cur[0] = (char)sy; cnt[sy]++;
    n += gen(F, 1, tc, s, 0, e_min, e_max, lhs);
    cnt[sy]--;
  }
  return n;
}

static double forms(form *F, int stack, int wmin, int wmax, int e_min, int e_max, int lhs) { /* gf_1 9599 */
  double n = 0; int s, blen = F->len, i;
  for (s = 0; s <= (nlst[2] ? 2 : 1); s += (nlst[1] ? 1 : 2)) {
    int nstack = stack + 1 - s, nmin = wmin + minw[s], nmax = wmax + maxw[s], rec = 1, ge = 0;
    F->cls[blen] = (char)('a' + s);
    if (nstack < 1 || nmin + s_minw > e_max || blen > MAXLEN || blen + 1 + nstack > MAXLEN + 1) rec = 0;
    F->stk[blen] = nstack; F->arg1[blen] = -1;
    if (s == 2) for (i = blen - 2; i >= 0; i--) if (F->stk[i] == nstack) { F->arg1[blen] = i; break; }
    if (nstack == 1 && nmin <= e_max && nmax >= e_min) ge = 1;
    if (ge && blen + 1 <= MAXLEN) { /* ge_1 9481: suffix weight bounds, then generate */
      cell st[STKMAX + 2]; int rm = 0, rx = 0;
      F->len = blen + 1; F->rmin[F->len] = F->rmax[F->len] = 0;
      for (i = F->len; i > 0;) { i--; rm += minw[F->cls[i] - 'a']; rx += maxw[F->cls[i] - 'a']; F->rmin[i] = rm; F->rmax[i] = rx; }
      g_ne += gen(F, 0, 0, st, 0, e_min, e_max, lhs);
    }
    if (rec) { F->len = blen + 1; n += forms(F, nstack, nmin, nmax, e_min, e_max, lhs); }
    F->len = blen;
    if (ge) n++;
  }
  return n;
}

/* ---------------- set-up: init1/parse_args/init2/main 10228-10744, 11117, 12111-12317 ---------------- */
static void setsyms(const char *s, int n) { for (; s && *s; s++) alwd[(unsigned char)*s] = n; }
static void restrict_rat(void) { restrict_sub = T_RAT; setsyms("pefqASCTlvLEW", 0); setsyms("+-*/nr", MAXLEN); setsyms("x", 1); setsyms("s^", 0); }
static void restrict_alg(int trig) { restrict_sub = T_ALG; setsyms("pelLEWA", 0); setsyms("+-*/nrsqfx", MAXLEN); setsyms("^v", MAXLEN); restrict_exp = T_RAT; setsyms("SCT", MAXLEN); if (trig) restrict_trig = T_RAT; }
static void bad_option(const char *a) { printf("%s: Unknown option '%s'\n\n", argv0, a); help(); print_end(-1); }

static int parse_target(const char *str) { /* 10758 */
  char s[256]; int j = 0, i, sec = 'b', dots = 0, init = 1, going = 1, last = -1; double t;
  for (i = 0; str[i] && j < 250; i++) if (str[i] != ' ') s[j++] = str[i] == ',' ? '.' : str[i];
  s[j] = 0;
  for (i = 0; s[i] && going; i++) {
    char c = s[i];
    if (sec == 'b') {
      if (c == '+' || c == '-') sec = 'm'; else if (init && c == '0') { sec = 'm'; last = i; }
      else if (c >= '0' && c <= '9') { sec = 'm'; last = i; init = 0; } else if (c == '.') { sec = 'm'; dots++; } else { s[i] = '0'; going = 0; }
    } else if (sec == 'm') {
      if (c == 'e' || c == 'E') sec = 'e'; else if (init && c == '0') last = i; else if (c >= '0' && c <= '9') { last = i; init = 0; }
      else if (!dots && c == '.') dots++; else { s[i] = '0'; going = 0; }
    } else going = 0;
  }
  if (last < 0 || sscanf(s, "%lf", &t) <= 0) return 0;
  target = t; return 1;
}

static void parse(int argc, char **argv) {
  int i; double t;
  for (i = 1; i < argc; i++) {
    const char *a = argv[i];
    if (!*a || !strcmp(a, "--ries-arguments-end")) break;
    if (!strcmp(a, "-")) continue;
    if (!strcmp(a, "--canon-simplify")) no_simplify = 0;
    else if (!strcmp(a, "--no-canon-simplify")) no_simplify = 1;
    else if (!strcmp(a, "--max-match-distance")) {
      if (++i < argc && sscanf(argv[i], "%lf", &t) > 0) {
        k_max_match_dist = t;
        if (t < 0) printf("First match must be closer than %g times your target value.\n", -t);
        else if (t == 0) printf("Only give an 'exact' match (if any) then exit.\n");
        else printf("First match must be closer than %g.\n", t);
      } else { printf("%s: --max-match-distance should be followed by a numeric argument.\n\n", argv0); help(); print_end(-1); }
    } else if (!strcmp(a, "--max-matches") || !strncmp(a, "-n", 2)) {
      const char *v = a[1] == 'n' ? a + 2 : (++i < argc ? argv[i] : 0);
      if (v && sscanf(v, "%lf", &t) > 0) {
        if (t >= 1) max_matches = t;
        else { printf("%s: -n or --max-matches argument should be 1 or more.\nexamples:  -n7  or  --max-matches 7\n\n", argv0); help(); print_end(-1); }
      } else { printf("%s: --max-matches should be followed by a numeric argument.\nexamples:  -n7  or  --max-matches 7\n\n", argv0); help(); print_end(-1); }
    }
    else if (!strcmp(a, "--no-slow-messages")) slow_msg = 0;
    else if (!strcmp(a, "--one-sided")) one_sided = 1;
    else if (!strcmp(a, "--numeric-anagram")) {
      if (i + 1 < argc && argv[i + 1][0] && argv[i + 1][0] != '-') { /* set_anagram 10202 */
        const char *ana = argv[++i], *as = "123456789efprs"; int k;
        for (k = 0; as[k]; k++) alwd[(unsigned char)as[k]] = 0;
        for (k = 0; ana[k]; k++) if (strchr(as, ana[k])) alwd[(unsigned char)ana[k]]++;
        anagram_opt = 1;
      } else { printf("%s: --numeric-anagram should be followed by a string of digits.\n\n", argv0); help(); print_end(-1); }
    }
    else if (!strcmp(a, "--symbol-weights")) {
      if (i + 1 < argc && argv[i + 1][0] && argv[i + 1][0] != '-') {
        while (i + 1 < argc && argv[i + 1][0] && argv[i + 1][0] != '-') {
          const char *arg = argv[++i]; char tmp[20]; int k = 0; double w;
          strncpy(tmp, arg, 19); tmp[19] = 0;
          while (tmp[k] == '.' || (tmp[k] >= '0' && tmp[k] <= '9')) k++;
          if (tmp[k] == ':' && tmp[k + 1]) {
            tmp[k++] = 0; w = strtod(tmp, 0);
            if (w <= 0.0) { printf("%s: Symbol weight may not be %f or less.\n", argv0, 0.0); print_end(-1); }
            if (w > 100.0) w = 100.0;
            preempt_w[(unsigned char)tmp[k]] = (int)floor(w + 0.5);
          } else {
            printf("%s: --symbol-weights argument syntax is NUMBER:<sym>, for example\n  12:^   to set weight of '^' to 12\n  Instead I got '%s'\n", argv0, arg);
            print_end(-1);
          }
        }
      } else {
        printf("%s: --symbol-weights should be followed by one or more tuples\n  of the form NUMBER:<sym>, for example 12:^   to set weight of '^' to 12\n", argv0);
        print_end(-1);
      }
    }
    else if (!strcmp(a, "--any-exponents")) restrict_exp = T_NONE;
    else if (!strcmp(a, "--any-subexpressions")) restrict_sub = T_NONE;
    else if (!strcmp(a, "--any-trig-args")) restrict_trig = T_NONE;
    else if (!strcmp(a, "--rational-exponents")) restrict_exp = T_RAT;
    else if (!strcmp(a, "--rational-trig-args")) restrict_trig = T_RAT;
    else if (!strcmp(a, "--wide-output") || !strcmp(a, "--wide")) { wide_out = 1; relative_x = 1; }
    else if (!strncmp(a, "-a", 2) || !strcmp(a, "--algebraic-subexpressions")) { restrict_alg(1); if (a[2] == 'e') exact_exit = 1; }
    else if (!strncmp(a, "-c", 2) || !strcmp(a, "--constructible-subexpressions")) {
      restrict_sub = T_CONS; if (a[2] == 'e') exact_exit = 1;
      setsyms("peASCTl^vLEW", 0); setsyms("+-*/nrsqf", MAXLEN); setsyms("x", 1);
    }
    else if (!strncmp(a, "-D", 2)) { if (strchr(a + 2, 'y')) debug_y = 1; }
    else if (!strncmp(a, "-E", 2)) { NOS_opt = 1; setsyms(a + 2, MAXLEN); }
    else if (!strncmp(a, "-F", 2)) {
      if (!a[2]) fmt = 3; else if (sscanf(a + 2, "%d", &fmt) <= 0) { printf("%s: -F parameter requires a number, e.g. '-F0'\n\n", argv0); help(); print_end(-1); }
    }
    else if (!strncmp(a, "-i", 2) || !strcmp(a, "--integer-subexpressions")) { restrict_sub = T_INT; setsyms("pefqASCTlvLEW", 0); if (a[2] == 'e') exact_exit = 1; }
    else if (sscanf(a, "-l%lf", &t) > 0) { levadj = t; tlevel = 2.0 + levadj; }
    else if (!strcmp(a, "-el") || !strcmp(a, "-l") || !strcmp(a, "--liouvillian-subexpressions") || !strcmp(a, "-le")) {
      restrict_alg(0); setsyms("eplEL", MAXLEN); restrict_sub = T_TRAN; restrict_exp = T_TCEL; restrict_trig = T_TCEL; if (a[2] == 'e') exact_exit = 1;
    }
    else if (!strncmp(a, "-l", 2)) { printf("%s: -l parameter requires a number, e.g. '-l3'\n\n", argv0); help(); print_end(-1); }
    else if (!strncmp(a, "-N", 2)) { NOS_opt = 1; setsyms(a + 2, 0); }
    else if (!strncmp(a, "-O", 2)) { NOS_opt = 1; setsyms(a + 2, 1); }
    else if (!strncmp(a, "-r", 2) || !strcmp(a, "--rational-subexpressions")) { restrict_rat(); if (a[2] == 'e') exact_exit = 1; }
    else if (!strncmp(a, "-S", 2)) {
      if (!a[2]) show_ss = 1;
      else { int c; S_opt = NOS_opt = 1; for (c = 0; c < 256; c++) if (c != 'x') alwd[c] = 0; setsyms(a + 2, MAXLEN); }
    }
    else if (!strcmp(a, "-x") || !strcmp(a, "--absolute-roots")) { if (wide_out) printf("Note: '%s' option with '--wide' is redundant.\n", a); else relative_x = 0; }
    else if ((a[0] >= '0' && a[0] <= '9') || a[0] == '-' || a[0] == '.') { if (parse_target(a)) got_target = 1; else bad_option(a); }
    else bad_option(a);
  }
}

static void rule(const char *req, char sy, int m) { /* add_rule 10014 */
  for (; *req; req++) if (alwd[(unsigned char)*req] < MAXLEN) return;
  mask[(unsigned char)sy] |= m;
}

static void init2(void) {
  int i; double k, pi;
  precision_ulp = 2.0 * ldexp(1.0, -54); k_min_best_match = 9.0072 * precision_ulp;   /* 7758-7763, 10067-10070 */
  k_phi = (sqrt(5.0) + 1.0) / 2.0;                                                     /* 10084 */
  for (k = 31.0, k_e = 1.0 / k; k > 1.0;) { k -= 1.0; k_e = 1.0 + k_e / k; }            /* 10107 */
  for (k = 103.0, pi = 1.57; k > 1.0;) { k -= 1.0; pi = 1.0 + (k / (1.0 + (2.0 * k))) * pi; } /* 10141 */
  k_pi = 2.0 * pi; scale_trig = k_pi;
  if (alwd['x'] != 1) alwd['x'] = MAXLEN;                                              /* 10336 */
  if (alwd['W']) { printf("%s: The Lambert W function 'W' requires the stand-alone maths library.\n", argv0); help(); print_end(1); }
  for (i = 0; i < NDEFS; i++) {
    unsigned char c = (unsigned char)DEFS[i].ch; int cl = DEFS[i].seft - 'a';
    seft_of[c] = DEFS[i].seft; wgt[c] = preempt_w[c] >= 0 ? preempt_w[c] : DEFS[i].w + 10; /* 9872 */ forth_of[c] = DEFS[i].forth; terse_of[c] = DEFS[i].terse;
    if (alwd[c]) lst[cl][nlst[cl]++] = c;
  }
  if (nlst[0] < 2) { printf("%s: You must allow at least one constant symbol.\n", argv0); if (S_opt) printf("  (Note that the -S option disables all other symbols)\n"); help(); print_end(1); }
  if (!nlst[1] && !nlst[2]) { printf("%s: You must allow at least one operator symbol.\n", argv0); help(); print_end(1); }
  for (i = 0; i < 3; i++) { minw[i] = 100; maxw[i] = 0; }                                /* setup_abc_mmw 9915 */
  for (i = 0; i < NDEFS; i++) { unsigned char c = (unsigned char)DEFS[i].ch; int cl = DEFS[i].seft - 'a';
    if (alwd[c]) { if (wgt[c] < minw[cl]) minw[cl] = wgt[c]; if (wgt[c] > maxw[cl]) maxw[cl] = wgt[c]; } }
  s_minw = minw[0]; if (minw[1] < s_minw) s_minw = minw[1]; if (minw[2] < s_minw) s_minw = minw[2];
  amkey['1'] = A1; amkey['2'] = A2; amkey['n'] = An; amkey['r'] = Ar; amkey['s'] = amkey['q'] = Asq; amkey['l'] = Al;
  amkey['E'] = AE; amkey['p'] = Api; amkey['+'] = Aplus; amkey['*'] = Amul; amkey['^'] = Apow;   /* 9822 */
  /* 10559-10665 */
  rule("", 'x', ARHS); rule("nr", 'n', Ar); rule("", 'n', An); rule("1", 'r', A1); rule("", 'r', Ar); rule("1", 's', A1);
  rule("4", 's', A2); rule("s", 's', An); rule("rs", 's', Ar); rule("4^", 's', Asq); rule("1", 'q', A1); rule("rq", 'q', Ar);
  rule("4v", 'q', Asq); rule("", 'l', A1); rule("ln", 'l', Ar); rule("", 'l', AE); rule("", 'E', Al); rule("Er", 'E', An);
  rule("Sn", 'S', An); rule("C", 'C', An); rule("2*", '+', KK); rule("*23456789", '+', A55); rule("+", '+', AjK);
  rule("-", '+', An); rule("", '-', KK); rule("1234n", '-', A55); rule("12345678n", '-', AjK); rule("+", '-', An);
  rule("s", '*', KK); rule("*", '*', AjK); rule("*", '*', A1); rule("*n", '*', An); rule("/", '*', Ar); rule("1", '/', KK);
  rule("r", '/', A1K); rule("", '/', A1); rule("/n", '/', An); rule("*", '/', Ar); rule("", '^', A1); rule("s", '^', A2);
  rule("^r", '^', An); rule("1", '^', A1K); rule("v", '^', Ar); rule("", 'v', A1); rule("q", 'v', A2); rule("vr", 'v', An);
  rule("1", 'v', A1K); rule("^", 'v', Ar); rule("1", 'L', KK); rule("", 'L', A1); rule("Ln", 'L', Ar); rule("", 'L', A1K);
  rule("", '/', AKxK); rule("", '-', AKpK); rule("", '+', Aplus); rule("", '*', Amul); rule("^", '^', Amul);
  rule("", 'q', Apow); rule("es", 'E', A2); rule("E", '^', Aa1e); rule("", '*', Aa11); rule("r", '/', Aa11);
  rule("", 'v', Aa1r); rule("", '^', Aa1r); rule("e", 'E', A1);
  exhaust_cpx = (MAXLEN / 2 + 1) * maxw[0] + (MAXLEN / 2) * maxw[2];                    /* 10669 */
  if (maxw[0] + (MAXLEN - 1) * maxw[1] > exhaust_cpx) exhaust_cpx = maxw[0] + (MAXLEN - 1) * maxw[1];
  if (target == 0) { fprintf(stderr, "%s: Target number cannot be zero.\n", argv0); exit(-1); }
}

static void show_symset(void) { /* 9969 */
  static const char *names[3] = {"Explicit values", "Functions of one argument", "Functions of two arguments"};
  int s, i;
  for (s = 0; s < 3; s++) {
    printf("%s:\n sym seft wght name description\n", names[s]);
    for (i = 0; i < 256; i++) if (seft_of[i] == 'a' + s && alwd[i]) {
      const char *d = terse_of[i] ? terse_of[i] : DESC[i];
      printf("  %c    %c   %2d   %-4s", i, seft_of[i], wgt[i], INFIX_NAME[i]);
      if (d) printf(" %s", d);
      printf("\n");
    }
    printf("\n");
  }
}

static void legend(void) { /* describe_symbols 7434 (postfix formats never mark atan2 as used) */
  int going = 1, left = 77, need_lf = 1, i;
  def_needed['A'] = 0;
  while (going) {
    going = 0;
    for (i = 0; i < 256; i++) if (def_needed[i] && !def_given[i]) { going = 1; if (need_lf) { printf("\n"); need_lf = 0; } }
    if (going) {
      int best = -1, bl = 0;
      for (i = 0; i < 256; i++) if (def_needed[i] && !def_given[i]) { int sl = (int)strlen(terse_of[i]); if (sl > bl && sl < left) { best = i; bl = sl; } }
      if (best >= 0) { printf("  %s", terse_of[best]); def_given[best] = 1; left -= bl + 2; } else { printf("\n"); left = 77; }
    }
  }
  if (left < 77) printf("\n");
}

static void print_end(int code) { /* 7574 */
  double tsec = now() - t0; long total = lhs_insert + rhs_insert;
  fflush(stdout);
  if (got_exact && restrict_sub < T_INT) printf("\n  NOTE: 'exact' match may result from floating-point roundoff error.\n");
  if (fmt <= 2) legend();
  if (total) {
    long blocks = (nodes_total + 909) / 910; double combos = (double)lhs_insert * (double)rhs_insert;
    printf("\n                         --LHS--      --RHS--      -Total-\n");
    printf("     max complexity: %11d  %11d  %11d\n", lmax, rmax, lmax + rmax);
    printf("          dead-ends: %11s  %11s  %11s  Time: %.3f\n", pfw(lhs_prune, 11), pfw(rhs_prune, 11), pfw(lhs_prune + rhs_prune, 11), tsec);
    printf("        expressions: %11s  %11s  %11s\n", pfw(lhs_gen, 11), pfw(rhs_gen, 11), pfw(gen_total, 11));
    printf("           distinct: %11ld  %11ld  %11ld", lhs_insert, rhs_insert, total);
    /* the reference's allocator: 72-byte nodes in 64 KiB blocks with an 8-byte link (4837-4900) */
    if (blocks * 64 > 1024) printf("  Memory: %ldKiB\n", blocks * 64); else printf("  Memory: %ld B\n", blocks * 8 + nodes_total * 72);
    printf("\n        Total equations tested:    %20s", pfw(combos, 20));
    if (combos > 9999) printf(" (%.4g)", combos);
    printf("\n");
  }
  fflush(stdout);
  if (code) exit(code);
}

int main(int argc, char **argv) {
  int i, going = 1, exhausted = 0; double lim_gen, lim_prune; char buf[64]; form F;
  t0 = now(); argv0 = argv[0];
  for (i = 0; i < 256; i++) { alwd[i] = MAXLEN; preempt_w[i] = -1; }
  alwd['W'] = 0;
  for (i = 0; i < NDEFS; i++) { INFIX_NAME[(unsigned char)DEFS[i].ch] = ""; }
  { static const char *nm[][3] = {{"1","1","integer"},{"f","phi",""},{"2","2","integer"},{"e","e",""},{"3","3","integer"},{"p","pi",""},
      {"4","4","integer"},{"5","5","integer"},{"6","6","integer"},{"7","7","integer"},{"8","8","integer"},{"9","9","integer"},
      {"x","x","the variable of the equation"},{"n","-","negate"},{"r","1/","reciprocal"},{"s","^2","square"},{"q","sqrt",""},
      {"l","ln",""},{"E","e^","natural exponent function"},{"S","sinpi","sinpi"},{"C","cospi","cospi"},{"T","tanpi","tanpi"},
      {"W","W","W"},{"+","+","add"},{"-","-","subtract"},{"*","*","multiply"},{"/","/","divide"},{"^","^","A ^ B = A to the power of B"},
      {"v","\"/",""},{"L","log_",""},{"A","atan2",""}};
    for (i = 0; i < (int)(sizeof nm / sizeof nm[0]); i++) { INFIX_NAME[(unsigned char)nm[i][0][0]] = nm[i][1]; DESC[(unsigned char)nm[i][0][0]] = nm[i][2]; } }
  if (argc > 1) parse(argc, argv);
  else { printf("%s: Please specify a target number.\n\n", argv0); help(); print_end(1); }
  if (anagram_opt) one_sided = 1; /* 12319 */
  init2();
  if (show_ss) {
    show_symset();
    printf("%s: %s", argv0, "Exiting now (to do an equation search, omit '-S' or include\n  symbol names, for example: 'ries 3.14159 -S123456789+/')\n");
    return 0;
  }
  if (got_target) targ_tag = target == floor(target) ? T_INT : (floor(target * 4.0) == target * 4.0 ? T_RAT : imin(T_RAT, restrict_sub)); /* 12043 */
  if (restrict_sub > tg_pi) tg_pi = imin(T_RAT, restrict_sub);
  if (restrict_sub > tg_phi) tg_phi = imin(T_RAT, restrict_sub);
  if (restrict_sub > tg_e) tg_e = imin(T_RAT, restrict_sub);
  if (!got_target) { printf("%s: Please specify a target number.\n\n", argv0); help(); print_end(1); }
  if (restrict_sub == T_INT) {                                                       /* 12181 */
    if (targ_tag != T_INT) { printf("ries: Replacing -i with -r because target isn't an integer.\n"); restrict_rat(); }
    else if (fabs(target) > 4503599627370496.0) { printf("ries: Replacing -i with -r because target is too large.\n"); restrict_rat(); }
  }
  if (fabs(target) * vanished_dx > 0.1) { vanished_dx = 0.1 / fabs(target); printf("Auto-setting --derivative-margin %g\n", vanished_dx); }
  else if (fabs(target) < 1.0e-8) { vanished_dx = 0.1 * fabs(target); printf("Auto-setting --derivative-margin %g\n", vanished_dx); }
  p_ovr = fabs(target) / (vanished_dx / 1.0e4); n_ovr = -p_ovr;                      /* 12249 */
  if (vanished_dx >= 1.0 / fabs(target)) {
    printf("%s: With this --derivative-margin option (%g) and\n  target value (%g), I cannot compute any expressions.  Either\n"
           "  use a smaller target value, or a derivative margin below %g\n\n", argv0, vanished_dx, target, 1.0 / fabs(target));
    print_end(-1);
  }
  if (fabs(target) / vanished_dx >= 1.0e15 && alwd['x'] > 1)
    printf("WARNING: Your target value is so large that RIES might report tautologies\n  like x/(x/3) = 3. To avoid this warning, use the option -Ox or give a\n"
           "  target value smaller than %g.\n\n", 10.0 / 1.0e-6);
  init_match_dist = k_max_match_dist >= 0 ? k_max_match_dist : fabs(target * (-k_max_match_dist));
  best_match = init_match_dist;
  if (fabs(target) > 1.0 || fabs(target) < 0.25) k_min_best_match = fabs(target) * 8.0 * precision_ulp;
  exec_x = target;
  if (fmt != 0) fmt = 3; /* infix formats are outside this oracle's scope */

  spfg(buf, 15, target);
  printf("\n   Your target value: T = %-21s  %29s\n\n", buf, "mrob.com/ries");          /* 12327 */
  lim_gen = 2000.0; if (tlevel >= -2.0) lim_gen *= pow(4.0, tlevel);
  lim_prune = lim_gen * 15.0;
  if (debug_y) printf("Will stop after generating searchmax=%g expressions\n", lim_gen);
  while (going) {                                                                    /* 12396 */
    double genf; int lhs = !((one_sided && lhs_insert >= 1) || (lhs_insert > rhs_insert)); /* 12425 */
    g_ne = 0; insert_count = 0; prune_count = 0;
    memset(&F, 0, sizeof F);
    if (!lhs) { rmin++; rmax++; genf = forms(&F, 0, 0, 0, rmin, rmax, 0);
      if (debug_y) printf("finished RHS from {%d} to {%d}: %g forms expanded, %g expressions.\n", rmin, rmax, genf, g_ne);
      rhs_gen += g_ne; rhs_insert += insert_count; rhs_prune += prune_count; }
    else { lmin++; lmax++; genf = forms(&F, 0, 0, 0, lmin, lmax, 1);
      if (debug_y) printf("finished LHS from {%d} to {%d}: %g forms expanded, %g expressions.\n\n", lmin, lmax, genf, g_ne);
      lhs_gen += g_ne; lhs_insert += insert_count; lhs_prune += prune_count; }
    gen_total += g_ne;
    if (lmax > exhaust_cpx && (one_sided || rmax > exhaust_cpx)) {
      if (num_matches == 0) {
        printf("exhaustion timeout (complexity LHS {%d}, RHS {%d}). Possible causes:\n  MAX_ELEN %d too small\n", lmax, rmax, MAXLEN);
        if (NOS_opt) printf("  too many symbols excluded via -N, -O or -S\n");
        if (anagram_opt) printf("  --numeric-anagram too restrictive\n");
        printf("  target too big or small for k_vanished_dx (use --derivative-margin)\n"); /* 12504: this restatement has no --derivative-margin option */
        exhausted = 1;
      }
      gen_total = lim_gen;
    }
    if (debug_y) { printf("gen_total %g (lhs %ld rhs %ld) searchmax %g gnm %ld\n", gen_total, lhs_insert, rhs_insert, lim_gen, (long)num_matches); bt_stats(); }
    if (num_matches == 0 && (debug_y || (now() - t0 > 2.0 && slow_msg))) {
      printf("  Still searching: %11s expr and %.3f sec so far...     \r", pfw(gen_total, 11), now() - t0); fflush(stdout); }
    if (lhs_prune + rhs_prune >= lim_prune) going = 0;
    if (gen_total >= lim_gen) going = 0;
  }
  if (debug_y) printf("Level %.g search complete (gen_total exceeded lim_gentotal).\n", levadj); /* 12550 */
  if (exhausted) { }
  else if (num_matches == 0) { printf("No solution was found (try using -l%d", (int)floor(levadj + 1.0)); if (init_match_dist < 0.001 * target) printf(" or a larger --max-match-distance"); printf(").  \n"); }
  else if (exact_exit) printf("               No 'exact' solution was found (try using -l%d).\n", (int)floor(levadj + 1.0));
  else printf("                  (for more results, use the option '-l%d')\n", (int)floor(levadj + 1.0));
  print_end(0);
  return 0;
}
