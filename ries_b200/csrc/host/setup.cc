// Command line, symbol table, pruning rules and thresholds: the host-side restatement of
// init1() / parse_args() / init2() / the parameter block of main()
// (/root/reference/public/ries.c:10228-10310, 11117-11964, 10314-10744, 12111-12317).
// The result is one RiesParams POD that is uploaded to the GPU's constant memory.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "host/ries_host.h"

namespace ries {

void Sink::write(const char* s, size_t n) const {
  if (fn) fn(ctx, s, n);
  else fwrite(s, 1, n, stdout);
}

void Sink::printf(const char* fmt, ...) const {
  char buf[2048];
  va_list ap;
  va_start(ap, fmt);
  int n = vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (n < 0) return;
  if ((size_t)n >= sizeof(buf)) n = (int)sizeof(buf) - 1;
  write(buf, (size_t)n);
}

/* init1(): every symbol allowed MAX_ELEN times except W (ries.c:10262-10266) */
Options::Options() {
  for (int i = 0; i < 256; i++) { sa_alwd[i] = RIES_MAX_ELEN; preempt_weight[i] = -1; sym_named[i] = false; }
  sa_alwd[(unsigned char)'W'] = 0;
}

void brief_help(const Sink& out) { /* ries.c:3755 */
  out.printf("%s",
"Usage:\n"
"  ries [options] target-value\n"
"\n"
"Target value is required and may be any number. Options include:\n"
"  -l3   Search further (``level-3 search'')\n"
"  -x    Show matched values as ``x = value'' rather than ``x = T + epsilon''\n"
"  -s    (Sort of) solve by transforming to ``x = ...'' form\n"
"  -N+-/ Do not use symbols +, - or /\n"
"  -S    Show list of all available symbols\n");
  out.printf("%s",
"  -l    Restrict to Liouvillian solutions\n"
"  -a    Restrict to algebraic solutions\n"
"  -c    Restrict to ``constructible'' (straightedge and compass) solutions\n"
"  -r    Restrict to rational solutions (-re for exact match)\n"
"  -i    Restrict to integer solutions (-ie for exact match)\n"
"\n"
"There are many more options; get the full manual at mrob.com/ries\n");
}

static void somesyms_set(Options& o, const char* s, short n) { /* ries.c:10193 */
  while (s && *s) {
    o.sa_alwd[(unsigned char)*s] = n;
    s++;
  }
}

static void allsyms_set(Options& o, short n, int include_x) { /* ries.c:10182 */
  for (int i = 0; i < 256; i++)
    if (include_x || ((char)i != 'x')) o.sa_alwd[i] = n;
}

static void set_restrict_rat(Options& o) { /* ries.c:10963 */
  o.restrict_subexpr = TYPE_RAT;
  somesyms_set(o, "pefqASCTlvLEW", 0);
  somesyms_set(o, "+-*/nr", RIES_MAX_ELEN);
  somesyms_set(o, "x", 1);
  somesyms_set(o, "s^", 0);
}

static void set_restrict_alg(Options& o, int restrict_trig) { /* ries.c:10978 */
  o.restrict_subexpr = TYPE_ALG;
  somesyms_set(o, "pelLEWA", 0);
  somesyms_set(o, "+-*/nrsqfx", RIES_MAX_ELEN);
  somesyms_set(o, "^v", RIES_MAX_ELEN);
  o.restrict_exponents = TYPE_RAT;
  somesyms_set(o, "SCT", RIES_MAX_ELEN);
  if (restrict_trig) {
    o.restrict_trig_args = TYPE_RAT;
    o.trig_scale = 3.14159265358979323846264338327950288419716; /* k_sincos_arg_scale = k_pi (static value, 3335) */
    o.trig_scale_default = true;
  }
}

/* parse_target ries.c:10758-10887 */
static int parse_target(const char* str, Options& o) {
  int nv = 0;
  o.mag_ulp = 0;
  size_t slen = strlen(str);
  std::string s;
  for (size_t i = 0; i < slen; i++) {
    if (str[i] == ' ') continue;
    s.push_back(str[i] == ',' ? '.' : str[i]);
  }
  int last_digit = -1;
  {
    char section = 'b';
    int num_dots = 0, init_sig = 1, going = 1;
    for (size_t i = 0; i < s.size() && going; i++) {
      char c = s[i];
      if (section == 'b') {
        if (c == '+' || c == '-') section = 'm';
        else if (init_sig && c == '0') { section = 'm'; last_digit = (int)i; }
        else if (c >= '0' && c <= '9') { section = 'm'; last_digit = (int)i; init_sig = 0; }
        else if (c == '.') { section = 'm'; num_dots++; }
        else { s[i] = '0'; going = 0; }
      } else if (section == 'm') {
        if (c == 'e' || c == 'E') section = 'e';
        else if (init_sig && c == '0') last_digit = (int)i;
        else if (c >= '0' && c <= '9') { last_digit = (int)i; init_sig = 0; }
        else if (num_dots == 0 && c == '.') num_dots++;
        else { s[i] = '0'; going = 0; }
      } else {
        going = 0;
      }
    }
  }
  if (last_digit >= 0) {
    double targ = 0, inc_ulp = 0;
    nv = sscanf(s.c_str(), "%lf", &targ);
    if (nv > 0) {
      if (s[last_digit] == '0') s[last_digit] = '1';
      else s[last_digit] = (char)(s[last_digit] - 1);
      if (sscanf(s.c_str(), "%lf", &inc_ulp) > 0) {
        o.mag_ulp = (inc_ulp > targ) ? (inc_ulp - targ) : (targ - inc_ulp);
      } else {
        nv = 0;
      }
      o.target = targ;
    }
    if (nv < 0) nv = 0;
  }
  return nv;
}

static void unknown_option(const Options& o, const char* arg, const Sink& out) {
  out.printf("%s: Unknown option '%s'\n\n", o.argv0.c_str(), arg);
  brief_help(out);
  throw ExitRequest{-1};
}

void parse_args(int argc, char** argv, Options& o, const Sink& out) {
  o.argv0 = argv[0] ? argv[0] : "ries";
  int i = 1;
  auto next_arg = [&]() -> const char* { return (i < argc) ? argv[i++] : (const char*)0; };
  while (i < argc) {
    const char* a = next_arg();
    if (a == 0 || a[0] == 0) break; /* a null string pops the argument stack (ries.c:11157) */
    double tmp_dbl;
    if (strcmp(a, "--ries-arguments-end") == 0) {
      break;
    } else if (strcmp(a, "-") == 0) {
      /* null option */
    } else if (strcmp(a, "--min-memory") == 0) { /* ries.c:11559 */
      /* steers only the reference's thrash watchdog (4768-4830), which stops a search that has become slow for
         lack of RAM: a matter of wall-clock time on the host, with no counterpart here.  The notice is printed. */
      const char* v = next_arg();
      double t;
      if (v && sscanf(v, "%lf", &t)) {
        out.printf("Memory-hogging safeguard disabled for the first %g bytes.\n", fabs(t));
      } else {
        out.printf("%s: --min-memory should be followed by a numeric argument.\n\n", o.argv0.c_str());
        brief_help(out);
        throw ExitRequest{-1};
      }
    } else if (strcmp(a, "--memory-abort-threshold") == 0) { /* ries.c:11459: same watchdog */
      const char* v = next_arg();
      double t;
      if (v && sscanf(v, "%lf", &t)) {
        if (t > 1.0) {
          out.printf("Memory slowness abort ratio: %g.\n", t);
        } else {
          out.printf("%s: --memory-abort-threshold should be at least 1.0, and values less than about 1.5 are"
                     " unlikely to be of much use.\n\n", o.argv0.c_str());
          brief_help(out);
          throw ExitRequest{-1};
        }
      } else {
        out.printf("%s: --memory-abort-threshold should be followed by a numeric\n"
                   "argument larger than 1.0 (larger than 1.5 is recommended).\n\n", o.argv0.c_str());
        brief_help(out);
        throw ExitRequest{-1};
      }
    } else if (strcmp(a, "--canon-reduction") == 0) { /* ries.c:11255 */
      if (i < argc && argv[i] && argv[i][0] && argv[i][0] != '-') {
        const char* v = next_arg();
        o.canon_ops = 0;
        for (; *v; v++) {
          if (*v == 'n') o.canon_ops |= 1;       /* CANONVAL_NEGATE */
          else if (*v == 'r') o.canon_ops |= 2;  /* CANONVAL_RECIPROCAL */
          else if (*v == '5') o.canon_ops |= 4;  /* CANONVAL_DIV2 */
          else if (*v == '2') o.canon_ops |= 8;  /* CANONVAL_MUL2 */
        }
      } else {
        out.printf("%s: --canon-reduction requires a set of reduction types, e.g. 'nr25'\n\n", o.argv0.c_str());
        brief_help(out);
        throw ExitRequest{-1};
      }
    } else if (strcmp(a, "--canon-simplify") == 0) {
      o.no_cv_simplify = false;
    } else if (strcmp(a, "--no-canon-simplify") == 0) {
      o.no_cv_simplify = true;
    } else if (strcmp(a, "--max-match-distance") == 0) { /* ries.c:11395 */
      const char* v = next_arg();
      double t;
      if (v && sscanf(v, "%lf", &t) > 0) {
        o.k_max_match_dist = t;
        o.match_all_digits = false; /* cannot use both options together (11400) */
        if (t < 0) out.printf("First match must be closer than %g times your target value.\n", -t);
        else if (t == 0) out.printf("Only give an 'exact' match (if any) then exit.\n");
        else out.printf("First match must be closer than %g.\n", t);
      } else {
        out.printf("%s: --max-match-distance should be followed by a numeric argument.\n\n",
                   o.argv0.c_str());
        brief_help(out);
        throw ExitRequest{-1};
      }
    } else if ((strcmp(a, "--max-matches") == 0) || (strncmp(a, "-n", 2) == 0)) { /* 11418 */
      const char* v = (a[1] == 'n') ? a + 2 : next_arg();
      double t;
      if (v && sscanf(v, "%lf", &t) > 0) {
        if (t >= 1) {
          o.max_matches = t;
        } else {
          out.printf("%s: -n or --max-matches argument should be 1 or more.\n"
                     "examples:  -n7  or  --max-matches 7\n\n", o.argv0.c_str());
          brief_help(out);
          throw ExitRequest{-1};
        }
      } else {
        out.printf("%s: --max-matches should be followed by a numeric argument.\n"
                   "examples:  -n7  or  --max-matches 7\n\n", o.argv0.c_str());
        brief_help(out);
        throw ExitRequest{-1};
      }
    } else if (strcmp(a, "--no-slow-messages") == 0) {
      o.allow_slow_message = false;
    } else if (strcmp(a, "--derivative-margin") == 0) { /* ries.c:11284 */
      const char* v = next_arg();
      double t;
      if (v && sscanf(v, "%lf", &t) > 0) {
        if (t < 0) {
          out.printf("%s: --derivative-margin must be positive (I got %g).\n\n", o.argv0.c_str(), t);
          brief_help(out);
          throw ExitRequest{-1};
        } else if (t > 1.0e-4) {
          out.printf("%s: --derivative-margin value should be at most 1.0e-4.\n\n", o.argv0.c_str());
          brief_help(out);
          throw ExitRequest{-1};
        } else {
          o.derivative_margin = t;
          out.printf("Allowing d/dx to be as small as value times %g.\n", t);
        }
      } else {
        out.printf("%s: --derivative-margin should be followed by a numeric argument.\n\n", o.argv0.c_str());
        brief_help(out);
        throw ExitRequest{-1};
      }
    } else if (strcmp(a, "--explicit-multiply") == 0) { /* 11338 */
      o.explicit_multiply = true;
    } else if ((strcmp(a, "--match-all-digits") == 0) || (strcmp(a, "--mad") == 0)) { /* 11366 */
      o.match_all_digits = true;
      o.k_max_match_dist = -0.01;
    } else if (strcmp(a, "--max-equate-value") == 0) { /* 11373 */
      const char* v = next_arg();
      double t;
      if (v && sscanf(v, "%lf", &t) > 0) {
        if (t <= o.min_equ_val) {
          out.printf("%s: --max-equate-value argument cannot be greater than --min-equate-value argument.\n",
                     o.argv0.c_str());
          brief_help(out);
          throw ExitRequest{-1};
        }
        o.max_equ_val = t;
        out.printf("Equations will have both sides at most %s\n", spfg(15, t).c_str());
      } else {
        out.printf("%s: --max-equate-value should be followed by a numeric argument.\n\n", o.argv0.c_str());
        brief_help(out);
        throw ExitRequest{-1};
      }
    } else if (strcmp(a, "--min-equate-value") == 0) { /* 11484 */
      const char* v = next_arg();
      double t;
      if (v && sscanf(v, "%lf", &t) > 0) {
        if (t >= o.max_equ_val) {
          out.printf("%s: --min-equate-value argument cannot be greater than --max-equate-value argument.\n",
                     o.argv0.c_str());
          brief_help(out);
          throw ExitRequest{-1};
        }
        o.min_equ_val = t;
        out.printf("Equations will have both sides at least %s\n", spfg(15, t).c_str());
      } else {
        out.printf("%s: --min-equate-value should be followed by a numeric argument.\n\n", o.argv0.c_str());
        brief_help(out);
        throw ExitRequest{-1};
      }
    } else if (strcmp(a, "--min-match-distance") == 0) { /* 11506 */
      const char* v = next_arg();
      double t;
      if (v && sscanf(v, "%lf", &t) > 0) {
        o.min_matchsize = t;
        if (t == 0.0) {
          out.printf("Will exit if an 'exact' match is found.\n");
          o.exact_exit = true;
        } else if (t < -0.1) {
          out.printf("%s: --min-match-distance argument can be negative but not less than -0.1\n\n",
                     o.argv0.c_str());
          brief_help(out);
          throw ExitRequest{-1};
        } else if (t < 0.0) {
          out.printf("Using a minimum match distance of %g times your target value.\n", t);
        } else {
          out.printf("Using minimum match distance: %g\n", t);
        }
      } else {
        out.printf("%s: --min-match-distance should be followed by a numeric argument.\n\n", o.argv0.c_str());
        brief_help(out);
        throw ExitRequest{-1};
      }
    } else if (strcmp(a, "--max-trig-cycles") == 0) { /* 11536 */
      const char* v = next_arg();
      double t;
      if (v && sscanf(v, "%lf", &t) > 0) {
        if ((t > 0.0) && (t <= 100.0)) {
          o.sincos_max_arg = t;
          out.printf("Arguments of trig functions will be restricted to %s cycles on either side of zero.\n",
                     spfg(15, t * 0.5).c_str());
        } else {
          out.printf("%s: --max-trig-cycles must be between 0.0 and 100.0\n\n", o.argv0.c_str());
          brief_help(out);
          throw ExitRequest{-1};
        }
      } else {
        out.printf("%s: --max-trig-cycles should be followed by a numeric argument.\n\n", o.argv0.c_str());
        brief_help(out);
        throw ExitRequest{-1};
      }
    } else if (strcmp(a, "--trig-argument-scale") == 0) { /* 11731 */
      const char* v = next_arg();
      double t;
      if (v && sscanf(v, "%lf", &t) > 0) {
        if ((t >= 0.0) && (t < 100.0)) {
          o.trig_scale = t;
          o.trig_scale_default = (t == 3.14159265358979323846264338327950288419716); /* k_pi's static value (3335) */
          out.printf("Argument of trig functions will be scaled by %s\n", spfg(15, t).c_str());
        } else {
          out.printf("%s: --trig-argument-scale must be between 0.0 and 100.0\n\n", o.argv0.c_str());
          brief_help(out);
          throw ExitRequest{-1};
        }
      } else {
        out.printf("%s: --trig-argument-scale should be followed by a numeric argument.\n\n", o.argv0.c_str());
        brief_help(out);
        throw ExitRequest{-1};
      }
    } else if (strcmp(a, "--relative-roots") == 0) { /* 11609 */
      if (o.match_all_digits) out.printf("Note: Ignoring '%s' because '--match-all-digits' is set.\n", a);
      else o.relative_x = true;
    } else if (strcmp(a, "--significance-loss-margin") == 0) { /* 11622 */
      const char* v = next_arg();
      double t;
      if (v && sscanf(v, "%lf", &t) > 0) {
        if ((t >= 0.0) && (t < 100.0)) {
          o.sig_loss = pow(10.0, -t);
          out.printf("Using significance loss margin: %g\n", o.sig_loss);
        } else {
          out.printf("%s: --significance-loss-margin must be between 0.0 and 100.0\n\n", o.argv0.c_str());
          brief_help(out);
          throw ExitRequest{-1};
        }
      } else {
        out.printf("%s: --significance-loss-margin should be followed by a numeric argument.\n\n",
                   o.argv0.c_str());
        brief_help(out);
        throw ExitRequest{-1};
      }
    } else if (strcmp(a, "--no-refinement") == 0) { /* ries.c:11577 */
      o.refinement = false;
    } else if (strcmp(a, "--one-sided") == 0) { /* ries.c:11600 */
      o.one_sided = true;
    } else if (strcmp(a, "--numeric-anagram") == 0) { /* ries.c:11586, set_anagram 10202 */
      if (i < argc && argv[i] && argv[i][0] && argv[i][0] != '-') {
        const char* ana = next_arg();
        const char* ana_syms = "123456789efprs";
        for (int k = 0; ana_syms[k]; k++) o.sa_alwd[(unsigned char)ana_syms[k]] = 0;
        for (int k = 0; ana[k]; k++)
          if (strchr(ana_syms, ana[k])) o.sa_alwd[(unsigned char)ana[k]]++;
        o.anagram = true;
      } else {
        out.printf("%s: --numeric-anagram should be followed by a string of digits.\n\n", o.argv0.c_str());
        brief_help(out);
        throw ExitRequest{-1};
      }
    } else if (strcmp(a, "--symbol-names") == 0) { /* ries.c:11648 */
      if (i < argc && argv[i] && argv[i][0] && argv[i][0] != '-') {
        char space_sym = ' ';
        while (i < argc && argv[i] && argv[i][0] && argv[i][0] != '-') {
          const char* arg = next_arg();
          if ((arg[0] == ':') && arg[1] && (arg[2] == ':') && (arg[3] == arg[1]) && (arg[4] == 0) && (space_sym == ' ')) {
            space_sym = arg[1]; /* :c:c defines the character that stands for a blank */
          } else if ((arg[0] == ':') && arg[1] && (arg[2] == ':')) {
            if (strlen(arg + 3) <= 24) {
              std::string nm(arg + 3);
              for (size_t k = 0; k < nm.size(); k++)
                if (nm[k] == space_sym) nm[k] = ' ';
              o.sym_name[(unsigned char)arg[1]] = nm;
              o.sym_named[(unsigned char)arg[1]] = true;
            } else {
              out.printf("%s: Symbol name can be at most %d characters\n(I got '%s')\n", o.argv0.c_str(), 24, arg + 3);
              throw ExitRequest{-1};
            }
          } else {
            out.printf("%s: --symbol-names argument syntax is :<sym>:name, for example\n"
                       "  :-:deme   to set name of '-' to 'deme'\n"
                       "  Instead I got '%s'\n", o.argv0.c_str(), arg);
            throw ExitRequest{-1};
          }
        }
      } else {
        out.printf("%s: --symbol-names should be followed by one or more tuples\n"
                   "  of the form <sym>:name, for example:\n"
                   "  :-:deme   to set name of '-' to 'deme'\n", o.argv0.c_str());
        throw ExitRequest{-1};
      }
    } else if (strcmp(a, "--symbol-weights") == 0) { /* ries.c:11689 */
      if (i < argc && argv[i] && argv[i][0] && argv[i][0] != '-') {
        while (i < argc && argv[i] && argv[i][0] && argv[i][0] != '-') {
          const char* arg = next_arg();
          char tmp[20];
          strncpy(tmp, arg, 19);
          tmp[19] = 0;
          int k = 0;
          while (tmp[k] == '.' || (tmp[k] >= '0' && tmp[k] <= '9')) k++;
          if (tmp[k] == ':' && tmp[k + 1]) {
            tmp[k++] = 0;
            double w = strtod(tmp, 0);
            if (w <= 0.0) {
              out.printf("%s: Symbol weight may not be %f or less.\n", o.argv0.c_str(), 0.0);
              throw ExitRequest{-1};
            }
            if (w > 100.0) w = 100.0;
            o.preempt_weight[(unsigned char)tmp[k]] = (short)floor(w + 0.5);
          } else {
            out.printf("%s: --symbol-weights argument syntax is NUMBER:<sym>, for example\n"
                       "  12:^   to set weight of '^' to 12\n"
                       "  Instead I got '%s'\n", o.argv0.c_str(), arg);
            throw ExitRequest{-1};
          }
        }
      } else {
        out.printf("%s: --symbol-weights should be followed by one or more tuples\n"
                   "  of the form NUMBER:<sym>, for example 12:^   to set weight of '^' to 12\n",
                   o.argv0.c_str());
        throw ExitRequest{-1};
      }
    } else if (strcmp(a, "--any-exponents") == 0) {
      o.restrict_exponents = TYPE_NONE;
    } else if (strcmp(a, "--any-subexpressions") == 0) {
      o.restrict_subexpr = TYPE_NONE;
    } else if (strcmp(a, "--any-trig-args") == 0) {
      o.restrict_trig_args = TYPE_NONE;
    } else if (strcmp(a, "--rational-exponents") == 0) {
      o.restrict_exponents = TYPE_RAT;
    } else if (strcmp(a, "--rational-trig-args") == 0) {
      o.restrict_trig_args = TYPE_RAT;
    } else if (strcmp(a, "--no-solve-for-x") == 0) {
      o.solve_for_x = false;
    } else if ((strncmp(a, "-s", 2) == 0) || (strcmp(a, "--try-solve-for-x") == 0)) {
      o.solve_for_x = true;
    } else if ((strcmp(a, "--wide-output") == 0) || (strcmp(a, "--wide") == 0)) {
      o.wide_output = true;
      o.relative_x = true;
    } else if ((strncmp(a, "-a", 2) == 0) || (strcmp(a, "--algebraic-subexpressions") == 0)) {
      set_restrict_alg(o, 1);
      if (a[2] == 'e') o.exact_exit = true;
    } else if ((strncmp(a, "-c", 2) == 0) || (strcmp(a, "--constructible-subexpressions") == 0)) {
      o.restrict_subexpr = TYPE_CONS;
      if (a[2] == 'e') o.exact_exit = true;
      somesyms_set(o, "peASCTl^vLEW", 0);
      somesyms_set(o, "+-*/nrsqf", RIES_MAX_ELEN);
      somesyms_set(o, "x", 1);
    } else if (strncmp(a, "-D", 2) == 0) {
      for (const char* d = a + 2; *d; d++)
        if (*d == 'y') o.debug_y = 1; /* only the per-pass trace is mirrored */
    } else if (strncmp(a, "-E", 2) == 0) {
      o.NOS_options = true;
      somesyms_set(o, a + 2, RIES_MAX_ELEN);
    } else if (strncmp(a, "-F", 2) == 0) {
      const char* v = a + 2;
      if (*v == 0) {
        o.out_format = OF_FORTH;
      } else if (sscanf(v, "%d", &o.out_format) <= 0) {
        out.printf("%s: -F parameter requires a number, e.g. '-F0'\n\n", o.argv0.c_str());
        brief_help(out);
        throw ExitRequest{-1};
      }
    } else if ((strncmp(a, "-i", 2) == 0) || (strcmp(a, "--integer-subexpressions") == 0)) {
      o.restrict_subexpr = TYPE_INT;
      somesyms_set(o, "pefqASCTlvLEW", 0);
      if (a[2] == 'e') o.exact_exit = true;
    } else if (sscanf(a, "-l%lf", &tmp_dbl) > 0) {
      o.levadj = tmp_dbl;
      o.tlevel = 2.0 + o.levadj;
    } else if ((strcmp(a, "-el") == 0) || (strcmp(a, "-l") == 0) ||
               (strcmp(a, "--liouvillian-subexpressions") == 0) || (strcmp(a, "-le") == 0)) {
      set_restrict_alg(o, 0);
      somesyms_set(o, "eplEL", RIES_MAX_ELEN);
      o.restrict_subexpr = TYPE_TRAN;
      o.restrict_exponents = TYPE_TCEL;
      o.restrict_trig_args = TYPE_TCEL;
      if (a[2] == 'e') o.exact_exit = true;
    } else if (strncmp(a, "-l", 2) == 0) {
      out.printf("%s: -l parameter requires a number, e.g. '-l3'\n\n", o.argv0.c_str());
      brief_help(out);
      throw ExitRequest{-1};
    } else if (strncmp(a, "-N", 2) == 0) {
      o.NOS_options = true;
      somesyms_set(o, a + 2, 0);
    } else if (strncmp(a, "-O", 2) == 0) {
      o.NOS_options = true;
      somesyms_set(o, a + 2, 1);
    } else if ((strncmp(a, "-r", 2) == 0) || (strcmp(a, "--rational-subexpressions") == 0)) {
      set_restrict_rat(o);
      if (a[2] == 'e') o.exact_exit = true;
    } else if (strncmp(a, "-S", 2) == 0) {
      const char* v = a + 2;
      if (*v == 0) {
        o.show_ss = true;
      } else {
        o.S_option = true;
        o.NOS_options = true;
        allsyms_set(o, 0, 0);
        somesyms_set(o, v, RIES_MAX_ELEN);
      }
    } else if ((strcmp(a, "-x") == 0) || (strcmp(a, "--absolute-roots") == 0)) {
      if (o.wide_output) out.printf("Note: '%s' option with '--wide' is redundant.\n", a);
      else o.relative_x = false;
    } else if (((a[0] >= '0') && (a[0] <= '9')) || (a[0] == '-') || (a[0] == '.')) {
      if (parse_target(a, o) == 1) o.got_target = true;
      else unknown_option(o, a, out);
    } else {
      unknown_option(o, a, out);
    }
  }
}

namespace {
struct SymDef {
  char ch;
  const char* forth;
  const char* infix;
  char seft;
  short weight;
  const char* desc;
};
/* add_symbol calls, ries.c:10375-10468 (weights are offsets from weight_base = 10) */
const SymDef kSymDefs[] = {
    {'1', 0, "1", 'a', 0, "integer"},
    {'f', "phi", "phi", 'a', 8, ""},
    {'2', 0, "2", 'a', 3, "integer"},
    {'e', "e", "e", 'a', 6, ""},
    {'3', 0, "3", 'a', 5, "integer"},
    {'p', "pi", "pi", 'a', 4, ""},
    {'4', 0, "4", 'a', 6, "integer"},
    {'5', 0, "5", 'a', 7, "integer"},
    {'6', 0, "6", 'a', 8, "integer"},
    {'7', 0, "7", 'a', 8, "integer"},
    {'8', 0, "8", 'a', 9, "integer"},
    {'9', 0, "9", 'a', 9, "integer"},
    {'x', 0, "x", 'a', 5, "the variable of the equation"},
    {'n', "neg", "-", 'b', -3, "negate"},
    {'r', "recip", "1/", 'b', -3, "reciprocal"},
    {'s', "dup*", "^2", 'b', -1, "square"},
    {'q', "sqrt", "sqrt", 'b', -1, ""},
    {'l', "ln", "ln", 'b', 3, ""},
    {'E', "exp", "e^", 'b', 3, "natural exponent function"},
    {'S', "sinpi", "sinpi", 'b', 3, "sinpi"},
    {'C', "cospi", "cospi", 'b', 3, "cospi"},
    {'T', "tanpi", "tanpi", 'b', 6, "tanpi"},
    {'W', "W", "W", 'b', 5, "W"},
    {'+', "+", "+", 'c', -6, "add"},
    {'-', "-", "-", 'c', -5, "subtract"},
    {'*', "*", "*", 'c', -6, "multiply"},
    {'/', "/", "/", 'c', -5, "divide"},
    {'^', "**", "^", 'c', -4, "A ^ B = A to the power of B"},
    {'v', "root", "\"/", 'c', -3, ""},
    {'L', "logN", "log_", 'c', -1, ""},
    {'A', "atan2", "atan2", 'c', -1, ""},
};
const int kNumSymDefs = (int)(sizeof(kSymDefs) / sizeof(kSymDefs[0]));

struct RuleDef {
  const char* req;
  char sym;
  int mask;
};
/* add_rule calls, ries.c:10559-10665, in source order (order is irrelevant: masks are OR-ed) */
const RuleDef kRules[] = {
    {"", 'x', AM_RHS},      {"nr", 'n', AM_r},      {"", 'n', AM_n},        {"1", 'r', AM_1},
    {"", 'r', AM_r},        {"1", 's', AM_1},       {"4", 's', AM_2},       {"s", 's', AM_n},
    {"rs", 's', AM_r},      {"4^", 's', AM_sq},     {"1", 'q', AM_1},       {"rq", 'q', AM_r},
    {"4v", 'q', AM_sq},     {"", 'l', AM_1},        {"ln", 'l', AM_r},      {"", 'l', AM_E},
    {"", 'E', AM_l},        {"Er", 'E', AM_n},      {"Sn", 'S', AM_n},      {"C", 'C', AM_n},
    {"2*", '+', AM_KK},     {"*23456789", '+', AM_55}, {"+", '+', AM_jK},   {"-", '+', AM_n},
    {"", '-', AM_KK},       {"1234n", '-', AM_55},  {"12345678n", '-', AM_jK}, {"+", '-', AM_n},
    {"s", '*', AM_KK},      {"*", '*', AM_jK},      {"*", '*', AM_1},       {"*n", '*', AM_n},
    {"/", '*', AM_r},       {"1", '/', AM_KK},      {"r", '/', AM_1K},      {"", '/', AM_1},
    {"/n", '/', AM_n},      {"*", '/', AM_r},       {"", '^', AM_1},        {"s", '^', AM_2},
    {"^r", '^', AM_n},      {"1", '^', AM_1K},      {"v", '^', AM_r},       {"", 'v', AM_1},
    {"q", 'v', AM_2},       {"vr", 'v', AM_n},      {"1", 'v', AM_1K},      {"^", 'v', AM_r},
    {"1", 'L', AM_KK},      {"", 'L', AM_1},        {"Ln", 'L', AM_r},      {"", 'L', AM_1K},
    /* (the two k_sincos_arg_scale == 1.0 rules, 10623-10628, never fire: the scale is pi) */
    {"", '/', AM_KxK},      {"", '-', AM_KpK},      {"", '+', AM_plus},     {"", '*', AM_mul},
    {"^", '^', AM_mul},     {"", 'q', AM_pow},      {"es", 'E', AM_2},      {"E", '^', AM_a1_e},
    {"", '*', AM_a1_1},     {"r", '/', AM_a1_1},    {"", 'v', AM_a1_r},     {"", '^', AM_a1_r},
    {"e", 'E', AM_1},
};
const int kNumRules = (int)(sizeof(kRules) / sizeof(kRules[0]));

int amkey_of(char c) { /* define_amkeys ries.c:9822 */
  switch (c) {
    case '1': return AM_1;
    case '2': return AM_2;
    case 'n': return AM_n;
    case 'r': return AM_r;
    case 's': case 'q': return AM_sq;
    case 'l': return AM_l;
    case 'E': return AM_E;
    case 'p': return AM_pi;
    case '+': return AM_plus;
    case '*': return AM_mul;
    case '^': return AM_pow;
    default: return 0;
  }
}

/* guess_valtype ries.c:12043 */
int guess_valtype(double v, int restrict_subexpr) {
  if (v == floor(v)) return TYPE_INT;
  if (floor(v * 4.0) == (v * 4.0)) return TYPE_RAT;
  return (TYPE_RAT < restrict_subexpr) ? TYPE_RAT : restrict_subexpr;
}
}  // namespace

/* Symbol definitions for the legend / -S table: add_symbol's def_terse / def_normal arguments
   (ries.c:10375-10468). `normal` selects the -F2 wording. */
const char* sym_defn(unsigned char ch, bool normal, bool pi_names) {
  if (!pi_names) { /* add_symbol(..., "S(x) = sine", 0, "sine") etc. (10436-10441): no -F2 definition */
    if (ch == 'S') return normal ? 0 : "S(x) = sine";
    if (ch == 'C') return normal ? 0 : "C(x) = cosine";
    if (ch == 'T') return normal ? 0 : "T(x) = tangent";
  }
  switch (ch) {
    case 'f': return normal ? "phi = the golden ratio, (1+sqrt(5))/2" : "f = phi, the golden ratio, (1+sqrt(5))/2";
    case 'e': return "e = base of natural logarithms, 2.71828...";
    case 'p': return normal ? "pi = 3.14159..." : "p = pi, 3.14159...";
    case 'n': return normal ? 0 : "n = negative";
    case 'q': return normal ? "sqrt(x) = square root" : "q = square root";
    case 'l': return normal ? "ln(x) = natural logarithm or log base e" : "l = ln, natural logarithm or log base e";
    case 'S': return normal ? "sinpi(X) = sin(pi * x)" : "S(x) = sinpi(x) = sin(pi * x)";
    case 'C': return normal ? "cospi(X) = cos(pi * x)" : "C(x) = cospi(x) = cos(pi * x)";
    case 'T': return normal ? "tanpi(X) = tan(pi * x)" : "T(x) = tanpi(x) = tan(pi * x)";
    case 'W': return "W(x) = LambertW(x) = inverse(x=w*e^w)";
    case 'v': return normal ? "A\"/B = Ath root of B" : "A v B = Ath root of B";
    case 'L': return normal ? "log_A(B) = logarithm to base A of B = ln(B) / ln(A)"
                            : "A L B = logarithm to base A of B = ln(B) / ln(A)";
    case 'A': return normal ? "atan2(s,c) = Angle of ray from origin through point (c,s)" : "s A c = arctan2(s,c)";
    default: return 0;
  }
}

void show_symset(const Options& o, const Setup& S, const Sink& out) { /* ries.c:9969 */
  static const char* seft_names[3] = {"Explicit values", "Functions of one argument",
                                      "Functions of two arguments"};
  for (int seft = 'a'; seft <= 'c'; seft++) {
    out.printf("%s:\n", seft_names[seft - 'a']);
    out.printf(" sym seft wght name description\n");
    for (int i = 0; i < 256; i++) { /* ASCII order, like the reference's scan of sym_attrs[] */
      int g = S.idx_of[i];
      if (g < 0) continue;
      if (S.info[g].seft != seft || !S.info[g].alwd) continue;
      out.printf("  %c    %c   %2d   %-4s", i, S.info[g].seft, S.info[g].wgt, S.info[g].name_infix);
      const char* def = sym_defn((unsigned char)i, o.out_format == OF_NORMAL, S.trig_pi_names);
      if (def == 0) {
        for (int k = 0; k < kNumSymDefs; k++)
          if (kSymDefs[k].ch == (char)i) def = kSymDefs[k].desc;
        if (!S.trig_pi_names && (i == 'S' || i == 'C' || i == 'T'))
          def = (i == 'S') ? "sine" : (i == 'C') ? "cosine" : "tangent";
      }
      if (def) out.printf(" %s", def);
      out.printf("\n");
    }
    out.printf("\n");
  }
}

void build_setup(Options& o, Setup& S, const Sink& out) {
  RiesParams& P = S.P;
  memset(&P, 0, sizeof(P));
  for (int i = 0; i < 256; i++) S.idx_of[i] = -1;

  /* init_formats() for IEEE binary64 (ries.c:7699-7789): 53 mantissa bits */
  S.k_nominal_digits = 17;
  S.k_usable_digits = 15;
  double k_ulp = ldexp(1.0, -54); /* 0.5 * 0.5^53 : loop 7760-7763 with k_mantissa_bits = 53 */
  S.k_precision_ulp = 2.0 * k_ulp;            /* 10067 */
  S.k_min_best_match = 9.0072 * S.k_precision_ulp; /* 10070 */

  /* init_numerics(): the series land on the correctly rounded doubles, recomputed here the same
     way so the bits are the reference's by construction (10084-10148). */
  P.k_phi = (sqrt(5.0) + 1.0) / 2.0;
  {
    double i = 31.0, e = 1.0 / i;
    while (i > 1.0) { i -= 1.0; e = 1.0 + e / i; }
    P.k_e = e;
  }
  {
    double i = 103.0, pi = 1.57;
    while (i > 1.0) { i -= 1.0; pi = 1.0 + (i / (1.0 + (2.0 * i))) * pi; }
    P.k_pi = 2.0 * pi;
  }
  /* init_numerics 10157-10161: an unset (<= 0) scale becomes pi and "default"; otherwise the flag is what
     parse_args computed */
  if (o.trig_scale <= 0) { P.sincos_scale = P.k_pi; P.trig_scale_default = 1; }
  else { P.sincos_scale = o.trig_scale; P.trig_scale_default = o.trig_scale_default ? 1 : 0; }
  S.trig_pi_names = (P.sincos_scale == P.k_pi);
  P.trig_max_arg = P.k_pi * o.sincos_max_arg; /* k_pi * k_sincos_max_arg (5503) */
  P.k_sin_clip = 0.99999;
  P.k_eXlim = 690.0;
  P.k_sig_loss = o.sig_loss;
  P.k_newton_settled = 1.0e-15;
  P.min_equ_val = o.min_equ_val;
  P.max_equ_val = o.max_equ_val;
  P.tg_phi = TYPE_CONS;
  P.tg_e = TYPE_TCEL;
  P.tg_pi = TYPE_TCEL;
  P.no_cv_simplify = o.no_cv_simplify ? 1 : 0;
  P.canon_ops = o.canon_ops;

  /* -Nx is prohibited; -Ox means "x once, LHS only" (10336-10340) */
  if (o.sa_alwd[(unsigned char)'x'] == 1) S.x_lhs_only = true;
  else { S.x_lhs_only = false; o.sa_alwd[(unsigned char)'x'] = RIES_MAX_ELEN; }

  if (o.sa_alwd[(unsigned char)'W']) { /* DUMMY_LAMBERT 10342-10349 */
    out.printf("%s: The Lambert W function 'W' requires the stand-alone maths library.\n",
               o.argv0.c_str());
    brief_help(out);
    throw ExitRequest{1};
  }

  /* the infix punctuation, which --symbol-names can rename too (10481-10485) */
  for (int c = 0; c < 256; c++) { S.any_name[c] = 0; S.any_forth[c] = 0; }
  S.any_name[(unsigned char)'('] = "("; S.any_name[(unsigned char)','] = ",";
  S.any_name[(unsigned char)')'] = ")"; S.any_name[(unsigned char)'='] = "=";
  {
    const char* punct = "(,)=";
    for (const char* q = punct; *q; q++) {
      unsigned char c = (unsigned char)*q;
      if (o.sym_named[c] && !o.sym_name[c].empty()) {
        S.name_store[c] = o.sym_name[c];
        S.any_name[c] = S.name_store[c].c_str();
      }
    }
  }

  /* symbol table */
  P.nsym = kNumSymDefs;
  for (int k = 0; k < 8; k++) P.canon_sym[k] = -1;
  for (int g = 0; g < kNumSymDefs; g++) {
    const SymDef& d = kSymDefs[g];
    int cls = d.seft - 'a';
    { const char* cs = "nr2/*"; const char* at = strchr(cs, d.ch); if (d.ch && at) P.canon_sym[at - cs] = (signed char)g; }
    P.sym_char[g] = (unsigned char)d.ch;
    P.sym_cls[g] = (unsigned char)cls;
    /* base complexity 10 + per-symbol weight, unless --symbol-weights overrides it (ries.c:9872) */
    P.sym_wgt[g] = (o.preempt_weight[(unsigned char)d.ch] >= 0) ? o.preempt_weight[(unsigned char)d.ch]
                                                                  : (short)(d.weight + 10);
    P.sym_alwd[g] = o.sa_alwd[(unsigned char)d.ch];
    P.sym_amkey[g] = amkey_of(d.ch);
    P.sym_mask[g] = 0;
    S.idx_of[(unsigned char)d.ch] = g;
    S.info[g].ch = (unsigned char)d.ch;
    S.info[g].seft = d.seft;
    S.info[g].wgt = P.sym_wgt[g];
    S.info[g].alwd = P.sym_alwd[g];
    S.info[g].name_forth = d.forth;
    S.info[g].name_infix = d.infix;
    if (!S.trig_pi_names && (d.ch == 'S' || d.ch == 'C' || d.ch == 'T')) /* 10436-10441 */
      S.info[g].name_forth = S.info[g].name_infix = (d.ch == 'S') ? "sin" : (d.ch == 'C') ? "cos" : "tan";
    if (o.sym_named[(unsigned char)d.ch]) {
      /* --symbol-names set sa_name = name_forth = the text; add_symbol keeps a non-empty user name and
         otherwise installs its defaults (9855-9862) */
      S.name_store[(unsigned char)d.ch] = o.sym_name[(unsigned char)d.ch];
      const char* given = S.name_store[(unsigned char)d.ch].c_str();
      if (given[0]) { S.info[g].name_infix = given; S.info[g].name_forth = given; }
      else if (S.info[g].name_forth == 0) S.info[g].name_forth = given; /* the empty string stays */
    }
    if (P.sym_alwd[g] != 0) {
      if (P.cls_n[cls] >= RIES_SLOTS) {
        out.printf("add_symbol: Too many seft '%c' symbols.\n", d.seft);
        throw ExitRequest{-1};
      }
      P.trial[cls][P.cls_n[cls]++] = (unsigned char)g;
      if (P.sym_alwd[g] < RIES_MAX_ELEN) P.any_limited = 1;
    }
  }

  if (P.cls_n[0] < 2) { /* 10493 */
    out.printf("%s: You must allow at least one constant symbol.\n", o.argv0.c_str());
    if (o.S_option) out.printf("  (Note that the -S option disables all other symbols)\n");
    brief_help(out);
    throw ExitRequest{1};
  }
  if (P.cls_n[1] == 0 && P.cls_n[2] == 0) { /* 10503 */
    out.printf("%s: You must allow at least one operator symbol.\n", o.argv0.c_str());
    brief_help(out);
    throw ExitRequest{1};
  }

  /* setup_abc_mmw (9915-9959) */
  short mn[3] = {100, 100, 100}, mx[3] = {0, 0, 0};
  for (int g = 0; g < P.nsym; g++) {
    if (!P.sym_alwd[g]) continue;
    int c = P.sym_cls[g];
    if (P.sym_wgt[g] < mn[c]) mn[c] = P.sym_wgt[g];
    if (P.sym_wgt[g] > mx[c]) mx[c] = P.sym_wgt[g];
  }
  P.a_minw = mn[0]; P.a_maxw = mx[0];
  P.b_minw = mn[1]; P.b_maxw = mx[1];
  P.c_minw = mn[2]; P.c_maxw = mx[2];
  P.s_minw = P.a_minw;
  if (P.b_minw < P.s_minw) P.s_minw = P.b_minw;
  if (P.c_minw < P.s_minw) P.s_minw = P.c_minw;

  /* add_rule (10014-10042): active iff every required symbol is unrestricted */
  for (int r = 0; r < kNumRules; r++) {
    bool allowed = true;
    for (const char* s = kRules[r].req; *s; s++)
      if (o.sa_alwd[(unsigned char)*s] < RIES_MAX_ELEN) allowed = false;
    if (allowed) {
      int g = S.idx_of[(unsigned char)kRules[r].sym];
      if (g >= 0) P.sym_mask[g] |= kRules[r].mask;
    }
  }

  if (P.sincos_scale == 1.0) { /* radians: [pS] => 0, [pC] => [1n] (10623-10628; no required symbols) */
    int gs = S.idx_of[(unsigned char)'S'], gc = S.idx_of[(unsigned char)'C'];
    if (gs >= 0) P.sym_mask[gs] |= AM_pi;
    if (gc >= 0 && o.sa_alwd[(unsigned char)'1'] >= RIES_MAX_ELEN && o.sa_alwd[(unsigned char)'n'] >= RIES_MAX_ELEN)
      P.sym_mask[gc] |= AM_pi;
  }

  /* g_exhaust_cpx (10669-10685), MAX_ELEN odd */
  {
    int el = RIES_MAX_ELEN >> 1;
    int ex = (el + 1) * P.a_maxw + el * P.c_maxw;
    int alt = P.a_maxw + (RIES_MAX_ELEN - 1) * P.b_maxw;
    if (alt > ex) ex = alt;
    P.exhaust_cpx = (short)ex;
  }

  if (o.target == 0) { /* 10740: fprintf(stderr) + exit(-1) */
    fprintf(stderr, "%s: Target number cannot be zero.\n", o.argv0.c_str());
    throw ExitRequest{-1};
  }
}

/* The part of main() between init2() and the header (ries.c:12111-12317). Split from build_setup so
   that the bare -S command can exit in between, as the reference does. */
void finish_setup(Options& o, Setup& S, const Sink& out) {
  RiesParams& P = S.P;
  if (o.got_target) P.targ_tags = guess_valtype(o.target, o.restrict_subexpr);
  if (o.restrict_subexpr > P.tg_pi) P.tg_pi = (TYPE_RAT < o.restrict_subexpr) ? TYPE_RAT : o.restrict_subexpr;
  if (o.restrict_subexpr > P.tg_phi) P.tg_phi = (TYPE_RAT < o.restrict_subexpr) ? TYPE_RAT : o.restrict_subexpr;
  if (o.restrict_subexpr > P.tg_e) P.tg_e = (TYPE_RAT < o.restrict_subexpr) ? TYPE_RAT : o.restrict_subexpr;

  if (!o.got_target) {
    out.printf("%s: Please specify a target number.\n\n", o.argv0.c_str());
    brief_help(out);
    throw ExitRequest{1};
  }

  /* -i is of no use for a non-integer target: demote to -r (12181-12195).  NOTE this happens AFTER
     init2() built the trial lists, class weight bounds and rule masks, so those keep their -i shape;
     only the per-symbol allowances change ('s' and '^' stay in the lists with allowance 0 and every
     attempt to use them is one "symcount" dead-end, ries.c:9440). */
  if (o.restrict_subexpr == TYPE_INT) {
    const double rv_maxint = 4503599627370496.0; /* init_formats 7726-7738: 2^54 / 4 */
    bool demote = false;
    if (P.targ_tags != TYPE_INT) {
      out.printf("ries: Replacing -i with -r because target isn't an integer.\n");
      demote = true;
    } else if (fabs(o.target) > rv_maxint) {
      out.printf("ries: Replacing -i with -r because target is too large.\n");
      demote = true;
    }
    if (demote) {
      set_restrict_rat(o);
      for (int g = 0; g < P.nsym; g++) {
        P.sym_alwd[g] = o.sa_alwd[P.sym_char[g]];
        S.info[g].alwd = P.sym_alwd[g];
      }
    }
  }
  P.any_limited = 0;
  for (int c = 0; c < 3; c++)
    for (int k = 0; k < P.cls_n[c]; k++)
      if (P.sym_alwd[P.trial[c][k]] < RIES_MAX_ELEN) P.any_limited = 1;
  P.restrict_subexpr = o.restrict_subexpr;
  P.restrict_exponents = o.restrict_exponents;
  P.restrict_trig_args = o.restrict_trig_args;

  S.tgt_to_print = o.target;
  if (o.match_all_digits) { /* 12199-12229 */
    if (o.mag_ulp > 0) {
      o.k_max_match_dist = o.mag_ulp * 0.5;
      if (o.target < 0) o.target -= o.k_max_match_dist;
      else o.target += o.k_max_match_dist;
      if (P.targ_tags == TYPE_INT) P.targ_tags = TYPE_RAT;
      if (!o.wide_output) o.relative_x = false;
    } else {
      out.printf("Ignoring --match-all-digits option because I could not determine\n"
                 "the magnitude of the ULP (unit in the last place).\n");
    }
  }
  P.target = o.target;

  /* k_vanished_dx: given, or auto-scaled (12231-12240) */
  P.k_vanished_dx = 1.0e-6;
  if (o.derivative_margin > 0) {
    P.k_vanished_dx = o.derivative_margin;
  } else if (fabs(o.target) * P.k_vanished_dx > 0.1) {
    P.k_vanished_dx = 0.1 / fabs(o.target);
    out.printf("Auto-setting --derivative-margin %g\n", P.k_vanished_dx);
  } else if (fabs(o.target) < 1.0e-8) {
    P.k_vanished_dx = 0.1 * fabs(o.target);
    out.printf("Auto-setting --derivative-margin %g\n", P.k_vanished_dx);
  }
  double k_prune_deriv = P.k_vanished_dx / 1.0e4; /* 12249 */
  P.p_ovr = fabs(o.target) / k_prune_deriv;
  P.n_ovr = -P.p_ovr;

  if (P.k_vanished_dx >= 1.0 / fabs(o.target)) { /* 12256 */
    out.printf("%s: With this --derivative-margin option (%g) and\n"
               "  target value (%g), I cannot compute any expressions.  Either\n"
               "  use a smaller target value, or a derivative margin below %g\n\n",
               o.argv0.c_str(), P.k_vanished_dx, o.target, 1.0 / fabs(o.target));
    throw ExitRequest{-1};
  }
  if (fabs(o.target) / P.k_vanished_dx >= 1.0e15) { /* 12264 */
    if (o.sa_alwd[(unsigned char)'x'] > 1 && o.derivative_margin > 0) {
      out.printf("WARNING: Your --derivative-margin option (%g)\n"
                 "  is so small that RIES might report tautologies like x/x = 1. To avoid\n"
                 "  this warning, use the option -Ox or give a target value smaller than\n"
                 "  %g.\n\n", o.derivative_margin, 10.0 / 1.0e-6);
    } else if (o.sa_alwd[(unsigned char)'x'] > 1) {
      out.printf("WARNING: Your target value is so large that RIES might report tautologies\n"
                 "  like x/(x/3) = 3. To avoid this warning, use the option -Ox or give a\n"
                 "  target value smaller than %g.\n\n", 10.0 / 1.0e-6);
    }
  }

  if (o.min_matchsize < 0.0) o.min_matchsize = fabs(o.target * (-o.min_matchsize)); /* 12283 */
  if (o.k_max_match_dist >= 0) S.init_match_dist = o.k_max_match_dist;             /* 12288 */
  else S.init_match_dist = fabs(o.target * (-o.k_max_match_dist));
  if ((fabs(o.target) > 1.0) || (fabs(o.target) < 0.25))                           /* 12303 */
    S.k_min_best_match = fabs(o.target) * 8.0 * S.k_precision_ulp;
}

}  // namespace ries