// Form enumeration for one pass: gen_forms()/gf_1() (/root/reference/public/ries.c:9599-9801) plus
// the per-form suffix weight tables ge_1() builds (9535-9566).  Host-side and tiny (a few thousand
// forms at -l6); the visiting order (a form is expanded BEFORE its extensions, classes tried a<b<c)
// defines the form index that orders expressions of one pass in time.
#include "host/ries_host.h"

namespace ries {

namespace {
struct FormState {
  int flen, stack, min_weight, max_weight;
  unsigned char cls[RIES_EXPR_ALLOC + 2];
  short stk[RIES_EXPR_ALLOC + 2];
  signed char arg1[RIES_EXPR_ALLOC + 2];
};

struct Ctx {
  const RiesParams* P;
  int e_minw, e_maxw;
  std::vector<FormDesc>* out;
  double n_leaf;
};

void emit(Ctx& c, const FormState& f) {
  const RiesParams& P = *c.P;
  FormDesc d;
  d.flen = (unsigned char)f.flen;
  for (int i = 0; i < RIES_EXPR_ALLOC + 1; i++) { d.cls[i] = 0; d.arg1[i] = -1; }
  for (int i = 0; i < f.flen; i++) { d.cls[i] = f.cls[i]; d.arg1[i] = f.arg1[i]; }
  int rminw = 0, rmaxw = 0;
  for (int i = 0; i < RIES_EXPR_ALLOC + 2; i++) { d.rminw[i] = 0; d.rmaxw[i] = 0; }
  for (int i = f.flen; i > 0;) {
    i--;
    switch (f.cls[i]) {
      case 0: rminw += P.a_minw; rmaxw += P.a_maxw; break;
      case 1: rminw += P.b_minw; rmaxw += P.b_maxw; break;
      default: rminw += P.c_minw; rmaxw += P.c_maxw; break;
    }
    d.rminw[i] = (short)rminw;
    d.rmaxw[i] = (short)rmaxw;
  }
  c.out->push_back(d);
}

void gf_1(Ctx& c, const FormState& base) {
  const RiesParams& P = *c.P;
  int blen = base.flen;
  int slim = (P.cls_n[2] == 0) ? 1 : 2;   /* no class-c symbols: only a and b (9638) */
  int sincr = (P.cls_n[1] == 0) ? 2 : 1;  /* no class-b symbols: skip b (9649) */
  for (int s = 0; s <= slim; s += sincr) {
    FormState next = base;
    next.flen = blen + 1;
    next.cls[blen] = (unsigned char)s;
    next.stack = base.stack + (1 - s);
    if (s == 0) { next.min_weight = base.min_weight + P.a_minw; next.max_weight = base.max_weight + P.a_maxw; }
    else if (s == 1) { next.min_weight = base.min_weight + P.b_minw; next.max_weight = base.max_weight + P.b_maxw; }
    else { next.min_weight = base.min_weight + P.c_minw; next.max_weight = base.max_weight + P.c_maxw; }

    bool recurse = true;
    if (next.stack < 1) recurse = false;
    else if (next.min_weight + P.s_minw > c.e_maxw) recurse = false;
    else if (blen > RIES_MAX_ELEN) recurse = false;
    else if ((next.flen + next.stack) > RIES_MAX_ELEN + 1) recurse = false;

    next.stk[blen] = (short)next.stack;
    next.arg1[blen] = -1;
    if (s == 2) {
      for (int i = blen - 2; i >= 0; i--) {
        if (next.stk[i] == next.stack) { next.arg1[blen] = (signed char)i; break; }
      }
    }

    bool gen_expr = false;
    if (next.stack == 1) {
      gen_expr = true;
      if (next.min_weight > c.e_maxw) gen_expr = false;
      else if (next.max_weight < c.e_minw) gen_expr = false;
    }
    /* The packed expression format holds MAX_ELEN symbols; the reference's own arrays overflow for
       the 22-symbol forms its length test lets through, which need complexity > 150. */
    if (gen_expr && next.flen <= RIES_MAX_ELEN) emit(c, next);
    if (recurse) gf_1(c, next);
    if (gen_expr) c.n_leaf += 1;
  }
}
}  // namespace

void gen_forms(const RiesParams& P, int e_minw, int e_maxw, std::vector<FormDesc>& forms,
               double* n_leaf_forms) {
  forms.clear();
  Ctx c{&P, e_minw, e_maxw, &forms, 0};
  FormState base;
  base.flen = 0; base.stack = 0; base.min_weight = 0; base.max_weight = 0;
  gf_1(c, base);
  if (n_leaf_forms) *n_leaf_forms = c.n_leaf;
}

}  // namespace ries
