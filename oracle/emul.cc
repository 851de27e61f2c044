// TEST INFRASTRUCTURE -- never linked into the product.
//
// Batch-semantics CPU emulation of the device pipeline: the same
// per-thread core functions the sm_100a kernels call (ries_b200/csrc/core/{expand,exec,refine}.h) are
// driven here by plain sequential loops, behind the same Engine interface the host driver uses for
// the GPU.  Purpose: prove on a machine without a GPU that the batch formulation
//   generate a whole pass -> stable sort by value -> dedupe -> merge -> gap-local neighbour test
//   -> replay candidates in insertion order
// reproduces the reference's streaming binary-tree semantics (bt_insert / check_sides,
// /root/reference/public/ries.c:8622-8724, 8488-8539) byte for byte.  Built two ways:
//   -DRIES_RMATH_LIBM : libm = glibc  -> must equal oracle/_ref/ries (the unmodified reference)
//   (default)         : libm = rmath  -> must equal oracle/_ref/ries_rm and the GPU
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "core/expand.h"
#include "core/refine.h"
#include "host/ries_host.h"

namespace {

struct EmulEngine : public ries::Engine {
  RiesParams P;
  std::vector<Row> table;            /* sorted by value, unique */
  std::vector<uint32_t> ins_seq;     /* seqs of rows inserted by the last pass */

  const char* name() const override { return "cpu-emulation"; }
  void begin(const RiesParams& p) override { P = p; table.clear(); }

  long inserted_before(const MatchEvent& ev, bool inclusive) override {
    const uint32_t seq = ev.seq;
    long n = 0;
    for (uint32_t s : ins_seq) n += inclusive ? (s <= seq) : (s < seq);
    return n;
  }

  void run_pass(int using_x, int cplx, const std::vector<FormDesc>& forms, double best_match0,
                bool skip_exact, ries::PassOutput& out) override {
    ins_seq.clear();
    /* K1: level-synchronous expansion over all forms of the pass */
    std::vector<Node> nodes;
    std::vector<std::vector<Row>> done_by_form(forms.size());
    size_t lvl_begin = 0;
    for (size_t f = 0; f < forms.size(); f++) {
      Node r;
      memset(&r, 0, sizeof(r));
      r.below = RIES_NO_NODE;
      r.form = (uint16_t)f;
      nodes.push_back(r);
    }
    size_t lvl_end = nodes.size();
    double dead = 0, gen = 0;
    while (lvl_end > lvl_begin) {
      for (size_t p = lvl_begin; p < lvl_end; p++) {
        Node par = nodes[p];
        const FormDesc& F = forms[par.form];
        int atts = prefix_attrs(P, F, par, using_x);
        for (int slot = 0; slot < RIES_SLOTS; slot++) {
          Node child;
          int oc = expand_one(P, F, nodes.data(), par, (uint32_t)p, slot, atts, using_x, cplx, cplx, child);
          if (oc == OUT_DEAD) dead += 1;
          else if (oc == OUT_LIVE) nodes.push_back(child);
          else if (oc == OUT_DONE) {
            gen += 1;
            Row r;
            r.val = child.val; r.der = child.der; r.sym = child.sym;
            done_by_form[child.form].push_back(r);
          }
        }
      }
      lvl_begin = lvl_end;
      lvl_end = nodes.size();
    }
    /* survivors in insertion order: (form index, lexicographic) */
    std::vector<Row> surv;
    for (auto& v : done_by_form) surv.insert(surv.end(), v.begin(), v.end());
    out.generated = gen;
    out.dead_ends = dead;

    /* K2: stable sort by value (seq = index in surv) */
    std::vector<uint32_t> order(surv.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = (uint32_t)i;
    std::stable_sort(order.begin(), order.end(),
                     [&](uint32_t a, uint32_t b) { return surv[a].val < surv[b].val; });

    /* dedupe inside the pass + against the table; EXACT events */
    std::vector<uint32_t> uniq; /* seqs of inserted rows, sorted by value */
    std::vector<size_t> gap;    /* lower_bound position in the old table */
    for (size_t i = 0; i < order.size(); i++) {
      uint32_t s = order[i];
      if (i > 0 && surv[order[i - 1]].val == surv[s].val) {
        /* duplicate of an earlier expression of the same pass, or of the same table row */
        size_t pos = std::lower_bound(table.begin(), table.end(), surv[s].val,
                                      [](const Row& r, double v) { return r.val < v; }) - table.begin();
        if (pos < table.size() && table[pos].val == surv[s].val) emit_exact(using_x, skip_exact, s, surv[s], table[pos], out);
        continue;
      }
      size_t pos = std::lower_bound(table.begin(), table.end(), surv[s].val,
                                    [](const Row& r, double v) { return r.val < v; }) - table.begin();
      if (pos < table.size() && table[pos].val == surv[s].val) {
        emit_exact(using_x, skip_exact, s, surv[s], table[pos], out);
        continue;
      }
      uniq.push_back(s);
      gap.push_back(pos);
    }
    out.inserted = (long)uniq.size();
    ins_seq = uniq;

    /* K3: gap-local neighbour test */
    for (size_t i = 0; i < uniq.size(); i++) {
      uint32_t s = uniq[i];
      bool test_prev = gap[i] > 0, test_next = gap[i] < table.size();
      for (size_t j = i; j-- > 0 && gap[j] == gap[i];)
        if (uniq[j] < s) { test_prev = false; break; }
      for (size_t j = i + 1; j < uniq.size() && gap[j] == gap[i]; j++)
        if (uniq[j] < s) { test_next = false; break; }
      if (test_prev) emit_neighbour(using_x, EV_PREV, s, surv[s], table[gap[i] - 1], best_match0, out);
      if (test_next) emit_neighbour(using_x, EV_NEXT, s, surv[s], table[gap[i]], best_match0, out);
    }

    /* merge */
    std::vector<Row> merged;
    merged.reserve(table.size() + uniq.size());
    size_t a = 0, b = 0;
    while (a < table.size() || b < uniq.size()) {
      if (b >= uniq.size() || (a < table.size() && table[a].val < surv[uniq[b]].val)) merged.push_back(table[a++]);
      else merged.push_back(surv[uniq[b++]]);
    }
    table.swap(merged);

    /* K4: refine + order */
    for (auto& ev : out.events)
      if (ev.kind != EV_EXACT) refine_event(P, ev);
    std::stable_sort(out.events.begin(), out.events.end(), [](const MatchEvent& x, const MatchEvent& y) {
      return x.seq != y.seq ? x.seq < y.seq : x.kind < y.kind;
    });
  }

  void emit_exact(int using_x, bool skip_exact, uint32_t seq, const Row& fresh, const Row& tree,
                  ries::PassOutput& out) {
    if (skip_exact) return;
    bool tree_is_lhs = (tree.der != 0.0);
    if (using_x ? tree_is_lhs : !tree_is_lhs) return; /* same side: discard (8552, 8583) */
    MatchEvent ev;
    memset(&ev, 0, sizeof(ev));
    ev.seq = seq; ev.kind = EV_EXACT;
    if (using_x) { ev.lhs = fresh; ev.rhs = tree; } else { ev.lhs = tree; ev.rhs = fresh; }
    out.events.push_back(ev);
  }

  void emit_neighbour(int using_x, int kind, uint32_t seq, const Row& fresh, const Row& other,
                      double best_match0, ries::PassOutput& out) {
    bool other_is_lhs = (other.der != 0.0);
    if (using_x ? other_is_lhs : !other_is_lhs) return; /* same side (8502, 8523) */
    const Row& L = using_x ? fresh : other;
    const Row& R = using_x ? other : fresh;
    double total = L.der - R.der;
    double score = fabs((R.val - L.val) / total);
    if (score > best_match0) return; /* superset filter: the threshold never grows (8270) */
    MatchEvent ev;
    memset(&ev, 0, sizeof(ev));
    ev.seq = seq; ev.kind = (uint32_t)kind;
    ev.lhs = L; ev.rhs = R;
    out.events.push_back(ev);
  }
};

void to_stdout(void*, const char* b, size_t n) { fwrite(b, 1, n, stdout); }

}  // namespace

int main(int argc, char** argv) {
  ries::Sink sink{to_stdout, 0};
  ries::Options opt;
  EmulEngine eng;
  int code = 0;
  try {
    if (argc > 1) {
      ries::parse_args(argc, argv, opt, sink);
    } else {
      sink.printf("%s: Please specify a target number.\n\n", argv[0]);
      ries::brief_help(sink);
      throw ries::ExitRequest{1};
    }
    code = ries::run_search(opt, eng, sink, 0);
  } catch (const ries::ExitRequest& e) {
    code = e.code;
  }
  fflush(stdout);
  return code & 0xff;
}
