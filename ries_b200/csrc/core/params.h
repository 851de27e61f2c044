// Plain-old-data shared by the host driver and the sm_100a kernels.
//
// Everything the device needs about one search lives in `RiesParams` (uploaded once into
// __constant__ memory) and, per complexity pass, an array of `FormDesc`.
//
// Reference data model being replaced (all /root/reference/public/ries.c):
//   symbol/sym_attr_block 2870,3179-3195   -> RiesParams::sym_* tables (dense, enabled symbols only)
//   form                  3166-3174        -> FormDesc
//   pe + metastack        3150-3156,3106   -> Node (cactus stack: one stack cell per live prefix)
//   expr (tree node)      2955-2968        -> Row (32 B, no pointers: the table is a sorted array)
#pragma once
#include <stdint.h>

#ifndef RIES_HD
#if defined(__CUDACC__)
#define RIES_HD __host__ __device__ __forceinline__
#else
#define RIES_HD inline
#endif
#endif

#define RIES_MAX_ELEN 21              /* ries.c:2779 */
#define RIES_EXPR_ALLOC (RIES_MAX_ELEN + 1)
#define RIES_STK_MAX ((RIES_MAX_ELEN + 1) >> 1) /* ries.c:3105 */
#define RIES_MAX_SYMS 48              /* enabled symbols, 6-bit codes 1..63 */
#define RIES_SLOTS 16                 /* candidate-symbol slots per live prefix (13 a / 10 b / 8 c max) */

/* number classes, ries.c:2927-2939 */
enum { TYPE_NONE = 0, TYPE_TRAN = 1, TYPE_TCEL = 2, TYPE_ALG = 3, TYPE_CONS = 4, TYPE_RAT = 5, TYPE_INT = 6 };

/* attribute bits, ries.c:3279-3299 */
enum {
  AM_KK = 0x0001, AM_1 = 0x0002, AM_2 = 0x0004, AM_n = 0x0008, AM_r = 0x0010, AM_55 = 0x0020,
  AM_jK = 0x0040, AM_RHS = 0x0080, AM_sq = 0x0100, AM_1K = 0x0200, AM_l = 0x0400, AM_E = 0x0800,
  AM_pi = 0x1000, AM_KxK = 0x2000, AM_KpK = 0x4000, AM_plus = 0x8000, AM_mul = 0x10000,
  AM_pow = 0x20000, AM_a1_e = 0x40000, AM_a1_1 = 0x80000, AM_a1_r = 0x100000
};

/* exec() error codes, ries.c:4517-4539 (only "nonzero" matters to the search) */
enum {
  ERR_EXEC_DIV_ZERO = -1, ERR_EXEC_ROOT_NEG = -2, ERR_EXEC_LOG_NEG = -3, ERR_EXEC_OVERFLOW = -4,
  ERR_EXEC_TRIG_RANGE = -5, ERR_EXEC_TRIG_LOW_DX = -6, ERR_EXEC_SIG_LOSS = -7,
  ERR_EXEC_POW_NEG_BASE = -8, ERR_EXEC_LOG_BAD_BASE = -9, ERR_EXEC_ILLEGAL_SYMBOL = -10,
  ERR_NEWTON_ZERO_DX = -11, ERR_NEWTON_NO_CONVERGE = -12, ERR_EVAL_TOO_LONG = -13,
  ERR_EVAL_STACK_OVERFLOW = -16, ERR_EVAL_STACK_UNDERFLOW = -17,
  ERR_EXEC_ILLEGAL_EXPONENT = -20, ERR_EXEC_TRIG_ARGTYPE = -21, ERR_EXEC_ILLEGAL_DERIV = -22,
  ERR_EXEC_ZERO_DERIV = -23
};

/* One value on the evaluation stack: value, d/dx, number class. */
struct Val {
  double v;
  double d;
  int tag;
};

/* 21 symbols x 6-bit codes (code = dense symbol index + 1, 0 terminates). */
struct Sym128 {
  uint64_t lo, hi;
};

RIES_HD unsigned sym_get(const Sym128& s, int i) {
  int off = 6 * i;
  if (off + 6 <= 64) return (unsigned)((s.lo >> off) & 63u);
  if (off >= 64) return (unsigned)((s.hi >> (off - 64)) & 63u);
  /* i == 10: bits 60..65 straddle */
  return (unsigned)(((s.lo >> 60) | (s.hi << 4)) & 63u);
}

RIES_HD void sym_set(Sym128& s, int i, unsigned code) {
  int off = 6 * i;
  uint64_t c = code & 63u;
  if (off + 6 <= 64) {
    s.lo = (s.lo & ~(63ull << off)) | (c << off);
  } else if (off >= 64) {
    s.hi = (s.hi & ~(63ull << (off - 64))) | (c << (off - 64));
  } else {
    s.lo = (s.lo & ~(15ull << 60)) | ((c & 15ull) << 60);
    s.hi = (s.hi & ~3ull) | (c >> 4);
  }
}

RIES_HD int sym_len(const Sym128& s) {
  int n = 0;
  while (n < RIES_MAX_ELEN + 1 && sym_get(s, n)) n++;
  return n;
}

/* Search-wide constants (replaces the reference's globals ries.c:3201-3383). */
struct RiesParams {
  /* symbol table: EVERY symbol the reference knows (add_symbol calls, ries.c:10375-10468), class a
     first, then b, then c; code = index + 1.  trial[c][0..cls_n[c]) lists the ENABLED symbols of class
     c in ge_2 trial order (= add_symbol order among enabled symbols, ries.c:9891-9909). */
  int nsym;
  int cls_n[3];
  unsigned char trial[3][RIES_SLOTS];
  unsigned char sym_char[RIES_MAX_SYMS];
  unsigned char sym_cls[RIES_MAX_SYMS]; /* 0 a, 1 b, 2 c */
  short sym_wgt[RIES_MAX_SYMS];
  short sym_alwd[RIES_MAX_SYMS];
  int sym_mask[RIES_MAX_SYMS];
  int sym_amkey[RIES_MAX_SYMS];
  int any_limited; /* some enabled symbol has sa_alwd < MAX_ELEN (-O) */

  /* class weight bounds (setup_abc_mmw ries.c:9915) */
  short a_minw, a_maxw, b_minw, b_maxw, c_minw, c_maxw, s_minw, exhaust_cpx;

  /* numerics */
  double target;       /* g_target */
  double k_phi, k_e, k_pi;
  double sincos_scale; /* k_sincos_arg_scale */
  double trig_max_arg; /* k_pi * k_sincos_max_arg */
  double k_sin_clip, k_eXlim, k_sig_loss, k_vanished_dx;
  double p_ovr, n_ovr;
  double min_equ_val, max_equ_val;
  double k_newton_settled;
  int tg_phi, tg_e, tg_pi, targ_tags;
  int restrict_subexpr, restrict_exponents, restrict_trig_args;
  int trig_scale_default;
  int no_cv_simplify;
  /* --canon-reduction (g_canon_ops, ries.c:8728-8732): bit 0 negate, 1 reciprocal, 2 divide by 2, 3 multiply
     by 2; canon_sym[] = dense indices of 'n' 'r' '2' '/' '*' (what canonval appends, 8786-8875) */
  int canon_ops;
  signed char canon_sym[8];
};

/* One a/b/c form of a pass (gf_1 ries.c:9599-9767 + ge_1's suffix tables 9535-9566). */
struct FormDesc {
  unsigned char flen;
  unsigned char cls[RIES_EXPR_ALLOC + 1];  /* 0 a, 1 b, 2 c */
  signed char arg1[RIES_EXPR_ALLOC + 1];   /* -1 = n/a */
  short rminw[RIES_EXPR_ALLOC + 2];
  short rmaxw[RIES_EXPR_ALLOC + 2];
};

/* A live prefix.  Its evaluation stack is a cactus stack: `top` is this prefix's top-of-stack
   and `below` is the index of the node whose `top` is the next cell down (NO_NODE if none). */
#define RIES_NO_NODE 0xffffffffu
struct __attribute__((aligned(16))) Node {
  double val;
  double der;
  Sym128 sym;
  uint32_t below;
  uint16_t form;
  int16_t cplx;
  uint8_t tag;
  uint8_t len;
  uint8_t sp; /* stack depth after this prefix */
  uint8_t pad0;
  uint32_t pad1;
};

/* One row of the value table (and one survivor of a pass): 32 bytes.
   der == 0  <=>  RHS (same convention as the reference, ries.c:8496). */
struct __attribute__((aligned(16))) Row {
  double val;
  double der;
  Sym128 sym;
};

/* Outcome of trying one symbol after one live prefix. */
enum { OUT_NONE = 0, OUT_DEAD = 1, OUT_ERR = 2, OUT_LIVE = 3, OUT_DONE = 4 };

/* A candidate the device hands back to the host replay (time order = (seq, kind)). */
enum { EV_EXACT = 0, EV_PREV = 1, EV_NEXT = 2 };
struct MatchEvent {
  uint32_t seq;      /* tiny passes: insertion number of the NEW expression inside its pass; general passes: its
                        form index (insertion order = (form, symbols), device/state.cuh SeqKey) */
  uint32_t kind;     /* EV_* ; replay order inside one seq is EXACT, PREV, NEXT */
  Row lhs;           /* as stored (before cv_simplify) */
  Row rhs;
  /* filled by the refine kernel (cv_simplify + newton, ries.c:7297,7139) for PREV/NEXT: */
  Sym128 lhs_s, rhs_s;  /* after cv_simplify */
  double root;
  double diff_dx;
  int32_t newton_err;
  int32_t pad;
};
