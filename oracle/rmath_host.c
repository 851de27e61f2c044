/* TEST INFRASTRUCTURE.  Host build of the repo's rmath (ries_b200/csrc/core/rmath.h) as ordinary
   extern functions, so that (a) the UNMODIFIED reference source can be linked with its libm calls
   interposed by them (oracle/_ref/ries_rm, see rmath_interpose.c), (b) the C restatement can do the
   same, and (c) tests can compare rmath against the platform libm value by value. */
#define RIES_HD static inline
#include "core/rmath.h"

double rmh_exp(double x) { return rm_exp(x); }
double rmh_log(double x) { return rm_log(x); }
double rmh_pow(double x, double y) { return rm_pow(x, y); }
void rmh_sincos(double x, double *s, double *c) { rm_sincos(x, s, c); }
double rmh_tan(double x) { return rm_tan(x); }
double rmh_atan2(double y, double x) { return rm_atan2(y, x); }
