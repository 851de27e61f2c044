/* TEST INFRASTRUCTURE.  Host build of the repo's rmath (ries_b200/csrc/core/rmath.h) as ordinary
   extern functions, so that (a) the UNMODIFIED reference source can be compiled with its libm calls
   redirected to them (oracle/_ref/ries_rm: -Dexp=rm_exp ...), (b) the C restatement can do the same,
   and (c) tests can compare rmath against the platform libm value by value (librmath_host.so). */
#define RIES_HD static inline
#include "core/rmath.h"

#undef rm_exp
double rmh_exp(double x) { return rm_exp(x); }
double rmh_log(double x) { return rm_log(x); }
double rmh_pow(double x, double y) { return rm_pow(x, y); }
double rmh_sin(double x) { return rm_sin(x); }
double rmh_cos(double x) { return rm_cos(x); }
double rmh_tan(double x) { return rm_tan(x); }
double rmh_atan2(double y, double x) { return rm_atan2(y, x); }
