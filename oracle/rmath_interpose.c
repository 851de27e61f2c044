/* TEST INFRASTRUCTURE.  Defines the libm entry points the reference calls (ries.c:2666-2701) in terms
   of the repo's rmath host twin, so that linking this object into an executable interposes them.
   sincos is included because gcc -O2 folds sin(a)+cos(a) into one sincos call. */
double rmh_exp(double), rmh_log(double), rmh_pow(double, double), rmh_sin(double), rmh_cos(double),
    rmh_tan(double), rmh_atan2(double, double);
double exp(double x) { return rmh_exp(x); }
double log(double x) { return rmh_log(x); }
double pow(double x, double y) { return rmh_pow(x, y); }
double sin(double x) { return rmh_sin(x); }
double cos(double x) { return rmh_cos(x); }
double tan(double x) { return rmh_tan(x); }
double atan2(double y, double x) { return rmh_atan2(y, x); }
void sincos(double x, double *s, double *c) { *s = rmh_sin(x); *c = rmh_cos(x); }
