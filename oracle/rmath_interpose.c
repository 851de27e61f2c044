/* TEST INFRASTRUCTURE.  Defines the libm entry points the reference calls (ries.c:2666-2701; after
   gcc -O2: exp log pow sincos tan atan2) in terms of the repo's rmath host twin, so that linking this
   object into an executable interposes them ahead of libm.so.6. */
double rmh_exp(double), rmh_log(double), rmh_pow(double, double), rmh_tan(double), rmh_atan2(double, double);
void rmh_sincos(double, double *, double *);
double exp(double x) { return rmh_exp(x); }
double log(double x) { return rmh_log(x); }
double pow(double x, double y) { return rmh_pow(x, y); }
double tan(double x) { return rmh_tan(x); }
double atan2(double y, double x) { return rmh_atan2(y, x); }
void sincos(double x, double *s, double *c) { rmh_sincos(x, s, c); }
double sin(double x) { double s, c; rmh_sincos(x, &s, &c); return s; }
double cos(double x) { double s, c; rmh_sincos(x, &s, &c); return c; }
