// TEMPORARY v0: platform math (glibc on host, libdevice on device). Replaced by the glibc-2.39
// restatement as each function is pinned against libm.so.6.
#pragma once
RIES_HD double rm_exp(double x) { return exp(x); }
RIES_HD double rm_log(double x) { return log(x); }
RIES_HD double rm_pow(double x, double y) { return pow(x, y); }
RIES_HD double rm_sin(double x) { return sin(x); }
RIES_HD double rm_cos(double x) { return cos(x); }
RIES_HD double rm_tan(double x) { return tan(x); }
RIES_HD double rm_atan2(double y, double x) { return atan2(y, x); }
