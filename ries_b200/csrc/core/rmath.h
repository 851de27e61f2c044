// rmath: the seven libm entry points the search calls (ries.c:2666-2701: exp log pow sin cos tan
// atan2; sin+cos always come as one sincos call), as ONE deterministic fp64 implementation compiled for both host and device.
//
// The reference links glibc's libm (2.39 on this image; not part of /root/reference), and the value
// table dedupes on `==` of doubles, so the device needs the same bits glibc produces.
// rmath_impl.h restates the algorithms glibc 2.39 ships for x86-64 with FMA (the variants its ifunc
// resolvers select on every FMA+AVX2 CPU): Szabolcs Nagy's exp/log/pow (sysdeps/ieee754/dbl-64/
// e_exp.c, e_log.c, e_pow.c) and IBM's Accurate Mathematical Library sin/cos/tan/atan2 (s_sin.c,
// s_tan.c, e_atan2.c), with every fused multiply-add written explicitly so that nvcc -fmad=false and
// gcc -ffp-contract=off produce identical results.
//
// Build switches:
//   RIES_RMATH_LIBM   host only: forward to the platform libm (used by the CPU emulation to compare
//                     the batch formulation against the unmodified reference bit for bit).
#pragma once
#include <math.h>
#include <stdint.h>

/* self-contained (valid C and C++/CUDA) so the host twin can be built with plain gcc */
#ifndef RIES_HD
#if defined(__CUDACC__)
#define RIES_HD __host__ __device__ __forceinline__
#elif defined(__cplusplus)
#define RIES_HD inline
#else
#define RIES_HD static inline
#endif
#endif

#if defined(RIES_RMATH_LIBM) && !defined(__CUDA_ARCH__)
RIES_HD double rm_exp(double x) { return exp(x); }
RIES_HD double rm_log(double x) { return log(x); }
RIES_HD double rm_pow(double x, double y) { return pow(x, y); }
RIES_HD void rm_sincos(double x, double* s, double* c) { *s = sin(x); *c = cos(x); } /* gcc: one sincos() */
RIES_HD double rm_tan(double x) { return tan(x); }
RIES_HD double rm_atan2(double y, double x) { return atan2(y, x); }
#else
#include "rmath_impl.h"
#endif
