/*
 * oracle/netlib_l1.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * The reference pins every Level-1 routine bit-exactly to Fortran netlib BLAS
 * 3.7.0 (`gfortran -c -O3 *.f`, .travis-setup.sh:19-23,44-51) through the FFI
 * bindings of test/Foreign.hs:32-81 and the properties of test/Main.hs:28-52.
 * netlib BLAS is a THIRD-PARTY dependency that is not vendored in
 * /root/reference and cannot be built here (no gfortran).  This file restates
 * the PUBLISHED algorithms of the 3.7.0 single-precision Level-1 routines
 * (sdot.f, sasum.f, snrm2.f, isamax.f, saxpy.f, sscal.f, scopy.f, sswap.f,
 * srot.f, srotm.f, srotg.f, srotmg.f, sdsdot.f and their d* twins) in C, in-place and with the
 * Fortran loop structure (clean-up loop + unrolled body, 1-based indices), so
 * tests/test_oracle_netlib.py can replay the reference's own property
 * "native == netlib" against oracle/lblas_oracle.c on this machine.
 *
 * Arguments are by value instead of Fortran's by-reference; arrays are the
 * 0-based C view of the 1-based Fortran arrays (X(i) is x[i-1]).
 */
#include <math.h>
#include <stdint.h>

typedef int64_t i64;
#define X(i) x[(i) - 1]
#define Y(i) y[(i) - 1]

#define REAL float
#define FN(n) nl_s##n
#define FNI(n) nl_is##n
#define FABS(v) fabsf(v)
#define SQRT(v) sqrtf(v)
#define COPYSIGN(a, b) copysignf(a, b)
#define LIT(v) v##f
#define NL_GAMSQ 1.67772E7f
#define NL_RGAMSQ 5.96046E-8f
#include "netlib_body.inc"
#undef REAL
#undef FN
#undef FNI
#undef FABS
#undef SQRT
#undef COPYSIGN
#undef LIT
#undef NL_GAMSQ
#undef NL_RGAMSQ

#define REAL double
#define FN(n) nl_d##n
#define FNI(n) nl_id##n
#define FABS(v) fabs(v)
#define SQRT(v) sqrt(v)
#define COPYSIGN(a, b) copysign(a, b)
#define LIT(v) v
#define NL_GAMSQ 16777216.0          /* drotmg.f: DATA GAM,GAMSQ,RGAMSQ/4096.D0,16777216.D0,5.9604645D-8/ */
#define NL_RGAMSQ 5.9604645E-8
#include "netlib_body.inc"

float nl_sdsdot(i64 n, float sb, const float *x, i64 incx, const float *y, i64 incy)
{
    double dsdot = (double)sb;
    if (n <= 0) return (float)dsdot;
    if (incx == incy && incx > 0) {
        i64 ns = n * incx, i;
        for (i = 1; i <= ns; i += incx) dsdot = dsdot + (double)X(i) * (double)Y(i);
    } else {
        i64 kx = 1, ky = 1, i;
        if (incx < 0) kx = 1 + (1 - n) * incx;
        if (incy < 0) ky = 1 + (1 - n) * incy;
        for (i = 1; i <= n; ++i) { dsdot = dsdot + (double)X(kx) * (double)Y(ky); kx += incx; ky += incy; }
    }
    return (float)dsdot;
}

