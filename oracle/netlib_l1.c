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
 * srot.f, srotm.f, srotg.f, srotmg.f, sdsdot.f) in C, in-place and with the
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

float nl_sdot(i64 n, const float *x, i64 incx, const float *y, i64 incy)
{
    float stemp = 0.0f;
    if (n <= 0) return 0.0f;
    if (incx == 1 && incy == 1) {
        i64 m = n % 5, i;
        if (m != 0) {
            for (i = 1; i <= m; ++i) stemp = stemp + X(i) * Y(i);
            if (n < 5) return stemp;
        }
        for (i = m + 1; i <= n; i += 5)
            stemp = stemp + X(i) * Y(i) + X(i + 1) * Y(i + 1) + X(i + 2) * Y(i + 2)
                          + X(i + 3) * Y(i + 3) + X(i + 4) * Y(i + 4);
    } else {
        i64 ix = 1, iy = 1, i;
        if (incx < 0) ix = (-n + 1) * incx + 1;
        if (incy < 0) iy = (-n + 1) * incy + 1;
        for (i = 1; i <= n; ++i) { stemp = stemp + X(ix) * Y(iy); ix += incx; iy += incy; }
    }
    return stemp;
}

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

float nl_sasum(i64 n, const float *x, i64 incx)
{
    float stemp = 0.0f;
    if (n <= 0 || incx <= 0) return 0.0f;
    if (incx == 1) {
        i64 m = n % 6, i;
        if (m != 0) {
            for (i = 1; i <= m; ++i) stemp = stemp + fabsf(X(i));
            if (n < 6) return stemp;
        }
        for (i = m + 1; i <= n; i += 6)
            stemp = stemp + fabsf(X(i)) + fabsf(X(i + 1)) + fabsf(X(i + 2))
                          + fabsf(X(i + 3)) + fabsf(X(i + 4)) + fabsf(X(i + 5));
    } else {
        i64 nincx = n * incx, i;
        for (i = 1; i <= nincx; i += incx) stemp = stemp + fabsf(X(i));
    }
    return stemp;
}

/* snrm2.f of 3.7.0: the scale/ssq ("slassq") formulation */
float nl_snrm2(i64 n, const float *x, i64 incx)
{
    if (n < 1 || incx < 1) return 0.0f;
    if (n == 1) return fabsf(X(1));
    float scale = 0.0f, ssq = 1.0f;
    i64 ix;
    for (ix = 1; ix <= 1 + (n - 1) * incx; ix += incx) {
        if (X(ix) != 0.0f) {
            float absxi = fabsf(X(ix));
            if (scale < absxi) {
                float q = scale / absxi;
                ssq = 1.0f + ssq * (q * q);
                scale = absxi;
            } else {
                float q = absxi / scale;
                ssq = ssq + q * q;
            }
        }
    }
    return scale * sqrtf(ssq);
}

/* 1-based result, 0 for n<1 or incx<=0 */
i64 nl_isamax(i64 n, const float *x, i64 incx)
{
    if (n < 1 || incx <= 0) return 0;
    i64 imax = 1, i;
    if (n == 1) return 1;
    float smax;
    if (incx == 1) {
        smax = fabsf(X(1));
        for (i = 2; i <= n; ++i)
            if (fabsf(X(i)) > smax) { imax = i; smax = fabsf(X(i)); }
    } else {
        i64 ix = 1;
        smax = fabsf(X(1));
        ix += incx;
        for (i = 2; i <= n; ++i) {
            if (fabsf(X(ix)) > smax) { imax = i; smax = fabsf(X(ix)); }
            ix += incx;
        }
    }
    return imax;
}

void nl_saxpy(i64 n, float sa, const float *x, i64 incx, float *y, i64 incy)
{
    if (n <= 0) return;
    if (sa == 0.0f) return;
    if (incx == 1 && incy == 1) {
        i64 m = n % 4, i;
        if (m != 0) for (i = 1; i <= m; ++i) Y(i) = Y(i) + sa * X(i);
        if (n < 4) return;
        for (i = m + 1; i <= n; i += 4) {
            Y(i) = Y(i) + sa * X(i);
            Y(i + 1) = Y(i + 1) + sa * X(i + 1);
            Y(i + 2) = Y(i + 2) + sa * X(i + 2);
            Y(i + 3) = Y(i + 3) + sa * X(i + 3);
        }
    } else {
        i64 ix = 1, iy = 1, i;
        if (incx < 0) ix = (-n + 1) * incx + 1;
        if (incy < 0) iy = (-n + 1) * incy + 1;
        for (i = 1; i <= n; ++i) { Y(iy) = Y(iy) + sa * X(ix); ix += incx; iy += incy; }
    }
}

void nl_sscal(i64 n, float sa, float *x, i64 incx)
{
    if (n <= 0 || incx <= 0) return;
    if (incx == 1) {
        i64 m = n % 5, i;
        if (m != 0) {
            for (i = 1; i <= m; ++i) X(i) = sa * X(i);
            if (n < 5) return;
        }
        for (i = m + 1; i <= n; i += 5) {
            X(i) = sa * X(i); X(i + 1) = sa * X(i + 1); X(i + 2) = sa * X(i + 2);
            X(i + 3) = sa * X(i + 3); X(i + 4) = sa * X(i + 4);
        }
    } else {
        i64 nincx = n * incx, i;
        for (i = 1; i <= nincx; i += incx) X(i) = sa * X(i);
    }
}

void nl_scopy(i64 n, const float *x, i64 incx, float *y, i64 incy)
{
    if (n <= 0) return;
    if (incx == 1 && incy == 1) {
        i64 m = n % 7, i;
        if (m != 0) {
            for (i = 1; i <= m; ++i) Y(i) = X(i);
            if (n < 7) return;
        }
        for (i = m + 1; i <= n; i += 7) {
            Y(i) = X(i); Y(i + 1) = X(i + 1); Y(i + 2) = X(i + 2); Y(i + 3) = X(i + 3);
            Y(i + 4) = X(i + 4); Y(i + 5) = X(i + 5); Y(i + 6) = X(i + 6);
        }
    } else {
        i64 ix = 1, iy = 1, i;
        if (incx < 0) ix = (-n + 1) * incx + 1;
        if (incy < 0) iy = (-n + 1) * incy + 1;
        for (i = 1; i <= n; ++i) { Y(iy) = X(ix); ix += incx; iy += incy; }
    }
}

void nl_sswap(i64 n, float *x, i64 incx, float *y, i64 incy)
{
    if (n <= 0) return;
    if (incx == 1 && incy == 1) {
        i64 i;   /* the 3-unrolled body is elementwise: order is irrelevant */
        for (i = 1; i <= n; ++i) { float t = X(i); X(i) = Y(i); Y(i) = t; }
    } else {
        i64 ix = 1, iy = 1, i;
        if (incx < 0) ix = (-n + 1) * incx + 1;
        if (incy < 0) iy = (-n + 1) * incy + 1;
        for (i = 1; i <= n; ++i) { float t = X(ix); X(ix) = Y(iy); Y(iy) = t; ix += incx; iy += incy; }
    }
}

void nl_srot(i64 n, float *x, i64 incx, float *y, i64 incy, float c, float s)
{
    if (n <= 0) return;
    i64 ix = 1, iy = 1, i;
    if (!(incx == 1 && incy == 1)) {
        if (incx < 0) ix = (-n + 1) * incx + 1;
        if (incy < 0) iy = (-n + 1) * incy + 1;
    }
    for (i = 1; i <= n; ++i) {
        float stemp = c * X(ix) + s * Y(iy);
        Y(iy) = c * Y(iy) - s * X(ix);
        X(ix) = stemp;
        ix += incx; iy += incy;
    }
}

void nl_srotm(i64 n, float *x, i64 incx, float *y, i64 incy, const float *sparam)
{
    float sflag = sparam[0];
    if (n <= 0 || (sflag + 2.0f == 0.0f)) return;
    i64 kx = 1, ky = 1, i;
    if (!(incx == incy && incx > 0)) {
        if (incx < 0) kx = 1 + (1 - n) * incx;
        if (incy < 0) ky = 1 + (1 - n) * incy;
    }
    if (sflag < 0.0f) {
        float sh11 = sparam[1], sh12 = sparam[3], sh21 = sparam[2], sh22 = sparam[4];
        for (i = 1; i <= n; ++i) {
            float w = X(kx), z = Y(ky);
            X(kx) = w * sh11 + z * sh12;
            Y(ky) = w * sh21 + z * sh22;
            kx += incx; ky += incy;
        }
    } else if (sflag == 0.0f) {
        float sh12 = sparam[3], sh21 = sparam[2];
        for (i = 1; i <= n; ++i) {
            float w = X(kx), z = Y(ky);
            X(kx) = w + z * sh12;
            Y(ky) = w * sh21 + z;
            kx += incx; ky += incy;
        }
    } else {
        float sh11 = sparam[1], sh22 = sparam[4];
        for (i = 1; i <= n; ++i) {
            float w = X(kx), z = Y(ky);
            X(kx) = w * sh11 + z;
            Y(ky) = -w + sh22 * z;
            kx += incx; ky += incy;
        }
    }
}

/* srotg.f: in/out sa -> r, sb -> z; out c, s */
void nl_srotg(float *sa, float *sb, float *c, float *s)
{
    float roe = *sb, r, z;
    if (fabsf(*sa) > fabsf(*sb)) roe = *sa;
    float scale = fabsf(*sa) + fabsf(*sb);
    if (scale == 0.0f) {
        *c = 1.0f; *s = 0.0f; r = 0.0f; z = 0.0f;
    } else {
        float qa = *sa / scale, qb = *sb / scale;
        r = scale * sqrtf(qa * qa + qb * qb);
        r = copysignf(1.0f, roe) * r;
        *c = *sa / r;
        *s = *sb / r;
        z = 1.0f;
        if (fabsf(*sa) > fabsf(*sb)) z = *s;
        if (fabsf(*sb) >= fabsf(*sa) && *c != 0.0f) z = 1.0f / *c;
    }
    *sa = r;
    *sb = z;
}

/* srotmg.f of 3.7.0.  Local H entries start at zero here (Fortran leaves them
 * uninitialised; the reference's wrapper reads only the ones the flag names,
 * test/Foreign.hs:248-253). */
void nl_srotmg(float *sd1, float *sd2, float *sx1, float sy1, float *sparam)
{
    const float zero = 0.0f, one = 1.0f, two = 2.0f;
    const float gam = 4096.0f, gamsq = 1.67772E7f, rgamsq = 5.96046E-8f;
    float sflag = zero, sh11 = zero, sh12 = zero, sh21 = zero, sh22 = zero;
    float sp1, sp2, sq1, sq2, su, stemp;
    int guard;

    if (*sd1 < zero) {
        sflag = -one;
        sh11 = zero; sh12 = zero; sh21 = zero; sh22 = zero;
        *sd1 = zero; *sd2 = zero; *sx1 = zero;
    } else {
        sp2 = *sd2 * sy1;
        if (sp2 == zero) {
            sflag = -two;
            sparam[0] = sflag;
            return;
        }
        sp1 = *sd1 * *sx1;
        sq2 = sp2 * sy1;
        sq1 = sp1 * *sx1;
        if (fabsf(sq1) > fabsf(sq2)) {
            sh21 = -sy1 / *sx1;
            sh12 = sp2 / sp1;
            su = one - sh12 * sh21;
            if (su > zero) {
                sflag = zero;
                *sd1 = *sd1 / su;
                *sd2 = *sd2 / su;
                *sx1 = *sx1 * su;
            }
        } else {
            if (sq2 < zero) {
                sflag = -one;
                sh11 = zero; sh12 = zero; sh21 = zero; sh22 = zero;
                *sd1 = zero; *sd2 = zero; *sx1 = zero;
            } else {
                sflag = one;
                sh11 = sp1 / sp2;
                sh22 = *sx1 / sy1;
                su = one + sh11 * sh22;
                stemp = *sd2 / su;
                *sd2 = *sd1 / su;
                *sd1 = stemp;
                *sx1 = sy1 * su;
            }
        }
        if (*sd1 != zero) {
            guard = 0;
            while ((*sd1 <= rgamsq || *sd1 >= gamsq) && guard++ < 4096) {
                if (sflag == zero) { sh11 = one; sh22 = one; sflag = -one; }
                else               { sh21 = -one; sh12 = one; sflag = -one; }
                if (*sd1 <= rgamsq) { *sd1 = *sd1 * (gam * gam); *sx1 = *sx1 / gam; sh11 = sh11 / gam; sh12 = sh12 / gam; }
                else                { *sd1 = *sd1 / (gam * gam); *sx1 = *sx1 * gam; sh11 = sh11 * gam; sh12 = sh12 * gam; }
            }
        }
        if (*sd2 != zero) {
            guard = 0;
            while ((fabsf(*sd2) <= rgamsq || fabsf(*sd2) >= gamsq) && guard++ < 4096) {
                if (sflag == zero) { sh11 = one; sh22 = one; sflag = -one; }
                else               { sh21 = -one; sh12 = one; sflag = -one; }
                if (fabsf(*sd2) <= rgamsq) { *sd2 = *sd2 * (gam * gam); sh21 = sh21 / gam; sh22 = sh22 / gam; }
                else                       { *sd2 = *sd2 / (gam * gam); sh21 = sh21 * gam; sh22 = sh22 * gam; }
            }
        }
    }
    if (sflag < zero) { sparam[1] = sh11; sparam[2] = sh21; sparam[3] = sh12; sparam[4] = sh22; }
    else if (sflag == zero) { sparam[2] = sh21; sparam[3] = sh12; }
    else { sparam[1] = sh11; sparam[4] = sh22; }
    sparam[0] = sflag;
}
