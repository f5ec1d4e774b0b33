/*
 * oracle/lblas_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C (scalar, strictly sequential, IEEE binary32) restatement of the
 * single-precision Level-1 routines of the reference, dlewissandy/lambda-blas:
 *     src/Numerical/BLAS/Single.hs   (the routines)
 *     src/Numerical/BLAS/Util.hs     (the stride / update index semantics)
 * Each function cites the reference lines it follows.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this file's shared object, and only as the
 * CHECKER or as the reported CPU baseline.  Nothing under lambda_blas_b200/
 * may import, link or execute it.
 *
 * PARITY UNPINNED (strict sense): the reference ships no golden vectors, no
 * known-answer tests and no fixtures for this path (its tests are unseeded
 * QuickCheck properties against Fortran netlib BLAS 3.7.0, test/Main.hs:28-52),
 * and neither GHC nor gfortran exist in this image, so the reference itself
 * cannot be run here.  What pins this file instead (tests/test_oracle_*.py):
 *   - the one worked example the reference documents (README.md:57-59, = 32);
 *   - the reference's own differential property, re-stated: this file must
 *     equal, bit for bit, oracle/netlib_l1.c (a restatement of the published
 *     netlib BLAS 3.7.0 routines the reference's tests bind in
 *     test/Foreign.hs:32-81) over the input domains of test/Main.hs;
 *   - scipy's OpenBLAS for the order-independent routines.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (see oracle/Makefile); no
 * -march=native, so no FMA contraction and no vectorised reassociation: every
 * product is rounded to binary32 before it is added, exactly like the code
 * GHC 8.0's native code generator emits for Float.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

typedef int64_t i64;

/* Single.hs:91-96  firstIndex n inc */
static inline i64 first_index(i64 n, i64 inc) { return inc > 0 ? 0 : (1 - n) * inc; }

/* Util.hs:75-88  sampleElems n v inc, element i  ==  v[firstIndex + i*inc]
 * (inc==1 take, inc>1 sampleIndices, inc==0 replicate v[0], inc==-1 reverse,
 *  inc<-1 sampleIndicesRev). */
#define ELEM(v, n, inc, i) ((v)[first_index((n), (inc)) + (i) * (inc)])

/* ------------------------------------------------------------------ */
/* Reductions                                                           */
/* ------------------------------------------------------------------ */

/* Single.hs:81-82  dot = foldl' (+) 0 . zipWith (*) : strict left-to-right,
 * product rounded before the add. */
float lbo_sdot(i64 n, const float *x, i64 incx, const float *y, i64 incy)
{
    float acc = 0.0f;
    if (n <= 0) return acc;
    i64 kx = first_index(n, incx), ky = first_index(n, incy);
    for (i64 i = 0; i < n; ++i) {
        float p = x[kx] * y[ky];
        acc = acc + p;
        kx += incx;
        ky += incy;
    }
    return acc;
}

/* Single.hs:161-163  asum: 0 when incx < 1, else foldl' (+) 0 . map abs */
float lbo_sasum(i64 n, const float *x, i64 incx)
{
    float acc = 0.0f;
    if (incx < 1 || n <= 0) return acc;
    for (i64 i = 0; i < n; ++i) acc = acc + fabsf(x[i * incx]);
    return acc;
}

/* Single.hs:180-199  nrm2: guards, then the slassq recurrence on (scale,ssq).
 * `compare xi 0` is EQ -> skip, LT -> f(-xi), GT (also NaN) -> f(xi). */
float lbo_snrm2(i64 n, const float *x, i64 incx)
{
    if (n < 1 || incx < 1) return 0.0f;
    if (n == 1) return fabsf(x[0]);
    float scale = 0.0f, ssq = 1.0f;
    i64 imax = incx * (n - 1);
    for (i64 i = 0; i <= imax; i += incx) {
        float xi = x[i];
        if (xi == 0.0f) continue;               /* EQ */
        float a = (xi < 0.0f) ? -xi : xi;       /* LT -> negate ; GT/NaN -> id */
        if (scale < a) {
            float q = scale / a;
            float q2 = q * q;
            float t = ssq * q2;
            ssq = 1.0f + t;
            scale = a;
        } else {
            float q = a / scale;
            float q2 = q * q;
            ssq = ssq + q2;
        }
    }
    return scale * sqrtf(ssq);
}

/* Single.hs:217-220  iamax: -1 | 0 | maxIndexBy (comparing abs).
 * vector-0.11 maxIndexBy keeps (i,x) unless `compare |x| |y| == LT`, and
 * Haskell's compare on Float yields LT only when |x| < |y| is true, so the
 * first maximum wins, a NaN at index 0 is never displaced and a NaN elsewhere
 * never wins.  Result is the 0-based LOGICAL index. */
i64 lbo_isamax(i64 n, const float *x, i64 incx)
{
    if (n < 1 || incx < 1) return -1;
    if (n == 1) return 0;
    i64 best = 0;
    float bv = fabsf(x[0]);
    for (i64 i = 1; i < n; ++i) {
        float v = fabsf(x[i * incx]);
        if (bv < v) { best = i; bv = v; }
    }
    return best;
}

/* Single.hs:238-262  sdsdot: double accumulator seeded with sb. */
float lbo_sdsdot(i64 n, float sb, const float *x, i64 incx, const float *y, i64 incy)
{
    double c = (double)sb;
    if (n < 1) return (float)c;
    if (incx == incy && incx > 0) {           /* Single.hs:240,245-254 */
        i64 imax = n * incx;
        for (i64 i = 0; i < imax; i += incx) c = c + (double)x[i] * (double)y[i];
        return (float)c;
    }
    i64 kx = first_index(n, incx), ky = first_index(n, incy);   /* Single.hs:241,257-262 */
    for (i64 i = 0; i < n; ++i) {
        c = c + (double)x[kx] * (double)y[ky];
        kx += incx;
        ky += incy;
    }
    return (float)c;
}

/* ------------------------------------------------------------------ */
/* updateElems family (pure: result = copy of the whole destination,   */
/* then writes applied in stream order; right-hand sides read the      */
/* ORIGINAL vectors).  Util.hs:125-142                                  */
/* ------------------------------------------------------------------ */

enum { F_COPY = 0, F_AXPY, F_ROT_X, F_ROT_Y, F_ROTM_N1_X, F_ROTM_N1_Y,
       F_ROTM_0_X, F_ROTM_0_Y, F_ROTM_1_X, F_ROTM_1_Y };

/* the lambdas handed to updateElems; `d` is the destination's old element,
 * `s` the element of the other vector */
static inline float apply_f(int f, float d, float s, const float *k)
{
    float t1, t2;
    switch (f) {
    case F_COPY:      return s;                                   /* Single.hs:127,144-145 */
    case F_AXPY:      t1 = k[0] * s; return d + t1;               /* Single.hs:439  y + a*x */
    case F_ROT_X:     t1 = k[0] * d; t2 = k[1] * s; return t1 + t2; /* :418 c*x + s*y */
    case F_ROT_Y:     t1 = k[0] * d; t2 = k[1] * s; return t1 - t2; /* :419 c*y - s*x */
    case F_ROTM_N1_X: t1 = k[0] * d; t2 = k[1] * s; return t1 + t2; /* :475 h11*x + h12*y */
    case F_ROTM_N1_Y: t1 = k[0] * s; t2 = k[1] * d; return t1 + t2; /* :476 h21*x + h22*y */
    case F_ROTM_0_X:  t1 = k[0] * s; return d + t1;               /* :478 x + h12*y */
    case F_ROTM_0_Y:  t1 = k[0] * s; return t1 + d;               /* :479 h21*x + y */
    case F_ROTM_1_X:  t1 = k[0] * d; return t1 + s;               /* :481 h11*x + y */
    default:          t1 = k[0] * d; return (-s) + t1;            /* :482 -x + h22*y */
    }
}

/* out (length lend) = updateElems f n dst incd src incs */
static void update_elems(int f, const float *k, i64 n,
                         const float *dst, i64 lend, i64 incd,
                         const float *src, i64 incs, float *out)
{
    if (out != dst) memcpy(out, dst, (size_t)lend * sizeof(float));  /* unsafeUpdate_ clones */
    if (n <= 0) return;                                              /* Util.hs:136 */
    i64 kd = first_index(n, incd), ks = first_index(n, incs);
    if (incd == 0) {
        /* index stream is all zeros: n sequential writes to slot 0, each from
         * the ORIGINAL dst[0]; the last one wins */
        float d0 = dst[0], last = d0;
        for (i64 i = 0; i < n; ++i) { last = apply_f(f, d0, src[ks], k); ks += incs; }
        out[0] = last;
        return;
    }
    if (incs == 0) {
        float s0 = src[0];
        for (i64 i = 0; i < n; ++i) { out[kd] = apply_f(f, dst[kd], s0, k); kd += incd; }
        return;
    }
    for (i64 i = 0; i < n; ++i) {
        out[kd] = apply_f(f, dst[kd], src[ks], k);
        kd += incd;
        ks += incs;
    }
}

/* NOTE on aliasing: `out` must not alias `src`; it may equal `dst` only when
 * the caller accepts in-place update (used by the timing leg, never by the
 * parity tests, which always take the pure form). */

/* Single.hs:108-111  scal = unsafeUpdate_ sx (sampleIndices n incx) (map (*a) (sampleElems n sx incx))
 * incx>=1: n strided products.  incx==0: every write is sx[0]*a at slot 0.
 * incx<0: sampleIndices ends at once for n>1 (imax<0) -> unchanged; for n==1
 * one write of sx[0]*a. */
void lbo_sscal(i64 n, float a, const float *x, i64 lenx, i64 incx, float *out)
{
    if (out != x) memcpy(out, x, (size_t)lenx * sizeof(float));
    if (n <= 0) return;
    if (incx > 0) {
        for (i64 i = 0; i < n; ++i) out[i * incx] = x[i * incx] * a;
    } else if (incx == 0 || n == 1) {
        out[0] = x[0] * a;
    }
}

/* Single.hs:127  copy n sx incx sy incy = updateElems (\_ x -> x) n sy incy sx incx */
void lbo_scopy(i64 n, const float *x, i64 incx, const float *y, i64 leny, i64 incy, float *yout)
{
    update_elems(F_COPY, 0, n, y, leny, incy, x, incx, yout);
}

/* Single.hs:143-145  swap: two updateElems, each from the originals */
void lbo_sswap(i64 n, const float *x, i64 lenx, i64 incx, const float *y, i64 leny, i64 incy,
               float *xout, float *yout)
{
    update_elems(F_COPY, 0, n, x, lenx, incx, y, incy, xout);
    update_elems(F_COPY, 0, n, y, leny, incy, x, incx, yout);
}

/* Single.hs:439  axpy n a sx incx sy incy = updateElems (\y x -> y + a*x) n sy incy sx incx
 * (no a==0 early-out) */
void lbo_saxpy(i64 n, float a, const float *x, i64 incx, const float *y, i64 leny, i64 incy,
               float *yout)
{
    float k[2] = { a, 0.0f };
    update_elems(F_AXPY, k, n, y, leny, incy, x, incx, yout);
}

/* Single.hs:417-419  rot */
void lbo_srot(i64 n, const float *x, i64 lenx, i64 incx, const float *y, i64 leny, i64 incy,
              float c, float s, float *xout, float *yout)
{
    float k[2] = { c, s };
    update_elems(F_ROT_X, k, n, x, lenx, incx, y, incy, xout);
    update_elems(F_ROT_Y, k, n, y, leny, incy, x, incx, yout);
}

/* Single.hs:472-482  rotm.  param = BLAS sparam [flag,h11,h21,h12,h22]
 * (test/Foreign.hs:248,277-280).  flag -2 returns the inputs as they are. */
void lbo_srotm(const float *param, i64 n, const float *x, i64 lenx, i64 incx,
               const float *y, i64 leny, i64 incy, float *xout, float *yout)
{
    float flag = param[0], h11 = param[1], h21 = param[2], h12 = param[3], h22 = param[4];
    float kx[2], ky[2];
    if (flag == -2.0f) {
        if (xout != x) memcpy(xout, x, (size_t)lenx * sizeof(float));
        if (yout != y) memcpy(yout, y, (size_t)leny * sizeof(float));
        return;
    }
    if (flag < 0.0f) {
        kx[0] = h11; kx[1] = h12; ky[0] = h21; ky[1] = h22;
        update_elems(F_ROTM_N1_X, kx, n, x, lenx, incx, y, incy, xout);
        update_elems(F_ROTM_N1_Y, ky, n, y, leny, incy, x, incx, yout);
    } else if (flag == 0.0f) {
        kx[0] = h12; kx[1] = 0; ky[0] = h21; ky[1] = 0;
        update_elems(F_ROTM_0_X, kx, n, x, lenx, incx, y, incy, xout);
        update_elems(F_ROTM_0_Y, ky, n, y, leny, incy, x, incx, yout);
    } else {
        kx[0] = h11; kx[1] = 0; ky[0] = h22; ky[1] = 0;
        update_elems(F_ROTM_1_X, kx, n, x, lenx, incx, y, incy, xout);
        update_elems(F_ROTM_1_Y, ky, n, y, leny, incy, x, incx, yout);
    }
}

/* ------------------------------------------------------------------ */
/* Host scalars                                                         */
/* ------------------------------------------------------------------ */

static inline float signumf(float v) { return v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : v); }

/* Single.hs:279-295  rotg -> out = (r, z, c, s)  (Types.hs:8) */
void lbo_srotg(float sa, float sb, float *out)
{
    float asa = fabsf(sa), asb = fabsf(sb);
    float scale = asa + asb;
    if (scale == 0.0f) { out[0] = 0; out[1] = 0; out[2] = 1; out[3] = 0; return; }
    float qa = sa / scale, qb = sb / scale;
    float qa2 = qa * qa, qb2 = qb * qb;
    float magnitude = scale * sqrtf(qa2 + qb2);
    if (asa > asb) {
        float r = signumf(sa) * magnitude;
        float s = sb / r;
        out[0] = r; out[1] = s; out[2] = sa / r; out[3] = s;
    } else {
        float r = signumf(sb) * magnitude;
        float c = sa / r;
        if (c == 0.0f) { out[0] = r; out[1] = 1; out[2] = 0; out[3] = 1; }
        else { out[0] = r; out[1] = 1.0f / c; out[2] = c; out[3] = sb / r; }
    }
}

typedef struct { int f; float d1, d2, x1, h11, h12, h21, h22; } rotmg_t;

#define LBO_SCALE_LOOP_CAP 4096   /* the reference loops forever on NaN / negative d1; we stop */

/* Single.hs:362-401  checkscale = checkscale2 . checkscale1 */
static void checkscale(rotmg_t *z)
{
    const float gam = 4096.0f, gamsq = 1.67772E7f, rgamsq = 5.96046E-8f;
    const float gam2 = gam * gam;
    int it;
    if (z->d1 != 0.0f) {
        for (it = 0; it < LBO_SCALE_LOOP_CAP && !(z->d1 > rgamsq && z->d1 < gamsq); ++it) {
            /* checkA :398-401 */
            if (z->f == 0) { z->h11 = 1.0f; z->h22 = 1.0f; } else { z->h12 = 1.0f; z->h21 = -1.0f; }
            z->f = -1;
            /* check1B :380-383 */
            if (fabsf(z->d1) <= rgamsq) { z->d1 = z->d1 * gam2; z->x1 = z->x1 / gam; z->h11 = z->h11 / gam; z->h12 = z->h12 / gam; }
            else                        { z->d1 = z->d1 / gam2; z->x1 = z->x1 * gam; z->h11 = z->h11 * gam; z->h12 = z->h12 * gam; }
        }
    }
    if (z->d2 != 0.0f) {
        for (it = 0; it < LBO_SCALE_LOOP_CAP; ++it) {
            float a = fabsf(z->d2);
            if (a > rgamsq && a < gamsq) break;
            if (z->f == 0) { z->h11 = 1.0f; z->h22 = 1.0f; } else { z->h12 = 1.0f; z->h21 = -1.0f; }
            z->f = -1;
            /* check2B :395-397 */
            if (a <= rgamsq) { z->d2 = z->d2 * gam2; z->h21 = z->h21 / gam; z->h22 = z->h22 / gam; }
            else             { z->d2 = z->d2 / gam2; z->h21 = z->h21 * gam; z->h22 = z->h22 * gam; }
        }
    }
}

/* Single.hs:324-352  rotmg.  out = [flag, d1, d2, x1, h11, h12, h21, h22];
 * flag in {-2,-1,0,1}; for -2 (FLAGNEG2, which carries no fields) the other
 * seven slots are zero. */
void lbo_srotmg(float sd1, float sd2, float sx1, float sy1, float *out)
{
    rotmg_t z; memset(&z, 0, sizeof z);
    if (sd1 < 0.0f) { z.f = -1; goto done; }                 /* :325 */
    {
        float sp2 = sd2 * sy1;
        if (sp2 == 0.0f) { z.f = -2; goto done; }            /* :326 */
        float sp1 = sd1 * sx1;
        float sq2 = sp2 * sy1;
        float sq1 = sp1 * sx1;
        if (fabsf(sq1) > fabsf(sq2)) {                       /* :327-333 */
            float sh21 = -(sy1 / sx1);
            float sh12 = sp2 / sp1;
            float t = sh12 * sh21;
            float su = 1.0f - t;
            z.f = 0; z.h12 = sh12; z.h21 = sh21;
            if (su > 0.0f) { z.d1 = sd1 / su; z.d2 = sd2 / su; z.x1 = sx1 * su; }
            else           { z.d1 = sd1;      z.d2 = sd2;      z.x1 = sx1; }
            checkscale(&z);
        } else if (sq2 < 0.0f) {                             /* :334-335 */
            z.f = -1;
            checkscale(&z);
        } else {                                             /* :336-339 */
            float sh11 = sp1 / sp2;
            float sh22 = sx1 / sy1;
            float t = sh11 * sh22;
            float su = 1.0f + t;
            z.f = 1; z.d1 = sd2 / su; z.d2 = sd1 / su; z.x1 = sy1 * su; z.h11 = sh11; z.h22 = sh22;
            checkscale(&z);
        }
    }
done:
    out[0] = (float)z.f; out[1] = z.d1; out[2] = z.d2; out[3] = z.x1;
    out[4] = z.h11; out[5] = z.h12; out[6] = z.h21; out[7] = z.h22;
}
