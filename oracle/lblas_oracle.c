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
 * PARITY PINNED TO THE REFERENCE'S SOURCE, EVALUATED (not to a GHC build).  The
 * reference ships no golden vectors, no known-answer tests and no fixtures for
 * this path (its tests are unseeded QuickCheck properties against Fortran netlib
 * BLAS 3.7.0, test/Main.hs:28-52), and neither GHC nor gfortran exist in this
 * image, so the reference cannot be COMPILED here.  Its three hot-path modules
 * are therefore evaluated from their text, where they lie under /root/reference,
 * by oracle/hs_eval.py (a Haskell-subset evaluator: lexer, layout rule, parser,
 * call-by-need evaluation; the primitives of base-4.9 / vector-0.11 it relies on
 * are restated there and listed in its header).  What pins this file:
 *   - tests/golden/ref_golden_{f32,f64}.json, the outputs of that evaluation over
 *     the reference's test domains (generator: tests/golden/make_ref_golden.py):
 *     this file must reproduce them BIT FOR BIT, reductions included
 *     (tests/test_golden.py), here and on the GPU box; in the build container
 *     tests/test_hs_eval.py re-evaluates the source against the fixtures and
 *     against this file on fresh random inputs;
 *   - the one worked example the reference documents (README.md:57-59, = 32);
 *   - the reference's own differential property, re-stated: this file must
 *     equal, bit for bit, oracle/netlib_l1.c (a restatement of the published
 *     netlib BLAS 3.7.0 routines the reference's tests bind in
 *     test/Foreign.hs:32-81) over the input domains of test/Main.hs;
 *   - a real compiled BLAS behind the same Fortran interface the reference's
 *     tests bind: scipy's OpenBLAS (tests/test_oracle_openblas.py, run here and
 *     on the GPU box) -- bit for bit for copy/swap/scal/iamax, within one
 *     rounding per term for axpy/rot/rotm, within the reduction bound for
 *     dot/sdsdot/asum/nrm2 (OpenBLAS vectorises sums and contracts FMAs, so it is
 *     not netlib; it is an independent implementation, not a restatement).
 * Not exercised by any of this: GHC's code generator itself (the evaluator
 * performs one correctly rounded IEEE operation per arithmetic primitive, which
 * is what GHC 8.0's native code generator emits for Float / Double on x86-64).
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


/* ---- Float instance: lbo_sdot, lbo_isamax, ... ---- */
#define REAL float
#define FN(n) lbo_s##n
#define FNI(n) lbo_is##n
#define GEN(n) n##_f32
#define FABS(v) fabsf(v)
#define SQRT(v) sqrtf(v)
#define LIT(v) v##f
#include "lblas_body.inc"
#undef REAL
#undef FN
#undef FNI
#undef GEN
#undef FABS
#undef SQRT
#undef LIT

/* ---- Double instance: lbo_ddot, lbo_idamax, ... ---- */
#define REAL double
#define FN(n) lbo_d##n
#define FNI(n) lbo_id##n
#define GEN(n) n##_f64
#define FABS(v) fabs(v)
#define SQRT(v) sqrt(v)
#define LIT(v) v
#include "lblas_body.inc"

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

