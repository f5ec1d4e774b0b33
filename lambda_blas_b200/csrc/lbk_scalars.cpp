// lbk_scalars.cpp -- the O(1) host routines of the path: rotg and rotmg (Float and Double instances).
//
// They stay on the CPU, as in the reference (src/Numerical/BLAS/Single.hs:274-300, 318-401):
// rotmg is how callers build the ModGivensRot handed to rotm (test/Main.hs:214,
// test/Bench.hs:41).  Compiled with -ffp-contract=off: every operation below is one IEEE
// operation in the routine's own precision, in the order the reference writes it.
#include "../../include/lbk.h"

#include <cmath>

namespace {

template <class R>
struct Params {                       // the 8-tuple threaded through checkscale (Single.hs:363-364)
    int flag;
    R d1, d2, x1, h11, h12, h21, h22;
};

// Single.hs:367-369: the reference is polymorphic, so BOTH instances use these literals (rounded to the
// instance's precision) -- unlike netlib, whose drotmg carries more digits (16777216, 5.9604645e-8).
template <class R> constexpr R gam() { return (R)4096.0; }
template <class R> constexpr R gamsq() { return (R)1.67772E7; }
template <class R> constexpr R rgamsq() { return (R)5.96046E-8; }
const int LOOP_CAP = 4096;            // the reference does not terminate on NaN; we do

// Single.hs:398-401: the first rescale turns any rotation into the general (flag -1) form
template <class R> inline void to_general_form(Params<R> &p)
{
    if (p.flag == 0) { p.h11 = (R)1; p.h22 = (R)1; }
    else { p.h12 = (R)1; p.h21 = (R)-1; }
    p.flag = -1;
}

// Single.hs:371-383: bring d1 into (rgamsq, gamsq)
template <class R> void rescale_d1(Params<R> &p)
{
    const R g = gam<R>(), g2 = g * g;
    if (p.d1 == (R)0) return;
    for (int k = 0; k < LOOP_CAP; ++k) {
        if (p.d1 > rgamsq<R>() && p.d1 < gamsq<R>()) return;
        to_general_form(p);
        if (std::fabs(p.d1) <= rgamsq<R>()) { p.d1 *= g2; p.x1 /= g; p.h11 /= g; p.h12 /= g; }
        else { p.d1 /= g2; p.x1 *= g; p.h11 *= g; p.h12 *= g; }
    }
}

// Single.hs:385-397: bring |d2| into (rgamsq, gamsq)
template <class R> void rescale_d2(Params<R> &p)
{
    const R g = gam<R>(), g2 = g * g;
    if (p.d2 == (R)0) return;
    for (int k = 0; k < LOOP_CAP; ++k) {
        const R mag = std::fabs(p.d2);
        if (mag > rgamsq<R>() && mag < gamsq<R>()) return;
        to_general_form(p);
        if (mag <= rgamsq<R>()) { p.d2 *= g2; p.h21 /= g; p.h22 /= g; }
        else { p.d2 /= g2; p.h21 *= g; p.h22 *= g; }
    }
}

template <class R> inline R sign_of(R v) { return v > (R)0 ? (R)1 : (v < (R)0 ? (R)-1 : v); }   // Haskell signum

template <class R> int rotg(R sa, R sb, R *out)
{
    if (!out) return LBK_ERR_ARG;
    const R mag_a = std::fabs(sa), mag_b = std::fabs(sb);
    const R scale = mag_a + mag_b;                           // Single.hs:292-294
    if (scale == (R)0) { out[0] = 0; out[1] = 0; out[2] = 1; out[3] = 0; return LBK_OK; }
    const R ua = sa / scale, ub = sb / scale;
    const R norm = scale * std::sqrt(ua * ua + ub * ub);     // Single.hs:295
    if (mag_a > mag_b) {                                     // Single.hs:283-285
        const R r = sign_of(sa) * norm;
        const R s = sb / r;
        out[0] = r; out[1] = s; out[2] = sa / r; out[3] = s;
    } else {                                                 // Single.hs:286-290
        const R r = sign_of(sb) * norm;
        const R c = sa / r;
        out[0] = r;
        if (c == (R)0) { out[1] = 1; out[2] = 0; out[3] = 1; }
        else { out[1] = (R)1 / c; out[2] = c; out[3] = sb / r; }
    }
    return LBK_OK;
}

template <class R> int rotmg(R sd1, R sd2, R sx1, R sy1, R *out, R *param)
{
    if (!out) return LBK_ERR_ARG;
    Params<R> p = {0, 0, 0, 0, 0, 0, 0, 0};
    if (sd1 < (R)0) {                                        // Single.hs:325
        p.flag = -1;
    } else {
        const R p2 = sd2 * sy1;
        if (p2 == (R)0) {                                    // Single.hs:326
            p.flag = -2;
        } else {
            const R p1 = sd1 * sx1;
            const R q2 = p2 * sy1, q1 = p1 * sx1;
            if (std::fabs(q1) > std::fabs(q2)) {             // Single.hs:327-333
                p.flag = 0;
                p.h21 = -(sy1 / sx1);
                p.h12 = p2 / p1;
                const R prod = p.h12 * p.h21;
                const R u = (R)1 - prod;
                if (u > (R)0) { p.d1 = sd1 / u; p.d2 = sd2 / u; p.x1 = sx1 * u; }
                else { p.d1 = sd1; p.d2 = sd2; p.x1 = sx1; }
            } else if (q2 < (R)0) {                          // Single.hs:334-335
                p.flag = -1;
            } else {                                         // Single.hs:336-339
                p.flag = 1;
                p.h11 = p1 / p2;
                p.h22 = sx1 / sy1;
                const R prod = p.h11 * p.h22;
                const R u = (R)1 + prod;
                p.d1 = sd2 / u; p.d2 = sd1 / u; p.x1 = sy1 * u;
            }
            rescale_d1(p);                                   // checkscale = checkscale2 . checkscale1
            rescale_d2(p);
        }
    }
    out[0] = (R)p.flag;
    out[1] = p.d1; out[2] = p.d2; out[3] = p.x1;
    out[4] = p.h11; out[5] = p.h12; out[6] = p.h21; out[7] = p.h22;
    if (param) {                                             // the sparam array test/Foreign.hs:276-280 pokes
        switch (p.flag) {
        case -2: param[0] = -2; param[1] = 1; param[2] = 0; param[3] = 0; param[4] = 1; break;
        case -1: param[0] = -1; param[1] = p.h11; param[2] = p.h21; param[3] = p.h12; param[4] = p.h22; break;
        case 0:  param[0] = 0; param[1] = 0; param[2] = p.h21; param[3] = p.h12; param[4] = 0; break;
        default: param[0] = 1; param[1] = p.h11; param[2] = 0; param[3] = 0; param[4] = p.h22; break;
        }
    }
    return LBK_OK;
}

}  // namespace

extern "C" int lbk_srotg(float sa, float sb, float out[4]) { return rotg<float>(sa, sb, out); }
extern "C" int lbk_drotg(double sa, double sb, double out[4]) { return rotg<double>(sa, sb, out); }
extern "C" int lbk_srotmg(float sd1, float sd2, float sx1, float sy1, float out[8], float param[5])
{ return rotmg<float>(sd1, sd2, sx1, sy1, out, param); }
extern "C" int lbk_drotmg(double sd1, double sd2, double sx1, double sy1, double out[8], double param[5])
{ return rotmg<double>(sd1, sd2, sx1, sy1, out, param); }
