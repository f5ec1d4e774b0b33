// lbk_scalars.cpp -- the O(1) host routines of the path: srotg and srotmg.
//
// They stay on the CPU, as in the reference (src/Numerical/BLAS/Single.hs:274-300, 318-401):
// srotmg is how callers build the ModGivensRot handed to srotm (test/Main.hs:214,
// test/Bench.hs:41).  Compiled with -ffp-contract=off: every operation below is one IEEE
// binary32 operation, in the order the reference writes it.
#include "../../include/lbk.h"

#include <cmath>

namespace {

struct Params {                       // the 8-tuple threaded through checkscale (Single.hs:363-364)
    int flag;
    float d1, d2, x1, h11, h12, h21, h22;
};

const float GAM = 4096.0f;            // Single.hs:367-369
const float GAMSQ = 1.67772E7f;
const float RGAMSQ = 5.96046E-8f;
const int LOOP_CAP = 4096;            // the reference does not terminate on NaN; we do

// Single.hs:398-401: the first rescale turns any rotation into the general (flag -1) form
inline void to_general_form(Params &p)
{
    if (p.flag == 0) { p.h11 = 1.0f; p.h22 = 1.0f; }
    else { p.h12 = 1.0f; p.h21 = -1.0f; }
    p.flag = -1;
}

// Single.hs:371-383: bring d1 into (rgamsq, gamsq)
void rescale_d1(Params &p)
{
    const float g2 = GAM * GAM;
    if (p.d1 == 0.0f) return;
    for (int k = 0; k < LOOP_CAP; ++k) {
        if (p.d1 > RGAMSQ && p.d1 < GAMSQ) return;
        to_general_form(p);
        if (std::fabs(p.d1) <= RGAMSQ) { p.d1 *= g2; p.x1 /= GAM; p.h11 /= GAM; p.h12 /= GAM; }
        else { p.d1 /= g2; p.x1 *= GAM; p.h11 *= GAM; p.h12 *= GAM; }
    }
}

// Single.hs:385-397: bring |d2| into (rgamsq, gamsq)
void rescale_d2(Params &p)
{
    const float g2 = GAM * GAM;
    if (p.d2 == 0.0f) return;
    for (int k = 0; k < LOOP_CAP; ++k) {
        const float mag = std::fabs(p.d2);
        if (mag > RGAMSQ && mag < GAMSQ) return;
        to_general_form(p);
        if (mag <= RGAMSQ) { p.d2 *= g2; p.h21 /= GAM; p.h22 /= GAM; }
        else { p.d2 /= g2; p.h21 *= GAM; p.h22 *= GAM; }
    }
}

inline float sign_of(float v) { return v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : v); }   // Haskell signum

}  // namespace

extern "C" int lbk_srotg(float sa, float sb, float out[4])
{
    if (!out) return LBK_ERR_ARG;
    const float mag_a = std::fabs(sa), mag_b = std::fabs(sb);
    const float scale = mag_a + mag_b;                       // Single.hs:292-294
    if (scale == 0.0f) { out[0] = 0.0f; out[1] = 0.0f; out[2] = 1.0f; out[3] = 0.0f; return LBK_OK; }
    const float ua = sa / scale, ub = sb / scale;
    const float norm = scale * std::sqrt(ua * ua + ub * ub); // Single.hs:295
    if (mag_a > mag_b) {                                     // Single.hs:283-285
        const float r = sign_of(sa) * norm;
        const float s = sb / r;
        out[0] = r; out[1] = s; out[2] = sa / r; out[3] = s;
    } else {                                                 // Single.hs:286-290
        const float r = sign_of(sb) * norm;
        const float c = sa / r;
        out[0] = r;
        if (c == 0.0f) { out[1] = 1.0f; out[2] = 0.0f; out[3] = 1.0f; }
        else { out[1] = 1.0f / c; out[2] = c; out[3] = sb / r; }
    }
    return LBK_OK;
}

extern "C" int lbk_srotmg(float sd1, float sd2, float sx1, float sy1, float out[8], float param[5])
{
    if (!out) return LBK_ERR_ARG;
    Params p = {0, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    if (sd1 < 0.0f) {                                        // Single.hs:325
        p.flag = -1;
    } else {
        const float p2 = sd2 * sy1;
        if (p2 == 0.0f) {                                    // Single.hs:326
            p.flag = -2;
        } else {
            const float p1 = sd1 * sx1;
            const float q2 = p2 * sy1, q1 = p1 * sx1;
            if (std::fabs(q1) > std::fabs(q2)) {             // Single.hs:327-333
                p.flag = 0;
                p.h21 = -(sy1 / sx1);
                p.h12 = p2 / p1;
                const float prod = p.h12 * p.h21;
                const float u = 1.0f - prod;
                if (u > 0.0f) { p.d1 = sd1 / u; p.d2 = sd2 / u; p.x1 = sx1 * u; }
                else { p.d1 = sd1; p.d2 = sd2; p.x1 = sx1; }
            } else if (q2 < 0.0f) {                          // Single.hs:334-335
                p.flag = -1;
            } else {                                         // Single.hs:336-339
                p.flag = 1;
                p.h11 = p1 / p2;
                p.h22 = sx1 / sy1;
                const float prod = p.h11 * p.h22;
                const float u = 1.0f + prod;
                p.d1 = sd2 / u; p.d2 = sd1 / u; p.x1 = sy1 * u;
            }
            rescale_d1(p);                                   // checkscale = checkscale2 . checkscale1
            rescale_d2(p);
        }
    }
    out[0] = (float)p.flag;
    out[1] = p.d1; out[2] = p.d2; out[3] = p.x1;
    out[4] = p.h11; out[5] = p.h12; out[6] = p.h21; out[7] = p.h22;
    if (param) {                                             // the sparam array test/Foreign.hs:276-280 pokes
        switch (p.flag) {
        case -2: param[0] = -2.0f; param[1] = 1.0f; param[2] = 0.0f; param[3] = 0.0f; param[4] = 1.0f; break;
        case -1: param[0] = -1.0f; param[1] = p.h11; param[2] = p.h21; param[3] = p.h12; param[4] = p.h22; break;
        case 0:  param[0] = 0.0f; param[1] = 0.0f; param[2] = p.h21; param[3] = p.h12; param[4] = 0.0f; break;
        default: param[0] = 1.0f; param[1] = p.h11; param[2] = 0.0f; param[3] = 0.0f; param[4] = p.h22; break;
        }
    }
    return LBK_OK;
}
