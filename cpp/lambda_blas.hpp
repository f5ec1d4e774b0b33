// lambda_blas.hpp -- header-only C++ mirror of Numerical.BLAS.Single over the C ABI (include/lbk.h).
//
// The reference is compiled Haskell and no GHC exists in this image, so the compiled host layer above
// the C ABI is C++ (the Haskell module of the same shape is hs/Numerical/BLAS/Device.hs).  Names,
// argument order and results follow src/Numerical/BLAS/Single.hs:84,113,129,147,165,201,222,421,441,484:
// the vector routines are PURE (they return new device vectors; Util.hs:134-142), degenerate arguments
// give the reference's sentinels, and anything the reference would turn into undefined behaviour
// (a span outside the vector) throws lbk::Error instead.
#pragma once

#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../include/lbk.h"

namespace lambda_blas {

struct Error : std::runtime_error {
    int status;
    Error(int s, const std::string &m) : std::runtime_error(m), status(s) {}
};

class Context {
  public:
    explicit Context(int device = 0)
    {
        if (int s = lbk_init(device, &c_)) throw Error(s, lbk_last_error(nullptr));
    }
    ~Context() { lbk_shutdown(c_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    lbk_ctx *get() const { return c_; }
    void check(int s) const { if (s) throw Error(s, lbk_last_error(c_)); }
    void sync() const { check(lbk_sync(c_)); }

  private:
    lbk_ctx *c_ = nullptr;
};

// Device-resident Vector Float (the representation added beside Numerical.BLAS.Types).
class DVector {
  public:
    DVector(const Context &ctx, int64_t n) : ctx_(&ctx), n_(n)
    {
        lbk_vec *v = nullptr;
        ctx.check(lbk_vec_alloc(ctx.get(), n, &v));
        v_.reset(v, [](lbk_vec *p) { lbk_vec_free(p); });
    }
    DVector(const Context &ctx, const std::vector<float> &host) : DVector(ctx, (int64_t)host.size())
    {
        ctx.check(lbk_vec_upload(v_.get(), host.data(), n_));       // explicit host -> device
    }
    std::vector<float> to_host() const                               // explicit device -> host
    {
        std::vector<float> h((size_t)n_);
        ctx_->check(lbk_vec_download(v_.get(), h.data(), n_));
        return h;
    }
    int64_t size() const { return n_; }
    lbk_vec *get() const { return v_.get(); }
    const Context &ctx() const { return *ctx_; }

  private:
    const Context *ctx_;
    int64_t n_;
    std::shared_ptr<lbk_vec> v_;
};

struct GivensRot { float r, z, c, s; };                              // Types.hs:8
struct ModGivensRot {                                                // Types.hs:15-20; flag names the constructor
    int flag = -2;                                                   // -2 FLAGNEG2, -1 FLAGNEG1, 0 FLAG0, 1 FLAG1
    float d1 = 0, d2 = 0, x1 = 0, h11 = 0, h12 = 0, h21 = 0, h22 = 0;
    void sparam(float p[5]) const                                    // test/Foreign.hs:276-280
    {
        const float t[4][5] = {{-2, 1, 0, 0, 1}, {-1, h11, h21, h12, h22}, {0, 0, h21, h12, 0}, {1, h11, 0, 0, h22}};
        for (int i = 0; i < 5; ++i) p[i] = t[flag + 2][i];
    }
};

// ---- products and norms ----------------------------------------------------------------------------
inline float sdot(int64_t n, const DVector &x, int64_t incx, const DVector &y, int64_t incy)
{ float o; x.ctx().check(lbk_sdot(x.ctx().get(), n, x.get(), incx, y.get(), incy, &o)); return o; }
inline float sdsdot(int64_t n, float sb, const DVector &x, int64_t incx, const DVector &y, int64_t incy)
{ float o; x.ctx().check(lbk_sdsdot(x.ctx().get(), n, sb, x.get(), incx, y.get(), incy, &o)); return o; }
inline float sasum(int64_t n, const DVector &x, int64_t incx)
{ float o; x.ctx().check(lbk_sasum(x.ctx().get(), n, x.get(), incx, &o)); return o; }
inline float snrm2(int64_t n, const DVector &x, int64_t incx)
{ float o; x.ctx().check(lbk_snrm2(x.ctx().get(), n, x.get(), incx, &o)); return o; }
inline int64_t isamax(int64_t n, const DVector &x, int64_t incx)
{ int64_t o; x.ctx().check(lbk_isamax(x.ctx().get(), n, x.get(), incx, &o)); return o; }

// ---- pure vector results -----------------------------------------------------------------------------
inline DVector sscal(int64_t n, float a, const DVector &x, int64_t incx)
{ DVector o(x.ctx(), x.size()); x.ctx().check(lbk_sscal_oop(x.ctx().get(), n, a, x.get(), incx, o.get())); return o; }
inline DVector scopy(int64_t n, const DVector &x, int64_t incx, const DVector &y, int64_t incy)
{ DVector o(y.ctx(), y.size()); x.ctx().check(lbk_scopy_oop(x.ctx().get(), n, x.get(), incx, y.get(), incy, o.get())); return o; }
inline DVector saxpy(int64_t n, float a, const DVector &x, int64_t incx, const DVector &y, int64_t incy)
{ DVector o(y.ctx(), y.size()); x.ctx().check(lbk_saxpy_oop(x.ctx().get(), n, a, x.get(), incx, y.get(), incy, o.get())); return o; }
inline std::pair<DVector, DVector> sswap(int64_t n, const DVector &x, int64_t incx, const DVector &y, int64_t incy)
{
    DVector xo(x.ctx(), x.size()), yo(y.ctx(), y.size());
    x.ctx().check(lbk_sswap_oop(x.ctx().get(), n, x.get(), incx, y.get(), incy, xo.get(), yo.get()));
    return {xo, yo};
}
inline std::pair<DVector, DVector> srot(int64_t n, const DVector &x, int64_t incx, const DVector &y, int64_t incy, float c, float s)
{
    DVector xo(x.ctx(), x.size()), yo(y.ctx(), y.size());
    x.ctx().check(lbk_srot_oop(x.ctx().get(), n, x.get(), incx, y.get(), incy, c, s, xo.get(), yo.get()));
    return {xo, yo};
}
inline std::pair<DVector, DVector> srotm(const ModGivensRot &flags, int64_t n, const DVector &x, int64_t incx, const DVector &y, int64_t incy)
{
    if (flags.flag == -2) return {x, y};                              // Single.hs:473: the arguments themselves
    float p[5];
    flags.sparam(p);
    DVector xo(x.ctx(), x.size()), yo(y.ctx(), y.size());
    x.ctx().check(lbk_srotm_oop(x.ctx().get(), n, x.get(), incx, y.get(), incy, p, xo.get(), yo.get()));
    return {xo, yo};
}

// ---- host scalars ---------------------------------------------------------------------------------------
inline GivensRot srotg(float a, float b)
{ float o[4]; lbk_srotg(a, b, o); return {o[0], o[1], o[2], o[3]}; }
inline ModGivensRot srotmg(float sd1, float sd2, float sx1, float sy1)
{
    float o[8];
    lbk_srotmg(sd1, sd2, sx1, sy1, o, nullptr);
    ModGivensRot m;
    m.flag = (int)o[0];
    if (m.flag != -2) { m.d1 = o[1]; m.d2 = o[2]; m.x1 = o[3]; m.h11 = o[4]; m.h12 = o[5]; m.h21 = o[6]; m.h22 = o[7]; }
    return m;
}

}  // namespace lambda_blas
