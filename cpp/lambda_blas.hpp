// lambda_blas.hpp -- header-only C++ mirror of Numerical.BLAS.Single over the C ABI (include/lbk.h):
// the polymorphic routines (dot, axpy, ...) as templates on the element type, plus the s*/d* aliases.
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
    // ONE host thread, many GPUs (lbk_init_multi): rank r is devices[r]; DVectorT::sharded(...) vectors are block
    // shards over the ranks and every routine below fans out over them.
    explicit Context(const std::vector<int> &devices)
    {
        if (int s = lbk_init_multi((int)devices.size(), devices.data(), &c_)) throw Error(s, lbk_last_error(nullptr));
    }
    // vectors that outlive the context keep the native context alive until they are freed (include/lbk.h)
    ~Context() { lbk_shutdown(c_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    lbk_ctx *get() const { return c_; }
    void check(int s) const { if (s) throw Error(s, lbk_last_error(c_)); }
    void sync() const { check(lbk_sync(c_)); }

  private:
    lbk_ctx *c_ = nullptr;
};

// A recorded sequence of calls (lbk_capture_begin / lbk_capture_end): everything the callable enqueues on the context is
// recorded into a CUDA graph instead of running; launch() replays it with one host call.  Only calls that neither allocate
// nor return a scalar to the host can be recorded (the in-place / *_oop routines of the C ABI, the *_async reductions).
class Graph {
  public:
    template <class F> Graph(const Context &ctx, F &&record) : ctx_(&ctx)
    {
        ctx.check(lbk_capture_begin(ctx.get()));
        try { record(); } catch (...) { lbk_graph *g = nullptr; if (!lbk_capture_end(ctx.get(), &g)) lbk_graph_destroy(g); throw; }
        ctx.check(lbk_capture_end(ctx.get(), &g_));
    }
    ~Graph() { lbk_graph_destroy(g_); }
    Graph(const Graph &) = delete;
    Graph &operator=(const Graph &) = delete;
    void launch() const { ctx_->check(lbk_graph_launch(g_)); }
    int64_t kernels() const { int64_t k = 0; ctx_->check(lbk_graph_kernels(g_, &k)); return k; }

  private:
    const Context *ctx_;
    lbk_graph *g_ = nullptr;
};

// The entry points of one instance (Float: lbk_s*, Double: lbk_d*) -- what the reference's type class
// resolves for `dot :: (Num a, Storable a) => ...` (Single.hs:73) at Float and at Double.
template <class T> struct Abi;
#define LB_ABI(T, P, IP, ALLOC)                                                                     \
    template <> struct Abi<T> {                                                                     \
        static int alloc(lbk_ctx *c, int64_t n, lbk_vec **v) { return ALLOC(c, n, v); }             \
        static constexpr auto dot = lbk_##P##dot;     static constexpr auto asum = lbk_##P##asum;   \
        static constexpr auto nrm2 = lbk_##P##nrm2;   static constexpr auto iamax = lbk_##IP##amax; \
        static constexpr auto axpy = lbk_##P##axpy;   static constexpr auto axpy_oop = lbk_##P##axpy_oop; \
        static constexpr auto scal_oop = lbk_##P##scal_oop; static constexpr auto copy_oop = lbk_##P##copy_oop; \
        static constexpr auto swap_oop = lbk_##P##swap_oop; static constexpr auto rot_oop = lbk_##P##rot_oop;   \
        static constexpr auto rotm_oop = lbk_##P##rotm_oop; static constexpr auto rotg = lbk_##P##rotg;         \
        static constexpr auto rotmg = lbk_##P##rotmg;                                                \
    };
LB_ABI(float, s, is, lbk_vec_alloc)
LB_ABI(double, d, id, lbk_vec_alloc_f64)
#undef LB_ABI

// Device-resident Vector a (the representation added beside Numerical.BLAS.Types).
template <class T> class DVectorT {
  public:
    DVectorT(const Context &ctx, int64_t n) : ctx_(&ctx), n_(n)
    {
        lbk_vec *v = nullptr;
        ctx.check(Abi<T>::alloc(ctx.get(), n, &v));
        v_.reset(v, [](lbk_vec *p) { lbk_vec_free(p); });
    }
    DVectorT(const Context &ctx, const std::vector<T> &host) : DVectorT(ctx, (int64_t)host.size())
    {
        ctx.check(lbk_vec_upload(v_.get(), host.data(), n_));       // explicit host -> device
    }
    // toDeviceSharded: ONE host buffer scattered over the ranks of a multi-device context (each rank uploads its own
    // block on its own stream); on a one-device context it is the same as the constructor above.
    static DVectorT sharded(const Context &ctx, const std::vector<T> &host)
    {
        DVectorT v(ctx, adopt(ctx, (int64_t)host.size(), true), (int64_t)host.size());
        ctx.check(lbk_vec_upload(v.get(), host.data(), v.size()));
        return v;
    }
    static DVectorT sharded(const Context &ctx, int64_t n) { return DVectorT(ctx, adopt(ctx, n, true), n); }
    // an uninitialised vector with this one's length, element type and shard layout: what a pure result is made of
    DVectorT like() const
    {
        lbk_vec *v = nullptr;
        ctx_->check(lbk_vec_alloc_like(v_.get(), &v));
        return DVectorT(*ctx_, v, n_);
    }
    bool is_sharded() const { return lbk_vec_is_sharded(v_.get()) != 0; }
    std::vector<T> to_host() const                                   // explicit device -> host
    {
        std::vector<T> h((size_t)n_);
        ctx_->check(lbk_vec_download(v_.get(), h.data(), n_));
        return h;
    }
    int64_t size() const { return n_; }
    lbk_vec *get() const { return v_.get(); }
    const Context &ctx() const { return *ctx_; }

  private:
    DVectorT(const Context &ctx, lbk_vec *v, int64_t n) : ctx_(&ctx), n_(n) { v_.reset(v, [](lbk_vec *p) { lbk_vec_free(p); }); }
    static lbk_vec *adopt(const Context &ctx, int64_t n, bool shard)
    {
        lbk_vec *v = nullptr;
        ctx.check(shard ? (sizeof(T) == 8 ? lbk_vec_alloc_sharded_f64(ctx.get(), n, &v) : lbk_vec_alloc_sharded(ctx.get(), n, &v)) : Abi<T>::alloc(ctx.get(), n, &v));
        return v;
    }
    const Context *ctx_;
    int64_t n_;
    std::shared_ptr<lbk_vec> v_;
};

// Page-locks a host buffer the caller owns for the lifetime of the object (lbk_host_register): uploads from it
// then run at the rate of a pinned copy.
class HostRegistration {
  public:
    HostRegistration(void *p, size_t nbytes) : p_(p) { if (int s = lbk_host_register(p, (int64_t)nbytes)) throw Error(s, lbk_last_error(nullptr)); }
    ~HostRegistration() { lbk_host_unregister(p_); }
    HostRegistration(const HostRegistration &) = delete;
    HostRegistration &operator=(const HostRegistration &) = delete;

  private:
    void *p_;
};
using DVector = DVectorT<float>;         // Vector Float
using DVectorD = DVectorT<double>;       // Vector Double

template <class T> struct GivensRotT { T r, z, c, s; };              // Types.hs:8
template <class T> struct ModGivensRotT {                            // Types.hs:15-20; flag names the constructor
    int flag = -2;                                                   // -2 FLAGNEG2, -1 FLAGNEG1, 0 FLAG0, 1 FLAG1
    T d1 = 0, d2 = 0, x1 = 0, h11 = 0, h12 = 0, h21 = 0, h22 = 0;
    void sparam(T p[5]) const                                        // test/Foreign.hs:276-280
    {
        const T t[4][5] = {{-2, 1, 0, 0, 1}, {-1, h11, h21, h12, h22}, {0, 0, h21, h12, 0}, {1, h11, 0, 0, h22}};
        for (int i = 0; i < 5; ++i) p[i] = t[flag + 2][i];
    }
};
using GivensRot = GivensRotT<float>;
using ModGivensRot = ModGivensRotT<float>;

// ---- products and norms (polymorphic like the reference's dot/asum/nrm2/iamax) -------------------------
template <class T> T dot(int64_t n, const DVectorT<T> &x, int64_t incx, const DVectorT<T> &y, int64_t incy)
{ T o; x.ctx().check(Abi<T>::dot(x.ctx().get(), n, x.get(), incx, y.get(), incy, &o)); return o; }
inline float sdsdot(int64_t n, float sb, const DVector &x, int64_t incx, const DVector &y, int64_t incy)
{ float o; x.ctx().check(lbk_sdsdot(x.ctx().get(), n, sb, x.get(), incx, y.get(), incy, &o)); return o; }
template <class T> T asum(int64_t n, const DVectorT<T> &x, int64_t incx)
{ T o; x.ctx().check(Abi<T>::asum(x.ctx().get(), n, x.get(), incx, &o)); return o; }
template <class T> T nrm2(int64_t n, const DVectorT<T> &x, int64_t incx)
{ T o; x.ctx().check(Abi<T>::nrm2(x.ctx().get(), n, x.get(), incx, &o)); return o; }
template <class T> int64_t iamax(int64_t n, const DVectorT<T> &x, int64_t incx)
{ int64_t o; x.ctx().check(Abi<T>::iamax(x.ctx().get(), n, x.get(), incx, &o)); return o; }

// ---- pure vector results -----------------------------------------------------------------------------
template <class T> DVectorT<T> scal(int64_t n, T a, const DVectorT<T> &x, int64_t incx)
{ DVectorT<T> o = x.like(); x.ctx().check(Abi<T>::scal_oop(x.ctx().get(), n, a, x.get(), incx, o.get())); return o; }
template <class T> DVectorT<T> copy(int64_t n, const DVectorT<T> &x, int64_t incx, const DVectorT<T> &y, int64_t incy)
{ DVectorT<T> o = y.like(); x.ctx().check(Abi<T>::copy_oop(x.ctx().get(), n, x.get(), incx, y.get(), incy, o.get())); return o; }
template <class T> DVectorT<T> axpy(int64_t n, T a, const DVectorT<T> &x, int64_t incx, const DVectorT<T> &y, int64_t incy)
{ DVectorT<T> o = y.like(); x.ctx().check(Abi<T>::axpy_oop(x.ctx().get(), n, a, x.get(), incx, y.get(), incy, o.get())); return o; }
template <class T> std::pair<DVectorT<T>, DVectorT<T>> swap(int64_t n, const DVectorT<T> &x, int64_t incx, const DVectorT<T> &y, int64_t incy)
{
    DVectorT<T> xo = x.like(), yo = y.like();
    x.ctx().check(Abi<T>::swap_oop(x.ctx().get(), n, x.get(), incx, y.get(), incy, xo.get(), yo.get()));
    return {xo, yo};
}
template <class T> std::pair<DVectorT<T>, DVectorT<T>> rot(int64_t n, const DVectorT<T> &x, int64_t incx, const DVectorT<T> &y, int64_t incy, T c, T s)
{
    DVectorT<T> xo = x.like(), yo = y.like();
    x.ctx().check(Abi<T>::rot_oop(x.ctx().get(), n, x.get(), incx, y.get(), incy, c, s, xo.get(), yo.get()));
    return {xo, yo};
}
template <class T> std::pair<DVectorT<T>, DVectorT<T>> rotm(const ModGivensRotT<T> &flags, int64_t n, const DVectorT<T> &x, int64_t incx, const DVectorT<T> &y, int64_t incy)
{
    if (flags.flag == -2) return {x, y};                              // Single.hs:473: the arguments themselves
    T p[5];
    flags.sparam(p);
    DVectorT<T> xo = x.like(), yo = y.like();
    x.ctx().check(Abi<T>::rotm_oop(x.ctx().get(), n, x.get(), incx, y.get(), incy, p, xo.get(), yo.get()));
    return {xo, yo};
}

// ---- host scalars ---------------------------------------------------------------------------------------
template <class T> GivensRotT<T> rotg(T a, T b)
{ T o[4]; Abi<T>::rotg(a, b, o); return {o[0], o[1], o[2], o[3]}; }
template <class T> ModGivensRotT<T> rotmg(T sd1, T sd2, T sx1, T sy1)
{
    T o[8];
    Abi<T>::rotmg(sd1, sd2, sx1, sy1, o, nullptr);
    ModGivensRotT<T> m;
    m.flag = (int)o[0];
    if (m.flag != -2) { m.d1 = o[1]; m.d2 = o[2]; m.x1 = o[3]; m.h11 = o[4]; m.h12 = o[5]; m.h21 = o[6]; m.h22 = o[7]; }
    return m;
}

// ---- the reference's deprecated monomorphic aliases (Single.hs:83-87, ...) ---------------------------------
#define LB_ALIASES(P, IP, T)                                                                                                   \
    inline T P##dot(int64_t n, const DVectorT<T> &x, int64_t ix, const DVectorT<T> &y, int64_t iy) { return dot<T>(n, x, ix, y, iy); }  \
    inline T P##asum(int64_t n, const DVectorT<T> &x, int64_t ix) { return asum<T>(n, x, ix); }                                  \
    inline T P##nrm2(int64_t n, const DVectorT<T> &x, int64_t ix) { return nrm2<T>(n, x, ix); }                                  \
    inline int64_t IP##amax(int64_t n, const DVectorT<T> &x, int64_t ix) { return iamax<T>(n, x, ix); }                          \
    inline DVectorT<T> P##scal(int64_t n, T a, const DVectorT<T> &x, int64_t ix) { return scal<T>(n, a, x, ix); }                \
    inline DVectorT<T> P##copy(int64_t n, const DVectorT<T> &x, int64_t ix, const DVectorT<T> &y, int64_t iy) { return copy<T>(n, x, ix, y, iy); } \
    inline DVectorT<T> P##axpy(int64_t n, T a, const DVectorT<T> &x, int64_t ix, const DVectorT<T> &y, int64_t iy) { return axpy<T>(n, a, x, ix, y, iy); } \
    inline std::pair<DVectorT<T>, DVectorT<T>> P##swap(int64_t n, const DVectorT<T> &x, int64_t ix, const DVectorT<T> &y, int64_t iy) { return swap<T>(n, x, ix, y, iy); } \
    inline std::pair<DVectorT<T>, DVectorT<T>> P##rot(int64_t n, const DVectorT<T> &x, int64_t ix, const DVectorT<T> &y, int64_t iy, T c, T s) { return rot<T>(n, x, ix, y, iy, c, s); } \
    inline std::pair<DVectorT<T>, DVectorT<T>> P##rotm(const ModGivensRotT<T> &f, int64_t n, const DVectorT<T> &x, int64_t ix, const DVectorT<T> &y, int64_t iy) { return rotm<T>(f, n, x, ix, y, iy); } \
    inline GivensRotT<T> P##rotg(T a, T b) { return rotg<T>(a, b); }                                                             \
    inline ModGivensRotT<T> P##rotmg(T d1, T d2, T x1, T y1) { return rotmg<T>(d1, d2, x1, y1); }
LB_ALIASES(s, is, float)
LB_ALIASES(d, id, double)
#undef LB_ALIASES

}  // namespace lambda_blas
