// lbk_kernels.cuh -- sm_100a device code for the Level-1 path.
//
// Every routine of Numerical.BLAS.Single that touches a vector (reference:
// src/Numerical/BLAS/Single.hs, index semantics src/Numerical/BLAS/Util.hs:39-142) is one of
// two shapes, both bounded by HBM bandwidth (O(1) flop per 4..16 bytes, no data reuse, hence no
// shared-memory staging and no tensor cores):
//
//   map     x[kx(i)], y[ky(i)]  ->  x[kx(i)], y[ky(i)]          saxpy sscal scopy sswap srot srotm
//   reduce  x[kx(i)] (, y[ky(i)]) -> one record                 sdot sdsdot sasum snrm2 isamax
//
// and each shape has two kernels:
//   *_unit     incx == incy == 1 : 128- or 256-bit coalesced global loads/stores (LDG.E.128 /
//              LDG.E.256 on sm_100a), UNROLL independent vectors in flight per thread per operand,
//              a persistent grid-stride loop over tiles with the grid sized to the resident wave;
//   *_strided  any increments (negative, zero, non-unit): one logical element per thread, 64-bit
//              index arithmetic, kx(i) = first(n,inc) + i*inc (Single.hs:91-96).
//
// Reductions are deterministic for a given launch shape: per-thread accumulation in a fixed
// order, warp shuffle tree, fixed-order block combine, one record per block, and the last block
// to arrive (ticket) folds the block records in index order ("two-pass in one launch").
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace lbk {

typedef long long i64;

constexpr int kMaxThreads = 512;        // launch bound of every streaming kernel
constexpr unsigned kNanKey = 0xFFFFFFFFu;

// ------------------------------------------------------------------------------------------------
// 128/256-bit global memory access
// ------------------------------------------------------------------------------------------------
template <int VEC> struct alignas(VEC * 4) Pack { float v[VEC]; };

// POLICY 0: default caching.  POLICY 1: streaming -- do not allocate in L1 (each byte is touched
// once), read-only operands through the non-coherent path.
template <int VEC, int POLICY, bool RO>
__device__ __forceinline__ void load_pack(Pack<VEC> &p, const float *ptr)
{
    if constexpr (VEC == 1) {
        if constexpr (POLICY >= 1 && RO)
            asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(p.v[0]) : "l"(ptr));
        else if constexpr (POLICY >= 1)
            asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(p.v[0]) : "l"(ptr) : "memory");
        else
            asm volatile("ld.global.f32 %0, [%1];" : "=f"(p.v[0]) : "l"(ptr) : "memory");
    } else if constexpr (VEC == 4) {
        if constexpr (POLICY >= 1 && RO)
            asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                         : "=f"(p.v[0]), "=f"(p.v[1]), "=f"(p.v[2]), "=f"(p.v[3]) : "l"(ptr));
        else if constexpr (POLICY >= 1)
            asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                         : "=f"(p.v[0]), "=f"(p.v[1]), "=f"(p.v[2]), "=f"(p.v[3]) : "l"(ptr) : "memory");
        else
            asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];"
                         : "=f"(p.v[0]), "=f"(p.v[1]), "=f"(p.v[2]), "=f"(p.v[3]) : "l"(ptr) : "memory");
    } else {
        static_assert(VEC == 8, "VEC is 1, 4 or 8");
#define LBK_LD8(QUAL, CLOBBER)                                                                      \
        asm volatile("ld.global" QUAL ".v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"                     \
                     : "=f"(p.v[0]), "=f"(p.v[1]), "=f"(p.v[2]), "=f"(p.v[3]),                         \
                       "=f"(p.v[4]), "=f"(p.v[5]), "=f"(p.v[6]), "=f"(p.v[7]) : "l"(ptr) CLOBBER)
        // experimental L2 policies (256-bit accesses only): 2/3 evict-first loads, 4 256-byte L2 prefetch
        if constexpr ((POLICY == 2 || POLICY == 3) && RO) LBK_LD8(".nc.L1::no_allocate.L2::evict_first", );
        else if constexpr (POLICY == 2 || POLICY == 3) LBK_LD8(".L1::no_allocate.L2::evict_first", : "memory");
        else if constexpr (POLICY == 4 && RO) LBK_LD8(".nc.L1::no_allocate.L2::256B", );
        else if constexpr (POLICY == 4) LBK_LD8(".L1::no_allocate.L2::256B", : "memory");
        else if constexpr (POLICY >= 1 && RO) LBK_LD8(".nc.L1::no_allocate", );
        else if constexpr (POLICY >= 1) LBK_LD8(".L1::no_allocate", : "memory");
        else LBK_LD8("", : "memory");
#undef LBK_LD8
    }
}

template <int VEC, int POLICY>
__device__ __forceinline__ void store_pack(float *ptr, const Pack<VEC> &p)
{
    if constexpr (VEC == 1) {
        asm volatile("st.global.f32 [%0], %1;" ::"l"(ptr), "f"(p.v[0]) : "memory");
    } else if constexpr (VEC == 4) {
        if constexpr (POLICY >= 1)
            asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(ptr),
                         "f"(p.v[0]), "f"(p.v[1]), "f"(p.v[2]), "f"(p.v[3]) : "memory");
        else
            asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(ptr),
                         "f"(p.v[0]), "f"(p.v[1]), "f"(p.v[2]), "f"(p.v[3]) : "memory");
    } else {
#define LBK_ST8(QUAL)                                                                               \
        asm volatile("st.global" QUAL ".v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(ptr),         \
                     "f"(p.v[0]), "f"(p.v[1]), "f"(p.v[2]), "f"(p.v[3]),                               \
                     "f"(p.v[4]), "f"(p.v[5]), "f"(p.v[6]), "f"(p.v[7]) : "memory")
        if constexpr (POLICY == 3 || POLICY == 5) LBK_ST8(".L1::no_allocate.L2::evict_first");
        else if constexpr (POLICY == 4) LBK_ST8(".cs");
        else if constexpr (POLICY >= 1) LBK_ST8(".L1::no_allocate");
        else LBK_ST8("");
#undef LBK_ST8
    }
}

// ------------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  A reduction ends with a latency-bound tail (block combine,
// fence, ticket, last-block fold: ~8 us, as much as 5% of a 4-byte-per-element kernel at n = 2^28).
// Once a block of a reduction has issued its last load it lets the NEXT kernel of the stream start
// (launch_dependents); that kernel streams through HBM while the tail drains and calls wait()
// before it touches anything the tail still owns (scratch, ticket) or before it exits.  The host only
// sets the launch attribute when the next kernel's operands do not overlap the tail's output.
// Both instructions are no-ops in a kernel launched without the attribute.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// map functors.  (x, y) are the OLD values; the functor overwrites them with the new ones.  Every
// product is rounded before it is added (__fmul_rn / __fadd_rn are never contracted into an FMA),
// which is what the reference's compiled code does, so results are bit-identical to it.
// ------------------------------------------------------------------------------------------------
struct AxpyF {   // Single.hs:439   y + a*x
    float a;
    static constexpr bool RX = true, RY = true, WX = false, WY = true;
    __device__ __forceinline__ void operator()(float &x, float &y) const { y = __fadd_rn(y, __fmul_rn(a, x)); }
};
struct ScalF {   // Single.hs:110   x*a
    float a;
    static constexpr bool RX = true, RY = false, WX = true, WY = false;
    __device__ __forceinline__ void operator()(float &x, float &) const { x = __fmul_rn(x, a); }
};
struct CopyF {   // Single.hs:127   \_ x -> x
    static constexpr bool RX = true, RY = false, WX = false, WY = true;
    __device__ __forceinline__ void operator()(float &x, float &y) const { y = x; }
};
struct SwapF {   // Single.hs:143-145
    static constexpr bool RX = true, RY = true, WX = true, WY = true;
    __device__ __forceinline__ void operator()(float &x, float &y) const { float t = x; x = y; y = t; }
};
struct RotF {    // Single.hs:418-419   c*x + s*y ; c*y - s*x
    float c, s;
    static constexpr bool RX = true, RY = true, WX = true, WY = true;
    __device__ __forceinline__ void operator()(float &x, float &y) const
    {
        float nx = __fadd_rn(__fmul_rn(c, x), __fmul_rn(s, y));
        float ny = __fsub_rn(__fmul_rn(c, y), __fmul_rn(s, x));
        x = nx; y = ny;
    }
};
struct RotmNeg1F {   // Single.hs:475-476   h11*x + h12*y ; h21*x + h22*y
    float h11, h12, h21, h22;
    static constexpr bool RX = true, RY = true, WX = true, WY = true;
    __device__ __forceinline__ void operator()(float &x, float &y) const
    {
        float nx = __fadd_rn(__fmul_rn(h11, x), __fmul_rn(h12, y));
        float ny = __fadd_rn(__fmul_rn(h21, x), __fmul_rn(h22, y));
        x = nx; y = ny;
    }
};
struct Rotm0F {      // Single.hs:478-479   x + h12*y ; h21*x + y
    float h12, h21;
    static constexpr bool RX = true, RY = true, WX = true, WY = true;
    __device__ __forceinline__ void operator()(float &x, float &y) const
    {
        float nx = __fadd_rn(x, __fmul_rn(h12, y));
        float ny = __fadd_rn(__fmul_rn(h21, x), y);
        x = nx; y = ny;
    }
};
struct Rotm1F {      // Single.hs:481-482   h11*x + y ; -x + h22*y
    float h11, h22;
    static constexpr bool RX = true, RY = true, WX = true, WY = true;
    __device__ __forceinline__ void operator()(float &x, float &y) const
    {
        float nx = __fadd_rn(__fmul_rn(h11, x), y);
        float ny = __fadd_rn(-x, __fmul_rn(h22, y));
        x = nx; y = ny;
    }
};

// ------------------------------------------------------------------------------------------------
// map, unit stride.  Reads come from (xr, yr), writes go to (xw, yw); in place when they are equal,
// out of place (the reference's pure result) when they differ.  `head` scalar elements precede the
// vector body so that xr+head (and yr/xw/yw+head: checked by the host) are VEC*4-byte aligned; they
// and the < VEC tail are handled by the first threads of block 0.
// ------------------------------------------------------------------------------------------------
//
// REV: the (incx, incy) = (1,-1) / (-1,1) pairing -- logical element i is x[i] and y[n-1-i] -- still moves
// whole aligned packs: the y pack for x[e..e+VEC) is y[n-e-VEC..n-e) with its lanes in reverse order.
template <class F, int VEC, int UNROLL, int POLICY, bool REV = false>
__global__ void __launch_bounds__(kMaxThreads)
map_unit_kernel(F f, const float *xr, const float *yr, float *xw, float *yw, i64 head, i64 n)
{
    const i64 body = n - head;
    const i64 npack = body / VEC;
    const i64 per_tile = (i64)UNROLL * blockDim.x;
    const i64 ntiles = (npack + per_tile - 1) / per_tile;
    const float *bxr = xr + head;
    float *bxw = xw + head;
    // y side: forward from y+head, or backward from one past the body's last y element
    const float *byr = REV ? yr + (n - head) - VEC : yr + head;
    float *byw = REV ? yw + (n - head) - VEC : yw + head;
    constexpr int YS = REV ? -1 : 1;

    for (i64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const i64 p0 = tile * per_tile + threadIdx.x;
        Pack<VEC> xv[UNROLL], yv[UNROLL];
        if (p0 + (i64)(UNROLL - 1) * blockDim.x < npack) {   // whole tile in range for this thread
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const i64 e = (p0 + (i64)u * blockDim.x) * VEC;
                if constexpr (F::RX) load_pack<VEC, POLICY, !F::WX>(xv[u], bxr + e);
                if constexpr (F::RY) load_pack<VEC, POLICY, !F::WY>(yv[u], byr + YS * e);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
                for (int l = 0; l < VEC; ++l) f(xv[u].v[l], yv[u].v[REV ? VEC - 1 - l : l]);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const i64 e = (p0 + (i64)u * blockDim.x) * VEC;
                if constexpr (F::WX) store_pack<VEC, POLICY>(bxw + e, xv[u]);
                if constexpr (F::WY) store_pack<VEC, POLICY>(byw + YS * e, yv[u]);
            }
        } else {
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const i64 p = p0 + (i64)u * blockDim.x;
                if (p < npack) {
                    const i64 e = p * VEC;
                    if constexpr (F::RX) load_pack<VEC, POLICY, !F::WX>(xv[u], bxr + e);
                    if constexpr (F::RY) load_pack<VEC, POLICY, !F::WY>(yv[u], byr + YS * e);
#pragma unroll
                    for (int l = 0; l < VEC; ++l) f(xv[u].v[l], yv[u].v[REV ? VEC - 1 - l : l]);
                    if constexpr (F::WX) store_pack<VEC, POLICY>(bxw + e, xv[u]);
                    if constexpr (F::WY) store_pack<VEC, POLICY>(byw + YS * e, yv[u]);
                }
            }
        }
    }
    // scalar head [0, head) and tail [head + npack*VEC, n)
    if (blockIdx.x == 0) {
        const i64 tail0 = head + npack * VEC;
        const i64 nscalar = head + (n - tail0);
        for (i64 t = threadIdx.x; t < nscalar; t += blockDim.x) {
            const i64 e = t < head ? t : tail0 + (t - head);
            const i64 ey = REV ? n - 1 - e : e;
            float xs = 0.f, ys = 0.f;
            if constexpr (F::RX) xs = xr[e];
            if constexpr (F::RY) ys = yr[ey];
            f(xs, ys);
            if constexpr (F::WX) xw[e] = xs;
            if constexpr (F::WY) yw[ey] = ys;
        }
    }
    pdl_wait();     // as a PDL secondary: do not retire before the preceding reduction has
}

// ------------------------------------------------------------------------------------------------
// map, any increments.  Logical element i lives at x[kx0 + i*incx], y[ky0 + i*incy] with
// kx0 = first(n,incx) (Single.hs:91-96, Util.hs:81-88,141-142).  Right-hand sides always read the
// ORIGINAL values (Util.hs:134-142): when an OUTPUT increment is 0 the host passes a snapshot of
// element 0 as the read pointer and sets *_last_only, so that only logical element n-1 (the last
// write of the reference's sequential scatter) stores.
// ------------------------------------------------------------------------------------------------
constexpr int kStridedUnroll = 4;   // independent 4-byte accesses in flight per operand per thread

template <class F>
__global__ void __launch_bounds__(256)
map_strided_kernel(F f, const float *xr, float *xw, i64 incx_r, i64 incx_w, i64 kx0_r, i64 kx0_w,
                   const float *yr, float *yw, i64 incy_r, i64 incy_w, i64 ky0_r, i64 ky0_w,
                   i64 i_begin, i64 n, int x_last_only, int y_last_only)
{
    constexpr int U = kStridedUnroll;
    const i64 per_tile = (i64)U * blockDim.x;
    const i64 count = n - i_begin;
    const i64 ntiles = (count + per_tile - 1) / per_tile;
    for (i64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const i64 i0 = i_begin + tile * per_tile + threadIdx.x;
        float xs[U], ys[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {            // all loads first: U independent requests per operand
            const i64 i = i0 + (i64)u * blockDim.x;
            xs[u] = 0.f; ys[u] = 0.f;
            if (i < n) {      // asm volatile loads: issued here, all U before the first use
                Pack<1> p;
                if constexpr (F::RX) { load_pack<1, 0, !F::WX>(p, xr + kx0_r + i * incx_r); xs[u] = p.v[0]; }
                if constexpr (F::RY) { load_pack<1, 0, !F::WY>(p, yr + ky0_r + i * incy_r); ys[u] = p.v[0]; }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) asm volatile("" : "+f"(xs[u]), "+f"(ys[u]));   // keep the U loads ahead of the first use
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const i64 i = i0 + (i64)u * blockDim.x;
            if (i < n) {
                f(xs[u], ys[u]);
                if constexpr (F::WX) { if (!x_last_only || i == n - 1) xw[kx0_w + i * incx_w] = xs[u]; }
                if constexpr (F::WY) { if (!y_last_only || i == n - 1) yw[ky0_w + i * incy_w] = ys[u]; }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// reductions
// ------------------------------------------------------------------------------------------------
// The record a block (and, across GPUs, a rank) contributes.  Layout shared with the host
// (lbk_record in include/lbk.h): f[0..2] sums, key/idx the isamax candidate.
struct Record {
    float f[3];
    unsigned key;
    i64 idx;
    i64 pad;
};
static_assert(sizeof(Record) == 32, "Record is 32 bytes");

__device__ __forceinline__ Record shfl_xor_record(const Record &r, int mask)
{
    Record o;
    o.f[0] = __shfl_xor_sync(0xffffffffu, r.f[0], mask);
    o.f[1] = __shfl_xor_sync(0xffffffffu, r.f[1], mask);
    o.f[2] = __shfl_xor_sync(0xffffffffu, r.f[2], mask);
    o.key = __shfl_xor_sync(0xffffffffu, r.key, mask);
    o.idx = __shfl_xor_sync(0xffffffffu, r.idx, mask);
    o.pad = 0;
    return o;
}

// Blue's scaled accumulation (as in LAPACK 3.10 snrm2, thresholds tightened so that 2^36 mid-range
// squares cannot overflow binary32): |x| > kBig is accumulated as (x*kSBig)^2, |x| < kSml as
// (x*kSSml)^2, everything else as x^2.  All three are plain sums, so the cross-block and
// cross-GPU combine is addition.
constexpr float kBig = 70368744177664.0f;            // 2^46
constexpr float kSml = 1.0842021724855044e-19f;      // 2^-63
constexpr float kSBig = 1.3234889800848443e-23f;     // 2^-76
constexpr float kSSml = 3.7778931862957162e+22f;     // 2^75

// Op concept:  Acc, zero(), step(acc,x,y,ord), to_record(acc, ord->index map), merge(a,b)
struct DotOp {       // Single.hs:81-82
    static constexpr bool TWO = true, IS_IAMAX = false, HAS_PACK_STEP = false;
    struct Acc { float s; };
    __device__ static __forceinline__ Acc zero() { return {0.f}; }
    __device__ static __forceinline__ void step(Acc &a, float x, float y, unsigned) { a.s = fmaf(x, y, a.s); }
    __device__ static __forceinline__ void to_record(Record &r, const Acc &a) { r.f[0] = a.s; }
    __device__ static __forceinline__ void single(Record &r, float x, float y, i64) { r.f[0] = fmaf(x, y, r.f[0]); }
    __host__ __device__ static __forceinline__ Record merge(const Record &a, const Record &b)
    { Record o = a; o.f[0] = a.f[0] + b.f[0]; return o; }
};
struct DsdotOp {     // Single.hs:238-262: products and sum in double (held as a double split over f[0],f[1] bits)
    static constexpr bool TWO = true, IS_IAMAX = false, HAS_PACK_STEP = false;
    struct Acc { double s; };
    __device__ static __forceinline__ Acc zero() { return {0.0}; }
    __device__ static __forceinline__ void step(Acc &a, float x, float y, unsigned) { a.s = fma((double)x, (double)y, a.s); }
    __device__ static __forceinline__ double get(const Record &r)
    { return __hiloint2double(__float_as_int(r.f[1]), __float_as_int(r.f[0])); }
    __device__ static __forceinline__ void put(Record &r, double d)
    { r.f[0] = __int_as_float(__double2loint(d)); r.f[1] = __int_as_float(__double2hiint(d)); }
    __device__ static __forceinline__ void to_record(Record &r, const Acc &a) { put(r, a.s); }
    __device__ static __forceinline__ void single(Record &r, float x, float y, i64) { put(r, fma((double)x, (double)y, get(r))); }
    __device__ static __forceinline__ Record merge(const Record &a, const Record &b)
    { Record o = a; put(o, get(a) + get(b)); return o; }
};
struct AsumOp {      // Single.hs:163
    static constexpr bool TWO = false, IS_IAMAX = false, HAS_PACK_STEP = false;
    struct Acc { float s; };
    __device__ static __forceinline__ Acc zero() { return {0.f}; }
    __device__ static __forceinline__ void step(Acc &a, float x, float, unsigned) { a.s += fabsf(x); }
    __device__ static __forceinline__ void to_record(Record &r, const Acc &a) { r.f[0] = a.s; }
    __device__ static __forceinline__ void single(Record &r, float x, float, i64) { r.f[0] += fabsf(x); }
    __host__ __device__ static __forceinline__ Record merge(const Record &a, const Record &b)
    { Record o = a; o.f[0] = a.f[0] + b.f[0]; return o; }
};
struct Nrm2Op {      // Single.hs:180-199, restated as Blue's three accumulators
    static constexpr bool TWO = false, IS_IAMAX = false, HAS_PACK_STEP = true;
    struct Acc { float big, med, sml; };
    __device__ static __forceinline__ Acc zero() { return {0.f, 0.f, 0.f}; }
    __device__ static __forceinline__ void step(Acc &a, float x, float, unsigned)
    {
        const float ax = fabsf(x);
        if (ax > kBig) { const float t = ax * kSBig; a.big = fmaf(t, t, a.big); }
        else if (ax < kSml) { const float t = ax * kSSml; a.sml = fmaf(t, t, a.sml); }
        else a.med = fmaf(ax, ax, a.med);      // NaN lands here and propagates
    }
    // One vector of lanes: when every lane is mid-range (the overwhelmingly common case) the pack costs
    // one FFMA and two FMNMX per element; a pack holding a huge, tiny or zero lane takes the exact
    // per-lane classification.  fmaxf/fminf skip NaN lanes, and the fast path's x*x propagates them.
    template <int VEC>
    __device__ static __forceinline__ void step_pack(Acc &a, const Pack<VEC> &x, const Pack<VEC> &, unsigned)
    {
        float mx = fabsf(x.v[0]), mn = mx;
#pragma unroll
        for (int l = 1; l < VEC; ++l) { mx = fmaxf(mx, fabsf(x.v[l])); mn = fminf(mn, fabsf(x.v[l])); }
        if (mx <= kBig && mn >= kSml) {
#pragma unroll
            for (int l = 0; l < VEC; ++l) a.med = fmaf(x.v[l], x.v[l], a.med);
        } else {
#pragma unroll
            for (int l = 0; l < VEC; ++l) step(a, x.v[l], 0.f, 0u);
        }
    }
    __device__ static __forceinline__ void to_record(Record &r, const Acc &a) { r.f[0] = a.big; r.f[1] = a.med; r.f[2] = a.sml; }
    __device__ static __forceinline__ void single(Record &r, float x, float, i64)
    { Acc a = {r.f[0], r.f[1], r.f[2]}; step(a, x, 0.f, 0u); to_record(r, a); }
    __host__ __device__ static __forceinline__ Record merge(const Record &a, const Record &b)
    { Record o = a; o.f[0] = a.f[0] + b.f[0]; o.f[1] = a.f[1] + b.f[1]; o.f[2] = a.f[2] + b.f[2]; return o; }
};
// isamax (Single.hs:217-220).  key = bit pattern of |x| (monotone for non-NaN), NaN -> 0 so that it
// can never displace an earlier element (`|best| < |NaN|` is false in the reference's fold); the
// "NaN at logical index 0 is never displaced" half of the rule is applied once, at the finish.
__device__ __forceinline__ unsigned amax_key(float x)
{
    const unsigned b = __float_as_uint(x) & 0x7fffffffu;
    return b > 0x7f800000u ? 0u : b;
}
struct IamaxOp {
    static constexpr bool TWO = false, IS_IAMAX = true, HAS_PACK_STEP = true;
    // running maximum of |x| as a float (-1 = nothing seen yet; any |x| >= 0 beats it) and the ordinal,
    // within this thread's element sequence, of the FIRST element that reached it
    struct Acc { float val; unsigned ord; };
    __device__ static __forceinline__ Acc zero() { return {-1.f, 0u}; }
    __device__ static __forceinline__ void step(Acc &a, float x, float, unsigned ord)
    {
        const float ax = fabsf(x);
        if (ax > a.val) { a.val = ax; a.ord = ord; }      // false for NaN: a NaN never displaces
    }
    // elements reach a thread in increasing index order, so strict '>' keeps the first maximum.  Per
    // pack: one FMNMX per element (fmaxf drops NaN lanes, which is the reference's "NaN never wins");
    // only a pack that raises the thread's maximum (O(log) times per thread) looks for the lane.
    template <int VEC>
    __device__ static __forceinline__ void step_pack(Acc &a, const Pack<VEC> &x, const Pack<VEC> &, unsigned ord0)
    {
        float m = fabsf(x.v[0]);
#pragma unroll
        for (int l = 1; l < VEC; ++l) m = fmaxf(m, fabsf(x.v[l]));
        if (m > a.val) {
            unsigned lane = 0;
#pragma unroll
            for (int l = VEC - 1; l >= 0; --l) if (fabsf(x.v[l]) == m) lane = l;
            a.val = m; a.ord = ord0 + lane;
        }
    }
    __device__ static __forceinline__ void to_record(Record &, const Acc &) {}
    __device__ static __forceinline__ void single(Record &r, float x, float, i64 idx)
    {
        const unsigned k = amax_key(x);
        if (r.idx < 0 || k > r.key || (k == r.key && idx < r.idx)) { r.key = k; r.idx = idx; }
    }
    // larger key wins; on a tie the lower index; idx < 0 means "no candidate"
    __host__ __device__ static __forceinline__ Record merge(const Record &a, const Record &b)
    {
        if (b.idx < 0) return a;
        if (a.idx < 0) return b;
        if (b.key > a.key || (b.key == a.key && b.idx < a.idx)) return b;
        return a;
    }
};

template <class Op>
__device__ __forceinline__ Record empty_record()
{
    Record r;
    r.f[0] = r.f[1] = r.f[2] = 0.f;
    r.key = 0u;
    r.idx = -1;
    r.pad = 0;
    return r;
}

// Block-wide fixed-order combine; the result is valid in thread 0.
template <class Op>
__device__ __forceinline__ Record block_combine(Record r, Record *smem)
{
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) r = Op::merge(r, shfl_xor_record(r, m));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = (blockDim.x + 31) >> 5;
    __syncthreads();                       // smem may still be in use by a previous combine
    if (lane == 0) smem[warp] = r;
    __syncthreads();
    if (warp == 0) {
        r = lane < nwarps ? smem[lane] : empty_record<Op>();
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) r = Op::merge(r, shfl_xor_record(r, m));
    }
    return r;
}

// What the last block does with the folded record.
//   mode 0: finish on this GPU and store the scalar to `out` (float, or int64 for isamax);
//   mode 1: store the record to `rec_out`; an NCCL all-gather and finish_records_kernel end the call;
//   mode 2: fused exchange -- push the record straight into every peer GPU's mailbox over NVLink (peer
//           memory mapped with CUDA IPC), wait for the peers' records in this GPU's mailbox, fold them in
//           rank order and finish, all inside this kernel.
struct FinishArgs {
    int kind;             // lbk_reduce_kind: 0 sum, 1 nrm2, 2 iamax, 3 dsdot
    int mode;
    const float *x0;      // logical element 0 if this rank owns it (isamax NaN rule, nrm2 n==1), else null
    i64 n_global;
    double sb;            // sdsdot's addend
    void *out;
    Record *rec_out;
    // mode 2
    unsigned long long *const *peers;   // device array: peers[r] = rank r's mailbox (peers[rank] is local)
    int nranks, rank;
    unsigned seq;                       // number of this exchange, identical on every rank, never 0
    int *error_flag;                    // host-mapped: set to 1 if a peer's record never arrived
};

// Mailbox: two parities x kMaxRanks senders x 8 words.  Each 64-bit word carries 32 bits of record and
// the 32-bit exchange number, so one store publishes data and "ready" together (no fence, no second
// round trip); a word is valid once its tag equals the current exchange number.  Exchange k+2 can only
// reuse exchange k's slots after every rank has consumed them (a rank sends k+1 only after finishing k).
constexpr int kMaxRanks = 64;
constexpr int kMailboxWords = 2 * kMaxRanks * 8;
constexpr unsigned long long kExchangeTimeoutNs = 10ull * 1000ull * 1000ull * 1000ull;

__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ Record fold_by_kind(const Record *recs, int nranks, int kind);

__host__ __device__ __forceinline__ float nrm2_finish(float abig, float amed, float asml)
{
    float scl, sumsq;
    if (abig > 0.f) {
        if (amed > 0.f || amed != amed) abig += (amed * kSBig) * kSBig;
        scl = 1.f / kSBig; sumsq = abig;
    } else if (asml > 0.f) {
        if (amed > 0.f || amed != amed) {
            const float m = sqrtf(amed), s = sqrtf(asml) / kSSml;
            const float ymin = fminf(m, s), ymax = fmaxf(m, s);
            const float q = ymin / ymax;
            scl = 1.f; sumsq = (ymax * ymax) * (1.f + q * q);
            if (amed != amed) sumsq = amed;
        } else { scl = 1.f / kSSml; sumsq = asml; }
    } else { scl = 1.f; sumsq = amed; }
    return scl * sqrtf(sumsq);
}

// Applies the rules that depend on logical element 0 to this rank's record (before any exchange).
__device__ __forceinline__ void localize_record(Record &r, const FinishArgs &fa, float x0)
{
    if (fa.x0 != nullptr) {
        if (fa.kind == 2 && x0 != x0) { r.key = kNanKey; r.idx = 0; }      // NaN at index 0 stays
        if (fa.kind == 1) r.key = __float_as_uint(fabsf(x0));              // nrm2 n==1 -> |x[0]|
    }
}

__device__ __forceinline__ void finish_record(const Record &r, const FinishArgs &fa)
{
    if (fa.kind == 2) { *reinterpret_cast<i64 *>(fa.out) = r.idx; }
    else if (fa.kind == 1) {
        *reinterpret_cast<float *>(fa.out) =
            fa.n_global == 1 ? __uint_as_float(r.key) : nrm2_finish(r.f[0], r.f[1], r.f[2]);
    }
    else if (fa.kind == 3) { *reinterpret_cast<float *>(fa.out) = (float)(fa.sb + DsdotOp::get(r)); }
    else { *reinterpret_cast<float *>(fa.out) = r.f[0]; }
}

template <class Op>
__device__ __forceinline__ void grid_finish(Record r, Record *partials, unsigned *ticket,
                                            const FinishArgs &fa, Record *smem, float x0)
{
    __shared__ bool is_last;
    r = block_combine<Op>(r, smem);
    pdl_wait();          // the scratch below may still be in use by the preceding reduction's tail
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = r;
        __threadfence();
        const unsigned t = atomicInc(ticket, gridDim.x - 1);   // wraps to 0: ready for the next launch
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    Record acc = empty_record<Op>();
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
        // other blocks' records: read through L2 (L1 is not coherent across SMs)
        union { int4 q[2]; Record r; } u;
        u.q[0] = __ldcg(reinterpret_cast<const int4 *>(&partials[b]));
        u.q[1] = __ldcg(reinterpret_cast<const int4 *>(&partials[b]) + 1);
        acc = Op::merge(acc, u.r);
    }
    acc = block_combine<Op>(acc, smem);
    if (fa.mode != 2) {
        if (threadIdx.x == 0) {
            localize_record(acc, fa, x0);
            if (fa.mode == 0) finish_record(acc, fa);
            else *fa.rec_out = acc;
        }
        return;
    }
    // ---- fused exchange over peer memory ----
    __shared__ Record mine;
    __shared__ Record all[kMaxRanks];
    __shared__ int timed_out;
    if (threadIdx.x == 0) { localize_record(acc, fa, x0); mine = acc; timed_out = 0; }
    __syncthreads();
    const unsigned par = fa.seq & 1u;
    const unsigned *words = reinterpret_cast<const unsigned *>(&mine);
    for (int t = threadIdx.x; t < fa.nranks * 8; t += blockDim.x) {      // push: one 8-byte store per word per peer
        const int peer = t >> 3, w = t & 7;
        unsigned long long *dst = fa.peers[peer] + ((par * kMaxRanks + fa.rank) << 3) + w;
        const unsigned long long v = ((unsigned long long)fa.seq << 32) | words[w];
        asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(dst), "l"(v) : "memory");
    }
    const unsigned long long t_start = global_timer_ns();
    for (int t = threadIdx.x; t < fa.nranks * 8; t += blockDim.x) {      // pull: poll this GPU's mailbox
        const int src = t >> 3, w = t & 7;
        const unsigned long long *slot = fa.peers[fa.rank] + ((par * kMaxRanks + src) << 3) + w;
        unsigned long long v;
        for (;;) {
            asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(slot) : "memory");
            if ((unsigned)(v >> 32) == fa.seq) break;
            if (global_timer_ns() - t_start > kExchangeTimeoutNs) { timed_out = 1; v = 0; break; }
        }
        reinterpret_cast<unsigned *>(&all[src])[w] = (unsigned)v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (timed_out) *fa.error_flag = 1;
        Record total = fold_by_kind(all, fa.nranks, fa.kind);
        finish_record(total, fa);
    }
}

template <class Op, int VEC, bool REV = false>
__device__ __forceinline__ void pack_step(typename Op::Acc &acc, const Pack<VEC> &x, const Pack<VEC> &y, unsigned ord0)
{
    if constexpr (Op::HAS_PACK_STEP) Op::template step_pack<VEC>(acc, x, y, ord0);
    else {
#pragma unroll
        for (int l = 0; l < VEC; ++l) Op::step(acc, x.v[l], Op::TWO ? y.v[REV ? VEC - 1 - l : l] : 0.f, ord0 + l);
    }
}

// reduce, unit stride.  idx_base = global logical index of element 0 of this shard (isamax).
template <class Op, int VEC, int UNROLL, int POLICY, bool REV = false>
__global__ void __launch_bounds__(kMaxThreads)
reduce_unit_kernel(const float *x, const float *y, i64 head, i64 n, i64 idx_base,
                   Record *partials, unsigned *ticket, FinishArgs fa)
{
    constexpr int YS = REV ? -1 : 1;     // REV: element i pairs x[i] with y[n-1-i] (see map_unit_kernel)
    __shared__ Record smem[32];
    const i64 body = n - head;
    const i64 npack = body / VEC;
    const i64 per_tile = (i64)UNROLL * blockDim.x;
    const i64 ntiles = (npack + per_tile - 1) / per_tile;
    const float *bx = x + head, *by = REV ? y + (n - head) - VEC : y + head;
    const float x0 = fa.x0 != nullptr ? *fa.x0 : 0.f;   // read now: a PDL secondary may overwrite x later

    typename Op::Acc acc = Op::zero();
    unsigned it = 0;
    for (i64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const i64 p0 = tile * per_tile + threadIdx.x;
        Pack<VEC> xv[UNROLL], yv[UNROLL];
        if (p0 + (i64)(UNROLL - 1) * blockDim.x < npack) {
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const i64 e = (p0 + (i64)u * blockDim.x) * VEC;
                load_pack<VEC, POLICY, true>(xv[u], bx + e);
                if constexpr (Op::TWO) load_pack<VEC, POLICY, true>(yv[u], by + YS * e);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) pack_step<Op, VEC, REV>(acc, xv[u], yv[u], it * (UNROLL * VEC) + u * VEC);
        } else {
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const i64 p = p0 + (i64)u * blockDim.x;
                if (p < npack) {
                    const i64 e = p * VEC;
                    load_pack<VEC, POLICY, true>(xv[u], bx + e);
                    if constexpr (Op::TWO) load_pack<VEC, POLICY, true>(yv[u], by + YS * e);
                    pack_step<Op, VEC, REV>(acc, xv[u], yv[u], it * (UNROLL * VEC) + u * VEC);
                }
            }
        }
    }
    Record r = empty_record<Op>();
    if (blockIdx.x == 0) {      // scalar head and tail
        const i64 tail0 = head + npack * VEC;
        const i64 nscalar = head + (n - tail0);
        for (i64 t = threadIdx.x; t < nscalar; t += blockDim.x) {
            const i64 e = t < head ? t : tail0 + (t - head);
            Op::single(r, x[e], Op::TWO ? y[REV ? n - 1 - e : e] : 0.f, idx_base + e);
        }
    }
    pdl_launch_dependents();    // every load of this block has been issued and consumed
    if constexpr (Op::IS_IAMAX) {
        if (acc.val >= 0.f) {    // saw >= 1 non-NaN element; ordinal -> global index
            const unsigned per = UNROLL * VEC;
            const unsigned iter = acc.ord / per, rem = acc.ord % per, u = rem / VEC, l = rem % VEC;
            const i64 tile = (i64)blockIdx.x + (i64)iter * gridDim.x;
            Record v = empty_record<Op>();
            v.key = __float_as_uint(acc.val);
            v.idx = idx_base + head + ((tile * UNROLL + u) * blockDim.x + threadIdx.x) * VEC + l;
            r = Op::merge(v, r);
        }
    } else {
        Record v = empty_record<Op>();
        Op::to_record(v, acc);
        r = Op::merge(v, r);
    }
    grid_finish<Op>(r, partials, ticket, fa, smem, x0);
}

// reduce, any increments: logical element i at x[kx0 + i*incx] (, y[ky0 + i*incy]).
template <class Op>
__global__ void __launch_bounds__(256)
reduce_strided_kernel(const float *x, i64 incx, i64 kx0, const float *y, i64 incy, i64 ky0, i64 n,
                      i64 idx_base, Record *partials, unsigned *ticket, FinishArgs fa)
{
    __shared__ Record smem[32];
    const float x0 = fa.x0 != nullptr ? *fa.x0 : 0.f;
    Record r = empty_record<Op>();
    // a plain grid-stride loop: the compiler unrolls it four times with the loads hoisted, which measured
    // faster than explicit tiling (ptxas interleaves predicated tile loads with their uses)
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        Op::single(r, x[kx0 + i * incx], Op::TWO ? y[ky0 + i * incy] : 0.f, idx_base + i);
    pdl_launch_dependents();
    grid_finish<Op>(r, partials, ticket, fa, smem, x0);
}

// Cross-GPU finish: every rank holds all ranks' records (rank order) and folds them identically.
template <class Op>
__host__ __device__ __forceinline__ Record fold_ranks(const Record *recs, int nranks)
{
    Record acc = recs[0];
    for (int r = 1; r < nranks; ++r) acc = Op::merge(acc, recs[r]);
    return acc;
}
__device__ __forceinline__ Record fold_by_kind(const Record *recs, int nranks, int kind)
{
    Record acc;
    if (kind == 2) acc = fold_ranks<IamaxOp>(recs, nranks);
    else if (kind == 1) { acc = fold_ranks<Nrm2Op>(recs, nranks); acc.key = recs[0].key; }   // |x[0]| travels in rank 0's key
    else if (kind == 3) acc = fold_ranks<DsdotOp>(recs, nranks);
    else acc = fold_ranks<DotOp>(recs, nranks);
    return acc;
}
__global__ void finish_records_kernel(const Record *recs, int nranks, FinishArgs fa)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    finish_record(fold_by_kind(recs, nranks, fa.kind), fa);
}

// ------------------------------------------------------------------------------------------------
// small utilities
// ------------------------------------------------------------------------------------------------
__global__ void store_f32_kernel(float *out, float v) { if (threadIdx.x == 0 && blockIdx.x == 0) *out = v; }
__global__ void store_i64_kernel(i64 *out, i64 v) { if (threadIdx.x == 0 && blockIdx.x == 0) *out = v; }

__host__ __device__ __forceinline__ unsigned long long splitmix64(unsigned long long z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// v[i] = lo + span * u,  u = top 24 bits of splitmix64(seed * 2^40 + global_i) scaled by 2^-24
__global__ void __launch_bounds__(256)
fill_hash_kernel(float *v, i64 n, i64 idx_base, unsigned long long seed, float lo, float span)
{
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long h = splitmix64((seed << 40) + (unsigned long long)(idx_base + i));
        const float u = (float)(h >> 40) * 5.9604644775390625e-8f;   // exact: 24 bits * 2^-24
        v[i] = __fadd_rn(lo, __fmul_rn(span, u));
    }
}
__global__ void __launch_bounds__(256) fill_kernel(float *v, i64 n, float value)
{
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) v[i] = value;
}

}  // namespace lbk
