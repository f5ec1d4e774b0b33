// lbk_kernels.cuh -- sm_100a device code for the Level-1 path.
//
// Every routine of Numerical.BLAS.Single that touches a vector (reference:
// src/Numerical/BLAS/Single.hs, index semantics src/Numerical/BLAS/Util.hs:39-142) is one of
// two shapes, both bounded by HBM bandwidth (O(1) flop per 4..16 bytes, no data reuse, hence no
// shared-memory staging and no tensor cores):
//
//   map     x[kx(i)], y[ky(i)]  ->  x[kx(i)], y[ky(i)]          axpy scal copy swap rot rotm
//   reduce  x[kx(i)] (, y[ky(i)]) -> one record                 dot sdsdot asum nrm2 iamax
//
// and each shape has two kernels:
//   *_unit     |incx| == |incy| == 1 : 128- or 256-bit coalesced global accesses (LDG.E.128 /
//              LDG.E.256 on sm_100a), UNROLL independent packs in flight per thread per operand,
//              a grid-stride loop over tiles; opposite signs pair x[i] with y[n-1-i] and still move
//              whole aligned packs (reversed lanes);
//   *_strided  any increments (negative, zero, non-unit): scalar accesses, 64-bit index
//              arithmetic, kx(i) = first(n,inc) + i*inc (Single.hs:91-96).
//
// The reference's routines are polymorphic (`dot :: Num a => ...`, Single.hs:73) with Float and
// Double instances (sdot/ddot, ...); every kernel here is a template on the element type T.
//
// Reductions are deterministic for a given launch shape: per-thread accumulation in a fixed
// order, warp shuffle tree, fixed-order block combine, one record per block, and the last block
// to arrive (ticket) folds the block records in index order ("two-pass in one launch").
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace lbk {

typedef long long i64;
typedef unsigned long long u64;

constexpr int kMaxThreads = 512;        // launch bound of every streaming kernel
// Resident 512-thread blocks per SM the reductions are compiled for (a register cap of 40).  Left to
// itself ptxas spends 45-60 registers on hoisting the tail's address arithmetic, which costs a resident
// block per SM and 3-6% of the streaming rate (profiles/r01_tune_reductions_minblocks{1,3}.csv); the shapes with
// 256 bytes in flight per operand per thread cannot fit and keep the plain bound.
#ifndef LBK_REDUCE_MINB
#define LBK_REDUCE_MINB 3
#endif
constexpr int reduce_min_blocks(int bytes_per_operand) { return bytes_per_operand >= 256 ? 1 : LBK_REDUCE_MINB; }
constexpr u64 kNanKey = ~0ull;
// The shifted-load shapes (shift_pack) are launched with at most 256 threads and compiled for 4 resident blocks of
// them (64 registers), or for 3 (80 registers) where two operands are loaded or the accumulator is a double: those
// spill at 64, and 24 bytes of spill cost sdot four points (profiles/r02_misaligned.md).  ptxas -v: no shape that is a
// default spills (tests/test_sass.py keeps it so).
constexpr int kShiftThreads = 256;
constexpr int shift_min_blocks(bool heavy) { return heavy ? 3 : 4; }

// ------------------------------------------------------------------------------------------------
// element-type helpers: one IEEE operation each, never contracted (the library is built with
// -fmad=false and these are the *_rn intrinsics), so elementwise results are bit-identical to the
// reference's compiled code; fma_() is used only where a fused multiply-add is wanted.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float fma_(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double fma_(double a, double b, double c) { return fma(a, b, c); }
__host__ __device__ __forceinline__ float abs_(float a) { return fabsf(a); }
__host__ __device__ __forceinline__ double abs_(double a) { return fabs(a); }
__device__ __forceinline__ float max_(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double max_(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float min_(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ double min_(double a, double b) { return fmin(a, b); }
// bit pattern of |x| as an integer: monotone in |x| for non-NaN values
__device__ __forceinline__ u64 abs_bits(float x) { return (u64)(__float_as_uint(x) & 0x7fffffffu); }
__device__ __forceinline__ u64 abs_bits(double x) { return (u64)__double_as_longlong(x) & 0x7fffffffffffffffull; }
template <class T> __device__ __forceinline__ T from_abs_bits(u64 b);
template <> __device__ __forceinline__ float from_abs_bits<float>(u64 b) { return __uint_as_float((unsigned)b); }
template <> __device__ __forceinline__ double from_abs_bits<double>(u64 b) { return __longlong_as_double((i64)b); }

// ------------------------------------------------------------------------------------------------
// 128/256-bit global memory access.  A Pack is VEC elements = 4/8 (scalar), 16 or 32 bytes.
// POLICY 0: default caching.  POLICY >= 1: streaming -- do not allocate in L1 (each byte is touched
// once).  Read-only operands do NOT go through the non-coherent path (ld.global.nc): every unit-stride
// kernel can be a programmatic-dependent-launch secondary that is resident while its predecessor is
// still writing what it will read after griddepcontrol.wait, and PTX defines .nc only for data that
// stays read-only for the kernel's whole lifetime.  A plain ld.global.L1::no_allocate is served from L2
// at the same rate (profiles/r02_nc_ab.md); LBK_USE_NC=1 rebuilds the old qualifier for A/B runs only.
// POLICY 2..5 add L2 hints on 256-bit
// accesses (2/3: evict-first loads, 3/5: evict-first stores, 4: 256-byte L2 prefetch + st.cs); they were
// measured to make no difference (profiles/r01_tune_l2policy_writers.csv) and exist for tools/tune.py.
// POLICY 6 = POLICY 1 plus, in the map kernel, a data dependency that holds a thread's stores back until
// ALL of its loads have landed (see map_unit_kernel).
// ------------------------------------------------------------------------------------------------
template <class T, int VEC> struct alignas(VEC * sizeof(T)) Pack { T v[VEC]; };

// qualifier class of a load: 0 plain, 1/2 no_allocate (read-only / read-write), 3/4 + L2::evict_first,
// 5/6 + L2::256B (the last four only on 32-byte accesses)
#define LBK_LDQ(POLICY, RO, WIDE)                                                                   \
    ((POLICY) == 0 ? 0 : ((WIDE) && ((POLICY) == 2 || (POLICY) == 3)) ? ((RO) ? 3 : 4)              \
                   : ((WIDE) && (POLICY) == 4) ? ((RO) ? 5 : 6) : ((RO) ? 1 : 2))

// the qualifier has to be a string literal inside the asm, so the shapes are spelled out by macro
#define LBK_LD_F32x1(Q) asm volatile("ld.global" Q ".f32 %0, [%1];" : "=f"(p.v[0]) : "l"(ptr) : "memory")
#define LBK_LD_F32x4(Q) asm volatile("ld.global" Q ".v4.f32 {%0,%1,%2,%3}, [%4];"                   \
                                     : "=f"(p.v[0]), "=f"(p.v[1]), "=f"(p.v[2]), "=f"(p.v[3]) : "l"(ptr) : "memory")
#define LBK_LD_F32x8(Q) asm volatile("ld.global" Q ".v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"       \
                                     : "=f"(p.v[0]), "=f"(p.v[1]), "=f"(p.v[2]), "=f"(p.v[3]),        \
                                       "=f"(p.v[4]), "=f"(p.v[5]), "=f"(p.v[6]), "=f"(p.v[7]) : "l"(ptr) : "memory")
#define LBK_LD_F64x1(Q) asm volatile("ld.global" Q ".f64 %0, [%1];" : "=d"(p.v[0]) : "l"(ptr) : "memory")
#define LBK_LD_F64x2(Q) asm volatile("ld.global" Q ".v2.f64 {%0,%1}, [%2];" : "=d"(p.v[0]), "=d"(p.v[1]) : "l"(ptr) : "memory")
#define LBK_LD_F64x4(Q) asm volatile("ld.global" Q ".v4.f64 {%0,%1,%2,%3}, [%4];"                   \
                                     : "=d"(p.v[0]), "=d"(p.v[1]), "=d"(p.v[2]), "=d"(p.v[3]) : "l"(ptr) : "memory")
#ifndef LBK_USE_NC
#define LBK_USE_NC 0
#endif
#if LBK_USE_NC
#define LBK_NC ".nc"
#else
#define LBK_NC ""
#endif
#define LBK_LD_DISPATCH(SHAPE)                                                                      \
    do {                                                                                            \
        constexpr int q = LBK_LDQ(POLICY, RO, sizeof(T) * VEC == 32);                               \
        if constexpr (q == 0) SHAPE("");                                                            \
        else if constexpr (q == 1) SHAPE(LBK_NC ".L1::no_allocate");                                \
        else if constexpr (q == 2) SHAPE(".L1::no_allocate");                                       \
        else if constexpr (q == 3) SHAPE(LBK_NC ".L1::no_allocate.L2::evict_first");                \
        else if constexpr (q == 4) SHAPE(".L1::no_allocate.L2::evict_first");                       \
        else if constexpr (q == 5) SHAPE(LBK_NC ".L1::no_allocate.L2::256B");                       \
        else SHAPE(".L1::no_allocate.L2::256B");                                                    \
    } while (0)

template <int VEC, int POLICY, bool RO>
__device__ __forceinline__ void load_pack(Pack<float, VEC> &p, const float *ptr)
{
    using T = float;
    if constexpr (VEC == 1) LBK_LD_DISPATCH(LBK_LD_F32x1);
    else if constexpr (VEC == 4) LBK_LD_DISPATCH(LBK_LD_F32x4);
    else { static_assert(VEC == 8, "float packs hold 1, 4 or 8 elements"); LBK_LD_DISPATCH(LBK_LD_F32x8); }
}
template <int VEC, int POLICY, bool RO>
__device__ __forceinline__ void load_pack(Pack<double, VEC> &p, const double *ptr)
{
    using T = double;
    if constexpr (VEC == 1) LBK_LD_DISPATCH(LBK_LD_F64x1);
    else if constexpr (VEC == 2) LBK_LD_DISPATCH(LBK_LD_F64x2);
    else { static_assert(VEC == 4, "double packs hold 1, 2 or 4 elements"); LBK_LD_DISPATCH(LBK_LD_F64x4); }
}

#define LBK_ST_F32x1(Q) asm volatile("st.global" Q ".f32 [%0], %1;" ::"l"(ptr), "f"(p.v[0]) : "memory")
#define LBK_ST_F32x4(Q) asm volatile("st.global" Q ".v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(ptr),       \
                                     "f"(p.v[0]), "f"(p.v[1]), "f"(p.v[2]), "f"(p.v[3]) : "memory")
#define LBK_ST_F32x8(Q) asm volatile("st.global" Q ".v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(ptr), \
                                     "f"(p.v[0]), "f"(p.v[1]), "f"(p.v[2]), "f"(p.v[3]),              \
                                     "f"(p.v[4]), "f"(p.v[5]), "f"(p.v[6]), "f"(p.v[7]) : "memory")
#define LBK_ST_F64x1(Q) asm volatile("st.global" Q ".f64 [%0], %1;" ::"l"(ptr), "d"(p.v[0]) : "memory")
#define LBK_ST_F64x2(Q) asm volatile("st.global" Q ".v2.f64 [%0], {%1,%2};" ::"l"(ptr), "d"(p.v[0]), "d"(p.v[1]) : "memory")
#define LBK_ST_F64x4(Q) asm volatile("st.global" Q ".v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(ptr),       \
                                     "d"(p.v[0]), "d"(p.v[1]), "d"(p.v[2]), "d"(p.v[3]) : "memory")
#define LBK_ST_DISPATCH(SHAPE)                                                                      \
    do {                                                                                            \
        constexpr bool wide = sizeof(T) * VEC == 32;                                                \
        if constexpr (POLICY == 0 || sizeof(T) * VEC <= 8) SHAPE("");                               \
        else if constexpr (wide && (POLICY == 3 || POLICY == 5)) SHAPE(".L1::no_allocate.L2::evict_first"); \
        else if constexpr (wide && POLICY == 4) SHAPE(".cs");                                       \
        else SHAPE(".L1::no_allocate");                                                             \
    } while (0)

template <int VEC, int POLICY>
__device__ __forceinline__ void store_pack(float *ptr, const Pack<float, VEC> &p)
{
    using T = float;
    if constexpr (VEC == 1) LBK_ST_DISPATCH(LBK_ST_F32x1);
    else if constexpr (VEC == 4) LBK_ST_DISPATCH(LBK_ST_F32x4);
    else LBK_ST_DISPATCH(LBK_ST_F32x8);
}
template <int VEC, int POLICY>
__device__ __forceinline__ void store_pack(double *ptr, const Pack<double, VEC> &p)
{
    using T = double;
    if constexpr (VEC == 1) LBK_ST_DISPATCH(LBK_ST_F64x1);
    else if constexpr (VEC == 2) LBK_ST_DISPATCH(LBK_ST_F64x2);
    else LBK_ST_DISPATCH(LBK_ST_F64x4);
}

// ------------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  Between two kernels of a stream the GPU idles for a few
// microseconds (drain of the last blocks, launch, ramp-up), and a reduction adds a latency-bound tail
// (block combine, ticket, last-block fold): together as much as 5% of a 4-byte-per-element kernel at
// n = 2^28.  Every unit-stride kernel therefore releases its dependents AT ONCE
// (griddepcontrol.launch_dependents as its first instruction): the next kernel of the stream, when the
// host launched it with the programmatic-serialization attribute, fills SMs as this one's blocks retire.
// What keeps that correct:
//   * the host only sets the attribute when the next kernel reads nothing that a still-running
//     predecessor writes (lbk_lib.cu, pdl_decide);
//   * a kernel that WRITES something a predecessor reads or writes is told to execute
//     griddepcontrol.wait (= all predecessors complete and visible) before its first store;
//   * a reduction executes it before it touches the shared scratch, its result or the mailbox;
//   * every kernel executes it before it exits, so completion stays in stream order.
// Both instructions are no-ops in a kernel launched without the attribute.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// Shifted loads.  Unit-stride operands that are misaligned DIFFERENTLY cannot share pack boundaries: with the body
// laid out on the other operands' boundaries, logical pack p of the read-only operand x starts XSH elements
// (0 < XSH < VEC) past a boundary.  Instead of falling back to one element per access, a thread loads ONE aligned
// pack and takes the few elements it lacks from a neighbouring lane's pack with a shuffle -- consecutive lanes own
// consecutive packs.  Whichever is fewer:
//   down (XSH <= VEC/2): the pack that holds the thread's first VEC-XSH elements; the last XSH come from lane+1;
//   up   (XSH >  VEC/2): the pack that holds its last XSH elements; the first VEC-XSH come from lane-1.
// The lane at the warp's edge (31 / 0) has no such neighbour and reads those elements -- its own -- one by one: at
// most VEC/2 four-byte requests per warp and row, inside sectors the neighbouring warp fetches anyway.  The host
// leaves at least VEC-XSH tail elements (and peels at least XSH head elements), so every aligned pack that is
// touched lies inside the vector.  XSH is a template argument: the lane renaming is free, what is added per pack
// is min(XSH, VEC-XSH) shuffles (profiles/r02_misaligned.md).
// ------------------------------------------------------------------------------------------------
template <int VEC, int XSH> struct Shift {
    static constexpr bool DOWN = XSH > 0 && 2 * XSH <= VEC;
    static constexpr int E = XSH <= 0 ? 1 : (DOWN ? XSH : VEC - XSH);     // elements taken from the neighbour
    static constexpr int PACK0 = DOWN ? -XSH : VEC - XSH;                // the aligned pack, relative to the thread's first element
    static constexpr int EDGE0 = DOWN ? VEC - XSH : 0;                   // the edge lane's own elements, likewise
    static constexpr unsigned EDGE_LANE = DOWN ? 31u : 0u;
};
// a: the aligned pack; ex: on the edge lane, the E elements it read itself
template <class T, int VEC, int XSH>
__device__ __forceinline__ void shift_pack(Pack<T, VEC> &out, const Pack<T, VEC> &a, const T (&ex)[Shift<VEC, XSH>::E], bool edge_lane)
{
    using S = Shift<VEC, XSH>;
    T nb[S::E];
#pragma unroll
    for (int j = 0; j < S::E; ++j) {
        const T v = S::DOWN ? __shfl_down_sync(0xffffffffu, a.v[j], 1) : __shfl_up_sync(0xffffffffu, a.v[VEC - S::E + j], 1);
        nb[j] = edge_lane ? ex[j] : v;
    }
#pragma unroll
    for (int l = 0; l < VEC; ++l) {
        if constexpr (S::DOWN) out.v[l] = (l + XSH < VEC) ? a.v[(l + XSH) % VEC] : nb[(l + XSH) % VEC];
        else out.v[l] = (l < S::E) ? nb[l % S::E] : a.v[(l + VEC - S::E) % VEC];
    }
}
// Loads of one row: the aligned pack and, on the edge lane, its own E elements.
template <class T, int VEC, int XSH, int POLICY>
__device__ __forceinline__ void shift_load(Pack<T, VEC> &a, T (&ex)[Shift<VEC, XSH>::E], const T *first, bool edge_lane)
{
    using S = Shift<VEC, XSH>;
    load_pack<VEC, POLICY, true>(a, first + S::PACK0);
    if (edge_lane) {
#pragma unroll
        for (int j = 0; j < S::E; ++j) { Pack<T, 1> p; load_pack<1, POLICY, true>(p, first + S::EDGE0 + j); ex[j] = p.v[0]; }
    }
}
// packs of the body when x is read through shifted loads: the last aligned pack touched must end inside the vector
template <int VEC, int XSH>
__host__ __device__ __forceinline__ i64 body_packs(i64 n, i64 head)
{
    const i64 body = n - head - (XSH > 0 ? VEC - XSH : 0);
    return body > 0 ? body / VEC : 0;
}

// ------------------------------------------------------------------------------------------------
// map functors.  (x, y) are the OLD values; the functor overwrites them with the new ones.
// ------------------------------------------------------------------------------------------------
template <class T> struct AxpyF {   // Single.hs:439   y + a*x
    using Scalar = T;
    T a;
    static constexpr bool RX = true, RY = true, WX = false, WY = true;
    static constexpr bool XSHIFT = true;      // built with shifted loads of x as well (see shift_pack)
    __device__ __forceinline__ void operator()(T &x, T &y) const { y = add_rn(y, mul_rn(a, x)); }
};
template <class T> struct ScalF {   // Single.hs:110   x*a
    using Scalar = T;
    T a;
    static constexpr bool RX = true, RY = false, WX = true, WY = false;
    static constexpr bool XSHIFT = true;      // out of place only: in place, source and destination are one pointer
    __device__ __forceinline__ void operator()(T &x, T &) const { x = mul_rn(x, a); }
};
template <class T> struct CopyF {   // Single.hs:127   \_ x -> x
    using Scalar = T;
    static constexpr bool RX = true, RY = false, WX = false, WY = true;
    static constexpr bool XSHIFT = true;
    __device__ __forceinline__ void operator()(T &x, T &y) const { y = x; }
};
template <class T> struct FillF {   // no reference counterpart: lbk_vec_fill (test / bench set-up), a pure store stream
    using Scalar = T;
    static constexpr bool XSHIFT = false;
    T value;
    static constexpr bool RX = false, RY = false, WX = true, WY = false;
    __device__ __forceinline__ void operator()(T &x, T &) const { x = value; }
};
template <class T> struct SwapF {   // Single.hs:143-145
    using Scalar = T;
    static constexpr bool XSHIFT = false;
    static constexpr bool RX = true, RY = true, WX = true, WY = true;
    __device__ __forceinline__ void operator()(T &x, T &y) const { T t = x; x = y; y = t; }
};
template <class T> struct RotF {    // Single.hs:418-419   c*x + s*y ; c*y - s*x
    using Scalar = T;
    static constexpr bool XSHIFT = false;
    T c, s;
    static constexpr bool RX = true, RY = true, WX = true, WY = true;
    __device__ __forceinline__ void operator()(T &x, T &y) const
    {
        T nx = add_rn(mul_rn(c, x), mul_rn(s, y));
        T ny = sub_rn(mul_rn(c, y), mul_rn(s, x));
        x = nx; y = ny;
    }
};
template <class T> struct RotmNeg1F {   // Single.hs:475-476   h11*x + h12*y ; h21*x + h22*y
    using Scalar = T;
    static constexpr bool XSHIFT = false;
    T h11, h12, h21, h22;
    static constexpr bool RX = true, RY = true, WX = true, WY = true;
    __device__ __forceinline__ void operator()(T &x, T &y) const
    {
        T nx = add_rn(mul_rn(h11, x), mul_rn(h12, y));
        T ny = add_rn(mul_rn(h21, x), mul_rn(h22, y));
        x = nx; y = ny;
    }
};
template <class T> struct Rotm0F {      // Single.hs:478-479   x + h12*y ; h21*x + y
    using Scalar = T;
    static constexpr bool XSHIFT = false;
    T h12, h21;
    static constexpr bool RX = true, RY = true, WX = true, WY = true;
    __device__ __forceinline__ void operator()(T &x, T &y) const
    {
        T nx = add_rn(x, mul_rn(h12, y));
        T ny = add_rn(mul_rn(h21, x), y);
        x = nx; y = ny;
    }
};
template <class T> struct Rotm1F {      // Single.hs:481-482   h11*x + y ; -x + h22*y
    using Scalar = T;
    static constexpr bool XSHIFT = false;
    T h11, h22;
    static constexpr bool RX = true, RY = true, WX = true, WY = true;
    __device__ __forceinline__ void operator()(T &x, T &y) const
    {
        T nx = add_rn(mul_rn(h11, x), y);
        T ny = add_rn(-x, mul_rn(h22, y));
        x = nx; y = ny;
    }
};

// ------------------------------------------------------------------------------------------------
// map, unit stride.  Reads come from (xr, yr), writes go to (xw, yw); in place when they are equal,
// out of place (the reference's pure result) when they differ.  `head` scalar elements precede the
// pack body so that xr+head (and yr/xw/yw+head: checked by the host) are pack aligned; they and the
// < VEC tail are handled by the first threads of block 0.
//
// REV: the (incx, incy) = (1,-1) / (-1,1) pairing -- logical element i is x[i] and y[n-1-i] -- still moves
// whole aligned packs: the y pack for x[e..e+VEC) is y[n-e-VEC..n-e) with its lanes in reverse order.
// ------------------------------------------------------------------------------------------------
// Load fence (POLICY 6).  A warp issues in order, so a store whose data has landed goes out while the
// same warp's later loads are still in flight; for the routines with no arithmetic between load and
// store (swap, copy) that interleaving costs up to 15% at some launch shapes, and holding every store
// until the thread's last load has returned never loses (tools/swap_experiment.cu,
// profiles/r01_swap_instruction_stream.log).  There is no "wait for my loads" instruction, so the fence
// is a true dependency: one element of every loaded pack is OR-ed together, masked with a kernel
// argument that is 0 at run time (the compiler cannot know), and added to every store address.
__device__ __forceinline__ unsigned low_bits(float v) { return __float_as_uint(v); }
__device__ __forceinline__ unsigned low_bits(double v) { return (unsigned)__double2loint(v); }

// The shapes that hold at most 32 bytes of loaded data per thread live on occupancy (2048 threads per SM):
// they are compiled for 4 resident 512-thread blocks, i.e. 32 registers -- two registers more cost a quarter of
// the resident threads and 5% of the bandwidth (dcopy 7020 -> 6540, drot 6900 -> 6600 GB/s when ptxas chose 34 / 40).
// Up to 128 bytes they are compiled for 2 such blocks (64 registers): left alone, ptxas has given saxpy's default
// shape anything from 58 to 72 registers from one build to the next, and 72 halves its resident threads
// (7130 -> 6590 GB/s).
constexpr int map_min_blocks(int loaded_bytes_per_thread) { return loaded_bytes_per_thread <= 32 ? 4 : loaded_bytes_per_thread <= 128 ? 2 : 1; }

template <class F, int VEC, int UNROLL, int POLICY, bool REV = false, int XSH = 0>
__global__ void __launch_bounds__(XSH > 0 ? kShiftThreads : kMaxThreads, XSH > 0 ? shift_min_blocks(F::RX && F::RY) : map_min_blocks(VEC * UNROLL * (int)sizeof(typename F::Scalar) * ((F::RX ? 1 : 0) + (F::RY ? 1 : 0))))
map_unit_kernel(F f, const typename F::Scalar *xr, const typename F::Scalar *yr, typename F::Scalar *xw,
                typename F::Scalar *yw, i64 head, i64 n, unsigned fence_mask, int wait_mode)
{
    using T = typename F::Scalar;
    static_assert(XSH == 0 || (XSH > 0 && XSH < VEC && !REV && F::RX), "shifted loads: 0 < XSH < VEC, forward pairing, x is read");
    constexpr bool FENCE = POLICY == 6 && (F::RX || F::RY);
    // wait_mode, low two bits (lbk_pdl.hpp): 0 wait for the predecessors only before exiting, 1 before the first store,
    // 2 before the first load -- and in that case release the dependents only afterwards.  Bit 2: walk the tiles from the
    // BACK of the vectors to the front.  An elementwise result cannot depend on the order, and the tail of what the previous
    // kernel streamed is what the 126 MB L2 still holds: starting there turns ~100 MB of a 2^28-element saxpy's DRAM reads
    // into L2 hits (lbk_lib.cu, Traversal; profiles/r02_serpentine.md).
    const bool backwards = (wait_mode & 4) != 0;
    wait_mode &= 3;
    const int wait_before_store = wait_mode == 1;
    if (wait_mode == 2) pdl_wait();
    pdl_launch_dependents();
    const i64 npack = XSH > 0 ? body_packs<VEC, XSH>(n, head) : (n - head) / VEC;
    const i64 per_tile = (i64)UNROLL * blockDim.x;
    const i64 ntiles = (npack + per_tile - 1) / per_tile;
    const T *bxr = xr + head;
    T *bxw = xw + head;
    // y side: forward from y+head, or backward from one past the body's last y element
    const T *byr = REV ? yr + (n - head) - VEC : yr + head;
    T *byw = REV ? yw + (n - head) - VEC : yw + head;
    constexpr int YS = REV ? -1 : 1;
    // shifted loads of x: a whole warp takes the fast path together (its last lane decides), the shuffles need every lane
    const bool edge_lane = (threadIdx.x & 31) == Shift<VEC, XSH>::EDGE_LANE;

    for (i64 t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const i64 tile = backwards ? ntiles - 1 - t : t;
        const i64 p0 = tile * per_tile + threadIdx.x;
        Pack<T, VEC> xv[UNROLL], yv[UNROLL];
        bool whole;                                          // whole tile in range for this thread (XSH: for its warp)
        if constexpr (XSH > 0) whole = (p0 | 31) + (i64)(UNROLL - 1) * blockDim.x < npack;
        else whole = p0 + (i64)(UNROLL - 1) * blockDim.x < npack;
        if (whole) {
            if constexpr (XSH > 0) {
                T ex[UNROLL][Shift<VEC, XSH>::E];
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) {
                    const i64 e = (p0 + (i64)u * blockDim.x) * VEC;
                    shift_load<T, VEC, XSH, POLICY>(xv[u], ex[u], bxr + e, edge_lane);
                    if constexpr (F::RY) load_pack<VEC, POLICY, !F::WY>(yv[u], byr + e);
                }
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) { const Pack<T, VEC> a = xv[u]; shift_pack<T, VEC, XSH>(xv[u], a, ex[u], edge_lane); }
            } else {
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) {
                    const i64 e = (p0 + (i64)u * blockDim.x) * VEC;
                    if constexpr (F::RX) load_pack<VEC, POLICY, !F::WX>(xv[u], bxr + e);
                    if constexpr (F::RY) load_pack<VEC, POLICY, !F::WY>(yv[u], byr + YS * e);
                }
            }
            i64 held = 0;
            if constexpr (FENCE) {
                unsigned dep = 0;
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) {
                    if constexpr (F::RX) dep |= low_bits(xv[u].v[VEC - 1]);
                    if constexpr (F::RY) dep |= low_bits(yv[u].v[VEC - 1]);
                }
                held = (i64)(dep & fence_mask);      // 0: fence_mask is 0
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
                for (int l = 0; l < VEC; ++l) f(xv[u].v[l], yv[u].v[REV ? VEC - 1 - l : l]);
            }
            if (wait_before_store) pdl_wait();      // a predecessor may still be reading / writing these addresses
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const i64 e = (p0 + (i64)u * blockDim.x) * VEC + held;
                if constexpr (F::WX) store_pack<VEC, POLICY>(bxw + e, xv[u]);
                if constexpr (F::WY) store_pack<VEC, POLICY>(byw + YS * e, yv[u]);
            }
        } else {
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const i64 p = p0 + (i64)u * blockDim.x;
                if (p < npack) {
                    const i64 e = p * VEC;
                    if constexpr (XSH > 0) {             // the ragged last tile: x one element at a time
#pragma unroll
                        for (int l = 0; l < VEC; ++l) xv[u].v[l] = bxr[e + l];
                    } else if constexpr (F::RX) load_pack<VEC, POLICY, !F::WX>(xv[u], bxr + e);
                    if constexpr (F::RY) load_pack<VEC, POLICY, !F::WY>(yv[u], byr + YS * e);
#pragma unroll
                    for (int l = 0; l < VEC; ++l) f(xv[u].v[l], yv[u].v[REV ? VEC - 1 - l : l]);
                    if (wait_before_store) pdl_wait();
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
            T xs = 0, ys = 0;
            if constexpr (F::RX) xs = xr[e];
            if constexpr (F::RY) ys = yr[ey];
            f(xs, ys);
            if (wait_before_store) pdl_wait();
            if constexpr (F::WX) xw[e] = xs;
            if constexpr (F::WY) yw[ey] = ys;
        }
    }
    pdl_wait();     // as a PDL secondary: do not retire before the predecessors have (completion stays in stream order)
}

// ------------------------------------------------------------------------------------------------
// map, any increments.  Logical element i lives at x[kx0 + i*incx], y[ky0 + i*incy] with
// kx0 = first(n,incx) (Single.hs:91-96, Util.hs:81-88,141-142).  Right-hand sides always read the
// ORIGINAL values (Util.hs:134-142): when an OUTPUT increment is 0 the host passes a snapshot of
// element 0 as the read pointer and sets *_last_only, so that only logical element n-1 (the last
// write of the reference's sequential scatter) stores.
// ------------------------------------------------------------------------------------------------
constexpr int kStridedUnroll = 4;   // independent scalar accesses in flight per operand per thread

template <class F, bool FENCED>
__global__ void __launch_bounds__(256)
map_strided_kernel(F f, const typename F::Scalar *xr, typename F::Scalar *xw, i64 incx_r, i64 incx_w, i64 kx0_r, i64 kx0_w,
                   const typename F::Scalar *yr, typename F::Scalar *yw, i64 incy_r, i64 incy_w, i64 ky0_r, i64 ky0_w,
                   i64 i_begin, i64 n, int x_last_only, int y_last_only, unsigned fence_mask, int wait_mode)
{
    using T = typename F::Scalar;
    constexpr int U = kStridedUnroll;
    constexpr bool FENCE = FENCED && (F::RX || F::RY);      // the load fence of map_unit_kernel: stores wait for all loads
    // programmatic dependent launch, as in map_unit_kernel: 0 wait only before exiting, 1 before the first store, 2 before the first load
    const int wait_before_store = wait_mode == 1;
    if (wait_mode == 2) pdl_wait();
    pdl_launch_dependents();
    const i64 per_tile = (i64)U * blockDim.x;
    const i64 count = n - i_begin;
    const i64 ntiles = (count + per_tile - 1) / per_tile;
    for (i64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const i64 i0 = i_begin + tile * per_tile + threadIdx.x;
        T xs[U], ys[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {            // all loads first: U independent requests per operand
            const i64 i = i0 + (i64)u * blockDim.x;
            xs[u] = 0; ys[u] = 0;
            if (i < n) {
                Pack<T, 1> p;
                if constexpr (F::RX) { load_pack<1, 0, !F::WX>(p, xr + kx0_r + i * incx_r); xs[u] = p.v[0]; }
                if constexpr (F::RY) { load_pack<1, 0, !F::WY>(p, yr + ky0_r + i * incy_r); ys[u] = p.v[0]; }
            }
        }
        i64 held = 0;
        if constexpr (FENCE) {
            unsigned dep = 0;
#pragma unroll
            for (int u = 0; u < U; ++u) dep |= low_bits(xs[u]) | low_bits(ys[u]);
            held = (i64)(dep & fence_mask);      // 0: fence_mask is 0 at run time
        }
        if (wait_before_store) pdl_wait();       // a predecessor may still be reading / writing these addresses
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const i64 i = i0 + (i64)u * blockDim.x;
            if (i < n) {
                f(xs[u], ys[u]);
                if constexpr (F::WX) { if (!x_last_only || i == n - 1) xw[kx0_w + i * incx_w + held] = xs[u]; }
                if constexpr (F::WY) { if (!y_last_only || i == n - 1) yw[ky0_w + i * incy_w + held] = ys[u]; }
            }
        }
    }
    pdl_wait();     // completion stays in stream order
}

// ------------------------------------------------------------------------------------------------
// reductions
// ------------------------------------------------------------------------------------------------
// The record a GPU contributes to a sharded reduction, and the slot a block's partial result lives in
// between the two passes.  Layout shared with the host (lbk_record in include/lbk.h): f[0..2] partial
// sums (held in double, so float partials are exact), key/idx the iamax candidate (or, for nrm2, the
// bits of |x[0]| for the n == 1 rule).
struct Record {
    double f[3];
    u64 key;
    i64 idx;
    i64 pad;
};
static_assert(sizeof(Record) == 48, "Record is 48 bytes");
constexpr int kRecordWords = sizeof(Record) / 4;

__host__ __device__ __forceinline__ Record empty_record()
{
    Record r;
    r.f[0] = r.f[1] = r.f[2] = 0.0;
    r.key = 0ull;
    r.idx = -1;
    r.pad = 0;
    return r;
}

// Blue's scaled accumulation (as in LAPACK 3.10 *nrm2): |x| > big is accumulated as (x*sbig)^2,
// |x| < sml as (x*ssml)^2, everything else as x^2.  All three are plain sums, so the cross-block and
// cross-GPU combine is addition.  The float thresholds are tightened so that 2^36 mid-range squares
// cannot overflow binary32; the double ones are LAPACK's.
template <class T> struct Blue;
template <> struct Blue<float> {
    static constexpr float big = 70368744177664.0f;            // 2^46
    static constexpr float sml = 1.0842021724855044e-19f;      // 2^-63
    static constexpr float sbig = 1.3234889800848443e-23f;     // 2^-76
    static constexpr float ssml = 3.7778931862957162e+22f;     // 2^75
};
template <> struct Blue<double> {
    static constexpr double big = 1.9979190722022350e+146;     // 2^486
    static constexpr double sml = 1.4916681462400413e-154;     // 2^-511
    static constexpr double sbig = 1.1113793747425387e-162;    // 2^-538
    static constexpr double ssml = 4.4989137945431964e+161;    // 2^537
};

// Op concept
//   Scalar, Acc, zero(), step(acc,x,y,ord) [or step_pack]      the streaming loop (registers only)
//   Partial, empty(), merge(a,b), shfl(p,mask)                  what a thread / warp / block contributes: as
//                                                               few registers as the routine needs, so that the
//                                                               tail does not set the kernel's register count
//   add_acc(p,acc), single(p,x,y,idx)                           accumulator / one scalar element -> partial
//   store(slot,p), load(slot)                                   a block's partial in its Record slot (only the
//                                                               fields the routine uses are moved)
// Partial sums are doubles for both element types: a float accumulator converts exactly, and the
// combine tree (warp, block, grid, GPUs) then adds without further float rounding.
struct SumPartial { double s; };
template <class Derived> struct SumOpBase {
    using Partial = SumPartial;
    __device__ static __forceinline__ Partial empty() { return {0.0}; }
    __device__ static __forceinline__ Partial merge(const Partial &a, const Partial &b) { return {a.s + b.s}; }
    __device__ static __forceinline__ Partial shfl(const Partial &p, int m) { return {__shfl_xor_sync(0xffffffffu, p.s, m)}; }
    __device__ static __forceinline__ void store(Record *slot, const Partial &p) { __stcg(&slot->f[0], p.s); }
    __device__ static __forceinline__ Partial load(const Record *slot) { return {__ldcg(&slot->f[0])}; }
    __device__ static __forceinline__ void to_record(Record &r, const Partial &p) { r.f[0] = p.s; }
};
template <class T> struct DotOp : SumOpBase<DotOp<T>> {       // Single.hs:81-82
    using Scalar = T;
    using Partial = SumPartial;
    static constexpr bool TWO = true, IS_IAMAX = false, HAS_PACK_STEP = false;
    struct Acc { T s; };
    __device__ static __forceinline__ Acc zero() { return {0}; }
    __device__ static __forceinline__ void step(Acc &a, T x, T y, unsigned) { a.s = fma_(x, y, a.s); }
    __device__ static __forceinline__ void add_acc(Partial &p, const Acc &a) { p.s += (double)a.s; }
    __device__ static __forceinline__ void single(Partial &p, T x, T y, i64) { p.s = fma((double)x, (double)y, p.s); }
};
struct DsdotOp : SumOpBase<DsdotOp> {     // Single.hs:238-262: float inputs, products and sum in double
    using Scalar = float;
    using Partial = SumPartial;
    static constexpr bool TWO = true, IS_IAMAX = false, HAS_PACK_STEP = false;
    struct Acc { double s; };
    __device__ static __forceinline__ Acc zero() { return {0.0}; }
    __device__ static __forceinline__ void step(Acc &a, float x, float y, unsigned) { a.s = fma((double)x, (double)y, a.s); }
    __device__ static __forceinline__ void add_acc(Partial &p, const Acc &a) { p.s += a.s; }
    __device__ static __forceinline__ void single(Partial &p, float x, float y, i64) { p.s = fma((double)x, (double)y, p.s); }
};
template <class T> struct AsumOp : SumOpBase<AsumOp<T>> {     // Single.hs:163
    using Scalar = T;
    using Partial = SumPartial;
    static constexpr bool TWO = false, IS_IAMAX = false, HAS_PACK_STEP = false;
    struct Acc { T s; };
    __device__ static __forceinline__ Acc zero() { return {0}; }
    __device__ static __forceinline__ void step(Acc &a, T x, T, unsigned) { a.s += abs_(x); }
    __device__ static __forceinline__ void add_acc(Partial &p, const Acc &a) { p.s += (double)a.s; }
    __device__ static __forceinline__ void single(Partial &p, T x, T, i64) { p.s += (double)abs_(x); }
};
template <class T> struct Nrm2Op {      // Single.hs:180-199, restated as Blue's three accumulators
    using Scalar = T;
    static constexpr bool TWO = false, IS_IAMAX = false, HAS_PACK_STEP = true;
    struct Acc { T big, med, sml; };
    struct Partial { double big, med, sml; };
    __device__ static __forceinline__ Acc zero() { return {0, 0, 0}; }
    __device__ static __forceinline__ void step(Acc &a, T x, T, unsigned)
    {
        const T ax = abs_(x);
        if (ax > Blue<T>::big) { const T t = ax * Blue<T>::sbig; a.big = fma_(t, t, a.big); }
        else if (ax < Blue<T>::sml) { const T t = ax * Blue<T>::ssml; a.sml = fma_(t, t, a.sml); }
        else a.med = fma_(ax, ax, a.med);      // NaN lands here and propagates
    }
    // One pack of lanes: when every lane is mid-range (the overwhelmingly common case) the pack costs
    // one FMA and two min/max per element; a pack holding a huge, tiny or zero lane takes the exact
    // per-lane classification.  max_/min_ skip NaN lanes, and the fast path's x*x propagates them.
    template <int VEC>
    __device__ static __forceinline__ void step_pack(Acc &a, const Pack<T, VEC> &x, const Pack<T, VEC> &, unsigned)
    {
        T mx = abs_(x.v[0]), mn = mx;
#pragma unroll
        for (int l = 1; l < VEC; ++l) { mx = max_(mx, abs_(x.v[l])); mn = min_(mn, abs_(x.v[l])); }
        if (mx <= Blue<T>::big && mn >= Blue<T>::sml) {
#pragma unroll
            for (int l = 0; l < VEC; ++l) a.med = fma_(x.v[l], x.v[l], a.med);
        } else {
#pragma unroll
            for (int l = 0; l < VEC; ++l) step(a, x.v[l], T(0), 0u);
        }
    }
    __device__ static __forceinline__ Partial empty() { return {0.0, 0.0, 0.0}; }
    __device__ static __forceinline__ Partial merge(const Partial &a, const Partial &b) { return {a.big + b.big, a.med + b.med, a.sml + b.sml}; }
    __device__ static __forceinline__ Partial shfl(const Partial &p, int m)
    { return {__shfl_xor_sync(0xffffffffu, p.big, m), __shfl_xor_sync(0xffffffffu, p.med, m), __shfl_xor_sync(0xffffffffu, p.sml, m)}; }
    __device__ static __forceinline__ void add_acc(Partial &p, const Acc &a) { p.big += (double)a.big; p.med += (double)a.med; p.sml += (double)a.sml; }
    __device__ static __forceinline__ void single(Partial &p, T x, T, i64)
    {   // the strided / scalar path classifies in T and accumulates in the partial's doubles
        const T ax = abs_(x);
        if (ax > Blue<T>::big) { const double t = (double)(ax * Blue<T>::sbig); p.big = fma(t, t, p.big); }
        else if (ax < Blue<T>::sml) { const double t = (double)(ax * Blue<T>::ssml); p.sml = fma(t, t, p.sml); }
        else { const double t = (double)ax; p.med = fma(t, t, p.med); }
    }
    __device__ static __forceinline__ void store(Record *slot, const Partial &p) { __stcg(&slot->f[0], p.big); __stcg(&slot->f[1], p.med); __stcg(&slot->f[2], p.sml); }
    __device__ static __forceinline__ Partial load(const Record *slot) { return {__ldcg(&slot->f[0]), __ldcg(&slot->f[1]), __ldcg(&slot->f[2])}; }
    __device__ static __forceinline__ void to_record(Record &r, const Partial &p) { r.f[0] = p.big; r.f[1] = p.med; r.f[2] = p.sml; }
};
// iamax (Single.hs:217-220).  key = bit pattern of |x| (monotone for non-NaN), NaN -> 0 so that it
// can never displace an earlier element (`|best| < |NaN|` is false in the reference's fold); the
// "NaN at logical index 0 is never displaced" half of the rule is one extra candidate (all-ones key,
// index 0) contributed by the thread that owns element 0.
template <class T> struct AmaxKey { typedef unsigned type; };
template <> struct AmaxKey<double> { typedef u64 type; };
template <class T> __device__ __forceinline__ typename AmaxKey<T>::type amax_key(T x)
{
    return (x != x) ? (typename AmaxKey<T>::type)0 : (typename AmaxKey<T>::type)abs_bits(x);
}
template <class T> struct IamaxOp {
    using Scalar = T;
    using Key = typename AmaxKey<T>::type;
    static constexpr bool TWO = false, IS_IAMAX = true, HAS_PACK_STEP = true;
    static constexpr Key kKeyMax = ~(Key)0;
    // running maximum of |x| (-1 = nothing seen yet; any |x| >= 0 beats it) and the ordinal, within this
    // thread's element sequence, of the FIRST element that reached it
    struct Acc { T val; unsigned ord; };
    struct Partial { i64 idx; Key key; };      // idx < 0: no candidate
    __device__ static __forceinline__ Acc zero() { return {T(-1), 0u}; }
    __device__ static __forceinline__ void step(Acc &a, T x, T, unsigned ord)
    {
        const T ax = abs_(x);
        if (ax > a.val) { a.val = ax; a.ord = ord; }      // false for NaN: a NaN never displaces
    }
    // elements reach a thread in increasing index order, so strict '>' keeps the first maximum.  Per
    // pack: one max per element (max_ drops NaN lanes, which is the reference's "NaN never wins");
    // only a pack that raises the thread's maximum (O(log) times per thread) looks for the lane.
    template <int VEC>
    __device__ static __forceinline__ void step_pack(Acc &a, const Pack<T, VEC> &x, const Pack<T, VEC> &, unsigned ord0)
    {
        T m = abs_(x.v[0]);
#pragma unroll
        for (int l = 1; l < VEC; ++l) m = max_(m, abs_(x.v[l]));
        if (m > a.val) {
            unsigned lane = 0;
#pragma unroll
            for (int l = VEC - 1; l >= 0; --l) if (abs_(x.v[l]) == m) lane = l;
            a.val = m; a.ord = ord0 + lane;
        }
    }
    __device__ static __forceinline__ Partial empty() { return {-1, 0}; }
    __device__ static __forceinline__ Partial nan_at_zero() { return {0, kKeyMax}; }
    __device__ static __forceinline__ Partial candidate(T absval, i64 idx) { return {idx, (Key)abs_bits(absval)}; }
    // larger key wins; on a tie the lower index
    __device__ static __forceinline__ Partial merge(const Partial &a, const Partial &b)
    {
        if (b.idx < 0) return a;
        if (a.idx < 0) return b;
        if (b.key > a.key || (b.key == a.key && b.idx < a.idx)) return b;
        return a;
    }
    __device__ static __forceinline__ Partial shfl(const Partial &p, int m)
    { return {__shfl_xor_sync(0xffffffffu, p.idx, m), __shfl_xor_sync(0xffffffffu, p.key, m)}; }
    __device__ static __forceinline__ void single(Partial &p, T x, T, i64 idx) { p = merge(p, Partial{idx, amax_key(x)}); }
    __device__ static __forceinline__ void store(Record *slot, const Partial &p) { __stcg(&slot->key, (u64)p.key); __stcg(&slot->idx, p.idx); }
    __device__ static __forceinline__ Partial load(const Record *slot) { return {__ldcg(&slot->idx), (Key)__ldcg(&slot->key)}; }
    __device__ static __forceinline__ void to_record(Record &r, const Partial &p) { r.key = p.key == kKeyMax ? kNanKey : (u64)p.key; r.idx = p.idx; }
};

// Block-wide fixed-order combine; the result is valid in thread 0.
template <class Op>
__device__ __forceinline__ typename Op::Partial block_combine(typename Op::Partial p)
{
    __shared__ typename Op::Partial smem[32];
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) p = Op::merge(p, Op::shfl(p, m));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = (blockDim.x + 31) >> 5;
    __syncthreads();                       // smem may still be in use by a previous combine
    if (lane == 0) smem[warp] = p;
    __syncthreads();
    if (warp == 0) {
        p = lane < nwarps ? smem[lane] : Op::empty();
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) p = Op::merge(p, Op::shfl(p, m));
    }
    return p;
}

// What the last block does with the folded record.
//   mode 0: finish on this GPU and store the scalar to `out` (T, or int64 for iamax);
//   mode 1: store the record to `rec_out`; an NCCL all-gather and finish_records_kernel end the call;
//   mode 2: fused exchange -- push the record straight into every peer GPU's mailbox over NVLink (peer
//           memory mapped with CUDA IPC), wait for the peers' records in this GPU's mailbox, fold them in
//           rank order and finish, all inside this kernel.
struct FinishArgs {
    int kind;             // 0 sum, 1 nrm2, 2 iamax, 3 dsdot
    int mode;
    int is_f64;           // element type of the vectors (and of the scalar result)
    const void *x0;       // logical element 0 if this rank owns it (iamax NaN rule, nrm2 n==1), else null
    i64 n_global;
    double sb;            // sdsdot's addend
    void *out;
    Record *rec_out;
    // mode 2
    u64 *const *peers;    // device array: peers[r] = rank r's mailbox (peers[rank] is local)
    int nranks, rank;
    unsigned seq;         // number of this exchange, identical on every rank, never 0
    int *error_flag;      // host-mapped: set to 1 if a peer's record never arrived
    u64 timeout_ns;       // how long the last block waits for the peers' records (lbk_set_exchange_timeout)
    // synchronous calls: `out` is host-mapped and the host polls done_flag (next to it) for done_seq instead
    // of synchronising the stream -- the scalar is back 5-7 us sooner (profiles/r01_call_latency.csv)
    unsigned *done_flag;
    unsigned done_seq;
};

// Publishes a host-mapped result: the value first, a system-scope fence, then the tag the host spins on.
__device__ __forceinline__ void signal_done(unsigned *done_flag, unsigned done_seq)
{
    if (done_flag != nullptr) {
        __threadfence_system();
        *reinterpret_cast<volatile unsigned *>(done_flag) = done_seq;
    }
}

// Mailbox: two parities x kMaxRanks senders x kMailboxSlot words.  Each 64-bit word carries 32 bits of
// record and the 32-bit exchange number, so one store publishes data and "ready" together (no fence, no
// second round trip); a word is valid once its tag equals the current exchange number.  Exchange k+2 can
// only reuse exchange k's slots after every rank has consumed them (a rank sends k+1 only after finishing k).
constexpr int kMaxRanks = 64;
constexpr int kMailboxSlot = 16;                               // words reserved per sender (>= kRecordWords)
constexpr int kMailboxWords = 2 * kMaxRanks * kMailboxSlot;
constexpr u64 kExchangeTimeoutNs = 5ull * 1000ull * 1000ull * 1000ull;   // default; lbk_set_exchange_timeout
static_assert(kRecordWords <= kMailboxSlot, "a record fits its mailbox slot");

__device__ __forceinline__ u64 global_timer_ns()
{
    u64 t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// scale * sqrt(sumsq) of Blue's combine, in double for both element types (SB, SS: the type's scalings)
__host__ __device__ __forceinline__ double nrm2_finish(double abig, double amed, double asml, double SB, double SS)
{
    double scl, sumsq;
    if (abig > 0.0) {
        if (amed > 0.0 || amed != amed) abig += (amed * SB) * SB;
        scl = 1.0 / SB; sumsq = abig;
    } else if (asml > 0.0) {
        if (amed > 0.0 || amed != amed) {
            const double m = sqrt(amed), s = sqrt(asml) / SS;
            const double ymin = m < s ? m : s, ymax = m < s ? s : m;
            const double q = ymin / ymax;
            scl = 1.0; sumsq = (ymax * ymax) * (1.0 + q * q);
            if (amed != amed) sumsq = amed;
        } else { scl = 1.0 / SS; sumsq = asml; }
    } else { scl = 1.0; sumsq = amed; }
    return scl * sqrt(sumsq);
}

__device__ __forceinline__ void store_result(const FinishArgs &fa, double v)
{
    if (fa.is_f64) *reinterpret_cast<double *>(fa.out) = v;
    else *reinterpret_cast<float *>(fa.out) = (float)v;
}

__device__ __forceinline__ void finish_value(const Record &r, const FinishArgs &fa)
{
    if (fa.kind == 2) { *reinterpret_cast<i64 *>(fa.out) = r.idx; return; }
    if (fa.kind == 1) {
        if (fa.n_global == 1) {           // |x[0]| exactly, Single.hs:182
            if (fa.is_f64) *reinterpret_cast<double *>(fa.out) = from_abs_bits<double>(r.key);
            else *reinterpret_cast<float *>(fa.out) = from_abs_bits<float>(r.key);
            return;
        }
        store_result(fa, fa.is_f64 ? nrm2_finish(r.f[0], r.f[1], r.f[2], (double)Blue<double>::sbig, (double)Blue<double>::ssml)
                                   : nrm2_finish(r.f[0], r.f[1], r.f[2], (double)Blue<float>::sbig, (double)Blue<float>::ssml));
        return;
    }
    store_result(fa, fa.kind == 3 ? fa.sb + r.f[0] : r.f[0]);
}
__device__ __forceinline__ void finish_record(const Record &r, const FinishArgs &fa)
{
    finish_value(r, fa);
    signal_done(fa.done_flag, fa.done_seq);
}

// Cross-GPU fold: every rank holds all ranks' records (rank order) and folds them identically.
// iamax: larger key, then lower index, idx < 0 = no candidate; everything else: the sums add, and the
// key (nrm2: bits of |x[0]|) is rank 0's.
__host__ __device__ __forceinline__ Record fold_by_kind(const Record *recs, int nranks, int kind)
{
    Record acc = recs[0];
    for (int r = 1; r < nranks; ++r) {
        const Record &b = recs[r];
        if (kind == 2) {
            if (b.idx >= 0 && (acc.idx < 0 || b.key > acc.key || (b.key == acc.key && b.idx < acc.idx))) acc = b;
        } else {
            acc.f[0] += b.f[0]; acc.f[1] += b.f[1]; acc.f[2] += b.f[2];
        }
    }
    return acc;
}

// Second pass and finish.  `p` is this thread's partial; x0_bits (block 0, thread 0 only) the bit pattern
// of |logical element 0|, read before this block released its PDL dependents.
template <class Op>
__device__ __forceinline__ void grid_finish(typename Op::Partial p, Record *partials, unsigned *ticket,
                                            const FinishArgs &fa, u64 x0_bits)
{
    __shared__ bool is_last;
    p = block_combine<Op>(p);
    pdl_wait();          // the scratch below may still be in use by the preceding reduction's tail
    typename Op::Partial acc = p;
    u64 x0_key = x0_bits;
    if (gridDim.x > 1) {                     // a one-block launch (small n) is its own last block: no second pass
        if (threadIdx.x == 0) {
            Op::store(&partials[blockIdx.x], p);
            if (blockIdx.x == 0 && fa.kind == 1) __stcg(&partials[0].key, x0_bits);
            // release this block's slot / acquire the others' with the ticket itself (no separate fence)
            unsigned t;
            asm volatile("atom.acq_rel.gpu.global.inc.u32 %0, [%1], %2;" : "=r"(t) : "l"(ticket), "r"(gridDim.x - 1) : "memory");
            is_last = (t == gridDim.x - 1);                        // the ticket wraps to 0: ready for the next launch
        }
        __syncthreads();
        if (!is_last) return;
        __threadfence();
        acc = Op::empty();
        for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x)
            acc = Op::merge(acc, Op::load(&partials[b]));          // other blocks' slots: read through L2
        acc = block_combine<Op>(acc);
        if (threadIdx.x == 0 && fa.kind == 1 && fa.x0 != nullptr) x0_key = __ldcg(&partials[0].key);
    }
    __shared__ Record mine;
    if (threadIdx.x == 0) {
        Record r = empty_record();
        Op::to_record(r, acc);
        if (fa.kind == 1 && fa.x0 != nullptr) r.key = x0_key;       // nrm2 n==1 -> |x[0]|
        if (fa.mode == 0) finish_record(r, fa);
        else if (fa.mode == 1) *fa.rec_out = r;
        else mine = r;
    }
    if (fa.mode != 2) return;
    // ---- fused exchange over peer memory ----
    __shared__ Record all[kMaxRanks];
    __shared__ int timed_out;
    if (threadIdx.x == 0) timed_out = 0;
    __syncthreads();
    const unsigned par = fa.seq & 1u;
    const unsigned *words = reinterpret_cast<const unsigned *>(&mine);
    for (int t = threadIdx.x; t < fa.nranks * kRecordWords; t += blockDim.x) {   // push: one 8-byte store per word per peer
        const int peer = t / kRecordWords, w = t % kRecordWords;
        u64 *dst = fa.peers[peer] + (par * kMaxRanks + fa.rank) * kMailboxSlot + w;
        const u64 v = ((u64)fa.seq << 32) | words[w];
        asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(dst), "l"(v) : "memory");
    }
    const u64 t_start = global_timer_ns();
    for (int t = threadIdx.x; t < fa.nranks * kRecordWords; t += blockDim.x) {   // pull: poll this GPU's mailbox
        const int src = t / kRecordWords, w = t % kRecordWords;
        const u64 *slot = fa.peers[fa.rank] + (par * kMaxRanks + src) * kMailboxSlot + w;
        u64 v;
        for (;;) {
            asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(slot) : "memory");
            if ((unsigned)(v >> 32) == fa.seq) break;
            if (global_timer_ns() - t_start > fa.timeout_ns) { timed_out = 1; v = 0; break; }
        }
        reinterpret_cast<unsigned *>(&all[src])[w] = (unsigned)v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (timed_out) *fa.error_flag = 1;
        finish_record(fold_by_kind(all, fa.nranks, fa.kind), fa);
    }
}

template <class Op, int VEC, bool REV = false>
__device__ __forceinline__ void pack_step(typename Op::Acc &acc, const Pack<typename Op::Scalar, VEC> &x,
                                          const Pack<typename Op::Scalar, VEC> &y, unsigned ord0)
{
    using T = typename Op::Scalar;
    if constexpr (Op::HAS_PACK_STEP) Op::template step_pack<VEC>(acc, x, y, ord0);
    else {
#pragma unroll
        for (int l = 0; l < VEC; ++l) Op::step(acc, x.v[l], Op::TWO ? y.v[REV ? VEC - 1 - l : l] : T(0), ord0 + l);
    }
}

// Rules that depend on logical element 0, applied by the one thread that owns it (block 0, thread 0).  (A PDL
// secondary that overwrites x holds its stores until this kernel has completed, so x[0] is still the input.)
template <class Op>
__device__ __forceinline__ u64 apply_x0(typename Op::Partial &p, const FinishArgs &fa)
{
    using T = typename Op::Scalar;
    u64 bits = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0 && fa.x0 != nullptr) {
        const T v = *reinterpret_cast<const T *>(fa.x0);
        bits = abs_bits(v);
        if constexpr (Op::IS_IAMAX) { if (v != v) p = Op::nan_at_zero(); }    // NaN at index 0 stays
    }
    return bits;
}

// reduce, unit stride.  idx_base = global logical index of element 0 of this shard (iamax).
template <class Op, int VEC, int UNROLL, int POLICY, bool REV = false, int XSH = 0>
__global__ void __launch_bounds__(XSH > 0 ? kShiftThreads : kMaxThreads, XSH > 0 ? shift_min_blocks(sizeof(typename Op::Acc) > sizeof(typename Op::Scalar)) : reduce_min_blocks(VEC * UNROLL * (int)sizeof(typename Op::Scalar)))
reduce_unit_kernel(const typename Op::Scalar *x, const typename Op::Scalar *y, i64 head, i64 n, i64 idx_base,
                   Record *partials, unsigned *ticket, FinishArgs fa, int late_trigger, int wait_first)
{
    using T = typename Op::Scalar;
    static_assert(XSH == 0 || (XSH > 0 && XSH < VEC && !REV && Op::TWO), "shifted loads of x: two operands, forward pairing");
    constexpr int YS = REV ? -1 : 1;     // REV: element i pairs x[i] with y[n-1-i] (see map_unit_kernel)
    const i64 npack = XSH > 0 ? body_packs<VEC, XSH>(n, head) : (n - head) / VEC;
    const bool edge_lane = (threadIdx.x & 31) == Shift<VEC, XSH>::EDGE_LANE;
    const i64 per_tile = (i64)UNROLL * blockDim.x;
    const i64 ntiles = (npack + per_tile - 1) / per_tile;
    const T *bx = x + head, *by = REV ? y + (n - head) - VEC : y + head;
    // Dependents are released at once, unless the host expects the next kernel to collide with this one's
    // operands or with a kernel still running before it (lbk_lib.cu, TriggerPredictor): then each block releases
    // them after its last load AND after its predecessors have completed, so that the next kernel may read and
    // write anything but this kernel's result as soon as it runs, instead of being launched the ordinary way.
    if (wait_first) pdl_wait();      // reads what a predecessor writes: an ordinary launch in effect, minus its latency
    if (!late_trigger) pdl_launch_dependents();

    typename Op::Acc acc = Op::zero();
    unsigned it = 0;
    for (i64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const i64 p0 = tile * per_tile + threadIdx.x;
        Pack<T, VEC> xv[UNROLL], yv[UNROLL];
        bool whole;                                          // whole tile in range for this thread (shifted loads: for its warp)
        if constexpr (XSH > 0) whole = (p0 | 31) + (i64)(UNROLL - 1) * blockDim.x < npack;
        else whole = p0 + (i64)(UNROLL - 1) * blockDim.x < npack;
        if (whole) {
            if constexpr (XSH > 0) {
                T ex[UNROLL][Shift<VEC, XSH>::E];
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) {
                    const i64 e = (p0 + (i64)u * blockDim.x) * VEC;
                    shift_load<T, VEC, XSH, POLICY>(xv[u], ex[u], bx + e, edge_lane);          // see shift_pack
                    load_pack<VEC, POLICY, true>(yv[u], by + e);
                }
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) { const Pack<T, VEC> a = xv[u]; shift_pack<T, VEC, XSH>(xv[u], a, ex[u], edge_lane); }
            } else {
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) {
                    const i64 e = (p0 + (i64)u * blockDim.x) * VEC;
                    load_pack<VEC, POLICY, true>(xv[u], bx + e);
                    if constexpr (Op::TWO) load_pack<VEC, POLICY, true>(yv[u], by + YS * e);
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) pack_step<Op, VEC, REV>(acc, xv[u], yv[u], it * (UNROLL * VEC) + u * VEC);
        } else {
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const i64 p = p0 + (i64)u * blockDim.x;
                if (p < npack) {
                    const i64 e = p * VEC;
                    if constexpr (XSH > 0) {
#pragma unroll
                        for (int l = 0; l < VEC; ++l) xv[u].v[l] = bx[e + l];
                    } else load_pack<VEC, POLICY, true>(xv[u], bx + e);
                    if constexpr (Op::TWO) load_pack<VEC, POLICY, true>(yv[u], by + YS * e);
                    pack_step<Op, VEC, REV>(acc, xv[u], yv[u], it * (UNROLL * VEC) + u * VEC);
                }
            }
        }
    }
    typename Op::Partial p = Op::empty();
    if (blockIdx.x == 0) {      // scalar head and tail
        const i64 tail0 = head + npack * VEC;
        const i64 nscalar = head + (n - tail0);
        for (i64 t = threadIdx.x; t < nscalar; t += blockDim.x) {
            const i64 e = t < head ? t : tail0 + (t - head);
            Op::single(p, x[e], Op::TWO ? y[REV ? n - 1 - e : e] : T(0), idx_base + e);
        }
    }
    const u64 x0_bits = apply_x0<Op>(p, fa);
    if (late_trigger) { pdl_wait(); pdl_launch_dependents(); }      // every load of this block has been consumed
    if constexpr (Op::IS_IAMAX) {
        if (acc.val >= T(0)) {    // saw >= 1 non-NaN element; ordinal -> global index
            const unsigned per = UNROLL * VEC;
            const unsigned iter = acc.ord / per, rem = acc.ord % per, u = rem / VEC, l = rem % VEC;
            const i64 tile = (i64)blockIdx.x + (i64)iter * gridDim.x;
            p = Op::merge(Op::candidate(acc.val, idx_base + head + ((tile * UNROLL + u) * blockDim.x + threadIdx.x) * VEC + l), p);
        }
    } else {
        Op::add_acc(p, acc);
    }
    grid_finish<Op>(p, partials, ticket, fa, x0_bits);
}

// reduce, any increments: logical element i at x[kx0 + i*incx] (, y[ky0 + i*incy]).
template <class Op>
__global__ void __launch_bounds__(256)
reduce_strided_kernel(const typename Op::Scalar *x, i64 incx, i64 kx0, const typename Op::Scalar *y, i64 incy, i64 ky0, i64 n,
                      i64 idx_base, Record *partials, unsigned *ticket, FinishArgs fa, int late_trigger, int wait_first)
{
    using T = typename Op::Scalar;
    if (wait_first) pdl_wait();                      // as in reduce_unit_kernel
    if (!late_trigger) pdl_launch_dependents();
    // A grid-stride loop in groups of kStridedUnroll elements: all loads of a group are issued before the first is used.
    // The loads are volatile asm (load_pack), which keeps them in that order: written as plain C++ loads the compiler
    // hoists them for the sums but serialises them for iamax -- LDG, FSETP, FSEL, LDG, ... on one register, one load in
    // flight per thread, 56% of the DRAM rate (profiles/r02_ncu_all_2p28.md, profiles/r02_strided_reductions.md).  The loop keeps the
    // routine's register accumulator (iamax: running maximum + ordinal of its first occurrence, one compare per element);
    // the ordinal becomes an index once, after the loop.
    constexpr int U = kStridedUnroll;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    const i64 first = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    typename Op::Acc acc = Op::zero();
    unsigned ord = 0;
    i64 i = first;
    for (; i + (U - 1) * stride < n; i += U * stride, ord += U) {
        Pack<T, 1> xv[U], yv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            load_pack<1, 0, true>(xv[u], x + kx0 + (i + u * stride) * incx);
            if constexpr (Op::TWO) load_pack<1, 0, true>(yv[u], y + ky0 + (i + u * stride) * incy);
            else yv[u].v[0] = T(0);
        }
        // the group as one pack: iamax and nrm2 first reduce the whole group (a max / min tree over all U values, which
        // also keeps every load ahead of the first use) and only look at single elements when the group matters
        Pack<T, U> gx, gy;
#pragma unroll
        for (int u = 0; u < U; ++u) { gx.v[u] = xv[u].v[0]; gy.v[u] = yv[u].v[0]; }
        pack_step<Op, U>(acc, gx, gy, ord);
    }
    for (; i < n; i += stride, ++ord)
        Op::step(acc, x[kx0 + i * incx], Op::TWO ? y[ky0 + i * incy] : T(0), ord);
    typename Op::Partial p = Op::empty();
    const u64 x0_bits = apply_x0<Op>(p, fa);
    if (late_trigger) { pdl_wait(); pdl_launch_dependents(); }      // every load of this block has been consumed
    if constexpr (Op::IS_IAMAX) {
        if (acc.val >= T(0)) p = Op::merge(Op::candidate(acc.val, idx_base + first + (i64)acc.ord * stride), p);
    } else {
        Op::add_acc(p, acc);
    }
    grid_finish<Op>(p, partials, ticket, fa, x0_bits);
}

static __global__ void finish_records_kernel(const Record *recs, int nranks, FinishArgs fa)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    finish_record(fold_by_kind(recs, nranks, fa.kind), fa);
}

// The single-process multi-device context's stream-ordered exchange (no kernel waits for another): every rank's
// reduction kernel leaves its record in its own ring slot (mode 1), the host makes every rank's stream wait for
// the other ranks' "record written" events, and this kernel then reads all ranks' slots -- peer memory, or mapped
// host memory between devices that are not peers -- and folds them in rank order.
static __global__ void finish_gather_kernel(const Record *const *srcs, int slot, int nranks, FinishArgs fa)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    Record all[kMaxRanks];
    for (int r = 0; r < nranks; ++r) {
        const volatile u64 *src = reinterpret_cast<const volatile u64 *>(srcs[r] + slot);
        u64 *dst = reinterpret_cast<u64 *>(&all[r]);
#pragma unroll
        for (int w = 0; w < (int)(sizeof(Record) / 8); ++w) dst[w] = src[w];
    }
    finish_record(fold_by_kind(all, nranks, fa.kind), fa);
}

// ------------------------------------------------------------------------------------------------
// small utilities
// ------------------------------------------------------------------------------------------------
template <class T> __global__ void store_scalar_kernel(T *out, T v, unsigned *done_flag, unsigned done_seq)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) { *out = v; signal_done(done_flag, done_seq); }
}

__host__ __device__ __forceinline__ u64 splitmix64(u64 z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// v[i] = lo + span * u,  u = top 24 bits of splitmix64(seed * 2^40 + global_i) scaled by 2^-24, in T arithmetic
template <class T>
__global__ void __launch_bounds__(256)
fill_hash_kernel(T *v, i64 n, i64 idx_base, u64 seed, T lo, T span)
{
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const u64 h = splitmix64((seed << 40) + (u64)(idx_base + i));
        const T u = (T)(h >> 40) * (T)5.9604644775390625e-8;   // exact: 24 bits * 2^-24
        v[i] = add_rn(lo, mul_rn(span, u));
    }
}
template <class T>
__global__ void __launch_bounds__(256) fill_kernel(T *v, i64 n, T value)
{
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) v[i] = value;
}

}  // namespace lbk
