// lbk_tables.hpp -- the launch-shape variants of the unit-stride kernels and the tables that map (routine, variant) to a kernel.
//
// The tables instantiate every (functor | op) x variant x element type -- some 700 kernels, most of them there for the tuning
// sweeps (tools/tune.py) -- so they are compiled in translation units of their own, in parallel (Makefile): lbk_tables_map_f32.cu,
// lbk_tables_map_f64.cu, lbk_tables_reduce_f32.cu, lbk_tables_reduce_f64.cu.  lbk_lib.cu only sees the declarations.
#pragma once

#include "lbk_kernels.cuh"

// X(id, elements per pack at Float (halved at Double), packs in flight per operand per thread, access policy)
#define LBK_VARIANTS(X) \
    X(0, 4, 4, 1) X(1, 4, 2, 1) X(2, 4, 8, 1) X(3, 8, 2, 1) X(4, 8, 4, 1) X(5, 4, 4, 0) \
    X(6, 8, 1, 1) X(7, 4, 1, 1) X(8, 1, 4, 1) X(9, 8, 2, 0) X(10, 8, 8, 1) \
    X(11, 8, 4, 2) X(12, 8, 4, 3) X(13, 8, 4, 4) X(14, 8, 4, 5) \
    X(15, 8, 4, 6) X(16, 8, 2, 6) X(17, 8, 1, 6) X(18, 8, 8, 6) X(19, 4, 4, 6) X(20, 4, 2, 6) X(21, 4, 8, 6) X(22, 4, 1, 6) \
    X(23, 1, 8, 6) X(24, 1, 16, 6) X(25, 1, 16, 1) X(26, 1, 8, 1)
constexpr int kNumVariants = 27;

template <class T> constexpr int vec_t(int vec_f32) { return sizeof(T) == 4 ? vec_f32 : (vec_f32 > 1 ? vec_f32 / 2 : 1); }

// the kernel of a unit-stride elementwise routine / reduction at a variant (nullptr: no such variant)
template <class F> const void *map_kernel_ptr(int variant);
template <class Op> const void *reduce_kernel_ptr(int variant);

#define LBK_MAP_TABLE_DEFINITION                                                                                         \
    template <class F> const void *map_kernel_ptr(int variant)                                                           \
    {                                                                                                                    \
        using T = typename F::Scalar;                                                                                    \
        switch (variant) {                                                                                               \
            LBK_VARIANTS(LBK_MAP_CASE)                                                                                    \
        }                                                                                                                \
        return nullptr;                                                                                                  \
    }
#define LBK_MAP_CASE(ID, VEC, UNROLL, POLICY) case ID: return (const void *)lbk::map_unit_kernel<F, vec_t<T>(VEC), UNROLL, POLICY>;
#define LBK_MAP_TABLES(T)                                                                                                \
    template const void *map_kernel_ptr<lbk::AxpyF<T>>(int); template const void *map_kernel_ptr<lbk::ScalF<T>>(int);    \
    template const void *map_kernel_ptr<lbk::CopyF<T>>(int); template const void *map_kernel_ptr<lbk::FillF<T>>(int);    \
    template const void *map_kernel_ptr<lbk::SwapF<T>>(int); template const void *map_kernel_ptr<lbk::RotF<T>>(int);     \
    template const void *map_kernel_ptr<lbk::RotmNeg1F<T>>(int); template const void *map_kernel_ptr<lbk::Rotm0F<T>>(int); \
    template const void *map_kernel_ptr<lbk::Rotm1F<T>>(int);

#define LBK_REDUCE_TABLE_DEFINITION                                                                                      \
    template <class Op> const void *reduce_kernel_ptr(int variant)                                                       \
    {                                                                                                                    \
        using T = typename Op::Scalar;                                                                                   \
        switch (variant) {                                                                                               \
            LBK_VARIANTS(LBK_REDUCE_CASE)                                                                                 \
        }                                                                                                                \
        return nullptr;                                                                                                  \
    }
// policy 6 (the load fence) is a map-only policy: the reductions of those variants are policy 1
#define LBK_REDUCE_CASE(ID, VEC, UNROLL, POLICY) case ID: return (const void *)lbk::reduce_unit_kernel<Op, vec_t<T>(VEC), UNROLL, (POLICY == 6 ? 1 : POLICY)>;
#define LBK_REDUCE_TABLES(T)                                                                                             \
    template const void *reduce_kernel_ptr<lbk::DotOp<T>>(int); template const void *reduce_kernel_ptr<lbk::AsumOp<T>>(int); \
    template const void *reduce_kernel_ptr<lbk::Nrm2Op<T>>(int); template const void *reduce_kernel_ptr<lbk::IamaxOp<T>>(int);
