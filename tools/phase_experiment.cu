// Experiment (round 2, VERDICT "writers"): can a COARSER separation of a writer's load phase from its store phase
// than the library's per-thread load fence lift saxpy / scopy above ~7.0-7.1 TB/s?  saxpy (2 reads : 1 write) and
// scopy (1 : 1) at n = 2^28, one tile per block, 256-bit accesses, back-to-back launches, CUDA events.
//
//   MODE 0  per-thread load fence (the library's POLICY 6): a thread's stores wait for all of ITS loads
//   MODE 1  CTA phase separation: fence, __syncthreads(), then every thread of the block stores
//   MODE 2  cluster phase separation: fence, barrier.cluster over CLUSTER blocks, then all of them store
//   MODE 3  hybrid: LDG into registers, result to shared memory, ONE bulk store per warp-sized slab
//           (cp.async.bulk.global.shared::cta) -- the store stream leaves the LSU and reaches L2 in 4 KB pieces
//   MODE 4  like 3 with one bulk store per BLOCK (the whole tile, up to 32 KB)
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/phase_exp tools/phase_experiment.cu && build/phase_exp
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

struct alignas(32) P8 { float v[8]; };
__device__ __forceinline__ void ld8(P8 &p, const float *a)
{
    asm volatile("ld.global.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(p.v[0]), "=f"(p.v[1]), "=f"(p.v[2]), "=f"(p.v[3]), "=f"(p.v[4]), "=f"(p.v[5]), "=f"(p.v[6]), "=f"(p.v[7]) : "l"(a) : "memory");
}
__device__ __forceinline__ void st8(float *a, const P8 &p)
{
    asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(a),
                 "f"(p.v[0]), "f"(p.v[1]), "f"(p.v[2]), "f"(p.v[3]), "f"(p.v[4]), "f"(p.v[5]), "f"(p.v[6]), "f"(p.v[7]) : "memory");
}
__device__ __forceinline__ void cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void bulk_store(float *dst, const void *smem_src, unsigned bytes)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_src);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_wait()
{
    asm volatile("cp.async.bulk.commit_group;\n\tcp.async.bulk.wait_group.read 0;" ::: "memory");
}

// AXPY: y += a*x (reads x and y, writes y); otherwise copy y = x.
template <int U, int MODE, bool AXPY>
__global__ void __launch_bounds__(512) writer_kernel(const float *x, float *y, float a, unsigned fence_mask)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const long long base = ((long long)blockIdx.x * U * blockDim.x) * 8;
    P8 xv[U], yv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const long long e = base + ((long long)u * blockDim.x + threadIdx.x) * 8;
        ld8(xv[u], x + e);
        if (AXPY) ld8(yv[u], y + e);
    }
    unsigned dep = 0;
#pragma unroll
    for (int u = 0; u < U; ++u) { dep |= __float_as_uint(xv[u].v[7]); if (AXPY) dep |= __float_as_uint(yv[u].v[7]); }
    const long long held = (long long)(dep & fence_mask);      // 0 at run time: every load of this thread has landed
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int l = 0; l < 8; ++l) yv[u].v[l] = AXPY ? __fadd_rn(yv[u].v[l], __fmul_rn(a, xv[u].v[l])) : xv[u].v[l];
    if (MODE == 1) { if (held == 1) y[0] = 0.f; __syncthreads(); }
    if (MODE == 2) { if (held == 1) y[0] = 0.f; cluster_sync(); }
    if (MODE <= 2) {
#pragma unroll
        for (int u = 0; u < U; ++u) st8(y + base + ((long long)u * blockDim.x + threadIdx.x) * 8 + held, yv[u]);
        return;
    }
    // hybrid: the tile goes to shared memory in its global layout, then leaves as bulk stores
    P8 *tile = reinterpret_cast<P8 *>(smem);
#pragma unroll
    for (int u = 0; u < U; ++u) tile[u * blockDim.x + threadIdx.x + held] = yv[u];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (MODE == 3) {
        __syncwarp();
        if ((threadIdx.x & 31) == 0) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int p0 = u * blockDim.x + (threadIdx.x & ~31);          // this warp's 32 packs of slab u: 1 KB
                bulk_store(y + base + (long long)p0 * 8, tile + p0, 32 * 32);
            }
            bulk_commit_wait();
        }
    } else {
        __syncthreads();
        if (threadIdx.x == 0) {
            bulk_store(y + base, tile, U * blockDim.x * 32);
            bulk_commit_wait();
        }
    }
}

template <int U, int MODE, bool AXPY>
static float run(const char *name, float *x, float *y, long long n, int threads, int cluster)
{
    const long long per_block = (long long)U * threads * 8;
    const int grid = (int)(n / per_block);
    const size_t smem = MODE >= 3 ? (size_t)U * threads * 32 : 0;
    auto k = writer_kernel<U, MODE, AXPY>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(grid); lc.blockDim = dim3(threads); lc.dynamicSmemBytes = smem; lc.stream = 0;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = MODE == 2 ? 1 : 0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const float a = 1.000123f;
    const unsigned mask = 0;
    for (int i = 0; i < 3; ++i) cudaLaunchKernelEx(&lc, k, (const float *)x, y, a, mask);
    cudaEventRecord(e0);
    const int reps = 20;
    for (int i = 0; i < reps; ++i) cudaLaunchKernelEx(&lc, k, (const float *)x, y, a, mask);
    cudaEventRecord(e1);
    cudaError_t err = cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= reps;
    const double bytes = (AXPY ? 12.0 : 8.0) * (double)n;
    if (err != cudaSuccess || cudaGetLastError() != cudaSuccess) { printf("%s,%d,%d,%d,%d,FAILED %s\n", name, MODE, U, threads, cluster, cudaGetErrorString(err)); return 0; }
    printf("%s,%d,%d,%d,%d,%.1f,%.1f,%.1f\n", name, MODE, U, threads, cluster, ms * 1e3, bytes / ms / 1e6, bytes / ms / 1e6 / 80.0);
    fflush(stdout);
    return ms;
}

template <bool AXPY> static void sweep(const char *name, float *x, float *y, long long n)
{
    run<2, 0, AXPY>(name, x, y, n, 512, 1);
    run<1, 0, AXPY>(name, x, y, n, 512, 1);
    run<4, 0, AXPY>(name, x, y, n, 256, 1);
    run<2, 1, AXPY>(name, x, y, n, 512, 1);
    run<1, 1, AXPY>(name, x, y, n, 512, 1);
    run<4, 1, AXPY>(name, x, y, n, 256, 1);
    run<2, 1, AXPY>(name, x, y, n, 256, 1);
    for (int c : {2, 4, 8}) {
        run<2, 2, AXPY>(name, x, y, n, 512, c);
        run<1, 2, AXPY>(name, x, y, n, 512, c);
        run<2, 2, AXPY>(name, x, y, n, 256, c);
    }
    run<2, 3, AXPY>(name, x, y, n, 512, 1);
    run<1, 3, AXPY>(name, x, y, n, 512, 1);
    run<2, 3, AXPY>(name, x, y, n, 256, 1);
    run<2, 4, AXPY>(name, x, y, n, 512, 1);
    run<1, 4, AXPY>(name, x, y, n, 512, 1);
    run<2, 4, AXPY>(name, x, y, n, 256, 1);
    run<4, 4, AXPY>(name, x, y, n, 256, 1);
}

int main(int argc, char **argv)
{
    const long long n = 1ll << (argc > 1 ? atoi(argv[1]) : 28);
    float *x, *y;
    cudaMalloc(&x, n * 4); cudaMalloc(&y, n * 4);
    cudaMemset(x, 0, n * 4); cudaMemset(y, 0, n * 4);
    printf("routine,mode,unroll,threads,cluster,us,GBps,pct_of_8TBs\n");
    sweep<true>("saxpy", x, y, n);
    sweep<false>("scopy", x, y, n);
    return 0;
}
