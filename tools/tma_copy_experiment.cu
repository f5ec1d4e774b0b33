// Experiment: does a TMA bulk-copy pipeline (cp.async.bulk global->shared->global, mbarrier completion, one
// issuing thread per CTA) move a 1 GiB fp32 vector faster than the LDG.256/STG.256 kernel scopy uses?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tma_copy tools/tma_copy_experiment.cu && /tmp/tma_copy
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count)); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity)
{
    asm volatile("{\n.reg .pred p;\nWAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT;\nDONE:\n}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *b)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void bulk_store(void *dst, const void *src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}

template <int STAGES, int TILE>
__global__ void __launch_bounds__(32) bulk_copy_kernel(const char *src, char *dst, long long ntiles)
{
    extern __shared__ __align__(128) char smem[];
    __shared__ uint64_t full[STAGES];
    if (threadIdx.x != 0) return;
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
    const long long mine = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    auto tile_of = [&](long long k) { return (long long)blockIdx.x + k * gridDim.x; };
    long long issued = 0;
    for (; issued < mine && issued < STAGES - 1; ++issued) {           // prologue: STAGES-1 loads in flight
        const int s = issued % STAGES;
        mbar_expect(&full[s], TILE);
        bulk_load(smem + (size_t)s * TILE, src + tile_of(issued) * TILE, TILE, &full[s]);
    }
    for (long long k = 0; k < mine; ++k) {
        const int s = k % STAGES;
        mbar_wait(&full[s], (k / STAGES) & 1);
        bulk_store(dst + tile_of(k) * TILE, smem + (size_t)s * TILE, TILE);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (issued < mine) {                                              // refill the buffer of tile k-1
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); // store k-1 has finished reading smem
            const int r = issued % STAGES;
            mbar_expect(&full[r], TILE);
            bulk_load(smem + (size_t)r * TILE, src + tile_of(issued) * TILE, TILE, &full[r]);
            ++issued;
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// Non-persistent form: one CTA moves NT consecutive tiles (a contiguous NT*TILE chunk) and exits; the
// hardware block scheduler walks the vector front to back (the shape that won for the LDG/STG maps).
template <int NT, int TILE>
__global__ void __launch_bounds__(32) bulk_copy_once_kernel(const char *src, char *dst)
{
    extern __shared__ __align__(128) char smem[];
    __shared__ uint64_t full[NT];
    if (threadIdx.x != 0) return;
    for (int s = 0; s < NT; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const size_t base = (size_t)blockIdx.x * NT * TILE;
    for (int s = 0; s < NT; ++s) {
        mbar_expect(&full[s], TILE);
        bulk_load(smem + (size_t)s * TILE, src + base + (size_t)s * TILE, TILE, &full[s]);
    }
    for (int s = 0; s < NT; ++s) {
        mbar_wait(&full[s], 0);
        bulk_store(dst + base + (size_t)s * TILE, smem + (size_t)s * TILE, TILE);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int NT, int TILE>
int run_once(const char *x, char *y, size_t bytes)
{
    const size_t smem = (size_t)NT * TILE;
    CK(cudaFuncSetAttribute(bulk_copy_once_kernel<NT, TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (int)(bytes / ((size_t)NT * TILE));
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) bulk_copy_once_kernel<NT, TILE><<<grid, 32, smem>>>(x, y);
    CK(cudaGetLastError());
    cudaEventRecord(a);
    for (int i = 0; i < 10; ++i) bulk_copy_once_kernel<NT, TILE><<<grid, 32, smem>>>(x, y);
    cudaEventRecord(b);
    CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("one-shot CTAs: %d tiles x %6d B per CTA, grid %d : %.1f us  %.1f GB/s (read+write)\n", NT, TILE, grid, ms * 100, 2.0 * bytes / (ms / 10) / 1e6);
    return 0;
}

template <int STAGES, int TILE>
int run(const char *x, char *y, size_t bytes, int sms, int ctas_per_sm)
{
    const size_t smem = (size_t)STAGES * TILE;
    CK(cudaFuncSetAttribute(bulk_copy_kernel<STAGES, TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long ntiles = bytes / TILE;
    const int grid = sms * ctas_per_sm;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) bulk_copy_kernel<STAGES, TILE><<<grid, 32, smem>>>(x, y, ntiles);
    CK(cudaGetLastError());
    cudaEventRecord(a);
    for (int i = 0; i < 10; ++i) bulk_copy_kernel<STAGES, TILE><<<grid, 32, smem>>>(x, y, ntiles);
    cudaEventRecord(b);
    CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("stages %d tile %6d B ctas/sm %d : %.1f us  %.1f GB/s (read+write)\n", STAGES, TILE, ctas_per_sm, ms * 100, 2.0 * bytes / (ms / 10) / 1e6);
    return 0;
}

int main()
{
    const size_t bytes = (size_t)1 << 30;
    char *x, *y;
    CK(cudaMalloc(&x, bytes)); CK(cudaMalloc(&y, bytes));
    CK(cudaMemset(x, 1, bytes)); CK(cudaMemset(y, 0, bytes));
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) cudaMemcpyAsync(y, x, bytes, cudaMemcpyDeviceToDevice);
    cudaEventRecord(a);
    for (int i = 0; i < 10; ++i) cudaMemcpyAsync(y, x, bytes, cudaMemcpyDeviceToDevice);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("cudaMemcpyAsync D2D: %.1f us %.1f GB/s (read+write)\n", ms * 100, 2.0 * bytes / (ms / 10) / 1e6);
    run<4, 16384>(x, y, bytes, sms, 1); run<8, 16384>(x, y, bytes, sms, 1); run<12, 16384>(x, y, bytes, sms, 1);
    run<6, 32768>(x, y, bytes, sms, 1); run<4, 16384>(x, y, bytes, sms, 2); run<6, 16384>(x, y, bytes, sms, 2);
    run<8, 8192>(x, y, bytes, sms, 2); run<4, 8192>(x, y, bytes, sms, 4); run<3, 65536>(x, y, bytes, sms, 1);
    run<2, 32768>(x, y, bytes, sms, 3); run<13, 16384>(x, y, bytes, sms, 1);
    run_once<1, 8192>(x, y, bytes); run_once<1, 16384>(x, y, bytes); run_once<1, 32768>(x, y, bytes); run_once<1, 65536>(x, y, bytes);
    run_once<2, 8192>(x, y, bytes); run_once<2, 16384>(x, y, bytes); run_once<4, 8192>(x, y, bytes); run_once<4, 16384>(x, y, bytes);
    run_once<2, 32768>(x, y, bytes); run_once<4, 4096>(x, y, bytes); run_once<8, 4096>(x, y, bytes); run_once<1, 4096>(x, y, bytes);
    CK(cudaMemset(y, 0, bytes));
    run_once<2, 16384>(x, y, bytes);
    unsigned char h[4];
    CK(cudaMemcpy(h, y + bytes - 4, 4, cudaMemcpyDeviceToHost));
    printf("check: %d %d\n", h[0], h[3]);
    return 0;
}
