/*
 * lbk.h -- C ABI of the B200 Level-1 kernel library ("lambda-blas kernels").
 *
 * This is the drop-in boundary for ONE path of dlewissandy/lambda-blas: the
 * single-precision Level-1 routines of Numerical.BLAS.Single
 * (src/Numerical/BLAS/Single.hs:41-61).  The reference has no plugin interface of
 * its own; the only FFI it contains is the test oracle binding of
 * test/Foreign.hs:32-81 (`foreign import ccall "sdot_" ...`) with the marshalling
 * helpers of test/Foreign.hs:363-396.  Each entry point below names the reference
 * function it replaces and the Fortran binding whose role it takes.  The Haskell
 * module a maintainer would add on top of it is hs/Numerical/BLAS/Device.hs and is
 * described in INTEGRATION.md.
 *
 * Conventions
 *   - plain C linkage, opaque handles, no torch / C++ types in any signature;
 *   - every function returns an int status, LBK_OK (0) on success; results come
 *     back through out-pointers; scalars are passed BY VALUE (unlike the Fortran
 *     by-reference convention of test/Foreign.hs);
 *   - all lengths, counts and increments are int64_t (Haskell `Int` is 64-bit);
 *   - degenerate arguments follow the reference, which "does not generally guard
 *     against invalid values" (Single.hs:32-39): they produce the reference's
 *     sentinel results (0, -1, unchanged vector) with status LBK_OK.  Non-zero
 *     statuses are reserved for CUDA/NCCL failures, null handles, spans that fall
 *     outside a vector (the reference would read out of bounds) and layouts the
 *     device path does not support (see LBK_ERR_UNSUPPORTED);
 *   - vector routines are IN PLACE on device vectors and only ENQUEUE work on the
 *     context's stream; routines returning a scalar return once that scalar has
 *     arrived on the host (the kernel publishes it in mapped memory and the host
 *     polls; earlier work of the stream has completed by then, later calls may
 *     be issued at once).  The reference's PURE results (new vectors, inputs untouched,
 *     Util.hs:134-142) are obtained by the host layer with lbk_vec_clone + the
 *     in-place routine, or in one pass with the *_oop ("out of place") variants;
 *   - one context per host thread, or calls on a context serialised by the caller.
 *     There is no CPU fallback: without a CUDA device lbk_init fails;
 *   - the library leaves the process as it found it: every call runs with the context's device current and
 *     restores the caller's current device before it returns, and vectors come from the context's OWN
 *     stream-ordered memory pool (the device's default pool and its release threshold are not touched);
 *   - handles keep their context alive: lbk_shutdown on a context that still has vectors or events only closes
 *     it (every further call fails with LBK_ERR_ARG), and the last lbk_vec_free / lbk_event_destroy -- e.g. a
 *     finaliser run by a garbage collector after the scope that closed the context -- releases it.
 */
#ifndef LBK_H
#define LBK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LBK_VERSION 200 /* 0.2.0 */

enum lbk_status {
    LBK_OK = 0,
    LBK_ERR_CUDA = 1,        /* a CUDA runtime call failed; see lbk_last_error */
    LBK_ERR_NCCL = 2,        /* an NCCL call failed */
    LBK_ERR_ARG = 3,         /* null handle / pointer, negative length, handle of another context */
    LBK_ERR_RANGE = 4,       /* 1+(n-1)|inc| exceeds the vector (reference: undefined behaviour) */
    LBK_ERR_ALIAS = 5,       /* the two operands overlap in a way an in-place parallel update cannot honour */
    LBK_ERR_UNSUPPORTED = 6, /* e.g. a zero, opposite-signed or unequal increment on sharded vectors (equal negative ones are fine) */
    LBK_ERR_NOMEM = 7
};

typedef struct lbk_ctx lbk_ctx;     /* device + stream + scratch (+ communicator) */
typedef struct lbk_vec lbk_vec;     /* device-resident fp32 vector, or one rank's shard of one */
typedef struct lbk_event lbk_event; /* CUDA event on the context's stream (timing) */

/* ---- context ------------------------------------------------------------------------------ */
int lbk_version(void);
int lbk_device_count(int *count);
int lbk_init(int device, lbk_ctx **ctx);
/* ONE host thread, many GPUs (the reference is a single-threaded library: lambda-blas.cabal:36-48, call surface
 * Single.hs:84-85).  Rank r of the context is device devs[r].  lbk_vec_alloc_sharded then makes block shards over
 * the ranks and EVERY routine of this header fans out over per-device streams -- from one launcher thread per
 * device (lbk_set_multi_threads, the default) or one after the other from the calling thread; lbk_vec_upload /
 * _download scatter / gather ONE host buffer, each rank moving its block over its own PCIe link.  Elementwise
 * routines need no communication; a reduction ends inside its kernel: the last block of every rank writes its
 * 48-byte record into every other rank's mailbox through plain peer pointers (cudaDeviceEnablePeerAccess over
 * NVLink; mapped host memory when the devices are not peers) and folds what it received in rank order, so all
 * ranks hold the same bits and the caller reads rank 0's.  lbk_vec_alloc / lbk_vec_wrap make UNSHARDED vectors:
 * whole on devs[0], with the full increment semantics of a one-device context; the vectors of one call are all
 * sharded or all unsharded.  The same device may appear more than once in devs[]: ranks then share a GPU, each with
 * its own stream -- the sharded path on a one-GPU box.  Kernels of ONE GPU must not wait for one another (nothing
 * guarantees they run at the same time), so such a context ends its reductions with a stream-ordered exchange
 * instead of the in-kernel one: every rank's kernel leaves its record in its ring slot, every stream waits for the
 * other ranks' "record written" events, and a finish kernel per rank folds the records in rank order -- the same
 * fold, the same bits.  lbk_set_fused_exchange(ctx, 0 / 1) selects that exchange / the in-kernel one on any
 * multi-device context (rank contexts can be driven by hand only with the in-kernel exchange). */
int lbk_init_multi(int ndev, const int *devs, lbk_ctx **ctx);
/* The rank context of rank r (borrowed; it goes with the multi-device context): a one-device context with
 * nranks = ndev whose sharded vectors are that rank's blocks -- for callers that drive the ranks themselves. */
int lbk_multi_rank(lbk_ctx *ctx, int rank, lbk_ctx **rank_ctx);
/* The scalar rank `rank` itself holds for the most recent SYNCHRONOUS reduction (8 bytes: the float / double in the
 * low bytes, or the int64 index).  All ranks fold the same records in the same order; this lets a test see it. */
int lbk_multi_result(lbk_ctx *ctx, int rank, void *out8);
/* enabled = 1: one launcher thread per device (they spin for 200 us after a call, then sleep); 0: the calling
 * thread launches rank after rank (each launch costs 2-3 us, so the last of 8 GPUs starts ~20 us late). */
int lbk_set_multi_threads(lbk_ctx *ctx, int enabled);
int lbk_shutdown(lbk_ctx *ctx);
int lbk_sync(lbk_ctx *ctx);                 /* wait for everything enqueued on the context */
const char *lbk_last_error(lbk_ctx *ctx);   /* text of the last failure ("" if none); ctx may be NULL */
void *lbk_stream(lbk_ctx *ctx);             /* the cudaStream_t all work is enqueued on */
int lbk_sm_count(lbk_ctx *ctx, int *sms);
/* Select the launch shape of a routine ("sdot", "saxpy", ..., or "all"): kernel variant (vector width /
 * unroll / cache policy, see DESIGN.md "Tuning"), threads per block and resident blocks per SM.  A
 * value <= 0 (variant < 0) restores the default.  Elementwise results never depend on the shape;
 * reductions are deterministic for a given shape. */
int lbk_set_launch(lbk_ctx *ctx, const char *routine, int variant, int threads, int blocks_per_sm);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
int lbk_launch_count(lbk_ctx *ctx, int64_t *launches);
/* Programmatic dependent launch (on by default): consecutive unit-stride kernels overlap at their boundary --
 * the next kernel fills SMs while the previous one's last blocks (and a reduction's latency-bound tail)
 * drain.  The library tracks what the kernels still in flight read and write and launches the next one
 * holding its stores, or its loads, whenever the two depend on each other,
 * so results are unaffected.  lbk_pdl_count reports how many launches took that path. */
/* enabled: 0 off, 1 on (default).  For experiments (tools/, cpp/latency.cpp): 2 = a kernel that overwrites what its
 * predecessors use waits before its first LOAD instead of before its first store, 3 = dependent kernels of any kind
 * are launched the ordinary way (only independent ones overlap), 4 = only true dependencies are. */
int lbk_set_pdl(lbk_ctx *ctx, int enabled);
int lbk_pdl_count(lbk_ctx *ctx, int64_t *launches);
/* Unit-stride operands that start at DIFFERENT offsets inside a 16-byte pack (views, wrapped foreign pointers) cannot
 * share 128/256-bit accesses.  When the read-only x is the only odd one out (axpy, copy, out-of-place scal, dot, sdsdot)
 * the kernels still read it in aligned 16-byte packs and shift the lanes; every other case moves
 * one element per access.  lbk_shifted_count reports how many launches read x that way; the routines "mapshifted" /
 * "redshifted" of lbk_set_launch choose their shape (a one-element variant turns the path off). */
int lbk_shifted_count(lbk_ctx *ctx, int64_t *launches);
/* Walk direction of the unit-stride ELEMENTWISE kernels over vectors larger than the L2 cache.  What the 126 MB L2 still holds
 * of such a vector is the end the previous kernel streamed last, so by default (1, adaptive) an elementwise kernel starts its
 * walk there -- backwards after a kernel that walked forwards, as every reduction does -- and ~100 MB of its reads hit the
 * L2.  Results cannot depend on it.  0 = always forwards, 2 = always backwards (experiments). */
int lbk_set_traversal(lbk_ctx *ctx, int mode);

/* CUDA graphs: record a fixed sequence of calls once, replay it with one host call.  The reference's own benchmark
 * (test/Bench.hs:17-132) lives at n <= 30 000, where a GPU call is all launch latency; a captured chain runs at the GPU's
 * node-to-node latency instead (profiles/r02_graph_latency.md).  Between lbk_capture_begin and lbk_capture_end every call on the
 * context is RECORDED, not executed: the in-place and *_oop elementwise routines, the *_async reductions, lbk_vec_fill*,
 * lbk_vec_upload_async / lbk_vec_download_async (pinned buffers).  Calls that cannot be recorded return
 * LBK_ERR_UNSUPPORTED and leave the capture intact: anything that allocates (the pure host mirrors, lbk_vec_alloc*,
 * lbk_vec_clone), reductions that return their scalar to the host, synchronous transfers, lbk_sync, events, reductions
 * over sharded vectors.  lbk_vec_free is allowed (the buffer is released when the capture ends).  Single-device contexts only.
 * A replay repeats the recorded calls on the recorded addresses with the recorded scalars; it is enqueue-only. */
typedef struct lbk_graph lbk_graph;
int lbk_capture_begin(lbk_ctx *ctx);
int lbk_capture_end(lbk_ctx *ctx, lbk_graph **graph);
int lbk_graph_launch(lbk_graph *graph);
int lbk_graph_kernels(const lbk_graph *graph, int64_t *kernels);   /* kernels one replay runs (added to lbk_launch_count per launch) */
int lbk_graph_destroy(lbk_graph *graph);                          /* safe after lbk_shutdown, like lbk_vec_free */

/* ---- multi-GPU: one process per GPU, NCCL over NVLink -------------------------------------- */
/* Rank 0 makes an id and hands the 128 bytes to every rank (any transport; the host layer uses
 * torch.distributed); then every rank calls lbk_comm_init.  After that, vectors made with
 * lbk_vec_alloc_sharded are contiguous block shards and the reductions finish with one exchange. */
int lbk_comm_unique_id(char id[128]);
int lbk_comm_init(lbk_ctx *ctx, int nranks, int rank, const char id[128]);
int lbk_comm_info(lbk_ctx *ctx, int *nranks, int *rank);
/* How a sharded reduction ends.  fused = 1: the reduction kernel's last block writes this rank's 48-byte
 * record straight into every peer GPU's mailbox over NVLink (peer memory mapped with CUDA IPC at
 * lbk_comm_init), waits for the peers' records and folds them in rank order -- compute and exchange in ONE
 * kernel.  fused = 0: record -> ncclAllGather -> a one-thread finish kernel.  lbk_comm_init picks fused when
 * every rank could map every mailbox; lbk_set_fused_exchange overrides it (all ranks must agree). */
int lbk_comm_transport(lbk_ctx *ctx, int *fused);
int lbk_set_fused_exchange(lbk_ctx *ctx, int enabled);
/* How long a reduction kernel's last block waits for the peers' records before it gives up and the call (or the
 * next lbk_sync) returns LBK_ERR_NCCL: a rank that never makes the matching call must not hang the others.
 * Default 5000 ms, or LBK_EXCHANGE_TIMEOUT_MS from the environment. */
int lbk_set_exchange_timeout(lbk_ctx *ctx, int64_t milliseconds);
/* Block partition of [0, global_len) used by lbk_vec_alloc_sharded: rank r owns [lo, hi). */
int lbk_shard_range(int64_t global_len, int nranks, int rank, int64_t *lo, int64_t *hi);
/* Sharded vectors take positive increments: logical element i of a call (n, inc) sits at global position i*inc.
 * Which of them rank r holds: the first one, how many, and the offset of the first inside r's block.  (Two
 * operands must have equal lengths and equal increments, so that every pairing stays on one rank.  Two equal NEGATIVE
 * increments pair the same elements as their absolute values and are run that way; zero, opposite-signed and unequal
 * increments on shards are LBK_ERR_UNSUPPORTED.) */
int lbk_shard_span(int64_t global_len, int nranks, int rank, int64_t n, int64_t inc, int64_t *first, int64_t *count, int64_t *offset);

/* ---- device vectors (the representation added beside Numerical.BLAS.Types, Types.hs:3) -----
 * The reference's routines are polymorphic with Float and Double instances (sdot/ddot, ..., Single.hs:83-87);
 * a device vector is single precision unless made by an *_f64 constructor.  The s* routines take single-,
 * the d* routines double-precision vectors (LBK_ERR_ARG otherwise).  Lengths are in ELEMENTS. */
/* Owning vectors come from the device's stream-ordered memory pool (cudaMallocAsync on the context's stream);
 * lbk_vec_free releases them in stream order: work already enqueued may still use the vector, and neither
 * call synchronises. */
int lbk_vec_alloc(lbk_ctx *ctx, int64_t len, lbk_vec **v);
int lbk_vec_alloc_f64(lbk_ctx *ctx, int64_t len, lbk_vec **v);
int lbk_vec_alloc_sharded(lbk_ctx *ctx, int64_t global_len, lbk_vec **v);
int lbk_vec_alloc_sharded_f64(lbk_ctx *ctx, int64_t global_len, lbk_vec **v);
int lbk_vec_wrap(lbk_ctx *ctx, void *device_ptr, int64_t len, lbk_vec **v);     /* non-owning */
int lbk_vec_wrap_f64(lbk_ctx *ctx, void *device_ptr, int64_t len, lbk_vec **v);
int lbk_vec_view(const lbk_vec *v, int64_t offset, int64_t len, lbk_vec **view); /* non-owning */
/* An uninitialised vector with v's length, element type and shard layout: what a pure routine's result is made of. */
int lbk_vec_alloc_like(const lbk_vec *v, lbk_vec **out);
/* Rank r's block of a multi-device context's vector, as a (borrowed, non-owning) handle of that rank's context. */
int lbk_vec_part(const lbk_vec *v, int rank, lbk_vec **part);
int lbk_vec_free(lbk_vec *v);
int lbk_vec_len(const lbk_vec *v, int64_t *local_len, int64_t *global_len, int64_t *shard_offset);
int lbk_vec_is_sharded(const lbk_vec *v);                                       /* 1: a block shard / sharded over ranks */
int lbk_vec_elem_size(const lbk_vec *v);                                        /* 4 or 8 */
void *lbk_vec_ptr(const lbk_vec *v);
int lbk_vec_upload(lbk_vec *v, const void *host, int64_t len);        /* host -> device, synchronous */
int lbk_vec_download(const lbk_vec *v, void *host, int64_t len);      /* device -> host, synchronous */
int lbk_vec_upload_async(lbk_vec *v, const void *host, int64_t len);  /* enqueue only (pin the host buffer) */
int lbk_vec_download_async(const lbk_vec *v, void *host, int64_t len);
int lbk_vec_clone(const lbk_vec *v, lbk_vec **out);
/* Synthetic data, generated on the device from a counter hash of (seed, GLOBAL index), so shards of
 * any world size hold the same logical vector: v[i] = lo + (hi-lo) * u24(splitmix64(seed, i)), computed in the
 * vector's own precision. */
int lbk_vec_fill_hash(lbk_vec *v, uint64_t seed, double lo, double hi);
int lbk_vec_fill(lbk_vec *v, double value);
/* Pinned host memory for the transfers above: lbk_host_alloc makes some; lbk_host_register page-locks memory
 * the caller already owns (a Data.Vector.Storable buffer, Single.hs:67) until lbk_host_unregister, so that
 * uploads from it run at the rate of a pinned copy instead of through the driver's staging buffer. */
int lbk_host_alloc(int64_t nbytes, void **host);
int lbk_host_free(void *host);
int lbk_host_register(void *host, int64_t nbytes);
int lbk_host_unregister(void *host);

/* ---- timing on the context's stream ---------------------------------------------------------- */
int lbk_event_create(lbk_ctx *ctx, lbk_event **ev);
int lbk_event_record(lbk_event *ev);
int lbk_event_elapsed_ms(lbk_event *start, lbk_event *stop, float *ms); /* synchronises on `stop` */
/* Everything enqueued on `ctx` after this call waits for `ev` (recorded on ANOTHER context of the same device):
 * lets a second context's stream download results while the first keeps uploading and computing. */
int lbk_wait_event(lbk_ctx *ctx, lbk_event *ev);
int lbk_event_destroy(lbk_event *ev);

/* ---- reductions ------------------------------------------------------------------------------ */
/* dot / sdot / ddot, Single.hs:73-87 (Fortran role: "sdot_"/"ddot_", test/Foreign.hs:48,50): sum of
 * x[kx(i)]*y[ky(i)], kx(i) = first(n,incx) + i*incx, first(n,inc) = inc>0 ? 0 : (1-n)*inc (Single.hs:91-96);
 * inc == 0 replicates element 0 (Util.hs:86).  n <= 0 gives 0. */
int lbk_sdot(lbk_ctx *ctx, int64_t n, const lbk_vec *x, int64_t incx, const lbk_vec *y, int64_t incy, float *out);
int lbk_ddot(lbk_ctx *ctx, int64_t n, const lbk_vec *x, int64_t incx, const lbk_vec *y, int64_t incy, double *out);
/* sdsdot, Single.hs:231-262 ("sdsdot_", Foreign.hs:52): sb + sum in double, rounded once to float. */
int lbk_sdsdot(lbk_ctx *ctx, int64_t n, float sb, const lbk_vec *x, int64_t incx, const lbk_vec *y, int64_t incy, float *out);
int lbk_sdsdot_async(lbk_ctx *ctx, int64_t n, float sb, const lbk_vec *x, int64_t incx, const lbk_vec *y, int64_t incy, lbk_vec *out);
/* asum / sasum / dasum, Single.hs:155-168 ("sasum_"/"dasum_", Foreign.hs:36,38): 0 when incx < 1. */
int lbk_sasum(lbk_ctx *ctx, int64_t n, const lbk_vec *x, int64_t incx, float *out);
int lbk_dasum(lbk_ctx *ctx, int64_t n, const lbk_vec *x, int64_t incx, double *out);
/* nrm2 / snrm2 / dnrm2, Single.hs:174-204 ("snrm2_"/"dnrm2_", Foreign.hs:54,56): 0 when n<1 or incx<1,
 * |x[0]| when n==1; overflow/underflow-safe scaled accumulation. */
int lbk_snrm2(lbk_ctx *ctx, int64_t n, const lbk_vec *x, int64_t incx, float *out);
int lbk_dnrm2(lbk_ctx *ctx, int64_t n, const lbk_vec *x, int64_t incx, double *out);
/* iamax / isamax / idamax, Single.hs:211-225 ("isamax_"/"idamax_", Foreign.hs:32,34): 0-based LOGICAL index
 * of the first element of largest |x|; -1 when n<1 or incx<1; 0 when n==1.  Bit-exact, including ties and NaN
 * (a NaN wins only at logical index 0). */
int lbk_isamax(lbk_ctx *ctx, int64_t n, const lbk_vec *x, int64_t incx, int64_t *out);
int lbk_idamax(lbk_ctx *ctx, int64_t n, const lbk_vec *x, int64_t incx, int64_t *out);
/* Enqueue-only forms: the result is written to out_dev[0] (i?amax: the int64 into the first 8 bytes) on the
 * stream; nothing is synchronised.  `out_dev` is a device vector of the routine's precision. */
int lbk_sdot_async(lbk_ctx *ctx, int64_t n, const lbk_vec *x, int64_t incx, const lbk_vec *y, int64_t incy, lbk_vec *out_dev);
int lbk_sasum_async(lbk_ctx *ctx, int64_t n, const lbk_vec *x, int64_t incx, lbk_vec *out_dev);
int lbk_snrm2_async(lbk_ctx *ctx, int64_t n, const lbk_vec *x, int64_t incx, lbk_vec *out_dev);
int lbk_isamax_async(lbk_ctx *ctx, int64_t n, const lbk_vec *x, int64_t incx, lbk_vec *out_dev);
int lbk_ddot_async(lbk_ctx *ctx, int64_t n, const lbk_vec *x, int64_t incx, const lbk_vec *y, int64_t incy, lbk_vec *out_dev);
int lbk_dasum_async(lbk_ctx *ctx, int64_t n, const lbk_vec *x, int64_t incx, lbk_vec *out_dev);
int lbk_dnrm2_async(lbk_ctx *ctx, int64_t n, const lbk_vec *x, int64_t incx, lbk_vec *out_dev);
int lbk_idamax_async(lbk_ctx *ctx, int64_t n, const lbk_vec *x, int64_t incx, lbk_vec *out_dev);

/* ---- elementwise (in place, enqueue only) ------------------------------------------------------ */
/* axpy, Single.hs:430-444 ("saxpy_"/"daxpy_", Foreign.hs:40,42): y[ky(i)] += a*x[kx(i)], product rounded
 * before the add, NO a==0 shortcut. */
int lbk_saxpy(lbk_ctx *ctx, int64_t n, float a, const lbk_vec *x, int64_t incx, lbk_vec *y, int64_t incy);
int lbk_daxpy(lbk_ctx *ctx, int64_t n, double a, const lbk_vec *x, int64_t incx, lbk_vec *y, int64_t incy);
/* scal, Single.hs:101-116 ("sscal_"/"dscal_", Foreign.hs:74,76): x[i*incx] *= a for incx >= 1; incx <= 0
 * follows the reference's index stream (unchanged for n>1 and incx<0; x[0]*a once otherwise). */
int lbk_sscal(lbk_ctx *ctx, int64_t n, float a, lbk_vec *x, int64_t incx);
int lbk_dscal(lbk_ctx *ctx, int64_t n, double a, lbk_vec *x, int64_t incx);
/* copy, Single.hs:119-132 ("scopy_"/"dcopy_", Foreign.hs:44,46): y[ky(i)] = x[kx(i)]; incy==0 leaves the
 * LAST sampled x in y[0]. */
int lbk_scopy(lbk_ctx *ctx, int64_t n, const lbk_vec *x, int64_t incx, lbk_vec *y, int64_t incy);
int lbk_dcopy(lbk_ctx *ctx, int64_t n, const lbk_vec *x, int64_t incx, lbk_vec *y, int64_t incy);
/* swap, Single.hs:135-150 ("sswap_"/"dswap_", Foreign.hs:78,80). */
int lbk_sswap(lbk_ctx *ctx, int64_t n, lbk_vec *x, int64_t incx, lbk_vec *y, int64_t incy);
int lbk_dswap(lbk_ctx *ctx, int64_t n, lbk_vec *x, int64_t incx, lbk_vec *y, int64_t incy);
/* rot, Single.hs:408-424 ("srot_"/"drot_", Foreign.hs:58,60): x' = c*x + s*y, y' = c*y - s*x, both from
 * the old values, each product rounded separately. */
int lbk_srot(lbk_ctx *ctx, int64_t n, lbk_vec *x, int64_t incx, lbk_vec *y, int64_t incy, float c, float s);
int lbk_drot(lbk_ctx *ctx, int64_t n, lbk_vec *x, int64_t incx, lbk_vec *y, int64_t incy, double c, double s);
/* rotm, Single.hs:463-487 ("srotm_"/"drotm_", Foreign.hs:66,68).  param is the BLAS sparam array
 * [flag,h11,h21,h12,h22] (Foreign.hs:276-280); flag -2 is a no-op that launches nothing. */
int lbk_srotm(lbk_ctx *ctx, int64_t n, lbk_vec *x, int64_t incx, lbk_vec *y, int64_t incy, const float param[5]);
int lbk_drotm(lbk_ctx *ctx, int64_t n, lbk_vec *x, int64_t incx, lbk_vec *y, int64_t incy, const double param[5]);

/* ---- out-of-place forms: the reference's pure results in one pass ------------------------------- */
/* yout (same length as y) = updateElems result of Util.hs:134-142: y with the touched slots replaced,
 * everything else copied.  Outputs must not overlap the inputs. */
int lbk_saxpy_oop(lbk_ctx *ctx, int64_t n, float a, const lbk_vec *x, int64_t incx, const lbk_vec *y, int64_t incy, lbk_vec *yout);
int lbk_sscal_oop(lbk_ctx *ctx, int64_t n, float a, const lbk_vec *x, int64_t incx, lbk_vec *xout);
int lbk_scopy_oop(lbk_ctx *ctx, int64_t n, const lbk_vec *x, int64_t incx, const lbk_vec *y, int64_t incy, lbk_vec *yout);
int lbk_sswap_oop(lbk_ctx *ctx, int64_t n, const lbk_vec *x, int64_t incx, const lbk_vec *y, int64_t incy, lbk_vec *xout, lbk_vec *yout);
int lbk_srot_oop(lbk_ctx *ctx, int64_t n, const lbk_vec *x, int64_t incx, const lbk_vec *y, int64_t incy, float c, float s, lbk_vec *xout, lbk_vec *yout);
int lbk_srotm_oop(lbk_ctx *ctx, int64_t n, const lbk_vec *x, int64_t incx, const lbk_vec *y, int64_t incy, const float param[5], lbk_vec *xout, lbk_vec *yout);
int lbk_daxpy_oop(lbk_ctx *ctx, int64_t n, double a, const lbk_vec *x, int64_t incx, const lbk_vec *y, int64_t incy, lbk_vec *yout);
int lbk_dscal_oop(lbk_ctx *ctx, int64_t n, double a, const lbk_vec *x, int64_t incx, lbk_vec *xout);
int lbk_dcopy_oop(lbk_ctx *ctx, int64_t n, const lbk_vec *x, int64_t incx, const lbk_vec *y, int64_t incy, lbk_vec *yout);
int lbk_dswap_oop(lbk_ctx *ctx, int64_t n, const lbk_vec *x, int64_t incx, const lbk_vec *y, int64_t incy, lbk_vec *xout, lbk_vec *yout);
int lbk_drot_oop(lbk_ctx *ctx, int64_t n, const lbk_vec *x, int64_t incx, const lbk_vec *y, int64_t incy, double c, double s, lbk_vec *xout, lbk_vec *yout);
int lbk_drotm_oop(lbk_ctx *ctx, int64_t n, const lbk_vec *x, int64_t incx, const lbk_vec *y, int64_t incy, const double param[5], lbk_vec *xout, lbk_vec *yout);

/* ---- host scalars (O(1); stay on the CPU like the reference's) ---------------------------------- */
/* rotg, Single.hs:274-300 ("srotg_"/"drotg_", Foreign.hs:62,64): out = (r, z, c, s) (Types.hs:8). */
int lbk_srotg(float sa, float sb, float out[4]);
int lbk_drotg(double sa, double sb, double out[4]);
/* rotmg + checkscale, Single.hs:318-401 ("srotmg_"/"drotmg_", Foreign.hs:70,72).
 * out = [flag, d1, d2, x1, h11, h12, h21, h22], flag in {-2,-1,0,1} = FLAGNEG2/FLAGNEG1/FLAG0/FLAG1
 * (Types.hs:15-20); param (may be NULL) receives the BLAS sparam array for lbk_?rotm. */
int lbk_srotmg(float sd1, float sd2, float sx1, float sy1, float out[8], float param[5]);
int lbk_drotmg(double sd1, double sd2, double sx1, double sy1, double out[8], double param[5]);

/* ---- host-side statement of the cross-rank combine (what the finish kernel computes) ------------- */
/* One 48-byte record per rank, in rank order; used by the CPU tests of the multi-GPU path.  f[0..2] are
 * partial sums held in double (float partials are exact in them); key/idx the i?amax candidate. */
typedef struct lbk_record { double f[3]; uint64_t key; int64_t idx; int64_t pad; } lbk_record;
enum lbk_reduce_kind { LBK_RED_SUM = 0, LBK_RED_NRM2 = 1, LBK_RED_IAMAX = 2 };
int lbk_combine_records(int kind, int is_f64, const lbk_record *recs, int nranks, double *fout, int64_t *iout);

/* ---- the launcher threads of a multi-device context on their own (CPU tests; lambda_blas_b200/csrc/lbk_team.hpp) ----
 * `jobs` jobs on a team of `nranks`; job j makes rank r add (j+1)*(r+1) to its counter and report status 100+r when
 * j == fail_at; *sum = total of the counters, *first_status = what the failing job returned. */
int lbk_team_selftest(int nranks, int jobs, int fail_at, int pause_us, int64_t *sum, int *first_status);

/* ---- the launch planner on its own (CPU tests of the programmatic-dependent-launch rules) ------------------
 * Feeds one unit-stride kernel at a time -- what it reads and writes as (address, bytes) pairs, whether it is a
 * reduction, its routine id -- to the planner the library uses (lambda_blas_b200/csrc/lbk_pdl.hpp) and reports
 * how it would be launched: mode 0 = ordinary launch, 1 = overlaps its predecessors, waits before it exits,
 * 2 = overlaps, but waits before its first store, 3 = resident early, but waits before its first load; late = 1: a reduction that releases its dependents only after
 * its loads.  lbk_pdl_sim_interrupt = any other operation on the stream. */
int lbk_pdl_sim_create(void **sim);
int lbk_pdl_sim_destroy(void *sim);
int lbk_pdl_sim_enable(void *sim, int enabled);
int lbk_pdl_sim_interrupt(void *sim);
int lbk_pdl_sim_op(void *sim, int reduction, int rid, const int64_t *rd, int nrd, const int64_t *wr, int nwr, int *mode, int *late);

#ifdef __cplusplus
}
#endif
#endif /* LBK_H */
