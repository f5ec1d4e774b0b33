// lbk_lib.cu -- the C ABI of include/lbk.h over the sm_100a kernels of lbk_kernels.cuh.
//
// Host-side responsibilities (everything the reference's index helpers decide before a loop
// runs, src/Numerical/BLAS/Util.hs:39-142 and the guards of src/Numerical/BLAS/Single.hs):
// sentinel results, span checks, choice of the unit-stride or the general-stride kernel,
// alignment peeling for the 128/256-bit path, snapshots for zero output increments, launch
// shape, and -- with a communicator -- the one exchange that ends a sharded reduction.
#include "../../include/lbk.h"
#include "lbk_kernels.cuh"
#include "lbk_pdl.hpp"
#include "lbk_tables.hpp"
#include "lbk_team.hpp"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>   // types only: the library is bound at run time (see NcclApi)

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

using lbk::i64;
using lbk::Record;
using lbk::FinishArgs;

static_assert(sizeof(lbk_record) == sizeof(Record), "host and device records share a layout");

// ------------------------------------------------------------------------------------------------
// launch variants: (VEC floats per access = 4/16/32 bytes, UNROLL accesses in flight per operand, cache POLICY);
// for double vectors the same variant moves the same bytes (VEC/2 elements)
// ------------------------------------------------------------------------------------------------
// (the variant table LBK_VARIANTS and the per-variant kernel tables live in lbk_tables.hpp / lbk_tables_*.cu)
static const int kVariantVec[kNumVariants] = {4, 4, 4, 8, 8, 4, 8, 4, 1, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 4, 4, 4, 4, 1, 1, 1, 1};
constexpr int kScalarVariant = 8;   // VEC == 1: any alignment (the reversed pairing's any-alignment shape)
constexpr int kVec4Fallback = 0;    // used when a 256-bit variant meets 16-byte-only alignment

// launch shapes are per routine and shared by the Float and Double instances (they are in bytes per access)
enum Routine { R_DOT, R_SDSDOT, R_ASUM, R_NRM2, R_IAMAX, R_AXPY, R_SCAL, R_COPY, R_SWAP, R_ROT, R_ROTM, R_FILL, R_MAPSTRIDED, R_REDSTRIDED,
               R_MAPMISALIGNED, R_REDMISALIGNED, R_MAPSHIFTED, R_REDSHIFTED, R_COUNT };
static const char *kRoutineNames[R_COUNT] = {"dot", "sdsdot", "asum", "nrm2", "iamax", "axpy", "scal", "copy", "swap", "rot", "rotm", "fill", "mapstrided", "redstrided",
                                            "mapmisaligned", "redmisaligned", "mapshifted", "redshifted"};
struct LaunchCfg { int variant; int threads; int blocks_per_sm; bool pinned = false; };   // <= 0: default / from occupancy; pinned: set by lbk_set_launch, taken literally
// Defaults from the n = 2^28 sweeps on B200 (profiles/r01_tune_*.csv; tools/tune.py), reductions tuned
// with their tails overlapped (PDL).  blocks_per_sm 0 = one resident wave (occupancy x SMs); larger
// values oversubscribe the SMs and let the hardware block scheduler balance the load; kOneTilePerBlock
// makes the grid as large as the tile count, so blocks retire front to back through the vector -- the
// shape every writer prefers (with the load fence of POLICY 6 and 32-64 bytes per thread per operand).
// (saxpy: 16-byte packs x 4 are 1% faster back to back, 32-byte packs x 2 are 1% faster inside the bench step,
// where saxpy runs under sdot's tail and snrm2 under saxpy's -- profiles/README.md; the step decides.)
constexpr int kOneTilePerBlock = 1 << 20;
static const LaunchCfg kDefaultCfg[R_COUNT] = {
    /* dot  */ {4, 256, 16},   /* sdsdot */ {3, 512, 8},    /* asum */ {3, 256, 16},
    /* nrm2 */ {3, 256, 16},   /* iamax  */ {3, 256, 16},   /* axpy */ {16, 512, kOneTilePerBlock},
    /* scal */ {20, 512, kOneTilePerBlock},  /* copy */ {20, 512, kOneTilePerBlock},  /* swap */ {22, 512, kOneTilePerBlock},
    /* rot  */ {22, 256, kOneTilePerBlock},  /* rotm */ {22, 256, kOneTilePerBlock},  /* fill */ {4, 256, 256},
    /* the general-stride map kernel: variant 0/1 = without/with the load fence (no measurable effect
       there); threads / blocks_per_sm 0 = chosen from the increments in launch_map_strided */ {0, 0, 0},
    /* the general-stride reduce kernel: threads (<= 256) and blocks per SM of its grid-stride loop (0: from the increments) */ {0, 256, 0},
    /* unit stride, operands misaligned DIFFERENTLY (x and y start at different offsets inside a 16-byte pack): 4-byte accesses,
       fully coalesced (a warp still moves 128 contiguous bytes per instruction), many of them in flight per thread and as large
       a tile per block as the wide shapes have -- profiles/r02_misaligned.md */
    /* map */ {26, 256, kOneTilePerBlock}, /* reduce */ {26, 512, 16},
    /* the same operands when only the READ-ONLY x is the odd one out (axpy, copy, out-of-place scal; dot, sdsdot): 16-byte
       aligned packs of x + a lane shift (lbk_kernels.cuh, shift_pack).  The variant gives the unroll (its packs must be 16
       bytes); a one-element variant here turns the shifted path off (A/B against the shapes above). */
    /* map */ {19, 256, kOneTilePerBlock}, /* reduce */ {0, 256, 16},
};

constexpr int kMaxGrid = 148 * 32 * 2;   // partial-record scratch (blocks)

struct lbk_ctx {
    int device = 0;
    int sms = 0;
    cudaStream_t stream = nullptr;
    cudaMemPool_t pool = nullptr;     // this context's own stream-ordered pool (the device's default pool is left alone)
    // Handles (vectors, events) keep their context alive: lbk_shutdown on a context that still has handles
    // only closes it, and the last lbk_vec_free / lbk_event_destroy releases it (a garbage-collected host
    // language runs finalisers after the scope that closed the context).
    std::atomic<int> refs{0};
    bool closed = false;
    // ---- single-process multi-device context (lbk_init_multi): `subs` are the rank contexts, one per entry of
    // devs[]; every routine fans out over them.  A rank context points back through `parent`.
    std::vector<lbk_ctx *> subs;
    lbk_ctx *parent = nullptr;
    std::unique_ptr<lbk_team::Team> team;   // launcher threads (rank >= 1), when threaded
    bool threaded = false;
    bool host_mailbox = false;        // the mailboxes live in mapped host memory (no peer access between the devices)
    void *sink = nullptr;             // 64 device bytes: where ranks >= 1 put the scalar of an enqueue-only reduction
    // the stream-ordered exchange of a multi-device context (ranks that share a device, or fused exchange off):
    // this rank's record ring (two parities), the table of all ranks' rings, the "record written" event, and the
    // finish this rank still owes for the reduction it launched last (lbk_lib.cu, exchange_finish)
    // pinned staging lanes for synchronous transfers from / to PAGEABLE host memory (staged_transfer), made on first use
    void *stage_buf[16] = {};         // two chunks per lane (lane t: slots t and t + 8): the copy engine works on one while the
    cudaEvent_t stage_ev[16] = {};    // lane's thread fills / drains the other
    size_t stage_chunk = 0;
    Record *rec_ring = nullptr;
    const Record **ring_table_dev = nullptr;
    cudaEvent_t xchg_ev = nullptr;
    bool finish_pending = false;
    bool in_multi_call = false;       // set by the multi-device context around a fanned-out reduction
    FinishArgs pending_fa;
    unsigned ring_seq = 0;
    Record *partials = nullptr;
    unsigned *ticket = nullptr;
    double *snap = nullptr;           // two slots: snapshots of x[0], y[0] for zero output increments
    void *result_host = nullptr;      // mapped pinned 64 bytes: where synchronous reductions land
    void *result_dev = nullptr;
    unsigned result_seq = 0;          // tag of the latest synchronous result (host-polled, see finish_sync)
    Record *rec_send = nullptr;       // this rank's record
    Record *rec_all = nullptr;        // all ranks' records after the exchange
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0;
    // fused exchange over peer memory (CUDA IPC): this rank's mailbox, the peers' mapped mailboxes
    unsigned long long *mailbox = nullptr;
    unsigned long long **peers_dev = nullptr;     // device copy of peer_ptrs
    void *peer_ptrs[lbk::kMaxRanks] = {};
    bool fused_exchange = false;
    unsigned exchange_seq = 0;
    int *error_flag_host = nullptr, *error_flag_dev = nullptr;   // mapped: a kernel reports an exchange timeout
    unsigned long long exchange_timeout_ns = lbk::kExchangeTimeoutNs;
    LaunchCfg cfg[R_COUNT];
    lbk_pdl::State pdl;               // which kernels of the stream may overlap: see lbk_pdl.hpp
    int64_t launches = 0;
    int64_t pdl_launches = 0;
    int64_t shifted_launches = 0;     // unit-stride kernels that read x through shifted loads (lbk_shifted_count)
    // Where did the last unit-stride kernel that streamed a vector END?  What the 126 MB L2 still holds of a vector much larger
    // than itself is the part streamed last, so an elementwise kernel (whose result cannot depend on the order) starts its walk
    // there: backwards after a kernel that walked forwards (every reduction does: their summation order is part of the result),
    // forwards after one that walked backwards.  traversal: 0 always forwards, 1 adaptive (default), 2 always backwards (A/B).
    struct Walk { const void *ptr; bool ended_at_head; };
    Walk walks[8] = {};
    int walk_next = 0;
    int traversal = 1;
    // stream capture (lbk_capture_begin .. lbk_capture_end): kernels and asynchronous copies are recorded into a CUDA
    // graph instead of running; whatever cannot be recorded is refused, and buffers released meanwhile are freed at the end
    bool capturing = false;
    int64_t capture_launches0 = 0;
    std::vector<void *> deferred_free;
    std::string err;
};

struct lbk_vec {
    lbk_ctx *ctx;
    char *ptr;
    i64 len;          // local length, in elements
    i64 global_len;   // == len unless sharded
    i64 shard_off;    // global index of local element 0
    bool owner;
    bool sharded;
    int dtype;        // 0: float (Vector Float), 1: double (Vector Double)
    // a vector of a multi-device context: its parts on the rank contexts, in rank order -- every rank's block
    // shard when `sharded`, otherwise ONE part on rank 0 (the whole vector on devs[0]).  ptr is then part 0's.
    std::vector<lbk_vec *> parts;
    bool borrowed = false;   // a handle lbk_vec_part gave out: freeing it must not free the part
};
static inline int esz(int dtype) { return dtype ? 8 : 4; }

struct lbk_event { lbk_ctx *ctx; cudaEvent_t ev; std::vector<lbk_event *> parts; };
struct lbk_graph { lbk_ctx *ctx; cudaGraph_t graph; cudaGraphExec_t exec; int64_t kernels; };

static thread_local std::string g_err;   // failures before a context exists

// NCCL is bound with dlopen on first use instead of at link time: a process that also hosts PyTorch
// must end up with ONE libnccl.so.2 (torch bundles a newer one than the system's and fails to import if
// the older was mapped first), and single-GPU users need no NCCL at all.  RTLD_NOLOAD picks up whichever
// copy the process already has.
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*CommGetAsyncError)(ncclComm_t, ncclResult_t *) = nullptr;   // optional
    std::string error;
};
static NcclApi *nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) { api.error = std::string("dlopen(libnccl.so.2): ") + dlerror(); return; }
        api.handle = h;
        api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
        api.AllGather = (decltype(api.AllGather))dlsym(h, "ncclAllGather");
        api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
        api.CommGetAsyncError = (decltype(api.CommGetAsyncError))dlsym(h, "ncclCommGetAsyncError");
        if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather || !api.GetErrorString) {
            api.error = "libnccl.so.2 lacks a required symbol";
            api.handle = nullptr;
        }
    });
    return &api;
}

static int fail(lbk_ctx *ctx, int code, const std::string &msg)
{
    if (ctx) ctx->err = msg; else g_err = msg;
    return code;
}
#define CU(ctx, call)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail((ctx), LBK_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));  \
    } while (0)
#define NC(ctx, call)                                                                              \
    do {                                                                                           \
        ncclResult_t r_ = (call);                                                                  \
        if (r_ != ncclSuccess)                                                                     \
            return fail((ctx), LBK_ERR_NCCL, std::string(#call) + ": " + nccl_api()->GetErrorString(r_)); \
    } while (0)
#define KERNEL_CHECK(ctx)                                                                          \
    do {                                                                                           \
        (ctx)->launches++;                                                                         \
        cudaError_t e_ = cudaGetLastError();                                                       \
        if (e_ != cudaSuccess)                                                                     \
            return fail((ctx), LBK_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e_)); \
    } while (0)

// Every entry point that launches or allocates runs with the context's device current and puts the caller's
// device back afterwards: the host application (torch, another library) keeps its own current device.
struct DevGuard {
    int prev = -1;
    bool changed = false;
    cudaError_t err;
    explicit DevGuard(int dev)
    {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != dev) {
            err = cudaSetDevice(dev);
            changed = err == cudaSuccess;
        }
    }
    ~DevGuard() { if (changed) cudaSetDevice(prev); }
    DevGuard(const DevGuard &) = delete;
    DevGuard &operator=(const DevGuard &) = delete;
};
#define ON_DEVICE(ctx)                                                                             \
    DevGuard dev_guard_((ctx)->device);                                                            \
    if (dev_guard_.err != cudaSuccess)                                                             \
        return fail((ctx), LBK_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(dev_guard_.err))

static inline bool is_multi(const lbk_ctx *c) { return c && !c->subs.empty(); }
#define LIVE(ctx, who)                                                                             \
    if ((ctx)->closed) return fail((ctx), LBK_ERR_ARG, std::string(who) + ": the context has been shut down")

#define NOT_CAPTURING(ctx, who)                                                                    \
    if ((ctx)->capturing) return fail((ctx), LBK_ERR_UNSUPPORTED, std::string(who) + ": not while the context is capturing a graph (lbk_capture_begin)")

static inline i64 first_index(i64 n, i64 inc) { return inc > 0 ? 0 : (1 - n) * inc; }   // Single.hs:91-96

// Device-to-device copy of whole elements on the context's stream with the copy kernel (scopy's): faster than
// cudaMemcpyAsync on this part (7.0 vs 6.4 TB/s) and, unlike it, a member of the chain of overlapping kernels.
static int device_copy(lbk_ctx *c, void *dst, const void *src, i64 count, int dtype);
static inline i64 iabs64(i64 v) { return v < 0 ? -v : v; }
// Does the span 1 + (n-1)*|inc| of n >= 1 logical elements exceed `len` elements?  Stated without the product,
// which overflows int64 for large n x inc (and |INT64_MIN| does not exist).
static inline bool span_exceeds(i64 n, i64 inc, i64 len)
{
    if (len < 1) return true;
    if (inc == INT64_MIN) return n > 1;
    const i64 a = iabs64(inc);
    return a != 0 && n - 1 > (len - 1) / a;
}

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
extern "C" int lbk_version(void) { return LBK_VERSION; }

extern "C" int lbk_device_count(int *count)
{
    if (!count) return fail(nullptr, LBK_ERR_ARG, "lbk_device_count: null out pointer");
    CU(nullptr, cudaGetDeviceCount(count));
    return LBK_OK;
}

static void destroy_one(lbk_ctx *c);

// One rank context: a device, a stream, a private memory pool and the reduction scratch.
static int init_one(int device, lbk_ctx **out)
{
    *out = nullptr;
    int count = 0;
    CU(nullptr, cudaGetDeviceCount(&count));
    if (count <= 0) return fail(nullptr, LBK_ERR_CUDA, "lbk_init: no CUDA device (there is no CPU fallback)");
    if (device < 0 || device >= count) return fail(nullptr, LBK_ERR_ARG, "lbk_init: no such device");
    lbk_ctx *c = new (std::nothrow) lbk_ctx();
    if (!c) return fail(nullptr, LBK_ERR_NOMEM, "lbk_init: out of host memory");
    c->device = device;
    for (int r = 0; r < R_COUNT; ++r) c->cfg[r] = kDefaultCfg[r];
    if (const char *t = getenv("LBK_EXCHANGE_TIMEOUT_MS")) {
        const long ms = atol(t);
        if (ms > 0) c->exchange_timeout_ns = (unsigned long long)ms * 1000000ull;
    }
    DevGuard guard(device);
    cudaError_t e = guard.err;
    if (e != cudaSuccess) { delete c; return fail(nullptr, LBK_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e)); }
#define INIT(call) if ((e = (call)) != cudaSuccess) { std::string m = std::string(#call) + ": " + cudaGetErrorString(e); cudaGetLastError(); destroy_one(c); return fail(nullptr, LBK_ERR_CUDA, m); }
    INIT(cudaDeviceGetAttribute(&c->sms, cudaDevAttrMultiProcessorCount, device));
    INIT(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    {   // vectors come from a stream-ordered pool: the reference's pure routines allocate a result per call
        // (Util.hs:134-142), which must not cost a cudaMalloc + a device-wide synchronise each time.  The pool is
        // this context's own: the device's default pool (and whoever else uses it in the process) is untouched.
        cudaMemPoolProps props;
        memset(&props, 0, sizeof props);
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        INIT(cudaMemPoolCreate(&c->pool, &props));
        unsigned long long keep = ~0ull;      // freed blocks stay cached in the pool until the context goes
        INIT(cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    INIT(cudaMalloc(&c->partials, sizeof(Record) * kMaxGrid));
    INIT(cudaMalloc(&c->ticket, sizeof(unsigned)));
    INIT(cudaMemset(c->ticket, 0, sizeof(unsigned)));
    INIT(cudaMalloc(&c->snap, 2 * sizeof(double)));
    INIT(cudaMalloc(&c->sink, 64));
    INIT(cudaHostAlloc(&c->result_host, 64, cudaHostAllocMapped));
    if (c->result_host) memset(c->result_host, 0, 64);     // the tag the host polls starts from a known value
    INIT(cudaHostGetDevicePointer(&c->result_dev, c->result_host, 0));
    INIT(cudaHostAlloc((void **)&c->error_flag_host, sizeof(int), cudaHostAllocMapped));
    if (c->error_flag_host) *c->error_flag_host = 0;
    INIT(cudaHostGetDevicePointer((void **)&c->error_flag_dev, c->error_flag_host, 0));
    INIT(cudaMalloc(&c->rec_send, sizeof(Record)));
    INIT(cudaMalloc(&c->rec_all, sizeof(Record) * 64));
    INIT(cudaDeviceSynchronize());
#undef INIT
    *out = c;
    return LBK_OK;
}

// Releases everything a rank context owns (the stream has been synchronised by the caller when it matters).
static void destroy_one(lbk_ctx *c)
{
    if (!c) return;
    {
        DevGuard guard(c->device);
        if (c->stream) cudaStreamSynchronize(c->stream);
        if (c->comm) {
            for (int r = 0; r < c->nranks; ++r)
                if (r != c->rank && c->peer_ptrs[r]) cudaIpcCloseMemHandle(c->peer_ptrs[r]);
            nccl_api()->CommDestroy(c->comm);
        }
        if (c->mailbox) { if (c->host_mailbox) cudaFreeHost(c->mailbox); else cudaFree(c->mailbox); }
        if (c->rec_ring) { if (c->host_mailbox) cudaFreeHost(c->rec_ring); else cudaFree(c->rec_ring); }
        cudaFree(c->ring_table_dev);
        if (c->xchg_ev) cudaEventDestroy(c->xchg_ev);
        for (int i = 0; i < 16; ++i) { if (c->stage_buf[i]) cudaFreeHost(c->stage_buf[i]); if (c->stage_ev[i]) cudaEventDestroy(c->stage_ev[i]); }
        cudaFree(c->peers_dev);
        if (c->error_flag_host) cudaFreeHost(c->error_flag_host);
        cudaFree(c->partials); cudaFree(c->ticket); cudaFree(c->snap); cudaFree(c->sink);
        cudaFree(c->rec_send); cudaFree(c->rec_all);
        if (c->result_host) cudaFreeHost(c->result_host);
        if (c->pool) cudaMemPoolDestroy(c->pool);
        if (c->stream) cudaStreamDestroy(c->stream);
        cudaGetLastError();
    }
    delete c;
}

static std::mutex g_life;      // context lifetime: shutdown and the last handle's release may come from different threads
static inline lbk_ctx *root_of(lbk_ctx *c) { return c->parent ? c->parent : c; }
static int live_handles(lbk_ctx *root)
{
    int n = root->refs.load();
    for (lbk_ctx *s : root->subs) n += s->refs.load();
    return n;
}
static void destroy_tree(lbk_ctx *root)
{
    root->team.reset();
    for (lbk_ctx *s : root->subs) destroy_one(s);
    root->subs.clear();
    if (root->stream || root->pool) destroy_one(root); else delete root;
}
static inline void ctx_ref(lbk_ctx *c) { c->refs.fetch_add(1); }
// A handle of `c` is gone; a context that was shut down while handles were alive goes with its last handle.
static void ctx_unref(lbk_ctx *c)
{
    std::lock_guard<std::mutex> lk(g_life);
    c->refs.fetch_sub(1);
    lbk_ctx *root = root_of(c);
    if (root->closed && live_handles(root) == 0) destroy_tree(root);
}

extern "C" int lbk_init(int device, lbk_ctx **out)
{
    if (!out) return fail(nullptr, LBK_ERR_ARG, "lbk_init: null out pointer");
    return init_one(device, out);
}

// Runs fn(rank context, rank) for the first `count` ranks of a multi-device context: on the launcher threads
// when all ranks take part and the context is threaded, otherwise one after the other on the calling thread.
static int fan_out(lbk_ctx *m, int count, const std::function<int(lbk_ctx *, int)> &fn)
{
    const int n = (int)m->subs.size();
    for (int r = 0; r < n; ++r) m->subs[r]->err.clear();
    int status = LBK_OK;
    if (count == n && n > 1 && m->threaded && m->team) {
        status = m->team->run([&](int r) { return fn(m->subs[r], r); });
    } else {
        for (int r = 0; r < count && !status; ++r) status = fn(m->subs[r], r);
    }
    if (status)
        for (int r = 0; r < n; ++r)
            if (!m->subs[r]->err.empty()) { m->err = "rank " + std::to_string(r) + " (device " + std::to_string(m->subs[r]->device) + "): " + m->subs[r]->err; break; }
    return status;
}
static inline int all_ranks(lbk_ctx *m, const std::function<int(lbk_ctx *, int)> &fn) { return fan_out(m, (int)m->subs.size(), fn); }

static void make_team(lbk_ctx *m)
{
    if (m->team || m->subs.size() < 2) return;
    std::vector<int> devs;
    for (lbk_ctx *s : m->subs) devs.push_back(s->device);
    m->team.reset(new lbk_team::Team((int)devs.size(), [devs](int r) { cudaSetDevice(devs[r]); }));
}

// lbk_init_multi: ONE host thread drives the GPUs devs[0..ndev).  Rank r of the context is devs[r]; vectors made
// with lbk_vec_alloc_sharded are block shards over the ranks, every routine fans out over per-device streams,
// and a reduction ends inside its kernel: the last block of every rank pushes its 48-byte record into every
// other rank's mailbox -- plain peer pointers (cudaDeviceEnablePeerAccess; no IPC, no NCCL) -- and folds what
// it received in rank order.  The same device may appear more than once (two ranks then share a GPU, each with
// its own stream): that is how the 2-shard path is exercised on a one-GPU box.
extern "C" int lbk_init_multi(int ndev, const int *devs, lbk_ctx **out)
{
    if (!out) return fail(nullptr, LBK_ERR_ARG, "lbk_init_multi: null out pointer");
    *out = nullptr;
    if (ndev < 1 || ndev > lbk::kMaxRanks || !devs) return fail(nullptr, LBK_ERR_ARG, "lbk_init_multi: needs 1..64 devices");
    lbk_ctx *m = new (std::nothrow) lbk_ctx();
    if (!m) return fail(nullptr, LBK_ERR_NOMEM, "lbk_init_multi: out of host memory");
    auto bail = [&](int code, const std::string &msg) { destroy_tree(m); return fail(nullptr, code, msg); };
    bool dup = false;
    for (int r = 0; r < ndev; ++r) {
        lbk_ctx *s = nullptr;
        int st = init_one(devs[r], &s);
        if (st) { std::string msg = g_err; return bail(st, msg); }
        s->parent = m; s->nranks = ndev; s->rank = r;
        m->subs.push_back(s);
        for (int q = 0; q < r; ++q) dup = dup || devs[q] == devs[r];
    }
    m->device = devs[0]; m->sms = m->subs[0]->sms; m->nranks = ndev; m->rank = 0;
    for (int r = 0; r < R_COUNT; ++r) m->cfg[r] = kDefaultCfg[r];
    if (ndev > 1) {
        // peer access between every pair of distinct devices; without it the mailboxes live in mapped host memory
        bool peer = true;
        for (int a = 0; a < ndev && peer; ++a)
            for (int b = 0; b < ndev && peer; ++b) {
                if (devs[a] == devs[b]) continue;
                int ok = 0;
                if (cudaDeviceCanAccessPeer(&ok, devs[a], devs[b]) != cudaSuccess || !ok) { cudaGetLastError(); peer = false; }
            }
        for (int a = 0; a < ndev && peer; ++a) {
            DevGuard g(devs[a]);
            for (int b = 0; b < ndev; ++b) {
                if (devs[a] == devs[b]) continue;
                cudaError_t e = cudaDeviceEnablePeerAccess(devs[b], 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) peer = false;
                cudaGetLastError();
            }
        }
        const size_t mb = sizeof(unsigned long long) * lbk::kMailboxWords;
        for (int r = 0; r < ndev; ++r) {
            lbk_ctx *s = m->subs[r];
            DevGuard g(s->device);
            cudaError_t e;
            s->host_mailbox = !peer;
            if (peer) { e = cudaMalloc(&s->mailbox, mb); if (e == cudaSuccess) e = cudaMemset(s->mailbox, 0, mb); }
            else { e = cudaHostAlloc((void **)&s->mailbox, mb, cudaHostAllocPortable | cudaHostAllocMapped); if (e == cudaSuccess) memset(s->mailbox, 0, mb); }
            if (e == cudaSuccess) e = cudaMalloc(&s->peers_dev, sizeof(void *) * lbk::kMaxRanks);
            if (e == cudaSuccess) {
                if (peer) { e = cudaMalloc(&s->rec_ring, 2 * sizeof(Record)); if (e == cudaSuccess) e = cudaMemset(s->rec_ring, 0, 2 * sizeof(Record)); }
                else { e = cudaHostAlloc((void **)&s->rec_ring, 2 * sizeof(Record), cudaHostAllocPortable | cudaHostAllocMapped); if (e == cudaSuccess) memset(s->rec_ring, 0, 2 * sizeof(Record)); }
            }
            if (e == cudaSuccess) e = cudaMalloc(&s->ring_table_dev, sizeof(void *) * lbk::kMaxRanks);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->xchg_ev, cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaDeviceSynchronize();
            if (e != cudaSuccess) return bail(LBK_ERR_CUDA, std::string("lbk_init_multi: mailbox: ") + cudaGetErrorString(e));
        }
        for (int r = 0; r < ndev; ++r) {
            lbk_ctx *s = m->subs[r];
            DevGuard g(s->device);
            for (int q = 0; q < ndev; ++q) s->peer_ptrs[q] = m->subs[q]->mailbox;     // unified addressing: the same pointer on every device
            cudaError_t e = cudaMemcpy(s->peers_dev, s->peer_ptrs, sizeof(void *) * lbk::kMaxRanks, cudaMemcpyHostToDevice);
            const Record *rings[lbk::kMaxRanks] = {};
            for (int q = 0; q < ndev; ++q) rings[q] = m->subs[q]->rec_ring;
            if (e == cudaSuccess) e = cudaMemcpy(s->ring_table_dev, rings, sizeof(void *) * lbk::kMaxRanks, cudaMemcpyHostToDevice);
            if (e != cudaSuccess) return bail(LBK_ERR_CUDA, std::string("lbk_init_multi: peer table: ") + cudaGetErrorString(e));
            // Ranks that SHARE a device must not wait for one another inside kernels: nothing guarantees that two
            // kernels of one GPU run at the same time (and the next kernel of a stream, resident early under
            // programmatic dependent launch, can hold the slots the awaited kernel needs).  There the exchange is
            // stream-ordered instead: records, events, a finish kernel (exchange_finish).
            s->fused_exchange = !dup;
            s->exchange_seq = 0;
        }
        m->host_mailbox = !peer;
        m->fused_exchange = !dup;
        const char *t = getenv("LBK_MULTI_THREADS");
        m->threaded = t ? atoi(t) != 0 : !dup;
        if (m->threaded) make_team(m);
    }
    *out = m;
    return LBK_OK;
}

extern "C" int lbk_multi_rank(lbk_ctx *c, int rank, lbk_ctx **out)
{
    if (!c || !out) return fail(c, LBK_ERR_ARG, "lbk_multi_rank: null argument");
    if (!is_multi(c) || rank < 0 || rank >= (int)c->subs.size()) return fail(c, LBK_ERR_ARG, "lbk_multi_rank: not a multi-device context, or no such rank");
    *out = c->subs[rank];
    return LBK_OK;
}

// The scalar rank `rank` itself holds for the context's most recent SYNCHRONOUS reduction (8 bytes: a float or double
// in the low bytes, or the int64 index): every rank computes the result from the same records in the same order, and
// this is how a test sees that they all hold the same bits.  Synchronises that rank's stream.
extern "C" int lbk_multi_result(lbk_ctx *c, int rank, void *out8)
{
    if (!c || !out8) return fail(c, LBK_ERR_ARG, "lbk_multi_result: null argument");
    if (!is_multi(c) || rank < 0 || rank >= (int)c->subs.size()) return fail(c, LBK_ERR_ARG, "lbk_multi_result: not a multi-device context, or no such rank");
    lbk_ctx *s = c->subs[rank];
    DevGuard guard(s->device);
    CU(c, cudaStreamSynchronize(s->stream));
    memcpy(out8, s->result_host, 8);
    return LBK_OK;
}

extern "C" int lbk_set_multi_threads(lbk_ctx *c, int enabled)
{
    if (!c) return fail(c, LBK_ERR_ARG, "lbk_set_multi_threads: null context");
    if (!is_multi(c)) return fail(c, LBK_ERR_ARG, "lbk_set_multi_threads: not a multi-device context");
    c->threaded = enabled != 0;
    if (c->threaded) make_team(c);
    return LBK_OK;
}

extern "C" int lbk_shutdown(lbk_ctx *c)
{
    if (!c) return LBK_OK;
    if (c->parent) return fail(c, LBK_ERR_ARG, "lbk_shutdown: a rank context goes with its multi-device context");
    std::lock_guard<std::mutex> lk(g_life);
    if (c->closed) return LBK_OK;
    for (lbk_ctx *s : c->subs) { DevGuard g(s->device); if (s->stream) cudaStreamSynchronize(s->stream); s->closed = true; }
    if (c->stream) {
        DevGuard g(c->device);
        if (c->capturing) {                    // shut down in the middle of a capture: drop what was recorded
            cudaGraph_t rec = nullptr;
            cudaStreamEndCapture(c->stream, &rec);
            if (rec) cudaGraphDestroy(rec);
            cudaGetLastError();
            c->capturing = false;
            for (void *p : c->deferred_free) cudaFreeAsync(p, c->stream);
            c->deferred_free.clear();
        }
        cudaStreamSynchronize(c->stream);
    }
    c->team.reset();
    c->closed = true;
    if (live_handles(c) == 0) destroy_tree(c);      // otherwise the last lbk_vec_free / lbk_event_destroy does it
    return LBK_OK;
}

static int check_exchange_error(lbk_ctx *c)
{
    if (c->error_flag_host && *reinterpret_cast<volatile int *>(c->error_flag_host)) {
        *c->error_flag_host = 0;
        return fail(c, LBK_ERR_NCCL, "peer exchange timed out: a rank never delivered its record (did every rank make the same call?)");
    }
    if (c->comm && nccl_api()->CommGetAsyncError) {      // a failure NCCL noticed on its own threads
        ncclResult_t async = ncclSuccess;
        if (nccl_api()->CommGetAsyncError(c->comm, &async) == ncclSuccess && async != ncclSuccess && async != ncclInProgress)
            return fail(c, LBK_ERR_NCCL, std::string("NCCL asynchronous error: ") + nccl_api()->GetErrorString(async));
    }
    return LBK_OK;
}

extern "C" int lbk_sync(lbk_ctx *c)
{
    if (!c) return fail(nullptr, LBK_ERR_ARG, "lbk_sync: null context");
    LIVE(c, "lbk_sync");
    if (is_multi(c)) return all_ranks(c, [](lbk_ctx *s, int) { return lbk_sync(s); });
    NOT_CAPTURING(c, "lbk_sync");
    ON_DEVICE(c);
    CU(c, cudaStreamSynchronize(c->stream));
    return check_exchange_error(c);
}

extern "C" const char *lbk_last_error(lbk_ctx *c) { return c ? c->err.c_str() : g_err.c_str(); }
extern "C" void *lbk_stream(lbk_ctx *c) { return !c ? nullptr : is_multi(c) ? (void *)c->subs[0]->stream : (void *)c->stream; }

extern "C" int lbk_sm_count(lbk_ctx *c, int *sms)
{
    if (!c || !sms) return fail(c, LBK_ERR_ARG, "lbk_sm_count: null argument");
    *sms = c->sms;
    return LBK_OK;
}

extern "C" int lbk_set_launch(lbk_ctx *c, const char *routine, int variant, int threads, int blocks_per_sm)
{
    if (!c || !routine) return fail(c, LBK_ERR_ARG, "lbk_set_launch: null argument");
    if (is_multi(c)) {
        for (lbk_ctx *s : c->subs) { int st = lbk_set_launch(s, routine, variant, threads, blocks_per_sm); if (st) { c->err = s->err; return st; } }
        return LBK_OK;
    }
    if (variant >= kNumVariants) return fail(c, LBK_ERR_ARG, "lbk_set_launch: no such variant");
    if (threads > lbk::kMaxThreads || (threads > 0 && threads % 32 != 0))
        return fail(c, LBK_ERR_ARG, "lbk_set_launch: threads must be a multiple of 32, at most 512");
    // the BLAS spellings name the same shapes: sdot/ddot -> dot, isamax/idamax -> iamax, ...
    std::string name(routine);
    if (name != "sdsdot" && name != "all") {
        if (name.size() > 2 && name[0] == 'i' && (name[1] == 's' || name[1] == 'd')) name = "i" + name.substr(2);
        else if (name.size() > 1 && (name[0] == 's' || name[0] == 'd') && name != "scal" && name != "swap" && name != "dot") name = name.substr(1);
    }
    routine = name.c_str();
    for (int r = 0; r < R_COUNT; ++r)
        if (!strcmp(routine, kRoutineNames[r]) || !strcmp(routine, "all")) {
            c->cfg[r] = LaunchCfg{variant >= 0 ? variant : kDefaultCfg[r].variant, threads > 0 ? threads : kDefaultCfg[r].threads,
                                  (variant < 0 && threads <= 0 && blocks_per_sm <= 0) ? kDefaultCfg[r].blocks_per_sm : blocks_per_sm,
                                  blocks_per_sm > 0};      // an explicit grid is taken literally (no size adaptation)
            if (strcmp(routine, "all")) return LBK_OK;
        }
    if (!strcmp(routine, "all")) return LBK_OK;
    return fail(c, LBK_ERR_ARG, std::string("lbk_set_launch: unknown routine ") + routine);
}

extern "C" int lbk_launch_count(lbk_ctx *c, int64_t *launches)
{
    if (!c || !launches) return fail(c, LBK_ERR_ARG, "lbk_launch_count: null argument");
    *launches = c->launches;
    for (lbk_ctx *s : c->subs) *launches += s->launches;
    return LBK_OK;
}

extern "C" int lbk_set_pdl(lbk_ctx *c, int enabled)
{
    if (!c) return fail(c, LBK_ERR_ARG, "lbk_set_pdl: null context");
    for (lbk_ctx *s : c->subs) lbk_set_pdl(s, enabled);
    c->pdl.enabled = enabled != 0;
    c->pdl.hold_stores = enabled != 2 && enabled != 3;      // 2: dependent kernels (of any kind) wait before their first load
    c->pdl.hold_loads = enabled != 3 && enabled != 4;       // 3: ... or get an ordinary launch; 4: only true dependencies do (tools/, experiments)
    return LBK_OK;
}

extern "C" int lbk_pdl_count(lbk_ctx *c, int64_t *launches)
{
    if (!c || !launches) return fail(c, LBK_ERR_ARG, "lbk_pdl_count: null argument");
    *launches = c->pdl_launches;
    for (lbk_ctx *s : c->subs) *launches += s->pdl_launches;
    return LBK_OK;
}

// ------------------------------------------------------------------------------------------------
// CUDA graphs.  A fixed sequence of small calls (the regime of the reference's own benchmark, test/Bench.hs: n <= 30 000)
// is bound by launch latency; recorded once into a graph it replays with one host call and the GPU's own node-to-node
// latency (profiles/r02_graph_latency.md).  Between begin and end every call on the context is RECORDED, not run: kernels (with
// their programmatic-dependent-launch edges -- the hazard planner works as usual, capture starts and ends a chain) and
// asynchronous copies.  Whatever cannot be recorded is refused with LBK_ERR_UNSUPPORTED before anything reaches the stream:
// allocation (pure forms: use the in-place / *_oop forms on vectors that exist), scalars returned to the host (use the _async
// reductions), synchronous transfers, lbk_sync, events, sharded reductions.  A replay repeats the recorded calls on the
// recorded addresses with the recorded scalars.
// ------------------------------------------------------------------------------------------------
extern "C" int lbk_capture_begin(lbk_ctx *c)
{
    if (!c) return fail(c, LBK_ERR_ARG, "lbk_capture_begin: null context");
    LIVE(c, "lbk_capture_begin");
    if (is_multi(c) || c->parent) return fail(c, LBK_ERR_UNSUPPORTED, "lbk_capture_begin: single-device contexts only");
    NOT_CAPTURING(c, "lbk_capture_begin");
    ON_DEVICE(c);
    lbk_pdl::interrupt(c->pdl);
    // relaxed: the library itself keeps everything it cannot record off the capturing stream, and a host language's finalisers
    // (cudaEventDestroy, cudaFreeHost, another graph's destruction ...) may run on this thread at any time
    CU(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed));
    c->capturing = true;
    c->capture_launches0 = c->launches;
    return LBK_OK;
}

extern "C" int lbk_capture_end(lbk_ctx *c, lbk_graph **out)
{
    if (!c || !out) return fail(c, LBK_ERR_ARG, "lbk_capture_end: null argument");
    if (!c->capturing) return fail(c, LBK_ERR_ARG, "lbk_capture_end: the context is not capturing");
    ON_DEVICE(c);
    c->capturing = false;
    lbk_pdl::interrupt(c->pdl);
    const int64_t recorded = c->launches - c->capture_launches0;
    c->launches = c->capture_launches0;        // recorded, not launched -- whatever becomes of the graph
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(c->stream, &g);
    for (void *p : c->deferred_free) cudaFreeAsync(p, c->stream);      // finalisers that ran during the capture
    c->deferred_free.clear();
    if (e != cudaSuccess || !g) { cudaGetLastError(); return fail(c, LBK_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e)); }
    cudaGraphExec_t x = nullptr;
    e = cudaGraphInstantiate(&x, g, 0);
    if (e != cudaSuccess) { cudaGraphDestroy(g); cudaGetLastError(); return fail(c, LBK_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e)); }
    lbk_graph *h = new (std::nothrow) lbk_graph{c, g, x, recorded};
    if (!h) { cudaGraphExecDestroy(x); cudaGraphDestroy(g); return fail(c, LBK_ERR_NOMEM, "out of host memory"); }
    ctx_ref(c);
    *out = h;
    return LBK_OK;
}

extern "C" int lbk_graph_launch(lbk_graph *g)
{
    if (!g) return fail(nullptr, LBK_ERR_ARG, "lbk_graph_launch: null graph");
    lbk_ctx *c = g->ctx;
    LIVE(c, "lbk_graph_launch");
    NOT_CAPTURING(c, "lbk_graph_launch");
    ON_DEVICE(c);
    lbk_pdl::interrupt(c->pdl);                // the graph is an ordinary stream operation: what follows is launched the ordinary way
    CU(c, cudaGraphLaunch(g->exec, c->stream));
    c->launches += g->kernels;
    return LBK_OK;
}

extern "C" int lbk_graph_kernels(const lbk_graph *g, int64_t *kernels)
{
    if (!g || !kernels) return fail(nullptr, LBK_ERR_ARG, "lbk_graph_kernels: null argument");
    *kernels = g->kernels;
    return LBK_OK;
}

extern "C" int lbk_graph_destroy(lbk_graph *g)
{
    if (!g) return LBK_OK;
    lbk_ctx *c = g->ctx;
    {
        DevGuard guard(c->device);
        cudaGraphExecDestroy(g->exec);         // a launch still in flight keeps what it needs (CUDA defers the release)
        cudaGraphDestroy(g->graph);
    }
    delete g;
    ctx_unref(c);
    return LBK_OK;
}

constexpr i64 kWalkMinBytes = (i64)32 << 20;     // below this an operand fits the L2 whole and the direction is moot
static int walk_vote(const lbk_ctx *c, const void *p)
{
    for (const auto &w : c->walks) if (w.ptr == p && p) return w.ended_at_head ? -1 : +1;
    return 0;
}
static void walk_note(lbk_ctx *c, const void *p, bool ended_at_head)
{
    if (!p) return;
    for (auto &w : c->walks) if (w.ptr == p) { w.ended_at_head = ended_at_head; return; }
    c->walks[c->walk_next] = {p, ended_at_head};
    c->walk_next = (c->walk_next + 1) % (int)(sizeof c->walks / sizeof c->walks[0]);
}
static void walk_forget(lbk_ctx *c) { for (auto &w : c->walks) w.ptr = nullptr; }

extern "C" int lbk_set_traversal(lbk_ctx *c, int mode)
{
    if (!c || mode < 0 || mode > 2) return fail(c, LBK_ERR_ARG, "lbk_set_traversal: mode is 0 (forwards), 1 (adaptive) or 2 (backwards)");
    for (lbk_ctx *s : c->subs) lbk_set_traversal(s, mode);
    c->traversal = mode;
    walk_forget(c);
    return LBK_OK;
}

extern "C" int lbk_shifted_count(lbk_ctx *c, int64_t *launches)
{
    if (!c || !launches) return fail(c, LBK_ERR_ARG, "lbk_shifted_count: null argument");
    *launches = c->shifted_launches;
    for (lbk_ctx *s : c->subs) *launches += s->shifted_launches;
    return LBK_OK;
}

extern "C" int lbk_set_exchange_timeout(lbk_ctx *c, int64_t milliseconds)
{
    if (!c || milliseconds <= 0) return fail(c, LBK_ERR_ARG, "lbk_set_exchange_timeout: needs a context and a positive time");
    for (lbk_ctx *s : c->subs) s->exchange_timeout_ns = (unsigned long long)milliseconds * 1000000ull;
    c->exchange_timeout_ns = (unsigned long long)milliseconds * 1000000ull;
    return LBK_OK;
}


// ------------------------------------------------------------------------------------------------
// communicator
// ------------------------------------------------------------------------------------------------
extern "C" int lbk_comm_unique_id(char id[128])
{
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    if (!id) return fail(nullptr, LBK_ERR_ARG, "lbk_comm_unique_id: null buffer");
    if (!nccl_api()->handle) return fail(nullptr, LBK_ERR_NCCL, nccl_api()->error);
    ncclUniqueId u;
    NC(nullptr, nccl_api()->GetUniqueId(&u));
    memcpy(id, &u, 128);
    return LBK_OK;
}

// Maps every rank's mailbox into this process (CUDA IPC over NVLink peer access).  All ranks then agree on
// whether the fused exchange is usable; if any rank could not map a peer, all use the NCCL all-gather.
static int setup_fused_exchange(lbk_ctx *c)
{
    const int n = c->nranks;
    struct Card { cudaIpcMemHandle_t h; int ok; int pad[3]; };
    Card mine;
    memset(&mine, 0, sizeof mine);
    Card *dev = nullptr;
    CU(c, cudaMalloc(&c->mailbox, sizeof(unsigned long long) * lbk::kMailboxWords));
    CU(c, cudaMemset(c->mailbox, 0, sizeof(unsigned long long) * lbk::kMailboxWords));
    CU(c, cudaMalloc(&c->peers_dev, sizeof(void *) * lbk::kMaxRanks));
    CU(c, cudaMalloc(&dev, sizeof(Card) * (n + 1)));
    mine.ok = cudaIpcGetMemHandle(&mine.h, c->mailbox) == cudaSuccess ? 1 : 0;
    cudaGetLastError();
    std::vector<Card> cards(n);
    // round 1: everybody's handle
    CU(c, cudaMemcpyAsync(dev + n, &mine, sizeof(Card), cudaMemcpyHostToDevice, c->stream));
    NC(c, nccl_api()->AllGather(dev + n, dev, sizeof(Card), ncclChar, c->comm, c->stream));
    CU(c, cudaMemcpyAsync(cards.data(), dev, sizeof(Card) * n, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    int ok = 1;
    for (int r = 0; r < n; ++r) {
        c->peer_ptrs[r] = nullptr;
        if (!cards[r].ok) { ok = 0; continue; }
        if (r == c->rank) { c->peer_ptrs[r] = c->mailbox; continue; }
        if (cudaIpcOpenMemHandle(&c->peer_ptrs[r], cards[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            c->peer_ptrs[r] = nullptr;
            ok = 0;
        }
    }
    // round 2: does every rank see every mailbox?
    mine.ok = ok;
    CU(c, cudaMemcpyAsync(dev + n, &mine, sizeof(Card), cudaMemcpyHostToDevice, c->stream));
    NC(c, nccl_api()->AllGather(dev + n, dev, sizeof(Card), ncclChar, c->comm, c->stream));
    CU(c, cudaMemcpyAsync(cards.data(), dev, sizeof(Card) * n, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    for (int r = 0; r < n; ++r) ok = ok && cards[r].ok;
    CU(c, cudaMemcpy(c->peers_dev, c->peer_ptrs, sizeof(void *) * lbk::kMaxRanks, cudaMemcpyHostToDevice));
    cudaFree(dev);
    c->fused_exchange = ok != 0;
    c->exchange_seq = 0;
    return LBK_OK;
}

extern "C" int lbk_comm_init(lbk_ctx *c, int nranks, int rank, const char id[128])
{
    if (!c || !id) return fail(c, LBK_ERR_ARG, "lbk_comm_init: null argument");
    if (nranks < 1 || nranks > 64 || rank < 0 || rank >= nranks) return fail(c, LBK_ERR_ARG, "lbk_comm_init: bad rank / world size");
    LIVE(c, "lbk_comm_init");
    if (is_multi(c) || c->parent) return fail(c, LBK_ERR_UNSUPPORTED, "lbk_comm_init: a multi-device context already spans its ranks (one process); NCCL joins one-device contexts");
    if (c->comm) return fail(c, LBK_ERR_ARG, "lbk_comm_init: communicator already attached");
    if (!nccl_api()->handle) return fail(c, LBK_ERR_NCCL, nccl_api()->error);
    ON_DEVICE(c);
    ncclUniqueId u;
    memcpy(&u, id, 128);
    NC(c, nccl_api()->CommInitRank(&c->comm, nranks, u, rank));
    c->nranks = nranks;
    c->rank = rank;
    return setup_fused_exchange(c);
}

extern "C" int lbk_comm_transport(lbk_ctx *c, int *fused)
{
    if (!c || !fused) return fail(c, LBK_ERR_ARG, "lbk_comm_transport: null argument");
    *fused = c->fused_exchange ? 1 : 0;
    return LBK_OK;
}

extern "C" int lbk_set_fused_exchange(lbk_ctx *c, int enabled)
{
    if (!c) return fail(c, LBK_ERR_ARG, "lbk_set_fused_exchange: null context");
    if (c->parent) return fail(c, LBK_ERR_ARG, "lbk_set_fused_exchange: set it on the multi-device context, for all ranks");
    if (is_multi(c)) {      // in-kernel over the peer mailboxes (1) or stream-ordered: records, events, a finish kernel (0)
        for (lbk_ctx *s : c->subs) { s->fused_exchange = enabled != 0 && c->subs.size() > 1; }
        c->fused_exchange = enabled != 0 && c->subs.size() > 1;
        return LBK_OK;
    }
    if (enabled && !c->mailbox) return fail(c, LBK_ERR_UNSUPPORTED, "lbk_set_fused_exchange: peer mailboxes were not mapped on this communicator");
    c->fused_exchange = enabled != 0;
    return LBK_OK;
}

extern "C" int lbk_comm_info(lbk_ctx *c, int *nranks, int *rank)
{
    if (!c) return fail(c, LBK_ERR_ARG, "lbk_comm_info: null context");
    if (nranks) *nranks = c->nranks;
    if (rank) *rank = c->rank;
    return LBK_OK;
}

// contiguous block partition: rank r owns [r*q + min(r,rem), ...) with q = len / nranks
extern "C" int lbk_shard_range(int64_t global_len, int nranks, int rank, int64_t *lo, int64_t *hi)
{
    if (global_len < 0 || nranks < 1 || rank < 0 || rank >= nranks || !lo || !hi)
        return fail(nullptr, LBK_ERR_ARG, "lbk_shard_range: bad argument");
    const i64 q = global_len / nranks, rem = global_len % nranks;
    *lo = rank * q + std::min<i64>(rank, rem);
    *hi = *lo + q + (rank < rem ? 1 : 0);
    return LBK_OK;
}

// ------------------------------------------------------------------------------------------------
// vectors
// ------------------------------------------------------------------------------------------------
static int make_vec(lbk_ctx *c, void *ptr, i64 len, i64 glen, i64 off, bool owner, bool sharded, int dtype, lbk_vec **out)
{
    lbk_vec *v = new (std::nothrow) lbk_vec{c, (char *)ptr, len, glen, off, owner, sharded, dtype, {}, false};
    if (!v) return fail(c, LBK_ERR_NOMEM, "out of host memory");
    ctx_ref(c);
    *out = v;
    return LBK_OK;
}

static int alloc_common(lbk_ctx *c, const char *who, i64 len, i64 glen, i64 off, bool sharded, int dtype, lbk_vec **out)
{
    LIVE(c, who);
    if (len > 0) NOT_CAPTURING(c, who);        // allocate before capturing: a replay reuses the recorded addresses
    ON_DEVICE(c);
    void *p = nullptr;
    if (len > 0) {
        if ((unsigned long long)len > (~0ull >> 4)) return fail(c, LBK_ERR_NOMEM, std::string(who) + ": length overflows the address space");
        // a stream-ordered allocation does not end the chain of overlapping kernels: hazards are tracked by address,
        // so a block the pool hands out again while its previous user is still running is handled like any other
#ifdef LBK_SYNC_MALLOC      // A/B builds only (profiles/r01_criterion_matrix*.csv): plain cudaMalloc / synchronise + cudaFree
        cudaError_t e = cudaMalloc(&p, (size_t)esz(dtype) * (size_t)len);
#else
        cudaError_t e = cudaMallocFromPoolAsync(&p, (size_t)esz(dtype) * (size_t)len, c->pool, c->stream);
#endif
        if (e != cudaSuccess) { cudaGetLastError(); return fail(c, LBK_ERR_NOMEM, std::string(who) + ": cudaMallocFromPoolAsync: " + cudaGetErrorString(e)); }
    }
    return make_vec(c, p, len, glen, off, true, sharded, dtype, out);
}

// The handle of a multi-device context's vector: `parts` on the rank contexts (all ranks when sharded, rank 0 only
// otherwise).  It takes the geometry of the logical vector; ptr is part 0's buffer.
static int make_multi_vec(lbk_ctx *m, std::vector<lbk_vec *> &&parts, i64 glen, bool sharded, int dtype, bool owner, lbk_vec **out)
{
    lbk_vec *v = new (std::nothrow) lbk_vec{m, parts.empty() ? nullptr : parts[0]->ptr, sharded ? glen : parts[0]->len, glen, 0, owner, sharded, dtype, {}, false};
    if (!v) { for (lbk_vec *p : parts) lbk_vec_free(p); return fail(m, LBK_ERR_NOMEM, "out of host memory"); }
    v->parts = std::move(parts);
    ctx_ref(m);
    *out = v;
    return LBK_OK;
}
// part of a multi-device vector on rank r (a vector that is not sharded lives on rank 0 alone)
static inline lbk_vec *part(const lbk_vec *v, int r) { return !v ? nullptr : v->sharded ? v->parts[r] : v->parts[0]; }

static int multi_alloc(lbk_ctx *m, const char *who, i64 len, bool sharded, int dtype, lbk_vec **out)
{
    LIVE(m, who);
    const int n = sharded ? (int)m->subs.size() : 1;
    std::vector<lbk_vec *> parts(n, nullptr);
    int s = fan_out(m, n, [&](lbk_ctx *sc, int r) {
        int64_t lo = 0, hi = len;
        if (sharded) lbk_shard_range(len, sc->nranks, r, &lo, &hi);
        return alloc_common(sc, who, hi - lo, len, lo, sharded && sc->nranks > 1, dtype, &parts[r]);
    });
    if (s) { for (lbk_vec *p : parts) lbk_vec_free(p); return s; }
    return make_multi_vec(m, std::move(parts), len, sharded, dtype, true, out);
}

static int vec_alloc(lbk_ctx *c, int64_t len, int dtype, lbk_vec **out)
{
    if (!c || !out || len < 0) return fail(c, LBK_ERR_ARG, "lbk_vec_alloc: bad argument");
    if (is_multi(c)) return multi_alloc(c, "lbk_vec_alloc", len, false, dtype, out);
    return alloc_common(c, "lbk_vec_alloc", len, len, 0, false, dtype, out);
}
static int vec_alloc_sharded(lbk_ctx *c, int64_t global_len, int dtype, lbk_vec **out)
{
    if (!c || !out || global_len < 0) return fail(c, LBK_ERR_ARG, "lbk_vec_alloc_sharded: bad argument");
    if (is_multi(c)) return multi_alloc(c, "lbk_vec_alloc_sharded", global_len, true, dtype, out);
    int64_t lo, hi;
    lbk_shard_range(global_len, c->nranks, c->rank, &lo, &hi);
    return alloc_common(c, "lbk_vec_alloc_sharded", hi - lo, global_len, lo, c->nranks > 1, dtype, out);
}
static int vec_wrap(lbk_ctx *c, void *device_ptr, int64_t len, int dtype, lbk_vec **out)
{
    if (!c || !out || len < 0 || (len > 0 && !device_ptr) || ((uintptr_t)device_ptr & (uintptr_t)(esz(dtype) - 1)))
        return fail(c, LBK_ERR_ARG, "lbk_vec_wrap: bad argument (null, negative or not element aligned)");
    LIVE(c, "lbk_vec_wrap");
    if (is_multi(c)) {      // foreign memory is one device's: the wrapped vector lives on rank 0 (devs[0])
        std::vector<lbk_vec *> parts(1, nullptr);
        int s = vec_wrap(c->subs[0], device_ptr, len, dtype, &parts[0]);
        if (s) { c->err = c->subs[0]->err; return s; }
        return make_multi_vec(c, std::move(parts), len, false, dtype, false, out);
    }
    return make_vec(c, device_ptr, len, len, 0, false, false, dtype, out);
}
extern "C" int lbk_vec_alloc(lbk_ctx *c, int64_t len, lbk_vec **out) { return vec_alloc(c, len, 0, out); }
extern "C" int lbk_vec_alloc_f64(lbk_ctx *c, int64_t len, lbk_vec **out) { return vec_alloc(c, len, 1, out); }
extern "C" int lbk_vec_alloc_sharded(lbk_ctx *c, int64_t glen, lbk_vec **out) { return vec_alloc_sharded(c, glen, 0, out); }
extern "C" int lbk_vec_alloc_sharded_f64(lbk_ctx *c, int64_t glen, lbk_vec **out) { return vec_alloc_sharded(c, glen, 1, out); }
extern "C" int lbk_vec_wrap(lbk_ctx *c, void *p, int64_t len, lbk_vec **out) { return vec_wrap(c, p, len, 0, out); }
extern "C" int lbk_vec_wrap_f64(lbk_ctx *c, void *p, int64_t len, lbk_vec **out) { return vec_wrap(c, p, len, 1, out); }

// An uninitialised vector with v's length, element type and layout (sharded like v, on the same ranks): what a
// pure routine's result is made of.
extern "C" int lbk_vec_alloc_like(const lbk_vec *v, lbk_vec **out)
{
    if (!v || !out) return fail(v ? v->ctx : nullptr, LBK_ERR_ARG, "lbk_vec_alloc_like: null argument");
    lbk_ctx *c = v->ctx;
    if (is_multi(c)) return multi_alloc(c, "lbk_vec_alloc_like", v->global_len, v->sharded, v->dtype, out);
    return alloc_common(c, "lbk_vec_alloc_like", v->len, v->global_len, v->shard_off, v->sharded, v->dtype, out);
}

extern "C" int lbk_vec_view(const lbk_vec *v, int64_t offset, int64_t len, lbk_vec **out)
{
    if (!v || !out || offset < 0 || len < 0 || offset > v->len || len > v->len - offset)
        return fail(v ? v->ctx : nullptr, LBK_ERR_ARG, "lbk_vec_view: window outside the vector");
    if (v->sharded) return fail(v->ctx, LBK_ERR_UNSUPPORTED, "lbk_vec_view: views of sharded vectors are not supported");
    LIVE(v->ctx, "lbk_vec_view");
    if (is_multi(v->ctx)) {
        std::vector<lbk_vec *> parts(1, nullptr);
        int s = lbk_vec_view(v->parts[0], offset, len, &parts[0]);
        if (s) { v->ctx->err = v->parts[0]->ctx->err; return s; }
        return make_multi_vec(v->ctx, std::move(parts), len, false, v->dtype, false, out);
    }
    return make_vec(v->ctx, v->ptr + offset * esz(v->dtype), len, len, 0, false, false, v->dtype, out);
}

// Rank r's part of a multi-device vector, as a handle of that rank's context (lbk_multi_rank): borrowed -- it
// stays valid as long as `v` and freeing it releases only the handle.
extern "C" int lbk_vec_part(const lbk_vec *v, int rank, lbk_vec **out)
{
    if (!v || !out) return fail(v ? v->ctx : nullptr, LBK_ERR_ARG, "lbk_vec_part: null argument");
    if (!is_multi(v->ctx) || rank < 0 || rank >= (int)v->parts.size()) return fail(v->ctx, LBK_ERR_ARG, "lbk_vec_part: not a multi-device vector, or no such part");
    const lbk_vec *p = v->parts[rank];
    int s = make_vec(p->ctx, p->ptr, p->len, p->global_len, p->shard_off, false, p->sharded, p->dtype, out);
    if (!s) (*out)->borrowed = true;
    return s;
}

extern "C" int lbk_vec_free(lbk_vec *v)
{
    if (!v) return LBK_OK;
    lbk_ctx *c = v->ctx;
    for (lbk_vec *p : v->parts) lbk_vec_free(p);
    if (v->parts.empty() && v->owner && v->ptr) {      // released in stream order: everything enqueued so far may still use it
        DevGuard guard(c->device);
#ifdef LBK_SYNC_MALLOC
        cudaStreamSynchronize(c->stream);
        cudaFree(v->ptr);
#else
        if (c->capturing) c->deferred_free.push_back(v->ptr);      // a finaliser ran inside a capture: released at its end
        else cudaFreeAsync(v->ptr, c->stream);
#endif
    }
    delete v;
    ctx_unref(c);          // the last handle of a context that was shut down takes the context with it
    return LBK_OK;
}

extern "C" int lbk_vec_len(const lbk_vec *v, int64_t *local_len, int64_t *global_len, int64_t *shard_offset)
{
    if (!v) return fail(nullptr, LBK_ERR_ARG, "lbk_vec_len: null vector");
    if (local_len) *local_len = v->len;
    if (global_len) *global_len = v->global_len;
    if (shard_offset) *shard_offset = v->shard_off;
    return LBK_OK;
}
extern "C" int lbk_vec_is_sharded(const lbk_vec *v) { return v && v->sharded ? 1 : 0; }
extern "C" int lbk_vec_elem_size(const lbk_vec *v) { return v ? esz(v->dtype) : 0; }
extern "C" void *lbk_vec_ptr(const lbk_vec *v) { return v ? (void *)v->ptr : nullptr; }

// A multi-device vector's transfers scatter / gather one host buffer: every rank moves its own block on its own
// stream (N copy engines, N PCIe links), `len` counting logical elements from the start of the vector.
static int multi_transfer(lbk_vec *v, const char *host_c, char *host, int64_t len, bool up)
{
    lbk_ctx *m = v->ctx;
    const int n = v->sharded ? (int)m->subs.size() : 1;
    return fan_out(m, n, [&](lbk_ctx *, int r) {
        lbk_vec *p = v->parts[r];
        const i64 lo = p->shard_off, cnt = std::max<i64>(0, std::min<i64>(len, lo + p->len) - lo);
        if (cnt <= 0) return (int)LBK_OK;
        return up ? lbk_vec_upload_async(p, host_c + lo * esz(v->dtype), cnt) : lbk_vec_download_async(p, host + lo * esz(v->dtype), cnt);
    });
}

extern "C" int lbk_vec_upload_async(lbk_vec *v, const void *host, int64_t len)
{
    if (!v || len < 0 || len > (is_multi(v->ctx) ? v->global_len : v->len) || (len > 0 && !host)) return fail(v ? v->ctx : nullptr, LBK_ERR_ARG, "lbk_vec_upload: bad argument");
    LIVE(v->ctx, "lbk_vec_upload");
    if (is_multi(v->ctx)) return multi_transfer(v, (const char *)host, nullptr, len, true);
    ON_DEVICE(v->ctx);
    if (len) { lbk_pdl::interrupt(v->ctx->pdl); CU(v->ctx, cudaMemcpyAsync(v->ptr, host, (size_t)esz(v->dtype) * (size_t)len, cudaMemcpyHostToDevice, v->ctx->stream)); }
    return LBK_OK;
}
// The synchronous transfers (what the reference-shaped `toDevice` / `fromDevice` of the host mirrors call) from or to
// PAGEABLE host memory -- a plain Data.Vector.Storable / numpy / std::vector buffer.  cudaMemcpy stages such a copy through
// the driver's own bounce buffer on one thread: 7-11 GB/s against 53-55 GB/s from pinned memory (bench.py `transfers`).
// Here a few host threads ("lanes") each own a pinned chunk: a lane copies its chunk of the caller's buffer into pinned memory
// while the other lanes' chunks are in flight on the copy engine (uploads), or copies a landed chunk out while the next ones
// arrive (downloads).  Pinned, registered and small buffers take the direct path.
constexpr int kMaxStageLanes = 8;
constexpr size_t kStageMinBytes = (size_t)16 << 20;
// Lanes and chunk size (tools/transfer_sweep.py, profiles/r02_transfers.md: one lane moves ~10 GB/s, 6 lanes x 8 MiB reach 47 GB/s up /
// 43 down of the 55 GB/s a pinned buffer gets; the driver's own pageable path: 10.7).  Default: 6 lanes, fewer when the ranks of a
// multi-device context (each with its own lanes) would oversubscribe the host's cores; LBK_STAGE_LANES / LBK_STAGE_CHUNK_MB override.
static int stage_lanes(const lbk_ctx *c)
{
    static const int forced = [] { const char *e = getenv("LBK_STAGE_LANES"); return e ? std::min(std::max(atoi(e), 1), kMaxStageLanes) : 0; }();
    if (forced) return forced;
    // ranks sharing this host's cores: the ranks of the multi-device context, or the processes torchrun / mpirun started here
    static const int local_world = [] {
        for (const char *name : {"LOCAL_WORLD_SIZE", "OMPI_COMM_WORLD_LOCAL_SIZE", "MPI_LOCALNRANKS"})
            if (const char *e = getenv(name)) { const int v = atoi(e); if (v > 0) return v; }
        return 1;
    }();
    const int ranks = std::max(c->parent ? (int)c->parent->subs.size() : 1, local_world);
    const int cores = (int)std::max(1u, std::thread::hardware_concurrency());
    return std::min(6, std::max(1, cores / (2 * ranks)));
}
static size_t stage_chunk_bytes()
{
    static const size_t bytes = [] { const char *e = getenv("LBK_STAGE_CHUNK_MB"); int v = e ? atoi(e) : 8; return (size_t)std::min(std::max(v, 1), 64) << 20; }();
    return bytes;
}

// the CPU half of a staged transfer: non-temporal stores (lbk_hostcopy.cpp); LBK_STAGE_MEMCPY=1 puts plain memcpy back (A/B)
extern "C" void lbk_host_stream_copy(void *dst, const void *src, size_t bytes);
static void stage_copy(void *dst, const void *src, size_t bytes)
{
    static const bool plain = [] { const char *e = getenv("LBK_STAGE_MEMCPY"); return e && atoi(e) != 0; }();
    if (plain) memcpy(dst, src, bytes);
    else lbk_host_stream_copy(dst, src, bytes);
}

static bool is_pageable(const void *host)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, host) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}

static int staged_transfer(lbk_ctx *c, char *dev, char *host, size_t bytes, bool up)
{
    ON_DEVICE(c);
    const int kStageLanes = stage_lanes(c);
    const size_t kStageChunk = stage_chunk_bytes();
    c->stage_chunk = kStageChunk;
    for (int lane = 0; lane < kStageLanes; ++lane)
        for (int i : {lane, lane + kMaxStageLanes}) {
            if (!c->stage_buf[i]) CU(c, cudaHostAlloc(&c->stage_buf[i], kStageChunk, cudaHostAllocDefault));
            if (!c->stage_ev[i]) CU(c, cudaEventCreateWithFlags(&c->stage_ev[i], cudaEventDisableTiming));
        }
    lbk_pdl::interrupt(c->pdl);
    const size_t nchunks = (bytes + kStageChunk - 1) / kStageChunk;
    std::atomic<int> failed{0};
    // Lane t moves chunks t, t + lanes, t + 2 lanes, ... through its two pinned buffers in turn: while the copy engine works on one,
    // the lane's thread copies the caller's bytes into / out of the other (non-temporal stores, lbk_hostcopy.cpp).
    auto lane = [&](int t) {
        if (cudaSetDevice(c->device) != cudaSuccess) { failed.store(1); return; }
        auto chunk = [&](size_t k, size_t *off, size_t *len) { *off = ((size_t)t + k * (size_t)kStageLanes) * kStageChunk; *len = *off < bytes ? std::min(kStageChunk, bytes - *off) : 0; };
        auto slot = [&](size_t k) { return t + (int)(k & 1) * kMaxStageLanes; };
        size_t off, len;
        cudaError_t e = cudaSuccess;
        if (up) {
            for (size_t k = 0; e == cudaSuccess && !failed.load(std::memory_order_relaxed); ++k) {
                chunk(k, &off, &len);
                if (!len) break;
                const int sl = slot(k);
                if (k >= 2) e = cudaEventSynchronize(c->stage_ev[sl]);          // chunk k-2 has left this buffer
                if (e == cudaSuccess) {
                    stage_copy(c->stage_buf[sl], host + off, len);
                    e = cudaMemcpyAsync(dev + off, c->stage_buf[sl], len, cudaMemcpyHostToDevice, c->stream);
                }
                if (e == cudaSuccess) e = cudaEventRecord(c->stage_ev[sl], c->stream);
            }
        } else {
            auto fetch = [&](size_t k) {                                       // enqueue the download of chunk k into its buffer
                size_t o, l;
                chunk(k, &o, &l);
                if (!l) return cudaSuccess;
                cudaError_t r = cudaMemcpyAsync(c->stage_buf[slot(k)], dev + o, l, cudaMemcpyDeviceToHost, c->stream);
                return r == cudaSuccess ? cudaEventRecord(c->stage_ev[slot(k)], c->stream) : r;
            };
            e = fetch(0);
            for (size_t k = 0; e == cudaSuccess && !failed.load(std::memory_order_relaxed); ++k) {
                chunk(k, &off, &len);
                if (!len) break;
                e = fetch(k + 1);                                              // the next chunk travels while this one is copied out
                if (e == cudaSuccess) e = cudaEventSynchronize(c->stage_ev[slot(k)]);
                if (e == cudaSuccess) stage_copy(host + off, c->stage_buf[slot(k)], len);
            }
        }
        if (e != cudaSuccess) failed.store((int)e);
    };
    (void)nchunks;
    std::thread workers[kMaxStageLanes - 1];
    for (int t = 1; t < kStageLanes; ++t) workers[t - 1] = std::thread(lane, t);
    lane(0);
    for (int t = 1; t < kStageLanes; ++t) workers[t - 1].join();
    if (int e = failed.load()) { cudaGetLastError(); return fail(c, LBK_ERR_CUDA, std::string("staged transfer: ") + cudaGetErrorString((cudaError_t)e)); }
    return LBK_OK;
}

// One rank's synchronous transfer: staged when the host side is large and pageable, direct otherwise.
static int sync_transfer_one(lbk_vec *v, char *host, int64_t len, bool up)
{
    lbk_ctx *c = v->ctx;
    NOT_CAPTURING(c, up ? "lbk_vec_upload" : "lbk_vec_download");
    const size_t bytes = (size_t)esz(v->dtype) * (size_t)len;
    int s;
    if (bytes >= kStageMinBytes && is_pageable(host)) s = staged_transfer(c, v->ptr, host, bytes, up);
    else s = up ? lbk_vec_upload_async(v, host, len) : lbk_vec_download_async(v, host, len);
    return s ? s : lbk_sync(c);
}

static int sync_transfer(lbk_vec *v, char *host, int64_t len, bool up, const char *who)
{
    if (!v || len < 0 || len > (is_multi(v->ctx) ? v->global_len : v->len) || (len > 0 && !host)) return fail(v ? v->ctx : nullptr, LBK_ERR_ARG, std::string(who) + ": bad argument");
    LIVE(v->ctx, who);
    if (!is_multi(v->ctx)) return sync_transfer_one(v, host, len, up);
    lbk_ctx *m = v->ctx;      // every rank moves its own block of the one host buffer, on its own stream (and staging lanes)
    return fan_out(m, v->sharded ? (int)m->subs.size() : 1, [&](lbk_ctx *, int r) {
        lbk_vec *p = v->parts[r];
        const i64 lo = p->shard_off, cnt = std::max<i64>(0, std::min<i64>(len, lo + p->len) - lo);
        return cnt > 0 ? sync_transfer_one(p, host + lo * esz(v->dtype), cnt, up) : (int)LBK_OK;
    });
}

extern "C" int lbk_vec_upload(lbk_vec *v, const void *host, int64_t len)
{
    return sync_transfer(v, const_cast<char *>(static_cast<const char *>(host)), len, true, "lbk_vec_upload");
}
extern "C" int lbk_vec_download_async(const lbk_vec *v, void *host, int64_t len)
{
    if (!v || len < 0 || len > (is_multi(v->ctx) ? v->global_len : v->len) || (len > 0 && !host)) return fail(v ? v->ctx : nullptr, LBK_ERR_ARG, "lbk_vec_download: bad argument");
    LIVE(v->ctx, "lbk_vec_download");
    if (is_multi(v->ctx)) return multi_transfer(const_cast<lbk_vec *>(v), nullptr, (char *)host, len, false);
    ON_DEVICE(v->ctx);
    if (len) { lbk_pdl::interrupt(v->ctx->pdl); CU(v->ctx, cudaMemcpyAsync(host, v->ptr, (size_t)esz(v->dtype) * (size_t)len, cudaMemcpyDeviceToHost, v->ctx->stream)); }
    return LBK_OK;
}
extern "C" int lbk_vec_download(const lbk_vec *v, void *host, int64_t len)
{
    return sync_transfer(const_cast<lbk_vec *>(v), static_cast<char *>(host), len, false, "lbk_vec_download");
}

extern "C" int lbk_vec_clone(const lbk_vec *v, lbk_vec **out)
{
    if (!v || !out) return fail(v ? v->ctx : nullptr, LBK_ERR_ARG, "lbk_vec_clone: null argument");
    LIVE(v->ctx, "lbk_vec_clone");
    if (is_multi(v->ctx)) {
        lbk_ctx *m = v->ctx;
        std::vector<lbk_vec *> parts(v->parts.size(), nullptr);
        int s = fan_out(m, (int)parts.size(), [&](lbk_ctx *, int r) { return lbk_vec_clone(v->parts[r], &parts[r]); });
        if (s) { for (lbk_vec *p : parts) lbk_vec_free(p); return s; }
        return make_multi_vec(m, std::move(parts), v->global_len, v->sharded, v->dtype, true, out);
    }
    lbk_vec *w = nullptr;
    int s = alloc_common(v->ctx, "lbk_vec_clone", v->len, v->global_len, v->shard_off, v->sharded, v->dtype, &w);
    if (s) return s;
    if (v->len) {
        s = device_copy(v->ctx, w->ptr, v->ptr, v->len, v->dtype);
        if (s) { lbk_vec_free(w); return s; }
    }
    *out = w;
    return LBK_OK;
}

static int stream_grid(lbk_ctx *c, i64 n, int threads) { return (int)std::max<i64>(1, std::min<i64>((n + threads - 1) / threads, (i64)c->sms * 16)); }

extern "C" int lbk_vec_fill_hash(lbk_vec *v, uint64_t seed, double lo, double hi)
{
    if (!v) return fail(nullptr, LBK_ERR_ARG, "lbk_vec_fill_hash: null vector");
    lbk_ctx *c = v->ctx;
    LIVE(c, "lbk_vec_fill_hash");
    if (is_multi(c)) return fan_out(c, (int)v->parts.size(), [&](lbk_ctx *, int r) { return lbk_vec_fill_hash(v->parts[r], seed, lo, hi); });
    if (!v->len) return LBK_OK;
    ON_DEVICE(c);
    lbk_pdl::interrupt(c->pdl);
    const int grid = stream_grid(c, v->len, 256);
    if (v->dtype) lbk::fill_hash_kernel<double><<<grid, 256, 0, c->stream>>>((double *)v->ptr, v->len, v->shard_off, seed, lo, hi - lo);
    else lbk::fill_hash_kernel<float><<<grid, 256, 0, c->stream>>>((float *)v->ptr, v->len, v->shard_off, seed, (float)lo, (float)hi - (float)lo);
    KERNEL_CHECK(c);
    return LBK_OK;
}

// Page-locked host memory for the transfers above: lbk_host_alloc makes some, lbk_host_register pins memory the
// caller already owns (a Data.Vector.Storable buffer, a numpy array) for the lifetime of the registration, so an
// upload from it runs at the rate of a pinned copy instead of being staged through the driver's bounce buffer.
extern "C" int lbk_host_alloc(int64_t nbytes, void **host)
{
    if (!host || nbytes < 0) return fail(nullptr, LBK_ERR_ARG, "lbk_host_alloc: bad argument");
    *host = nullptr;
    if (nbytes == 0) return LBK_OK;
    cudaError_t e = cudaHostAlloc(host, (size_t)nbytes, cudaHostAllocPortable);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(nullptr, LBK_ERR_NOMEM, std::string("cudaHostAlloc: ") + cudaGetErrorString(e)); }
    return LBK_OK;
}
extern "C" int lbk_host_free(void *host)
{
    if (host) CU(nullptr, cudaFreeHost(host));
    return LBK_OK;
}
extern "C" int lbk_host_register(void *host, int64_t nbytes)
{
    if (!host || nbytes <= 0) return fail(nullptr, LBK_ERR_ARG, "lbk_host_register: bad argument");
    cudaError_t e = cudaHostRegister(host, (size_t)nbytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(nullptr, e == cudaErrorMemoryAllocation ? LBK_ERR_NOMEM : LBK_ERR_CUDA, std::string("cudaHostRegister: ") + cudaGetErrorString(e)); }
    return LBK_OK;
}
extern "C" int lbk_host_unregister(void *host)
{
    if (!host) return fail(nullptr, LBK_ERR_ARG, "lbk_host_unregister: null pointer");
    CU(nullptr, cudaHostUnregister(host));
    return LBK_OK;
}

// ------------------------------------------------------------------------------------------------
// events
// ------------------------------------------------------------------------------------------------
extern "C" int lbk_event_create(lbk_ctx *c, lbk_event **out)
{
    if (!c || !out) return fail(c, LBK_ERR_ARG, "lbk_event_create: null argument");
    LIVE(c, "lbk_event_create");
    lbk_event *e = new (std::nothrow) lbk_event{c, nullptr, {}};
    if (!e) return fail(c, LBK_ERR_NOMEM, "out of host memory");
    if (is_multi(c)) {         // one event per rank, each on its own device's stream
        for (lbk_ctx *s : c->subs) {
            lbk_event *pe = nullptr;
            int st = lbk_event_create(s, &pe);
            if (st) { c->err = s->err; for (lbk_event *q : e->parts) lbk_event_destroy(q); delete e; return st; }
            e->parts.push_back(pe);
        }
    } else {
        DevGuard guard(c->device);
        cudaError_t ce = guard.err == cudaSuccess ? cudaEventCreate(&e->ev) : guard.err;
        if (ce != cudaSuccess) { delete e; return fail(c, LBK_ERR_CUDA, std::string("cudaEventCreate: ") + cudaGetErrorString(ce)); }
    }
    ctx_ref(c);
    *out = e;
    return LBK_OK;
}
extern "C" int lbk_event_record(lbk_event *e)
{
    if (!e) return fail(nullptr, LBK_ERR_ARG, "lbk_event_record: null event");
    LIVE(e->ctx, "lbk_event_record");
    if (!e->parts.empty()) return all_ranks(e->ctx, [&](lbk_ctx *, int r) { return lbk_event_record(e->parts[r]); });
    NOT_CAPTURING(e->ctx, "lbk_event_record");
    ON_DEVICE(e->ctx);
    lbk_pdl::interrupt(e->ctx->pdl);
    CU(e->ctx, cudaEventRecord(e->ev, e->ctx->stream));
    return LBK_OK;
}
// Everything enqueued on `c` after this call waits for `e`.  With multi-device contexts rank r waits for the
// event's rank-r part (both sides must then be multi-device contexts over the same devices).
extern "C" int lbk_wait_event(lbk_ctx *c, lbk_event *e)
{
    if (!c || !e) return fail(c, LBK_ERR_ARG, "lbk_wait_event: null argument");
    LIVE(c, "lbk_wait_event");
    if (is_multi(c) != !e->parts.empty() || (is_multi(c) && c->subs.size() != e->parts.size()))
        return fail(c, LBK_ERR_ARG, "lbk_wait_event: context and event must both be single-device, or multi-device over the same ranks");
    if (is_multi(c)) return all_ranks(c, [&](lbk_ctx *s, int r) { return lbk_wait_event(s, e->parts[r]); });
    NOT_CAPTURING(c, "lbk_wait_event");
    ON_DEVICE(c);
    lbk_pdl::interrupt(c->pdl);
    CU(c, cudaStreamWaitEvent(c->stream, e->ev, 0));
    return LBK_OK;
}
// Multi-device events: the LONGEST of the ranks' own device times (every rank's pair brackets the same calls).
extern "C" int lbk_event_elapsed_ms(lbk_event *a, lbk_event *b, float *ms)
{
    if (!a || !b || !ms) return fail(nullptr, LBK_ERR_ARG, "lbk_event_elapsed_ms: null argument");
    if (a->parts.size() != b->parts.size()) return fail(b->ctx, LBK_ERR_ARG, "lbk_event_elapsed_ms: events of different contexts");
    if (!a->parts.empty()) {
        float worst = 0.f;
        for (size_t r = 0; r < a->parts.size(); ++r) {
            float t = 0.f;
            int s = lbk_event_elapsed_ms(a->parts[r], b->parts[r], &t);
            if (s) { b->ctx->err = b->parts[r]->ctx->err; return s; }
            worst = std::max(worst, t);
        }
        *ms = worst;
        return LBK_OK;
    }
    ON_DEVICE(b->ctx);
    CU(b->ctx, cudaEventSynchronize(b->ev));
    CU(b->ctx, cudaEventElapsedTime(ms, a->ev, b->ev));
    return LBK_OK;
}
extern "C" int lbk_event_destroy(lbk_event *e)
{
    if (!e) return LBK_OK;
    lbk_ctx *c = e->ctx;
    for (lbk_event *p : e->parts) lbk_event_destroy(p);
    if (e->ev) cudaEventDestroy(e->ev);
    delete e;
    ctx_unref(c);
    return LBK_OK;
}

// ------------------------------------------------------------------------------------------------
// argument checks shared by all routines
// ------------------------------------------------------------------------------------------------
struct Operand {
    const lbk_vec *v;
    i64 inc;
    i64 n_local;   // logical elements of this call that live on this rank
    i64 i0;        // first logical element on this rank (0 unless sharded)
    i64 base;      // where that element sits in this rank's buffer (0 unless sharded with an increment > 1)
    const char *ptr() const { return v->ptr + base * (v->dtype ? 8 : 4); }
};

// Which logical elements i in [0, n), sitting at global positions i*inc (inc >= 1), fall into the block
// [lo, hi) of positions: i0 = the first one, count of them, and where the first one sits inside the block.
static void shard_span(i64 lo, i64 hi, i64 n, i64 inc, i64 *i0, i64 *count, i64 *base)
{
    const i64 i_lo = (lo + inc - 1) / inc, i_hi = std::min<i64>(n, (hi + inc - 1) / inc);
    *count = i_hi > i_lo ? i_hi - i_lo : 0;
    *i0 = i_lo;
    *base = *count > 0 ? i_lo * inc - lo : 0;
}

extern "C" int lbk_shard_span(int64_t global_len, int nranks, int rank, int64_t n, int64_t inc,
                              int64_t *first, int64_t *count, int64_t *offset)
{
    int64_t lo, hi;
    int s = lbk_shard_range(global_len, nranks, rank, &lo, &hi);
    if (s) return s;
    if (n < 0 || inc < 1 || !first || !count || !offset) return fail(nullptr, LBK_ERR_ARG, "lbk_shard_span: bad argument");
    if (n > 0 && span_exceeds(n, inc, global_len)) return fail(nullptr, LBK_ERR_RANGE, "lbk_shard_span: 1+(n-1)*inc exceeds the vector");
    i64 a = 0, b = 0, c2 = 0;
    if (n > 0) shard_span(lo, hi, n, inc, &a, &b, &c2);
    *first = a; *count = b; *offset = c2;
    return LBK_OK;
}

// Validates one operand for a call over n logical elements and resolves sharding.
static int check_operand(lbk_ctx *c, const char *who, const lbk_vec *v, int dtype, i64 inc, i64 n, Operand *op)
{
    if (!v) return fail(c, LBK_ERR_ARG, std::string(who) + ": null vector");
    if (v->ctx != c) return fail(c, LBK_ERR_ARG, std::string(who) + ": vector belongs to another context");
    if (v->dtype != dtype) return fail(c, LBK_ERR_ARG, std::string(who) + (dtype ? ": needs a double-precision vector" : ": needs a single-precision vector"));
    LIVE(c, who);
    op->v = v; op->inc = inc; op->i0 = 0; op->n_local = n; op->base = 0;
    if (n <= 0) { op->n_local = 0; return LBK_OK; }
    if (v->sharded) {
        // Logical element i sits at global position i*inc; this rank holds positions [lo, hi).  Positive increments
        // keep every pairing on one rank as long as both operands are partitioned alike (checked by the caller);
        // negative ones pair x[i] with y[n-1-i] across devices and zero replicates one rank's element: refused.
        if (inc < 1) return fail(c, LBK_ERR_UNSUPPORTED, std::string(who) + ": sharded vectors take equal increments, positive or both negative (not zero, not opposite signs)");
        if (span_exceeds(n, inc, v->global_len)) return fail(c, LBK_ERR_RANGE, std::string(who) + ": 1+(n-1)*inc exceeds the vector");
        shard_span(v->shard_off, v->shard_off + v->len, n, n == 1 ? 1 : inc, &op->i0, &op->n_local, &op->base);   // n == 1: the increment is never applied
        return LBK_OK;
    }
    if (span_exceeds(n, inc, v->len)) return fail(c, LBK_ERR_RANGE, std::string(who) + ": 1+(n-1)|inc| exceeds the vector");
    return LBK_OK;
}

// A call on a multi-device context: its vectors must all be that context's, and all sharded (the call runs on
// every rank) or all whole (it runs on rank 0, with the full increment semantics of a single device).
static int multi_count(lbk_ctx *m, const char *who, std::initializer_list<const lbk_vec *> vs, int *count)
{
    LIVE(m, who);
    int sharded = 0, whole = 0;
    for (const lbk_vec *v : vs) {
        if (!v) continue;
        if (v->ctx != m) return fail(m, LBK_ERR_ARG, std::string(who) + ": vector belongs to another context");
        (v->sharded ? sharded : whole)++;
    }
    if (sharded && whole) return fail(m, LBK_ERR_UNSUPPORTED, std::string(who) + ": sharded and unsharded vectors cannot be mixed in one call");
    *count = sharded ? (int)m->subs.size() : 1;
    return LBK_OK;
}

// touched byte range [lo, hi) of an operand
static void touched(const Operand &o, const char **lo, const char **hi)
{
    const i64 n = o.n_local;
    *lo = o.ptr();
    *hi = o.ptr() + (n > 0 ? 1 + (n - 1) * iabs64(o.inc) : 0) * esz(o.v->dtype);
}
template <class T> static inline int misalign(const T *p, int vec) { return (int)(((uintptr_t)p / sizeof(T)) % (uintptr_t)vec); }
// elements per access of a variant for element type T: the table is in floats (4/16/32 bytes)
template <class T> static inline int variant_vec(int variant)
{
    const int f = kVariantVec[(variant >= 0 && variant < kNumVariants) ? variant : 0];
    return sizeof(T) == 4 ? f : (f > 1 ? f / 2 : 1);
}

// Picks the variant that the pointers' alignment allows and the scalar head count.
template <class T>
static bool resolve_alignment(int *variant, i64 *head, const T *const *ptrs, int nptr, i64 n, int scalar_variant)
{
    for (;;) {
        const int vec = variant_vec<T>(*variant);
        if (vec == 1) {
            // 4- or 8-byte accesses need no alignment, but a warp's 128 contiguous bytes should not straddle two 128-byte
            // lines on the WRITTEN side: peel up to a line of scalar elements so that the last pointer (a destination
            // whenever the routine has one) starts on a line boundary (profiles/r02_misaligned.md)
            const i64 per_line = 128 / (i64)sizeof(T);
            const i64 m = (i64)(((uintptr_t)ptrs[nptr - 1] / sizeof(T)) % (uintptr_t)per_line);
            *head = std::min<i64>((per_line - m) % per_line, n);
            return *variant == scalar_variant;
        }
        const int m = misalign(ptrs[0], vec);
        bool same = true;
        for (int i = 1; i < nptr; ++i) same = same && misalign(ptrs[i], vec) == m;
        if (same) { *head = std::min<i64>((vec - m) % vec, n); return false; }
        *variant = (kVariantVec[*variant] == 8) ? kVec4Fallback : scalar_variant;
    }
}
// Reversed pairing: x+head and (y+n)-head must both be pack aligned; otherwise the scalar shape is used.
template <class T>
static void resolve_alignment_rev(int *variant, i64 *head, const T *const *xptrs, int nx, const T *const *yptrs, int ny, i64 n)
{
    const int vec = variant_vec<T>(kVec4Fallback);
    const int m = misalign(xptrs[0], vec);
    const i64 h = (vec - m) % vec;          // x + h is aligned; y + n - h must be too  <=>  misalign(y + n) == h
    bool ok = h <= n;
    for (int i = 0; i < nx; ++i) ok = ok && misalign(xptrs[i], vec) == m;
    for (int i = 0; i < ny; ++i) ok = ok && misalign(yptrs[i] + n, vec) == (int)h;
    if (ok) { *variant = kVec4Fallback; *head = h; }      // 16-byte packs, UNROLL 4
    else { *variant = kScalarVariant; *head = 0; }
}

// ------------------------------------------------------------------------------------------------
// map launchers
// ------------------------------------------------------------------------------------------------
// the reversed pairing is built for two shapes only: 16-byte packs x UNROLL 4, and the any-alignment scalar one
template <class F>
static const void *map_kernel_ptr_rev(int variant)
{
    using T = typename F::Scalar;
    return variant == kScalarVariant ? (const void *)lbk::map_unit_kernel<F, 1, 4, 1, true>
                                     : (const void *)lbk::map_unit_kernel<F, vec_t<T>(4), 4, 1, true>;
}
// shifted loads of x (lbk_kernels.cuh, shift_pack): 16-byte packs, unroll 2 or 4, shift 1..VEC-1 elements
template <class F, int UNROLL>
static const void *map_kernel_ptr_shifted_u(int shift)
{
    using T = typename F::Scalar;
    if constexpr (F::XSHIFT) {
        constexpr int V = vec_t<T>(4);
        if (shift == 1) return (const void *)lbk::map_unit_kernel<F, V, UNROLL, 6, false, 1>;
        if constexpr (V == 4) {
            if (shift == 2) return (const void *)lbk::map_unit_kernel<F, V, UNROLL, 6, false, 2>;
            if (shift == 3) return (const void *)lbk::map_unit_kernel<F, V, UNROLL, 6, false, 3>;
        }
    }
    return nullptr;
}
template <class F>
static const void *map_kernel_ptr_shifted(int unroll, int shift)
{
    return unroll == 2 ? map_kernel_ptr_shifted_u<F, 2>(shift) : unroll == 4 ? map_kernel_ptr_shifted_u<F, 4>(shift) : nullptr;
}
template <class Op, int UNROLL>
static const void *reduce_kernel_ptr_shifted_u(int shift)
{
    using T = typename Op::Scalar;
    if constexpr (Op::TWO) {
        constexpr int V = vec_t<T>(4);
        if (shift == 1) return (const void *)lbk::reduce_unit_kernel<Op, V, UNROLL, 1, false, 1>;
        if constexpr (V == 4) {
            if (shift == 2) return (const void *)lbk::reduce_unit_kernel<Op, V, UNROLL, 1, false, 2>;
            if (shift == 3) return (const void *)lbk::reduce_unit_kernel<Op, V, UNROLL, 1, false, 3>;
        }
    }
    return nullptr;
}
template <class Op>
static const void *reduce_kernel_ptr_shifted(int unroll, int shift)
{
    return unroll == 2 ? reduce_kernel_ptr_shifted_u<Op, 2>(shift) : unroll == 4 ? reduce_kernel_ptr_shifted_u<Op, 4>(shift) : nullptr;
}
// x is the only pointer whose offset inside a 16-byte pack differs from the others' (which agree): the body is laid out on
// the others' boundaries and x is read through shifted loads.  Returns the shift (0: not this case) and the head to peel.
template <class T>
static int shifted_layout(const T *x, const T *const *others, int nothers, i64 n, i64 *head)
{
    const int vec = variant_vec<T>(kVec4Fallback);
    if (nothers < 1 || n < 16 * vec) return 0;
    const int m = misalign(others[0], vec);
    for (int i = 1; i < nothers; ++i) if (misalign(others[i], vec) != m) return 0;
    i64 h = (vec - m) % vec;
    const int shift = (int)((misalign(x, vec) + h) % vec);
    if (shift == 0) return 0;
    if (h < shift) h += vec;            // the first aligned pack of x that is touched starts inside the vector
    *head = h;
    return shift;
}

static int variant_unroll(int variant)
{
    switch (variant) {
#define X(ID, VEC, UNROLL, POLICY) case ID: return UNROLL;
        LBK_VARIANTS(X)
#undef X
    }
    return 1;
}

// adapt_tile_bytes > 0: a reduction's grid-stride loop wants 384 KiB per operand per block (24 tiles of 16 KiB).  The default
// 16 (8) blocks per SM were tuned at n = 2^28, where a block streams 453 KiB; at the sizes a rank sees when n_global = 2^28 is
// sharded over 8 GPUs (2^25) the same grid leaves 3.5 tiles per block and a ragged, half-empty last wave: sasum 22.6 -> 18.3 us,
// isamax 26.3 -> 18.4 us with a quarter of the blocks (profiles/r02_midsize_launch.md).  So the blocks per SM are halved, down to
// 1024 threads per SM, until a block has its 384 KiB (LBK_TILES_PER_BLOCK = that figure in 16-KiB tiles, for the sweeps).
// 1024 threads per SM also leave room for the NEXT kernel's first blocks while this one is still resident, which is what hides a
// sharded reduction's record exchange under its successor: 2 GPUs, 2^26 per rank, isamax 41.2 -> 36.5 us (36.4 without any
// exchange).  The measure is bytes, not tiles: sdot's tiles are twice sasum's, and shrinking ITS grid at 2^28 cost the bench step 1.1 %.
static int grid_for(lbk_ctx *c, const void *kernel, const LaunchCfg &cfg, int threads, i64 ntiles, int *grid, i64 max_grid = kMaxGrid,
                    i64 adapt_tile_bytes = 0)
{
    int bps = cfg.blocks_per_sm;
    if (bps <= 0) {
        CU(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kernel, threads, 0));
        if (bps < 1) bps = 1;
    } else if (adapt_tile_bytes > 0 && bps < kOneTilePerBlock) {
        const int floor_bps = std::max(1, 1024 / threads);          // 1024 threads per SM
        static const i64 want_bytes = [] { const char *e = getenv("LBK_TILES_PER_BLOCK"); const int v = e ? atoi(e) : 0; return (i64)(v > 0 ? v : 24) << 14; }();
        while (bps > floor_bps && ntiles * adapt_tile_bytes < want_bytes * c->sms * bps) bps /= 2;
        if (bps < floor_bps) bps = floor_bps;
    }
    i64 g = std::min<i64>(std::max<i64>(ntiles, 1), (i64)c->sms * bps);
    *grid = (int)std::min<i64>(g, max_grid);
    return LBK_OK;
}

// Launches a unit-stride kernel the way lbk_pdl planned it and records it in the group of kernels in flight.
static int launch_stream_kernel(lbk_ctx *c, const void *k, int grid, int threads, void **args, const lbk_pdl::Op &op,
                                const lbk_pdl::Plan &plan)
{
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(grid); lc.blockDim = dim3(threads); lc.dynamicSmemBytes = 0; lc.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at; lc.numAttrs = plan.mode != lbk_pdl::NONE ? 1 : 0;
    lbk_pdl::interrupt(c->pdl);                          // stays closed if the launch fails
    CU(c, cudaLaunchKernelExC(&lc, k, args));
    if (plan.mode != lbk_pdl::NONE) c->pdl_launches++;
    KERNEL_CHECK(c);
    lbk_pdl::commit(c->pdl, op, plan);
    return LBK_OK;
}

template <class F>
static int launch_map_unit(lbk_ctx *c, int rid, const F &f, const typename F::Scalar *xr, const typename F::Scalar *yr,
                           typename F::Scalar *xw, typename F::Scalar *yw, i64 n, bool rev = false)
{
    using T = typename F::Scalar;
    const LaunchCfg &cfg = c->cfg[rid];
    const LaunchCfg *pcfg = &cfg;
    int variant = cfg.variant;
    // the reversed pairing has one shape: 16-byte packs x 4, 128 threads, no fence (config-4 sweep, profiles/)
    int threads = rev ? 128 : (cfg.threads > 0 ? cfg.threads : 256);
    i64 head = 0;
    int xshift = 0;             // > 0: x is read through shifted loads (lbk_kernels.cuh, shift_pack)
    const void *k;
    if (rev) {
        const T *xp[2], *yp[2]; int nx = 0, ny = 0;
        if (F::RX) xp[nx++] = xr;
        if (F::WX) xp[nx++] = xw;
        if (F::RY) yp[ny++] = yr;
        if (F::WY) yp[ny++] = yw;
        resolve_alignment_rev<T>(&variant, &head, xp, nx, yp, ny, n);
        k = map_kernel_ptr_rev<F>(variant);
    } else {
        const T *ptrs[4]; int np = 0;
        if (F::RX) ptrs[np++] = xr;
        if (F::RY) ptrs[np++] = yr;
        if (F::WX) ptrs[np++] = xw;
        if (F::WY) ptrs[np++] = yw;
        // x and y misaligned differently: 8 accesses of one element in flight per operand, or 4 for the routines that write
        // both vectors; 256 threads (profiles/r02_misaligned.md)
        LaunchCfg mis = c->cfg[R_MAPMISALIGNED];
        if (mis.variant == kDefaultCfg[R_MAPMISALIGNED].variant && F::WX && F::WY) mis.variant = kScalarVariant;
        const bool aligned = !resolve_alignment<T>(&variant, &head, ptrs, np, n, kVariantVec[mis.variant] == 1 ? mis.variant : kScalarVariant);
        if (!aligned && variant == mis.variant) {
            pcfg = &c->cfg[R_MAPMISALIGNED];
            if (mis.threads > 0) threads = mis.threads;
        }
        k = map_kernel_ptr<F>(variant);
        if constexpr (F::XSHIFT) {
            const LaunchCfg &sh = c->cfg[R_MAPSHIFTED];
            if (variant_vec<T>(variant) == 1 && kVariantVec[sh.variant] == 4) {      // only x is the odd one out?
                i64 h = 0;
                const int shift = shifted_layout<T>(xr, ptrs + 1, np - 1, n, &h);
                const void *ks = shift ? map_kernel_ptr_shifted<F>(variant_unroll(sh.variant), shift) : nullptr;
                if (ks) { k = ks; head = h; variant = sh.variant; pcfg = &sh; threads = std::min(sh.threads > 0 ? sh.threads : 256, lbk::kShiftThreads); xshift = shift; }
            }
        }
    }
    const i64 npack = xshift ? std::max<i64>(0, (n - head - (variant_vec<T>(variant) - xshift)) / variant_vec<T>(variant))
                             : (n - head) / variant_vec<T>(variant);
    const i64 per_tile = (i64)variant_unroll(variant) * threads;
    int grid;
    int s = grid_for(c, k, *pcfg, threads, (npack + per_tile - 1) / per_tile, &grid, 0x7fffffff);   // maps keep no per-block scratch
    if (s) return s;
    F fc = f;
    unsigned fence_mask = 0;     // see map_unit_kernel: always 0, but only known at run time
    const i64 nb = n * (i64)sizeof(T);
    lbk_pdl::Op op;
    if (F::RX) { op.rd[op.nrd] = xr; op.rdn[op.nrd++] = nb; }
    if (F::RY) { op.rd[op.nrd] = yr; op.rdn[op.nrd++] = nb; }
    if (F::WX) { op.wr[op.nwr] = xw; op.wrn[op.nwr++] = nb; }
    if (F::WY) { op.wr[op.nwr] = yw; op.wrn[op.nwr++] = nb; }
    const lbk_pdl::Plan plan = lbk_pdl::plan(c->pdl, op);
    int wait_mode = plan.mode == lbk_pdl::WAIT_BEFORE_STORE ? 1 : plan.mode == lbk_pdl::WAIT_BEFORE_LOAD ? 2 : 0;
    // walk the vectors from the end the previous kernel left in the L2 (see lbk_ctx::walks)
    bool backwards = c->traversal == 2;
    if (c->traversal == 1 && nb >= kWalkMinBytes && !rev)
        backwards = (F::RX ? walk_vote(c, xr) : 0) + (F::RY ? walk_vote(c, yr) : 0) > 0;
    if (backwards) wait_mode |= 4;
    if (F::RX) walk_note(c, xr, backwards);
    if (F::RY) walk_note(c, yr, backwards);
    if (F::WX) walk_note(c, xw, backwards);
    if (F::WY) walk_note(c, yw, backwards);
    void *args[] = {(void *)&fc, (void *)&xr, (void *)&yr, (void *)&xw, (void *)&yw, (void *)&head, (void *)&n,
                    (void *)&fence_mask, (void *)&wait_mode};
    if (xshift) c->shifted_launches++;
    return launch_stream_kernel(c, k, grid, threads, args, op, plan);
}

extern "C" int lbk_vec_fill(lbk_vec *v, double value)
{
    if (!v) return fail(nullptr, LBK_ERR_ARG, "lbk_vec_fill: null vector");
    lbk_ctx *c = v->ctx;
    LIVE(c, "lbk_vec_fill");
    if (is_multi(c)) return fan_out(c, (int)v->parts.size(), [&](lbk_ctx *, int r) { return lbk_vec_fill(v->parts[r], value); });
    if (!v->len) return LBK_OK;
    ON_DEVICE(c);
    if (v->dtype) return launch_map_unit(c, R_FILL, lbk::FillF<double>{value}, (const double *)nullptr, (const double *)nullptr, (double *)v->ptr, (double *)nullptr, v->len);
    return launch_map_unit(c, R_FILL, lbk::FillF<float>{(float)value}, (const float *)nullptr, (const float *)nullptr, (float *)v->ptr, (float *)nullptr, v->len);
}

static int device_copy(lbk_ctx *c, void *dst, const void *src, i64 count, int dtype)
{
    if (count <= 0) return LBK_OK;
    ON_DEVICE(c);
    if (dtype) return launch_map_unit(c, R_COPY, lbk::CopyF<double>{}, (const double *)src, (const double *)nullptr, (double *)nullptr, (double *)dst, count);
    return launch_map_unit(c, R_COPY, lbk::CopyF<float>{}, (const float *)src, (const float *)nullptr, (float *)nullptr, (float *)dst, count);
}

template <class F>
static int launch_map_strided(lbk_ctx *c, const F &f, const typename F::Scalar *xr, typename F::Scalar *xw, i64 incx_r, i64 incx_w,
                              i64 kx0_r, i64 kx0_w, const typename F::Scalar *yr, typename F::Scalar *yw, i64 incy_r, i64 incy_w,
                              i64 ky0_r, i64 ky0_w, i64 i_begin, i64 n, int x_last, int y_last)
{
    // Launch shape (profiles/r01_strided_launch_sweep.md): with an increment of 4 or more every element
    // sits in its own 32-byte sector and one small tile per block (blocks retire front to back) is best;
    // with increments of 1-2 on either side, fewer, longer-lived blocks of 128 threads win by 5-20%.
    const LaunchCfg &cfg = c->cfg[R_MAPSTRIDED];
    const bool wide = std::max(std::max(iabs64(incx_r), iabs64(incx_w)), std::max(iabs64(incy_r), iabs64(incy_w))) >= 4;
    const int threads = (cfg.threads > 0 && cfg.threads <= 256) ? cfg.threads : (wide ? 256 : 128);
    const i64 per_tile = (i64)lbk::kStridedUnroll * threads;
    const i64 ntiles = (n - i_begin + per_tile - 1) / per_tile;
    const i64 cap = cfg.blocks_per_sm > 0 ? (i64)c->sms * cfg.blocks_per_sm : (wide ? ntiles : (i64)c->sms * 256);
    const int grid = (int)std::max<i64>(1, std::min<i64>(std::min<i64>(ntiles, cap), 0x7fffffff));
    // The general-stride kernels join the chain of overlapping kernels like the unit-stride ones (lbk_pdl.hpp); what they touch
    // is stated as the whole span of each operand, 1 + (n-1)|inc| elements from its first touched one.
    using T = typename F::Scalar;
    auto span_of = [&](const T *base, i64 k0, i64 inc, const T **lo) -> i64 {
        const i64 a = inc < 0 ? k0 + (n - 1) * inc : k0, b = inc < 0 ? k0 : k0 + (n - 1) * inc;     // lowest / highest index touched
        *lo = base + a;
        return (b - a + 1) * (i64)sizeof(T);
    };
    lbk_pdl::Op op;
    const T *lo;
    if (F::RX) { op.rdn[op.nrd] = span_of(xr, kx0_r, incx_r, &lo); op.rd[op.nrd++] = lo; }
    if (F::RY) { op.rdn[op.nrd] = span_of(yr, ky0_r, incy_r, &lo); op.rd[op.nrd++] = lo; }
    if (F::WX) { op.wrn[op.nwr] = span_of(xw, kx0_w, incx_w, &lo); op.wr[op.nwr++] = lo; }
    if (F::WY) { op.wrn[op.nwr] = span_of(yw, ky0_w, incy_w, &lo); op.wr[op.nwr++] = lo; }
    const lbk_pdl::Plan plan = lbk_pdl::plan(c->pdl, op);
    int wait_mode = plan.mode == lbk_pdl::WAIT_BEFORE_STORE ? 1 : plan.mode == lbk_pdl::WAIT_BEFORE_LOAD ? 2 : 0;
    F fc = f;
    unsigned fence_mask = 0;
    const void *k = cfg.variant == 0 ? (const void *)lbk::map_strided_kernel<F, false> : (const void *)lbk::map_strided_kernel<F, true>;
    void *args[] = {(void *)&fc, (void *)&xr, (void *)&xw, (void *)&incx_r, (void *)&incx_w, (void *)&kx0_r, (void *)&kx0_w,
                    (void *)&yr, (void *)&yw, (void *)&incy_r, (void *)&incy_w, (void *)&ky0_r, (void *)&ky0_w,
                    (void *)&i_begin, (void *)&n, (void *)&x_last, (void *)&y_last, (void *)&fence_mask, (void *)&wait_mode};
    walk_forget(c);          // which end of these vectors the L2 holds is anybody's guess now
    return launch_stream_kernel(c, k, grid, threads, args, op, plan);
}

// In-place (xw == xr, yw == yr) or out-of-place update of n logical elements.  A destination that
// differs from its source must already hold a copy of the source wherever this call does not write
// (map_entry arranges that).
template <class F>
static int map_dispatch(lbk_ctx *c, int rid, const F &f, i64 n, const Operand &X, typename F::Scalar *xw,
                        const Operand &Y, typename F::Scalar *yw)
{
    using T = typename F::Scalar;
    if (n <= 0) return LBK_OK;
    ON_DEVICE(c);
    const bool use_x = F::RX || F::WX, use_y = F::RY || F::WY;
    const T *xr = use_x ? (const T *)X.ptr() : nullptr;
    const T *yr = use_y ? (const T *)Y.ptr() : nullptr;
    if (xw) xw += X.base;       // the destinations are laid out like their sources
    if (yw) yw += Y.base;
    const i64 incx = use_x ? X.inc : 1, incy = use_y ? Y.inc : 1;
    const i64 nl = use_x ? X.n_local : Y.n_local;   // sharded operands agree (checked by the caller)
    if (nl <= 0) return LBK_OK;
    // unit stride; (-1,-1) visits the same pairs in the opposite order, which an elementwise update cannot see
    const bool unit = (incx == 1 && incy == 1) || (use_x && use_y && incx == -1 && incy == -1) ||
                      (!use_y && incx == 1) || (!use_x && incy == 1);
    if (unit) return launch_map_unit<F>(c, rid, f, xr, yr, xw, yw, nl);
    // (1,-1) and (-1,1) pair x[k] with y[n-1-k]: the same set of pairs, handled by the reversed-lane unit kernel
    if (use_x && use_y && ((incx == 1 && incy == -1) || (incx == -1 && incy == 1)))
        return launch_map_unit<F>(c, rid, f, xr, yr, xw, yw, nl, true);

    i64 kx0 = first_index(nl, incx), ky0 = first_index(nl, incy);
    i64 kx0_r = kx0, ky0_r = ky0;
    int x_last = 0, y_last = 0;
    T *snap = (T *)c->snap;
    if (F::WX && incx == 0) {            // n sequential writes to x[0]: the last wins, all read the original
        x_last = 1;
        if (F::RX) { lbk_pdl::interrupt(c->pdl); CU(c, cudaMemcpyAsync(snap, xr, sizeof(T), cudaMemcpyDeviceToDevice, c->stream)); xr = snap; kx0_r = 0; }
    }
    if (F::WY && incy == 0) {
        y_last = 1;
        if (F::RY) { lbk_pdl::interrupt(c->pdl); CU(c, cudaMemcpyAsync(snap + 1, yr, sizeof(T), cudaMemcpyDeviceToDevice, c->stream)); yr = snap + 1; ky0_r = 0; }
    }
    const bool only_last = (!F::WX || x_last) && (!F::WY || y_last);
    return launch_map_strided<F>(c, f, xr, xw, incx, incx, kx0_r, kx0, yr, yw, incy, incy, ky0_r, ky0,
                                 only_last ? nl - 1 : 0, nl, x_last, y_last);
}

// Shared front end of the elementwise routines.  xo / yo: optional out-of-place destinations.
template <class F>
static int map_entry(lbk_ctx *c, const char *who, int rid, const F &f, i64 n,
                     const lbk_vec *x, i64 incx, const lbk_vec *y, i64 incy, lbk_vec *xo, lbk_vec *yo)
{
    using T = typename F::Scalar;
    constexpr int dtype = sizeof(T) == 8 ? 1 : 0;
    if (!c) return fail(nullptr, LBK_ERR_ARG, std::string(who) + ": null context");
    const bool use_x = F::RX || F::WX, use_y = F::RY || F::WY;
    if (is_multi(c)) {
        int count, st = multi_count(c, who, {use_x ? x : nullptr, use_y ? y : nullptr, xo, yo}, &count);
        if (st) return st;
        if ((use_x && !x) || (use_y && !y)) return fail(c, LBK_ERR_ARG, std::string(who) + ": null vector");
        return fan_out(c, count, [&](lbk_ctx *sc, int r) { return map_entry<F>(sc, who, rid, f, n, part(x, r), incx, part(y, r), incy, part(xo, r), part(yo, r)); });
    }
    // Two equal NEGATIVE increments on shards pair the same elements as their absolute values do (logical i sits at position
    // (n-1-i)|inc| in BOTH vectors), and an elementwise update cannot see the order: run it forwards, every pairing on one rank.
    if (use_x && use_y && x && y && x->sharded && y->sharded && incx < 0 && incx == incy && incx != INT64_MIN) { incx = -incx; incy = -incy; }
    Operand X{}, Y{};
    int s;
    if (use_x && (s = check_operand(c, who, x, dtype, incx, n, &X))) return s;
    if (use_y && (s = check_operand(c, who, y, dtype, incy, n, &Y))) return s;
    if (use_x && use_y && (x->sharded != y->sharded || (x->sharded && (x->global_len != y->global_len || (n > 0 && incx != incy)))))
        return fail(c, LBK_ERR_UNSUPPORTED, std::string(who) + ": sharded operands need equal lengths and equal increments");
    T *xw = nullptr, *yw = nullptr;
    if (F::WX) {
        lbk_vec *d = xo ? xo : const_cast<lbk_vec *>(x);
        if (d->ctx != c || d->len != x->len || d->dtype != dtype || d->sharded != x->sharded || d->shard_off != x->shard_off)
            return fail(c, LBK_ERR_ARG, std::string(who) + ": output vector must match the input's length, type and shard layout");
        xw = (T *)d->ptr;
    }
    if (F::WY) {
        lbk_vec *d = yo ? yo : const_cast<lbk_vec *>(y);
        if (d->ctx != c || d->len != y->len || d->dtype != dtype || d->sharded != y->sharded || d->shard_off != y->shard_off)
            return fail(c, LBK_ERR_ARG, std::string(who) + ": output vector must match the input's length, type and shard layout");
        yw = (T *)d->ptr;
    }
    // A written span must not overlap any OTHER span this call touches (reads may alias each other freely,
    // and a destination may be its own source: that is the in-place form).
    if (n > 0) {
        const char *xl = nullptr, *xh = nullptr, *yl = nullptr, *yh = nullptr;
        if (use_x) touched(X, &xl, &xh);
        if (use_y) touched(Y, &yl, &yh);
        auto hits = [](const char *al, const char *ah, const char *bl, const char *bh) { return al && bl && al < bh && bl < ah; };
        const char *xwl = xw ? (const char *)(xw + X.base) : nullptr, *xwh = xwl ? xwl + (xh - xl) : nullptr;   // laid out like the sources
        const char *ywl = yw ? (const char *)(yw + Y.base) : nullptr, *ywh = ywl ? ywl + (yh - yl) : nullptr;
        bool bad = false;
        if (F::WX) bad = bad || hits(xwl, xwh, yl, yh) || (xwl != xl && hits(xwl, xwh, xl, xh)) || (F::WY && hits(xwl, xwh, ywl, ywh));
        if (F::WY) bad = bad || hits(ywl, ywh, xl, xh) || (ywl != yl && hits(ywl, ywh, yl, yh));
        if (bad) return fail(c, LBK_ERR_ALIAS, std::string(who) + ": a written span overlaps another operand");
    }
    // out of place: everything this call does not write is a copy of the source (Util.hs:134-142:
    // unsafeUpdate_ clones the whole destination first).  Fused into the update when the update covers
    // the vector; otherwise the remainder (unit stride) or the whole vector (strided) is copied.
    auto prepare = [&](const Operand &O, T *w, bool writes) -> int {
        if (!writes || (char *)w == O.v->ptr || O.v->len == 0) return LBK_OK;
        const i64 nl = n > 0 ? O.n_local : 0;
        const bool unit_cover = nl > 0 && (O.inc == 1 || O.inc == -1);
        const i64 from = unit_cover ? nl : 0;
        if (from < O.v->len) return device_copy(c, w + from, (const T *)O.v->ptr + from, O.v->len - from, dtype);
        return LBK_OK;
    };
    if ((s = prepare(X, xw, F::WX))) return s;
    if ((s = prepare(Y, yw, F::WY))) return s;
    // a strided out-of-place update reads the copy it just made (same values as the source)
    Operand Xs = X, Ys = Y;
    lbk_vec xtmp, ytmp;
    const bool x_unit = use_x && (X.inc == 1 || X.inc == -1), y_unit = use_y && (Y.inc == 1 || Y.inc == -1);
    if (F::WX && (char *)xw != x->ptr && !x_unit) { xtmp = *x; xtmp.ptr = (char *)xw; Xs.v = &xtmp; }
    if (F::WY && (char *)yw != y->ptr && !y_unit) { ytmp = *y; ytmp.ptr = (char *)yw; Ys.v = &ytmp; }
    return map_dispatch<F>(c, rid, f, n, Xs, xw, Ys, yw);
}

// ------------------------------------------------------------------------------------------------
// reduce launchers
// ------------------------------------------------------------------------------------------------
// Runs the local reduction and, on a communicator with sharded operands, the exchange + finish.
// `out` is a device-visible address that receives the scalar (T, or the int64 for iamax).
template <class Op>
static int reduce_dispatch(lbk_ctx *c, int rid, int kind, i64 n, const Operand &X, const Operand *Y, double sb, void *out)
{
    using T = typename Op::Scalar;
    if (c->capturing && (out == c->result_dev || (c->nranks > 1 && X.v->sharded)))
        return fail(c, LBK_ERR_UNSUPPORTED, "a reduction that returns its scalar to the host, or exchanges records between GPUs, cannot be captured: "
                                            "use the _async form on unsharded vectors (lbk_capture_begin)");
    ON_DEVICE(c);
    const bool exchange = c->nranks > 1 && X.v->sharded;
    const bool fused = exchange && c->fused_exchange;
    const bool by_events = exchange && !fused && c->parent != nullptr;      // a rank of a multi-device context: exchange_finish follows
    const i64 nl = X.n_local;
    const T *x = (const T *)X.ptr(), *y = Y ? (const T *)Y->ptr() : nullptr;
    const i64 incx = X.inc, incy = Y ? Y->inc : 1;
    FinishArgs fa;
    memset(&fa, 0, sizeof fa);
    fa.kind = kind;
    fa.mode = exchange ? (fused ? 2 : 1) : 0;
    fa.is_f64 = (kind != 3 && sizeof(T) == 8) ? 1 : 0;
    if (fused) {
        if (++c->exchange_seq == 0) ++c->exchange_seq;    // tag 0 means "empty slot"
        fa.peers = c->peers_dev; fa.nranks = c->nranks; fa.rank = c->rank; fa.seq = c->exchange_seq;
        fa.error_flag = c->error_flag_dev;
        fa.timeout_ns = c->exchange_timeout_ns;
    }
    if (exchange && !fused && !by_events && !c->comm) return fail(c, LBK_ERR_UNSUPPORTED, "sharded reduction: this context has neither peer mailboxes nor an NCCL communicator");
    if (by_events && !c->in_multi_call) return fail(c, LBK_ERR_UNSUPPORTED, "sharded reduction on a rank context: with the stream-ordered exchange (ranks sharing a device, or fused exchange off) the ranks are driven through the multi-device context");
    fa.n_global = n;
    fa.sb = sb;
    fa.out = out;
    fa.rec_out = by_events ? c->rec_ring + (++c->ring_seq & 1u) : c->rec_send;
    fa.x0 = (X.i0 == 0 && nl > 0) ? x : nullptr;     // logical element 0 lives at x[0] for incx >= 1
    if (out == c->result_dev) {                        // a synchronous call: the host polls for this tag
        if (++c->result_seq == 0) ++c->result_seq;
        fa.done_flag = reinterpret_cast<unsigned *>(static_cast<char *>(c->result_dev) + 8);
        fa.done_seq = c->result_seq;
    }
    const i64 idx_base = X.i0;
    const bool rev = nl > 0 && Y && ((incx == 1 && incy == -1) || (incx == -1 && incy == 1));
    const bool unit = nl > 0 && ((incx == 1 && (!Y || incy == 1)) || (Y && incx == -1 && incy == -1) || rev);
    if (unit) {
        const LaunchCfg &cfg = c->cfg[rid];
        const LaunchCfg *pcfg = &cfg;
        int variant = cfg.variant;
        int threads = cfg.threads > 0 ? cfg.threads : 256;
        const T *ptrs[2] = {x, y};
        i64 head = 0;
        int xshift = 0;
        const void *k;
        if (rev) {
            resolve_alignment_rev<T>(&variant, &head, &x, 1, &y, 1, nl);
            k = variant == kScalarVariant ? (const void *)lbk::reduce_unit_kernel<Op, 1, 4, 1, true>
                                          : (const void *)lbk::reduce_unit_kernel<Op, vec_t<T>(4), 4, 1, true>;
        } else {
            const LaunchCfg &mis = c->cfg[R_REDMISALIGNED];
            if (resolve_alignment<T>(&variant, &head, ptrs, Y ? 2 : 1, nl, kVariantVec[mis.variant] == 1 ? mis.variant : kScalarVariant) && variant == mis.variant) {
                pcfg = &mis;
                if (mis.threads > 0) threads = mis.threads;
            }
            k = reduce_kernel_ptr<Op>(variant);
            if constexpr (Op::TWO) {
                const LaunchCfg &sh = c->cfg[R_REDSHIFTED];
                if (Y && variant_vec<T>(variant) == 1 && kVariantVec[sh.variant] == 4) {     // x and y misaligned differently
                    i64 h = 0;
                    const int shift = shifted_layout<T>(x, &y, 1, nl, &h);
                    // the Double instance of the default shape runs unroll 2: ptxas cannot fit unroll 4 of 8-byte elements
                    // plus the shifted lanes into 64 registers (variant 1 = 16-byte packs x 2)
                    const int sv = (sizeof(T) == 8 && sh.variant == kDefaultCfg[R_REDSHIFTED].variant) ? 1 : sh.variant;
                    const void *ks = shift ? reduce_kernel_ptr_shifted<Op>(variant_unroll(sv), shift) : nullptr;
                    if (ks) { k = ks; head = h; variant = sv; pcfg = &sh; threads = std::min(sh.threads > 0 ? sh.threads : 256, lbk::kShiftThreads); xshift = shift; }
                }
            }
        }
        const i64 npack = xshift ? std::max<i64>(0, (nl - head - (variant_vec<T>(variant) - xshift)) / variant_vec<T>(variant))
                                 : (nl - head) / variant_vec<T>(variant);
        const i64 per_tile = (i64)variant_unroll(variant) * threads;
        int grid;
        int s = grid_for(c, k, *pcfg, threads, (npack + per_tile - 1) / per_tile, &grid, kMaxGrid, pcfg->pinned ? 0 : per_tile * variant_vec<T>(variant) * (i64)sizeof(T));
        if (s) return s;
        i64 nn = nl, ib = idx_base;
        lbk_pdl::Op op;
        op.reduction = true; op.rid = rid;
        op.rd[0] = x; op.rdn[0] = nl * (i64)sizeof(T);
        op.rd[1] = y; op.rdn[1] = y ? nl * (i64)sizeof(T) : 0;
        op.nrd = 2;
        op.wr[0] = out; op.wrn[0] = 8; op.nwr = 1;
        const lbk_pdl::Plan plan = lbk_pdl::plan(c->pdl, op);
        int late_trigger = plan.late ? 1 : 0, wait_first = plan.mode == lbk_pdl::WAIT_BEFORE_LOAD ? 1 : 0;
        void *args[] = {(void *)&x, (void *)&y, (void *)&head, (void *)&nn, (void *)&ib, (void *)&c->partials, (void *)&c->ticket,
                        (void *)&fa, (void *)&late_trigger, (void *)&wait_first};
        if (xshift) c->shifted_launches++;
        walk_note(c, x, false); walk_note(c, y, false);        // a reduction walks forwards: its summation order is part of its result
        s = launch_stream_kernel(c, k, grid, threads, args, op, plan);
        if (s) return s;
        // with the NCCL transport the all-gather and the finish kernel follow on the stream: no chain through them
        if (exchange && !fused) lbk_pdl::interrupt(c->pdl);
    } else {
        const LaunchCfg &scfg = c->cfg[R_REDSTRIDED];
        const int threads = (scfg.threads > 0 && scfg.threads <= 256) ? scfg.threads : 256;
        // Every block does the same share of the work (grid-stride), so the grid is a whole number of resident waves:
        // one wave for increments of 1-2, four when an increment of 4 or more puts every element in its own sector
        // (config-4 sweep, profiles/r01_strided_launch_sweep.md).  The wave comes from the kernel's occupancy, not from
        // a constant: isamax's instance needs 40 registers, 6 resident blocks of 256 -- on a grid of 8 per SM its last
        // quarter ran as a second, mostly empty wave at 2.5x the time (profiles/r02_strided_reductions.md).
        const bool wide = std::max(iabs64(incx), Y ? iabs64(incy) : 0) >= 4;
        static std::atomic<int> cached{0};      // per instantiation: (threads << 8) | resident blocks
        int resident = cached.load(std::memory_order_relaxed);
        if ((resident >> 8) != threads) {
            int r = 0;
            CU(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r, lbk::reduce_strided_kernel<Op>, threads, 0));
            resident = (threads << 8) | std::min(std::max(r, 1), 255);
            cached.store(resident, std::memory_order_relaxed);
        }
        resident &= 255;
        const i64 cap = (i64)c->sms * (scfg.blocks_per_sm > 0 ? scfg.blocks_per_sm : (wide ? 4 * resident : resident));
        const int grid = (int)std::max<i64>(1, std::min<i64>(std::min<i64>((std::max<i64>(nl, 1) + threads - 1) / threads, cap), kMaxGrid));
        // joins the chain of overlapping kernels like the unit-stride reductions (spans of the operands, lbk_pdl.hpp)
        i64 kx0 = first_index(nl, incx), ky0 = first_index(nl, incy), nn = nl, ib = idx_base, ix = incx, iy = incy;
        lbk_pdl::Op op;
        op.reduction = true; op.rid = R_REDSTRIDED;
        op.rd[0] = x; op.rdn[0] = (1 + (nl - 1) * iabs64(incx)) * (i64)sizeof(T);
        op.rd[1] = y; op.rdn[1] = y ? (1 + (nl - 1) * iabs64(incy)) * (i64)sizeof(T) : 0;
        op.nrd = 2;
        op.wr[0] = out; op.wrn[0] = 8; op.nwr = 1;
        const lbk_pdl::Plan plan = lbk_pdl::plan(c->pdl, op);
        int late_trigger = plan.late ? 1 : 0, wait_first = plan.mode == lbk_pdl::WAIT_BEFORE_LOAD ? 1 : 0;
        void *args[] = {(void *)&x, (void *)&ix, (void *)&kx0, (void *)&y, (void *)&iy, (void *)&ky0, (void *)&nn, (void *)&ib,
                        (void *)&c->partials, (void *)&c->ticket, (void *)&fa, (void *)&late_trigger, (void *)&wait_first};
        walk_forget(c);
        int s = launch_stream_kernel(c, (const void *)lbk::reduce_strided_kernel<Op>, grid, threads, args, op, plan);
        if (s) return s;
        if (exchange && !fused) lbk_pdl::interrupt(c->pdl);
    }
    if (by_events) {           // the record is in this rank's ring slot once this event has fired
        lbk_pdl::interrupt(c->pdl);
        CU(c, cudaEventRecord(c->xchg_ev, c->stream));
        c->pending_fa = fa;
        c->finish_pending = true;
        return LBK_OK;
    }
    if (exchange && !fused) {
        lbk_pdl::interrupt(c->pdl);
        NC(c, nccl_api()->AllGather(c->rec_send, c->rec_all, sizeof(Record), ncclChar, c->comm, c->stream));
        fa.mode = 0;
        lbk::finish_records_kernel<<<1, 32, 0, c->stream>>>(c->rec_all, c->nranks, fa);
        KERNEL_CHECK(c);
    }
    return LBK_OK;
}

template <class T>
static int store_value(lbk_ctx *c, void *out, T v)
{
    if (out == c->result_dev) NOT_CAPTURING(c, "a reduction that returns its scalar to the host");
    ON_DEVICE(c);
    lbk_pdl::interrupt(c->pdl);
    unsigned *flag = nullptr;
    if (out == c->result_dev) {
        if (++c->result_seq == 0) ++c->result_seq;
        flag = reinterpret_cast<unsigned *>(static_cast<char *>(c->result_dev) + 8);
    }
    lbk::store_scalar_kernel<T><<<1, 32, 0, c->stream>>>((T *)out, v, flag, c->result_seq);
    KERNEL_CHECK(c);
    return LBK_OK;
}

static int check_out_vec(lbk_ctx *c, const char *who, lbk_vec *o, int dtype, i64 need_bytes)
{
    if (!o || o->ctx != c || o->dtype != dtype || o->len * esz(o->dtype) < need_bytes)
        return fail(c, LBK_ERR_ARG, std::string(who) + ": result vector is null, foreign, of the wrong type or too short");
    if ((uintptr_t)o->ptr & 7) return fail(c, LBK_ERR_ARG, std::string(who) + ": result vector must be 8-byte aligned");
    return LBK_OK;
}

// dot (kind 0) for T, sdsdot (kind 3) for float
template <class T>
static int dot_common(lbk_ctx *c, const char *who, int rid, int kind, i64 n, float sb, const lbk_vec *x, i64 incx,
                      const lbk_vec *y, i64 incy, void *out_dev)
{
    constexpr int dtype = sizeof(T) == 8 ? 1 : 0;
    LIVE(c, who);
    // equal negative increments on shards: the same set of products as with their absolute values (a sum has no order here)
    if (x && y && x->sharded && y->sharded && incx < 0 && incx == incy && incx != INT64_MIN) { incx = -incx; incy = -incy; }
    Operand X{}, Y{};
    int s;
    if ((s = check_operand(c, who, x, dtype, incx, n, &X))) return s;
    if ((s = check_operand(c, who, y, dtype, incy, n, &Y))) return s;
    if (x->sharded != y->sharded || (x->sharded && (x->global_len != y->global_len || (n > 0 && incx != incy))))
        return fail(c, LBK_ERR_UNSUPPORTED, std::string(who) + ": sharded operands need equal lengths and equal increments");
    if (n <= 0) return store_value<T>(c, out_dev, kind == 3 ? (T)sb : (T)0);       // Single.hs:81 (empty fold), :239
    if constexpr (sizeof(T) == 4) {
        if (kind == 3) return reduce_dispatch<lbk::DsdotOp>(c, rid, 3, n, X, &Y, (double)sb, out_dev);
    }
    return reduce_dispatch<lbk::DotOp<T>>(c, rid, 0, n, X, &Y, 0.0, out_dev);
}

template <class R>
static int finish_sync(lbk_ctx *c, int s, R *out)
{
    if (s) return s;
    // The kernel publishes the scalar into mapped host memory and then the tag; polling for the tag returns
    // as soon as the value has crossed PCIe instead of after the stream has drained and signalled.  A result
    // that has not arrived within the spin budget (a long kernel, or a failed one) goes through
    // cudaStreamSynchronize, which also reports the error.
    volatile unsigned *tag = reinterpret_cast<volatile unsigned *>(static_cast<char *>(c->result_host) + 8);
    const unsigned want = c->result_seq;
    bool seen = false;
    const auto t0 = std::chrono::steady_clock::now();
    for (int spin = 0; !seen; ++spin) {
        seen = (*tag == want);
        if (!seen && (spin & 63) == 63 &&
            std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(200)) break;
    }
    if (!seen) { ON_DEVICE(c); CU(c, cudaStreamSynchronize(c->stream)); }
    std::atomic_thread_fence(std::memory_order_acquire);
    *out = *reinterpret_cast<volatile R *>(c->result_host);
    return check_exchange_error(c);
}

// asum (kind 0 with AsumOp), nrm2 (kind 1), iamax (kind 2): one vector, positive increments
template <class T>
static int ivi_common(lbk_ctx *c, const char *who, int rid, i64 n, const lbk_vec *x, i64 incx, void *out_dev)
{
    constexpr int dtype = sizeof(T) == 8 ? 1 : 0;
    LIVE(c, who);
    if (!x) return fail(c, LBK_ERR_ARG, std::string(who) + ": null vector");
    if (x->ctx != c) return fail(c, LBK_ERR_ARG, std::string(who) + ": vector belongs to another context");
    if (x->dtype != dtype) return fail(c, LBK_ERR_ARG, std::string(who) + ": vector has the wrong element type");
    // the reference's guards come first: Single.hs:162 (asum), :181 (nrm2), :218 (iamax)
    if (n < 1 || incx < 1) return rid == R_IAMAX ? store_value<i64>(c, out_dev, -1) : store_value<T>(c, out_dev, (T)0);
    Operand X{};
    int s;
    if ((s = check_operand(c, who, x, dtype, incx, n, &X))) return s;
    if (rid == R_ASUM) return reduce_dispatch<lbk::AsumOp<T>>(c, rid, 0, n, X, nullptr, 0.0, out_dev);
    if (rid == R_NRM2) return reduce_dispatch<lbk::Nrm2Op<T>>(c, rid, 1, n, X, nullptr, 0.0, out_dev);
    return reduce_dispatch<lbk::IamaxOp<T>>(c, rid, 2, n, X, nullptr, 0.0, out_dev);
}

// Second half of a multi-device context's stream-ordered exchange, after EVERY rank has launched its reduction and
// recorded its event: this rank's stream waits for the other ranks' records, then folds all of them.
static int exchange_finish(lbk_ctx *c)
{
    if (!c->finish_pending) return LBK_OK;
    c->finish_pending = false;
    ON_DEVICE(c);
    lbk_pdl::interrupt(c->pdl);
    for (lbk_ctx *q : c->parent->subs)
        if (q != c) CU(c, cudaStreamWaitEvent(c->stream, q->xchg_ev, 0));
    FinishArgs fa = c->pending_fa;
    fa.mode = 0;
    lbk::finish_gather_kernel<<<1, 32, 0, c->stream>>>(c->ring_table_dev, (int)(c->ring_seq & 1u), c->nranks, fa);
    KERNEL_CHECK(c);
    return LBK_OK;
}

// A reduction on a multi-device context.  call(rank context, rank, out) enqueues the routine on one rank with
// `out` as the device-visible address of its scalar: every rank computes the same bits (the records are folded
// in rank order everywhere); the caller gets rank 0's -- through rank 0's mapped result slot (synchronous
// forms, `out_host`) or in out_vec[0] (enqueue-only forms; out_vec is an unsharded vector, i.e. on rank 0,
// and the other ranks park their copy in a private sink).
template <class R, class Call>
static int multi_reduce(lbk_ctx *m, const char *who, std::initializer_list<const lbk_vec *> vs, lbk_vec *out_vec, int out_dtype,
                        i64 out_bytes, R *out_host, Call call)
{
    int count, s = multi_count(m, who, vs, &count);
    if (s) return s;
    for (const lbk_vec *v : vs)
        if (!v) return fail(m, LBK_ERR_ARG, std::string(who) + ": null vector");
    void *out0 = nullptr;
    if (!out_host) {
        if (!out_vec || out_vec->ctx != m || out_vec->sharded) return fail(m, LBK_ERR_ARG, std::string(who) + ": the result vector must be an unsharded vector of this context");
        if ((s = check_out_vec(m->subs[0], who, out_vec->parts[0], out_dtype, out_bytes))) { m->err = m->subs[0]->err; return s; }
        out0 = out_vec->parts[0]->ptr;
    }
    for (lbk_ctx *sc : m->subs) { sc->in_multi_call = true; sc->finish_pending = false; }
    s = fan_out(m, count, [&](lbk_ctx *sc, int r) { return call(sc, r, out_host ? sc->result_dev : (r == 0 ? out0 : sc->sink)); });
    for (lbk_ctx *sc : m->subs) sc->in_multi_call = false;
    if (!s && count > 1 && !m->fused_exchange) s = fan_out(m, count, [&](lbk_ctx *sc, int) { return exchange_finish(sc); });
    if (!out_host) return s;
    s = finish_sync(m->subs[0], s, out_host);
    if (s && m->err.empty()) m->err = m->subs[0]->err;
    return s;
}

#define LBK_REDUCTIONS(P, T, DT, IPFX)                                                                                         \
    extern "C" int lbk_##P##dot(lbk_ctx *c, int64_t n, const lbk_vec *x, int64_t incx, const lbk_vec *y, int64_t incy, T *out)  \
    {                                                                                                                          \
        if (!c || !out) return fail(c, LBK_ERR_ARG, "lbk_" #P "dot: null argument");                                           \
        if (is_multi(c)) return multi_reduce<T>(c, "lbk_" #P "dot", {x, y}, nullptr, DT, sizeof(T), out, [&](lbk_ctx *sc, int r, void *o) { \
            return dot_common<T>(sc, "lbk_" #P "dot", R_DOT, 0, n, 0.f, part(x, r), incx, part(y, r), incy, o); });             \
        return finish_sync(c, dot_common<T>(c, "lbk_" #P "dot", R_DOT, 0, n, 0.f, x, incx, y, incy, c->result_dev), out);       \
    }                                                                                                                          \
    extern "C" int lbk_##P##dot_async(lbk_ctx *c, int64_t n, const lbk_vec *x, int64_t incx, const lbk_vec *y, int64_t incy, lbk_vec *o) \
    {                                                                                                                          \
        if (!c) return fail(c, LBK_ERR_ARG, "lbk_" #P "dot_async: null context");                                              \
        if (is_multi(c)) return multi_reduce<T>(c, "lbk_" #P "dot_async", {x, y}, o, DT, sizeof(T), nullptr, [&](lbk_ctx *sc, int r, void *od) { \
            return dot_common<T>(sc, "lbk_" #P "dot_async", R_DOT, 0, n, 0.f, part(x, r), incx, part(y, r), incy, od); });      \
        int s = check_out_vec(c, "lbk_" #P "dot_async", o, DT, sizeof(T));                                                     \
        return s ? s : dot_common<T>(c, "lbk_" #P "dot_async", R_DOT, 0, n, 0.f, x, incx, y, incy, o->ptr);                     \
    }                                                                                                                          \
    LBK_IVI(P##asum, T, T, DT, sizeof(T), R_ASUM)                                                                              \
    LBK_IVI(P##nrm2, T, T, DT, sizeof(T), R_NRM2)                                                                              \
    LBK_IVI(IPFX##amax, T, int64_t, DT, 8, R_IAMAX)
// one-vector reductions: NAME(ctx, n, x, incx, &out) and NAME_async(ctx, n, x, incx, out_vec)
#define LBK_IVI(NAME, T, RT, DT, OUTBYTES, RID)                                                                                \
    extern "C" int lbk_##NAME(lbk_ctx *c, int64_t n, const lbk_vec *x, int64_t incx, RT *out)                                   \
    {                                                                                                                          \
        if (!c || !out) return fail(c, LBK_ERR_ARG, "lbk_" #NAME ": null argument");                                           \
        if (is_multi(c)) return multi_reduce<RT>(c, "lbk_" #NAME, {x}, nullptr, DT, OUTBYTES, out, [&](lbk_ctx *sc, int r, void *o) { \
            return ivi_common<T>(sc, "lbk_" #NAME, RID, n, part(x, r), incx, o); });                                           \
        return finish_sync(c, ivi_common<T>(c, "lbk_" #NAME, RID, n, x, incx, c->result_dev), out);                             \
    }                                                                                                                          \
    extern "C" int lbk_##NAME##_async(lbk_ctx *c, int64_t n, const lbk_vec *x, int64_t incx, lbk_vec *o)                        \
    {                                                                                                                          \
        if (!c) return fail(c, LBK_ERR_ARG, "lbk_" #NAME "_async: null context");                                              \
        if (is_multi(c)) return multi_reduce<RT>(c, "lbk_" #NAME "_async", {x}, o, DT, OUTBYTES, nullptr, [&](lbk_ctx *sc, int r, void *od) { \
            return ivi_common<T>(sc, "lbk_" #NAME "_async", RID, n, part(x, r), incx, od); });                                 \
        int s = check_out_vec(c, "lbk_" #NAME "_async", o, DT, OUTBYTES);                                                      \
        return s ? s : ivi_common<T>(c, "lbk_" #NAME "_async", RID, n, x, incx, o->ptr);                                        \
    }
LBK_REDUCTIONS(s, float, 0, is)
LBK_REDUCTIONS(d, double, 1, id)

extern "C" int lbk_sdsdot(lbk_ctx *c, int64_t n, float sb, const lbk_vec *x, int64_t incx, const lbk_vec *y, int64_t incy, float *out)
{
    if (!c || !out) return fail(c, LBK_ERR_ARG, "lbk_sdsdot: null argument");
    if (is_multi(c)) return multi_reduce<float>(c, "lbk_sdsdot", {x, y}, nullptr, 0, sizeof(float), out, [&](lbk_ctx *sc, int r, void *o) {
        return dot_common<float>(sc, "lbk_sdsdot", R_SDSDOT, 3, n, sb, part(x, r), incx, part(y, r), incy, o); });
    return finish_sync(c, dot_common<float>(c, "lbk_sdsdot", R_SDSDOT, 3, n, sb, x, incx, y, incy, c->result_dev), out);
}

extern "C" int lbk_sdsdot_async(lbk_ctx *c, int64_t n, float sb, const lbk_vec *x, int64_t incx, const lbk_vec *y, int64_t incy, lbk_vec *o)
{
    if (!c) return fail(c, LBK_ERR_ARG, "lbk_sdsdot_async: null context");
    if (is_multi(c)) return multi_reduce<float>(c, "lbk_sdsdot_async", {x, y}, o, 0, sizeof(float), nullptr, [&](lbk_ctx *sc, int r, void *od) {
        return dot_common<float>(sc, "lbk_sdsdot_async", R_SDSDOT, 3, n, sb, part(x, r), incx, part(y, r), incy, od); });
    int s = check_out_vec(c, "lbk_sdsdot_async", o, 0, sizeof(float));
    return s ? s : dot_common<float>(c, "lbk_sdsdot_async", R_SDSDOT, 3, n, sb, x, incx, y, incy, o->ptr);
}

// ------------------------------------------------------------------------------------------------
// elementwise entry points
// ------------------------------------------------------------------------------------------------
// Single.hs:108-111: the index stream `sampleIndices n incx` is empty for incx < 0 and n > 1, and
// yields slot 0 (n times for incx == 0, once for n == 1), each write being x[0]*a.
template <class T>
static int scal_common(lbk_ctx *c, const char *who, i64 n, T a, const lbk_vec *x, i64 incx, lbk_vec *xout)
{
    if (!c) return fail(nullptr, LBK_ERR_ARG, std::string(who) + ": null context");
    if (!x) return fail(c, LBK_ERR_ARG, std::string(who) + ": null vector");
    if (is_multi(c)) {
        int count, st = multi_count(c, who, {x, xout}, &count);
        if (st) return st;
        return fan_out(c, count, [&](lbk_ctx *sc, int r) { return scal_common<T>(sc, who, n, a, part(x, r), incx, part(xout, r)); });
    }
    LIVE(c, who);
    if (x->ctx != c || (xout && xout->ctx != c)) return fail(c, LBK_ERR_ARG, std::string(who) + ": vector belongs to another context");
    i64 n_eff = n, inc_eff = incx;
    if (n > 0 && incx <= 0) {
        if (incx == 0 || n == 1) { n_eff = 1; inc_eff = 1; }
        else n_eff = 0;
        if (x->sharded) return fail(c, LBK_ERR_UNSUPPORTED, std::string(who) + ": sharded vectors take positive increments only");
    }
    if (n_eff <= 0 && xout && xout->ptr != x->ptr) {      // nothing to scale: the pure result is a copy
        if (xout->ctx != c || xout->len != x->len || xout->dtype != x->dtype) return fail(c, LBK_ERR_ARG, std::string(who) + ": output vector must match the input's length and type");
        return x->len ? device_copy(c, xout->ptr, x->ptr, x->len, x->dtype) : LBK_OK;
    }
    return map_entry(c, who, R_SCAL, lbk::ScalF<T>{a}, n_eff, x, inc_eff, nullptr, 1, xout, nullptr);
}

// Single.hs:472-482; param = [flag,h11,h21,h12,h22] (test/Foreign.hs:276-280)
template <class T>
static int rotm_common(lbk_ctx *c, const char *who, i64 n, const lbk_vec *x, i64 incx, const lbk_vec *y, i64 incy,
                       const T *p, lbk_vec *xout, lbk_vec *yout)
{
    if (!c) return fail(nullptr, LBK_ERR_ARG, std::string(who) + ": null context");
    if (!p) return fail(c, LBK_ERR_ARG, std::string(who) + ": null param");
    if (is_multi(c)) {
        int count, st = multi_count(c, who, {x, y, xout, yout}, &count);
        if (st) return st;
        if (!x || !y) return fail(c, LBK_ERR_ARG, std::string(who) + ": null vector");
        return fan_out(c, count, [&](lbk_ctx *sc, int r) { return rotm_common<T>(sc, who, n, part(x, r), incx, part(y, r), incy, p, part(xout, r), part(yout, r)); });
    }
    LIVE(c, who);
    const T flag = p[0];
    if (flag == (T)-2) {          // FLAGNEG2: the inputs come back as they are (Single.hs:473)
        const lbk_vec *src[2] = {x, y};
        lbk_vec *dst[2] = {xout, yout};
        for (int i = 0; i < 2; ++i)
            if (dst[i] && src[i] && dst[i]->ptr != src[i]->ptr && src[i]->len) {
                if (dst[i]->ctx != c || src[i]->ctx != c || dst[i]->len != src[i]->len || dst[i]->dtype != src[i]->dtype) return fail(c, LBK_ERR_ARG, std::string(who) + ": output vector must match the input's length and type");
                int s = device_copy(c, dst[i]->ptr, src[i]->ptr, src[i]->len, src[i]->dtype);
                if (s) return s;
            }
        return LBK_OK;
    }
    if (flag < (T)0) return map_entry(c, who, R_ROTM, lbk::RotmNeg1F<T>{p[1], p[3], p[2], p[4]}, n, x, incx, y, incy, xout, yout);
    if (flag == (T)0) return map_entry(c, who, R_ROTM, lbk::Rotm0F<T>{p[3], p[2]}, n, x, incx, y, incy, xout, yout);
    return map_entry(c, who, R_ROTM, lbk::Rotm1F<T>{p[1], p[4]}, n, x, incx, y, incy, xout, yout);
}

#define LBK_NEED(p, name) if (!(p)) return fail(c, LBK_ERR_ARG, name ": null output")
#define LBK_MAPS(P, T)                                                                                                          \
    extern "C" int lbk_##P##axpy(lbk_ctx *c, int64_t n, T a, const lbk_vec *x, int64_t incx, lbk_vec *y, int64_t incy)          \
    { return map_entry(c, "lbk_" #P "axpy", R_AXPY, lbk::AxpyF<T>{a}, n, x, incx, y, incy, nullptr, nullptr); }                 \
    extern "C" int lbk_##P##axpy_oop(lbk_ctx *c, int64_t n, T a, const lbk_vec *x, int64_t incx, const lbk_vec *y, int64_t incy, lbk_vec *yout) \
    { LBK_NEED(yout, "lbk_" #P "axpy_oop"); return map_entry(c, "lbk_" #P "axpy_oop", R_AXPY, lbk::AxpyF<T>{a}, n, x, incx, y, incy, nullptr, yout); } \
    extern "C" int lbk_##P##scal(lbk_ctx *c, int64_t n, T a, lbk_vec *x, int64_t incx)                                          \
    { return scal_common<T>(c, "lbk_" #P "scal", n, a, x, incx, nullptr); }                                                     \
    extern "C" int lbk_##P##scal_oop(lbk_ctx *c, int64_t n, T a, const lbk_vec *x, int64_t incx, lbk_vec *xout)                 \
    { LBK_NEED(xout, "lbk_" #P "scal_oop"); return scal_common<T>(c, "lbk_" #P "scal_oop", n, a, x, incx, xout); }              \
    extern "C" int lbk_##P##copy(lbk_ctx *c, int64_t n, const lbk_vec *x, int64_t incx, lbk_vec *y, int64_t incy)               \
    { return map_entry(c, "lbk_" #P "copy", R_COPY, lbk::CopyF<T>{}, n, x, incx, y, incy, nullptr, nullptr); }                  \
    extern "C" int lbk_##P##copy_oop(lbk_ctx *c, int64_t n, const lbk_vec *x, int64_t incx, const lbk_vec *y, int64_t incy, lbk_vec *yout) \
    { LBK_NEED(yout, "lbk_" #P "copy_oop"); return map_entry(c, "lbk_" #P "copy_oop", R_COPY, lbk::CopyF<T>{}, n, x, incx, y, incy, nullptr, yout); } \
    extern "C" int lbk_##P##swap(lbk_ctx *c, int64_t n, lbk_vec *x, int64_t incx, lbk_vec *y, int64_t incy)                     \
    { return map_entry(c, "lbk_" #P "swap", R_SWAP, lbk::SwapF<T>{}, n, x, incx, y, incy, nullptr, nullptr); }                  \
    extern "C" int lbk_##P##swap_oop(lbk_ctx *c, int64_t n, const lbk_vec *x, int64_t incx, const lbk_vec *y, int64_t incy, lbk_vec *xout, lbk_vec *yout) \
    { LBK_NEED(xout && yout, "lbk_" #P "swap_oop"); return map_entry(c, "lbk_" #P "swap_oop", R_SWAP, lbk::SwapF<T>{}, n, x, incx, y, incy, xout, yout); } \
    extern "C" int lbk_##P##rot(lbk_ctx *c, int64_t n, lbk_vec *x, int64_t incx, lbk_vec *y, int64_t incy, T cs, T sn)          \
    { return map_entry(c, "lbk_" #P "rot", R_ROT, lbk::RotF<T>{cs, sn}, n, x, incx, y, incy, nullptr, nullptr); }               \
    extern "C" int lbk_##P##rot_oop(lbk_ctx *c, int64_t n, const lbk_vec *x, int64_t incx, const lbk_vec *y, int64_t incy, T cs, T sn, lbk_vec *xout, lbk_vec *yout) \
    { LBK_NEED(xout && yout, "lbk_" #P "rot_oop"); return map_entry(c, "lbk_" #P "rot_oop", R_ROT, lbk::RotF<T>{cs, sn}, n, x, incx, y, incy, xout, yout); } \
    extern "C" int lbk_##P##rotm(lbk_ctx *c, int64_t n, lbk_vec *x, int64_t incx, lbk_vec *y, int64_t incy, const T param[5])   \
    { return rotm_common<T>(c, "lbk_" #P "rotm", n, x, incx, y, incy, param, nullptr, nullptr); }                               \
    extern "C" int lbk_##P##rotm_oop(lbk_ctx *c, int64_t n, const lbk_vec *x, int64_t incx, const lbk_vec *y, int64_t incy, const T param[5], lbk_vec *xout, lbk_vec *yout) \
    { LBK_NEED(xout && yout, "lbk_" #P "rotm_oop"); return rotm_common<T>(c, "lbk_" #P "rotm_oop", n, x, incx, y, incy, param, xout, yout); }
LBK_MAPS(s, float)
LBK_MAPS(d, double)

// ------------------------------------------------------------------------------------------------
// host statement of the cross-rank combine (tests of the multi-GPU path on CPU)
// ------------------------------------------------------------------------------------------------
extern "C" int lbk_combine_records(int kind, int is_f64, const lbk_record *recs, int nranks, double *fout, int64_t *iout)
{
    if (!recs || nranks < 1) return fail(nullptr, LBK_ERR_ARG, "lbk_combine_records: bad argument");
    const Record *r = reinterpret_cast<const Record *>(recs);
    if (kind != LBK_RED_SUM && kind != LBK_RED_NRM2 && kind != LBK_RED_IAMAX) return fail(nullptr, LBK_ERR_ARG, "lbk_combine_records: unknown kind");
    const Record a = lbk::fold_by_kind(r, nranks, kind);
    if (kind == LBK_RED_IAMAX) {
        if (!iout) return fail(nullptr, LBK_ERR_ARG, "lbk_combine_records: null out pointer");
        *iout = a.idx;
        return LBK_OK;
    }
    if (!fout) return fail(nullptr, LBK_ERR_ARG, "lbk_combine_records: null out pointer");
    double v = a.f[0];
    if (kind == LBK_RED_NRM2)
        v = is_f64 ? lbk::nrm2_finish(a.f[0], a.f[1], a.f[2], (double)lbk::Blue<double>::sbig, (double)lbk::Blue<double>::ssml)
                   : lbk::nrm2_finish(a.f[0], a.f[1], a.f[2], (double)lbk::Blue<float>::sbig, (double)lbk::Blue<float>::ssml);
    *fout = is_f64 ? v : (double)(float)v;
    return LBK_OK;
}

// ------------------------------------------------------------------------------------------------
// the launcher threads of lbk_team.hpp on their own (no device needed): tests/test_team.py
// ------------------------------------------------------------------------------------------------
// Runs `jobs` jobs on a team of `nranks`; job j makes rank r add (j+1)*(r+1) to its own counter and fail with
// status 100+r when j == fail_at (-1: never).  Returns through *sum the total of all counters, through
// *first_status the status the failing job reported (0 if none), and checks that every job ran exactly once
// per rank (LBK_ERR_ARG otherwise).  pause_us > 0 idles between jobs so that the members go to sleep and
// have to be woken.
extern "C" int lbk_team_selftest(int nranks, int jobs, int fail_at, int pause_us, int64_t *sum, int *first_status)
{
    if (nranks < 1 || nranks > 64 || jobs < 0 || !sum || !first_status) return fail(nullptr, LBK_ERR_ARG, "lbk_team_selftest: bad argument");
    std::vector<int64_t> counter(nranks, 0), runs(nranks, 0);
    std::atomic<int> started{0};
    *first_status = 0;
    {
        lbk_team::Team team(nranks, [&](int) { started.fetch_add(1); });
        for (int j = 0; j < jobs; ++j) {
            const int st = team.run([&](int r) { counter[r] += (int64_t)(j + 1) * (r + 1); runs[r]++; return j == fail_at ? 100 + r : 0; });
            if (st && !*first_status) *first_status = st;
            if (pause_us > 0) std::this_thread::sleep_for(std::chrono::microseconds(pause_us));
        }
    }
    *sum = 0;
    for (int r = 0; r < nranks; ++r) {
        *sum += counter[r];
        if (runs[r] != jobs) return fail(nullptr, LBK_ERR_ARG, "lbk_team_selftest: a rank missed or repeated a job");
    }
    if (started.load() != nranks - 1) return fail(nullptr, LBK_ERR_ARG, "lbk_team_selftest: on_start did not run once per member");
    return LBK_OK;
}

// ------------------------------------------------------------------------------------------------
// the launch planner of lbk_pdl.hpp on its own (no device needed): tests/test_pdl_plan.py
// ------------------------------------------------------------------------------------------------
extern "C" int lbk_pdl_sim_create(void **sim)
{
    if (!sim) return fail(nullptr, LBK_ERR_ARG, "lbk_pdl_sim_create: null out pointer");
    *sim = new (std::nothrow) lbk_pdl::State();
    return *sim ? LBK_OK : fail(nullptr, LBK_ERR_NOMEM, "out of host memory");
}
extern "C" int lbk_pdl_sim_destroy(void *sim)
{
    delete static_cast<lbk_pdl::State *>(sim);
    return LBK_OK;
}
extern "C" int lbk_pdl_sim_enable(void *sim, int enabled)
{
    if (!sim) return fail(nullptr, LBK_ERR_ARG, "lbk_pdl_sim_enable: null simulator");
    static_cast<lbk_pdl::State *>(sim)->enabled = enabled != 0;
    return LBK_OK;
}
extern "C" int lbk_pdl_sim_interrupt(void *sim)
{
    if (!sim) return fail(nullptr, LBK_ERR_ARG, "lbk_pdl_sim_interrupt: null simulator");
    lbk_pdl::interrupt(*static_cast<lbk_pdl::State *>(sim));
    return LBK_OK;
}
// rd / wr: (address, bytes) pairs of what the kernel reads / writes; mode: 0 ordinary launch, 1 PDL with the
// wait at exit, 2 PDL with the wait before the first store; late: a reduction releases its dependents late
extern "C" int lbk_pdl_sim_op(void *sim, int reduction, int rid, const int64_t *rd, int nrd, const int64_t *wr, int nwr,
                              int *mode, int *late)
{
    if (!sim || nrd < 0 || nrd > 2 || nwr < 0 || nwr > 2 || (nrd && !rd) || (nwr && !wr) || !mode || !late)
        return fail(nullptr, LBK_ERR_ARG, "lbk_pdl_sim_op: bad argument");
    lbk_pdl::State &st = *static_cast<lbk_pdl::State *>(sim);
    lbk_pdl::Op op;
    op.reduction = reduction != 0; op.rid = rid;
    op.nrd = nrd; op.nwr = nwr;
    for (int i = 0; i < nrd; ++i) { op.rd[i] = reinterpret_cast<const void *>((uintptr_t)rd[2 * i]); op.rdn[i] = rd[2 * i + 1]; }
    for (int i = 0; i < nwr; ++i) { op.wr[i] = reinterpret_cast<const void *>((uintptr_t)wr[2 * i]); op.wrn[i] = wr[2 * i + 1]; }
    const lbk_pdl::Plan plan = lbk_pdl::plan(st, op);
    lbk_pdl::commit(st, op, plan);
    *mode = (int)plan.mode;
    *late = plan.late ? 1 : 0;
    return LBK_OK;
}

