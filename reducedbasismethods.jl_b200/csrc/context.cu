// context.cu -- rbm_ctx life cycle, error reporting, pointer staging, and the run-time
// binding to NCCL (one process per GPU; the communicator is only used to sum the
// charge density / diagnostics per step and the Gram matrix once).
#include <dlfcn.h>

#include <cstring>

#include "common.cuh"

static thread_local std::string g_init_error;
static void xch_release(rbm_ctx *ctx);

int rbm_fail(rbm_ctx *ctx, int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    else g_init_error = buf;
    return code;
}

bool rbm_is_device_ptr(const void *p)
{
    if (!p) return false;
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

int DevIn::stage(rbm_ctx *ctx, const void *src, size_t bytes, DevBuf *persist)
{
    if (!src || bytes == 0) {
        ptr = src;
        return RBM_OK;
    }
    if (rbm_is_device_ptr(src)) {
        ptr = src;
        return RBM_OK;
    }
    DevBuf &b = persist ? *persist : buf;
    if (persist) {
        if (!(b.p && b.bytes >= bytes)) RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // the old buffer may be in use
        RBM_CUDA(ctx, b.reserve(bytes));
    } else {
        RBM_CUDA(ctx, b.alloc(bytes));
    }
    RBM_CUDA(ctx, cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    ptr = b.p;
    return RBM_OK;
}

int DevOut::prepare(rbm_ctx *ctx, void *dst, size_t nbytes, bool zero, DevBuf *persist)
{
    bytes = nbytes;
    host = nullptr;
    ptr = dst;
    if (!dst || nbytes == 0) return RBM_OK;
    if (!rbm_is_device_ptr(dst)) {
        DevBuf &b = persist ? *persist : buf;
        if (persist) {
            if (!(b.p && b.bytes >= nbytes)) RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            RBM_CUDA(ctx, b.reserve(nbytes));
        } else {
            RBM_CUDA(ctx, b.alloc(nbytes));
        }
        ptr = b.p;
        host = dst;
    }
    if (zero) RBM_CUDA(ctx, cudaMemsetAsync(ptr, 0, nbytes, ctx->stream));
    return RBM_OK;
}

int DevOut::flush(rbm_ctx *ctx)
{
    if (host && bytes)
        RBM_CUDA(ctx, cudaMemcpyAsync(host, ptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return RBM_OK;
}

extern "C" int rbm_version(void) { return RBM_VERSION; }

extern "C" int rbm_init(int device, rbm_ctx **out)
{
    if (!out) return rbm_fail(nullptr, RBM_ERR_INVALID, "rbm_init: out == NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return rbm_fail(nullptr, RBM_ERR_UNSUPPORTED,
                        "rbm_init: no CUDA device (%s); librbm_b200 has no CPU fallback",
                        e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    }
    if (device < 0 || device >= ndev)
        return rbm_fail(nullptr, RBM_ERR_INVALID, "rbm_init: device %d out of range [0,%d)", device,
                        ndev);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return rbm_fail(nullptr, RBM_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return rbm_fail(nullptr, RBM_ERR_UNSUPPORTED,
                        "rbm_init: device %d is sm_%d%d; librbm_b200 is built for sm_100a only",
                        device, prop.major, prop.minor);
    if ((e = cudaSetDevice(device)) != cudaSuccess)
        return rbm_fail(nullptr, RBM_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    rbm_ctx *ctx = new (std::nothrow) rbm_ctx();
    if (!ctx) return rbm_fail(nullptr, RBM_ERR_NOMEM, "rbm_init: out of host memory");
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = (int)prop.sharedMemPerBlockOptin;
    // a blocking stream: ordered with the legacy default stream a host (torch, CUDA.jl) produces its device arrays on
    if ((e = cudaStreamCreate(&ctx->own_stream)) != cudaSuccess) {
        delete ctx;
        return rbm_fail(nullptr, RBM_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    }
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return RBM_OK;
}

extern "C" const char *rbm_last_error(rbm_ctx *ctx)
{
    return ctx ? ctx->err.c_str() : g_init_error.c_str();
}

extern "C" int rbm_set_stream(rbm_ctx *ctx, void *cuda_stream)
{
    if (!ctx) return RBM_ERR_INVALID;
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return RBM_OK;
}

extern "C" int rbm_synchronize(rbm_ctx *ctx)
{
    if (!ctx) return RBM_ERR_INVALID;
    RBM_CUDA(ctx, cudaSetDevice(ctx->device));
    RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RBM_OK;
}

extern "C" int rbm_device_info(rbm_ctx *ctx, int *sm_count, int64_t *free_bytes, int64_t *total_bytes)
{
    if (!ctx) return RBM_ERR_INVALID;
    RBM_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t f = 0, t = 0;
    RBM_CUDA(ctx, cudaMemGetInfo(&f, &t));
    if (sm_count) *sm_count = ctx->sm_count;
    if (free_bytes) *free_bytes = (int64_t)f;
    if (total_bytes) *total_bytes = (int64_t)t;
    return RBM_OK;
}

extern "C" int rbm_device_malloc(rbm_ctx *ctx, int64_t bytes, void **out)
{
    if (!ctx || !out || bytes < 0) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_device_malloc: bad argument");
    RBM_CUDA(ctx, cudaSetDevice(ctx->device));
    RBM_CUDA(ctx, cudaMalloc(out, (size_t)bytes));
    return RBM_OK;
}

extern "C" int rbm_device_free(rbm_ctx *ctx, void *ptr)
{
    if (!ctx) return RBM_ERR_INVALID;
    RBM_CUDA(ctx, cudaSetDevice(ctx->device));
    RBM_CUDA(ctx, cudaFree(ptr));
    return RBM_OK;
}

extern "C" int rbm_memcpy(rbm_ctx *ctx, void *dst, const void *src, int64_t bytes)
{
    if (!ctx || bytes < 0 || (bytes > 0 && (!dst || !src)))
        return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_memcpy: bad argument");
    RBM_CUDA(ctx, cudaSetDevice(ctx->device));
    RBM_CUDA(ctx, cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, ctx->stream));
    RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RBM_OK;
}

extern "C" int rbm_host_register(rbm_ctx *ctx, void *ptr, int64_t bytes)
{
    if (!ctx || !ptr || bytes <= 0) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_host_register: bad argument");
    RBM_CUDA(ctx, cudaSetDevice(ctx->device));
    RBM_CUDA(ctx, cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable));
    return RBM_OK;
}

extern "C" int rbm_host_unregister(rbm_ctx *ctx, void *ptr)
{
    if (!ctx || !ptr) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_host_unregister: bad argument");
    RBM_CUDA(ctx, cudaSetDevice(ctx->device));
    RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // no copy may still target the range
    RBM_CUDA(ctx, cudaHostUnregister(ptr));
    return RBM_OK;
}

extern "C" int64_t rbm_launch_count(rbm_ctx *ctx) { return ctx ? ctx->launches : 0; }

static void prof_clear(rbm_ctx *ctx)
{
    for (auto &r : ctx->prof) {
        cudaEventDestroy(r.beg);
        cudaEventDestroy(r.end);
    }
    ctx->prof.clear();
}

extern "C" int rbm_profile_enable(rbm_ctx *ctx, int on)
{
    if (!ctx) return RBM_ERR_INVALID;
    RBM_CUDA(ctx, cudaSetDevice(ctx->device));
    RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    prof_clear(ctx);
    ctx->profiling = on != 0;
    return RBM_OK;
}

extern "C" int rbm_profile_query(rbm_ctx *ctx, const char *kernel, int64_t *launches, double *total_ms)
{
    if (!ctx || !kernel) return RBM_ERR_INVALID;
    RBM_CUDA(ctx, cudaSetDevice(ctx->device));
    RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int64_t n = 0;
    double ms = 0.0;
    for (auto &r : ctx->prof) {
        if (strcmp(r.name, kernel) != 0) continue;
        float t = 0.f;
        RBM_CUDA(ctx, cudaEventElapsedTime(&t, r.beg, r.end));
        ms += t;
        ++n;
    }
    if (launches) *launches = n;
    if (total_ms) *total_ms = ms;
    return RBM_OK;
}

// ---------------------------------------------------------------- NCCL (dlopen)

namespace {
typedef struct { char internal[128]; } nccl_uid;
typedef int (*fn_getuid)(nccl_uid *);
typedef int (*fn_initrank)(void **, int, nccl_uid, int);
typedef int (*fn_allreduce)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef int (*fn_allgather)(const void *, void *, size_t, int, void *, cudaStream_t);
typedef int (*fn_destroy)(void *);
typedef const char *(*fn_errstr)(int);

struct Nccl {
    void *h = nullptr;
    fn_getuid get_uid = nullptr;
    fn_initrank init_rank = nullptr;
    fn_allreduce all_reduce = nullptr;
    fn_allgather all_gather = nullptr;
    fn_destroy destroy = nullptr;
    fn_errstr errstr = nullptr;
    bool tried = false;
} g_nccl;

const int NCCL_F64 = 8;  // ncclFloat64
const int NCCL_SUM = 0;  // ncclSum

bool nccl_load()
{
    if (g_nccl.tried) return g_nccl.h != nullptr;
    g_nccl.tried = true;
    const char *env = getenv("RBM_NCCL_LIB");
    void *h = nullptr;
    if (env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // already in the process (torch)
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return false;
    g_nccl.get_uid = (fn_getuid)dlsym(h, "ncclGetUniqueId");
    g_nccl.init_rank = (fn_initrank)dlsym(h, "ncclCommInitRank");
    g_nccl.all_reduce = (fn_allreduce)dlsym(h, "ncclAllReduce");
    g_nccl.all_gather = (fn_allgather)dlsym(h, "ncclAllGather");
    g_nccl.destroy = (fn_destroy)dlsym(h, "ncclCommDestroy");
    g_nccl.errstr = (fn_errstr)dlsym(h, "ncclGetErrorString");
    if (!g_nccl.get_uid || !g_nccl.init_rank || !g_nccl.all_reduce || !g_nccl.all_gather ||
        !g_nccl.destroy) {
        dlclose(h);
        return false;
    }
    g_nccl.h = h;
    return true;
}
const char *nccl_err(int rc) { return g_nccl.errstr ? g_nccl.errstr(rc) : "nccl error"; }
}  // namespace

extern "C" int rbm_comm_unique_id(void *id128)
{
    if (!id128) return rbm_fail(nullptr, RBM_ERR_INVALID, "rbm_comm_unique_id: NULL");
    if (!nccl_load()) return rbm_fail(nullptr, RBM_ERR_COMM, "NCCL library not found (set RBM_NCCL_LIB)");
    nccl_uid id;
    int rc = g_nccl.get_uid(&id);
    if (rc) return rbm_fail(nullptr, RBM_ERR_COMM, "ncclGetUniqueId: %s", nccl_err(rc));
    memcpy(id128, &id, sizeof(id));
    return RBM_OK;
}

extern "C" int rbm_comm_init(rbm_ctx *ctx, int nranks, int rank, const void *id128)
{
    if (!ctx || !id128 || nranks < 1 || rank < 0 || rank >= nranks)
        return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_comm_init: bad argument");
    if (ctx->nccl_comm) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_comm_init: communicator already attached");
    if (!nccl_load()) return rbm_fail(ctx, RBM_ERR_COMM, "NCCL library not found (set RBM_NCCL_LIB)");
    RBM_CUDA(ctx, cudaSetDevice(ctx->device));
    nccl_uid id;
    memcpy(&id, id128, sizeof(id));
    void *comm = nullptr;
    int rc = g_nccl.init_rank(&comm, nranks, id, rank);
    if (rc) return rbm_fail(ctx, RBM_ERR_COMM, "ncclCommInitRank: %s", nccl_err(rc));
    ctx->nccl_comm = comm;
    ctx->nranks = nranks;
    ctx->rank = rank;
    return RBM_OK;
}

extern "C" int rbm_comm_destroy(rbm_ctx *ctx)
{
    if (!ctx) return RBM_ERR_INVALID;
    if (ctx->nccl_comm) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        xch_release(ctx);
        g_nccl.destroy(ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
    ctx->nranks = 1;
    ctx->rank = 0;
    return RBM_OK;
}

static size_t xch_bytes()
{
    return sizeof(double) * 2 * rbm_ctx::XCH_MAX_RANKS * rbm_ctx::XCH_CAPR + sizeof(unsigned long long) * (2 * rbm_ctx::XCH_MAX_RANKS + 2);
}

extern "C" int rbm_comm_peer_handle(rbm_ctx *ctx, void *handle64)
{
    if (!ctx || !handle64) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_comm_peer_handle: NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
    RBM_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->xch_local) {
        RBM_CUDA(ctx, cudaMalloc(&ctx->xch_local, xch_bytes()));
        RBM_CUDA(ctx, cudaMemset(ctx->xch_local, 0, xch_bytes()));
    }
    cudaIpcMemHandle_t h;
    RBM_CUDA(ctx, cudaIpcGetMemHandle(&h, ctx->xch_local));
    memcpy(handle64, &h, sizeof(h));
    return RBM_OK;
}

extern "C" int rbm_comm_peer_attach(rbm_ctx *ctx, const void *handles)
{
    if (!ctx) return RBM_ERR_INVALID;
    if (!handles) {   // switch the fast path off (e.g. another rank could not map its peers)
        ctx->xch_ready = false;
        return RBM_OK;
    }
    if (!ctx->nccl_comm || ctx->nranks < 2) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_comm_peer_attach: call rbm_comm_init first");
    if (ctx->nranks > rbm_ctx::XCH_MAX_RANKS) return rbm_fail(ctx, RBM_ERR_UNSUPPORTED, "rbm_comm_peer_attach: more than %d ranks", rbm_ctx::XCH_MAX_RANKS);
    if (!ctx->xch_local) return rbm_fail(ctx, RBM_ERR_INVALID, "rbm_comm_peer_attach: call rbm_comm_peer_handle first");
    RBM_CUDA(ctx, cudaSetDevice(ctx->device));
    for (int r = 0; r < ctx->nranks; ++r) {
        if (r == ctx->rank) {
            ctx->xch_peer[r] = ctx->xch_local;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)handles + 64 * (size_t)r, sizeof(h));
        cudaError_t e = cudaIpcOpenMemHandle(&ctx->xch_peer[r], h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return rbm_fail(ctx, RBM_ERR_COMM, "cudaIpcOpenMemHandle(rank %d): %s (the NCCL path stays in use)", r, cudaGetErrorString(e));
        }
    }
    ctx->xch_ready = true;
    return RBM_OK;
}

static void xch_release(rbm_ctx *ctx)
{
    for (int r = 0; r < rbm_ctx::XCH_MAX_RANKS; ++r) {
        if (ctx->xch_peer[r] && ctx->xch_peer[r] != ctx->xch_local) cudaIpcCloseMemHandle(ctx->xch_peer[r]);
        ctx->xch_peer[r] = nullptr;
    }
    if (ctx->xch_local) cudaFree(ctx->xch_local);
    ctx->xch_local = nullptr;
    ctx->xch_ready = false;
}

int rbm_allreduce_sum_f64(rbm_ctx *ctx, double *dev, size_t count)
{
    if (!ctx->nccl_comm || ctx->nranks == 1 || count == 0) return RBM_OK;
    int rc = g_nccl.all_reduce(dev, dev, count, NCCL_F64, NCCL_SUM, ctx->nccl_comm, ctx->stream);
    if (rc) return rbm_fail(ctx, RBM_ERR_COMM, "ncclAllReduce: %s", nccl_err(rc));
    return RBM_OK;
}

__global__ void k_argmax_pairs(const double *pairs, int n, double *out)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double bv = pairs[0], bi = pairs[1];
        for (int r = 1; r < n; ++r) {
            double v = pairs[2 * r], i = pairs[2 * r + 1];
            if (v > bv || (v == bv && i < bi)) {
                bv = v;
                bi = i;
            }
        }
        out[0] = bv;
        out[1] = bi;
    }
}

int rbm_allreduce_argmax(rbm_ctx *ctx, double *pair_dev)
{
    if (!ctx->nccl_comm || ctx->nranks == 1) return RBM_OK;
    DevBuf all;
    RBM_CUDA(ctx, all.alloc(sizeof(double) * 2 * ctx->nranks));
    int rc = g_nccl.all_gather(pair_dev, all.p, 2, NCCL_F64, ctx->nccl_comm, ctx->stream);
    if (rc) return rbm_fail(ctx, RBM_ERR_COMM, "ncclAllGather: %s", nccl_err(rc));
    k_argmax_pairs<<<1, 32, 0, ctx->stream>>>(all.as<double>(), ctx->nranks, pair_dev);
    RBM_LAUNCH_CHECK(ctx);
    RBM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `all` is freed on return
    return RBM_OK;
}

extern "C" void rbm_finalize(rbm_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    prof_clear(ctx);
    rbm_comm_destroy(ctx);
    rbm_pod_release(ctx);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}
