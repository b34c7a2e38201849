// common.cuh -- context, error plumbing and pointer staging shared by the translation
// units of librbm_b200.so.  sm_100a only; there is no CPU fallback anywhere.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/rbm_b200.h"

struct rbm_ctx {
    int device = 0;
    int sm_count = 0;
    int smem_optin = 0;        // max opt-in dynamic shared memory per block (bytes)
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;  // stream all work is issued on
    std::string err;
    int64_t launches = 0;
    // communicator (NCCL, resolved at run time)
    void *nccl_comm = nullptr;
    int nranks = 1, rank = 0;
    // Peer-memory exchange (NVLink P2P through CUDA IPC) for the one latency-bound collective of the path:
    // the per-step sum of the charge-density vector over ranks.  xch_local is this rank's buffer
    // [2][16][XCH_CAPR] doubles + [2][16] sequence flags + status word + epoch word (layout in pic_kernels.cuh);
    // xch_peer[r] maps rank r's.
    static constexpr int XCH_MAX_RANKS = 16;    // == rbm::XCH_R
    static constexpr int XCH_CAPR = 16384;      // == rbm::XCH_CAPR, doubles per (buffer, source rank): 256 samples x 64 bins
    void *xch_local = nullptr;
    void *xch_peer[XCH_MAX_RANKS] = {};
    bool xch_ready = false;
    // cuSOLVER handle, created lazily by the POD path
    void *cusolver = nullptr;
    // split-K workspace of the GEMM kernel, kept between calls (cudaMalloc/cudaFree are slow once peer mappings exist)
    void *gemm_work = nullptr;
    size_t gemm_work_bytes = 0;
    // optional per-kernel timing (rbm_profile_*): event pairs around every launch of a named kernel
    bool profiling = false;
    struct ProfRec { const char *name; cudaEvent_t beg, end; };
    std::vector<ProfRec> prof;
};

// Wraps one kernel launch in a CUDA-event pair when profiling is on (no-op otherwise).
struct ProfScope {
    rbm_ctx *ctx;
    cudaEvent_t end = nullptr;
    ProfScope(rbm_ctx *c, const char *name) : ctx(c)
    {
        if (!c->profiling) return;
        cudaEvent_t b = nullptr;
        if (cudaEventCreate(&b) != cudaSuccess || cudaEventCreate(&end) != cudaSuccess) { end = nullptr; return; }
        cudaEventRecord(b, c->stream);
        c->prof.push_back({name, b, end});
    }
    ~ProfScope()
    {
        if (end) cudaEventRecord(end, ctx->stream);
    }
};

int rbm_fail(rbm_ctx *ctx, int code, const char *fmt, ...);

#define RBM_CUDA(ctx, call)                                                                  \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            cudaGetLastError(); /* a non-sticky error must not surface again in a later, unrelated call */ \
            return rbm_fail((ctx), e__ == cudaErrorMemoryAllocation ? RBM_ERR_NOMEM          \
                                                                    : RBM_ERR_CUDA,          \
                            "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),         \
                            __FILE__, __LINE__);                                             \
        }                                                                                    \
    } while (0)

#define RBM_TRY(expr)                  \
    do {                               \
        int rc__ = (expr);             \
        if (rc__ != RBM_OK) return rc__; \
    } while (0)

#define RBM_LAUNCH_CHECK(ctx)                         \
    do {                                              \
        (ctx)->launches++;                            \
        RBM_CUDA((ctx), cudaPeekAtLastError());       \
    } while (0)

// true when p is device-accessible memory on this context's GPU (device or managed)
bool rbm_is_device_ptr(const void *p);

// RAII device buffer (freed on scope exit; stream-ordered frees are not used so that
// lifetime is independent of the user's stream)
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    cudaError_t alloc(size_t n) {
        release();
        bytes = n;
        if (n == 0) return cudaSuccess;
        return cudaMalloc(&p, n);
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    // persistent workspace: (re)allocate only when too small; *grew reports a reallocation (pointer changed)
    cudaError_t reserve(size_t n, bool *grew = nullptr) {
        if (grew) *grew = false;
        if (p && bytes >= n) return cudaSuccess;
        if (grew) *grew = true;
        return alloc(n);
    }
    template <class T> T *as() const { return static_cast<T *>(p); }
};

// An input array that must be readable by kernels: either the caller's device pointer
// or a staged device copy of the caller's host array.
struct DevIn {
    DevBuf buf;
    const void *ptr = nullptr;
    // persist != NULL: a host array is staged in that long-lived buffer (grown on demand) instead of a per-call cudaMalloc
    int stage(rbm_ctx *ctx, const void *src, size_t bytes, DevBuf *persist = nullptr);
    template <class T> const T *as() const { return static_cast<const T *>(ptr); }
};

// An output array written by kernels: the caller's device pointer, or a device staging
// buffer that flush() copies back to the caller's host array.
struct DevOut {
    DevBuf buf;
    void *ptr = nullptr;
    void *host = nullptr;
    size_t bytes = 0;
    int prepare(rbm_ctx *ctx, void *dst, size_t nbytes, bool zero = false, DevBuf *persist = nullptr);
    int flush(rbm_ctx *ctx);  // async on ctx->stream; caller synchronises
    template <class T> T *as() const { return static_cast<T *>(ptr); }
};

// releases the cuSOLVER handle of the POD path (pod.cu)
void rbm_pod_release(rbm_ctx *ctx);

// Dense FP64 DMMA GEMM on device matrices described by arrays of column pointers (pod.cu):
//   trans_a: D (M x N) = A' B, contraction over the K rows of A (M columns) and B (N columns);
//   else   : D (M x N) = A B with A given by its K columns (length M) and B by its N columns (length K).
// rbm_make_cols builds the device array of column pointers of a column-major matrix.
int rbm_make_cols(rbm_ctx *ctx, const double *base, int64_t ld, int64_t ncols, DevBuf &out, bool *vec_ok);
int rbm_run_gemm(rbm_ctx *ctx, bool trans_a, const double *const *acol, const double *const *bcol, int64_t M, int64_t N,
                 int64_t K, bool vec, double *D, int64_t ldd, const double *a_base = nullptr, int64_t a_ld = 0);
// a_base/a_ld (A B only): A is one column-major matrix with column k at a_base + k*a_ld (skips the pointer array)
// Reduced model: deposit of X = A Z without storing X (recon.cu); RBM_ERR_UNSUPPORTED = shape not suited, nothing launched
namespace rbm { struct Geom; }
int rbm_recon_deposit_max_chunks(const rbm_ctx *ctx);
int rbm_recon_deposit(rbm_ctx *ctx, const rbm::Geom &g, const double *A, int64_t lda, const double *Z, int64_t ldz, int64_t N,
                      int S, int K, double *part, int *nchunks);
// inverse of a small dense matrix (n x n, column-major, device) through cuSOLVER LU: Ainv = A^-1 (A is destroyed)
int rbm_invert(rbm_ctx *ctx, double *A, int n, double *Ainv);

// NCCL all-reduce (sum, f64) in place on a device buffer; no-op without a communicator.
int rbm_allreduce_sum_f64(rbm_ctx *ctx, double *dev, size_t count);
// all-reduce of (value, index) pairs keeping the maximum value (ties: smallest index)
int rbm_allreduce_argmax(rbm_ctx *ctx, double *val_idx_pair_dev);
