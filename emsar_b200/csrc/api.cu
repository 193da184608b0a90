// Context management and error reporting of libemsar_cuda.so.
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_err[1024] = "";

void emsar_set_err(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *emsar_cuda_last_error(void) { return g_err; }

extern "C" const char *emsar_cuda_strerror(int status)
{
    switch (status) {
    case EMSAR_OK: return "ok";
    case EMSAR_ERR_NO_DEVICE: return "no usable sm_100 CUDA device (there is no CPU fallback)";
    case EMSAR_ERR_CUDA: return "CUDA runtime error";
    case EMSAR_ERR_BAD_ARG: return "bad argument";
    case EMSAR_ERR_BAD_INDEX: return "malformed rsh index (not in scan order)";
    case EMSAR_ERR_UNSUPPORTED: return "documented limit exceeded";
    case EMSAR_ERR_STATE: return "call order violated";
    case EMSAR_ERR_NOMEM: return "out of memory";
    case EMSAR_ERR_COMM: return "multi-GPU communication error";
    default: return "unknown status";
    }
}

// the context whose stream / memory pool the allocation helpers of this host thread use (set by every API entry)
thread_local emsar_ctx *g_cur_ctx = nullptr;

int ctx_use(emsar_ctx *ctx)
{
    CU(cudaSetDevice(ctx->device));
    g_cur_ctx = ctx;
    return EMSAR_OK;
}

int dev_alloc_bytes(void **p, size_t bytes)
{
    emsar_ctx *ctx = g_cur_ctx;
    if (bytes == 0) bytes = 1;
    cudaError_t e = (ctx && ctx->pool) ? cudaMallocFromPoolAsync(p, bytes, (cudaMemPool_t)ctx->pool, ctx->stream) : cudaMalloc(p, bytes);
    if (e != cudaSuccess) { *p = nullptr; emsar_set_err("device allocation of %zu bytes: %s", bytes, cudaGetErrorString(e)); cudaGetLastError(); return EMSAR_ERR_NOMEM; }
    return EMSAR_OK;
}

void dev_free(void *p)
{
    if (!p) return;
    emsar_ctx *ctx = g_cur_ctx;
    if (ctx && ctx->pool) cudaFreeAsync(p, ctx->stream);       // ordered after everything already enqueued on the stream
    else cudaFree(p);
}

extern "C" int emsar_cuda_open(int device, emsar_ctx **out)
{
    CHECK_ARG(out != nullptr, "emsar_cuda_open: ctx is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        emsar_set_err("emsar_cuda_open: no CUDA device (%s); this library has no CPU path", cudaGetErrorString(e));
        return EMSAR_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n) { emsar_set_err("emsar_cuda_open: device %d out of range (0..%d)", device, n - 1); return EMSAR_ERR_BAD_ARG; }
    CU(cudaSetDevice(device));
    emsar_ctx *ctx = new emsar_ctx();
    memset(ctx, 0, sizeof(*ctx));
    ctx->device = device;
    CU(cudaGetDeviceProperties(&ctx->prop, device));
    if (ctx->prop.major != 10) {
        emsar_set_err("emsar_cuda_open: device %d is sm_%d%d; this library carries sm_100a code only", device, ctx->prop.major, ctx->prop.minor);
        delete ctx;
        return EMSAR_ERR_NO_DEVICE;
    }
    CU(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    {   // stream-ordered allocations from a pool that never trims: the per-sample arrays of a -M list cost a cudaMalloc once
        cudaMemPoolProps pp;
        memset(&pp, 0, sizeof(pp));
        pp.allocType = cudaMemAllocationTypePinned;
        pp.handleTypes = cudaMemHandleTypeNone;
        pp.location.type = cudaMemLocationTypeDevice;
        pp.location.id = device;
        cudaMemPool_t pool = nullptr;
        if (!getenv("EMSAR_NO_POOL") && cudaMemPoolCreate(&pool, &pp) == cudaSuccess) {
            unsigned long long keep = ~0ULL;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            ctx->pool = pool;
        }
        cudaGetLastError();
    }
    g_cur_ctx = ctx;
    CU(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 9; i++) CU(cudaEventCreateWithFlags(&ctx->copy_ev[i], cudaEventDisableTiming));
    CU(cudaEventCreate(&ctx->ev0));
    CU(cudaEventCreate(&ctx->ev1));
    CU(cudaEventCreate(&ctx->tev0));
    CU(cudaEventCreate(&ctx->tev1));
    // [0..63] scalars of the EM kernel (delta slots, iteration count), then one 128-byte barrier line per CTA
    const size_t bar_bytes = 256 + (size_t)ctx->prop.multiProcessorCount * 4 * 128;
    CU(cudaMalloc(&ctx->d_barrier, bar_bytes));
    CU(cudaMemset(ctx->d_barrier, 0, bar_bytes));
    // L2 persistence carve-out for theta|q (north star: keep the theta vector L2-resident)
    size_t want = (size_t)ctx->prop.persistingL2CacheMaxSize;
    if (want > (size_t)64 << 20) want = (size_t)64 << 20;
    if (want > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) ctx->l2_persist_bytes = want;
    cudaGetLastError();
    int rc = em_query_occupancy(ctx);
    if (rc != EMSAR_OK) return rc;
    *out = ctx;
    return EMSAR_OK;
}

extern "C" int emsar_cuda_close(emsar_ctx *ctx)
{
    if (!ctx) return EMSAR_OK;
    ctx_use(ctx);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->d_barrier);
    cudaFree(ctx->d_scratch);
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    cudaEventDestroy(ctx->tev0);
    cudaEventDestroy(ctx->tev1);
    for (int i = 0; i < 9; i++) if (ctx->copy_ev[i]) cudaEventDestroy(ctx->copy_ev[i]);
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
    if (ctx->pool) cudaMemPoolDestroy((cudaMemPool_t)ctx->pool);
    cudaStreamDestroy(ctx->stream);
    if (g_cur_ctx == ctx) g_cur_ctx = nullptr;
    delete ctx;
    return EMSAR_OK;
}

extern "C" int emsar_cuda_launch_count(emsar_ctx *ctx, int64_t *launches)
{
    CHECK_ARG(ctx && launches, "emsar_cuda_launch_count: NULL argument");
    *launches = ctx->launches;
    return EMSAR_OK;
}

extern "C" int emsar_cuda_synchronize(emsar_ctx *ctx)
{
    CHECK_ARG(ctx, "emsar_cuda_synchronize: NULL ctx");
    TRY(ctx_use(ctx));
    CU(cudaStreamSynchronize(ctx->stream));
    return EMSAR_OK;
}

extern "C" int emsar_cuda_timer_start(emsar_ctx *ctx)
{
    CHECK_ARG(ctx, "emsar_cuda_timer_start: NULL ctx");
    TRY(ctx_use(ctx));
    CU(cudaEventRecord(ctx->tev0, ctx->stream));
    return EMSAR_OK;
}

extern "C" int emsar_cuda_timer_stop(emsar_ctx *ctx, double *elapsed_ms)
{
    CHECK_ARG(ctx && elapsed_ms, "emsar_cuda_timer_stop: NULL argument");
    TRY(ctx_use(ctx));
    CU(cudaEventRecord(ctx->tev1, ctx->stream));
    CU(cudaEventSynchronize(ctx->tev1));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, ctx->tev0, ctx->tev1));
    *elapsed_ms = ms;
    return EMSAR_OK;
}

extern "C" int emsar_cuda_device_info(emsar_ctx *ctx, emsar_device_info *info)
{
    CHECK_ARG(ctx && info, "emsar_cuda_device_info: NULL argument");
    memset(info, 0, sizeof(*info));
    info->sm_count = ctx->prop.multiProcessorCount;
    info->cc_major = ctx->prop.major;
    info->cc_minor = ctx->prop.minor;
    info->l2_bytes = ctx->prop.l2CacheSize;
    info->hbm_bytes = (int64_t)ctx->prop.totalGlobalMem;
    info->em_blocks_per_sm = ctx->em_blocks_per_sm;
    info->em_block_threads = EM_BLOCK;
    strncpy(info->name, ctx->prop.name, sizeof(info->name) - 1);
    return EMSAR_OK;
}

int ctx_scratch(emsar_ctx *ctx, size_t bytes, void **p)
{
    if (bytes > ctx->scratch_bytes) {
        CU(cudaStreamSynchronize(ctx->stream));
        if (ctx->d_scratch) cudaFree(ctx->d_scratch);
        ctx->d_scratch = nullptr;
        ctx->scratch_bytes = 0;
        size_t want = bytes + (bytes >> 2) + 4096;
        cudaError_t e = cudaMalloc(&ctx->d_scratch, want);
        if (e != cudaSuccess) { emsar_set_err("scratch cudaMalloc(%zu): %s", want, cudaGetErrorString(e)); return EMSAR_ERR_NOMEM; }
        ctx->scratch_bytes = want;
    }
    *p = ctx->d_scratch;
    return EMSAR_OK;
}
