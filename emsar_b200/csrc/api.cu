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
    CU(cudaEventCreate(&ctx->ev0));
    CU(cudaEventCreate(&ctx->ev1));
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
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->d_barrier);
    cudaFree(ctx->d_scratch);
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    cudaStreamDestroy(ctx->stream);
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
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
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
        if (ctx->d_scratch) CU(cudaFree(ctx->d_scratch));
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
