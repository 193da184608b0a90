// Multi-GPU plumbing for class-sharded samples: an NCCL communicator per context. libnccl is loaded with dlopen so that
// libemsar_cuda.so itself has no link-time dependency on it (inside a PyTorch process the already loaded copy is reused).
#include <dlfcn.h>

#include "common.cuh"

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclSum = 0 };
enum { ncclInt32 = 2, ncclFloat64 = 8 };

static struct {
    void *h;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
    const char *(*GetErrorString)(ncclResult_t);
} N;

static int nccl_load()
{
    if (N.h) return EMSAR_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
    for (int i = 0; names[i] && !N.h; i++) N.h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!N.h) { emsar_set_err("cannot load libnccl.so.2: %s", dlerror()); return EMSAR_ERR_COMM; }
    N.GetUniqueId = (ncclResult_t(*)(ncclUniqueId *))dlsym(N.h, "ncclGetUniqueId");
    N.CommInitRank = (ncclResult_t(*)(ncclComm_t *, int, ncclUniqueId, int))dlsym(N.h, "ncclCommInitRank");
    N.CommDestroy = (ncclResult_t(*)(ncclComm_t))dlsym(N.h, "ncclCommDestroy");
    N.AllReduce = (ncclResult_t(*)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(N.h, "ncclAllReduce");
    N.GetErrorString = (const char *(*)(ncclResult_t))dlsym(N.h, "ncclGetErrorString");
    if (!N.GetUniqueId || !N.CommInitRank || !N.CommDestroy || !N.AllReduce || !N.GetErrorString) {
        emsar_set_err("libnccl lacks an expected symbol");
        N.h = nullptr;
        return EMSAR_ERR_COMM;
    }
    return EMSAR_OK;
}

#define NC(call)                                                                                   \
    do {                                                                                           \
        ncclResult_t r_ = (call);                                                                  \
        if (r_ != 0) { emsar_set_err("%s -> %s", #call, N.GetErrorString(r_)); return EMSAR_ERR_COMM; } \
    } while (0)

extern "C" int emsar_comm_unique_id(uint8_t id[128])
{
    CHECK_ARG(id, "emsar_comm_unique_id: NULL id");
    TRY(nccl_load());
    ncclUniqueId u;
    NC(N.GetUniqueId(&u));
    memcpy(id, u.internal, 128);
    return EMSAR_OK;
}

extern "C" int emsar_comm_init(emsar_ctx *ctx, int32_t rank, int32_t nranks, const uint8_t id[128])
{
    CHECK_ARG(ctx && id && nranks >= 1 && rank >= 0 && rank < nranks, "emsar_comm_init: bad argument");
    TRY(nccl_load());
    CU(cudaSetDevice(ctx->device));
    if (ctx->nccl_comm) { emsar_set_err("emsar_comm_init: the context already has a communicator"); return EMSAR_ERR_STATE; }
    ncclUniqueId u;
    memcpy(u.internal, id, 128);
    ncclComm_t c = nullptr;
    NC(N.CommInitRank(&c, nranks, u, rank));
    ctx->nccl_comm = c; ctx->rank = rank; ctx->nranks = nranks;
    return EMSAR_OK;
}

extern "C" int emsar_comm_destroy(emsar_ctx *ctx)
{
    if (!ctx || !ctx->nccl_comm) return EMSAR_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    N.CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr; ctx->nranks = 0; ctx->rank = 0;
    return EMSAR_OK;
}

int comm_allreduce_f64(emsar_ctx *ctx, const double *in, double *out, size_t n)
{
    if (!ctx->nccl_comm) { emsar_set_err("no communicator: call emsar_comm_init first"); return EMSAR_ERR_STATE; }
    NC(N.AllReduce(in, out, n, ncclFloat64, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    return EMSAR_OK;
}

int comm_allreduce_i32(emsar_ctx *ctx, int32_t *inout, size_t n)
{
    if (!ctx->nccl_comm) { emsar_set_err("no communicator: call emsar_comm_init first"); return EMSAR_ERR_STATE; }
    NC(N.AllReduce(inout, inout, n, ncclInt32, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
    return EMSAR_OK;
}

extern "C" int emsar_sample_counts_allreduce(emsar_sample *s)
{
    CHECK_ARG(s, "emsar_sample_counts_allreduce: NULL sample");
    emsar_index *ix = s->index;
    CU(cudaSetDevice(s->ctx->device));
    TRY(comm_allreduce_i32(s->ctx, s->d_R, (size_t)ix->C));
    TRY(comm_allreduce_i32(s->ctx, s->d_hist, (size_t)ix->max_fl + 1));
    TRY(comm_allreduce_i32(s->ctx, s->d_flags, 1));         // error flags (bit-or would be exact; a sum stays non-zero)
    CU(cudaStreamSynchronize(s->ctx->stream));
    s->have_counts = true; s->prepared = false;
    return EMSAR_OK;
}

// rank r gets items out[r] .. out[r+1]; weight_prefix[i] = total weight of items 0..i-1 (n+1 entries)
extern "C" int emsar_shard_ranges(int64_t n, const int64_t *weight_prefix, int32_t nranks, int64_t *out)
{
    CHECK_ARG(n >= 0 && weight_prefix && nranks >= 1 && out, "emsar_shard_ranges: bad argument");
    const int64_t total = weight_prefix[n];
    out[0] = 0;
    for (int r = 1; r < nranks; r++) {
        const int64_t target = (int64_t)((__int128)total * r / nranks);
        int64_t lo = out[r - 1], hi = n;
        while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (weight_prefix[mid] >= target) hi = mid; else lo = mid + 1; }
        out[r] = lo;
    }
    out[nranks] = n;
    return EMSAR_OK;
}
