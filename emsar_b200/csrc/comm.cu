// Multi-GPU plumbing for class-sharded samples: an NCCL communicator per context. libnccl is loaded with dlopen so that
// libemsar_cuda.so itself has no link-time dependency on it (inside a PyTorch process the already loaded copy is reused).
#include <dlfcn.h>
#include <unistd.h>

#include "common.cuh"

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclSum = 0 };
enum { ncclInt8 = 0, ncclInt32 = 2, ncclFloat64 = 8 };

static struct {
    void *h;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t);
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
    N.AllGather = (ncclResult_t(*)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t))dlsym(N.h, "ncclAllGather");
    N.GetErrorString = (const char *(*)(ncclResult_t))dlsym(N.h, "ncclGetErrorString");
    if (!N.GetUniqueId || !N.AllGather || !N.CommInitRank || !N.CommDestroy || !N.AllReduce || !N.GetErrorString) {
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
    CHECK_ARG(ctx && id && nranks >= 1 && nranks <= EMSAR_MAX_RANKS && rank >= 0 && rank < nranks, "emsar_comm_init: bad argument (1..8 ranks)");
    TRY(nccl_load());
    TRY(ctx_use(ctx));
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
    ctx_use(ctx);
    cudaStreamSynchronize(ctx->stream);
    comm_window_release(ctx);
    N.CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr; ctx->nranks = 0; ctx->rank = 0;
    return EMSAR_OK;
}

extern "C" int emsar_comm_info(emsar_ctx *ctx, int32_t *rank, int32_t *nranks, int32_t *peer_memory)
{
    CHECK_ARG(ctx, "emsar_comm_info: NULL context");
    if (rank) *rank = ctx->rank;
    if (nranks) *nranks = ctx->nccl_comm ? ctx->nranks : 0;
    if (peer_memory) *peer_memory = ctx->win_state;
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

// stream-ordered barrier over the ranks: every rank's earlier work on its stream is complete on every rank's device
// before anything enqueued after it starts (an all-reduce cannot finish before every rank has joined it)
int comm_barrier(emsar_ctx *ctx)
{
    return comm_allreduce_i32(ctx, (int32_t *)(ctx->d_barrier + 60), 1);
}

// ---- peer-memory window of the fused sharded EM kernel ------------------------------------------------------
// Every rank allocates one block, publishes {IPC handle, pid, raw pointer, device} through an NCCL all-gather and maps the
// other ranks' blocks: cudaIpcOpenMemHandle between processes (bench.py / tests: one process per GPU), plain peer access
// between the contexts of one process (the emsar CLI: one host thread per GPU). If any rank cannot map any peer, every
// rank agrees (all-reduce of the status) to stay on the NCCL path.
struct WinInfo {
    cudaIpcMemHandle_t handle;      // 64 bytes
    long long pid;
    unsigned long long ptr;
    int device;
    int ok;
    char pad[128 - 64 - 8 - 8 - 4 - 4];
};
static_assert(sizeof(WinInfo) == 128, "WinInfo is exchanged as 128 raw bytes");

void comm_window_release(emsar_ctx *ctx)
{
    for (int r = 0; r < EMSAR_MAX_RANKS; r++) {
        if (ctx->peer_win[r] && ctx->peer_ipc[r]) cudaIpcCloseMemHandle(ctx->peer_win[r]);
        ctx->peer_win[r] = nullptr; ctx->peer_ipc[r] = false;
    }
    if (ctx->win) cudaFree(ctx->win);
    ctx->win = nullptr; ctx->win_rows = 0; ctx->win_state = 0; ctx->win_bytes = 0;
    cudaGetLastError();
}

// `min_bytes`: the window must hold at least this many bytes (k_em_psum lays its own slot arrays out in it); `rows`: capacity in rows of
// the legacy fused kernel's layout. Collective: every rank passes the same numbers.
int comm_window_ensure(emsar_ctx *ctx, int64_t rows, size_t min_bytes)
{
    if (!ctx->nccl_comm) { emsar_set_err("no communicator: call emsar_comm_init first"); return EMSAR_ERR_STATE; }
    if (ctx->win_state == -1) return EMSAR_OK;
    if (ctx->win_state == 1 && rows <= ctx->win_rows && min_bytes <= ctx->win_bytes) return EMSAR_OK;
    const char *mode = getenv("EMSAR_SHARD_MODE");
    if (mode && !strcmp(mode, "nccl")) { ctx->win_state = -1; return EMSAR_OK; }     // every rank reads the same environment
    const int R = ctx->nranks, me = ctx->rank;
    cudaStream_t st = ctx->stream;
    CU(cudaStreamSynchronize(st));
    TRY(comm_barrier(ctx));                     // nobody is still inside a kernel that writes into a window being replaced
    CU(cudaStreamSynchronize(st));
    comm_window_release(ctx);
    const int64_t cap = rows + (rows >> 3) + 1024;
    size_t bytes = WIN_HDR_BYTES + 32 * (size_t)cap + 16 * (size_t)(win_slice_rows(cap, R) * R + 64);       // dm | 2 x theta | xbuf
    if (bytes < min_bytes + (min_bytes >> 3)) bytes = min_bytes + (min_bytes >> 3);
    ctx->win_bytes = bytes;
    WinInfo mine;
    memset(&mine, 0, sizeof(mine));
    mine.pid = (long long)getpid(); mine.device = ctx->device;
    if (cudaMalloc(&ctx->win, bytes) == cudaSuccess && cudaMemset(ctx->win, 0, bytes) == cudaSuccess &&
        cudaIpcGetMemHandle(&mine.handle, ctx->win) == cudaSuccess) {
        mine.ok = 1; mine.ptr = (unsigned long long)(uintptr_t)ctx->win;
    }
    cudaGetLastError();
    WinInfo *d_info = nullptr;
    TRY(dev_alloc(&d_info, (size_t)R + 1));
    std::vector<WinInfo> all((size_t)R);
    CU(cudaMemcpyAsync(d_info + R, &mine, sizeof(mine), cudaMemcpyHostToDevice, st));
    NC(N.AllGather(d_info + R, d_info, sizeof(WinInfo), ncclInt8, (ncclComm_t)ctx->nccl_comm, st));
    CU(cudaMemcpyAsync(all.data(), d_info, sizeof(WinInfo) * (size_t)R, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    int32_t good = mine.ok;
    for (int r = 0; r < R && good; r++) {
        if (!all[(size_t)r].ok) { good = 0; break; }
        if (r == me) { ctx->peer_win[r] = ctx->win; continue; }
        if (all[(size_t)r].pid == mine.pid) {
            // another context of this process: its pointer is valid here once peer access is on
            int can = 0;
            cudaDeviceCanAccessPeer(&can, ctx->device, all[(size_t)r].device);
            cudaError_t e = can ? cudaDeviceEnablePeerAccess(all[(size_t)r].device, 0) : cudaErrorPeerAccessUnsupported;
            if (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled) ctx->peer_win[r] = (void *)(uintptr_t)all[(size_t)r].ptr;
            else good = 0;
        } else {
            void *p = nullptr;
            if (cudaIpcOpenMemHandle(&p, all[(size_t)r].handle, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess) { ctx->peer_win[r] = p; ctx->peer_ipc[r] = true; }
            else good = 0;
        }
        cudaGetLastError();
    }
    // min over the ranks: 1 only if every rank mapped every window
    int32_t *d_flag = (int32_t *)d_info;
    int32_t neg = good ? 0 : 1;
    CU(cudaMemcpyAsync(d_flag, &neg, 4, cudaMemcpyHostToDevice, st));
    TRY(comm_allreduce_i32(ctx, d_flag, 1));
    CU(cudaMemcpyAsync(&neg, d_flag, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    dev_free(d_info);
    if (neg != 0) {
        comm_window_release(ctx);
        ctx->win_state = -1;
        if (getenv("EMSAR_VERBOSE")) fprintf(stderr, "emsar_cuda: rank %d: peer memory unavailable, class-sharded samples use the NCCL path\n", me);
        return EMSAR_OK;
    }
    ctx->win_rows = cap;
    ctx->win_state = 1;
    return EMSAR_OK;
}

extern "C" int emsar_sample_counts_allreduce(emsar_sample *s)
{
    CHECK_ARG(s, "emsar_sample_counts_allreduce: NULL sample");
    emsar_index *ix = s->index;
    TRY(ctx_use(s->ctx));
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
