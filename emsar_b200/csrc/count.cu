// Read -> rsh-class counting (bit-exact integers).
// Replaces update_ReadCounts (reference emsar_functions.c:838-943) with its two lookups
// update_rshbucket_single 'r' (:1528-1536) and update_rshbucket 'r' (:1597-1624, cmptarr :1677-1684),
// clear_readcounts_in_rshbucket (:1726-1752) and the FraglengthCounts / TotalReadCount bookkeeping (:940-941).
#include <limits.h>
#include <cub/cub.cuh>

#include "common.cuh"

constexpr int CNT_BLOCK = 256;
constexpr int CNT_WARPS = CNT_BLOCK / 32;
constexpr int CNT_HIST_SMEM = 4096;   // fragment-length bins kept in shared memory
constexpr int CNT_SMALL = 8;          // lists up to this length are sorted in registers by one thread

struct CountParams {
    int64_t n_reads;
    const int64_t *rd_ptr;
    const int32_t *rd_tid;
    const int32_t *rd_fl;
    int32_t T;
    const uint32_t *cls_off;
    const int32_t *cls_tid;
    const uint8_t *has_node;
    const unsigned long long *hash;
    uint64_t hash_mask;
    int32_t max_t_size, min_fl, max_fl;
    int32_t *R;
    int32_t *hist;
    int32_t *flags;
    int64_t n_tids;          // size of rd_tid when the caller states it (compact form: offsets are rebuilt from lengths), else INT64_MAX
};

#define CE(a, b) { int lo_ = min(v[a], v[b]); int hi_ = max(v[a], v[b]); v[a] = lo_; v[b] = hi_; }

// Probe for a sorted key of k tids held in registers; returns cid or -1.
__device__ __forceinline__ int64_t probe_small(const CountParams &p, const int (&v)[CNT_SMALL], int k)
{
    uint64_t sum = 0;
#pragma unroll
    for (int i = 0; i < CNT_SMALL; i++) if (i < k) sum += key_elem(i, v[i]);
    uint64_t h = key_finish(sum, k);
    uint32_t fp = (uint32_t)(h >> 32);
    uint64_t s = h & p.hash_mask;
    for (;;) {
        unsigned long long e = p.hash[s];
        if (e == 0ULL) return -1;
        if ((uint32_t)(e >> 32) == fp) {
            int64_t cid = (int64_t)(uint32_t)e - 1;
            uint32_t o = p.cls_off[cid];
            if ((int)(p.cls_off[cid + 1] - o) == k) {
                bool eq = true;
#pragma unroll
                for (int i = 0; i < CNT_SMALL; i++) if (i < k) eq = eq && (p.cls_tid[o + i] == v[i]);
                if (eq) return cid;
            }
        }
        s = (s + 1) & p.hash_mask;
    }
}

__global__ void __launch_bounds__(CNT_BLOCK) k_count(CountParams p)
{
    __shared__ int s_sort[CNT_WARPS][EMSAR_MAX_READ_TIDS];
    __shared__ int s_hist[CNT_HIST_SMEM];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool hist_in_smem = (p.max_fl + 1) <= CNT_HIST_SMEM;
    if (hist_in_smem) {
        for (int i = threadIdx.x; i <= p.max_fl; i += CNT_BLOCK) s_hist[i] = 0;
        __syncthreads();
    }
    const int64_t n_chunks = (p.n_reads + 31) / 32;
    for (int64_t chunk = (int64_t)blockIdx.x * CNT_WARPS + warp; chunk < n_chunks; chunk += (int64_t)gridDim.x * CNT_WARPS) {
        const int64_t r = chunk * 32 + lane;
        int64_t off = 0; int k = 0, fl = 0; bool ok = false;
        if (r < p.n_reads) {
            off = p.rd_ptr[r];
            int64_t k64 = p.rd_ptr[r + 1] - off;
            fl = p.rd_fl[r];
            if (k64 > EMSAR_MAX_READ_TIDS) { atomicOr(p.flags, 1); k64 = 0; }
            if (off + k64 > p.n_tids) { atomicOr(p.flags, 4); k64 = 0; }          // lengths and tid count of a compact batch disagree
            k = (int)k64;
            ok = k > 0 && fl <= p.max_fl && fl >= p.min_fl;          // :849
        }
        // fragment-length histogram + TotalReadCount (:940-941): one atomic per distinct length per warp
        {
            unsigned act = __ballot_sync(0xffffffffu, ok);
            if (ok) {
                unsigned same = __match_any_sync(act, fl);
                if (lane == __ffs(same) - 1) {
                    if (hist_in_smem) atomicAdd(&s_hist[fl], __popc(same)); else atomicAdd(&p.hist[fl], __popc(same));
                }
            }
        }
        if (ok && k == 1) {
            int t = p.rd_tid[off];
            if (t < 0 || t >= p.T) atomicOr(p.flags, 2);
            else if (p.has_node[t]) atomicAdd(&p.R[t], 1);            // :1530-1534, cid == tid
        } else if (ok && k <= CNT_SMALL) {
            int v[CNT_SMALL];
            bool bad = false;
#pragma unroll
            for (int i = 0; i < CNT_SMALL; i++) {
                v[i] = (i < k) ? p.rd_tid[off + i] : INT_MAX;
                if (i < k && (v[i] < 0 || v[i] >= p.T)) bad = true;
            }
            if (bad) atomicOr(p.flags, 2);
            else if (k <= p.max_t_size) {                             // :1599
                // 19-comparator sorting network for 8 keys (INT_MAX padding sinks to the end)
                CE(0, 2) CE(1, 3) CE(4, 6) CE(5, 7)
                CE(0, 4) CE(1, 5) CE(2, 6) CE(3, 7)
                CE(0, 1) CE(2, 3) CE(4, 5) CE(6, 7)
                CE(2, 4) CE(3, 5)
                CE(1, 4) CE(3, 6)
                CE(1, 2) CE(3, 4) CE(5, 6)
                int64_t cid = probe_small(p, v, k);
                if (cid >= 0) atomicAdd(&p.R[cid], 1);
            }
        }
        // long lists: the whole warp sorts one list at a time in shared memory
        unsigned longm = __ballot_sync(0xffffffffu, ok && k > CNT_SMALL);
        while (longm) {
            const int src = __ffs(longm) - 1;
            longm &= longm - 1;
            const int64_t o = __shfl_sync(0xffffffffu, off, src);
            const int kk = __shfl_sync(0xffffffffu, k, src);
            int *buf = s_sort[warp];
            int P = 32; while (P < kk) P <<= 1;
            bool bad = false;
            for (int i = lane; i < P; i += 32) {
                int t = (i < kk) ? p.rd_tid[o + i] : INT_MAX;
                if (i < kk && (t < 0 || t >= p.T)) bad = true;
                buf[i] = t;
            }
            bad = __any_sync(0xffffffffu, bad);
            __syncwarp();
            if (bad) { if (lane == 0) atomicOr(p.flags, 2); continue; }
            if (kk > p.max_t_size) continue;                          // :1599
            for (int size = 2; size <= P; size <<= 1)
                for (int stride = size >> 1; stride > 0; stride >>= 1) {
                    for (int i = lane; i < (P >> 1); i += 32) {
                        int lo = 2 * i - (i & (stride - 1));          // index with bit `stride` cleared
                        int hi = lo + stride;
                        bool up = ((lo & size) == 0);
                        int a = buf[lo], b = buf[hi];
                        if ((a > b) == up) { buf[lo] = b; buf[hi] = a; }
                    }
                    __syncwarp();
                }
            uint64_t sum = 0;
            for (int i = lane; i < kk; i += 32) sum += key_elem(i, buf[i]);
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
            const uint64_t h = key_finish(sum, kk);
            const uint32_t fp = (uint32_t)(h >> 32);
            uint64_t s = h & p.hash_mask;
            for (;;) {
                unsigned long long e = p.hash[s];
                if (e == 0ULL) break;
                if ((uint32_t)(e >> 32) == fp) {
                    int64_t cid = (int64_t)(uint32_t)e - 1;
                    uint32_t co = p.cls_off[cid];
                    if ((int)(p.cls_off[cid + 1] - co) == kk) {
                        bool eq = true;
                        for (int i = lane; i < kk; i += 32) eq = eq && (p.cls_tid[co + i] == buf[i]);
                        if (__all_sync(0xffffffffu, eq)) { if (lane == 0) atomicAdd(&p.R[cid], 1); break; }
                    }
                }
                s = (s + 1) & p.hash_mask;
            }
            __syncwarp();
        }
    }
    if (hist_in_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i <= p.max_fl; i += CNT_BLOCK) { int c = s_hist[i]; if (c) atomicAdd(&p.hist[i], c); }
    }
}

static int launch_count(emsar_sample *s, int64_t n_reads, const int64_t *d_ptr, const int32_t *d_tid, const int32_t *d_fl, int64_t n_tids = INT64_MAX)
{
    emsar_index *ix = s->index;
    emsar_ctx *ctx = s->ctx;
    if (n_reads <= 0) return EMSAR_OK;
    CountParams p;
    p.n_reads = n_reads; p.rd_ptr = d_ptr; p.rd_tid = d_tid; p.rd_fl = d_fl;
    p.T = ix->T; p.cls_off = ix->d_cls_off; p.cls_tid = ix->d_cls_tid; p.has_node = ix->d_has_node;
    p.hash = ix->d_hash; p.hash_mask = ix->hash_mask; p.max_t_size = ix->max_t_size; p.min_fl = ix->min_fl; p.max_fl = ix->max_fl;
    p.R = s->d_R; p.hist = s->d_hist; p.flags = s->d_flags; p.n_tids = n_tids;
    int64_t chunks = (n_reads + 31) / 32;
    int64_t blocks = (chunks + CNT_WARPS - 1) / CNT_WARPS;
    int64_t cap = (int64_t)ctx->prop.multiProcessorCount * 8;        // grid-stride: 8 CTAs (of 8 resident) per SM
    if (blocks > cap) blocks = cap;
    k_count<<<(unsigned)blocks, CNT_BLOCK, 0, ctx->stream>>>(p);
    LAUNCHED(ctx);
    CU(cudaGetLastError());
    s->have_counts = true;
    s->prepared = false;
    return EMSAR_OK;
}

static int grow(void **p, size_t *cap, size_t bytes, emsar_ctx *ctx)
{
    if (bytes <= *cap) return EMSAR_OK;
    (void)ctx;
    if (*p) dev_free(*p);
    *p = nullptr; *cap = 0;
    size_t want = bytes + (bytes >> 3) + 256;
    TRY(dev_alloc_bytes(p, want));
    *cap = want;
    return EMSAR_OK;
}

extern "C" int emsar_sample_count(emsar_sample *s, int64_t n_reads, const int64_t *read_ptr, const int32_t *read_tid,
                                  const int32_t *read_fraglen)
{
    CHECK_ARG(s && n_reads >= 0, "emsar_sample_count: bad argument");
    if (n_reads == 0) { s->have_counts = true; return EMSAR_OK; }   // an empty alignment file is still a sample
    CHECK_ARG(read_ptr && read_tid && read_fraglen, "emsar_sample_count: NULL read arrays");
    emsar_ctx *ctx = s->ctx;
    TRY(ctx_use(ctx));
    const int64_t base = read_ptr[0];
    const int64_t ntid = read_ptr[n_reads] - base;
    CHECK_ARG(ntid >= 0, "emsar_sample_count: read_ptr not monotone");
    TRY(grow(&s->d_rd_ptr, &s->cap_rd_ptr, (size_t)(n_reads + 1) * 8, ctx));
    TRY(grow(&s->d_rd_tid, &s->cap_rd_tid, (size_t)(ntid > 0 ? ntid : 1) * 4, ctx));
    TRY(grow(&s->d_rd_fl, &s->cap_rd_fl, (size_t)n_reads * 4, ctx));
    if (!s->count_ev[0]) for (int i = 0; i < 4; i++) CU(cudaEventCreateWithFlags(&s->count_ev[i], cudaEventDisableTiming));
    const int32_t *d_tid0 = (const int32_t *)s->d_rd_tid - base;     // offsets stay absolute: shift the tid base pointer instead of rewriting read_ptr
    constexpr int NCH = 8;
    if (n_reads >= (int64_t)NCH << 20 && ctx->copy_stream) {
        // a large batch: copy it in NCH pieces on a second stream and count every piece as soon as it has landed, so that the
        // counting kernel (3.4 ms per 30 M reads) hides behind the PCIe transfer (14 ms) instead of following it
        cudaStream_t cs = ctx->copy_stream;
        CU(cudaEventRecord(ctx->copy_ev[NCH], ctx->stream));               // staging buffers: allocated / no longer read by earlier work
        CU(cudaStreamWaitEvent(cs, ctx->copy_ev[NCH], 0));
        for (int c = 0; c < NCH; c++) {
            const int64_t r0 = n_reads * c / NCH, r1 = n_reads * (c + 1) / NCH;
            const int64_t t0 = read_ptr[r0] - base, t1 = read_ptr[r1] - base;
            const int64_t p0 = c == 0 ? r0 : r0 + 1;                      // entry r0 already went with the previous piece
            CU(cudaMemcpyAsync((int64_t *)s->d_rd_ptr + p0, read_ptr + p0, (size_t)(r1 + 1 - p0) * 8, cudaMemcpyHostToDevice, cs));
            if (t1 > t0) CU(cudaMemcpyAsync((int32_t *)s->d_rd_tid + t0, read_tid + base + t0, (size_t)(t1 - t0) * 4, cudaMemcpyHostToDevice, cs));
            CU(cudaMemcpyAsync((int32_t *)s->d_rd_fl + r0, read_fraglen + r0, (size_t)(r1 - r0) * 4, cudaMemcpyHostToDevice, cs));
            CU(cudaEventRecord(ctx->copy_ev[c], cs));
            CU(cudaStreamWaitEvent(ctx->stream, ctx->copy_ev[c], 0));
            TRY(launch_count(s, r1 - r0, (const int64_t *)s->d_rd_ptr + r0, d_tid0, (const int32_t *)s->d_rd_fl + r0));
        }
        CU(cudaEventRecord(s->count_ev[s->count_seq & 3], cs));            // the host arrays are free once the copies are done
        s->count_seq++;
        return EMSAR_OK;
    }
    CU(cudaMemcpyAsync(s->d_rd_ptr, read_ptr, (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (ntid > 0) CU(cudaMemcpyAsync(s->d_rd_tid, read_tid + base, (size_t)ntid * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(s->d_rd_fl, read_fraglen, (size_t)n_reads * 4, cudaMemcpyHostToDevice, ctx->stream));
    // the host arrays are free again once these copies are done (matters for page-locked arrays: the copies are asynchronous)
    CU(cudaEventRecord(s->count_ev[s->count_seq & 3], ctx->stream));
    s->count_seq++;
    return launch_count(s, n_reads, (const int64_t *)s->d_rd_ptr, d_tid0, (const int32_t *)s->d_rd_fl);
}

// ---- compact wire form: lengths instead of offsets, 16-bit fragment lengths or none ----
__global__ void k_len_widen(int64_t n, const uint16_t *__restrict__ len, int64_t *__restrict__ out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) out[i] = i < n ? (int64_t)len[i] : 0;
}
__global__ void k_fl_widen(int64_t n, const uint16_t *__restrict__ fl, int32_t cfl, int32_t *__restrict__ out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = fl ? (int32_t)fl[i] : cfl;
}

extern "C" int emsar_sample_count_compact(emsar_sample *s, int64_t n_reads, int64_t n_tids, const uint16_t *read_len, const int32_t *read_tid,
                                          const uint16_t *read_fraglen, int32_t const_fraglen)
{
    CHECK_ARG(s && n_reads >= 0, "emsar_sample_count_compact: bad argument");
    if (n_reads == 0) { s->have_counts = true; return EMSAR_OK; }
    CHECK_ARG(read_len && read_tid, "emsar_sample_count_compact: NULL read arrays");
    CHECK_ARG(n_reads < ((int64_t)1 << 31), "emsar_sample_count_compact: more than 2^31 read groups in one batch");
    emsar_ctx *ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    TRY(ctx_use(ctx));
    CHECK_ARG(n_tids >= 0, "emsar_sample_count_compact: negative tid count");
    const int64_t ntid = n_tids;            // must be the sum of read_len; the counting kernel checks every list against it
    // staging: [int64 offsets n+1] [tids] [int32 fragment lengths n] as for the wide form, plus the 16-bit arrays and the scan scratch
    TRY(grow(&s->d_rd_ptr, &s->cap_rd_ptr, (size_t)(n_reads + 1) * 8, ctx));
    TRY(grow(&s->d_rd_tid, &s->cap_rd_tid, (size_t)(ntid > 0 ? ntid : 1) * 4, ctx));
    TRY(grow(&s->d_rd_fl, &s->cap_rd_fl, (size_t)n_reads * 4, ctx));
    size_t scan_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (int64_t *)nullptr, (int64_t *)nullptr, (int)(n_reads + 1));
    const size_t need = (size_t)(n_reads + 1) * 2 * 2 + (size_t)(n_reads + 1) * 8 + scan_bytes + 2048;
    TRY(grow(&s->d_rd_aux, &s->cap_rd_aux, need, ctx));
    if (!s->count_ev[0]) for (int i = 0; i < 4; i++) CU(cudaEventCreateWithFlags(&s->count_ev[i], cudaEventDisableTiming));
    char *aux = (char *)s->d_rd_aux;
    uint16_t *d_len = (uint16_t *)aux;                                   aux += (((size_t)(n_reads + 1) * 2 + 255) / 256) * 256;
    uint16_t *d_fl16 = (uint16_t *)aux;                                  aux += (((size_t)(n_reads + 1) * 2 + 255) / 256) * 256;
    int64_t *d_len64 = (int64_t *)aux;                                   aux += (((size_t)(n_reads + 1) * 8 + 255) / 256) * 256;
    void *d_scan = aux;
    CU(cudaMemcpyAsync(d_len, read_len, (size_t)n_reads * 2, cudaMemcpyHostToDevice, st));
    if (ntid > 0) CU(cudaMemcpyAsync(s->d_rd_tid, read_tid, (size_t)ntid * 4, cudaMemcpyHostToDevice, st));
    if (read_fraglen) CU(cudaMemcpyAsync(d_fl16, read_fraglen, (size_t)n_reads * 2, cudaMemcpyHostToDevice, st));
    CU(cudaEventRecord(s->count_ev[s->count_seq & 3], st));           // the host arrays are free once the copies are done
    s->count_seq++;
    const unsigned nb = (unsigned)((n_reads + 1 + 255) / 256);
    k_len_widen<<<nb, 256, 0, st>>>(n_reads, d_len, d_len64);
    CU(cub::DeviceScan::ExclusiveSum(d_scan, scan_bytes, d_len64, (int64_t *)s->d_rd_ptr, (int)(n_reads + 1), st));
    k_fl_widen<<<nb, 256, 0, st>>>(n_reads, read_fraglen ? d_fl16 : nullptr, const_fraglen, (int32_t *)s->d_rd_fl);
    ctx->launches += 3;
    return launch_count(s, n_reads, (const int64_t *)s->d_rd_ptr, (const int32_t *)s->d_rd_tid, (const int32_t *)s->d_rd_fl, ntid);
}

extern "C" int emsar_sample_count_wait(emsar_sample *s, int32_t lag)
{
    CHECK_ARG(s && lag >= 0 && lag < 4, "emsar_sample_count_wait: lag must be 0..3");
    if ((int64_t)s->count_seq - 1 - lag < 0) return EMSAR_OK;
    TRY(ctx_use(s->ctx));
    CU(cudaEventSynchronize(s->count_ev[(s->count_seq - 1 - (unsigned)lag) & 3]));
    return EMSAR_OK;
}

extern "C" int emsar_host_alloc(emsar_ctx *ctx, size_t bytes, void **p)
{
    CHECK_ARG(ctx && p, "emsar_host_alloc: NULL argument");
    TRY(ctx_use(ctx));
    cudaError_t e = cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocPortable);
    if (e != cudaSuccess) { *p = nullptr; emsar_set_err("cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e)); cudaGetLastError(); return EMSAR_ERR_NOMEM; }
    return EMSAR_OK;
}

extern "C" int emsar_host_free(emsar_ctx *ctx, void *p)
{
    CHECK_ARG(ctx, "emsar_host_free: NULL context");
    if (!p) return EMSAR_OK;
    TRY(ctx_use(ctx));
    CU(cudaFreeHost(p));
    return EMSAR_OK;
}

extern "C" int emsar_sample_count_device(emsar_sample *s, int64_t n_reads, const void *d_read_ptr, const void *d_read_tid,
                                         const void *d_read_fraglen)
{
    CHECK_ARG(s && n_reads >= 0, "emsar_sample_count_device: bad argument");
    if (n_reads == 0) { s->have_counts = true; return EMSAR_OK; }
    CHECK_ARG(d_read_ptr && d_read_tid && d_read_fraglen, "emsar_sample_count_device: NULL read arrays");
    TRY(ctx_use(s->ctx));
    return launch_count(s, n_reads, (const int64_t *)d_read_ptr, (const int32_t *)d_read_tid, (const int32_t *)d_read_fraglen);
}

extern "C" int emsar_sample_begin(emsar_index *ix, emsar_sample **out)
{
    CHECK_ARG(ix && out, "emsar_sample_begin: NULL argument");
    *out = nullptr;
    emsar_ctx *ctx = ix->ctx;
    TRY(ctx_use(ctx));
    emsar_sample *s = new emsar_sample();
    s->index = ix; s->ctx = ctx;
    int rc;
    if ((rc = dev_alloc(&s->d_R, (size_t)ix->C)) || (rc = dev_alloc(&s->d_hist, (size_t)ix->max_fl + 1)) || (rc = dev_alloc(&s->d_flags, 4))) {
        emsar_sample_end(s); return rc;
    }
    CU(cudaMemsetAsync(s->d_R, 0, (size_t)ix->C * 4, ctx->stream));                    // clear_readcounts_in_rshbucket
    CU(cudaMemsetAsync(s->d_hist, 0, ((size_t)ix->max_fl + 1) * 4, ctx->stream));      // calloc FraglengthCounts
    CU(cudaMemsetAsync(s->d_flags, 0, 16, ctx->stream));
    *out = s;
    return EMSAR_OK;
}

static int check_flags(emsar_sample *s)
{
    int32_t f = 0;
    CU(cudaMemcpyAsync(&f, s->d_flags, 4, cudaMemcpyDeviceToHost, s->ctx->stream));
    CU(cudaStreamSynchronize(s->ctx->stream));
    if (f & 1) { emsar_set_err("a read group carried more than %d alignments (use -k <= %d)", EMSAR_MAX_READ_TIDS, EMSAR_MAX_READ_TIDS); return EMSAR_ERR_UNSUPPORTED; }
    if (f & 2) { emsar_set_err("a read carried a transcript id outside 0..T-1"); return EMSAR_ERR_BAD_ARG; }
    if (f & 4) { emsar_set_err("emsar_sample_count_compact: the lengths add up to more tids than the caller passed"); return EMSAR_ERR_BAD_ARG; }
    return EMSAR_OK;
}
int sample_check_flags(emsar_sample *s) { return check_flags(s); }

extern "C" int emsar_sample_counts_set(emsar_sample *s, const int32_t *ReadCount, const int32_t *FraglengthCounts)
{
    CHECK_ARG(s && ReadCount && FraglengthCounts, "emsar_sample_counts_set: NULL argument");
    emsar_index *ix = s->index;
    TRY(ctx_use(s->ctx));
    CU(cudaMemcpyAsync(s->d_R, ReadCount, (size_t)ix->C * 4, cudaMemcpyHostToDevice, s->ctx->stream));
    CU(cudaMemcpyAsync(s->d_hist, FraglengthCounts, ((size_t)ix->max_fl + 1) * 4, cudaMemcpyHostToDevice, s->ctx->stream));
    CU(cudaStreamSynchronize(s->ctx->stream));
    s->have_counts = true; s->prepared = false;
    return EMSAR_OK;
}

extern "C" int emsar_sample_counts_get(emsar_sample *s, int32_t *ReadCount, int32_t *FraglengthCounts, int64_t *TotalReadCount)
{
    CHECK_ARG(s, "emsar_sample_counts_get: NULL sample");
    emsar_index *ix = s->index;
    TRY(ctx_use(s->ctx));
    TRY(check_flags(s));
    std::vector<int32_t> hist((size_t)ix->max_fl + 1);
    CU(cudaMemcpyAsync(hist.data(), s->d_hist, hist.size() * 4, cudaMemcpyDeviceToHost, s->ctx->stream));
    if (ReadCount) CU(cudaMemcpyAsync(ReadCount, s->d_R, (size_t)ix->C * 4, cudaMemcpyDeviceToHost, s->ctx->stream));
    CU(cudaStreamSynchronize(s->ctx->stream));
    if (FraglengthCounts) memcpy(FraglengthCounts, hist.data(), hist.size() * 4);
    if (TotalReadCount) { int64_t n = 0; for (int32_t v : hist) n += v; *TotalReadCount = n; }
    return EMSAR_OK;
}
