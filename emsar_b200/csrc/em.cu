// The EM loop: one persistent cooperative kernel, one CTA per SM. Each CTA owns a contiguous range of transcripts
// (rows) and the classes whose first member lies in it, and keeps theta / q of what it owns in SHARED MEMORY: the
// gathers of both phases hit 32 independent banks instead of one L1 line per wavefront; only references that leave
// the range (halo) go through the L2-resident global copies. Per iteration:
//   E-phase  q_c = R_c / sum_{t in c} theta_t        class-major, binned by cardinality
//   -- grid barrier --
//   M-phase  theta_t' = (Rs_t + theta_t * sum_{c∋t} q_c) / A_t   transposed CSR, deterministic segmented reduction,
//            fused with the convergence measure  max_t |dtheta_t| A_t / (eps_abs + eps_rel n_t)
//   -- grid barrier --  (every CTA reads the reduced delta and decides to stop: no host round trip)
// until convergence or max_iter.  It replaces run_MLE_threads / MLE_range / MLE / Fp / lambdap of the reference
// (emsar_functions.c:2946-3126), which reach the same Poisson-likelihood optimum by a randomized pattern search.
// No floating-point atomics anywhere: every sum has an order fixed by the packed layout alone.
#include <cooperative_groups.h>
#include <math.h>

#include "common.cuh"

int sample_ensure_sets(emsar_sample *s);

struct EmParams {
    EmModel m;
    double eps_abs, eps_rel;
    int max_iter, stop_on_conv;
    unsigned *bar;                 // [0] count, [1] generation
    unsigned long long *dmax;      // [2] alternating slots for the reduced delta (bit pattern of a double >= 0)
    int *iters_done;
    double *final_delta;
};

// Sense-reversing grid barrier. All CTAs are co-resident (cooperative launch). The __threadfence() pair makes the
// writes of the phase visible device-wide and drops stale L1 lines before the next phase gathers.
__device__ __forceinline__ void grid_barrier(unsigned *bar, unsigned nblocks)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        volatile unsigned *gen = bar + 1;
        const unsigned g = *gen;
        __threadfence();
        if (atomicAdd(bar, 1u) == nblocks - 1) {
            bar[0] = 0;
            __threadfence();
            atomicAdd(bar + 1, 1u);
        } else {
            while (*gen == g) { }
        }
        __threadfence();
    }
    __syncthreads();
}

// ---- shared-memory resident slices ---------------------------------------------------------------------
struct BlockView {
    double *sm_theta;          // theta of the rows this CTA owns
    double *sm_q;              // q of the resident classes this CTA owns; sm_q[nres] == 0 (padding target)
    const int4 *sm_etiles;     // this CTA's E tile descriptors
    const int4 *sm_mitems;     // this CTA's M items
    int row0, nrows, cls0, nres;
};

// branch-free: one generic load from either the CTA's shared slice or the global (halo) copy
__device__ __forceinline__ double load_theta(const EmParams &p, const BlockView &v, int enc)
{
    const double *ptr = enc >= 0 ? v.sm_theta + enc : p.m.theta + ~enc;
    return *ptr;
}
__device__ __forceinline__ double load_q(const EmParams &p, const BlockView &v, int enc)
{
    const double *ptr = enc >= 0 ? v.sm_q + enc : p.m.q + ~enc;
    return *ptr;
}
__device__ __forceinline__ void store_q(const EmParams &p, const BlockView &v, int j, uint32_t rflag, double s)
{
    const double r = (double)(rflag & 0x7fffffffu);
    const double val = s > 0 ? r / s : 0.0;
    const int loc = j - v.cls0;
    if (loc < v.nres) v.sm_q[loc] = val;
    if (rflag & 0x80000000u) p.m.q[j] = val;        // a row of another CTA reads it, or it does not fit in shared memory
}

// ---- E-phase: q_c = R_c / sum of theta over the class members --------------------------------------------------
__device__ __forceinline__ void e_phase(const EmParams &p, const BlockView &v, int n_tiles, int warp, int lane)
{
    // tiles are ordered by cardinality: walk them from the heaviest so that the tail of the phase is made of light tiles
    for (int g = n_tiles - 1 - warp; g >= 0; g -= EM_WARPS) {
        const int4 tile = v.sm_etiles[g];
        const int k = tile.w & 0xffff, mode = tile.w >> 16;
        if (mode == 0) {
            // one thread per class; member j of the 32 classes of the tile is one coalesced 128-byte line
            const int32_t *__restrict__ tids = p.m.e_tid + (uint32_t)tile.z + lane;
            uint32_t rf = 0;
            if (lane < tile.y) rf = __ldg(p.m.e_R + tile.x + lane);
            double s = 0;
            int j = 0;
            for (; j + 4 <= k; j += 4) {
                const int t0 = __ldg(tids + j * 32), t1 = __ldg(tids + j * 32 + 32), t2 = __ldg(tids + j * 32 + 64), t3 = __ldg(tids + j * 32 + 96);
                const double x0 = load_theta(p, v, t0), x1 = load_theta(p, v, t1), x2 = load_theta(p, v, t2), x3 = load_theta(p, v, t3);
                s += x0; s += x1; s += x2; s += x3;          // sequential member order
            }
            for (; j < k; j++) s += load_theta(p, v, __ldg(tids + j * 32));
            if (lane < tile.y) store_q(p, v, tile.x + lane, rf, s);
        } else {
            // long classes: one warp per class, members row-major
            for (int cl = 0; cl < tile.y; cl++) {
                const int32_t *__restrict__ tids = p.m.e_tid + (uint32_t)tile.z + (uint32_t)cl * (uint32_t)k;
                double s = 0;
#pragma unroll 4
                for (int i = lane; i < k; i += 32) s += load_theta(p, v, __ldg(tids + i));
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
                if (lane == 0) store_q(p, v, tile.x + cl, __ldg(p.m.e_R + tile.x + cl), s);
            }
        }
    }
}

// ---- M-phase: theta_t' = (Rs_t + theta_t * sum of q over the row) / A_t, fused convergence measure --------------
__device__ __forceinline__ double m_update(const EmParams &p, const BlockView &v, int slot, double Q)
{
    const double2 ra = p.m.row_RsA[v.row0 + slot];
    const double th = v.sm_theta[slot];
    const double n = ra.x + th * Q;
    const double thn = n / ra.y;
    v.sm_theta[slot] = thn;
    p.m.theta[v.row0 + slot] = thn;                 // write-through: halo readers and the final result
    return fabs(thn - th) * ra.y / (p.eps_abs + p.eps_rel * n);
}

__device__ __forceinline__ double m_phase(const EmParams &p, const BlockView &v, int n_items, int warp, int lane)
{
    double dmax = 0;
    // items are ordered longest first (long rows, then slices by decreasing length)
    for (int g = warp; g < n_items; g += EM_WARPS) {
        const int4 it = v.sm_mitems[g];
        const int len = it.w & 0x3fffffff;
        if ((it.w >> 30) == 0) {
            // a slice of 32 rows stored transposed: one thread per row, sequential sum in ascending class order
            const int32_t *__restrict__ ent = p.m.m_cls + (uint32_t)it.z + lane;
            double Q = 0;
            int j = 0;
            for (; j + 4 <= len; j += 4) {
                const int c0 = __ldg(ent + j * 32), c1 = __ldg(ent + j * 32 + 32), c2 = __ldg(ent + j * 32 + 64), c3 = __ldg(ent + j * 32 + 96);
                const double x0 = load_q(p, v, c0), x1 = load_q(p, v, c1), x2 = load_q(p, v, c2), x3 = load_q(p, v, c3);
                Q += x0; Q += x1; Q += x2; Q += x3;
            }
            for (; j < len; j++) Q += load_q(p, v, __ldg(ent + j * 32));
            if (lane < it.y) dmax = fmax(dmax, m_update(p, v, it.x + lane, Q));
        } else {
            // a long row: the whole warp, fixed shuffle tree
            const int32_t *__restrict__ ent = p.m.m_cls + (uint32_t)it.z;
            double s = 0;
#pragma unroll 4
            for (int e = lane; e < len; e += 32) s += load_q(p, v, __ldg(ent + e));
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
            if (lane == 0) dmax = fmax(dmax, m_update(p, v, it.x, s));
        }
    }
    return dmax;
}

template <int MINB>
__global__ void __launch_bounds__(EM_BLOCK, MINB) k_em_persistent(EmParams p)
{
    extern __shared__ __align__(16) unsigned char sm_dyn[];
    __shared__ double sm_red[EM_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x;
    const int et0 = p.m.blk_etile0[b], n_et = p.m.blk_etile0[b + 1] - et0;
    const int mi0 = p.m.blk_mitem0[b], n_mi = p.m.blk_mitem0[b + 1] - mi0;
    BlockView v;
    v.row0 = p.m.blk_row0[b]; v.nrows = p.m.blk_row0[b + 1] - v.row0;
    v.cls0 = p.m.blk_cls0[b]; v.nres = p.m.blk_nres[b];
    const SmemPlan pl = em_smem_plan(n_et, n_mi, v.nrows);
    int4 *s_et = (int4 *)(sm_dyn + pl.off_etiles);
    int4 *s_mi = (int4 *)(sm_dyn + pl.off_mitems);
    v.sm_theta = (double *)(sm_dyn + pl.off_theta);
    v.sm_q = (double *)(sm_dyn + pl.off_q);
    v.sm_etiles = s_et; v.sm_mitems = s_mi;
    // per-CTA constants and the CTA's slice of theta -> shared memory, once
    for (int i = threadIdx.x; i < n_et; i += EM_BLOCK) s_et[i] = p.m.e_tiles[et0 + i];
    for (int i = threadIdx.x; i < n_mi; i += EM_BLOCK) s_mi[i] = p.m.m_items[mi0 + i];
    for (int i = threadIdx.x; i < v.nrows; i += EM_BLOCK) v.sm_theta[i] = p.m.theta[v.row0 + i];
    for (int i = threadIdx.x; i <= v.nres; i += EM_BLOCK) v.sm_q[i] = 0.0;
    __syncthreads();
    int it = 0;
    double d = INFINITY;
    while (it < p.max_iter) {
        e_phase(p, v, n_et, warp, lane);
        grid_barrier(p.bar, gridDim.x);
        double dm = m_phase(p, v, n_mi, warp, lane);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dm = fmax(dm, __shfl_xor_sync(0xffffffffu, dm, o));
        if (lane == 0) sm_red[warp] = dm;
        __syncthreads();
        if (threadIdx.x == 0) {
            double bm = 0;
            for (int w = 0; w < EM_WARPS; w++) bm = fmax(bm, sm_red[w]);
            atomicMax(p.dmax + (it & 1), (unsigned long long)__double_as_longlong(bm));
        }
        grid_barrier(p.bar, gridDim.x);
        d = __longlong_as_double((long long)*((volatile unsigned long long *)(p.dmax + (it & 1))));
        if (blockIdx.x == 0 && threadIdx.x == 0) p.dmax[(it + 1) & 1] = 0ULL;
        it++;
        if (p.stop_on_conv && d <= 1.0) break;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { *p.iters_done = it; *p.final_delta = d; }
}

int em_query_occupancy(emsar_ctx *ctx)
{
    // one CTA per SM owns a row range; all the shared memory an SM can give goes to the theta | q slices
    ctx->em_minb = 1;
    int smem = (int)ctx->prop.sharedMemPerBlockOptin - 2048;      // static (sm_red) + reserve
    const char *e = getenv("EMSAR_EM_SMEM_KB");
    if (e && atoi(e) > 0 && atoi(e) * 1024 < smem) smem = atoi(e) * 1024;
    smem &= ~255;
    CU(cudaFuncSetAttribute(k_em_persistent<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int nb = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_em_persistent<1>, EM_BLOCK, smem));
    if (nb < 1) { emsar_set_err("EM kernel does not fit on an SM (%d bytes of shared memory)", smem); return EMSAR_ERR_CUDA; }
    ctx->em_blocks_per_sm = 1;
    ctx->em_smem_bytes = smem;
    return EMSAR_OK;
}

int em_launch(emsar_sample *s, int max_iter, int stop_on_conv, int *iters_done, double *final_delta, double *ms_out)
{
    emsar_ctx *ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    EmParams p;
    p.m = s->m;
    p.eps_abs = s->opts.eps_abs; p.eps_rel = s->opts.eps_rel;
    p.max_iter = max_iter; p.stop_on_conv = stop_on_conv;
    p.bar = ctx->d_barrier;
    p.dmax = (unsigned long long *)(ctx->d_barrier + 4);
    p.iters_done = (int *)(ctx->d_barrier + 8);
    p.final_delta = (double *)(ctx->d_barrier + 10);
    CU(cudaMemsetAsync(ctx->d_barrier, 0, 64, st));
    const int grid = s->m.B;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(EM_BLOCK);
    cfg.dynamicSmemBytes = (size_t)ctx->em_smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attrs[2];
    int na = 0;
    attrs[na].id = cudaLaunchAttributeCooperative;
    attrs[na].val.cooperative = 1;
    na++;
    if (ctx->l2_persist_bytes > 0 && s->state_bytes > 0) {
        // the global copies of theta | q (halo traffic) stay resident in L2 while the index streams through
        size_t win = s->state_bytes;
        if (win > (size_t)ctx->prop.accessPolicyMaxWindowSize) win = (size_t)ctx->prop.accessPolicyMaxWindowSize;
        attrs[na].id = cudaLaunchAttributeAccessPolicyWindow;
        attrs[na].val.accessPolicyWindow.base_ptr = s->d_state;
        attrs[na].val.accessPolicyWindow.num_bytes = win;
        attrs[na].val.accessPolicyWindow.hitRatio = win <= ctx->l2_persist_bytes ? 1.0f : (float)ctx->l2_persist_bytes / (float)win;
        attrs[na].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attrs[na].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        na++;
    }
    cfg.attrs = attrs;
    cfg.numAttrs = na;
    CU(cudaEventRecord(ctx->ev0, st));
    CU(cudaLaunchKernelEx(&cfg, k_em_persistent<1>, p));
    LAUNCHED(ctx);
    CU(cudaEventRecord(ctx->ev1, st));
    int it = 0; double fd = 0;
    CU(cudaMemcpyAsync(&it, p.iters_done, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&fd, p.final_delta, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    if (iters_done) *iters_done = it;
    if (final_delta) *final_delta = fd;
    if (ms_out) *ms_out = ms;
    return EMSAR_OK;
}

__global__ void k_fill_double2(double *p, int64_t n, double v)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

extern "C" int emsar_sample_em_run(emsar_sample *s, int32_t max_iter, int32_t stop_on_conv, int32_t reset_theta,
                                   int32_t *iters_done, double *final_delta, double *elapsed_ms)
{
    CHECK_ARG(s, "emsar_sample_em_run: NULL sample");
    if (!s->prepared) { emsar_set_err("emsar_sample_em_run: call emsar_sample_prepare first"); return EMSAR_ERR_STATE; }
    CU(cudaSetDevice(s->ctx->device));
    const int P = s->m.P;
    if (reset_theta) {
        if (P > 0) { k_fill_double2<<<(unsigned)((P + 255) / 256), 256, 0, s->ctx->stream>>>(s->m.theta, P, 1.0); LAUNCHED(s->ctx); }
        s->n_iter = 0;
    }
    if (max_iter <= 0) max_iter = s->opts.max_iter;
    int it = 0; double fd = 0, ms = 0;
    TRY(em_launch(s, max_iter, stop_on_conv, &it, &fd, &ms));
    s->n_iter += it; s->final_delta = fd; s->em_ms += ms;
    if (iters_done) *iters_done = it;
    if (final_delta) *final_delta = fd;
    if (elapsed_ms) *elapsed_ms = ms;
    return EMSAR_OK;
}

// ---- outputs ---------------------------------------------------------------------------------------
// FPKM in transcript order, including the closed cases of MLE() (:3054-3066) for transcripts outside the EM.
__global__ void k_fpkm(int32_t T, const int32_t *__restrict__ pos, const double *__restrict__ theta, const uint8_t *__restrict__ lone,
                       const int32_t *__restrict__ R, const double *__restrict__ adj, double nscale, double p10, double *__restrict__ fpkm)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const int p = pos[t];
    double f;
    if (p >= 0) f = theta[p];
    else if (lone[t] && R[t] > 0) f = (double)R[t] / (adj[t] / 1E3 * nscale * p10);   // lone-singleton set: R / EUMAps
    else f = 0.0;
    fpkm[t] = f;
}

// iReadcount, Round_off, per-CTA partial sums of FPKM and iReadcount_int (fixed tree: deterministic).
constexpr int FIN_BLOCK = 256;
__global__ void __launch_bounds__(FIN_BLOCK) k_finalize1(int32_t T, const double *__restrict__ fpkm, const double *__restrict__ iE, double nscale,
                                                         double *__restrict__ ireadcount, int32_t *__restrict__ iri, double *__restrict__ part_f,
                                                         long long *__restrict__ part_i)
{
    __shared__ double sf[FIN_BLOCK];
    __shared__ long long si[FIN_BLOCK];
    int t = blockIdx.x * FIN_BLOCK + threadIdx.x;
    double f = 0; long long ri = 0;
    if (t < T) {
        f = fpkm[t];
        double ir = (iE[t] / 1E3) * f * nscale;                    // print_FPKMfinal :3203
        int r = (ir - (int)ir >= 0.5) ? (int)ir + 1 : (int)ir;     // Round_off :3215-3217
        ireadcount[t] = ir; iri[t] = r; ri = r;
    }
    sf[threadIdx.x] = f; si[threadIdx.x] = ri;
    __syncthreads();
    for (int s = FIN_BLOCK / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) { sf[threadIdx.x] += sf[threadIdx.x + s]; si[threadIdx.x] += si[threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { part_f[blockIdx.x] = sf[0]; part_i[blockIdx.x] = si[0]; }
}
__global__ void k_finalize2(int nparts, const double *__restrict__ part_f, const long long *__restrict__ part_i, double *tot_f, long long *tot_i)
{
    if (threadIdx.x || blockIdx.x) return;
    double f = 0; long long i = 0;
    for (int b = 0; b < nparts; b++) { f += part_f[b]; i += part_i[b]; }
    *tot_f = f; *tot_i = i;
}
__global__ void k_tpm(int32_t T, const double *__restrict__ fpkm, const double *tot_f, double *__restrict__ tpm)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < T) tpm[t] = fpkm[t] * 1E6 / *tot_f;                    // :3207
}

// Per class: log-likelihood term exactly as Fp/lambdap (:2946-2975) and expected_Readcount (print_aEUMA_3 :2289-2296).
constexpr double NEAR_LOWEST = -9.9E307;
__global__ void __launch_bounds__(FIN_BLOCK) k_class_eval(int64_t C, const uint32_t *__restrict__ cls_off, const int32_t *__restrict__ cls_tid,
                                                          const double *__restrict__ fpkm, const double *__restrict__ amodel,
                                                          const double *__restrict__ adj, const int32_t *__restrict__ R, double nscale,
                                                          double *__restrict__ expected, double *__restrict__ part_ll, int *__restrict__ bad)
{
    __shared__ double sl[FIN_BLOCK];
    int64_t c = (int64_t)blockIdx.x * FIN_BLOCK + threadIdx.x;
    double ll = 0;
    if (c < C) {
        const uint32_t o = cls_off[c], e = cls_off[c + 1];
        double s = 0, ex = 0;
        const double w = adj[c] / 1E3;
        for (uint32_t j = o; j < e; j++) { double f = fpkm[cls_tid[j]]; s += f; ex += f * w * nscale; }
        if (expected) expected[c] = ex;
        const double a = amodel[c];
        if (a != 0) {
            const double lamb = a * s;
            if (lamb == 0) { if (R[c] != 0) *bad = 1; }
            else if (lamb < 0) *bad = 1;
            else ll = (double)R[c] * log(lamb) - lamb;
        }
    }
    sl[threadIdx.x] = ll;
    __syncthreads();
    for (int s = FIN_BLOCK / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) sl[threadIdx.x] += sl[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) part_ll[blockIdx.x] = sl[0];
}
__global__ void k_sum_parts(int nparts, const double *__restrict__ part, double *out)
{
    if (threadIdx.x || blockIdx.x) return;
    double f = 0;
    for (int b = 0; b < nparts; b++) f += part[b];
    *out = f;
}

static int class_eval(emsar_sample *s, const double *d_fpkm, double *d_expected, double *loglik)
{
    emsar_index *ix = s->index; emsar_ctx *ctx = s->ctx; cudaStream_t st = ctx->stream;
    const int nb = (int)((ix->C + FIN_BLOCK - 1) / FIN_BLOCK);
    void *scr = nullptr;
    // scratch layout: [fpkm T doubles (caller)] is NOT here; only partials
    size_t need = (size_t)nb * 8 + 64;
    double *d_part = nullptr; int *d_bad = nullptr; double *d_out = nullptr;
    TRY(dev_alloc(&d_part, (size_t)nb + 8));
    (void)scr; (void)need;
    d_out = d_part + nb; d_bad = (int *)(d_part + nb + 1);
    CU(cudaMemsetAsync(d_part + nb, 0, 64, st));
    const double nscale = (double)s->N / 1E6;
    k_class_eval<<<nb, FIN_BLOCK, 0, st>>>(ix->C, ix->d_cls_off, ix->d_cls_tid, d_fpkm, s->d_amodel, s->d_adj, s->d_R, nscale, d_expected, d_part, d_bad);
    LAUNCHED(ctx);
    k_sum_parts<<<1, 32, 0, st>>>(nb, d_part, d_out);
    LAUNCHED(ctx);
    CU(cudaGetLastError());
    double ll = 0; int bad = 0;
    CU(cudaMemcpyAsync(&ll, d_out, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    cudaFree(d_part);
    if (bad || ll < NEAR_LOWEST) ll = NEAR_LOWEST;
    if (loglik) *loglik = ll;
    return EMSAR_OK;
}

extern "C" int emsar_sample_finalize(emsar_sample *s, emsar_solve_out *out)
{
    CHECK_ARG(s && out, "emsar_sample_finalize: NULL argument");
    if (!s->prepared) { emsar_set_err("emsar_sample_finalize: sample not prepared"); return EMSAR_ERR_STATE; }
    emsar_index *ix = s->index; emsar_ctx *ctx = s->ctx; cudaStream_t st = ctx->stream;
    CU(cudaSetDevice(ctx->device));
    const int32_t T = ix->T;
    const int nb = (T + FIN_BLOCK - 1) / FIN_BLOCK;
    char *buf = nullptr;
    const size_t tb = (((size_t)T * 8 + 255) / 256) * 256;
    TRY(dev_alloc(&buf, tb * 4 + (size_t)nb * 16 + 256));
    double *d_fpkm = (double *)buf, *d_ir = (double *)(buf + tb), *d_tpm = (double *)(buf + 2 * tb);
    int32_t *d_iri = (int32_t *)(buf + 3 * tb);
    double *d_pf = (double *)(buf + 4 * tb);
    long long *d_pi = (long long *)(d_pf + nb);
    double *d_totf = (double *)(d_pi + nb);
    long long *d_toti = (long long *)(d_totf + 1);
    const double nscale = (double)s->N / 1E6, p10 = pow(10, s->delta);
    k_fpkm<<<(T + 255) / 256, 256, 0, st>>>(T, s->d_pos, s->m.theta, s->d_lone, s->d_R, s->d_adj, nscale, p10, d_fpkm);
    LAUNCHED(ctx);
    k_finalize1<<<nb, FIN_BLOCK, 0, st>>>(T, d_fpkm, s->d_iE, nscale, d_ir, d_iri, d_pf, d_pi);
    LAUNCHED(ctx);
    k_finalize2<<<1, 32, 0, st>>>(nb, d_pf, d_pi, d_totf, d_toti);
    LAUNCHED(ctx);
    k_tpm<<<(T + 255) / 256, 256, 0, st>>>(T, d_fpkm, d_totf, d_tpm);
    LAUNCHED(ctx);
    CU(cudaGetLastError());
    if (out->fpkm) CU(cudaMemcpyAsync(out->fpkm, d_fpkm, (size_t)T * 8, cudaMemcpyDeviceToHost, st));
    if (out->efflen) CU(cudaMemcpyAsync(out->efflen, s->d_iE, (size_t)T * 8, cudaMemcpyDeviceToHost, st));
    if (out->ireadcount) CU(cudaMemcpyAsync(out->ireadcount, d_ir, (size_t)T * 8, cudaMemcpyDeviceToHost, st));
    if (out->ireadcount_int) CU(cudaMemcpyAsync(out->ireadcount_int, d_iri, (size_t)T * 4, cudaMemcpyDeviceToHost, st));
    if (out->tpm) CU(cudaMemcpyAsync(out->tpm, d_tpm, (size_t)T * 8, cudaMemcpyDeviceToHost, st));
    long long toti = 0;
    CU(cudaMemcpyAsync(&toti, d_toti, 8, cudaMemcpyDeviceToHost, st));
    double ll = 0;
    int rc = class_eval(s, d_fpkm, nullptr, &ll);      // synchronizes the stream
    cudaFree(buf);
    if (rc != EMSAR_OK) return rc;
    out->n_iter = s->n_iter; out->final_delta = s->final_delta; out->loglik = ll;
    out->total_ireadcount = toti; out->total_readcount = s->N; out->eumacut = s->eumacut; out->max_sid = s->max_sid;
    out->em_ms = s->em_ms; out->prep_ms = s->prep_ms;
    return EMSAR_OK;
}

extern "C" int emsar_sample_solve(emsar_sample *s, const emsar_solve_opts *opts, emsar_solve_out *out)
{
    CHECK_ARG(s && out, "emsar_sample_solve: NULL argument");
    TRY(emsar_sample_prepare(s, opts));
    int it = 0; double fd = 0, ms = 0;
    TRY(emsar_sample_em_run(s, s->opts.max_iter, 1, 0, &it, &fd, &ms));
    return emsar_sample_finalize(s, out);
}

extern "C" int emsar_sample_theta_get(emsar_sample *s, double *theta)
{
    CHECK_ARG(s && theta, "emsar_sample_theta_get: NULL argument");
    if (!s->prepared) { emsar_set_err("emsar_sample_theta_get: sample not prepared"); return EMSAR_ERR_STATE; }
    emsar_index *ix = s->index; emsar_ctx *ctx = s->ctx; cudaStream_t st = ctx->stream;
    CU(cudaSetDevice(ctx->device));
    double *d_fpkm = nullptr;
    TRY(dev_alloc(&d_fpkm, (size_t)ix->T));
    const double nscale = (double)s->N / 1E6, p10 = pow(10, s->delta);
    k_fpkm<<<(ix->T + 255) / 256, 256, 0, st>>>(ix->T, s->d_pos, s->m.theta, s->d_lone, s->d_R, s->d_adj, nscale, p10, d_fpkm);
    LAUNCHED(ctx);
    CU(cudaMemcpyAsync(theta, d_fpkm, (size_t)ix->T * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    cudaFree(d_fpkm);
    return EMSAR_OK;
}

extern "C" int emsar_sample_segments_get(emsar_sample *s, double *adjEUMA, double *expected, int32_t *set_id)
{
    CHECK_ARG(s, "emsar_sample_segments_get: NULL sample");
    if (!s->prepared) { emsar_set_err("emsar_sample_segments_get: sample not prepared"); return EMSAR_ERR_STATE; }
    emsar_index *ix = s->index; emsar_ctx *ctx = s->ctx; cudaStream_t st = ctx->stream;
    CU(cudaSetDevice(ctx->device));
    if (adjEUMA) CU(cudaMemcpyAsync(adjEUMA, s->d_adj, (size_t)ix->C * 8, cudaMemcpyDeviceToHost, st));
    if (expected) {
        double *d_fpkm = nullptr, *d_ex = nullptr;
        TRY(dev_alloc(&d_fpkm, (size_t)ix->T));
        TRY(dev_alloc(&d_ex, (size_t)ix->C));
        const double nscale = (double)s->N / 1E6, p10 = pow(10, s->delta);
        k_fpkm<<<(ix->T + 255) / 256, 256, 0, st>>>(ix->T, s->d_pos, s->m.theta, s->d_lone, s->d_R, s->d_adj, nscale, p10, d_fpkm);
        LAUNCHED(ctx);
        int rc = class_eval(s, d_fpkm, d_ex, nullptr);
        if (rc != EMSAR_OK) return rc;
        CU(cudaMemcpyAsync(expected, d_ex, (size_t)ix->C * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        cudaFree(d_fpkm); cudaFree(d_ex);
    }
    CU(cudaStreamSynchronize(st));
    if (set_id) {
        TRY(sample_ensure_sets(s));
        memcpy(set_id, s->h_CS.data(), (size_t)ix->C * 4);
    }
    return EMSAR_OK;
}

extern "C" int emsar_sample_end(emsar_sample *s)
{
    if (!s) return EMSAR_OK;
    cudaSetDevice(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    cudaFree(s->d_R); cudaFree(s->d_hist); cudaFree(s->d_flags);
    cudaFree(s->d_rd_ptr); cudaFree(s->d_rd_tid); cudaFree(s->d_rd_fl);
    cudaFree(s->d_Wf); cudaFree(s->d_adj); cudaFree(s->d_amodel); cudaFree(s->d_in_model);
    cudaFree(s->d_A); cudaFree(s->d_Rs); cudaFree(s->d_iE); cudaFree(s->d_lone); cudaFree(s->d_pos);
    cudaFree(s->d_state); cudaFree(s->d_pack); cudaFree(s->d_mcls);
    delete s;
    return EMSAR_OK;
}
