// The EM loop: one persistent cooperative kernel (one wave of CTAs resident on all 148 SMs) that runs
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

// ---- E-phase ---------------------------------------------------------------------------------------
template <int K>
__device__ __forceinline__ double esum_tp(const int32_t *__restrict__ tids, const double *theta, int lane)
{
    int t[K];
#pragma unroll
    for (int j = 0; j < K; j++) t[j] = __ldg(tids + j * 32 + lane);
    double v[K];
#pragma unroll
    for (int j = 0; j < K; j++) v[j] = theta[t[j]];
    double s = 0;
#pragma unroll
    for (int j = 0; j < K; j++) s += v[j];      // sequential member order
    return s;
}

template <int G>
__device__ __forceinline__ void etile_group(const EmParams &p, int4 tile, int k, int lane)
{
    constexpr int CPP = 32 / G;          // classes per pass
    const int sub = lane / G, l = lane % G;
    const double *theta = p.m.theta;
    for (int c0 = 0; c0 < tile.y; c0 += CPP) {
        const int cl = c0 + sub;
        const bool valid = cl < tile.y;
        double s = 0;
        if (valid) {
            const int32_t *tids = p.m.e_tid + (uint32_t)tile.z + (uint32_t)cl * (uint32_t)k;
#pragma unroll 4
            for (int i = l; i < k; i += G) s += theta[__ldg(tids + i)];
        }
#pragma unroll
        for (int d = G / 2; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        if (valid && l == 0) {
            const int j = tile.x + cl;
            const double r = (double)__ldg(p.m.e_R + j);
            p.m.q[j] = s > 0 ? r / s : 0.0;
        }
    }
}

__device__ __forceinline__ void e_phase(const EmParams &p, int gwarp, int nwarps, int lane)
{
    for (int g = gwarp; g < p.m.n_etiles; g += nwarps) {
        const int4 tile = __ldg(p.m.e_tiles + g);
        const int k = tile.w & 0xffff, mode = tile.w >> 16;
        if (mode == 0) {
            const int32_t *tids = p.m.e_tid + (uint32_t)tile.z;
            double s;
            switch (k) {
            case 2: s = esum_tp<2>(tids, p.m.theta, lane); break;
            case 3: s = esum_tp<3>(tids, p.m.theta, lane); break;
            case 4: s = esum_tp<4>(tids, p.m.theta, lane); break;
            case 5: s = esum_tp<5>(tids, p.m.theta, lane); break;
            case 6: s = esum_tp<6>(tids, p.m.theta, lane); break;
            case 7: s = esum_tp<7>(tids, p.m.theta, lane); break;
            default: s = esum_tp<8>(tids, p.m.theta, lane); break;
            }
            if (lane < tile.y) {
                const int j = tile.x + lane;
                const double r = (double)__ldg(p.m.e_R + j);
                p.m.q[j] = s > 0 ? r / s : 0.0;
            }
        } else if (mode == 1) etile_group<8>(p, tile, k, lane);
        else etile_group<32>(p, tile, k, lane);
    }
}

// ---- M-phase ---------------------------------------------------------------------------------------
__device__ __forceinline__ double m_update(const EmParams &p, int row, double Q)
{
    const double2 ra = p.m.row_RsA[row];
    const double th = p.m.theta[row];
    const double n = ra.x + th * Q;
    const double thn = n / ra.y;
    p.m.theta[row] = thn;
    return fabs(thn - th) * ra.y / (p.eps_abs + p.eps_rel * n);
}

__device__ __forceinline__ double m_phase(const EmParams &p, double *sm_block, int gwarp, int nwarps, int lane)
{
    double dmax = 0;
    const double *q = p.m.q;
    const uint32_t *row_off = p.m.row_off;
    // (1) hub rows: one CTA per row
    const int hub0 = p.m.n_short + p.m.n_long;
    for (int h = blockIdx.x; h < p.m.n_hub; h += gridDim.x) {
        const int row = hub0 + h;
        const uint32_t e0 = row_off[row], e1 = row_off[row + 1];
        double s = 0;
        for (uint32_t e = e0 + threadIdx.x; e < e1; e += EM_BLOCK) s += q[__ldg(p.m.m_cls + e)];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        if (lane == 0) sm_block[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            double Q = 0;
            for (int w = 0; w < EM_WARPS; w++) Q += sm_block[w];
            dmax = fmax(dmax, m_update(p, row, Q));
        }
        __syncthreads();
    }
    // (2) long rows: one warp per row
    for (int r = gwarp; r < p.m.n_long; r += nwarps) {
        const int row = p.m.n_short + r;
        const uint32_t e0 = row_off[row], e1 = row_off[row + 1];
        double s = 0;
#pragma unroll 4
        for (uint32_t e = e0 + lane; e < e1; e += 32) s += q[__ldg(p.m.m_cls + e)];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        if (lane == 0) dmax = fmax(dmax, m_update(p, row, s));
    }
    // (3) short rows: a warp takes one tile of <= 32 consecutive rows; 8 lanes reduce one row at a time (4 rows per
    //     pass, fixed shuffle tree), then all 32 lanes update their row together (coalesced Rs/A/theta traffic)
    for (int g = nwarps - 1 - gwarp; g < p.m.n_mtiles; g += nwarps) {
        const int2 tile = __ldg(p.m.m_tiles + g);
        const int nrows = tile.y - tile.x;
        if (nrows <= 0) continue;
        const int sub = lane >> 3, l = lane & 7;
        double Qmine = 0;
        for (int pass = 0; pass * 4 < nrows; pass++) {
            const int r = pass * 4 + sub;
            double s = 0;
            if (r < nrows) {
                const uint32_t a = row_off[tile.x + r], b = row_off[tile.x + r + 1];
#pragma unroll 4
                for (uint32_t e = a + l; e < b; e += 8) s += q[__ldg(p.m.m_cls + e)];
            }
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            const double got = __shfl_sync(0xffffffffu, s, (lane & 3) * 8);
            if ((lane >> 2) == pass) Qmine = got;
        }
        if (lane < nrows) dmax = fmax(dmax, m_update(p, tile.x + lane, Qmine));
    }
    return dmax;
}

template <int MINB>
__global__ void __launch_bounds__(EM_BLOCK, MINB) k_em_persistent(EmParams p)
{
    __shared__ double sm_block[EM_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = gridDim.x * EM_WARPS;
    const int gwarp = blockIdx.x * EM_WARPS + warp;
    int it = 0;
    double d = INFINITY;
    while (it < p.max_iter) {
        e_phase(p, gwarp, nwarps, lane);
        grid_barrier(p.bar, gridDim.x);
        double dm = m_phase(p, sm_block, gwarp, nwarps, lane);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dm = fmax(dm, __shfl_xor_sync(0xffffffffu, dm, o));
        __syncthreads();
        if (lane == 0) sm_block[warp] = dm;
        __syncthreads();
        if (threadIdx.x == 0) {
            double b = 0;
            for (int w = 0; w < EM_WARPS; w++) b = fmax(b, sm_block[w]);
            atomicMax(p.dmax + (it & 1), (unsigned long long)__double_as_longlong(b));
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
    int nb = 0;
    // register budget variant: 2, 3 or 4 resident CTAs per SM (EMSAR_EM_MINB overrides the default for tuning)
    int minb = EM_MIN_BLOCKS;
    const char *e = getenv("EMSAR_EM_MINB");
    if (e && atoi(e) >= 2 && atoi(e) <= 4) minb = atoi(e);
    ctx->em_minb = minb;
    if (minb == 2) CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_em_persistent<2>, EM_BLOCK, 0));
    else if (minb == 3) CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_em_persistent<3>, EM_BLOCK, 0));
    else CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_em_persistent<4>, EM_BLOCK, 0));
    if (nb < 1) { emsar_set_err("EM kernel does not fit on an SM"); return EMSAR_ERR_CUDA; }
    ctx->em_blocks_per_sm = nb;
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
    const int grid = ctx->prop.multiProcessorCount * ctx->em_blocks_per_sm;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(EM_BLOCK);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attrs[2];
    int na = 0;
    attrs[na].id = cudaLaunchAttributeCooperative;
    attrs[na].val.cooperative = 1;
    na++;
    if (ctx->l2_persist_bytes > 0 && s->state_bytes > 0) {
        // keep theta | q resident in L2 while the index streams through (access-policy window)
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
    if (ctx->em_minb == 2) CU(cudaLaunchKernelEx(&cfg, k_em_persistent<2>, p));
    else if (ctx->em_minb == 3) CU(cudaLaunchKernelEx(&cfg, k_em_persistent<3>, p));
    else CU(cudaLaunchKernelEx(&cfg, k_em_persistent<4>, p));
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
    const int P = s->m.n_short + s->m.n_long + s->m.n_hub;
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
    cudaFree(s->d_state); cudaFree(s->d_pack);
    delete s;
    return EMSAR_OK;
}
