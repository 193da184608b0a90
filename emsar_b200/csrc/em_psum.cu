// k_em_psum (EM variant 5): the class-owner-centric persistent EM kernel. Layout and the idea: psum.cuh.
// One cooperative launch runs all iterations; nothing in it waits for the whole grid: a CTA depends only on the CTAs it exchanges rows
// with, through tagged 16-byte slots, and the convergence measure of iteration i is read behind the E-phase of iteration i + 1 (theta has
// not changed by then, so stopping there leaves exactly the state of iteration i - stopping rule and iteration count are the oracle's).
// Replaces run_MLE_threads / MLE_range / MLE / Fp / lambdap (reference emsar_functions.c:2946-3126) like k_em_persistent does.
#include "em_common.cuh"
#include "psum.cuh"

struct PsParams {
    PsModel m;
    double eps_abs, eps_rel;
    int max_iter, stop_on_conv;
    unsigned tag0;
    int *abort_flag;
    int *iters_done;
    double *final_delta;
    unsigned long long *trace;     // optional [B * 8] globaltimer stamps of the last iteration (tuning aid)
};

struct PsView {
    double *theta, *q, *Q;
    const unsigned char *cache;    // resident index cache (shared memory)
    int nrows, ncls, hr0;
};

template <bool RES> __device__ __forceinline__ uint4 ps_ld128(const unsigned char *sm, const unsigned char *g, int off16)
{
    if (RES) return *(const uint4 *)(sm + 16 * (size_t)off16);
    return __ldg((const uint4 *)g + off16);
}
template <bool RES> __device__ __forceinline__ uint2 ps_ld64(const unsigned char *sm, const unsigned char *g, int off8)
{
    if (RES) return *(const uint2 *)(sm + 8 * (size_t)off8);
    return __ldg((const uint2 *)g + off8);
}
template <bool RES> __device__ __forceinline__ uint32_t ps_ld32(const unsigned char *sm, const unsigned char *g, int off4)
{
    if (RES) return *(const uint32_t *)(sm + 4 * (size_t)off4);
    return __ldg((const uint32_t *)g + off4);
}

__device__ __forceinline__ double ps_q_of(uint32_t r, double s) { return s > 0 ? fast_div((double)r, s) : 0.0; }

// ---- staging: the index data of the NEXT tile / item of a warp travels from L2 / HBM into the warp's private staging buffer (cp.async)
// while the warp works on the current one, so that a tile that is not resident costs no exposed round trip. One buffer per warp is enough:
// a tile's words are pulled into registers first, then the buffer is handed to the next copy.
constexpr int PS_STG = 1152;            // bytes per warp: 1 KB of index data + 128 bytes of read counts / destinations (small tiles: 512 + 512)
__device__ __forceinline__ uint32_t ps_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ps_cp16(void *dst, const void *src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ps_smem_u32(dst)), "l"(src) : "memory"); }
__device__ __forceinline__ void ps_cp4(void *dst, const void *src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(ps_smem_u32(dst)), "l"(src) : "memory"); }
__device__ __forceinline__ void ps_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void ps_cp_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ double ps_gather4(const double *a, uint2 w, double s)
{
    const double x0 = a[w.x & 0xffffu], x1 = a[w.x >> 16], x2 = a[w.y & 0xffffu], x3 = a[w.y >> 16];
    s += x0; s += x1; s += x2; s += x3;            // member order
    return s;
}

// members of a long tile / slice (more than 4 chunks per lane): chunk c of all lanes = 256 bytes; four loads kept in flight
template <bool RES>
__device__ __forceinline__ double ps_sum_chunks(const double *a, const unsigned char *sm, const unsigned char *g, int o8, int n4)
{
    uint2 w[4];
#pragma unroll
    for (int u = 0; u < 4; u++) w[u] = ps_ld64<RES>(sm, g, o8 + min(u, n4 - 1) * 32);
    double s = 0;
    for (int c = 0; c < n4; c += 4) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (c + u < n4) {
                const uint2 cur = w[u];
                if (c + u + 4 < n4) w[u] = ps_ld64<RES>(sm, g, o8 + (c + u + 4) * 32);
                s = ps_gather4(a, cur, s);
            }
        }
    }
    return s;
}

__device__ __forceinline__ void ps_emit(const PsParams &p, const PsView &v, uint32_t dst, double S, unsigned tag)
{
    if (!(dst & PS_REMOTE)) v.Q[dst] = S;
    else ll_store(p.m.win[(dst >> 28) & 7u] + p.m.part_off + 16 * (size_t)(dst & 0x0fffffffu), S, tag);     // the row's owner collects it, in its rank
}

// a group of long rows: the warp reduces one row at a time (lane-strided partial sums, four loads in flight, fixed shuffle tree)
template <bool RES>
__device__ __forceinline__ void ps_m_group(const PsParams &p, const PsView &v, const int4 t, const unsigned char *gdat, int lane, unsigned tag)
{
    const int off16 = t.z, n = t.y;
    const double *q = v.q;
    const uint2 hw = lane < n ? ps_ld64<RES>(v.cache, gdat, off16 * 2 + lane) : make_uint2(0u, PS_NONE);      // {length, destination}
    const int mywords = (int)((hw.x + 1) >> 1);
    int start = mywords;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, start, o); if (lane >= o) start += y; }
    start += ((2 * n + 3) & ~3) - mywords;             // exclusive prefix (32-bit words), behind the header
    double mine = 0;
    for (int r = 0; r < n; r++) {
        const int a = off16 * 4 + __shfl_sync(0xffffffffu, start, r), W = __shfl_sync(0xffffffffu, mywords, r);
        double s = 0;
        int e = lane;
        for (; e + 96 < W; e += 128) {
            const uint32_t w0 = ps_ld32<RES>(v.cache, gdat, a + e), w1 = ps_ld32<RES>(v.cache, gdat, a + e + 32);
            const uint32_t w2 = ps_ld32<RES>(v.cache, gdat, a + e + 64), w3 = ps_ld32<RES>(v.cache, gdat, a + e + 96);
            const double x0 = q[w0 & 0xffffu], x1 = q[w0 >> 16], x2 = q[w1 & 0xffffu], x3 = q[w1 >> 16];
            const double x4 = q[w2 & 0xffffu], x5 = q[w2 >> 16], x6 = q[w3 & 0xffffu], x7 = q[w3 >> 16];
            s += x0; s += x1; s += x2; s += x3; s += x4; s += x5; s += x6; s += x7;
        }
        for (; e < W; e += 32) {
            const uint32_t w0 = ps_ld32<RES>(v.cache, gdat, a + e);
            const double x0 = q[w0 & 0xffffu], x1 = q[w0 >> 16];
            s += x0; s += x1;
        }
#pragma unroll
        for (int dd = 16; dd > 0; dd >>= 1) s += __shfl_xor_sync(0xffffffffu, s, dd);
        if (lane == r) mine = s;
    }
    if (lane < n) ps_emit(p, v, hw.y, mine, tag);
}

#define PS_TRACE(slot) do { if (p.trace && it == p.max_iter - 1 && threadIdx.x == 0) p.trace[blockIdx.x * 8 + (slot)] = gtime(); } while (0)

__global__ void __launch_bounds__(EM_BLOCK, 1) k_em_psum(PsParams p)
{
    extern __shared__ __align__(16) unsigned char sm_dyn[];
    __shared__ double sm_red[EM_WARPS];
    __shared__ double sm_bc;
    __shared__ int sm_ctr[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = (int)blockIdx.x + p.m.block0;
    const int row0 = p.m.blk_row0[b], nrows = p.m.blk_row0[b + 1] - row0;
    const int cls0 = p.m.blk_cls0[b], ncls = p.m.blk_cls0[b + 1] - cls0;
    const int et0 = p.m.blk_etile0[b], n_et = p.m.blk_etile0[b + 1] - et0;
    const int mi0 = p.m.blk_mitem0[b], n_mi = p.m.blk_mitem0[b + 1] - mi0;
    const int hr0 = p.m.blk_hr0[b], nhr = p.m.blk_hr0[b + 1] - hr0;
    const int desc_smem = p.m.blk_desc_smem[b];
    const int in0 = p.m.inc_off[row0], nin = p.m.inc_off[row0 + nrows] - in0;       // partial sums this CTA receives per iteration
    const PsPlan pl = ps_smem_plan(desc_smem, n_et, n_mi, nrows, nhr, ncls, nin, p.m.stage);
    const bool stage_e = p.m.stage & 1, stage_m = p.m.stage & 2;
    double *const s_in = (double *)(sm_dyn + pl.off_in);
    unsigned char *const stg = sm_dyn + pl.off_stg + PS_STG * warp;          // this warp's staging buffer
    PsView v;
    v.theta = (double *)(sm_dyn + pl.off_theta);
    v.q = (double *)(sm_dyn + pl.off_q);
    v.Q = (double *)(sm_dyn + pl.off_Q);
    v.cache = sm_dyn + pl.off_cache;
    v.nrows = nrows; v.ncls = ncls; v.hr0 = hr0;
    const int4 *et = desc_smem ? (const int4 *)(sm_dyn + pl.off_et) : p.m.e_tiles + et0;
    const int4 *mi = desc_smem ? (const int4 *)(sm_dyn + pl.off_mi) : p.m.m_items + mi0;
    if (desc_smem) {
        int4 *s_et = (int4 *)(sm_dyn + pl.off_et), *s_mi = (int4 *)(sm_dyn + pl.off_mi);
        for (int i = threadIdx.x; i < n_et; i += EM_BLOCK) s_et[i] = p.m.e_tiles[et0 + i];
        for (int i = threadIdx.x; i < n_mi; i += EM_BLOCK) s_mi[i] = p.m.m_items[mi0 + i];
    }
    for (int i = threadIdx.x; i < nrows; i += EM_BLOCK) { v.theta[i] = p.m.theta[row0 + i]; v.Q[i] = 0.0; }
    for (int i = threadIdx.x; i <= ncls; i += EM_BLOCK) v.q[i] = 0.0;             // q[ncls] = 0.0: padding target of the M items
    if (threadIdx.x == 0) v.theta[nrows + nhr] = 0.0;                             // zero-theta slot: padding target of the E tiles
    {
        // resident index cache: the data of the chosen tiles (+ their read counts) and items stays in shared memory for the whole kernel
        unsigned char *cache = sm_dyn + pl.off_cache;
        for (int i = warp; i < n_et; i += EM_WARPS) {
            const int4 t = p.m.e_tiles[et0 + i];
            if (!((t.w >> 30) & 1)) continue;
            const int steps = t.w & 0xfff, lg = (t.w >> 12) & 0xf;
            const int n16 = (lg == 0 && steps <= 4) ? 32 : 16 * ((steps + 3) >> 2);
            const uint4 *src = (const uint4 *)p.m.e_data + p.m.e_src[et0 + i];
            uint4 *dst = (uint4 *)cache + t.z;
            for (int j = lane; j < n16; j += 32) dst[j] = __ldg(src + j);
            uint32_t *rd = (uint32_t *)(dst + n16);
            for (int j = lane; j < t.y; j += 32) rd[j] = __ldg(p.m.e_R + cls0 + t.x + j);
        }
        for (int i = warp; i < n_mi; i += EM_WARPS) {
            const int4 t = p.m.m_items[mi0 + i];
            if (!((t.w >> 29) & 1)) continue;
            const int n16 = t.x;                    // an item's x carries its size in 16-byte units
            const uint4 *src = (const uint4 *)p.m.m_data + p.m.m_src[mi0 + i];
            uint4 *dst = (uint4 *)cache + t.z;
            for (int j = lane; j < n16; j += 32) dst[j] = __ldg(src + j);
        }
    }
    __syncthreads();
    const uint32_t *gR = p.m.e_R + cls0;
    const int Bt = p.m.Bt;
    unsigned char *const my_win = p.m.win[p.m.rank];
    const unsigned char *const th_slots = my_win + p.m.th_off, *const part_slots = my_win + p.m.part_off, *const dm_slots = my_win + p.m.dm_off;
    // the convergence measure of iteration j: the maximum over every CTA (slots alternate by iteration parity)
    auto read_dm = [&](int j) -> double {
        const unsigned tg = p.tag0 + (unsigned)j + 1u;
        double x = 0;
        for (int i = threadIdx.x; i < Bt; i += EM_BLOCK) x = fmax(x, ll_load(dm_slots + 16 * (size_t)((j & 1) * Bt + i), tg, p.abort_flag));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, o));
        if (lane == 0) sm_red[warp] = x;
        __syncthreads();
        if (warp == 0) {
            double y = sm_red[lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) y = fmax(y, __shfl_xor_sync(0xffffffffu, y, o));
            if (lane == 0) sm_bc = y;
        }
        __syncthreads();
        return sm_bc;
    };
    int it = 0;
    double d = INFINITY;
    bool stopped = false;
    while (it < p.max_iter) {
        const unsigned tag = p.tag0 + (unsigned)it + 1u;
        PS_TRACE(0);
        // theta of the halo rows, as their owners published it
        for (int i = threadIdx.x; i < nhr; i += EM_BLOCK) v.theta[nrows + i] = ll_load(th_slots + 16 * (size_t)__ldg(p.m.halo_rows + hr0 + i), tag, p.abort_flag);
        if (threadIdx.x == 0) { sm_ctr[0] = 0; sm_ctr[1] = 0; }
        __syncthreads();
        PS_TRACE(1);
        {
            // ---- E-phase: q_c = R_c / sum of theta over the members. Tiles come from the CTA's queue, heaviest first; a warp keeps the copy of
            // its next tile and the descriptor of the one after in flight while it computes ----
            auto e_desc = [&](int tk) -> int4 { return tk < n_et ? et[n_et - 1 - tk] : make_int4(0, 0, 0, 0); };
            auto e_stage = [&](const int4 &t) {
                const int steps = t.w & 0xfff, lg = (t.w >> 12) & 0xf, s4 = (steps + 3) >> 2;
                const bool small = lg == 0 && steps <= 4;
                if (!stage_e || ((t.w >> 30) & 1) || (!small && s4 > 4)) return;            // resident, or too long for the buffer (read directly)
                const int n16 = small ? 32 : 16 * s4;
                const uint4 *src = (const uint4 *)p.m.e_data + t.z;
                for (int j = lane; j < n16; j += 32) ps_cp16(stg + 16 * j, src + j);
                for (int j = lane; j < t.y; j += 32) ps_cp4(stg + 16 * n16 + 4 * j, gR + t.x + j);
                ps_cp_commit();
            };
            int tkA = next_item(&sm_ctr[0], lane);
            int4 tA = e_desc(tkA);
            if (tkA < n_et) e_stage(tA);
            int tkB = tkA < n_et ? next_item(&sm_ctr[0], lane) : n_et;
            int4 tB = e_desc(tkB);
            while (tkA < n_et) {
                const int steps = tA.w & 0xfff, lg = (tA.w >> 12) & 0xf, s4 = (steps + 3) >> 2;
                const bool small = lg == 0 && steps <= 4, res = (tA.w >> 30) & 1, staged = stage_e && !res && (small || s4 <= 4);
                const unsigned char *dsm = res ? v.cache + 16 * (size_t)tA.z : stg;     // where the tile's words are when they are in shared memory
                int tkC = n_et;
                int4 tC = make_int4(0, 0, 0, 0);
                auto advance = [&]() {          // tile A's words sit in registers (or it does not use the buffer): start the copy of tile B, ask for C
                    if (staged) __syncwarp();
                    if (tkB < n_et) { e_stage(tB); tkC = next_item(&sm_ctr[0], lane); tC = e_desc(tkC); }
                };
                if (staged) { ps_cp_wait(); __syncwarp(); }
                const double *th = v.theta;
                const int last = tA.y - 1;
                if (small) {
                    const bool insm = res || staged;
                    const uint4 w = insm ? *(const uint4 *)(dsm + 16 * lane) : __ldg((const uint4 *)p.m.e_data + tA.z + lane);
                    const uint32_t *rr = insm ? (const uint32_t *)(dsm + 512) : gR + tA.x;      // read counts: behind the index data in shared memory, else the compact class array
                    if (steps == 2) {
                        uint32_t r[4];
#pragma unroll
                        for (int g = 0; g < 4; g++) r[g] = insm ? rr[min(g * 32 + lane, last)] : __ldg(rr + min(g * 32 + lane, last));
                        advance();
                        const double a0 = th[w.x & 0xffffu], a1 = th[w.x >> 16], b0 = th[w.y & 0xffffu], b1 = th[w.y >> 16];
                        const double c0 = th[w.z & 0xffffu], c1 = th[w.z >> 16], d0 = th[w.w & 0xffffu], d1 = th[w.w >> 16];
                        const double sum[4] = {a0 + a1, b0 + b1, c0 + c1, d0 + d1};
#pragma unroll
                        for (int g = 0; g < 4; g++) {
                            const int c = g * 32 + lane;
                            if (c < tA.y) v.q[tA.x + c] = ps_q_of(r[g], sum[g]);
                        }
                    } else {
                        const uint32_t r0 = insm ? rr[min(lane, last)] : __ldg(rr + min(lane, last)), r1 = insm ? rr[min(32 + lane, last)] : __ldg(rr + min(32 + lane, last));
                        advance();
                        const double s0 = ps_gather4(th, make_uint2(w.x, w.y), 0.0), s1 = ps_gather4(th, make_uint2(w.z, w.w), 0.0);     // the pad slot of a 3-member class holds 0.0
                        if (lane < tA.y) v.q[tA.x + lane] = ps_q_of(r0, s0);
                        if (32 + lane < tA.y) v.q[tA.x + 32 + lane] = ps_q_of(r1, s1);
                    }
                } else {
                    // G = 1 << lg lanes per class; a lane's members come in chunks of 4 (chunk c of all lanes = 256 bytes)
                    const int G = 1 << lg, cls = lane >> lg;
                    double sum;
                    uint32_t r;
                    if (s4 <= 4 && (res || staged)) {
                        uint2 w[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) w[u] = *(const uint2 *)(dsm + 256 * min(u, s4 - 1) + 8 * lane);
                        r = ((const uint32_t *)(dsm + 256 * s4))[min(cls, last)];
                        advance();
                        sum = ps_gather4(th, w[0], 0.0);
                        if (s4 > 1) sum = ps_gather4(th, w[1], sum);
                        if (s4 > 2) sum = ps_gather4(th, w[2], sum);
                        if (s4 > 3) sum = ps_gather4(th, w[3], sum);
                    } else if (res) {
                        advance();
                        r = ((const uint32_t *)(dsm + 256 * s4))[min(cls, last)];
                        sum = ps_sum_chunks<true>(th, v.cache, nullptr, tA.z * 2 + lane, s4);
                    } else {
                        r = __ldg(gR + tA.x + min(cls, last));
                        advance();
                        sum = ps_sum_chunks<false>(th, nullptr, p.m.e_data, tA.z * 2 + lane, s4);       // four chunk loads in flight from the start
                    }
                    for (int dd = G >> 1; dd > 0; dd >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, dd);
                    if ((lane & (G - 1)) == 0 && cls < tA.y) v.q[tA.x + cls] = ps_q_of(r, sum);
                }
                tkA = tkB; tA = tB; tkB = tkC; tB = tC;
            }
        }
        PS_TRACE(2);
        if (it > 0) {
            d = read_dm(it - 1);
            if (*((volatile int *)p.abort_flag) != 0 || (p.stop_on_conv && d <= 1.0)) { stopped = true; break; }
        } else __syncthreads();
        {
            // ---- M-phase: partial row sums over the CTA's own classes; items longest first, same look-ahead as the E-phase ----
            auto m_desc = [&](int tk) -> int4 { return tk < n_mi ? mi[tk] : make_int4(0, 0, 0, 0); };
            auto m_stage = [&](const int4 &t) {
                const int len4 = ((t.w & 0x1fffffff) + 3) >> 2;
                if (!stage_m || ((t.w >> 29) & 1) || ((t.w >> 30) & 1) || len4 > 4) return;   // resident, a group of long rows, or a slice too long for the buffer
                const int n16 = 8 + 16 * len4;
                const uint4 *src = (const uint4 *)p.m.m_data + t.z;
                for (int j = lane; j < n16; j += 32) ps_cp16(stg + 16 * j, src + j);
                ps_cp_commit();
            };
            int tkA = next_item(&sm_ctr[1], lane);
            int4 tA = m_desc(tkA);
            if (tkA < n_mi) m_stage(tA);
            int tkB = tkA < n_mi ? next_item(&sm_ctr[1], lane) : n_mi;
            int4 tB = m_desc(tkB);
            while (tkA < n_mi) {
                const int len4 = ((tA.w & 0x1fffffff) + 3) >> 2;
                const bool res = (tA.w >> 29) & 1, group = (tA.w >> 30) & 1, staged = stage_m && !res && !group && len4 <= 4;
                const unsigned char *dsm = res ? v.cache + 16 * (size_t)tA.z : stg;
                int tkC = n_mi;
                int4 tC = make_int4(0, 0, 0, 0);
                auto advance = [&]() {
                    if (staged) __syncwarp();
                    if (tkB < n_mi) { m_stage(tB); tkC = next_item(&sm_ctr[1], lane); tC = m_desc(tkC); }
                };
                if (staged) { ps_cp_wait(); __syncwarp(); }
                if (group) {
                    advance();
                    if (res) ps_m_group<true>(p, v, tA, nullptr, lane, tag);
                    else ps_m_group<false>(p, v, tA, p.m.m_data, lane, tag);
                } else {
                    // a slice of 32 rows: 32 destinations, then chunks of 4 entries per lane
                    const double *q = v.q;
                    uint32_t dst;
                    double S;
                    if (len4 <= 4 && (res || staged)) {
                        uint2 w[4];
                        dst = ((const uint32_t *)dsm)[lane];
#pragma unroll
                        for (int u = 0; u < 4; u++) w[u] = *(const uint2 *)(dsm + 128 + 256 * min(u, len4 - 1) + 8 * lane);
                        advance();
                        S = ps_gather4(q, w[0], 0.0);                  // ascending class order
                        if (len4 > 1) S = ps_gather4(q, w[1], S);
                        if (len4 > 2) S = ps_gather4(q, w[2], S);
                        if (len4 > 3) S = ps_gather4(q, w[3], S);
                    } else if (res) {
                        advance();
                        dst = ((const uint32_t *)dsm)[lane];
                        S = ps_sum_chunks<true>(q, v.cache, nullptr, tA.z * 2 + 16 + lane, len4);
                    } else {
                        dst = __ldg((const uint32_t *)p.m.m_data + (size_t)tA.z * 4 + lane);
                        advance();
                        S = ps_sum_chunks<false>(q, nullptr, p.m.m_data, tA.z * 2 + 16 + lane, len4);
                    }
                    if (dst != PS_NONE) ps_emit(p, v, dst, S, tag);
                }
                tkA = tkB; tA = tB; tkB = tkC; tB = tC;
            }
        }
        __syncthreads();
        PS_TRACE(3);
        // owner update: own partial sum + the contributions of the other CTAs, in CTA order. The constants of a thread's first two rows are
        // requested first, then every incoming partial sum is fetched by its own thread (one round trip for all of them) and staged
        double dm = 0;
        const int i0 = threadIdx.x, i1 = threadIdx.x + EM_BLOCK;
        double2 ra0 = make_double2(0.0, 1.0), ra1 = ra0;
        int a0 = 0, b0 = 0, a1 = 0, b1 = 0;
        unsigned mk0 = 0, mk1 = 0;
        if (i0 < nrows) { ra0 = __ldg(p.m.row_RsA + row0 + i0); a0 = __ldg(p.m.inc_off + row0 + i0); b0 = __ldg(p.m.inc_off + row0 + i0 + 1); mk0 = (unsigned)__ldg(p.m.row_mask + row0 + i0); }
        if (i1 < nrows) { ra1 = __ldg(p.m.row_RsA + row0 + i1); a1 = __ldg(p.m.inc_off + row0 + i1); b1 = __ldg(p.m.inc_off + row0 + i1 + 1); mk1 = (unsigned)__ldg(p.m.row_mask + row0 + i1); }
        for (int e = threadIdx.x; e < nin; e += EM_BLOCK) s_in[e] = ll_load(part_slots + 16 * (size_t)(in0 + e), tag, p.abort_flag);
        __syncthreads();
        auto update = [&](int i, double2 ra, int e0, int e1, unsigned mk) {
            double Q = v.Q[i];
            for (int e = e0; e < e1; e++) Q += s_in[e - in0];
            const double th = v.theta[i];
            const double n = ra.x + th * Q;
            const double thn = fast_div(n, ra.y);
            v.theta[i] = thn;
            while (mk) {                                        // its readers are exactly its contributors: one store per rank that holds any
                const int r = __ffs(mk) - 1;
                mk &= mk - 1;
                ll_store(p.m.win[r] + p.m.th_off + 16 * (size_t)(row0 + i), thn, tag + 1u);
            }
            dm = fmax(dm, fast_div(fabs(thn - th) * ra.y, p.eps_abs + p.eps_rel * n));
        };
        if (i0 < nrows) update(i0, ra0, a0, b0, mk0);
        if (i1 < nrows) update(i1, ra1, a1, b1, mk1);
        for (int i = threadIdx.x + 2 * EM_BLOCK; i < nrows; i += EM_BLOCK)
            update(i, __ldg(p.m.row_RsA + row0 + i), __ldg(p.m.inc_off + row0 + i), __ldg(p.m.inc_off + row0 + i + 1), (unsigned)__ldg(p.m.row_mask + row0 + i));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dm = fmax(dm, __shfl_xor_sync(0xffffffffu, dm, o));
        if (lane == 0) sm_red[warp] = dm;
        __syncthreads();
        if ((int)threadIdx.x < p.m.nranks) {                     // one store per rank: every CTA of every rank reads every CTA's measure
            double bm = 0;
            for (int w = 0; w < EM_WARPS; w++) bm = fmax(bm, sm_red[w]);
            ll_store(p.m.win[threadIdx.x] + p.m.dm_off + 16 * (size_t)((it & 1) * Bt + b), bm, tag);
        }
        __syncthreads();                                         // sm_red is reused by read_dm
        PS_TRACE(4);
        it++;
    }
    if (it > 0 && !stopped) d = read_dm(it - 1);
    if (*((volatile int *)p.abort_flag) != 0) d = INFINITY;
    __syncthreads();
    for (int i = threadIdx.x; i < nrows; i += EM_BLOCK) p.m.theta[row0 + i] = v.theta[i];      // the copy the output kernels read
    if (blockIdx.x == 0 && threadIdx.x == 0) { *p.iters_done = it; *p.final_delta = d; }
}

__global__ void k_ps_theta_to_slots(int32_t P, const double *__restrict__ theta, unsigned char *__restrict__ th_slots, unsigned tag)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < P) ll_store(th_slots + 16 * (size_t)p, theta[p], tag);
}

int em_psum_attr(emsar_ctx *ctx)
{
    CU(cudaFuncSetAttribute(k_em_psum, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->em_smem_bytes));
    int nb = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_em_psum, EM_BLOCK, ctx->em_smem_bytes));
    if (nb < 1) { emsar_set_err("k_em_psum does not fit on an SM (%d bytes of shared memory)", ctx->em_smem_bytes); return EMSAR_ERR_CUDA; }
    return EMSAR_OK;
}

__global__ void k_ps_zero_foreign(int32_t P, int32_t lo, int32_t hi, double *__restrict__ theta)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < P && (p < lo || p >= hi)) theta[p] = 0.0;
}

int em_psum_launch(emsar_sample *s, int max_iter, int stop_on_conv, int *iters_done, double *final_delta, double *ms_out)
{
    emsar_ctx *ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    PsParams p;
    p.m = s->ps;
    p.eps_abs = s->opts.eps_abs; p.eps_rel = s->opts.eps_rel;
    p.max_iter = max_iter; p.stop_on_conv = stop_on_conv;
    p.iters_done = (int *)(ctx->d_barrier + 8);
    p.final_delta = (double *)(ctx->d_barrier + 10);
    p.abort_flag = (int *)(ctx->d_barrier + 14);
    p.trace = s->d_trace;
    const bool multi = s->ps.nranks > 1;
    CU(cudaMemsetAsync(ctx->d_barrier, 0, 256, st));                                     // scalars, abort flag
    const size_t slot_bytes = (size_t)p.m.dm_off + 2 * (size_t)s->ps.Bt * 16;
    void *slot_base = nullptr;
    if (multi) {
        // one sample over several GPUs: the slot arrays live in this rank's peer-mapped window; every launch starts from clean tags
        if (ctx->win_state != 1 || ctx->win_bytes < slot_bytes) { emsar_set_err("k_em_psum: the peer window is gone (call emsar_sample_prepare again)"); return EMSAR_ERR_STATE; }
        for (int r = 0; r < s->ps.nranks; r++) p.m.win[r] = (unsigned char *)ctx->peer_win[r];
        slot_base = ctx->win;
        CU(cudaMemsetAsync(ctx->win, 0, slot_bytes, st));
        p.tag0 = 0;
    } else {
        slot_base = s->d_slots;
        p.m.win[0] = (unsigned char *)s->d_slots;
        // the convergence slots are cleared at every launch (a launch that ended after 1-2 iterations leaves low tags behind); the
        // theta / partial-sum slots never are: their tags only grow within a sample
        CU(cudaMemsetAsync((unsigned char *)s->d_slots + p.m.dm_off, 0, 2 * (size_t)s->ps.Bt * 16, st));
        if ((unsigned)(s->slot_tag + (unsigned)max_iter + 4u) < s->slot_tag) {              // tag wrap: start over from clean slots
            CU(cudaMemsetAsync(s->d_slots, 0, s->slots_bytes, st));
            s->slot_tag = 0;
        }
        p.tag0 = s->slot_tag;
        s->slot_tag += (unsigned)max_iter + 2u;
    }
    if (s->ps.P > 0) { k_ps_theta_to_slots<<<(s->ps.P + 255) / 256, 256, 0, st>>>(s->ps.P, s->ps.theta, (unsigned char *)slot_base + p.m.th_off, p.tag0 + 1u); LAUNCHED(ctx); }
    if (multi) TRY(comm_barrier(ctx));          // every window is reset and carries theta(0) before any rank's kernel writes into it
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)s->ps.B);
    cfg.blockDim = dim3(EM_BLOCK);
    cfg.dynamicSmemBytes = (size_t)ctx->em_smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attrs[2];
    int na = 0;
    attrs[na].id = cudaLaunchAttributeCooperative;       // co-residency of all CTAs: they wait for each other's slots
    attrs[na].val.cooperative = 1;
    na++;
    if (ctx->l2_persist_bytes > 0 && slot_bytes > 0) {
        // the exchange slots (theta of shared rows, partial sums) stay resident in L2 while the index streams through
        size_t win = slot_bytes;
        if (win > (size_t)ctx->prop.accessPolicyMaxWindowSize) win = (size_t)ctx->prop.accessPolicyMaxWindowSize;
        attrs[na].id = cudaLaunchAttributeAccessPolicyWindow;
        attrs[na].val.accessPolicyWindow.base_ptr = slot_base;
        attrs[na].val.accessPolicyWindow.num_bytes = win;
        attrs[na].val.accessPolicyWindow.hitRatio = win <= ctx->l2_persist_bytes ? 1.0f : (float)ctx->l2_persist_bytes / (float)win;
        attrs[na].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attrs[na].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        na++;
    }
    cfg.attrs = attrs;
    cfg.numAttrs = na;
    CU(cudaFuncSetAttribute(k_em_psum, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->em_smem_bytes));    // per device, and other contexts may have changed it
    CU(cudaEventRecord(ctx->ev0, st));
    CU(cudaLaunchKernelEx(&cfg, k_em_psum, p));
    LAUNCHED(ctx);
    CU(cudaEventRecord(ctx->ev1, st));
    if (multi && s->ps.P > 0) {
        // every rank holds the final theta of its own rows: zero the others and add up (x + 0 is exact: all ranks end with the same bits)
        k_ps_zero_foreign<<<(s->ps.P + 255) / 256, 256, 0, st>>>(s->ps.P, s->ps_row_lo, s->ps_row_hi, s->ps.theta);
        LAUNCHED(ctx);
        TRY(comm_allreduce_f64(ctx, s->ps.theta, s->ps.theta, (size_t)s->ps.P));
    }
    int it = 0, aborted = 0; double fd = 0;
    CU(cudaMemcpyAsync(&it, p.iters_done, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&fd, p.final_delta, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&aborted, p.abort_flag, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (aborted) {
        emsar_set_err(multi ? "k_em_psum: a wait on peer memory timed out (a rank died or the ranks disagree on the call sequence)" : "k_em_psum: a wait on a tagged slot timed out (internal error)");
        return multi ? EMSAR_ERR_COMM : EMSAR_ERR_STATE;
    }
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    if (iters_done) *iters_done = it;
    if (final_delta) *final_delta = fd;
    if (ms_out) *ms_out = ms;
    return EMSAR_OK;
}
