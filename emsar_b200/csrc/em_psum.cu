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

// maximum over the warp of a value >= +0 that is not NaN: such doubles order like their bit patterns, two integer warp reductions
// (high word, then the low words of the lanes that hold the maximal high word) replace ten shuffles. A NaN wins.
__device__ __forceinline__ double ps_warp_max(double x)
{
    const unsigned hi = (unsigned)__double2hiint(x), lo = (unsigned)__double2loint(x);
    const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    return __hiloint2double((int)mh, (int)ml);
}

__device__ __forceinline__ double ps_q_of(uint32_t r, double s) { return s > 0 ? fast_div((double)r, s) : 0.0; }

__device__ __forceinline__ double ps_gather4(const double *a, uint2 w, double s)
{
    const double x0 = a[w.x & 0xffffu], x1 = a[w.x >> 16], x2 = a[w.y & 0xffffu], x3 = a[w.y >> 16];
    s += x0; s += x1; s += x2; s += x3;            // member order
    return s;
}

// the chunks of one lane (chunk c of all lanes = 256 bytes), four loads kept in flight: index data that is not resident comes from L2 or,
// for a model larger than L2, from HBM, and a lane that waits for every chunk in turn leaves the memory system idle
template <bool RES>
__device__ __forceinline__ double ps_sum_chunks(const double *a, const unsigned char *sm, const unsigned char *g, int o8, int n4)
{
    uint2 w[4];
#pragma unroll
    for (int u = 0; u < 4; u++) w[u] = u < n4 ? ps_ld64<RES>(sm, g, o8 + u * 32) : make_uint2(0u, 0u);
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

// ---- E-phase: one tile ----------------------------------------------------------------------------------------------------------------
template <bool RES>
__device__ __forceinline__ void ps_e_tile(const PsView &v, const int4 t, const unsigned char *gdat, const uint32_t *gR, int lane)
{
    const int steps = t.w & 0xfff, lg = (t.w >> 12) & 0xf;
    const int off16 = t.z;
    const double *th = v.theta;
    if (lg == 0 && steps <= 4) {
        const uint4 w = ps_ld128<RES>(v.cache, gdat, off16 + lane);
        // read counts: behind the 512 bytes of index data in a resident copy, in the compact class array otherwise. They are fetched together
        // with the members (one round trip to L2 for a tile that is not resident), for lanes beyond the tile's classes from a clamped index
        const int r4 = off16 * 4 + 128, last = t.y - 1;
        if (steps == 2) {
            uint32_t r[4];
#pragma unroll
            for (int g = 0; g < 4; g++) { const int c = min(g * 32 + lane, last); r[g] = RES ? ps_ld32<true>(v.cache, nullptr, r4 + c) : __ldg(gR + t.x + c); }
            const double a0 = th[w.x & 0xffffu], a1 = th[w.x >> 16], b0 = th[w.y & 0xffffu], b1 = th[w.y >> 16];
            const double c0 = th[w.z & 0xffffu], c1 = th[w.z >> 16], d0 = th[w.w & 0xffffu], d1 = th[w.w >> 16];
            const double s[4] = {a0 + a1, b0 + b1, c0 + c1, d0 + d1};
#pragma unroll
            for (int g = 0; g < 4; g++) {
                const int c = g * 32 + lane;
                if (c < t.y) v.q[t.x + c] = ps_q_of(r[g], s[g]);
            }
        } else {
            const int c0i = min(lane, last), c1i = min(32 + lane, last);
            const uint32_t r0 = RES ? ps_ld32<true>(v.cache, nullptr, r4 + c0i) : __ldg(gR + t.x + c0i);
            const uint32_t r1 = RES ? ps_ld32<true>(v.cache, nullptr, r4 + c1i) : __ldg(gR + t.x + c1i);
            const double a0 = th[w.x & 0xffffu], a1 = th[w.x >> 16], a2 = th[w.y & 0xffffu], a3 = th[w.y >> 16];
            const double b0 = th[w.z & 0xffffu], b1 = th[w.z >> 16], b2 = th[w.w & 0xffffu], b3 = th[w.w >> 16];
            double s0 = a0; s0 += a1; s0 += a2; s0 += a3;           // sequential member order (the pad slot holds 0.0)
            double s1 = b0; s1 += b1; s1 += b2; s1 += b3;
            if (lane < t.y) v.q[t.x + lane] = ps_q_of(r0, s0);
            if (32 + lane < t.y) v.q[t.x + 32 + lane] = ps_q_of(r1, s1);
        }
        return;
    }
    // G = 1 << lg lanes per class; a lane's members come in chunks of 4 (chunk c of all lanes = 256 bytes)
    const int steps4 = (steps + 3) >> 2;
    const int G = 1 << lg, cls = lane >> lg;
    if (lg == 0) {
        const uint32_t r = RES ? ps_ld32<true>(v.cache, nullptr, off16 * 4 + steps4 * 64 + min(cls, t.y - 1)) : __ldg(gR + t.x + min(cls, t.y - 1));
        const double s = ps_sum_chunks<RES>(th, v.cache, gdat, off16 * 2 + lane, steps4);
        if (cls < t.y) v.q[t.x + cls] = ps_q_of(r, s);
        return;
    }
    // several lanes per class: the tile covers nb consecutive row blocks (32 / G classes each, at most 64 classes in all), streamed as one run of
    // chunks with four loads in flight. Lane l fetches the read counts of classes l and 32 + l up front; a block's heads get theirs by shuffle.
    const int nb = max(1, (t.w >> 16) & 0xff), n4 = nb * steps4, cpb = 32 >> lg, o8 = off16 * 2 + lane;
    const uint32_t ra = RES ? ps_ld32<true>(v.cache, nullptr, off16 * 4 + n4 * 64 + min(lane, t.y - 1)) : __ldg(gR + t.x + min(lane, t.y - 1));
    const uint32_t rb = RES ? ps_ld32<true>(v.cache, nullptr, off16 * 4 + n4 * 64 + min(32 + lane, t.y - 1)) : __ldg(gR + t.x + min(32 + lane, t.y - 1));
    uint2 w[4];
#pragma unroll
    for (int u = 0; u < 4; u++) w[u] = ps_ld64<RES>(v.cache, gdat, o8 + min(u, n4 - 1) * 32);
    double s = 0;
    int left = steps4, blk = 0;                // chunks left in the current block
    for (int c = 0; c < n4; c += 4) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (c + u < n4) {
                const uint2 cur = w[u];
                if (c + u + 4 < n4) w[u] = ps_ld64<RES>(v.cache, gdat, o8 + (c + u + 4) * 32);
                s = ps_gather4(th, cur, s);
                if (--left == 0) {              // the block is complete: one class per group of G lanes
                    for (int d = G >> 1; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
                    const int ci = blk * cpb + cls;
                    const uint32_t r = __shfl_sync(0xffffffffu, ci < 32 ? ra : rb, ci & 31);
                    if ((lane & (G - 1)) == 0 && ci < t.y) v.q[t.x + ci] = ps_q_of(r, s);
                    s = 0; left = steps4; blk++;
                }
            }
        }
    }
}

// ---- M-phase: one item (partial row sums over the CTA's own classes) ------------------------------------------------------------------------
__device__ __forceinline__ void ps_emit(const PsParams &p, const PsView &v, uint32_t dst, double S, unsigned tag)
{
    if (!(dst & PS_REMOTE)) v.Q[dst] = S;
    else ll_store(p.m.win[(dst >> 28) & 7u] + p.m.part_off + 16 * (size_t)(dst & 0x0fffffffu), S, tag);     // the row's owner collects it, in its rank
}

template <bool RES>
__device__ __forceinline__ void ps_m_item(const PsParams &p, const PsView &v, const int4 t, const unsigned char *gdat, int lane, unsigned tag)
{
    const int len = t.w & 0x1fffffff;
    const int off16 = t.z;
    const double *q = v.q;
    if (((t.w >> 30) & 1) == 0) {
        const uint32_t dst = ps_ld32<RES>(v.cache, gdat, off16 * 4 + lane);
        const double S = ps_sum_chunks<RES>(q, v.cache, gdat, off16 * 2 + 16 + lane, (len + 3) >> 2);      // ascending class order
        if (dst != PS_NONE) ps_emit(p, v, dst, S, tag);
    } else {
        // a group of long rows, stored back to back: the warp streams the group's 32-bit words (two entries each) with eight loads per lane in
        // flight - header and first words are requested together - and closes a row where it ends: lane-strided partial sums, fixed shuffle
        // tree, lane r keeps row r's sum. A long row has more than 32 words, so at most one row ends inside a block of 32 words.
        const int n = t.y, hdr = (2 * n + 3) & ~3;
        const int base4 = off16 * 4 + hdr, Wt = t.x * 4 - hdr;           // the tail of the last 16 bytes is padded with zero-q pairs
        const int nch = (Wt + 31) >> 5;
        const uint32_t zz = (uint32_t)v.ncls | ((uint32_t)v.ncls << 16);
        const uint2 hw = lane < n ? ps_ld64<RES>(v.cache, gdat, off16 * 2 + lane) : make_uint2(0u, PS_NONE);      // {length, destination}
        uint32_t ring[8];
#pragma unroll
        for (int u = 0; u < 8; u++) ring[u] = (u * 32 + lane < Wt) ? ps_ld32<RES>(v.cache, gdat, base4 + u * 32 + lane) : zz;
        const int mywords = (int)((hw.x + 1) >> 1);
        int start = mywords;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, start, o); if (lane >= o) start += y; }
        start -= mywords;                                  // exclusive prefix: first word of the lane's row
        int r = 0, end_r = __shfl_sync(0xffffffffu, start, 0) + __shfl_sync(0xffffffffu, mywords, 0);
        double acc = 0, mine = 0;
        for (int c = 0; c < nch; c += 8) {
#pragma unroll
            for (int u = 0; u < 8; u++) {
                if (c + u < nch) {
                    const uint32_t cur = ring[u];
                    const int nx = (c + u + 8) * 32 + lane;
                    if (c + u + 8 < nch) ring[u] = nx < Wt ? ps_ld32<RES>(v.cache, gdat, base4 + nx) : zz;
                    const double x0 = q[cur & 0xffffu], x1 = q[cur >> 16];
                    if (r < n && end_r <= (c + u) * 32 + 32) {             // row r ends inside this block of words (warp-uniform)
                        const bool in_r = (c + u) * 32 + lane < end_r;
                        double sum = acc;
                        if (in_r) { sum += x0; sum += x1; }
#pragma unroll
                        for (int dd = 16; dd > 0; dd >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, dd);
                        if (lane == r) mine = sum;
                        acc = 0;
                        if (!in_r) { acc += x0; acc += x1; }
                        r++;
                        const int rr = min(r, n - 1);
                        end_r = __shfl_sync(0xffffffffu, start, rr) + __shfl_sync(0xffffffffu, mywords, rr);
                    } else { acc += x0; acc += x1; }
                }
            }
        }
        if (lane < n) ps_emit(p, v, hw.y, mine, tag);
    }
}

#define PS_TRACE(slot) do { if (p.trace && it == p.max_iter - 1 && threadIdx.x == 0) p.trace[blockIdx.x * 8 + (slot)] = gtime(); } while (0)

// SCHED bit 0: the E tiles are dealt to the warps in a fixed boustrophedon order over the cost-sorted list (no ticket, the next descriptor is
// requested while the current tile is processed) instead of being taken from the CTA's work queue.
template <int SCHED>
__global__ void __launch_bounds__(EM_BLOCK, 1) k_em_psum(PsParams p)
{
    extern __shared__ __align__(16) unsigned char sm_dyn[];
    __shared__ double sm_red[EM_WARPS];
    __shared__ double sm_bc;
    __shared__ int sm_ctr[2];
    __shared__ int sm_flag;          // bit 0: some CTA's convergence measure of the previous iteration is above 1; bit 1: abort
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = (int)blockIdx.x + p.m.block0;
    const int row0 = p.m.blk_row0[b], nrows = p.m.blk_row0[b + 1] - row0;
    const int cls0 = p.m.blk_cls0[b], ncls = p.m.blk_cls0[b + 1] - cls0;
    const int et0 = p.m.blk_etile0[b], n_et = p.m.blk_etile0[b + 1] - et0;
    const int mi0 = p.m.blk_mitem0[b], n_mi = p.m.blk_mitem0[b + 1] - mi0;
    const int hr0 = p.m.blk_hr0[b], nhr = p.m.blk_hr0[b + 1] - hr0;
    const int desc_smem = p.m.blk_desc_smem[b];
    const int in0 = p.m.inc_off[row0], nin = p.m.inc_off[row0 + nrows] - in0;       // partial sums this CTA receives per iteration
    const PsPlan pl = ps_smem_plan(desc_smem, n_et, n_mi, nrows, nhr, ncls, nin, 0);
    double *const s_in = (double *)(sm_dyn + pl.off_in);
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
            const int n16 = (lg == 0 && steps <= 4) ? 32 : 16 * ((steps + 3) >> 2) * max(1, (t.w >> 16) & 0xff);
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
    // the slots of a thread's first two halo rows stay in registers: one dependent round trip less at the top of every iteration
    const int hs0 = (int)threadIdx.x < nhr ? __ldg(p.m.halo_rows + hr0 + threadIdx.x) : 0;
    const int hs1 = (int)threadIdx.x + EM_BLOCK < nhr ? __ldg(p.m.halo_rows + hr0 + threadIdx.x + EM_BLOCK) : 0;
    while (it < p.max_iter) {
        const unsigned tag = p.tag0 + (unsigned)it + 1u;
        PS_TRACE(0);
        // theta of the halo rows, as their owners published it
        if ((int)threadIdx.x < nhr) v.theta[nrows + threadIdx.x] = ll_load(th_slots + 16 * (size_t)hs0, tag, p.abort_flag);
        if ((int)threadIdx.x + EM_BLOCK < nhr) v.theta[nrows + threadIdx.x + EM_BLOCK] = ll_load(th_slots + 16 * (size_t)hs1, tag, p.abort_flag);
        for (int i = threadIdx.x + 2 * EM_BLOCK; i < nhr; i += EM_BLOCK) v.theta[nrows + i] = ll_load(th_slots + 16 * (size_t)__ldg(p.m.halo_rows + hr0 + i), tag, p.abort_flag);
        if (threadIdx.x == 0) { sm_ctr[0] = 0; sm_ctr[1] = 0; sm_flag = 0; }
        __syncthreads();
        PS_TRACE(1);
        if (SCHED & 1) {
            int pos = warp;
            int4 t = pos < n_et ? et[n_et - 1 - pos] : make_int4(0, 0, 0, 0);
            for (int j = 0; pos < n_et; ) {
                j++;
                const int pos2 = j * 32 + ((j & 1) ? 31 - warp : warp);
                // a position beyond the list in an odd stripe belongs to no tile: the warp's list ends there (stripe j + 1 is beyond it as well)
                const int4 t2 = pos2 < n_et ? et[n_et - 1 - pos2] : make_int4(0, 0, 0, 0);
                if ((t.w >> 30) & 1) ps_e_tile<true>(v, t, nullptr, nullptr, lane);
                else ps_e_tile<false>(v, t, p.m.e_data, gR, lane);
                pos = pos2; t = t2;
            }
        } else if (desc_smem) {
            for (int tk = next_item(&sm_ctr[0], lane); tk < n_et; tk = next_item(&sm_ctr[0], lane)) {
                const int4 t = et[n_et - 1 - tk];               // tiles are ordered by cardinality: heaviest first
                if ((t.w >> 30) & 1) ps_e_tile<true>(v, t, nullptr, nullptr, lane);
                else ps_e_tile<false>(v, t, p.m.e_data, gR, lane);
            }
        } else {
            // descriptors in global memory (many tiles): the next ticket's descriptor is requested before the current tile is processed
            int tk = next_item(&sm_ctr[0], lane);
            int4 t = tk < n_et ? __ldg(et + (n_et - 1 - tk)) : make_int4(0, 0, 0, 0);
            while (tk < n_et) {
                const int tk2 = next_item(&sm_ctr[0], lane);
                const int4 t2 = tk2 < n_et ? __ldg(et + (n_et - 1 - tk2)) : make_int4(0, 0, 0, 0);
                if ((t.w >> 30) & 1) ps_e_tile<true>(v, t, nullptr, nullptr, lane);
                else ps_e_tile<false>(v, t, p.m.e_data, gR, lane);
                tk = tk2; t = t2;
            }
        }
        PS_TRACE(2);
        __syncthreads();
        // the convergence measure of the previous iteration: every CTA's slot is polled by one thread, by the first warps, before they join the
        // M-phase queue - the other warps start on the items at once, and nobody waits for the result before the end of the M-phase (theta is
        // untouched until the U-phase, so that stopping after the M-phase leaves exactly the previous iteration's state)
        if (it > 0 && (int)threadIdx.x < Bt) {
            double x = 0;
            for (int i = threadIdx.x; i < Bt; i += EM_BLOCK) x = fmax(x, ll_load(dm_slots + 16 * (size_t)(((it - 1) & 1) * Bt + i), tag - 1u, p.abort_flag));
            if (!(x <= 1.0)) atomicOr(&sm_flag, 1);
        }
        if (threadIdx.x == 0 && *((volatile int *)p.abort_flag) != 0) atomicOr(&sm_flag, 2);
        if (desc_smem) {
            for (int tk = next_item(&sm_ctr[1], lane); tk < n_mi; tk = next_item(&sm_ctr[1], lane)) {
                const int4 t = mi[tk];                          // items: groups of long rows, then slices, longest first
                if ((t.w >> 29) & 1) ps_m_item<true>(p, v, t, nullptr, lane, tag);
                else ps_m_item<false>(p, v, t, p.m.m_data, lane, tag);
            }
        } else {
            int tk = next_item(&sm_ctr[1], lane);
            int4 t = tk < n_mi ? __ldg(mi + tk) : make_int4(0, 0, 0, 0);
            while (tk < n_mi) {
                const int tk2 = next_item(&sm_ctr[1], lane);
                const int4 t2 = tk2 < n_mi ? __ldg(mi + tk2) : make_int4(0, 0, 0, 0);
                if ((t.w >> 29) & 1) ps_m_item<true>(p, v, t, nullptr, lane, tag);
                else ps_m_item<false>(p, v, t, p.m.m_data, lane, tag);
                tk = tk2; t = t2;
            }
        }
        __syncthreads();
        PS_TRACE(3);
        {
            const int fl = sm_flag;
            if ((fl & 2) || (p.stop_on_conv && it > 0 && !(fl & 1))) break;      // the same decision in every CTA: they all read the same slots
        }
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
        dm = ps_warp_max(dm);
        if (lane == 0) sm_red[warp] = dm;
        __syncthreads();
        if (warp == 0) {                                         // one store per rank: every CTA of every rank reads every CTA's measure
            const double bm = ps_warp_max(sm_red[lane]);
            if (lane < p.m.nranks) ll_store(p.m.win[lane] + p.m.dm_off + 16 * (size_t)((it & 1) * Bt + b), bm, tag);
        }
        PS_TRACE(4);                                             // (sm_red is written again three barriers from here)
        it++;
    }
    __syncthreads();
    if (it > 0) d = read_dm(it - 1);
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

// EMSAR_PS_SCHED: 0 (default) = E tiles taken from the CTA's work queue, 1 = dealt to the warps in a fixed order (measured slower, profiles/r2o:
// 23.6 vs 22.0 us at config #2, 194 vs 171 at config #5 - the queue's balance inside a CTA is worth more than the 20 instructions of a ticket)
static int ps_sched()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("EMSAR_PS_SCHED"); v = e ? (atoi(e) & 1) : 0; }
    return v;
}
typedef void (*ps_kernel_t)(PsParams);
static ps_kernel_t ps_kernel() { return ps_sched() ? k_em_psum<1> : k_em_psum<0>; }

int em_psum_attr(emsar_ctx *ctx)
{
    CU(cudaFuncSetAttribute(ps_kernel(), cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->em_smem_bytes));
    int nb = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, ps_kernel(), EM_BLOCK, ctx->em_smem_bytes));
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
    CU(cudaFuncSetAttribute(ps_kernel(), cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->em_smem_bytes));    // per device, and other contexts may have changed it
    CU(cudaEventRecord(ctx->ev0, st));
    CU(cudaLaunchKernelEx(&cfg, ps_kernel(), p));
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

// tuning aid (profiles/trace_psum.py): per CTA, 16 numbers that describe its share of the packed model -
// {rows, halo rows, classes, incoming partial sums, E tiles, M items, resident 16-byte units, descriptors in smem,
//  cardinality-2 tiles, cardinality-3/4 tiles, chunk steps of one-lane tiles, chunk steps of multi-lane tiles, resident E tiles,
//  chunk steps of slices, 32-word blocks of groups, resident M items}
extern "C" int emsar_debug_psum_blocks(emsar_sample *s, int32_t *out, int *n_blocks)
{
    if (!s || !s->prepared || !s->use_psum) return EMSAR_ERR_STATE;
    TRY(ctx_use(s->ctx));
    const PsModel &m = s->ps;
    const int B = m.B, b0 = m.block0;
    CU(cudaStreamSynchronize(s->ctx->stream));
    std::vector<int32_t> row0(B + 1), cls0(B + 1), et0(B + 1), mi0(B + 1), hr0(B + 1), res(B + 1), desc(B + 1);
    CU(cudaMemcpy(row0.data(), m.blk_row0 + b0, (size_t)(B + 1) * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(cls0.data(), m.blk_cls0 + b0, (size_t)(B + 1) * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(et0.data(), m.blk_etile0 + b0, (size_t)(B + 1) * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(mi0.data(), m.blk_mitem0 + b0, (size_t)(B + 1) * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(hr0.data(), m.blk_hr0 + b0, (size_t)(B + 1) * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(res.data(), m.blk_res16 + b0, (size_t)B * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(desc.data(), m.blk_desc_smem + b0, (size_t)B * 4, cudaMemcpyDeviceToHost));
    std::vector<int4> et((size_t)(et0[B] - et0[0]) + 1), mi((size_t)(mi0[B] - mi0[0]) + 1);
    if (et0[B] > et0[0]) CU(cudaMemcpy(et.data(), m.e_tiles + et0[0], (size_t)(et0[B] - et0[0]) * 16, cudaMemcpyDeviceToHost));
    if (mi0[B] > mi0[0]) CU(cudaMemcpy(mi.data(), m.m_items + mi0[0], (size_t)(mi0[B] - mi0[0]) * 16, cudaMemcpyDeviceToHost));
    for (int b = 0; b < B; b++) {
        int32_t *o = out + 16 * (size_t)b;
        int32_t in_lo = 0, in_hi = 0;
        CU(cudaMemcpy(&in_lo, m.inc_off + row0[b], 4, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(&in_hi, m.inc_off + row0[b + 1], 4, cudaMemcpyDeviceToHost));
        o[0] = row0[b + 1] - row0[b]; o[1] = hr0[b + 1] - hr0[b]; o[2] = cls0[b + 1] - cls0[b]; o[3] = in_hi - in_lo;
        o[4] = et0[b + 1] - et0[b]; o[5] = mi0[b + 1] - mi0[b]; o[6] = res[b]; o[7] = desc[b];
        for (int k = 8; k < 16; k++) o[k] = 0;
        for (int i = et0[b] - et0[0]; i < et0[b + 1] - et0[0]; i++) {
            const int4 t = et[(size_t)i];
            const int steps = t.w & 0xfff, lg = (t.w >> 12) & 0xf, nb = std::max(1, (t.w >> 16) & 0xff);
            if (lg == 0 && steps <= 4) o[steps == 2 ? 8 : 9]++;
            else if (lg == 0) o[10] += (steps + 3) >> 2;
            else o[11] += nb * ((steps + 3) >> 2);
            if ((t.w >> 30) & 1) o[12]++;
        }
        for (int i = mi0[b] - mi0[0]; i < mi0[b + 1] - mi0[0]; i++) {
            const int4 t = mi[(size_t)i];
            const int len = t.w & 0x1fffffff;
            if ((t.w >> 30) & 1) o[14] += (t.x * 4 + 31) / 32; else o[13] += (len + 3) >> 2;
            if ((t.w >> 29) & 1) o[15]++;
        }
    }
    *n_blocks = B;
    return EMSAR_OK;
}
