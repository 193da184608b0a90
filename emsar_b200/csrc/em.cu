// The EM loop: one persistent cooperative kernel, one CTA per SM. Each CTA owns a contiguous range of transcripts
// (rows) and the classes whose first member lies in it, and keeps theta / q of what it owns in SHARED MEMORY: the
// gathers of both phases hit 32 independent banks instead of one L1 line per wavefront; only references that leave
// the range (halo) go through L2-resident global memory. Per iteration:
//   E-phase  q_c = R_c / sum_{t in c} theta_t        class-major, binned by cardinality
//   M-phase  theta_t' = (Rs_t + theta_t * sum_{c∋t} q_c) / A_t   transposed CSR, deterministic segmented reduction,
//            fused with the convergence measure  max_t |dtheta_t| A_t / (eps_abs + eps_rel n_t)
// until convergence or max_iter, with no host round trip.  It replaces run_MLE_threads / MLE_range / MLE / Fp / lambdap of the
// reference (emsar_functions.c:2946-3126), which reach the same Poisson-likelihood optimum by a randomized pattern search.
// No floating-point atomics anywhere: every sum has an order fixed by the packed layout alone.
// Instantiations (k_em_persistent<MODE>):
//   3  default: no grid barrier; halo theta / q and the convergence measure travel through tagged 16-byte slots (ll_store / ll_load)
//   1  two grid barriers per iteration (overflow: some halo row / class has no shared-memory slot; EMSAR_EM_MODE=barrier)
//   0  as 1 with TMA-pipelined index streams instead of the resident index cache (EMSAR_EM_MODE=pipe)
//   2  one sample class-sharded over several GPUs: the per-iteration all-reduce runs inside the kernel over NVLink peer memory
#include <cooperative_groups.h>
#include <math.h>

#include "common.cuh"
#include "em_common.cuh"

int sample_ensure_sets(emsar_sample *s);

struct EmParams {
    EmModel m;
    double eps_abs, eps_rel;
    int max_iter, stop_on_conv;
    unsigned *bar;                 // one generation flag per CTA, 128 bytes apart
    unsigned long long *dmax;      // [2] alternating slots for the reduced delta (bit pattern of a double >= 0)
    int *iters_done;
    double *final_delta;
    unsigned long long *trace;     // optional [B*8] globaltimer stamps of the last iteration (tuning aid)
    // barrier-free single-GPU kernel (k_em_persistent<3>): halo theta / q travel through tagged slots instead of grid barriers
    unsigned char *th_slots, *q_slots, *dm_slots;    // [P], [C_a], [2 * B] 16-byte tagged slots
    unsigned df_tag0;
    int *df_abort;
    // class-sharded sample (k_em_persistent<2> only): this rank holds a class range; per-row sums are exchanged in natural order
    struct Shard {
        int rank, nranks;          // nranks == 0: not sharded
        int S;                     // rows per owner slice: rank r updates the natural rows [r*S, min(P, (r+1)*S))
        int fused;                 // 1: all-reduce inside the kernel over peer memory; 0: one pass, sums to q_out (NCCL path)
        unsigned char *win[EMSAR_MAX_RANKS];   // every rank's window (comm.cu): dm | theta | xbuf, 16-byte tagged slots
        long long theta_off, xbuf_off;         // byte offsets inside a window
        long long theta_cap;                   // slots per theta parity array
        unsigned tag0;             // tags of this launch start above tag0 (windows are zeroed before every launch)
        int *abort_flag;           // set when a wait ran out of patience (a peer died): every wait then falls through
        double *q_out;             // NCCL path: [P] partial row sums in natural order
    } sh;
};

// tuning aid: globaltimer stamps of the last iteration of a launch, per CTA (see emsar_debug_em_trace)
#define TRACE(slot) do { if (p.trace && it == p.max_iter - 1 && threadIdx.x == 0) p.trace[blockIdx.x * 8 + (slot)] = gtime(); } while (0)

// Grid barrier for the co-resident CTAs (cooperative launch), without atomics: every CTA publishes the barrier's
// generation in its own 128-byte line with a release store; thread i of every CTA polls CTA i's line with acquire
// loads. The release / acquire pairs make the phase's global writes (theta / q write-through) visible to the CTAs that
// read them as halo. Latency: one store + one poll round trip after the last arrival.
__device__ __forceinline__ void grid_barrier(unsigned *flags, unsigned nblocks, unsigned gen)
{
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("st.release.gpu.u32 [%0], %1;" ::"l"(flags + blockIdx.x * 32), "r"(gen) : "memory");
    for (unsigned i = threadIdx.x; i < nblocks; i += blockDim.x) {
        unsigned cur;
        do { asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(cur) : "l"(flags + i * 32) : "memory"); } while ((int)(cur - gen) < 0);
    }
    __syncthreads();
}

// ---- shared-memory resident slices ---------------------------------------------------------------------
struct BlockView {
    double *sm_theta;          // theta of the rows this CTA owns, then its halo rows
    double *sm_q;              // q of the resident classes this CTA owns; sm_q[nres] == 0 (padding target); then halo classes
    const double2 *sm_rsa;     // {Rs, A} of the rows this CTA owns
    const int4 *sm_etiles;     // this CTA's E tile descriptors
    const int4 *sm_mitems;     // this CTA's M items
    double *sm_stage;          // sharded mode: partial row sums of this CTA in natural order (reuses the {Rs,A} area)
    const int *sm_noff;        // sharded mode: natural offset (inside the CTA's range) of every row slot
    int row0, nrows, cls0, nres, nhr, nhc;
};

// branch-free: one generic load from either the CTA's shared slice or the global copy (overflow only)
template <bool SH>
__device__ __forceinline__ double load_theta(const EmParams &p, const BlockView &v, int enc)
{
    if (SH && enc < 0) return __ldcg(p.m.theta + ~enc);      // sharded overflow path: the permuted global copy is rewritten every iteration
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
    const double val = s > 0 ? fast_div(r, s) : 0.0;
    const int loc = j - v.cls0;
    if (loc < v.nres) v.sm_q[loc] = val;
    if (rflag & 0x80000000u) p.m.q[j] = val;        // a row of another CTA reads it, or it does not fit in shared memory
}

// ---- chunk staging: a contiguous piece of an int32 stream -> shared memory with one TMA bulk copy (cp.async.bulk), -----
// completion signalled on an mbarrier. [g, g+n) is fetched from its 16-byte-aligned floor; element i then sits at
// buf[shift + i]. Geometry (shift, bytes) is computed by every thread; only thread 0 issues the copy.
__device__ __forceinline__ int stage_shift(const void *g) { return (int)(((uintptr_t)g & 15) >> 2); }
__device__ __forceinline__ uint32_t stage_bytes(const void *g, int n) { return (uint32_t)((n + stage_shift(g) + 3) >> 2) * 16u; }
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *g, uint32_t bytes, unsigned long long *bar)
{
    const uintptr_t ga = (uintptr_t)g & ~(uintptr_t)15;
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(ga), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity)
{
    asm volatile("{\n.reg .pred p;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra LAB_DONE;\nbra LAB_WAIT;\nLAB_DONE:\n}" ::"r"(smem_u32(bar)),
                 "r"(parity)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// barrier among the consumer warps only (the producer warp never joins it)
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"((EM_WARPS - 1) * 32) : "memory"); }
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- E-phase: q_c = R_c / sum of theta over the class members --------------------------------------------------
// tids / rfl point at the tile's member indices / read counts (shared memory when staged, global otherwise)
template <bool SH, int K, int G>
__device__ __forceinline__ void etile_small(const EmParams &p, const BlockView &v, int4 tile, const int *tids, const uint32_t *rfl, int lane)
{
    int t[K * G];
    uint32_t rf[G];
#pragma unroll
    for (int u = 0; u < K * G; u++) t[u] = (u / K) * 32 < tile.y ? tids[u * 32 + lane] : 0;
#pragma unroll
    for (int g = 0; g < G; g++) rf[g] = (g * 32 + lane < tile.y) ? rfl[g * 32 + lane] : 0u;
    double x[K * G];
#pragma unroll
    for (int u = 0; u < K * G; u++) x[u] = load_theta<SH>(p, v, t[u]);
#pragma unroll
    for (int g = 0; g < G; g++) {
        double s = 0;
#pragma unroll
        for (int j = 0; j < K; j++) s += x[g * K + j];       // sequential member order
        if (g * 32 + lane < tile.y) store_q(p, v, tile.x + g * 32 + lane, rf[g], s);
    }
}

template <bool SH>
__device__ __forceinline__ void e_tile(const EmParams &p, const BlockView &v, int4 tile, const int *tids, const uint32_t *rfl, int lane)
{
    const int steps = tile.w & 0xfff, lg = (tile.w >> 12) & 0xf;
    if (lg == 0) {
        if (steps == 2) { etile_small<SH, 2, 4>(p, v, tile, tids, rfl, lane); return; }
        if (steps == 3) { etile_small<SH, 3, 2>(p, v, tile, tids, rfl, lane); return; }
        if (steps == 4) { etile_small<SH, 4, 2>(p, v, tile, tids, rfl, lane); return; }
    }
    // G = 1 << lg lanes per class; step j of all 32 lanes is one 128-byte line
    const int cls = lane >> lg, G = 1 << lg;
    const bool head = cls < tile.y && (lane & (G - 1)) == 0;
    uint32_t rf = 0;
    if (head) rf = rfl[cls];
    tids += lane;
    double s = 0;
    int j = 0;
    for (; j + 4 <= steps; j += 4) {
        int t[4];
#pragma unroll
        for (int u = 0; u < 4; u++) t[u] = tids[(j + u) * 32];
        double x[4];
#pragma unroll
        for (int u = 0; u < 4; u++) x[u] = load_theta<SH>(p, v, t[u]);
#pragma unroll
        for (int u = 0; u < 4; u++) s += x[u];
    }
    for (; j < steps; j++) s += load_theta<SH>(p, v, tids[j * 32]);
    for (int d = G >> 1; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (head) store_q(p, v, tile.x + cls, rf, s);
}

// ---- M-phase: theta_t' = (Rs_t + theta_t * sum of q over the row) / A_t, fused convergence measure --------------
template <bool SH>
__device__ __forceinline__ double m_update(const EmParams &p, const BlockView &v, int slot, double Q)
{
    if (SH) { v.sm_stage[v.sm_noff[slot]] = Q; return 0.0; }     // sharded: stage the partial sum in natural order
    const double2 ra = v.sm_rsa[slot];
    const double th = v.sm_theta[slot];
    const double n = ra.x + th * Q;
    const double thn = fast_div(n, ra.y);
    v.sm_theta[slot] = thn;
    p.m.theta[v.row0 + slot] = thn;                 // write-through: halo readers and the final result
    return fast_div(fabs(thn - th) * ra.y, p.eps_abs + p.eps_rel * n);
}

template <bool SH>
__device__ __forceinline__ double m_item(const EmParams &p, const BlockView &v, int4 it, const int *ent, int lane)
{
    const int len = it.w & 0x3fffffff;
    double d = 0;
    if ((it.w >> 30) == 0) {
        // a slice of 32 rows stored transposed: one thread per row, sequential sum in ascending class order
        ent += lane;
        double Q = 0;
        int j = 0;
        for (; j + 4 <= len; j += 4) {
            int c[4];
#pragma unroll
            for (int u = 0; u < 4; u++) c[u] = ent[(j + u) * 32];
            double x[4];
#pragma unroll
            for (int u = 0; u < 4; u++) x[u] = load_q(p, v, c[u]);
#pragma unroll
            for (int u = 0; u < 4; u++) Q += x[u];       // ascending class order
        }
        for (; j < len; j++) Q += load_q(p, v, ent[j * 32]);
        if (lane < it.y) d = m_update<SH>(p, v, it.x + lane, Q);
    } else {
        // a group of long rows: header = the rows' lengths, then their entries; the warp reduces one row at a time with a
        // fixed shuffle tree, lane r keeps row r's sum, then all rows are updated together
        const int n = it.y;
        const int mylen = lane < n ? ent[lane] : 0;
        int start = mylen;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, start, o); if (lane >= o) start += y; }
        start += n - mylen;                          // exclusive prefix, after the header
        double mine = 0;
        for (int r = 0; r < n; r++) {
            const int a = __shfl_sync(0xffffffffu, start, r), L = __shfl_sync(0xffffffffu, mylen, r);
            double s = 0;
#pragma unroll 4
            for (int e = lane; e < L; e += 32) s += load_q(p, v, ent[a + e]);
#pragma unroll
            for (int dd = 16; dd > 0; dd >>= 1) s += __shfl_xor_sync(0xffffffffu, s, dd);
            if (lane == r) mine = s;
        }
        if (lane < n) d = m_update<SH>(p, v, it.x + lane, mine);
    }
    return d;
}

// ---- fast path: the chunk is staged and every index of this CTA is a shared-memory slot. All addresses are 32-bit
// offsets from the CTA's dynamic shared memory, so the loops are LDS + LDS.64 + DADD with almost no address arithmetic.
struct SmView {
    unsigned char *base;   // sm_dyn
    int theta8, q8, rsa16; // element offsets of theta / q / {Rs,A} inside sm_dyn
    int noff4;             // sharded mode: int offset of the slot -> natural offset table (second half of the {Rs,A} area)
    unsigned tag;          // data-flow mode: tag of the current iteration
    int zslot;             // theta slot that holds 0.0 (padding target)
    int row0, cls0, nres;
};
#define S32(v) ((const int *)(v).base)
#define S64(v) ((double *)(v).base)
// where an item's int32 indices come from: the staged chunk in shared memory, or straight from global memory (L2)
struct IdxS { const unsigned char *base; __device__ __forceinline__ int operator()(int i) const { return ((const int *)base)[i]; } };
struct IdxG { const int32_t *g; __device__ __forceinline__ int operator()(int i) const { return __ldg(g + i); } };

template <int DF>
__device__ __forceinline__ void f_store_q(const EmParams &p, const SmView &v, int j, uint32_t rflag, double s)
{
    const double r = (double)(rflag & 0x7fffffffu);
    const double val = s > 0 ? fast_div(r, s) : 0.0;
    S64(v)[v.q8 + (j - v.cls0)] = val;              // fast path: every owned class is resident
    if (rflag & 0x80000000u) {                      // a row of another CTA reads it
        if (DF) ll_store(p.q_slots + 16 * (size_t)j, val, v.tag);
        else p.m.q[j] = val;
    }
}

template <int K, int G, int DF, class IT, class IR>
__device__ __forceinline__ void f_etile_small(const EmParams &p, const SmView &v, int4 tile, IT T_, IR R_, int ti, int ri, int lane)
{
    int t[K * G];
    uint32_t rf[G];
#pragma unroll
    for (int u = 0; u < K * G; u++) t[u] = (u / K) * 32 < tile.y ? T_(ti + u * 32 + lane) : 0;
#pragma unroll
    for (int g = 0; g < G; g++) rf[g] = (g * 32 + lane < tile.y) ? (uint32_t)R_(ri + g * 32 + lane) : 0u;
    double x[K * G];
#pragma unroll
    for (int u = 0; u < K * G; u++) x[u] = S64(v)[v.theta8 + t[u]];
#pragma unroll
    for (int g = 0; g < G; g++) {
        double s = 0;
#pragma unroll
        for (int j = 0; j < K; j++) s += x[g * K + j];       // sequential member order
        if (g * 32 + lane < tile.y) f_store_q<DF>(p, v, tile.x + g * 32 + lane, rf[g], s);
    }
}

template <int DF = 0, class IT, class IR>
__device__ __forceinline__ void f_e_tile(const EmParams &p, const SmView &v, int4 tile, IT T_, IR R_, int ti, int ri, int lane)
{
    const int steps = tile.w & 0xfff, lg = (tile.w >> 12) & 0xf;
    if (lg == 0) {
        if (steps == 2) { f_etile_small<2, 4, DF>(p, v, tile, T_, R_, ti, ri, lane); return; }
        if (steps == 3) { f_etile_small<3, 2, DF>(p, v, tile, T_, R_, ti, ri, lane); return; }
        if (steps == 4) { f_etile_small<4, 2, DF>(p, v, tile, T_, R_, ti, ri, lane); return; }
    }
    // G = 1 << lg lanes per class; step j of all 32 lanes is one 128-byte line
    const int cls = lane >> lg, G = 1 << lg;
    const bool head = cls < tile.y && (lane & (G - 1)) == 0;
    uint32_t rf = 0;
    if (head) rf = (uint32_t)R_(ri + cls);
    ti += lane;
    double s = 0;
    int j = 0;
    for (; j + 8 <= steps; j += 8) {
        int t[8];
#pragma unroll
        for (int u = 0; u < 8; u++) t[u] = T_(ti + (j + u) * 32);
        double x[8];
#pragma unroll
        for (int u = 0; u < 8; u++) x[u] = S64(v)[v.theta8 + t[u]];
#pragma unroll
        for (int u = 0; u < 8; u++) s += x[u];
    }
    for (; j + 4 <= steps; j += 4) {
        const int t0 = T_(ti + j * 32), t1 = T_(ti + j * 32 + 32), t2 = T_(ti + j * 32 + 64), t3 = T_(ti + j * 32 + 96);
        const double x0 = S64(v)[v.theta8 + t0], x1 = S64(v)[v.theta8 + t1], x2 = S64(v)[v.theta8 + t2], x3 = S64(v)[v.theta8 + t3];
        s += x0; s += x1; s += x2; s += x3;
    }
    if (j < steps) {                       // 1..3 rows left: one predicated chunk, the missing rows read the zero-theta slot
        const int r = steps - j;
        const int t0 = T_(ti + j * 32), t1 = r > 1 ? T_(ti + j * 32 + 32) : v.zslot, t2 = r > 2 ? T_(ti + j * 32 + 64) : v.zslot;
        const double x0 = S64(v)[v.theta8 + t0], x1 = S64(v)[v.theta8 + t1], x2 = S64(v)[v.theta8 + t2];
        s += x0; s += x1; s += x2;
    }
    for (int d = G >> 1; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (head) f_store_q<DF>(p, v, tile.x + cls, rf, s);
}

template <int MD>
__device__ __forceinline__ double f_m_update(const EmParams &p, const SmView &v, int slot, double Q)
{
    if (MD == 1) { S64(v)[v.rsa16 * 2 + ((const int *)v.base)[v.noff4 + slot]] = Q; return 0.0; }   // stage = the {Rs,A} area (unused when sharded)
    const double2 ra = ((const double2 *)v.base)[v.rsa16 + slot];
    const double th = S64(v)[v.theta8 + slot];
    const double n = ra.x + th * Q;
    const double thn = fast_div(n, ra.y);
    S64(v)[v.theta8 + slot] = thn;
    if (MD == 2) ll_store(p.th_slots + 16 * (size_t)(v.row0 + slot), thn, v.tag + 1);   // halo readers of the next iteration
    else p.m.theta[v.row0 + slot] = thn;            // write-through: halo readers and the final result
    return fast_div(fabs(thn - th) * ra.y, p.eps_abs + p.eps_rel * n);
}

template <int MD, class IT>
__device__ __forceinline__ double f_m_item(const EmParams &p, const SmView &v, int4 it, IT T_, int ei, int lane)
{
    const int len = it.w & 0x3fffffff;
    double d = 0;
    if ((it.w >> 30) == 0) {
        ei += lane;
        double Q = 0;
        int j = 0;
        for (; j + 8 <= len; j += 8) {
            int c[8];
#pragma unroll
            for (int u = 0; u < 8; u++) c[u] = T_(ei + (j + u) * 32);
            double x[8];
#pragma unroll
            for (int u = 0; u < 8; u++) x[u] = S64(v)[v.q8 + c[u]];
#pragma unroll
            for (int u = 0; u < 8; u++) Q += x[u];       // ascending class order
        }
        for (; j + 4 <= len; j += 4) {
            const int c0 = T_(ei + j * 32), c1 = T_(ei + j * 32 + 32), c2 = T_(ei + j * 32 + 64), c3 = T_(ei + j * 32 + 96);
            const double x0 = S64(v)[v.q8 + c0], x1 = S64(v)[v.q8 + c1], x2 = S64(v)[v.q8 + c2], x3 = S64(v)[v.q8 + c3];
            Q += x0; Q += x1; Q += x2; Q += x3;          // ascending class order
        }
        if (j < len) {                     // 1..3 entries left: one predicated chunk, the missing ones read the zero slot of q
            const int r = len - j;
            const int c0 = T_(ei + j * 32), c1 = r > 1 ? T_(ei + j * 32 + 32) : v.nres, c2 = r > 2 ? T_(ei + j * 32 + 64) : v.nres;
            const double x0 = S64(v)[v.q8 + c0], x1 = S64(v)[v.q8 + c1], x2 = S64(v)[v.q8 + c2];
            Q += x0; Q += x1; Q += x2;             // ascending class order
        }
        if (lane < it.y) d = f_m_update<MD>(p, v, it.x + lane, Q);
    } else {
        const int n = it.y;
        const int mylen = lane < n ? T_(ei + lane) : 0;
        int start = mylen;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, start, o); if (lane >= o) start += y; }
        start += n - mylen;
        double mine = 0;
        for (int r = 0; r < n; r++) {
            const int a = ei + __shfl_sync(0xffffffffu, start, r), L = __shfl_sync(0xffffffffu, mylen, r);
            double s = 0;
#pragma unroll 4
            for (int e = lane; e < L; e += 32) s += S64(v)[v.q8 + T_(a + e)];
#pragma unroll
            for (int dd = 16; dd > 0; dd >>= 1) s += __shfl_xor_sync(0xffffffffu, s, dd);
            if (lane == r) mine = s;
        }
        if (lane < n) d = f_m_update<MD>(p, v, it.x + lane, mine);
    }
    return d;
}

// MODE 0: TMA-pipelined index streams; 1: direct (resident index cache + L2); 2: direct, class-sharded over several GPUs;
// 3: direct without grid barriers (halo theta / q and the convergence measure travel through tagged slots)
template <int MODE>
__global__ void __launch_bounds__(EM_BLOCK, 1) k_em_persistent(EmParams p)
{
    constexpr bool DIRECT = MODE != 0;
    constexpr bool SHARDED = MODE == 2;
    extern __shared__ __align__(16) unsigned char sm_dyn[];
    __shared__ double sm_red[EM_WARPS];
    __shared__ int sm_ctr[NSTAGE];     // work-queue tickets, one counter per pipeline stage
    __shared__ __align__(8) unsigned long long sm_full[NSTAGE], sm_empty[NSTAGE];   // TMA landed / consumers done
    constexpr bool direct = DIRECT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool producer = warp == EM_WARPS - 1;
    const int b = blockIdx.x;
    const int et0 = p.m.blk_etile0[b], n_et = p.m.blk_etile0[b + 1] - et0;
    const int mi0 = p.m.blk_mitem0[b], n_mi = p.m.blk_mitem0[b + 1] - mi0;
    const int ec0 = p.m.blk_ech0[b], n_ech = p.m.blk_ech0[b + 1] - ec0;
    const int mc0 = p.m.blk_mch0[b], n_mch = p.m.blk_mch0[b + 1] - mc0;
    BlockView v;
    v.row0 = p.m.blk_row0[b]; v.nrows = p.m.blk_row0[b + 1] - v.row0;
    v.cls0 = p.m.blk_cls0[b]; v.nres = p.m.blk_nres[b];
    v.nhr = p.m.blk_nhr[b]; v.nhc = p.m.blk_nhc[b];
    const int hr0 = p.m.blk_hr0[b], hc0 = p.m.blk_hc0[b];
    const SmemPlan pl = em_smem_plan(direct ? 0 : NSTAGE * CH_BYTES, n_et, n_mi, direct ? 0 : n_ech, direct ? 0 : n_mch, direct ? n_et + n_mi : 0, v.nrows, v.nhr, v.nres, v.nhc);
    int4 *s_et = (int4 *)(sm_dyn + pl.off_etiles);
    int4 *s_mi = (int4 *)(sm_dyn + pl.off_mitems);
    int4 *s_ech = (int4 *)(sm_dyn + pl.off_ech);
    int4 *s_mch = (int4 *)(sm_dyn + pl.off_mch);
    int32_t *s_hrl = (int32_t *)(sm_dyn + pl.off_hrl);
    int32_t *s_hcl = (int32_t *)(sm_dyn + pl.off_hcl);
    int32_t *s_eres = s_hcl + v.nhc, *s_mres = s_eres + n_et;      // direct mode: resident-cache offset of every tile / item (-1: none)
    double2 *s_rsa = (double2 *)(sm_dyn + pl.off_rsa);
    v.sm_theta = (double *)(sm_dyn + pl.off_theta);
    v.sm_q = (double *)(sm_dyn + pl.off_q);
    v.sm_etiles = s_et; v.sm_mitems = s_mi; v.sm_rsa = s_rsa;
    v.sm_stage = nullptr; v.sm_noff = nullptr;
    SmView f;
    f.base = sm_dyn; f.theta8 = pl.off_theta / 8; f.q8 = pl.off_q / 8; f.rsa16 = pl.off_rsa / 16;
    f.row0 = v.row0; f.cls0 = v.cls0; f.nres = v.nres; f.zslot = v.nrows + v.nhr; f.tag = 0; f.noff4 = 0;
    // every index of this CTA is a shared-memory slot (the host plan gave all its halo rows / classes a slot)
    const bool all_local = v.nhr == p.m.blk_hr0[b + 1] - hr0 && v.nhc == p.m.blk_hc0[b + 1] - hc0 && v.nres == p.m.blk_cls0[b + 1] - v.cls0;
    // per-CTA constants and the CTA's slice of theta -> shared memory, once
    for (int i = threadIdx.x; i < n_et; i += EM_BLOCK) s_et[i] = p.m.e_tiles[et0 + i];
    for (int i = threadIdx.x; i < n_mi; i += EM_BLOCK) s_mi[i] = p.m.m_items[mi0 + i];
    if (!direct) {
        for (int i = threadIdx.x; i < n_ech; i += EM_BLOCK) s_ech[i] = p.m.e_chunks[ec0 + i];
        for (int i = threadIdx.x; i < n_mch; i += EM_BLOCK) s_mch[i] = p.m.m_chunks[mc0 + i];
    } else {
        for (int i = threadIdx.x; i < n_et; i += EM_BLOCK) s_eres[i] = p.m.e_res[et0 + i];
        for (int i = threadIdx.x; i < n_mi; i += EM_BLOCK) s_mres[i] = p.m.m_res[mi0 + i];
    }
    for (int i = threadIdx.x; i < v.nhr; i += EM_BLOCK) s_hrl[i] = p.m.halo_rows[hr0 + i];
    for (int i = threadIdx.x; i < v.nhc; i += EM_BLOCK) s_hcl[i] = p.m.halo_cls[hc0 + i];
    for (int i = threadIdx.x; i < v.nrows; i += EM_BLOCK) { s_rsa[i] = p.m.row_RsA[v.row0 + i]; v.sm_theta[i] = p.m.theta[v.row0 + i]; }
    for (int i = threadIdx.x; i <= v.nres + v.nhc; i += EM_BLOCK) v.sm_q[i] = 0.0;
    if (threadIdx.x == 0) v.sm_theta[v.nrows + v.nhr] = 0.0;            // zero-theta slot: padding target of the E tiles
    if (threadIdx.x == 0)
        for (int sg = 0; sg < NSTAGE; sg++) { mbar_init(&sm_full[sg], 1); mbar_init(&sm_empty[sg], EM_WARPS - 1); }
    __syncthreads();
    if (direct) {
        // fill the resident index cache: [members | read counts] of the chosen E tiles, entries of the chosen M items
        int *res = (int *)(sm_dyn + pl.off_res);
        for (int i = warp; i < n_et; i += EM_WARPS) {
            const int o = s_eres[i];
            if (o < 0) continue;
            const int4 t = s_et[i];
            const int cpb = 32 >> ((t.w >> 12) & 0xf);
            const int ints = ((t.y + cpb - 1) / cpb) * 32 * (t.w & 0xfff);
            for (int j = lane; j < ints; j += 32) res[o + j] = p.m.e_tid[(uint32_t)t.z + j];
            for (int j = lane; j < t.y; j += 32) res[o + ints + j] = (int)p.m.e_R[t.x + j];
        }
        for (int i = warp; i < n_mi; i += EM_WARPS) {
            const int o = s_mres[i];
            if (o < 0) continue;
            const int4 t = s_mi[i];
            const int len = t.w & 0x3fffffff;
            const int ints = (t.w >> 30) == 0 ? 32 * len : len + t.y;
            for (int j = lane; j < ints; j += 32) res[o + j] = p.m.m_cls[(uint32_t)t.z + j];
        }
        __syncthreads();
    }
    // chunk sequence number: chunk g lives in stage g % NSTAGE, its barriers are in phase (g / NSTAGE) & 1
    int gseq = 0;            // consumers: next chunk to consume; producer: next chunk to issue
    int m_pre = 0;           // producer: M chunks of this iteration already issued before the grid barrier
    int it = 0;
    double d = INFINITY;

    // ---- producer: wait until the stage is free, reset its ticket counter, launch the TMA copies of one chunk ----
    auto issue_e = [&](int ci, int g) {
        const int sg = g % NSTAGE;
        mbar_wait(&sm_empty[sg], ((g / NSTAGE) & 1) ^ 1);
        const int4 c = s_ech[ci];
        if (lane == 0) sm_ctr[sg] = 0;
        if (c.w >= 0) {
            const int j0 = s_et[c.x].x, j1 = s_et[c.y - 1].x + s_et[c.y - 1].y;
            const int n_idx = c.w - (j1 - j0);
            const int32_t *gi = p.m.e_tid + (uint32_t)c.z;
            const uint32_t *gr = p.m.e_R + j0;
            const int ro = (stage_shift(gi) + n_idx + 3) & ~3;
            __syncwarp();
            if (lane == 0) mbar_expect_tx(&sm_full[sg], stage_bytes(gi, n_idx) + stage_bytes(gr, j1 - j0));
            __syncwarp();
            if (lane == 0) bulk_g2s(sm_dyn + sg * CH_BYTES, gi, stage_bytes(gi, n_idx), &sm_full[sg]);     // one bulk copy per stream piece: issuing costs ~55 ns each
            if (lane == 0) bulk_g2s((int *)(sm_dyn + sg * CH_BYTES) + ro, gr, stage_bytes(gr, j1 - j0), &sm_full[sg]);
        } else {
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm_full[sg]);       // oversized item: read straight from global
        }
    };
    auto issue_m = [&](int ci, int g) {
        const int sg = g % NSTAGE;
        mbar_wait(&sm_empty[sg], ((g / NSTAGE) & 1) ^ 1);
        const int4 c = s_mch[ci];
        if (lane == 0) sm_ctr[sg] = 0;
        __syncwarp();
        if (c.w >= 0) {
            const int32_t *gi = p.m.m_cls + (uint32_t)c.z;
            if (lane == 0) mbar_expect_tx(&sm_full[sg], stage_bytes(gi, c.w));
            __syncwarp();
            if (lane == 0) bulk_g2s(sm_dyn + sg * CH_BYTES, gi, stage_bytes(gi, c.w), &sm_full[sg]);
        } else if (lane == 0) mbar_arrive(&sm_full[sg]);
    };

    const int res4 = pl.off_res / 4;
    while (it < p.max_iter && MODE == 1) {
        // ---- direct mode: one work queue per phase, all 32 warps, indices straight from global memory (L2) ----
        TRACE(0);
        for (int i = threadIdx.x; i < v.nhr; i += EM_BLOCK) v.sm_theta[v.nrows + i] = __ldcg(p.m.theta + s_hrl[i]);
        if (threadIdx.x == 0) { sm_ctr[0] = 0; sm_ctr[1] = 0; }
        __syncthreads();
        if (all_local) {
            for (int tk = next_item(&sm_ctr[0], lane); tk < n_et; tk = next_item(&sm_ctr[0], lane)) {
                const int ti = n_et - 1 - tk;
                const int4 tile = s_et[ti];
                const int ro = s_eres[ti];
                if (ro >= 0) {
                    const int cpb = 32 >> ((tile.w >> 12) & 0xf), ints = ((tile.y + cpb - 1) / cpb) * 32 * (tile.w & 0xfff);
                    f_e_tile(p, f, tile, IdxS{sm_dyn}, IdxS{sm_dyn}, res4 + ro, res4 + ro + ints, lane);
                } else f_e_tile(p, f, tile, IdxG{p.m.e_tid}, IdxG{(const int32_t *)p.m.e_R}, tile.z, tile.x, lane);
            }
        } else {
            for (int tk = next_item(&sm_ctr[0], lane); tk < n_et; tk = next_item(&sm_ctr[0], lane)) {
                const int4 tile = s_et[n_et - 1 - tk];
                e_tile<false>(p, v, tile, p.m.e_tid + (uint32_t)tile.z, p.m.e_R + tile.x, lane);
            }
        }
        TRACE(1);
        grid_barrier(p.bar, gridDim.x, 2 * it + 1);
        TRACE(2);
        for (int i = threadIdx.x; i < v.nhc; i += EM_BLOCK) v.sm_q[v.nres + 1 + i] = __ldcg(p.m.q + s_hcl[i]);
        __syncthreads();
        double dm = 0;
        if (all_local) {
            for (int tk = next_item(&sm_ctr[1], lane); tk < n_mi; tk = next_item(&sm_ctr[1], lane)) {
                const int4 itm = s_mi[tk];
                const int ro = s_mres[tk];
                if (ro >= 0) dm = fmax(dm, f_m_item<0>(p, f, itm, IdxS{sm_dyn}, res4 + ro, lane));
                else dm = fmax(dm, f_m_item<0>(p, f, itm, IdxG{p.m.m_cls}, itm.z, lane));
            }
        } else {
            for (int tk = next_item(&sm_ctr[1], lane); tk < n_mi; tk = next_item(&sm_ctr[1], lane)) {
                const int4 itm = s_mi[tk];
                dm = fmax(dm, m_item<false>(p, v, itm, p.m.m_cls + (uint32_t)itm.z, lane));
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dm = fmax(dm, __shfl_xor_sync(0xffffffffu, dm, o));
        if (lane == 0) sm_red[warp] = dm;
        __syncthreads();
        if (threadIdx.x == 0) {
            double bm = 0;
            for (int w = 0; w < EM_WARPS; w++) bm = fmax(bm, sm_red[w]);
            atomicMax(p.dmax + (it & 1), (unsigned long long)__double_as_longlong(bm));
        }
        TRACE(3);
        grid_barrier(p.bar, gridDim.x, 2 * it + 2);
        TRACE(4);
        d = __longlong_as_double((long long)*((volatile unsigned long long *)(p.dmax + (it & 1))));
        if (blockIdx.x == 0 && threadIdx.x == 0) p.dmax[(it + 1) & 1] = 0ULL;
        it++;
        if (p.stop_on_conv && d <= 1.0) break;
    }
    if (MODE == 3) {
        // ---- barrier-free iteration (needs every halo row / class of every CTA in a shared-memory slot) --------------------
        // A CTA depends only on the CTAs it shares classes with, so nothing here waits for the whole grid: the owner of a row /
        // class publishes theta / q in a tagged 16-byte slot (ll_store) and the readers poll exactly the slots they need
        // (ll_load) at the start of a phase. A slot is rewritten one iteration later; its readers are done by then because the
        // dependency is symmetric: the owner b of class c needs theta of every member row of c before it can recompute q_c, and
        // the owner n of such a row publishes that theta only after it has read q_c (and the same the other way round).
        // The convergence measure of iteration i is read after the E-phase of iteration i+1 (theta has not changed by then, so
        // stopping there leaves exactly the state of iteration i); that read is the only all-to-all dependency and bounds the
        // drift between CTAs to one iteration, which is what lets the delta slots alternate by parity.
        auto read_dm = [&](int j) -> double {
            const unsigned tg = p.df_tag0 + (unsigned)j + 1u;
            double x = 0;
            for (int i = threadIdx.x; i < (int)gridDim.x; i += EM_BLOCK)
                x = fmax(x, ll_load(p.dm_slots + 16 * (size_t)((j & 1) * (int)gridDim.x + i), tg, p.df_abort));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, o));
            __syncthreads();
            if (lane == 0) sm_red[warp] = x;
            __syncthreads();
            x = 0;
            for (int w = 0; w < EM_WARPS; w++) x = fmax(x, sm_red[w]);
            __syncthreads();
            return x;
        };
        bool stopped = false;
        while (it < p.max_iter) {
            const unsigned tag = p.df_tag0 + (unsigned)it + 1u;
            f.tag = tag;
            TRACE(0);
            for (int i = threadIdx.x; i < v.nhr; i += EM_BLOCK) v.sm_theta[v.nrows + i] = ll_load(p.th_slots + 16 * (size_t)s_hrl[i], tag, p.df_abort);
            if (threadIdx.x == 0) { sm_ctr[0] = 0; sm_ctr[1] = 0; }
            __syncthreads();
            for (int tk = next_item(&sm_ctr[0], lane); tk < n_et; tk = next_item(&sm_ctr[0], lane)) {
                const int ti = n_et - 1 - tk;
                const int4 tile = s_et[ti];
                const int ro = s_eres[ti];
                if (ro >= 0) {
                    const int cpb = 32 >> ((tile.w >> 12) & 0xf), ints = ((tile.y + cpb - 1) / cpb) * 32 * (tile.w & 0xfff);
                    f_e_tile<1>(p, f, tile, IdxS{sm_dyn}, IdxS{sm_dyn}, res4 + ro, res4 + ro + ints, lane);
                } else f_e_tile<1>(p, f, tile, IdxG{p.m.e_tid}, IdxG{(const int32_t *)p.m.e_R}, tile.z, tile.x, lane);
            }
            TRACE(1);
            if (it > 0) {
                d = read_dm(it - 1);
                if (*((volatile int *)p.df_abort) != 0 || (p.stop_on_conv && d <= 1.0)) { stopped = true; break; }
            }
            for (int i = threadIdx.x; i < v.nhc; i += EM_BLOCK) v.sm_q[v.nres + 1 + i] = ll_load(p.q_slots + 16 * (size_t)s_hcl[i], tag, p.df_abort);
            __syncthreads();                       // the CTA's E-phase is complete and the halo q are in place
            TRACE(2);
            double dm = 0;
            for (int tk = next_item(&sm_ctr[1], lane); tk < n_mi; tk = next_item(&sm_ctr[1], lane)) {
                const int4 itm = s_mi[tk];
                const int ro = s_mres[tk];
                if (ro >= 0) dm = fmax(dm, f_m_item<2>(p, f, itm, IdxS{sm_dyn}, res4 + ro, lane));
                else dm = fmax(dm, f_m_item<2>(p, f, itm, IdxG{p.m.m_cls}, itm.z, lane));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) dm = fmax(dm, __shfl_xor_sync(0xffffffffu, dm, o));
            if (lane == 0) sm_red[warp] = dm;
            __syncthreads();
            if (threadIdx.x == 0) {
                double bm = 0;
                for (int w = 0; w < EM_WARPS; w++) bm = fmax(bm, sm_red[w]);
                ll_store(p.dm_slots + 16 * (size_t)((it & 1) * (int)gridDim.x + b), bm, tag);
            }
            TRACE(3);
            TRACE(4);
            it++;
        }
        if (it > 0 && !stopped) d = read_dm(it - 1);
        if (*((volatile int *)p.df_abort) != 0) d = INFINITY;
        __syncthreads();
        for (int i = threadIdx.x; i < v.nrows; i += EM_BLOCK) p.m.theta[v.row0 + i] = v.sm_theta[i];      // the copy the output kernels read
    }
    if (SHARDED) {
        // ---- class-sharded sample: this rank's classes only. Per iteration: [theta of my rows + halo rows from my window] ->
        // E-phase -> grid barrier -> partial row sums, pushed (coalesced, natural order, tagged slots) into the xbuf of the rank
        // that owns the row -> the owner adds the nranks partials in rank order as they arrive, updates theta of its slice and
        // pushes it into EVERY rank's theta window -> every owner CTA pushes its convergence measure to every rank; all CTAs
        // read all of them (the only all-to-all wait of the iteration). theta is bit-identical on all ranks by construction.
        // Hazards: a slot is rewritten one iteration later; by then its readers are done because (a) a CTA pushes partials of
        // iteration i+1 only after the wait at the end of iteration i, which every owner CTA joins after reading its xbuf
        // slots, and (b) an owner rewrites theta only after partials of iteration i+1 arrived from every rank, i.e. after that
        // rank's E -> M grid barrier, which every reader of theta(i) has passed. The dm slots alternate by parity.
        const EmParams::Shard &sh = p.sh;
        const int R = sh.nranks, S = sh.S, me = sh.rank;
        const unsigned char *my_xbuf = sh.win[me] + sh.xbuf_off;
        // theta lives in two slot arrays used in turn (iteration parity): without the E -> M grid barrier (df below) a reader of
        // theta(i) may still be at it while the owner already publishes theta(i+1)
        auto theta_slot = [&](int r, int iter, int n) -> unsigned char * { return sh.win[r] + sh.theta_off + 16 * ((size_t)(iter & 1) * sh.theta_cap + (size_t)n); };
        // df: no grid barrier between the phases either - q of the classes other CTAs read goes through local tagged slots as in
        // k_em_persistent<3> (needs every halo row / class of every CTA of this rank in shared memory)
        const bool df = sh.fused && p.m.all_local && p.q_slots != nullptr;
        const int n0 = v.row0;                                                   // the CTA's rows are a contiguous natural range
        int *s_noff = (int *)(s_rsa + 0) + 2 * v.nrows;                          // second half of the {Rs,A} area
        double *s_stage = (double *)s_rsa;
        v.sm_stage = s_stage; v.sm_noff = s_noff;
       
        f.noff4 = pl.off_rsa / 4 + 2 * v.nrows;
        for (int i = threadIdx.x; i < v.nrows; i += EM_BLOCK) s_noff[i] = p.m.row_n[v.row0 + i] - n0;
        if (sh.fused)
            for (int i = threadIdx.x; i < v.nhr; i += EM_BLOCK) s_hrl[i] = p.m.row_n[s_hrl[i]];    // halo rows: natural index
        __syncthreads();
        // the slice this rank owns, cut over the CTAs
        const int own0 = me * S, own1 = min(p.m.P, own0 + S);
        const int per = (max(own1 - own0, 0) + (int)gridDim.x - 1) / (int)gridDim.x;
        const int u0 = min(own1, own0 + b * per), u1 = min(own1, u0 + per);
        // the convergence measure of iteration j: the maximum over every owner CTA of every rank (slots alternate by parity)
        auto read_dm = [&](int j) -> double {
            const unsigned tg = sh.tag0 + (unsigned)j + 1u;
            double x = 0;
            for (int i = threadIdx.x; i < R * (int)gridDim.x; i += EM_BLOCK) {
                const int r = i / (int)gridDim.x, c = i - r * (int)gridDim.x;
                x = fmax(x, ll_load(sh.win[me] + 16 * ((size_t)((j & 1) * R + r) * WIN_MAX_CTAS + c), tg, sh.abort_flag));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, o));
            __syncthreads();
            if (lane == 0) sm_red[warp] = x;
            __syncthreads();
            x = 0;
            for (int w = 0; w < EM_WARPS; w++) x = fmax(x, sm_red[w]);
            __syncthreads();
            return x;
        };
        bool stopped = false;
        while (it < p.max_iter) {
            const unsigned tag = sh.tag0 + (unsigned)it + 1u;             // theta(it) and partials / dm of iteration it carry it
            f.tag = p.df_tag0 + (unsigned)it + 1u;                        // tag of the local q slots (df)
            TRACE(0);
            if (sh.fused) {
                for (int i = threadIdx.x; i < v.nrows; i += EM_BLOCK) v.sm_theta[i] = ll_load(theta_slot(me, it, n0 + s_noff[i]), tag, sh.abort_flag);
                for (int i = threadIdx.x; i < v.nhr; i += EM_BLOCK) v.sm_theta[v.nrows + i] = ll_load(theta_slot(me, it, s_hrl[i]), tag, sh.abort_flag);
                if (!p.m.all_local)    // overflow path (some CTA): halo rows without a slot are read from the permuted global copy during the E-phase
                    for (int i = threadIdx.x; i < v.nrows; i += EM_BLOCK) p.m.theta[v.row0 + i] = v.sm_theta[i];
            } else {
                for (int i = threadIdx.x; i < v.nhr; i += EM_BLOCK) v.sm_theta[v.nrows + i] = __ldcg(p.m.theta + s_hrl[i]);
            }
            if (threadIdx.x == 0) { sm_ctr[0] = 0; sm_ctr[1] = 0; }
            if (sh.fused && !p.m.all_local) grid_barrier(p.bar, gridDim.x, 2 * it + 1);  // the global copy is complete before anybody gathers from it
            else __syncthreads();
            if (df) {
                for (int tk = next_item(&sm_ctr[0], lane); tk < n_et; tk = next_item(&sm_ctr[0], lane)) {
                    const int ti = n_et - 1 - tk;
                    const int4 tile = s_et[ti];
                    const int ro = s_eres[ti];
                    if (ro >= 0) {
                        const int cpb = 32 >> ((tile.w >> 12) & 0xf), ints = ((tile.y + cpb - 1) / cpb) * 32 * (tile.w & 0xfff);
                        f_e_tile<1>(p, f, tile, IdxS{sm_dyn}, IdxS{sm_dyn}, res4 + ro, res4 + ro + ints, lane);
                    } else f_e_tile<1>(p, f, tile, IdxG{p.m.e_tid}, IdxG{(const int32_t *)p.m.e_R}, tile.z, tile.x, lane);
                }
            } else if (all_local) {
                for (int tk = next_item(&sm_ctr[0], lane); tk < n_et; tk = next_item(&sm_ctr[0], lane)) {
                    const int ti = n_et - 1 - tk;
                    const int4 tile = s_et[ti];
                    const int ro = s_eres[ti];
                    if (ro >= 0) {
                        const int cpb = 32 >> ((tile.w >> 12) & 0xf), ints = ((tile.y + cpb - 1) / cpb) * 32 * (tile.w & 0xfff);
                        f_e_tile(p, f, tile, IdxS{sm_dyn}, IdxS{sm_dyn}, res4 + ro, res4 + ro + ints, lane);
                    } else f_e_tile(p, f, tile, IdxG{p.m.e_tid}, IdxG{(const int32_t *)p.m.e_R}, tile.z, tile.x, lane);
                }
            } else {
                for (int tk = next_item(&sm_ctr[0], lane); tk < n_et; tk = next_item(&sm_ctr[0], lane)) {
                    const int4 tile = s_et[n_et - 1 - tk];
                    e_tile<true>(p, v, tile, p.m.e_tid + (uint32_t)tile.z, p.m.e_R + tile.x, lane);
                }
            }
            TRACE(1);
            if (sh.fused && it > 0) {
                // The delta of the PREVIOUS iteration is read here, not at its end: by now the values have arrived, so the only
                // all-to-all dependency of an iteration costs nothing. theta has not changed since (this E-phase only wrote q),
                // so stopping here leaves exactly the state of the iteration that met the rule. Every CTA of every rank takes
                // the same decision from the same numbers.
                d = read_dm(it - 1);
                if (*((volatile int *)sh.abort_flag) != 0) {
                    // leaving early: park this CTA's barrier flag at the last generation so that no CTA still running waits for it
                    if (threadIdx.x == 0) asm volatile("st.release.gpu.u32 [%0], %1;" ::"l"(p.bar + blockIdx.x * 32), "r"(0x7fffffffu) : "memory");
                    stopped = true;
                    break;
                }
                if (p.stop_on_conv && d <= 1.0) { stopped = true; break; }
            }
            if (df) {
                for (int i = threadIdx.x; i < v.nhc; i += EM_BLOCK) v.sm_q[v.nres + 1 + i] = ll_load(p.q_slots + 16 * (size_t)s_hcl[i], f.tag, sh.abort_flag);
            } else {
                grid_barrier(p.bar, gridDim.x, 2 * it + 2);
                for (int i = threadIdx.x; i < v.nhc; i += EM_BLOCK) v.sm_q[v.nres + 1 + i] = __ldcg(p.m.q + s_hcl[i]);
            }
            TRACE(2);
            __syncthreads();
            if (all_local) {
                for (int tk = next_item(&sm_ctr[1], lane); tk < n_mi; tk = next_item(&sm_ctr[1], lane)) {
                    const int4 itm = s_mi[tk];
                    const int ro = s_mres[tk];
                    if (ro >= 0) f_m_item<1>(p, f, itm, IdxS{sm_dyn}, res4 + ro, lane);
                    else f_m_item<1>(p, f, itm, IdxG{p.m.m_cls}, itm.z, lane);
                }
            } else {
                for (int tk = next_item(&sm_ctr[1], lane); tk < n_mi; tk = next_item(&sm_ctr[1], lane)) {
                    const int4 itm = s_mi[tk];
                    m_item<true>(p, v, itm, p.m.m_cls + (uint32_t)itm.z, lane);
                }
            }
            __syncthreads();
            if (!sh.fused) {
                // NCCL path: one pass; the host all-reduces q_out and runs the update kernel
                for (int i = threadIdx.x; i < v.nrows; i += EM_BLOCK) sh.q_out[n0 + i] = s_stage[i];
                it++;
                break;
            }
            // push: natural row n goes to rank n / S, slot me * S + (n - owner * S)
            for (int i = threadIdx.x; i < v.nrows; i += EM_BLOCK) {
                const int n = n0 + i, o = n / S;
                ll_store(sh.win[o] + sh.xbuf_off + 16 * ((size_t)me * S + (n - o * S)), s_stage[i], tag);
            }
            TRACE(3);
            // owner update of this CTA's share of the slice, as the partials arrive
            double dm = 0;
            for (int n = u0 + threadIdx.x; n < u1; n += EM_BLOCK) {
                double Q = 0;
                for (int r = 0; r < R; r++) Q += ll_load(my_xbuf + 16 * ((size_t)r * S + (n - own0)), tag, sh.abort_flag);   // rank order: deterministic
                const double2 ra = p.m.rsa_nat[n];
                const double th = ll_load(theta_slot(me, it, n), tag, sh.abort_flag);
                const double nn = ra.x + th * Q;
                const double thn = fast_div(nn, ra.y);
                dm = fmax(dm, fast_div(fabs(thn - th) * ra.y, p.eps_abs + p.eps_rel * nn));
                for (int r = 0; r < R; r++) ll_store(theta_slot(r, it + 1, n), thn, tag + 1);
            }
            TRACE(4);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) dm = fmax(dm, __shfl_xor_sync(0xffffffffu, dm, o));
            if (lane == 0) sm_red[warp] = dm;
            __syncthreads();
            if ((int)threadIdx.x < R) {
                double bm = 0;
                for (int w = 0; w < EM_WARPS; w++) bm = fmax(bm, sm_red[w]);
                ll_store(sh.win[threadIdx.x] + 16 * ((size_t)((it & 1) * R + me) * WIN_MAX_CTAS + b), bm, tag);
            }
            TRACE(5);
            it++;
        }
        if (sh.fused && it > 0 && !stopped) d = read_dm(it - 1);         // ran out of iterations: the last delta is still in flight
        if (sh.fused && *((volatile int *)sh.abort_flag) != 0) d = INFINITY;
        if (sh.fused) {      // the permuted copy the output kernels read
            const unsigned tag = sh.tag0 + (unsigned)it + 1u;
            for (int i = threadIdx.x; i < v.nrows; i += EM_BLOCK) p.m.theta[v.row0 + i] = ll_load(theta_slot(me, it, n0 + s_noff[i]), tag, sh.abort_flag);
        }
    }
    while (it < p.max_iter && MODE == 0) {
        TRACE(0);
        double dm = 0;
        if (producer) {
            // ================= producer warp =================
            for (int ci = 0; ci < n_ech; ci++) issue_e(ci, gseq + ci);
            gseq += n_ech;
            // the M stream does not depend on the E results: prefetch its first chunks before the grid barrier. Only
            // chunks whose stage was last used by an E chunk may be issued here (the consumers release those without
            // waiting for anybody), i.e. at most NSTAGE of them.
            m_pre = min(NSTAGE, n_mch);
            for (int ci = 0; ci < m_pre; ci++) issue_m(ci, gseq + ci);
        } else {
            // ================= consumer warps: E-phase =================
            for (int i = threadIdx.x; i < v.nhr; i += EM_BLOCK - 32) v.sm_theta[v.nrows + i] = __ldcg(p.m.theta + s_hrl[i]);
            if (v.nhr > 0) consumer_sync();             // halo theta visible to every consumer
            for (int ci = 0; ci < n_ech; ci++) {
                const int g = gseq + ci, sg = g % NSTAGE;
                const int4 c = s_ech[ci];
                mbar_wait(&sm_full[sg], (g / NSTAGE) & 1);
                const int n_items = c.y - c.x;
                const bool staged = c.w >= 0;
                const int jbase = s_et[c.x].x;
                const int32_t *gi = p.m.e_tid + (uint32_t)c.z;
                const int sh = stage_shift(gi);
                const int n_idx = c.w - (s_et[c.y - 1].x + s_et[c.y - 1].y - jbase);
                const int ro = (sh + n_idx + 3) & ~3, sr = stage_shift(p.m.e_R + jbase);
                // tiles are ordered by cardinality: take them from the heaviest end
                if (staged && all_local) {
                    const int tbase = sg * (CH_BYTES / 4) + sh - c.z, rbase = sg * (CH_BYTES / 4) + ro + sr - jbase;
                    for (int tk = next_item(&sm_ctr[sg], lane); tk < n_items; tk = next_item(&sm_ctr[sg], lane)) {
                        const int4 tile = s_et[c.y - 1 - tk];
                        f_e_tile(p, f, tile, IdxS{sm_dyn}, IdxS{sm_dyn}, tbase + tile.z, rbase + tile.x, lane);
                    }
                } else {
                    int *buf = (int *)(sm_dyn + sg * CH_BYTES);
                    for (int tk = next_item(&sm_ctr[sg], lane); tk < n_items; tk = next_item(&sm_ctr[sg], lane)) {
                        const int4 tile = s_et[c.y - 1 - tk];
                        const int *tids = staged ? buf + sh + (tile.z - c.z) : p.m.e_tid + (uint32_t)tile.z;
                        const uint32_t *rfl = staged ? (const uint32_t *)(buf + ro + sr + (tile.x - jbase)) : p.m.e_R + tile.x;
                        e_tile<false>(p, v, tile, tids, rfl, lane);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm_empty[sg]);   // this warp is done with the stage
            }
            gseq += n_ech;
        }
        TRACE(1);
        grid_barrier(p.bar, gridDim.x, 2 * it + 1);
        TRACE(2);
        if (producer) {
            for (int ci = m_pre; ci < n_mch; ci++) issue_m(ci, gseq + ci);
            gseq += n_mch;
        } else {
            // ================= consumer warps: M-phase =================
            for (int i = threadIdx.x; i < v.nhc; i += EM_BLOCK - 32) v.sm_q[v.nres + 1 + i] = __ldcg(p.m.q + s_hcl[i]);
            if (v.nhc > 0) consumer_sync();
            for (int ci = 0; ci < n_mch; ci++) {
                const int g = gseq + ci, sg = g % NSTAGE;
                const int4 c = s_mch[ci];
                mbar_wait(&sm_full[sg], (g / NSTAGE) & 1);
                const int n_items = c.y - c.x;
                const bool staged = c.w >= 0;
                const int sh = stage_shift(p.m.m_cls + (uint32_t)c.z);
                // items are ordered longest first
                if (staged && all_local) {
                    const int ebase = sg * (CH_BYTES / 4) + sh - c.z;
                    for (int tk = next_item(&sm_ctr[sg], lane); tk < n_items; tk = next_item(&sm_ctr[sg], lane)) {
                        const int4 itm = s_mi[c.x + tk];
                        dm = fmax(dm, f_m_item<0>(p, f, itm, IdxS{sm_dyn}, ebase + itm.z, lane));
                    }
                } else {
                    int *buf = (int *)(sm_dyn + sg * CH_BYTES);
                    for (int tk = next_item(&sm_ctr[sg], lane); tk < n_items; tk = next_item(&sm_ctr[sg], lane)) {
                        const int4 itm = s_mi[c.x + tk];
                        const int *ent = staged ? buf + sh + (itm.z - c.z) : p.m.m_cls + (uint32_t)itm.z;
                        dm = fmax(dm, m_item<false>(p, v, itm, ent, lane));
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm_empty[sg]);
            }
            gseq += n_mch;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dm = fmax(dm, __shfl_xor_sync(0xffffffffu, dm, o));
        if (lane == 0) sm_red[warp] = dm;
        __syncthreads();
        if (threadIdx.x == 0) {
            double bm = 0;
            for (int w = 0; w < EM_WARPS; w++) bm = fmax(bm, sm_red[w]);
            atomicMax(p.dmax + (it & 1), (unsigned long long)__double_as_longlong(bm));
        }
        TRACE(3);
        grid_barrier(p.bar, gridDim.x, 2 * it + 2);
        TRACE(4);
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
    CU(cudaFuncSetAttribute(k_em_persistent<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CU(cudaFuncSetAttribute(k_em_persistent<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CU(cudaFuncSetAttribute(k_em_persistent<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CU(cudaFuncSetAttribute(k_em_persistent<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int nb = 0, nb2 = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_em_persistent<1>, EM_BLOCK, smem));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb2, k_em_persistent<2>, EM_BLOCK, smem));
    if (nb2 < nb) nb = nb2;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb2, k_em_persistent<3>, EM_BLOCK, smem));
    if (nb2 < nb) nb = nb2;
    if (nb < 1) { emsar_set_err("EM kernel does not fit on an SM (%d bytes of shared memory)", smem); return EMSAR_ERR_CUDA; }
    ctx->em_blocks_per_sm = 1;
    ctx->em_smem_bytes = smem;
    TRY(em_psum_attr(ctx));
    return EMSAR_OK;
}

__global__ void k_theta_to_slots(int32_t P, const double *__restrict__ theta, unsigned char *__restrict__ th_slots, unsigned tag)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < P) ll_store(th_slots + 16 * (size_t)p, theta[p], tag);
}

int em_launch(emsar_sample *s, int max_iter, int stop_on_conv, int *iters_done, double *final_delta, double *ms_out, bool fused)
{
    emsar_ctx *ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    EmParams p;
    p.m = s->m;
    p.eps_abs = s->opts.eps_abs; p.eps_rel = s->opts.eps_rel;
    p.max_iter = max_iter; p.stop_on_conv = stop_on_conv;
    p.bar = ctx->d_barrier + 64;                                   // one 128-byte line per CTA
    p.dmax = (unsigned long long *)(ctx->d_barrier + 4);
    p.iters_done = (int *)(ctx->d_barrier + 8);
    p.final_delta = (double *)(ctx->d_barrier + 10);
    p.trace = s->d_trace;
    memset(&p.sh, 0, sizeof(p.sh));
    if (s->sharded) {
        if (!s->m.direct) { emsar_set_err("sharded samples need the direct EM mode"); return EMSAR_ERR_UNSUPPORTED; }
        p.sh.rank = ctx->rank; p.sh.nranks = ctx->nranks;
        p.sh.S = (int)win_slice_rows(s->m.P > 0 ? s->m.P : 1, ctx->nranks);
        p.sh.fused = fused ? 1 : 0;
        p.sh.q_out = s->d_qpart;
        p.sh.abort_flag = (int *)(ctx->d_barrier + 14);
        p.sh.tag0 = 0;
        if (fused) {
            if (s->m.B > WIN_MAX_CTAS) { emsar_set_err("sharded EM: more than %d CTAs", WIN_MAX_CTAS); return EMSAR_ERR_UNSUPPORTED; }
            for (int r = 0; r < ctx->nranks; r++) p.sh.win[r] = (unsigned char *)ctx->peer_win[r];
            p.sh.theta_off = (long long)WIN_HDR_BYTES;
            p.sh.theta_cap = (long long)ctx->win_rows;
            p.sh.xbuf_off = (long long)(WIN_HDR_BYTES + 32 * (size_t)ctx->win_rows);
        }
    }
    // scalars, barrier flags and the delta slots of the barrier-free kernel (behind the two flag arrays): the slots must be clean at every
    // launch because their tags restart with every sample (s->slot_tag), and a launch that ended after 1-2 iterations leaves tags 1 and 2 behind
    CU(cudaMemsetAsync(ctx->d_barrier, 0, 256 + 2 * (size_t)ctx->prop.multiProcessorCount * 128 + 2 * (size_t)s->m.B * 16, st));
    // barrier-free variant: whenever every CTA holds its whole halo in shared memory (EMSAR_EM_MODE=barrier keeps the grid barriers)
    const char *em_mode = getenv("EMSAR_EM_MODE");
    const bool dataflow = !s->sharded && s->m.direct && s->m.all_local && s->d_slots && !(em_mode && !strcmp(em_mode, "barrier"));
    p.th_slots = p.q_slots = p.dm_slots = nullptr; p.df_tag0 = 0; p.df_abort = (int *)(ctx->d_barrier + 14);
    if (s->sharded && fused && s->m.all_local && s->d_slots && !(em_mode && !strcmp(em_mode, "barrier"))) {
        // the sharded kernel's local q exchange: same tagged slots, a fresh tag range per launch
        p.q_slots = (unsigned char *)s->d_slots + 16 * (size_t)(s->m.P + 1);
        if ((unsigned)(s->slot_tag + (unsigned)max_iter + 4u) < s->slot_tag) { CU(cudaMemsetAsync(s->d_slots, 0, s->slots_bytes, st)); s->slot_tag = 0; }
        p.df_tag0 = s->slot_tag;
        s->slot_tag += (unsigned)max_iter + 2u;
    }
    if (dataflow) {
        p.th_slots = (unsigned char *)s->d_slots;
        p.q_slots = p.th_slots + 16 * (size_t)(s->m.P + 1);
        p.dm_slots = (unsigned char *)ctx->d_barrier + 256 + 2 * (size_t)ctx->prop.multiProcessorCount * 128;     // behind the two flag arrays
        if ((unsigned)(s->slot_tag + (unsigned)max_iter + 4u) < s->slot_tag) {         // tag wrap: start over from clean slots
            CU(cudaMemsetAsync(s->d_slots, 0, s->slots_bytes, st));
            s->slot_tag = 0;
        }
        p.df_tag0 = s->slot_tag;
        s->slot_tag += (unsigned)max_iter + 2u;
        if (s->m.P > 0) { k_theta_to_slots<<<(s->m.P + 255) / 256, 256, 0, st>>>(s->m.P, s->m.theta, p.th_slots, p.df_tag0 + 1u); LAUNCHED(ctx); }
    }
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
        // the global copies of theta | q (halo traffic; in the barrier-free kernel their tagged slots) stay resident in L2 while
        // the index streams through
        size_t win = dataflow ? s->slots_bytes : s->state_bytes;
        if (win > (size_t)ctx->prop.accessPolicyMaxWindowSize) win = (size_t)ctx->prop.accessPolicyMaxWindowSize;
        attrs[na].id = cudaLaunchAttributeAccessPolicyWindow;
        attrs[na].val.accessPolicyWindow.base_ptr = dataflow ? s->d_slots : (void *)s->d_state;
        attrs[na].val.accessPolicyWindow.num_bytes = win;
        attrs[na].val.accessPolicyWindow.hitRatio = win <= ctx->l2_persist_bytes ? 1.0f : (float)ctx->l2_persist_bytes / (float)win;
        attrs[na].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attrs[na].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        na++;
    }
    cfg.attrs = attrs;
    cfg.numAttrs = na;
    CU(cudaEventRecord(ctx->ev0, st));
    // the limit is a per-device attribute of the function, not of this context: another context of the process (another GPU, or a
    // differently sized test context) may have changed it since
    const void *fn = s->sharded ? (const void *)k_em_persistent<2> : dataflow ? (const void *)k_em_persistent<3>
                     : s->m.direct ? (const void *)k_em_persistent<1> : (const void *)k_em_persistent<0>;
    CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->em_smem_bytes));
    if (s->sharded) CU(cudaLaunchKernelEx(&cfg, k_em_persistent<2>, p));
    else if (dataflow) CU(cudaLaunchKernelEx(&cfg, k_em_persistent<3>, p));
    else if (s->m.direct) CU(cudaLaunchKernelEx(&cfg, k_em_persistent<1>, p));
    else CU(cudaLaunchKernelEx(&cfg, k_em_persistent<0>, p));
    LAUNCHED(ctx);
    CU(cudaEventRecord(ctx->ev1, st));
    int it = 0; double fd = 0;
    int aborted = 0;
    CU(cudaMemcpyAsync(&it, p.iters_done, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&fd, p.final_delta, 8, cudaMemcpyDeviceToHost, st));
    if (s->sharded || dataflow) CU(cudaMemcpyAsync(&aborted, p.df_abort, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (aborted && dataflow) { emsar_set_err("EM kernel: a wait on a tagged slot timed out (internal error)"); return EMSAR_ERR_STATE; }
    if (aborted) { emsar_set_err("sharded EM: a wait on peer memory timed out (a rank died or the ranks disagree on the call sequence)"); return EMSAR_ERR_COMM; }
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

// tuning aid (not part of include/emsar_cuda.h): run `iters` iterations and return, per CTA, the globaltimer stamps (ns) of
// the last iteration: [0] E start, [1] E end, [2] after barrier 1, [3] M end, [4] after barrier 2
extern "C" int emsar_sample_em_run(emsar_sample *s, int32_t max_iter, int32_t stop_on_conv, int32_t reset_theta,
                                   int32_t *iters_done, double *final_delta, double *elapsed_ms);
extern "C" int emsar_debug_em_trace(emsar_sample *s, int iters, unsigned long long *out, int *n_blocks)
{
    if (!s || !s->prepared) return EMSAR_ERR_STATE;
    TRY(ctx_use(s->ctx));
    const int B = s->use_psum ? s->ps.B : s->m.B;
    TRY(dev_alloc(&s->d_trace, (size_t)B * 8 + 64 + 1600));
    CU(cudaMemsetAsync(s->d_trace, 0, (size_t)B * 64 + 512 + 12800, s->ctx->stream));
    int it = 0; double fd = 0, ms = 0;
    int rc = emsar_sample_em_run(s, iters, 0, 0, &it, &fd, &ms);       // the sharded runner when the sample is sharded
    if (rc == EMSAR_OK) {
        CU(cudaMemcpy(out, s->d_trace, (size_t)B * 64 + 512 + 12800, cudaMemcpyDeviceToHost));
        *n_blocks = B;
    }
    dev_free(s->d_trace);
    s->d_trace = nullptr;
    return rc;
}

// ---- sharded mode, NCCL path: theta update from the all-reduced per-row sums (identical on every rank) ----
__global__ void k_update_sharded(int32_t P, const int32_t *__restrict__ row_n, const double2 *__restrict__ row_RsA, const double *__restrict__ qsum,
                                 double *__restrict__ theta, double eps_abs, double eps_rel, unsigned long long *dmax)
{
    __shared__ double red[8];
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    double d = 0;
    if (p < P) {
        const double2 ra = row_RsA[p];
        const double th = theta[p];
        const double n = ra.x + th * qsum[row_n[p]];
        const double thn = fast_div(n, ra.y);
        theta[p] = thn;
        d = fast_div(fabs(thn - th) * ra.y, eps_abs + eps_rel * n);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d = fmax(d, __shfl_xor_sync(0xffffffffu, d, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d;
    __syncthreads();
    if (threadIdx.x == 0) {
        double b = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) b = fmax(b, red[w]);
        atomicMax(dmax, (unsigned long long)__double_as_longlong(b));
    }
}

// fused path: theta in natural order into this rank's window before the kernel starts
__global__ void k_theta_to_nat(int32_t P, const int32_t *__restrict__ row_n, const double *__restrict__ theta, unsigned char *__restrict__ th_slots, unsigned tag)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < P) ll_store(th_slots + 16 * (size_t)row_n[p], theta[p], tag);
}

// One sample whose active classes are range-sharded over the ranks (BASELINE.json configs[2]).
//  * fused path (default): ONE persistent kernel per rank; the per-iteration all-reduce of the per-row sums runs inside it
//    over NVLink peer memory (reduce-scatter by pushes, owner update, all-gather by pushes; two cross-GPU barriers).
//  * NCCL path (EMSAR_SHARD_MODE=nccl, or when peer memory cannot be mapped): one kernel pass, ncclAllReduce of the fp64
//    sums, update kernel, host check of the convergence measure - per iteration.
static int em_run_sharded(emsar_sample *s, int max_iter, int stop_on_conv, int *iters_done, double *final_delta, double *ms_out)
{
    emsar_ctx *ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    const int32_t P = s->m.P;
    if (!s->m.direct) { emsar_set_err("sharded samples need the direct EM mode"); return EMSAR_ERR_UNSUPPORTED; }
    TRY(comm_window_ensure(ctx, P));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    int it = 0; double d = INFINITY;
    if (ctx->win_state == 1) {
        CU(cudaMemsetAsync(ctx->win, 0, ctx->win_bytes, st));          // every tag back to 0: the tags of a launch start at 1
        if (P > 0) { k_theta_to_nat<<<(P + 255) / 256, 256, 0, st>>>(P, s->m.row_n, s->m.theta, (unsigned char *)ctx->win + WIN_HDR_BYTES, 1u); LAUNCHED(ctx); }
        TRY(comm_barrier(ctx));                  // every window is reset before any rank's kernel can write into it
        CU(cudaEventRecord(e0, st));
        TRY(em_launch(s, max_iter, stop_on_conv, &it, &d, nullptr, true));
        CU(cudaEventRecord(e1, st));
    } else {
        if (!s->d_qpart) { TRY(dev_alloc(&s->d_qpart, 2 * (size_t)s->index->T + 2)); }
        CU(cudaMemsetAsync(s->d_qpart, 0, (2 * (size_t)s->index->T + 2) * 8, st));
        double *d_sum = s->d_qpart + s->index->T;
        unsigned long long *d_dmax = (unsigned long long *)(s->d_qpart + 2 * (size_t)s->index->T);
        CU(cudaEventRecord(e0, st));
        while (it < max_iter) {
            int one = 0; double fd = 0;
            TRY(em_launch(s, 1, 0, &one, &fd, nullptr, false));          // E-phase + partial M-phase of this rank's classes
            TRY(comm_allreduce_f64(ctx, s->d_qpart, d_sum, (size_t)(P > 0 ? P : 1)));
            CU(cudaMemsetAsync(d_dmax, 0, 8, st));
            if (P > 0) { k_update_sharded<<<(P + 255) / 256, 256, 0, st>>>(P, s->m.row_n, s->m.row_RsA, d_sum, s->m.theta, s->opts.eps_abs, s->opts.eps_rel, d_dmax); LAUNCHED(ctx); }
            unsigned long long bits = 0;
            CU(cudaMemcpyAsync(&bits, d_dmax, 8, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            memcpy(&d, &bits, 8);
            it++;
            if (stop_on_conv && d <= 1.0) break;
        }
        CU(cudaEventRecord(e1, st));
    }
    CU(cudaStreamSynchronize(st));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (iters_done) *iters_done = it;
    if (final_delta) *final_delta = d;
    if (ms_out) *ms_out = ms;
    return EMSAR_OK;
}

extern "C" int emsar_sample_em_run(emsar_sample *s, int32_t max_iter, int32_t stop_on_conv, int32_t reset_theta,
                                   int32_t *iters_done, double *final_delta, double *elapsed_ms)
{
    CHECK_ARG(s, "emsar_sample_em_run: NULL sample");
    if (!s->prepared) { emsar_set_err("emsar_sample_em_run: call emsar_sample_prepare first"); return EMSAR_ERR_STATE; }
    TRY(ctx_use(s->ctx));
    const int P = s->m.P;
    if (reset_theta) {
        if (P > 0) { k_fill_double2<<<(unsigned)((P + 255) / 256), 256, 0, s->ctx->stream>>>(s->m.theta, P, 1.0); LAUNCHED(s->ctx); }
        s->n_iter = 0;
    }
    if (max_iter <= 0) max_iter = s->opts.max_iter;
    int it = 0; double fd = 0, ms = 0;
    if (s->use_psum) TRY(em_psum_launch(s, max_iter, stop_on_conv, &it, &fd, &ms));
    else if (s->sharded) TRY(em_run_sharded(s, max_iter, stop_on_conv, &it, &fd, &ms));
    else TRY(em_launch(s, max_iter, stop_on_conv, &it, &fd, &ms, false));
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
    dev_free(d_part);
    if (bad || ll < NEAR_LOWEST) ll = NEAR_LOWEST;
    if (loglik) *loglik = ll;
    return EMSAR_OK;
}

extern "C" int emsar_sample_finalize(emsar_sample *s, emsar_solve_out *out)
{
    CHECK_ARG(s && out, "emsar_sample_finalize: NULL argument");
    if (!s->prepared) { emsar_set_err("emsar_sample_finalize: sample not prepared"); return EMSAR_ERR_STATE; }
    emsar_index *ix = s->index; emsar_ctx *ctx = s->ctx; cudaStream_t st = ctx->stream;
    TRY(ctx_use(ctx));
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
    dev_free(buf);
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
    TRY(ctx_use(ctx));
    double *d_fpkm = nullptr;
    TRY(dev_alloc(&d_fpkm, (size_t)ix->T));
    const double nscale = (double)s->N / 1E6, p10 = pow(10, s->delta);
    k_fpkm<<<(ix->T + 255) / 256, 256, 0, st>>>(ix->T, s->d_pos, s->m.theta, s->d_lone, s->d_R, s->d_adj, nscale, p10, d_fpkm);
    LAUNCHED(ctx);
    CU(cudaMemcpyAsync(theta, d_fpkm, (size_t)ix->T * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    dev_free(d_fpkm);
    return EMSAR_OK;
}

__global__ void k_theta_random(int32_t T, const int32_t *__restrict__ pos, unsigned long long seed, double *__restrict__ theta)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const int p = pos[t];
    if (p < 0) return;
    // keyed by the transcript id, not by the row: the start does not depend on how the rows are laid out (one GPU or several)
    const uint64_t h = mix64(seed * 0x9E3779B97F4A7C15ULL + (uint64_t)t + 1);
    const double u = (double)(h >> 11) * (1.0 / 9007199254740992.0);          // [0, 1)
    theta[p] = exp((2.0 * u - 1.0) * 2.302585092994046);                        // 0.1 .. 10
}

extern "C" int emsar_sample_theta_randomize(emsar_sample *s, uint64_t seed)
{
    CHECK_ARG(s, "emsar_sample_theta_randomize: NULL sample");
    if (!s->prepared) { emsar_set_err("emsar_sample_theta_randomize: sample not prepared"); return EMSAR_ERR_STATE; }
    TRY(ctx_use(s->ctx));
    const int32_t T = s->index->T;
    k_theta_random<<<(T + 255) / 256, 256, 0, s->ctx->stream>>>(T, s->d_pos, (unsigned long long)seed, s->m.theta);
    LAUNCHED(s->ctx);
    CU(cudaGetLastError());
    s->n_iter = 0; s->final_delta = INFINITY;
    return EMSAR_OK;
}

extern "C" int emsar_sample_segments_get(emsar_sample *s, double *adjEUMA, double *expected, int32_t *set_id)
{
    CHECK_ARG(s, "emsar_sample_segments_get: NULL sample");
    if (!s->prepared) { emsar_set_err("emsar_sample_segments_get: sample not prepared"); return EMSAR_ERR_STATE; }
    emsar_index *ix = s->index; emsar_ctx *ctx = s->ctx; cudaStream_t st = ctx->stream;
    TRY(ctx_use(ctx));
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
        dev_free(d_fpkm); dev_free(d_ex);
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
    ctx_use(s->ctx);
    cudaStreamSynchronize(s->ctx->stream);
    for (int i = 0; i < 4; i++) if (s->count_ev[i]) cudaEventDestroy(s->count_ev[i]);
    dev_free(s->d_R); dev_free(s->d_hist); dev_free(s->d_flags);
    dev_free(s->d_rd_ptr); dev_free(s->d_rd_tid); dev_free(s->d_rd_fl); dev_free(s->d_rd_aux);
    dev_free(s->d_Wf); dev_free(s->d_adj); dev_free(s->d_amodel); dev_free(s->d_in_model);
    dev_free(s->d_A); dev_free(s->d_Rs); dev_free(s->d_iE); dev_free(s->d_lone); dev_free(s->d_pos);
    for (void *q : s->ps_allocs) dev_free(q);
    dev_free(s->d_state); dev_free(s->d_pack); dev_free(s->d_mcls); dev_free(s->d_halo); dev_free(s->d_chunks); dev_free(s->d_qpart); dev_free(s->d_slots);
    delete s;
    return EMSAR_OK;
}
