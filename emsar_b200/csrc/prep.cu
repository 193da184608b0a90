// Per-sample model build on the device (one-off per alignment file):
//   Wf (transfer_fraglendist_to_Wf, reference emsar_functions.c:2503-2513), adjEUMA (compute_adjEUMA :2517-2523 as
//   scan_rshbucket :2135-2192 applies it), EUMAps (construct_EUMAps :3148-3154), iEUMA (compute_iEUMA :3218-3232),
//   the set decomposition with EUMAcut (build_TC_from_CT_2 :2201-2227, propagate_2 :2234-2259, emsar_main.c:411-425),
//   and the packed ACTIVE model the EM kernel streams: classes that are modelled (in a set, EUMAps > 0) and hold
//   reads, binned by cardinality, plus its transposed CSR with rows binned by length.
// The deterministic fp64 pre-steps use the reference's summation order and no FMA contraction (-fmad=false), so
// Wf, adjEUMA, EUMAps and iEUMA are bit-identical to the CPU reference.
#include <algorithm>
#include <stdarg.h>
#include <cub/cub.cuh>
#include <numeric>

#include "common.cuh"
#include "psum.cuh"

int sample_check_flags(emsar_sample *s);

// ---------------------------------------------------------------------------------------------------
__global__ void k_wf(const int32_t *__restrict__ hist, int frag_min, int nF, int max_fl, double *Wf, long long *N_out)
{
    if (threadIdx.x || blockIdx.x) return;
    double sum = 0;
    for (int i = 0; i < nF; i++) { double w = (double)hist[i + frag_min]; Wf[i] = w; sum += w; }
    for (int i = 0; i < nF; i++) Wf[i] /= sum;
    long long n = 0;
    for (int f = 0; f <= max_fl; f++) n += hist[f];
    *N_out = n;
}

// adjEUMA[c] = sum_i Wf[i] * EUMA[c][i], i ascending. 128 rows per CTA, 32 columns per pass staged through
// shared memory so that the global reads are coalesced while each thread keeps the sequential order.
constexpr int ADJ_ROWS = 128;
__global__ void __launch_bounds__(ADJ_ROWS) k_adjeuma(int64_t C, int nF, const int32_t *__restrict__ euma,
                                                      const uint8_t *__restrict__ has_node, const double *__restrict__ Wf,
                                                      double *__restrict__ adj)
{
    __shared__ int tile[ADJ_ROWS][33];
    const int64_t row0 = (int64_t)blockIdx.x * ADJ_ROWS;
    const int64_t row = row0 + threadIdx.x;
    double acc = 0;
    if (nF == 1) {
        if (row < C) adj[row] = has_node[row] ? 0.0 + Wf[0] * (double)euma[row] : 0.0;
        return;
    }
    for (int c0 = 0; c0 < nF; c0 += 32) {
        for (int idx = threadIdx.x; idx < ADJ_ROWS * 32; idx += ADJ_ROWS) {
            int r = idx >> 5, col = idx & 31;
            int64_t gr = row0 + r;
            tile[r][col] = (gr < C && c0 + col < nF) ? euma[gr * nF + c0 + col] : 0;
        }
        __syncthreads();
        int lim = min(32, nF - c0);
        for (int col = 0; col < lim; col++) acc += Wf[c0 + col] * (double)tile[threadIdx.x][col];
        __syncthreads();
    }
    if (row < C) adj[row] = has_node[row] ? acc : 0.0;
}


// Streaming version for nF a multiple of 4 (row starts are then 16-byte aligned): the only kernel of the path that is bound by
// HBM proper - 4*C*nF bytes, 3.2 GB at config #3. A CTA owns 128 consecutive rows (one contiguous piece of the matrix) and walks
// their columns in chunks of AJ_COLS through an AJ_STAGES-deep cp.async pipeline (16-byte copies straight into shared memory,
// no registers in between), so that ~150 KB per SM are in flight while thread r adds up row r. The sum itself stays strictly sequential in
// the column index (compute_adjEUMA :2517-2523; no FMA contraction), which is what makes the result bit-identical to the
// reference; rows are padded to AJ_COLS + 4 words in shared memory, ((AJ_COLS + 4) / 4) odd, so the 128-bit reads of a
// quarter-warp hit distinct banks.
template <int AJ_ROWS, int AJ_COLS, int AJ_STAGES>
__global__ void __launch_bounds__(AJ_ROWS) k_adjeuma_stream(int64_t C, int nF, const int32_t *__restrict__ euma, const uint8_t *__restrict__ has_node,
                                                             const double *__restrict__ Wf, double *__restrict__ adj)
{
    constexpr int AJ_STRIDE = AJ_COLS + 4, Q = AJ_COLS / 4;      // (AJ_STRIDE / 4) is odd for AJ_COLS = 64, 128
    extern __shared__ __align__(16) int aj_sm[];
    const int tid = threadIdx.x;
    const int64_t row0 = (int64_t)blockIdx.x * AJ_ROWS;
    const int nch = (nF + AJ_COLS - 1) / AJ_COLS;
    auto issue = [&](int ch) {
        if (ch < nch) {
            int *stage = aj_sm + (ch % AJ_STAGES) * (AJ_ROWS * AJ_STRIDE);
#pragma unroll 4
            for (int k = 0; k < Q; k++) {
                const int idx = tid + k * AJ_ROWS, r = idx / Q, q = idx % Q, col = ch * AJ_COLS + q * 4;
                const int64_t gr = row0 + r;
                if (gr < C && col < nF) {
                    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(stage + r * AJ_STRIDE + q * 4);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(euma + gr * nF + col) : "memory");
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    for (int c = 0; c < AJ_STAGES - 1; c++) issue(c);
    double acc = 0;
    for (int ch = 0; ch < nch; ch++) {
        issue(ch + AJ_STAGES - 1);
        asm volatile("cp.async.wait_group %0;" ::"n"(AJ_STAGES - 1) : "memory");
        __syncthreads();
        const int4 *rowp = (const int4 *)(aj_sm + (ch % AJ_STAGES) * (AJ_ROWS * AJ_STRIDE) + tid * AJ_STRIDE);
        const int c0 = ch * AJ_COLS, lim = min(AJ_COLS, nF - c0) >> 2;
        for (int q = 0; q < lim; q++) {
            const int4 v = rowp[q];
            const double *w = Wf + c0 + q * 4;
            acc += w[0] * (double)v.x;
            acc += w[1] * (double)v.y;
            acc += w[2] * (double)v.z;
            acc += w[3] * (double)v.w;
        }
        __syncthreads();                       // the stage is refilled AJ_STAGES - 1 iterations from now
    }
    const int64_t row = row0 + tid;
    if (row < C) adj[row] = has_node[row] ? acc : 0.0;
}
template <int R, int Cc, int S>
static int launch_adjeuma_stream(emsar_sample *s, cudaStream_t st)
{
    emsar_index *ix = s->index;
    constexpr int smem = S * R * (Cc + 4) * 4;
    CU(cudaFuncSetAttribute(k_adjeuma_stream<R, Cc, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));       // per device; cheap
    k_adjeuma_stream<R, Cc, S><<<(unsigned)((ix->C + R - 1) / R), R, smem, st>>>(ix->C, ix->nF, ix->d_euma, ix->d_has_node, s->d_Wf, s->d_adj);
    return EMSAR_OK;
}

static int launch_adjeuma(emsar_sample *s, cudaStream_t st)
{
    emsar_index *ix = s->index;
    if (ix->nF >= 16 && ix->nF % 4 == 0 && !getenv("EMSAR_ADJEUMA_SIMPLE")) {
        // 128 rows x 32 columns x 3 stages = 54 KB per CTA -> 4 CTAs (512 row sums in flight) per SM: the best of the
        // {64,128,256} x {16,32,64,128} x {2,3,4} sweep (profiles/r1i_adjeuma.txt); fewer resident rows starve the fp64 add chains
        TRY((launch_adjeuma_stream<128, 32, 3>(s, st)));
    } else {
        k_adjeuma<<<(unsigned)((ix->C + ADJ_ROWS - 1) / ADJ_ROWS), ADJ_ROWS, 0, st>>>(ix->C, ix->nF, ix->d_euma, ix->d_has_node, s->d_Wf, s->d_adj);
    }
    LAUNCHED(s->ctx);
    CU(cudaGetLastError());
    return EMSAR_OK;
}

extern "C" int emsar_sample_time_adjeuma(emsar_sample *s, int32_t reps, double *ms_per_launch)
{
    CHECK_ARG(s && ms_per_launch && reps > 0, "emsar_sample_time_adjeuma: bad argument");
    if (!s->prepared) { emsar_set_err("emsar_sample_time_adjeuma: sample not prepared"); return EMSAR_ERR_STATE; }
    emsar_ctx *ctx = s->ctx;
    TRY(ctx_use(ctx));
    TRY(launch_adjeuma(s, ctx->stream));                 // warm-up
    CU(cudaEventRecord(ctx->tev0, ctx->stream));
    for (int i = 0; i < reps; i++) TRY(launch_adjeuma(s, ctx->stream));
    CU(cudaEventRecord(ctx->tev1, ctx->stream));
    CU(cudaEventSynchronize(ctx->tev1));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, ctx->tev0, ctx->tev1));
    *ms_per_launch = (double)ms / reps;
    return EMSAR_OK;
}

// EUMAps, modelled / active flags.  nscale = (double)N / 1e6, p10 = pow(10, DELTA) (both formed on the host exactly
// as construct_EUMAps does).
__global__ void k_class_model(int64_t C, int32_t T, const double *__restrict__ adj, const uint8_t *__restrict__ in_model,
                              const int32_t *__restrict__ R, double nscale, double p10, double *__restrict__ amodel,
                              int32_t *__restrict__ act)
{
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c > C) return;
    if (c == C) { act[C - T] = 0; return; }
    double ps = adj[c] / 1E3 * nscale * p10;
    bool modelled = (in_model == nullptr || in_model[c]) && ps > 0;
    amodel[c] = modelled ? ps : 0.0;
    if (c >= T) act[c - T] = (modelled && R[c] > 0) ? 1 : 0;
}

// One warp per transcript over its row of the static multi-class transpose: iEUMA (all classes, ascending cid,
// multiplicity, sequential order), A_t (modelled classes), Rs_t, active degree, lone-singleton flag, participation flag.
__global__ void k_row_stats(int32_t T, const uint32_t *__restrict__ txm_off, const int32_t *__restrict__ txm_cid,
                            const double *__restrict__ adj, const double *__restrict__ amodel, const int32_t *__restrict__ act,
                            const uint8_t *__restrict__ in_model, const int32_t *__restrict__ R,
                            double *__restrict__ iE, double *__restrict__ A, double *__restrict__ Rs, int32_t *__restrict__ deg,
                            uint8_t *__restrict__ lone, uint32_t *__restrict__ rflag)
{
    const int lane = threadIdx.x & 31;
    const int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (t >= T) return;
    double acc_i = adj[t];        // the singleton class cid == tid comes first in ascending cid order
    double acc_a = amodel[t];
    int d = 0; bool any_model = false;
    const uint32_t e0 = txm_off[t], e1 = txm_off[t + 1];
    for (uint32_t e = e0; e < e1; e += 32) {
        const int n = min(32u, e1 - e);
        double vi = 0, va = 0; int a = 0; bool im = false;
        if (lane < n) {
            int cid = txm_cid[e + lane];
            vi = adj[cid]; va = amodel[cid]; a = act[cid - T];
            im = (in_model == nullptr) || in_model[cid];
        }
        for (int l = 0; l < n; l++) {
            acc_i += __shfl_sync(0xffffffffu, vi, l);
            acc_a += __shfl_sync(0xffffffffu, va, l);
        }
        d += __popc(__ballot_sync(0xffffffffu, a != 0));
        any_model = any_model || __any_sync(0xffffffffu, im);
    }
    if (lane == 0) {
        iE[t] = acc_i; A[t] = acc_a;
        Rs[t] = amodel[t] > 0 ? (double)R[t] : 0.0;
        deg[t] = d;
        lone[t] = any_model ? 0 : 1;
        rflag[t] = acc_a > 0 ? 1u : 0u;
        if (t == T - 1) rflag[T] = 0u;
    }
}

// ---- class-range sharding (one sample over several GPUs): rank r keeps the active classes whose member-count prefix
// falls into its nnz-balanced range; everything downstream (degrees, ownership, packing) then sees only those classes
__global__ void k_shard_weights(int64_t n_multi, int32_t T, const uint32_t *__restrict__ cls_off, const int32_t *__restrict__ act, uint32_t *__restrict__ w)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n_multi) return;
    w[i] = (i < n_multi && act[i]) ? cls_off[T + i + 1] - cls_off[T + i] : 0u;
}
__global__ void k_shard_apply(int64_t n_multi, const uint32_t *__restrict__ wpre, int rank, int nranks, int32_t *__restrict__ act)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_multi || !act[i]) return;
    // same rule as emsar_shard_ranges: rank r owns [lower_bound(total*r/R), lower_bound(total*(r+1)/R))
    const unsigned long long total = wpre[n_multi], me = wpre[i];
    const unsigned long long lo = total * (unsigned long long)rank / (unsigned long long)nranks;
    const unsigned long long hi = total * (unsigned long long)(rank + 1) / (unsigned long long)nranks;
    const bool mine = (rank == 0 || me >= lo) && (rank == nranks - 1 || me < hi);
    if (!mine) act[i] = 0;
}

// natural order = the index's locality order (components kept together): flags gathered in that order, scanned, scattered back
__global__ void k_order_gather(int32_t T, const int32_t *__restrict__ order, const uint32_t *__restrict__ rflag, uint32_t *__restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > T) return;
    out[i] = i < T ? rflag[order[i]] : 0u;
}
__global__ void k_order_scatter(int32_t T, const int32_t *__restrict__ order, const uint32_t *__restrict__ pre, uint32_t *__restrict__ nat)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > T) return;
    if (i < T) nat[order[i]] = pre[i]; else nat[T] = pre[T];
}

// participating transcripts in natural order: n = rank, with their active row length
__global__ void k_nat_fill(int32_t T, const uint32_t *__restrict__ rflag, const uint32_t *__restrict__ nat, const int32_t *__restrict__ deg,
                           int32_t *__restrict__ pos, uint32_t *__restrict__ degn, int32_t *__restrict__ tn, int32_t *__restrict__ ecost, int32_t P)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) { degn[P] = 0; ecost[P] = 0; }
    if (t >= T) return;
    if (rflag[t]) { const int n = (int)nat[t]; degn[n] = (uint32_t)deg[t]; tn[n] = t; ecost[n] = 0; }
    else pos[t] = -1;
}

// The member of a class whose row (and so whose CTA) owns it. rule 0: the first member (classes of a CTA are then contiguous in
// cid order). rule 1 (EMSAR_OWNER=light): the member with the fewest active entries, first one on ties - the classes of a hub
// transcript then spread over its partners' CTAs instead of piling their E-phase work onto the hub's CTA.
__global__ void k_class_owner(int64_t n_multi, int32_t T, const uint32_t *__restrict__ cls_off, const int32_t *__restrict__ cls_tid,
                              const int32_t *__restrict__ act, const int32_t *__restrict__ deg, int rule, int32_t *__restrict__ owner)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_multi || !act[i]) return;
    const uint32_t o = cls_off[T + i], e = cls_off[T + i + 1];
    int best = cls_tid[o];
    if (rule == 2) best = cls_tid[o + ((e - o) >> 1)];          // the median member: the classes of a wide module spread over the CTAs that hold its rows
    if (rule == 1) {
        int bd = deg[best];
        for (uint32_t j = o + 1; j < e; j++) { const int u = cls_tid[j], d = deg[u]; if (d < bd) { bd = d; best = u; } }
    }
    owner[i] = best;
}

// E-phase cost lands on the row that owns the class
__global__ void k_class_cost(int64_t n_multi, int32_t T, const uint32_t *__restrict__ cls_off, const int32_t *__restrict__ owner,
                             const int32_t *__restrict__ act, const uint32_t *__restrict__ nat, int32_t *__restrict__ ecost, int per_class, int per_member)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_multi || !act[i]) return;
    atomicAdd(&ecost[nat[owner[i]]], per_member * (int)(cls_off[T + i + 1] - cls_off[T + i]) + per_class);       // gathers + the class's own work (divide, store)
}

__global__ void k_row_cost(int32_t P, const uint32_t *__restrict__ degn, const int32_t *__restrict__ ecost, uint32_t *__restrict__ cost, int per_row, int per_entry)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p > P) return;
    cost[p] = p < P ? (uint32_t)per_entry * degn[p] + (uint32_t)ecost[p] + (uint32_t)per_row : 0u;
}

// cut the rows (natural order) into B ranges of equal cost
__global__ void k_block_bounds(int B, int32_t P, const uint32_t *__restrict__ costp, int32_t *__restrict__ row0)
{
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > B) return;
    if (b == 0) { row0[0] = 0; return; }
    if (b == B) { row0[B] = P; return; }
    const unsigned long long target = (unsigned long long)costp[P] * (unsigned long long)b / (unsigned long long)B;
    int lo = 0, hi = P;
    while (lo < hi) { int mid = (lo + hi) >> 1; if ((unsigned long long)costp[mid] >= target) hi = mid; else lo = mid + 1; }
    row0[b] = lo;
}

__device__ __forceinline__ int seg_of_cid(const int64_t *kseg_cid0, int n_kseg, int64_t cid)
{
    int lo = 0, hi = n_kseg - 1;
    while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (kseg_cid0[mid] <= cid) lo = mid; else hi = mid - 1; }
    return lo;
}
__device__ __forceinline__ int block_of_row(const int32_t *row0, int B, int p)
{
    int lo = 0, hi = B - 1;       // largest b with row0[b] <= p (that CTA is non-empty because p < row0[b+1])
    while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (row0[mid] <= p) lo = mid; else hi = mid - 1; }
    return lo;
}

// sort key: (owner CTA, longest row first); the radix sort is stable, so ties keep their natural order
__global__ void k_sort_keys(int32_t P, int B, const int32_t *__restrict__ row0, const uint32_t *__restrict__ degn,
                            unsigned long long *__restrict__ key, int32_t *__restrict__ val)
{
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= P) return;
    key[n] = ((unsigned long long)block_of_row(row0, B, n) << 32) | (unsigned long long)(0xFFFFFFFFu - degn[n]);
    val[n] = n;
}

__global__ void k_apply_perm(int32_t P, const int32_t *__restrict__ perm, const int32_t *__restrict__ tn, const uint32_t *__restrict__ degn,
                             const double *__restrict__ Rs, const double *__restrict__ A, int32_t *__restrict__ pos,
                             uint32_t *__restrict__ degp, double2 *__restrict__ row_RsA, int32_t *__restrict__ row_n, double2 *__restrict__ rsa_nat)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const int n = perm[p], t = tn[n];
    pos[t] = p;
    row_n[p] = n;
    rsa_nat[n] = make_double2(Rs[t], A[t]);
    degp[p] = degn[n];
    row_RsA[p] = make_double2(Rs[t], A[t]);
}

// M items of every CTA: first groups of its long rows (<= 32 rows and <= M_GROUP_ENTRIES entries each, one warp per
// group), then slices of 32 rows. Pass 1 counts the items of each CTA (greedy grouping), pass 2 (after the prefix over
// CTAs) writes them.
__device__ __forceinline__ int count_long(const uint32_t *degp, int r0, int r1)
{
    int lo = r0, hi = r1;            // rows are sorted longest first: first row with degp <= M_LONG
    while (lo < hi) { int mid = (lo + hi) >> 1; if (degp[mid] <= (uint32_t)M_LONG) hi = mid; else lo = mid + 1; }
    return lo - r0;
}
__global__ void k_block_items_count(int B, const int32_t *__restrict__ row0, const uint32_t *__restrict__ degp, int32_t *__restrict__ nlong,
                                    int32_t *__restrict__ nitems, int32_t *__restrict__ ngroups)
{
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int r0 = row0[b], r1 = row0[b + 1];
    const int nl = count_long(degp, r0, r1);
    int groups = 0, rows = 0; uint32_t ent = 0;
    for (int i = 0; i < nl; i++) {
        const uint32_t d = degp[r0 + i];
        if (rows > 0 && (rows == M_GROUP_ROWS || ent + d > (uint32_t)M_GROUP_ENTRIES)) { groups++; rows = 0; ent = 0; }
        rows++; ent += d;
    }
    if (rows > 0) groups++;
    nlong[b] = nl;
    ngroups[b] = groups;
    nitems[b] = groups + ((r1 - r0 - nl) + 31) / 32;
}
__global__ void k_block_items_prefix(int B, const int32_t *__restrict__ nitems, int32_t *__restrict__ item0)
{
    if (threadIdx.x || blockIdx.x) return;
    int acc = 0;
    for (int b = 0; b < B; b++) { item0[b] = acc; acc += nitems[b]; }
    item0[B] = acc;
}
// one thread per CTA: item descriptors (offsets come later), item sizes, and for every long row its offset inside the group
__global__ void k_block_items_fill(int B, const int32_t *__restrict__ row0, const uint32_t *__restrict__ degp, const int32_t *__restrict__ nlong,
                                   const int32_t *__restrict__ item0, uint32_t *__restrict__ size, int4 *__restrict__ items,
                                   uint32_t *__restrict__ rowbase, int32_t *__restrict__ rowitem)
{
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b == 0) size[item0[B]] = 0;
    if (b >= B) return;
    const int r0 = row0[b], r1 = row0[b + 1], nl = nlong[b];
    int g = item0[b];
    int rows = 0, first = 0; uint32_t ent = 0;
    for (int i = 0; i <= nl; i++) {
        const uint32_t d = i < nl ? degp[r0 + i] : 0;
        if (rows > 0 && (i == nl || rows == M_GROUP_ROWS || ent + d > (uint32_t)M_GROUP_ENTRIES)) {
            items[g] = make_int4(first, rows, 0, (int)(ent | (1u << 30)));
            size[g] = ent + (uint32_t)rows;          // header + entries
            g++; rows = 0; ent = 0;
        }
        if (i == nl) break;
        if (rows == 0) first = i;
        rowbase[r0 + i] = ent;                      // entries of the earlier rows of the group
        rowitem[r0 + i] = g;
        rows++; ent += d;
    }
    const int nrows = r1 - r0;
    for (int slot0 = nl; slot0 < nrows; slot0 += 32, g++) {
        const uint32_t d = degp[r0 + slot0];
        items[g] = make_int4(slot0, min(32, nrows - slot0), 0, (int)d);
        size[g] = 32u * d;
    }
}

__global__ void k_item_offsets(int n_items, const uint32_t *__restrict__ item_off, int4 *__restrict__ items)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n_items) items[g].z = (int)item_off[g];
}
// group headers (row lengths); the unused lanes of a CTA's last (partial) slice point at the zero slot
__global__ void k_item_finish(int n_items, int B, const int32_t *__restrict__ item0, const int32_t *__restrict__ row0, const int32_t *__restrict__ nres,
                              const uint32_t *__restrict__ degp, const int4 *__restrict__ items, int32_t *__restrict__ m_cls)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_items) return;
    const int4 it = items[g];
    int lo = 0, hi = B - 1;
    while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (item0[mid] <= g) lo = mid; else hi = mid - 1; }
    if ((it.w >> 30) == 1) {
        for (int r = 0; r < it.y; r++) m_cls[(uint32_t)it.z + r] = (int32_t)degp[row0[lo] + it.x + r];
    } else if (it.y < 32) {
        const int zero = nres[lo], len = it.w;
        for (int j = 0; j < len; j++)
            for (int l = it.y; l < 32; l++) m_cls[(uint32_t)it.z + (uint32_t)j * 32u + (uint32_t)l] = zero;
    }
}

// cell = (owner CTA, cardinality segment). The new compact id of a class is its rank in the order (cell, old compact id): the
// classes of a cell stay in cid order whatever the ownership rule is, so the layout is a function of the model alone.
__global__ void k_class_cells(int64_t n_multi, int32_t T, int n_kseg, int B, const int64_t *__restrict__ kseg_cid0,
                              const int32_t *__restrict__ owner, const int32_t *__restrict__ act,
                              const int32_t *__restrict__ newid, const int32_t *__restrict__ pos, const int32_t *__restrict__ row0,
                              int32_t *__restrict__ cellof, int32_t *__restrict__ cell_cnt, unsigned long long *__restrict__ keys, int32_t *__restrict__ vals)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_multi || !act[i]) return;
    const int ob = block_of_row(row0, B, pos[owner[i]]);
    const int cell = ob * n_kseg + seg_of_cid(kseg_cid0, n_kseg, T + i);
    const int j = newid[i];
    cellof[j] = cell;
    atomicAdd(&cell_cnt[cell], 1);
    keys[j] = ((unsigned long long)(uint32_t)cell << 32) | (unsigned long long)(uint32_t)j;
    vals[j] = (int32_t)i;
}

// after sorting (cell, old id): the class at sorted position r gets the new compact id r
__global__ void k_class_newid(int64_t C_a, const int32_t *__restrict__ sorted_vals, int32_t *__restrict__ newid2)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < C_a) newid2[sorted_vals[r]] = (int32_t)r;
}

__global__ void k_cell_sizes(int n_cells, int n_kseg, const int32_t *__restrict__ kseg_k, const int32_t *__restrict__ cell_cnt,
                             uint32_t *__restrict__ cell_ints, int32_t *__restrict__ cell_tiles)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > n_cells) return;
    if (c == n_cells) { cell_ints[c] = 0; cell_tiles[c] = 0; return; }
    const int k = kseg_k[c % n_kseg], cnt = cell_cnt[c], cpt = e_cls_per_tile(k), cpb = e_cls_per_block(k);
    cell_ints[c] = (uint32_t)((cnt + cpb - 1) / cpb) * 32u * (uint32_t)e_steps(k);
    cell_tiles[c] = (cnt + cpt - 1) / cpt;
}

__global__ void k_block_tables(int B, int n_kseg, const int32_t *__restrict__ clsbase, const int32_t *__restrict__ tilebase,
                               int32_t *__restrict__ cls0, int32_t *__restrict__ etile0)
{
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > B) return;
    cls0[b] = clsbase[b * n_kseg];
    etile0[b] = tilebase[b * n_kseg];
}

// ---- halo lists -----------------------------------------------------------------------------------------
// every reference that leaves the owner's range is appended as (CTA << 32 | global index); sorting + unique gives each
// CTA its list of distinct remote rows (E side) / classes (M side), in a deterministic order
__global__ void k_halo_collect_e(int64_t n_multi, int32_t T, int B, const uint32_t *__restrict__ cls_off, const int32_t *__restrict__ cls_tid,
                                 const int32_t *__restrict__ owner, const int32_t *__restrict__ act, const int32_t *__restrict__ pos, const int32_t *__restrict__ row0,
                                 unsigned long long *__restrict__ keys, unsigned int *__restrict__ count)
{
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n_multi || !act[i]) return;
    const uint32_t o = cls_off[T + i], e = cls_off[T + i + 1];
    const int ob = block_of_row(row0, B, pos[owner[i]]);
    const int r0 = row0[ob], r1 = row0[ob + 1];
    for (uint32_t j = o + lane; j < e; j += 32) {
        const int p = pos[cls_tid[j]];
        if (p < r0 || p >= r1) keys[atomicAdd(count, 1u)] = ((unsigned long long)ob << 32) | (unsigned long long)(uint32_t)p;
    }
}

__global__ void k_halo_collect_m(int32_t T, int B, const uint32_t *__restrict__ txm_off, const int32_t *__restrict__ txm_cid,
                                 const int32_t *__restrict__ act, const int32_t *__restrict__ newid2, const int32_t *__restrict__ pos,
                                 const int32_t *__restrict__ row0, const int32_t *__restrict__ cls0,
                                 unsigned long long *__restrict__ keys, unsigned int *__restrict__ count)
{
    const int lane = threadIdx.x & 31;
    const int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (t >= T) return;
    const int p = pos[t];
    if (p < 0) return;
    const int b = block_of_row(row0, B, p);
    const int c0 = cls0[b], c1 = cls0[b + 1];
    for (uint32_t e = txm_off[t] + lane; e < txm_off[t + 1]; e += 32) {
        const int i = txm_cid[e] - T;
        if (!act[i]) continue;
        const int id = newid2[i];
        if (id < c0 || id >= c1) keys[atomicAdd(count, 1u)] = ((unsigned long long)b << 32) | (unsigned long long)(uint32_t)id;
    }
}

__global__ void k_halo_ranges(int B, unsigned int n, const unsigned long long *__restrict__ uniq, int32_t *__restrict__ h0)
{
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > B) return;
    const unsigned long long target = (unsigned long long)b << 32;
    unsigned int lo = 0, hi = n;
    while (lo < hi) { unsigned int mid = (lo + hi) >> 1; if (uniq[mid] >= target) hi = mid; else lo = mid + 1; }
    h0[b] = (int32_t)lo;
}
__global__ void k_halo_list(unsigned int n, const unsigned long long *__restrict__ uniq, int32_t *__restrict__ list)
{
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) list[i] = (int32_t)(uint32_t)uniq[i];
}
__device__ __forceinline__ int halo_find(const unsigned long long *uniq, int h0, int h1, unsigned long long key)
{
    int lo = h0, hi = h1;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (uniq[mid] >= key) hi = mid; else lo = mid + 1; }
    return lo - h0;       // the key is present by construction
}

// One warp per multi-tid class: members of an active class as encoded rows (>= 0: slot in the owner's shared theta,
// < 0: ~global row), its read count, and the flag that q must also go to global memory.
__global__ void k_pack_classes(int64_t n_multi, int32_t T, int n_kseg, const int32_t *__restrict__ blk_nres, const int32_t *__restrict__ blk_hr0,
                               const int32_t *__restrict__ blk_nhr, const unsigned long long *__restrict__ uniq_e, const int32_t *__restrict__ kseg_k,
                               const uint32_t *__restrict__ cls_off, const int32_t *__restrict__ cls_tid, const int32_t *__restrict__ act,
                               const int32_t *__restrict__ newid, const int32_t *__restrict__ cellof, const int32_t *__restrict__ newid2,
                               const int32_t *__restrict__ clsbase, const uint32_t *__restrict__ intbase, const int32_t *__restrict__ R,
                               const int32_t *__restrict__ pos, const int32_t *__restrict__ row0, const int32_t *__restrict__ cls0,
                               int32_t *__restrict__ e_tid, uint32_t *__restrict__ e_R)
{
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n_multi || !act[i]) return;
    const int jo = newid[i];
    const int cell = cellof[jo];
    const int ob = cell / n_kseg, k = kseg_k[cell % n_kseg];
    const int jn = newid2[i];
    const int jl = jn - clsbase[cell];
    const int r0 = row0[ob], r1 = row0[ob + 1];
    const int nres = blk_nres[ob];
    const uint32_t o = cls_off[T + i];
    const uint32_t base = intbase[cell];
    const int lg = e_lgG(k), G = 1 << lg, steps = e_steps(k), cpb = 32 >> lg;
    const uint32_t rb = (uint32_t)(jl / cpb), cb = (uint32_t)(jl % cpb);
    const int zero_enc = (r1 - r0) + blk_nhr[ob];            // the zero-theta slot follows the CTA's own and halo rows
    bool remote = false;
    for (int jj = lane; jj < steps * G; jj += 32) {
        int enc = zero_enc;
        if (jj < k) {
            const int p = pos[cls_tid[o + jj]];
            const bool local = p >= r0 && p < r1;
            remote = remote || !local;
            enc = p - r0;
            if (!local) {
                const int h = halo_find(uniq_e, blk_hr0[ob], blk_hr0[ob + 1], ((unsigned long long)ob << 32) | (unsigned long long)(uint32_t)p);
                enc = h < blk_nhr[ob] ? (r1 - r0) + h : ~p;
            }
        }
        e_tid[base + rb * 32u * (uint32_t)steps + (uint32_t)(jj >> lg) * 32u + cb * (uint32_t)G + (uint32_t)(jj & (G - 1))] = enc;
    }
    remote = __any_sync(0xffffffffu, remote);
    if (lane == 0) {
        const bool resident = (jn - cls0[ob]) < nres;
        e_R[jn] = (uint32_t)R[T + i] | ((remote || !resident) ? 0x80000000u : 0u);
    }
}

// One warp per transcript: the ACTIVE entries of its transposed row, in ascending cid order, as encoded classes
// (>= 0: slot in the row owner's shared q, < 0: ~global compact id), written into the row's slice column (padded with
// the zero slot up to the slice's length) or, for a long row, contiguously.
__global__ void k_scatter_rows(int32_t T, int B, const int32_t *__restrict__ blk_nres, const int32_t *__restrict__ blk_hc0, const int32_t *__restrict__ blk_nhc,
                               const unsigned long long *__restrict__ uniq_m, const uint32_t *__restrict__ txm_off, const int32_t *__restrict__ txm_cid,
                               const int32_t *__restrict__ act, const int32_t *__restrict__ newid2, const int32_t *__restrict__ pos,
                               const int32_t *__restrict__ row0, const int32_t *__restrict__ cls0, const int32_t *__restrict__ nlong,
                               const int32_t *__restrict__ item0, const uint32_t *__restrict__ rowbase,
                               const int32_t *__restrict__ rowitem, const int32_t *__restrict__ ngroups, const int4 *__restrict__ items,
                               int32_t *__restrict__ m_cls)
{
    const int lane = threadIdx.x & 31;
    const int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (t >= T) return;
    const int p = pos[t];
    if (p < 0) return;
    const int b = block_of_row(row0, B, p);
    const int c0 = cls0[b], c1 = cls0[b + 1], nres = blk_nres[b];
    const int slot = p - row0[b], nl = nlong[b];
    uint32_t base, stride, len;
    if (slot < nl) {
        const int4 it = items[rowitem[p]];
        base = (uint32_t)it.z + (uint32_t)it.y + rowbase[p]; stride = 1; len = 0;      // long rows are never padded
    } else {
        const int4 it = items[item0[b] + ngroups[b] + ((slot - nl) >> 5)];
        base = (uint32_t)it.z + (uint32_t)((slot - nl) & 31); stride = 32; len = (uint32_t)it.w & 0x3fffffffu;
    }
    uint32_t out = 0;
    const uint32_t e0 = txm_off[t], e1 = txm_off[t + 1];
    for (uint32_t e = e0; e < e1; e += 32) {
        int a = 0, id = 0;
        if (e + lane < e1) { int i = txm_cid[e + lane] - T; a = act[i]; if (a) id = newid2[i]; }
        unsigned m = __ballot_sync(0xffffffffu, a != 0);
        if (a) {
            int enc;
            if (id >= c0 && id < c1) enc = (id - c0) < nres ? id - c0 : ~id;
            else {
                const int h = halo_find(uniq_m, blk_hc0[b], blk_hc0[b + 1], ((unsigned long long)b << 32) | (unsigned long long)(uint32_t)id);
                enc = h < blk_nhc[b] ? nres + 1 + h : ~id;
            }
            m_cls[base + (out + (uint32_t)__popc(m & ((1u << lane) - 1))) * stride] = enc;
        }
        out += __popc(m);
    }
    for (uint32_t j = out + lane; j < len; j += 32) m_cls[base + j * stride] = nres;    // slice padding -> the zero slot
}

__global__ void k_etiles(int n_tiles, int n_cells, int n_kseg, const int32_t *__restrict__ kseg_k, const int32_t *__restrict__ tilebase,
                         const int32_t *__restrict__ clsbase, const int32_t *__restrict__ cell_cnt, const uint32_t *__restrict__ intbase,
                         int4 *__restrict__ tiles)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_tiles) return;
    int lo = 0, hi = n_cells - 1;      // largest cell with tilebase[cell] <= g (non-empty by construction)
    while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (tilebase[mid] <= g) lo = mid; else hi = mid - 1; }
    const int c = lo, k = kseg_k[c % n_kseg], cpt = e_cls_per_tile(k), cpb = e_cls_per_block(k), steps = e_steps(k), lt = g - tilebase[c];
    int4 t;
    t.x = clsbase[c] + lt * cpt;
    t.y = min(cpt, cell_cnt[c] - lt * cpt);
    t.z = (int)(intbase[c] + (uint32_t)lt * (uint32_t)(cpt / cpb) * 32u * (uint32_t)steps);
    t.w = steps | (e_lgG(k) << 12);
    tiles[g] = t;
}

__global__ void k_fill_double(double *p, int64_t n, double v)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void k_fill_int(int32_t *p, int64_t n, int32_t v)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ---------------------------------------------------------------------------------------------------
// Host: sequence-sharing sets with the EUMAcut loop (emsar_main.c:411-425). Union-find over transcripts using
// every class that survives the cut; set ids are numbered in first-seen cid order like the reference's DFS.
static int uf_find(std::vector<int32_t> &p, int x)
{
    while (p[x] != x) { p[x] = p[p[x]]; x = p[x]; }
    return x;
}

static void compute_sets(const emsar_index *ix, const std::vector<double> &adj, double *eumacut, int max_ntid,
                         std::vector<int32_t> *CS, int32_t *max_sid)
{
    const int32_t T = ix->T;
    const int64_t C = ix->C;
    std::vector<int32_t> par((size_t)T), sz((size_t)T);
    for (;;) {
        std::iota(par.begin(), par.end(), 0);
        for (int64_t c = T; c < C; c++) {
            if (adj[(size_t)c] < *eumacut) continue;                        // propagate_2 :2242
            uint32_t o = ix->h_cls_off[(size_t)c], e = ix->h_cls_off[(size_t)c + 1];
            int r0 = uf_find(par, ix->h_cls_tid[o]);
            for (uint32_t j = o + 1; j < e; j++) { int r = uf_find(par, ix->h_cls_tid[j]); if (r != r0) par[(size_t)r] = r0; }
        }
        std::fill(sz.begin(), sz.end(), 0);
        int mx = 0;
        for (int32_t t = 0; t < T; t++) { int r = uf_find(par, t); if (++sz[(size_t)r] > mx) mx = sz[(size_t)r]; }
        if (mx > max_ntid) { *eumacut += 2; continue; }                     // EUMACUT_INCREMENT, emsar_main.c:417-423
        break;
    }
    std::vector<int32_t> label((size_t)T, -1);
    int32_t next = 0;
    CS->assign((size_t)C, -1);
    for (int64_t c = 0; c < C; c++) {
        if (c >= T && adj[(size_t)c] < *eumacut) continue;
        int r = uf_find(par, ix->h_cls_tid[ix->h_cls_off[(size_t)c]]);
        if (label[(size_t)r] < 0) label[(size_t)r] = next++;
        (*CS)[(size_t)c] = label[(size_t)r];
    }
    *max_sid = next - 1;
}

// ---------------------------------------------------------------------------------------------------
// Device: the same decomposition (propagate_2 :2234-2259 with the EUMAcut loop of emsar_main.c:411-425) as label propagation. Every
// transcript starts as its own label; a class that survives the cut pulls all its members to the smallest label among them
// (atomicMin), labels are shortened by pointer jumping, until nothing changes: a transcript's label is then the smallest tid of its
// set. The reference numbers the sets in the order its scan first meets them, and it meets the singleton class of every transcript
// (cid == tid) before any multi-tid class: set ids are the ranks of those smallest tids. Integer work only: the result is exact.
__global__ void k_cc_init(int32_t T, int32_t *__restrict__ label)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < T) label[t] = t;
}
__global__ void k_cc_hook(int64_t n_multi, int32_t T, const uint32_t *__restrict__ cls_off, const int32_t *__restrict__ cls_tid, const double *__restrict__ adj,
                          double cut, int32_t *__restrict__ label, int *__restrict__ changed)
{
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n_multi) return;
    if (adj[T + i] < cut) return;                                   // propagate_2 :2242
    const uint32_t o = cls_off[T + i], e = cls_off[T + i + 1];
    int m = 0x7fffffff;
    for (uint32_t j = o + lane; j < e; j += 32) m = min(m, label[cls_tid[j]]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, d));
    bool ch = false;
    for (uint32_t j = o + lane; j < e; j += 32) {
        const int t = cls_tid[j];
        if (label[t] > m) { atomicMin(&label[t], m); ch = true; }
        const int r = label[m];                                     // keep the trees shallow: the root of the minimum, too
        if (r < m) atomicMin(&label[t], r);
    }
    if (__any_sync(0xffffffffu, ch) && lane == 0) *changed = 1;
}
__global__ void k_cc_jump(int32_t T, int32_t *__restrict__ label, int *__restrict__ changed)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    int l = label[t];
    int r = label[l];
    if (r != l) {
        while (label[r] != r) r = label[r];
        label[t] = r;
        *changed = 1;
    }
}
__global__ void k_cc_sizes(int32_t T, const int32_t *__restrict__ label, int32_t *__restrict__ size, uint32_t *__restrict__ isroot)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > T) return;
    if (t == T) { isroot[T] = 0; return; }
    atomicAdd(&size[label[t]], 1);
    isroot[t] = label[t] == t ? 1u : 0u;
}
__global__ void k_cc_max(int32_t T, const int32_t *__restrict__ size, int32_t *__restrict__ mx)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int v = t < T ? size[t] : 0;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, d));
    if ((threadIdx.x & 31) == 0 && v > 0) atomicMax(mx, v);
}
__global__ void k_cc_class_sets(int64_t C, int32_t T, const uint32_t *__restrict__ cls_off, const int32_t *__restrict__ cls_tid, const double *__restrict__ adj,
                                double cut, const int32_t *__restrict__ label, const uint32_t *__restrict__ sid_of_root, int32_t *__restrict__ CS,
                                uint8_t *__restrict__ in_model)
{
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    int cs = -1;
    if (c < T || adj[c] >= cut) cs = (int)sid_of_root[label[cls_tid[cls_off[c]]]];
    CS[c] = cs;
    in_model[c] = cs >= 0 ? 1 : 0;
}

// sets on the device: fills d_CS [C] and d_in_model [C], raises *eumacut by 2 until no set has more than max_ntid transcripts
static int device_sets(emsar_sample *s, double *eumacut, int max_ntid, int32_t *d_CS, uint8_t *d_in_model, int32_t *max_sid)
{
    emsar_index *ix = s->index; emsar_ctx *ctx = s->ctx; cudaStream_t st = ctx->stream;
    const int32_t T = ix->T;
    const int64_t C = ix->C, nm = ix->n_multi;
    int32_t *d_label = nullptr, *d_size = nullptr, *d_flag = nullptr;
    uint32_t *d_isroot = nullptr, *d_sid = nullptr;
    TRY(dev_alloc(&d_label, (size_t)T + 1)); TRY(dev_alloc(&d_size, (size_t)T + 1)); TRY(dev_alloc(&d_flag, 4));
    TRY(dev_alloc(&d_isroot, (size_t)T + 1)); TRY(dev_alloc(&d_sid, (size_t)T + 1));
    struct G { int32_t *a, *b, *c; uint32_t *d, *e; ~G() { dev_free(a); dev_free(b); dev_free(c); dev_free(d); dev_free(e); } } g{d_label, d_size, d_flag, d_isroot, d_sid};
    size_t scan_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (uint32_t *)nullptr, (uint32_t *)nullptr, T + 1);
    void *d_tmp = nullptr;
    TRY(ctx_scratch(ctx, scan_bytes + 256, &d_tmp));
    const unsigned bt = (unsigned)((T + 255) / 256), bc = (unsigned)((nm * 32 + 255) / 256);
    for (;;) {
        k_cc_init<<<bt, 256, 0, st>>>(T, d_label);
        LAUNCHED(ctx);
        for (int round = 0; round < 4096; round++) {
            CU(cudaMemsetAsync(d_flag, 0, 4, st));
            if (nm > 0) k_cc_hook<<<bc, 256, 0, st>>>(nm, T, ix->d_cls_off, ix->d_cls_tid, s->d_adj, *eumacut, d_label, d_flag);
            k_cc_jump<<<bt, 256, 0, st>>>(T, d_label, d_flag);
            ctx->launches += 2;
            int changed = 0;
            CU(cudaMemcpyAsync(&changed, d_flag, 4, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            if (!changed) break;
        }
        CU(cudaMemsetAsync(d_size, 0, ((size_t)T + 1) * 4, st));
        CU(cudaMemsetAsync(d_flag, 0, 4, st));
        k_cc_sizes<<<(unsigned)((T + 1 + 255) / 256), 256, 0, st>>>(T, d_label, d_size, d_isroot);
        k_cc_max<<<bt, 256, 0, st>>>(T, d_size, d_flag);
        ctx->launches += 2;
        int mx = 0;
        CU(cudaMemcpyAsync(&mx, d_flag, 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (mx > max_ntid) { *eumacut += 2; continue; }             // EUMACUT_INCREMENT, emsar_main.c:417-423
        break;
    }
    CU(cub::DeviceScan::ExclusiveSum(d_tmp, scan_bytes, d_isroot, d_sid, T + 1, st));
    k_cc_class_sets<<<(unsigned)((C + 255) / 256), 256, 0, st>>>(C, T, ix->d_cls_off, ix->d_cls_tid, s->d_adj, *eumacut, d_label, d_sid, d_CS, d_in_model);
    ctx->launches += 2;
    uint32_t nsets = 0;
    CU(cudaMemcpyAsync(&nsets, d_sid + T, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    *max_sid = (int32_t)nsets - 1;
    return EMSAR_OK;
}

// ---------------------------------------------------------------------------------------------------
template <class T> static T *arena_take(char *&cur, size_t n)
{
    T *p = (T *)cur;
    cur += ((n * sizeof(T) + 255) / 256) * 256;
    return p;
}

#include "prep_psum.inl"

extern "C" int emsar_sample_prepare(emsar_sample *s, const emsar_solve_opts *opts_in)
{
    CHECK_ARG(s, "emsar_sample_prepare: NULL sample");
    if (!s->have_counts) { emsar_set_err("emsar_sample_prepare: no counts yet (call emsar_sample_count / counts_set first)"); return EMSAR_ERR_STATE; }
    emsar_index *ix = s->index;
    emsar_ctx *ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    TRY(ctx_use(ctx));
    emsar_solve_opts o;
    memset(&o, 0, sizeof(o));
    if (opts_in) o = *opts_in;
    if (!(o.eps_abs > 0)) o.eps_abs = 1e-7;
    if (!(o.eps_rel > 0)) o.eps_rel = 1e-10;
    if (o.max_iter <= 0) o.max_iter = 200000;
    if (o.max_ntid_per_sid <= 0) o.max_ntid_per_sid = 5000;
    s->opts = o;
    s->opts.in_model = nullptr;
    TRY(sample_check_flags(s));
    cudaEvent_t e0 = ctx->ev0, e1 = ctx->ev1;
    CU(cudaEventRecord(e0, st));
    const int32_t T = ix->T;
    const int64_t C = ix->C, nm = ix->n_multi;
    const int n_kseg = (int)ix->kseg.size();
    // ---- per-sample dense arrays (allocated once per sample) ----
    if (!s->d_Wf) {
        TRY(dev_alloc(&s->d_Wf, (size_t)ix->nF + 1));
        TRY(dev_alloc(&s->d_adj, (size_t)C));
        TRY(dev_alloc(&s->d_amodel, (size_t)C));
        TRY(dev_alloc(&s->d_in_model, (size_t)C));
        TRY(dev_alloc(&s->d_A, (size_t)T));
        TRY(dev_alloc(&s->d_Rs, (size_t)T));
        TRY(dev_alloc(&s->d_iE, (size_t)T));
        TRY(dev_alloc(&s->d_lone, (size_t)T));
        TRY(dev_alloc(&s->d_pos, (size_t)T + 1));
    }
    long long *d_N = (long long *)(s->d_Wf + ix->nF);
    k_wf<<<1, 32, 0, st>>>(s->d_hist, ix->frag_min, ix->nF, ix->max_fl, s->d_Wf, d_N);
    LAUNCHED(ctx);
    TRY(launch_adjeuma(s, st));
    long long N = 0;
    CU(cudaMemcpyAsync(&N, d_N, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    s->N = N;
    s->delta = o.delta;
    // ---- sets / EUMAcut ----
    s->eumacut = o.eumacut;
    s->have_CS = false;
    const uint8_t *d_in_model = nullptr;
    if (o.in_model) {
        CU(cudaMemcpyAsync(s->d_in_model, o.in_model, (size_t)C, cudaMemcpyHostToDevice, st));
        d_in_model = s->d_in_model;
        s->max_sid = -1;
    } else if (s->eumacut > 0 || ix->max_set_tids > o.max_ntid_per_sid) {
        if (getenv("EMSAR_SETS_HOST")) {          // the host union-find (cross-check of the device decomposition)
            std::vector<double> adj((size_t)C);
            CU(cudaMemcpyAsync(adj.data(), s->d_adj, (size_t)C * 8, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            compute_sets(ix, adj, &s->eumacut, o.max_ntid_per_sid, &s->h_CS, &s->max_sid);
            s->have_CS = true;
            std::vector<uint8_t> im((size_t)C);
            for (int64_t c = 0; c < C; c++) im[(size_t)c] = s->h_CS[(size_t)c] >= 0;
            CU(cudaMemcpyAsync(s->d_in_model, im.data(), (size_t)C, cudaMemcpyHostToDevice, st));
            CU(cudaStreamSynchronize(st));
        } else {
            int32_t *d_CS = nullptr;
            TRY(dev_alloc(&d_CS, (size_t)C));
            int rc_ = device_sets(s, &s->eumacut, o.max_ntid_per_sid, d_CS, s->d_in_model, &s->max_sid);
            if (rc_ == EMSAR_OK) {
                s->h_CS.resize((size_t)C);
                rc_ = cudaMemcpyAsync(s->h_CS.data(), d_CS, (size_t)C * 4, cudaMemcpyDeviceToHost, st) == cudaSuccess && cudaStreamSynchronize(st) == cudaSuccess ? EMSAR_OK : EMSAR_ERR_CUDA;
                s->have_CS = rc_ == EMSAR_OK;
            }
            dev_free(d_CS);
            if (rc_ != EMSAR_OK) return rc_;
        }
        d_in_model = s->d_in_model;
    } else {
        s->max_sid = ix->n_sets_nocut - 1;
    }
    // ---- scratch carve-up ----
    const int B = ctx->prop.multiProcessorCount * ctx->em_blocks_per_sm;
    const int n_cells = B * (n_kseg > 0 ? n_kseg : 1);
    size_t cub_bytes = 0, b1 = 0, b2 = 0, b3 = 0, b4 = 0;
    const int scan_max = (int)std::max<int64_t>(std::max<int64_t>(nm + 1, (int64_t)T + 1), (int64_t)n_cells + 1);
    cub::DeviceScan::ExclusiveSum(nullptr, b1, (int32_t *)nullptr, (int32_t *)nullptr, scan_max);
    cub::DeviceScan::ExclusiveSum(nullptr, b2, (uint32_t *)nullptr, (uint32_t *)nullptr, scan_max);
    cub::DeviceScan::ExclusiveSum(nullptr, b3, (uint32_t *)nullptr, (int32_t *)nullptr, scan_max);
    cub::DeviceRadixSort::SortPairs(nullptr, b4, (unsigned long long *)nullptr, (unsigned long long *)nullptr, (int32_t *)nullptr, (int32_t *)nullptr, T + 1, 0, 44);
    const size_t href_max = (size_t)ix->nnz_multi + 1;          // remote references of one side: at most every member entry
    size_t b5 = 0, b6 = 0, b7 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b7, (unsigned long long *)nullptr, (unsigned long long *)nullptr, (int32_t *)nullptr, (int32_t *)nullptr, (int)(nm + 1), 0, 64);
    cub::DeviceRadixSort::SortKeys(nullptr, b5, (unsigned long long *)nullptr, (unsigned long long *)nullptr, (int)href_max, 0, 44);
    cub::DeviceSelect::Unique(nullptr, b6, (unsigned long long *)nullptr, (unsigned long long *)nullptr, (unsigned int *)nullptr, (int)href_max);
    cub_bytes = std::max(std::max(std::max(b1, b2), std::max(b3, b4)), std::max(std::max(b5, b6), b7));
    size_t need = ((cub_bytes + 255) / 256) * 256 + (size_t)(nm + 1) * 20 + (size_t)(T + 1) * 72 + (size_t)(2 * (size_t)T + B + 64) * 8 + (size_t)(B + 1) * 16 + (size_t)(n_cells + 1) * 28 + 4 * href_max * 8 + 64 * 256;
    void *scr = nullptr;
    TRY(ctx_scratch(ctx, need, &scr));
    char *cur = (char *)scr;
    void *d_cub = arena_take<char>(cur, cub_bytes);
    int32_t *d_act = arena_take<int32_t>(cur, (size_t)nm + 1);
    int32_t *d_newid = arena_take<int32_t>(cur, (size_t)nm + 1);
    int32_t *d_newid2 = arena_take<int32_t>(cur, (size_t)nm + 1);
    int32_t *d_cellof = arena_take<int32_t>(cur, (size_t)nm + 1);
    int32_t *d_owner = arena_take<int32_t>(cur, (size_t)nm + 1);
    uint32_t *d_rflag = arena_take<uint32_t>(cur, (size_t)T + 1);
    uint32_t *d_nat = arena_take<uint32_t>(cur, (size_t)T + 1);
    int32_t *d_deg = arena_take<int32_t>(cur, (size_t)T + 1);
    uint32_t *d_degn = arena_take<uint32_t>(cur, (size_t)T + 1);
    uint32_t *d_degp = arena_take<uint32_t>(cur, (size_t)T + 1);
    int32_t *d_tn = arena_take<int32_t>(cur, (size_t)T + 1);
    int32_t *d_ecost = arena_take<int32_t>(cur, (size_t)T + 1);
    uint32_t *d_cost = arena_take<uint32_t>(cur, (size_t)T + 1);
    uint32_t *d_costp = arena_take<uint32_t>(cur, (size_t)T + 1);
    unsigned long long *d_key = arena_take<unsigned long long>(cur, (size_t)T + 1);
    unsigned long long *d_key2 = arena_take<unsigned long long>(cur, (size_t)T + 1);
    int32_t *d_val = arena_take<int32_t>(cur, (size_t)T + 1);
    int32_t *d_perm = arena_take<int32_t>(cur, (size_t)T + 1);
    int32_t *d_cell_cnt = arena_take<int32_t>(cur, (size_t)n_cells + 1);
    int32_t *d_cell_first = arena_take<int32_t>(cur, (size_t)n_cells + 1);
    uint32_t *d_cell_ints = arena_take<uint32_t>(cur, (size_t)n_cells + 1);
    int32_t *d_cell_tiles = arena_take<int32_t>(cur, (size_t)n_cells + 1);
    int32_t *d_clsbase = arena_take<int32_t>(cur, (size_t)n_cells + 1);
    uint32_t *d_intbase = arena_take<uint32_t>(cur, (size_t)n_cells + 1);
    int32_t *d_tilebase = arena_take<int32_t>(cur, (size_t)n_cells + 1);
    int32_t *d_nlong = arena_take<int32_t>(cur, (size_t)B + 1);
    int32_t *d_nitems = arena_take<int32_t>(cur, (size_t)B + 1);
    int32_t *d_ngroups = arena_take<int32_t>(cur, (size_t)B + 1);
    uint32_t *d_rowbase = arena_take<uint32_t>(cur, (size_t)T + 1);
    int32_t *d_rowitem = arena_take<int32_t>(cur, (size_t)T + 1);
    uint32_t *d_isize = arena_take<uint32_t>(cur, 2 * (size_t)T + B + 64);
    uint32_t *d_ioff = arena_take<uint32_t>(cur, 2 * (size_t)T + B + 64);
    unsigned long long *d_hkeys = arena_take<unsigned long long>(cur, href_max);
    unsigned long long *d_hsort = arena_take<unsigned long long>(cur, href_max);
    unsigned long long *d_uniq_e = arena_take<unsigned long long>(cur, href_max);
    unsigned long long *d_uniq_m = arena_take<unsigned long long>(cur, href_max);
    unsigned int *d_hcount = (unsigned int *)arena_take<unsigned int>(cur, 8);
    int32_t *d_pos = s->d_pos;   // t -> row (kept for finalize)
    // ---- class model + active scan ----
    const double nscale = (double)N / 1E6;
    const double p10 = pow(10, o.delta);
    k_class_model<<<(unsigned)((C + 1 + 255) / 256), 256, 0, st>>>(C, T, s->d_adj, d_in_model, s->d_R, nscale, p10, s->d_amodel, d_act);
    LAUNCHED(ctx);
    s->sharded = o.sharded != 0 && ctx->nranks > 1;
    if (o.sharded && !ctx->nccl_comm) { emsar_set_err("sharded solve without a communicator (emsar_comm_init)"); return EMSAR_ERR_STATE; }
    // k_em_psum wherever the sample is eligible (EMSAR_EM_MODE=legacy|barrier|pipe keeps the older kernels). A class-sharded sample is then
    // NOT cut by class range: every rank packs the whole model for nranks x B virtual CTAs and runs its own B of them.
    const char *em_mode_ = getenv("EMSAR_EM_MODE");
    const char *sh_mode_ = getenv("EMSAR_SHARD_MODE");
    const bool want_psum = (!em_mode_ || !strcmp(em_mode_, "psum")) && !getenv("EMSAR_OWNER") && !s->force_legacy &&
                           !(s->sharded && (ctx->win_state == -1 || (sh_mode_ && !strcmp(sh_mode_, "nccl"))));
    if (s->sharded && nm > 0 && !want_psum) {
        uint32_t *d_w = (uint32_t *)d_newid2, *d_wpre = (uint32_t *)d_cellof;      // scratch reuse: both are written later
        k_shard_weights<<<(unsigned)((nm + 1 + 255) / 256), 256, 0, st>>>(nm, T, ix->d_cls_off, d_act, d_w);
        CU(cub::DeviceScan::ExclusiveSum(d_cub, cub_bytes, d_w, d_wpre, (int)(nm + 1), st));
        k_shard_apply<<<(unsigned)((nm + 255) / 256), 256, 0, st>>>(nm, d_wpre, ctx->rank, ctx->nranks, d_act);
        ctx->launches += 3;
    }
    CU(cub::DeviceScan::ExclusiveSum(d_cub, cub_bytes, d_act, d_newid, (int)(nm + 1), st));
    LAUNCHED(ctx);
    // ---- row statistics and the natural-order numbering of the participating rows ----
    k_row_stats<<<(unsigned)(((int64_t)T * 32 + 255) / 256), 256, 0, st>>>(T, ix->d_txm_off, ix->d_txm_cid, s->d_adj, s->d_amodel, d_act, d_in_model,
                                                                          s->d_R, s->d_iE, s->d_A, s->d_Rs, d_deg, s->d_lone, d_rflag);
    LAUNCHED(ctx);
    k_order_gather<<<(unsigned)((T + 1 + 255) / 256), 256, 0, st>>>(T, ix->d_order, d_rflag, d_degn);      // d_degn / d_degp: free until k_nat_fill
    CU(cub::DeviceScan::ExclusiveSum(d_cub, cub_bytes, d_degn, d_degp, T + 1, st));
    k_order_scatter<<<(unsigned)((T + 1 + 255) / 256), 256, 0, st>>>(T, ix->d_order, d_degp, d_nat);
    ctx->launches += 3;
    CU(cudaGetLastError());
    uint32_t P32 = 0; int32_t C_a32 = 0;
    CU(cudaMemcpyAsync(&P32, d_nat + T, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&C_a32, d_newid + nm, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const int32_t P = (int32_t)P32;
    const int64_t C_a = C_a32;
    // ---- state: theta | q in one allocation ----
    size_t theta_bytes = (((size_t)(P > 0 ? P : 1) * 8 + 255) / 256) * 256;
    size_t q_bytes = (((size_t)(C_a > 0 ? C_a : 1) * 8 + 255) / 256) * 256;
    if (theta_bytes + q_bytes > s->state_bytes) {
        if (s->d_state) dev_free(s->d_state);
        s->d_state = nullptr;
        double *p = nullptr;
        TRY(dev_alloc(&p, (theta_bytes + q_bytes) / 8));
        s->d_state = p;
        s->state_bytes = theta_bytes + q_bytes;
    }
    {   // tagged slots of the barrier-free kernel
        const size_t need = 16 * ((size_t)P + 1 + (size_t)C_a + 1);
        if (need > s->slots_bytes) {
            if (s->d_slots) dev_free(s->d_slots);
            s->d_slots = nullptr;
            char *ps = nullptr;
            TRY(dev_alloc(&ps, need + (need >> 3)));
            s->d_slots = ps;
            s->slots_bytes = need + (need >> 3);
            CU(cudaMemsetAsync(s->d_slots, 0, s->slots_bytes, st));
            s->slot_tag = 0;
        }
    }
    EmModel &m = s->m;
    memset(&m, 0, sizeof(m));
    m.T = T; m.P = P; m.B = B; m.C_a = C_a; m.smem_bytes = ctx->em_smem_bytes;
    m.theta = s->d_state;
    m.q = (double *)((char *)s->d_state + theta_bytes);
    // ---- the class-owner-centric model (k_em_psum) whenever the sample is eligible ----
    s->use_psum = false;
    if (want_psum) {
        PsPrepIn pin;
        pin.T = T; pin.P = P; pin.nm = nm; pin.C_a = C_a; pin.n_kseg = n_kseg; pin.d_cub = d_cub; pin.cub_bytes = cub_bytes;
        pin.nranks = s->sharded ? ctx->nranks : 1; pin.rank = s->sharded ? ctx->rank : 0; pin.B_local = B; pin.B = B * pin.nranks;
        pin.d_act = d_act; pin.d_newid = d_newid; pin.d_deg = d_deg; pin.d_rflag = d_rflag; pin.d_nat = d_nat; pin.d_pos = d_pos;
        TRY(sample_build_psum(s, pin));
        if (s->use_psum) {
            CU(cudaEventRecord(e1, st));
            CU(cudaStreamSynchronize(st));
            float ms = 0;
            CU(cudaEventElapsedTime(&ms, e0, e1));
            s->prep_ms = ms;
            s->prepared = true;
            s->n_iter = 0; s->final_delta = INFINITY; s->em_ms = 0;
            return EMSAR_OK;
        }
        if (s->sharded) {            // not eligible: the legacy sharded kernel cuts the classes by range, which has to happen before the row statistics
            s->force_legacy = true;
            return emsar_sample_prepare(s, opts_in);
        }
    } else ps_free(s);
    // ---- arena part 1: everything whose size is known now ----
    // padding: a class is padded to steps*G members (< 1.25 k + 31), a cell to whole row blocks (< 32*steps ints per (CTA, cardinality) cell)
    const size_t e_ints_max = (size_t)ix->nnz_multi + (size_t)ix->nnz_multi / 4 + 32 * (size_t)nm + (size_t)n_cells * 32 * 64 + 64;
    const size_t e_tiles_max = (size_t)C_a + (size_t)n_cells + 1;
    const size_t m_items_max = (size_t)P / 32 + 2 * (size_t)B + std::min<size_t>((size_t)P, (size_t)ix->nnz_multi / M_LONG + 1) + 64;   // slices + long rows
    auto rnd = [](size_t b) { return ((b + 255) / 256) * 256; };
    size_t arena1 = rnd(e_ints_max * 4) + rnd((size_t)(C_a + 1) * 4) + rnd(e_tiles_max * 16) + 2 * rnd((size_t)(P + 1) * 16) + rnd((size_t)(P + 1) * 4) + rnd(m_items_max * 16) + 12 * rnd((size_t)(B + 1) * 4);
    if (arena1 > s->pack_bytes) {
        if (s->d_pack) dev_free(s->d_pack);
        s->d_pack = nullptr;
        char *p = nullptr;
        TRY(dev_alloc(&p, arena1));
        s->d_pack = p;
        s->pack_bytes = arena1;
    }
    char *ac = (char *)s->d_pack;
    m.e_tid = arena_take<int32_t>(ac, e_ints_max);
    m.e_R = arena_take<uint32_t>(ac, (size_t)C_a + 1);
    m.e_tiles = arena_take<int4>(ac, e_tiles_max);
    m.row_RsA = arena_take<double2>(ac, (size_t)P + 1);
    m.row_n = arena_take<int32_t>(ac, (size_t)P + 1);
    m.rsa_nat = arena_take<double2>(ac, (size_t)P + 1);
    m.m_items = arena_take<int4>(ac, m_items_max);
    m.blk_row0 = arena_take<int32_t>(ac, (size_t)B + 1);
    m.blk_cls0 = arena_take<int32_t>(ac, (size_t)B + 1);
    m.blk_etile0 = arena_take<int32_t>(ac, (size_t)B + 1);
    m.blk_mitem0 = arena_take<int32_t>(ac, (size_t)B + 1);
    m.blk_nres = arena_take<int32_t>(ac, (size_t)B + 1);
    m.blk_hr0 = arena_take<int32_t>(ac, (size_t)B + 1);
    m.blk_hc0 = arena_take<int32_t>(ac, (size_t)B + 1);
    m.blk_nhr = arena_take<int32_t>(ac, (size_t)B + 1);
    m.blk_nhc = arena_take<int32_t>(ac, (size_t)B + 1);
    // ---- rows: costs, ownership ranges, length-sorted order inside each CTA ----
    // cost of a row = its M entries + the members of the classes it owns (one gather each) + fixed work per row / per class,
    // in units of one gather (tuning knobs: EMSAR_COST_ROW / EMSAR_COST_CLASS)
    const int cost_row = getenv("EMSAR_COST_ROW") ? atoi(getenv("EMSAR_COST_ROW")) : 2;
    const int cost_class = getenv("EMSAR_COST_CLASS") ? atoi(getenv("EMSAR_COST_CLASS")) : 0;
    const int owner_rule = (getenv("EMSAR_OWNER") && !strcmp(getenv("EMSAR_OWNER"), "light")) ? 1 : 0;
    k_nat_fill<<<(unsigned)((T + 255) / 256), 256, 0, st>>>(T, d_rflag, d_nat, d_deg, d_pos, d_degn, d_tn, d_ecost, P);
    LAUNCHED(ctx);
    if (nm > 0) {
        k_class_owner<<<(unsigned)((nm + 255) / 256), 256, 0, st>>>(nm, T, ix->d_cls_off, ix->d_cls_tid, d_act, d_deg, owner_rule, d_owner);
        k_class_cost<<<(unsigned)((nm + 255) / 256), 256, 0, st>>>(nm, T, ix->d_cls_off, d_owner, d_act, d_nat, d_ecost, cost_class, 1);
        ctx->launches += 2;
    }
    k_row_cost<<<(unsigned)((P + 1 + 255) / 256), 256, 0, st>>>(P, d_degn, d_ecost, d_cost, cost_row, 1);
    LAUNCHED(ctx);
    CU(cub::DeviceScan::ExclusiveSum(d_cub, cub_bytes, d_cost, d_costp, P + 1, st));
    LAUNCHED(ctx);
    k_block_bounds<<<(unsigned)((B + 1 + 255) / 256), 256, 0, st>>>(B, P, d_costp, m.blk_row0);
    LAUNCHED(ctx);
    if (P > 0) {
        k_sort_keys<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(P, B, m.blk_row0, d_degn, d_key, d_val);
        LAUNCHED(ctx);
        CU(cub::DeviceRadixSort::SortPairs(d_cub, cub_bytes, d_key, d_key2, d_val, d_perm, P, 0, 44, st));
        ctx->launches += 4;
        k_apply_perm<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(P, d_perm, d_tn, d_degn, s->d_Rs, s->d_A, d_pos, d_degp, m.row_RsA, m.row_n, m.rsa_nat);
        LAUNCHED(ctx);
    }
    k_block_items_count<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(B, m.blk_row0, d_degp, d_nlong, d_nitems, d_ngroups);
    k_block_items_prefix<<<1, 32, 0, st>>>(B, d_nitems, m.blk_mitem0);
    ctx->launches += 2;
    // ---- classes: (owner CTA, cardinality) cells ----
    CU(cudaMemsetAsync(d_cell_cnt, 0, (size_t)(n_cells + 1) * 4, st));
    k_fill_int<<<(unsigned)((n_cells + 1 + 255) / 256), 256, 0, st>>>(d_cell_first, n_cells + 1, 0x7fffffff);
    LAUNCHED(ctx);
    if (nm > 0 && n_kseg > 0) {
        k_class_cells<<<(unsigned)((nm + 255) / 256), 256, 0, st>>>(nm, T, n_kseg, B, ix->d_kseg_cid0, d_owner, d_act, d_newid,
                                                                   d_pos, m.blk_row0, d_cellof, d_cell_cnt, d_hkeys, (int32_t *)d_uniq_e);
        LAUNCHED(ctx);
    }
    k_cell_sizes<<<(unsigned)((n_cells + 1 + 255) / 256), 256, 0, st>>>(n_cells, n_kseg > 0 ? n_kseg : 1, ix->d_kseg_k, d_cell_cnt, d_cell_ints, d_cell_tiles);
    LAUNCHED(ctx);
    CU(cub::DeviceScan::ExclusiveSum(d_cub, cub_bytes, d_cell_cnt, d_clsbase, n_cells + 1, st));
    CU(cub::DeviceScan::ExclusiveSum(d_cub, cub_bytes, d_cell_ints, d_intbase, n_cells + 1, st));
    CU(cub::DeviceScan::ExclusiveSum(d_cub, cub_bytes, d_cell_tiles, d_tilebase, n_cells + 1, st));
    ctx->launches += 3;
    k_block_tables<<<(unsigned)((B + 1 + 255) / 256), 256, 0, st>>>(B, n_kseg > 0 ? n_kseg : 1, d_clsbase, d_tilebase, m.blk_cls0, m.blk_etile0);
    LAUNCHED(ctx);
    CU(cudaGetLastError());
    uint32_t e_ints = 0;
    int32_t n_etiles = 0, n_mitems = 0;
    CU(cudaMemcpyAsync(&e_ints, d_intbase + n_cells, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&n_etiles, d_tilebase + n_cells, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&n_mitems, m.blk_mitem0 + B, 4, cudaMemcpyDeviceToHost, st));
    // ---- halo lists: distinct remote rows (E side) and remote classes (M side) per CTA ----
    unsigned int h_cnt[4] = {0, 0, 0, 0};    // collected e, unique e, collected m, unique m
    CU(cudaMemsetAsync(d_hcount, 0, 32, st));
    if (nm > 0 && n_kseg > 0) {
        if (C_a > 0) {
            // new compact ids = rank in (cell, old id) order
            CU(cub::DeviceRadixSort::SortPairs(d_cub, cub_bytes, d_hkeys, d_hsort, (int32_t *)d_uniq_e, (int32_t *)d_uniq_m, (int)C_a, 0, 64, st));
            k_class_newid<<<(unsigned)((C_a + 255) / 256), 256, 0, st>>>(C_a, (const int32_t *)d_uniq_m, d_newid2);
            ctx->launches += 4;
        }
        k_halo_collect_e<<<(unsigned)((nm * 32 + 255) / 256), 256, 0, st>>>(nm, T, B, ix->d_cls_off, ix->d_cls_tid, d_owner, d_act, d_pos, m.blk_row0, d_hkeys, d_hcount);
        LAUNCHED(ctx);
    }
    CU(cudaMemcpyAsync(&h_cnt[0], d_hcount, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (h_cnt[0] > 0) {
        CU(cub::DeviceRadixSort::SortKeys(d_cub, cub_bytes, d_hkeys, d_hsort, (int)h_cnt[0], 0, 44, st));
        CU(cub::DeviceSelect::Unique(d_cub, cub_bytes, d_hsort, d_uniq_e, d_hcount + 1, (int)h_cnt[0], st));
        ctx->launches += 4;
    }
    if (nm > 0 && n_kseg > 0) {
        k_halo_collect_m<<<(unsigned)(((int64_t)T * 32 + 255) / 256), 256, 0, st>>>(T, B, ix->d_txm_off, ix->d_txm_cid, d_act, d_newid2, d_pos, m.blk_row0,
                                                                                   m.blk_cls0, d_hkeys, d_hcount + 2);
        LAUNCHED(ctx);
    }
    CU(cudaMemcpyAsync(&h_cnt[1], d_hcount + 1, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&h_cnt[2], d_hcount + 2, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (h_cnt[2] > 0) {
        CU(cub::DeviceRadixSort::SortKeys(d_cub, cub_bytes, d_hkeys, d_hsort, (int)h_cnt[2], 0, 44, st));
        CU(cub::DeviceSelect::Unique(d_cub, cub_bytes, d_hsort, d_uniq_m, d_hcount + 3, (int)h_cnt[2], st));
        ctx->launches += 4;
        CU(cudaMemcpyAsync(&h_cnt[3], d_hcount + 3, 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    const unsigned int n_ue = h_cnt[0] ? h_cnt[1] : 0, n_um = h_cnt[2] ? h_cnt[3] : 0;
    if ((size_t)(n_ue + n_um + 2) * 4 > s->halo_bytes) {
        if (s->d_halo) dev_free(s->d_halo);
        s->d_halo = nullptr;
        int32_t *p = nullptr;
        TRY(dev_alloc(&p, (size_t)n_ue + n_um + 64));
        s->d_halo = p;
        s->halo_bytes = ((size_t)n_ue + n_um + 64) * 4;
    }
    m.halo_rows = s->d_halo;
    m.halo_cls = s->d_halo + n_ue;
    k_halo_ranges<<<(unsigned)((B + 1 + 255) / 256), 256, 0, st>>>(B, n_ue, d_uniq_e, m.blk_hr0);
    k_halo_ranges<<<(unsigned)((B + 1 + 255) / 256), 256, 0, st>>>(B, n_um, d_uniq_m, m.blk_hc0);
    ctx->launches += 2;
    if (n_ue) { k_halo_list<<<(n_ue + 255) / 256, 256, 0, st>>>(n_ue, d_uniq_e, m.halo_rows); LAUNCHED(ctx); }
    if (n_um) { k_halo_list<<<(n_um + 255) / 256, 256, 0, st>>>(n_um, d_uniq_m, m.halo_cls); LAUNCHED(ctx); }
    CU(cudaGetLastError());
    if ((size_t)e_ints > e_ints_max || (size_t)n_etiles > e_tiles_max || (size_t)n_mitems > m_items_max) {
        emsar_set_err("internal: packed model exceeds its bounds (%u ints, %d tiles, %d items)", e_ints, n_etiles, n_mitems);
        return EMSAR_ERR_STATE;
    }
    // ---- M items: sizes -> offsets; E tile descriptors ----
    if ((size_t)n_mitems + 1 > 2 * (size_t)T + B + 64) { emsar_set_err("internal: item scratch too small"); return EMSAR_ERR_STATE; }
    k_block_items_fill<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(B, m.blk_row0, d_degp, d_nlong, m.blk_mitem0, d_isize, m.m_items, d_rowbase, d_rowitem);
    LAUNCHED(ctx);
    CU(cub::DeviceScan::ExclusiveSum(d_cub, cub_bytes, d_isize, d_ioff, n_mitems + 1, st));
    LAUNCHED(ctx);
    if (n_mitems > 0) { k_item_offsets<<<(unsigned)((n_mitems + 255) / 256), 256, 0, st>>>(n_mitems, d_ioff, m.m_items); LAUNCHED(ctx); }
    if (n_etiles > 0) {
        k_etiles<<<(unsigned)((n_etiles + 255) / 256), 256, 0, st>>>(n_etiles, n_cells, n_kseg, ix->d_kseg_k, d_tilebase, d_clsbase, d_cell_cnt, d_intbase, m.e_tiles);
        LAUNCHED(ctx);
    }
    // ---- host: cut each CTA's two index streams into staged chunks and plan its shared memory ----
    std::vector<int4> h_et((size_t)n_etiles), h_mi((size_t)n_mitems);
    std::vector<int32_t> h_row0(B + 1), h_cls0(B + 1), h_et0(B + 1), h_mi0(B + 1), h_hr0(B + 1), h_hc0(B + 1);
    uint32_t m_ints = 0;
    CU(cudaMemcpyAsync(&m_ints, d_ioff + n_mitems, 4, cudaMemcpyDeviceToHost, st));
    if (n_etiles) CU(cudaMemcpyAsync(h_et.data(), m.e_tiles, (size_t)n_etiles * 16, cudaMemcpyDeviceToHost, st));
    if (n_mitems) CU(cudaMemcpyAsync(h_mi.data(), m.m_items, (size_t)n_mitems * 16, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_row0.data(), m.blk_row0, (size_t)(B + 1) * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_cls0.data(), m.blk_cls0, (size_t)(B + 1) * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_et0.data(), m.blk_etile0, (size_t)(B + 1) * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_mi0.data(), m.blk_mitem0, (size_t)(B + 1) * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_hr0.data(), m.blk_hr0, (size_t)(B + 1) * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_hc0.data(), m.blk_hc0, (size_t)(B + 1) * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    std::vector<int4> h_ech, h_mch;
    std::vector<int32_t> h_ech0(B + 1), h_mch0(B + 1), h_nhr(B + 1, 0), h_nres(B + 1, 0), h_nhc(B + 1, 0);
    const int lim = CH_INTS - 16;
    for (int b = 0; b < B; b++) {
        h_ech0[b] = (int32_t)h_ech.size();
        {   // E: tiles in order; a chunk carries the tiles' member indices followed by their read counts
            int i0 = h_et0[b], cur = 0;
            for (int i = h_et0[b]; i < h_et0[b + 1]; i++) {
                const int4 t = h_et[(size_t)i];
                const int steps = t.w & 0xfff, cpb = 32 >> ((t.w >> 12) & 0xf);
                const int ints = ((t.y + cpb - 1) / cpb) * 32 * steps;
                const int cost = ints + t.y;
                if (cost > lim) {            // oversized tile: its own, unstaged chunk
                    if (i > i0) h_ech.push_back(make_int4(i0 - h_et0[b], i - h_et0[b], h_et[(size_t)i0].z, cur));
                    h_ech.push_back(make_int4(i - h_et0[b], i + 1 - h_et0[b], t.z, -1));
                    i0 = i + 1; cur = 0;
                    continue;
                }
                if (cur + cost > lim) { h_ech.push_back(make_int4(i0 - h_et0[b], i - h_et0[b], h_et[(size_t)i0].z, cur)); i0 = i; cur = 0; }
                cur += cost;
            }
            if (h_et0[b + 1] > i0) h_ech.push_back(make_int4(i0 - h_et0[b], h_et0[b + 1] - h_et0[b], h_et[(size_t)i0].z, cur));
        }
        h_mch0[b] = (int32_t)h_mch.size();
        {
            int i0 = h_mi0[b], cur = 0;
            for (int i = h_mi0[b]; i < h_mi0[b + 1]; i++) {
                const int4 t = h_mi[(size_t)i];
                const int len = t.w & 0x3fffffff;
                const int cost = (t.w >> 30) == 0 ? 32 * len : len + t.y;
                if (cost > lim) {
                    if (i > i0) h_mch.push_back(make_int4(i0 - h_mi0[b], i - h_mi0[b], h_mi[(size_t)i0].z, cur));
                    h_mch.push_back(make_int4(i - h_mi0[b], i + 1 - h_mi0[b], t.z, -1));
                    i0 = i + 1; cur = 0;
                    continue;
                }
                if (cur + cost > lim) { h_mch.push_back(make_int4(i0 - h_mi0[b], i - h_mi0[b], h_mi[(size_t)i0].z, cur)); i0 = i; cur = 0; }
                cur += cost;
            }
            if (h_mi0[b + 1] > i0) h_mch.push_back(make_int4(i0 - h_mi0[b], h_mi0[b + 1] - h_mi0[b], h_mi[(size_t)i0].z, cur));
        }
    }
    h_ech0[B] = (int32_t)h_ech.size(); h_mch0[B] = (int32_t)h_mch.size();
    if (getenv("EMSAR_DEBUG_PLAN")) {      // tuning aid: what every CTA owns
        for (int b = 0; b < B; b++) {
            int groups = 0, unst_m = 0, unst_e = 0; long long gent = 0, sent = 0;
            for (int i = h_mi0[b]; i < h_mi0[b + 1]; i++) { const int4 t = h_mi[(size_t)i]; if (t.w >> 30) { groups++; gent += t.w & 0x3fffffff; } else sent += 32LL * (t.w & 0x3fffffff); }
            for (int i = h_mch0[b]; i < h_mch0[b + 1]; i++) unst_m += h_mch[(size_t)i].w < 0;
            for (int i = h_ech0[b]; i < h_ech0[b + 1]; i++) unst_e += h_ech[(size_t)i].w < 0;
            fprintf(stderr, "cta %3d rows %5d cls %6d etiles %4d echunks %3d (unstaged %d) mitems %4d groups %4d group_ent %7lld slice_ent %7lld mchunks %3d (unstaged %d)\n", b,
                    h_row0[b + 1] - h_row0[b], h_cls0[b + 1] - h_cls0[b], h_et0[b + 1] - h_et0[b], h_ech0[b + 1] - h_ech0[b], unst_e, h_mi0[b + 1] - h_mi0[b], groups,
                    gent, sent, h_mch0[b + 1] - h_mch0[b], unst_m);
        }
    }
    const bool direct = !(getenv("EMSAR_EM_MODE") && !strcmp(getenv("EMSAR_EM_MODE"), "pipe"));
    std::vector<int32_t> h_eres((size_t)n_etiles + 1, -1), h_mres((size_t)n_mitems + 1, -1), h_resints(B + 1, 0);
    int32_t grid_all_local = 1;
    int64_t resident_ints = 0, rows_long = 0;
    for (int i = 0; i < n_mitems; i++) if ((h_mi[(size_t)i].w >> 30) == 1) rows_long += h_mi[(size_t)i].y;
    for (int b = 0; b < B; b++) {
        // what of a CTA's state gets a shared-memory slot: its rows (always), then halo rows, then its classes, then halo classes
        const int nrows = h_row0[b + 1] - h_row0[b];
        const int n_et = h_et0[b + 1] - h_et0[b], n_mi = h_mi0[b + 1] - h_mi0[b];
        const int n_ech = direct ? 0 : h_ech0[b + 1] - h_ech0[b], n_mch = direct ? 0 : h_mch0[b + 1] - h_mch0[b];
        int left = ctx->em_smem_bytes - em_smem_plan(direct ? 0 : NSTAGE * CH_BYTES, n_et, n_mi, n_ech, n_mch, direct ? n_et + n_mi : 0, nrows, 0, 0, 0).total - 64;
        if (left < 0) { emsar_set_err("a CTA's rows and tables do not fit in shared memory (%d rows)", nrows); return EMSAR_ERR_UNSUPPORTED; }
        int a = std::max(0, std::min(h_hr0[b + 1] - h_hr0[b], left / 12)); left -= a * 12;
        int r = std::max(0, std::min(h_cls0[b + 1] - h_cls0[b], left / 8)); left -= r * 8;
        int c = std::max(0, std::min(h_hc0[b + 1] - h_hc0[b], left / 12)); left -= c * 12;
        h_nhr[b] = a; h_nres[b] = r; h_nhc[b] = c;
        if (a != h_hr0[b + 1] - h_hr0[b] || r != h_cls0[b + 1] - h_cls0[b] || c != h_hc0[b + 1] - h_hc0[b]) grid_all_local = 0;
        if (direct && left > 64) {
            // resident index cache: the index data of the items with the longest dependent chains stays in shared memory for
            // the whole kernel (it never changes); everything else is read from L2 every iteration
            struct Cand { int steps, ints, idx; bool e; };
            std::vector<Cand> cand;
            for (int i = h_et0[b]; i < h_et0[b + 1]; i++) {
                const int4 t = h_et[(size_t)i];
                const int steps = t.w & 0xfff, cpb = 32 >> ((t.w >> 12) & 0xf);
                const int ints = ((t.y + cpb - 1) / cpb) * 32 * steps + t.y;                    // members + read counts
                cand.push_back({t.y > 32 ? 2 : steps, ints, i, true});
            }
            for (int i = h_mi0[b]; i < h_mi0[b + 1]; i++) {
                const int4 t = h_mi[(size_t)i];
                const int len = t.w & 0x3fffffff;
                if ((t.w >> 30) == 0) cand.push_back({len, 32 * len, i, false});
                else cand.push_back({t.y * 6 + len / 32, len + t.y, i, false});
            }
            std::stable_sort(cand.begin(), cand.end(), [](const Cand &x, const Cand &y) { return x.steps > y.steps; });
            int used = 0;
            const int cap_ints = (left - 64) / 4;
            for (const Cand &cd : cand) {
                const int need = (cd.ints + 3) & ~3;
                if (used + need > cap_ints) continue;
                (cd.e ? h_eres : h_mres)[(size_t)cd.idx] = used;
                used += need;
            }
            h_resints[b] = used;
            resident_ints += used;
        }
    }
    {
        const size_t nb = (h_ech.size() + h_mch.size() + 3) * 16 + 4 * (size_t)(B + 1) * 4 + ((size_t)n_etiles + n_mitems + 2) * 4 + 2048;
        if (nb > s->chunk_bytes) {
            if (s->d_chunks) dev_free(s->d_chunks);
            s->d_chunks = nullptr;
            char *p = nullptr;
            TRY(dev_alloc(&p, nb + (nb >> 2)));
            s->d_chunks = p;
            s->chunk_bytes = nb + (nb >> 2);
        }
        char *cc = (char *)s->d_chunks;
        m.e_chunks = arena_take<int4>(cc, h_ech.size() + 1);
        m.m_chunks = arena_take<int4>(cc, h_mch.size() + 1);
        m.blk_ech0 = arena_take<int32_t>(cc, (size_t)B + 1);
        m.blk_mch0 = arena_take<int32_t>(cc, (size_t)B + 1);
        m.blk_res_ints = arena_take<int32_t>(cc, (size_t)B + 1);
        m.e_res = arena_take<int32_t>(cc, (size_t)n_etiles + 1);
        m.m_res = arena_take<int32_t>(cc, (size_t)n_mitems + 1);
        m.direct = direct ? 1 : 0;
        m.all_local = grid_all_local;
        CU(cudaMemcpyAsync(m.blk_res_ints, h_resints.data(), (size_t)(B + 1) * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(m.e_res, h_eres.data(), ((size_t)n_etiles + 1) * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(m.m_res, h_mres.data(), ((size_t)n_mitems + 1) * 4, cudaMemcpyHostToDevice, st));
        if (!h_ech.empty()) CU(cudaMemcpyAsync(m.e_chunks, h_ech.data(), h_ech.size() * 16, cudaMemcpyHostToDevice, st));
        if (!h_mch.empty()) CU(cudaMemcpyAsync(m.m_chunks, h_mch.data(), h_mch.size() * 16, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(m.blk_ech0, h_ech0.data(), (size_t)(B + 1) * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(m.blk_mch0, h_mch0.data(), (size_t)(B + 1) * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(m.blk_nhr, h_nhr.data(), (size_t)(B + 1) * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(m.blk_nres, h_nres.data(), (size_t)(B + 1) * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(m.blk_nhc, h_nhc.data(), (size_t)(B + 1) * 4, cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));     // the host vectors go out of scope below
    }
    if ((size_t)(m_ints + 1) * 4 > s->mcls_bytes) {
        if (s->d_mcls) dev_free(s->d_mcls);
        s->d_mcls = nullptr;
        int32_t *p = nullptr;
        TRY(dev_alloc(&p, (size_t)m_ints + (m_ints >> 3) + 64));
        s->d_mcls = p;
        s->mcls_bytes = ((size_t)m_ints + (m_ints >> 3) + 64) * 4;
    }
    m.m_cls = s->d_mcls;
    m.n_etiles = n_etiles; m.n_mitems = n_mitems; m.m_ints = m_ints;
    k_item_finish<<<(unsigned)((n_mitems + 255) / 256 + 1), 256, 0, st>>>(n_mitems, B, m.blk_mitem0, m.blk_row0, m.blk_nres, d_degp, m.m_items, m.m_cls);
    LAUNCHED(ctx);
    CU(cudaMemsetAsync(m.e_tid, 0, e_ints_max * 4, st));
    if (nm > 0 && n_kseg > 0) {
        k_pack_classes<<<(unsigned)((nm * 32 + 255) / 256), 256, 0, st>>>(nm, T, n_kseg, m.blk_nres, m.blk_hr0, m.blk_nhr, d_uniq_e, ix->d_kseg_k, ix->d_cls_off,
                                                                          ix->d_cls_tid, d_act, d_newid, d_cellof, d_newid2, d_clsbase, d_intbase, s->d_R,
                                                                          d_pos, m.blk_row0, m.blk_cls0, m.e_tid, m.e_R);
        LAUNCHED(ctx);
    }
    k_scatter_rows<<<(unsigned)(((int64_t)T * 32 + 255) / 256), 256, 0, st>>>(T, B, m.blk_nres, m.blk_hc0, m.blk_nhc, d_uniq_m, ix->d_txm_off, ix->d_txm_cid, d_act, d_newid2, d_pos, m.blk_row0,
                                                                             m.blk_cls0, d_nlong, m.blk_mitem0, d_rowbase, d_rowitem, d_ngroups, m.m_items, m.m_cls);
    LAUNCHED(ctx);
    // start point: theta = 1 for every row that takes part (A_t > 0)
    if (P > 0) { k_fill_double<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(m.theta, P, 1.0); LAUNCHED(ctx); }
    CU(cudaGetLastError());
    uint32_t nnz_a = 0;
    {   // nnz_a = sum of the active row lengths
        CU(cub::DeviceScan::ExclusiveSum(d_cub, cub_bytes, d_degn, d_cost, P + 1, st));
        LAUNCHED(ctx);
        CU(cudaMemcpyAsync(&nnz_a, d_cost + P, 4, cudaMemcpyDeviceToHost, st));
    }
    CU(cudaEventRecord(e1, st));
    CU(cudaStreamSynchronize(st));
    m.nnz_a = nnz_a;
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    s->prep_ms = ms;
    emsar_model_stats &ms_ = s->stats;
    memset(&ms_, 0, sizeof(ms_));
    ms_.T = T; ms_.C_a = C_a; ms_.nnz_a = m.nnz_a;
    ms_.rows_short = P; ms_.rows_long = rows_long; ms_.rows_hub = 0; ms_.rows_fixed = T - P;
    {
        const char *em_mode = getenv("EMSAR_EM_MODE");
        const bool want_barrier = em_mode && !strcmp(em_mode, "barrier");
        ms_.em_variant = s->sharded ? 2 : (m.direct && m.all_local && s->d_slots && !want_barrier) ? 3 : m.direct ? 1 : 0;      // as em_launch decides
    }
    ms_.all_local = m.all_local; ms_.halo_rows = n_ue; ms_.halo_classes = n_um;
    ms_.resident_index_bytes = 4 * resident_ints;
    ms_.index_bytes = 4 * ((int64_t)e_ints + C_a + (int64_t)m_ints);
    // fused class-sharded kernel: partial row sums pushed to the slice owners + the new theta of this rank's slice pushed to every rank (16-byte slots)
    ms_.peer_bytes_per_iter = s->sharded ? (int64_t)(32.0 * P * (ctx->nranks - 1) / ctx->nranks) : 0;
    ms_.e_tiles = n_etiles; ms_.m_tiles = n_mitems;
    ms_.bytes_per_iter = 8 * m.nnz_a + 24 * C_a + 44 * (int64_t)T;
    // what the kernel streams per iteration: E: encoded members (padded) + R (+ q to global for halo classes);
    // M: encoded classes (padded) + {Rs,A} + theta write-through; tile descriptors live in shared memory
    ms_.stream_bytes_per_iter = 4 * (int64_t)e_ints + 4 * C_a + 4 * (int64_t)m_ints + 16 * (int64_t)P + 8 * (int64_t)P;
    s->prepared = true;
    s->n_iter = 0; s->final_delta = INFINITY; s->em_ms = 0;
    return EMSAR_OK;
}

extern "C" int emsar_sample_model_stats(emsar_sample *s, emsar_model_stats *st)
{
    CHECK_ARG(s && st, "emsar_sample_model_stats: NULL argument");
    if (!s->prepared) { emsar_set_err("emsar_sample_model_stats: sample not prepared"); return EMSAR_ERR_STATE; }
    *st = s->stats;
    return EMSAR_OK;
}

extern "C" int emsar_sample_wf_get(emsar_sample *s, double *Wf)
{
    CHECK_ARG(s && Wf, "emsar_sample_wf_get: NULL argument");
    if (!s->prepared) { emsar_set_err("emsar_sample_wf_get: sample not prepared"); return EMSAR_ERR_STATE; }
    TRY(ctx_use(s->ctx));
    CU(cudaMemcpyAsync(Wf, s->d_Wf, (size_t)s->index->nF * 8, cudaMemcpyDeviceToHost, s->ctx->stream));
    CU(cudaStreamSynchronize(s->ctx->stream));
    return EMSAR_OK;
}

// set ids for the -g output (computed lazily when the cut loop was not needed)
int sample_ensure_sets(emsar_sample *s)
{
    if (s->have_CS) return EMSAR_OK;
    emsar_index *ix = s->index;
    TRY(ctx_use(s->ctx));
    double cut = s->eumacut;
    const int cap = s->opts.max_ntid_per_sid > 0 ? s->opts.max_ntid_per_sid : 5000;
    if (getenv("EMSAR_SETS_HOST")) {
        std::vector<double> adj((size_t)ix->C);
        CU(cudaMemcpyAsync(adj.data(), s->d_adj, (size_t)ix->C * 8, cudaMemcpyDeviceToHost, s->ctx->stream));
        CU(cudaStreamSynchronize(s->ctx->stream));
        compute_sets(ix, adj, &cut, cap, &s->h_CS, &s->max_sid);
        s->have_CS = true;
        return EMSAR_OK;
    }
    int32_t *d_CS = nullptr; uint8_t *d_im = nullptr;
    TRY(dev_alloc(&d_CS, (size_t)ix->C)); TRY(dev_alloc(&d_im, (size_t)ix->C));
    int rc = device_sets(s, &cut, cap, d_CS, d_im, &s->max_sid);
    if (rc == EMSAR_OK) {
        s->h_CS.resize((size_t)ix->C);
        rc = cudaMemcpyAsync(s->h_CS.data(), d_CS, (size_t)ix->C * 4, cudaMemcpyDeviceToHost, s->ctx->stream) == cudaSuccess && cudaStreamSynchronize(s->ctx->stream) == cudaSuccess ? EMSAR_OK : EMSAR_ERR_CUDA;
        s->have_CS = rc == EMSAR_OK;
    }
    dev_free(d_CS); dev_free(d_im);
    return rc;
}
