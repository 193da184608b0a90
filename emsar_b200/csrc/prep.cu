// Per-sample model build on the device (one-off per alignment file):
//   Wf (transfer_fraglendist_to_Wf, reference emsar_functions.c:2503-2513), adjEUMA (compute_adjEUMA :2517-2523 as
//   scan_rshbucket :2135-2192 applies it), EUMAps (construct_EUMAps :3148-3154), iEUMA (compute_iEUMA :3218-3232),
//   the set decomposition with EUMAcut (build_TC_from_CT_2 :2201-2227, propagate_2 :2234-2259, emsar_main.c:411-425),
//   and the packed ACTIVE model the EM kernel streams: classes that are modelled (in a set, EUMAps > 0) and hold
//   reads, binned by cardinality, plus its transposed CSR with rows binned by length.
// The deterministic fp64 pre-steps use the reference's summation order and no FMA contraction (-fmad=false), so
// Wf, adjEUMA, EUMAps and iEUMA are bit-identical to the CPU reference.
#include <algorithm>
#include <cub/cub.cuh>
#include <numeric>

#include "common.cuh"

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
// multiplicity, sequential order), A_t (modelled classes), Rs_t, active degree, lone-singleton flag, row class.
__global__ void k_row_stats(int32_t T, const uint32_t *__restrict__ txm_off, const int32_t *__restrict__ txm_cid,
                            const double *__restrict__ adj, const double *__restrict__ amodel, const int32_t *__restrict__ act,
                            const uint8_t *__restrict__ in_model, const int32_t *__restrict__ R,
                            double *__restrict__ iE, double *__restrict__ A, double *__restrict__ Rs, int32_t *__restrict__ deg,
                            uint8_t *__restrict__ lone, unsigned long long *__restrict__ rkey)
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
        int rc = !(acc_a > 0) ? 3 : (d <= M_SHORT_MAX ? 0 : (d < M_HUB_MIN ? 1 : 2));
        rkey[t] = rc < 3 ? (1ULL << (21 * rc)) : 0ULL;
        if (t == T - 1) rkey[T] = 0ULL;
    }
}

__global__ void k_row_perm(int32_t T, const unsigned long long *__restrict__ rpre, const unsigned long long *__restrict__ rkey,
                           const int32_t *__restrict__ deg, const double *__restrict__ Rs, const double *__restrict__ A,
                           int32_t *__restrict__ pos, int32_t *__restrict__ row_t, uint32_t *__restrict__ degp, double2 *__restrict__ row_RsA)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const unsigned long long tot = rpre[T];
    const int n0 = (int)(tot & 0x1FFFFF), n1 = (int)((tot >> 21) & 0x1FFFFF), n2 = (int)((tot >> 42) & 0x1FFFFF);
    unsigned long long key = rkey[t], pre = rpre[t];
    int p = -1;
    if (key == 1ULL) p = (int)(pre & 0x1FFFFF);
    else if (key == (1ULL << 21)) p = n0 + (int)((pre >> 21) & 0x1FFFFF);
    else if (key == (1ULL << 42)) p = n0 + n1 + (int)((pre >> 42) & 0x1FFFFF);
    pos[t] = p;
    if (p >= 0) { row_t[p] = t; degp[p] = (uint32_t)deg[t]; row_RsA[p] = make_double2(Rs[t], A[t]); }
    if (t == 0) degp[n0 + n1 + n2] = 0;
}

// One warp per transcript: copy the ACTIVE entries of its transposed row, in order, as compact class ids.
__global__ void k_scatter_rows(int32_t T, const uint32_t *__restrict__ txm_off, const int32_t *__restrict__ txm_cid,
                               const int32_t *__restrict__ act, const int32_t *__restrict__ newid, const int32_t *__restrict__ pos,
                               const uint32_t *__restrict__ row_off, int32_t *__restrict__ m_cls)
{
    const int lane = threadIdx.x & 31;
    const int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (t >= T) return;
    const int p = pos[t];
    if (p < 0) return;
    uint32_t out = row_off[p];
    const uint32_t e0 = txm_off[t], e1 = txm_off[t + 1];
    for (uint32_t e = e0; e < e1; e += 32) {
        int a = 0, id = 0;
        if (e + lane < e1) { int i = txm_cid[e + lane] - T; a = act[i]; id = newid[i]; }
        unsigned m = __ballot_sync(0xffffffffu, a != 0);
        if (a) m_cls[out + __popc(m & ((1u << lane) - 1))] = id;
        out += __popc(m);
    }
}

// Cardinality segments of the active classes: counts, tid-layout offsets and tile ranges. <= ~1000 segments.
struct SegTab {          // device arrays of n_kseg (+1) entries
    int32_t *j0;         // first compact id
    int32_t *cnt;
    uint32_t *tid_off;   // offset (ints) of the segment's tid block
    int32_t *tile0;      // [n_kseg+1] first tile
    int32_t *cpt;        // classes per tile
    long long *totals;   // [0] n_etiles, [1] e_tid ints, [2] C_a
};
__global__ void k_seg_tables(int n_kseg, int32_t T, const int64_t *__restrict__ kseg_cid0, const int32_t *__restrict__ kseg_k,
                             const int32_t *__restrict__ newid, SegTab st)
{
    if (threadIdx.x || blockIdx.x) return;
    long long tiles = 0, ints = 0;
    for (int s = 0; s < n_kseg; s++) {
        int j0 = newid[kseg_cid0[s] - T], j1 = newid[kseg_cid0[s + 1] - T];
        int cnt = j1 - j0, k = kseg_k[s];
        int cpt;
        long long sz;
        if (k <= KT) { cpt = 32; sz = (long long)((cnt + 31) / 32) * 32 * k; }
        else if (k <= KSUB) { cpt = max(4, (E_TILE_TARGET / k) & ~3); cpt = min(cpt, 32); sz = (long long)cnt * k; }
        else { cpt = max(1, E_TILE_TARGET / k); sz = (long long)cnt * k; }
        st.j0[s] = j0; st.cnt[s] = cnt; st.tid_off[s] = (uint32_t)ints; st.tile0[s] = (int32_t)tiles; st.cpt[s] = cpt;
        tiles += (cnt + cpt - 1) / cpt;
        ints += sz;
    }
    st.tile0[n_kseg] = (int32_t)tiles;
    st.totals[0] = tiles; st.totals[1] = ints;
    st.totals[2] = n_kseg ? newid[kseg_cid0[n_kseg] - T] : 0;
}

__device__ __forceinline__ int seg_of_cid(const int64_t *kseg_cid0, int n_kseg, int64_t cid)
{
    int lo = 0, hi = n_kseg - 1;
    while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (kseg_cid0[mid] <= cid) lo = mid; else hi = mid - 1; }
    return lo;
}

// One warp per multi-tid class: write the member list of an active class into the packed E layout
// (as PERMUTED ROW indices, so that theta is stored in row order) and its read count.
__global__ void k_pack_classes(int64_t n_multi, int32_t T, int n_kseg, const int64_t *__restrict__ kseg_cid0,
                               const int32_t *__restrict__ kseg_k, const uint32_t *__restrict__ cls_off,
                               const int32_t *__restrict__ cls_tid, const int32_t *__restrict__ act,
                               const int32_t *__restrict__ newid, const int32_t *__restrict__ R, const int32_t *__restrict__ pos,
                               SegTab st, int32_t *__restrict__ e_tid, int32_t *__restrict__ e_R)
{
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n_multi || !act[i]) return;
    const int64_t cid = T + i;
    const int j = newid[i];
    const int s = seg_of_cid(kseg_cid0, n_kseg, cid);
    const int k = kseg_k[s];
    const int jl = j - st.j0[s];
    const uint32_t o = cls_off[cid];
    const uint32_t base = st.tid_off[s];
    for (int jj = lane; jj < k; jj += 32) {
        int p = pos[cls_tid[o + jj]];
        uint32_t dst = (k <= KT) ? base + (uint32_t)(jl >> 5) * 32u * k + (uint32_t)jj * 32u + (uint32_t)(jl & 31)
                                 : base + (uint32_t)jl * k + jj;
        e_tid[dst] = p;
    }
    if (lane == 0) e_R[j] = R[cid];
}

__global__ void k_etiles(int n_tiles, int n_kseg, const int32_t *__restrict__ kseg_k, SegTab st, int4 *__restrict__ tiles)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_tiles) return;
    int lo = 0, hi = n_kseg - 1;
    while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (st.tile0[mid] <= g) lo = mid; else hi = mid - 1; }
    // skip empty segments that share the same tile0
    while (lo + 1 < n_kseg && st.tile0[lo + 1] <= g) lo++;
    const int s = lo, k = kseg_k[s], cpt = st.cpt[s], lt = g - st.tile0[s];
    const int mode = k <= KT ? 0 : (k <= KSUB ? 1 : 2);
    int4 t;
    t.x = st.j0[s] + lt * cpt;
    t.y = min(cpt, st.cnt[s] - lt * cpt);
    t.z = (int)(st.tid_off[s] + (uint32_t)lt * (uint32_t)cpt * (uint32_t)k);  // mode 0: cpt == 32 -> lt*32*k
    t.w = k | (mode << 16);
    tiles[g] = t;
}

// Short-row tiles: windows of M_WINDOW over cost(p) = row_off[p] + M_ROW_COST * p.
__global__ void k_mtiles(int n_tiles, int n_short, const uint32_t *__restrict__ row_off, int2 *__restrict__ tiles)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_tiles) return;
    int2 out;
#pragma unroll
    for (int side = 0; side < 2; side++) {
        unsigned long long target = (unsigned long long)(g + side) * M_WINDOW;
        int lo = 0, hi = n_short;   // first p in [0, n_short] with cost(p) >= target
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            unsigned long long c = (unsigned long long)row_off[mid] + (unsigned long long)M_ROW_COST * mid;
            if (c >= target) hi = mid; else lo = mid + 1;
        }
        if (side == 0) out.x = lo; else out.y = lo;
    }
    tiles[g] = out;
}

__global__ void k_fill_double(double *p, int64_t n, double v)
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
template <class T> static T *arena_take(char *&cur, size_t n)
{
    T *p = (T *)cur;
    cur += ((n * sizeof(T) + 255) / 256) * 256;
    return p;
}

extern "C" int emsar_sample_prepare(emsar_sample *s, const emsar_solve_opts *opts_in)
{
    CHECK_ARG(s, "emsar_sample_prepare: NULL sample");
    if (!s->have_counts) { emsar_set_err("emsar_sample_prepare: no counts yet (call emsar_sample_count / counts_set first)"); return EMSAR_ERR_STATE; }
    emsar_index *ix = s->index;
    emsar_ctx *ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    CU(cudaSetDevice(ctx->device));
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
    k_adjeuma<<<(unsigned)((C + ADJ_ROWS - 1) / ADJ_ROWS), ADJ_ROWS, 0, st>>>(C, ix->nF, ix->d_euma, ix->d_has_node, s->d_Wf, s->d_adj);
    LAUNCHED(ctx);
    CU(cudaGetLastError());
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
        std::vector<double> adj((size_t)C);
        CU(cudaMemcpyAsync(adj.data(), s->d_adj, (size_t)C * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        compute_sets(ix, adj, &s->eumacut, o.max_ntid_per_sid, &s->h_CS, &s->max_sid);
        s->have_CS = true;
        std::vector<uint8_t> im((size_t)C);
        for (int64_t c = 0; c < C; c++) im[(size_t)c] = s->h_CS[(size_t)c] >= 0;
        CU(cudaMemcpyAsync(s->d_in_model, im.data(), (size_t)C, cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));
        d_in_model = s->d_in_model;
    } else {
        s->max_sid = ix->n_sets_nocut - 1;
    }
    // ---- scratch carve-up (ints / flags / scans) ----
    size_t cub_bytes = 0, b1 = 0, b2 = 0, b3 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, b1, (int32_t *)nullptr, (int32_t *)nullptr, (int)(nm + 1));
    cub::DeviceScan::ExclusiveSum(nullptr, b2, (unsigned long long *)nullptr, (unsigned long long *)nullptr, T + 1);
    cub::DeviceScan::ExclusiveSum(nullptr, b3, (uint32_t *)nullptr, (uint32_t *)nullptr, T + 1);
    cub_bytes = std::max(b1, std::max(b2, b3));
    size_t need = ((cub_bytes + 255) / 256) * 256 + (size_t)(nm + 1) * 8 + (size_t)(T + 1) * (8 + 8 + 4 + 4 + 4) + (size_t)(n_kseg + 2) * 24 + 4096 + 256 * 16;
    void *scr = nullptr;
    TRY(ctx_scratch(ctx, need, &scr));
    char *cur = (char *)scr;
    void *d_cub = arena_take<char>(cur, cub_bytes);
    int32_t *d_act = arena_take<int32_t>(cur, (size_t)nm + 1);
    int32_t *d_newid = arena_take<int32_t>(cur, (size_t)nm + 1);
    unsigned long long *d_rkey = arena_take<unsigned long long>(cur, (size_t)T + 1);
    unsigned long long *d_rpre = arena_take<unsigned long long>(cur, (size_t)T + 1);
    int32_t *d_deg = arena_take<int32_t>(cur, (size_t)T + 1);
    int32_t *d_pos = s->d_pos;   // t -> permuted row (kept for finalize)
    uint32_t *d_degp = arena_take<uint32_t>(cur, (size_t)T + 1);
    SegTab stab;
    stab.j0 = arena_take<int32_t>(cur, (size_t)n_kseg + 1);
    stab.cnt = arena_take<int32_t>(cur, (size_t)n_kseg + 1);
    stab.tid_off = arena_take<uint32_t>(cur, (size_t)n_kseg + 1);
    stab.tile0 = arena_take<int32_t>(cur, (size_t)n_kseg + 2);
    stab.cpt = arena_take<int32_t>(cur, (size_t)n_kseg + 1);
    stab.totals = arena_take<long long>(cur, 4);
    // ---- class model + active scan ----
    const double nscale = (double)N / 1E6;
    const double p10 = pow(10, o.delta);
    k_class_model<<<(unsigned)((C + 1 + 255) / 256), 256, 0, st>>>(C, T, s->d_adj, d_in_model, s->d_R, nscale, p10, s->d_amodel, d_act);
    LAUNCHED(ctx);
    CU(cub::DeviceScan::ExclusiveSum(d_cub, cub_bytes, d_act, d_newid, (int)(nm + 1), st));
    LAUNCHED(ctx);
    // ---- row statistics, row classes, permutation ----
    k_row_stats<<<(unsigned)(((int64_t)T * 32 + 255) / 256), 256, 0, st>>>(T, ix->d_txm_off, ix->d_txm_cid, s->d_adj, s->d_amodel, d_act, d_in_model,
                                                                          s->d_R, s->d_iE, s->d_A, s->d_Rs, d_deg, s->d_lone, d_rkey);
    LAUNCHED(ctx);
    CU(cub::DeviceScan::ExclusiveSum(d_cub, cub_bytes, d_rkey, d_rpre, T + 1, st));
    LAUNCHED(ctx);
    k_seg_tables<<<1, 32, 0, st>>>(n_kseg, T, ix->d_kseg_cid0, ix->d_kseg_k, d_newid, stab);
    LAUNCHED(ctx);
    CU(cudaGetLastError());
    unsigned long long rtot = 0;
    long long tot[3] = {0, 0, 0};
    CU(cudaMemcpyAsync(&rtot, d_rpre + T, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(tot, stab.totals, 24, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const int n_short = (int)(rtot & 0x1FFFFF), n_long = (int)((rtot >> 21) & 0x1FFFFF), n_hub = (int)((rtot >> 42) & 0x1FFFFF);
    const int P = n_short + n_long + n_hub;
    const int64_t n_etiles = tot[0], e_ints = tot[1], C_a = tot[2];
    // ---- state: theta | q in one allocation ----
    size_t theta_bytes = (((size_t)(P > 0 ? P : 1) * 8 + 255) / 256) * 256;
    size_t q_bytes = (((size_t)(C_a > 0 ? C_a : 1) * 8 + 255) / 256) * 256;
    if (theta_bytes + q_bytes > s->state_bytes) {
        if (s->d_state) CU(cudaFree(s->d_state));
        s->d_state = nullptr;
        double *p = nullptr;
        TRY(dev_alloc(&p, (theta_bytes + q_bytes) / 8));
        s->d_state = p;
        s->state_bytes = theta_bytes + q_bytes;
    }
    EmModel &m = s->m;
    memset(&m, 0, sizeof(m));
    m.T = T; m.C_a = C_a;
    m.theta = s->d_state;
    m.q = (double *)((char *)s->d_state + theta_bytes);
    // ---- packed arena (needs nnz_a: upper bound first, exact after the degree scan) ----
    // row permutation writes degp; scan gives row_off; nnz_a = row_off[P]
    size_t arena_bytes = 0;
    {
        auto rnd = [](size_t b) { return ((b + 255) / 256) * 256; };
        arena_bytes += rnd((size_t)(e_ints > 0 ? e_ints : 1) * 4);        // e_tid
        arena_bytes += rnd((size_t)(C_a > 0 ? C_a : 1) * 4);              // e_R
        arena_bytes += rnd((size_t)(n_etiles > 0 ? n_etiles : 1) * 16);   // e_tiles
        arena_bytes += rnd((size_t)(e_ints > 0 ? e_ints : 1) * 4);        // m_cls (nnz_a <= e_ints)
        arena_bytes += rnd((size_t)(P + 1) * 4) * 2;                      // row_off, row_t
        arena_bytes += rnd((size_t)(P + 1) * 16);                         // row_RsA
        size_t mt_max = ((size_t)(e_ints > 0 ? e_ints : 0) + (size_t)M_ROW_COST * (size_t)(P + 1)) / M_WINDOW + 2;
        arena_bytes += rnd(mt_max * 8);                                   // m_tiles
    }
    if (arena_bytes > s->pack_bytes) {
        if (s->d_pack) CU(cudaFree(s->d_pack));
        s->d_pack = nullptr;
        char *p = nullptr;
        TRY(dev_alloc(&p, arena_bytes));
        s->d_pack = p;
        s->pack_bytes = arena_bytes;
    }
    char *ac = (char *)s->d_pack;
    m.e_tid = arena_take<int32_t>(ac, (size_t)(e_ints > 0 ? e_ints : 1));
    m.e_R = arena_take<int32_t>(ac, (size_t)(C_a > 0 ? C_a : 1));
    m.e_tiles = arena_take<int4>(ac, (size_t)(n_etiles > 0 ? n_etiles : 1));
    m.m_cls = arena_take<int32_t>(ac, (size_t)(e_ints > 0 ? e_ints : 1));
    m.row_off = arena_take<uint32_t>(ac, (size_t)P + 1);
    m.row_t = arena_take<int32_t>(ac, (size_t)P + 1);
    m.row_RsA = arena_take<double2>(ac, (size_t)P + 1);
    m.m_tiles = (int2 *)ac;
    m.n_etiles = (int32_t)n_etiles;
    m.n_short = n_short; m.n_long = n_long; m.n_hub = n_hub;
    CU(cudaMemsetAsync(m.e_tid, 0, (size_t)(e_ints > 0 ? e_ints : 1) * 4, st));
    k_row_perm<<<(unsigned)((T + 255) / 256), 256, 0, st>>>(T, d_rpre, d_rkey, d_deg, s->d_Rs, s->d_A, d_pos, m.row_t, d_degp, m.row_RsA);
    LAUNCHED(ctx);
    CU(cub::DeviceScan::ExclusiveSum(d_cub, cub_bytes, d_degp, m.row_off, P + 1, st));
    LAUNCHED(ctx);
    k_scatter_rows<<<(unsigned)(((int64_t)T * 32 + 255) / 256), 256, 0, st>>>(T, ix->d_txm_off, ix->d_txm_cid, d_act, d_newid, d_pos, m.row_off, m.m_cls);
    LAUNCHED(ctx);
    if (nm > 0) {
        k_pack_classes<<<(unsigned)((nm * 32 + 255) / 256), 256, 0, st>>>(nm, T, n_kseg, ix->d_kseg_cid0, ix->d_kseg_k, ix->d_cls_off, ix->d_cls_tid,
                                                                          d_act, d_newid, s->d_R, d_pos, stab, m.e_tid, m.e_R);
        LAUNCHED(ctx);
    }
    if (n_etiles > 0) {
        k_etiles<<<(unsigned)((n_etiles + 255) / 256), 256, 0, st>>>((int)n_etiles, n_kseg, ix->d_kseg_k, stab, m.e_tiles);
        LAUNCHED(ctx);
    }
    CU(cudaGetLastError());
    uint32_t off_short = 0, off_all = 0;
    CU(cudaMemcpyAsync(&off_short, m.row_off + n_short, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&off_all, m.row_off + P, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    m.nnz_a = off_all;
    const int64_t n_mtiles = ((int64_t)off_short + (int64_t)M_ROW_COST * n_short + M_WINDOW - 1) / M_WINDOW;
    m.n_mtiles = (int32_t)n_mtiles;
    if (n_mtiles > 0) {
        k_mtiles<<<(unsigned)((n_mtiles + 255) / 256), 256, 0, st>>>((int)n_mtiles, n_short, m.row_off, m.m_tiles);
        LAUNCHED(ctx);
    }
    // start point: theta = 1 for every row that takes part (A_t > 0)
    if (P > 0) { k_fill_double<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(m.theta, P, 1.0); LAUNCHED(ctx); }
    CU(cudaGetLastError());
    CU(cudaEventRecord(e1, st));
    CU(cudaStreamSynchronize(st));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    s->prep_ms = ms;
    emsar_model_stats &ms_ = s->stats;
    memset(&ms_, 0, sizeof(ms_));
    ms_.T = T; ms_.C_a = C_a; ms_.nnz_a = m.nnz_a;
    ms_.rows_short = n_short; ms_.rows_long = n_long; ms_.rows_hub = n_hub; ms_.rows_fixed = T - P;
    ms_.e_tiles = n_etiles; ms_.m_tiles = n_mtiles;
    ms_.bytes_per_iter = 8 * m.nnz_a + 24 * C_a + 44 * (int64_t)T;
    // what the kernels stream: E: tids (padded) + R + q write + tiles; M: m_cls + row_off + RsA + theta r/w + tiles
    ms_.stream_bytes_per_iter = 4 * e_ints + 4 * C_a + 8 * C_a + 16 * n_etiles + 4 * m.nnz_a + 4 * (int64_t)P + 16 * (int64_t)P + 16 * (int64_t)P + 8 * n_mtiles;
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
    CU(cudaSetDevice(s->ctx->device));
    CU(cudaMemcpyAsync(Wf, s->d_Wf, (size_t)s->index->nF * 8, cudaMemcpyDeviceToHost, s->ctx->stream));
    CU(cudaStreamSynchronize(s->ctx->stream));
    return EMSAR_OK;
}

// set ids for the -g output (computed lazily when the cut loop was not needed)
int sample_ensure_sets(emsar_sample *s)
{
    if (s->have_CS) return EMSAR_OK;
    emsar_index *ix = s->index;
    std::vector<double> adj((size_t)ix->C);
    CU(cudaMemcpyAsync(adj.data(), s->d_adj, (size_t)ix->C * 8, cudaMemcpyDeviceToHost, s->ctx->stream));
    CU(cudaStreamSynchronize(s->ctx->stream));
    double cut = s->eumacut;
    compute_sets(ix, adj, &cut, s->opts.max_ntid_per_sid > 0 ? s->opts.max_ntid_per_sid : 5000, &s->h_CS, &s->max_sid);
    s->have_CS = true;
    return EMSAR_OK;
}
