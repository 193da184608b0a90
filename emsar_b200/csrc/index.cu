// rsh index on the device: class-major CSR, transpose of the multi-tid classes, cardinality segments,
// and the open-addressing hash (class key -> cid) that replaces the rshbucket chains
// (reference emsar.h:76-83,139-145; emsar_functions.c:1334-1347, 1432-1510, 1542-1625).
#include <algorithm>
#include <numeric>

#include "common.cuh"

// One thread per multi-tid class: hash the sorted multiset and claim a slot by linear probing.
// Only "reachable" classes are inserted: a chain node that an earlier node of the same (cardinality,
// first tid) chain shadows can never be hit by update_rshbucket's walk (:1603-1622).
__global__ void k_hash_build(int64_t n_multi, int32_t T, const uint32_t *__restrict__ cls_off,
                             const int32_t *__restrict__ cls_tid, const uint8_t *__restrict__ insertable,
                             unsigned long long *table, uint64_t mask, unsigned long long *n_inserted)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_multi) return;
    if (!insertable[i]) return;
    int64_t cid = T + i;
    uint32_t o = cls_off[cid];
    int k = (int)(cls_off[cid + 1] - o);
    uint64_t sum = 0;
    for (int j = 0; j < k; j++) sum += key_elem(j, cls_tid[o + j]);
    uint64_t h = key_finish(sum, k);
    unsigned long long entry = ((h >> 32) << 32) | (unsigned long long)(uint32_t)(cid + 1);
    uint64_t s = h & mask;
    for (;;) {
        unsigned long long prev = atomicCAS(&table[s], 0ULL, entry);
        if (prev == 0ULL) break;
        s = (s + 1) & mask;
    }
    atomicAdd(n_inserted, 1ULL);
}

int index_build_hash(emsar_index *ix, const std::vector<uint8_t> &insertable)
{
    emsar_ctx *ctx = ix->ctx;
    uint64_t slots = 1024;
    while (slots < (uint64_t)ix->n_multi * 2) slots <<= 1;
    ix->hash_mask = slots - 1;
    TRY(dev_alloc(&ix->d_hash, slots));
    ix->device_bytes += slots * 8;
    CU(cudaMemsetAsync(ix->d_hash, 0, slots * 8, ctx->stream));
    ix->hash_inserted = 0;
    if (ix->n_multi == 0) return EMSAR_OK;
    uint8_t *d_ins = nullptr;
    unsigned long long *d_cnt = nullptr;
    TRY(dev_alloc(&d_ins, (size_t)ix->n_multi));
    TRY(dev_alloc(&d_cnt, 1));
    CU(cudaMemcpyAsync(d_ins, insertable.data(), (size_t)ix->n_multi, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(d_cnt, 0, 8, ctx->stream));
    int bs = 256;
    k_hash_build<<<(unsigned)((ix->n_multi + bs - 1) / bs), bs, 0, ctx->stream>>>(ix->n_multi, ix->T, ix->d_cls_off, ix->d_cls_tid, d_ins,
                                                                                  ix->d_hash, ix->hash_mask, d_cnt);
    LAUNCHED(ctx);
    CU(cudaGetLastError());
    unsigned long long cnt = 0;
    CU(cudaMemcpyAsync(&cnt, d_cnt, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ix->hash_inserted = (int64_t)cnt;
    dev_free(d_ins);
    dev_free(d_cnt);
    return EMSAR_OK;
}

static int uf_find(std::vector<int32_t> &p, int x)
{
    while (p[x] != x) { p[x] = p[p[x]]; x = p[x]; }
    return x;
}

// ---- locality order of the transcripts (host only; no device needed) -----------------------------------------------------------
// The EM kernel cuts the participating rows into contiguous per-CTA ranges in THIS order, so what matters is that transcripts which
// share classes sit close together however the fasta was sorted (by name, randomly) and wherever paralogs lie.
//   level 1: connected components of the class <-> transcript graph (in the order of their smallest tid) are kept together;
//   level 2: inside a component, transcripts that co-occur in at least TWO small classes (cardinality <= 8) form a cluster - the isoforms
//            of a gene / a tight gene family share many segments, whereas two paralogs or a repeat typically share one - and the
//            clusters follow each other in the order of their smallest tid; inside a cluster the transcripts are laid out breadth-first
//            over those strong links from a pseudo-peripheral transcript (a chain of overlapping isoform groups becomes a band).
// mode 3 = level 1, plus level 2 when that lowers the share of member references leaving an SM's range (what the library uses),
// 2 = both levels, 1 = components only, 0 = plain tid order (EMSAR_ORDER=cluster / component / tid: A-B runs).
// share of the member references that leave the block of their class's first member when the transcripts are cut, in this order, into
// nblocks equal ranges: the quantity the per-SM ownership of the EM kernel wants small
static double order_remote_share(int32_t T, int64_t C, const int64_t *class_ptr, const int32_t *class_tid, const int32_t *order, int nblocks)
{
    std::vector<int32_t> blk((size_t)T);
    for (int32_t i = 0; i < T; i++) blk[(size_t)order[i]] = (int32_t)((int64_t)i * nblocks / T);
    int64_t remote = 0, all = 0;
    for (int64_t c = T; c < C; c++) {
        const int64_t o = class_ptr[c], e = class_ptr[c + 1];
        const int32_t b0 = blk[(size_t)class_tid[o]];
        for (int64_t j = o + 1; j < e; j++) remote += blk[(size_t)class_tid[j]] != b0;
        all += e - o;
    }
    return all ? (double)remote / (double)all : 0.0;
}

extern "C" int emsar_locality_order(int32_t T, int64_t C, const int64_t *class_ptr, const int32_t *class_tid, int32_t mode, int32_t *order)
{
    CHECK_ARG(T > 0 && C >= T && class_ptr && class_tid && order, "emsar_locality_order: bad argument");
    if (mode == 0) { std::iota(order, order + T, 0); return EMSAR_OK; }
    if (mode == 3) {
        // automatic: the component order is kept when it is already local (a fasta sorted by gene); otherwise the cluster order is
        // tried and the better of the two wins
        constexpr int NB = 148;
        TRY(emsar_locality_order(T, C, class_ptr, class_tid, 1, order));
        const double s1 = order_remote_share(T, C, class_ptr, class_tid, order, NB);
        if (s1 < 0.15) return EMSAR_OK;
        std::vector<int32_t> o2((size_t)T);
        TRY(emsar_locality_order(T, C, class_ptr, class_tid, 2, o2.data()));
        if (order_remote_share(T, C, class_ptr, class_tid, o2.data(), NB) < s1) memcpy(order, o2.data(), sizeof(int32_t) * (size_t)T);
        return EMSAR_OK;
    }
    std::vector<int32_t> par((size_t)T), cmin((size_t)T), smin((size_t)T);
    std::iota(par.begin(), par.end(), 0);
    for (int64_t c = T; c < C; c++) {
        const int r0 = uf_find(par, class_tid[class_ptr[c]]);
        for (int64_t j = class_ptr[c] + 1; j < class_ptr[c + 1]; j++) { const int r = uf_find(par, class_tid[j]); if (r != r0) par[(size_t)r] = r0; }
    }
    std::fill(cmin.begin(), cmin.end(), -1);
    for (int32_t t = 0; t < T; t++) { const int r = uf_find(par, t); if (cmin[(size_t)r] < 0) cmin[(size_t)r] = t; }      // ascending t: first seen = smallest
    for (int32_t t = 0; t < T; t++) cmin[(size_t)t] = cmin[(size_t)uf_find(par, t)];
    if (mode == 2) {
        constexpr int KP = 8;
        std::vector<uint64_t> pairs;
        for (int64_t c = T; c < C; c++) {
            const int64_t o = class_ptr[c], k = class_ptr[c + 1] - o;
            if (k > KP) break;                                  // classes are ordered by cardinality
            for (int64_t i = 0; i < k; i++)
                for (int64_t j = i + 1; j < k; j++)
                    if (class_tid[o + i] != class_tid[o + j]) pairs.push_back(((uint64_t)(uint32_t)class_tid[o + i] << 32) | (uint32_t)class_tid[o + j]);
        }
        std::sort(pairs.begin(), pairs.end());
        // strong edges = pairs seen at least twice; adjacency in CSR form (both directions)
        std::vector<uint64_t> strong;
        for (size_t i = 0; i + 1 < pairs.size(); i++)
            if (pairs[i] == pairs[i + 1] && (strong.empty() || strong.back() != pairs[i])) strong.push_back(pairs[i]);
        pairs.clear(); pairs.shrink_to_fit();
        std::vector<uint32_t> adj_off((size_t)T + 1, 0);
        for (uint64_t e : strong) { adj_off[(size_t)(e >> 32) + 1]++; adj_off[(size_t)(uint32_t)e + 1]++; }
        for (int32_t t = 0; t < T; t++) adj_off[(size_t)t + 1] += adj_off[(size_t)t];
        std::vector<int32_t> adj((size_t)adj_off[(size_t)T]);
        {
            std::vector<uint32_t> cur(adj_off.begin(), adj_off.end() - 1);
            for (uint64_t e : strong) {                              // sorted by (a, b): every adjacency list comes out in ascending order
                const int a = (int)(e >> 32), b = (int)(uint32_t)e;
                adj[cur[(size_t)a]++] = b; adj[cur[(size_t)b]++] = a;
            }
        }
        std::iota(par.begin(), par.end(), 0);
        for (uint64_t e : strong) {
            const int a = uf_find(par, (int)(e >> 32)), b = uf_find(par, (int)(uint32_t)e);
            if (a != b) par[(size_t)b] = a;
        }
        std::fill(smin.begin(), smin.end(), -1);
        for (int32_t t = 0; t < T; t++) { const int r = uf_find(par, t); if (smin[(size_t)r] < 0) smin[(size_t)r] = t; }
        for (int32_t t = 0; t < T; t++) smin[(size_t)t] = smin[(size_t)uf_find(par, t)];
        // inside a cluster: breadth-first order over the strong edges from a pseudo-peripheral transcript (the last one a first sweep from
        // the smallest tid reaches), which lays a chain of overlapping isoform groups out as a band instead of in name order
        std::vector<int32_t> pos((size_t)T, -1), queue;
        std::vector<uint8_t> seen((size_t)T, 0);
        queue.reserve(1024);
        auto bfs = [&](int32_t start, uint8_t mark) {
            queue.clear(); queue.push_back(start); seen[(size_t)start] = mark;
            for (size_t h = 0; h < queue.size(); h++) {
                const int32_t u = queue[h];
                for (uint32_t e = adj_off[(size_t)u]; e < adj_off[(size_t)u + 1]; e++) {
                    const int32_t v = adj[e];
                    if (seen[(size_t)v] != mark) { seen[(size_t)v] = mark; queue.push_back(v); }
                }
            }
        };
        for (int32_t t = 0; t < T; t++) {
            if (smin[(size_t)t] != t) continue;                     // t = smallest tid of its cluster
            bfs(t, 1);
            const int32_t far = queue.back();
            bfs(far, 2);
            for (size_t i = 0; i < queue.size(); i++) pos[(size_t)queue[i]] = (int32_t)i;
        }
        std::iota(order, order + T, 0);
        std::sort(order, order + T, [&](int32_t a, int32_t b) {
            if (cmin[(size_t)a] != cmin[(size_t)b]) return cmin[(size_t)a] < cmin[(size_t)b];
            if (smin[(size_t)a] != smin[(size_t)b]) return smin[(size_t)a] < smin[(size_t)b];
            return pos[(size_t)a] < pos[(size_t)b];
        });
        return EMSAR_OK;
    }
    std::iota(order, order + T, 0);
    std::sort(order, order + T, [&](int32_t a, int32_t b) {
        if (cmin[(size_t)a] != cmin[(size_t)b]) return cmin[(size_t)a] < cmin[(size_t)b];
        return a < b;
    });
    return EMSAR_OK;
}

extern "C" int emsar_index_create(emsar_ctx *ctx, const emsar_index_desc *d, emsar_index **out)
{
    CHECK_ARG(ctx && d && out, "emsar_index_create: NULL argument");
    *out = nullptr;
    CHECK_ARG(d->T > 0 && d->C >= d->T && d->class_ptr && d->class_tid && d->euma && d->nF >= 1,
              "emsar_index_create: empty or incomplete descriptor");
    CHECK_ARG(d->T < (1 << 21), "emsar_index_create: T = %d exceeds the supported 2^21 - 1 transcripts", d->T);
    const int32_t T = d->T;
    const int64_t C = d->C;
    const int64_t nnz = d->class_ptr[C];
    if (nnz >= ((int64_t)1 << 31) || C >= ((int64_t)1 << 31)) { emsar_set_err("index too large: nnz=%lld C=%lld", (long long)nnz, (long long)C); return EMSAR_ERR_UNSUPPORTED; }
    // ---- validate the scan order (scan_rshbucket :2149-2191) ----
    for (int32_t t = 0; t < T; t++)
        if (d->class_ptr[t] != t || d->class_tid[t] != t) { emsar_set_err("class %d is not the singleton of tid %d", t, t); return EMSAR_ERR_BAD_INDEX; }
    if (d->class_ptr[T] != T) { emsar_set_err("class_ptr[T] != T"); return EMSAR_ERR_BAD_INDEX; }
    std::vector<KSeg> kseg;
    std::vector<uint8_t> insertable((size_t)(C - T), 1);
    int32_t max_card = 1;
    const emsar_index_aux *aux = d->aux;
    if (aux) {
        // packed image: the table was validated when the image was made; only the cardinality segments are rebuilt (one pass over class_ptr)
        CHECK_ARG(aux->txm_off && aux->order && aux->insertable && (aux->txm_cid || aux->nnz_multi == 0) && aux->nnz_multi == nnz - T,
                  "emsar_index_create: the packed image's derived arrays do not match the class table");
        int prev_k = 1;
        for (int64_t c = T; c < C; c++) {
            const int k = (int)(d->class_ptr[c + 1] - d->class_ptr[c]);
            if (k < 2 || k < prev_k || k > d->max_t_size) { emsar_set_err("class %lld out of scan order (packed image)", (long long)c); return EMSAR_ERR_BAD_INDEX; }
            if (k != prev_k) kseg.push_back(KSeg{k, c, c});
            kseg.back().cid1 = c + 1;
            prev_k = k;
            if (k > max_card) max_card = k;
        }
        memcpy(insertable.data(), aux->insertable, (size_t)(C - T));
    } else {
        int prev_k = 1, prev_t0 = -1;
        int64_t chain_max = -1; // cid holding the running maximum key of the current chain
        for (int64_t c = T; c < C; c++) {
            int64_t o = d->class_ptr[c];
            int64_t k64 = d->class_ptr[c + 1] - o;
            if (k64 < 2) { emsar_set_err("class %lld has cardinality %lld (multi-tid classes need >= 2)", (long long)c, (long long)k64); return EMSAR_ERR_BAD_INDEX; }
            int k = (int)k64;
            if (k > d->max_t_size) { emsar_set_err("class %lld has %d tids > header max_t_size %d", (long long)c, k, d->max_t_size); return EMSAR_ERR_BAD_INDEX; }
            const int32_t *t = d->class_tid + o;
            for (int j = 0; j < k; j++) {
                if (t[j] < 0 || t[j] >= T) { emsar_set_err("class %lld: tid %d out of range", (long long)c, t[j]); return EMSAR_ERR_BAD_INDEX; }
                if (j && t[j] < t[j - 1]) { emsar_set_err("class %lld: tids not sorted", (long long)c); return EMSAR_ERR_BAD_INDEX; }
            }
            if (k < prev_k || (k == prev_k && t[0] < prev_t0)) {
                emsar_set_err("class %lld out of scan order (cardinality, first tid)", (long long)c);
                return EMSAR_ERR_BAD_INDEX;
            }
            if (k != prev_k) { kseg.push_back(KSeg{k, c, c}); }
            kseg.back().cid1 = c + 1;
            if (k != prev_k || t[0] != prev_t0) chain_max = c; // new chain: head is always reachable
            else {
                const int32_t *m = d->class_tid + d->class_ptr[chain_max];
                int cmp = 0;
                for (int j = 1; j < k && !cmp; j++) cmp = (t[j] < m[j]) ? -1 : (t[j] > m[j] ? 1 : 0);
                if (cmp > 0) chain_max = c; else insertable[(size_t)(c - T)] = 0; // shadowed: unreachable by the chain walk
            }
            prev_k = k; prev_t0 = t[0];
            if (k > max_card) max_card = k;
        }
    }
    emsar_index *ix = new emsar_index();
    ix->ctx = ctx; ix->T = T; ix->C = C; ix->nnz = nnz; ix->n_multi = C - T; ix->nnz_multi = nnz - T;
    ix->nF = d->nF; ix->min_fl = d->min_fraglength; ix->max_fl = d->max_fraglength; ix->readlength = d->readlength;
    ix->max_t_size = d->max_t_size; ix->max_card = max_card;
    ix->frag_min = d->min_fraglength > d->readlength ? d->min_fraglength : d->readlength;   // determine_fraglength_range :2471-2475
    ix->frag_max = d->max_fraglength >= ix->frag_min ? d->max_fraglength : ix->frag_min;
    if (ix->frag_max - ix->frag_min + 1 != d->nF) {
        emsar_set_err("nF = %d does not match the fragment length range %d..%d", d->nF, ix->frag_min, ix->frag_max);
        delete ix; return EMSAR_ERR_BAD_INDEX;
    }
    if (ix->frag_max > ix->max_fl) { // the reference would index FraglengthCounts out of bounds (:2511)
        emsar_set_err("read length %d exceeds Max_Fraglength %d", d->readlength, d->max_fraglength);
        delete ix; return EMSAR_ERR_BAD_INDEX;
    }
    ix->kseg = kseg;
    ix->device_bytes = 0;
    TRY(ctx_use(ctx));
    // ---- host copies (32-bit) ----
    ix->h_cls_off.resize((size_t)C + 1);
    for (int64_t c = 0; c <= C; c++) ix->h_cls_off[(size_t)c] = (uint32_t)d->class_ptr[c];
    ix->h_cls_tid.assign(d->class_tid, d->class_tid + nnz);
    // ---- transpose of the multi-tid classes (build_TC_from_CT_2 :2201-2227 without the singleton entries) ----
    std::vector<uint32_t> &txm_off = ix->h_txm_off;
    std::vector<int32_t> &txm_cid = ix->h_txm_cid;
    ix->h_insertable = insertable;
    if (aux) {
        txm_off.assign(aux->txm_off, aux->txm_off + T + 1);
        txm_cid.assign(aux->txm_cid, aux->txm_cid + (nnz - T));
        ix->h_order.assign(aux->order, aux->order + T);
        ix->n_sets_nocut = aux->n_sets_nocut; ix->max_set_tids = aux->max_set_tids;
    } else {
    txm_off.assign((size_t)T + 1, 0);
    for (int64_t j = T; j < nnz; j++) txm_off[(size_t)d->class_tid[j] + 1]++;
    for (int32_t t = 0; t < T; t++) txm_off[(size_t)t + 1] += txm_off[(size_t)t];
    txm_cid.resize((size_t)(nnz - T));
    {
        std::vector<uint32_t> cur(txm_off.begin(), txm_off.end() - 1);
        for (int64_t c = T; c < C; c++)
            for (int64_t j = d->class_ptr[c]; j < d->class_ptr[c + 1]; j++) txm_cid[cur[(size_t)d->class_tid[j]]++] = (int32_t)c;
    }
    // ---- sets without EUMAcut (union-find over transcripts); decides whether the cut loop can ever trigger ----
    {
        std::vector<int32_t> par((size_t)T);
        std::iota(par.begin(), par.end(), 0);
        for (int64_t c = T; c < C; c++) {
            int r0 = uf_find(par, d->class_tid[d->class_ptr[c]]);
            for (int64_t j = d->class_ptr[c] + 1; j < d->class_ptr[c + 1]; j++) {
                int r = uf_find(par, d->class_tid[j]);
                if (r != r0) par[(size_t)r] = r0;
            }
        }
        std::vector<int32_t> sz((size_t)T, 0);
        int32_t ns = 0, mx = 0;
        for (int32_t t = 0; t < T; t++) { int r = uf_find(par, t); if (sz[(size_t)r]++ == 0) ns++; if (sz[(size_t)r] > mx) mx = sz[(size_t)r]; }
        ix->n_sets_nocut = ns; ix->max_set_tids = mx;
    }
    // ---- locality order of the transcripts (emsar_locality_order below) ----
    {
        ix->h_order.resize((size_t)T);
        const char *eo = getenv("EMSAR_ORDER");
        const int mode = (eo && !strcmp(eo, "tid")) ? 0 : (eo && !strcmp(eo, "component")) ? 1 : (eo && !strcmp(eo, "cluster")) ? 2 : 3;
        int orc = emsar_locality_order(T, C, d->class_ptr, d->class_tid, mode, ix->h_order.data());
        if (orc != EMSAR_OK) { delete ix; return orc; }
    }
    }   // !aux
    // ---- upload ----
    int rc;
#define UP(dst, src, n, type)                                                                         \
    if ((rc = dev_alloc(&dst, (size_t)(n))) != EMSAR_OK) { emsar_index_destroy(ix); return rc; }      \
    ix->device_bytes += (int64_t)(n) * sizeof(type);                                                  \
    CU(cudaMemcpyAsync(dst, src, (size_t)(n) * sizeof(type), cudaMemcpyHostToDevice, ctx->stream));
    UP(ix->d_cls_off, ix->h_cls_off.data(), C + 1, uint32_t);
    UP(ix->d_cls_tid, ix->h_cls_tid.data(), nnz, int32_t);
    UP(ix->d_euma, d->euma, C * (int64_t)d->nF, int32_t);
    std::vector<uint8_t> hn((size_t)C, 1);
    if (d->has_node) memcpy(hn.data(), d->has_node, (size_t)C);
    UP(ix->d_has_node, hn.data(), C, uint8_t);
    UP(ix->d_txm_off, txm_off.data(), T + 1, uint32_t);
    UP(ix->d_txm_cid, txm_cid.data(), nnz - T, int32_t);
    UP(ix->d_order, ix->h_order.data(), T, int32_t);
    std::vector<int64_t> kc0; std::vector<int32_t> kk;
    for (auto &s : kseg) { kc0.push_back(s.cid0); kk.push_back(s.k); }
    kc0.push_back(C);
    UP(ix->d_kseg_cid0, kc0.data(), kc0.size(), int64_t);
    UP(ix->d_kseg_k, kk.data(), kk.size() ? kk.size() : 1, int32_t);
#undef UP
    CU(cudaStreamSynchronize(ctx->stream));
    rc = index_build_hash(ix, insertable);
    if (rc != EMSAR_OK) { emsar_index_destroy(ix); return rc; }
    *out = ix;
    return EMSAR_OK;
}

extern "C" int emsar_index_info_get(const emsar_index *ix, emsar_index_info *info)
{
    CHECK_ARG(ix && info, "emsar_index_info_get: NULL argument");
    memset(info, 0, sizeof(*info));
    info->T = ix->T; info->C = ix->C; info->nnz = ix->nnz; info->n_multi = ix->n_multi; info->nnz_multi = ix->nnz_multi;
    info->n_kseg = (int32_t)ix->kseg.size(); info->max_card = ix->max_card;
    info->hash_slots = (int64_t)ix->hash_mask + 1; info->hash_inserted = ix->hash_inserted;
    info->n_sets_nocut = ix->n_sets_nocut; info->max_set_tids = ix->max_set_tids;
    info->device_bytes = ix->device_bytes; info->frag_min = ix->frag_min; info->frag_max = ix->frag_max;
    return EMSAR_OK;
}

extern "C" int emsar_index_aux_get(const emsar_index *ix, emsar_index_aux *aux)
{
    CHECK_ARG(ix && aux, "emsar_index_aux_get: NULL argument");
    memset(aux, 0, sizeof(*aux));
    aux->nnz_multi = ix->nnz_multi;
    aux->txm_off = ix->h_txm_off.data(); aux->txm_cid = ix->h_txm_cid.data(); aux->order = ix->h_order.data();
    aux->insertable = ix->h_insertable.data();
    aux->n_sets_nocut = ix->n_sets_nocut; aux->max_set_tids = ix->max_set_tids;
    return EMSAR_OK;
}

extern "C" int emsar_index_destroy(emsar_index *ix)
{
    if (!ix) return EMSAR_OK;
    ctx_use(ix->ctx);
    cudaStreamSynchronize(ix->ctx->stream);
    dev_free(ix->d_cls_off); dev_free(ix->d_cls_tid); dev_free(ix->d_euma); dev_free(ix->d_has_node);
    dev_free(ix->d_txm_off); dev_free(ix->d_txm_cid); dev_free(ix->d_order); dev_free(ix->d_hash); dev_free(ix->d_kseg_cid0); dev_free(ix->d_kseg_k);
    delete ix;
    return EMSAR_OK;
}
