// Packer of the class-owner-centric EM model (k_em_psum, psum.cuh). Included by prep.cu (same translation unit: it reuses the
// ownership kernels and device helpers defined there). One-off per sample, everything on the device except the shared-memory plan.

struct PsPrepIn {
    int32_t T, P;
    int64_t nm, C_a;
    int n_kseg, B;                 // B: virtual CTAs over all ranks (= B_local * nranks)
    int B_local, rank, nranks;     // class-sharded sample: this device runs the CTAs rank * B_local ..; 1 rank otherwise
    void *d_cub;
    size_t cub_bytes;
    const int32_t *d_act, *d_newid, *d_deg;
    const uint32_t *d_rflag, *d_nat;
    int32_t *d_pos;
};

__global__ void k_ps_cell_sizes(int n_cells, int n_kseg, const int32_t *__restrict__ kseg_k, const int32_t *__restrict__ cell_cnt,
                                uint32_t *__restrict__ cell_u16, int32_t *__restrict__ cell_tiles)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > n_cells) return;
    if (c == n_cells) { cell_u16[c] = 0; cell_tiles[c] = 0; return; }
    const int k = kseg_k[c % n_kseg], cnt = cell_cnt[c], cpt = e_cls_per_tile(k), nbt = ps_blocks_per_tile(k);
    const int nblk = (cnt + cpt - 1) / cpt;                 // row blocks (for cardinality <= 4: tiles of 128 / 64 classes)
    cell_tiles[c] = (nblk + nbt - 1) / nbt;
    cell_u16[c] = (uint32_t)nblk * (uint32_t)ps_tile_u16(k);
}

__global__ void k_ps_etiles(int n_tiles, int n_cells, int n_kseg, const int32_t *__restrict__ kseg_k, const int32_t *__restrict__ tilebase,
                            const int32_t *__restrict__ clsbase, const int32_t *__restrict__ cell_cnt, const uint32_t *__restrict__ u16base,
                            const int32_t *__restrict__ cls0, int4 *__restrict__ tiles)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_tiles) return;
    int lo = 0, hi = n_cells - 1;      // largest cell with tilebase[cell] <= g (non-empty by construction)
    while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (tilebase[mid] <= g) lo = mid; else hi = mid - 1; }
    const int c = lo, k = kseg_k[c % n_kseg], cpt = e_cls_per_tile(k), nbt = ps_blocks_per_tile(k), lt = g - tilebase[c], b = c / n_kseg;
    const int nblk = (cell_cnt[c] + cpt - 1) / cpt, nb = min(nbt, nblk - lt * nbt);
    int4 t;
    t.x = clsbase[c] + lt * nbt * cpt - cls0[b];              // local class id
    t.y = min(nb * cpt, cell_cnt[c] - lt * nbt * cpt);
    t.z = (int)((u16base[c] + (uint32_t)(lt * nbt) * (uint32_t)ps_tile_u16(k)) >> 3);      // 16-byte units
    t.w = e_steps(k) | (e_lgG(k) << 12) | (nb << 16);
    tiles[g] = t;
}

// weight of a class in the member-pair stream: its cardinality when it is active
__global__ void k_ps_weights(int64_t n_multi, int32_t T, const uint32_t *__restrict__ cls_off, const int32_t *__restrict__ act, uint32_t *__restrict__ w)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n_multi) return;
    w[i] = (i < n_multi && act[i]) ? cls_off[T + i + 1] - cls_off[T + i] : 0u;
}

// One warp per active class: its members as 16-bit theta slots of the owner CTA in the tile layout of psum.cuh, its read count, and one
// (CTA, row slot, local class) key per member for the M side.
__global__ void k_ps_pack_classes(int64_t n_multi, int32_t T, int n_kseg, const int32_t *__restrict__ blk_hr0, const unsigned long long *__restrict__ uniq_e,
                                  const int32_t *__restrict__ kseg_k, const uint32_t *__restrict__ cls_off, const int32_t *__restrict__ cls_tid,
                                  const int32_t *__restrict__ act, const int32_t *__restrict__ newid, const int32_t *__restrict__ cellof,
                                  const int32_t *__restrict__ newid2, const int32_t *__restrict__ clsbase, const uint32_t *__restrict__ u16base,
                                  const int32_t *__restrict__ R, const int32_t *__restrict__ pos, const int32_t *__restrict__ row0,
                                  const int32_t *__restrict__ cls0, const uint32_t *__restrict__ apos, uint16_t *__restrict__ e_data,
                                  uint32_t *__restrict__ e_R, unsigned long long *__restrict__ pairs)
{
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n_multi || !act[i]) return;
    const int cell = cellof[newid[i]];
    const int ob = cell / n_kseg, k = kseg_k[cell % n_kseg];
    const int jn = newid2[i];
    const int jl = jn - clsbase[cell];                 // index inside the cell
    const int lc = jn - cls0[ob];                      // local class id of the owner CTA
    const int r0 = row0[ob], r1 = row0[ob + 1];
    const int nrows = r1 - r0, nhr = blk_hr0[ob + 1] - blk_hr0[ob];
    const int zslot = nrows + nhr;
    const uint32_t o = cls_off[T + i];
    const uint32_t pb = apos[i];
    const int cpt = e_cls_per_tile(k);
    const uint32_t tbase = u16base[cell] + (uint32_t)(jl / cpt) * (uint32_t)ps_tile_u16(k);
    const int ci = jl % cpt;
    const int lg = e_lgG(k), G = 1 << lg, steps = e_steps(k), steps4 = (steps + 3) >> 2;
    const int n_ent = k <= 4 ? (k == 2 ? 2 : 4) : steps4 * 4 * G;
    for (int jj = lane; jj < n_ent; jj += 32) {
        int slot = zslot;
        if (jj < k) {
            const int p = pos[cls_tid[o + jj]];
            if (p >= r0 && p < r1) slot = p - r0;
            else slot = nrows + halo_find(uniq_e, blk_hr0[ob], blk_hr0[ob + 1], ((unsigned long long)ob << 32) | (unsigned long long)(uint32_t)p);
            pairs[pb + jj] = ((unsigned long long)ob << 32) | ((unsigned long long)(uint32_t)slot << 16) | (unsigned long long)(uint32_t)lc;
        }
        uint32_t at;
        if (k <= 4) {
            const int g = ci >> 5, l = ci & 31, W = k == 2 ? 2 : 4;
            at = (uint32_t)(l * 8 + g * W + jj);
        } else {
            const int cb = ci;                                   // one row block per tile
            const int l = cb * G + (jj & (G - 1)), step = jj >> lg;
            at = (uint32_t)((step >> 2) * 128 + l * 4 + (step & 3));
        }
        e_data[tbase + at] = (uint16_t)slot;
    }
    if (lane == 0) e_R[jn] = (uint32_t)R[T + i];
}

__global__ void k_ps_heads(int64_t n, const unsigned long long *__restrict__ keys, uint32_t *__restrict__ flag)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    flag[i] = (i < n && (i == 0 || (keys[i] >> 16) != (keys[i - 1] >> 16))) ? 1u : 0u;
}
__global__ void k_ps_touched(int64_t n, const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ flag, const uint32_t *__restrict__ tidx,
                             uint32_t n_tr, unsigned long long *__restrict__ tr_key, uint32_t *__restrict__ tr_start)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) tr_start[n_tr] = (uint32_t)n;
    if (i >= n || !flag[i]) return;
    tr_key[tidx[i]] = keys[i] >> 16;            // (CTA << 16) | row slot
    tr_start[tidx[i]] = (uint32_t)i;
}
// Order of a CTA's touched rows: long rows (groups) first, then the short rows, longest first. With remote_first the short rows whose owner is
// another CTA come before its own, so that the partial sums other CTAs wait for are sent in the first part of the M-phase.
__global__ void k_ps_touched_keys(uint32_t n_tr, const unsigned long long *__restrict__ tr_key, const uint32_t *__restrict__ tr_start,
                                  const int32_t *__restrict__ row0, int remote_first, unsigned long long *__restrict__ key, int32_t *__restrict__ val)
{
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_tr) return;
    const uint32_t deg = tr_start[j + 1] - tr_start[j];
    const int b = (int)(tr_key[j] >> 16), slot = (int)(tr_key[j] & 0xffffu);
    const uint32_t part = deg > (uint32_t)M_LONG ? 2u : ((remote_first && slot >= row0[b + 1] - row0[b]) ? 1u : 0u);
    key[j] = ((tr_key[j] >> 16) << 32) | (unsigned long long)(0xFFFFFFFFu - ((part << 28) | (deg & 0x0FFFFFFFu)));
    val[j] = (int32_t)j;
}
__global__ void k_ps_sorted_deg(uint32_t n_tr, const unsigned long long *__restrict__ skey, uint32_t *__restrict__ tdeg)
{
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j <= n_tr) tdeg[j] = j < n_tr ? ((0xFFFFFFFFu - (uint32_t)skey[j]) & 0x0FFFFFFFu) : 0u;
}

// M items of every CTA over its touched rows (sorted longest first): groups of its long rows, then slices of 32 rows.
__global__ void k_ps_items_count(int B, const int32_t *__restrict__ tr0, const uint32_t *__restrict__ tdeg, int32_t *__restrict__ nlong,
                                 int32_t *__restrict__ nitems, int32_t *__restrict__ ngroups)
{
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int r0 = tr0[b], r1 = tr0[b + 1];
    const int nl = count_long(tdeg, r0, r1);
    int groups = 0, rows = 0; uint32_t ent = 0;
    for (int i = 0; i < nl; i++) {
        const uint32_t d = tdeg[r0 + i];
        if (rows > 0 && (rows == M_GROUP_ROWS || ent + d > (uint32_t)M_GROUP_ENTRIES)) { groups++; rows = 0; ent = 0; }
        rows++; ent += d;
    }
    if (rows > 0) groups++;
    nlong[b] = nl; ngroups[b] = groups;
    nitems[b] = groups + ((r1 - r0 - nl) + 31) / 32;
}
// one thread per CTA: item descriptors {size16, rows, -, length | group << 30}, their sizes in 16-byte units, and for every long row its
// index inside the group and its entry offset (u16) behind the group header
__global__ void k_ps_items_fill(int B, const int32_t *__restrict__ tr0, const uint32_t *__restrict__ tdeg, const int32_t *__restrict__ nlong,
                                const int32_t *__restrict__ item0, uint32_t *__restrict__ size16, int4 *__restrict__ items,
                                uint32_t *__restrict__ rowbase, int32_t *__restrict__ rowitem, int32_t *__restrict__ rowidx)
{
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b == 0) size16[item0[B]] = 0;
    if (b >= B) return;
    const int r0 = tr0[b], r1 = tr0[b + 1], nl = nlong[b];
    int g = item0[b];
    int rows = 0; uint32_t ent = 0, ent16 = 0;       // ent: entries (grouping rule), ent16: padded u16 entries written
    for (int i = 0; i <= nl; i++) {
        const uint32_t d = i < nl ? tdeg[r0 + i] : 0;
        if (rows > 0 && (i == nl || rows == M_GROUP_ROWS || ent + d > (uint32_t)M_GROUP_ENTRIES)) {
            const uint32_t u16s = (uint32_t)ps_group_hdr_u16(rows) + ent16;
            const uint32_t s16 = (u16s + 7) >> 3;
            items[g] = make_int4((int)s16, rows, 0, (int)(ent | (1u << 30)));
            size16[g] = s16;
            g++; rows = 0; ent = 0; ent16 = 0;
        }
        if (i == nl) break;
        rowbase[r0 + i] = ent16;
        rowitem[r0 + i] = g;
        rowidx[r0 + i] = rows;
        rows++; ent += d; ent16 += (d + 1u) & ~1u;
    }
    const int nrows = r1 - r0;
    for (int s0 = nl; s0 < nrows; s0 += 32, g++) {
        uint32_t d = 0;                                 // the slice that straddles the remote / own boundary is not sorted
        for (int i = s0; i < min(s0 + 32, nrows); i++) d = max(d, tdeg[r0 + i]);
        const uint32_t s16 = (uint32_t)ps_slice_u16((int)d) >> 3;
        items[g] = make_int4((int)s16, min(32, nrows - s0), 0, (int)d);
        size16[g] = s16;
    }
}
__global__ void k_ps_item_offsets(int n_items, const uint32_t *__restrict__ off16, int4 *__restrict__ items)
{
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n_items) items[g].z = (int)off16[g];
}
// one warp per item: slices start out as "no row, every entry = the zero-q slot"; group headers are cleared
__global__ void k_ps_item_init(int n_items, int B, const int32_t *__restrict__ item0, const int32_t *__restrict__ cls0, const int4 *__restrict__ items,
                               uint16_t *__restrict__ m_data)
{
    const int lane = threadIdx.x & 31;
    const int g = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (g >= n_items) return;
    const int4 it = items[g];
    int lo = 0, hi = B - 1;
    while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (item0[mid] <= g) lo = mid; else hi = mid - 1; }
    const uint16_t zq = (uint16_t)(cls0[lo + 1] - cls0[lo]);
    uint16_t *d = m_data + (size_t)it.z * 8;
    const int n = it.x * 8;
    if ((it.w >> 30) & 1) { for (int j = lane; j < n; j += 32) d[j] = zq; for (int j = lane; j < ps_group_hdr_u16(it.y); j += 32) d[j] = 0; }
    else for (int j = lane; j < n; j += 32) d[j] = j < 64 ? (uint16_t)0xFFFF : zq;            // 32 destinations = PS_NONE
}
// one warp per touched row (in the CTA's sorted order): its local classes, ascending, into its slice column or group row
__global__ void k_ps_scatter_rows(uint32_t n_tr, int B, const unsigned long long *__restrict__ skey, const int32_t *__restrict__ tperm,
                                  const unsigned long long *__restrict__ tr_key, const uint32_t *__restrict__ tr_start,
                                  const unsigned long long *__restrict__ pairs, const int32_t *__restrict__ tr0, const int32_t *__restrict__ nlong,
                                  const int32_t *__restrict__ ngroups, const int32_t *__restrict__ item0, const uint32_t *__restrict__ rowbase,
                                  const int32_t *__restrict__ rowitem, const int32_t *__restrict__ rowidx, const int4 *__restrict__ items,
                                  const int32_t *__restrict__ row0, const int32_t *__restrict__ hr0, const int32_t *__restrict__ halo_tgt,
                                  uint16_t *__restrict__ m_data)
{
    const int lane = threadIdx.x & 31;
    const uint32_t j = (uint32_t)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (j >= n_tr) return;
    const int b = (int)(skey[j] >> 32);
    const int ti = tperm[j];
    const uint32_t s0 = tr_start[ti], deg = tr_start[ti + 1] - s0;
    const int slot = (int)(tr_key[ti] & 0xffffu), nrows = row0[b + 1] - row0[b];
    // where the row's partial sum goes: the CTA's own array, or the slot its owner reads (in the owner's rank)
    const uint32_t dst = slot < nrows ? (uint32_t)slot : (PS_REMOTE | (uint32_t)halo_tgt[hr0[b] + (slot - nrows)]);
    const int jj = (int)j - tr0[b], nl = nlong[b];
    if (jj < nl) {
        const int4 it = items[rowitem[j]];
        uint16_t *d = m_data + (size_t)it.z * 8;
        if (lane == 0) { ((uint32_t *)d)[2 * rowidx[j]] = deg; ((uint32_t *)d)[2 * rowidx[j] + 1] = dst; }
        uint16_t *ent = d + ps_group_hdr_u16(it.y) + rowbase[j];
        for (uint32_t e = lane; e < deg; e += 32) ent[e] = (uint16_t)(pairs[s0 + e] & 0xffffu);
    } else {
        const int4 it = items[item0[b] + ngroups[b] + ((jj - nl) >> 5)];
        const int l = (jj - nl) & 31;
        uint16_t *d = m_data + (size_t)it.z * 8;
        if (lane == 0) ((uint32_t *)d)[l] = dst;
        for (uint32_t e = lane; e < deg; e += 32) d[64 + (e >> 2) * 128 + l * 4 + (e & 3)] = (uint16_t)(pairs[s0 + e] & 0xffffu);
    }
}

// incidences (row, contributing CTA): the halo list re-sorted by row gives every owner the contiguous run of partial-sum slots of its row
__global__ void k_ps_inc_keys(unsigned int n, const unsigned long long *__restrict__ uniq_e, unsigned long long *__restrict__ key, int32_t *__restrict__ val)
{
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    key[i] = ((uniq_e[i] & 0xffffffffULL) << 32) | (uniq_e[i] >> 32);       // (row, CTA)
    val[i] = (int32_t)i;
}
// incidence e = (row p, contributing CTA b): the contributor's halo slot learns where its partial sum goes (slot e in the rank of p's
// owner), and p learns which ranks read its theta
__global__ void k_ps_inc_tgt(unsigned int n, const unsigned long long *__restrict__ skey, const int32_t *__restrict__ sval, const int32_t *__restrict__ row0, int B, int B_local,
                             int32_t *__restrict__ tgt, int32_t *__restrict__ row_mask, unsigned long long *__restrict__ peer_stores, int my_rank)
{
    unsigned int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int p = (int)(skey[e] >> 32), b = (int)(uint32_t)skey[e];
    const int owner_rank = block_of_row(row0, B, p) / B_local, reader_rank = b / B_local;
    tgt[sval[e]] = (int32_t)(e | ((unsigned)owner_rank << 28));
    const int old = atomicOr(&row_mask[p], 1 << reader_rank);
    // stores of THIS rank that cross NVLink per iteration: its partial sums for rows of other ranks, theta of its rows to every other reader rank
    if (reader_rank == my_rank && owner_rank != my_rank) atomicAdd(peer_stores, 1ULL);
    if (owner_rank == my_rank && reader_rank != my_rank && !(old & (1 << reader_rank))) atomicAdd(peer_stores, 1ULL);
}
__global__ void k_ps_inc_off(int32_t P, unsigned int n, const unsigned long long *__restrict__ skey, int32_t *__restrict__ inc_off)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p > P) return;
    const unsigned long long target = (unsigned long long)(uint32_t)p << 32;
    unsigned int lo = 0, hi = n;
    while (lo < hi) { unsigned int mid = (lo + hi) >> 1; if (skey[mid] >= target) hi = mid; else lo = mid + 1; }
    inc_off[p] = (int32_t)lo;
}

// why a sample did not get the psum model (EMSAR_VERBOSE=1)
static void ps_note(const char *fmt, ...)
{
    if (!getenv("EMSAR_VERBOSE")) return;
    va_list ap;
    va_start(ap, fmt);
    fprintf(stderr, "emsar_cuda: k_em_psum not used: ");
    vfprintf(stderr, fmt, ap);
    fprintf(stderr, "\n");
    va_end(ap);
}

static void ps_free(emsar_sample *s)
{
    for (void *p : s->ps_allocs) dev_free(p);
    s->ps_allocs.clear();
    s->use_psum = false;
}
template <class TT> static int ps_alloc(emsar_sample *s, TT **p, size_t n)
{
    TRY(dev_alloc(p, n));
    s->ps_allocs.push_back((void *)*p);
    return EMSAR_OK;
}

// Builds s->ps. Returns EMSAR_OK with s->use_psum = true, or EMSAR_OK with s->use_psum = false when the sample is not eligible (a CTA's
// state does not fit in shared memory / 16-bit slots): the caller then packs the legacy model.
static int sample_build_psum(emsar_sample *s, const PsPrepIn &in)
{
    emsar_index *ix = s->index;
    emsar_ctx *ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    ps_free(s);
    const int32_t T = in.T, P = in.P;
    const int64_t nm = in.nm, C_a = in.C_a;
    const int B = in.B, B_local = in.B_local, n_kseg = in.n_kseg > 0 ? in.n_kseg : 1;
    if (P <= 0 || C_a <= 0 || nm <= 0 || in.n_kseg <= 0) { ps_note("empty model"); return EMSAR_OK; }         // nothing to iterate on: the legacy path handles the degenerate cases
    const int n_cells = B * n_kseg;
    void *d_cub = in.d_cub; size_t cub_bytes = in.cub_bytes;
    auto rnd = [](size_t b) { return ((b + 255) / 256) * 256; };
    // ---- scratch (stream-ordered pool memory, released at the end) ----
    uint32_t *d_w = nullptr, *d_apos = nullptr;
    TRY(dev_alloc(&d_w, (size_t)nm + 2)); TRY(dev_alloc(&d_apos, (size_t)nm + 2));
    k_ps_weights<<<(unsigned)((nm + 1 + 255) / 256), 256, 0, st>>>(nm, T, ix->d_cls_off, in.d_act, d_w);
    CU(cub::DeviceScan::ExclusiveSum(d_cub, cub_bytes, d_w, d_apos, (int)(nm + 1), st));
    ctx->launches += 2;
    uint32_t nnz32 = 0;
    CU(cudaMemcpyAsync(&nnz32, d_apos + nm, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const size_t nnz_a = nnz32;
    char *scr = nullptr;
    const size_t tb = rnd(((size_t)T + 1) * 8), mb = rnd(((size_t)nm + 1) * 4), cb = rnd(((size_t)n_cells + 1) * 4), pb = rnd((nnz_a + 2) * 8), bb = rnd(((size_t)B + 2) * 4);
    const size_t scr_bytes = 12 * tb + 4 * mb + 6 * cb + 4 * pb + 8 * bb + 4096;
    TRY(dev_alloc(&scr, scr_bytes));
    struct Guard { char *p; uint32_t *a, *b; ~Guard() { dev_free(p); dev_free(a); dev_free(b); } } guard{scr, d_w, d_apos};
    char *cur = scr;
    uint32_t *d_degn = arena_take<uint32_t>(cur, (size_t)T + 1), *d_degp = arena_take<uint32_t>(cur, (size_t)T + 1);
    int32_t *d_tn = arena_take<int32_t>(cur, (size_t)T + 1), *d_ecost = arena_take<int32_t>(cur, (size_t)T + 1);
    uint32_t *d_cost = arena_take<uint32_t>(cur, (size_t)T + 1), *d_costp = arena_take<uint32_t>(cur, (size_t)T + 1);
    unsigned long long *d_key = arena_take<unsigned long long>(cur, (size_t)T + 1), *d_key2 = arena_take<unsigned long long>(cur, (size_t)T + 1);
    int32_t *d_val = arena_take<int32_t>(cur, (size_t)T + 1), *d_perm = arena_take<int32_t>(cur, (size_t)T + 1);
    int32_t *d_rown = arena_take<int32_t>(cur, (size_t)T + 1);
    double2 *d_rsan = arena_take<double2>(cur, (size_t)T + 1);
    int32_t *d_owner = arena_take<int32_t>(cur, (size_t)nm + 1), *d_cellof = arena_take<int32_t>(cur, (size_t)nm + 1), *d_newid2 = arena_take<int32_t>(cur, (size_t)nm + 1);
    int32_t *d_cell_cnt = arena_take<int32_t>(cur, (size_t)n_cells + 1), *d_cell_tiles = arena_take<int32_t>(cur, (size_t)n_cells + 1);
    uint32_t *d_cell_u16 = arena_take<uint32_t>(cur, (size_t)n_cells + 1), *d_u16base = arena_take<uint32_t>(cur, (size_t)n_cells + 1);
    int32_t *d_clsbase = arena_take<int32_t>(cur, (size_t)n_cells + 1), *d_tilebase = arena_take<int32_t>(cur, (size_t)n_cells + 1);
    unsigned long long *d_pa = arena_take<unsigned long long>(cur, nnz_a + 2), *d_pb = arena_take<unsigned long long>(cur, nnz_a + 2);
    unsigned long long *d_pc = arena_take<unsigned long long>(cur, nnz_a + 2), *d_uniq_e = arena_take<unsigned long long>(cur, nnz_a + 2);
    int32_t *d_nlong = arena_take<int32_t>(cur, (size_t)B + 1), *d_nitems = arena_take<int32_t>(cur, (size_t)B + 1), *d_ngroups = arena_take<int32_t>(cur, (size_t)B + 1);
    int32_t *d_tr0 = arena_take<int32_t>(cur, (size_t)B + 1);
    unsigned int *d_hcount = arena_take<unsigned int>(cur, 8);
    // ---- model arrays that are sized now ----
    PsModel &m = s->ps;
    memset(&m, 0, sizeof(m));
    m.P = P; m.B = B_local; m.Bt = B; m.block0 = in.rank * B_local; m.rank = in.rank; m.nranks = in.nranks;
    m.C_a = C_a; m.nnz_a = (int64_t)nnz_a; m.smem_bytes = ctx->em_smem_bytes;
    m.theta = s->d_state;
    TRY(ps_alloc(s, &m.blk_row0, (size_t)B + 1)); TRY(ps_alloc(s, &m.blk_cls0, (size_t)B + 1)); TRY(ps_alloc(s, &m.blk_etile0, (size_t)B + 1));
    TRY(ps_alloc(s, &m.blk_mitem0, (size_t)B + 1)); TRY(ps_alloc(s, &m.blk_hr0, (size_t)B + 1)); TRY(ps_alloc(s, &m.blk_desc_smem, (size_t)B + 1));
    TRY(ps_alloc(s, &m.blk_res16, (size_t)B + 1));
    TRY(ps_alloc(s, &m.row_RsA, (size_t)P + 1)); TRY(ps_alloc(s, &m.inc_off, (size_t)P + 2)); TRY(ps_alloc(s, &m.row_mask, (size_t)P + 2));
    CU(cudaMemsetAsync(m.row_mask, 0, ((size_t)P + 2) * 4, st));
    uint32_t *d_eR = nullptr;
    TRY(ps_alloc(s, &d_eR, (size_t)C_a + 1));
    m.e_R = d_eR;
    // ---- rows: ownership by the median member, cost-balanced ranges ----
    const int cost_row = getenv("EMSAR_PS_COST_ROW") ? atoi(getenv("EMSAR_PS_COST_ROW")) : 6;
    const int cost_class = getenv("EMSAR_PS_COST_CLASS") ? atoi(getenv("EMSAR_PS_COST_CLASS")) : 6;
    k_nat_fill<<<(unsigned)((T + 255) / 256), 256, 0, st>>>(T, in.d_rflag, in.d_nat, in.d_deg, in.d_pos, d_degn, d_tn, d_ecost, P);
    k_class_owner<<<(unsigned)((nm + 255) / 256), 256, 0, st>>>(nm, T, ix->d_cls_off, ix->d_cls_tid, in.d_act, in.d_deg, 2, d_owner);
    k_class_cost<<<(unsigned)((nm + 255) / 256), 256, 0, st>>>(nm, T, ix->d_cls_off, d_owner, in.d_act, in.d_nat, d_ecost, cost_class, 2);
    k_row_cost<<<(unsigned)((P + 1 + 255) / 256), 256, 0, st>>>(P, d_degn, d_ecost, d_cost, cost_row, 0);
    {
        // Row ranges on the host (P numbers each way): equal cost per CTA, but a range also ends where the state its CTA must hold in shared
        // memory - theta and the partial sum of every row (16 bytes), q of every class a row owns (8 bytes) - reaches the budget; a region
        // of many small classes would otherwise pile more q into one CTA than an SM can hold. The remaining cost is re-divided over the
        // remaining CTAs after every cut.
        CU(cudaMemsetAsync(d_costp, 0, ((size_t)P + 1) * 4, st));
        k_class_cost<<<(unsigned)((nm + 255) / 256), 256, 0, st>>>(nm, T, ix->d_cls_off, d_owner, in.d_act, in.d_nat, (int32_t *)d_costp, 1, 0);      // classes owned per row
        std::vector<uint32_t> h_cost((size_t)P + 1), h_ncls((size_t)P + 1);
        CU(cudaMemcpyAsync(h_cost.data(), d_cost, ((size_t)P + 1) * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(h_ncls.data(), d_costp, ((size_t)P + 1) * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        const long long cap = (long long)(getenv("EMSAR_PS_STATE_KB") ? atoi(getenv("EMSAR_PS_STATE_KB")) * 1024 : (ctx->em_smem_bytes * 6) / 10);   // the rest: halo theta, descriptors, index cache
        unsigned long long left = 0;
        for (int32_t i = 0; i < P; i++) left += h_cost[(size_t)i];
        std::vector<int32_t> h_r0((size_t)B + 1, P);
        h_r0[0] = 0;
        int b = 0;
        unsigned long long acc = 0, target = left / (unsigned long long)B;
        long long state = 0;
        for (int32_t i = 0; i < P && b < B - 1; i++) {
            const long long sw = 16 + 8 * (long long)h_ncls[(size_t)i];
            if (i > h_r0[(size_t)b] && (acc + h_cost[(size_t)i] / 2 > target || state + sw > cap)) {       // row i opens the next range
                b++;
                h_r0[(size_t)b] = i;
                left -= acc;
                acc = 0; state = 0;
                target = left / (unsigned long long)(B - b);
            }
            acc += h_cost[(size_t)i];
            state += sw;
        }
        for (int bb = b + 1; bb <= B; bb++) h_r0[(size_t)bb] = P;
        CU(cudaMemcpyAsync(m.blk_row0, h_r0.data(), ((size_t)B + 1) * 4, cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));
    }
    k_sort_keys<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(P, B, m.blk_row0, d_degn, d_key, d_val);
    CU(cub::DeviceRadixSort::SortPairs(d_cub, cub_bytes, d_key, d_key2, d_val, d_perm, P, 0, 44, st));
    k_apply_perm<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(P, d_perm, d_tn, d_degn, s->d_Rs, s->d_A, in.d_pos, d_degp, m.row_RsA, d_rown, d_rsan);
    ctx->launches += 12;
    // ---- classes: (owner CTA, cardinality) cells -> compact ids, tile geometry ----
    CU(cudaMemsetAsync(d_cell_cnt, 0, (size_t)(n_cells + 1) * 4, st));
    k_class_cells<<<(unsigned)((nm + 255) / 256), 256, 0, st>>>(nm, T, n_kseg, B, ix->d_kseg_cid0, d_owner, in.d_act, in.d_newid, in.d_pos, m.blk_row0, d_cellof, d_cell_cnt,
                                                               d_pa, (int32_t *)d_pc);
    k_ps_cell_sizes<<<(unsigned)((n_cells + 1 + 255) / 256), 256, 0, st>>>(n_cells, n_kseg, ix->d_kseg_k, d_cell_cnt, d_cell_u16, d_cell_tiles);
    CU(cub::DeviceScan::ExclusiveSum(d_cub, cub_bytes, d_cell_cnt, d_clsbase, n_cells + 1, st));
    CU(cub::DeviceScan::ExclusiveSum(d_cub, cub_bytes, d_cell_u16, d_u16base, n_cells + 1, st));
    CU(cub::DeviceScan::ExclusiveSum(d_cub, cub_bytes, d_cell_tiles, d_tilebase, n_cells + 1, st));
    k_block_tables<<<(unsigned)((B + 1 + 255) / 256), 256, 0, st>>>(B, n_kseg, d_clsbase, d_tilebase, m.blk_cls0, m.blk_etile0);
    CU(cub::DeviceRadixSort::SortPairs(d_cub, cub_bytes, d_pa, d_pb, (int32_t *)d_pc, (int32_t *)d_uniq_e, (int)C_a, 0, 64, st));
    k_class_newid<<<(unsigned)((C_a + 255) / 256), 256, 0, st>>>(C_a, (const int32_t *)d_uniq_e, d_newid2);
    ctx->launches += 12;
    uint32_t e_u16 = 0; int32_t n_etiles = 0;
    CU(cudaMemcpyAsync(&e_u16, d_u16base + n_cells, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&n_etiles, d_tilebase + n_cells, 4, cudaMemcpyDeviceToHost, st));
    // ---- halo rows: every row a CTA's classes touch outside its own range ----
    CU(cudaMemsetAsync(d_hcount, 0, 32, st));
    k_halo_collect_e<<<(unsigned)((nm * 32 + 255) / 256), 256, 0, st>>>(nm, T, B, ix->d_cls_off, ix->d_cls_tid, d_owner, in.d_act, in.d_pos, m.blk_row0, d_pa, d_hcount);
    LAUNCHED(ctx);
    unsigned int h_cnt = 0, n_ue = 0;
    CU(cudaMemcpyAsync(&h_cnt, d_hcount, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (h_cnt > 0) {
        CU(cub::DeviceRadixSort::SortKeys(d_cub, cub_bytes, d_pa, d_pb, (int)h_cnt, 0, 44, st));
        CU(cub::DeviceSelect::Unique(d_cub, cub_bytes, d_pb, d_uniq_e, d_hcount + 1, (int)h_cnt, st));
        ctx->launches += 4;
        CU(cudaMemcpyAsync(&n_ue, d_hcount + 1, 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    TRY(ps_alloc(s, &m.halo_rows, (size_t)n_ue + 1)); TRY(ps_alloc(s, &m.halo_tgt, (size_t)n_ue + 1));
    m.n_inc = (int32_t)n_ue;
    if (n_ue >= (1u << 28)) { ps_note("%u shared rows", n_ue); ps_free(s); return EMSAR_OK; }
    k_halo_ranges<<<(unsigned)((B + 1 + 255) / 256), 256, 0, st>>>(B, n_ue, d_uniq_e, m.blk_hr0);
    if (n_ue) k_halo_list<<<(n_ue + 255) / 256, 256, 0, st>>>(n_ue, d_uniq_e, m.halo_rows);
    ctx->launches += 2;
    std::vector<int32_t> h_row0(B + 1), h_cls0(B + 1), h_et0(B + 1), h_hr0(B + 1);
    CU(cudaMemcpyAsync(h_row0.data(), m.blk_row0, (size_t)(B + 1) * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_cls0.data(), m.blk_cls0, (size_t)(B + 1) * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_et0.data(), m.blk_etile0, (size_t)(B + 1) * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_hr0.data(), m.blk_hr0, (size_t)(B + 1) * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    for (int b = 0; b < B; b++) {
        const int nrows = h_row0[b + 1] - h_row0[b], nhr = h_hr0[b + 1] - h_hr0[b], ncls = h_cls0[b + 1] - h_cls0[b];
        if (nrows + nhr + 1 > PS_MAX_SLOT || ncls + 1 > PS_MAX_SLOT) {       // 16-bit slots do not reach: legacy layout
            ps_note("CTA %d: %d rows + %d halo rows, %d classes exceed 16-bit slots", b, nrows, nhr, ncls);
            ps_free(s); return EMSAR_OK;
        }
        if (ps_smem_plan(0, 0, 0, nrows, nhr, ncls, nhr, 0).total + 1024 > ctx->em_smem_bytes) {
            ps_note("CTA %d: state of %d rows + %d halo rows, %d classes = %d bytes does not fit in %d bytes of shared memory", b, nrows, nhr, ncls,
                    ps_smem_plan(0, 0, 0, nrows, nhr, ncls, nhr, 0).total, ctx->em_smem_bytes);
            ps_free(s); return EMSAR_OK;
        }
    }
    // ---- incidences: partial-sum slots per row, and where every halo slot sends its contribution ----
    if (n_ue > 0) {
        unsigned long long *d_ik = d_pa, *d_isk = d_pb;                                   // both free here: the halo keys have been reduced to d_uniq_e
        int32_t *d_iv = (int32_t *)d_pc, *d_isv = (int32_t *)d_pc + (nnz_a + 2);          // n_ue <= nnz_a
        k_ps_inc_keys<<<(n_ue + 255) / 256, 256, 0, st>>>(n_ue, d_uniq_e, d_ik, d_iv);
        CU(cub::DeviceRadixSort::SortPairs(d_cub, cub_bytes, d_ik, d_isk, d_iv, d_isv, (int)n_ue, 0, 64, st));
        CU(cudaMemsetAsync(d_hcount + 4, 0, 8, st));
        k_ps_inc_tgt<<<(n_ue + 255) / 256, 256, 0, st>>>(n_ue, d_isk, d_isv, m.blk_row0, B, B_local, m.halo_tgt, m.row_mask, (unsigned long long *)(d_hcount + 4), in.rank);
        k_ps_inc_off<<<(unsigned)((P + 1 + 255) / 256), 256, 0, st>>>(P, n_ue, d_isk, m.inc_off);
        ctx->launches += 6;
    } else CU(cudaMemsetAsync(m.inc_off, 0, ((size_t)P + 2) * 4, st));
    // ---- E side: tile descriptors, member slots, read counts; one key per member for the M side ----
    uint16_t *d_edata = nullptr;
    TRY(ps_alloc(s, &d_edata, (size_t)e_u16 + 64));
    m.e_data = (const unsigned char *)d_edata;
    TRY(ps_alloc(s, &m.e_tiles, (size_t)n_etiles + 1)); TRY(ps_alloc(s, &m.e_src, (size_t)n_etiles + 1));
    CU(cudaMemsetAsync(d_edata, 0, ((size_t)e_u16 + 64) * 2, st));          // lanes of a partial tile that carry no class gather slot 0 and drop the result
    k_ps_etiles<<<(unsigned)((n_etiles + 255) / 256), 256, 0, st>>>(n_etiles, n_cells, n_kseg, ix->d_kseg_k, d_tilebase, d_clsbase, d_cell_cnt, d_u16base, m.blk_cls0, m.e_tiles);
    k_ps_pack_classes<<<(unsigned)((nm * 32 + 255) / 256), 256, 0, st>>>(nm, T, n_kseg, m.blk_hr0, d_uniq_e, ix->d_kseg_k, ix->d_cls_off, ix->d_cls_tid, in.d_act, in.d_newid,
                                                                         d_cellof, d_newid2, d_clsbase, d_u16base, s->d_R, in.d_pos, m.blk_row0, m.blk_cls0, d_apos,
                                                                         d_edata, d_eR, d_pa);
    ctx->launches += 2;
    // ---- M side: members sorted by (CTA, row slot, class) -> touched rows -> items ----
    int cta_bits = 1;
    while ((1 << cta_bits) < B) cta_bits++;
    CU(cub::DeviceRadixSort::SortKeys(d_cub, cub_bytes, d_pa, d_pb, (int)nnz_a, 0, 32 + cta_bits, st));
    uint32_t *d_flag = (uint32_t *)d_pa, *d_tidx = (uint32_t *)d_pa + (nnz_a + 2);       // d_pa is free again (nnz_a + 2 words each)
    k_ps_heads<<<(unsigned)((nnz_a + 1 + 255) / 256), 256, 0, st>>>((int64_t)nnz_a, d_pb, d_flag);
    CU(cub::DeviceScan::ExclusiveSum(d_cub, cub_bytes, d_flag, d_tidx, (int)(nnz_a + 1), st));
    ctx->launches += 5;
    uint32_t n_tr = 0;
    CU(cudaMemcpyAsync(&n_tr, d_tidx + nnz_a, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    // per touched row: key, start; sorted order inside the CTA
    char *tscr = nullptr;
    const size_t trb = rnd(((size_t)n_tr + 2) * 8);
    TRY(dev_alloc(&tscr, 8 * trb + 4096));
    struct Guard2 { char *p, *q; ~Guard2() { dev_free(p); dev_free(q); } } guard2{tscr, nullptr};
    char *tc = tscr;
    unsigned long long *d_trkey = arena_take<unsigned long long>(tc, (size_t)n_tr + 2), *d_k2 = arena_take<unsigned long long>(tc, (size_t)n_tr + 2);
    unsigned long long *d_sk2 = arena_take<unsigned long long>(tc, (size_t)n_tr + 2);
    uint32_t *d_trstart = arena_take<uint32_t>(tc, (size_t)n_tr + 2), *d_tdeg = arena_take<uint32_t>(tc, (size_t)n_tr + 2);
    int32_t *d_v2 = arena_take<int32_t>(tc, (size_t)n_tr + 2), *d_tperm = arena_take<int32_t>(tc, (size_t)n_tr + 2);
    uint32_t *d_rowbase = arena_take<uint32_t>(tc, (size_t)n_tr + 2);
    int32_t *d_rowitem = arena_take<int32_t>(tc, (size_t)n_tr + 2), *d_rowidx = arena_take<int32_t>(tc, (size_t)n_tr + 2);
    uint32_t *d_size16 = arena_take<uint32_t>(tc, (size_t)n_tr + 2), *d_off16 = arena_take<uint32_t>(tc, (size_t)n_tr + 2);
    {   // the CUB scratch of the caller is sized for the legacy packer's sorts: make sure the two sorts over touched rows / incidences fit
        size_t need1 = 0, need2 = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, need1, (unsigned long long *)nullptr, (unsigned long long *)nullptr, (int32_t *)nullptr, (int32_t *)nullptr, (int)n_tr + 1, 0, 64);
        cub::DeviceScan::ExclusiveSum(nullptr, need2, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)n_tr + 2);
        if (std::max(need1, need2) > cub_bytes) {
            char *bigger = nullptr;
            TRY(dev_alloc(&bigger, std::max(need1, need2) + 256));
            guard2.q = bigger;
            d_cub = bigger; cub_bytes = std::max(need1, need2);
        }
    }
    k_ps_touched<<<(unsigned)((nnz_a + 255) / 256), 256, 0, st>>>((int64_t)nnz_a, d_pb, d_flag, d_tidx, n_tr, d_trkey, d_trstart);
    // EMSAR_PS_MORDER=1: remote-owner rows first in the M-phase. Measured (profiles/r2p): no gain on one GPU (21.8 us both ways at config #2,
    // 167.2 vs 165.0 us at config #5: the mixed slice costs padding), so the plain longest-first order is the default
    static const int remote_first = []() { const char *e = getenv("EMSAR_PS_MORDER"); return e ? atoi(e) : 0; }();
    k_ps_touched_keys<<<(n_tr + 255) / 256, 256, 0, st>>>(n_tr, d_trkey, d_trstart, m.blk_row0, remote_first, d_k2, d_v2);
    CU(cub::DeviceRadixSort::SortPairs(d_cub, cub_bytes, d_k2, d_sk2, d_v2, d_tperm, (int)n_tr, 0, 32 + cta_bits, st));
    k_ps_sorted_deg<<<(n_tr + 1 + 255) / 256, 256, 0, st>>>(n_tr, d_sk2, d_tdeg);
    k_halo_ranges<<<(unsigned)((B + 1 + 255) / 256), 256, 0, st>>>(B, n_tr, d_sk2, d_tr0);
    k_ps_items_count<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(B, d_tr0, d_tdeg, d_nlong, d_nitems, d_ngroups);
    k_block_items_prefix<<<1, 32, 0, st>>>(B, d_nitems, m.blk_mitem0);
    ctx->launches += 10;
    int32_t n_mitems = 0;
    CU(cudaMemcpyAsync(&n_mitems, m.blk_mitem0 + B, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if ((size_t)n_mitems + 1 > (size_t)n_tr + 2) { emsar_set_err("internal: more M items than touched rows"); return EMSAR_ERR_STATE; }
    TRY(ps_alloc(s, &m.m_items, (size_t)n_mitems + 1)); TRY(ps_alloc(s, &m.m_src, (size_t)n_mitems + 1));
    k_ps_items_fill<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(B, d_tr0, d_tdeg, d_nlong, m.blk_mitem0, d_size16, m.m_items, d_rowbase, d_rowitem, d_rowidx);
    CU(cub::DeviceScan::ExclusiveSum(d_cub, cub_bytes, d_size16, d_off16, n_mitems + 1, st));
    if (n_mitems > 0) k_ps_item_offsets<<<(unsigned)((n_mitems + 255) / 256), 256, 0, st>>>(n_mitems, d_off16, m.m_items);
    ctx->launches += 3;
    uint32_t m16 = 0;
    CU(cudaMemcpyAsync(&m16, d_off16 + n_mitems, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    uint16_t *d_mdata = nullptr;
    TRY(ps_alloc(s, &d_mdata, (size_t)m16 * 8 + 64));
    m.m_data = (const unsigned char *)d_mdata;
    if (n_mitems > 0) {
        k_ps_item_init<<<(unsigned)(((int64_t)n_mitems * 32 + 255) / 256), 256, 0, st>>>(n_mitems, B, m.blk_mitem0, m.blk_cls0, m.m_items, d_mdata);
        k_ps_scatter_rows<<<(unsigned)(((int64_t)n_tr * 32 + 255) / 256), 256, 0, st>>>(n_tr, B, d_sk2, d_tperm, d_trkey, d_trstart, d_pb, d_tr0, d_nlong, d_ngroups, m.blk_mitem0,
                                                                                     d_rowbase, d_rowitem, d_rowidx, m.m_items, m.blk_row0, m.blk_hr0, m.halo_tgt, d_mdata);
        ctx->launches += 2;
    }
    // ---- host: shared-memory plan of every CTA, resident index cache ----
    std::vector<int4> h_et((size_t)n_etiles), h_mi((size_t)n_mitems);
    std::vector<int32_t> h_mi0(B + 1), h_inc((size_t)P + 2);
    CU(cudaMemcpyAsync(h_inc.data(), m.inc_off, ((size_t)P + 1) * 4, cudaMemcpyDeviceToHost, st));
    if (n_etiles) CU(cudaMemcpyAsync(h_et.data(), m.e_tiles, (size_t)n_etiles * 16, cudaMemcpyDeviceToHost, st));
    if (n_mitems) CU(cudaMemcpyAsync(h_mi.data(), m.m_items, (size_t)n_mitems * 16, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_mi0.data(), m.blk_mitem0, (size_t)(B + 1) * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    std::vector<int32_t> h_esrc((size_t)n_etiles + 1), h_msrc((size_t)n_mitems + 1), h_desc(B + 1, 0), h_res16(B + 1, 0);
    int64_t resident16 = 0, rows_long = 0;
    for (int i = 0; i < n_etiles; i++) h_esrc[(size_t)i] = h_et[(size_t)i].z;
    for (int i = 0; i < n_mitems; i++) { h_msrc[(size_t)i] = h_mi[(size_t)i].z; if ((h_mi[(size_t)i].w >> 30) & 1) rows_long += h_mi[(size_t)i].y; }
    // look-ahead staging of the items that are not resident (E tiles: bit 0, M items: bit 1; EMSAR_PS_STAGE overrides). It costs 36 KB of
    // shared memory per CTA, so it is dropped when the state would not fit beside it.
    int stage = 0;      // measured (profiles/r2h): routing the index words through shared memory costs more LSU work than the hidden latency is worth
    for (int pass = 0; pass < 2; pass++) {
        bool ok = true;
        for (int b = 0; b < B && ok; b++) {
            const int nrows = h_row0[b + 1] - h_row0[b], nhr = h_hr0[b + 1] - h_hr0[b], ncls = h_cls0[b + 1] - h_cls0[b];
            const int nin = h_inc[(size_t)h_row0[b + 1]] - h_inc[(size_t)h_row0[b]];
            if (ctx->em_smem_bytes - ps_smem_plan(0, 0, 0, nrows, nhr, ncls, nin, stage).total - 256 < 8 * 1024) ok = false;
        }
        if (ok || stage == 0) break;
        stage = 0;
    }
    m.stage = stage;
    bool fits = true;
    for (int b = 0; b < B && fits; b++) {
        const int nrows = h_row0[b + 1] - h_row0[b], nhr = h_hr0[b + 1] - h_hr0[b], ncls = h_cls0[b + 1] - h_cls0[b];
        const int n_et = h_et0[b + 1] - h_et0[b], n_mi = h_mi0[b + 1] - h_mi0[b];
        const int nin = h_inc[(size_t)h_row0[b + 1]] - h_inc[(size_t)h_row0[b]];        // partial sums it receives per iteration (staged in shared memory)
        int desc = 1;
        int left = ctx->em_smem_bytes - ps_smem_plan(1, n_et, n_mi, nrows, nhr, ncls, nin, stage).total - 256;
        if (left < 16 * 1024) {         // many tiles (high cardinalities): descriptors stay in global memory, the space goes to the state
            desc = 0;
            left = ctx->em_smem_bytes - ps_smem_plan(0, n_et, n_mi, nrows, nhr, ncls, nin, stage).total - 256;
        }
        if (left < 0) { ps_note("CTA %d: %d rows + %d halo rows, %d classes, %d incoming partial sums, %d tiles, %d items: %d bytes short", b, nrows, nhr, ncls, nin, n_et, n_mi, -left); fits = false; break; }
        h_desc[b] = desc;
        // resident index cache: the data of the items with the longest dependent chains stays in shared memory for the whole kernel
        struct Cand { int prio, n16, idx; bool e; };
        std::vector<Cand> cand;
        for (int i = h_et0[b]; i < h_et0[b + 1]; i++) {
            const int4 t = h_et[(size_t)i];
            const int steps = t.w & 0xfff, lg = (t.w >> 12) & 0xf;
            const int d16 = (lg == 0 && steps <= 4) ? 32 : 16 * ((steps + 3) >> 2) * max(1, (t.w >> 16) & 0xff);
            cand.push_back({(lg == 0 && steps <= 4) ? 2 : steps, d16 + (t.y * 4 + 15) / 16, i, true});       // index data + read counts
        }
        for (int i = h_mi0[b]; i < h_mi0[b + 1]; i++) {
            const int4 t = h_mi[(size_t)i];
            const int len = t.w & 0x1fffffff;
            cand.push_back({((t.w >> 30) & 1) ? t.y * 6 + len / 64 : len, t.x, i, false});
        }
        std::stable_sort(cand.begin(), cand.end(), [](const Cand &x, const Cand &y) { return x.prio > y.prio; });
        int used = 0;
        const int cap16 = left / 16;
        for (const Cand &cd : cand) {
            if (used + cd.n16 > cap16) continue;
            if (cd.e) { h_et[(size_t)cd.idx].z = used; h_et[(size_t)cd.idx].w |= 1 << 30; }
            else { h_mi[(size_t)cd.idx].z = used; h_mi[(size_t)cd.idx].w |= 1 << 29; }
            used += cd.n16;
        }
        h_res16[b] = used;
        resident16 += used;
    }
    if (!fits) { ps_free(s); return EMSAR_OK; }
    if (n_etiles) CU(cudaMemcpyAsync(m.e_tiles, h_et.data(), (size_t)n_etiles * 16, cudaMemcpyHostToDevice, st));
    if (n_mitems) CU(cudaMemcpyAsync(m.m_items, h_mi.data(), (size_t)n_mitems * 16, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(m.e_src, h_esrc.data(), ((size_t)n_etiles + 1) * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(m.m_src, h_msrc.data(), ((size_t)n_mitems + 1) * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(m.blk_desc_smem, h_desc.data(), (size_t)(B + 1) * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(m.blk_res16, h_res16.data(), (size_t)(B + 1) * 4, cudaMemcpyHostToDevice, st));
    // ---- exchange slots: theta [P + 1] | partial sums [n_inc + 1] | convergence [2 * Bt] ----
    unsigned long long peer_stores = 0;
    {
        m.th_off = 0;
        m.part_off = 16 * ((long long)P + 1);
        m.dm_off = m.part_off + 16 * ((long long)n_ue + 1);
        const size_t need = (size_t)m.dm_off + 16 * (2 * (size_t)B + 8);
        if (in.nranks > 1) {
            // one sample over several GPUs: the slots live in a window every rank maps into the others (comm.cu); collective call
            if (n_ue) CU(cudaMemcpyAsync(&peer_stores, d_hcount + 4, 8, cudaMemcpyDeviceToHost, st));
            TRY(comm_window_ensure(ctx, 0, need));
            if (ctx->win_state != 1) { ps_note("no peer memory between the ranks"); ps_free(s); return EMSAR_OK; }          // the NCCL path of the legacy kernel
            s->ps_row_lo = h_row0[in.rank * B_local]; s->ps_row_hi = h_row0[(in.rank + 1) * B_local];
        } else {
            if (need > s->slots_bytes) {
                if (s->d_slots) dev_free(s->d_slots);
                s->d_slots = nullptr;
                char *psl = nullptr;
                TRY(dev_alloc(&psl, need + (need >> 3)));
                s->d_slots = psl;
                s->slots_bytes = need + (need >> 3);
                CU(cudaMemsetAsync(s->d_slots, 0, s->slots_bytes, st));
                s->slot_tag = 0;
            }
            s->ps_row_lo = 0; s->ps_row_hi = P;
        }
    }
    k_fill_double<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(m.theta, P, 1.0);
    LAUNCHED(ctx);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(st));           // the host vectors go out of scope
    s->use_psum = true;
    emsar_model_stats &ms_ = s->stats;
    memset(&ms_, 0, sizeof(ms_));
    ms_.T = T; ms_.C_a = C_a; ms_.nnz_a = (int64_t)nnz_a;
    ms_.rows_short = P; ms_.rows_long = rows_long; ms_.rows_hub = 0; ms_.rows_fixed = T - P;
    ms_.e_tiles = n_etiles; ms_.m_tiles = n_mitems;
    ms_.bytes_per_iter = 8 * (int64_t)nnz_a + 24 * C_a + 44 * (int64_t)T;
    ms_.index_bytes = 2 * (int64_t)e_u16 + 4 * C_a + 16 * (int64_t)m16;
    // what the kernel streams per iteration: index data + read counts, {Rs, A} and the exchange slots of shared rows
    ms_.stream_bytes_per_iter = ms_.index_bytes + 16 * (int64_t)P + 3 * 16 * (int64_t)n_ue;
    ms_.em_variant = 5; ms_.all_local = 1; ms_.halo_rows = n_ue; ms_.halo_classes = 0;
    ms_.resident_index_bytes = 16 * resident16;
    ms_.peer_bytes_per_iter = in.nranks > 1 ? 16 * (int64_t)peer_stores + 16 * (int64_t)B_local * (in.nranks - 1) : 0;
    return EMSAR_OK;
}
