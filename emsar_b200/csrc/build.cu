// Index construction on the device: the read-sharing classes of a transcriptome and their per-fragment-length counts.
// Replaces, in the reference (parklab/emsar v2.0.1, src/emsar_functions.c): the suffix-array sorts initialize_suffixarray_* / sort_* (:949-1230),
// the run scans construct_rshbucket_2 (:1758-1816) and construct_rshbucket_PE_3 (:1902-1974), and the mate clustering
// process_mate1_cluster_by_mate_3 (:2784-2934). Only the RESULT of those is specified - which transcripts share a read (fragment), how often -
// so nothing here is a suffix array:
//   1. every read-length window of the concatenated transcriptome gets a 128-bit polynomial hash (two 64-bit hashes with different bases);
//   2. occurrences (SE: one per forward position, represented by the smaller of itself and its reverse complement when the library is
//      unstranded; PE: one per admissible (mate 1, mate 2) pair, with the reference's flip rule) are keyed by the hash of their bases and
//      radix-sorted by (hash, tid) - three stable CUB passes over a permutation;
//   3. runs of equal keys are the reference's runs of equal substrings: they are VERIFIED base by base against their predecessor (a hash
//      collision is an error, never a wrong class), then turned into singleton counts (atomic int adds) or into class candidates;
//   4. class candidates (sorted tid multisets) are hashed, sorted and counted the same way, and verified tid by tid.
// The host (emsar_b200/host/build_index.c) folds the unique classes of every read length into its class store and orders them for print_rsh.
// Paired-end candidate sets larger than the entry buffer are processed in partitions of the mate-1 hash (equal fragments share a partition).
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

#include "common.cuh"

namespace {

constexpr uint64_t HB1 = 0x100000001B3ULL * 31 + 2;          // odd multipliers of the two polynomial hashes (mod 2^64)
constexpr uint64_t HB2 = 0x9E3779B97F4A7C15ULL;
constexpr long long BUILD_CAP_DEFAULT = 1LL << 27;           // entries per partition (56 bytes each)

__device__ __forceinline__ uint64_t base_code(char ch) { return ch == 'A' ? 1u : ch == 'C' ? 2u : ch == 'G' ? 3u : ch == 'T' ? 4u : 0u; }
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t h) { h ^= h >> 33; h *= 0xff51afd7ed558ccdULL; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL; h ^= h >> 33; return h; }

// H1[p], H2[p] = hashes of S[p, p + L), ok[p] = 1 where the window is pure ACGT (mark_noncanonical :2642-2660) and inside the text
__global__ void k_window_hash(int64_t n, int L, const char *__restrict__ S, uint64_t *__restrict__ H1, uint64_t *__restrict__ H2, uint8_t *__restrict__ ok)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    uint64_t h1 = 0, h2 = 0;
    bool good = p + L <= n;
    for (int j = 0; j < L && good; j++) {
        const uint64_t c = base_code(S[p + j]);
        if (!c) good = false;
        h1 = h1 * HB1 + c;
        h2 = h2 * HB2 + c;
    }
    H1[p] = good ? h1 : 0; H2[p] = good ? h2 : 0; ok[p] = good ? 1 : 0;
}

__device__ __forceinline__ int dev_memcmp(const char *a, const char *b, int L)
{
    for (int j = 0; j < L; j++) { const int d = (int)(unsigned char)a[j] - (int)(unsigned char)b[j]; if (d) return d; }
    return 0;
}

// transcript of a forward position (start[t] <= i < start[t + 1])
__device__ __forceinline__ int tid_of(const int64_t *__restrict__ start, int T, int64_t i)
{
    int lo = 0, hi = T - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (start[mid] <= i) lo = mid; else hi = mid - 1; }
    return lo;
}

struct Entries {
    uint64_t *k1, *k2;
    int32_t *tid, *d;
    int64_t *pos;
    unsigned long long *count;     // entries appended so far
    long long cap;
    int *overflow;
};

__device__ __forceinline__ void append(const Entries &e, uint64_t k1, uint64_t k2, int tid, int d, int64_t pos)
{
    const unsigned long long j = atomicAdd(e.count, 1ULL);
    if ((long long)j >= e.cap) { *e.overflow = 1; return; }
    e.k1[j] = k1; e.k2[j] = k2; e.tid[j] = tid; e.d[j] = d; e.pos[j] = pos;
}

// SE occurrences (initialize_suffixarray_NS_5 :1001-1027 / _SS): one per forward window, at its canonical position
__global__ void k_occ_se(int64_t border, int64_t end, int L, int stranded, int T, const char *__restrict__ S, const int64_t *__restrict__ start,
                         const uint64_t *__restrict__ H1, const uint64_t *__restrict__ H2, const uint8_t *__restrict__ ok, Entries e)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= border || !ok[i]) return;
    int64_t p = i;
    if (!stranded) {
        const int64_t fl = end - i - L;
        if (dev_memcmp(S + i, S + fl, L) > 0) p = fl;
    }
    append(e, H1[p], H2[p], tid_of(start, T, i), 0, p);
}

// PE occurrences (process_mate1_cluster_by_mate_3 :2852-2874): thread = (forward window i, strand copy s); mate 2 at every admissible distance
template <bool COUNT_ONLY>
__global__ void k_occ_pe(int64_t border, int64_t end, int L, int stranded, int T, int dmin, int dmax, int part, int nparts, uint64_t pw1, uint64_t pw2,
                         const char *__restrict__ S, const int64_t *__restrict__ start, const uint64_t *__restrict__ H1, const uint64_t *__restrict__ H2,
                         const uint8_t *__restrict__ ok, Entries e, unsigned long long *__restrict__ part_count)
{
    const int reps = stranded ? 1 : 2;
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = g / reps;
    const int s = (int)(g - i * reps);
    if (i >= border || !ok[i]) return;
    const int64_t p = s == 0 ? i : end - i - L;
    const uint64_t h1p = H1[p];
    const int mypart = nparts > 1 ? (int)(((unsigned __int128)mix64(h1p) * (unsigned)nparts) >> 64) : 0;
    if (!COUNT_ONLY && mypart != part) return;
    const int t = tid_of(start, T, i);
    int64_t lo, hi;                                  // mate 2 must start in [lo, hi]: the member's transcript, in the half it lies in
    if (p < border) { lo = start[t]; hi = start[t + 1] - 1 - L; }
    else { lo = end - (start[t + 1] - 1); hi = end - start[t] - L; }
    const uint64_t h2p = H2[p];
    unsigned long long found = 0;
    for (int d = dmin; d <= dmax; d++) {
        const int64_t q = p + d;
        if (q < lo || q > hi || !ok[q]) continue;
        if (!stranded) {
            const char *a = S + p, *b = S + (end - q - L);
            int cr = dev_memcmp(a, b, L);            // strcmp_pe :2674-2678
            if (!cr) cr = dev_memcmp(a + d, b + d, L);
            if (!((p < border && cr <= 0) || (p > border && cr < 0))) continue;
        }
        if (COUNT_ONLY) found++;
        else append(e, h1p * pw1 + H1[q], h2p * pw2 + H2[q], t, d, p);
    }
    if (COUNT_ONLY && found) atomicAdd(part_count + mypart, found);
}

__global__ void k_iota(int64_t n, uint32_t *__restrict__ idx)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) idx[j] = (uint32_t)j;
}
template <class TT> __global__ void k_gather(int64_t n, const uint32_t *__restrict__ idx, const TT *__restrict__ src, TT *__restrict__ dst)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) dst[j] = src[idx[j]];
}
__global__ void k_gather_tid(int64_t n, const uint32_t *__restrict__ idx, const int32_t *__restrict__ src, uint32_t *__restrict__ dst)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) dst[j] = (uint32_t)src[idx[j]];
}

// sorted order: heads of runs; every other entry is compared with its predecessor base by base (both mates for PE)
__global__ void k_run_heads(int64_t n, int L, int pe, const char *__restrict__ S, const uint32_t *__restrict__ idx, const uint64_t *__restrict__ k1s,
                            const uint64_t *__restrict__ k2, const int64_t *__restrict__ pos, const int32_t *__restrict__ d, uint8_t *__restrict__ head,
                            int *__restrict__ collision)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    if (j == 0) { head[0] = 1; return; }
    const uint32_t a = idx[j], b = idx[j - 1];
    const bool same = k1s[j] == k1s[j - 1] && k2[a] == k2[b];
    head[j] = same ? 0 : 1;
    if (same) {
        const char *x = S + pos[a], *y = S + pos[b];
        bool eq = dev_memcmp(x, y, L) == 0;
        if (eq && pe) eq = dev_memcmp(x + d[a], y + d[b], L) == 0;
        if (!eq) *collision = 1;
    }
}

// one thread per run (construct_rshbucket_2 / construct_rshbucket_PE_3): r = 1 -> the singleton count of (tid, d); 1 < r < max_repeat (PE: all
// members at one distance) -> class candidate with a 128-bit key over (r, tids); else dropped
__global__ void k_runs(int64_t R, int64_t n, int pe, int max_repeat, int dmin, int nD, const int64_t *__restrict__ run_start, const uint32_t *__restrict__ idx,
                       const int32_t *__restrict__ tid, const int32_t *__restrict__ d, int32_t *__restrict__ single, uint8_t *__restrict__ is_class,
                       uint64_t *__restrict__ ck1, uint64_t *__restrict__ ck2)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const int64_t a = run_start[r], b = r + 1 < R ? run_start[r + 1] : n;
    const int64_t len = b - a;
    is_class[r] = 0;
    const int d0 = d[idx[a]];
    if (len == 1) { atomicAdd(&single[(int64_t)tid[idx[a]] * nD + (d0 - dmin)], 1); return; }
    if (len >= max_repeat) return;
    uint64_t h1 = 0x9E3779B97F4A7C15ULL * (uint64_t)len, h2 = 0xC2B2AE3D27D4EB4FULL + (uint64_t)len;
    for (int64_t j = a; j < b; j++) {
        const uint32_t e = idx[j];
        if (pe && d[e] != d0) return;                 // multi_d (:1925-1928)
        const uint64_t t = (uint64_t)(uint32_t)tid[e];
        h1 = mix64(h1 ^ (t + 0x632BE59BD9B4E019ULL * (uint64_t)(j - a + 1)));
        h2 = (h2 + t) * HB1 + 0x51ED270B7F4A7C15ULL;
    }
    is_class[r] = 1;
    ck1[r] = h1; ck2[r] = mix64(h2);
}

__global__ void k_class_d(int64_t nc, const int64_t *__restrict__ crun, const int64_t *__restrict__ run_start, const uint32_t *__restrict__ idx,
                          const int32_t *__restrict__ d, int dmin, uint32_t *__restrict__ key)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < nc) key[c] = (uint32_t)(d[idx[run_start[crun[c]]]] - dmin);
}
template <class TT> __global__ void k_gather_via(int64_t n, const int64_t *__restrict__ via, const TT *__restrict__ src, TT *__restrict__ dst)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) dst[j] = src[via[j]];
}
template <class TT> __global__ void k_gather64(int64_t n, const uint32_t *__restrict__ idx, const int64_t *__restrict__ via, const TT *__restrict__ src, TT *__restrict__ dst)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) dst[j] = src[via[idx[j]]];
}

// class candidates in sorted order (d, key): heads of groups of equal classes; a non-head is compared with its predecessor tid by tid
__global__ void k_class_heads(int64_t nc, int64_t n, int64_t R, const uint32_t *__restrict__ cidx, const int64_t *__restrict__ crun, const uint32_t *__restrict__ dkey,
                              const uint64_t *__restrict__ ck1, const uint64_t *__restrict__ ck2, const int64_t *__restrict__ run_start,
                              const uint32_t *__restrict__ idx, const int32_t *__restrict__ tid, uint8_t *__restrict__ head, uint32_t *__restrict__ klen,
                              int *__restrict__ collision)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nc) return;
    const int64_t ra = crun[cidx[c]];
    const int64_t a0 = run_start[ra], a1 = ra + 1 < R ? run_start[ra + 1] : n;
    klen[c] = 0;
    if (c == 0) { head[0] = 1; klen[0] = (uint32_t)(a1 - a0); return; }
    const int64_t rb = crun[cidx[c - 1]];
    const bool same = dkey[cidx[c]] == dkey[cidx[c - 1]] && ck1[ra] == ck1[rb] && ck2[ra] == ck2[rb];
    head[c] = same ? 0 : 1;
    if (!same) { klen[c] = (uint32_t)(a1 - a0); return; }
    const int64_t b0 = run_start[rb], b1 = rb + 1 < R ? run_start[rb + 1] : n;
    bool eq = (a1 - a0) == (b1 - b0);
    for (int64_t j = 0; eq && j < a1 - a0; j++) eq = tid[idx[a0 + j]] == tid[idx[b0 + j]];
    if (!eq) *collision = 1;
}

// one thread per unique class: its tids, its distance index and the number of runs that produced it
__global__ void k_class_out(int64_t nu, int64_t nc, int64_t n, int64_t R, const int64_t *__restrict__ uhead, const uint32_t *__restrict__ cidx,
                            const int64_t *__restrict__ crun, const uint32_t *__restrict__ dkey, const int64_t *__restrict__ run_start,
                            const uint32_t *__restrict__ idx, const int32_t *__restrict__ tid, const uint32_t *__restrict__ off, int32_t *__restrict__ out_tid,
                            int32_t *__restrict__ out_d, int32_t *__restrict__ out_count)
{
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= nu) return;
    const int64_t c0 = uhead[u], c1 = u + 1 < nu ? uhead[u + 1] : nc;
    const int64_t ra = crun[cidx[c0]];
    const int64_t a0 = run_start[ra], a1 = ra + 1 < R ? run_start[ra + 1] : n;
    for (int64_t j = a0; j < a1; j++) out_tid[off[u] + (j - a0)] = tid[idx[a0 + (j - a0)]];
    out_d[u] = (int32_t)dkey[cidx[c0]];
    out_count[u] = (int32_t)(c1 - c0);
}

struct DevBufs {
    std::vector<void *> p;
    ~DevBufs() { for (void *q : p) dev_free(q); }
    template <class TT> int take(TT **out, size_t n) { TRY(dev_alloc(out, n)); p.push_back((void *)*out); return EMSAR_OK; }
};

struct HostOut {
    std::vector<int32_t> single;
    std::vector<int64_t> class_off;
    std::vector<int32_t> class_tid, class_d, class_count;
};

inline unsigned nblk(int64_t n, int b = 256) { return (unsigned)((n + b - 1) / b); }

// sort the permutation idx[0..n) by key[idx] (stable), bits [0, end_bit)
template <class K>
int sort_by(emsar_ctx *ctx, int64_t n, const K *key_src, K *key_a, K *key_b, uint32_t *&idx, uint32_t *&idx_alt, int end_bit, void *tmp, size_t tmp_bytes)
{
    cudaStream_t st = ctx->stream;
    k_gather<K><<<nblk(n), 256, 0, st>>>(n, idx, key_src, key_a);
    CU(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, key_a, key_b, idx, idx_alt, (int)n, 0, end_bit, st));
    std::swap(idx, idx_alt);
    ctx->launches += 2;
    return EMSAR_OK;
}

} // namespace

// one pass over the partition's entries: sort, runs, singles, unique classes -> appended to `ho`
static int build_partition(emsar_ctx *ctx, const emsar_build_desc *bd, const char *d_S, const Entries &e, int64_t n, int nD, int32_t *d_single, int *d_flags, HostOut &ho,
                           int64_t *n_runs_out)
{
    cudaStream_t st = ctx->stream;
    if (n == 0) return EMSAR_OK;
    if (n >= (1LL << 31) - 1) { emsar_set_err("index construction: %lld occurrences in one partition (limit 2^31)", (long long)n); return EMSAR_ERR_BAD_ARG; }
    DevBufs B;
    uint32_t *idx = nullptr, *idx2 = nullptr, *k32a = nullptr, *k32b = nullptr;
    uint64_t *k64a = nullptr, *k64b = nullptr;
    uint8_t *head = nullptr;
    int64_t *run_start = nullptr, *d_num = nullptr;
    TRY(B.take(&idx, (size_t)n)); TRY(B.take(&idx2, (size_t)n)); TRY(B.take(&k32a, (size_t)n)); TRY(B.take(&k32b, (size_t)n));
    TRY(B.take(&k64a, (size_t)n)); TRY(B.take(&k64b, (size_t)n)); TRY(B.take(&head, (size_t)n + 1));
    TRY(B.take(&run_start, (size_t)n + 1)); TRY(B.take(&d_num, 4));
    size_t tmp_bytes = 0, t2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (uint64_t *)nullptr, (uint64_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)n, 0, 64);
    cub::DeviceSelect::Flagged(nullptr, t2, thrust::counting_iterator<int64_t>(0), (uint8_t *)nullptr, (int64_t *)nullptr, (int64_t *)nullptr, (int)n);
    tmp_bytes = std::max(tmp_bytes, t2) + 256;
    void *tmp = nullptr;
    { char *t = nullptr; TRY(B.take(&t, tmp_bytes)); tmp = t; }
    int tbits = 1;
    while ((1LL << tbits) < (long long)bd->T) tbits++;
    k_iota<<<nblk(n), 256, 0, st>>>(n, idx);
    // stable passes, least significant key first: tid, second hash, first hash
    k_gather_tid<<<nblk(n), 256, 0, st>>>(n, idx, e.tid, k32a);
    CU(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k32a, k32b, idx, idx2, (int)n, 0, tbits, st));
    std::swap(idx, idx2);
    TRY(sort_by<uint64_t>(ctx, n, e.k2, k64a, k64b, idx, idx2, 64, tmp, tmp_bytes));
    TRY(sort_by<uint64_t>(ctx, n, e.k1, k64a, k64b, idx, idx2, 64, tmp, tmp_bytes));
    const uint64_t *k1s = k64b;                      // first hash in sorted order
    k_run_heads<<<nblk(n), 256, 0, st>>>(n, bd->readlen, bd->pe, d_S, idx, k1s, e.k2, e.pos, e.d, head, d_flags + 1);
    CU(cub::DeviceSelect::Flagged(tmp, tmp_bytes, thrust::counting_iterator<int64_t>(0), head, run_start, d_num, (int)n, st));
    ctx->launches += 5;
    int64_t R = 0;
    CU(cudaMemcpyAsync(&R, d_num, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (n_runs_out) *n_runs_out += R;
    uint8_t *is_class = nullptr;
    uint64_t *ck1 = nullptr, *ck2 = nullptr;
    int64_t *crun = nullptr;
    TRY(B.take(&is_class, (size_t)R + 1)); TRY(B.take(&ck1, (size_t)R + 1)); TRY(B.take(&ck2, (size_t)R + 1)); TRY(B.take(&crun, (size_t)R + 1));
    k_runs<<<nblk(R), 256, 0, st>>>(R, n, bd->pe, bd->max_repeat, bd->d_min, nD, run_start, idx, e.tid, e.d, d_single, is_class, ck1, ck2);
    CU(cub::DeviceSelect::Flagged(tmp, tmp_bytes, thrust::counting_iterator<int64_t>(0), is_class, crun, d_num, (int)R, st));
    ctx->launches += 2;
    int64_t nc = 0;
    CU(cudaMemcpyAsync(&nc, d_num, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (nc == 0) return EMSAR_OK;
    // class candidates: sort by (distance, key2, key1), count equal classes
    uint32_t *cidx = idx2, *cidx2 = k32a, *dkey = nullptr, *dk_a = k32b, *dk_b = nullptr, *klen = nullptr, *koff = nullptr;   // idx2, k32a, k32b are free again (nc <= n)
    int64_t *uhead = nullptr;
    TRY(B.take(&dkey, (size_t)nc + 1)); TRY(B.take(&dk_b, (size_t)nc + 1)); TRY(B.take(&klen, (size_t)nc + 1)); TRY(B.take(&koff, (size_t)nc + 2)); TRY(B.take(&uhead, (size_t)nc + 1));
    k_class_d<<<nblk(nc), 256, 0, st>>>(nc, crun, run_start, idx, e.d, bd->d_min, dkey);
    k_iota<<<nblk(nc), 256, 0, st>>>(nc, cidx);
    k_gather64<uint64_t><<<nblk(nc), 256, 0, st>>>(nc, cidx, crun, ck1, k64a);
    CU(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k64a, k64b, cidx, cidx2, (int)nc, 0, 64, st));
    std::swap(cidx, cidx2);
    k_gather64<uint64_t><<<nblk(nc), 256, 0, st>>>(nc, cidx, crun, ck2, k64a);
    CU(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k64a, k64b, cidx, cidx2, (int)nc, 0, 64, st));
    std::swap(cidx, cidx2);
    int dbits = 1;
    while ((1 << dbits) < nD) dbits++;
    k_gather<uint32_t><<<nblk(nc), 256, 0, st>>>(nc, cidx, dkey, dk_a);
    CU(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, dk_a, dk_b, cidx, cidx2, (int)nc, 0, dbits, st));
    std::swap(cidx, cidx2);
    k_class_heads<<<nblk(nc), 256, 0, st>>>(nc, n, R, cidx, crun, dkey, ck1, ck2, run_start, idx, e.tid, head, klen, d_flags + 2);
    CU(cub::DeviceSelect::Flagged(tmp, tmp_bytes, thrust::counting_iterator<int64_t>(0), head, uhead, d_num, (int)nc, st));
    ctx->launches += 10;
    int64_t nu = 0;
    CU(cudaMemcpyAsync(&nu, d_num, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    // tid offsets of the unique classes: lengths of the heads, compacted in head order, then scanned
    uint32_t *ulen = nullptr;
    int32_t *o_tid = nullptr, *o_d = nullptr, *o_count = nullptr;
    TRY(B.take(&ulen, (size_t)nu + 2)); TRY(B.take(&o_d, (size_t)nu + 1)); TRY(B.take(&o_count, (size_t)nu + 1));
    k_gather_via<uint32_t><<<nblk(nu), 256, 0, st>>>(nu, uhead, klen, ulen);
    CU(cudaMemsetAsync(ulen + nu, 0, 4, st));
    CU(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, ulen, koff, (int)(nu + 1), st));
    uint32_t total = 0;
    CU(cudaMemcpyAsync(&total, koff + nu, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    TRY(B.take(&o_tid, (size_t)total + 1));
    k_class_out<<<nblk(nu), 256, 0, st>>>(nu, nc, n, R, uhead, cidx, crun, dkey, run_start, idx, e.tid, koff, o_tid, o_d, o_count);
    ctx->launches += 3;
    const size_t u0 = ho.class_d.size(), t0 = ho.class_tid.size();
    std::vector<uint32_t> h_off((size_t)nu + 1);
    ho.class_d.resize(u0 + (size_t)nu); ho.class_count.resize(u0 + (size_t)nu); ho.class_tid.resize(t0 + (size_t)total);
    CU(cudaMemcpyAsync(h_off.data(), koff, ((size_t)nu + 1) * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(ho.class_d.data() + u0, o_d, (size_t)nu * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(ho.class_count.data() + u0, o_count, (size_t)nu * 4, cudaMemcpyDeviceToHost, st));
    if (total) CU(cudaMemcpyAsync(ho.class_tid.data() + t0, o_tid, (size_t)total * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    for (int64_t u = 0; u < nu; u++) ho.class_off.push_back((int64_t)t0 + (int64_t)h_off[(size_t)u + 1]);
    return EMSAR_OK;
}

struct emsar_build_owner { HostOut ho; };

extern "C" void emsar_build_classes_free(emsar_build_classes *out)
{
    if (!out) return;
    delete (emsar_build_owner *)out->owner;
    memset(out, 0, sizeof(*out));
}

extern "C" int emsar_build_classes_run(emsar_ctx *ctx, const emsar_build_desc *bd, emsar_build_classes *out)
{
    CHECK_ARG(ctx && bd && out, "emsar_build_classes_run: NULL argument");
    CHECK_ARG(bd->seq && bd->start && bd->T > 0 && bd->border > 0 && bd->end == 2 * bd->border + 1, "emsar_build_classes_run: bad transcriptome layout");
    CHECK_ARG(bd->readlen >= 1 && bd->max_repeat >= 1, "emsar_build_classes_run: bad read length / max_repeat");
    CHECK_ARG(!bd->pe || (bd->d_min >= 0 && bd->d_max >= bd->d_min), "emsar_build_classes_run: bad mate distance range");
    memset(out, 0, sizeof(*out));
    TRY(ctx_use(ctx));
    cudaStream_t st = ctx->stream;
    const int64_t n = bd->end + 1, border = bd->border;
    const int L = bd->readlen, T = bd->T;
    const int dmin = bd->pe ? bd->d_min : 0, dmax = bd->pe ? bd->d_max : 0, nD = dmax - dmin + 1;
    DevBufs B;
    char *d_S = nullptr; int64_t *d_start = nullptr; uint64_t *H1 = nullptr, *H2 = nullptr; uint8_t *ok = nullptr;
    int32_t *d_single = nullptr; int *d_flags = nullptr; unsigned long long *d_count = nullptr, *d_pc = nullptr;
    TRY(B.take(&d_S, (size_t)n + 1)); TRY(B.take(&d_start, (size_t)T + 1)); TRY(B.take(&H1, (size_t)n)); TRY(B.take(&H2, (size_t)n)); TRY(B.take(&ok, (size_t)n));
    TRY(B.take(&d_single, (size_t)T * nD)); TRY(B.take(&d_flags, 4)); TRY(B.take(&d_count, 1));
    CU(cudaEventRecord(ctx->ev0, st));
    CU(cudaMemcpyAsync(d_S, bd->seq, (size_t)n, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_start, bd->start, ((size_t)T + 1) * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(d_single, 0, (size_t)T * nD * 4, st));
    CU(cudaMemsetAsync(d_flags, 0, 16, st));
    k_window_hash<<<nblk(n), 256, 0, st>>>(n, L, d_S, H1, H2, ok);
    LAUNCHED(ctx);
    uint64_t pw1 = 1, pw2 = 1;
    for (int j = 0; j < L; j++) { pw1 *= HB1; pw2 *= HB2; }
    const long long cap_env = getenv("EMSAR_BUILD_CAP") ? atoll(getenv("EMSAR_BUILD_CAP")) : BUILD_CAP_DEFAULT;
    const long long cap_lim = cap_env > 1024 ? cap_env : 1024;
    const int reps = bd->stranded ? 1 : 2;
    // partitions: the SE list has at most `border` entries; the PE candidates are counted per partition of the mate-1 hash first
    int nparts = 1;
    long long cap = 0;
    std::vector<unsigned long long> h_pc;
    if (!bd->pe) {
        if (border >= (1LL << 31) - 1) { emsar_set_err("index construction: transcriptome of %lld bases (limit 2^31)", (long long)border); return EMSAR_ERR_BAD_ARG; }
        cap = border;
    } else {
        const long double worst = (long double)border * reps * nD;
        nparts = (int)std::min<long double>(65536.0L, worst / (long double)cap_lim + 1.0L);
        TRY(B.take(&d_pc, (size_t)nparts));
        CU(cudaMemsetAsync(d_pc, 0, (size_t)nparts * 8, st));
        Entries none{};
        k_occ_pe<true><<<nblk(border * reps), 256, 0, st>>>(border, bd->end, L, bd->stranded, T, dmin, dmax, 0, nparts, pw1, pw2, d_S, d_start, H1, H2, ok, none, d_pc);
        LAUNCHED(ctx);
        h_pc.resize((size_t)nparts);
        CU(cudaMemcpyAsync(h_pc.data(), d_pc, (size_t)nparts * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        for (unsigned long long c : h_pc) cap = std::max<long long>(cap, (long long)c);
    }
    Entries e{};
    e.cap = cap > 0 ? cap : 1;
    TRY(B.take(&e.k1, (size_t)e.cap)); TRY(B.take(&e.k2, (size_t)e.cap)); TRY(B.take(&e.tid, (size_t)e.cap)); TRY(B.take(&e.d, (size_t)e.cap)); TRY(B.take(&e.pos, (size_t)e.cap));
    e.count = d_count; e.overflow = d_flags;
    emsar_build_owner *own = new emsar_build_owner();
    struct Guard { emsar_build_owner *o; ~Guard() { delete o; } } guard{own};
    HostOut &ho = own->ho;
    ho.class_off.push_back(0);
    int64_t occ = 0, runs = 0;
    for (int part = 0; part < nparts; part++) {
        if (bd->pe && h_pc[(size_t)part] == 0) continue;
        CU(cudaMemsetAsync(d_count, 0, 8, st));
        if (!bd->pe) k_occ_se<<<nblk(border), 256, 0, st>>>(border, bd->end, L, bd->stranded, T, d_S, d_start, H1, H2, ok, e);
        else k_occ_pe<false><<<nblk(border * reps), 256, 0, st>>>(border, bd->end, L, bd->stranded, T, dmin, dmax, part, nparts, pw1, pw2, d_S, d_start, H1, H2, ok, e, nullptr);
        LAUNCHED(ctx);
        unsigned long long cnt = 0;
        int flags[4] = {0, 0, 0, 0};
        CU(cudaMemcpyAsync(&cnt, d_count, 8, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(flags, d_flags, 16, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (flags[0] || (long long)cnt > e.cap) { emsar_set_err("index construction: internal error (partition %d holds %llu occurrences, counted %lld)", part, cnt, (long long)e.cap); return EMSAR_ERR_STATE; }
        occ += (int64_t)cnt;
        TRY(build_partition(ctx, bd, d_S, e, (int64_t)cnt, nD, d_single, d_flags, ho, &runs));
    }
    int flags[4] = {0, 0, 0, 0};
    ho.single.resize((size_t)T * nD);
    CU(cudaMemcpyAsync(flags, d_flags, 16, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(ho.single.data(), d_single, (size_t)T * nD * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaEventRecord(ctx->ev1, st));
    CU(cudaStreamSynchronize(st));
    float dev_ms = 0;
    CU(cudaEventElapsedTime(&dev_ms, ctx->ev0, ctx->ev1));
    if (flags[1] || flags[2]) {
        emsar_set_err("index construction: two different %s share a 128-bit hash (EMSAR_BUILD_HOST=1 selects the host builder)", flags[1] ? "substrings" : "classes");
        return EMSAR_ERR_STATE;
    }
    out->T = T; out->n_d = nD;
    out->single_count = ho.single.data();
    out->n_class = (int64_t)ho.class_d.size();
    out->class_off = ho.class_off.data();
    out->class_tid = ho.class_tid.data();
    out->class_d = ho.class_d.data();
    out->class_count = ho.class_count.data();
    out->occurrences = occ; out->runs = runs; out->partitions = nparts; out->device_ms = dev_ms;
    out->owner = own;
    guard.o = nullptr;
    return EMSAR_OK;
}
