// Internal declarations shared by the translation units of libemsar_cuda.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "emsar_cuda.h"

void emsar_set_err(const char *fmt, ...);

#define CU(call)                                                                                    \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            emsar_set_err("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));    \
            return EMSAR_ERR_CUDA;                                                                  \
        }                                                                                           \
    } while (0)

#define CHECK_ARG(cond, ...)                                                                        \
    do {                                                                                            \
        if (!(cond)) { emsar_set_err(__VA_ARGS__); return EMSAR_ERR_BAD_ARG; }                      \
    } while (0)

#define TRY(call)                                                                                   \
    do { int rc_ = (call); if (rc_ != EMSAR_OK) return rc_; } while (0)

// ---- tunables of the EM kernels ---------------------------------------------------------------------
#ifndef EM_BLOCK
#define EM_BLOCK 1024         // threads per CTA of the persistent EM kernel (one CTA per SM owns a row range)
#endif
#ifndef EM_MIN_BLOCKS
#define EM_MIN_BLOCKS 1       // launch-bounds hint: resident CTAs per SM
#endif
constexpr int EM_WARPS = EM_BLOCK / 32;
// E side: G lanes share a class (G = 1, 4, 16, 32 by cardinality) so that a lane's dependent chain stays <= ~16 members;
// a tile is one 32-lane row block of 32/G classes stored transposed (step j of every lane is one 128-byte line), classes
// shorter than steps*G are padded with the zero-theta slot
constexpr int KT = 16;            // cardinality <= KT: one thread per class
__host__ __device__ __forceinline__ int e_lgG(int k) { return k <= KT ? 0 : (k <= 64 ? 2 : (k <= 256 ? 4 : 5)); }
__host__ __device__ __forceinline__ int e_steps(int k) { const int lg = e_lgG(k); return (k + (1 << lg) - 1) >> lg; }
// classes per tile: small cardinalities are grouped (k=2: 4 row blocks, k=3,4: 2) so that a lane keeps ~8 loads in flight
__host__ __device__ __forceinline__ int e_cls_per_tile(int k) { return k == 2 ? 128 : (k <= 4 ? 64 : (32 >> e_lgG(k))); }
__host__ __device__ __forceinline__ int e_cls_per_block(int k) { return 32 >> e_lgG(k); }
constexpr int M_LONG = 64;        // transposed rows with more active entries: one warp per row; the rest go into SELL-32 slices
constexpr int M_GROUP_ENTRIES = 768;  // a group of long rows (<= M_GROUP_ROWS rows, one warp) holds at most this many entries unless a single row is longer
constexpr int M_GROUP_ROWS = 8;
constexpr int CH_INTS = 4096;     // ints per staged chunk of an index stream (indices + read counts): 16 KB per pipeline stage
constexpr int CH_BYTES = CH_INTS * 4 + 64;
constexpr int NSTAGE = 4;         // stages of the TMA pipeline (producer warp -> consumer warps)

constexpr int EMSAR_MAX_RANKS = 8;     // one NVSwitch domain

struct KSeg {          // multi-tid classes of one cardinality, contiguous in cid order
    int32_t k;
    int64_t cid0, cid1;
};

struct emsar_ctx {
    int device;
    cudaStream_t stream;
    cudaStream_t copy_stream;  // H2D of large read batches, overlapped with the counting kernel
    cudaEvent_t copy_ev[9];
    cudaDeviceProp prop;
    void *pool;               // cudaMemPool_t of the stream-ordered allocations (NULL: plain cudaMalloc / cudaFree)
    int64_t launches;
    int em_blocks_per_sm;
    int em_minb;              // launch-bounds variant of the EM kernel in use
    int em_smem_bytes;        // dynamic shared memory per CTA (theta | q slices)
    unsigned *d_barrier;      // [0] arrival count, [1] generation   (persistent-kernel grid barrier)
    void *d_scratch;          // CUB temp storage (grow-only)
    size_t scratch_bytes;
    cudaEvent_t ev0, ev1;
    cudaEvent_t tev0, tev1;   // emsar_cuda_timer_start / _stop
    size_t l2_persist_bytes;
    // multi-GPU (class-sharded samples): NCCL communicator loaded at run time
    void *nccl_comm;
    int rank, nranks;
    // peer-memory window of the fused class-sharded EM kernel (comm.cu): one cudaMalloc block per rank, mapped into every
    // other rank with CUDA IPC (or plain peer access inside one process). Layout: [flags | dmax | theta_nat | xbuf].
    void *win;                 // this rank's window
    void *peer_win[EMSAR_MAX_RANKS];   // every rank's window as seen from this device (peer_win[rank] == win)
    bool peer_ipc[EMSAR_MAX_RANKS];    // mapping opened with cudaIpcOpenMemHandle (must be closed)
    int64_t win_rows;          // capacity: participating rows
    size_t win_bytes;
    int win_state;             // 0 = not tried, 1 = usable, -1 = peer memory unavailable (NCCL path only)
};
// Window layout (bytes): [dm: 2 parities x nranks x WIN_MAX_CTAS slots][theta: 2 parities x win_rows slots][xbuf: nranks x S slots].
// A slot is 16 bytes: two 64-bit words {low half of the double | tag << 32}, {high half | tag << 32}. Each word is written
// with one 8-byte store, so a reader that polls until both tags match has the value - no fence, no separate flag.
constexpr int WIN_MAX_CTAS = 256;
constexpr size_t WIN_HDR_BYTES = 2 * (size_t)EMSAR_MAX_RANKS * WIN_MAX_CTAS * 16;
__host__ __device__ __forceinline__ int64_t win_slice_rows(int64_t rows, int nranks) { return (rows + nranks - 1) / nranks; }

struct emsar_index {
    emsar_ctx *ctx;
    int32_t T;
    int64_t C, nnz, n_multi, nnz_multi;
    int32_t nF, min_fl, max_fl, readlength, max_t_size, frag_min, frag_max, max_card;
    // device
    uint32_t *d_cls_off;   // [C+1]
    int32_t *d_cls_tid;    // [nnz]
    int32_t *d_euma;       // [C*nF]
    uint8_t *d_has_node;   // [C]
    uint32_t *d_txm_off;   // [T+1]   transpose of the multi-tid classes, ascending cid, multiplicity kept
    int32_t *d_txm_cid;    // [nnz_multi]
    int32_t *d_order;      // [T] locality order: connected components kept together (the "natural order" of the EM rows follows it)
    unsigned long long *d_hash; // open addressing: (fingerprint << 32) | (cid + 1), 0 = empty
    uint64_t hash_mask;
    int64_t hash_inserted;
    int64_t *d_kseg_cid0;  // [n_kseg+1] first cid of each cardinality segment (last = C)
    int32_t *d_kseg_k;     // [n_kseg]
    std::vector<KSeg> kseg;
    // host copies kept for the set decomposition (emsar_main.c:411-425)
    std::vector<uint32_t> h_cls_off;
    std::vector<int32_t> h_cls_tid;
    std::vector<int32_t> h_order;
    std::vector<uint32_t> h_txm_off;       // kept for emsar_index_aux_get (the packed image)
    std::vector<int32_t> h_txm_cid;
    std::vector<uint8_t> h_insertable;
    int32_t n_sets_nocut, max_set_tids;
    int64_t device_bytes;
};

// per-sample packed model the EM kernel streams (all device pointers).
// Ownership: the participating rows (transcripts with A_t > 0) are cut, in natural order, into B contiguous ranges of
// equal cost, one per CTA of the persistent kernel; a class belongs to the CTA that owns its FIRST member row. Each
// CTA keeps theta of its rows and q of (as many as fit of) its classes in shared memory; references that leave the
// CTA's range ("halo") go through the global copies. Indices are pre-encoded: >= 0 = slot in the owner's shared memory,
// < 0 = ~(global index).
// Both directions are stored as SELL-32 slices: inside a CTA the rows are sorted by length, 32 rows form a slice stored
// transposed and padded to its longest row (one thread per row, sequential sum); classes are sorted by cardinality
// (no padding needed), 32 classes form a tile stored transposed (one thread per class).
struct EmModel {
    int32_t T, P, B;
    int64_t C_a, nnz_a;
    int32_t *blk_row0;     // [B+1] first row of each CTA
    int32_t *blk_cls0;     // [B+1] first (compact) class of each CTA
    int32_t *blk_etile0;   // [B+1]
    int32_t *blk_mitem0;   // [B+1]
    int32_t *blk_ech0, *blk_mch0;  // [B+1] chunk ranges
    int4 *e_chunks, *m_chunks;     // {first item, end item, stream offset (ints), ints}; ints < 0: not staged (oversized item)
    int32_t *e_res, *m_res;        // per E tile / M item: int offset of its resident copy in the CTA's shared-memory index cache, -1 = none
    int32_t *blk_res_ints;         // [B] ints of the CTA's resident index cache
    int32_t direct;                // 1: direct mode (index cache, no staging pipeline)
    int32_t all_local;             // 1: every halo row / class and every q of every CTA has a shared-memory slot
    int32_t *blk_nres;     // [B]   classes of the CTA whose q lives in shared memory (slot nres holds 0.0: padding target)
    // halo: distinct remote rows / classes a CTA references; copied into its shared memory at the start of each phase
    int32_t *blk_hr0, *blk_hc0;   // [B+1] ranges in halo_rows / halo_cls
    int32_t *blk_nhr, *blk_nhc;   // [B]   how many of them got a shared-memory slot (the rest falls back to global loads)
    int32_t *halo_rows;    // global row of each halo slot
    int32_t *halo_cls;     // global compact class of each halo slot
    int32_t smem_bytes;    // dynamic shared memory per CTA
    // E side
    int32_t *e_tid;        // encoded member rows (tile-transposed for k<=KT, row-major otherwise)
    uint32_t *e_R;         // [C_a] read count | bit31: q must also be stored to global (halo / not resident)
    int4 *e_tiles;         // {j0, cnt, tid_off, steps | log2(G) << 12}
    int32_t n_etiles;
    // M side
    int32_t *m_cls;        // encoded classes: slices transposed + padded, long rows row-major
    int4 *m_items;         // {first row slot, rows, entry offset, length | mode<<30}: mode 0 = slice (length = longest row),
                           // 1 = group of long rows (length = all entries; a header of `rows` lengths precedes them)
    double2 *row_RsA;      // [P] {Rs, A}
    int32_t *row_n;        // [P] natural index (rank among the participating transcripts) of row p: the row order differs
                           // from rank to rank in sharded mode, the natural order does not, so sums are exchanged in it
    double2 *rsa_nat;      // [P] {Rs, A} in natural order (sharded mode: the owner of a slice updates theta from it)
    int32_t n_mitems;
    int64_t m_ints;        // entries stored (with padding)
    // state (global copies; the shared-memory copies are loaded from / written through to these)
    double *theta;         // [P]
    double *q;             // [C_a]
};

#include "psum_model.cuh"

struct emsar_sample {
    emsar_index *index;
    emsar_ctx *ctx;
    // counts
    int32_t *d_R;          // [C]
    int32_t *d_hist;       // [max_fl+1]
    int32_t *d_flags;      // [0] error flags raised by kernels
    bool have_counts;
    // staging for host read batches (grow-only)
    void *d_rd_ptr, *d_rd_tid, *d_rd_fl, *d_rd_aux;
    size_t cap_rd_ptr, cap_rd_tid, cap_rd_fl, cap_rd_aux;
    cudaEvent_t count_ev[4];   // "host arrays of batch k are free again" (emsar_sample_count_wait)
    unsigned count_seq;
    // model
    bool prepared;
    int64_t N;
    double eumacut, delta;
    int32_t max_sid;
    double *d_Wf;          // [nF]
    double *d_adj;         // [C]
    double *d_amodel;      // [C]  EUMAps if the class is modelled else 0
    uint8_t *d_in_model;   // [C]
    std::vector<int32_t> h_CS; // set ids (filled when computed)
    bool have_CS;
    double *d_A, *d_Rs, *d_iE; // [T]
    uint8_t *d_lone;       // [T] 1 = no in-model multi-tid class contains t (the lone-singleton set of MLE())
    int32_t *d_pos;        // [T] permuted row of transcript t, -1 when A_t == 0
    double *d_state;       // theta | q (one allocation: one L2 access-policy window)
    size_t state_bytes;
    void *d_pack;          // packed model arena
    size_t pack_bytes;
    int32_t *d_mcls;       // M-side entries (sized after the slices are known)
    size_t mcls_bytes;
    int32_t *d_halo;       // halo_rows | halo_cls
    size_t halo_bytes;
    void *d_chunks;        // chunk tables
    size_t chunk_bytes;
    unsigned long long *d_trace;   // tuning aid
    void *d_slots;         // barrier-free EM kernel: tagged 16-byte slots of theta [P+1] and q [C_a+1]
    size_t slots_bytes;
    bool force_legacy;     // the psum model was not eligible for a sharded sample: pack the legacy one
    unsigned slot_tag;     // tags handed out so far (every launch takes a fresh range, so the slots are never cleared)
    double *d_qpart;       // sharded mode, NCCL path: [2*P] partial / reduced per-row sums in natural order
    bool sharded;
    EmModel m;
    // class-owner-centric model of k_em_psum (psum.cuh): used instead of `m` whenever the sample is eligible
    bool use_psum;
    PsModel ps;
    std::vector<void *> ps_allocs;
    int32_t ps_row_lo, ps_row_hi;  // class-sharded: the rows this rank's CTAs own
    emsar_model_stats stats;
    // solve bookkeeping
    int32_t n_iter;
    double final_delta;
    double em_ms, prep_ms;
    emsar_solve_opts opts;
};

// ---- helpers implemented across the .cu files ---------------------------------------------------
int ctx_scratch(emsar_ctx *ctx, size_t bytes, void **p);
// Device memory of the current context (ctx_use): stream-ordered (cudaMallocFromPoolAsync / cudaFreeAsync on the context's
// stream) from a pool that keeps what it is given back, so the second sample of a run allocates nothing from the driver.
int ctx_use(emsar_ctx *ctx);
int dev_alloc_bytes(void **p, size_t bytes);
void dev_free(void *p);
template <class T> static inline int dev_alloc(T **p, size_t n)
{
    void *q = nullptr;
    int rc = dev_alloc_bytes(&q, (n ? n : 1) * sizeof(T));
    *p = (T *)q;
    return rc;
}
#define LAUNCHED(ctx) ((ctx)->launches++)

int index_build_hash(emsar_index *ix, const std::vector<uint8_t> &insertable);
int sample_build_model(emsar_sample *s);
int em_psum_attr(emsar_ctx *ctx);
int em_psum_launch(emsar_sample *s, int max_iter, int stop_on_conv, int *iters_done, double *final_delta, double *ms_out);
int em_launch(emsar_sample *s, int max_iter, int stop_on_conv, int *iters_done, double *final_delta, double *ms, bool fused = false);
int em_query_occupancy(emsar_ctx *ctx);
int sample_finalize_device(emsar_sample *s, emsar_solve_out *out);
int comm_allreduce_f64(emsar_ctx *ctx, const double *in, double *out, size_t n);
int comm_allreduce_i32(emsar_ctx *ctx, int32_t *inout, size_t n);
int comm_barrier(emsar_ctx *ctx);
int comm_window_ensure(emsar_ctx *ctx, int64_t rows, size_t min_bytes = 0);      // collective; leaves ctx->win_state at 1 (peer memory) or -1
void comm_window_release(emsar_ctx *ctx);

// ---- device helpers ------------------------------------------------------------------------------
#ifdef __CUDACC__
// Shared-memory plan of one CTA of k_em_persistent:
// [stage buffer 0][stage buffer 1][E tile descriptors][M items][E chunks][M chunks][halo row list][halo class list]
// [{Rs,A} of its rows][theta: own rows | halo rows][q: resident classes | 0.0 | halo classes]
struct SmemPlan { int off_etiles, off_mitems, off_ech, off_mch, off_hrl, off_hcl, off_rsa, off_theta, off_q, off_res, total; };
// `stage_bytes`: NSTAGE * CH_BYTES in pipelined mode, 0 in direct mode; `n_res_tab`: tiles + items with a resident-cache entry table;
// in direct mode the chunk-table slots (n_ech) hold the CTA's super-tile rounds instead
__host__ __device__ __forceinline__ SmemPlan em_smem_plan(int stage_bytes, int n_et, int n_mi, int n_ech, int n_mch, int n_res_tab, int nrows, int nhr, int nres, int nhc)
{
    SmemPlan p;
    p.off_etiles = stage_bytes;
    p.off_mitems = p.off_etiles + n_et * 16;
    p.off_ech = p.off_mitems + n_mi * 16;
    p.off_mch = p.off_ech + n_ech * 16;
    p.off_hrl = p.off_mch + n_mch * 16;
    p.off_hcl = p.off_hrl + nhr * 4;
    p.off_rsa = (p.off_hcl + nhc * 4 + n_res_tab * 4 + 15) & ~15;      // the resident-offset tables sit right after the halo lists
    p.off_theta = p.off_rsa + nrows * 16;
    p.off_q = p.off_theta + (nrows + nhr + 1) * 8;       // + the zero-theta slot (padding target of the E tiles)
    p.off_res = p.off_q + (nres + 1 + nhc) * 8;
    p.total = p.off_res;
    return p;
}
// chunks a CTA can need at most for a stream of `ints` ints cut greedily at CH_INTS (every two consecutive chunks hold > CH_INTS)
__host__ __device__ __forceinline__ int em_max_chunks(long long ints, int n_items) { long long c = 2 * (ints / CH_INTS) + 2; return (int)(c < n_items + 1 ? c : n_items + 1); }
__device__ __forceinline__ uint64_t mix64(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
// order-sensitive, but summable in any order: h(key) = mix(sum_i elem(i, tid_i) + k*phi)
__device__ __forceinline__ uint64_t key_elem(int i, int tid) { return mix64(((uint64_t)(uint32_t)tid << 32) | (uint32_t)i); }
__device__ __forceinline__ uint64_t key_finish(uint64_t sum, int k) { return mix64(sum + (uint64_t)k * 0x9E3779B97F4A7C15ULL); }
#endif
