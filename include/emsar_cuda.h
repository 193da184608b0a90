/*
 * emsar_cuda.h — C ABI of libemsar_cuda.so: the B200 (sm_100a) replacement for EMSAR's per-sample
 * quantification hot path.  Plain C, plain pointers and sizes; no CUDA or torch types in any signature.
 *
 * The reference (parklab/emsar v2.0.1, paths relative to its src/) has no plugin API; its only designed
 * seams are three global function pointers (emsar.h:219-221) and the global arrays the estimator reads and
 * writes (emsar.h:90-195).  Each entry point below names the reference code it replaces, so that the
 * reference's own main loop (emsar_main.c:380-488) can call this library instead.
 *
 * Conventions
 *   - every function returns an int status: EMSAR_OK (0) or an EMSAR_ERR_* code; emsar_cuda_strerror()
 *     names the code and emsar_cuda_last_error() gives the detail of the calling thread's last failure.
 *     (The reference prints to stderr and exit(1)s; the host program turns a non-zero status into that.)
 *   - host buffers stay owned by the caller; device memory lives behind the opaque handles.
 *   - a context is bound to ONE device and ONE CUDA stream; handles are not thread-safe: use one host
 *     thread (or one process) per device.  There is no CPU fallback: without a usable sm_100 device
 *     emsar_cuda_open() fails with EMSAR_ERR_NO_DEVICE.
 *   - class ids follow the reference's scan order (scan_rshbucket, emsar_functions.c:2149-2191): cid
 *     0..T-1 are the singleton classes (cid == tid), multi-tid classes follow ordered by cardinality,
 *     then first tid, then chain (file) order.  A class is a sorted multiset of tids.
 */
#ifndef EMSAR_CUDA_H
#define EMSAR_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EMSAR_OK 0
#define EMSAR_ERR_NO_DEVICE 1   /* no CUDA device / not compute capability 10.x */
#define EMSAR_ERR_CUDA 2        /* a CUDA runtime call failed (see emsar_cuda_last_error) */
#define EMSAR_ERR_BAD_ARG 3     /* NULL / out-of-range argument */
#define EMSAR_ERR_BAD_INDEX 4   /* class table not in the reference's scan order / malformed */
#define EMSAR_ERR_UNSUPPORTED 5 /* a documented limit was exceeded (e.g. a read with > 1024 alignments) */
#define EMSAR_ERR_STATE 6       /* call order violated (e.g. solve before any counts) */
#define EMSAR_ERR_NOMEM 7       /* host or device allocation failed */
#define EMSAR_ERR_COMM 8        /* NCCL / multi-GPU failure */

#define EMSAR_MAX_READ_TIDS 1024 /* longest tid list one read may carry (reference: MAX_REPEAT, default 100) */

typedef struct emsar_ctx emsar_ctx;       /* one device + stream */
typedef struct emsar_index emsar_index;   /* packed rsh index resident on the device */
typedef struct emsar_sample emsar_sample; /* per-alignment-file state: counts, model, estimates */

/* -------- context ------------------------------------------------------------------------------- */
int emsar_cuda_open(int device, emsar_ctx **ctx);
int emsar_cuda_close(emsar_ctx *ctx);
const char *emsar_cuda_strerror(int status);
const char *emsar_cuda_last_error(void);
/* number of kernels this library has launched on the context so far (bench.py's gpu_launches) */
int emsar_cuda_launch_count(emsar_ctx *ctx, int64_t *launches);
int emsar_cuda_synchronize(emsar_ctx *ctx);

typedef struct {
    int32_t sm_count;
    int32_t cc_major, cc_minor;
    int64_t l2_bytes;
    int64_t hbm_bytes;
    int32_t em_blocks_per_sm; /* resident CTAs/SM of the persistent EM kernel */
    int32_t em_block_threads;
    char name[64];
} emsar_device_info;
int emsar_cuda_device_info(emsar_ctx *ctx, emsar_device_info *info);

/* -------- index ---------------------------------------------------------------------------------
 * Replaces the product of construct_rsh_from_rshfile (emsar_functions.c:1351-1378): rshbucket,
 * rshbucket_single (emsar.h:139-145), initialize_rshbucket (:1334-1347), delete_rshbucket (:1687-1723). */
typedef struct {
    int32_t T;                 /* max_tid + 1 */
    int64_t C;                 /* max_cid + 1 (T singleton classes first) */
    const int64_t *class_ptr;  /* [C+1] CSR offsets, class_ptr[c] == c for c <= T */
    const int32_t *class_tid;  /* [class_ptr[C]] sorted within a class, duplicates kept */
    int32_t nF;                /* nFraglen = Fraglengths.max - Fraglengths.min + 1 (:2471-2475) */
    const int32_t *euma;       /* [C * nF] row-major EUMA counts (node1.EUMA, emsar.h:79) */
    const uint8_t *has_node;   /* [C] 0 = singleton line without EUMA: no node exists (:1486), NULL = all 1 */
    int32_t min_fraglength;    /* rsh header field 3 -> Min_Fraglength (:1419) */
    int32_t max_fraglength;    /* rsh header field 4 -> Max_Fraglength (:1420) */
    int32_t readlength;        /* rsh header field 5, -1 for SE (:1421) */
    int32_t max_t_size;        /* rsh header field 2 -> rshbucket_max_t_size (:1418) */
    const struct emsar_index_aux *aux; /* optional: the derived structures from a packed image (below); NULL = derive them here */
} emsar_index_desc;

/* What emsar_index_create derives from the class table - the transpose of the multi-tid classes (build_TC_from_CT_2 :2201-2227), which
 * classes the chain walk can reach (hash insertion), the set statistics without EUMAcut and the locality order of the transcripts. A packed
 * index image carries them so that loading an index costs no pass over its members: emsar_index_aux_get hands out
 * the arrays of a created index (valid until it is destroyed), emsar_index_desc.aux takes them back. They are trusted, not re-validated. */
typedef struct emsar_index_aux {
    int64_t nnz_multi;         /* members of the multi-tid classes */
    const uint32_t *txm_off;   /* [T+1] */
    const int32_t *txm_cid;    /* [nnz_multi] ascending cid per transcript */
    const int32_t *order;      /* [T] emsar_locality_order mode 3 */
    const uint8_t *insertable; /* [C-T] 1 = reachable by the reference's chain walk (:1603-1622) */
    int32_t n_sets_nocut, max_set_tids;
} emsar_index_aux;

typedef struct {
    int32_t T;
    int64_t C, nnz;            /* all classes */
    int64_t n_multi, nnz_multi;
    int32_t n_kseg;            /* distinct cardinalities among multi-tid classes */
    int32_t max_card;
    int64_t hash_slots, hash_inserted;
    int32_t n_sets_nocut;      /* sequence-sharing sets with EUMAcut = 0 */
    int32_t max_set_tids;      /* largest set (transcripts) with EUMAcut = 0 */
    int64_t device_bytes;
    int32_t frag_min, frag_max; /* Fraglengths.min / .max */
} emsar_index_info;

int emsar_index_create(emsar_ctx *ctx, const emsar_index_desc *desc, emsar_index **index);
/* The order in which emsar_index_create lays the transcripts out for the EM kernel's per-SM row ranges (host-only helper, no
 * device needed): order[i] = tid of the i-th transcript. mode 1: connected components of the class <-> transcript graph kept
 * together, tid order inside; 2: additionally, inside a component, the clusters of transcripts that co-occur in at least two
 * small classes (the isoforms of a gene, a tight family) kept together and laid out breadth-first; 3 (what the library uses):
 * 1, or 2 where that leaves fewer member references outside an SM's range; 0: tid order. */
int emsar_locality_order(int32_t T, int64_t C, const int64_t *class_ptr, const int32_t *class_tid, int32_t mode, int32_t *order);
int emsar_index_info_get(const emsar_index *index, emsar_index_info *info);
int emsar_index_aux_get(const emsar_index *index, emsar_index_aux *aux);
int emsar_index_destroy(emsar_index *index);

/* -------- per-sample ----------------------------------------------------------------------------
 * emsar_sample_begin   : clear_readcounts_in_rshbucket_PTR + calloc FraglengthCounts (emsar_main.c:383-384)
 * emsar_sample_count   : update_ReadCounts -> update_rshbucket[_single]_PTR(...,'r',...)
 *                        (emsar_functions.c:838-943, 1514-1537, 1597-1624) for a batch of read groups that
 *                        already passed the reader-side filters (alignment.c:29-60, 85-95; size <= MAX_REPEAT).
 *                        tids may arrive unsorted; asynchronous with respect to the host.
 * emsar_sample_counts_get : ReadCount[] / FraglengthCounts[] / TotalReadCount as scan_rshbucket (:2135-2192)
 *                        would flatten them (bit-exact integers).
 * emsar_sample_solve   : transfer_fraglendist_to_Wf ... compute_iEUMA (emsar_main.c:396-454) and the numeric
 *                        part of print_FPKMfinal (:3163-3212).
 * emsar_sample_segments_get : numeric columns of print_aEUMA_3 (:2262-2300).
 * emsar_sample_end     : the frees at emsar_main.c:478-486. */
int emsar_sample_begin(emsar_index *index, emsar_sample **sample);
int emsar_sample_count(emsar_sample *s, int64_t n_reads, const int64_t *read_ptr, const int32_t *read_tid,
                       const int32_t *read_fraglen);
/* The same batch in its compact wire form - what crosses PCIe is what the counting needs and nothing else: one uint16 length per read group
 * (<= EMSAR_MAX_READ_TIDS) instead of an int64 offset, the tids, and one uint16 fragment length per group, or none at all when every group
 * of the batch has the same fragment length (read_fraglen == NULL: const_fraglen applies; an SE index with one fragment length).
 * 765 MB -> 465 MB per 30M-read sample. Offsets are rebuilt on the device by a prefix sum. Same semantics and asynchrony as emsar_sample_count. */
int emsar_sample_count_compact(emsar_sample *s, int64_t n_reads, int64_t n_tids, const uint16_t *read_len, const int32_t *read_tid,
                               const uint16_t *read_fraglen, int32_t const_fraglen);     /* n_tids = sum of read_len = entries of read_tid */
/* same, but the three arrays already live in device memory of the context's device (bench: resident inputs) */
int emsar_sample_count_device(emsar_sample *s, int64_t n_reads, const void *d_read_ptr, const void *d_read_tid,
                              const void *d_read_fraglen);
/* install counts computed elsewhere (class-sharded mode, tests of the estimator alone) */
int emsar_sample_counts_set(emsar_sample *s, const int32_t *ReadCount, const int32_t *FraglengthCounts);
int emsar_sample_counts_get(emsar_sample *s, int32_t *ReadCount, int32_t *FraglengthCounts, int64_t *TotalReadCount);

typedef struct {
    double eps_abs;        /* reads; <= 0 selects the default 1e-7 */
    double eps_rel;        /* relative; <= 0 selects the default 1e-10 */
    int32_t max_iter;      /* <= 0 selects the default 200000 (reference -i, MAX_NITER_MLE) */
    double delta;          /* reference -d (DELTA): lambda scaled by 10^delta */
    double eumacut;        /* in: EUMAcut carried over from the previous sample (emsar.h:94 is never reset) */
    int32_t max_ntid_per_sid; /* <= 0 selects MAX_NTID_PER_SID = 5000 (emsar.h:17) */
    const uint8_t *in_model;  /* optional [C]: caller-supplied set membership (CS[c] != -1); NULL = computed here */
    int32_t sharded;          /* != 0: the active classes of THIS sample are range-sharded over the ranks of the context's
                                 communicator (emsar_comm_init); every rank must hold the same counts and make the same calls */
} emsar_solve_opts;

typedef struct {
    /* caller-provided output buffers, each [T] (any may be NULL) */
    double *fpkm;            /* FPKM[tid] (one deterministic round; reference averages NUM_ROUND random rounds) */
    double *efflen;          /* iEUMA[tid]  (.fpkm column eff.length) */
    double *ireadcount;      /* iEUMA/1e3 * FPKM * N/1e6 */
    int32_t *ireadcount_int; /* Round_off() */
    double *tpm;             /* FPKM * 1e6 / sum FPKM */
    /* scalars filled by the call */
    int32_t n_iter;
    double final_delta;      /* max_t |dtheta_t| A_t / (eps_abs + eps_rel n_t) of the last iteration; converged iff <= 1 */
    double loglik;           /* sum_c R_c log(lambda_c) - lambda_c over modelled classes (Fp, :2946-2964) */
    int64_t total_ireadcount;
    int64_t total_readcount; /* TotalReadCount */
    double eumacut;          /* EUMAcut after the set-size loop (emsar_main.c:411-425) */
    int32_t max_sid;
    double em_ms;            /* device time of the EM loop (CUDA events) */
    double prep_ms;          /* device+host time of the per-sample model build */
} emsar_solve_out;

int emsar_sample_solve(emsar_sample *s, const emsar_solve_opts *opts, emsar_solve_out *out);
/* [C] each, any may be NULL: adjEUMA (eff.length), expected_Readcount, set id (CS, -1 = cut by EUMAcut) */
int emsar_sample_segments_get(emsar_sample *s, double *adjEUMA, double *expected, int32_t *set_id);
/* [nF]: Wf (normalized.Fragment.length.sampling.prob of .fraglength_effect) */
int emsar_sample_wf_get(emsar_sample *s, double *Wf);
int emsar_sample_end(emsar_sample *s);

/* -------- ingestion pipeline: page-locked batch buffers and asynchronous counting ----------
 * emsar_sample_count returns as soon as the copies and the kernel are enqueued when its arrays live in page-locked memory
 * (emsar_host_alloc); the arrays must then stay untouched until emsar_sample_count_wait(s, lag) has returned, which blocks
 * until the batch submitted `lag` calls before the latest one has been consumed (lag 0 = the latest). A reader with two
 * buffer sets calls emsar_sample_count(batch k) and then emsar_sample_count_wait(s, 1) before it refills the set of batch
 * k-1 (emsar_b200/host/emsar_main.c). The reference parses and counts on one thread (emsar_functions.c:323-836). */
int emsar_host_alloc(emsar_ctx *ctx, size_t bytes, void **p);
int emsar_host_free(emsar_ctx *ctx, void *p);
int emsar_sample_count_wait(emsar_sample *s, int32_t lag);

/* -------- one sample sharded over several GPUs (BASELINE.json configs[2]) ------------------------
 * One process (or thread) per GPU, each with its own context holding the SAME index. The multi-tid classes that are
 * active in the sample are cut into nnz-balanced contiguous ranges, one per rank; every EM iteration each rank computes
 * the per-transcript sums of its classes and an all-reduce (fp64, one value per participating transcript) over NVLink adds
 * them up; theta stays replicated and bit-identical on every rank. The all-reduce runs INSIDE the persistent EM kernel over
 * peer memory (pushes to the owner of a transcript slice, owner update, pushes of the new theta; two cross-GPU barriers per
 * iteration, no host involvement), or, where peer memory cannot be mapped, as one ncclAllReduce per iteration.
 * The reference has no counterpart (single process, emsar_main.c).
 * The 128-byte id is NCCL's unique id: rank 0 makes it, the caller ships it to the other ranks (bench.py and the tests
 * use torch.distributed for that), then every rank calls emsar_comm_init. */
int emsar_comm_unique_id(uint8_t id[128]);
int emsar_comm_init(emsar_ctx *ctx, int32_t rank, int32_t nranks, const uint8_t id[128]);
int emsar_comm_destroy(emsar_ctx *ctx);
/* peer_memory: 1 = the fused kernel exchanges the sums over NVLink peer memory (CUDA IPC between processes, peer access inside
 * one process), -1 = peer memory could not be mapped (or EMSAR_SHARD_MODE=nccl): ncclAllReduce per iteration, 0 = not decided
 * yet (decided collectively at the first sharded solve) */
int emsar_comm_info(emsar_ctx *ctx, int32_t *rank, int32_t *nranks, int32_t *peer_memory);
/* each rank counted a different slice of the read groups: sum ReadCount / FraglengthCounts over the ranks (exact: integers) */
int emsar_sample_counts_allreduce(emsar_sample *s);
/* contiguous ranges of equal weight: out[r] .. out[r+1] is rank r's range of the n items (host-only helper, no device needed) */
int emsar_shard_ranges(int64_t n, const int64_t *weight_prefix, int32_t nranks, int64_t *out);

/* -------- finer-grained steps of emsar_sample_solve (tests, bench.py, profiling) ----------------- */
typedef struct {
    int32_t T;
    int64_t C_a, nnz_a;       /* active multi-tid classes (modelled, R > 0) and their members */
    int64_t rows_short;       /* participating transcripts (A_t > 0): the rows of the EM */
    int64_t rows_long;        /* of those: rows with more than 64 active entries (one warp per row instead of a slice lane) */
    int64_t rows_hub;         /* of those: rows whose entries are split over several CTAs (0 = none) */
    int64_t rows_fixed;       /* transcripts outside the EM (A_t == 0) */
    int64_t e_tiles, m_tiles;
    int64_t bytes_per_iter;   /* algorithmic bytes of one EM iteration: 8 nnz_a + 24 C_a + 44 T */
    int64_t stream_bytes_per_iter; /* bytes the kernels actually stream per iteration (index + state) */
    int32_t em_variant;       /* k_em_persistent<V> this sample runs: 3 barrier-free, 1 grid barriers (overflow), 0 TMA-pipelined, 2 class-sharded */
    int32_t all_local;        /* 1: every CTA holds its whole halo and all its q in shared memory */
    int64_t halo_rows, halo_classes;   /* distinct (CTA, remote row) / (CTA, remote class) references */
    int64_t resident_index_bytes;      /* index data kept in shared memory for the whole kernel, summed over the CTAs */
    int64_t index_bytes;               /* packed index data of the model (E members + read counts + M entries, with padding) */
    int64_t peer_bytes_per_iter;       /* class-sharded sample: bytes this rank writes into other GPUs' memory per iteration (0 otherwise) */
} emsar_model_stats;
/* Wf, adjEUMA, EUMAps, sets/EUMAcut, A_t, iEUMA and the packed active model; theta := start point */
int emsar_sample_prepare(emsar_sample *s, const emsar_solve_opts *opts);
int emsar_sample_model_stats(emsar_sample *s, emsar_model_stats *st);
/* run up to max_iter EM iterations from the current theta; stop_on_conv != 0 stops once delta <= 1.
   elapsed_ms is measured with CUDA events on the context's stream. */
int emsar_sample_em_run(emsar_sample *s, int32_t max_iter, int32_t stop_on_conv, int32_t reset_theta,
                        int32_t *iters_done, double *final_delta, double *elapsed_ms);
int emsar_sample_theta_get(emsar_sample *s, double *theta);
/* Restart rounds (reference -n, NUM_ROUND: emsar_main.c:444-450 runs the estimator from rand() starting points and prints mean and sd):
 * sets theta of every participating transcript to a seeded pseudo-random positive value (log-uniform over two decades around 1; the
 * same seed gives the same start on every device) and resets the iteration count. emsar_sample_em_run + emsar_sample_finalize then give
 * that round's estimate. EM reaches the same optimum from any positive start where the optimum is unique; rounds differ (sd > 0) exactly
 * on transcripts that the data cannot tell apart. */
int emsar_sample_theta_randomize(emsar_sample *s, uint64_t seed);
/* Measurement helpers (bench.py): CUDA events on the context's own stream around whatever the caller enqueues in between
 * (torch.cuda.Event only sees torch's stream). emsar_cuda_timer_stop synchronizes and returns the milliseconds since _start. */
int emsar_cuda_timer_start(emsar_ctx *ctx);
int emsar_cuda_timer_stop(emsar_ctx *ctx, double *elapsed_ms);
/* re-runs the adjEUMA kernel of a prepared sample `reps` times (same inputs, same outputs) and returns the mean duration of one
 * launch: the HBM-bound one-off of a PE index (4 * C * nF bytes) */
int emsar_sample_time_adjeuma(emsar_sample *s, int32_t reps, double *ms_per_launch);
int emsar_sample_finalize(emsar_sample *s, emsar_solve_out *out);

/* ---- index construction on the device ---------------------------------------------------------------------------
 * Replaces the reference's suffix-array construction for one read length: initialize_suffixarray_* / sort (emsar_functions.c:949-1230),
 * construct_rshbucket_2 (:1758-1816), construct_rshbucket_PE_3 (:1902-1974), process_mate1_cluster_by_mate_3 (:2784-2934). The caller owns
 * the fasta reader and the class store (emsar_b200/host/build_index.c): it passes the concatenated transcriptome exactly as read_raw_fasta
 * (:31-196) lays it out and folds the returned counts into its store (update_rshbucket[_single] 'e', :1514-1596). */
typedef struct emsar_build_desc {
    const char *seq;          /* S[0 .. end]: transcripts (upper case, other letters = 'N') joined by '@', '$' at `border`, the reverse
                                 complement of S[0 .. border), '$' at `end` = 2 * border + 1 */
    int64_t border, end;
    int32_t T;
    const int64_t *start;     /* [T + 1] first position of every transcript in the forward half; start[T] = border + 1 */
    int32_t pe;               /* paired-end: fragments (mate 1, mate 2 at distance d on the same strand string) instead of single reads */
    int32_t stranded;         /* 0: an occurrence stands for the smaller of itself and its reverse complement / flipped fragment */
    int32_t readlen;          /* read length of this pass */
    int32_t d_min, d_max;     /* pe: mate distance range = fragment length - read length (>= 0); ignored otherwise */
    int32_t max_repeat;       /* MAX_REPEAT (-k): a sequence shared by this many occurrences or more is dropped */
} emsar_build_desc;
typedef struct emsar_build_classes {     /* arrays owned by the library until emsar_build_classes_free */
    int32_t T, n_d;           /* n_d = d_max - d_min + 1 (1 for single-end) */
    const int32_t *single_count;   /* [T * n_d] sequences seen in exactly one place: transcript t, distance index d - d_min */
    int64_t n_class;               /* distinct (tid multiset, distance) pairs shared by 2 .. max_repeat - 1 occurrences */
    const int64_t *class_off;      /* [n_class + 1] into class_tid */
    const int32_t *class_tid;      /* ascending tids, a transcript repeats when the sequence occurs in it more than once */
    const int32_t *class_d;        /* [n_class] distance index */
    const int32_t *class_count;    /* [n_class] number of distinct sequences with exactly this tid multiset at this distance */
    int64_t occurrences, runs;     /* diagnostics: occurrences keyed, distinct sequences among them */
    int32_t partitions;
    double device_ms;              /* CUDA-event time of the call on the context's stream (copies, kernels, sorts) */
    void *owner;
} emsar_build_classes;
int emsar_build_classes_run(emsar_ctx *ctx, const emsar_build_desc *desc, emsar_build_classes *out);
void emsar_build_classes_free(emsar_build_classes *out);

#ifdef __cplusplus
}
#endif
#endif /* EMSAR_CUDA_H */
