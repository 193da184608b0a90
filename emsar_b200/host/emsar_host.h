/*
 * emsar_host.h — host side of the B200 EMSAR hot path, in C like the reference: the `.rsh` text loader that packs
 * the class store into CSR, the alignment readers (bowtie / SAM / BAM) with the reference's read-group semantics,
 * and the output writers. No CUDA here; the numeric work is behind include/emsar_cuda.h.
 * Reference citations are relative to parklab/emsar v2.0.1 src/.
 */
#ifndef EMSAR_HOST_H
#define EMSAR_HOST_H

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EMSAR_HOST_ERRLEN 512

/* ---- rsh index in host memory (product of construct_rsh_from_rshfile, emsar_functions.c:1351-1510),
 *      flattened in the reference's scan order (scan_rshbucket :2149-2191) ---- */
typedef struct {
    int32_t T;              /* max_tid + 1 */
    int64_t C;              /* max_cid + 1 */
    int64_t *class_ptr;     /* [C+1] */
    int32_t *class_tid;     /* [class_ptr[C]] */
    int32_t nF;             /* nFraglen */
    int32_t *euma;          /* [C * nF] */
    uint8_t *has_node;      /* [C] */
    int32_t min_fraglength, max_fraglength, readlength, max_t_size; /* header fields 3,4,5,2 */
    int32_t frag_min, frag_max; /* Fraglengths.min/.max (determine_fraglength_range :2471-2475) */
    char **names;           /* IndexTable[tid] */
    /* tname -> tid (replaces the reference's character trie, stringhash.c) */
    uint32_t *name_slots;
    uint32_t name_mask;
    /* optional: the structures the device library derives from the class table (emsar_index_aux of include/emsar_cuda.h), carried by a
     * packed image so that creating the index costs no pass over the members. Owned by this struct when `aux_owned`. */
    int has_aux, aux_owned;
    int64_t aux_nnz_multi;
    uint32_t *aux_txm_off;       /* [T+1] */
    int32_t *aux_txm_cid;        /* [aux_nnz_multi] */
    int32_t *aux_order;          /* [T] */
    uint8_t *aux_insertable;     /* [C-T] */
    int32_t aux_n_sets_nocut, aux_max_set_tids;
} emsar_rsh;

int emsar_rsh_load(const char *path, emsar_rsh **out, char *err);
void emsar_rsh_free(emsar_rsh *r);
int emsar_rsh_tid(const emsar_rsh *r, const char *name);           /* -1 when absent (search_treehash) */
int emsar_rsh_write(const emsar_rsh *r, int pe, const char *path, char *err); /* print_rsh :2071-2130 */
/* packed binary image of a loaded index: same arrays, no parsing; tied to the text file it was made from */
int emsar_rsh_save_packed(const emsar_rsh *r, const char *path, const char *src_path, char *err);
int emsar_rsh_load_packed(const char *path, const char *src_path, emsar_rsh **out, char *err);   /* 2 = stale w.r.t. src_path */
int emsar_rsh_load_auto(const char *path, emsar_rsh **out, int *from_cache, char *err);          /* <path>.pack if fresh, else text */

struct emsar_build_desc;
struct emsar_build_classes;
/* ---- index construction from a fasta (what emsar-build / emsar -x do; emsar_b200/host/build_index.c) ---- */
typedef struct {
    int pe;                        /* -P */
    int stranded;                  /* library strand type other than "ns" */
    int readlength;                /* PE: the read length */
    int readlen_min, readlen_max;  /* SE: read length range ("50" or "48-52") */
    int min_fraglength, max_fraglength; /* PE: -f / -F (reference defaults 1 / 400) */
    int max_repeat;                /* -k MAX_REPEAT: substrings occurring this often or more are dropped (default 100) */
    char header;                   /* 'E' Ensembl header (name up to the first blank, default), 'R' RefSeq (4th '|' field) */
    int threads;                   /* worker threads of the paired-end construction (-p); results do not depend on it */
    /* optional: the classes are constructed on the device (libemsar_cuda: emsar_build_classes_run / _free / emsar_cuda_last_error with an
     * open context in device_ctx; include/emsar_cuda.h) instead of by the host code in build_index.c. Fasta reader, class store and the
     * print order stay here; the file is the same byte for byte. NULL device_run = host construction. */
    int (*device_run)(void *device_ctx, const struct emsar_build_desc *d, struct emsar_build_classes *out);
    void (*device_free)(struct emsar_build_classes *out);
    const char *(*device_error)(void);
    void *device_ctx;
} emsar_build_opts;
int emsar_rsh_build(const char *fasta_path, const emsar_build_opts *o, emsar_rsh **out, char *err);
/* read length(s) of an alignment file, as emsar -x learns them (emsar_main.c:306-316): PE -> the first aligned record,
 * SE -> minimum and maximum over the whole file */
int emsar_sniff_readlengths(const char *path, char format, int pe, int *rl_min, int *rl_max, char *err);

/* ---- alignment readers ------------------------------------------------------------------------ */
typedef struct {
    int pe;            /* -P */
    char strand;       /* library_strand_type: 0, '+', '-' (set_library_strand_type :16-22) */
    int max_repeat;    /* -k MAX_REPEAT */
    char format;       /* 0 = default bowtie output, 's' = SAM, 'b' = BAM (bamflag) */
    int64_t batch_reads; /* read groups per batch handed to the callback (0 = 1<<20) */
    /* ingestion pipeline; all optional, zero = the plain single-threaded reader */
    int io_threads;      /* BAM: BGZF blocks are inflated by this many worker threads ahead of the parser (0: zlib gz* inline) */
    int nbuf;            /* 2: two batch buffer sets used in turn. The callback may then return while the device is still
                            copying the batch; it must only make sure that the PREVIOUS batch has been consumed */
    void *(*buf_alloc)(void *hook_user, size_t bytes);   /* batch buffers from here (e.g. pinned host memory) instead of malloc */
    void (*buf_free)(void *hook_user, void *p);
    void *hook_user;
} emsar_reader_opts;

/* A batch of read groups that passed the reader-side filters (add_alignment_to_list alignment.c:29-60, size <=
 * MAX_REPEAT, PE: check_fraglen_discrepancy alignment.c:85-95): per group the tids of the kept alignments in file
 * order and the fragment length of the first one — exactly the argument of update_ReadCounts (:838). */
typedef int (*emsar_batch_fn)(void *user, int64_t n_reads, const int64_t *read_ptr, const int32_t *read_tid,
                              const int32_t *read_fraglen);

/* `readlength` is the PE read length (in: rsh header field 5 or -1; out: the value seen). Empty path = stdin.
 * Returns 0, or non-zero with `err` filled (the reference prints the same text to stderr and exits). */
int emsar_read_alignments(const emsar_rsh *r, const char *path, const emsar_reader_opts *o, int *readlength,
                          emsar_batch_fn fn, void *user, char *err);

/* one read group through the reference's list logic; exposed for tests. alignments: arrays of n; keep: out indices.
 * returns the kept count or -1 when the group is dropped */
int emsar_filter_group(int n, const int *tid, const int *mm, const int *fraglen, const int *pos, int max_repeat, int pe,
                       int *keep);
int emsar_parse_mmstr(const char *s);       /* alignment.c:101-108 */
int emsar_parse_sam_mmstr(const char *s);   /* emsar_functions.c:418-424 */
int emsar_check_mate_readid_matching(const char *id1, const char *id2); /* alignment.c:113-126 */

/* ---- writers (print_FPKMfinal :3163-3212, print_FraglengthDist :2477-2493, print_aEUMA_3 :2262-2300) ---- */
int emsar_write_fpkm(const char *path, const emsar_rsh *r, const double *fpkm, const double *sd, const double *efflen,
                     const double *ireadcount, const int32_t *ireadcount_int, const double *tpm, char *err);
int emsar_write_fraglength(const char *path, const emsar_rsh *r, const int32_t *FraglengthCounts, const double *Wf, char *err);
int emsar_write_segments(const char *path, const emsar_rsh *r, const int32_t *set_id, const double *adjEUMA,
                         const int32_t *ReadCount, const double *expected, char *err);

#ifdef __cplusplus
}
#endif
#endif
