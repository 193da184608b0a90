/* emsar — command-line driver of the B200 quantification path. Same flags, positional arguments and output files
 * as the reference driver (parklab/emsar v2.0.1 src/emsar_main.c:4-511); the per-file loop (:380-488) calls
 * libemsar_cuda (include/emsar_cuda.h) where the reference calls update_ReadCounts / scan_rshbucket / run_MLE_threads.
 *
 * Differences that are visible and deliberate:
 *   - the estimator is a deterministic EM run from theta = 1; with -n R (R > 1) R - 1 further runs start from seeded random points
 *     and the files carry the mean and sd.of.FPKM over the R runs like the reference's rounds (without -n: one run, sd 0.000000);
 *     -e / -r / -i map to eps_abs / eps_rel / max EM iterations when given;
 *   - -x builds the index on the host first (emsar_b200/host/build_index.c: same classes and counts as the reference's
 *     suffix-array construction, the read length(s) are learnt from the first alignment file like emsar_main.c:306-316);
 *     -T (print suffix array) is rejected: there is no suffix array;  -m/-W/-w (positional bias, undocumented and
 *     half-implemented in the reference) are rejected;
 *   - -k above 1024 is rejected (device sort limit, EMSAR_MAX_READ_TIDS);
 *   - EMSAR_DEVICES=0,1,.. spreads the files of a -M list over several GPUs (one host thread per GPU taking files from a shared
 *     queue, largest first, no communication). EUMAcut carries over from file to file on one GPU as in the reference; on several
 *     GPUs every file starts from the initial cut, so that the output does not depend on which GPU took which file.
 *   - EMSAR_DEVICES=0,1,.. with EMSAR_SHARD=classes puts EVERY sample on all the GPUs: the sample's active classes are
 *     range-sharded, theta is all-reduced over NVLink every iteration (BASELINE.json configs[2]); GPU 0 reads and counts,
 *     the other GPUs join the collectives, the first GPU writes the files.
 */
#define _GNU_SOURCE
#include <getopt.h>
#include <sys/stat.h>
#include <math.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "emsar_cuda.h"
#include "emsar_host.h"

#define FILENAMEMAX 1000
#define MAX_nALNFILES 1000

typedef struct {
    char rshfile[FILENAMEMAX], fasta[FILENAMEMAX], strand_str[8];
    int pe, multisample, print_segments, print_rsh, verbose, max_repeat, nthread, max_iter;
    int min_fl, max_fl, num_round, rounds_set;
    char bamflag, strand, fasta_header;
    double eps_abs, eps_rel, delta;
    const char *outdir, *outprefix;
    char **aln; int naln;
} options;

static void die(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    fputc('\n', stderr);
    exit(1);
}

static void stamp(const options *o, const char *what)
{
    if (o->verbose <= 0) return;
    time_t t = time(NULL);
    struct tm tm;
    localtime_r(&t, &tm);
    char b[32];
    strftime(b, sizeof b, "%m/%d,%T", &tm);
    fprintf(stdout, "%s :%s\n", what, b);
    fflush(stdout);
}

static void usage(const char *p)
{
    fprintf(stdout,
            "usage: %s <options> -I rshfile | -x fastafile  outdir outprefix [alignmentfile | listfile (with -M)]\n"
            "  -x fastafile build the index from the transcriptome first (read lengths are taken from the first alignment file;\n"
            "               -h E|R fasta header style, -F/-f fragment length range for -P, -k also caps repeated substrings)\n"
            "  -I rshfile   rsh index built by emsar-build          -M  third argument is a list of alignment files\n"
            "  -P           paired-end                              -s  ns|ssf|ssr (SE)  ns|ssfr|ssrf (PE)\n"
            "  -S / -B      SAM / BAM input (default: bowtie out)   -k  max alignments per read (default 100)\n"
            "  -g           also write prefix.i.segments            -R  write the rsh index to outdir/prefix.rsh\n"
            "  -i n         max EM iterations (default 200000)      -e/-r  absolute (reads) / relative EM tolerance\n"
            "  -d n         delta (lambda scaled by 10^n)           -q / -v  quiet / verbose\n"
            "  -p n         BGZF inflate threads for BAM input (default 4)\n"
            "  -n -b -t -h -F -f are accepted for compatibility; -F/-f are overwritten by the rsh header.\n",
            p);
}

/* Ingestion pipeline: BGZF blocks are inflated by -p worker threads, the parser fills one of two page-locked batch buffers
 * while the device still copies / counts the other one (emsar_sample_count is asynchronous on page-locked arrays). */
typedef struct { emsar_sample *s; emsar_ctx *ctx; int rc; int64_t batches, groups; uint16_t *len16[2], *fl16[2]; int64_t cap16; } count_ctx;
static int on_batch(void *user, int64_t n, const int64_t *ptr, const int32_t *tid, const int32_t *fl)
{
    /* the batch goes to the device in its compact wire form (emsar_sample_count_compact): 16-bit lengths instead of 64-bit offsets, 16-bit
     * fragment lengths or none when the whole batch has one - 40 % fewer bytes over PCIe, which is what eight GPUs on one host contend for */
    count_ctx *c = (count_ctx *)user;
    const int k = (int)(c->batches & 1);
    if (n > c->cap16) {
        if (c->batches > 0 && (c->rc = emsar_sample_count_wait(c->s, 0))) return c->rc;       /* nothing in flight reads the old arrays */
        for (int b = 0; b < 2; b++) { emsar_host_free(c->ctx, c->len16[b]); emsar_host_free(c->ctx, c->fl16[b]); c->len16[b] = c->fl16[b] = NULL; }
        c->cap16 = n + (n >> 2) + 1024;
        for (int b = 0; b < 2; b++)
            if (emsar_host_alloc(c->ctx, sizeof(uint16_t) * (size_t)c->cap16, (void **)&c->len16[b]) || emsar_host_alloc(c->ctx, sizeof(uint16_t) * (size_t)c->cap16, (void **)&c->fl16[b]))
                return c->rc = EMSAR_ERR_NOMEM;
    }
    int wide = 0, same = 1;
    for (int64_t r = 0; r < n; r++) {
        const int64_t l = ptr[r + 1] - ptr[r];
        if (l > 65535 || fl[r] < 0 || fl[r] > 65535) { wide = 1; break; }
        c->len16[k][r] = (uint16_t)l; c->fl16[k][r] = (uint16_t)fl[r];
        same &= fl[r] == fl[0];
    }
    if (wide) c->rc = emsar_sample_count(c->s, n, ptr, tid, fl);
    else c->rc = emsar_sample_count_compact(c->s, n, ptr[n] - ptr[0], c->len16[k], tid + ptr[0], same ? NULL : c->fl16[k], n > 0 ? fl[0] : 0);
    if (!c->rc) c->rc = emsar_sample_count_wait(c->s, 1);      /* the other buffer set (batch k-1) is free again */
    c->batches++; c->groups += n;
    return c->rc;
}
static void *pinned_alloc(void *user, size_t bytes)
{
    void *p = NULL;
    if (emsar_host_alloc((emsar_ctx *)user, bytes, &p)) die("%s", emsar_cuda_last_error());
    return p;
}
static void pinned_free(void *user, void *p) { emsar_host_free((emsar_ctx *)user, p); }

typedef struct {
    const options *o; const emsar_rsh *rsh; int device, worker, nworker; int rc;
    int by_class;                      /* EMSAR_SHARD=classes */
    uint8_t *comm_id;                  /* shared: NCCL unique id made by worker 0 */
    pthread_barrier_t *bar;
    int *next; const int *order; int serial;      /* -M work queue: shared ticket, files sorted by size (largest first) */
} worker_arg;

static int run_file(const options *o, const emsar_rsh *rsh, emsar_ctx *ctx, emsar_index *ix, int i, double *eumacut, int shard_rank, int shard_n)
{
    /* shard_n > 1: this sample runs on shard_n GPUs at once; rank 0 reads, counts and writes, the others only take part in the
     * collectives (counts all-reduce, per-iteration all-reduce inside emsar_sample_solve) */
    const int sharded = shard_n > 1, lead = shard_rank == 0;
    char err[EMSAR_HOST_ERRLEN] = "";
    emsar_sample *s = NULL;
    int rc = emsar_sample_begin(ix, &s);
    if (rc) die("%s: %s", emsar_cuda_strerror(rc), emsar_cuda_last_error());
    if (lead) fprintf(stdout, "alnfile[%d]=%s\n", i, o->aln[i]);
    emsar_reader_opts ro;
    memset(&ro, 0, sizeof ro);
    ro.pe = o->pe; ro.strand = o->strand; ro.max_repeat = o->max_repeat; ro.format = o->bamflag; ro.batch_reads = 1 << 20;
    ro.io_threads = o->nthread > 0 ? o->nthread : 4;            /* -p: BGZF inflate threads (the reference's -p sizes its MLE thread team) */
    ro.nbuf = 2; ro.buf_alloc = pinned_alloc; ro.buf_free = pinned_free; ro.hook_user = ctx;
    int readlength = rsh->readlength;
    count_ctx cc;
    memset(&cc, 0, sizeof cc);
    cc.s = s; cc.ctx = ctx;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    if (lead) {
        if (emsar_read_alignments(rsh, o->aln[i], &ro, &readlength, on_batch, &cc, err)) {
            if (cc.rc) die("%s: %s", emsar_cuda_strerror(cc.rc), emsar_cuda_last_error());
            die("%s", err);
        }
        if ((rc = emsar_sample_count_wait(s, 0))) die("%s: %s", emsar_cuda_strerror(rc), emsar_cuda_last_error());
    } else if ((rc = emsar_sample_count(s, 0, NULL, NULL, NULL))) die("%s: %s", emsar_cuda_strerror(rc), emsar_cuda_last_error());
    for (int b = 0; b < 2; b++) { emsar_host_free(ctx, cc.len16[b]); emsar_host_free(ctx, cc.fl16[b]); }
    if (sharded && (rc = emsar_sample_counts_allreduce(s))) die("%s: %s", emsar_cuda_strerror(rc), emsar_cuda_last_error());   /* everybody gets rank 0's counts */
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (o->verbose > 0 && lead) {
        const double sec = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
        fprintf(stdout, "alignments read: %lld read groups in %lld batches, %.2f s (%.2f M groups/s, %d inflate threads)\n", (long long)cc.groups,
                (long long)cc.batches, sec, sec > 0 ? 1e-6 * (double)cc.groups / sec : 0.0, ro.io_threads);
    }
    if (lead) stamp(o, "\nscanning rsh array and constructing EUMA, ReadCount and CT array...");
    emsar_solve_opts so;
    memset(&so, 0, sizeof so);
    so.sharded = sharded;
    so.eps_abs = o->eps_abs; so.eps_rel = o->eps_rel; so.max_iter = o->max_iter; so.delta = o->delta; so.eumacut = *eumacut;
    emsar_solve_out out;
    memset(&out, 0, sizeof out);
    const int32_t T = rsh->T;
    out.fpkm = (double *)malloc(sizeof(double) * T);
    out.efflen = (double *)malloc(sizeof(double) * T);
    out.ireadcount = (double *)malloc(sizeof(double) * T);
    out.ireadcount_int = (int32_t *)malloc(sizeof(int32_t) * T);
    out.tpm = (double *)malloc(sizeof(double) * T);
    rc = emsar_sample_solve(s, &so, &out);
    if (rc) die("%s: %s", emsar_cuda_strerror(rc), emsar_cuda_last_error());
    if (out.eumacut != *eumacut && o->verbose > 0 && lead) fprintf(stdout, "module size too big. EUMAcut is readjusted to %.0f\n", out.eumacut);
    *eumacut = out.eumacut;   /* EUMAcut is a global that is never reset between files (emsar.h:94) */
    if (lead && out.final_delta > 1.0)       /* whatever -q says: the reference prints "FPKM not converging, reinitializing.." (:3110) */
        fprintf(stderr, "WARNING: %s: the estimator did not converge in %d iterations (delta %.3g > 1); the results are written but not final. "
                        "Raise -i or loosen the tolerances.\n", o->aln[i], out.n_iter, out.final_delta);
    if (!lead) {               /* the other ranks hold the same result; only the first one writes it */
        free(out.fpkm); free(out.efflen); free(out.ireadcount); free(out.ireadcount_int); free(out.tpm);
        emsar_sample_end(s);
        return 0;
    }
    if (sharded && o->verbose > 0) {
        int32_t pm = 0;
        emsar_comm_info(ctx, NULL, NULL, &pm);
        fprintf(stdout, "sample sharded by class range over %d GPUs (%s)\n", shard_n, pm == 1 ? "all-reduce inside the EM kernel, NVLink peer memory" : "ncclAllReduce per iteration");
    }
    if (o->verbose > 0) {
        fprintf(stdout, "EM finished: %d iterations, delta %.3g, %.1f ms on the device (model build %.1f ms), logL %.10g\n",
                out.n_iter, out.final_delta, out.em_ms, out.prep_ms, out.loglik);
        emsar_model_stats ms;
        if (emsar_sample_model_stats(s, &ms) == 0 && out.n_iter > 0 && out.em_ms > 0)     /* the per-phase profile line */
            fprintf(stdout, "EM model: %lld active classes, %lld members, %d transcripts; %.1f us per iteration, %.0f GB/s of algorithmic bytes (%lld per iteration)\n",
                    (long long)ms.C_a, (long long)ms.nnz_a, ms.T, 1e3 * out.em_ms / out.n_iter,
                    (double)ms.bytes_per_iter * out.n_iter / (out.em_ms * 1e-3) / 1e9, (long long)ms.bytes_per_iter);
    }
    int32_t *F = (int32_t *)malloc(sizeof(int32_t) * ((size_t)rsh->max_fraglength + 1));
    int32_t *R = (int32_t *)malloc(sizeof(int32_t) * (size_t)rsh->C);
    double *Wf = (double *)malloc(sizeof(double) * rsh->nF);
    int64_t N = 0;
    if ((rc = emsar_sample_counts_get(s, R, F, &N)) || (rc = emsar_sample_wf_get(s, Wf))) die("%s: %s", emsar_cuda_strerror(rc), emsar_cuda_last_error());
    char p1[FILENAMEMAX * 2 + 64], p2[FILENAMEMAX * 2 + 64], p3[FILENAMEMAX * 2 + 64];
    snprintf(p1, sizeof p1, "%s/%s.%d.fpkm", o->outdir, o->outprefix, i);
    snprintf(p2, sizeof p2, "%s/%s.%d.fraglength_effect", o->outdir, o->outprefix, i);
    snprintf(p3, sizeof p3, "%s/%s.%d.segments", o->outdir, o->outprefix, i);
    /* the .segments file (expected counts per class) comes from the standard run, before any restart round moves theta */
    if (o->print_segments) {
        double *adj = (double *)malloc(sizeof(double) * (size_t)rsh->C), *ex = (double *)malloc(sizeof(double) * (size_t)rsh->C);
        int32_t *cs = (int32_t *)malloc(sizeof(int32_t) * (size_t)rsh->C);
        if ((rc = emsar_sample_segments_get(s, adj, ex, cs))) die("%s: %s", emsar_cuda_strerror(rc), emsar_cuda_last_error());
        if (emsar_write_segments(p3, rsh, cs, adj, R, ex, err)) die("%s", err);
        free(adj); free(ex); free(cs);
    }
    double *sd = NULL;
    if (o->rounds_set && o->num_round > 1 && !sharded) {
        /* -n R given explicitly: R - 1 restart rounds from seeded random starting points next to the standard run (reference emsar_main.c:444-450);
         * the file then carries the mean over the rounds and sd.of.FPKM = sqrt(sum (x - m)^2 / (R - 1)) / R as print_FPKMfinal (:3186-3208) does.
         * Without -n there is one deterministic run and the column prints 0. */
        const int R_ = o->num_round;
        double *sum = (double *)calloc((size_t)T, sizeof(double)), *sq = (double *)calloc((size_t)T, sizeof(double)), *fr = (double *)malloc(sizeof(double) * (size_t)T);
        double **all = (double **)malloc(sizeof(double *) * (size_t)R_);
        all[0] = (double *)malloc(sizeof(double) * (size_t)T);
        memcpy(all[0], out.fpkm, sizeof(double) * (size_t)T);
        for (int rd = 1; rd < R_; rd++) {
            emsar_solve_out o2;
            memset(&o2, 0, sizeof o2);
            all[rd] = (double *)malloc(sizeof(double) * (size_t)T);
            o2.fpkm = all[rd];
            int32_t it2 = 0; double fd2 = 0, ms2 = 0;
            if ((rc = emsar_sample_theta_randomize(s, (uint64_t)rd)) || (rc = emsar_sample_em_run(s, 0, 1, 0, &it2, &fd2, &ms2)) || (rc = emsar_sample_finalize(s, &o2)))
                die("%s: %s", emsar_cuda_strerror(rc), emsar_cuda_last_error());
            if (o->verbose > 0) fprintf(stdout, "round %d/%d: %d iterations, delta %.3g, logL %.10g\n", rd + 1, R_, o2.n_iter, o2.final_delta, o2.loglik);
        }
        sd = (double *)malloc(sizeof(double) * (size_t)T);
        double tot = 0;
        for (int32_t t = 0; t < T; t++) {
            double m = 0;
            for (int rd = 0; rd < R_; rd++) m += all[rd][t];
            m /= R_;
            double v = 0;
            for (int rd = 0; rd < R_; rd++) v += (all[rd][t] - m) * (all[rd][t] - m);
            sd[t] = sqrt(v / (R_ - 1)) / R_;
            out.fpkm[t] = m;
            tot += m;
        }
        const double nscale = (double)out.total_readcount / 1E6;
        for (int32_t t = 0; t < T; t++) {
            const double ir = (out.efflen[t] / 1E3) * out.fpkm[t] * nscale;                  /* print_FPKMfinal :3203 */
            out.ireadcount[t] = ir;
            out.ireadcount_int[t] = (ir - (int)ir >= 0.5) ? (int)ir + 1 : (int)ir;           /* Round_off :3215-3217 */
            out.tpm[t] = out.fpkm[t] * 1E6 / tot;
        }
        for (int rd = 0; rd < R_; rd++) free(all[rd]);
        free(all); free(sum); free(sq); free(fr);
    }
    if (emsar_write_fpkm(p1, rsh, out.fpkm, sd, out.efflen, out.ireadcount, out.ireadcount_int, out.tpm, err)) die("%s", err);
    free(sd);
    if (o->verbose > 0) fprintf(stdout, "Total inferred readcount=%lld\n", (long long)out.total_ireadcount);
    if (emsar_write_fraglength(p2, rsh, F, Wf, err)) die("%s", err);
    fprintf(stdout, "Complete: Output file :\n  %s\n  %s\n", p1, p2);
    if (o->print_segments) fprintf(stdout, "  %s\n", p3);
    fflush(stdout);
    free(F); free(R); free(Wf);
    free(out.fpkm); free(out.efflen); free(out.ireadcount); free(out.ireadcount_int); free(out.tpm);
    emsar_sample_end(s);
    return 0;
}

static void *worker(void *p)
{
    worker_arg *w = (worker_arg *)p;
    const options *o = w->o;
    emsar_ctx *ctx = NULL;
    int rc = emsar_cuda_open(w->device, &ctx);
    if (rc) die("%s: %s", emsar_cuda_strerror(rc), emsar_cuda_last_error());
    emsar_index_desc d;
    memset(&d, 0, sizeof d);
    const emsar_rsh *r = w->rsh;
    d.T = r->T; d.C = r->C; d.class_ptr = r->class_ptr; d.class_tid = r->class_tid; d.nF = r->nF; d.euma = r->euma; d.has_node = r->has_node;
    d.min_fraglength = r->min_fraglength; d.max_fraglength = r->max_fraglength; d.readlength = r->readlength; d.max_t_size = r->max_t_size;
    emsar_index *ix = NULL;
    emsar_index_aux ax;
    if (w->rsh->has_aux && !getenv("EMSAR_RSH_NO_AUX")) {        /* a packed image with the derived arrays: no pass over the members */
        memset(&ax, 0, sizeof ax);
        ax.nnz_multi = w->rsh->aux_nnz_multi; ax.txm_off = w->rsh->aux_txm_off; ax.txm_cid = w->rsh->aux_txm_cid; ax.order = w->rsh->aux_order;
        ax.insertable = w->rsh->aux_insertable; ax.n_sets_nocut = w->rsh->aux_n_sets_nocut; ax.max_set_tids = w->rsh->aux_max_set_tids;
        d.aux = &ax;
    }
    rc = emsar_index_create(ctx, &d, &ix);
    if (rc) die("%s: %s", emsar_cuda_strerror(rc), emsar_cuda_last_error());
    if (w->worker == 0 && d.aux && o->verbose > 0) fprintf(stdout, "index image carries transpose, locality order and reachable classes: nothing derived at load\n");
    if (w->worker == 0 && getenv("EMSAR_RSH_CACHE") && !r->has_aux && o->rshfile[0] && !(strlen(o->rshfile) > 5 && !strcmp(o->rshfile + strlen(o->rshfile) - 5, ".pack"))) {
        /* complete the packed image: the arrays the library just derived go into <rshfile>.pack, so that the next run
         * of this index creates it without a pass over the members */
        emsar_index_aux got;
        if (emsar_index_aux_get(ix, &got) == 0) {
            emsar_rsh tmp = *r;
            tmp.has_aux = 1; tmp.aux_owned = 0; tmp.aux_nnz_multi = got.nnz_multi;
            tmp.aux_txm_off = (uint32_t *)got.txm_off; tmp.aux_txm_cid = (int32_t *)got.txm_cid; tmp.aux_order = (int32_t *)got.order;
            tmp.aux_insertable = (uint8_t *)got.insertable; tmp.aux_n_sets_nocut = got.n_sets_nocut; tmp.aux_max_set_tids = got.max_set_tids;
            char pk[FILENAMEMAX + 16], e2[EMSAR_HOST_ERRLEN];
            snprintf(pk, sizeof pk, "%s.pack", o->rshfile);
            if (emsar_rsh_save_packed(&tmp, pk, o->rshfile, e2) == 0 && o->verbose > 0) fprintf(stdout, "packed index image written: %s\n", pk);
        }
    }
    double eumacut = 0;
    if (w->by_class) {
        if (w->worker == 0 && (rc = emsar_comm_unique_id(w->comm_id))) die("%s: %s", emsar_cuda_strerror(rc), emsar_cuda_last_error());
        pthread_barrier_wait(w->bar);
        if ((rc = emsar_comm_init(ctx, w->worker, w->nworker, w->comm_id))) die("%s: %s", emsar_cuda_strerror(rc), emsar_cuda_last_error());
        for (int i = 0; i < o->naln; i++) run_file(o, r, ctx, ix, i, &eumacut, w->worker, w->nworker);      /* every file on all GPUs */
        emsar_comm_destroy(ctx);
    } else {
        /* shared work queue, largest file first (LPT): a GPU takes the next file the moment it is free, so a list of uneven samples
         * finishes together. Output file numbers are the positions in the list, whoever computes them. One GPU: list order, and EUMAcut
         * carries over from file to file as in the reference (emsar.h:94 is never reset); several GPUs: every file starts from the
         * initial EUMAcut, so that the result does not depend on which GPU took which file. */
        for (;;) {
            const int k = w->nworker > 1 ? __atomic_fetch_add(w->next, 1, __ATOMIC_RELAXED) : w->serial++;
            if (k >= o->naln) break;
            if (w->nworker > 1) eumacut = 0;
            run_file(o, r, ctx, ix, w->nworker > 1 ? w->order[k] : k, &eumacut, 0, 1);
        }
    }
    emsar_index_destroy(ix);
    emsar_cuda_close(ctx);
    return NULL;
}

int main(int argc, char *argv[])
{
    options o;
    memset(&o, 0, sizeof o);
    if (argc < 3) { usage(argv[0]); return 0; }
    static struct option long_options[] = {
        {"rsh", required_argument, 0, 'I'}, {"fasta", required_argument, 0, 'x'}, {"print_segments", no_argument, 0, 'g'},
        {"print_sfa", no_argument, 0, 'T'}, {"print_rsh", no_argument, 0, 'R'}, {"BAM", no_argument, 0, 'B'}, {"SAM", no_argument, 0, 'S'},
        {"PE", no_argument, 0, 'P'}, {"strand_type", required_argument, 0, 's'}, {"multisample", no_argument, 0, 'M'},
        {"bias_model", required_argument, 0, 'm'}, {"posbias_training_len", required_argument, 0, 'W'},
        {"posbias_impute_len", required_argument, 0, 'w'}, {"binsize", required_argument, 0, 'b'}, {"maxthread", required_argument, 0, 'p'},
        {"header", required_argument, 0, 'h'}, {"taglen", required_argument, 0, 't'}, {"maxfraglen", required_argument, 0, 'F'},
        {"minfraglen", required_argument, 0, 'f'}, {"max_repeat", required_argument, 0, 'k'}, {"nround", required_argument, 0, 'n'},
        {"epsilon", required_argument, 0, 'e'}, {"precision", required_argument, 0, 'r'}, {"delta", required_argument, 0, 'd'},
        {"max_niter_mle", required_argument, 0, 'i'}, {"max_nloop_mle", required_argument, 0, 'l'}, {"verbose", no_argument, 0, 'v'},
        {"no_verbose", no_argument, 0, 'q'}, {0, 0, 0, 0}};
    /* defaults (emsar_main.c:64-91) */
    strcpy(o.strand_str, "ns");
    o.max_fl = 400; o.min_fl = 1; o.max_repeat = 100; o.verbose = 1; o.nthread = 0; o.num_round = 4;
    int c, oi;
    while ((c = getopt_long(argc, argv, "vqPs:b:p:h:t:F:f:n:e:r:p:d:gm:MHBSW:w:k:i:l:TRI:x:", long_options, &oi)) != -1) {
        switch (c) {
        case 'I': strncpy(o.rshfile, optarg, FILENAMEMAX - 1); break;
        case 'x': strncpy(o.fasta, optarg, FILENAMEMAX - 1); break;
        case 'P': o.pe = 1; break;
        case 's': strncpy(o.strand_str, optarg, sizeof(o.strand_str) - 1); break;
        case 'h': o.fasta_header = optarg[0]; if (o.fasta_header != 'E' && o.fasta_header != 'R') die("error: invalid fasta option."); break;
        case 'b': case 't': case 'l': case 'H': break;                           /* suffix-array / MLE-loop knobs: accepted, unused */
        case 'p': o.nthread = atoi(optarg); if (o.nthread < 1) die("error: number of threads must be at least 1 (option -p)."); break;
        case 'F': o.max_fl = atoi(optarg); break;
        case 'f': o.min_fl = atoi(optarg); break;
        case 'k': o.max_repeat = atoi(optarg); break;
        case 'n': o.rounds_set = 1; o.num_round = atoi(optarg); if (o.num_round <= 0) { fprintf(stderr, "option -n must be a natural number.\n"); return 0; } break;
        case 'e': o.eps_abs = atof(optarg); if (o.eps_abs <= 0) { fprintf(stderr, "option -e must be positive.\n"); return 0; } break;
        case 'r': o.eps_rel = atof(optarg); if (o.eps_rel <= 0) { fprintf(stderr, "option -p must be positive.\n"); return 0; } break;
        case 'i': o.max_iter = atoi(optarg); if (o.max_iter <= 0) { fprintf(stderr, "option -i must be positive.\n"); return 0; } break;
        case 'd': o.delta = atoi(optarg); break;
        case 'g': o.print_segments = 1; break;
        case 'm': if (optarg[0] != '0') die("positional bias model (-m 1) is not supported by this build."); break;
        case 'W': case 'w': break;
        case 'M': o.multisample = 1; break;
        case 'B': if (o.bamflag == 's') { fprintf(stderr, "error: Options -B(--BAM) and -S(--SAM) cannot be used simultaneously.\n"); return 0; } o.bamflag = 'b'; break;
        case 'S': if (o.bamflag == 'b') { fprintf(stderr, "error: Options -B(--BAM) and -S(--SAM) cannot be used simultaneously.\n"); return 0; } o.bamflag = 's'; break;
        case 'T': die("-T (print suffix array) belongs to index construction: use emsar-build.");
        case 'R': o.print_rsh = 1; break;
        case 'v': o.verbose = 2; break;
        case 'q': o.verbose = 0; break;
        case '?': fprintf(stderr, "error: unknown option?\n"); /* fallthrough */
        default: return 0;
        }
    }
    if (strlen(o.rshfile) == 0 && strlen(o.fasta) == 0) die("error: either fasta file or an rsh file must be used as an input.");
    if (o.min_fl > o.max_fl || o.min_fl < 1 || o.max_fl < 1) die("error: invalid fragment length range.");
    /* set_library_strand_type (:16-22); unlike the reference an unknown type IS an error here */
    if (!strcmp(o.strand_str, "ns")) o.strand = 0;
    else if (!strcmp(o.strand_str, "ssf") && !o.pe) o.strand = '+';
    else if (!strcmp(o.strand_str, "ssr") && !o.pe) o.strand = '-';
    else if (!strcmp(o.strand_str, "ssfr") && o.pe) o.strand = '+';
    else if (!strcmp(o.strand_str, "ssrf") && o.pe) o.strand = '-';
    else die("error: invalid strand type.");
    if (o.max_repeat > EMSAR_MAX_READ_TIDS) die("error: -k %d exceeds the supported maximum of %d alignments per read.", o.max_repeat, EMSAR_MAX_READ_TIDS);
    if (optind + 1 >= argc) { usage(argv[0]); return 0; }
    o.outdir = argv[optind]; o.outprefix = argv[optind + 1];
    const char *third = optind + 2 < argc ? argv[optind + 2] : "";
    o.aln = (char **)malloc(sizeof(char *) * MAX_nALNFILES);
    if (!o.multisample) { o.aln[0] = strdup(third); o.naln = 1; }
    else {
        FILE *lf = fopen(third, "r");
        if (!lf) { fprintf(stderr, "Can't open alignment list file.\n"); return 1; }
        char line[FILENAMEMAX];
        while (fgets(line, sizeof line, lf) && o.naln < MAX_nALNFILES) {
            size_t n = strlen(line);
            if (n && line[n - 1] == '\n') line[n - 1] = 0;
            o.aln[o.naln++] = strdup(line);
        }
        fclose(lf);
    }
    if (o.naln == 0) { fprintf(stderr, "No alignment files in the alignment list\n"); return 1; }
    if (o.verbose > 0) {
        if (strlen(o.rshfile) == 0) fprintf(stdout, "input fastafile name= %s\n", o.fasta);
        fprintf(stdout, "input rshfile name= %s\nInput type= %s\nPaired-end= %c\nstrand type= %s\nMultisample= %c\nMAX_REPEAT= %d\n", o.rshfile,
                o.bamflag == 0 ? "default bowtie output" : (o.bamflag == 's' ? "SAM" : "BAM"), o.pe ? 'y' : 'n', o.strand_str, o.multisample ? 'y' : 'n', o.max_repeat);
        fprintf(stdout, "print segments = %c\nprint rsh structure = %c\nfinished reading options and arguments..\n", o.print_segments ? 'y' : 'n', o.print_rsh ? 'y' : 'n');
    }
    char cmd[FILENAMEMAX + 16];
    snprintf(cmd, sizeof cmd, "mkdir -p %s", o.outdir);
    fflush(stdout);
    if (system(cmd) != 0) die("can't create output directory %s", o.outdir);
    stamp(&o, "reading rsh array...");
    char err[EMSAR_HOST_ERRLEN] = "";
    emsar_rsh *rsh = NULL;
    int from_cache = 0;       /* <rshfile>.pack (written when EMSAR_RSH_CACHE is set) replaces the text parse while it is fresh */
    if (strlen(o.rshfile) == 0) {
        /* -x: learn the read length(s) from the first alignment file, then build the index from the fasta (emsar_main.c:300-346) */
        emsar_build_opts bo;
        memset(&bo, 0, sizeof bo);
        bo.pe = o.pe; bo.stranded = o.strand != 0; bo.max_repeat = o.max_repeat; bo.header = o.fasta_header ? o.fasta_header : 'E';
        bo.min_fraglength = o.min_fl; bo.max_fraglength = o.max_fl; bo.threads = o.nthread > 0 ? o.nthread : 4;
        int rl0 = 0, rl1 = 0;
        if (emsar_sniff_readlengths(o.aln[0], o.bamflag, o.pe, &rl0, &rl1, err)) die("%s", err);
        if (o.pe) bo.readlength = rl0; else { bo.readlen_min = rl0; bo.readlen_max = rl1; }
        if (o.verbose > 0) {
            if (o.pe) fprintf(stdout, "read length : %d\n", rl0);
            else fprintf(stdout, "read length range : %d - %d\n", rl0, rl1);
        }
        stamp(&o, "building the rsh index from the fasta file...");
        /* the classes are constructed on the first device (emsar_build_classes_run); EMSAR_BUILD_HOST=1 keeps the construction
         * on the host (emsar_b200/host/build_index.c) - the index is the same file byte for byte */
        emsar_ctx *bctx = NULL;
        const char *bh = getenv("EMSAR_BUILD_HOST");
        if (!(bh && atoi(bh))) {
            int dev0 = 0;
            const char *denv = getenv("EMSAR_DEVICES");
            if (denv && *denv) dev0 = atoi(denv);
            int brc = emsar_cuda_open(dev0, &bctx);
            if (brc) die("%s: %s", emsar_cuda_strerror(brc), emsar_cuda_last_error());
            bo.device_run = (int (*)(void *, const struct emsar_build_desc *, struct emsar_build_classes *))emsar_build_classes_run;
            bo.device_free = emsar_build_classes_free;
            bo.device_error = emsar_cuda_last_error;
            bo.device_ctx = bctx;
        }
        if (emsar_rsh_build(o.fasta, &bo, &rsh, err)) { printf("%s\n", err); exit(1); }
        if (bctx) emsar_cuda_close(bctx);
    } else if (emsar_rsh_load_auto(o.rshfile, &rsh, &from_cache, err)) { printf("%s\n", err); exit(1); }
    if (from_cache && o.verbose > 0) fprintf(stdout, "rsh index taken from its packed image\n");
    fprintf(stderr, "done reading rsh. rshsize=%lld\n", (long long)(rsh->C - rsh->T));
    if (o.verbose > 0) fprintf(stdout, "max_tid=%d, rshsize=%lld, max_cid=%lld\n", rsh->T - 1, (long long)(rsh->C - rsh->T), (long long)rsh->C - 1);
    if (o.print_rsh) {
        char p[FILENAMEMAX * 2 + 16];
        snprintf(p, sizeof p, "%s/%s.rsh", o.outdir, o.outprefix);
        if (emsar_rsh_write(rsh, o.pe, p, err)) die("%s", err);
    }
    stamp(&o, "reading alignment file(s)...");
    /* devices: EMSAR_DEVICES="0,1,2" (default "0") — files of a -M list are spread over them, no communication */
    int devs[64], ndev = 0;
    const char *env = getenv("EMSAR_DEVICES");
    if (env && *env) { char *dup = strdup(env), *sv = NULL; for (char *t = strtok_r(dup, ",", &sv); t && ndev < 64; t = strtok_r(NULL, ",", &sv)) devs[ndev++] = atoi(t); free(dup); }
    if (ndev == 0) devs[ndev++] = 0;
    const char *shard = getenv("EMSAR_SHARD");
    const int by_class = shard && !strcmp(shard, "classes") && ndev > 1;
    if (by_class && ndev > 8) die("error: EMSAR_SHARD=classes supports at most 8 GPUs.");
    if (!by_class && ndev > o.naln) ndev = o.naln;
    pthread_t th[64];
    worker_arg wa[64];
    uint8_t comm_id[128];
    /* the -M work queue: files by decreasing size (a proxy for the number of reads), ties in list order */
    int queue_next = 0;
    int *order = (int *)malloc(sizeof(int) * (size_t)(o.naln > 0 ? o.naln : 1));
    {
        long long *sz = (long long *)malloc(sizeof(long long) * (size_t)(o.naln > 0 ? o.naln : 1));
        for (int i = 0; i < o.naln; i++) { struct stat sb; sz[i] = stat(o.aln[i], &sb) == 0 ? (long long)sb.st_size : 0; order[i] = i; }
        for (int i = 1; i < o.naln; i++) {          /* insertion sort: stable, the list is short */
            const int x = order[i]; int j = i - 1;
            while (j >= 0 && sz[order[j]] < sz[x]) { order[j + 1] = order[j]; j--; }
            order[j + 1] = x;
        }
        free(sz);
    }
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, NULL, (unsigned)ndev);
    for (int w = 0; w < ndev; w++) {
        wa[w].o = &o; wa[w].rsh = rsh; wa[w].device = devs[w]; wa[w].worker = w; wa[w].nworker = ndev; wa[w].rc = 0;
        wa[w].next = &queue_next; wa[w].order = order; wa[w].serial = 0;
        wa[w].by_class = by_class; wa[w].comm_id = comm_id; wa[w].bar = &bar;
    }
    for (int w = 1; w < ndev; w++) pthread_create(&th[w], NULL, worker, &wa[w]);
    worker(&wa[0]);
    for (int w = 1; w < ndev; w++) pthread_join(th[w], NULL);
    free(order);
    stamp(&o, "freeing rsh array ...");
    emsar_rsh_free(rsh);
    return 0;
}
