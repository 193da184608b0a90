/* Alignment readers with the reference's read-group semantics:
 *   read_bowtie_SE/PE (reference emsar_functions.c:707-836), parse_bowtieline[_PE] (:552-587, :612-703),
 *   read_BAM_SE/PE (:323-388, :474-548), convert_bam_alignment_2_alignment[_PE] (:391-469),
 *   add_alignment_to_list / check_fraglen_discrepancy / parse_mmstr / check_mate_readid_matching (alignment.c).
 * Output: batches of read groups (tids of the kept alignments + the first fragment length) for the device counter.
 * SAM text and BGZF/BAM are decoded here directly over zlib (the reference vendors samtools 0.1.19 for this). */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <zlib.h>

#include "emsar_host.h"

static int fail(char *err, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    if (err) vsnprintf(err, EMSAR_HOST_ERRLEN, fmt, ap);
    va_end(ap);
    return 1;
}

/* ---- small helpers shared with the tests ------------------------------------------------------------ */
int emsar_parse_mmstr(const char *s)
{
    int mm = 0;
    size_t n = strlen(s);
    if (n > 0) mm++;
    for (size_t i = 0; i < n; i++) if (s[i] == ',') mm++;
    return mm;
}

int emsar_parse_sam_mmstr(const char *s)
{
    int mm = 0;
    for (; *s; s++) if (*s < '0' || *s > '9') mm++;   /* any non-numeric character adds to mm (:421) */
    return mm;
}

int emsar_check_mate_readid_matching(const char *a, const char *b)
{
    size_t la = strlen(a), lb = strlen(b);
    if (la != lb) return 0;
    int slen = (int)la;
    /* operator precedence kept as written in alignment.c:119: A && B && (C || (D && E)) */
    if (slen >= 2 && a[slen - 2] == '/' && b[slen - 2] == '/' &&
        ((a[slen - 1] == '1' && b[slen - 1] == '2') || ((a[slen - 1] == '2' && b[slen - 1] == '1') && strncmp(a, b, (size_t)slen - 2) == 0)))
        return slen - 2;
    for (int i = 0; i < slen; i++) {
        if (a[i] == ' ' && b[i] == ' ') return i;
        if (a[i] != b[i]) return 0;
    }
    return slen;
}

int emsar_filter_group(int n, const int *tid, const int *mm, const int *fraglen, const int *pos, int max_repeat, int pe, int *keep)
{
    int size = 0, cur = 10000;
    for (int i = 0; i < n; i++) {
        int dup = 0;
        for (int j = 0; j < size; j++) {
            int q = keep[j];
            if (tid[i] == tid[q] && pos[i] == pos[q] && fraglen[i] == fraglen[q]) { dup = 1; break; }
        }
        if (dup) continue;
        if (mm[i] > cur) continue;
        if (mm[i] < cur) { size = 0; cur = mm[i]; }
        keep[size++] = i;
    }
    if (size > max_repeat) return -1;
    if (pe) for (int j = 1; j < size; j++) if (fraglen[keep[j]] != fraglen[keep[0]]) return -1;
    return size;
}

/* ---- group accumulator: the alignment_list of one read id, streamed --------------------------------- */
typedef struct {
    const emsar_reader_opts *o;
    emsar_batch_fn fn;
    void *user;
    int64_t batch;
    /* current group */
    int *tid, *pos, *fl;
    int size, cap, cur_min;
    char *prev_id; size_t prev_cap; int have_prev;
    /* output batch: `nbuf` buffer sets used in turn (2 = the parser fills one while the device still copies the other) */
    int64_t *rptr; int32_t *rtid; int32_t *rfl;
    int64_t nr, ntid, cap_r, cap_t;
    struct { int64_t *rptr; int32_t *rtid; int32_t *rfl; int64_t cap_t; } set[2];
    int nbuf, cur_set;
    int rc;
} grouper;

static void *g_alloc(grouper *g, size_t bytes) { return g->o->buf_alloc ? g->o->buf_alloc(g->o->hook_user, bytes) : malloc(bytes); }
static void g_free(grouper *g, void *p) { if (!p) return; if (g->o->buf_free) g->o->buf_free(g->o->hook_user, p); else free(p); }

static void g_init(grouper *g, const emsar_reader_opts *o, emsar_batch_fn fn, void *user)
{
    memset(g, 0, sizeof(*g));
    g->o = o; g->fn = fn; g->user = user;
    g->batch = o->batch_reads > 0 ? o->batch_reads : (1 << 20);
    g->cur_min = 10000;
    g->cap_r = g->batch + 1;
    g->nbuf = (o->nbuf == 2) ? 2 : 1;
    for (int b = 0; b < g->nbuf; b++) {
        g->set[b].cap_t = g->batch * 4;
        g->set[b].rptr = (int64_t *)g_alloc(g, sizeof(int64_t) * (size_t)(g->cap_r + 1));
        g->set[b].rfl = (int32_t *)g_alloc(g, sizeof(int32_t) * (size_t)g->cap_r);
        g->set[b].rtid = (int32_t *)g_alloc(g, sizeof(int32_t) * (size_t)g->set[b].cap_t);
    }
    g->cur_set = 0;
    g->rptr = g->set[0].rptr; g->rfl = g->set[0].rfl; g->rtid = g->set[0].rtid; g->cap_t = g->set[0].cap_t;
    g->rptr[0] = 0;
}

static void g_emit_batch(grouper *g)
{
    if (g->nr > 0 && !g->rc) {
        g->rc = g->fn(g->user, g->nr, g->rptr, g->rtid, g->rfl);
        if (g->nbuf == 2) {        /* the callback may still be copying this set: go on in the other one */
            g->set[g->cur_set].rtid = g->rtid; g->set[g->cur_set].cap_t = g->cap_t;
            g->cur_set ^= 1;
            g->rptr = g->set[g->cur_set].rptr; g->rfl = g->set[g->cur_set].rfl; g->rtid = g->set[g->cur_set].rtid; g->cap_t = g->set[g->cur_set].cap_t;
        }
    }
    g->nr = 0; g->ntid = 0; g->rptr[0] = 0;
}

/* end of a read group: size <= MAX_REPEAT (and, PE, one fragment length) -> update_ReadCounts */
static void g_flush_group(grouper *g)
{
    if (g->size > 0 && g->size <= g->o->max_repeat) {
        int ok = 1;
        if (g->o->pe) for (int j = 1; j < g->size; j++) if (g->fl[j] != g->fl[0]) { ok = 0; break; }
        if (ok) {
            if (g->ntid + g->size > g->cap_t) {
                while (g->ntid + g->size > g->cap_t) g->cap_t *= 2;
                int32_t *nt = (int32_t *)g_alloc(g, sizeof(int32_t) * (size_t)g->cap_t);
                memcpy(nt, g->rtid, sizeof(int32_t) * (size_t)g->ntid);
                g_free(g, g->rtid);
                g->rtid = nt;
            }
            for (int j = 0; j < g->size; j++) g->rtid[g->ntid + j] = g->tid[j];
            g->ntid += g->size;
            g->rfl[g->nr] = g->fl[0];
            g->nr++;
            g->rptr[g->nr] = g->ntid;
            if (g->nr >= g->batch) g_emit_batch(g);
        }
    }
    g->size = 0;
}

/* add_alignment_to_list (alignment.c:29-60) */
static void g_add(grouper *g, int tid, int mm, int fl, int pos)
{
    for (int j = 0; j < g->size; j++)
        if (g->tid[j] == tid && g->pos[j] == pos && g->fl[j] == fl) return;   /* duplicate: dropped before the mm test */
    if (mm > g->cur_min) return;
    if (mm < g->cur_min) { g->size = 0; g->cur_min = mm; }
    if (g->size == g->cap) {
        g->cap = g->cap ? g->cap * 2 : 64;
        g->tid = (int *)realloc(g->tid, sizeof(int) * g->cap);
        g->pos = (int *)realloc(g->pos, sizeof(int) * g->cap);
        g->fl = (int *)realloc(g->fl, sizeof(int) * g->cap);
    }
    g->tid[g->size] = tid; g->pos[g->size] = pos; g->fl[g->size] = fl; g->size++;
}

/* one non-NULL alignment with its read id (the loop body shared by all four readers, e.g. :748-759) */
static void g_alignment(grouper *g, const char *read_id, int tid, int mm, int fl, int pos)
{
    if (g->have_prev && strcmp(g->prev_id, read_id) == 0) { g_add(g, tid, mm, fl, pos); }
    else {
        if (g->have_prev) g_flush_group(g);
        g->cur_min = 10000;
        g->size = 0;
        g_add(g, tid, mm, fl, pos);
    }
    size_t n = strlen(read_id) + 1;
    if (n > g->prev_cap) { g->prev_cap = n * 2; g->prev_id = (char *)realloc(g->prev_id, g->prev_cap); }
    memcpy(g->prev_id, read_id, n);
    g->have_prev = 1;
}

static int g_finish(grouper *g)
{
    g_flush_group(g);      /* the last group is flushed after EOF (:761) */
    g_emit_batch(g);
    int rc = g->rc;
    g->set[g->cur_set].rtid = g->rtid;
    free(g->tid); free(g->pos); free(g->fl); free(g->prev_id);
    for (int b = 0; b < g->nbuf; b++) { g_free(g, g->set[b].rptr); g_free(g, g->set[b].rtid); g_free(g, g->set[b].rfl); }
    return rc;
}

/* ---- bowtie default output ---------------------------------------------------------------------------- */
typedef struct { char *id; char strand; char *tname; int pos; int seqlen; char *mm; int nfield; } btline;

static void bt_parse(char *line, btline *b)
{
    /* fields split on tabs; 1 id, 2 strand, 3 tname, 4 pos, 5 seq, 8 mismatches (:561-577) */
    static char empty[1] = "";
    char *f[9];
    int n = 0;
    char *p = line;
    f[n++] = p;
    for (; *p; p++) if (*p == '\t') { *p = 0; if (n < 9) f[n++] = p + 1; else { n++; } }
    b->nfield = n;
    b->id = f[0];
    b->strand = n > 1 ? f[1][0] : 0;
    b->tname = n > 2 ? f[2] : empty;
    b->pos = n > 3 ? atoi(f[3]) : 0;
    b->seqlen = n > 4 ? (int)strlen(f[4]) : 0;
    b->mm = n > 7 ? f[7] : empty;
}

static int read_bowtie(const emsar_rsh *r, FILE *fp, const emsar_reader_opts *o, int *readlength, grouper *g, char *err)
{
    char *l1 = NULL, *l2 = NULL, *id = NULL;
    size_t c1 = 0, c2 = 0, cid = 0;
    ssize_t n1;
    int rc = 0;
    while (!rc && (n1 = getline(&l1, &c1, fp)) >= 0) {
        if (n1 > 0 && l1[n1 - 1] == '\n') l1[--n1] = 0;
        btline a;
        if (!o->pe) {
            bt_parse(l1, &a);
            if (o->strand != 0 && a.nfield > 1 && o->strand != a.strand) continue;   /* strand filter precedes everything (:568) */
            if (a.nfield < 7) { rc = fail(err, "Error: input alignment file doesn't look like bowtieout file."); break; }
            int tid = emsar_rsh_tid(r, a.tname);
            if (tid < 0) { rc = fail(err, "error: unexisting tid in the bowtie output file. Check bowtieout file."); break; }
            g_alignment(g, a.id, tid, emsar_parse_mmstr(a.mm), a.seqlen, a.pos);
        } else {
            ssize_t n2 = getline(&l2, &c2, fp);
            if (n2 < 0) { if (l2) l2[0] = 0; n2 = 0; }
            if (n2 > 0 && l2[n2 - 1] == '\n') l2[--n2] = 0;
            btline b;
            bt_parse(l1, &a);
            if (a.nfield < 7) { rc = fail(err, "Error: input alignment file doesn't look like bowtieout file."); break; }
            size_t need = strlen(a.id) + 1;
            if (need > cid) { cid = need * 2; id = (char *)realloc(id, cid); }
            memcpy(id, a.id, need);
            bt_parse(l2, &b);
            int adj = emsar_check_mate_readid_matching(id, b.id);
            if (adj == 0) { rc = fail(err, "Error: mate read ID's don't match. Check bowtie out format."); break; }
            /* (*read_id)[len-1]==1 compares a char with the integer 1 (:652): order_reversed is always 1, so
             * the FIRST line is treated as mate 2 and the second as mate 1 */
            id[adj] = 0;
            if (b.nfield < 7) { rc = fail(err, "Error: input alignment file doesn't look like bowtieout file."); break; }
            if (strcmp(a.tname, b.tname) != 0) continue;                               /* :667 */
            if (*readlength == -1) *readlength = a.seqlen;
            if (*readlength != a.seqlen || *readlength != b.seqlen) {
                rc = fail(err, "Error: Paired-end data with variable read length is not supported. Check your bowtieout file."); break;
            }
            int tid = emsar_rsh_tid(r, a.tname);
            if (tid < 0) { rc = fail(err, "error: unexisting tid in the bowtie output file. Check bowtieout file."); break; }
            int pos1 = b.pos, pos2 = a.pos;              /* swapped (order_reversed) */
            char s1 = b.strand, s2 = a.strand;
            int mm = emsar_parse_mmstr(a.mm) + emsar_parse_mmstr(b.mm);
            int fl, pos;
            if (pos2 > pos1) {
                fl = pos2 - pos1 + *readlength; pos = pos1;
                if (o->strand == '-') continue;
                if (!(s1 == '+' && s2 == '-')) continue;
            } else {
                fl = pos1 - pos2 + *readlength; pos = pos2;
                if (o->strand == '+') continue;
                if (!(s1 == '-' && s2 == '+')) continue;
            }
            g_alignment(g, id, tid, mm, fl, pos);
        }
    }
    free(l1); free(l2); free(id);
    return rc;
}


/* ---- bowtie default output, SE, parsed by a pool of threads -------------------------------------------------
 * The reference parses and groups on one thread (read_bowtie_SE :707-762). Splitting a line into fields, the transcript-name lookup and
 * the mismatch string are independent per line; only the grouping by read id needs the file order. The reading thread cuts the file into
 * blocks that end on a line break, `nthr` workers turn every line of a block into a small record, and the reading thread consumes the
 * blocks in file order, feeding the records to the same grouper the single-threaded reader uses. Results are identical by construction
 * (tests/test_host_ingest_cpu.py). */
typedef struct { uint32_t id_off; int32_t tid, mm, fl, pos; int8_t status; } bt_rec;        /* status: 0 ok, 1 dropped by the strand filter, 2 malformed, 3 unknown transcript */
typedef struct { char *buf; size_t len, cap; bt_rec *rec; size_t nrec, rec_cap; int state; } bt_job;   /* state: 0 free, 1 queued, 2 running, 3 done */
typedef struct {
    const emsar_rsh *r; char strand;
    int nthr, njobs;
    pthread_t *thr;
    pthread_mutex_t mu;
    pthread_cond_t cv_work, cv_done;
    bt_job *jobs;
    unsigned long long n_read, n_taken;
    int stop;
} bt_mt;

static void bt_parse_block(const emsar_rsh *r, char strand, bt_job *j)
{
    j->nrec = 0;
    char *p = j->buf, *end = j->buf + j->len;
    while (p < end) {
        char *nl = (char *)memchr(p, '\n', (size_t)(end - p));
        if (!nl) nl = end;
        *nl = 0;
        if (j->nrec == j->rec_cap) { j->rec_cap = j->rec_cap ? j->rec_cap * 2 : 65536; j->rec = (bt_rec *)realloc(j->rec, sizeof(bt_rec) * j->rec_cap); }
        bt_rec *q = &j->rec[j->nrec++];
        btline a;
        bt_parse(p, &a);
        q->id_off = (uint32_t)(a.id - j->buf);
        q->status = 0; q->tid = -1; q->mm = 0; q->fl = 0; q->pos = 0;
        if (strand != 0 && a.nfield > 1 && strand != a.strand) q->status = 1;
        else if (a.nfield < 7) q->status = 2;
        else {
            q->tid = emsar_rsh_tid(r, a.tname);
            if (q->tid < 0) q->status = 3;
            else { q->mm = emsar_parse_mmstr(a.mm); q->fl = a.seqlen; q->pos = a.pos; }
        }
        p = nl + 1;
    }
}

static void *bt_worker(void *arg)
{
    bt_mt *h = (bt_mt *)arg;
    pthread_mutex_lock(&h->mu);
    for (;;) {
        while (!h->stop && h->n_taken == h->n_read) pthread_cond_wait(&h->cv_work, &h->mu);
        if (h->stop) break;
        bt_job *j = &h->jobs[h->n_taken % (unsigned long long)h->njobs];
        h->n_taken++;
        j->state = 2;
        pthread_mutex_unlock(&h->mu);
        bt_parse_block(h->r, h->strand, j);
        pthread_mutex_lock(&h->mu);
        j->state = 3;
        pthread_cond_broadcast(&h->cv_done);
    }
    pthread_mutex_unlock(&h->mu);
    return NULL;
}

static int bt_consume(bt_job *j, grouper *g, char *err)
{
    for (size_t i = 0; i < j->nrec; i++) {
        const bt_rec *q = &j->rec[i];
        if (q->status == 1) continue;
        if (q->status == 2) return fail(err, "Error: input alignment file doesn't look like bowtieout file.");
        if (q->status == 3) return fail(err, "error: unexisting tid in the bowtie output file. Check bowtieout file.");
        g_alignment(g, j->buf + q->id_off, q->tid, q->mm, q->fl, q->pos);
    }
    return 0;
}

static int read_bowtie_se_mt(const emsar_rsh *r, FILE *fp, const emsar_reader_opts *o, grouper *g, char *err)
{
    enum { BLOCK = 4 << 20 };
    bt_mt h;
    memset(&h, 0, sizeof h);
    h.r = r; h.strand = o->strand; h.nthr = o->io_threads; h.njobs = 2 * o->io_threads + 2;
    h.jobs = (bt_job *)calloc((size_t)h.njobs, sizeof(bt_job));
    h.thr = (pthread_t *)calloc((size_t)h.nthr, sizeof(pthread_t));
    pthread_mutex_init(&h.mu, NULL); pthread_cond_init(&h.cv_work, NULL); pthread_cond_init(&h.cv_done, NULL);
    for (int i = 0; i < h.nthr; i++) pthread_create(&h.thr[i], NULL, bt_worker, &h);
    char *carry = NULL; size_t ncarry = 0, carry_cap = 0;
    unsigned long long n_consumed = 0;
    int rc = 0, eof = 0;
    while (!rc && (!eof || n_consumed < h.n_read)) {
        /* keep the ring full: read blocks ahead */
        while (!eof && h.n_read - n_consumed < (unsigned long long)h.njobs) {
            bt_job *j = &h.jobs[h.n_read % (unsigned long long)h.njobs];
            if (j->cap < ncarry + BLOCK + 1) { j->cap = ncarry + BLOCK + 1; j->buf = (char *)realloc(j->buf, j->cap); }
            if (ncarry) memcpy(j->buf, carry, ncarry);
            size_t got = fread(j->buf + ncarry, 1, BLOCK, fp);
            size_t len = ncarry + got;
            ncarry = 0;
            if (got < (size_t)BLOCK) eof = 1;
            if (!eof) {                     /* cut at the last line break; the tail goes in front of the next block */
                size_t cut = len;
                while (cut > 0 && j->buf[cut - 1] != '\n') cut--;
                if (cut == 0) { rc = fail(err, "bowtie file: a line longer than %d bytes", (int)BLOCK); break; }
                ncarry = len - cut;
                if (ncarry > carry_cap) { carry_cap = ncarry * 2; carry = (char *)realloc(carry, carry_cap); }
                memcpy(carry, j->buf + cut, ncarry);
                len = cut;
            }
            if (len > 0 && j->buf[len - 1] == '\n') len--;          /* the last line of the block needs no terminator */
            j->len = len;
            if (len == 0 && eof) break;
            pthread_mutex_lock(&h.mu);
            j->state = 1;
            h.n_read++;
            pthread_cond_signal(&h.cv_work);
            pthread_mutex_unlock(&h.mu);
        }
        if (rc || n_consumed == h.n_read) continue;
        bt_job *j = &h.jobs[n_consumed % (unsigned long long)h.njobs];
        pthread_mutex_lock(&h.mu);
        while (j->state != 3) pthread_cond_wait(&h.cv_done, &h.mu);
        pthread_mutex_unlock(&h.mu);
        rc = bt_consume(j, g, err);
        j->state = 0;
        n_consumed++;
    }
    pthread_mutex_lock(&h.mu);
    h.stop = 1;
    pthread_cond_broadcast(&h.cv_work);
    pthread_mutex_unlock(&h.mu);
    for (int i = 0; i < h.nthr; i++) pthread_join(h.thr[i], NULL);
    for (int i = 0; i < h.njobs; i++) { free(h.jobs[i].buf); free(h.jobs[i].rec); }
    free(h.jobs); free(h.thr); free(carry);
    pthread_mutex_destroy(&h.mu); pthread_cond_destroy(&h.cv_work); pthread_cond_destroy(&h.cv_done);
    return rc;
}

/* ---- BGZF with a pool of inflate threads ------------------------------------------------------
 * A BAM file is a sequence of independent gzip members of <= 64 KB (BGZF); the reference inflates them one at a time on
 * the thread that also parses (samtools 0.1.19 bgzf.c). Here the parsing thread only reads the compressed blocks ahead
 * (the block length is in the `BC` extra field) into a ring of jobs; `nthr` workers inflate them out of order and the
 * parser consumes them in file order. CRC32 and ISIZE of every block are checked. */
typedef struct { unsigned char *c, *u; int clen, ulen, isize, state, err; unsigned crc; } bgzf_job;   /* state: 0 free, 1 queued, 2 running, 3 done */
typedef struct {
    FILE *fp;
    int nthr, njobs;
    pthread_t *thr;
    pthread_mutex_t mu;
    pthread_cond_t cv_work, cv_done;
    bgzf_job *jobs;
    unsigned long long n_read, n_taken, n_consumed;     /* counters; slot = counter % njobs */
    int eof, stop, failed, have_cur;
    const unsigned char *cur; int cur_len, cur_off;
    unsigned char first[18]; int first_len;             /* bytes consumed by the format probe */
} bgzf_mt;

static void *bgzf_worker(void *arg)
{
    bgzf_mt *h = (bgzf_mt *)arg;
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    inflateInit2(&zs, -15);
    pthread_mutex_lock(&h->mu);
    for (;;) {
        while (!h->stop && h->n_taken == h->n_read) pthread_cond_wait(&h->cv_work, &h->mu);
        if (h->stop) break;
        bgzf_job *j = &h->jobs[h->n_taken % (unsigned long long)h->njobs];
        h->n_taken++;
        j->state = 2;
        pthread_mutex_unlock(&h->mu);
        inflateReset(&zs);
        zs.next_in = j->c; zs.avail_in = (uInt)j->clen;
        zs.next_out = j->u; zs.avail_out = 65536;
        int rc = inflate(&zs, Z_FINISH);
        j->ulen = (int)(65536 - zs.avail_out);
        j->err = !(rc == Z_STREAM_END && j->ulen == j->isize && (unsigned)crc32(crc32(0L, Z_NULL, 0), j->u, (uInt)j->ulen) == j->crc);
        pthread_mutex_lock(&h->mu);
        j->state = 3;
        pthread_cond_broadcast(&h->cv_done);
    }
    pthread_mutex_unlock(&h->mu);
    inflateEnd(&zs);
    return NULL;
}

/* reads the next compressed block into job slot j; returns 1 = block, 0 = clean EOF, -1 = malformed */
static int bgzf_fetch(bgzf_mt *h, bgzf_job *j)
{
    unsigned char hd[18];
    size_t got;
    if (h->first_len) { memcpy(hd, h->first, (size_t)h->first_len); got = (size_t)h->first_len; h->first_len = 0; }
    else got = fread(hd, 1, 12, h->fp);
    if (got == 0) return 0;
    if (got < 12) { got += fread(hd + got, 1, 12 - got, h->fp); }
    if (got < 12 || hd[0] != 31 || hd[1] != 139 || hd[2] != 8 || !(hd[3] & 4)) return -1;
    int xlen = hd[10] | (hd[11] << 8), bsize = -1;       /* 10 fixed bytes, then XLEN */
    unsigned char extra[256];
    if (xlen > (int)sizeof extra || fread(extra, 1, (size_t)xlen, h->fp) != (size_t)xlen) return -1;
    for (int o = 0; o + 4 <= xlen;) {
        int slen = extra[o + 2] | (extra[o + 3] << 8);
        if (extra[o] == 'B' && extra[o + 1] == 'C' && slen == 2 && o + 6 <= xlen) bsize = extra[o + 4] | (extra[o + 5] << 8);
        o += 4 + slen;
    }
    if (bsize < 0) return -1;
    int clen = bsize + 1 - 12 - xlen - 8;
    if (clen < 0 || clen > 65536) return -1;
    unsigned char tail[8];
    if (fread(j->c, 1, (size_t)clen, h->fp) != (size_t)clen || fread(tail, 1, 8, h->fp) != 8) return -1;
    j->clen = clen;
    j->crc = (unsigned)tail[0] | ((unsigned)tail[1] << 8) | ((unsigned)tail[2] << 16) | ((unsigned)tail[3] << 24);
    j->isize = (int)((unsigned)tail[4] | ((unsigned)tail[5] << 8) | ((unsigned)tail[6] << 16) | ((unsigned)tail[7] << 24));
    if (j->isize < 0 || j->isize > 65536) return -1;
    return 1;
}

static void bgzf_mt_close(bgzf_mt *h)
{
    if (!h) return;
    pthread_mutex_lock(&h->mu);
    h->stop = 1;
    pthread_cond_broadcast(&h->cv_work);
    pthread_mutex_unlock(&h->mu);
    for (int i = 0; i < h->nthr; i++) pthread_join(h->thr[i], NULL);
    for (int i = 0; i < h->njobs; i++) { free(h->jobs[i].c); free(h->jobs[i].u); }
    free(h->jobs); free(h->thr);
    if (h->fp && h->fp != stdin) fclose(h->fp);
    pthread_mutex_destroy(&h->mu); pthread_cond_destroy(&h->cv_work); pthread_cond_destroy(&h->cv_done);
    free(h);
}

/* NULL when the file cannot be opened or does not start with a BGZF block (plain gzip: the caller falls back to zlib's gz*) */
static bgzf_mt *bgzf_mt_open(const char *path, int nthr)
{
    FILE *fp = (path[0] == 0 || strcmp(path, "-") == 0) ? stdin : fopen(path, "rb");
    if (!fp) return NULL;
    if (fp == stdin) return NULL;                 /* stdin cannot be rewound for the fallback: leave it to gzdopen */
    unsigned char hd[12];
    if (fread(hd, 1, 12, fp) != 12 || hd[0] != 31 || hd[1] != 139 || hd[2] != 8 || !(hd[3] & 4)) { fclose(fp); return NULL; }
    bgzf_mt *h = (bgzf_mt *)calloc(1, sizeof *h);
    memcpy(h->first, hd, 12); h->first_len = 12;
    h->fp = fp;
    h->nthr = nthr < 1 ? 1 : (nthr > 64 ? 64 : nthr);
    h->njobs = 8 * h->nthr + 8;
    h->jobs = (bgzf_job *)calloc((size_t)h->njobs, sizeof(bgzf_job));
    for (int i = 0; i < h->njobs; i++) { h->jobs[i].c = (unsigned char *)malloc(65536); h->jobs[i].u = (unsigned char *)malloc(65536); }
    pthread_mutex_init(&h->mu, NULL); pthread_cond_init(&h->cv_work, NULL); pthread_cond_init(&h->cv_done, NULL);
    h->thr = (pthread_t *)calloc((size_t)h->nthr, sizeof(pthread_t));
    for (int i = 0; i < h->nthr; i++) pthread_create(&h->thr[i], NULL, bgzf_worker, h);
    return h;
}

/* like gzread: up to n bytes, fewer only at the end of the file; -1 on a malformed / corrupt block */
static int bgzf_mt_read(bgzf_mt *h, void *buf, unsigned n)
{
    unsigned done = 0;
    while (done < n) {
        if (h->have_cur && h->cur_off < h->cur_len) {
            unsigned take = (unsigned)(h->cur_len - h->cur_off);
            if (take > n - done) take = n - done;
            memcpy((char *)buf + done, h->cur + h->cur_off, take);
            h->cur_off += (int)take; done += take;
            continue;
        }
        if (h->failed) return -1;
        if (h->have_cur) {                        /* release the block just finished */
            pthread_mutex_lock(&h->mu);
            h->jobs[h->n_consumed % (unsigned long long)h->njobs].state = 0;
            h->n_consumed++;
            pthread_mutex_unlock(&h->mu);
            h->have_cur = 0;
        }
        /* keep the ring full: only this thread touches the file and n_read */
        while (!h->eof && h->n_read - h->n_consumed < (unsigned long long)h->njobs) {
            bgzf_job *j = &h->jobs[h->n_read % (unsigned long long)h->njobs];
            int st = bgzf_fetch(h, j);
            if (st == 0) { h->eof = 1; break; }
            if (st < 0) { h->failed = 1; h->eof = 1; break; }
            pthread_mutex_lock(&h->mu);
            j->state = 1;
            h->n_read++;
            pthread_cond_signal(&h->cv_work);
            pthread_mutex_unlock(&h->mu);
        }
        if (h->n_consumed == h->n_read) { if (h->failed) return -1; break; }      /* end of file */
        bgzf_job *j = &h->jobs[h->n_consumed % (unsigned long long)h->njobs];
        pthread_mutex_lock(&h->mu);
        while (j->state != 3) pthread_cond_wait(&h->cv_done, &h->mu);
        pthread_mutex_unlock(&h->mu);
        if (j->err) { h->failed = 1; return -1; }
        h->cur = j->u; h->cur_len = j->ulen; h->cur_off = 0; h->have_cur = 1;
    }
    return (int)done;
}

/* ---- SAM / BAM records ----------------------------------------------------------------------------------- */
typedef struct { char *qname; int flag; int ref; int pos; int l_qseq; const char *md; } samrec;   /* md NULL = no MD tag */

typedef struct {
    /* header */
    char **ref_names; int n_ref, cap_ref;
    int *ref_tid;       /* lazily resolved tid of each reference, -2 = not looked up yet */
    /* text SAM */
    FILE *fp; char *line; size_t cap;
    /* BAM */
    gzFile gz; int is_bam;
    bgzf_mt *mt;        /* threaded BGZF reader (NULL: zlib's gz* on a plain gzip stream or stdin) */
    unsigned char *blk; size_t blk_cap;
    char *md_buf; size_t md_cap;
    char *qbuf; size_t qcap;
    char *pending; /* first alignment line of a SAM file read while scanning the header */
    uint32_t *ref_slots; uint32_t ref_mask;   /* RNAME -> reference index (text SAM) */
} samfile;

static uint32_t fnv1a_(const char *s)
{
    uint32_t h = 2166136261u;
    for (; *s; s++) { h ^= (unsigned char)*s; h *= 16777619u; }
    return h;
}

static void sam_add_ref(samfile *s, const char *name)
{
    if (s->n_ref == s->cap_ref) {
        s->cap_ref = s->cap_ref ? s->cap_ref * 2 : 1024;
        s->ref_names = (char **)realloc(s->ref_names, sizeof(char *) * s->cap_ref);
        s->ref_tid = (int *)realloc(s->ref_tid, sizeof(int) * s->cap_ref);
    }
    s->ref_names[s->n_ref] = strdup(name);
    s->ref_tid[s->n_ref] = -2;
    s->n_ref++;
}

static void sam_build_ref_map(samfile *s)
{
    uint32_t slots = 16;
    while (slots < (uint32_t)s->n_ref * 2u) slots <<= 1;
    s->ref_mask = slots - 1;
    s->ref_slots = (uint32_t *)calloc(slots, sizeof(uint32_t));
    for (int i = 0; i < s->n_ref; i++) {
        uint32_t h = fnv1a_(s->ref_names[i]) & s->ref_mask;
        while (s->ref_slots[h] && strcmp(s->ref_names[s->ref_slots[h] - 1], s->ref_names[i]) != 0) h = (h + 1) & s->ref_mask;
        if (!s->ref_slots[h]) s->ref_slots[h] = (uint32_t)i + 1;      /* first @SQ of a name wins */
    }
}

static int sam_ref_index(samfile *s, const char *name)
{
    uint32_t h = fnv1a_(name) & s->ref_mask;
    for (;;) {
        uint32_t v = s->ref_slots[h];
        if (!v) return -1;
        if (strcmp(s->ref_names[v - 1], name) == 0) return (int)(v - 1);
        h = (h + 1) & s->ref_mask;
    }
}

static int bam_read(samfile *s, void *buf, unsigned n)
{
    return s->mt ? bgzf_mt_read(s->mt, buf, n) : gzread(s->gz, buf, n);
}

static int sam_open(samfile *s, const char *path, char fmt, int io_threads, char *err)
{
    memset(s, 0, sizeof(*s));
    s->is_bam = (fmt == 'b');
    if (s->is_bam) {
        if (io_threads > 0) s->mt = bgzf_mt_open(path, io_threads);
        if (!s->mt) {
            s->gz = (path[0] == 0 || strcmp(path, "-") == 0) ? gzdopen(0, "rb") : gzopen(path, "rb");   /* BGZF is a valid multi-member gzip stream */
            if (!s->gz) return fail(err, "can't open BAM file.");
            gzbuffer(s->gz, 1 << 20);
        }
        char magic[4];
        int32_t l_text, n_ref;
        if (bam_read(s, magic, 4) != 4 || memcmp(magic, "BAM\1", 4) != 0) return fail(err, "can't open BAM file.");
        if (bam_read(s, &l_text, 4) != 4) return fail(err, "truncated BAM header");
        if (l_text < 0) return fail(err, "corrupt BAM header (negative l_text)");
        char *text = (char *)malloc((size_t)l_text + 1);
        if (!text) return fail(err, "out of memory reading the BAM header");
        if (bam_read(s, text, (unsigned)l_text) != l_text) { free(text); return fail(err, "truncated BAM header"); }
        free(text);
        if (bam_read(s, &n_ref, 4) != 4) return fail(err, "truncated BAM header");
        if (n_ref < 0) return fail(err, "corrupt BAM header (negative n_ref)");
        for (int i = 0; i < n_ref; i++) {
            int32_t l_name, l_ref;
            if (bam_read(s, &l_name, 4) != 4) return fail(err, "truncated BAM header");
            if (l_name < 0) return fail(err, "corrupt BAM header (negative l_name)");
            char *nm = (char *)malloc((size_t)l_name + 1);
            if (!nm) return fail(err, "out of memory reading the BAM header");
            if (bam_read(s, nm, (unsigned)l_name) != l_name) { free(nm); return fail(err, "truncated BAM header"); }
            nm[l_name] = 0;
            if (bam_read(s, &l_ref, 4) != 4) { free(nm); return fail(err, "truncated BAM header"); }
            sam_add_ref(s, nm);
            free(nm);
        }
    } else {
        s->fp = (path[0] == 0 || strcmp(path, "-") == 0) ? stdin : fopen(path, "r");
        if (!s->fp) return fail(err, "can't open SAM file.");
        ssize_t n;
        while ((n = getline(&s->line, &s->cap, s->fp)) >= 0) {
            if (s->line[0] != '@') { s->pending = s->line; break; }
            if (strncmp(s->line, "@SQ", 3) == 0) {
                char *p = strstr(s->line, "\tSN:");
                if (p) {
                    p += 4;
                    char *e = p;
                    while (*e && *e != '\t' && *e != '\n') e++;
                    char c = *e; *e = 0;
                    sam_add_ref(s, p);
                    *e = c;
                }
            }
        }
        /* samtools 0.1.19 (the reference's reader) aborts on a text SAM without @SQ lines ("missing header"); every RNAME would resolve
           to -1 here and the run would "succeed" with zero counts */
        if (s->n_ref == 0) return fail(err, "SAM file has no @SQ header lines (missing header)");
        sam_build_ref_map(s);
    }
    return 0;
}

static void sam_close(samfile *s)
{
    if (s->gz) gzclose(s->gz);
    if (s->mt) bgzf_mt_close(s->mt);
    if (s->fp && s->fp != stdin) fclose(s->fp);
    for (int i = 0; i < s->n_ref; i++) free(s->ref_names[i]);
    free(s->ref_names); free(s->ref_tid); free(s->ref_slots); free(s->line); free(s->blk); free(s->md_buf); free(s->qbuf);
}

/* returns 1 = record, 0 = EOF, -1 = error */
static int sam_next(samfile *s, samrec *r, char *err)
{
    if (s->is_bam) {
        int32_t bs;
        int got = bam_read(s, &bs, 4);
        if (got == 0) return 0;
        if (got != 4 || bs < 32) { fail(err, "truncated BAM record"); return -1; }
        if ((size_t)bs + 1 > s->blk_cap) { s->blk_cap = (size_t)bs * 2 + 64; s->blk = (unsigned char *)realloc(s->blk, s->blk_cap); }
        if (bam_read(s, s->blk, (unsigned)bs) != bs) { fail(err, "truncated BAM record"); return -1; }
        const unsigned char *b = s->blk;
        int32_t refID, pos, l_seq;
        uint8_t l_read_name; uint16_t n_cigar, flag;
        memcpy(&refID, b, 4); memcpy(&pos, b + 4, 4);
        l_read_name = b[8];
        memcpy(&n_cigar, b + 12, 2); memcpy(&flag, b + 14, 2); memcpy(&l_seq, b + 16, 4);
        r->ref = refID; r->pos = pos; r->flag = flag; r->l_qseq = l_seq;
        r->qname = (char *)(b + 32);
        s->blk[bs] = 0;                      /* terminator first: nothing below can run past the record */
        if (l_seq < 0 || l_read_name == 0) { fail(err, "corrupt BAM record"); return -1; }
        size_t off = 32 + (size_t)l_read_name + 4 * (size_t)n_cigar + ((size_t)l_seq + 1) / 2 + (size_t)l_seq;
        if (off > (size_t)bs || b[32 + (size_t)l_read_name - 1] != 0) { fail(err, "corrupt BAM record (fields exceed the record)"); return -1; }
        r->md = NULL;
        /* aux fields: tag[2] type value; every length is checked against what is left of the record */
        while (off + 3 <= (size_t)bs) {
            const unsigned char *t = b + off;
            const size_t left = (size_t)bs - off - 3;
            char ty = (char)t[2];
            size_t vlen;
            if (ty == 'A' || ty == 'c' || ty == 'C') vlen = 1;
            else if (ty == 's' || ty == 'S') vlen = 2;
            else if (ty == 'i' || ty == 'I' || ty == 'f') vlen = 4;
            else if (ty == 'd') vlen = 8;
            else if (ty == 'Z' || ty == 'H') {
                const void *z = memchr(t + 3, 0, left);
                if (!z) { fail(err, "corrupt BAM record (unterminated aux string)"); return -1; }
                vlen = (size_t)((const unsigned char *)z - (t + 3)) + 1;
            } else if (ty == 'B') {
                if (left < 5) { fail(err, "corrupt BAM record (aux array)"); return -1; }
                char st = (char)t[3]; int32_t cnt; memcpy(&cnt, t + 4, 4);
                size_t es = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4;
                if (cnt < 0) { fail(err, "corrupt BAM record (aux array)"); return -1; }
                vlen = 5 + es * (size_t)cnt;
            } else break;
            if (vlen > left) { fail(err, "corrupt BAM record (aux field exceeds the record)"); return -1; }
            if (t[0] == 'M' && t[1] == 'D' && ty == 'Z') r->md = (const char *)t + 3;
            off += 3 + vlen;
        }
        s->blk[bs] = 0;
        return 1;
    }
    /* text SAM */
    ssize_t n;
    if (s->pending) { s->pending = NULL; n = (ssize_t)strlen(s->line); }
    else n = getline(&s->line, &s->cap, s->fp);
    if (n < 0) return 0;
    if (n > 0 && s->line[n - 1] == '\n') s->line[--n] = 0;
    char *f[12];
    int nf = 0;
    char *p = s->line;
    f[nf++] = p;
    char *rest = NULL;
    for (; *p; p++)
        if (*p == '\t') { *p = 0; if (nf < 11) f[nf++] = p + 1; else { rest = p + 1; break; } }
    if (nf < 11) { fail(err, "truncated SAM line"); return -1; }
    r->qname = f[0];
    r->flag = atoi(f[1]);
    r->ref = (f[2][0] == '*' && f[2][1] == 0) ? -1 : sam_ref_index(s, f[2]);
    if (r->ref == -1 && !(f[2][0] == '*' && f[2][1] == 0)) { /* unknown reference name: samtools maps it to tid -1 */ }
    r->pos = atoi(f[3]) - 1;
    r->l_qseq = (f[9][0] == '*' && f[9][1] == 0) ? 0 : (int)strlen(f[9]);
    r->md = NULL;
    for (char *t = rest; t && *t;) {
        char *e = strchr(t, '\t');
        if (e) *e = 0;
        if (strncmp(t, "MD:Z:", 5) == 0) r->md = t + 5;
        t = e ? e + 1 : NULL;
    }
    return 1;
}

static int sam_tid(const emsar_rsh *rs, samfile *s, int ref)
{
    if (s->ref_tid[ref] == -2) s->ref_tid[ref] = emsar_rsh_tid(rs, s->ref_names[ref]);
    return s->ref_tid[ref];
}

static int read_sam(const emsar_rsh *rs, const char *path, const emsar_reader_opts *o, int *readlength, grouper *g, char *err)
{
    samfile s;
    if (sam_open(&s, path, o->format, o->io_threads, err)) { sam_close(&s); return 1; }
    int rc = 0, st;
    samrec a, b;
    char *qn = NULL; size_t qcap = 0;
    const char *what = o->format == 'b' ? "bam/sam" : "bam/sam";
    (void)what;
    while (!rc && (st = sam_next(&s, &a, err)) != 0) {
        if (st < 0) { rc = 1; break; }
        if (a.ref == -1) continue;                                       /* skip unaligned reads (:359, :515) */
        if (!o->pe) {
            int tid = sam_tid(rs, &s, a.ref);
            if (tid < 0) { rc = fail(err, "error: unexisting tid in the bowtie output file. Check bowtieout file."); break; }
            char strand = (a.flag & 0x10) ? '-' : '+';
            if (o->strand != 0 && o->strand != strand) continue;         /* :402 */
            if (!a.md) { rc = fail(err, "error: alignment without MD tag (the reference requires MD:Z)."); break; }
            g_alignment(g, a.qname, tid, emsar_parse_sam_mmstr(a.md), a.l_qseq, a.pos);
        } else {
            /* the mate is the NEXT record whatever it is (:517); the pair uses the first record's name and reference */
            size_t need = strlen(a.qname) + 1;
            if (need > qcap) { qcap = need * 2; qn = (char *)realloc(qn, qcap); }
            memcpy(qn, a.qname, need);
            int a_flag = a.flag, a_ref = a.ref, a_pos = a.pos, a_lq = a.l_qseq;
            int a_mm = a.md ? emsar_parse_sam_mmstr(a.md) : -1;
            st = sam_next(&s, &b, err);
            if (st < 0) { rc = 1; break; }
            if (st == 0) break;                                          /* a trailing unpaired record contributes nothing */
            int tid = sam_tid(rs, &s, a_ref);
            if (tid < 0) { rc = fail(err, "error: unexisting tid in the bowtie output file. Check bam/sam file."); break; }
            if (*readlength == -1) *readlength = a_lq;
            if (*readlength != a_lq || *readlength != b.l_qseq) {
                rc = fail(err, "Error: Paired-end data with variable read length is not supported. Check your bam/sam file."); break;
            }
            int f1, f2, p1, p2, mm1, mm2;
            int b_mm = b.md ? emsar_parse_sam_mmstr(b.md) : -1;
            if ((a_flag & 0x40) && (b.flag & 0x80)) { f1 = a_flag; p1 = a_pos; mm1 = a_mm; f2 = b.flag; p2 = b.pos; mm2 = b_mm; }
            else if ((b.flag & 0x40) && (a_flag & 0x80)) { f1 = b.flag; p1 = b.pos; mm1 = b_mm; f2 = a_flag; p2 = a_pos; mm2 = a_mm; }
            else { rc = fail(err, "error: mates are not grouped in the BAM/SAM file."); break; }
            if (mm1 < 0 || mm2 < 0) { rc = fail(err, "error: alignment without MD tag (the reference requires MD:Z)."); break; }
            char s1 = (f1 & 0x10) ? '-' : '+', s2 = (f2 & 0x10) ? '-' : '+';
            int fl, pos;
            if (p2 > p1) {
                fl = p2 - p1 + *readlength; pos = p1;
                if (o->strand == '-') continue;
                if (!(s1 == '+' && s2 == '-')) continue;
            } else {
                fl = p1 - p2 + *readlength; pos = p2;
                if (o->strand == '+') continue;
                if (!(s1 == '-' && s2 == '+')) continue;
            }
            g_alignment(g, qn, tid, mm1 + mm2, fl, pos);
        }
    }
    free(qn);
    sam_close(&s);
    return rc;
}

int emsar_sniff_readlengths(const char *path, char format, int pe, int *rl_min, int *rl_max, char *err)
{
    int mn = 30000, mx = 0, seen = 0, rc = 0;           /* the reference's initial values (:272-273) */
    if (format == 0) {
        FILE *fp = (path[0] == 0) ? stdin : fopen(path, "r");
        if (!fp) return fail(err, "can't open bowtie file.");
        char *line = NULL; size_t cap = 0; ssize_t n;
        while ((n = getline(&line, &cap, fp)) >= 0) {
            if (n > 0 && line[n - 1] == '\n') line[--n] = 0;
            btline a;
            bt_parse(line, &a);
            if (a.nfield < 7) { rc = fail(err, "Error: input alignment file doesn't look like bowtieout file."); break; }
            if (a.seqlen < mn) mn = a.seqlen;
            if (a.seqlen > mx) mx = a.seqlen;
            seen = 1;
            if (pe) break;
        }
        free(line);
        if (fp != stdin) fclose(fp);
    } else {
        samfile s;
        if (sam_open(&s, path, format, 2, err)) { sam_close(&s); return 1; }
        samrec a;
        int st;
        while ((st = sam_next(&s, &a, err)) != 0) {
            if (st < 0) { rc = 1; break; }
            if (a.ref == -1) continue;
            if (a.l_qseq < mn) mn = a.l_qseq;
            if (a.l_qseq > mx) mx = a.l_qseq;
            seen = 1;
            if (pe) break;
        }
        sam_close(&s);
    }
    if (rc) return rc;
    if (!seen) return fail(err, "no aligned read in %s: cannot learn the read length", path);
    *rl_min = mn; *rl_max = mx;
    return 0;
}

int emsar_read_alignments(const emsar_rsh *r, const char *path, const emsar_reader_opts *o, int *readlength,
                          emsar_batch_fn fn, void *user, char *err)
{
    grouper g;
    g_init(&g, o, fn, user);
    int rc;
    if (o->format == 0) {
        FILE *fp = (path[0] == 0) ? stdin : fopen(path, "r");
        if (!fp) { g_finish(&g); return fail(err, "can't open bowtie file."); }
        if (!o->pe && o->io_threads > 1) rc = read_bowtie_se_mt(r, fp, o, &g, err);
        else rc = read_bowtie(r, fp, o, readlength, &g, err);
        if (fp != stdin) fclose(fp);
    } else {
        rc = read_sam(r, path, o, readlength, &g, err);
    }
    if (rc) { g.rc = 0; g.fn = NULL; g.nr = 0; g.size = 0; g_finish(&g); return rc; }
    rc = g_finish(&g);
    if (rc) return fail(err, "read batch callback failed (%d)", rc);
    return 0;
}
