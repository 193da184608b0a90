/* Host-side rsh index construction from a transcriptome fasta: what `emsar-build` / `emsar -x` of the reference produce
 * (parklab/emsar v2.0.1: read_raw_fasta emsar_functions.c:31-196, preprocess_SE / preprocess_PE :3243-3350,
 * initialize_suffixarray_* :949-1110, construct_rshbucket_2 :1758-1816, process_mate1_cluster_by_mate_3 :2820-2934,
 * construct_rshbucket_PE_3 :1902-1974, update_rshbucket[_single] 'e' :1514-1596, print_rsh :2071-2130).
 * index construction is outside the GPU hot path but the command line has to honour -x.
 *
 * The reference sorts suffix arrays of the concatenated transcriptome (one pass per 2-letter tag) and scans runs of equal
 * read-length substrings. Only the RESULT is specified - classes (sorted tid multisets) and their per-fragment-length counts -
 * so this implementation groups substrings by a 64-bit rolling hash and settles equal hashes by comparing the bases: one
 * qsort per read length (SE) or per mate-1 cluster (PE), no tags, no suffix arrays. The output is byte-identical to
 * `emsar-build` (tests/test_build_index_cpu.py runs both).
 *
 * Semantics kept (see the reference lines above):
 *  - sequence characters other than ACGT (either case) become N; substrings containing one, or crossing a transcript end,
 *    are ignored; transcripts shorter than the read length contribute nothing;
 *  - SE unstranded: an occurrence is represented by the smaller of itself and its reverse complement (tie: forward);
 *    SE stranded: forward only. A run of r equal substrings adds 1 to EUMA[class][len - min]: r = 1 -> the singleton class
 *    of that transcript, 2 <= r < MAX_REPEAT -> the class of the r tids (a transcript may repeat), else dropped;
 *  - PE: fragments (mate 1 at i, mate 2 at i + d on the same strand string, d = fragment length - read length) are keyed by
 *    (mate-1 bases, mate-2 bases); unstranded libraries keep the smaller of a fragment and its flip (tie: forward); a run with
 *    one member -> singleton, with several members -> class only if all share d and r < MAX_REPEAT, else dropped. */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "emsar_host.h"
#include "emsar_cuda.h"       /* emsar_build_desc / emsar_build_classes: the layout of the device builder's interface (no link dependency) */

static int fail(char *err, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    if (err) vsnprintf(err, EMSAR_HOST_ERRLEN, fmt, ap);
    va_end(ap);
    return 1;
}

/* ---- the transcriptome: S = f0 @ f1 @ ... @ fn $ rc(fn) @ ... @ rc(f0) $ ------------------------------------------- */
typedef struct {
    char *S;                 /* concatenated sequence, NUL terminated */
    int64_t border, end;     /* position of the first '$' and of the last one (= seqlength of the reference) */
    int32_t T;
    int64_t *start;          /* [T+1] first position of transcript t in the forward half; start[T] = border + 1 */
    char **names;
} txome;

static char up(int c)
{
    switch (c) {
    case 'A': case 'a': return 'A';
    case 'C': case 'c': return 'C';
    case 'G': case 'g': return 'G';
    case 'T': case 't': return 'T';
    case '@': return '@';
    case '$': return '$';
    default: return 'N';
    }
}
static char rc_of(char c)
{
    switch (c) {
    case 'A': return 'T';
    case 'C': return 'G';
    case 'G': return 'C';
    case 'T': return 'A';
    case '@': return '@';
    case '$': return '$';
    default: return 'N';
    }
}

static void parse_header(const char *h, char option, char *out)
{
    if (option == 'R') {                    /* >xx|xx|xx|name|xx : the 4th '|' field */
        int bars = 0, j = 0;
        for (const char *p = h; *p; p++) {
            if (*p == '|') { if (++bars == 4) break; }
            else if (bars == 3) out[j++] = *p;
        }
        out[j] = 0;
    } else {                                /* Ensembl: up to the first blank */
        int j = 0;
        for (const char *p = h; *p && *p != '\t' && *p != ' '; p++) out[j++] = *p;
        out[j] = 0;
    }
}

static void txome_free(txome *x)
{
    free(x->S); free(x->start);
    if (x->names) { for (int32_t t = 0; t < x->T; t++) free(x->names[t]); free(x->names); }
    memset(x, 0, sizeof *x);
}

static int txome_read(const char *path, char option, txome *x, char *err)
{
    memset(x, 0, sizeof *x);
    FILE *f = fopen(path, "r");
    if (!f) return fail(err, "can't open fasta file.");
    /* the reader of the reference is a character state machine: any '>' opens a header, the header ends at the newline,
     * blanks inside the sequence are skipped */
    size_t cap = 1 << 20, n = 0, hcap = 4096, hn = 0;
    char *seq = (char *)malloc(cap), *hdr = (char *)malloc(hcap);
    int64_t *start = NULL; char **names = NULL; int32_t T = 0, capT = 0;
    int c, prev = 0, mode = 'h', rc = 0;
    while ((c = fgetc(f)) != EOF) {
        if (prev == 0 && c != '>') { rc = fail(err, "ERROR: wrong fasta file format."); break; }
        if (c == '>') {
            mode = 'h';
            if (n != 0) { if (n + 2 > cap) { cap *= 2; seq = (char *)realloc(seq, cap); } seq[n++] = '@'; }
            else if (T > 0) { rc = fail(err, "fasta: the first sequence is empty"); break; }
        } else if (prev == '\n' && mode == 'h') mode = 's';
        if (mode == 'h') {
            if (c == '\n') {
                hdr[hn] = 0;
                if (T == capT) { capT = capT ? capT * 2 : 1024; start = (int64_t *)realloc(start, sizeof(int64_t) * (size_t)(capT + 1)); names = (char **)realloc(names, sizeof(char *) * (size_t)capT); }
                char *nm = (char *)malloc(hn + 1);
                parse_header(hdr, option, nm);
                names[T] = nm;
                start[T] = (int64_t)n;
                T++;
                hn = 0;
            } else if (c != '>') { if (hn + 2 > hcap) { hcap *= 2; hdr = (char *)realloc(hdr, hcap); } hdr[hn++] = (char)c; }
        } else if (c != '\n' && c != ' ' && c != '\t') {
            if (n + 2 > cap) { cap *= 2; seq = (char *)realloc(seq, cap); }
            seq[n++] = up(c);
        }
        prev = c;
    }
    fclose(f);
    free(hdr);
    if (!rc && T == 0) rc = fail(err, "fasta file holds no sequence");
    if (rc) { free(seq); free(start); for (int32_t t = 0; t < T; t++) free(names[t]); free(names); return rc; }
    /* a header that follows an empty sequence gets the position behind the '@' that was written for it */
    x->T = T; x->names = names; x->start = start;
    x->border = (int64_t)n;
    x->S = (char *)malloc(2 * n + 3);
    memcpy(x->S, seq, n);
    free(seq);
    x->S[n] = '$';
    for (size_t j = 0; j < n; j++) x->S[n + 1 + j] = rc_of(x->S[n - 1 - j]);
    x->end = (int64_t)(2 * n + 1);
    x->S[x->end] = '$';
    x->S[x->end + 1] = 0;
    /* transcript starts exactly as the reference derives them: the position after every '@' of the forward half */
    int32_t k = 1;
    x->start[0] = 0;
    for (int64_t i = 0; i < x->border && k < T; i++) if (x->S[i] == '@') x->start[k++] = i + 1;
    if (k != T) { txome_free(x); return fail(err, "fasta: empty sequences are not supported"); }
    x->start[T] = x->border + 1;
    return 0;
}

/* ---- class store ------------------------------------------------------------------------------------------------------ */
typedef struct { int32_t k; int64_t tid_off; int64_t euma_row; } bclass;
typedef struct {
    int nF; int32_t T;
    int32_t *s_euma; uint8_t *s_node;          /* singletons */
    bclass *cls; int64_t ncls, capcls;
    int32_t *tids; int64_t ntids, captids;
    int32_t *euma; int64_t nrows, caprows;      /* multi-class EUMA rows [nrows * nF] */
    uint64_t *slots; uint64_t mask;             /* open addressing: class index + 1 */
    int32_t max_k;
} cstore;

static uint64_t mix(uint64_t h) { h ^= h >> 33; h *= 0xff51afd7ed558ccdULL; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL; h ^= h >> 33; return h; }
static uint64_t key_hash(const int32_t *t, int k)
{
    uint64_t h = 0x9E3779B97F4A7C15ULL * (uint64_t)k;
    for (int i = 0; i < k; i++) h = mix(h ^ ((uint64_t)(uint32_t)t[i] + 0x632BE59BD9B4E019ULL * (uint64_t)(i + 1)));
    return h;
}
static void cs_init(cstore *c, int32_t T, int nF)
{
    memset(c, 0, sizeof *c);
    c->T = T; c->nF = nF;
    c->s_euma = (int32_t *)calloc((size_t)T * nF, sizeof(int32_t));
    c->s_node = (uint8_t *)calloc((size_t)T, 1);
    c->mask = (1u << 16) - 1;
    c->slots = (uint64_t *)calloc(c->mask + 1, sizeof(uint64_t));
    c->max_k = 10;                              /* INIT_RSHBUCKET_MAX_T_SIZE (emsar.h:16): the header never shows less */
}
static void cs_free(cstore *c) { free(c->s_euma); free(c->s_node); free(c->cls); free(c->tids); free(c->euma); free(c->slots); }
static void cs_rehash(cstore *c)
{
    uint64_t nmask = c->mask * 2 + 1;
    uint64_t *ns = (uint64_t *)calloc(nmask + 1, sizeof(uint64_t));
    for (int64_t i = 0; i < c->ncls; i++) {
        uint64_t s = key_hash(c->tids + c->cls[i].tid_off, c->cls[i].k) & nmask;
        while (ns[s]) s = (s + 1) & nmask;
        ns[s] = (uint64_t)i + 1;
    }
    free(c->slots); c->slots = ns; c->mask = nmask;
}
/* the EUMA row of the class of the sorted tid multiset t[0..k), created (all zero) when it is new */
static int32_t *cs_row(cstore *c, const int32_t *t, int k)
{
    if (k > c->max_k) c->max_k = k;
    uint64_t s = key_hash(t, k) & c->mask;
    for (;;) {
        uint64_t v = c->slots[s];
        if (!v) break;
        const bclass *b = &c->cls[v - 1];
        if (b->k == k && memcmp(c->tids + b->tid_off, t, sizeof(int32_t) * (size_t)k) == 0) return c->euma + b->euma_row * c->nF;
        s = (s + 1) & c->mask;
    }
    if (c->ncls == c->capcls) { c->capcls = c->capcls ? c->capcls * 2 : 4096; c->cls = (bclass *)realloc(c->cls, sizeof(bclass) * (size_t)c->capcls); }
    if (c->ntids + k > c->captids) { c->captids = c->captids ? c->captids * 2 : 1 << 16; while (c->ntids + k > c->captids) c->captids *= 2; c->tids = (int32_t *)realloc(c->tids, sizeof(int32_t) * (size_t)c->captids); }
    if (c->nrows == c->caprows) { c->caprows = c->caprows ? c->caprows * 2 : 4096; c->euma = (int32_t *)realloc(c->euma, sizeof(int32_t) * (size_t)c->caprows * c->nF); }
    bclass *b = &c->cls[c->ncls];
    b->k = k; b->tid_off = c->ntids; b->euma_row = c->nrows;
    memcpy(c->tids + c->ntids, t, sizeof(int32_t) * (size_t)k);
    memset(c->euma + c->nrows * c->nF, 0, sizeof(int32_t) * (size_t)c->nF);
    const int64_t row = c->nrows;
    c->ntids += k; c->nrows++;
    c->slots[s] = (uint64_t)(++c->ncls);
    if ((uint64_t)c->ncls * 2 > c->mask) cs_rehash(c);
    return c->euma + row * c->nF;
}
/* update_rshbucket(..., 'e', fi): the class gains one substring at fragment-length index fi */
static void cs_add(cstore *c, const int32_t *t, int k, int fi) { cs_row(c, t, k)[fi]++; }
static void cs_add_single(cstore *c, int32_t tid, int fi) { c->s_node[tid] = 1; c->s_euma[(size_t)tid * c->nF + fi]++; }

/* fold the store of a worker thread into dst (counts add up; the final order does not depend on who found a class first) */
static void cs_merge(cstore *dst, const cstore *src)
{
    for (int32_t t = 0; t < src->T; t++) {
        if (!src->s_node[t]) continue;
        dst->s_node[t] = 1;
        for (int i = 0; i < src->nF; i++) dst->s_euma[(size_t)t * dst->nF + i] += src->s_euma[(size_t)t * src->nF + i];
    }
    for (int64_t j = 0; j < src->ncls; j++) {
        const bclass *b = &src->cls[j];
        const int32_t *row = src->euma + b->euma_row * src->nF;
        int32_t *dst_row = cs_row(dst, src->tids + b->tid_off, b->k);
        for (int i = 0; i < src->nF; i++) dst_row[i] += row[i];
    }
}

/* ---- substrings: rolling hash + exact comparison ---------------------------------------------------------------------- */
static const uint64_t HB = 0x100000001B3ULL * 31 + 2;      /* odd multiplier of the polynomial hash (mod 2^64) */
static inline uint64_t code(char ch) { return ch == 'A' ? 1 : ch == 'C' ? 2 : ch == 'G' ? 3 : ch == 'T' ? 4 : 0; }

/* H[p] = hash of S[p, p+L) for every p whose substring is pure ACGT, ok[p] = 1 there (mark_noncanonical :2642-2660) */
static void hash_all(const txome *x, int L, uint64_t *H, uint8_t *ok)
{
    const int64_t n = x->end + 1;
    uint64_t pw = 1;
    for (int i = 0; i < L - 1; i++) pw *= HB;
    memset(ok, 0, (size_t)n);
    int64_t run = 0;          /* ACGT characters ending at i */
    uint64_t h = 0;
    for (int64_t i = 0; i < n; i++) {
        const uint64_t cd = code(x->S[i]);
        if (!cd) { run = 0; h = 0; continue; }
        if (run >= L) h -= code(x->S[i - L]) * pw;
        h = h * HB + cd;
        run++;
        if (run >= L) { H[i - L + 1] = h; ok[i - L + 1] = 1; }
    }
}

typedef struct { uint64_t h; int64_t pos; int32_t tid; int32_t d; } sub;
static const char *g_S;       /* qsort context (single-threaded build) */
static int g_L;
static int sub_cmp(const void *a, const void *b)
{
    const sub *x = (const sub *)a, *y = (const sub *)b;
    if (x->h != y->h) return x->h < y->h ? -1 : 1;
    return memcmp(g_S + x->pos, g_S + y->pos, (size_t)g_L);
}
static int i32_cmp(const void *a, const void *b) { int32_t x = *(const int32_t *)a, y = *(const int32_t *)b; return x < y ? -1 : x > y; }

/* a run e[0..r) of equal substrings -> the class store (construct_rshbucket_2 / construct_rshbucket_PE_3) */
static void emit_run(cstore *c, const sub *e, int64_t r, int max_repeat, int fi_base, int pe, int32_t **tbuf, int64_t *tcap)
{
    if (r == 1) { cs_add_single(c, e[0].tid, pe ? e[0].d + fi_base : fi_base); return; }
    if (pe) for (int64_t j = 1; j < r; j++) if (e[j].d != e[0].d) return;          /* multi_d (:1925-1928) */
    if (r >= max_repeat) return;
    if (r > *tcap) { *tcap = r * 2; *tbuf = (int32_t *)realloc(*tbuf, sizeof(int32_t) * (size_t)*tcap); }
    for (int64_t j = 0; j < r; j++) (*tbuf)[j] = e[j].tid;
    qsort(*tbuf, (size_t)r, sizeof(int32_t), i32_cmp);
    cs_add(c, *tbuf, (int)r, pe ? e[0].d + fi_base : fi_base);
}
/* g_S / g_L must be set by the caller (they are the same for every thread of a PE build) */
static void scan_runs(cstore *c, sub *e, int64_t n, int L, const char *S, int max_repeat, int fi_base, int pe, int32_t **tbuf, int64_t *tcap)
{
    qsort(e, (size_t)n, sizeof(sub), sub_cmp);
    for (int64_t a = 0; a < n;) {
        int64_t b = a + 1;
        while (b < n && e[b].h == e[a].h && memcmp(S + e[b].pos, S + e[a].pos, (size_t)L) == 0) b++;
        emit_run(c, e + a, b - a, max_repeat, fi_base, pe, tbuf, tcap);
        a = b;
    }
}

/* ---- the two library layouts -------------------------------------------------------------------------------------------- */
typedef struct { sub *part; const int64_t *cnt; int nb; int *next; pthread_mutex_t *mu; int L; const char *S; int max_repeat, fi; cstore *c; } se_job;
static void *se_worker(void *arg)
{
    se_job *jb = (se_job *)arg;
    int32_t *tbuf = NULL; int64_t tcap = 0;
    for (;;) {
        pthread_mutex_lock(jb->mu);
        const int b = (*jb->next)++;
        pthread_mutex_unlock(jb->mu);
        if (b >= jb->nb) break;
        const int64_t a0 = jb->cnt[b], a1 = jb->cnt[b + 1];
        if (a1 > a0) scan_runs(jb->c, jb->part + a0, a1 - a0, jb->L, jb->S, jb->max_repeat, jb->fi, 0, &tbuf, &tcap);
    }
    free(tbuf);
    return NULL;
}

static void build_se(const txome *x, const emsar_build_opts *o, cstore *c)
{
    const int64_t n = x->end + 1;
    uint64_t *H = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)n);
    uint8_t *ok = (uint8_t *)malloc((size_t)n);
    sub *e = NULL; int64_t cap = 0;
    int32_t *tbuf = NULL; int64_t tcap = 0;
    for (int L = o->readlen_min; L <= o->readlen_max; L++) {
        hash_all(x, L, H, ok);
        g_S = x->S; g_L = L;
        int64_t m = 0;
        for (int32_t t = 0; t < x->T; t++) {
            const int64_t s0 = x->start[t], len = x->start[t + 1] - 1 - s0;          /* start[t+1] - 1 is the '@' / '$' */
            for (int64_t i = s0; i + L <= s0 + len; i++) {
                if (!ok[i]) continue;
                int64_t p = i;
                if (!o->stranded) {                                                  /* initialize_suffixarray_NS_5 :1001-1027 */
                    const int64_t fl = x->end - i - L;
                    if (memcmp(x->S + i, x->S + fl, (size_t)L) > 0) p = fl;
                }
                if (m == cap) { cap = cap ? cap * 2 : 1 << 20; e = (sub *)realloc(e, sizeof(sub) * (size_t)cap); }
                e[m].h = H[p]; e[m].pos = p; e[m].tid = t; e[m].d = 0;
                m++;
            }
        }
        const int nthr = (o->threads > 1 && m >= 200000) ? (o->threads > 64 ? 64 : o->threads) : 1;
        if (nthr == 1) { scan_runs(c, e, m, L, x->S, o->max_repeat, L - o->readlen_min, 0, &tbuf, &tcap); continue; }
        /* equal substrings share their hash: partition the occurrences by its top bits and let every thread sort and scan whole
         * partitions into a class store of its own (merged afterwards; the result does not depend on the thread count) */
        enum { NB = 256 };
        int64_t cnt[NB + 1];
        memset(cnt, 0, sizeof cnt);
        for (int64_t i = 0; i < m; i++) cnt[(e[i].h >> 56) + 1]++;
        for (int b = 0; b < NB; b++) cnt[b + 1] += cnt[b];
        sub *part = (sub *)malloc(sizeof(sub) * (size_t)m);
        int64_t fill[NB];
        memcpy(fill, cnt, sizeof fill);
        for (int64_t i = 0; i < m; i++) part[fill[e[i].h >> 56]++] = e[i];
        se_job jobs[64];
        pthread_t th[64];
        int next_bucket = 0;
        pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
        for (int w = 0; w < nthr; w++) {
            jobs[w].part = part; jobs[w].cnt = cnt; jobs[w].nb = NB; jobs[w].next = &next_bucket; jobs[w].mu = &mu;
            jobs[w].L = L; jobs[w].S = x->S; jobs[w].max_repeat = o->max_repeat; jobs[w].fi = L - o->readlen_min;
            jobs[w].c = w == 0 ? c : (cstore *)malloc(sizeof(cstore));
            if (w > 0) cs_init(jobs[w].c, c->T, c->nF);
        }
        for (int w = 1; w < nthr; w++) pthread_create(&th[w], NULL, se_worker, &jobs[w]);
        se_worker(&jobs[0]);
        for (int w = 1; w < nthr; w++) { pthread_join(th[w], NULL); cs_merge(c, jobs[w].c); cs_free(jobs[w].c); free(jobs[w].c); }
        free(part);
    }
    free(H); free(ok); free(e); free(tbuf);
}

static int cmp_pe(const char *a, const char *b, int d, int L)       /* strcmp_pe :2674-2678 */
{
    int r = memcmp(a, b, (size_t)L);
    return r ? r : memcmp(a + d, b + d, (size_t)L);
}

typedef struct { const txome *x; const emsar_build_opts *o; const sub *m1; int64_t a, b; const uint8_t *ok; const uint64_t *H; int fmin, fmax; cstore *c; } pe_job;

/* the clusters inside m1[a..b): every member with every admissible mate 2 (process_mate1_cluster_by_mate_3 :2852-2874) */
static void *pe_worker(void *arg)
{
    pe_job *jb = (pe_job *)arg;
    const txome *x = jb->x;
    const emsar_build_opts *o = jb->o;
    const sub *m1 = jb->m1;
    const int L = o->readlength, dmin = jb->fmin - L, dmax = jb->fmax - L;
    sub *e = NULL; int64_t cap = 0;
    int32_t *tbuf = NULL; int64_t tcap = 0;
    for (int64_t a = jb->a; a < jb->b;) {
        int64_t b = a + 1;
        while (b < jb->b && m1[b].h == m1[a].h && memcmp(x->S + m1[b].pos, x->S + m1[a].pos, (size_t)L) == 0) b++;
        int64_t m = 0;
        for (int64_t j = a; j < b; j++) {
            const int64_t p = m1[j].pos;
            const int32_t t = m1[j].tid;
            int64_t lo, hi;                                  /* mate 2 must start in [lo, hi]: the member's transcript, in the half it lies in */
            if (p < x->border) { lo = x->start[t]; hi = x->start[t + 1] - 1 - L; }
            else { lo = x->end - (x->start[t + 1] - 1); hi = x->end - x->start[t] - L; }
            for (int d = dmin; d <= dmax; d++) {
                const int64_t q = p + d;
                if (q < lo || q > hi || !jb->ok[q]) continue;
                if (!o->stranded) {
                    const int cr = cmp_pe(x->S + p, x->S + (x->end - q - L), d, L);
                    if (!((p < x->border && cr <= 0) || (p > x->border && cr < 0))) continue;
                }
                if (m == cap) { cap = cap ? cap * 2 : 1 << 16; e = (sub *)realloc(e, sizeof(sub) * (size_t)cap); }
                e[m].h = jb->H[q]; e[m].pos = q; e[m].tid = t; e[m].d = d;
                m++;
            }
        }
        if (m > 0) scan_runs(jb->c, e, m, L, x->S, o->max_repeat, L - jb->fmin, 1, &tbuf, &tcap);
        a = b;
    }
    free(e); free(tbuf);
    return NULL;
}

static void build_pe(const txome *x, const emsar_build_opts *o, cstore *c, int fmin, int fmax)
{
    const int L = o->readlength;
    const int64_t n = x->end + 1;
    uint64_t *H = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)n);
    uint8_t *ok = (uint8_t *)malloc((size_t)n);
    hash_all(x, L, H, ok);
    /* mate-1 occurrences: forward substrings, and their reverse complements when the library is unstranded (:1054-1103) */
    sub *m1 = NULL; int64_t nm1 = 0, cap1 = 0;
    for (int32_t t = 0; t < x->T; t++) {
        const int64_t s0 = x->start[t], len = x->start[t + 1] - 1 - s0;
        for (int64_t i = s0; i + L <= s0 + len; i++) {
            if (!ok[i]) continue;
            const int reps = o->stranded ? 1 : 2;
            for (int k = 0; k < reps; k++) {
                const int64_t p = k == 0 ? i : x->end - i - L;
                if (nm1 == cap1) { cap1 = cap1 ? cap1 * 2 : 1 << 20; m1 = (sub *)realloc(m1, sizeof(sub) * (size_t)cap1); }
                m1[nm1].h = H[p]; m1[nm1].pos = p; m1[nm1].tid = t; m1[nm1].d = 0;
                nm1++;
            }
        }
    }
    g_S = x->S; g_L = L;
    qsort(m1, (size_t)nm1, sizeof(sub), sub_cmp);
    /* clusters of equal mate 1 are independent: cut the sorted list at cluster boundaries into one piece per thread */
    int nthr = o->threads > 1 ? (o->threads > 64 ? 64 : o->threads) : 1;
    if (nm1 < 2000) nthr = 1;
    pe_job jobs[64];
    pthread_t th[64];
    int64_t cut = 0;
    for (int w = 0; w < nthr; w++) {
        int64_t nxt = w == nthr - 1 ? nm1 : nm1 * (w + 1) / nthr;
        while (nxt > cut && nxt < nm1 && m1[nxt].h == m1[nxt - 1].h && memcmp(x->S + m1[nxt].pos, x->S + m1[nxt - 1].pos, (size_t)L) == 0) nxt++;
        jobs[w].x = x; jobs[w].o = o; jobs[w].m1 = m1; jobs[w].a = cut; jobs[w].b = nxt; jobs[w].ok = ok; jobs[w].H = H; jobs[w].fmin = fmin; jobs[w].fmax = fmax;
        jobs[w].c = w == 0 ? c : (cstore *)malloc(sizeof(cstore));
        if (w > 0) cs_init(jobs[w].c, c->T, c->nF);
        cut = nxt;
    }
    for (int w = 1; w < nthr; w++) pthread_create(&th[w], NULL, pe_worker, &jobs[w]);
    pe_worker(&jobs[0]);
    for (int w = 1; w < nthr; w++) {
        pthread_join(th[w], NULL);
        cs_merge(c, jobs[w].c);
        cs_free(jobs[w].c);
        free(jobs[w].c);
    }
    free(H); free(ok); free(m1);
}

/* ---- construction on the device (include/emsar_cuda.h: emsar_build_classes_run) ---------------------------------------------- */
static double g_device_ms;      /* diagnostics of the last device construction (EMSAR_BUILD_TIMING) */
static long long g_device_occ, g_device_runs;
static int g_device_parts;
static int build_device(const txome *x, const emsar_build_opts *o, cstore *c, int fmin, int fmax, char *err)
{
    g_device_ms = 0; g_device_occ = 0; g_device_runs = 0; g_device_parts = 0;
    const int L0 = o->pe ? o->readlength : o->readlen_min, L1 = o->pe ? o->readlength : o->readlen_max;
    for (int L = L0; L <= L1; L++) {
        emsar_build_desc d;
        memset(&d, 0, sizeof d);
        d.seq = x->S; d.border = x->border; d.end = x->end; d.T = x->T; d.start = x->start;
        d.pe = o->pe; d.stranded = o->stranded; d.readlen = L; d.max_repeat = o->max_repeat;
        if (o->pe) { d.d_min = fmin - L; d.d_max = fmax - L; }
        emsar_build_classes r;
        memset(&r, 0, sizeof r);
        if (o->device_run(o->device_ctx, &d, &r)) return fail(err, "index construction on the device: %s", o->device_error ? o->device_error() : "failed");
        const int fi0 = o->pe ? 0 : L - fmin;               /* PE: fragment length index = distance index (d_min = fmin - L) */
        for (int32_t t = 0; t < r.T; t++)
            for (int j = 0; j < r.n_d; j++) {
                const int32_t n = r.single_count[(size_t)t * r.n_d + j];
                if (n) { c->s_node[t] = 1; c->s_euma[(size_t)t * c->nF + fi0 + j] += n; }
            }
        for (int64_t u = 0; u < r.n_class; u++)
            cs_row(c, r.class_tid + r.class_off[u], (int)(r.class_off[u + 1] - r.class_off[u]))[fi0 + r.class_d[u]] += r.class_count[u];
        g_device_ms += r.device_ms; g_device_occ += r.occurrences; g_device_runs += r.runs; g_device_parts += r.partitions;
        if (o->device_free) o->device_free(&r);
    }
    return 0;
}

/* ---- class store -> emsar_rsh in the reference's scan / print order ------------------------------------------------------ */
static const cstore *g_cs;
static int cls_cmp(const void *a, const void *b)
{
    const bclass *x = &g_cs->cls[*(const int64_t *)a], *y = &g_cs->cls[*(const int64_t *)b];
    if (x->k != y->k) return x->k < y->k ? -1 : 1;
    const int32_t *tx = g_cs->tids + x->tid_off, *ty = g_cs->tids + y->tid_off;
    for (int i = 0; i < x->k; i++) if (tx[i] != ty[i]) return tx[i] < ty[i] ? -1 : 1;
    return 0;
}

void emsar_rsh_name_index(emsar_rsh *r);

int emsar_rsh_build(const char *fasta_path, const emsar_build_opts *o, emsar_rsh **out, char *err)
{
    if (!fasta_path || !o || !out) return fail(err, "emsar_rsh_build: NULL argument");
    if (o->max_repeat < 1) return fail(err, "emsar_rsh_build: max_repeat must be positive");
    int fmin, fmax, nF;
    if (o->pe) {
        if (o->readlength < 1) return fail(err, "emsar_rsh_build: paired-end needs the read length");
        if (o->min_fraglength > o->max_fraglength || o->min_fraglength < 1 || o->max_fraglength < 1) return fail(err, "error: invalid fragment length range.");
        fmin = o->min_fraglength > o->readlength ? o->min_fraglength : o->readlength;          /* determine_fraglength_range :2471-2475 */
        fmax = o->max_fraglength >= fmin ? o->max_fraglength : fmin;
    } else {
        if (o->readlen_min < 1 || o->readlen_max < o->readlen_min) return fail(err, "emsar_rsh_build: bad read length range");
        fmin = o->readlen_min; fmax = o->readlen_max;
    }
    nF = fmax - fmin + 1;
    /* EMSAR_BUILD_TIMING=1: seconds per stage on stderr (profiles/build_bench.py) */
    const int timing = getenv("EMSAR_BUILD_TIMING") && atoi(getenv("EMSAR_BUILD_TIMING"));
    struct timespec ts0, ts1, ts2, ts3;
    clock_gettime(CLOCK_MONOTONIC, &ts0);
    txome x;
    if (txome_read(fasta_path, o->header == 'R' ? 'R' : 'E', &x, err)) return 1;
    clock_gettime(CLOCK_MONOTONIC, &ts1);
    cstore c;
    cs_init(&c, x.T, nF);
    if (o->device_run) {
        if (build_device(&x, o, &c, fmin, fmax, err)) { cs_free(&c); txome_free(&x); return 1; }
    } else if (o->pe) build_pe(&x, o, &c, fmin, fmax);
    else build_se(&x, o, &c);
    clock_gettime(CLOCK_MONOTONIC, &ts2);
    /* flatten: singletons in tid order, then by (cardinality, tids) - the order print_rsh walks the buckets and their sorted chains */
    int64_t *ord = (int64_t *)malloc(sizeof(int64_t) * (size_t)(c.ncls ? c.ncls : 1));
    for (int64_t i = 0; i < c.ncls; i++) ord[i] = i;
    g_cs = &c;
    qsort(ord, (size_t)c.ncls, sizeof(int64_t), cls_cmp);
    emsar_rsh *r = (emsar_rsh *)calloc(1, sizeof(emsar_rsh));
    r->T = x.T; r->C = (int64_t)x.T + c.ncls; r->nF = nF;
    r->min_fraglength = fmin; r->max_fraglength = fmax; r->frag_min = fmin; r->frag_max = fmax;
    r->readlength = o->pe ? o->readlength : -1;
    r->max_t_size = c.max_k;
    r->class_ptr = (int64_t *)malloc(sizeof(int64_t) * (size_t)(r->C + 1));
    r->class_tid = (int32_t *)malloc(sizeof(int32_t) * (size_t)(x.T + c.ntids + 1));
    r->euma = (int32_t *)malloc(sizeof(int32_t) * (size_t)r->C * nF);
    r->has_node = (uint8_t *)malloc((size_t)r->C);
    for (int32_t t = 0; t < x.T; t++) { r->class_ptr[t] = t; r->class_tid[t] = t; r->has_node[t] = c.s_node[t]; }
    memcpy(r->euma, c.s_euma, sizeof(int32_t) * (size_t)x.T * nF);
    int64_t off = x.T;
    for (int64_t i = 0; i < c.ncls; i++) {
        const bclass *b = &c.cls[ord[i]];
        const int64_t cid = x.T + i;
        r->class_ptr[cid] = off;
        memcpy(r->class_tid + off, c.tids + b->tid_off, sizeof(int32_t) * (size_t)b->k);
        memcpy(r->euma + (size_t)cid * nF, c.euma + b->euma_row * nF, sizeof(int32_t) * (size_t)nF);
        r->has_node[cid] = 1;
        off += b->k;
    }
    r->class_ptr[r->C] = off;
    r->names = x.names; x.names = NULL;              /* ownership moves to the index */
    emsar_rsh_name_index(r);
    free(ord);
    cs_free(&c);
    txome_free(&x);
    *out = r;
    clock_gettime(CLOCK_MONOTONIC, &ts3);
    if (timing) {
        #define SECS(a, b) ((double)((b).tv_sec - (a).tv_sec) + 1e-9 * (double)((b).tv_nsec - (a).tv_nsec))
        fprintf(stderr, "build timing: fasta %.3f s, classes (%s) %.3f s, order %.3f s\n", SECS(ts0, ts1), o->device_run ? "device" : "host", SECS(ts1, ts2), SECS(ts2, ts3));
        if (o->device_run) fprintf(stderr, "build timing: device stream %.1f ms for %lld occurrences, %lld distinct sequences, %d partition(s)\n", g_device_ms, g_device_occ, g_device_runs, g_device_parts);
        #undef SECS
    }
    return 0;
}
