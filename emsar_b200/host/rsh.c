/* `.rsh` text index: loader (reference construct_rsh_from_rshfile, emsar_functions.c:1351-1510), writer (print_rsh
 * :2071-2130) and the tname -> tid map (the reference uses a character trie, stringhash.c; any map will do).
 * The loader flattens the class store directly into the reference's scan order (scan_rshbucket :2149-2191):
 * all singletons in tid order, then multi-tid classes by cardinality, first tid and chain (file) order. */
#define _GNU_SOURCE
#include <errno.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include "emsar_host.h"

static int fail(char *err, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    if (err) vsnprintf(err, EMSAR_HOST_ERRLEN, fmt, ap);
    va_end(ap);
    return 1;
}

static uint32_t fnv1a(const char *s)
{
    uint32_t h = 2166136261u;
    for (; *s; s++) { h ^= (unsigned char)*s; h *= 16777619u; }
    return h;
}

int emsar_rsh_tid(const emsar_rsh *r, const char *name)
{
    uint32_t s = fnv1a(name) & r->name_mask;
    for (;;) {
        uint32_t v = r->name_slots[s];
        if (v == 0) return -1;
        if (r->names[v - 1] && strcmp(r->names[v - 1], name) == 0) return (int)(v - 1);
        s = (s + 1) & r->name_mask;
    }
}

static void name_insert(emsar_rsh *r, int tid)
{
    /* insert_key (stringhash.c): a later identical name overwrites the value, so lookups return the LAST tid */
    uint32_t s = fnv1a(r->names[tid]) & r->name_mask;
    for (;;) {
        uint32_t v = r->name_slots[s];
        if (v == 0) { r->name_slots[s] = (uint32_t)tid + 1; return; }
        if (strcmp(r->names[v - 1], r->names[tid]) == 0) { r->name_slots[s] = (uint32_t)tid + 1; return; }
        s = (s + 1) & r->name_mask;
    }
}

/* (re)build the tname -> tid map of an index whose names[] are filled in */
void emsar_rsh_name_index(emsar_rsh *r)
{
    free(r->name_slots);
    uint32_t slots = 16;
    while (slots < (uint32_t)r->T * 2u) slots <<= 1;
    r->name_mask = slots - 1;
    r->name_slots = (uint32_t *)calloc(slots, sizeof(uint32_t));
    for (int32_t t = 0; t < r->T; t++) name_insert(r, t);
}

void emsar_rsh_free(emsar_rsh *r)
{
    if (!r) return;
    if (r->names) { for (int32_t t = 0; t < r->T; t++) free(r->names[t]); free(r->names); }
    free(r->class_ptr); free(r->class_tid); free(r->euma); free(r->has_node); free(r->name_slots);
    if (r->aux_owned) { free(r->aux_txm_off); free(r->aux_txm_cid); free(r->aux_order); free(r->aux_insertable); }
    free(r);
}

/* split `line` in place at `sep`; returns the number of fields (pointers into line) */
static int split(char *line, char sep, char **f, int maxf)
{
    int n = 0;
    char *p = line;
    f[n++] = p;
    for (; *p; p++)
        if (*p == sep) { *p = 0; if (n < maxf) f[n++] = p + 1; else return n; }
    return n;
}

typedef struct { int32_t k, tid0; int64_t seq, tid_off, euma_off; } mrec;
static int mrec_cmp(const void *a, const void *b)
{
    const mrec *x = (const mrec *)a, *y = (const mrec *)b;
    if (x->k != y->k) return x->k < y->k ? -1 : 1;
    if (x->tid0 != y->tid0) return x->tid0 < y->tid0 ? -1 : 1;
    return x->seq < y->seq ? -1 : (x->seq > y->seq ? 1 : 0);
}

/* values are terminated by ',' like the reference's loops (:1459-1467, :1475-1483): a last value without a
 * trailing comma is never stored */
static int parse_commas(const char *s, int32_t *out, int maxn)
{
    int n = 0;
    const char *start = s;
    for (const char *p = s; *p; p++)
        if (*p == ',') { if (n < maxn) out[n] = atoi(start); n++; start = p + 1; }
    return n;
}

int emsar_rsh_load(const char *path, emsar_rsh **out, char *err)
{
    FILE *f = fopen(path, "r");
    if (!f) return fail(err, "can't open input rsh file.");
    emsar_rsh *r = (emsar_rsh *)calloc(1, sizeof(emsar_rsh));
    char *line = NULL;
    size_t cap = 0;
    ssize_t len;
    int have_header = 0;
    mrec *m = NULL; int64_t nm = 0, capm = 0;
    int32_t *tpool = NULL; int64_t ntp = 0, captp = 0;
    int32_t *epool = NULL; int64_t nep = 0, capep = 0;
    int32_t *s_euma = NULL; /* singleton EUMA rows [T*nF] */
    int rc = 0;
    int64_t lineno = 0;
    while ((len = getline(&line, &cap, f)) >= 0) {
        lineno++;
        if (len > 0 && line[len - 1] == '\n') line[--len] = 0;
        if (line[0] == '#') {                                   /* parse_rsh_headerline :1406-1430 */
            char *fld[8];
            int n = split(line, ',', fld, 8);
            if (n < 5) { rc = fail(err, "rsh header line has %d fields, need 5", n); break; }
            r->T = atoi(fld[0] + 1) + 1;
            r->max_t_size = atoi(fld[1]);
            r->min_fraglength = atoi(fld[2]);
            r->max_fraglength = atoi(fld[3]);
            r->readlength = atoi(fld[4]);
            if (r->T <= 0) { rc = fail(err, "rsh header: bad max_tid"); break; }
            r->frag_min = r->min_fraglength > r->readlength ? r->min_fraglength : r->readlength;
            r->frag_max = r->max_fraglength >= r->frag_min ? r->max_fraglength : r->frag_min;
            r->nF = r->frag_max - r->frag_min + 1;
            r->names = (char **)calloc((size_t)r->T, sizeof(char *));
            r->has_node = NULL;
            s_euma = (int32_t *)calloc((size_t)r->T * r->nF, sizeof(int32_t));
            r->has_node = (uint8_t *)calloc((size_t)r->T, 1);
            have_header = 1;
        } else if (line[0] == '@') {                            /* parse_rsh_indexline :1381-1403 */
            if (!have_header) { rc = fail(err, "rsh: index line before the header"); break; }
            char *fld[3];
            int n = split(line, '\t', fld, 3);
            int tid = atoi(fld[0] + 1);
            if (n < 2 || tid < 0 || tid >= r->T) { rc = fail(err, "rsh line %lld: bad index line", (long long)lineno); break; }
            free(r->names[tid]);
            r->names[tid] = strdup(fld[1]);
        } else if (line[0] != 'c') {                            /* parse_rsh_mainline :1432-1510 */
            if (!have_header) { rc = fail(err, "rsh: class line before the header"); break; }
            if (line[0] == 0) continue;
            char *fld[6];
            int n = split(line, '\t', fld, 6);
            if (n < 3) { rc = fail(err, "rsh line %lld: too few fields", (long long)lineno); break; }
            int k = atoi(fld[1]), tid0 = atoi(fld[2]);
            const char *others = n > 3 ? fld[3] : "", *eu = n > 4 ? fld[4] : "";
            if (k < 1 || tid0 < 0 || tid0 >= r->T) { rc = fail(err, "rsh line %lld: bad cardinality or tid", (long long)lineno); break; }
            if (strlen(eu) == 0) continue;                      /* no EUMA: no node is created (:1486) */
            if (k == 1) {
                int32_t *row = s_euma + (size_t)tid0 * r->nF;
                memset(row, 0, sizeof(int32_t) * r->nF);
                parse_commas(eu, row, r->nF);
                r->has_node[tid0] = 1;                          /* a later line replaces an earlier one (:1488) */
                continue;
            }
            if (k > r->max_t_size) { rc = fail(err, "rsh line %lld: %d tids exceed header max_t_size %d", (long long)lineno, k, r->max_t_size); break; }
            if (nm == capm) { capm = capm ? capm * 2 : 1 << 16; m = (mrec *)realloc(m, sizeof(mrec) * capm); }
            if (ntp + k > captp) { captp = captp ? captp * 2 : 1 << 18; while (ntp + k > captp) captp *= 2; tpool = (int32_t *)realloc(tpool, sizeof(int32_t) * captp); }
            if (nep + r->nF > capep) { capep = capep ? capep * 2 : 1 << 18; while (nep + r->nF > capep) capep *= 2; epool = (int32_t *)realloc(epool, sizeof(int32_t) * capep); }
            tpool[ntp] = tid0;
            int got = parse_commas(others, tpool + ntp + 1, k - 1);
            if (got != k - 1) { rc = fail(err, "rsh line %lld: %d other tids, expected %d", (long long)lineno, got, k - 1); break; }
            for (int j = 0; j < k; j++) {
                if (tpool[ntp + j] < 0 || tpool[ntp + j] >= r->T) { rc = fail(err, "rsh line %lld: tid out of range", (long long)lineno); break; }
                if (j && tpool[ntp + j] < tpool[ntp + j - 1]) { rc = fail(err, "rsh line %lld: tids not sorted", (long long)lineno); break; }
            }
            if (rc) break;
            memset(epool + nep, 0, sizeof(int32_t) * r->nF);
            parse_commas(eu, epool + nep, r->nF);
            m[nm].k = k; m[nm].tid0 = tid0; m[nm].seq = nm; m[nm].tid_off = ntp; m[nm].euma_off = nep;
            nm++; ntp += k; nep += r->nF;
        }
    }
    free(line);
    fclose(f);
    if (!rc && !have_header) rc = fail(err, "rsh file has no header line");
    if (!rc) {
        /* chains are appended behind the LAST node read (lastp, :1495-1502): a (cardinality, first tid) chain that is
         * revisited after another chain was started cannot be represented by the reference; reject such files */
        qsort(m, (size_t)nm, sizeof(mrec), mrec_cmp);
        for (int64_t i = 1; i < nm && !rc; i++)
            if (m[i].k == m[i - 1].k && m[i].tid0 == m[i - 1].tid0 && m[i].seq != m[i - 1].seq + 1)
                rc = fail(err, "rsh: classes with %d tids starting at tid %d are not contiguous in the file", m[i].k, m[i].tid0);
    }
    if (!rc) {
        for (int32_t t = 0; t < r->T; t++)
            if (!r->names[t]) { char b[32]; snprintf(b, sizeof b, "tid%d", t); r->names[t] = strdup(b); }
        r->C = (int64_t)r->T + nm;
        r->class_ptr = (int64_t *)malloc(sizeof(int64_t) * (size_t)(r->C + 1));
        r->class_tid = (int32_t *)malloc(sizeof(int32_t) * (size_t)(r->T + ntp > 0 ? r->T + ntp : 1));
        r->euma = (int32_t *)malloc(sizeof(int32_t) * (size_t)r->C * r->nF);
        r->has_node = (uint8_t *)realloc(r->has_node, (size_t)r->C);
        for (int32_t t = 0; t < r->T; t++) { r->class_ptr[t] = t; r->class_tid[t] = t; }
        memcpy(r->euma, s_euma, sizeof(int32_t) * (size_t)r->T * r->nF);
        int64_t o = r->T;
        for (int64_t i = 0; i < nm; i++) {
            int64_t c = r->T + i;
            r->class_ptr[c] = o;
            memcpy(r->class_tid + o, tpool + m[i].tid_off, sizeof(int32_t) * m[i].k);
            memcpy(r->euma + (size_t)c * r->nF, epool + m[i].euma_off, sizeof(int32_t) * r->nF);
            r->has_node[c] = 1;
            o += m[i].k;
        }
        r->class_ptr[r->C] = o;
        uint32_t slots = 16;
        while (slots < (uint32_t)r->T * 2u) slots <<= 1;
        r->name_mask = slots - 1;
        r->name_slots = (uint32_t *)calloc(slots, sizeof(uint32_t));
        for (int32_t t = 0; t < r->T; t++) name_insert(r, t);
    }
    free(m); free(tpool); free(epool); free(s_euma);
    if (rc) { emsar_rsh_free(r); return rc; }
    *out = r;
    return 0;
}

/* print_rsh writes hundreds of millions of small integers for a paired-end index (C x nF counts): they go through a
 * hand-rolled decimal formatter into a 1 MB buffer instead of one fprintf each */
typedef struct { FILE *f; char *buf; size_t n, cap; int failed; } wbuf;
static void wb_flush(wbuf *w) { if (w->n && fwrite(w->buf, 1, w->n, w->f) != w->n) w->failed = 1; w->n = 0; }
static inline void wb_room(wbuf *w, size_t need) { if (w->n + need > w->cap) wb_flush(w); }
static inline void wb_ch(wbuf *w, char c) { wb_room(w, 1); w->buf[w->n++] = c; }
static void wb_str(wbuf *w, const char *s)
{
    size_t l = strlen(s);
    if (l >= w->cap) { wb_flush(w); if (fwrite(s, 1, l, w->f) != l) w->failed = 1; return; }
    wb_room(w, l);
    memcpy(w->buf + w->n, s, l);
    w->n += l;
}
static void wb_int(wbuf *w, long long v)
{
    char t[24];
    int k = 0;
    unsigned long long u = v < 0 ? 0ULL - (unsigned long long)v : (unsigned long long)v;
    do { t[k++] = (char)('0' + u % 10); u /= 10; } while (u);
    wb_room(w, (size_t)k + 1);
    if (v < 0) w->buf[w->n++] = '-';
    while (k) w->buf[w->n++] = t[--k];
}

int emsar_rsh_write(const emsar_rsh *r, int pe, const char *path, char *err)
{
    FILE *f = fopen(path, "w");
    if (!f) return fail(err, "Can't write to output rsh file %s", path);
    wbuf w = {f, (char *)malloc(1 << 20), 0, 1 << 20, 0};
    wb_ch(&w, '#'); wb_int(&w, r->T - 1); wb_ch(&w, ','); wb_int(&w, r->max_t_size); wb_ch(&w, ','); wb_int(&w, r->frag_min); wb_ch(&w, ',');
    wb_int(&w, r->frag_max); wb_ch(&w, ','); wb_int(&w, pe ? r->readlength : -1); wb_ch(&w, '\n');
    for (int32_t t = 0; t < r->T; t++) { wb_ch(&w, '@'); wb_int(&w, t); wb_ch(&w, '\t'); wb_str(&w, r->names[t]); wb_ch(&w, '\n'); }
    wb_str(&w, "cid\tno.tids\tfirst.tid\tother.tids\tsegment.length\n");
    for (int64_t c = 0; c < r->C; c++) {
        int64_t o = r->class_ptr[c];
        int k = (int)(r->class_ptr[c + 1] - o);
        wb_int(&w, c); wb_ch(&w, '\t'); wb_int(&w, k); wb_ch(&w, '\t'); wb_int(&w, r->class_tid[o]); wb_ch(&w, '\t');
        if (k == 1 && !r->has_node[c]) { wb_str(&w, "\t\t\n"); continue; }
        for (int j = 1; j < k; j++) { wb_int(&w, r->class_tid[o + j]); wb_ch(&w, ','); }
        wb_ch(&w, '\t');
        const int32_t *e = r->euma + (size_t)c * r->nF;
        for (int i = 0; i < r->nF; i++) { wb_int(&w, e[i]); wb_ch(&w, ','); }
        wb_ch(&w, '\n');
    }
    wb_flush(&w);
    free(w.buf);
    if (fclose(f) != 0 || w.failed) return fail(err, "Can't write to output rsh file %s", path);
    return 0;
}

/* ---- packed binary image of a loaded index ------------------------------------------------------
 * The text `.rsh` stays the interchange format (emsar-build writes it, -R writes it); parsing it costs a getline + atoi
 * pass over every class line each run. The packed image holds the arrays exactly as emsar_rsh keeps them (scan order
 * already applied), so loading it is a handful of large reads. It records size and mtime of the text file it was made
 * from and is ignored when they no longer match. */
typedef struct {
    char magic[8];                      /* "EMSARPK2" */
    int32_t T, nF, min_fraglength, max_fraglength, readlength, max_t_size, frag_min, frag_max;
    int64_t C, nnz, names_bytes, src_size, src_mtime_ns;
    int64_t aux_nnz_multi;      /* EMSARPK2: > 0 (or aux_present) = the derived arrays follow the names */
    int32_t aux_present, aux_n_sets_nocut, aux_max_set_tids, reserved;
} pack_header;

static void src_stamp(const char *src, int64_t *size, int64_t *mtime_ns)
{
    struct stat st;
    *size = -1; *mtime_ns = -1;
    if (src && stat(src, &st) == 0) { *size = (int64_t)st.st_size; *mtime_ns = (int64_t)st.st_mtim.tv_sec * 1000000000LL + st.st_mtim.tv_nsec; }
}

int emsar_rsh_save_packed(const emsar_rsh *r, const char *path, const char *src_path, char *err)
{
    char tmp[4096];
    snprintf(tmp, sizeof tmp, "%s.tmp%d", path, (int)getpid());
    FILE *f = fopen(tmp, "wb");
    if (!f) return fail(err, "can't write packed rsh image %s", tmp);
    pack_header h;
    memset(&h, 0, sizeof h);
    memcpy(h.magic, "EMSARPK2", 8);
    if (r->has_aux) { h.aux_present = 1; h.aux_nnz_multi = r->aux_nnz_multi; h.aux_n_sets_nocut = r->aux_n_sets_nocut; h.aux_max_set_tids = r->aux_max_set_tids; }
    h.T = r->T; h.nF = r->nF; h.min_fraglength = r->min_fraglength; h.max_fraglength = r->max_fraglength; h.readlength = r->readlength;
    h.max_t_size = r->max_t_size; h.frag_min = r->frag_min; h.frag_max = r->frag_max;
    h.C = r->C; h.nnz = r->class_ptr[r->C];
    for (int32_t t = 0; t < r->T; t++) h.names_bytes += (int64_t)strlen(r->names[t]) + 1;
    src_stamp(src_path, &h.src_size, &h.src_mtime_ns);
    int ok = fwrite(&h, sizeof h, 1, f) == 1;
    ok = ok && fwrite(r->class_ptr, sizeof(int64_t), (size_t)r->C + 1, f) == (size_t)r->C + 1;
    ok = ok && (h.nnz == 0 || fwrite(r->class_tid, sizeof(int32_t), (size_t)h.nnz, f) == (size_t)h.nnz);
    ok = ok && fwrite(r->euma, sizeof(int32_t), (size_t)r->C * r->nF, f) == (size_t)r->C * r->nF;
    ok = ok && fwrite(r->has_node, 1, (size_t)r->C, f) == (size_t)r->C;
    for (int32_t t = 0; ok && t < r->T; t++) ok = fwrite(r->names[t], 1, strlen(r->names[t]) + 1, f) == strlen(r->names[t]) + 1;
    if (ok && r->has_aux) {       /* transpose, locality order, reachable classes: what emsar_index_create would otherwise derive */
        ok = fwrite(r->aux_txm_off, sizeof(uint32_t), (size_t)r->T + 1, f) == (size_t)r->T + 1;
        ok = ok && (r->aux_nnz_multi == 0 || fwrite(r->aux_txm_cid, sizeof(int32_t), (size_t)r->aux_nnz_multi, f) == (size_t)r->aux_nnz_multi);
        ok = ok && fwrite(r->aux_order, sizeof(int32_t), (size_t)r->T, f) == (size_t)r->T;
        ok = ok && (r->C == r->T || fwrite(r->aux_insertable, 1, (size_t)(r->C - r->T), f) == (size_t)(r->C - r->T));
    }
    ok = (fclose(f) == 0) && ok;
    if (!ok || rename(tmp, path) != 0) { remove(tmp); return fail(err, "can't write packed rsh image %s", path); }
    return 0;
}

/* src_path != NULL: the image must have been made from exactly that file (size + mtime), else -> 2 ("stale"), no error text */
int emsar_rsh_load_packed(const char *path, const char *src_path, emsar_rsh **out, char *err)
{
    FILE *f = fopen(path, "rb");
    if (!f) return fail(err, "can't open packed rsh image %s", path);
    pack_header h;
    if (fread(&h, sizeof h, 1, f) != 1 || memcmp(h.magic, "EMSARPK2", 8) != 0 || h.T <= 0 || h.C < h.T || h.nF <= 0 || h.nnz < h.C || h.names_bytes < h.T) {
        fclose(f);
        return fail(err, "%s is not a packed rsh image", path);
    }
    if (src_path) {
        int64_t sz, mt;
        src_stamp(src_path, &sz, &mt);
        if (sz != h.src_size || mt != h.src_mtime_ns) { fclose(f); return 2; }
    }
    emsar_rsh *r = (emsar_rsh *)calloc(1, sizeof(emsar_rsh));
    r->T = h.T; r->nF = h.nF; r->min_fraglength = h.min_fraglength; r->max_fraglength = h.max_fraglength; r->readlength = h.readlength;
    r->max_t_size = h.max_t_size; r->frag_min = h.frag_min; r->frag_max = h.frag_max; r->C = h.C;
    r->class_ptr = (int64_t *)malloc(sizeof(int64_t) * ((size_t)h.C + 1));
    r->class_tid = (int32_t *)malloc(sizeof(int32_t) * (size_t)(h.nnz > 0 ? h.nnz : 1));
    r->euma = (int32_t *)malloc(sizeof(int32_t) * (size_t)h.C * h.nF);
    r->has_node = (uint8_t *)malloc((size_t)h.C);
    char *blob = (char *)malloc((size_t)h.names_bytes);
    r->names = (char **)calloc((size_t)h.T, sizeof(char *));
    int ok = r->class_ptr && r->class_tid && r->euma && r->has_node && blob && r->names;
    ok = ok && fread(r->class_ptr, sizeof(int64_t), (size_t)h.C + 1, f) == (size_t)h.C + 1;
    ok = ok && fread(r->class_tid, sizeof(int32_t), (size_t)h.nnz, f) == (size_t)h.nnz;
    ok = ok && fread(r->euma, sizeof(int32_t), (size_t)h.C * h.nF, f) == (size_t)h.C * h.nF;
    ok = ok && fread(r->has_node, 1, (size_t)h.C, f) == (size_t)h.C;
    ok = ok && fread(blob, 1, (size_t)h.names_bytes, f) == (size_t)h.names_bytes;
    if (ok && h.aux_present) {
        ok = h.aux_nnz_multi == h.nnz - h.T;
        r->aux_txm_off = (uint32_t *)malloc(sizeof(uint32_t) * ((size_t)h.T + 1));
        r->aux_txm_cid = (int32_t *)malloc(sizeof(int32_t) * (size_t)(h.aux_nnz_multi > 0 ? h.aux_nnz_multi : 1));
        r->aux_order = (int32_t *)malloc(sizeof(int32_t) * (size_t)h.T);
        r->aux_insertable = (uint8_t *)malloc((size_t)(h.C - h.T > 0 ? h.C - h.T : 1));
        r->aux_owned = 1;
        ok = ok && r->aux_txm_off && r->aux_txm_cid && r->aux_order && r->aux_insertable;
        ok = ok && fread(r->aux_txm_off, sizeof(uint32_t), (size_t)h.T + 1, f) == (size_t)h.T + 1;
        ok = ok && (h.aux_nnz_multi == 0 || fread(r->aux_txm_cid, sizeof(int32_t), (size_t)h.aux_nnz_multi, f) == (size_t)h.aux_nnz_multi);
        ok = ok && fread(r->aux_order, sizeof(int32_t), (size_t)h.T, f) == (size_t)h.T;
        ok = ok && (h.C == h.T || fread(r->aux_insertable, 1, (size_t)(h.C - h.T), f) == (size_t)(h.C - h.T));
        ok = ok && r->aux_txm_off[0] == 0 && (int64_t)r->aux_txm_off[h.T] == h.aux_nnz_multi;
        if (ok) { r->has_aux = 1; r->aux_nnz_multi = h.aux_nnz_multi; r->aux_n_sets_nocut = h.aux_n_sets_nocut; r->aux_max_set_tids = h.aux_max_set_tids; }
    }
    fclose(f);
    ok = ok && r->class_ptr[0] == 0 && r->class_ptr[h.C] == h.nnz && blob[h.names_bytes - 1] == 0;
    if (ok) {
        const char *p = blob, *end = blob + h.names_bytes;
        for (int32_t t = 0; t < h.T; t++) {
            if (p >= end) { ok = 0; break; }
            r->names[t] = strdup(p);           /* emsar_rsh_free frees names one by one */
            p += strlen(p) + 1;
        }
    }
    free(blob);
    if (!ok) { emsar_rsh_free(r); return fail(err, "packed rsh image %s is truncated or corrupt", path); }
    uint32_t slots = 16;
    while (slots < (uint32_t)r->T * 2u) slots <<= 1;
    r->name_mask = slots - 1;
    r->name_slots = (uint32_t *)calloc(slots, sizeof(uint32_t));
    for (int32_t t = 0; t < r->T; t++) name_insert(r, t);
    *out = r;
    return 0;
}

/* What the command line uses for -I: `<path>.pack` when it is fresh, else the text (and, with EMSAR_RSH_CACHE set, the image is
 * (re)written for the next run). A path that itself ends in ".pack" is loaded as an image. `from_cache`: 1 when no text was parsed. */
int emsar_rsh_load_auto(const char *path, emsar_rsh **out, int *from_cache, char *err)
{
    if (from_cache) *from_cache = 0;
    const size_t n = strlen(path);
    if (n > 5 && strcmp(path + n - 5, ".pack") == 0) {
        int rc = emsar_rsh_load_packed(path, NULL, out, err);
        if (!rc && from_cache) *from_cache = 1;
        return rc;
    }
    char pk[4096];
    snprintf(pk, sizeof pk, "%s.pack", path);
    struct stat st;
    if (stat(pk, &st) == 0) {
        char e2[EMSAR_HOST_ERRLEN];
        if (emsar_rsh_load_packed(pk, path, out, e2) == 0) { if (from_cache) *from_cache = 1; return 0; }
    }
    int rc = emsar_rsh_load(path, out, err);
    if (!rc && getenv("EMSAR_RSH_CACHE")) { char e2[EMSAR_HOST_ERRLEN]; emsar_rsh_save_packed(*out, pk, path, e2); }
    return rc;
}
