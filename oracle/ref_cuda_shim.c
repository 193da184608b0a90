/* ref_cuda_shim.c - TEST INFRASTRUCTURE: the binding INTEGRATION.md describes, compiled.
 *
 * oracle/Makefile target `ref_cuda` compiles the reference's OWN sources from where they lie under /root/reference/src into
 * oracle/_ref/emsar_cuda, with this file linked in and three one-line edits applied to a temporary copy of its emsar_main.c by
 * sed (the edits are the three seams of SURVEY.md section 8b; no reference source is stored in this repository):
 *   emsar_main.c:288-290   after the hook pointers are bound          -> emsar_shim_bind()      (the 'r' role of the hooks moves to the GPU)
 *   emsar_main.c:~404      after scan_rshbucket() flattened the store -> emsar_shim_counts()    (ReadCount[] comes from the device)
 *   emsar_main.c:446       run_MLE_threads()                          -> emsar_shim_estimate()  (FPKM[] comes from emsar_sample_solve)
 * Everything else - option parsing, the readers and their read-group filters, Wf, adjEUMA, the set decomposition, EUMAps, iEUMA and the
 * writers - stays the reference's code. The test (tests/test_integration_gpu.py) runs it next to the unmodified binary on a golden fixture.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "emsar.h"          /* the reference's header: node1, rshbucket, rshbucket_single, the globals and the hook pointers */
#include "emsar_cuda.h"

static emsar_ctx *g_ctx;
static emsar_index *g_index;
static emsar_sample *g_sample;
static int64_t *b_ptr; static int32_t *b_tid, *b_fl;        /* the current batch of read groups */
static int64_t b_n, b_ntid, b_cap_n, b_cap_tid;
static char (*ref_update)(int, int *, char, int, char *);
static char (*ref_update_single)(int, char, int, char *);
static void (*ref_clear)(void);

static void ok(int rc, const char *what)
{
    if (rc) { fprintf(stderr, "%s: %s: %s\n", what, emsar_cuda_strerror(rc), emsar_cuda_last_error()); exit(1); }
}

/* the class store flattened in the order scan_rshbucket (emsar_functions.c:2149-2191) walks it */
static void build_index(void)
{
    const int T = (int)max_tid + 1;
    int64_t C = 0, nnz = 0;
    for (int j = 0; j < T; j++) { node1 *p = rshbucket_single[j]; if (!p) { C++; nnz++; } for (; p; p = p->next) { C++; nnz++; } }
    for (int k = 2; k <= rshbucket_max_t_size; k++)
        if (rshbucket[k - 2]) for (int j = 0; j < T; j++) for (node1 *p = rshbucket[k - 2][j]; p; p = p->next) { C++; nnz += k; }
    int64_t *cp = malloc(sizeof(int64_t) * (size_t)(C + 1));
    int32_t *ct = malloc(sizeof(int32_t) * (size_t)nnz), *eu = calloc((size_t)C * (size_t)nFraglen, sizeof(int32_t));
    uint8_t *hn = malloc((size_t)C);
    int64_t c = 0, o = 0;
    for (int j = 0; j < T; j++) {
        node1 *p = rshbucket_single[j];
        if (!p) { cp[c] = o; ct[o++] = j; hn[c++] = 0; continue; }
        for (; p; p = p->next) { cp[c] = o; ct[o++] = j; hn[c] = 1; memcpy(eu + c * nFraglen, p->EUMA, sizeof(int) * (size_t)nFraglen); c++; }
    }
    if (c != T) { fprintf(stderr, "shim: a transcript with several singleton nodes is not supported\n"); exit(1); }
    for (int k = 2; k <= rshbucket_max_t_size; k++)
        if (rshbucket[k - 2]) for (int j = 0; j < T; j++) for (node1 *p = rshbucket[k - 2][j]; p; p = p->next) {
            cp[c] = o; ct[o++] = j;
            for (int m = 1; m < k; m++) ct[o++] = p->tarr[m - 1];
            hn[c] = 1; memcpy(eu + c * nFraglen, p->EUMA, sizeof(int) * (size_t)nFraglen); c++;
        }
    cp[C] = o;
    emsar_index_desc d;
    memset(&d, 0, sizeof d);
    d.T = T; d.C = C; d.class_ptr = cp; d.class_tid = ct; d.nF = nFraglen; d.euma = eu; d.has_node = hn;
    d.min_fraglength = Min_Fraglength; d.max_fraglength = Max_Fraglength; d.readlength = readlength; d.max_t_size = rshbucket_max_t_size;
    ok(emsar_cuda_open(0, &g_ctx), "emsar_cuda_open");
    ok(emsar_index_create(g_ctx, &d, &g_index), "emsar_index_create");
    free(cp); free(ct); free(eu); free(hn);
}

static void flush(void)
{
    if (b_n > 0) { b_ptr[b_n] = b_ntid; ok(emsar_sample_count(g_sample, b_n, b_ptr, b_tid, b_fl), "emsar_sample_count"); ok(emsar_sample_count_wait(g_sample, 0), "emsar_sample_count_wait"); }
    b_n = 0; b_ntid = 0;
}

static void push(int n, const int *t, int t0, int fraglen)
{
    if (b_n + 2 > b_cap_n) { b_cap_n = b_cap_n ? 2 * b_cap_n : 1 << 16; b_ptr = realloc(b_ptr, sizeof(int64_t) * (size_t)(b_cap_n + 1)); b_fl = realloc(b_fl, sizeof(int32_t) * (size_t)b_cap_n); }
    if (b_ntid + n + 1 > b_cap_tid) { b_cap_tid = 2 * (b_cap_tid + n) + 1024; b_tid = realloc(b_tid, sizeof(int32_t) * (size_t)b_cap_tid); }
    b_ptr[b_n] = b_ntid; b_fl[b_n] = fraglen; b_n++;
    if (t) for (int i = 0; i < n; i++) b_tid[b_ntid++] = t[i]; else b_tid[b_ntid++] = t0;
    if (b_n >= (1 << 20)) flush();
}

/* update_rshbucket_PTR / update_rshbucket_single_PTR (emsar.h:219-221): 'e' builds the index (stays the reference's), 'r' counts a read */
static char shim_update(int t_size, int *tarray, char type, int fraglen, char *poscat)
{
    if (type != 'r') return ref_update(t_size, tarray, type, fraglen, poscat);
    push(t_size, tarray, 0, fraglen);
    return 0;
}
static char shim_update_single(int tid, char type, int fraglen, char *poscat)
{
    if (type != 'r') return ref_update_single(tid, type, fraglen, poscat);
    push(1, NULL, tid, fraglen);
    return 0;
}
/* clear_readcounts_in_rshbucket_PTR (emsar_main.c:384): top of the per-file loop -> a new sample */
static void shim_clear(void)
{
    ref_clear();
    if (!g_index) build_index();
    if (g_sample) ok(emsar_sample_end(g_sample), "emsar_sample_end");
    ok(emsar_sample_begin(g_index, &g_sample), "emsar_sample_begin");
    b_n = 0; b_ntid = 0;
}

void emsar_shim_bind(void)
{
    ref_update = update_rshbucket_PTR; ref_update_single = update_rshbucket_single_PTR; ref_clear = clear_readcounts_in_rshbucket_PTR;
    update_rshbucket_PTR = shim_update; update_rshbucket_single_PTR = shim_update_single; clear_readcounts_in_rshbucket_PTR = shim_clear;
}

/* after scan_rshbucket(): the reference's ReadCount[] (all zero: its chains were never incremented) is replaced by the device's counts */
void emsar_shim_counts(void)
{
    flush();
    int32_t *F = malloc(sizeof(int32_t) * (size_t)(Max_Fraglength + 1));
    int64_t N = 0;
    ok(emsar_sample_counts_get(g_sample, ReadCount, F, &N), "emsar_sample_counts_get");
    for (int f = 0; f <= Max_Fraglength; f++)
        if (F[f] != FraglengthCounts[f]) { fprintf(stderr, "shim: FraglengthCounts[%d] differs: device %d, host %d\n", f, F[f], FraglengthCounts[f]); exit(1); }
    if (N != TotalReadCount) { fprintf(stderr, "shim: TotalReadCount differs: device %lld, host %d\n", (long long)N, TotalReadCount); exit(1); }
    free(F);
}

/* run_MLE_threads() (emsar_main.c:446): FPKM[] of every transcript from the device */
void emsar_shim_estimate(void)
{
    emsar_solve_opts so;
    emsar_solve_out out;
    memset(&so, 0, sizeof so); memset(&out, 0, sizeof out);
    so.delta = DELTA; so.eumacut = EUMAcut;
    out.fpkm = FPKM;
    ok(emsar_sample_solve(g_sample, &so, &out), "emsar_sample_solve");
    if (out.eumacut != EUMAcut) { fprintf(stderr, "shim: EUMAcut differs: device %.0f, host %.0f\n", out.eumacut, EUMAcut); exit(1); }
    fprintf(stdout, "emsar_cuda: EM finished: %d iterations, %.2f ms on the device\n", out.n_iter, out.em_ms);
}
