/*
 * emsar_oracle.c — TEST INFRASTRUCTURE ONLY. Not part of the product path.
 *
 * A plain-C, CPU restatement of the EMSAR v2.0.1 quantification hot path, used by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs as the CHECKER for
 * the CUDA path. The product (emsar_b200/, include/) never links, imports or executes this file.
 *
 * Pinning: the reference ships no tests or golden vectors for this path. The
 * restatement is therefore pinned against OUTPUTS OF THE REFERENCE ITSELF: oracle/_ref/emsar (built
 * from the unmodified sources by oracle/Makefile) is run on generated fixtures and its .segments /
 * .fraglength_effect / .fpkm files are committed under tests/golden/ (script: tests/golden/make_golden.py).
 * Integer results (ReadCount, FraglengthCounts, N) and the deterministic fp64 pre-steps (Wf, adjEUMA,
 * iEUMA) must equal the reference exactly / at print precision; the estimator is compared under the
 * tolerance policy of SURVEY.md §8(c) because the reference's own optimiser is seed dependent.
 *
 * Every function cites the reference lines (relative to /root/reference/src) it follows.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <time.h>

#define NEAR_LOWEST_NUMBER (-9.9E307) /* emsar.h:21 */
#define MAX_NTID_PER_SID 5000         /* emsar.h:17 */
#define EUMACUT_INCREMENT 2           /* emsar.h:18 */

/* ------------------------------------------------------------------------------------------------
 * a4. Read-group filter: add_alignment_to_list (alignment.c:29-60), check_fraglen_discrepancy
 * (alignment.c:85-95) and the `size <= MAX_REPEAT` gate of the readers (emsar_functions.c:372,532,752,819).
 * Input: the non-NULL alignments of ONE read group in file order. Output: indices of the kept ones.
 * Returns the kept count, or -1 when the group is rejected (too many best hits / PE fraglen discrepancy).
 * ---------------------------------------------------------------------------------------------- */
int orc_filter_group(int n, const int *tid, const int *mm, const int *fraglen, const int *pos,
                     int max_repeat, int pe, int *keep)
{
    int size = 0, cur_min = 10000, i, j;
    for (i = 0; i < n; i++) {
        int dup = 0;
        for (j = 0; j < size; j++) { /* duplicate scan runs BEFORE the mismatch test (alignment.c:36-40) */
            int q = keep[j];
            if (tid[i] == tid[q] && pos[i] == pos[q] && fraglen[i] == fraglen[q]) { dup = 1; break; }
        }
        if (dup) continue;
        if (mm[i] > cur_min) continue;                   /* alignment.c:42 */
        if (mm[i] < cur_min) { size = 0; cur_min = mm[i]; } /* alignment.c:43-46 */
        keep[size++] = i;
    }
    if (size > max_repeat) return -1;
    if (pe && size > 0) {
        for (j = 1; j < size; j++) if (fraglen[keep[j]] != fraglen[keep[0]]) return -1;
    }
    return size;
}

/* parse_mmstr (alignment.c:101-108): bowtie mismatch column -> count; parse_SAM_mmstr (:418-424). */
int orc_parse_mmstr(const char *s)
{
    int mm = 0; size_t i, n = strlen(s);
    if (n > 0) mm++;
    for (i = 0; i <= n; i++) if (s[i] == ',') mm++;
    return mm;
}
int orc_parse_sam_mmstr(const char *s)
{
    int mm = 0; size_t i, n = strlen(s);
    for (i = 0; i < n; i++) if (s[i] < '0' || s[i] > '9') mm++;
    return mm;
}

/* ------------------------------------------------------------------------------------------------
 * a1/a5/a6. Class store + read->class counting.
 * Chains are rebuilt exactly as parse_rsh_mainline appends them (emsar_functions.c:1486-1505): a node is
 * created only when the line carries EUMA values; multi-tid nodes are linked in FILE order behind `lastp`;
 * a later singleton line for the same tid replaces the earlier one. Lookup walks the chain like
 * update_rshbucket's 'r' branch (:1597-1624) with cmptarr (:1677-1684), including its early-out on cmp<0.
 * ---------------------------------------------------------------------------------------------- */
static int cmp_key(const int *q, const int *node, int k) /* cmptarr on tids [1..k) (first tid equal by bucket) */
{
    int i;
    for (i = 1; i < k; i++) { if (q[i] < node[i]) return -1; else if (q[i] > node[i]) return 1; }
    return 0;
}

int orc_count(int T, int64_t C, const int64_t *class_ptr, const int32_t *class_tid, const uint8_t *has_node,
              int max_t_size, int min_fraglength, int max_fraglength,
              int64_t n_reads, const int64_t *read_ptr, const int32_t *read_tid, const int32_t *read_fraglen,
              int32_t *ReadCount, int32_t *FraglengthCounts, int64_t *TotalReadCount)
{
    int64_t c, r;
    int nb = max_t_size - 1; /* rshbucket has max_t_size-1 slots (:1339) */
    int64_t **head = (int64_t **)calloc(nb > 0 ? nb : 1, sizeof(int64_t *));
    int64_t *next = (int64_t *)malloc(sizeof(int64_t) * (C > 0 ? C : 1));
    int64_t *single = (int64_t *)malloc(sizeof(int64_t) * (T > 0 ? T : 1));
    int64_t lastp = -1;
    int maxk = 16, *tarr = (int *)malloc(sizeof(int) * maxk);
    if (!head || !next || !single || !tarr) return -1;
    for (c = 0; c < T; c++) single[c] = -1;
    for (c = 0; c < C; c++) next[c] = -1;
    memset(ReadCount, 0, sizeof(int32_t) * C);                        /* clear_readcounts_in_rshbucket :1726 */
    memset(FraglengthCounts, 0, sizeof(int32_t) * (max_fraglength + 1)); /* emsar_main.c:383 */
    *TotalReadCount = 0;
    for (c = 0; c < C; c++) {
        int k = (int)(class_ptr[c + 1] - class_ptr[c]);
        int tid0 = class_tid[class_ptr[c]];
        if (!has_node[c]) continue;                     /* no EUMA on the line: no node (:1486) */
        if (k == 1) { single[tid0] = c; continue; }     /* :1488 */
        if (k - 2 >= nb) continue;                      /* would overflow rshbucket in the reference; malformed */
        if (!head[k - 2]) {
            int t; head[k - 2] = (int64_t *)malloc(sizeof(int64_t) * T);
            for (t = 0; t < T; t++) head[k - 2][t] = -1;
        }
        if (head[k - 2][tid0] < 0) { head[k - 2][tid0] = c; lastp = c; }  /* :1495-1498 */
        else { next[lastp] = c; lastp = c; }                              /* :1499-1502 */
    }
    for (r = 0; r < n_reads; r++) {
        int k = (int)(read_ptr[r + 1] - read_ptr[r]);
        const int32_t *t = read_tid + read_ptr[r];
        int fl = read_fraglen[r];
        if (k <= 0) continue;
        if (!(fl <= max_fraglength && fl >= min_fraglength)) continue;    /* :849 */
        if (k == 1) {
            if (single[t[0]] >= 0) ReadCount[single[t[0]]]++;             /* :1528-1536 */
        } else {
            int tsize = 0, j, ti2, i;
            if (k > maxk) { maxk = k; tarr = (int *)realloc(tarr, sizeof(int) * maxk); }
            for (i = 0; i < k; i++) {                                     /* insertion sort, dups kept (:886-902) */
                int nopush = 0;
                for (j = 0; j < tsize; j++) {
                    if (tarr[j] >= t[i]) {
                        for (ti2 = tsize - 1; ti2 >= j; ti2--) tarr[ti2 + 1] = tarr[ti2];
                        tarr[j] = t[i]; tsize++; nopush = 1; break;
                    }
                }
                if (!nopush) tarr[tsize++] = t[i];
            }
            if (tsize <= max_t_size && head[tsize - 2]) {                 /* :1599-1600 */
                int64_t p = head[tsize - 2][tarr[0]];
                while (p >= 0) {                                          /* :1603-1622 */
                    int cmp = cmp_key(tarr, class_tid + class_ptr[p], tsize);
                    if (cmp < 0) break;
                    if (cmp == 0) { ReadCount[p]++; break; }
                    p = next[p];
                }
            }
        }
        FraglengthCounts[fl]++;                                           /* :940 */
        (*TotalReadCount)++;                                              /* :941 */
    }
    for (c = 0; c < nb; c++) free(head[c]);
    free(head); free(next); free(single); free(tarr);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * a7/a8/a11/a15. Wf (transfer_fraglendist_to_Wf :2503-2513), adjEUMA (compute_adjEUMA :2517-2523 as
 * used by scan_rshbucket :2135-2192: 0 for a singleton without node), EUMAps (construct_EUMAps
 * :3148-3154), iEUMA (compute_iEUMA :3218-3232: all classes, multiplicity, ascending cid).
 * ---------------------------------------------------------------------------------------------- */
void orc_prepare(int T, int64_t C, const int64_t *class_ptr, const int32_t *class_tid, int nF, const int32_t *euma,
                 const uint8_t *has_node, int frag_min, const int32_t *FraglengthCounts, int64_t N, double delta,
                 double *Wf, double *adjEUMA, double *EUMAps, double *iEUMA)
{
    int i; int64_t c, j;
    double sumWf = 0;
    for (i = 0; i < nF; i++) { Wf[i] = FraglengthCounts[i + frag_min]; sumWf += Wf[i]; }
    for (i = 0; i < nF; i++) Wf[i] /= sumWf;
    for (c = 0; c < C; c++) {
        double a = 0;
        if (has_node[c]) for (i = 0; i < nF; i++) a += Wf[i] * (double)euma[c * nF + i];
        adjEUMA[c] = a;
    }
    {
        double p10 = pow(10, delta);
        for (c = 0; c < C; c++) EUMAps[c] = adjEUMA[c] / 1E3 * ((double)N / 1E6) * p10;
    }
    for (i = 0; i < T; i++) iEUMA[i] = 0;
    for (c = 0; c < C; c++) for (j = class_ptr[c]; j < class_ptr[c + 1]; j++) iEUMA[class_tid[j]] += adjEUMA[c];
}

/* ------------------------------------------------------------------------------------------------
 * a9/a10. Sequence-sharing sets with the EUMAcut escape hatch: build_TC_from_CT_2 (:2201-2227),
 * propagate_2 (:2234-2259), the loop at emsar_main.c:411-425. Returns max_sid (number of sets - 1);
 * *eumacut is raised in place by EUMACUT_INCREMENT until every set has <= MAX_NTID_PER_SID transcripts.
 * Set ids are assigned in first-seen cid order; cut classes keep CS = -1. The DFS is iterative here
 * (the reference recurses) — visiting order does not change the labelling.
 * ---------------------------------------------------------------------------------------------- */
int orc_components(int T, int64_t C, const int64_t *class_ptr, const int32_t *class_tid, const double *adjEUMA,
                   double *eumacut, int max_ntid, int32_t *CS, int32_t *TS)
{
    int64_t nnz = class_ptr[C], c, j;
    int64_t *tptr = (int64_t *)calloc(T + 2, sizeof(int64_t));
    int64_t *tcid = (int64_t *)malloc(sizeof(int64_t) * (nnz > 0 ? nnz : 1));
    int64_t *stack = (int64_t *)malloc(sizeof(int64_t) * (C > 0 ? C : 1));
    int nsets;
    if (max_ntid <= 0) max_ntid = MAX_NTID_PER_SID;
    for (j = 0; j < nnz; j++) tptr[class_tid[j] + 2]++;
    for (j = 0; j < T; j++) tptr[j + 2] += tptr[j + 1];
    for (c = 0; c < C; c++) for (j = class_ptr[c]; j < class_ptr[c + 1]; j++) tcid[tptr[class_tid[j] + 1]++] = c;
    /* now tptr[t]..tptr[t+1] is row t */
    for (;;) {
        int reiterate = 0;
        for (c = 0; c < C; c++) CS[c] = -1;
        for (j = 0; j < T; j++) TS[j] = -1;
        nsets = 0;
        for (c = 0; c < C && !reiterate; c++) {
            int ntid = 0; int64_t sp = 0;
            if (CS[c] != -1) continue;
            if (class_ptr[c + 1] - class_ptr[c] > 1 && adjEUMA[c] < *eumacut) continue; /* :2242 */
            CS[c] = nsets; stack[sp++] = c;
            while (sp > 0) {
                int64_t cc = stack[--sp];
                for (j = class_ptr[cc]; j < class_ptr[cc + 1]; j++) {
                    int t = class_tid[j]; int64_t e;
                    if (TS[t] != -1) continue;
                    TS[t] = nsets; ntid++;
                    for (e = tptr[t]; e < tptr[t + 1]; e++) {
                        int64_t c2 = tcid[e];
                        if (CS[c2] != -1) continue;
                        if (class_ptr[c2 + 1] - class_ptr[c2] > 1 && adjEUMA[c2] < *eumacut) continue;
                        CS[c2] = nsets; stack[sp++] = c2;
                    }
                }
            }
            nsets++;
            if (ntid > max_ntid) { *eumacut += EUMACUT_INCREMENT; reiterate = 1; } /* emsar_main.c:417-423 */
        }
        if (!reiterate) break;
    }
    free(tptr); free(tcid); free(stack);
    return nsets - 1;
}

/* ------------------------------------------------------------------------------------------------
 * a13 (objective). Log-likelihood exactly as Fp/lambdap evaluate it (:2946-2975), summed over every class
 * that sits in a set (CS != -1  <=>  in_model != 0): classes with EUMAps == 0 are skipped, lambda == 0 with
 * reads -> NEAR_LOWEST_NUMBER.
 * ---------------------------------------------------------------------------------------------- */
double orc_loglik(int64_t C, const int64_t *class_ptr, const int32_t *class_tid, const int32_t *ReadCount,
                  const double *EUMAps, const uint8_t *in_model, const double *FPKM)
{
    double sum = 0; int64_t c, j;
    for (c = 0; c < C; c++) {
        double s = 0, lamb;
        if (in_model && !in_model[c]) continue;
        if (EUMAps[c] == 0) continue;
        for (j = class_ptr[c]; j < class_ptr[c + 1]; j++) { if (FPKM[class_tid[j]] < 0) return NEAR_LOWEST_NUMBER; s += FPKM[class_tid[j]]; }
        lamb = EUMAps[c] * s;
        if (lamb == 0) { if (ReadCount[c] != 0) return NEAR_LOWEST_NUMBER; }
        else if (lamb < 0) return NEAR_LOWEST_NUMBER;
        else sum += (double)ReadCount[c] * log(lamb) - lamb;
    }
    if (sum < NEAR_LOWEST_NUMBER) sum = NEAR_LOWEST_NUMBER;
    return sum;
}

/* ------------------------------------------------------------------------------------------------
 * a12-a14 replacement. The EM / Richardson-Lucy update that the CUDA path implements (same fixed point
 * as the reference's Poisson MLE):
 *      S_c   = sum_{t in c} theta_t                       (multiplicity kept)
 *      q_c   = R_c / S_c          for multi-tid classes that are modelled (in_model, EUMAps>0) and have R_c>0
 *      n_t   = Rs_t + theta_t * sum_{c∋t} q_c             (ascending cid, multiplicity kept)
 *      theta_t' = n_t / A_t ,   A_t = sum_{c∋t modelled} EUMAps_c ,  Rs_t = R of the modelled singleton
 * Start: theta = 1 where A_t > 0, else 0. Stop after the first iteration with
 *      max_t |theta_t' - theta_t| * A_t / (eps_abs + eps_rel * n_t)  <= 1,   or at max_iter.
 * Closed cases copied from MLE() (:3054-3066): a transcript with A_t == 0 gets 0, except the lone-singleton
 * set, where FPKM = R/EUMAps is evaluated literally (inf when EUMAps == 0 and R > 0).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int T; int64_t C;
    const int64_t *class_ptr; const int32_t *class_tid;
    int64_t *tptr; int64_t *tcid;          /* transpose of the ACTIVE multi classes, ascending cid */
    int64_t n_act; int64_t *act;           /* active multi class ids */
    double *q, *A, *Rs;
} em_ws;

static void em_build(em_ws *w, int T, int64_t C, const int64_t *class_ptr, const int32_t *class_tid,
                     const int32_t *ReadCount, const double *EUMAps, const uint8_t *in_model)
{
    int64_t c, j, nnz = class_ptr[C];
    w->T = T; w->C = C; w->class_ptr = class_ptr; w->class_tid = class_tid;
    w->A = (double *)calloc(T > 0 ? T : 1, sizeof(double));
    w->Rs = (double *)calloc(T > 0 ? T : 1, sizeof(double));
    w->q = (double *)calloc(C > 0 ? C : 1, sizeof(double));
    w->act = (int64_t *)malloc(sizeof(int64_t) * (C > 0 ? C : 1));
    w->tptr = (int64_t *)calloc(T + 2, sizeof(int64_t));
    w->tcid = (int64_t *)malloc(sizeof(int64_t) * (nnz > 0 ? nnz : 1));
    w->n_act = 0;
    for (c = 0; c < C; c++) {
        int k = (int)(class_ptr[c + 1] - class_ptr[c]);
        int modelled = (!in_model || in_model[c]) && EUMAps[c] > 0;
        if (!modelled) continue;
        for (j = class_ptr[c]; j < class_ptr[c + 1]; j++) w->A[class_tid[j]] += EUMAps[c];
        if (k == 1) w->Rs[class_tid[class_ptr[c]]] = (double)ReadCount[c];
        else if (ReadCount[c] > 0) {
            w->act[w->n_act++] = c;
            for (j = class_ptr[c]; j < class_ptr[c + 1]; j++) w->tptr[class_tid[j] + 2]++;
        }
    }
    for (j = 0; j < T; j++) w->tptr[j + 2] += w->tptr[j + 1];
    for (j = 0; j < w->n_act; j++) {
        int64_t e; c = w->act[j];
        for (e = class_ptr[c]; e < class_ptr[c + 1]; e++) w->tcid[w->tptr[class_tid[e] + 1]++] = c;
    }
}
static void em_free(em_ws *w) { free(w->A); free(w->Rs); free(w->q); free(w->act); free(w->tptr); free(w->tcid); }

/* One EM iteration restricted to classes [j0,j1) of the active list (E-step) / transcripts [t0,t1) (M-step). */
static void em_estep(em_ws *w, const int32_t *ReadCount, const double *theta, int64_t j0, int64_t j1)
{
    int64_t j;
    for (j = j0; j < j1; j++) {
        int64_t c = w->act[j], e; double s = 0;
        for (e = w->class_ptr[c]; e < w->class_ptr[c + 1]; e++) s += theta[w->class_tid[e]];
        w->q[c] = s > 0 ? (double)ReadCount[c] / s : 0.0;
    }
}
static double em_mstep(em_ws *w, double *theta, double eps_abs, double eps_rel, int t0, int t1)
{
    int t; double dmax = 0;
    for (t = t0; t < t1; t++) {
        int64_t e; double Q = 0, n, th, d;
        if (!(w->A[t] > 0)) continue;
        for (e = w->tptr[t]; e < w->tptr[t + 1]; e++) Q += w->q[w->tcid[e]];
        n = w->Rs[t] + theta[t] * Q;
        th = n / w->A[t];
        d = fabs(th - theta[t]) * w->A[t] / (eps_abs + eps_rel * n);
        if (d > dmax) dmax = d;
        theta[t] = th;
    }
    return dmax;
}

/* pthread team: static contiguous ranges, two barriers per iteration (no OpenMP runtime in this image). */
typedef struct {
    em_ws *w; const int32_t *R; double *theta; double eps_abs, eps_rel;
    int nthr, max_iter, n_steps; double *steps_out;
    pthread_barrier_t bar; double *dpart; int iters_done; double last_delta; int stop_on_conv;
} em_team;
typedef struct { em_team *tm; int id; } em_arg;

static void *em_worker(void *p)
{
    em_arg *a = (em_arg *)p; em_team *tm = a->tm; em_ws *w = tm->w; int id = a->id, n = tm->nthr, it, k;
    int64_t j0 = w->n_act * id / n, j1 = w->n_act * (id + 1) / n;
    int t0 = (int)((int64_t)w->T * id / n), t1 = (int)((int64_t)w->T * (id + 1) / n);
    for (it = 0; it < tm->max_iter; it++) {
        double d = 0;
        em_estep(w, tm->R, tm->theta, j0, j1);
        pthread_barrier_wait(&tm->bar);
        tm->dpart[id] = em_mstep(w, tm->theta, tm->eps_abs, tm->eps_rel, t0, t1);
        pthread_barrier_wait(&tm->bar);
        for (k = 0; k < n; k++) if (tm->dpart[k] > d) d = tm->dpart[k];
        if (id == 0) {
            if (it < tm->n_steps && tm->steps_out) memcpy(tm->steps_out + (size_t)it * w->T, tm->theta, sizeof(double) * w->T);
            tm->iters_done = it + 1; tm->last_delta = d;
        }
        pthread_barrier_wait(&tm->bar); /* dpart is rewritten next iteration */
        if (tm->stop_on_conv && d <= 1.0) break;
    }
    return NULL;
}

static int em_run(em_ws *w, const int32_t *R, double *theta, double eps_abs, double eps_rel, int max_iter, int nthreads,
                  int stop_on_conv, int n_steps, double *steps_out, double *final_delta)
{
    em_team tm; em_arg *args; pthread_t *th; int i;
    if (nthreads < 1) nthreads = 1;
    tm.w = w; tm.R = R; tm.theta = theta; tm.eps_abs = eps_abs; tm.eps_rel = eps_rel; tm.nthr = nthreads;
    tm.max_iter = max_iter; tm.n_steps = n_steps; tm.steps_out = steps_out; tm.iters_done = 0; tm.last_delta = INFINITY;
    tm.stop_on_conv = stop_on_conv;
    tm.dpart = (double *)calloc(nthreads, sizeof(double));
    args = (em_arg *)malloc(sizeof(em_arg) * nthreads); th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    pthread_barrier_init(&tm.bar, NULL, nthreads);
    for (i = 0; i < nthreads; i++) { args[i].tm = &tm; args[i].id = i; }
    for (i = 1; i < nthreads; i++) pthread_create(&th[i], NULL, em_worker, &args[i]);
    em_worker(&args[0]);
    for (i = 1; i < nthreads; i++) pthread_join(th[i], NULL);
    pthread_barrier_destroy(&tm.bar);
    if (final_delta) *final_delta = tm.last_delta;
    free(tm.dpart); free(args); free(th);
    return tm.iters_done;
}

/* Runs EM to convergence. theta: out (T). Returns the number of iterations performed.
   If `steps_out` is non-NULL, theta after each of the first n_steps iterations is stored there (n_steps*T). */
int orc_em(int T, int64_t C, const int64_t *class_ptr, const int32_t *class_tid, const int32_t *ReadCount,
           const double *EUMAps, const uint8_t *in_model, int max_iter, double eps_abs, double eps_rel,
           int nthreads, double *theta, double *final_delta, int n_steps, double *steps_out)
{
    em_ws w; int it, t; double d = INFINITY;
    int64_t c;
    int32_t *nmodel = (int32_t *)calloc(T > 0 ? T : 1, sizeof(int32_t));
    em_build(&w, T, C, class_ptr, class_tid, ReadCount, EUMAps, in_model);
    for (t = 0; t < T; t++) theta[t] = w.A[t] > 0 ? 1.0 : 0.0;
    it = em_run(&w, ReadCount, theta, eps_abs, eps_rel, max_iter, nthreads, 1, n_steps, steps_out, &d);
    /* closed cases of MLE() (:3054-3066) for transcripts the EM does not touch (A_t == 0) */
    for (c = 0; c < C; c++) {
        int64_t j;
        if (in_model && !in_model[c]) continue;
        for (j = class_ptr[c]; j < class_ptr[c + 1]; j++)
            if (j == class_ptr[c] || class_tid[j] != class_tid[j - 1]) nmodel[class_tid[j]]++; /* sorted: count once */
    }
    for (t = 0; t < T; t++) {
        if (w.A[t] > 0) continue;
        theta[t] = 0.0;
        if (nmodel[t] == 1 && ReadCount[t] > 0) theta[t] = (double)ReadCount[t] / EUMAps[t]; /* :3062-3066, cid == tid */
    }
    if (final_delta) *final_delta = d;
    free(nmodel);
    em_free(&w);
    return it;
}

/* Time `iters` EM iterations on `nthreads` threads (cpu_baseline leg of bench.py). Returns seconds. */
double orc_em_time(int T, int64_t C, const int64_t *class_ptr, const int32_t *class_tid, const int32_t *ReadCount,
                   const double *EUMAps, int iters, int nthreads, double *theta)
{
    em_ws w; int t; struct timespec a, b;
    em_build(&w, T, C, class_ptr, class_tid, ReadCount, EUMAps, NULL);
    for (t = 0; t < T; t++) theta[t] = w.A[t] > 0 ? 1.0 : 0.0;
    em_run(&w, ReadCount, theta, 1e-7, 1e-10, 1, nthreads, 0, 0, NULL, NULL); /* warm */
    clock_gettime(CLOCK_MONOTONIC, &a);
    em_run(&w, ReadCount, theta, 1e-7, 1e-10, iters, nthreads, 0, 0, NULL, NULL);
    clock_gettime(CLOCK_MONOTONIC, &b);
    em_free(&w);
    return (b.tv_sec - a.tv_sec) + 1e-9 * (b.tv_nsec - a.tv_nsec);
}

/* ------------------------------------------------------------------------------------------------
 * a16 (numeric part). print_FPKMfinal (:3163-3212) for one deterministic round, Round_off (:3215-3217),
 * and the expected_Readcount column of print_aEUMA_3 (:2289-2296).
 * ---------------------------------------------------------------------------------------------- */
static int round_off(double x) { return (x - (int)x >= 0.5 ? (int)x + 1 : (int)x); }

void orc_finalize(int T, int64_t C, const int64_t *class_ptr, const int32_t *class_tid, const double *FPKM,
                  const double *adjEUMA, const double *iEUMA, int64_t N,
                  double *iReadcount, int32_t *iReadcount_int, double *TPM, double *expected, int64_t *total_ireadcount)
{
    double totalFPKM = 0; int t; int64_t c, j, tot = 0;
    for (t = 0; t < T; t++) totalFPKM += FPKM[t];
    for (t = 0; t < T; t++) {
        iReadcount[t] = (iEUMA[t] / 1E3) * FPKM[t] * ((double)N / 1E6);
        iReadcount_int[t] = round_off(iReadcount[t]);
        tot += iReadcount_int[t];
        TPM[t] = FPKM[t] * 1E6 / totalFPKM;
    }
    if (expected) for (c = 0; c < C; c++) {
        double e = 0;
        for (j = class_ptr[c]; j < class_ptr[c + 1]; j++) e += FPKM[class_tid[j]] * (adjEUMA[c] / 1E3) * ((double)N / 1E6);
        expected[c] = e;
    }
    if (total_ireadcount) *total_ireadcount = tot;
}
