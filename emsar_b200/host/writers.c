/* Output files, byte-compatible with the reference's writers: print_FPKMfinal (emsar_functions.c:3163-3212),
 * print_FraglengthDist (:2477-2493), print_aEUMA_3 (:2262-2300). */
#include <stdarg.h>
#include <string.h>

#include "emsar_host.h"

static int fail(char *err, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    if (err) vsnprintf(err, EMSAR_HOST_ERRLEN, fmt, ap);
    va_end(ap);
    return 1;
}

/* a full disk shows up in ferror() or in fclose()'s flush: a truncated result file must not pass for a complete one */
static int finish(FILE *f, const char *path, char *err)
{
    const int bad = ferror(f);
    if (fclose(f) != 0 || bad) return fail(err, "write error on %s (disk full?)", path);
    return 0;
}

int emsar_write_fpkm(const char *path, const emsar_rsh *r, const double *fpkm, const double *sd, const double *efflen,
                     const double *ireadcount, const int32_t *ireadcount_int, const double *tpm, char *err)
{
    FILE *f = fopen(path, "w");
    if (!f) return fail(err, "Can't write to FPKMfile %s", path);
    fprintf(f, "transcriptID\tFPKM\tsd.of.FPKM\teff.length\tiReadcount\tiReadcount.int\tTPM\n");
    for (int32_t t = 0; t < r->T; t++)
        fprintf(f, "%s\t%lf\t%lf\t%lf\t%lf\t%d\t%lf\n", r->names[t], fpkm[t], sd ? sd[t] : 0.0, efflen[t], ireadcount[t], ireadcount_int[t], tpm[t]);
    return finish(f, path, err);
}

int emsar_write_fraglength(const char *path, const emsar_rsh *r, const int32_t *FraglengthCounts, const double *Wf, char *err)
{
    FILE *f = fopen(path, "w");
    if (!f) return fail(err, "Can't write to fraglength file %s", path);
    fprintf(f, "Fragment.length\tObs.Counts\tnormalized.Fragment.length.sampling.prob\n");
    for (int i = 0; i < r->nF; i++) fprintf(f, "%d\t%d\t%lg\n", i + r->frag_min, FraglengthCounts[i + r->frag_min], Wf[i]);
    return finish(f, path, err);
}

int emsar_write_segments(const char *path, const emsar_rsh *r, const int32_t *set_id, const double *adjEUMA,
                         const int32_t *ReadCount, const double *expected, char *err)
{
    FILE *f = fopen(path, "w");
    if (!f) return fail(err, "Can't write to output aEUMA file %s", path);
    fprintf(f, "segment_id\tsequence_sharing_set_id\ttranscript_id\ttranscript_names\teff.length\tReadcount\texpected_Readcount\n");
    for (int64_t c = 0; c < r->C; c++) {
        int64_t o = r->class_ptr[c], e = r->class_ptr[c + 1];
        fprintf(f, "c%lld\ts%d\t", (long long)c, set_id[c]);
        for (int64_t j = o; j < e; j++) fprintf(f, "%st%d", j > o ? "," : "", r->class_tid[j]);
        fprintf(f, "\t");
        for (int64_t j = o; j < e; j++) fprintf(f, "%s%s", j > o ? "+" : "", r->names[r->class_tid[j]]);
        fprintf(f, "\t%lf\t%d\t%f\n", adjEUMA[c], ReadCount[c], expected[c]);
    }
    return finish(f, path, err);
}
