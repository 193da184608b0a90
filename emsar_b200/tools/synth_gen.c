/* synth_gen.c - fast seeded generators for the human-scale synthetic workloads.
 *
 * BENCH / TEST TOOLING, not part of the product path: emsar_b200/synth.py calls it to build the inputs of bench.py and of the
 * full-size parity tests (100 M reads in seconds instead of minutes of numpy). Every item (class, read) draws from its own
 * counter-based generator keyed by (seed, item index), so the output does not depend on the number of OpenMP threads.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t splitmix(uint64_t *s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static inline uint64_t item_state(uint64_t seed, uint64_t item)
{
    uint64_t s = seed * 0xD1342543DE82EF95ULL + item * 0x9E3779B97F4A7C15ULL + 0x2545F4914F6CDD1DULL;
    splitmix(&s);
    return s;
}
static inline double u01(uint64_t *s) { return (double)(splitmix(s) >> 11) * (1.0 / 9007199254740992.0); }
static inline uint64_t below(uint64_t *s, uint64_t n) { return (uint64_t)(u01(s) * (double)n) % (n ? n : 1); }

static int cmp_i32(const void *a, const void *b)
{
    const int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    return x < y ? -1 : x > y;
}
static void sort_i32(int32_t *v, int n)
{
    if (n <= 24) {
        for (int i = 1; i < n; i++) { int32_t x = v[i]; int j = i - 1; while (j >= 0 && v[j] > x) { v[j + 1] = v[j]; j--; } v[j + 1] = x; }
    } else qsort(v, (size_t)n, sizeof(int32_t), cmp_i32);
}

/* n classes of cardinality k: class i takes k distinct positions of the virtual window [vlo[i], vlo[i] + w[i]) (selection sampling,
 * Knuth 3.4.2 S), maps them to transcript ids through vmap and sorts them. With probability p_dup one member is replaced by a copy of
 * another (a transcript holding the same k-mer twice, reference emsar_functions.c:1792). Hub transcripts: hub_ptr/hub_tid list the
 * hubs of every group (indexed by grp[i]); each is forced into the class with probability hub_p[grp[i]]. rows: int32[n * k]. */
int synth_gen_classes(int64_t n, int32_t k, const int64_t *vlo, const int64_t *w, const int32_t *vmap, uint64_t seed, double p_dup,
                      const int32_t *grp, const int64_t *hub_ptr, const int32_t *hub_tid, const double *hub_p, int32_t *rows)
{
    int bad = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(| : bad)
    for (int64_t i = 0; i < n; i++) {
        uint64_t st = item_state(seed, (uint64_t)i);
        int32_t *r = rows + i * (int64_t)k;
        const int64_t W = w[i], lo = vlo[i];
        if (W < k) { bad |= 1; for (int j = 0; j < k; j++) r[j] = 0; continue; }
        int got = 0;
        for (int64_t j = 0; j < W && got < k; j++)
            if (u01(&st) * (double)(W - j) < (double)(k - got)) r[got++] = vmap[lo + j];
        while (got < k) r[got++] = vmap[lo + W - 1];          /* rounding at the very end of the window (never in practice) */
        if (hub_ptr && grp) {
            const int g = grp[i];
            for (int64_t h = hub_ptr[g]; h < hub_ptr[g + 1]; h++) {
                if (u01(&st) >= hub_p[g]) continue;
                const int32_t ht = hub_tid[h];
                int present = 0;
                for (int j = 0; j < k; j++) if (r[j] == ht) { present = 1; break; }
                if (present) continue;
                /* replace a member that is not itself a hub of the group */
                for (int tries = 0; tries < 8; tries++) {
                    const int j = (int)below(&st, (uint64_t)k);
                    int is_hub = 0;
                    for (int64_t h2 = hub_ptr[g]; h2 < hub_ptr[g + 1]; h2++) if (hub_tid[h2] == r[j]) { is_hub = 1; break; }
                    if (!is_hub) { r[j] = ht; break; }
                }
            }
        }
        if (k > 1 && u01(&st) < p_dup) {
            const int src = (int)below(&st, (uint64_t)k);
            const int dst = (src + 1 + (int)below(&st, (uint64_t)(k - 1))) % k;
            r[dst] = r[src];
        }
        sort_i32(r, k);
    }
    return bad;
}

/* Walker / Vose alias tables for p[0..n) (any non-negative weights). prob[i] in [0,1], alias[i] in [0,n). */
int synth_alias_build(int64_t n, const double *p, double *prob, int64_t *alias)
{
    double sum = 0;
    for (int64_t i = 0; i < n; i++) sum += p[i];
    if (!(sum > 0)) return 1;
    int64_t *small = (int64_t *)malloc(sizeof(int64_t) * (size_t)n), *large = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    if (!small || !large) { free(small); free(large); return 2; }
    int64_t ns = 0, nl = 0;
    const double scale = (double)n / sum;
    for (int64_t i = 0; i < n; i++) {
        prob[i] = p[i] * scale;
        alias[i] = i;
        if (prob[i] < 1.0) small[ns++] = i; else large[nl++] = i;
    }
    while (ns > 0 && nl > 0) {
        const int64_t s = small[--ns], l = large[nl - 1];
        alias[s] = l;
        prob[l] = (prob[l] + prob[s]) - 1.0;
        if (prob[l] < 1.0) { nl--; small[ns++] = l; }
    }
    while (nl > 0) prob[large[--nl]] = 1.0;
    while (ns > 0) prob[small[--ns]] = 1.0;
    free(small); free(large);
    return 0;
}

/* Pass 1: the class of every read (alias draw; -1 = a tid pair that matches no class, with probability p_un) and its list length
 * into read_ptr[r + 1]; then the prefix sum. Returns the total number of tids. */
int64_t synth_reads_pass1(int64_t n, uint64_t seed, int64_t C, const int64_t *class_ptr, const double *prob, const int64_t *alias, double p_un,
                          int64_t *cls, int64_t *read_ptr)
{
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n; r++) {
        uint64_t st = item_state(seed, (uint64_t)r);
        if (u01(&st) < p_un) { cls[r] = -1; read_ptr[r + 1] = 2; continue; }
        const double x = u01(&st) * (double)C;
        int64_t c = (int64_t)x;
        if (c >= C) c = C - 1;
        if ((x - (double)c) >= prob[c]) c = alias[c];
        cls[r] = c;
        read_ptr[r + 1] = class_ptr[c + 1] - class_ptr[c];
    }
    read_ptr[0] = 0;
    for (int64_t r = 0; r < n; r++) read_ptr[r + 1] += read_ptr[r];
    return read_ptr[n];
}

/* Pass 2: the tid lists (the class's members rotated by a random offset, reversed half of the time: the readers hand
 * update_ReadCounts unsorted lists) and the fragment lengths (inverse CDF over nF lengths). */
void synth_reads_pass2(int64_t n, uint64_t seed, int32_t T, const int64_t *class_ptr, const int32_t *class_tid, const int64_t *cls,
                       const int64_t *read_ptr, int32_t nF, int32_t frag_min, const double *frag_cdf, int32_t *read_tid, int32_t *read_fraglen)
{
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n; r++) {
        uint64_t st = item_state(seed ^ 0xA5A5A5A5DEADBEEFULL, (uint64_t)r);
        int32_t *out = read_tid + read_ptr[r];
        const int64_t c = cls[r];
        if (c < 0) {
            const int32_t a = (int32_t)below(&st, (uint64_t)T);
            out[0] = a;
            out[1] = (int32_t)(((int64_t)a + T / 2 + (int64_t)below(&st, (uint64_t)(T / 4 > 0 ? T / 4 : 1))) % T);
        } else {
            const int64_t o = class_ptr[c], k = class_ptr[c + 1] - o;
            const int64_t rot = (int64_t)below(&st, (uint64_t)k);
            const int rev = u01(&st) < 0.5;
            for (int64_t j = 0; j < k; j++) {
                int64_t p = (j + rot) % k;
                if (rev) p = k - 1 - p;
                out[j] = class_tid[o + p];
            }
        }
        if (nF <= 1) read_fraglen[r] = frag_min;
        else {
            const double u = u01(&st);
            int lo = 0, hi = nF - 1;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (frag_cdf[mid] > u) hi = mid; else lo = mid + 1; }
            read_fraglen[r] = frag_min + lo;
        }
    }
}

/* EUMA[c][f] = max(0, floor(a_c - slope_c * f)) for nF fragment lengths (row-major int32), in parallel. */
void synth_euma_fill(int64_t C, int32_t nF, const double *a, const double *slope, int32_t *euma)
{
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < C; c++) {
        int32_t *row = euma + c * (int64_t)nF;
        for (int f = 0; f < nF; f++) {
            const double v = a[c] - slope[c] * (double)f;
            row[f] = v > 0 ? (int32_t)v : 0;
        }
    }
}
