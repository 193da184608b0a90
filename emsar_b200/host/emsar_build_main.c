/* emsar-build — rsh index from a transcriptome fasta. Same command line and output file as the reference's emsar-build
 * (parklab/emsar v2.0.1 src/emsar_build_main.c): `emsar-build <options> fastafile readlength outdir outprefix` writes
 * outdir/outprefix.rsh. The construction is emsar_b200/host/build_index.c on the CPU, or - with --device N (or EMSAR_BUILD_DEVICE=N) -
 * emsar_build_classes_run of libemsar_cuda.so on GPU N (radix sort of hashed read-length windows). The library is loaded
 * at run time only when asked for, so that the tool keeps working on a machine without CUDA; the file is the same either way. */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <getopt.h>
#include <libgen.h>
#include <unistd.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "emsar_host.h"

static void usage(const char *p)
{
    printf("usage: %s <options> fastafile readlength outdir outprefix\n"
           "  readlength   a number (PE), or a number / range such as 48-52 (SE)\n"
           "  -P  paired-end        -s ns|ssf|ssr|ssfr|ssrf        -F / -f  max / min fragment length (PE; default 400 / 1)\n"
           "  -k  MAX_REPEAT (default 100)        -h E|R  fasta header: Ensembl (default) or RefSeq        -q / -v\n"
           "  -p  threads (paired-end construction)        -b -t are accepted for compatibility (bin size, tag length of the reference's suffix arrays)\n"
           "  --device N  construct the classes on GPU N (libemsar_cuda.so) instead of on the CPU; same output\n", p);
}

int main(int argc, char *argv[])
{
    emsar_build_opts o;
    memset(&o, 0, sizeof o);
    o.min_fraglength = 1; o.max_fraglength = 400; o.max_repeat = 100; o.header = 'E';
    char strand[8] = "ns";
    int verbose = 1, c, oi;
    static struct option lo[] = {{"PE", no_argument, 0, 'P'}, {"strand_type", required_argument, 0, 's'}, {"maxfraglen", required_argument, 0, 'F'},
                                 {"minfraglen", required_argument, 0, 'f'}, {"max_repeat", required_argument, 0, 'k'}, {"header", required_argument, 0, 'h'},
                                 {"maxthread", required_argument, 0, 'p'}, {"binsize", required_argument, 0, 'b'}, {"taglen", required_argument, 0, 't'},
                                 {"bias_model", required_argument, 0, 'm'}, {"print_sfa", no_argument, 0, 'T'}, {"verbose", no_argument, 0, 'v'},
                                 {"no_verbose", no_argument, 0, 'q'}, {"device", required_argument, 0, 1000}, {0, 0, 0, 0}};
    int device = getenv("EMSAR_BUILD_DEVICE") && *getenv("EMSAR_BUILD_DEVICE") ? atoi(getenv("EMSAR_BUILD_DEVICE")) : -1;
    while ((c = getopt_long(argc, argv, "Ps:F:f:k:h:p:b:t:m:W:w:Tvq", lo, &oi)) != -1) {
        switch (c) {
        case 'P': o.pe = 1; break;
        case 's': strncpy(strand, optarg, sizeof strand - 1); break;
        case 'F': o.max_fraglength = atoi(optarg); break;
        case 'f': o.min_fraglength = atoi(optarg); break;
        case 'k': o.max_repeat = atoi(optarg); break;
        case 'h': o.header = optarg[0]; if (o.header != 'E' && o.header != 'R') { fprintf(stderr, "error: invalid fasta option.\n"); return 0; } break;
        case 'm': if (optarg[0] != '0') { fprintf(stderr, "error: the positional bias model (-m 1) is not supported.\n"); return 1; } break;
        case 'T': fprintf(stderr, "error: -T (print suffix array) is not supported: this builder has no suffix array.\n"); return 1;
        case 'p': o.threads = atoi(optarg); break;
        case 'b': case 't': case 'W': case 'w': break;
        case 'v': verbose = 2; break;
        case 'q': verbose = 0; break;
        case 1000: device = atoi(optarg); break;
        default: return 0;
        }
    }
    if (o.min_fraglength > o.max_fraglength || o.min_fraglength < 1 || o.max_fraglength < 1) { fprintf(stderr, "error: invalid fragment length range.\n"); return 1; }
    if (!strcmp(strand, "ns")) o.stranded = 0;
    else if ((!strcmp(strand, "ssf") || !strcmp(strand, "ssr")) && !o.pe) o.stranded = 1;
    else if ((!strcmp(strand, "ssfr") || !strcmp(strand, "ssrf")) && o.pe) o.stranded = 1;
    else { fprintf(stderr, "error: invalid strand type.\n"); return 1; }
    if (optind + 3 >= argc) { usage(argv[0]); return 0; }
    const char *fasta = argv[optind], *outdir = argv[optind + 2], *prefix = argv[optind + 3];
    char *rl = argv[optind + 1];
    if (o.pe) o.readlength = atoi(rl);
    else {                                              /* parse_readlength_range (:2461-2469) */
        char *dash = strchr(rl, '-');
        if (dash) { o.readlen_max = atoi(dash + 1); *dash = 0; o.readlen_min = atoi(rl); }
        else o.readlen_min = o.readlen_max = atoi(rl);
    }
    char cmd[4200], path[4200], err[EMSAR_HOST_ERRLEN] = "";
    snprintf(cmd, sizeof cmd, "mkdir -p %s", outdir);
    if (system(cmd) != 0) { fprintf(stderr, "can't create output directory %s\n", outdir); return 1; }
    emsar_rsh *r = NULL;
    void *cuda = NULL, *dctx = NULL;
    int (*dclose)(void *) = NULL;
    if (device >= 0) {
        /* libemsar_cuda.so lies one directory above this binary (emsar_b200/bin/emsar-build) */
        char self[4096], lib[4300];
        ssize_t n = readlink("/proc/self/exe", self, sizeof self - 1);
        if (n <= 0) { fprintf(stderr, "error: cannot locate the executable\n"); return 1; }
        self[n] = 0;
        snprintf(lib, sizeof lib, "%s/../libemsar_cuda.so", dirname(self));
        cuda = dlopen(lib, RTLD_NOW | RTLD_GLOBAL);
        if (!cuda) cuda = dlopen("libemsar_cuda.so", RTLD_NOW | RTLD_GLOBAL);
        if (!cuda) { fprintf(stderr, "error: --device needs libemsar_cuda.so (%s)\n", dlerror()); return 1; }
        int (*dopen)(int, void **) = (int (*)(int, void **))dlsym(cuda, "emsar_cuda_open");
        dclose = (int (*)(void *))dlsym(cuda, "emsar_cuda_close");
        o.device_run = (int (*)(void *, const struct emsar_build_desc *, struct emsar_build_classes *))dlsym(cuda, "emsar_build_classes_run");
        o.device_free = (void (*)(struct emsar_build_classes *))dlsym(cuda, "emsar_build_classes_free");
        o.device_error = (const char *(*)(void))dlsym(cuda, "emsar_cuda_last_error");
        if (!dopen || !dclose || !o.device_run || !o.device_free || !o.device_error) { fprintf(stderr, "error: libemsar_cuda.so lacks the index construction entry points\n"); return 1; }
        if (dopen(device, &dctx)) { fprintf(stderr, "error: %s\n", o.device_error()); return 1; }
        o.device_ctx = dctx;
    }
    if (emsar_rsh_build(fasta, &o, &r, err)) { fprintf(stderr, "%s\n", err); return 1; }
    if (dctx) dclose(dctx);
    snprintf(path, sizeof path, "%s/%s.rsh", outdir, prefix);
    if (emsar_rsh_write(r, o.pe, path, err)) { fprintf(stderr, "%s\n", err); return 1; }
    if (verbose > 0) printf("max_tid=%d, rshsize=%lld, max_cid=%lld\nwrote %s\n", r->T - 1, (long long)(r->C - r->T), (long long)r->C - 1, path);
    emsar_rsh_free(r);
    return 0;
}
