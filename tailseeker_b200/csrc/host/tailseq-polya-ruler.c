/*
 * tailseq-polya-ruler (B200) -- drop-in for src/polyaruler/polyaruler.c: nine positional arguments
 * (polyaruler.c:402-418), corrected taginfo on stdout, positive-sample distribution written to
 * argv[9]; exit codes 1 (usage / cutoffs), 2 (ruling), 3 (taginfo) as in polyaruler.c:402-444.
 * measure_polya_length() and the resampling run on the GPU through tsk_ruler_records().
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include "tsk_host.h"

#define SCORE_LINEBUF_SIZE   16384
#define TAGINFO_LINEBUF_SIZE 2048

/* load_score_cutoffs(), polyaruler.c:43-102 */
static int load_cutoffs(const char *filename, const char *tileid, uint32_t *cutoffs, int *ncycles)
{
    char line[SCORE_LINEBUF_SIZE], *bptr, *tok;
    const size_t idlen = strlen(tileid);
    FILE *fp = fopen(filename, "r");
    int i;
    if (!fp) { fprintf(stderr, "Cannot open %s.\n", filename); return -1; }
    for (;;) {
        if (!fgets(line, sizeof(line), fp)) {
            fprintf(stderr, "No poly(A) score cutoffs found for tile %s.\n", tileid);
            fclose(fp);
            return -1;
        }
        if (strlen(line) >= idlen && memcmp(line, tileid, idlen) == 0 && line[idlen] == '\t') break;
    }
    fclose(fp);
    bptr = line + idlen + 1;
    for (i = 0; (tok = strsep(&bptr, "\t\r\n, ")) != NULL; i++) {
        if (*tok == 0) break;
        if (i >= TSK_MAX_CUTOFF_CYCLES) {
            fprintf(stderr, "Exceeded the maximum allowed number of cycles.\n");
            return -1;
        }
        cutoffs[i] = tsk_cutoff_to_packed(atof(tok));
    }
    *ncycles = i;
    return 0;
}

static uint32_t *read_sigpack(const char *filename, uint32_t *total_clusters, uint32_t *max_cycles, size_t *nrecords)
{
    uint32_t header[3], *recs = NULL;
    size_t cap = 0, n = 0, words;
    gzFile fp = gzopen(filename, "rb");
    if (!fp) { fprintf(stderr, "Cannot open the input file: %s\n", filename); return NULL; }
    if (gzread(fp, header, sizeof(header)) != (int)sizeof(header)) {
        fprintf(stderr, "Failed to read the header from %s.\n", filename);
        gzclose(fp);
        return NULL;
    }
    if (header[0] != 4) {
        fprintf(stderr, "The file %s was written in a machine with different architecture.\n", filename);
        gzclose(fp);
        return NULL;
    }
    *total_clusters = header[1]; *max_cycles = header[2];
    words = 2 + (size_t)header[2];
    for (;;) {
        int got;
        if (n == cap) {
            cap = cap ? cap * 2 : 65536;
            recs = realloc(recs, sizeof(uint32_t) * words * cap);
            if (!recs) { perror("process_polya_ruling"); gzclose(fp); return NULL; }
        }
        got = gzread(fp, recs + n * words, (unsigned)(words * 4));
        if (got == 0) break;
        if (got < (int)(words * 4)) { fprintf(stderr, "Unexpected end of file.\n"); free(recs); gzclose(fp); return NULL; }
        n++;
    }
    gzclose(fp);
    *nrecords = n;
    return recs;
}

/* output_corrected_polya_measurements(), polyaruler.c:316-388 */
static int rewrite_taginfo(const char *taginfo_file, const int16_t *meas, const char *tile_id)
{
    gzFile fp = gzopen(taginfo_file, "rt");
    if (!fp) return -1;
    for (;;) {
        char line[TAGINFO_LINEBUF_SIZE], *bptr, *tok, *mods = NULL, *umi = NULL;
        uint32_t clusterno, flags = 0;
        int i;
        if (gzgets(fp, line, TAGINFO_LINEBUF_SIZE) == NULL) {
            int err;
            (void)gzerror(fp, &err);
            if (err == 0) break;
            fprintf(stderr, "Error occurred on reading %s.\n", taginfo_file);
            gzclose(fp);
            return -1;
        }
        bptr = line;
        tok = strsep(&bptr, "\t\n\r");
        if (!tok) continue;
        clusterno = atoi(tok);
        if (meas[clusterno] < 0) {                 /* not revised: pass the line through */
            size_t term;
            tok[strlen(tok)] = '\t';
            term = strlen(tok);
            if (term > 0) tok[term - 1] = '\n';
            fputs(tile_id, stdout); fputc('\t', stdout); fputs(line, stdout);
            continue;
        }
        for (i = 0; (tok = strsep(&bptr, "\t\n\r")) != NULL; i++)
            switch (i) {
            case 0: flags = atoi(tok); break;
            case 2: mods = tok; break;
            case 3: umi = tok; break;
            default: break;
            }
        flags |= TSK_PAFLAG_MEASURED_FROM_FLUORESCENCE;
        printf("%s\t%d\t%d\t%d\t%s\t%s\n", tile_id, clusterno, flags, meas[clusterno],
               mods ? mods : "(null)", umi ? umi : "(null)");
    }
    gzclose(fp);
    return 0;
}

/* Creating the CUDA context costs far more than measuring one sample's records (0.7 s against
 * milliseconds): it is created on its own thread while the signal pack is read and inflated, and
 * torn down on one while the taginfo stream is rewritten. */
struct ctx_boot { pthread_t thread; const tsk_params *params; int device, rc; tsk_ctx *ctx; char error[512]; };

static void *boot_context(void *arg)
{
    struct ctx_boot *b = arg;
    b->rc = tsk_ctx_create(&b->ctx, b->device, b->params);
    if (b->rc != 0) snprintf(b->error, sizeof(b->error), "%s", tsk_last_error());
    return NULL;
}

static void *drop_context(void *arg)
{
    tsk_ctx_destroy((tsk_ctx *)arg);
    return NULL;
}

int main(int argc, char *argv[])
{
    static uint32_t cutoffs[TSK_MAX_CUTOFF_CYCLES];
    uint32_t total_clusters = 0, max_cycles = 0, *recs, *hist;
    int ncycles = 0, device = 0;
    size_t nrecords = 0;
    int16_t *lens, *meas;
    tsk_ctx *ctx = NULL;
    tsk_params *p;
    const char *dev = getenv("TAILSEEKER_DEVICE");

    if (argc < 10) {
        fprintf(stderr, "Usage: %s {tile id} {signals} {cutoffs} {min polya} {downhill ext weight} {taginfo} "
                        "{sampling bin count} {positive sampling gap} {positive sampling output}\n", argv[0]);
        return 1;
    }
    const char *tile_id = argv[1], *signals_file = argv[2], *cutoffs_file = argv[3], *taginfo_file = argv[6];
    const int min_len = atoi(argv[4]), bins = atoi(argv[7]);
    const float weight = atof(argv[5]), gap = atof(argv[8]);
    const char *sigdist_file = argv[9];

    if (load_cutoffs(cutoffs_file, tile_id, cutoffs, &ncycles) < 0) return 1;

    /* the ruler needs no importer parameters: a minimal context */
    p = calloc(1, sizeof(*p));
    p->abi_version = TSK_ABI_VERSION;
    p->total_cycles = ncycles > 0 ? ncycles : 1; p->threep_start = 0; p->threep_length = p->total_cycles;
    p->index_length = 0; p->balancer_num_positive_samples = 2; p->balancer_num_negative_samples = 4;
    p->dist_sampling_bins = bins > 0 ? bins : 1; p->fair_sampling_hash_space_size = 1;
    p->max_cctr_scan_left_space = p->max_cctr_scan_right_space = 1;
    p->num_samples = 1; p->samples[0].delimiter_pos = -1;
    if (dev) device = atoi(dev);
    struct ctx_boot boot;
    int threaded;
    memset(&boot, 0, sizeof(boot));
    boot.params = p; boot.device = device;
    threaded = pthread_create(&boot.thread, NULL, boot_context, &boot) == 0;
    if (!threaded) boot_context(&boot);
    recs = read_sigpack(signals_file, &total_clusters, &max_cycles, &nrecords);
    if (threaded) pthread_join(boot.thread, NULL);
    if (!recs) return 2;
    if (boot.rc != 0) { fprintf(stderr, "%s\n", boot.error); return 2; }
    ctx = boot.ctx;

    lens = malloc(sizeof(int16_t) * (nrecords ? nrecords : 1));
    meas = malloc(sizeof(int16_t) * (total_clusters ? total_clusters : 1));
    hist = calloc((size_t)(ncycles > 0 ? ncycles : 1) * (bins > 0 ? bins : 1), sizeof(uint32_t));
    if (!lens || !meas || !hist) { perror("process_polya_ruling"); return 2; }
    memset(meas, 0xff, sizeof(int16_t) * (total_clusters ? total_clusters : 1));
    if (ncycles > 0 && bins > 0) {
        if (tsk_ruler_hist_reset(ctx, ncycles, bins) != 0 ||
            tsk_ruler_records(ctx, recs, nrecords, (int)max_cycles, cutoffs, ncycles, min_len, weight, bins, gap,
                              lens, 0) != 0 ||
            tsk_ruler_get_hist(ctx, hist) != 0) {
            fprintf(stderr, "%s\n", tsk_last_error());
            return 2;
        }
        for (size_t r = 0; r < nrecords; r++)
            if (lens[r] >= 0) meas[recs[r * (2 + (size_t)max_cycles)]] = lens[r];
    }
    {
        const uint32_t h[3] = {4, (uint32_t)ncycles, (uint32_t)bins};
        gzFile f = gzopen(sigdist_file, "wb");
        if (!f) { fprintf(stderr, "Cannot open %s to write.\n", sigdist_file); return 2; }
        if (gzwrite(f, h, sizeof(h)) <= 0 ||
            (ncycles > 0 && bins > 0 && gzwrite(f, hist, (unsigned)(sizeof(uint32_t) * (size_t)ncycles * bins)) <= 0)) {
            gzclose(f);
            return 2;
        }
        gzclose(f);
    }
    {
        pthread_t th;
        const int bg = pthread_create(&th, NULL, drop_context, ctx) == 0;
        if (!bg) tsk_ctx_destroy(ctx);
        const int rc = rewrite_taginfo(taginfo_file, meas, tile_id);
        fflush(stdout);
        if (bg) pthread_join(th, NULL);
        if (rc < 0) return 3;
    }
    return 0;
}
