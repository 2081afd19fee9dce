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
static int rewrite_taginfo(const char *taginfo_file, const int16_t *meas, uint32_t total_clusters,
                           const char *tile_id, FILE *out)
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
        if (clusterno >= total_clusters) {          /* a taginfo of another tile (the reference reads out of bounds) */
            fprintf(stderr, "Cluster %u in %s is beyond the %u clusters of the signal pack.\n",
                    clusterno, taginfo_file, total_clusters);
            gzclose(fp);
            return -1;
        }
        if (meas[clusterno] < 0) {                 /* not revised: pass the line through */
            size_t term;
            tok[strlen(tok)] = '\t';
            term = strlen(tok);
            if (term > 0) tok[term - 1] = '\n';
            fputs(tile_id, out); fputc('\t', out); fputs(line, out);
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
        fprintf(out, "%s\t%d\t%d\t%d\t%s\t%s\n", tile_id, clusterno, flags, meas[clusterno],
               mods ? mods : "(null)", umi ? umi : "(null)");
    }
    gzclose(fp);
    return 0;
}

/* Creating the CUDA context costs far more than measuring one sample's records (0.7 s or more
 * against milliseconds): it is created on its own thread while the first signal pack is read and
 * inflated, and ONE context serves every job of a batch (`--batch`, below). */
struct ctx_boot { pthread_t thread; tsk_params params; int device, rc, threaded, ncycles, bins; tsk_ctx *ctx; char error[512]; };

static void *boot_context(void *arg)
{
    struct ctx_boot *b = arg;
    b->rc = tsk_ctx_create(&b->ctx, b->device, &b->params);
    if (b->rc != 0) snprintf(b->error, sizeof(b->error), "%s", tsk_last_error());
    return NULL;
}

/* (re)start a context for histograms of ncycles x bins; the ruler needs no importer parameters */
static void start_context(struct ctx_boot *b, int ncycles, int bins)
{
    tsk_params *p = &b->params;
    if (b->ctx) { tsk_ctx_destroy(b->ctx); b->ctx = NULL; }
    memset(p, 0, sizeof(*p));
    p->abi_version = TSK_ABI_VERSION;
    p->total_cycles = ncycles > 0 ? ncycles : 1; p->threep_start = 0; p->threep_length = p->total_cycles;
    p->index_length = 0; p->balancer_num_positive_samples = 2; p->balancer_num_negative_samples = 4;
    p->dist_sampling_bins = bins > 0 ? bins : 1; p->fair_sampling_hash_space_size = 1;
    p->max_cctr_scan_left_space = p->max_cctr_scan_right_space = 1;
    p->num_samples = 1; p->samples[0].delimiter_pos = -1;
    b->ncycles = ncycles; b->bins = bins; b->rc = 0;
    b->threaded = pthread_create(&b->thread, NULL, boot_context, b) == 0;
    if (!b->threaded) boot_context(b);
}

static tsk_ctx *await_context(struct ctx_boot *b)
{
    if (b->threaded) { pthread_join(b->thread, NULL); b->threaded = 0; }
    if (b->rc != 0) { fprintf(stderr, "%s\n", b->error); return NULL; }
    return b->ctx;
}

/* One job = the nine positional arguments of polyaruler.c:402-418; corrected taginfo to `out`.
 * Returns the reference's exit code (0, 1 cutoffs, 2 ruling, 3 taginfo). */
static int run_job(struct ctx_boot *boot, char *const a[9], FILE *out)
{
    static uint32_t cutoffs[TSK_MAX_CUTOFF_CYCLES];
    uint32_t total_clusters = 0, max_cycles = 0, *recs = NULL, *hist = NULL;
    int ncycles = 0, rc = 2;
    size_t nrecords = 0;
    int16_t *lens = NULL, *meas = NULL;
    const char *tile_id = a[0], *signals_file = a[1], *cutoffs_file = a[2], *taginfo_file = a[5];
    const int min_len = atoi(a[3]), bins = atoi(a[6]);
    const float weight = atof(a[4]), gap = atof(a[7]);
    const char *sigdist_file = a[8];
    tsk_ctx *ctx;

    if (load_cutoffs(cutoffs_file, tile_id, cutoffs, &ncycles) < 0) return 1;
    if (!boot->ctx && !boot->threaded) start_context(boot, ncycles, bins);
    recs = read_sigpack(signals_file, &total_clusters, &max_cycles, &nrecords);
    ctx = await_context(boot);
    if (ctx && (boot->ncycles != ncycles || boot->bins != bins)) {      /* another histogram shape */
        start_context(boot, ncycles, bins);
        ctx = await_context(boot);
    }
    if (!recs || !ctx) goto done;

    lens = malloc(sizeof(int16_t) * (nrecords ? nrecords : 1));
    meas = malloc(sizeof(int16_t) * (total_clusters ? total_clusters : 1));
    hist = calloc((size_t)(ncycles > 0 ? ncycles : 1) * (bins > 0 ? bins : 1), sizeof(uint32_t));
    if (!lens || !meas || !hist) { perror("process_polya_ruling"); goto done; }
    memset(meas, 0xff, sizeof(int16_t) * (total_clusters ? total_clusters : 1));
    if (ncycles > 0 && bins > 0) {
        if (tsk_ruler_hist_reset(ctx, ncycles, bins) != 0 ||
            tsk_ruler_records(ctx, recs, nrecords, (int)max_cycles, cutoffs, ncycles, min_len, weight, bins, gap,
                              lens, 0) != 0 ||
            tsk_ruler_get_hist(ctx, hist) != 0) {
            fprintf(stderr, "%s\n", tsk_last_error());
            goto done;
        }
        for (size_t r = 0; r < nrecords; r++) {
            const uint32_t clusterno = recs[r * (2 + (size_t)max_cycles)];
            if (clusterno >= total_clusters) {     /* truncated header / records of another tile */
                fprintf(stderr, "Cluster %u in %s is beyond the %u clusters of its header.\n",
                        clusterno, signals_file, total_clusters);
                goto done;
            }
            if (lens[r] >= 0) meas[clusterno] = lens[r];
        }
    }
    {
        const uint32_t h[3] = {4, (uint32_t)ncycles, (uint32_t)bins};
        gzFile f = gzopen(sigdist_file, "wb");
        if (!f) { fprintf(stderr, "Cannot open %s to write.\n", sigdist_file); goto done; }
        if (gzwrite(f, h, sizeof(h)) <= 0 ||
            (ncycles > 0 && bins > 0 && gzwrite(f, hist, (unsigned)(sizeof(uint32_t) * (size_t)ncycles * bins)) <= 0)) {
            gzclose(f);
            goto done;
        }
        gzclose(f);
    }
    rc = rewrite_taginfo(taginfo_file, meas, total_clusters, tile_id, out) < 0 ? 3 : 0;
done:
    free(recs); free(lens); free(meas); free(hist);
    return rc;
}

int main(int argc, char *argv[])
{
    struct ctx_boot *boot = calloc(1, sizeof(*boot));
    const char *dev = getenv("TAILSEEKER_DEVICE");
    int rc = 0;
    if (!boot) return 2;
    if (dev) boot->device = atoi(dev);

    if (argc == 3 && strcmp(argv[1], "--batch") == 0) {
        /* not in the reference: one job per line of argv[2] -- the nine arguments and the file the
         * corrected taginfo goes to, tab separated -- all on one CUDA context */
        char line[8192];
        FILE *jobs = fopen(argv[2], "r");
        if (!jobs) { fprintf(stderr, "Cannot open %s.\n", argv[2]); return 1; }
        while (rc == 0 && fgets(line, sizeof(line), jobs)) {
            char *a[10], *bptr = line, *tok;
            int n = 0;
            while (n < 10 && (tok = strsep(&bptr, "\t\r\n")) != NULL) if (*tok) a[n++] = tok;
            if (n == 0) continue;
            if (n != 10) { fprintf(stderr, "A batch line needs 10 fields.\n"); rc = 1; break; }
            FILE *out = fopen(a[9], "w");
            if (!out) { fprintf(stderr, "Cannot open %s to write.\n", a[9]); rc = 3; break; }
            rc = run_job(boot, a, out);
            if (fclose(out) != 0 && rc == 0) rc = 3;
        }
        fclose(jobs);
    } else if (argc >= 10) {
        rc = run_job(boot, argv + 1, stdout);
        fflush(stdout);
    } else {
        fprintf(stderr, "Usage: %s {tile id} {signals} {cutoffs} {min polya} {downhill ext weight} {taginfo} "
                        "{sampling bin count} {positive sampling gap} {positive sampling output}\n"
                        "       %s --batch {job list}\n", argv[0], argv[0]);
        return 1;
    }
    if (boot->threaded) pthread_join(boot->thread, NULL);
    if (boot->ctx) tsk_ctx_destroy(boot->ctx);
    free(boot);
    return rc;
}
