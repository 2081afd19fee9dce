/*
 * tailseq-import (B200) -- drop-in for the reference program of the same name
 * (src/importer/tailseq-import.c:717-757): `tailseq-import {config.ini}`, same INI schema, same
 * input files, same output files.  main() / process() keep the reference's structure; the body of
 * distribute_processing() (tailseq-import.c:516-565) is one call sequence into the CUDA library.
 */
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <time.h>
#include "tsk_host.h"

/* TSK_HOST_TIMING=1: wall time of each phase on stderr */
static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
static int timing_on = -1;
static double t_last;
static void phase(const char *what)
{
    if (timing_on < 0) { timing_on = getenv("TSK_HOST_TIMING") != NULL; t_last = now_s(); }
    if (timing_on) { const double t = now_s(); fprintf(stderr, "[timing] %-28s %.3f s\n", what, t - t_last); t_last = t; }
}

static void usage(const char *prog)
{
    printf("tailseq-import 3.0 (B200)"
           "\n - Import Illumina .cif and .bcl files into TAIL-seq internal formats"
           "\n\nUsage: %s {config.ini} [more config.ini ...]\n\n", prog);
}

/* The CUDA context (0.7 s to create) comes up on its own thread while the main thread opens the
 * output files and reads the first block of CIF / BCL data. */
struct ctx_boot {
    pthread_t thread;
    const tsk_params *params;
    int device, rc, started;
    int threaded;                       /* boot_context() runs on its own thread (set before it starts) */
    tsk_ctx *ctx;
    char error[512];
    /* the thread goes on after the context is up: it sets up the output encoder and page-locks its
     * host buffer (0.3 s) while the main thread uploads and processes the first block */
    pthread_mutex_t lock;
    pthread_cond_t changed;
    int ctx_ready;                      /* tsk_ctx_create() has returned */
    int hint_set;                       /* block_clusters is final (0: no encoder warm-up) */
    uint32_t block_clusters;
    const tsk_encode_layout *layout;    /* NULL: host writers */
    int warm_rc;
    char warm_error[512];
};

/* A process that handles several tiles keeps its context -- device buffers, the encoder's pinned
 * staging buffer -- from one tile to the next when the parameters are the same (they are, within
 * a run): creating and sizing them again costs more than the tile's kernels.  One per host thread. */
static __thread tsk_ctx *kept_ctx = NULL;
static __thread tsk_params *kept_params = NULL;
static __thread int kept_device = -1;

static void drop_kept_context(void)
{
    if (kept_ctx) tsk_ctx_destroy(kept_ctx);
    free(kept_params);
    kept_ctx = NULL; kept_params = NULL; kept_device = -1;
}

static void *boot_context(void *arg)
{
    struct ctx_boot *b = arg;
    uint32_t n;
    b->rc = tsk_ctx_create(&b->ctx, b->device, b->params);
    if (b->rc != 0) snprintf(b->error, sizeof(b->error), "%s", tsk_last_error());
    if (!b->threaded) { b->ctx_ready = 1; return NULL; }       /* called in line: nothing to overlap with */
    pthread_mutex_lock(&b->lock);
    b->ctx_ready = 1;
    pthread_cond_broadcast(&b->changed);
    while (!b->hint_set) pthread_cond_wait(&b->changed, &b->lock);
    n = b->block_clusters;
    pthread_mutex_unlock(&b->lock);
    if (b->rc == 0 && b->layout && n > 0) {
        b->warm_rc = tsk_encode_configure(b->ctx, b->layout);
        if (b->warm_rc == 0) b->warm_rc = tsk_encode_reserve(b->ctx, n);
        if (b->warm_rc != 0) snprintf(b->warm_error, sizeof(b->warm_error), "%s", tsk_last_error());
    }
    return NULL;
}

/* the size of the tile's read blocks is known (or there will be none): lets the start-up thread go on */
static void boot_hint(struct ctx_boot *b, uint32_t block_clusters)
{
    pthread_mutex_lock(&b->lock);
    if (!b->hint_set) { b->block_clusters = block_clusters; b->hint_set = 1; }
    pthread_cond_broadcast(&b->changed);
    pthread_mutex_unlock(&b->lock);
}

/* the context, as soon as it exists (the start-up thread may still be warming the encoder up) */
static tsk_ctx *await_context(struct ctx_boot *b)
{
    if (b->started) {
        pthread_mutex_lock(&b->lock);
        while (!b->ctx_ready) pthread_cond_wait(&b->changed, &b->lock);
        pthread_mutex_unlock(&b->lock);
    }
    if (b->rc != 0) { fprintf(stderr, "%s\n", b->error); return NULL; }
    return b->ctx;
}

/* the end of the start-up thread: before the encoder of its context is used, and before leaving */
static int finish_boot(struct ctx_boot *b)
{
    if (b->started) { boot_hint(b, 0); pthread_join(b->thread, NULL); b->started = 0; }
    if (b->warm_rc != 0) { fprintf(stderr, "%s\n", b->warm_error); b->warm_rc = 0; return -1; }
    return 0;
}

/* --devices: one worker (host thread + CUDA context) per GPU, the configuration files dealt round
 * robin (one tile set per GPU); every device sums the histograms and counters of its tiles in a
 * tsk_merge accumulator and the sums are all-reduced over NCCL at the end of the run. */
struct run_merge {
    tsk_merge *m;
    int cycles, bins, nsamples;
    char **sample_names;         /* list order, taken from the first tile */
    char **sample_rows;          /* the constant CSV columns of each sample */
};

static int merge_tile(struct run_merge *rm, const struct host_config *cfg, tsk_ctx *ctx, int device, int per_block)
{
    if (!rm) return 0;
    if (!rm->m) {
        if (tsk_merge_create(&rm->m, device, cfg->total_cycles, cfg->bins, cfg->num_samples, 0) != 0) {
            fprintf(stderr, "%s\n", tsk_last_error());
            return -1;
        }
        rm->cycles = cfg->total_cycles; rm->bins = cfg->bins; rm->nsamples = cfg->num_samples;
        rm->sample_names = calloc(cfg->num_samples, sizeof(char *));
        rm->sample_rows = calloc(cfg->num_samples, sizeof(char *));
        for (const struct host_sample *s = cfg->samples; s; s = s->next) {
            char row[512];
            snprintf(row, sizeof(row), "%s,%d,%s,%d,%d", s->index ? s->index : "(null)", s->maximum_index_mismatches,
                     s->delimiter ? s->delimiter : "(null)", s->delimiter_pos, s->maximum_delimiter_mismatches);
            rm->sample_names[s->numindex] = strdup(s->name);
            rm->sample_rows[s->numindex] = strdup(row);
        }
    }
    if ((per_block ? tsk_merge_add_tile(rm->m, ctx) : tsk_merge_add_totals(rm->m, ctx)) != 0) {
        fprintf(stderr, "%s\n", tsk_last_error());
        return -1;
    }
    return 0;
}

static int process(struct host_config *cfg, struct ctx_boot *boot, const float *invmatrix, struct run_merge *rm)
{
    tsk_ctx *ctx = NULL;
    char msgprefix[BUFSIZ];
    struct tile_files *tf = NULL;
    uint8_t *bcl = NULL;
    int16_t *cif = NULL;
    tsk_call *calls = NULL;
    uint32_t **records = NULL, *nslots = NULL, *pos = NULL, *neg = NULL;
    tsk_counters *counters = NULL;
    const int S = cfg->num_samples, T = cfg->total_cycles, L3 = cfg->threep_length;
    uint32_t nclusters, togo, blocksize, blockno, totalblocks;
    int rc = -1;

    /* The block's text and records are formatted and deflated on the GPU (tsk_tile_encode) and the
     * host only appends the BGZF blocks; TSK_HOST_WRITERS=1 keeps the host formatter + zlib, for
     * timing one against the other. */
    const int device_encode = getenv("TSK_HOST_WRITERS") == NULL;
    int encoder_ready = 0;

    tsk_writers_device_encoding(device_encode);
    snprintf(msgprefix, BUFSIZ, "[%s%d] ", cfg->laneid, cfg->tile);
    if (tsk_open_writers(cfg) < 0) return -1;
    printf("%sOpening input sources\n", msgprefix);
    tf = tsk_open_tile(cfg, msgprefix);
    if (!tf) goto done;

    nclusters = togo = tsk_tile_nclusters(tf);
    blocksize = cfg->read_buffer_entry_count > 0 ? (uint32_t)cfg->read_buffer_entry_count : 1;
    if (blocksize > nclusters && nclusters > 0) blocksize = nclusters;
    totalblocks = nclusters / blocksize + ((nclusters % blocksize > 0) ? 1 : 0);
    printf("%sProcessing %u clusters.\n", msgprefix, nclusters);
    boot_hint(boot, blocksize);
    if (tsk_write_signal_headers(cfg, nclusters) < 0) goto done;

    /* plain page-aligned buffers: pinning 2.5 GB for a single pass costs more than the staged copy */
    if (posix_memalign((void **)&bcl, 4096, (size_t)T * blocksize + 4096) != 0 ||
        posix_memalign((void **)&cif, 4096, sizeof(int16_t) * 4 * (size_t)L3 * blocksize + 4096) != 0) {
        fprintf(stderr, "Cannot allocate the read buffers.\n");
        goto done;
    }
    calls = malloc(sizeof(tsk_call) * (size_t)blocksize);
    records = calloc(S, sizeof(*records));
    nslots = calloc(S, sizeof(*nslots));
    pos = malloc(sizeof(uint32_t) * (size_t)T * cfg->bins);
    neg = malloc(sizeof(uint32_t) * (size_t)T * cfg->bins);
    counters = calloc(S, sizeof(*counters));
    if (!calls || !records || !nslots || !pos || !neg || !counters) { perror("process"); goto done; }
    phase("open files + buffers");

    for (blockno = 0; togo > 0; blockno++) {
        const uint32_t count = togo >= blocksize ? blocksize : togo;
        snprintf(msgprefix, BUFSIZ, "[%s%d#%d/%d] ", cfg->laneid, cfg->tile, blockno + 1, totalblocks);
        printf("%sLoading CIF and BCL files\n", msgprefix);
        /* The cycle files go to the device row by row while they are read (through the library's
         * page-locked staging slots) and .bcl.gz files are inflated there.  On the first tile of a
         * process that means waiting for the context first -- the files' read-ahead was started when
         * they were opened -- which still beats reading into pageable arrays meanwhile and uploading
         * those (0.76 s against 1.3 s for the first 1.07 M-cluster tile).  TSK_DIRECT_INGEST=0 keeps
         * the whole-block host arrays (the reference's order of things, and what the host writers need). */
        int direct = 0;
        const char *want_direct = getenv("TSK_DIRECT_INGEST");
        const int never_direct = want_direct && want_direct[0] == '0';
        if (device_encode && !never_direct && !ctx && !(ctx = await_context(boot))) goto done;
        direct = device_encode && !never_direct && ctx != NULL;
        if ((direct ? tsk_read_block_to_device(tf, count, bcl, cif, ctx, invmatrix) : tsk_read_block(tf, count, bcl, cif)) < 0) goto done;
        phase(direct ? "read CIF/BCL -> device" : "read CIF/BCL");
        if (!ctx && !(ctx = await_context(boot))) goto done;
        phase("wait for the CUDA context");

        printf("%sAnalyzing and writing out\n", msgprefix);
        if ((!direct && tsk_tile_upload(ctx, bcl, count, cif, count, invmatrix, count, nclusters - togo, 0) != 0) ||
            tsk_tile_process(ctx) != 0 || (!device_encode && tsk_tile_get_calls(ctx, calls) != 0)) {
            fprintf(stderr, "%s\n", tsk_last_error());
            goto done;
        }
        if (device_encode && !encoder_ready) {
            tsk_encode_layout *layout = calloc(1, sizeof(*layout));
            if (!layout || tsk_build_encode_layout(cfg, layout) != 0) {
                fprintf(stderr, "Cannot describe the output layout.\n");
                free(layout);
                goto done;
            }
            /* the start-up thread of a new context has set the encoder up meanwhile, a context kept from
             * the previous tile is set up already: the call then returns at once */
            if (finish_boot(boot) != 0 || tsk_encode_configure(ctx, layout) != 0) {
                fprintf(stderr, "%s\n", tsk_last_error());
                free(layout);
                goto done;
            }
            free(layout);
            encoder_ready = 1;
        }
        phase("upload + kernels + calls");
        if (device_encode) {
            if (tsk_tile_encode(ctx) != 0) { fprintf(stderr, "%s\n", tsk_last_error()); goto done; }
            phase("format + deflate on the GPU");
            if (tsk_write_encoded_outputs(cfg, ctx) < 0) goto done;
            phase("write");
        } else {
        for (struct host_sample *s = cfg->samples; s; s = s->next) {
            const int i = s->numindex;
            free(records[i]); records[i] = NULL; nslots[i] = 0;
            if (!s->stream_signal) continue;
            if (tsk_tile_get_record_count(ctx, i, &nslots[i]) != 0) { fprintf(stderr, "%s\n", tsk_last_error()); goto done; }
            const size_t words = (size_t)nslots[i] * (2 + (size_t)s->signal_dump_length);
            records[i] = malloc(sizeof(uint32_t) * (words ? words : 1));
            if (!records[i] || tsk_tile_get_records(ctx, i, records[i], words) != 0) {
                fprintf(stderr, "%s\n", tsk_last_error());
                goto done;
            }
        }
        phase("fetch records");
        if (tsk_write_block_outputs(cfg, calls, count, nclusters - togo, bcl, records, nslots) < 0) goto done;
        phase("format + compress + write");
        }
        /* like the reference, the pos/neg distributions are rewritten after every block
         * (tailseq-import.c:358-370,559): a multi-block tile keeps its last block's */
        if (cfg->signal_dists_output) {
            if (tsk_get_hist(ctx, pos, neg) != 0) { fprintf(stderr, "%s\n", tsk_last_error()); goto done; }
            if (tsk_write_sigdists(cfg->signal_dists_output, "pos", pos, T, cfg->bins) ||
                tsk_write_sigdists(cfg->signal_dists_output, "neg", neg, T, cfg->bins)) goto done;
        }
        togo -= count;
    }
    printf("%sClearing\n", msgprefix);
    if (tsk_check_tile_end(tf) < 0) goto done;
    if (!ctx && !(ctx = await_context(boot))) goto done;      /* a tile without clusters */
    if (cfg->stats_output) {
        if (tsk_get_counters(ctx, counters) != 0) { fprintf(stderr, "%s\n", tsk_last_error()); goto done; }
        if (tsk_write_stats(cfg->stats_output, cfg, counters) < 0) goto done;
    }
    /* run-wide sums: what the tile's .sigdists files hold (the LAST block's histograms, like the
     * reference's per-block rewrite, tailseq-import.c:358-370,559) and its counters */
    if (merge_tile(rm, cfg, ctx, boot->device, 1) != 0 || merge_tile(rm, cfg, ctx, boot->device, 0) != 0) goto done;
    phase("sigdists + stats");
    printf("[%s%d] Finished.\n", cfg->laneid, cfg->tile);
    rc = 0;
done:
    if (records) for (int i = 0; i < S; i++) free(records[i]);
    free(records); free(nslots); free(calls); free(pos); free(neg); free(counters);
    free(bcl);
    free(cif);
    tsk_close_tile(tf);
    tsk_close_writers(cfg);
    return rc;
}

/* one tile: the body of the reference's main() (tailseq-import.c:736-757) */
static int run_tile(const char *inifile, int device, struct run_merge *rm)
{
    struct host_config *cfg;
    tsk_params *params;
    tsk_ctx *ctx = NULL;
    struct ctx_boot boot;
    tsk_encode_layout *layout = NULL;
    float inv[16];
    int r;

    cfg = tsk_parse_config(inifile);
    if (cfg == NULL) return -1;
    if (tsk_load_color_matrix(inv, cfg->threep_colormatrix_filename) < 0) { tsk_free_config(cfg); return -1; }
    params = malloc(sizeof(*params));
    if (!params || tsk_build_params(cfg, params) < 0) { free(params); tsk_free_config(cfg); return -1; }
    memset(&boot, 0, sizeof(boot));
    boot.params = params; boot.device = device;
    pthread_mutex_init(&boot.lock, NULL);
    pthread_cond_init(&boot.changed, NULL);
    if (kept_ctx && kept_device == device && memcmp(kept_params, params, sizeof(*params)) == 0) {
        boot.ctx = kept_ctx;                        /* same parameters as the previous tile */
        if (tsk_counters_reset(kept_ctx) != 0) { fprintf(stderr, "%s\n", tsk_last_error()); free(params); tsk_free_config(cfg); return -1; }
    } else {
        drop_kept_context();
        if (getenv("TSK_HOST_WRITERS") == NULL && (layout = calloc(1, sizeof(*layout))) != NULL &&
            tsk_build_encode_layout(cfg, layout) == 0)
            boot.layout = layout;                   /* else: process() reports what is wrong with it */
        boot.threaded = 1;
        if (pthread_create(&boot.thread, NULL, boot_context, &boot) == 0) boot.started = 1;
        else { boot.threaded = 0; boot_context(&boot); }
    }
    r = process(cfg, &boot, inv, rm);
    phase("close files");
    if (finish_boot(&boot) != 0) r = -1;
    ctx = boot.rc == 0 ? boot.ctx : NULL;
    pthread_mutex_destroy(&boot.lock);
    pthread_cond_destroy(&boot.changed);
    free(layout);
    if (ctx && ctx != kept_ctx) {                   /* keep it for the next tile of this thread */
        kept_ctx = ctx; kept_device = device;
        kept_params = params; params = NULL;
    }
    if (r != 0) drop_kept_context();
    free(params);
    tsk_free_config(cfg);
    return r;
}

/* ---- several devices ---- */
struct dev_worker {
    pthread_t thread;
    int device, index, ndev, nfiles, rc;
    char **files;
    struct run_merge rm;
};

static void *device_worker(void *arg)
{
    struct dev_worker *w = arg;
    for (int i = w->index; i < w->nfiles; i += w->ndev) {
        w->rc = run_tile(w->files[i], w->device, &w->rm);
        if (w->rc != 0) break;
    }
    drop_kept_context();
    return NULL;
}

static int write_merged(const struct run_merge *rm, const char *sigdists_fmt, const char *stats_file)
{
    const size_t hw = (size_t)rm->cycles * rm->bins;
    uint64_t *pos = malloc(sizeof(uint64_t) * hw), *neg = malloc(sizeof(uint64_t) * hw);
    uint64_t *cnt = malloc(sizeof(uint64_t) * 6 * (size_t)rm->nsamples);
    uint32_t *h32 = malloc(sizeof(uint32_t) * hw);
    int rc = -1;
    if (!pos || !neg || !cnt || !h32) { perror("write_merged"); goto done; }
    if (tsk_merge_get(rm->m, NULL, pos, neg, NULL, cnt) != 0) { fprintf(stderr, "%s\n", tsk_last_error()); goto done; }
    if (sigdists_fmt) {
        /* the .sigdists layout holds u32 counts (tailseq-import.c:290-296); a run-wide count beyond that saturates */
        for (int k = 0; k < 2; k++) {
            const uint64_t *src = k ? neg : pos;
            for (size_t i = 0; i < hw; i++) h32[i] = src[i] > 0xffffffffull ? 0xffffffffu : (uint32_t)src[i];
            if (tsk_write_sigdists(sigdists_fmt, k ? "neg" : "pos", h32, rm->cycles, rm->bins) != 0) goto done;
        }
    }
    if (stats_file) {
        /* the "(total)" rows of scripts/stats-merge-demultiplexing-counts.py:33-53 (sorted by name) */
        FILE *fp = fopen(stats_file, "w");
        int *order = malloc(sizeof(int) * rm->nsamples);
        if (!fp || !order) { perror("write_merged"); if (fp) fclose(fp); free(order); goto done; }
        for (int i = 0; i < rm->nsamples; i++) order[i] = i;
        for (int i = 1; i < rm->nsamples; i++)
            for (int j = i; j > 0 && strcmp(rm->sample_names[order[j - 1]], rm->sample_names[order[j]]) > 0; j--) {
                const int t = order[j]; order[j] = order[j - 1]; order[j - 1] = t;
            }
        fprintf(fp, "tile,name,index,max_index_mm,delim,delim_pos,max_delim_mm,cln_no_index_mm,cln_1_index_mm,"
                    "cln_2+_index_mm,cln_fp_mm,cln_qc_fail,cln_no_delim,cln_passed\n");
        for (int k = 0; k < rm->nsamples; k++) {
            const int i = order[k];
            const uint64_t *c = cnt + 6 * (size_t)i;      /* mm0 mm1 mm2+ nodelim fpmm qcfail */
            fprintf(fp, "(total),%s,%s,%llu,%llu,%llu,%llu,%llu,%llu,%lld\n", rm->sample_names[i], rm->sample_rows[i],
                    (unsigned long long)c[0], (unsigned long long)c[1], (unsigned long long)c[2], (unsigned long long)c[4],
                    (unsigned long long)c[5], (unsigned long long)c[3],
                    (long long)(c[0] + c[1] + c[2]) - (long long)(c[3] + c[4] + c[5]));
        }
        free(order);
        fclose(fp);
    }
    rc = 0;
done:
    free(pos); free(neg); free(cnt); free(h32);
    return rc;
}

static int run_devices(const char *devlist, char **files, int nfiles, const char *sigdists_fmt, const char *stats_file)
{
    struct dev_worker w[64];
    int devices[64], ndev = 0, rc = 0;
    void *comms[64];
    char *list = strdup(devlist), *save = NULL;
    for (char *tok = strtok_r(list, ",", &save); tok && ndev < 64; tok = strtok_r(NULL, ",", &save)) devices[ndev++] = atoi(tok);
    free(list);
    if (ndev < 1) { fprintf(stderr, "--devices needs a comma-separated list of GPU numbers\n"); return 2; }
    memset(w, 0, sizeof(w));
    for (int d = 0; d < ndev; d++) {
        w[d].device = devices[d]; w[d].index = d; w[d].ndev = ndev; w[d].nfiles = nfiles; w[d].files = files;
        if (pthread_create(&w[d].thread, NULL, device_worker, &w[d]) != 0) { perror("pthread_create"); return -1; }
    }
    for (int d = 0; d < ndev; d++) { pthread_join(w[d].thread, NULL); if (w[d].rc != 0 && rc == 0) rc = w[d].rc; }
    if (rc != 0 || (!sigdists_fmt && !stats_file)) return rc;
    /* the run-wide sums: one all-reduce over the devices that processed tiles, written by the first */
    int nact = 0, act[64];
    for (int d = 0; d < ndev; d++) if (w[d].rm.m) { act[nact] = d; devices[nact] = w[d].device; nact++; }
    if (nact == 0) return 0;
    for (int k = 1; k < nact; k++)
        if (w[act[k]].rm.cycles != w[act[0]].rm.cycles || w[act[k]].rm.bins != w[act[0]].rm.bins ||
            w[act[k]].rm.nsamples != w[act[0]].rm.nsamples) {
            fprintf(stderr, "The tiles of a merged run must share cycles, bins and sample list.\n");
            return -1;
        }
    if (nact > 1) {
        if (tsk_nccl_init_all(comms, nact, devices) != 0 || tsk_nccl_group_start() != 0) { fprintf(stderr, "%s\n", tsk_last_error()); return -1; }
        for (int k = 0; k < nact; k++)
            if (tsk_merge_allreduce(w[act[k]].rm.m, NULL, comms[k]) != 0) { fprintf(stderr, "%s\n", tsk_last_error()); return -1; }
        if (tsk_nccl_group_end() != 0) { fprintf(stderr, "%s\n", tsk_last_error()); return -1; }
    }
    rc = write_merged(&w[act[0]].rm, sigdists_fmt, stats_file);
    if (nact > 1) for (int k = 0; k < nact; k++) tsk_nccl_destroy(comms[k]);
    for (int d = 0; d < ndev; d++) tsk_merge_destroy(w[d].rm.m);
    return rc;
}

/* `tailseq-import {config.ini}` as in the reference.  Further configuration files are processed one
 * after the other in the same process, so a lane's tiles pay for the CUDA start-up (0.4-2 s) once
 * instead of once per tile; the run stops at the first tile that fails, like a chain of commands.
 * `tailseq-import --devices 0,1,...,7 [--merged-sigdists FMT{posneg}] [--merged-stats FILE] a.ini b.ini ...`
 * deals the tiles over several GPUs and writes the run-wide sums next to the per-tile outputs. */
int main(int argc, char *argv[])
{
    int device = 0, first = 1;
    const char *dev = getenv("TAILSEEKER_DEVICE"), *devlist = NULL, *msig = NULL, *mstats = NULL;

    if (argc < 2) { usage(argv[0]); return 0; }
    phase("start");
    if (dev) device = atoi(dev);
    while (first + 1 < argc && argv[first][0] == '-' && argv[first][1] == '-') {
        if (!strcmp(argv[first], "--devices")) devlist = argv[first + 1];
        else if (!strcmp(argv[first], "--merged-sigdists")) msig = argv[first + 1];
        else if (!strcmp(argv[first], "--merged-stats")) mstats = argv[first + 1];
        else { usage(argv[0]); return 2; }
        first += 2;
    }
    if (devlist || msig || mstats) {
        char one[16];
        snprintf(one, sizeof(one), "%d", device);
        return run_devices(devlist ? devlist : one, argv + first, argc - first, msig, mstats);
    }
    for (int i = first; i < argc; i++) {
        const int r = run_tile(argv[i], device, NULL);
        if (r != 0) return r;
    }
    drop_kept_context();
    return 0;
}
