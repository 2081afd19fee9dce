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
    tsk_ctx *ctx;
    char error[512];
};

static void *boot_context(void *arg)
{
    struct ctx_boot *b = arg;
    b->rc = tsk_ctx_create(&b->ctx, b->device, b->params);
    if (b->rc != 0) snprintf(b->error, sizeof(b->error), "%s", tsk_last_error());
    return NULL;
}

static tsk_ctx *await_context(struct ctx_boot *b)
{
    if (b->started) { pthread_join(b->thread, NULL); b->started = 0; }
    if (b->rc != 0) { fprintf(stderr, "%s\n", b->error); return NULL; }
    return b->ctx;
}

static int process(struct host_config *cfg, struct ctx_boot *boot, const float *invmatrix)
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
        if (tsk_read_block(tf, count, bcl, cif) < 0) goto done;
        phase("read CIF/BCL");
        if (!ctx && !(ctx = await_context(boot))) goto done;
        phase("wait for the CUDA context");

        printf("%sAnalyzing and writing out\n", msgprefix);
        if (device_encode && !encoder_ready) {
            tsk_encode_layout *layout = malloc(sizeof(*layout));
            if (!layout || tsk_build_encode_layout(cfg, layout) != 0 || tsk_encode_configure(ctx, layout) != 0) {
                fprintf(stderr, "%s\n", layout ? tsk_last_error() : "out of memory");
                free(layout);
                goto done;
            }
            free(layout);
            encoder_ready = 1;
        }
        if (tsk_tile_upload(ctx, bcl, count, cif, count, invmatrix, count, nclusters - togo, 0) != 0 ||
            tsk_tile_process(ctx) != 0 || (!device_encode && tsk_tile_get_calls(ctx, calls) != 0)) {
            fprintf(stderr, "%s\n", tsk_last_error());
            goto done;
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
static int run_tile(const char *inifile, int device)
{
    struct host_config *cfg;
    tsk_params *params;
    tsk_ctx *ctx = NULL;
    struct ctx_boot boot;
    float inv[16];
    int r;

    cfg = tsk_parse_config(inifile);
    if (cfg == NULL) return -1;
    if (tsk_load_color_matrix(inv, cfg->threep_colormatrix_filename) < 0) { tsk_free_config(cfg); return -1; }
    params = malloc(sizeof(*params));
    if (!params || tsk_build_params(cfg, params) < 0) { free(params); tsk_free_config(cfg); return -1; }
    memset(&boot, 0, sizeof(boot));
    boot.params = params; boot.device = device;
    if (pthread_create(&boot.thread, NULL, boot_context, &boot) == 0) boot.started = 1;
    else boot_context(&boot);
    r = process(cfg, &boot, inv);
    phase("close files");
    ctx = await_context(&boot);
    if (ctx) tsk_ctx_destroy(ctx);
    free(params);
    tsk_free_config(cfg);
    return r;
}

/* `tailseq-import {config.ini}` as in the reference.  Further configuration files are processed one
 * after the other in the same process, so a lane's tiles pay for the CUDA start-up (0.4-2 s) once
 * instead of once per tile; the run stops at the first tile that fails, like a chain of commands. */
int main(int argc, char *argv[])
{
    int device = 0;
    const char *dev = getenv("TAILSEEKER_DEVICE");

    if (argc < 2) { usage(argv[0]); return 0; }
    phase("start");
    if (dev) device = atoi(dev);
    for (int i = 1; i < argc; i++) {
        const int r = run_tile(argv[i], device);
        if (r != 0) return r;
    }
    return 0;
}
