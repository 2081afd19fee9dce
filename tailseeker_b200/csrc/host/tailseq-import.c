/*
 * tailseq-import (B200) -- drop-in for the reference program of the same name
 * (src/importer/tailseq-import.c:717-757): `tailseq-import {config.ini}`, same INI schema, same
 * input files, same output files.  main() / process() keep the reference's structure; the body of
 * distribute_processing() (tailseq-import.c:516-565) is one call sequence into the CUDA library.
 */
#include <stdlib.h>
#include <string.h>
#include <cuda_runtime_api.h>
#include "tsk_host.h"

static void usage(const char *prog)
{
    printf("tailseq-import 3.0 (B200)"
           "\n - Import Illumina .cif and .bcl files into TAIL-seq internal formats"
           "\n\nUsage: %s {config.ini}\n\n", prog);
}

static int process(struct host_config *cfg, tsk_ctx *ctx, const float *invmatrix)
{
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

    /* pinned host buffers: the upload is an asynchronous DMA from these */
    if (cudaMallocHost((void **)&bcl, (size_t)T * blocksize) != cudaSuccess ||
        cudaMallocHost((void **)&cif, sizeof(int16_t) * 4 * (size_t)L3 * blocksize) != cudaSuccess) {
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

    for (blockno = 0; togo > 0; blockno++) {
        const uint32_t count = togo >= blocksize ? blocksize : togo;
        snprintf(msgprefix, BUFSIZ, "[%s%d#%d/%d] ", cfg->laneid, cfg->tile, blockno + 1, totalblocks);
        printf("%sLoading CIF and BCL files\n", msgprefix);
        if (tsk_read_block(tf, count, bcl, cif) < 0) goto done;

        printf("%sAnalyzing and writing out\n", msgprefix);
        if (tsk_tile_upload(ctx, bcl, count, cif, count, invmatrix, count, nclusters - togo, 0) != 0 ||
            tsk_tile_process(ctx) != 0 || tsk_tile_get_calls(ctx, calls) != 0) {
            fprintf(stderr, "%s\n", tsk_last_error());
            goto done;
        }
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
        if (tsk_write_block_outputs(cfg, calls, count, nclusters - togo, bcl, records, nslots) < 0) goto done;
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
    if (cfg->stats_output) {
        if (tsk_get_counters(ctx, counters) != 0) { fprintf(stderr, "%s\n", tsk_last_error()); goto done; }
        if (tsk_write_stats(cfg->stats_output, cfg, counters) < 0) goto done;
    }
    printf("[%s%d] Finished.\n", cfg->laneid, cfg->tile);
    rc = 0;
done:
    if (records) for (int i = 0; i < S; i++) free(records[i]);
    free(records); free(nslots); free(calls); free(pos); free(neg); free(counters);
    if (bcl) cudaFreeHost(bcl);
    if (cif) cudaFreeHost(cif);
    tsk_close_tile(tf);
    tsk_close_writers(cfg);
    return rc;
}

int main(int argc, char *argv[])
{
    struct host_config *cfg;
    tsk_params *params;
    tsk_ctx *ctx = NULL;
    float inv[16];
    int r, device = 0;
    const char *dev = getenv("TAILSEEKER_DEVICE");

    if (argc < 2) { usage(argv[0]); return 0; }
    cfg = tsk_parse_config(argv[1]);
    if (cfg == NULL) return -1;
    if (cfg->altcalls_given) {
        fprintf(stderr, "[alternative_calls] (third-party base-caller override, altcalls.c) is not supported.\n");
        return -1;
    }
    if (tsk_load_color_matrix(inv, cfg->threep_colormatrix_filename) < 0) return -1;
    params = malloc(sizeof(*params));
    if (!params || tsk_build_params(cfg, params) < 0) return -1;
    if (dev) device = atoi(dev);
    if (tsk_ctx_create(&ctx, device, params) != 0) {
        fprintf(stderr, "%s\n", tsk_last_error());
        return -1;
    }
    r = process(cfg, ctx, inv);
    tsk_ctx_destroy(ctx);
    free(params);
    tsk_free_config(cfg);
    return r;
}
