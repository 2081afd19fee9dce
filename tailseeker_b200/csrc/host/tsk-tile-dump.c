/*
 * tsk-tile-dump -- writes the arrays tailseq-import uploads for a tile, block after block:
 *     tsk-tile-dump {config.ini} {bcl.out} [cif.out]
 * bcl.out holds, for every read block, u32 count then u8[T][count]; cif.out i16[L3][4][count].
 * Same configuration parser, readers and block split as the importer (tailseq-import.c here,
 * src/importer/tailseq-import.c:597-631 in the reference), no GPU involved: a maintainer's view of
 * the input side (.bcl / .bcl.gz / [alternative_calls] FASTQ overrides), and what the CPU tests
 * check against tailseeker_b200/altcalls.py.
 */
#include <stdlib.h>
#include <string.h>
#include "tsk_host.h"

int main(int argc, char *argv[])
{
    struct host_config *cfg;
    struct tile_files *tf;
    FILE *fb, *fc = NULL;
    uint8_t *bcl;
    int16_t *cif;
    uint32_t nclusters, togo, blocksize;
    char msgprefix[BUFSIZ];
    int rc = 1;

    if (argc < 3) { fprintf(stderr, "Usage: %s {config.ini} {bcl.out} [cif.out]\n", argv[0]); return 2; }
    cfg = tsk_parse_config(argv[1]);
    if (!cfg) return 1;
    snprintf(msgprefix, BUFSIZ, "[%s%d] ", cfg->laneid, cfg->tile);
    tf = tsk_open_tile(cfg, msgprefix);
    if (!tf) return 1;
    nclusters = togo = tsk_tile_nclusters(tf);
    blocksize = cfg->read_buffer_entry_count > 0 ? (uint32_t)cfg->read_buffer_entry_count : 1;
    if (blocksize > nclusters && nclusters > 0) blocksize = nclusters;
    bcl = malloc((size_t)cfg->total_cycles * blocksize + 1);
    cif = malloc(sizeof(int16_t) * 4 * (size_t)cfg->threep_length * blocksize + 1);
    fb = fopen(argv[2], "wb");
    if (argc > 3) fc = fopen(argv[3], "wb");
    if (!bcl || !cif || !fb || (argc > 3 && !fc)) { perror("tsk-tile-dump"); return 1; }
    printf("%sProcessing %u clusters.\n", msgprefix, nclusters);
    while (togo > 0) {
        const uint32_t count = togo >= blocksize ? blocksize : togo;
        if (tsk_read_block(tf, count, bcl, cif) < 0) goto done;
        fwrite(&count, 4, 1, fb);
        fwrite(bcl, 1, (size_t)cfg->total_cycles * count, fb);
        if (fc) fwrite(cif, sizeof(int16_t), 4 * (size_t)cfg->threep_length * count, fc);
        togo -= count;
    }
    if (tsk_check_tile_end(tf) < 0) goto done;
    rc = 0;
done:
    fclose(fb);
    if (fc) fclose(fc);
    free(bcl); free(cif);
    tsk_close_tile(tf);
    tsk_free_config(cfg);
    return rc;
}
