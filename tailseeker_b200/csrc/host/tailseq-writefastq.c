/*
 * tailseq-writefastq (B200 build) -- drop-in for the reference program of the same name
 * (src/exporter/tailseq-writefastq.c:43-420; called by tailseeker/main.py:291-307):
 *
 *   tailseq-writefastq --taginfo {taginfo.txt.gz} --seqqual 'seqqual/{sample}_@tile@.txt.gz'
 *                      --fastq5 {R5.fastq.gz} --fastq3 {R3.fastq.gz} [--fastq-id-verbose] [--threads N]
 *
 * Joins the (deduplicated, tile-ordered) taginfo lines `tile cluster flags polyA mods duplicates`
 * with the per-tile seqqual lines `cluster seq5 qual5 seq3 qual3` and writes the two FASTQ files.
 * The join is host work (two text streams inflated and matched line by line, no arithmetic); the
 * output side -- where the reference hands `--threads` to htslib's BGZF writer -- is the GPU's BGZF
 * deflater of the importer (tsk_bgzf_deflate, csrc/tsk_encode.cu): the FASTQ text collected for both
 * mates is deflated on the device and the program appends the blocks.  TSK_HOST_WRITERS=1 keeps the
 * host path (1 MiB pieces deflated by all threads as gzip members, writers.c) for comparison and
 * for machines without a GPU.  Same options, same record layout, same diagnostics, same exit
 * status; the files decompress to the same bytes.
 */
#include <getopt.h>
#include <stdlib.h>
#include <string.h>
#include "tsk_host.h"

#define LINE_MAX_BYTES 8192            /* the reference's LINE_BUFFER_SIZE: longer lines are split there too */
#define NAME_MAX_BYTES 255
#define PIECE_BYTES    (1u << 20)

struct tag {
    char tile[64], mods[51];
    unsigned long cluster;
    long flags, polya, duplicates;
};

/* taginfo line -> fields (parse_taginfo_line, tailseq-writefastq.c:56-108); numbers as strtol reads them */
static int parse_tag(struct tag *t, const char *line)
{
    const char *p = strchr(line, '\t'), *q;
    char *end;
    if (!p || (size_t)(p - line) >= sizeof(t->tile)) return -1;
    memcpy(t->tile, line, p - line);
    t->tile[p - line] = '\0';
    t->cluster = strtoul(p + 1, &end, 10);
    t->flags = strtol(end + (*end != '\0'), &end, 10);
    t->polya = strtol(end + (*end != '\0'), &end, 10);
    p = end + (*end != '\0');
    q = strchr(p, '\t');
    if (!q || (size_t)(q - p) >= sizeof(t->mods)) return -1;
    memcpy(t->mods, p, q - p);
    t->mods[q - p] = '\0';
    t->duplicates = strtol(q + 1, &end, 10);
    return 0;
}

/* ---- output: FASTQ text collected per file, deflated piecewise by all threads ---- */
struct outfile {
    FILE *fp;
    char *text;
    size_t used, cap;
};

struct piece { const char *src; size_t n; unsigned char *gz; size_t gzn; };

static int deflate_piece(void *arg)
{
    struct piece *pc = arg;
    return tsk_gz_member(pc->src, pc->n, &pc->gz, &pc->gzn);
}

static tsk_bgzf *gpu_deflater = NULL;       /* NULL: host zlib (TSK_HOST_WRITERS=1) */

static int flush_outputs_gpu(struct outfile *out, int nout)
{
    for (int f = 0; f < nout; f++) {
        const uint8_t *z;
        size_t zn;
        if (out[f].used == 0) continue;
        if (tsk_bgzf_deflate(gpu_deflater, out[f].text, out[f].used, &z, &zn) != 0) { fprintf(stderr, "%s\n", tsk_last_error()); return -1; }
        if (fwrite(z, 1, zn, out[f].fp) != zn) return -1;
        out[f].used = 0;
    }
    return 0;
}

static int flush_outputs(struct outfile *out, int nout, int threads)
{
    if (gpu_deflater) return flush_outputs_gpu(out, nout);
    struct piece *pcs;
    struct tsk_task *tasks;
    int n = 0, rc = 0;
    if (nout <= 0) return 0;
    for (int f = 0; f < nout; f++) n += (int)((out[f].used + PIECE_BYTES - 1) / PIECE_BYTES);
    if (n == 0) return 0;
    pcs = calloc(n, sizeof(*pcs));
    tasks = calloc(n, sizeof(*tasks));
    if (!pcs || !tasks) { perror("tailseq-writefastq"); free(pcs); free(tasks); return -1; }
    n = 0;
    for (int f = 0; f < nout; f++)
        for (size_t at = 0; at < out[f].used; at += PIECE_BYTES, n++) {
            pcs[n].src = out[f].text + at;
            pcs[n].n = out[f].used - at < PIECE_BYTES ? out[f].used - at : PIECE_BYTES;
            tasks[n].fn = deflate_piece; tasks[n].arg = &pcs[n];
        }
    if (tsk_run_tasks(tasks, n, threads) != 0) rc = -1;
    n = 0;
    for (int f = 0; f < nout; f++) {
        for (size_t at = 0; at < out[f].used; at += PIECE_BYTES, n++) {
            if (rc == 0 && fwrite(pcs[n].gz, 1, pcs[n].gzn, out[f].fp) != pcs[n].gzn) rc = -1;
            free(pcs[n].gz);
        }
        out[f].used = 0;
    }
    free(pcs); free(tasks);
    return rc;
}

static inline char *put(char *dst, const char *src, size_t n) { memcpy(dst, src, n); return dst + n; }

/* one FASTQ record: @name, bases, +name, qualities */
static void put_record(struct outfile *o, const char *name, int namelen, const char *seq, const char *qual, int len)
{
    char *p = o->text + o->used;
    *p++ = '@'; p = put(p, name, namelen); *p++ = '\n';
    p = put(p, seq, len); *p++ = '\n';
    *p++ = '+'; p = put(p, name, namelen); *p++ = '\n';
    p = put(p, qual, len); *p++ = '\n';
    o->used = p - o->text;
}

/* write_fastq_entry (tailseq-writefastq.c:148-240): both mates of one cluster */
static int put_entry(struct outfile *out, const struct tag *t, const char *sq, int verbose)
{
    char name[NAME_MAX_BYTES + 1];
    const char *f[5];
    int namelen, len5, len3;
    if (verbose)
        namelen = snprintf(name, NAME_MAX_BYTES, "%s:%08u:%04x:%d:%d:%s", t->tile, (unsigned)t->cluster,
                           (unsigned)(int)t->flags, (int)t->duplicates, (int)t->polya, t->mods);
    else
        namelen = snprintf(name, NAME_MAX_BYTES, "%s:%08u", t->tile, (unsigned)t->cluster);
    if (namelen >= NAME_MAX_BYTES) namelen = NAME_MAX_BYTES - 1;
    /* cluster \t seq5 \t qual5 \t seq3 \t qual3 \n */
    f[0] = sq;
    for (int i = 1; i < 5; i++) {
        const char *tab = strchr(f[i - 1], '\t');
        if (!tab) return -1;
        f[i] = tab + 1;
    }
    const char *eol = strchr(f[4], '\n');
    if (!eol) return -1;
    len5 = (int)(f[2] - f[1]) - 1;
    len3 = (int)(f[4] - f[3]) - 1;
    if (len5 != (int)(f[3] - f[2]) - 1 || len3 != (int)(eol - f[4])) return -1;
    put_record(&out[0], name, namelen, f[1], f[2], len5);
    put_record(&out[1], name, namelen, f[3], f[4], len3);
    return 0;
}

static int convert(gzFile taginfo, const char *seqqual_format, struct outfile *out, int verbose, int threads)
{
    char tile_now[64] = "", line[LINE_MAX_BYTES], sq[LINE_MAX_BYTES];
    gzFile seqqual = NULL;
    struct tag t;
    const size_t high_water = (size_t)PIECE_BYTES * (threads > 1 ? threads : 1);
    int rc = -1, errnum;

    while (gzgets(taginfo, line, LINE_MAX_BYTES) != NULL) {
        long have;
        if (parse_tag(&t, line) < 0) { fprintf(stderr, "Failed to parse a line: %s", line); goto done; }
        if (tile_now[0] == '\0' || strcmp(tile_now, t.tile) != 0) {       /* next tile: next seqqual file */
            char *fn = tsk_replace_placeholder(seqqual_format, "@tile@", t.tile);
            strcpy(tile_now, t.tile);
            if (seqqual) gzclose(seqqual);
            seqqual = fn ? gzopen(fn, "rt") : NULL;
            if (!seqqual) { fprintf(stderr, "Failed to open %s.\n", fn ? fn : seqqual_format); free(fn); goto done; }
            gzbuffer(seqqual, 1 << 20);
            free(fn);
        }
        do {                               /* the taginfo lines are a subset of the seqqual lines, in order */
            char *end;
            if (gzgets(seqqual, sq, LINE_MAX_BYTES) == NULL) have = -1;
            else { have = (int)strtol(sq, &end, 10); if (*end != '\t') have = -1; }
            if (have < 0) { fprintf(stderr, "Unexpected format in a seqqual file.\n"); goto done; }
            if (have > (int)t.cluster) { fprintf(stderr, "The taginfo file must be a subset of seqqual.\n"); goto done; }
        } while (have < (int)t.cluster);
        if (put_entry(out, &t, sq, verbose) < 0) { fprintf(stderr, "Failed to write an entry in %s.\n", tile_now); goto done; }
        if (out[0].used >= high_water || out[1].used >= high_water)
            if (flush_outputs(out, 2, threads) < 0) { fprintf(stderr, "Failed to write an entry in %s.\n", tile_now); goto done; }
    }
    {
        const char *msg = gzerror(taginfo, &errnum);
        if (errnum != 0) { fprintf(stderr, "process_write_fastq: %s\n", msg); goto done; }
    }
    rc = flush_outputs(out, 2, threads);
done:
    if (seqqual) gzclose(seqqual);
    return rc;
}

int main(int argc, char *argv[])
{
    int verbose = 0, threads = 1, r;
    const char *taginfo_fn = NULL, *seqqual_fn = NULL, *fastq_fn[2] = {NULL, NULL};
    struct option long_options[] = {
        {"fastq-id-verbose", no_argument, &verbose, 1},
        {"threads", required_argument, 0, 't'},
        {"taginfo", required_argument, 0, 'i'},
        {"seqqual", required_argument, 0, 's'},
        {"fastq5", required_argument, 0, '5'},
        {"fastq3", required_argument, 0, '3'},
        {0, 0, 0, 0}};
    struct outfile out[2];
    gzFile taginfo;

    for (;;) {
        int idx = 0;
        const int c = getopt_long(argc, argv, "t:i:s:5:3:", long_options, &idx);
        if (c == -1) break;
        switch (c) {
        case 't': threads = atoi(optarg); break;
        case 'i': taginfo_fn = optarg; break;
        case 's': seqqual_fn = optarg; break;
        case '5': fastq_fn[0] = optarg; break;
        case '3': fastq_fn[1] = optarg; break;
        }
    }
    if (!taginfo_fn || !seqqual_fn || !fastq_fn[0] || !fastq_fn[1]) {
        fprintf(stderr, "One or more required arguments are not specified.\n");
        return -1;
    }
    if (!getenv("TSK_HOST_WRITERS")) {
        const char *dev = getenv("TAILSEEKER_DEVICE");
        if (tsk_bgzf_create(&gpu_deflater, dev ? atoi(dev) : 0) != 0) { fprintf(stderr, "%s\n", tsk_last_error()); return -1; }
        if (threads < 16) threads = 16;         /* the device takes 16 MiB of text per stream and call */
    }
    taginfo = gzopen(taginfo_fn, "rt");
    if (!taginfo) { fprintf(stderr, "Failed to open %s.\n", taginfo_fn); return -1; }
    gzbuffer(taginfo, 1 << 20);
    memset(out, 0, sizeof(out));
    for (int f = 0; f < 2; f++) {
        /* room for the high-water mark plus one more record of two maximal lines and two names */
        out[f].cap = (size_t)PIECE_BYTES * (threads > 1 ? threads : 1) + 4 * LINE_MAX_BYTES;
        out[f].text = malloc(out[f].cap);
        out[f].fp = fopen(fastq_fn[f], "wb");
        if (!out[f].fp || !out[f].text) { fprintf(stderr, "Failed to open %s.\n", fastq_fn[f]); return -1; }
    }
    r = convert(taginfo, seqqual_fn, out, verbose, threads);
    gzclose(taginfo);
    for (int f = 0; f < 2; f++) {
        if (gpu_deflater) {                                     /* htslib's end-of-file marker block */
            const uint8_t *eof; size_t n;
            if (tsk_bgzf_eof_block(&eof, &n) == 0) fwrite(eof, 1, n, out[f].fp);
        } else if (out[f].used == 0 && ftell(out[f].fp) == 0) {        /* nothing written: an empty gzip member */
            unsigned char *gz; size_t gzn;
            if (tsk_gz_member("", 0, &gz, &gzn) == 0) { fwrite(gz, 1, gzn, out[f].fp); free(gz); }
        }
        if (fclose(out[f].fp) != 0) r = -1;
        free(out[f].text);
    }
    tsk_bgzf_destroy(gpu_deflater);
    return r;
}
