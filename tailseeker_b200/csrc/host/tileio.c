/*
 * tileio.c -- Illumina per-cycle file readers, kept in the files' native layouts:
 *   CIF  "CIF", u8 version=1, u8 datasize=2, u16 first_cycle, u16 ncycles, u32 N, then four planes
 *        int16[N] (A,C,G,T)              (reference src/importer/cifreader.c:38-102,118-170,235)
 *   BCL  u32 N, u8[N]; ".bcl" or ".bcl.gz" (reference src/importer/bclreader.c:43-80,95-119,206)
 * Blocks are read straight into bcl[T][count] / cif[L3][4][count]: no AoS transpose.
 *   [alternative_calls]  gzip'ed or plain FASTQ, one record per cluster in tile order, overriding the
 *        base-call bytes of cycles first_cycle .. first_cycle+readlength-1; the BCL files of those
 *        cycles are never opened      (reference src/importer/altcalls.c:51-196, bclreader.c:185-191,
 *        tailseq-import.c:203-230)
 * Colour matrix: 16 numbers, inverted like src/contrib/misc.c:11-58 (float cofactors, same order).
 */
#include <endian.h>
#include <fcntl.h>
#include <limits.h>
#include <stdlib.h>
#include <string.h>
#include "tsk_host.h"

struct fastq_override {
    gzFile fp;
    const char *filename;               /* owned by the configuration */
    int first_cycle, ncycles;
};

struct tile_files {
    int T, L3;
    uint32_t nclusters, read;
    FILE **cif;
    gzFile *bcl;                        /* NULL for a cycle taken from a FASTQ override */
    char **bcl_gz_path;                 /* path of the cycle's file when it is a .bcl.gz, else NULL */
    uint8_t **gz_raw;                   /* the whole compressed file, read once (device-side inflate) */
    size_t *gz_len;
    struct fastq_override *alt;         /* in the configuration's list order */
    int nalt;
    int threads;                        /* [options] threads: cycle files read side by side */
};

/* One FASTQ line into buf (BUFSIZ like the reference, which therefore splits longer lines too).
 * Returns the length without the last character -- the reference takes strlen - 1 whether or not
 * that character is a newline (altcalls.c:72,154,177) -- or -1 at the end of the file. */
static int fastq_line(gzFile fp, char *buf)
{
    if (gzgets(fp, buf, BUFSIZ) == NULL) return -1;
    return (int)strlen(buf) - 1;
}

static int open_fastq_override(struct fastq_override *o, const struct host_altcall *ac)
{
    char line[BUFSIZ];
    o->filename = ac->filename;
    o->first_cycle = ac->first_cycle;
    o->fp = gzopen(ac->filename, "rt");
    if (!o->fp) { fprintf(stderr, "open_alternative_calls: Can't open file %s.\n", ac->filename); return -1; }
    /* the read length is that of the first record */
    if (fastq_line(o->fp, line) < 0) { fprintf(stderr, "No sequence available in %s.\n", ac->filename); return -1; }
    if (line[0] != '@') { fprintf(stderr, "%s is not in FASTQ format.\n", ac->filename); return -1; }
    if ((o->ncycles = fastq_line(o->fp, line)) < 0) {
        fprintf(stderr, "Unexpected EOF while reading %s\n", ac->filename);
        return -1;
    }
    if (gzrewind(o->fp) < 0) { fprintf(stderr, "Failed to rewind the FASTQ reader for %s\n", ac->filename); return -1; }
    return 0;
}

/* base-call byte of one FASTQ position: C 1, G 2, T/U 3, everything else (A, N, ...) 0 in bits 0-1,
 * phred in bits 2-7 truncated to a byte (altcalls.c:35-44,186-188) */
static inline uint8_t fastq_call(char base, char qual)
{
    uint8_t code = 0;
    switch (base) {
    case 'C': case 'c': code = 1; break;
    case 'G': case 'g': code = 2; break;
    case 'T': case 't': case 'U': case 'u': code = 3; break;
    }
    return (uint8_t)(code | (uint8_t)((uint8_t)(qual - 33) << 2));
}

/* the next `count` records of one override into rows first_cycle.. of bcl[T][count] */
static int read_fastq_override(struct fastq_override *o, uint32_t count, uint8_t *bcl)
{
    char head[BUFSIZ], seq[BUFSIZ], qual[BUFSIZ];
    for (uint32_t k = 0; k < count; k++) {
        int n;
        if (fastq_line(o->fp, head) < 0) { fprintf(stderr, "Not enough sequences from %s\n", o->filename); return -1; }
        if (head[0] != '@') { fprintf(stderr, "%s is not in FASTQ format.\n", o->filename); return -1; }
        if ((n = fastq_line(o->fp, seq)) < 0) goto eof;
        if (n != o->ncycles) goto ragged;
        if (fastq_line(o->fp, head) < 0) goto eof;
        if (head[0] != '+') { fprintf(stderr, "%s is not in FASTQ format.\n", o->filename); return -1; }
        if ((n = fastq_line(o->fp, qual)) < 0) goto eof;
        if (n != o->ncycles) goto ragged;
        for (int j = 0; j < n; j++)
            bcl[(size_t)(o->first_cycle + j) * count + k] = fastq_call(seq[j], qual[j]);
    }
    return 0;
eof:
    fprintf(stderr, "Unexpected EOF while reading %s\n", o->filename);
    return -1;
ragged:
    fprintf(stderr, "Sequence length in the FASTQ is inconsistent.\n");
    return -1;
}

static float cofactor_entry(int i, int j, const float *m)
{
    static const signed char term[6][6] = {
        {+1, -1, 0, 0, -1, +1}, {+1, +1, 0, -1, -1, 0}, {-1, -1, +1, 0, 0, +1},
        {-1, -1, 0, 0, +1, +1}, {-1, +1, 0, -1, +1, 0}, {+1, -1, -1, 0, 0, +1},
    };
    const int o = 2 + (j - i), ci = i + 4 + o, rj = j + 4 - o;
    float acc = 0.f;
    for (int t = 0; t < 6; t++) {
        const float prod = m[((rj + term[t][1]) % 4) * 4 + ((ci + term[t][0]) % 4)] *
                           m[((rj + term[t][3]) % 4) * 4 + ((ci + term[t][2]) % 4)] *
                           m[((rj + term[t][5]) % 4) * 4 + ((ci + term[t][4]) % 4)];
        if (t == 0) acc = prod;
        else if (t < 3) acc = acc + prod;
        else acc = acc - prod;
    }
    return (o % 2) ? acc : -acc;
}

int tsk_invert_matrix(const float m[16], float out[16])
{
    float adj[16], det = 0.f;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            adj[j * 4 + i] = cofactor_entry(i, j, m);
    for (int k = 0; k < 4; k++) det += m[k] * adj[k * 4];
    if (det == 0) return -1;
    det = 1. / det;
    for (int i = 0; i < 16; i++) out[i] = adj[i] * det;
    return 0;
}

int tsk_load_color_matrix(float inv[16], const char *filename)
{
    float m[16];
    FILE *fp = fopen(filename, "rt");
    if (!fp) { fprintf(stderr, "Color matrix <%s> could not be opened.\n", filename); return -1; }
    for (int i = 0; i < 16; i++)
        if (fscanf(fp, "%f", &m[i]) < 1) {
            fclose(fp);
            fprintf(stderr, "Color matrix <%s> could not be loaded.\n", filename);
            return -1;
        }
    fclose(fp);
    if (tsk_invert_matrix(m, inv) < 0) { fprintf(stderr, "Color matrix could not be inverted.\n"); return -1; }
    return 0;
}

static FILE *open_cif(const char *path, uint32_t *nclusters)
{
    unsigned char h[13];
    FILE *fp = fopen(path, "rb");
    if (!fp) { fprintf(stderr, "open_cif_file: Can't open file %s.\n", path); return NULL; }
    if (fread(h, 1, 13, fp) != 13) { fprintf(stderr, "Unexpected EOF %s.\n", path); fclose(fp); return NULL; }
    if (memcmp(h, "CIF", 3) != 0) { fprintf(stderr, "File %s does not seem to be a CIF file.", path); fclose(fp); return NULL; }
    if (h[3] != 1) { fprintf(stderr, "Unsupported CIF version: %d.\n", h[3]); fclose(fp); return NULL; }
    if (h[4] != 2) { fprintf(stderr, "Unsupported data size (%d).\n", h[4]); fclose(fp); return NULL; }
    *nclusters = (uint32_t)h[9] | ((uint32_t)h[10] << 8) | ((uint32_t)h[11] << 16) | ((uint32_t)h[12] << 24);
    return fp;
}

struct tile_files *tsk_open_tile(const struct host_config *cfg, const char *msgprefix)
{
    char path[PATH_MAX];
    struct tile_files *tf = calloc(1, sizeof(*tf));
    if (!tf) return NULL;
    tf->T = cfg->total_cycles; tf->L3 = cfg->threep_length;
    tf->threads = cfg->threads > 0 ? cfg->threads : 1;
    tf->cif = calloc(tf->L3, sizeof(FILE *));
    tf->bcl = calloc(tf->T, sizeof(gzFile));
    tf->bcl_gz_path = calloc(tf->T, sizeof(char *));
    tf->gz_raw = calloc(tf->T, sizeof(uint8_t *));
    tf->gz_len = calloc(tf->T, sizeof(size_t));
    unsigned char *overridden = calloc(tf->T > 0 ? tf->T : 1, 1);
    for (const struct host_altcall *ac = cfg->altcalls; ac; ac = ac->next) tf->nalt++;
    tf->alt = calloc(tf->nalt ? tf->nalt : 1, sizeof(*tf->alt));
    if (!tf->cif || !tf->bcl || !tf->bcl_gz_path || !tf->gz_raw || !tf->gz_len || !overridden || !tf->alt) { perror("tsk_open_tile"); free(overridden); tsk_close_tile(tf); return NULL; }
    {
        int a = 0;
        for (const struct host_altcall *ac = cfg->altcalls; ac; ac = ac->next, a++) {
            struct fastq_override *o = &tf->alt[a];
            if (open_fastq_override(o, ac) < 0) { free(overridden); tsk_close_tile(tf); return NULL; }
            printf("%sUsing FASTQ %s for cycles %d-%d.\n", msgprefix, o->filename, o->first_cycle + 1,
                   o->first_cycle + o->ncycles);
            /* the reference writes past its cycle array here (tailseq-import.c:212-214) */
            if (o->first_cycle + o->ncycles > tf->T) {
                fprintf(stderr, "Alternative calls in %s run past the last cycle (%d).\n", o->filename, tf->T);
                free(overridden); tsk_close_tile(tf);
                return NULL;
            }
            memset(overridden + o->first_cycle, 1, o->ncycles);
        }
    }
    printf("%sUsing cluster intensities for cycle %d-%d from %s.\n", msgprefix, cfg->threep_start + 1,
           cfg->threep_start + cfg->threep_length, cfg->datadir);
    for (int i = 0; i < tf->L3; i++) {
        uint32_t n = 0;
        snprintf(path, PATH_MAX, "%s/L%03d/C%d.1/s_%d_%04d.cif", cfg->datadir, cfg->lane,
                 cfg->threep_start + i + 1, cfg->lane, cfg->tile);
        tf->cif[i] = open_cif(path, &n);
        if (!tf->cif[i]) { free(overridden); tsk_close_tile(tf); return NULL; }
        /* 8 bytes per cluster and cycle: start the read-ahead now, the readers come later (after the
         * CUDA context is up on the first tile of a process) */
        posix_fadvise(fileno(tf->cif[i]), 0, 0, POSIX_FADV_WILLNEED);
        if (i == 0) tf->nclusters = n;
        else if (n != tf->nclusters) {
            fprintf(stderr, "Inconsistent number of clusters in CIF cycle %d.\n", cfg->threep_length + i + 1);
            free(overridden); tsk_close_tile(tf);
            return NULL;
        }
    }
    /* one message per stretch of cycles that still comes from BCL files (bclreader.c:193-228) */
    for (int c = 0, run = -1; c <= tf->T; c++) {
        uint32_t n;
        if (c == tf->T || overridden[c]) {
            if (run >= 0) printf("%sUsing BCL base calls for cycle %d-%d.\n", msgprefix, run + 1, c);
            run = -1;
            continue;
        }
        snprintf(path, PATH_MAX, "%s/BaseCalls/L%03d/C%d.1/s_%d_%04d.bcl", cfg->datadir, cfg->lane, c + 1,
                 cfg->lane, cfg->tile);
        tf->bcl[c] = gzopen(path, "rb");
        if (!tf->bcl[c]) {
            strncat(path, ".gz", PATH_MAX - strlen(path) - 1);
            tf->bcl[c] = gzopen(path, "rb");
            if (tf->bcl[c]) tf->bcl_gz_path[c] = strdup(path);
        }
        if (!tf->bcl[c]) { fprintf(stderr, "load_bcl_file: Can't open file %s.\n", path); goto bad_bcl; }
        if (gzread(tf->bcl[c], &n, 4) < 4) { fprintf(stderr, "Unexpected EOF %s.\n", path); goto bad_bcl; }
        n = le32toh(n);
        if (n != tf->nclusters) { fprintf(stderr, "Inconsistent number of clusters in cycle %d.\n", c + 1); goto bad_bcl; }
        if (run < 0) run = c;
    }
    free(overridden);
    return tf;
bad_bcl:
    free(overridden);
    tsk_close_tile(tf);
    return NULL;
}

uint32_t tsk_tile_nclusters(const struct tile_files *tf) { return tf->nclusters; }

/* One cycle file is one task: every file has its own handle and its own rows of the block, so the
 * [options] threads read them side by side (a page-cached 1 M-cluster tile is 2.5 GB of copying,
 * the longest host phase of the importer after start-up when it runs on one thread). */
struct read_job {
    struct tile_files *tf;
    uint32_t count;
    uint8_t *bcl;
    int16_t *cif;
    int index;                          /* CIF cycle, BCL cycle, or unused for the overrides */
    tsk_ctx *ctx;                       /* not NULL: hand the rows to the device as soon as they are read */
};

static int put_failed(const char *what, int cycle)
{
    fprintf(stderr, "%s (cycle %d): %s\n", what, cycle + 1, tsk_last_error());
    return -1;
}

static int read_cif_cycle(void *arg)
{
    const struct read_job *j = arg;
    struct tile_files *tf = j->tf;
    const int i = j->index;
    /* on the way to the device the planes are read into a page-locked staging slot of the library
     * (copied at the PCIe rate, and the reader thread is free at once); otherwise into the block array */
    int16_t *planes = j->cif + (size_t)i * 4 * j->count;
    int slot = -1, rc = 0;
    if (j->ctx) {
        void *p = NULL;
        if (tsk_tile_stage_acquire(j->ctx, sizeof(int16_t) * 4 * (size_t)j->count, &p, &slot) != 0)
            return put_failed("tsk_tile_stage_acquire", tf->T - tf->L3 + i);
        planes = p;
    }
    for (int ch = 0; ch < 4 && rc == 0; ch++) {
        const long at = 13 + ((long)tf->nclusters * ch + tf->read) * 2;
        int16_t *dst = planes + (size_t)ch * j->count;
        if (fseek(tf->cif[i], at, SEEK_SET) != 0) {
            fprintf(stderr, "Failed to seek cluster intensity values in a CIF.\n");
            rc = -1;
        } else if (fread(dst, 2, j->count, tf->cif[i]) != j->count) {
            fprintf(stderr, "Not all data were loaded from a CIF.\n");
            rc = -1;
        }
#if __BYTE_ORDER != __LITTLE_ENDIAN
        for (uint32_t k = 0; rc == 0 && k < j->count; k++) dst[k] = (int16_t)le16toh((uint16_t)dst[k]);
#endif
    }
    if (j->ctx) {
        if (rc == 0 && tsk_tile_put_cif(j->ctx, i, planes, j->count) != 0) rc = put_failed("tsk_tile_put_cif", tf->T - tf->L3 + i);
        if (tsk_tile_stage_release(j->ctx, slot) != 0 && rc == 0) rc = put_failed("tsk_tile_stage_release", tf->T - tf->L3 + i);
    }
    return rc;
}

/* a .bcl.gz for the device: the whole compressed file once, then every block names its window */
static int send_bcl_gz(const struct read_job *j, int c)
{
    struct tile_files *tf = j->tf;
    if (!tf->gz_raw[c]) {
        FILE *fp = fopen(tf->bcl_gz_path[c], "rb");
        long size;
        if (!fp || fseek(fp, 0, SEEK_END) != 0 || (size = ftell(fp)) < 0 || fseek(fp, 0, SEEK_SET) != 0) {
            fprintf(stderr, "load_bcl_file: Can't read file %s.\n", tf->bcl_gz_path[c]);
            if (fp) fclose(fp);
            return -1;
        }
        tf->gz_raw[c] = malloc(size > 0 ? (size_t)size : 1);
        if (!tf->gz_raw[c] || fread(tf->gz_raw[c], 1, (size_t)size, fp) != (size_t)size) {
            fprintf(stderr, "Unexpected EOF %s.\n", tf->bcl_gz_path[c]);
            fclose(fp);
            return -1;
        }
        fclose(fp);
        tf->gz_len[c] = (size_t)size;
    }
    if (tsk_tile_put_bcl_gz(j->ctx, c, tf->gz_raw[c], tf->gz_len[c], 4 + (uint64_t)tf->nclusters, 4 + (uint64_t)tf->read) != 0)
        return put_failed("tsk_tile_put_bcl_gz", c);
    return 0;
}

static int read_bcl_cycle(void *arg)
{
    const struct read_job *j = arg;
    const int c = j->index;
    uint8_t *dst = j->bcl + (size_t)c * j->count;
    uint32_t got = 0;
    int slot = -1, rc = 0;
    if (j->ctx && j->tf->bcl_gz_path[c]) return send_bcl_gz(j, c);      /* inflated on the device */
    if (j->ctx) {                                                       /* see read_cif_cycle() */
        void *p = NULL;
        if (tsk_tile_stage_acquire(j->ctx, j->count ? j->count : 1, &p, &slot) != 0) return put_failed("tsk_tile_stage_acquire", c);
        dst = p;
    }
    while (got < j->count) {          /* gzread takes an unsigned int length */
        const uint32_t want = j->count - got > (1u << 30) ? (1u << 30) : j->count - got;
        const int r = gzread(j->tf->bcl[c], dst + got, want);
        if (r < 1) { fprintf(stderr, "Unexpected EOF in a BCL (cycle %d).\n", c + 1); rc = -1; break; }
        got += (uint32_t)r;
    }
    if (j->ctx) {
        if (rc == 0 && tsk_tile_put_bcl(j->ctx, c, dst) != 0) rc = put_failed("tsk_tile_put_bcl", c);
        if (tsk_tile_stage_release(j->ctx, slot) != 0 && rc == 0) rc = put_failed("tsk_tile_stage_release", c);
    }
    return rc;
}

/* the overrides stay on one thread, in list order: where two of them overlap the later entry of
 * the list wins (their cycles have no BCL task, so nothing else writes those rows) */
static int read_overrides(void *arg)
{
    const struct read_job *j = arg;
    for (int a = 0; a < j->tf->nalt; a++)
        if (read_fastq_override(&j->tf->alt[a], j->count, j->bcl) < 0) return -1;
    if (j->ctx)
        for (int a = 0; a < j->tf->nalt; a++)
            for (int c = j->tf->alt[a].first_cycle; c < j->tf->alt[a].first_cycle + j->tf->alt[a].ncycles; c++)
                if (tsk_tile_put_bcl(j->ctx, c, j->bcl + (size_t)c * j->count) != 0) return put_failed("tsk_tile_put_bcl", c);
    return 0;
}

static int read_block(struct tile_files *tf, uint32_t count, uint8_t *bcl, int16_t *cif, tsk_ctx *ctx);

int tsk_read_block(struct tile_files *tf, uint32_t count, uint8_t *bcl, int16_t *cif)
{
    return read_block(tf, count, bcl, cif, NULL);
}

/* The same block straight to the device: every cycle file's rows are sent as soon as its reader
 * thread has them (tsk_tile_put_*), compressed base-call files travel compressed, and the tile is
 * complete (waiting for tsk_tile_process) when this returns.  The block arrays are NOT filled on the
 * host (the rows pass through the library's page-locked staging slots); only the rows of FASTQ
 * overrides are written to `bcl`. */
int tsk_read_block_to_device(struct tile_files *tf, uint32_t count, uint8_t *bcl, int16_t *cif, tsk_ctx *ctx,
                             const float invmatrix[16])
{
    if (tsk_tile_begin(ctx, count, tf->read, invmatrix) != 0) { fprintf(stderr, "%s\n", tsk_last_error()); return -1; }
    const int rc = read_block(tf, count, bcl, cif, ctx);
    if (tsk_tile_end(ctx) != 0) { fprintf(stderr, "%s\n", tsk_last_error()); return -1; }
    return rc;
}

static int read_block(struct tile_files *tf, uint32_t count, uint8_t *bcl, int16_t *cif, tsk_ctx *ctx)
{
    const int maxtasks = tf->L3 + tf->T + 1;
    struct read_job *jobs = calloc(maxtasks, sizeof(*jobs));
    struct tsk_task *tasks = calloc(maxtasks, sizeof(*tasks));
    int n = 0, rc;
    if (!jobs || !tasks) { perror("tsk_read_block"); free(jobs); free(tasks); return -1; }
#define ADD_TASK(function, idx) do { jobs[n] = (struct read_job){tf, count, bcl, cif, (idx), ctx}; \
                                     tasks[n].fn = (function); tasks[n].arg = &jobs[n]; n++; } while (0)
    if (tf->nalt > 0) ADD_TASK(read_overrides, 0);          /* the slowest task first */
    for (int i = 0; i < tf->L3; i++) ADD_TASK(read_cif_cycle, i);
    for (int c = 0; c < tf->T; c++) if (tf->bcl[c]) ADD_TASK(read_bcl_cycle, c);
#undef ADD_TASK
    rc = tsk_run_tasks(tasks, n, tf->threads);
    free(jobs); free(tasks);
    if (rc != 0) return -1;
    tf->read += count;
    return 0;
}

/* after the last block: a FASTQ override must be used up (altcalls.c:116-124, tailseq-import.c:648) */
int tsk_check_tile_end(struct tile_files *tf)
{
    char line[BUFSIZ];
    for (int a = 0; a < tf->nalt; a++)
        if (gzgets(tf->alt[a].fp, line, BUFSIZ) != NULL) {
            fprintf(stderr, "Extra sequences found in %s\n", tf->alt[a].filename);
            return -1;
        }
    return 0;
}

void tsk_close_tile(struct tile_files *tf)
{
    if (!tf) return;
    if (tf->alt) for (int a = 0; a < tf->nalt; a++) if (tf->alt[a].fp) gzclose(tf->alt[a].fp);
    free(tf->alt);
    if (tf->cif) for (int i = 0; i < tf->L3; i++) if (tf->cif[i]) fclose(tf->cif[i]);
    if (tf->bcl) for (int c = 0; c < tf->T; c++) if (tf->bcl[c]) gzclose(tf->bcl[c]);
    if (tf->bcl_gz_path) for (int c = 0; c < tf->T; c++) free(tf->bcl_gz_path[c]);
    if (tf->gz_raw) for (int c = 0; c < tf->T; c++) free(tf->gz_raw[c]);
    free(tf->bcl_gz_path); free(tf->gz_raw); free(tf->gz_len);
    free(tf->cif); free(tf->bcl); free(tf);
}
