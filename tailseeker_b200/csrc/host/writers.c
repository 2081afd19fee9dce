/*
 * writers.c -- the importer's output files, driven by the per-cluster calls from the GPU:
 *   taginfo / seqqual text   (reference src/importer/spotanalyzer.c:206-352)
 *   .sigpack                 (src/importer/tailseq-import.c:98-119, spotanalyzer.c:389-438)
 *   .sigdists                (tailseq-import.c:268-300)
 *   demultiplexing CSV       (tailseq-import.c:685-714)
 * Streams are gzip (every consumer opens them with gzopen / gzip.open / zcat).  Like the
 * reference's BGZF files they are a series of independent gzip members: each member is
 * compressed on its own, so a block's text chunks and record chunks are deflated by all
 * [options] threads at once and then appended to the files in cluster order.
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include "tsk_host.h"

/* zlib's default level like the reference's bgzf_open(fn, "w"), or the level in TSK_GZ_LEVEL
 * (TSK_SHIM_GZ_LEVEL, the variable of the test harness's BGZF shim, is honoured too so that both
 * programs compress alike when they are timed side by side). */
static int gz_level(void)
{
    static int level = -2;
    if (level == -2) {
        const char *lv = getenv("TSK_GZ_LEVEL");
        if (!lv) lv = getenv("TSK_SHIM_GZ_LEVEL");
        level = (lv && lv[0] >= '0' && lv[0] <= '9' && lv[1] == '\0') ? lv[0] - '0' : Z_DEFAULT_COMPRESSION;
    }
    return level;
}

/* one gzip member holding src[0..n); *out is malloc'ed */
static int gz_member(const void *src, size_t n, unsigned char **out, size_t *outn)
{
    z_stream z;
    memset(&z, 0, sizeof(z));
    if (deflateInit2(&z, gz_level(), Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) return -1;
    const size_t cap = deflateBound(&z, (uLong)n) + 32;
    unsigned char *buf = malloc(cap);
    if (!buf) { deflateEnd(&z); return -1; }
    z.next_in = (Bytef *)src; z.avail_in = (uInt)n;
    z.next_out = buf; z.avail_out = (uInt)cap;
    const int r = deflate(&z, Z_FINISH);
    deflateEnd(&z);
    if (r != Z_STREAM_END) { free(buf); return -1; }
    *out = buf; *outn = cap - z.avail_out;
    return 0;
}
#define GZ_MEMBER_MAX (1u << 26)        /* bytes of input per member (uInt limits of zlib) */

/* ---- a few worker threads over an array of independent tasks ([options] threads) ---- */
struct task { int (*fn)(void *); void *arg; };
struct pool { struct task *tasks; int ntasks, next, failed; };

static void *pool_worker(void *arg)
{
    struct pool *pl = arg;
    for (;;) {
        const int i = __atomic_fetch_add(&pl->next, 1, __ATOMIC_RELAXED);
        if (i >= pl->ntasks) break;
        if (pl->tasks[i].fn(pl->tasks[i].arg) != 0) __atomic_store_n(&pl->failed, 1, __ATOMIC_RELAXED);
    }
    return NULL;
}

static int run_tasks(struct task *tasks, int ntasks, int nthreads)
{
    struct pool pl = {tasks, ntasks, 0, 0};
    pthread_t th[64];
    int started = 0;
    if (nthreads > ntasks) nthreads = ntasks;
    if (nthreads > 64) nthreads = 64;
    for (int i = 1; i < nthreads; i++)
        if (pthread_create(&th[started], NULL, pool_worker, &pl) == 0) started++;
    pool_worker(&pl);                              /* the calling thread works too */
    for (int i = 0; i < started; i++) pthread_join(th[i], NULL);
    return pl.failed ? -1 : 0;
}

static FILE *open_out(const char *fmt, const char *key, const char *val)
{
    char *fn = tsk_replace_placeholder(fmt, key, val);
    FILE *f;
    if (!fn) return NULL;
    f = fopen(fn, "wb");
    if (!f) { perror("open_writers"); fprintf(stderr, "Failed to write to %s\n", fn); }
    free(fn);
    return f;
}

/* src[0..n) as one gzip member appended to f */
static int put_member(FILE *f, const void *src, size_t n)
{
    unsigned char *z = NULL;
    size_t zn = 0;
    if (gz_member(src, n, &z, &zn) != 0) return -1;
    const int ok = fwrite(z, 1, zn, f) == zn;
    free(z);
    return ok ? 0 : -1;
}

/* ---- BGZF framing for the few bytes the host itself adds to device-encoded files ---- */
static int device_encoding = 0;      /* files are series of BGZF blocks made by tsk_tile_encode() */

void tsk_writers_device_encoding(int on) { device_encoding = on; }

/* src[0..n) (n <= 65000) as one BGZF member holding a stored deflate block */
static int put_bgzf_stored(FILE *f, const void *src, size_t n)
{
    unsigned char m[18 + 5 + 65000 + 8];
    static const unsigned char hdr[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
    const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), src, (uInt)n), len = (uint32_t)n;
    const size_t total = 18 + 5 + n + 8;
    if (n > 65000) return -1;
    memcpy(m, hdr, 16);
    m[16] = (unsigned char)((total - 1) & 0xff); m[17] = (unsigned char)((total - 1) >> 8);
    m[18] = 1; m[19] = len & 0xff; m[20] = len >> 8; m[21] = ~len & 0xff; m[22] = (~len >> 8) & 0xff;
    memcpy(m + 23, src, n);
    for (int k = 0; k < 4; k++) { m[23 + n + k] = (crc >> (8 * k)) & 0xff; m[27 + n + k] = (len >> (8 * k)) & 0xff; }
    return fwrite(m, 1, total, f) == total ? 0 : -1;
}

int tsk_open_writers(struct host_config *cfg)
{
    for (struct host_sample *s = cfg->samples; s; s = s->next) {
        if (s->index[0] == 'X') continue;          /* no files for Unknown / control */
        if (!(s->stream_seqqual = open_out(cfg->seqqual_output, "{name}", s->name))) return -1;
        if (cfg->taginfo_output && !(s->stream_taginfo = open_out(cfg->taginfo_output, "{name}", s->name))) return -1;
        if (!(s->stream_signal = open_out(cfg->signal_output, "{name}", s->name))) return -1;
    }
    return 0;
}

void tsk_close_writers(struct host_config *cfg)
{
    for (struct host_sample *s = cfg->samples; s; s = s->next) {
        FILE **fs[3] = {&s->stream_seqqual, &s->stream_taginfo, &s->stream_signal};
        for (int w = 0; w < 3; w++)
            if (*fs[w]) {
                if (device_encoding) {                      /* htslib's end-of-file marker block */
                    const uint8_t *eof; size_t n;
                    if (tsk_bgzf_eof_block(&eof, &n) == 0) fwrite(eof, 1, n, *fs[w]);
                } else if (ftell(*fs[w]) == 0) put_member(*fs[w], "", 0);   /* an empty stream is still a gzip file */
                fclose(*fs[w]);
                *fs[w] = NULL;
            }
    }
}

int tsk_write_signal_headers(struct host_config *cfg, uint32_t total_clusters)
{
    for (struct host_sample *s = cfg->samples; s; s = s->next)
        if (s->stream_signal) {
            const uint32_t h[3] = {4, total_clusters, (uint32_t)s->signal_dump_length};
            if ((device_encoding ? put_bgzf_stored(s->stream_signal, h, sizeof(h)) : put_member(s->stream_signal, h, sizeof(h))) != 0) {
                perror("write_output_file_headers");
                return -1;
            }
        }
    return 0;
}

struct outbuf { char *p; size_t n, cap; };

static int ob_reserve(struct outbuf *b, size_t extra)
{
    if (b->n + extra <= b->cap) return 0;
    size_t cap = b->cap ? b->cap * 2 : (1u << 20);
    while (cap < b->n + extra) cap *= 2;
    char *q = realloc(b->p, cap);
    if (!q) return -1;
    b->p = q; b->cap = cap;
    return 0;
}

/* replace the text in b by its gzip member (nothing for an empty buffer) */
static int ob_compress(struct outbuf *b)
{
    unsigned char *z = NULL;
    size_t zn = 0;
    if (b->n == 0) return 0;
    if (b->n > GZ_MEMBER_MAX) {                    /* chunks are far smaller (<= 32768 clusters) */
        fprintf(stderr, "write_block_outputs: a text chunk of %zu bytes exceeds one gzip member\n", b->n);
        return -1;
    }
    if (gz_member(b->p, b->n, &z, &zn) != 0) return -1;
    free(b->p);
    b->p = (char *)z; b->n = zn; b->cap = zn;
    return 0;
}

/* One chunk of consecutive clusters formatted into per-sample text buffers. */
struct fmt_job {
    const struct host_config *cfg;
    struct host_sample **byidx;
    const tsk_call *calls;
    const uint8_t *bcl;
    uint32_t nclusters, firstclusterno, begin, end;
    struct outbuf *tag, *sq;            /* [num_samples] */
};

static int format_chunk(void *arg)
{
    static const char letters[4] = {'A', 'C', 'G', 'T'};
    struct fmt_job *j = arg;
    const struct host_config *cfg = j->cfg;
    const int T = cfg->total_cycles, ts = cfg->threep_start, tl = cfg->threep_length;
    char *seq = malloc(T + 1), *qual = malloc(T + 1);
    int rc = -1;
    if (!seq || !qual) goto done;
    for (uint32_t n = j->begin; n < j->end; n++) {
        const tsk_call *c = &j->calls[n];
        struct host_sample *s;
        if (!c->written) continue;
        s = j->byidx[c->sample];
        if (!s->stream_seqqual && !s->stream_taginfo) continue;
        for (int k = 0; k < T; k++) {               /* format_basecalls(), bclreader.c:146-166 */
            const uint8_t b = j->bcl[(size_t)k * j->nclusters + n];
            if (b == 0) { seq[k] = 'N'; qual[k] = 2 + 33; }
            else { seq[k] = letters[b & 3]; qual[k] = (char)((b >> 2) + 33); }
        }
        if (s->stream_seqqual) {
            struct outbuf *b = &j->sq[c->sample];
            int start3, len3;
            if (c->delimiter_end >= 0) { start3 = c->delimiter_end; len3 = ts + tl - c->delimiter_end; }
            else { start3 = ts; len3 = tl; }
            if (len3 > cfg->threep_seqqual_output_length) len3 = cfg->threep_seqqual_output_length;
            if (ob_reserve(b, 32 + 2 * (size_t)cfg->fivep_length + 2 * (size_t)len3)) goto done;
            b->n += sprintf(b->p + b->n, "%u\t", (unsigned)(j->firstclusterno + n));
            memcpy(b->p + b->n, seq + cfg->fivep_start, cfg->fivep_length); b->n += cfg->fivep_length; b->p[b->n++] = '\t';
            memcpy(b->p + b->n, qual + cfg->fivep_start, cfg->fivep_length); b->n += cfg->fivep_length; b->p[b->n++] = '\t';
            memcpy(b->p + b->n, seq + start3, len3); b->n += len3; b->p[b->n++] = '\t';
            memcpy(b->p + b->n, qual + start3, len3); b->n += len3; b->p[b->n++] = '\n';
        }
        if (s->stream_taginfo) {
            struct outbuf *b = &j->tag[c->sample];
            size_t need = 64 + (c->terminal_mods > 0 ? c->terminal_mods : 0);
            for (int i = 0; i < s->umi_count; i++) need += s->umi[i].length > 0 ? s->umi[i].length : 0;
            if (ob_reserve(b, need)) goto done;
            b->n += sprintf(b->p + b->n, "%u\t%d\t%d\t", (unsigned)(j->firstclusterno + n), (int)c->flags, (int)c->polya_len);
            for (int i = 0; i < c->terminal_mods; i++) {       /* reverse complement of the modifications */
                char ch;
                switch (seq[c->delimiter_end + c->terminal_mods - 1 - i]) {
                case 'A': ch = 'T'; break; case 'C': ch = 'G'; break;
                case 'G': ch = 'C'; break; case 'T': ch = 'A'; break; default: ch = 'N'; break;
                }
                b->p[b->n++] = ch;
            }
            b->p[b->n++] = '\t';
            for (int i = 0; i < s->umi_count; i++)
                if (s->umi[i].length > 0) { memcpy(b->p + b->n, seq + s->umi[i].start, s->umi[i].length); b->n += s->umi[i].length; }
            b->p[b->n++] = '\n';
        }
    }
    for (int i = 0; i < cfg->num_samples; i++)
        if (ob_compress(&j->sq[i]) || ob_compress(&j->tag[i])) goto done;
    rc = 0;
done:
    free(seq); free(qual);
    return rc;
}

/* A run of record slots of one sample -> one gzip member (withdrawn slots, valid < 0, skipped). */
struct rec_job {
    const uint32_t *records; uint32_t nslots; size_t words;
    struct outbuf z;
};

static int compress_records(void *arg)
{
    struct rec_job *j = arg;
    const uint32_t *r = j->records;
    if (ob_reserve(&j->z, (size_t)j->nslots * j->words * 4 + 1)) return -1;
    for (uint32_t k = 0; k < j->nslots; k++, r += j->words)
        if ((int16_t)(r[1] >> 16) >= 0) { memcpy(j->z.p + j->z.n, r, j->words * 4); j->z.n += j->words * 4; }
    return ob_compress(&j->z);
}

/* One output stream: the chunks' members in cluster order. */
struct out_job {
    FILE *f;
    struct outbuf **parts; int nparts;
};

static int write_stream(void *arg)
{
    struct out_job *o = arg;
    for (int c = 0; c < o->nparts; c++)
        if (o->parts[c]->n && fwrite(o->parts[c]->p, 1, o->parts[c]->n, o->f) != o->parts[c]->n) {
            perror("write_block_outputs");
            return -1;
        }
    return 0;
}

#define REC_SLOTS_PER_MEMBER 8192       /* ~7.6 MB of records per gzip member */

int tsk_write_block_outputs(struct host_config *cfg, const tsk_call *calls, uint32_t nclusters,
                            uint32_t firstclusterno, const uint8_t *bcl, uint32_t **records,
                            const uint32_t *nslots)
{
    const int S = cfg->num_samples;
    const int nthreads = cfg->threads > 0 ? cfg->threads : 1;
    /* a few chunks per thread even out samples of very different sizes; and never more than 32768
     * clusters in a chunk, so that one sample's text of a chunk (< 700 B a line) stays far below
     * what a single gzip member may hold, whatever [options] threads and read-buffer-size are */
    int nchunks = nthreads > 1 ? nthreads * 4 : 1;
    if ((uint32_t)nchunks < (nclusters + 32767u) / 32768u) nchunks = (int)((nclusters + 32767u) / 32768u);
    struct host_sample **byidx = calloc(S, sizeof(*byidx));
    struct fmt_job *jobs = NULL;
    struct rec_job *recs = NULL;
    struct out_job *outs = NULL;
    struct outbuf **parts = NULL;
    struct task *tasks = NULL;
    int rc = -1, nout = 0, nrec = 0, ntask = 0;
    size_t nparts = 0, at = 0;
    if ((uint32_t)nchunks > nclusters) nchunks = nclusters ? (int)nclusters : 1;
    for (struct host_sample *s = cfg->samples; s; s = s->next)
        if (s->stream_signal) nrec += (int)((nslots[s->numindex] + REC_SLOTS_PER_MEMBER - 1) / REC_SLOTS_PER_MEMBER);
    jobs = calloc(nchunks, sizeof(*jobs));
    recs = calloc(nrec ? nrec : 1, sizeof(*recs));
    outs = calloc(3 * (size_t)S, sizeof(*outs));
    nparts = 2 * (size_t)S * nchunks + nrec;
    parts = calloc(nparts ? nparts : 1, sizeof(*parts));
    tasks = calloc((size_t)nchunks + nrec + 3 * S, sizeof(*tasks));
    if (!byidx || !jobs || !recs || !outs || !parts || !tasks) {
        fprintf(stderr, "write_block_outputs: out of memory\n");
        goto done;
    }
    for (struct host_sample *s = cfg->samples; s; s = s->next) byidx[s->numindex] = s;

    /* phase 1, all threads: format + deflate the text chunks, deflate the record chunks */
    for (int c = 0; c < nchunks; c++) {
        struct fmt_job *j = &jobs[c];
        j->cfg = cfg; j->byidx = byidx; j->calls = calls; j->bcl = bcl;
        j->nclusters = nclusters; j->firstclusterno = firstclusterno;
        j->begin = (uint32_t)((uint64_t)nclusters * c / nchunks);
        j->end = (uint32_t)((uint64_t)nclusters * (c + 1) / nchunks);
        j->tag = calloc(S, sizeof(struct outbuf));
        j->sq = calloc(S, sizeof(struct outbuf));
        if (!j->tag || !j->sq) goto done;
        tasks[ntask].fn = format_chunk; tasks[ntask].arg = j; ntask++;
    }
    nrec = 0;
    for (struct host_sample *s = cfg->samples; s; s = s->next) {
        if (!s->stream_signal) continue;
        const size_t words = 2 + (size_t)s->signal_dump_length;
        for (uint32_t k0 = 0; k0 < nslots[s->numindex]; k0 += REC_SLOTS_PER_MEMBER) {
            struct rec_job *j = &recs[nrec++];
            j->records = records[s->numindex] + (size_t)k0 * words;
            j->nslots = nslots[s->numindex] - k0 < REC_SLOTS_PER_MEMBER ? nslots[s->numindex] - k0 : REC_SLOTS_PER_MEMBER;
            j->words = words;
            tasks[ntask].fn = compress_records; tasks[ntask].arg = j; ntask++;
        }
    }
    if (run_tasks(tasks, ntask, nthreads) != 0) {
        fprintf(stderr, "write_block_outputs: formatting / compressing the block's outputs failed\n");
        goto done;
    }

    /* phase 2: append the members to each file in cluster order (one task per file) */
    nrec = 0;
    for (struct host_sample *s = cfg->samples; s; s = s->next) {
        const int i = s->numindex;
        if (s->stream_seqqual) {
            struct out_job *o = &outs[nout];
            o->f = s->stream_seqqual; o->parts = parts + at; o->nparts = nchunks;
            for (int c = 0; c < nchunks; c++) parts[at++] = &jobs[c].sq[i];
            tasks[nout].fn = write_stream; tasks[nout].arg = o; nout++;
        }
        if (s->stream_taginfo) {
            struct out_job *o = &outs[nout];
            o->f = s->stream_taginfo; o->parts = parts + at; o->nparts = nchunks;
            for (int c = 0; c < nchunks; c++) parts[at++] = &jobs[c].tag[i];
            tasks[nout].fn = write_stream; tasks[nout].arg = o; nout++;
        }
        if (s->stream_signal) {
            struct out_job *o = &outs[nout];
            const int n = (int)((nslots[i] + REC_SLOTS_PER_MEMBER - 1) / REC_SLOTS_PER_MEMBER);
            o->f = s->stream_signal; o->parts = parts + at; o->nparts = n;
            for (int c = 0; c < n; c++) parts[at++] = &recs[nrec++].z;
            tasks[nout].fn = write_stream; tasks[nout].arg = o; nout++;
        }
    }
    if (run_tasks(tasks, nout, nthreads) != 0) {
        fprintf(stderr, "write_block_outputs: writing the block's outputs failed\n");
        goto done;
    }
    rc = 0;
done:
    if (jobs)
        for (int c = 0; c < nchunks; c++) {
            if (jobs[c].tag) for (int i = 0; i < S; i++) free(jobs[c].tag[i].p);
            if (jobs[c].sq) for (int i = 0; i < S; i++) free(jobs[c].sq[i].p);
            free(jobs[c].tag); free(jobs[c].sq);
        }
    if (recs) for (int i = 0; i < nrec; i++) free(recs[i].z.p);
    free(jobs); free(recs); free(outs); free(parts); free(tasks); free(byidx);
    return rc;
}

/* The block's outputs as the GPU encoded them (tsk_tile_encode): whole BGZF blocks per stream,
 * appended to the sample's files, one writer task per file. */
struct enc_job { FILE *f; const uint8_t *data; size_t n; };

static int write_encoded(void *arg)
{
    struct enc_job *j = arg;
    if (j->n && fwrite(j->data, 1, j->n, j->f) != j->n) { perror("write_block_outputs"); return -1; }
    return 0;
}

int tsk_write_encoded_outputs(struct host_config *cfg, tsk_ctx *ctx)
{
    const int S = cfg->num_samples;
    struct enc_job *jobs = calloc(3 * (size_t)S, sizeof(*jobs));
    struct task *tasks = calloc(3 * (size_t)S, sizeof(*tasks));
    int n = 0, rc = -1;
    if (!jobs || !tasks) { fprintf(stderr, "write_block_outputs: out of memory\n"); goto done; }
    for (struct host_sample *s = cfg->samples; s; s = s->next) {
        FILE *fs[3] = {s->stream_seqqual, s->stream_taginfo, s->stream_signal};
        for (int kind = 0; kind < 3; kind++) {
            if (!fs[kind]) continue;
            jobs[n].f = fs[kind];
            if (tsk_tile_get_encoded(ctx, s->numindex, kind, &jobs[n].data, &jobs[n].n) != 0) {
                fprintf(stderr, "%s\n", tsk_last_error());
                goto done;
            }
            tasks[n].fn = write_encoded; tasks[n].arg = &jobs[n];
            n++;
        }
    }
    if (n && run_tasks(tasks, n, cfg->threads > 0 ? cfg->threads : 1) != 0) {
        fprintf(stderr, "write_block_outputs: writing the block's outputs failed\n");
        goto done;
    }
    rc = 0;
done:
    free(jobs); free(tasks);
    return rc;
}

int tsk_write_sigdists(const char *filename_format, const char *type, const uint32_t *counts,
                       int total_cycles, int bins)
{
    const uint32_t h[3] = {4, (uint32_t)total_cycles, (uint32_t)bins};
    char *fn = tsk_replace_placeholder(filename_format, "{posneg}", type);
    gzFile f;
    size_t bytes = sizeof(uint32_t) * (size_t)total_cycles * bins, at = 0;
    int rc = 0;
    if (!fn) return -1;
    {
        char mode[8] = "wb";
        if (gz_level() >= 0) snprintf(mode, sizeof(mode), "wb%d", gz_level());
        f = gzopen(fn, mode);
    }
    if (!f) { fprintf(stderr, "Cannot open %s to write.\n", fn); free(fn); return -1; }
    free(fn);
    if (gzwrite(f, h, sizeof(h)) <= 0) rc = -1;
    while (rc == 0 && at < bytes) {
        const unsigned chunk = bytes - at > (1u << 30) ? (1u << 30) : (unsigned)(bytes - at);
        if (gzwrite(f, (const char *)counts + at, chunk) <= 0) rc = -1;
        at += chunk;
    }
    gzclose(f);
    return rc;
}

int tsk_write_stats(const char *output, const struct host_config *cfg, const tsk_counters *counters)
{
    FILE *fp = fopen(output, "w");
    if (!fp) {
        perror("write_demultiplexing_statistics");
        fprintf(stderr, "Failed to write the demultiplexing statistics: %s\n", output);
        return -1;
    }
    fprintf(fp, "name,index,max_index_mm,delim,delim_pos,max_delim_mm,cln_no_index_mm,cln_1_index_mm,"
                "cln_2+_index_mm,cln_fp_mm,cln_qc_fail,cln_no_delim\n");
    for (const struct host_sample *s = cfg->samples; s; s = s->next) {
        const tsk_counters *c = &counters[s->numindex];
        fprintf(fp, "%s,%s,%d,%s,%d,%d,%d,%d,%d,%d,%d,%d\n", s->name, s->index ? s->index : "(null)",
                s->maximum_index_mismatches, s->delimiter ? s->delimiter : "(null)", s->delimiter_pos,
                s->maximum_delimiter_mismatches, c->clusters_mm0, c->clusters_mm1, c->clusters_mm2plus,
                c->clusters_fpmismatch, c->clusters_qcfailed, c->clusters_nodelim);
    }
    fclose(fp);
    return 0;
}

/* ---- shared with the other host programs ------------------------------------------------ */
int tsk_gz_member(const void *src, size_t n, unsigned char **out, size_t *outn)
{
    return gz_member(src, n, out, outn);
}

int tsk_run_tasks(struct tsk_task *tasks, int ntasks, int nthreads)
{
    return run_tasks((struct task *)tasks, ntasks, nthreads);      /* same layout */
}
