/*
 * writers.c -- the importer's output files, driven by the per-cluster calls from the GPU:
 *   taginfo / seqqual text   (reference src/importer/spotanalyzer.c:206-352)
 *   .sigpack                 (src/importer/tailseq-import.c:98-119, spotanalyzer.c:389-438)
 *   .sigdists                (tailseq-import.c:268-300)
 *   demultiplexing CSV       (tailseq-import.c:685-714)
 * Streams are gzip (every consumer opens them with gzopen / gzip.open / zcat); the reference
 * writes BGZF blocks, which are gzip members too.
 */
#include <stdlib.h>
#include <string.h>
#include "tsk_host.h"

static gzFile open_out(const char *fmt, const char *key, const char *val)
{
    char *fn = tsk_replace_placeholder(fmt, key, val);
    gzFile f;
    if (!fn) return NULL;
    f = gzopen(fn, "wb");
    if (!f) { perror("open_writers"); fprintf(stderr, "Failed to write to %s\n", fn); }
    free(fn);
    return f;
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
        if (s->stream_seqqual) { gzclose(s->stream_seqqual); s->stream_seqqual = NULL; }
        if (s->stream_taginfo) { gzclose(s->stream_taginfo); s->stream_taginfo = NULL; }
        if (s->stream_signal) { gzclose(s->stream_signal); s->stream_signal = NULL; }
    }
}

int tsk_write_signal_headers(struct host_config *cfg, uint32_t total_clusters)
{
    for (struct host_sample *s = cfg->samples; s; s = s->next)
        if (s->stream_signal) {
            const uint32_t h[3] = {4, total_clusters, (uint32_t)s->signal_dump_length};
            if (gzwrite(s->stream_signal, h, sizeof(h)) <= 0) { perror("write_output_file_headers"); return -1; }
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

static int ob_flush(struct outbuf *b, gzFile f)
{
    size_t at = 0;
    while (at < b->n) {
        const unsigned chunk = b->n - at > (1u << 30) ? (1u << 30) : (unsigned)(b->n - at);
        if (gzwrite(f, b->p + at, chunk) <= 0) return -1;
        at += chunk;
    }
    b->n = 0;
    return 0;
}

int tsk_write_block_outputs(struct host_config *cfg, const tsk_call *calls, uint32_t nclusters,
                            uint32_t firstclusterno, const uint8_t *bcl, uint32_t **records,
                            const uint32_t *nslots)
{
    static const char letters[4] = {'A', 'C', 'G', 'T'};
    const int T = cfg->total_cycles;
    const int ts = cfg->threep_start, tl = cfg->threep_length;
    struct host_sample **byidx = calloc(cfg->num_samples, sizeof(*byidx));
    struct outbuf *tag = calloc(cfg->num_samples, sizeof(*tag)), *sq = calloc(cfg->num_samples, sizeof(*sq));
    char *seq = malloc(T + 1), *qual = malloc(T + 1);
    int rc = -1;
    if (!byidx || !tag || !sq || !seq || !qual) goto done;
    for (struct host_sample *s = cfg->samples; s; s = s->next) byidx[s->numindex] = s;

    for (uint32_t n = 0; n < nclusters; n++) {
        const tsk_call *c = &calls[n];
        struct host_sample *s;
        if (!c->written) continue;
        s = byidx[c->sample];
        if (!s->stream_seqqual && !s->stream_taginfo) continue;
        for (int k = 0; k < T; k++) {               /* format_basecalls(), bclreader.c:146-166 */
            const uint8_t b = bcl[(size_t)k * nclusters + n];
            if (b == 0) { seq[k] = 'N'; qual[k] = 2 + 33; }
            else { seq[k] = letters[b & 3]; qual[k] = (char)((b >> 2) + 33); }
        }
        if (s->stream_seqqual) {
            struct outbuf *b = &sq[c->sample];
            int start3, len3;
            if (c->delimiter_end >= 0) { start3 = c->delimiter_end; len3 = ts + tl - c->delimiter_end; }
            else { start3 = ts; len3 = tl; }
            if (len3 > cfg->threep_seqqual_output_length) len3 = cfg->threep_seqqual_output_length;
            if (ob_reserve(b, 32 + 2 * (size_t)cfg->fivep_length + 2 * (size_t)len3)) goto done;
            b->n += sprintf(b->p + b->n, "%u\t", (unsigned)(firstclusterno + n));
            memcpy(b->p + b->n, seq + cfg->fivep_start, cfg->fivep_length); b->n += cfg->fivep_length; b->p[b->n++] = '\t';
            memcpy(b->p + b->n, qual + cfg->fivep_start, cfg->fivep_length); b->n += cfg->fivep_length; b->p[b->n++] = '\t';
            memcpy(b->p + b->n, seq + start3, len3); b->n += len3; b->p[b->n++] = '\t';
            memcpy(b->p + b->n, qual + start3, len3); b->n += len3; b->p[b->n++] = '\n';
        }
        if (s->stream_taginfo) {
            struct outbuf *b = &tag[c->sample];
            size_t need = 64 + (c->terminal_mods > 0 ? c->terminal_mods : 0);
            for (int i = 0; i < s->umi_count; i++) need += s->umi[i].length > 0 ? s->umi[i].length : 0;
            if (ob_reserve(b, need)) goto done;
            b->n += sprintf(b->p + b->n, "%u\t%d\t%d\t", (unsigned)(firstclusterno + n), (int)c->flags, (int)c->polya_len);
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
    for (struct host_sample *s = cfg->samples; s; s = s->next) {
        if (s->stream_seqqual && ob_flush(&sq[s->numindex], s->stream_seqqual)) goto done;
        if (s->stream_taginfo && ob_flush(&tag[s->numindex], s->stream_taginfo)) goto done;
        if (s->stream_signal) {
            /* record slots in cluster order; withdrawn slots (valid < 0) are skipped */
            const size_t words = 2 + (size_t)s->signal_dump_length;
            const uint32_t *r = records[s->numindex];
            for (uint32_t k = 0; k < nslots[s->numindex]; k++, r += words)
                if ((int16_t)(r[1] >> 16) >= 0 && gzwrite(s->stream_signal, r, (unsigned)(words * 4)) <= 0) {
                    perror("write_polya_score");
                    goto done;
                }
        }
    }
    rc = 0;
done:
    if (tag) for (int i = 0; i < cfg->num_samples; i++) free(tag[i].p);
    if (sq) for (int i = 0; i < cfg->num_samples; i++) free(sq[i].p);
    free(tag); free(sq); free(seq); free(qual); free(byidx);
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
    f = gzopen(fn, "wb");
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
