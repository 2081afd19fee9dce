/*
 * tailseq-dedup-perfect (B200) -- drop-in for the reference program of the same name
 * (src/deduplicator/tailseq-dedup-perfect.c): taginfo lines sorted by UMI on stdin, one line per
 * UMI group on stdout, optional duplicate trace (gzip) named by argv[1].
 *
 *   tailseq-dedup-perfect [trace.gz] < sorted.txt > collapsed.txt        the reference's argv
 *   tailseq-dedup-perfect --sort [trace.gz] < tiles.txt > collapsed.txt
 *        stdin = the tiles' taginfo files concatenated in (tile, cluster) order, NOT sorted: the
 *        UMI sort runs on the GPU and stdout comes in input order, i.e. the whole
 *        `sort -k6,6 -k1,1 -k2,2n | tailseq-dedup-perfect | sort -k1,1 -k2,2n` of
 *        tailseeker/main.py:272-278 in one process.
 *
 * The host parses the lines (parse_line(), :216-274), packs the UMIs and prints; grouping,
 * priorities, means and representatives come from libtailseeker_b200.so (tsk_dedup_*).
 */
#include <errno.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "tsk_host.h"

#define MAX_LINE_LEN 511          /* the reference's fgets() buffer (:42,369) */

/* TSK_HOST_TIMING=1: wall time of each phase on stderr */
#include <time.h>
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

struct rec_table {
    const char *buf;
    size_t nlines;
    const char **line;            /* start of every line */
    int *cluster, *flags, *polya;
    uint16_t *tile_len, *head_len, *mods_off, *mods_len;
    uint64_t *khi, *klo;
    long bad_line;                /* first line that does not parse, or -1 */
    char error[256];
};

static int nthreads(void)
{
    const char *e = getenv("TSK_THREADS");
    long n = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
    if (n < 1) n = 1;
    if (n > 64) n = 64;
    return (int)n;
}

static char *read_all(FILE *f, size_t *len)
{
    size_t cap = 1 << 24, n = 0;
    char *buf = malloc(cap + 1);
    if (!buf) return NULL;
    for (;;) {
        if (n == cap) {
            cap += cap / 2;
            char *nb = realloc(buf, cap + 1);
            if (!nb) { free(buf); return NULL; }
            buf = nb;
        }
        const size_t r = fread(buf + n, 1, cap - n, f);
        n += r;
        if (r == 0) break;
    }
    buf[n] = '\0';
    *len = n;
    return buf;
}

static int parse_int(const char **pp, const char *end, int *out)
{
    const char *p = *pp;
    int neg = 0;
    long v = 0;
    if (p < end && (*p == '-' || *p == '+')) { neg = *p == '-'; p++; }
    if (p >= end || *p < '0' || *p > '9') return -1;
    while (p < end && *p >= '0' && *p <= '9') { v = v * 10 + (*p - '0'); if (v > 0x7fffffffL) return -1; p++; }
    if (p >= end || *p != '\t') return -1;
    *out = (int)(neg ? -v : v);
    *pp = p + 1;
    return 0;
}

struct parse_job { struct rec_table *t; size_t beg, end; long bad; char error[256]; };

static int parse_range(void *arg)
{
    struct parse_job *j = arg;
    struct rec_table *t = j->t;
    j->bad = -1;
    for (size_t i = j->beg; i < j->end; i++) {
        const char *ln = t->line[i], *end = t->line[i + 1] - 1;        /* end -> the '\n' */
        const char *p = memchr(ln, '\t', (size_t)(end - ln));
        if (!p || end - ln > MAX_LINE_LEN - 1) { j->bad = (long)i; return 0; }
        t->tile_len[i] = (uint16_t)(p - ln);
        p++;
        if (parse_int(&p, end, &t->cluster[i]) || parse_int(&p, end, &t->flags[i]) ||
            parse_int(&p, end, &t->polya[i])) { j->bad = (long)i; return 0; }
        const char *m = memchr(p, '\t', (size_t)(end - p));
        if (!m) { j->bad = (long)i; return 0; }
        t->mods_off[i] = (uint16_t)(p - ln);
        t->mods_len[i] = (uint16_t)(m - p);
        t->head_len[i] = (uint16_t)(m - ln);
        uint64_t key[2];
        if (tsk_dedup_pack_umi(m + 1, (size_t)(end - (m + 1)), key) != 0) {
            j->bad = (long)i;
            snprintf(j->error, sizeof(j->error), "%s", tsk_last_error());
            return 0;
        }
        t->khi[i] = key[0]; t->klo[i] = key[1];
    }
    return 0;
}

static void free_table(struct rec_table *t)
{
    free(t->line); free(t->cluster); free(t->flags); free(t->polya); free(t->tile_len); free(t->head_len);
    free(t->mods_off); free(t->mods_len); free(t->khi); free(t->klo);
}

static int build_table(struct rec_table *t, const char *buf, size_t len)
{
    size_t cap = len / 40 + 16, n = 0;
    memset(t, 0, sizeof(*t));
    t->buf = buf; t->bad_line = -1;
    t->line = malloc(sizeof(*t->line) * (cap + 1));
    if (!t->line) return -1;
    const char *p = buf, *end = buf + len;
    while (p < end) {
        const char *nl = memchr(p, '\n', (size_t)(end - p));
        if (n == cap) {
            cap += cap / 2;
            const char **nn = realloc(t->line, sizeof(*t->line) * (cap + 1));
            if (!nn) return -1;
            t->line = nn;
        }
        t->line[n] = p;
        if (!nl) { t->bad_line = (long)n; break; }        /* no newline: parse_line() fails (:264-266) */
        n++;
        p = nl + 1;
    }
    t->line[n] = p;
    t->nlines = n;
    const size_t m = n ? n : 1;
    t->cluster = malloc(sizeof(int) * m); t->flags = malloc(sizeof(int) * m); t->polya = malloc(sizeof(int) * m);
    t->tile_len = malloc(2 * m); t->head_len = malloc(2 * m); t->mods_off = malloc(2 * m); t->mods_len = malloc(2 * m);
    t->khi = malloc(8 * m); t->klo = malloc(8 * m);
    if (!t->cluster || !t->flags || !t->polya || !t->tile_len || !t->head_len || !t->mods_off || !t->mods_len ||
        !t->khi || !t->klo) return -1;

    enum { JOBS = 64 };
    struct parse_job jobs[JOBS];
    struct tsk_task tasks[JOBS];
    for (int k = 0; k < JOBS; k++) {
        jobs[k].t = t; jobs[k].beg = n * k / JOBS; jobs[k].end = n * (k + 1) / JOBS; jobs[k].error[0] = '\0';
        tasks[k].fn = parse_range; tasks[k].arg = &jobs[k];
    }
    if (tsk_run_tasks(tasks, JOBS, nthreads()) != 0) return -1;
    for (int k = 0; k < JOBS; k++)
        if (jobs[k].bad >= 0) {
            t->bad_line = jobs[k].bad;
            snprintf(t->error, sizeof(t->error), "%s", jobs[k].error);
            t->nlines = (size_t)jobs[k].bad;              /* like the reference: stop there, finish the rest */
            break;
        }
    return 0;
}

/* ---- output -------------------------------------------------------------------------------- */
struct outbuf { char *p; size_t n, cap; };

static int ob_room(struct outbuf *b, size_t extra)
{
    if (b->n + extra <= b->cap) return 0;
    size_t cap = b->cap ? b->cap : 1 << 16;
    while (cap < b->n + extra) cap += cap / 2;
    char *np = realloc(b->p, cap);
    if (!np) return -1;
    b->p = np; b->cap = cap;
    return 0;
}

struct fmt_ctx {
    const struct rec_table *t;
    const uint32_t *order, *group, *slot_at;      /* slot_at[slot] = position */
    const int32_t *tval;
    const uint32_t *rep, *ndup, *kept, *emit;     /* emit[k] = group printed k-th */
    const int32_t *gpolya;
    const uint8_t *lost;                          /* per record (input index), or NULL */
};

struct fmt_job { const struct fmt_ctx *c; size_t beg, end; int trace; struct outbuf out; unsigned char *z; size_t zn; };

static int format_range(void *arg)
{
    struct fmt_job *j = arg;
    const struct fmt_ctx *c = j->c;
    const struct rec_table *t = c->t;
    struct outbuf *b = &j->out;
    for (size_t k = j->beg; k < j->end; k++) {
        if (ob_room(b, MAX_LINE_LEN + 64) != 0) return -1;
        char *w = b->p + b->n;
        if (j->trace) {
            const uint32_t p = c->slot_at[k], r = c->order[p];
            if (c->lost && c->lost[r]) b->n += (size_t)sprintf(w, "\t0\t%u\t%d\n", c->group[p], c->tval[p]);
            else {
                memcpy(w, t->line[r], t->tile_len[r]);
                b->n += t->tile_len[r] + (size_t)sprintf(w + t->tile_len[r], "\t%d\t%u\t%d\n", t->cluster[r], c->group[p], c->tval[p]);
            }
        } else {
            const uint32_t g = c->emit[k], r = c->order[c->rep[g]];
            if (c->kept[g] == 1) {                                        /* :286-287 */
                memcpy(w, t->line[r], t->head_len[r]);
                b->n += t->head_len[r] + (size_t)sprintf(w + t->head_len[r], "\t%u\n", c->ndup[g]);
            } else if (c->lost && c->lost[r]) {
                b->n += (size_t)sprintf(w, "\t0\t0\t%d\t\t%u\n", c->gpolya[g], c->ndup[g]);
            } else {                                                      /* :324-326 */
                memcpy(w, t->line[r], t->tile_len[r]);
                size_t n = t->tile_len[r];
                n += (size_t)sprintf(w + n, "\t%d\t%d\t%d\t", t->cluster[r], t->flags[r], c->gpolya[g]);
                memcpy(w + n, t->line[r] + t->mods_off[r], t->mods_len[r]);
                n += t->mods_len[r];
                n += (size_t)sprintf(w + n, "\t%u\n", c->ndup[g]);
                b->n += n;
            }
        }
    }
    if (j->trace && b->n > 0 && tsk_gz_member(b->p, b->n, &j->z, &j->zn) != 0) return -1;
    return 0;
}

static int emit_all(const struct fmt_ctx *c, size_t count, int trace, FILE *f)
{
    enum { PER_JOB = 1 << 17 };
    int rc = 0;
    for (size_t base = 0; base < count && rc == 0; base += (size_t)PER_JOB * 64) {
        struct fmt_job jobs[64];
        struct tsk_task tasks[64];
        int nj = 0;
        for (size_t b = base; b < count && nj < 64; b += PER_JOB, nj++) {
            memset(&jobs[nj], 0, sizeof(jobs[nj]));
            jobs[nj].c = c; jobs[nj].beg = b; jobs[nj].end = b + PER_JOB < count ? b + PER_JOB : count; jobs[nj].trace = trace;
            tasks[nj].fn = format_range; tasks[nj].arg = &jobs[nj];
        }
        rc = tsk_run_tasks(tasks, nj, nthreads());
        for (int k = 0; k < nj; k++) {
            if (rc == 0) {
                const void *src = trace ? (const void *)jobs[k].z : (const void *)jobs[k].out.p;
                const size_t n = trace ? jobs[k].zn : jobs[k].out.n;
                if (n && fwrite(src, 1, n, f) != n) rc = -1;
            }
            free(jobs[k].out.p); free(jobs[k].z);
        }
    }
    return rc;
}

struct boot { pthread_t th; int device, rc, started; char error[512]; };
static void *boot_cuda(void *arg)
{
    struct boot *b = arg;
    tsk_dedup *d = NULL;
    b->rc = tsk_dedup_create(&d, b->device, 1);          /* brings the CUDA context up while stdin is read */
    if (b->rc != 0) snprintf(b->error, sizeof(b->error), "%s", tsk_last_error());
    else tsk_dedup_destroy(d);
    return NULL;
}

int main(int argc, char *argv[])
{
    int sort = 0, argi = 1, rc = -1;
    const char *tracefn = NULL, *dev = getenv("TAILSEEKER_DEVICE");
    FILE *tracef = NULL;
    struct rec_table t;
    struct boot boot;
    tsk_dedup *d = NULL;
    char *buf = NULL;
    size_t len = 0;
    uint32_t *order = NULL, *group = NULL, *tslot = NULL, *slot_at = NULL, *rep = NULL, *ndup = NULL, *kept = NULL, *emit = NULL;
    int32_t *tval = NULL, *gpolya = NULL;
    uint8_t *lost = NULL;

    if (argi < argc && strcmp(argv[argi], "--sort") == 0) { sort = 1; argi++; }
    if (argi < argc) tracefn = argv[argi];
    memset(&t, 0, sizeof(t));
    memset(&boot, 0, sizeof(boot));
    boot.device = dev ? atoi(dev) : 0;
    if (pthread_create(&boot.th, NULL, boot_cuda, &boot) == 0) boot.started = 1;
    else boot_cuda(&boot);

    if (tracefn) {
        tracef = fopen(tracefn, "wb");
        if (!tracef) { fprintf(stderr, "ERROR: Cannot open %s.", tracefn); goto done; }
    }
    phase("start");
    buf = read_all(stdin, &len);
    phase("read stdin");
    if (!buf) { perror("tailseq-dedup-perfect"); goto done; }
    if (ferror(stdin)) { perror("fgets"); goto done; }
    if (build_table(&t, buf, len) != 0) { perror("tailseq-dedup-perfect"); goto done; }
    if (t.bad_line >= 0) {
        const char *ln = t.line[t.bad_line];
        const char *e = memchr(ln, '\n', (size_t)(buf + len - ln));
        fprintf(stderr, "Could not parse line %ld: %.*s%s%s\n", t.bad_line, (int)((e ? e : buf + len) - ln), ln,
                t.error[0] ? " -- " : "", t.error);
    }
    phase("index + parse lines");
    if (boot.started) { pthread_join(boot.th, NULL); boot.started = 0; }
    if (boot.rc != 0) { fprintf(stderr, "%s\n", boot.error); goto done; }
    phase("wait for the CUDA context");

    const size_t n = t.nlines;
    uint64_t ngroups = 0;
    if (n > 0) {
        if (tsk_dedup_create(&d, boot.device, n) != 0 || tsk_dedup_upload(d, n, t.khi, t.klo, t.flags, t.polya) != 0 ||
            tsk_dedup_run(d, sort, &ngroups) != 0) { fprintf(stderr, "%s\n", tsk_last_error()); goto done; }
        phase("upload + sort + collapse");
        order = malloc(4 * n); group = malloc(4 * n); tval = malloc(4 * n); tslot = malloc(4 * n);
        rep = malloc(4 * ngroups); gpolya = malloc(4 * ngroups); ndup = malloc(4 * ngroups); kept = malloc(4 * ngroups);
        emit = malloc(4 * ngroups);
        if (!order || !group || !tval || !tslot || !rep || !gpolya || !ndup || !kept || !emit) { perror("tailseq-dedup-perfect"); goto done; }
        if (tsk_dedup_get_records(d, order, group, tval, tracef ? tslot : NULL) != 0 ||
            tsk_dedup_get_groups(d, rep, gpolya, ndup, kept) != 0) { fprintf(stderr, "%s\n", tsk_last_error()); goto done; }
        uint32_t lostpos[64], nlost = 0;
        if (tsk_dedup_get_lost(d, lostpos, 64, &nlost) != 0) { fprintf(stderr, "%s\n", tsk_last_error()); goto done; }
        if (nlost > 0) {
            lost = calloc(n, 1);
            if (!lost) { perror("tailseq-dedup-perfect"); goto done; }
            for (uint32_t i = 0; i < nlost; i++) lost[order[lostpos[i]]] = 1;
        }
        phase("fetch results");
        /* stdout order: groups as they come (reference), or by the representative's place in the input */
        if (sort) {
            /* one pass over the input positions: group whose representative sits there, if any */
            uint32_t *at = calloc(n, 4);
            if (!at) { perror("tailseq-dedup-perfect"); goto done; }
            for (uint64_t g = 0; g < ngroups; g++) at[order[rep[g]]] = (uint32_t)g + 1;
            uint64_t k = 0;
            for (size_t i = 0; i < n; i++) if (at[i]) emit[k++] = at[i] - 1;
            free(at);
        } else
            for (uint64_t g = 0; g < ngroups; g++) emit[g] = (uint32_t)g;
        phase("output order");
        struct fmt_ctx c = {&t, order, group, NULL, tval, rep, ndup, kept, emit, gpolya, lost};
        if (emit_all(&c, ngroups, 0, stdout) != 0) { perror("tailseq-dedup-perfect"); goto done; }
        phase("format + write stdout");
        if (tracef) {
            slot_at = malloc(4 * n);
            if (!slot_at) { perror("tailseq-dedup-perfect"); goto done; }
            for (size_t p = 0; p < n; p++) slot_at[tslot[p]] = (uint32_t)p;
            c.slot_at = slot_at;
            if (emit_all(&c, n, 1, tracef) != 0) { perror("tailseq-dedup-perfect"); goto done; }
            phase("format + compress trace");
        }
    }
    if (tracef && n == 0) {                                   /* an empty gzip stream, like bgzf_close() on nothing */
        unsigned char *z = NULL; size_t zn = 0;
        if (tsk_gz_member("", 0, &z, &zn) != 0 || fwrite(z, 1, zn, tracef) != zn) { free(z); goto done; }
        free(z);
    }
    if (fflush(stdout) != 0) goto done;
    rc = 0;
done:
    if (boot.started) pthread_join(boot.th, NULL);
    if (tracef && fclose(tracef) != 0) rc = -1;
    if (d) tsk_dedup_destroy(d);
    free(order); free(group); free(tval); free(tslot); free(slot_at); free(rep); free(gpolya); free(ndup); free(kept);
    free(emit); free(lost);
    free_table(&t);
    free(buf);
    phase("clean up");
    return rc;
}
