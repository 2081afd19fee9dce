/*
 * config.c -- importer configuration: INI sections of scripts/generate-signalproc-conf.py:31-221 into
 * host_config, defaults and derived values as src/importer/parseconfig.c:39-671 computes them,
 * and the POD tsk_params handed to the CUDA library.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include "tsk_host.h"

#define IS(n) (strcasecmp(name, n) == 0)

static int yesno(const char *name, const char *value, int *out)
{
    if (strcasecmp(value, "yes") == 0 || strcmp(value, "1") == 0) *out = 1;
    else if (strcasecmp(value, "no") == 0 || strcmp(value, "0") == 0) *out = 0;
    else { fprintf(stderr, "\"%s\" must be either yes or no.\n", name); return -1; }
    return 0;
}

static int weight_slot(char c)
{
    if (c >= 'a' && c <= 'z') c = (char)(c - 'a' + 'A');
    switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3;
                 case 'N': return 4; default: return -1; }
}

static struct host_sample *sample_named(struct host_config *cfg, const char *sname, const char *key)
{
    struct host_sample *s;
    for (s = cfg->samples; s != NULL; s = s->next)
        if (*key != '\0' && strcmp(s->name, sname) == 0)
            return s;
    s = calloc(1, sizeof(*s));
    if (!s) return NULL;
    strncpy(s->name, sname, TSK_NAME_MAX - 1);
    s->next = cfg->samples;            /* LIFO: the list ends up in reverse file order */
    cfg->samples = s;
    return s;
}

static int feed_sample(struct host_config *cfg, const char *sname, const char *name, const char *value)
{
    struct host_sample *s = sample_named(cfg, sname, name);
    if (!s) return -1;
    if (IS("index")) s->index = strdup(value);
    else if (IS("maximum-index-mismatch")) s->maximum_index_mismatches = atoi(value);
    else if (IS("delimiter-seq")) s->delimiter = strdup(value);
    else if (IS("delimiter-start")) s->delimiter_pos = atoi(value) - 1;
    else if (IS("maximum-delimiter-mismatch")) s->maximum_delimiter_mismatches = atoi(value);
    else if (IS("fingerprint-seq")) s->fingerprint = strdup(value);
    else if (IS("fingerprint-start")) s->fingerprint_pos = atoi(value) - 1;
    else if (IS("maximum-fingerprint-mismatch")) s->maximum_fingerprint_mismatches = atoi(value);
    else if (IS("limit-threep-processing")) s->limit_threep_processing = atoi(value);
    else if (strncasecmp(name, "umi-", 4) == 0) {
        const char *colon = strrchr(name, ':');
        int no;
        if (!colon) { fprintf(stderr, "umi qualifiers must passed with a number.\n"); return -1; }
        no = atoi(colon + 1);
        if (no < 1) { fprintf(stderr, "umi number must start from 1.\n"); return -1; }
        if (no > TSK_MAX_UMI) { fprintf(stderr, "too many umi ranges in sample %s.\n", s->name); return -1; }
        if (s->umi_count < no) s->umi_count = no;
        if (strncasecmp(name, "umi-start:", 10) == 0) s->umi[no - 1].start = atoi(value) - 1;
        else if (strncasecmp(name, "umi-length:", 11) == 0) s->umi[no - 1].length = atoi(value);
        else { fprintf(stderr, "Unknown key \"%s\" in sample %s.\n", name, s->name); return -1; }
    } else {
        fprintf(stderr, "Unknown key \"%s\" in sample %s.\n", name, s->name);
        return -1;
    }
    return 0;
}

static int feed(void *user, const char *section, const char *name, const char *value)
{
    struct host_config *c = user;
#define SEC(s) (strcasecmp(section, s) == 0)
#define UNKNOWN(sec) do { fprintf(stderr, "Unknown key \"%s\" in [" sec "].\n", name); return -1; } while (0)
    if (SEC("source")) {
        if (IS("data-dir")) c->datadir = strdup(value);
        else if (IS("laneid")) c->laneid = strdup(value);
        else if (IS("lane")) c->lane = atoi(value);
        else if (IS("tile")) c->tile = atoi(value);
        else if (IS("threep-colormatrix")) c->threep_colormatrix_filename = strdup(value);
        else UNKNOWN("source");
    } else if (SEC("read_format")) {
        if (IS("total-cycles")) c->total_cycles = atoi(value);
        else if (IS("fivep-start")) c->fivep_start = atoi(value) - 1;
        else if (IS("fivep-length")) c->fivep_length = atoi(value);
        else if (IS("index-start")) c->index_start = atoi(value) - 1;
        else if (IS("index-length")) c->index_length = atoi(value);
        else if (IS("threep-start")) c->threep_start = atoi(value) - 1;
        else if (IS("threep-length")) c->threep_length = atoi(value);
        else if (IS("threep-seqqual-output-length")) c->threep_seqqual_output_length = atoi(value);
        else UNKNOWN("read_format");
    } else if (SEC("options")) {
        if (IS("keep-no-delimiter")) return yesno(name, value, &c->keep_no_delimiter);
        else if (IS("keep-low-quality-balancer")) return yesno(name, value, &c->keep_low_quality_balancer);
        else if (IS("threads")) c->threads = atoi(value);
        else if (IS("read-buffer-size")) c->read_buffer_size = (size_t)atoll(value);
        else UNKNOWN("options");
    } else if (SEC("output")) {
        if (IS("seqqual")) c->seqqual_output = strdup(value);
        else if (IS("taginfo")) c->taginfo_output = strdup(value);
        else if (IS("signal")) c->signal_output = strdup(value);
        else if (IS("signal-dists")) c->signal_dists_output = strdup(value);
        else if (IS("stats")) c->stats_output = strdup(value);
        else if (IS("length-dists")) c->length_dists_output = strdup(value);
        else UNKNOWN("output");
    } else if (SEC("alternative_calls")) {
        struct host_altcall *ac;
        if (atoi(name) <= 0) { fprintf(stderr, "Alternative call position is out of range.\n"); return -1; }
        ac = calloc(1, sizeof(*ac));
        if (!ac || !(ac->filename = strdup(value))) { perror("alternative_calls"); free(ac); return -1; }
        ac->first_cycle = atoi(name) - 1;
        ac->next = c->altcalls;
        c->altcalls = ac;
    } else if (SEC("control")) {
        if (IS("phix-match-name")) { strncpy(c->control_name, value, TSK_NAME_MAX - 1); }
        else if (IS("phix-match-start")) c->control_first_cycle = atoi(value) - 1;
        else if (IS("phix-match-length")) c->control_read_length = atoi(value);
        else UNKNOWN("control");
    } else if (SEC("balancer")) {
        if (IS("start")) c->balancer_start = atoi(value) - 1;
        else if (IS("length")) c->balancer_length = atoi(value);
        else if (IS("minimum-occurrence")) c->balancer_minimum_occurrence = atoi(value);
        else if (IS("num-positive-samples")) c->balancer_npos = atoi(value);
        else if (IS("num-negative-samples")) c->balancer_nneg = atoi(value);
        else if (IS("minimum-quality")) c->balancer_min_quality = atoi(value);
        else if (IS("minimum-qcpass-percent")) c->balancer_min_fraction_passes = atof(value) * 0.01;
        else UNKNOWN("balancer");
    } else if (SEC("polyA_seeder")) {
        if (IS("seed-trigger-polya-length")) c->seed_trigger = atoi(value);
        else if (IS("negative-sample-polya-length")) c->neg_polya_len = atoi(value);
        else if (IS("max-cctr-scan-left-space")) c->cctr_left = atoi(value);
        else if (IS("max-cctr-scan-right-space")) c->cctr_right = atoi(value);
        else if (IS("required-cdf-contrast")) c->required_cdf_contrast = atof(value);
        else if (IS("polya-boundary-pos")) c->boundary_pos = atoi(value);
        else if (IS("polya-sampling-gap")) c->sampling_gap = atoi(value);
        else if (IS("dist-sampling-bins")) c->bins = atoi(value);
        else if (IS("fair-sampling-fingerprint-length")) c->fair_fp_len = atoi(value);
        else if (IS("fair-sampling-hash-space-size")) c->fair_hash_size = atoi(value);
        else if (IS("fair-sampling-max-count")) c->fair_max_count = atoi(value) & 0xff;   /* uint8_t */
        else UNKNOWN("polyA_seeder");
    } else if (SEC("polyA_finder")) {
        if (strncasecmp(name, "polyA-weight-", 13) == 0 && weight_slot(name[13]) >= 0 && name[14] == '\0')
            c->w_polyA[weight_slot(name[13])] = atoi(value);
        else if (strncasecmp(name, "nonA-weight-", 12) == 0 && weight_slot(name[12]) >= 0 && name[13] == '\0')
            c->w_nonA[weight_slot(name[12])] = atoi(value);
        else if (IS("minimum-polya-length")) c->min_polya_length = atoi(value);
        else if (IS("maximum-modifications")) c->max_terminal_modifications = atoi(value);
        else if (IS("signal-analysis-trigger")) c->sigproc_trigger = atoi(value);
        else UNKNOWN("polyA_finder");
    } else if (SEC("polyA_ruler")) {
        if (IS("dark-cycles-threshold")) c->dark_cycles_threshold = (float)atof(value);
        else if (IS("maximum-dark-cycles")) c->max_dark_cycles = atoi(value);
        else if (IS("t-intensity-k")) c->t_intensity_k = (float)atof(value);
        else if (IS("t-intensity-center")) c->t_intensity_center = (float)atof(value);
        else UNKNOWN("polyA_ruler");
    } else if (strncasecmp(section, "sample:", 7) == 0) {
        return feed_sample(c, section + 7, name, value);
    } else {
        fprintf(stderr, "Unknown section [%s] in the configuration.", section);
        return -1;
    }
    return 0;
}

static void defaults(struct host_config *c)          /* parseconfig.c:482-534 */
{
    c->lane = c->tile = c->total_cycles = c->index_start = c->threep_start = c->threep_length =
        c->fivep_start = c->fivep_length = -1;
    c->threep_seqqual_output_length = 100;
    c->threads = 1;
    c->index_length = 6;
    c->read_buffer_size = 536870912;
    c->balancer_start = 0; c->balancer_length = 20;
    c->balancer_minimum_occurrence = 2; c->balancer_npos = 2; c->balancer_nneg = 4;
    c->balancer_min_quality = 25; c->balancer_min_fraction_passes = .70f;
    c->seed_trigger = 10; c->neg_polya_len = 0; c->cctr_left = 20; c->cctr_right = 20;
    c->required_cdf_contrast = .35f; c->boundary_pos = 120; c->sampling_gap = 3; c->bins = 1000;
    c->fair_fp_len = 30; c->fair_hash_size = 1048576; c->fair_max_count = 5;
    c->max_terminal_modifications = 20; c->min_polya_length = 5; c->sigproc_trigger = 10;
    /* A, C, G, T, N */
    c->w_polyA[0] = -9; c->w_polyA[1] = -9; c->w_polyA[2] = -9; c->w_polyA[3] = 2; c->w_polyA[4] = -1;
    c->w_nonA[0] = 0; c->w_nonA[1] = -4; c->w_nonA[2] = -4; c->w_nonA[3] = -1; c->w_nonA[4] = 0;
    c->dark_cycles_threshold = 10; c->max_dark_cycles = 5;
    c->t_intensity_k = 20.f; c->t_intensity_center = .75f;
}

static struct host_sample *pseudo_sample(const struct host_config *c, const char *name)
{
    struct host_sample *s = calloc(1, sizeof(*s));
    strncpy(s->name, name, TSK_NAME_MAX - 1);
    s->index = malloc(c->index_length + 1);
    memset(s->index, 'X', c->index_length);
    s->index[c->index_length] = '\0';
    s->delimiter = strdup("");
    s->delimiter_pos = -1;
    s->maximum_index_mismatches = c->index_length;
    s->maximum_delimiter_mismatches = -1;
    s->fingerprint = strdup("");
    return s;
}

static void derive(struct host_config *c)             /* parseconfig.c:537-671 */
{
    struct host_sample *s;
    int n = 0, longest_umi = 0;
    for (s = c->samples; s; s = s->next) {
        int total = 0;
        if (s->delimiter) s->delimiter_length = (int)strlen(s->delimiter);
        if (s->fingerprint) s->fingerprint_length = (int)strlen(s->fingerprint);
        for (int i = 0; i < s->umi_count; i++)
            if (s->umi[i].length > 0) total += s->umi[i].length;
        if (total > longest_umi) longest_umi = total;
    }
    if (c->control_name[0] != '\0') {
        s = pseudo_sample(c, c->control_name);
        s->next = c->samples;
        c->samples = s;
    }
    s = pseudo_sample(c, "Unknown");
    s->next = c->samples;
    c->samples = s;

    for (s = c->samples; s; s = s->next) {
        s->numindex = n++;
        s->signal_dump_length = c->threep_length;
        if (s->delimiter) s->signal_dump_length -= s->delimiter_pos + s->delimiter_length;
        if (s->limit_threep_processing > 0 && s->signal_dump_length > s->limit_threep_processing)
            s->signal_dump_length = s->limit_threep_processing;
        if (s->delimiter_pos >= 0) s->delimiter_pos += c->threep_start;
    }
    c->num_samples = n;
    if (c->threep_seqqual_output_length > c->threep_length)
        c->threep_seqqual_output_length = c->threep_length;
    {
        /* read blocks: same arithmetic as the reference, so multi-block tiles split identically */
        const size_t per_seqqual = (size_t)n * (10 + c->fivep_length * 2 + c->threep_seqqual_output_length * 2 + 6);
        const size_t per_taginfo = (size_t)n * (5 + 30 + c->max_terminal_modifications + longest_umi);
        const int per_entry = 2 * c->total_cycles + 8 * c->threep_length;
        const size_t wbuf = (size_t)c->threads * 512 * (per_seqqual + per_taginfo);
        c->read_buffer_entry_count = (int)((c->read_buffer_size - wbuf) / per_entry);
    }
}

struct host_config *tsk_parse_config(const char *filename)
{
    struct host_config *c = calloc(1, sizeof(*c));
    if (!c) return NULL;
    defaults(c);
    if (tsk_ini_parse(filename, feed, c) < 0) {
        fprintf(stderr, "Failed to parse %s.\n", filename);
        free(c);
        return NULL;
    }
    derive(c);
    return c;
}

void tsk_free_config(struct host_config *c)
{
    if (!c) return;
    free(c->datadir); free(c->laneid); free(c->threep_colormatrix_filename);
    free(c->seqqual_output); free(c->taginfo_output); free(c->signal_output);
    free(c->signal_dists_output); free(c->stats_output); free(c->length_dists_output);
    while (c->altcalls) {
        struct host_altcall *n = c->altcalls->next;
        free(c->altcalls->filename); free(c->altcalls);
        c->altcalls = n;
    }
    while (c->samples) {
        struct host_sample *n = c->samples->next;
        free(c->samples->index); free(c->samples->delimiter); free(c->samples->fingerprint);
        free(c->samples);
        c->samples = n;
    }
    free(c);
}

/* replace_placeholder(), src/utils.c:34-53 -- including its quirk: when the placeholder is absent
 * the result is the PLACEHOLDER text, not the format. */
char *tsk_replace_placeholder(const char *format, const char *old, const char *repl)
{
    const char *found = strstr(format, old);
    char *out;
    if (!found) return strdup(old);
    out = malloc(strlen(format) - strlen(old) + strlen(repl) + 1);
    memcpy(out, format, found - format);
    strcpy(out + (found - format), repl);
    strcat(out, found + strlen(old));
    return out;
}

static int copy_seq(char *dst, const char *src)
{
    if (!src) { dst[0] = '\0'; return 0; }
    if (strlen(src) >= TSK_MAX_SEQ) return -1;
    strcpy(dst, src);
    return 0;
}

int tsk_build_params(const struct host_config *c, tsk_params *p)
{
    const struct host_sample *s;
    int i;
    memset(p, 0, sizeof(*p));
    p->abi_version = TSK_ABI_VERSION;
    p->total_cycles = c->total_cycles;
    p->index_start = c->index_start; p->index_length = c->index_length;
    p->threep_start = c->threep_start; p->threep_length = c->threep_length;
    p->keep_no_delimiter = c->keep_no_delimiter;
    p->keep_low_quality_balancer = c->keep_low_quality_balancer;
    p->has_control = c->control_name[0] != '\0';
    p->control_first_cycle = c->control_first_cycle;
    p->control_read_length = c->control_read_length;
    p->balancer_start = c->balancer_start; p->balancer_length = c->balancer_length;
    p->balancer_min_quality = c->balancer_min_quality;
    p->balancer_min_bases_passes = (int)(c->balancer_length * c->balancer_min_fraction_passes);
    p->balancer_minimum_occurrence = c->balancer_minimum_occurrence;
    p->balancer_num_positive_samples = c->balancer_npos;
    p->balancer_num_negative_samples = c->balancer_nneg;
    p->seed_trigger_polya_length = c->seed_trigger;
    p->negative_sample_polya_length = c->neg_polya_len;
    p->max_cctr_scan_left_space = c->cctr_left; p->max_cctr_scan_right_space = c->cctr_right;
    p->required_cdf_contrast = c->required_cdf_contrast;
    p->polya_boundary_pos = c->boundary_pos; p->polya_sampling_gap = c->sampling_gap;
    p->dist_sampling_bins = c->bins;
    p->fair_sampling_fingerprint_length = c->fair_fp_len;
    p->fair_sampling_hash_space_size = c->fair_hash_size;
    p->fair_sampling_max_count = c->fair_max_count;
    for (i = 0; i < 5; i++) { p->weights_polyA[i] = c->w_polyA[i]; p->weights_nonA[i] = c->w_nonA[i]; }
    p->max_terminal_modifications = c->max_terminal_modifications;
    p->min_polya_length = c->min_polya_length;
    p->sigproc_trigger_polya_length = c->sigproc_trigger;
    p->dark_cycles_threshold = c->dark_cycles_threshold;
    p->max_dark_cycles = c->max_dark_cycles;
    {
        /* precalc_score_tables(), signalproc.c:54-82: evaluated with the HOST libm, never on the GPU */
        float prob = 1. / 4, ent = 0.f, k = c->t_intensity_k, center = c->t_intensity_center, v;
        for (i = 0; i < 4; i++) ent -= prob * logf(prob);
        p->maximum_entropy = ent;
        for (i = 0; i < TSK_T_SCORE_BINS; i++) {
            v = (1. * (.5 + i)) / TSK_T_SCORE_BINS;
            p->t_intensity_score[i] = 1. / (1. + expf(-k * (v - center)));
        }
        p->t_intensity_score[TSK_T_SCORE_BINS] = 1.f;
    }
    if (c->num_samples > TSK_MAX_SAMPLES) { fprintf(stderr, "Too many samples.\n"); return -1; }
    p->num_samples = c->num_samples;
    for (s = c->samples; s; s = s->next) {
        tsk_sample *o = &p->samples[s->numindex];
        if (copy_seq(o->index, s->index) || copy_seq(o->delimiter, s->delimiter) ||
            copy_seq(o->fingerprint, s->fingerprint)) {
            fprintf(stderr, "Sample %s: index/delimiter/fingerprint too long.\n", s->name);
            return -1;
        }
        o->maximum_index_mismatches = s->maximum_index_mismatches;
        o->delimiter_pos = s->delimiter_pos; o->delimiter_length = s->delimiter_length;
        o->maximum_delimiter_mismatches = s->maximum_delimiter_mismatches;
        o->fingerprint_pos = s->fingerprint_pos; o->fingerprint_length = s->fingerprint_length;
        o->maximum_fingerprint_mismatches = s->maximum_fingerprint_mismatches;
        o->umi_ranges_count = s->umi_count;
        o->limit_threep_processing = s->limit_threep_processing;
        o->signal_dump_length = s->signal_dump_length;
    }
    return 0;
}

/* What tsk_encode_configure() needs besides the parameters: which files each sample has (none for
 * the 'X...' pseudo-samples, tailseq-import.c:46-48), the 5' read and the UMI ranges. */
int tsk_build_encode_layout(const struct host_config *cfg, tsk_encode_layout *out)
{
    memset(out, 0, sizeof(*out));
    out->fivep_start = cfg->fivep_start;
    out->fivep_length = cfg->fivep_length;
    out->threep_seqqual_output_length = cfg->threep_seqqual_output_length;
    for (const struct host_sample *s = cfg->samples; s; s = s->next) {
        tsk_encode_sample *e;
        if (s->numindex < 0 || s->numindex >= TSK_MAX_SAMPLES) return -1;
        e = &out->samples[s->numindex];
        if (s->index && s->index[0] == 'X') continue;
        e->want_seqqual = 1;
        e->want_taginfo = cfg->taginfo_output != NULL;
        e->want_signal = 1;
        if (s->umi_count > TSK_ENCODE_MAX_UMI) return -1;
        e->umi_count = s->umi_count;
        for (int i = 0; i < s->umi_count; i++) { e->umi_start[i] = s->umi[i].start; e->umi_length[i] = s->umi[i].length; }
    }
    return 0;
}
