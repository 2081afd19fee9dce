/*
 * tsk_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference algorithm.
 *
 * A plain-C, single-threaded restatement of tailseeker 3.2.1's per-cluster signal path,
 * written against the behaviour spec in SURVEY.md appendix A and the reference sources
 * cited at each function.  It exists to CHECK the CUDA path: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.  The product
 * (tailseeker_b200/csrc) never links or calls it.
 *
 * Parity pin: this file is validated against the UNMODIFIED reference binaries built
 * into oracle/_ref/ (oracle/Makefile, oracle/refrun.py) on seeded synthetic tiles, and
 * against the golden outputs of those binaries committed under tests/golden/
 * (tests/test_oracle_vs_reference.py, tests/test_oracle_golden.py).  The reference
 * itself ships no tests or golden vectors.
 *
 * Build: gcc -O2 -ffp-contract=off (no FMA contraction, like the reference's
 * -O3 build for baseline x86-64, src/Makefile:3), host libm logf.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "tsk_oracle.h"

#define MAX_CYCLES 2048
#define MAX_RANK   16

/* ----------------------------------------------------------------------------------- */
/* base-call decoding: format_basecalls(), bclreader.c:146-166                          */
/* ----------------------------------------------------------------------------------- */
static void decode_read(const uint8_t *bcl, size_t pitch, uint32_t n, int ncycles,
                        char *seq, char *qual)
{
    static const char letters[4] = {'A', 'C', 'G', 'T'};
    for (int c = 0; c < ncycles; c++) {
        uint8_t b = bcl[(size_t)c * pitch + n];
        if (b == 0) {                 /* no-call: 'N', quality 2 */
            seq[c] = 'N';
            qual[c] = 2 + 33;
        } else {
            seq[c] = letters[b & 3];
            qual[c] = (char)((b >> 2) + 33);
        }
    }
    seq[ncycles] = qual[ncycles] = '\0';
}

/* ----------------------------------------------------------------------------------- */
/* assign_barcode(), spotanalyzer.c:67-103: unique minimum Hamming distance over every  */
/* list entry after "Unknown", accepted if within that entry's own mismatch budget.      */
/* ----------------------------------------------------------------------------------- */
static int pick_sample(const tsk_params *p, const char *indexread, int *mismatches)
{
    int best = -1, bestmm = p->index_length + 1, tie = 0;
    for (int s = 1; s < p->num_samples; s++) {
        int mm = 0;
        for (int k = 0; k < p->index_length; k++)
            mm += indexread[k] != p->samples[s].index[k];
        if (mm < bestmm) {
            best = s; bestmm = mm; tie = 0;
        } else if (mm == bestmm) {
            tie = 1;
        }
    }
    if (best < 0 || tie || bestmm > p->samples[best].maximum_index_mismatches) {
        *mismatches = -1;
        return -1;
    }
    *mismatches = bestmm;
    return best;
}

/* ----------------------------------------------------------------------------------- */
/* try_alignment_to_control(), controlaligner.c:107-145.                                 */
/*  1. fast path (:117-119): the first read_length characters of the WHOLE read (cycle 0  */
/*     onwards, not first_cycle -- SURVEY quirk 6) occur verbatim in either PhiX strand;   */
/*  2. otherwise Smith-Waterman of read[first_cycle .. +read_length) against               */
/*     forward + 20 x N + reverse (load_control_sequence, :74-104): match +1, mismatch -2, */
/*     any N 0 (:54-66), a gap of length l costs 4 + (l - 1) (ssw.c:226-233,              */
/*     tailseq-import.h:152-155); aligned when the best local score reaches                */
/*     (int)(read_length * 0.65) (:162-163, :135).                                         */
/* ssw.c's striped byte kernel returns the exact optimum here (scores <= read_length,      */
/* far from the 8-bit saturation point), so a plain Gotoh recurrence restates it.          */
/* ----------------------------------------------------------------------------------- */
#include "../tailseeker_b200/csrc/phix174.h"

#define CONTROL_SPACING 20
#define CONTROL_TARGET_LEN (2 * TSK_PHIX174_LENGTH + CONTROL_SPACING)

static char phix_fwd[TSK_PHIX174_LENGTH + 1], phix_rev[TSK_PHIX174_LENGTH + 1];
static int8_t phix_target[CONTROL_TARGET_LEN];

static void control_prepare(void)
{
    static const char letters[4] = {'A', 'C', 'G', 'T'};
    if (phix_fwd[0] != '\0') return;
    for (int i = 0; i < TSK_PHIX174_LENGTH; i++) {
        int b = (TSK_PHIX174_2BIT[i / 16] >> (2 * (i % 16))) & 3;
        phix_fwd[i] = letters[b];
        phix_rev[TSK_PHIX174_LENGTH - 1 - i] = letters[3 - b];
    }
    for (int i = 0; i < TSK_PHIX174_LENGTH; i++) {
        phix_target[i] = (int8_t)((TSK_PHIX174_2BIT[i / 16] >> (2 * (i % 16))) & 3);
        phix_target[2 * TSK_PHIX174_LENGTH + CONTROL_SPACING - 1 - i] = (int8_t)(3 - phix_target[i]);
    }
    for (int i = 0; i < CONTROL_SPACING; i++) phix_target[TSK_PHIX174_LENGTH + i] = 4;
}

static int contains(const char *hay, int haylen, const char *needle, int n)
{
    for (int i = 0; i + n <= haylen; i++)
        if (memcmp(hay + i, needle, (size_t)n) == 0) return 1;
    return 0;
}

int tsk_oracle_control_score(const char *read, int n)
{
    int h[TSK_MAX_CONTROL_READ + 1], e[TSK_MAX_CONTROL_READ + 1], best = 0;
    int8_t r[TSK_MAX_CONTROL_READ];
    control_prepare();
    for (int i = 0; i < n; i++)
        r[i] = read[i] == 'A' ? 0 : read[i] == 'C' ? 1 : read[i] == 'G' ? 2 : read[i] == 'T' ? 3 : 4;
    for (int i = 0; i <= n; i++) h[i] = e[i] = 0;
    for (int j = 0; j < CONTROL_TARGET_LEN; j++) {
        int diag = 0, f = 0, t = phix_target[j];
        for (int i = 1; i <= n; i++) {
            int sub = (t == 4 || r[i - 1] == 4) ? 0 : (t == r[i - 1] ? 1 : -2);
            int v = diag + sub;                    /* H(i-1, j-1) + s */
            int ee = e[i] - 1 > h[i] - 4 ? e[i] - 1 : h[i] - 4;      /* gap along the target */
            int ff = f - 1 > h[i - 1] - 4 ? f - 1 : h[i - 1] - 4;    /* gap along the read   */
            if (ee < 0) ee = 0;
            if (ff < 0) ff = 0;
            if (v < ee) v = ee;
            if (v < ff) v = ff;
            if (v < 0) v = 0;
            diag = h[i];
            h[i] = v; e[i] = ee; f = ff;
            if (v > best) best = v;
        }
    }
    return best;
}

int tsk_oracle_control_try(const char *seq, int first_cycle, int n)
{
    control_prepare();
    if (contains(phix_fwd, TSK_PHIX174_LENGTH, seq, n) || contains(phix_rev, TSK_PHIX174_LENGTH, seq, n))
        return 1;
    return tsk_oracle_control_score(seq + first_cycle, n) >= (int)(n * 0.65);
}

static int control_aligned(const tsk_params *p, const char *seq)
{
    return tsk_oracle_control_try(seq, p->control_first_cycle, p->control_read_length);
}

/* IUPAC sets, spotanalyzer.c:35-62 (note: a read 'N' matches nothing, not even N/X). */
static const char *iupac_set(char code)
{
    switch (code) {
    case 'A': return "A";    case 'B': return "CGT";  case 'C': return "C";
    case 'D': return "AGT";  case 'G': return "G";    case 'H': return "ACT";
    case 'K': return "GT";   case 'M': return "AC";   case 'N': return "GATC";
    case 'R': return "AG";   case 'S': return "CG";   case 'T': return "T";
    case 'U': return "T";    case 'V': return "ACG";  case 'W': return "AT";
    case 'X': return "GATC"; case 'Y': return "CT";
    default:  return "";
    }
}

/* find_delimiter_end_position(), spotanalyzer.c:106-152: offsets 0, -1, +1 in that order. */
static int locate_delimiter(const char *seq, const tsk_sample *s, int *flags)
{
    static const int shifts[3] = {0, -1, 1};
    for (int t = 0; t < 3; t++) {
        const char *at = seq + s->delimiter_pos + shifts[t];
        int ndiff = 0;
        for (int k = 0; k < s->delimiter_length; k++)
            ndiff += (strchr(iupac_set(s->delimiter[k]), at[k]) == NULL) || at[k] == '\0';
        if (ndiff <= s->maximum_delimiter_mismatches) {
            if (ndiff > 0) *flags |= TSK_PAFLAG_DELIMITER_HAS_MISMATCH;
            if (shifts[t] != 0) *flags |= TSK_PAFLAG_DELIMITER_IS_SHIFTED;
            return s->delimiter_pos + s->delimiter_length + shifts[t];
        }
    }
    *flags |= TSK_PAFLAG_DELIMITER_NOT_FOUND;
    return -1;
}

static int base_slot(char c)   /* A,C,G,T,N -> 0..4 (weights tables of tsk_params) */
{
    switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2;
                 case 'T': return 3; default: return 4; }
}

/* find_polya(), findpolya.c:35-84: exhaustive scan in (i asc, j asc) order with strict
 * improvement; returns (start << 16) | end, or 0. */
static uint32_t seed_polya(const tsk_params *p, const char *seq, int seqlen)
{
    int maxmod = p->max_terminal_modifications;
    int minlen = p->min_polya_length;
    int best_i = -1, best_j = -1, best_len = -1, best_score = -1;
    int prefix_penalty = 0;

    if (seqlen < maxmod)
        maxmod = seqlen;
    for (int i = 0; i < maxmod; i++) {
        int run = prefix_penalty + p->weights_polyA[base_slot(seq[i])];
        if (minlen <= 1 && run > best_score) {
            best_i = best_j = i; best_len = 1; best_score = run;
        }
        for (int j = i + 1; j < seqlen; j++) {
            run += p->weights_polyA[base_slot(seq[j])];
            if (run > best_score && j - i + 1 >= minlen) {
                best_i = i; best_j = j; best_len = j - i + 1; best_score = run;
            }
        }
        prefix_penalty += p->weights_nonA[base_slot(seq[i])];
    }
    if (best_len < minlen || best_score < 1)
        return 0;
    return ((uint32_t)best_i << 16) | (uint32_t)(best_j + 1);
}

/* decrosstalk_intensity(), signalproc.c:85-103: four products summed left to right. */
static void decrosstalk(float out[4], const int16_t *cif, size_t pitch, int r3cycle,
                        uint32_t n, const float *inv)
{
    const int16_t *plane = cif + (size_t)r3cycle * 4 * pitch + n;
    int16_t v0 = plane[0], v1 = plane[pitch], v2 = plane[2 * pitch], v3 = plane[3 * pitch];
    for (int j = 0; j < 4; j++) {
        const float *row = inv + 4 * j;
        out[j] = row[0] * v0 + row[1] * v1 + row[2] * v2 + row[3] * v3;
    }
}

/* check_balancer(): check_balance_minimum() + probe_signal_ranges(),
 * signalproc.c:106-129,132-201,204-243.  Returns -1 when the balancer is biased. */
static int probe_balancer(const tsk_params *p, const char *r3seq,
                          const int16_t *cif, size_t pitch, uint32_t n, const float *inv,
                          int balancer_len, float low[4], float bandwidth[4], int *flags)
{
    int seen[256];
    int npos = p->balancer_num_positive_samples, nneg = p->balancer_num_negative_samples;
    int from = p->balancer_start, to = p->balancer_start + balancer_len;
    float top[4][MAX_RANK], bottom[4][MAX_RANK];

    memset(seen, 0, sizeof(seen));
    for (int c = from; c < to; c++)
        seen[(unsigned char)r3seq[c]]++;
    if (seen['A'] < p->balancer_minimum_occurrence || seen['C'] < p->balancer_minimum_occurrence ||
        seen['G'] < p->balancer_minimum_occurrence || seen['T'] < p->balancer_minimum_occurrence) {
        *flags |= TSK_PAFLAG_BALANCER_BIASED;
        return -1;
    }

    for (int ch = 0; ch < 4; ch++) {
        for (int r = 0; r < npos; r++) top[ch][r] = -INFINITY;
        for (int r = 0; r < nneg; r++) bottom[ch][r] = INFINITY;
    }
    for (int c = from; c < to; c++) {
        float sig[4];
        decrosstalk(sig, cif, pitch, c, n, inv);
        for (int ch = 0; ch < 4; ch++) {
            float carry = sig[ch];          /* bubble into the sorted top-npos list */
            for (int r = 0; r < npos; r++)
                if (carry > top[ch][r]) { float t = top[ch][r]; top[ch][r] = carry; carry = t; }
            carry = sig[ch];                /* and into the sorted bottom-nneg list */
            for (int r = 0; r < nneg; r++)
                if (carry < bottom[ch][r]) { float t = bottom[ch][r]; bottom[ch][r] = carry; carry = t; }
        }
    }
    for (int ch = 0; ch < 4; ch++) {
        float acc = 0.f;
        for (int r = 0; r < nneg; r++) acc += bottom[ch][r];
        low[ch] = acc / nneg;
        acc = 0.f;
        for (int r = 0; r < npos; r++) acc += top[ch][r];
        bandwidth[ch] = acc / npos - low[ch];
    }
    return 0;
}

/* compute_polya_score() + normalize_signals() + shannon_entropy(),
 * signalproc.c:269-342,246-266,39-51.  Returns -1 on a dark-cycle abort. */
static int score_cycles(const tsk_params *p, const int16_t *cif, size_t pitch, uint32_t n,
                        const float *inv, int first_r3cycle, int ncycles,
                        const float low[4], const float bandwidth[4],
                        float *scores, char *downhill, int *flags)
{
    int ndark = 0;
    float prev_entropy = NAN;

    for (int c = 0; c < ncycles; c++) {
        float sig[4], norm[4], total;

        decrosstalk(sig, cif, pitch, first_r3cycle + c, n, inv);

        total = 0.f;
        for (int ch = 0; ch < 4; ch++)
            if (sig[ch] > 0.f) total += sig[ch];
        if (total < p->dark_cycles_threshold) {
            ndark++; scores[c] = prev_entropy = NAN; downhill[c] = 0;
            continue;
        }

        total = 0.f;
        for (int ch = 0; ch < 4; ch++) {
            norm[ch] = (sig[ch] - low[ch]) / bandwidth[ch];
            if (norm[ch] < 0.f) norm[ch] = 0.f;
            total += norm[ch];
        }
        for (int ch = 0; ch < 4; ch++)
            norm[ch] /= total;
        if (isnan(norm[0])) {
            ndark++; scores[c] = prev_entropy = NAN; downhill[c] = 0;
            continue;
        }

        float h = 0.f;
        for (int ch = 0; ch < 4; ch++)
            if (norm[ch] > 0.f) h -= norm[ch] * logf(norm[ch]);
        float entropy_score = 1.f - h / p->maximum_entropy;
        /* the reference indexes unchecked: (int)(norm[3] * 200) is undefined behaviour (and an
         * out-of-bounds read) when norm[3] is NaN/Inf, which needs a zero signal bandwidth
         * (SURVEY quirk 7).  Defined here, and in the CUDA kernels, as NaN -> 0 and a clamp to
         * [0, 200]; identical to the reference wherever the reference is defined. */
        float tpos = norm[3] * TSK_T_SCORE_BINS;
        int ti = (tpos >= 0.f) ? (tpos > (float)TSK_T_SCORE_BINS ? TSK_T_SCORE_BINS : (int)tpos) : 0;
        float tscore = p->t_intensity_score[ti];

        scores[c] = entropy_score * tscore;
        downhill[c] = (char)(entropy_score < prev_entropy);
        prev_entropy = entropy_score;
    }

    if (ndark > 0) {
        *flags |= TSK_PAFLAG_DARKCYCLE_EXISTS;
        if (ndark >= p->max_dark_cycles) {
            *flags |= TSK_PAFLAG_DARKCYCLE_OVER_THRESHOLD;
            return -1;
        }
    }
    return 0;
}

/* find_max_cumulative_contrast(), signalproc.c:378-412 (double cumsum; note the i / len-i
 * denominators and the first candidate at boundary left-1). */
static int max_contrast(const float *scores, int length, int left, int right, float *best_out)
{
    double cum[MAX_CYCLES], total = 0.;
    for (int i = 0; i < length; i++) { total += scores[i]; cum[i] = total; }
    if (left + right >= length)
        return -1;
    int at = left - 1;
    double best = cum[at] / at - (total - cum[at]) / (length - at);
    for (int i = left; i <= length - right; i++) {
        double v = cum[i] / i - (total - cum[i]) / (length - i);
        if (v > best) { best = v; at = i; }
    }
    *best_out = (float)best;
    return at + 1;
}

/* add_polya_score_sample(), spotanalyzer.c:441-457 */
static void hist_add(uint32_t *hist, const float *scores, int length, int firstcycle, int bins)
{
    for (int i = 0; i < length; i++) {
        if (isfinite(scores[i])) {
            int b = (int)(scores[i] * bins);
            if (b >= bins) b = bins - 1;
            hist[(size_t)(firstcycle + i) * bins + b]++;
        }
    }
}

/* check_sequence_for_fair_sampling(), spotanalyzer.c:460-489 */
static int fair_sampling_accept(const tsk_params *p, uint8_t *table, const char *seq)
{
    uint64_t v = 0;
    for (int k = 0; k < p->fair_sampling_fingerprint_length; k++)
        v = ((v << 3) | (((uint64_t)seq[k] & 7) ^ (v >> 60))) & ((1ULL << 63) - 1ULL);
    v %= (uint64_t)p->fair_sampling_hash_space_size;
    if (p->fair_sampling_max_count == 0 || table[v] < (uint8_t)p->fair_sampling_max_count) {
        table[v]++;
        return 1;
    }
    return 0;
}

/* write_polya_score(), spotanalyzer.c:389-438: header + packets, zero padded.
 * `downhill` is the UN-offset array (spotanalyzer.c:604-606). */
static void pack_record(uint32_t *dst, const float *scores, int length, const char *downhill,
                        uint32_t clusterno, int first_cycle, int dump_len)
{
    if (length > dump_len) length = dump_len;
    dst[0] = clusterno;
    dst[1] = ((uint32_t)(uint16_t)(int16_t)first_cycle) | ((uint32_t)(uint16_t)(int16_t)length << 16);
    memset(dst + 2, 0, sizeof(uint32_t) * (size_t)dump_len);
    for (int i = 0; i < length; i++) {
        if (!isnan(scores[i])) {
            unsigned s = 1 + (int)(scores[i] * (float)(TSK_SCORE_MAX - 1));
            if (s >= TSK_SCORE_MAX) s = TSK_SCORE_MAX;
            dst[2 + i] = (s << 1) | (uint32_t)(downhill[i] & 1);
        }
    }
}

struct pending_record { uint32_t cluster; int sample; size_t at; };

int tsk_oracle_process(const tsk_params *p,
                       const uint8_t *bcl, size_t bcl_pitch,
                       const int16_t *cif, size_t cif_pitch,
                       const float invmatrix[16],
                       uint32_t nclusters, uint32_t firstclusterno,
                       tsk_call *calls,
                       uint32_t *records, size_t capacity_words,
                       uint64_t *rec_offset, uint32_t *rec_count,
                       tsk_counters *counters,
                       uint32_t *pos, uint32_t *neg,
                       float *scores_out)
{
    const int T = p->total_cycles, ts = p->threep_start, tl = p->threep_length;
    const int bins = p->dist_sampling_bins;
    char seq[MAX_CYCLES + 1], qual[MAX_CYCLES + 1];
    float scores[MAX_CYCLES];
    char downhill[MAX_CYCLES];
    uint8_t *fair;
    uint32_t *scratch;          /* records in cluster order, regrouped per sample at the end */
    struct pending_record *pend;
    size_t npend = 0, used = 0;

    if (T > MAX_CYCLES || p->balancer_num_positive_samples > MAX_RANK ||
        p->balancer_num_negative_samples > MAX_RANK || p->num_samples > TSK_MAX_SAMPLES)
        return -1;

    fair = calloc((size_t)p->fair_sampling_hash_space_size, 1);
    scratch = malloc(sizeof(uint32_t) * (capacity_words ? capacity_words : 1));
    pend = malloc(sizeof(*pend) * (nclusters ? nclusters : 1));
    if (!fair || !scratch || !pend) { free(fair); free(scratch); free(pend); return -1; }
    memset(pos, 0, sizeof(uint32_t) * (size_t)T * bins);
    memset(neg, 0, sizeof(uint32_t) * (size_t)T * bins);
    for (int s = 0; s < p->num_samples; s++) rec_count[s] = 0;

    for (uint32_t n = 0; n < nclusters; n++) {
        tsk_call *call = &calls[n];
        int flags = 0, mismatches, sidx, delimiter_end = -1, status = -1, mods = -1;
        const tsk_sample *s;

        memset(call, 0, sizeof(*call));
        if (scores_out)
            for (int c = 0; c < tl; c++) scores_out[(size_t)n * tl + c] = NAN;

        decode_read(bcl, bcl_pitch, n, T, seq, qual);

        /* process_spots(), spotanalyzer.c:648-679 */
        sidx = pick_sample(p, seq + p->index_start, &mismatches);
        if (sidx < 0)            /* :650-666: Unknown (entry 0) or the control pseudo-sample (entry 1) */
            sidx = (p->has_control && control_aligned(p, seq)) ? 1 : 0;
        s = &p->samples[sidx];
        if (mismatches <= 0) counters[sidx].clusters_mm0++;
        else {
            flags |= TSK_PAFLAG_BARCODE_HAS_MISMATCHES;
            if (mismatches == 1) counters[sidx].clusters_mm1++;
            else counters[sidx].clusters_mm2plus++;
        }
        call->sample = (int16_t)sidx;

        /* fingerprint, spotanalyzer.c:682-691,155-170 */
        if (s->fingerprint_length > 0) {
            int mm = 0;
            for (int k = 0; s->fingerprint[k] != '\0'; k++)
                mm += s->fingerprint[k] != seq[s->fingerprint_pos + k];
            if (mm > s->maximum_fingerprint_mismatches) {
                counters[sidx].clusters_fpmismatch++;
                call->flags = (uint16_t)flags; call->delimiter_end = -1;
                call->polya_len = -1; call->terminal_mods = -1;
                continue;
            }
        }

        /* balancer base-call quality, spotanalyzer.c:697-700,173-203 -- indexes the
         * WHOLE-READ quality string with the balancer range. */
        if (s->umi_ranges_count > 0 && p->balancer_min_bases_passes > 0) {
            int pass = 0;
            for (int j = p->balancer_start; j < p->balancer_start + p->balancer_length; j++)
                pass += (qual[j] - 33) >= p->balancer_min_quality;
            if (pass < p->balancer_min_bases_passes) {
                flags |= TSK_PAFLAG_BALANCER_CALL_QUALITY_BAD;
                counters[sidx].clusters_qcfailed++;
                if (!p->keep_low_quality_balancer) {
                    call->flags = (uint16_t)flags; call->delimiter_end = -1;
                    call->polya_len = -1; call->terminal_mods = -1;
                    continue;
                }
            }
        }

        if (s->delimiter_length > 0) {
            delimiter_end = locate_delimiter(seq, s, &flags);
            if (delimiter_end < 0) {
                counters[sidx].clusters_nodelim++;
                if (!p->keep_no_delimiter) {
                    call->flags = (uint16_t)flags; call->delimiter_end = -1;
                    call->polya_len = -1; call->terminal_mods = -1;
                    continue;
                }
            } else {
                /* process_polya_signal(), spotanalyzer.c:492-614 */
                uint32_t hit = seed_polya(p, seq + delimiter_end, ts + tl - delimiter_end);
                int polya_start = (int)(hit >> 16), polya_len = (int)(hit & 0xffff) - polya_start;
                int balancer_len = delimiter_end - ts;
                float low[4], bandwidth[4];

                mods = polya_start;
                if (balancer_len > p->balancer_length) balancer_len = p->balancer_length;
                if (polya_start > 0) flags |= TSK_PAFLAG_HAVE_3P_MODIFICATION;
                if (polya_len > 0) flags |= TSK_PAFLAG_POLYA_DETECTED;

                status = -1;
                if (probe_balancer(p, seq + ts, cif, cif_pitch, n, invmatrix, balancer_len,
                                   low, bandwidth, &flags) == 0) {
                    int insert_len = ts + tl - delimiter_end;
                    if (s->limit_threep_processing > 0 && s->limit_threep_processing < insert_len)
                        insert_len = s->limit_threep_processing;

                    if (score_cycles(p, cif, cif_pitch, n, invmatrix, delimiter_end - ts,
                                     insert_len, low, bandwidth, scores, downhill, &flags) == 0) {
                        int scan_len = insert_len - polya_start;
                        status = polya_len;
                        if (scores_out)
                            memcpy(scores_out + (size_t)n * tl + (delimiter_end - ts), scores,
                                   sizeof(float) * (size_t)insert_len);
                        if (scan_len > 0) {
                            if (polya_len >= p->seed_trigger_polya_length) {
                                float contrast = 0.f;
                                int at = max_contrast(scores + polya_start, scan_len,
                                                      p->max_cctr_scan_left_space,
                                                      p->max_cctr_scan_right_space, &contrast);
                                if (at >= p->polya_boundary_pos &&
                                    contrast >= p->required_cdf_contrast) {
                                    int cap = p->polya_boundary_pos - p->polya_sampling_gap;
                                    int take = -polya_start + (scan_len < cap ? scan_len : cap);
                                    if (take > 0)
                                        hist_add(pos, scores + polya_start, take,
                                                 delimiter_end + polya_start, bins);
                                }
                            } else if (polya_len <= p->negative_sample_polya_length &&
                                       scan_len >= p->fair_sampling_fingerprint_length &&
                                       fair_sampling_accept(p, fair,
                                                            seq + delimiter_end + polya_start)) {
                                hist_add(neg, scores + polya_start, scan_len,
                                         delimiter_end + polya_start, bins);
                            }
                            if (polya_len >= p->sigproc_trigger_polya_length) {
                                size_t words = 2 + (size_t)s->signal_dump_length;
                                if (used + words > capacity_words) {
                                    free(fair); free(scratch); free(pend);
                                    return -1;
                                }
                                pack_record(scratch + used, scores + polya_start, scan_len,
                                            downhill, firstclusterno + n,
                                            delimiter_end + polya_start, s->signal_dump_length);
                                pend[npend].cluster = n; pend[npend].sample = sidx;
                                pend[npend].at = used;
                                npend++; used += words;
                                call->emitted = 1;
                                call->record_slot = rec_count[sidx]++;
                            }
                        }
                    }
                }
            }
        }

        call->flags = (uint16_t)flags;
        call->delimiter_end = (int16_t)delimiter_end;
        call->polya_len = (int16_t)status;
        call->terminal_mods = (int16_t)mods;
        call->written = 1;
    }

    /* regroup: sample after sample, cluster order inside */
    {
        uint64_t at = 0;
        for (int s = 0; s < p->num_samples; s++) {
            size_t words = 2 + (size_t)p->samples[s].signal_dump_length;
            rec_offset[s] = at;
            for (size_t k = 0; k < npend; k++)
                if (pend[k].sample == s) {
                    memcpy(records + at, scratch + pend[k].at, sizeof(uint32_t) * words);
                    at += words;
                }
        }
    }
    free(fair); free(scratch); free(pend);
    return 0;
}

/* ----------------------------------------------------------------------------------- */
/* inverse_4x4_matrix(), contrib/misc.c:11-58                                           */
/* ----------------------------------------------------------------------------------- */
static float cofactor_entry(int i, int j, const float *m)
{
    /* six signed triple products of the 3x3 minor, walked with wrap-around offsets
     * (a = column offset, b = row offset) in the reference's term order */
    static const signed char term[6][6] = {
        {+1, -1, 0, 0, -1, +1}, {+1, +1, 0, -1, -1, 0}, {-1, -1, +1, 0, 0, +1},
        {-1, -1, 0, 0, +1, +1}, {-1, +1, 0, -1, +1, 0}, {+1, -1, -1, 0, 0, +1},
    };
    int o = 2 + (j - i);
    int ci = i + 4 + o, rj = j + 4 - o;
    float acc = 0.f;
    for (int t = 0; t < 6; t++) {
        float prod = m[((rj + term[t][1]) % 4) * 4 + ((ci + term[t][0]) % 4)] *
                     m[((rj + term[t][3]) % 4) * 4 + ((ci + term[t][2]) % 4)] *
                     m[((rj + term[t][5]) % 4) * 4 + ((ci + term[t][4]) % 4)];
        if (t == 0) acc = prod;
        else if (t < 3) acc = acc + prod;
        else acc = acc - prod;
    }
    return (o % 2) ? acc : -acc;
}

int tsk_oracle_invert_matrix(const float m[16], float out[16])
{
    float adj[16], det = 0.f;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            adj[j * 4 + i] = cofactor_entry(i, j, m);
    for (int k = 0; k < 4; k++)
        det += m[k] * adj[k * 4];
    if (det == 0)
        return -1;
    det = (float)(1. / det);
    for (int i = 0; i < 16; i++)
        out[i] = adj[i] * det;
    return 0;
}

/* ----------------------------------------------------------------------------------- */
/* ruler, polyaruler.c                                                                  */
/* ----------------------------------------------------------------------------------- */
uint32_t tsk_oracle_cutoff_to_packed(double cutoff)
{
    /* polyaruler.c:93-96.  (int)NaN is 0x80000000 on x86-64, so "nan" becomes
     * 0x80000001 and is clamped to SCORE_MAX; spell that out instead of
     * relying on UB. */
    uint32_t v;
    double scaled = cutoff * (double)(TSK_SCORE_MAX - 1);
    if (!(scaled > -2147483649.0 && scaled < 2147483648.0))
        v = 1u + 0x80000000u;
    else
        v = 1u + (uint32_t)(int)scaled;
    if (v >= TSK_SCORE_MAX) v = TSK_SCORE_MAX;
    return v;
}

/* measure_polya_length(), polyaruler.c:104-152 */
static int ruler_length(const uint32_t *packets, int valid, int first_cycle,
                        const uint32_t *cutoffs, int ncutoffs, float weight)
{
    int cum = 0, best = 0, argmax = -1, ext = -1;
    for (int i = 0, cyc = first_cycle; i < valid && cyc < ncutoffs; i++, cyc++) {
        uint32_t cut = cutoffs[cyc];
        if (cut == 0) continue;
        cum += (((packets[i] >> 1) & TSK_SCORE_MAX) >= cut) ? 1 : -1;
        if (cum > best) { best = cum; argmax = i; }
    }
    if (argmax >= 0)
        for (int i = argmax + 1; i < valid; i++) {
            if (((packets[i] >> 1) & TSK_SCORE_MAX) == 0) continue;   /* dark */
            if (packets[i] & 1) ext = i; else break;
        }
    if (ext >= 0)
        return argmax + 1 + (int)roundf((ext - argmax) * weight);
    return argmax + 1;
}

int tsk_oracle_ruler(const uint32_t *records, size_t nrecords, int dump_len,
                     const uint32_t *cutoffs, int ncutoffs,
                     int minimum_polya_len, float downhill_ext_weight,
                     int bins, float gap,
                     int16_t *polya_len_out, uint32_t *pos)
{
    size_t words = 2 + (size_t)dump_len;
    for (size_t r = 0; r < nrecords; r++) {
        const uint32_t *rec = records + r * words;
        int first_cycle = (int16_t)(rec[1] & 0xffff), valid = (int16_t)(rec[1] >> 16);
        int len = ruler_length(rec + 2, valid, first_cycle, cutoffs, ncutoffs,
                               downhill_ext_weight);
        polya_len_out[r] = -1;
        if (len >= minimum_polya_len) {
            int take = len - (int)(len * gap);
            polya_len_out[r] = (int16_t)len;
            /* polyaruler.c:154-172 */
            for (int i = 0; i < take; i++) {
                uint32_t sc = (rec[2 + i] >> 1) & TSK_SCORE_MAX;
                if (sc > 0) {
                    int b = (int)(((sc - 1.) / (TSK_SCORE_MAX - 1)) * bins);
                    if (b >= bins) b = bins - 1;
                    pos[(size_t)(first_cycle + i) * bins + b]++;
                }
            }
        }
    }
    return 0;
}

float tsk_oracle_logf(float x) { return logf(x); }
