/*
 * tsk_capi.cu -- the extern "C" boundary of libtailseeker_b200.so (include/tailseeker_b200.h).
 * Host-side plumbing only: parameter validation/upload, HBM buffers, stream, result copies.
 */
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tsk_internal.h"

static thread_local char g_err[512] = "";

void tsk_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *tsk_last_error(void) { return g_err; }

extern "C" int tsk_probe(int device)
{
    int count = 0;
    cudaDeviceProp prop;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
        tsk_set_error("no CUDA device %d (found %d); this library has no CPU path", device, count);
        return -1;
    }
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) {
        tsk_set_error("device %d is not an sm_100-class GPU; kernels are built for sm_100a only", device);
        return -1;
    }
    return TSK_ABI_VERSION;
}

/* ---- parameter translation ----------------------------------------------------------- */
static int base_code_of(char c)
{
    switch (c) {
    case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3;
    case 'N': return 4; default: return 7;
    }
}

static int iupac_mask(char c)    /* bits over A,C,G,T; spotanalyzer.c:35-62 */
{
    switch (c) {
    case 'A': return 1;  case 'C': return 2;  case 'G': return 4;  case 'T': case 'U': return 8;
    case 'B': return 14; case 'D': return 13; case 'H': return 11; case 'V': return 7;
    case 'K': return 12; case 'M': return 3;  case 'R': return 5;  case 'S': return 6;
    case 'W': return 9;  case 'Y': return 10; case 'N': case 'X': return 15;
    default:  return 0;
    }
}

static int build_device_params(const tsk_params *p, DevParams *d, DevSample *ds)
{
    if (p->abi_version != TSK_ABI_VERSION) { tsk_set_error("tsk_params.abi_version mismatch"); return -1; }
    if (p->num_samples < 1 || p->num_samples > TSK_MAX_SAMPLES) { tsk_set_error("bad num_samples"); return -1; }
    if (p->total_cycles < 1 || p->total_cycles > TSK_MAX_CUTOFF_CYCLES) { tsk_set_error("bad total_cycles"); return -1; }
    if (p->threep_start < 0 || p->threep_length < 1 ||
        p->threep_start + p->threep_length > p->total_cycles) { tsk_set_error("3' read outside the run"); return -1; }
    if (p->index_length < 0 || p->index_length >= TSK_MAX_SEQ || p->index_start < 0 ||
        p->index_start + p->index_length > p->total_cycles) { tsk_set_error("bad index read"); return -1; }
    if (p->balancer_start != 0) {
        /* the reference indexes a balancer_len-long array with absolute balancer.start
         * (signalproc.c:152-156 vs spotanalyzer.c:534-536): only start = 1 is defined */
        tsk_set_error("balancer start != 1 is undefined behaviour in the reference; refused");
        return -1;
    }
    if (p->balancer_length < 0 || p->balancer_length > p->threep_length ||
        p->balancer_length > p->total_cycles) { tsk_set_error("bad balancer length"); return -1; }
    if (p->balancer_num_positive_samples < 1 || p->balancer_num_positive_samples > TSK_MAX_RANK ||
        p->balancer_num_negative_samples < 1 || p->balancer_num_negative_samples > TSK_MAX_RANK) {
        tsk_set_error("balancer num-positive/negative-samples must be 1..%d", TSK_MAX_RANK); return -1;
    }
    if (p->max_terminal_modifications < 0 || p->max_terminal_modifications >= TSK_MAX_MODS) {
        tsk_set_error("maximum-modifications must be < %d", TSK_MAX_MODS); return -1;
    }
    if (p->dist_sampling_bins < 1 || p->dist_sampling_bins > 65536) { tsk_set_error("bad dist-sampling-bins"); return -1; }
    if (p->fair_sampling_hash_space_size < 1) { tsk_set_error("bad fair-sampling-hash-space-size"); return -1; }
    if (p->fair_sampling_fingerprint_length < 0) { tsk_set_error("bad fair-sampling-fingerprint-length"); return -1; }
    if (p->max_cctr_scan_left_space < 1 || p->max_cctr_scan_right_space < 1) { tsk_set_error("bad cctr scan space"); return -1; }
    /* has_control: list entry 1 is the control pseudo-sample ("XXXXXX", no files); the
     * barcode-unassigned clusters are classified by tsk_control.cu (controlaligner.c:107-145).
     * The reference reads control_read_length characters from cycle 0 and from first_cycle. */
    if (p->has_control) {
        if (p->num_samples < 2) { tsk_set_error("has_control needs the control pseudo-sample as list entry 1"); return -1; }
        if (p->control_read_length < 8 || p->control_read_length > TSK_MAX_CONTROL_READ) {
            tsk_set_error("phix-match-length must be 8..%d", TSK_MAX_CONTROL_READ); return -1;
        }
        if (p->control_first_cycle < 0 || p->control_first_cycle + p->control_read_length > p->total_cycles) {
            tsk_set_error("control window outside the read"); return -1;
        }
    }

    memset(d, 0, sizeof(*d));
    d->total_cycles = p->total_cycles; d->index_start = p->index_start; d->index_length = p->index_length;
    d->threep_start = p->threep_start; d->threep_length = p->threep_length;
    d->keep_no_delimiter = p->keep_no_delimiter; d->keep_low_quality_balancer = p->keep_low_quality_balancer;
    d->balancer_start = p->balancer_start; d->balancer_length = p->balancer_length;
    d->balancer_min_quality = p->balancer_min_quality;
    d->balancer_min_bases_passes = p->balancer_min_bases_passes;
    d->balancer_minimum_occurrence = p->balancer_minimum_occurrence;
    d->npos = p->balancer_num_positive_samples; d->nneg = p->balancer_num_negative_samples;
    d->seed_trigger = p->seed_trigger_polya_length; d->neg_polya_len = p->negative_sample_polya_length;
    d->cctr_left = p->max_cctr_scan_left_space; d->cctr_right = p->max_cctr_scan_right_space;
    d->required_contrast = p->required_cdf_contrast;
    d->boundary_pos = p->polya_boundary_pos; d->sampling_gap = p->polya_sampling_gap;
    d->bins = p->dist_sampling_bins;
    d->fair_fp_len = p->fair_sampling_fingerprint_length;
    d->fair_hash_size = p->fair_sampling_hash_space_size;
    d->fair_max_count = p->fair_sampling_max_count & 0xff;
    for (int i = 0; i < 5; i++) { d->w_polyA[i] = p->weights_polyA[i]; d->w_nonA[i] = p->weights_nonA[i]; }
    d->w_small = 1;
    for (int i = 0; i < 5; i++) d->w_small = d->w_small && p->weights_polyA[i] >= -128 && p->weights_polyA[i] <= 127;
    d->max_mods = p->max_terminal_modifications; d->min_polya_len = p->min_polya_length;
    d->sigproc_trigger = p->sigproc_trigger_polya_length;
    d->dark_threshold = p->dark_cycles_threshold; d->max_dark_cycles = p->max_dark_cycles;
    d->maximum_entropy = p->maximum_entropy;
    d->num_samples = p->num_samples;
    d->has_control = p->has_control != 0;
    d->index_words = (p->index_length + 9) / 10;
    if (d->index_words > 4) { tsk_set_error("index too long"); return -1; }

    for (int s = 0; s < p->num_samples; s++) {
        const tsk_sample *hs = &p->samples[s];
        DevSample *o = &ds[s];
        memset(o, 0, sizeof(*o));
        if ((int)strnlen(hs->index, TSK_MAX_SEQ) != p->index_length) {
            tsk_set_error("sample %d: index length differs from index-length", s); return -1;
        }
        for (int k = 0; k < p->index_length; k++)
            o->index_pk[k / 10] |= (uint32_t)(base_code_of(hs->index[k]) & 7) << (3 * (k % 10));
        if (hs->delimiter_length < 0 || hs->delimiter_length >= TSK_MAX_SEQ ||
            hs->fingerprint_length < 0 || hs->fingerprint_length >= TSK_MAX_SEQ) {
            tsk_set_error("sample %d: delimiter/fingerprint too long", s); return -1;
        }
        if (hs->delimiter_length > 0 &&
            (hs->delimiter_pos < 1 || hs->delimiter_pos + hs->delimiter_length + 1 > p->total_cycles ||
             hs->delimiter_pos + hs->delimiter_length - 1 < p->threep_start)) {
            tsk_set_error("sample %d: delimiter window outside the read", s); return -1;
        }
        if (hs->fingerprint_length > 0 &&
            (hs->fingerprint_pos < 0 || hs->fingerprint_pos + hs->fingerprint_length > p->total_cycles)) {
            tsk_set_error("sample %d: fingerprint outside the read", s); return -1;
        }
        for (int k = 0; k < hs->delimiter_length; k++) o->delim_mask[k] = (uint8_t)iupac_mask(hs->delimiter[k]);
        for (int k = 0; k < hs->fingerprint_length; k++) o->fingerprint[k] = (uint8_t)base_code_of(hs->fingerprint[k]);
        o->max_index_mm = (int16_t)hs->maximum_index_mismatches;
        o->delimiter_pos = (int16_t)hs->delimiter_pos; o->delimiter_length = (int16_t)hs->delimiter_length;
        o->max_delim_mm = (int16_t)hs->maximum_delimiter_mismatches;
        o->fp_pos = (int16_t)hs->fingerprint_pos; o->fp_len = (int16_t)hs->fingerprint_length;
        o->max_fp_mm = (int16_t)hs->maximum_fingerprint_mismatches;
        o->has_umi = (int16_t)(hs->umi_ranges_count > 0);
        o->limit_threep = (int16_t)hs->limit_threep_processing;
        if (hs->signal_dump_length < 0 || hs->signal_dump_length > p->threep_length + 1) {
            tsk_set_error("sample %d: bad signal_dump_length", s); return -1;
        }
        o->dump_len = (int16_t)hs->signal_dump_length;
    }
    return 0;
}

/* ---- per-stage timing without host waits ------------------------------------------------- */
static void prof_destroy(tsk_ctx *ctx)
{
    for (int i = 0; i < TSK_PROF_RING; i++)
        for (int k = 0; k < 5; k++) cudaEventDestroy(ctx->prof[i].ev[k]);
}

/* adds the oldest pending event set to the stage totals (waits for it if it has not finished) */
static int prof_harvest_one(tsk_ctx *ctx)
{
    tsk_ctx::ProfSlot &p = ctx->prof[ctx->prof_head];
    TSK_CUDA(cudaEventSynchronize(p.ev[p.nev - 1]));
    for (int i = 0; i + 1 < p.nev; i++) {
        float ms = 0.f;
        TSK_CUDA(cudaEventElapsedTime(&ms, p.ev[i], p.ev[i + 1]));
        ctx->stage_ms[p.first_stage + i] += ms;
        ctx->stage_runs[p.first_stage + i]++;
    }
    ctx->prof_head = (ctx->prof_head + 1) % TSK_PROF_RING;
    ctx->prof_count--;
    return 0;
}

static int prof_harvest_all(tsk_ctx *ctx)
{
    while (ctx->prof_count > 0)
        if (prof_harvest_one(ctx) != 0) return -1;
    return 0;
}

/* next free event set (NULL when timing is off); the caller records ev[0..nev-1] in order */
static cudaEvent_t *prof_take(tsk_ctx *ctx, int nev, int first_stage)
{
    if (!ctx->profile) return NULL;
    if (ctx->prof_count == TSK_PROF_RING && prof_harvest_one(ctx) != 0) return NULL;
    tsk_ctx::ProfSlot &p = ctx->prof[(ctx->prof_head + ctx->prof_count) % TSK_PROF_RING];
    p.nev = nev; p.first_stage = first_stage;
    ctx->prof_count++;
    return p.ev;
}

/* ---- context --------------------------------------------------------------------------- */
extern "C" int tsk_ctx_create(tsk_ctx **out, int device, const tsk_params *params)
{
    if (!out || !params) { tsk_set_error("null argument"); return -1; }
    *out = NULL;
    if (tsk_probe(device) < 0) return -1;
    DevSample hs[TSK_MAX_SAMPLES];
    tsk_ctx *ctx = (tsk_ctx *)calloc(1, sizeof(tsk_ctx));
    if (!ctx) { tsk_set_error("out of memory"); return -1; }
    if (build_device_params(params, &ctx->dp, hs) != 0) { free(ctx); return -1; }
    ctx->device = device;
    ctx->hp = *params;
    ctx->maxdump = 0;
    for (int s = 0; s < params->num_samples; s++)
        if (params->samples[s].signal_dump_length > ctx->maxdump) ctx->maxdump = params->samples[s].signal_dump_length;

    TSK_CUDA(cudaSetDevice(device));
    TSK_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->own_stream = 1;
    TSK_CUDA(cudaStreamCreateWithFlags(&ctx->cstream, cudaStreamNonBlocking));
    TSK_CUDA(cudaEventCreateWithFlags(&ctx->ev_up[0], cudaEventDisableTiming));
    TSK_CUDA(cudaEventCreateWithFlags(&ctx->ev_up[1], cudaEventDisableTiming));
    TSK_CUDA(cudaEventCreateWithFlags(&ctx->ev_done[0], cudaEventDisableTiming));
    TSK_CUDA(cudaEventCreateWithFlags(&ctx->ev_done[1], cudaEventDisableTiming));
    ctx->tile_slot = -1;
    const int S = params->num_samples;
    const size_t histbytes = sizeof(uint32_t) * (size_t)params->total_cycles * params->dist_sampling_bins;
    TSK_CUDA(cudaMalloc(&ctx->d_samples, sizeof(DevSample) * S));
    TSK_CUDA(cudaMemcpy(ctx->d_samples, hs, sizeof(DevSample) * S, cudaMemcpyHostToDevice));
    TSK_CUDA(cudaMalloc(&ctx->d_tscore, sizeof(float) * (TSK_T_SCORE_BINS + 1)));
    TSK_CUDA(cudaMemcpy(ctx->d_tscore, params->t_intensity_score, sizeof(float) * (TSK_T_SCORE_BINS + 1),
                        cudaMemcpyHostToDevice));
    {
        double rcp[288];
        rcp[0] = 0.;
        for (int i = 1; i < 288; i++) rcp[i] = 1.0 / (double)i;     /* IEEE RN on the host */
        TSK_CUDA(cudaMalloc(&ctx->d_rcp, sizeof(rcp)));
        TSK_CUDA(cudaMemcpy(ctx->d_rcp, rcp, sizeof(rcp), cudaMemcpyHostToDevice));
    }
    if (tsk_score_setup(ctx) != 0) return -1;
    TSK_CUDA(cudaMalloc(&ctx->d_pos, histbytes));
    TSK_CUDA(cudaMalloc(&ctx->d_neg, histbytes));
    TSK_CUDA(cudaMemset(ctx->d_pos, 0, histbytes));
    TSK_CUDA(cudaMemset(ctx->d_neg, 0, histbytes));
    TSK_CUDA(cudaMalloc(&ctx->d_counters, sizeof(uint32_t) * S * TSK_NCOUNT));
    TSK_CUDA(cudaMemset(ctx->d_counters, 0, sizeof(uint32_t) * S * TSK_NCOUNT));
    TSK_CUDA(cudaMalloc(&ctx->d_fair, sizeof(uint32_t) * (size_t)params->fair_sampling_hash_space_size));
    TSK_CUDA(cudaMalloc(&ctx->d_totals, sizeof(uint32_t) * (TSK_MAX_SAMPLES + 1)));
    TSK_CUDA(cudaMalloc(&ctx->d_recbase, sizeof(uint64_t) * (TSK_MAX_SAMPLES + 1)));
    TSK_CUDA(cudaMalloc(&ctx->d_cutoffs, sizeof(uint32_t) * TSK_MAX_CUTOFF_CYCLES));
    TSK_CUDA(cudaStreamCreateWithFlags(&ctx->rstream, cudaStreamNonBlocking));
    TSK_CUDA(cudaEventCreateWithFlags(&ctx->ev_rul[0], cudaEventDisableTiming));
    TSK_CUDA(cudaEventCreateWithFlags(&ctx->ev_rul[1], cudaEventDisableTiming));
    TSK_CUDA(cudaMalloc(&ctx->d_rtotals, sizeof(uint32_t) * (TSK_MAX_SAMPLES + 1)));
    TSK_CUDA(cudaMalloc(&ctx->d_rrecbase, sizeof(uint64_t) * (TSK_MAX_SAMPLES + 1)));
    TSK_CUDA(cudaMalloc(&ctx->d_ruler_called, sizeof(unsigned long long)));
    TSK_CUDA(cudaMemset(ctx->d_ruler_called, 0, sizeof(unsigned long long)));
    if (params->has_control) {
        if (tsk_control_setup(ctx) != 0) return -1;
        TSK_CUDA(cudaMalloc(&ctx->d_ctl_queue, sizeof(uint32_t) * 4));
        TSK_CUDA(cudaStreamCreateWithFlags(&ctx->xstream, cudaStreamNonBlocking));
        TSK_CUDA(cudaEventCreateWithFlags(&ctx->ev_ctl[0], cudaEventDisableTiming));
        TSK_CUDA(cudaEventCreateWithFlags(&ctx->ev_ctl[1], cudaEventDisableTiming));
    }
    *out = ctx;
    return 0;
}

extern "C" int tsk_ctx_destroy(tsk_ctx *ctx)
{
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    tsk_encode_free(ctx);
    tsk_ingest_free(ctx);
    if (ctx->rstream) {
        cudaStreamSynchronize(ctx->rstream);
        cudaStreamDestroy(ctx->rstream);
        cudaEventDestroy(ctx->ev_rul[0]); cudaEventDestroy(ctx->ev_rul[1]);
    }
    void *ptrs[] = {ctx->d_logki, ctx->d_rtotals, ctx->d_rrecbase, ctx->d_samples, ctx->d_tscore, ctx->d_rcp, ctx->d_bcl_own[0], ctx->d_cif_own[0], ctx->d_bcl_own[1],
                    ctx->d_cif_own[1], ctx->d_calls, ctx->d_kind,
                    ctx->d_bucket, ctx->d_worklist, ctx->d_blkcnt, ctx->d_totals, ctx->d_recbase,
                    ctx->d_records, ctx->d_scratch, ctx->d_pos, ctx->d_neg, ctx->d_counters, ctx->d_fair, ctx->d_fairslots,
                    ctx->d_lists, ctx->d_ruler_pos, ctx->d_cutoffs, ctx->d_rlen, ctx->d_rrec,
                    ctx->d_fit_hist, ctx->d_fit_out, ctx->d_ctl_queue, ctx->d_ctl_index, ctx->d_ruler_called};
    for (size_t i = 0; i < sizeof(ptrs) / sizeof(ptrs[0]); i++)
        if (ptrs[i]) cudaFree(ptrs[i]);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    cudaStreamDestroy(ctx->cstream);
    if (ctx->xstream) { cudaStreamDestroy(ctx->xstream); cudaEventDestroy(ctx->ev_ctl[0]); cudaEventDestroy(ctx->ev_ctl[1]); }
    cudaEventDestroy(ctx->ev_up[0]); cudaEventDestroy(ctx->ev_up[1]);
    cudaEventDestroy(ctx->ev_done[0]); cudaEventDestroy(ctx->ev_done[1]);
    if (ctx->profile) prof_destroy(ctx);
    free(ctx);
    return 0;
}

template <typename T>
static int ensure(T **ptr, size_t *cap, size_t need)
{
    if (*cap >= need && *ptr) return 0;
    if (*ptr) cudaFree(*ptr);
    *ptr = NULL; *cap = 0;
    TSK_CUDA(cudaMalloc(ptr, need * sizeof(T)));
    *cap = need;
    return 0;
}

static int ensure_cluster_buffers(tsk_ctx *ctx, uint32_t n)
{
    if (n <= ctx->cap_clusters && ctx->d_calls) return 0;
    /* round the capacity up so consecutive tiles of similar size do not reallocate */
    uint32_t cap = (n + 4095u) & ~4095u;
    if (cap < 4096) cap = 4096;
    const int S = ctx->dp.num_samples;
    void *old[] = {ctx->d_calls, ctx->d_kind, ctx->d_bucket, ctx->d_worklist, ctx->d_blkcnt,
                   ctx->d_records, ctx->d_scratch, ctx->d_lists};
    for (size_t i = 0; i < sizeof(old) / sizeof(old[0]); i++) if (old[i]) cudaFree(old[i]);
    ctx->d_calls = NULL; ctx->d_kind = NULL; ctx->d_bucket = NULL; ctx->d_worklist = NULL;
    ctx->d_blkcnt = NULL; ctx->d_records = NULL; ctx->d_scratch = NULL; ctx->d_lists = NULL;
    ctx->cap_clusters = 0;
    ctx->processed = 0;          /* the previous tile's result buffers are gone */
    const size_t nblk = (cap + TSK_K1_THREADS - 1) / TSK_K1_THREADS;
    const int maxinsert = ctx->dp.threep_length + 1;
    TSK_CUDA(cudaMalloc(&ctx->d_calls, sizeof(tsk_call) * (size_t)cap));
    TSK_CUDA(cudaMalloc(&ctx->d_kind, (size_t)cap));
    TSK_CUDA(cudaMalloc(&ctx->d_bucket, sizeof(uint32_t) * (size_t)cap));
    TSK_CUDA(cudaMalloc(&ctx->d_worklist, sizeof(uint32_t) * (size_t)cap));
    TSK_CUDA(cudaMalloc(&ctx->d_blkcnt, sizeof(uint32_t) * nblk * (S + 1)));
    ctx->rec_cap_words = (size_t)cap * (size_t)(2 + ctx->maxdump);
    TSK_CUDA(cudaMalloc(&ctx->d_records, sizeof(uint32_t) * ctx->rec_cap_words));
    ctx->scratch_cap = (size_t)cap * (size_t)maxinsert;
    TSK_CUDA(cudaMalloc(&ctx->d_scratch, sizeof(float) * 2 * ctx->scratch_cap));   /* second half: handed-over clusters */
    TSK_CUDA(cudaMalloc(&ctx->d_lists, sizeof(uint32_t) * (16 + 6 * (size_t)cap)));      /* tsk_score.cu: six arrays */
    ctx->cap_clusters = cap;
    return 0;
}

/* ---- upload slots (shared by tsk_tile_upload and the row-wise ingestion of tsk_ingest.cu) ---- */
/* Takes the next upload slot for a tile of `nclusters`: sizes the slot's HBM arrays (rows padded to
 * 128 clusters), makes the copy stream wait for the kernels of the tile that used the slot before,
 * and fills the pending-tile descriptor.  The tile becomes visible to tsk_tile_process() with
 * tsk_upload_slot_commit(). */
int tsk_upload_slot_acquire(tsk_ctx *ctx, uint32_t nclusters, uint32_t firstclusterno, const float invmatrix[16],
                            int *slot_out, size_t *bpitch_out, size_t *cpitch_out)
{
    if (nclusters >= 0x7fffffffu) { tsk_set_error("too many clusters in one block"); return -1; }
    TSK_CUDA(cudaSetDevice(ctx->device));
    if (ctx->npending >= 2) { tsk_set_error("two tiles are already uploaded; call tsk_tile_process() first"); return -1; }
    if (ensure_cluster_buffers(ctx, nclusters) != 0) return -1;
    const int T = ctx->dp.total_cycles, L3 = ctx->dp.threep_length;
    auto &slot = ctx->pending[ctx->npending];
    memcpy(slot.inv, invmatrix, sizeof(float) * 16);
    slot.tv.nclusters = nclusters;
    slot.tv.firstclusterno = firstclusterno;
    const int u = ctx->up_slot;
    const size_t pitch = ((size_t)nclusters + 127) & ~(size_t)127;
    if (ensure(&ctx->d_bcl_own[u], &ctx->bcl_cap[u], pitch * T + 128) != 0) return -1;
    if (ensure(&ctx->d_cif_own[u], &ctx->cif_cap[u], pitch * 4 * L3 + 128) != 0) return -1;
    if (ctx->slot_busy[u]) {
        TSK_CUDA(cudaStreamWaitEvent(ctx->cstream, ctx->ev_done[u], 0));
        ctx->slot_busy[u] = 0;
    }
    slot.tv.bcl = ctx->d_bcl_own[u]; slot.tv.bcl_pitch = pitch;
    slot.tv.cif = ctx->d_cif_own[u]; slot.tv.cif_pitch = pitch;
    slot.slot = u;
    *slot_out = u; *bpitch_out = pitch; *cpitch_out = pitch;
    return 0;
}

int tsk_upload_slot_commit(tsk_ctx *ctx, int u)
{
    TSK_CUDA(cudaEventRecord(ctx->ev_up[u], ctx->cstream));
    ctx->up_slot ^= 1;
    ctx->npending++;
    return 0;
}

/* ---- importer path ----------------------------------------------------------------------- */
extern "C" int tsk_tile_upload(tsk_ctx *ctx, const uint8_t *bcl, size_t bcl_pitch,
                               const int16_t *cif, size_t cif_pitch, const float invmatrix[16],
                               uint32_t nclusters, uint32_t firstclusterno, int on_device)
{
    if (!ctx || !invmatrix || (nclusters && (!bcl || !cif))) { tsk_set_error("null argument"); return -1; }
    if (bcl_pitch < nclusters || cif_pitch < nclusters) { tsk_set_error("pitch smaller than nclusters"); return -1; }
    if (nclusters >= 0x7fffffffu) { tsk_set_error("too many clusters in one block"); return -1; }
    TSK_CUDA(cudaSetDevice(ctx->device));
    if (ctx->npending >= 2) { tsk_set_error("two tiles are already uploaded; call tsk_tile_process() first"); return -1; }
    if (ensure_cluster_buffers(ctx, nclusters) != 0) return -1;
    const int T = ctx->dp.total_cycles, L3 = ctx->dp.threep_length;
    auto &slot = ctx->pending[ctx->npending];
    memcpy(slot.inv, invmatrix, sizeof(float) * 16);
    slot.tv.nclusters = nclusters;
    slot.tv.firstclusterno = firstclusterno;
    if (on_device) {
        if (nclusters && (((uintptr_t)bcl % 16) != 0 || (bcl_pitch % 16) != 0)) {
            tsk_set_error("device base calls must be 16-byte aligned with a pitch that is a multiple of 16");
            return -1;
        }
        slot.tv.bcl = bcl; slot.tv.bcl_pitch = bcl_pitch;
        slot.tv.cif = cif; slot.tv.cif_pitch = cif_pitch;
        slot.slot = -1;
        ctx->npending++;
        return 0;
    }
    /* own HBM copy (rows padded to 128 clusters), on the copy stream: it overlaps the processing
     * of the previously uploaded tile */
    int u;
    size_t bpitch, cpitch;
    if (tsk_upload_slot_acquire(ctx, nclusters, firstclusterno, invmatrix, &u, &bpitch, &cpitch) != 0) return -1;
    if (nclusters) {
        TSK_CUDA(cudaMemcpy2DAsync(ctx->d_bcl_own[u], bpitch, bcl, bcl_pitch, nclusters, T,
                                   cudaMemcpyHostToDevice, ctx->cstream));
        TSK_CUDA(cudaMemcpy2DAsync(ctx->d_cif_own[u], cpitch * sizeof(int16_t), cif, cif_pitch * sizeof(int16_t),
                                   (size_t)nclusters * sizeof(int16_t), (size_t)4 * L3,
                                   cudaMemcpyHostToDevice, ctx->cstream));
    }
    return tsk_upload_slot_commit(ctx, u);
}

/* next uploaded tile becomes the resident one */
static int next_tile(tsk_ctx *ctx)
{
    if (ctx->npending == 0) {
        if (ctx->d_calls && ctx->tile_valid) return 0;      /* re-process the resident tile */
        tsk_set_error("no tile uploaded");
        return -1;
    }
    ctx->tile = ctx->pending[0].tv;
    memcpy(ctx->dp.inv, ctx->pending[0].inv, sizeof(float) * 16);
    if (ctx->pending[0].slot >= 0)
        TSK_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_up[ctx->pending[0].slot], 0));
    ctx->tile_slot = ctx->pending[0].slot;
    ctx->pending[0] = ctx->pending[1];
    ctx->npending--;
    ctx->tile_valid = 1;
    ctx->processed = 0;
    return 0;
}

int tsk_ruler_join(tsk_ctx *ctx)
{
    if (ctx->ruler_inflight) {
        TSK_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_rul[1], 0));
        ctx->ruler_inflight = 0;
    }
    return 0;
}

static int fetch_totals(tsk_ctx *ctx)
{
    const int S = ctx->dp.num_samples;
    TSK_CUDA(cudaMemcpyAsync(ctx->h_totals, ctx->d_totals, sizeof(uint32_t) * (S + 1),
                             cudaMemcpyDeviceToHost, ctx->stream));
    TSK_CUDA(cudaMemcpyAsync(ctx->h_recbase, ctx->d_recbase, sizeof(uint64_t) * (S + 1),
                             cudaMemcpyDeviceToHost, ctx->stream));
    TSK_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int tsk_tile_process_async(tsk_ctx *ctx)
{
    if (!ctx || !ctx->d_calls) { tsk_set_error("no tile uploaded"); return -1; }
    TSK_CUDA(cudaSetDevice(ctx->device));
    if (next_tile(ctx) != 0) return -1;
    if (tsk_launch_import(ctx, prof_take(ctx, 5, 0)) != 0) return -1;
    ctx->processed = 2;          /* enqueued; the host mirrors of the slot counts are not fetched yet */
    ctx->rlen_resident = 0;
    return 0;
}

extern "C" int tsk_tile_process(tsk_ctx *ctx)
{
    if (tsk_tile_process_async(ctx) != 0) return -1;
    if (fetch_totals(ctx) != 0) return -1;
    ctx->processed = 1;
    return 0;
}

extern "C" int tsk_set_stream(tsk_ctx *ctx, void *cuda_stream)
{
    if (!ctx) { tsk_set_error("null ctx"); return -1; }
    TSK_CUDA(cudaSetDevice(ctx->device));
    if (tsk_ruler_join(ctx) != 0) return -1;
    TSK_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->own_stream = 0;
    return 0;
}

extern "C" int tsk_debug_force_generic(tsk_ctx *ctx, int on)
{
    if (!ctx) { tsk_set_error("null ctx"); return -1; }
    ctx->force_generic = on;     /* 0: default, 1: literal kernel for every cluster, 2: row form of the fast kernel */
    return 0;
}

extern "C" int tsk_profile_enable(tsk_ctx *ctx, int on)
{
    if (!ctx) { tsk_set_error("null ctx"); return -1; }
    TSK_CUDA(cudaSetDevice(ctx->device));
    if (ctx->profile && prof_harvest_all(ctx) != 0) return -1;
    if (on && !ctx->profile)
        for (int i = 0; i < TSK_PROF_RING; i++)
            for (int k = 0; k < 5; k++) TSK_CUDA(cudaEventCreate(&ctx->prof[i].ev[k]));
    if (!on && ctx->profile) prof_destroy(ctx);
    ctx->profile = on ? 1 : 0;
    ctx->prof_head = ctx->prof_count = 0;
    for (int i = 0; i < 6; i++) { ctx->stage_ms[i] = 0.; ctx->stage_runs[i] = 0; }
    return 0;
}

extern "C" int tsk_profile_get(tsk_ctx *ctx, double stage_ms[6], uint64_t stage_runs[6])
{
    if (!ctx) { tsk_set_error("null ctx"); return -1; }
    TSK_CUDA(cudaSetDevice(ctx->device));
    if (ctx->profile && prof_harvest_all(ctx) != 0) return -1;      /* waits for the timed work */
    for (int i = 0; i < 6; i++) { stage_ms[i] = ctx->stage_ms[i]; stage_runs[i] = ctx->stage_runs[i]; }
    return 0;
}

extern "C" int tsk_tile_process_timed(tsk_ctx *ctx, int iters, float ms_out[5])
{
    if (!ctx || !ctx->d_calls || iters < 1) { tsk_set_error("bad argument"); return -1; }
    TSK_CUDA(cudaSetDevice(ctx->device));
    if (next_tile(ctx) != 0) return -1;
    cudaEvent_t ev[5];
    for (int i = 0; i < 5; i++) TSK_CUDA(cudaEventCreate(&ev[i]));
    const int S = ctx->dp.num_samples;
    double acc[5] = {0, 0, 0, 0, 0};
    uint32_t *saved = (uint32_t *)malloc(sizeof(uint32_t) * S * TSK_NCOUNT);
    for (int it = 0; it < iters; it++) {
        /* counters accumulate across calls by contract; keep them unchanged by the re-runs */
        TSK_CUDA(cudaMemcpyAsync(saved, ctx->d_counters, sizeof(uint32_t) * S * TSK_NCOUNT,
                                 cudaMemcpyDeviceToHost, ctx->stream));
        TSK_CUDA(cudaStreamSynchronize(ctx->stream));
        if (tsk_launch_import(ctx, ev) != 0) { free(saved); return -1; }
        TSK_CUDA(cudaStreamSynchronize(ctx->stream));
        float a, b, c, d, t;
        cudaEventElapsedTime(&a, ev[0], ev[1]);
        cudaEventElapsedTime(&b, ev[1], ev[2]);
        cudaEventElapsedTime(&c, ev[2], ev[3]);
        cudaEventElapsedTime(&d, ev[3], ev[4]);
        cudaEventElapsedTime(&t, ev[0], ev[4]);
        acc[0] += a; acc[1] += b; acc[2] += c; acc[3] += d; acc[4] += t;
        if (ctx->processed || it > 0) {
            TSK_CUDA(cudaMemcpyAsync(ctx->d_counters, saved, sizeof(uint32_t) * S * TSK_NCOUNT,
                                     cudaMemcpyHostToDevice, ctx->stream));
            TSK_CUDA(cudaStreamSynchronize(ctx->stream));
        }
    }
    free(saved);
    for (int i = 0; i < 5; i++) { ms_out[i] = (float)(acc[i] / iters); cudaEventDestroy(ev[i]); }
    if (fetch_totals(ctx) != 0) return -1;
    ctx->processed = 1;
    return 0;
}

static int need_processed(tsk_ctx *ctx)
{
    if (!ctx || !ctx->processed) { tsk_set_error("tsk_tile_process() has not run on this tile"); return -1; }
    if (cudaSetDevice(ctx->device) != cudaSuccess) return -1;
    if (ctx->processed == 2) {   /* enqueued by tsk_tile_process_async(): wait and fetch the slot counts */
        if (fetch_totals(ctx) != 0) return -1;
        ctx->processed = 1;
    }
    return 0;
}

extern "C" int tsk_tile_get_calls(tsk_ctx *ctx, tsk_call *calls)
{
    if (need_processed(ctx) != 0) return -1;
    TSK_CUDA(cudaMemcpyAsync(calls, ctx->d_calls, sizeof(tsk_call) * (size_t)ctx->tile.nclusters,
                             cudaMemcpyDeviceToHost, ctx->stream));
    TSK_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int tsk_tile_get_record_count(tsk_ctx *ctx, int sample, uint32_t *nslots)
{
    if (need_processed(ctx) != 0) return -1;
    if (sample < 0 || sample >= ctx->dp.num_samples) { tsk_set_error("bad sample"); return -1; }
    *nslots = ctx->h_totals[sample];
    return 0;
}

extern "C" int tsk_tile_get_records(tsk_ctx *ctx, int sample, uint32_t *records, size_t capacity_words)
{
    if (need_processed(ctx) != 0) return -1;
    if (sample < 0 || sample >= ctx->dp.num_samples) { tsk_set_error("bad sample"); return -1; }
    const size_t words = (size_t)ctx->h_totals[sample] * (size_t)(2 + ctx->hp.samples[sample].signal_dump_length);
    if (capacity_words < words) { tsk_set_error("record buffer too small"); return -1; }
    if (words) {
        TSK_CUDA(cudaMemcpyAsync(records, ctx->d_records + ctx->h_recbase[sample], sizeof(uint32_t) * words,
                                 cudaMemcpyDeviceToHost, ctx->stream));
        TSK_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return 0;
}

extern "C" int tsk_get_counters(tsk_ctx *ctx, tsk_counters *counters)
{
    if (!ctx) { tsk_set_error("null ctx"); return -1; }
    TSK_CUDA(cudaSetDevice(ctx->device));
    TSK_CUDA(cudaMemcpyAsync(counters, ctx->d_counters, sizeof(uint32_t) * ctx->dp.num_samples * TSK_NCOUNT,
                             cudaMemcpyDeviceToHost, ctx->stream));
    TSK_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int tsk_counters_reset(tsk_ctx *ctx)
{
    if (!ctx) { tsk_set_error("null ctx"); return -1; }
    TSK_CUDA(cudaSetDevice(ctx->device));
    TSK_CUDA(cudaMemsetAsync(ctx->d_counters, 0, sizeof(uint32_t) * ctx->dp.num_samples * TSK_NCOUNT, ctx->stream));
    return 0;
}

extern "C" int tsk_get_hist(tsk_ctx *ctx, uint32_t *pos, uint32_t *neg)
{
    if (!ctx) { tsk_set_error("null ctx"); return -1; }
    TSK_CUDA(cudaSetDevice(ctx->device));
    const size_t histbytes = sizeof(uint32_t) * (size_t)ctx->dp.total_cycles * ctx->dp.bins;
    if (pos) TSK_CUDA(cudaMemcpyAsync(pos, ctx->d_pos, histbytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (neg) TSK_CUDA(cudaMemcpyAsync(neg, ctx->d_neg, histbytes, cudaMemcpyDeviceToHost, ctx->stream));
    TSK_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int tsk_get_device_ptrs(tsk_ctx *ctx, void **pos_hist, void **neg_hist, void **counters,
                                   void **ruler_pos_hist)
{
    if (!ctx) { tsk_set_error("null ctx"); return -1; }
    if (pos_hist) *pos_hist = ctx->d_pos;
    if (neg_hist) *neg_hist = ctx->d_neg;
    if (counters) *counters = ctx->d_counters;
    if (ruler_pos_hist) { if (tsk_ruler_join(ctx) != 0) return -1; *ruler_pos_hist = ctx->d_ruler_pos; }
    return 0;
}

extern "C" int tsk_synchronize(tsk_ctx *ctx)
{
    if (!ctx) { tsk_set_error("null ctx"); return -1; }
    TSK_CUDA(cudaSetDevice(ctx->device));
    if (tsk_ruler_join(ctx) != 0) return -1;
    TSK_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" uint64_t tsk_launch_count(tsk_ctx *ctx) { return ctx ? ctx->launches : 0; }

/* ---- ruler path ---------------------------------------------------------------------------- */
extern "C" uint32_t tsk_cutoff_to_packed(double cutoff)
{
    /* polyaruler.c:93-96; "nan" -> (int)NaN = INT_MIN on x86-64 -> 0x80000001 -> clamped */
    const double scaled = cutoff * (double)(TSK_SCORE_MAX - 1);
    uint32_t v;
    if (!(scaled > -2147483649.0 && scaled < 2147483648.0)) v = 0x80000001u;
    else v = 1u + (uint32_t)(int)scaled;
    if (v >= TSK_SCORE_MAX) v = TSK_SCORE_MAX;
    return v;
}

extern "C" int tsk_ruler_hist_reset(tsk_ctx *ctx, int ncutoffs, int dist_sampling_bins)
{
    if (!ctx || ncutoffs < 1 || ncutoffs > TSK_MAX_CUTOFF_CYCLES || dist_sampling_bins < 1) {
        tsk_set_error("bad argument"); return -1;
    }
    TSK_CUDA(cudaSetDevice(ctx->device));
    if (tsk_ruler_join(ctx) != 0) return -1;
    const size_t words = (size_t)ncutoffs * dist_sampling_bins;
    if (ensure(&ctx->d_ruler_pos, &ctx->ruler_pos_words, words) != 0) return -1;
    TSK_CUDA(cudaMemsetAsync(ctx->d_ruler_pos, 0, sizeof(uint32_t) * words, ctx->stream));
    TSK_CUDA(cudaMemsetAsync(ctx->d_ruler_called, 0, sizeof(unsigned long long), ctx->stream));
    ctx->ruler_ncut = ncutoffs;
    ctx->ruler_bins = dist_sampling_bins;
    return 0;
}

extern "C" int tsk_ruler_get_hist(tsk_ctx *ctx, uint32_t *pos)
{
    if (!ctx || !ctx->d_ruler_pos) { tsk_set_error("no ruler histogram"); return -1; }
    TSK_CUDA(cudaSetDevice(ctx->device));
    if (tsk_ruler_join(ctx) != 0) return -1;
    TSK_CUDA(cudaMemcpyAsync(pos, ctx->d_ruler_pos, sizeof(uint32_t) * (size_t)ctx->ruler_ncut * ctx->ruler_bins,
                             cudaMemcpyDeviceToHost, ctx->stream));
    TSK_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

/* histogram shape and cutoffs on the device; the cutoffs are re-sent only when they change */
static int ruler_prepare(tsk_ctx *ctx, const uint32_t *cutoffs, int ncutoffs, int bins)
{
    if (ncutoffs < 1 || ncutoffs > TSK_MAX_CUTOFF_CYCLES || bins < 1 || !cutoffs) {
        tsk_set_error("bad ruler argument"); return -1;
    }
    if (tsk_ruler_join(ctx) != 0) return -1;      /* cutoffs / histogram shape may change below */
    if (!ctx->d_ruler_pos || ctx->ruler_ncut != ncutoffs || ctx->ruler_bins != bins)
        if (tsk_ruler_hist_reset(ctx, ncutoffs, bins) != 0) return -1;
    if (ctx->h_ncut != ncutoffs || memcmp(ctx->h_cutoffs, cutoffs, sizeof(uint32_t) * ncutoffs) != 0) {
        /* h_cutoffs outlives the call, so the copy may complete later */
        memcpy(ctx->h_cutoffs, cutoffs, sizeof(uint32_t) * ncutoffs);
        ctx->h_ncut = ncutoffs;
        TSK_CUDA(cudaMemcpyAsync(ctx->d_cutoffs, ctx->h_cutoffs, sizeof(uint32_t) * ncutoffs, cudaMemcpyHostToDevice,
                                 ctx->stream));
    }
    return 0;
}

static int ruler_common(tsk_ctx *ctx, const uint32_t *d_records, size_t nrecords, int dump_len,
                        const uint32_t *cutoffs, int ncutoffs, int min_len, float weight, int bins,
                        float gap, int16_t *polya_len_out)
{
    if (dump_len < 0) { tsk_set_error("bad ruler argument"); return -1; }
    if (ruler_prepare(ctx, cutoffs, ncutoffs, bins) != 0) return -1;
    if (ensure(&ctx->d_rlen, &ctx->rlen_cap, nrecords ? nrecords : 1) != 0) return -1;
    ctx->rlen_resident = 0;
    if (nrecords) {
        cudaEvent_t *ev = prof_take(ctx, 2, 4);
        if (ev) TSK_CUDA(cudaEventRecord(ev[0], ctx->stream));
        if (tsk_launch_ruler(ctx, d_records, nrecords, dump_len, ncutoffs, min_len, weight, bins, gap,
                             ctx->d_rlen) != 0) return -1;
        if (ev) TSK_CUDA(cudaEventRecord(ev[1], ctx->stream));
        TSK_CUDA(cudaMemcpyAsync(polya_len_out, ctx->d_rlen, sizeof(int16_t) * nrecords,
                                 cudaMemcpyDeviceToHost, ctx->stream));
    }
    TSK_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int tsk_ruler_records(tsk_ctx *ctx, const uint32_t *records, size_t nrecords, int dump_len,
                                 const uint32_t *cutoffs, int ncutoffs, int minimum_polya_len,
                                 float downhill_ext_weight, int dist_sampling_bins, float dist_sampling_gap,
                                 int16_t *polya_len_out, int on_device)
{
    if (!ctx || (nrecords && (!records || !polya_len_out))) { tsk_set_error("null argument"); return -1; }
    TSK_CUDA(cudaSetDevice(ctx->device));
    const uint32_t *d_rec = records;
    if (!on_device && nrecords) {
        const size_t words = nrecords * (size_t)(2 + dump_len);
        if (ensure(&ctx->d_rrec, &ctx->rrec_cap, words) != 0) return -1;
        TSK_CUDA(cudaMemcpyAsync(ctx->d_rrec, records, sizeof(uint32_t) * words, cudaMemcpyHostToDevice, ctx->stream));
        d_rec = ctx->d_rrec;
    }
    return ruler_common(ctx, d_rec, nrecords, dump_len, cutoffs, ncutoffs, minimum_polya_len,
                        downhill_ext_weight, dist_sampling_bins, dist_sampling_gap, polya_len_out);
}

extern "C" int tsk_ruler_tile(tsk_ctx *ctx, int sample, const uint32_t *cutoffs, int ncutoffs,
                              int minimum_polya_len, float downhill_ext_weight, int dist_sampling_bins,
                              float dist_sampling_gap, int16_t *polya_len_out)
{
    if (need_processed(ctx) != 0) return -1;
    if (sample < 0 || sample >= ctx->dp.num_samples) { tsk_set_error("bad sample"); return -1; }
    return ruler_common(ctx, ctx->d_records + ctx->h_recbase[sample], ctx->h_totals[sample],
                        ctx->hp.samples[sample].signal_dump_length, cutoffs, ncutoffs, minimum_polya_len,
                        downhill_ext_weight, dist_sampling_bins, dist_sampling_gap, polya_len_out);
}

extern "C" int tsk_ruler_tile_all_async(tsk_ctx *ctx, const uint32_t *cutoffs, int ncutoffs, int minimum_polya_len,
                                        float downhill_ext_weight, int dist_sampling_bins, float dist_sampling_gap)
{
    if (!ctx || !ctx->processed) { tsk_set_error("tsk_tile_process() has not run on this tile"); return -1; }
    TSK_CUDA(cudaSetDevice(ctx->device));
    if (ruler_prepare(ctx, cutoffs, ncutoffs, dist_sampling_bins) != 0) return -1;
    /* a tile has at most one record slot per cluster */
    if (ensure(&ctx->d_rlen, &ctx->rlen_cap, (size_t)ctx->cap_clusters) != 0) return -1;
    /* the ruler runs on its own stream from here on: the next tile's decode and scan kernels
     * (main stream) do not wait for it; tsk_launch_import joins before the record slots are reused */
    TSK_CUDA(cudaMemcpyAsync(ctx->d_rtotals, ctx->d_totals, sizeof(uint32_t) * (TSK_MAX_SAMPLES + 1),
                             cudaMemcpyDeviceToDevice, ctx->stream));
    TSK_CUDA(cudaMemcpyAsync(ctx->d_rrecbase, ctx->d_recbase, sizeof(uint64_t) * (TSK_MAX_SAMPLES + 1),
                             cudaMemcpyDeviceToDevice, ctx->stream));
    TSK_CUDA(cudaEventRecord(ctx->ev_rul[0], ctx->stream));
    TSK_CUDA(cudaStreamWaitEvent(ctx->rstream, ctx->ev_rul[0], 0));
    cudaEvent_t *ev = prof_take(ctx, 2, 4);
    if (ev) TSK_CUDA(cudaEventRecord(ev[0], ctx->rstream));
    if (tsk_launch_ruler_resident(ctx, ctx->rstream, ncutoffs, minimum_polya_len, downhill_ext_weight, dist_sampling_bins,
                                  dist_sampling_gap, ctx->d_rlen) != 0) return -1;
    if (ev) TSK_CUDA(cudaEventRecord(ev[1], ctx->rstream));
    TSK_CUDA(cudaEventRecord(ctx->ev_rul[1], ctx->rstream));
    ctx->ruler_inflight = 1;
    ctx->rlen_resident = 1;
    return 0;
}

extern "C" int tsk_ruler_get_lengths(tsk_ctx *ctx, int16_t *polya_len_out, uint32_t *slot_offsets)
{
    if (need_processed(ctx) != 0) return -1;
    if (!ctx->rlen_resident) { tsk_set_error("tsk_ruler_tile_all_async() has not run on this tile"); return -1; }
    if (tsk_ruler_join(ctx) != 0) return -1;
    if (!slot_offsets) { tsk_set_error("null argument"); return -1; }
    const int S = ctx->dp.num_samples;
    uint32_t at = 0;
    for (int s = 0; s < S; s++) { slot_offsets[s] = at; at += ctx->h_totals[s]; }
    slot_offsets[S] = at;
    if (at) {
        if (!polya_len_out) { tsk_set_error("null output"); return -1; }
        TSK_CUDA(cudaMemcpyAsync(polya_len_out, ctx->d_rlen, sizeof(int16_t) * at, cudaMemcpyDeviceToHost, ctx->stream));
    }
    TSK_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int tsk_ruler_tile_all(tsk_ctx *ctx, const uint32_t *cutoffs, int ncutoffs, int minimum_polya_len,
                                  float downhill_ext_weight, int dist_sampling_bins, float dist_sampling_gap,
                                  int16_t *polya_len_out, uint32_t *slot_offsets)
{
    if (tsk_ruler_tile_all_async(ctx, cutoffs, ncutoffs, minimum_polya_len, downhill_ext_weight, dist_sampling_bins,
                                 dist_sampling_gap) != 0) return -1;
    return tsk_ruler_get_lengths(ctx, polya_len_out, slot_offsets);
}

extern "C" int tsk_ruler_get_called(tsk_ctx *ctx, uint64_t *ncalled)
{
    if (!ctx || !ncalled) { tsk_set_error("null argument"); return -1; }
    TSK_CUDA(cudaSetDevice(ctx->device));
    if (tsk_ruler_join(ctx) != 0) return -1;
    unsigned long long v = 0;
    TSK_CUDA(cudaMemcpyAsync(&v, ctx->d_ruler_called, sizeof(v), cudaMemcpyDeviceToHost, ctx->stream));
    TSK_CUDA(cudaStreamSynchronize(ctx->stream));
    *ncalled = v;
    return 0;
}

/* ---- cutoff estimation ----------------------------------------------------------------------- */
extern "C" int tsk_fit_cutoffs(tsk_ctx *ctx, const uint64_t *pos, const uint64_t *neg, int cycles, int bins,
                               uint64_t min_spots, double kde_bandwidth, double search_low,
                               const double *search_high, int nsearch_high, double *cutoffs_out)
{
    if (!ctx || !pos || !neg || !cutoffs_out || cycles < 1 || bins < 2 || nsearch_high < 1 ||
        nsearch_high > 8 || !search_high) { tsk_set_error("bad argument"); return -1; }
    TSK_CUDA(cudaSetDevice(ctx->device));
    const size_t hw = (size_t)cycles * bins;
    const size_t hb = sizeof(uint64_t) * hw;
    if (ensure(&ctx->d_fit_hist, &ctx->fit_cap, 2 * hw) != 0) return -1;
    if (ensure(&ctx->d_fit_out, &ctx->fit_out_cap, (size_t)cycles) != 0) return -1;
    uint64_t *d_pos = ctx->d_fit_hist, *d_neg = ctx->d_fit_hist + hw;
    double *d_out = ctx->d_fit_out;
    TSK_CUDA(cudaMemcpyAsync(d_pos, pos, hb, cudaMemcpyHostToDevice, ctx->stream));
    TSK_CUDA(cudaMemcpyAsync(d_neg, neg, hb, cudaMemcpyHostToDevice, ctx->stream));
    cudaEvent_t *pev = prof_take(ctx, 2, 5);
    if (pev) cudaEventRecord(pev[0], ctx->stream);
    int rc = tsk_launch_fit(ctx, d_pos, d_neg, cycles, bins, min_spots, kde_bandwidth, search_low,
                            search_high, nsearch_high, d_out);
    if (pev) cudaEventRecord(pev[1], ctx->stream);
    if (rc == 0) {
        cudaError_t e = cudaMemcpyAsync(cutoffs_out, d_out, sizeof(double) * cycles, cudaMemcpyDeviceToHost,
                                        ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { tsk_set_error("fit copy back: %s", cudaGetErrorString(e)); rc = -1; }
    }
    return rc;
}
