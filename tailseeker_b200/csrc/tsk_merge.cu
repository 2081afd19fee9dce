/*
 * tsk_merge.cu -- the cross-tile sums of the pipeline, on the devices.
 *
 * The reference merges per-tile FILES in Python: pos/neg score histograms before cutoff estimation
 * (scripts/calculate-optimal-parameters.py:139-140), the ruler's positive histogram per round
 * (scripts/measure-polya-lengths.py:64-70) and the demultiplexing counters
 * (scripts/stats-merge-demultiplexing-counts.py:33-53).  Here a tile set lives on each GPU:
 *   tsk_merge_*          run-wide u64 accumulators on one device (sum over the tiles it processed)
 *                        and the per-tile histograms kept for the per-tile fits;
 *   tsk_merge_allreduce  ONE ncclAllReduce(uint64, sum) over the packed accumulators;
 *   tsk_hist_allreduce   the same on a context's own u32 histograms / counters, in place, as one
 *                        grouped collective (§8b);
 *   tsk_merge_fit        cutoffs of the run-wide histograms and of every kept tile in one launch,
 *                        reading device memory (no round trip through the host).
 * Integer sums: order-independent, bit-exact, equal to the single-device result.
 *
 * NCCL is loaded at run time (dlopen "libnccl.so.2", or the file named by TSK_NCCL_LIB): the
 * library and the drop-in programs do not depend on it until a collective is asked for.
 */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <nccl.h>
#include "tsk_internal.h"

/* ---- NCCL through dlopen ----------------------------------------------------------------------- */
struct NcclApi {
    void *handle;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)(void);
    ncclResult_t (*GroupEnd)(void);
    const char *(*GetErrorString)(ncclResult_t);
};

static NcclApi *nccl_api(void)
{
    static NcclApi api;
    static int state = 0;            /* 0 not tried, 1 ok, -1 failed */
    if (state == 1) return &api;
    if (state == -1) { tsk_set_error("NCCL is not available (libnccl.so.2 / TSK_NCCL_LIB)"); return NULL; }
    const char *name = getenv("TSK_NCCL_LIB");
    void *h = dlopen(name && *name ? name : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { state = -1; tsk_set_error("cannot load NCCL: %s", dlerror()); return NULL; }
    api.handle = h;
#define TSK_NCCL_SYM(field, sym) *(void **)(&api.field) = dlsym(h, sym); if (!api.field) { state = -1; tsk_set_error("NCCL lacks %s", sym); return NULL; }
    TSK_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    TSK_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    TSK_NCCL_SYM(CommInitAll, "ncclCommInitAll")
    TSK_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    TSK_NCCL_SYM(AllReduce, "ncclAllReduce")
    TSK_NCCL_SYM(GroupStart, "ncclGroupStart")
    TSK_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    TSK_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef TSK_NCCL_SYM
    state = 1;
    return &api;
}

#define TSK_NCCL(api, call)                                                                  \
    do {                                                                                     \
        ncclResult_t r_ = (call);                                                            \
        if (r_ != ncclSuccess) { tsk_set_error("%s failed: %s", #call, (api)->GetErrorString(r_)); return -1; } \
    } while (0)

extern "C" int tsk_nccl_unique_id(uint8_t id[TSK_NCCL_ID_BYTES])
{
    NcclApi *a = nccl_api();
    if (!a || !id) return -1;
    static_assert(sizeof(ncclUniqueId) == TSK_NCCL_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId u;
    TSK_NCCL(a, a->GetUniqueId(&u));
    memcpy(id, &u, sizeof(u));
    return 0;
}

extern "C" int tsk_nccl_init_rank(void **comm, int device, int nranks, int rank, const uint8_t id[TSK_NCCL_ID_BYTES])
{
    NcclApi *a = nccl_api();
    if (!a || !comm || !id) return -1;
    TSK_CUDA(cudaSetDevice(device));
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    ncclComm_t c;
    TSK_NCCL(a, a->CommInitRank(&c, nranks, u, rank));
    *comm = (void *)c;
    return 0;
}

extern "C" int tsk_nccl_init_all(void **comms, int ndev, const int *devices)
{
    NcclApi *a = nccl_api();
    if (!a || !comms || ndev < 1 || ndev > 64) return -1;
    ncclComm_t c[64];
    TSK_NCCL(a, a->CommInitAll(c, ndev, devices));
    for (int i = 0; i < ndev; i++) comms[i] = (void *)c[i];
    return 0;
}

extern "C" int tsk_nccl_destroy(void *comm)
{
    NcclApi *a = nccl_api();
    if (!a) return -1;
    if (comm) TSK_NCCL(a, a->CommDestroy((ncclComm_t)comm));
    return 0;
}

extern "C" int tsk_nccl_group_start(void) { NcclApi *a = nccl_api(); if (!a) return -1; TSK_NCCL(a, a->GroupStart()); return 0; }
extern "C" int tsk_nccl_group_end(void) { NcclApi *a = nccl_api(); if (!a) return -1; TSK_NCCL(a, a->GroupEnd()); return 0; }

/* §8b: the context's pos / neg histograms, demultiplexing counters and (if one exists) the ruler
 * histogram summed over all ranks, in place, as one grouped collective on the context's stream. */
extern "C" int tsk_hist_allreduce(tsk_ctx *ctx, void *nccl_comm)
{
    NcclApi *a = nccl_api();
    if (!a) return -1;
    if (!ctx || !nccl_comm) { tsk_set_error("null argument"); return -1; }
    TSK_CUDA(cudaSetDevice(ctx->device));
    if (tsk_ruler_join(ctx) != 0) return -1;
    const size_t hw = (size_t)ctx->dp.total_cycles * ctx->dp.bins;
    ncclComm_t c = (ncclComm_t)nccl_comm;
    TSK_NCCL(a, a->GroupStart());
    TSK_NCCL(a, a->AllReduce(ctx->d_pos, ctx->d_pos, hw, ncclUint32, ncclSum, c, ctx->stream));
    TSK_NCCL(a, a->AllReduce(ctx->d_neg, ctx->d_neg, hw, ncclUint32, ncclSum, c, ctx->stream));
    TSK_NCCL(a, a->AllReduce(ctx->d_counters, ctx->d_counters, (size_t)ctx->dp.num_samples * TSK_NCOUNT, ncclUint32, ncclSum, c, ctx->stream));
    if (ctx->d_ruler_pos)
        TSK_NCCL(a, a->AllReduce(ctx->d_ruler_pos, ctx->d_ruler_pos, (size_t)ctx->ruler_ncut * ctx->ruler_bins, ncclUint32, ncclSum, c, ctx->stream));
    TSK_NCCL(a, a->GroupEnd());
    ctx->launches++;
    return 0;
}

/* ---- run-wide accumulators ------------------------------------------------------------------------ */
struct tsk_merge {
    int device, cycles, bins, nsamples, max_tiles, ntiles;
    size_t hw;                       /* cycles * bins */
    /* one packed buffer: pos | neg | ruler pos | counters [nsamples][6] | called records | clusters */
    unsigned long long *d_acc;
    size_t acc_words;
    /* per-tile histograms of the tiles this device processed, for the per-tile fits */
    unsigned long long *d_tile_pos, *d_tile_neg;          /* [max_tiles][cycles][bins] */
    double *d_fit;  size_t fit_cap;
    cudaStream_t own;                /* used when a call is made without a context (end of a run) */
};

__global__ void k_merge_add_u32(unsigned long long *__restrict__ acc, unsigned long long *__restrict__ keep,
                                const uint32_t *__restrict__ src, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long v = src[i];
        acc[i] += v;
        if (keep) keep[i] = v;
    }
}

__global__ void k_merge_set_u64(unsigned long long *__restrict__ dst, const unsigned long long *__restrict__ src, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

extern "C" int tsk_merge_create(tsk_merge **out, int device, int cycles, int bins, int nsamples, int max_tiles)
{
    if (!out || cycles < 1 || bins < 1 || nsamples < 1 || nsamples > TSK_MAX_SAMPLES || max_tiles < 0) { tsk_set_error("bad argument"); return -1; }
    *out = NULL;
    if (tsk_probe(device) < 0) return -1;
    TSK_CUDA(cudaSetDevice(device));
    tsk_merge *m = (tsk_merge *)calloc(1, sizeof(tsk_merge));
    if (!m) { tsk_set_error("out of memory"); return -1; }
    m->device = device; m->cycles = cycles; m->bins = bins; m->nsamples = nsamples; m->max_tiles = max_tiles;
    m->hw = (size_t)cycles * bins;
    m->acc_words = 3 * m->hw + (size_t)nsamples * TSK_NCOUNT + 2;
    TSK_CUDA(cudaStreamCreateWithFlags(&m->own, cudaStreamNonBlocking));
    TSK_CUDA(cudaMalloc(&m->d_acc, sizeof(unsigned long long) * m->acc_words));
    TSK_CUDA(cudaMemset(m->d_acc, 0, sizeof(unsigned long long) * m->acc_words));
    if (max_tiles > 0) {
        TSK_CUDA(cudaMalloc(&m->d_tile_pos, sizeof(unsigned long long) * m->hw * max_tiles));
        TSK_CUDA(cudaMalloc(&m->d_tile_neg, sizeof(unsigned long long) * m->hw * max_tiles));
    }
    *out = m;
    return 0;
}

extern "C" void tsk_merge_destroy(tsk_merge *m)
{
    if (!m) return;
    cudaSetDevice(m->device);
    if (m->d_acc) cudaFree(m->d_acc);
    if (m->d_tile_pos) cudaFree(m->d_tile_pos);
    if (m->d_tile_neg) cudaFree(m->d_tile_neg);
    if (m->d_fit) cudaFree(m->d_fit);
    if (m->own) cudaStreamDestroy(m->own);
    free(m);
}

/* stream of a call that may come without a context (ctx == NULL: everything the device still has
 * queued is waited for first, then the accumulator's own stream is used) */
static int merge_stream(tsk_merge *m, tsk_ctx *ctx, cudaStream_t *st)
{
    if (ctx) { *st = ctx->stream; return 0; }
    TSK_CUDA(cudaSetDevice(m->device));
    TSK_CUDA(cudaDeviceSynchronize());
    *st = m->own;
    return 0;
}

static int merge_matches(const tsk_merge *m, const tsk_ctx *ctx, bool ctx_optional = false)
{
    if (!m || (!ctx && !ctx_optional)) { tsk_set_error("null argument"); return -1; }
    if (!ctx) return 0;
    if (m->device != ctx->device || m->cycles != ctx->dp.total_cycles || m->bins != ctx->dp.bins ||
        m->nsamples != ctx->dp.num_samples) { tsk_set_error("accumulator and context differ in device or shape"); return -1; }
    return 0;
}

/* Adds the histograms of the tile last processed by `ctx` (and its size) to the run-wide sums; keeps
 * a copy for the per-tile fit when the accumulator was created with room for tiles.  Enqueued on the
 * context's stream, so it may follow tsk_tile_process_async() directly. */
extern "C" int tsk_merge_add_tile(tsk_merge *m, tsk_ctx *ctx)
{
    if (merge_matches(m, ctx) != 0) return -1;
    TSK_CUDA(cudaSetDevice(m->device));
    unsigned long long *kp = NULL, *kn = NULL;
    if (m->max_tiles > 0) {
        if (m->ntiles >= m->max_tiles) { tsk_set_error("more tiles than the accumulator was created for"); return -1; }
        kp = m->d_tile_pos + m->hw * m->ntiles;
        kn = m->d_tile_neg + m->hw * m->ntiles;
    }
    k_merge_add_u32<<<148 * 2, 256, 0, ctx->stream>>>(m->d_acc, kp, ctx->d_pos, m->hw);
    k_merge_add_u32<<<148 * 2, 256, 0, ctx->stream>>>(m->d_acc + m->hw, kn, ctx->d_neg, m->hw);
    ctx->launches += 2;
    m->ntiles++;
    TSK_CUDA(cudaGetLastError());
    return 0;
}

/* Adds the context's ruler histogram and demultiplexing counters (both accumulate inside the
 * context across tiles, so: once, after the last tile) and the called-records counter. */
extern "C" int tsk_merge_add_totals(tsk_merge *m, tsk_ctx *ctx)
{
    if (merge_matches(m, ctx) != 0) return -1;
    TSK_CUDA(cudaSetDevice(m->device));
    if (tsk_ruler_join(ctx) != 0) return -1;
    if (ctx->d_ruler_pos) {
        if (ctx->ruler_ncut != m->cycles || ctx->ruler_bins != m->bins) { tsk_set_error("ruler histogram has another shape"); return -1; }
        k_merge_add_u32<<<148 * 2, 256, 0, ctx->stream>>>(m->d_acc + 2 * m->hw, NULL, ctx->d_ruler_pos, m->hw);
        ctx->launches++;
    }
    k_merge_add_u32<<<1, 256, 0, ctx->stream>>>(m->d_acc + 3 * m->hw, NULL, ctx->d_counters, (size_t)m->nsamples * TSK_NCOUNT);
    ctx->launches++;
    TSK_CUDA(cudaGetLastError());
    return 0;
}

/* One collective: the packed accumulators summed over the communicator's ranks, in place. */
extern "C" int tsk_merge_allreduce(tsk_merge *m, tsk_ctx *ctx, void *nccl_comm)
{
    NcclApi *a = nccl_api();
    if (!a) return -1;
    TskRange nvtx("tsk:merge_allreduce");
    if (merge_matches(m, ctx, true) != 0 || !nccl_comm) { if (!nccl_comm) tsk_set_error("null communicator"); return -1; }
    TSK_CUDA(cudaSetDevice(m->device));
    cudaStream_t st;
    if (merge_stream(m, ctx, &st) != 0) return -1;
    TSK_NCCL(a, a->AllReduce(m->d_acc, m->d_acc, m->acc_words, ncclUint64, ncclSum, (ncclComm_t)nccl_comm, st));
    if (ctx) ctx->launches++;
    return 0;
}

extern "C" int tsk_merge_get(tsk_merge *m, tsk_ctx *ctx, uint64_t *pos, uint64_t *neg, uint64_t *ruler_pos, uint64_t *counters)
{
    if (merge_matches(m, ctx, true) != 0) return -1;
    TSK_CUDA(cudaSetDevice(m->device));
    cudaStream_t st;
    if (merge_stream(m, ctx, &st) != 0) return -1;
    const size_t hb = sizeof(uint64_t) * m->hw;
    if (pos) TSK_CUDA(cudaMemcpyAsync(pos, m->d_acc, hb, cudaMemcpyDeviceToHost, st));
    if (neg) TSK_CUDA(cudaMemcpyAsync(neg, m->d_acc + m->hw, hb, cudaMemcpyDeviceToHost, st));
    if (ruler_pos) TSK_CUDA(cudaMemcpyAsync(ruler_pos, m->d_acc + 2 * m->hw, hb, cudaMemcpyDeviceToHost, st));
    if (counters) TSK_CUDA(cudaMemcpyAsync(counters, m->d_acc + 3 * m->hw, sizeof(uint64_t) * m->nsamples * TSK_NCOUNT,
                                           cudaMemcpyDeviceToHost, st));
    TSK_CUDA(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int tsk_merge_reset(tsk_merge *m, tsk_ctx *ctx, int what)
{
    if (merge_matches(m, ctx) != 0) return -1;
    TSK_CUDA(cudaSetDevice(m->device));
    if (what & 1) { TSK_CUDA(cudaMemsetAsync(m->d_acc, 0, sizeof(unsigned long long) * 2 * m->hw, ctx->stream)); m->ntiles = 0; }
    if (what & 2) TSK_CUDA(cudaMemsetAsync(m->d_acc + 2 * m->hw, 0, sizeof(unsigned long long) * m->hw, ctx->stream));
    if (what & 4) TSK_CUDA(cudaMemsetAsync(m->d_acc + 3 * m->hw, 0, sizeof(unsigned long long) * ((size_t)m->nsamples * TSK_NCOUNT + 2), ctx->stream));
    return 0;
}

/* find_all_eqodds_point() (calculate-optimal-parameters.py:57-78) for the run-wide histograms
 * (row block 0) and for every kept tile (row blocks 1 .. ntiles): two launches on device memory,
 * one CTA per (tile, cycle) -- not one launch and two copies per tile.
 * positive: 0 = the importer's positive histogram, 1 = the ruler's (round >= 1, main.py:214-215;
 * per-tile rows then use the tiles' importer positives: the ruler's are only summed).
 * cutoffs_out: double [(1 + ntiles) * cycles], NaN where a cycle has too few spots or no root. */
extern "C" int tsk_merge_fit(tsk_merge *m, tsk_ctx *ctx, int positive, uint64_t min_spots, double kde_bandwidth,
                             double search_low, const double *search_high, int nsearch_high, double *cutoffs_out)
{
    if (merge_matches(m, ctx) != 0) return -1;
    if (!cutoffs_out || nsearch_high < 1 || nsearch_high > 8 || !search_high || positive < 0 || positive > 1) { tsk_set_error("bad argument"); return -1; }
    TSK_CUDA(cudaSetDevice(m->device));
    const int rows = (1 + (m->max_tiles > 0 ? m->ntiles : 0)) * m->cycles;
    if ((size_t)rows > m->fit_cap) {
        if (m->d_fit) cudaFree(m->d_fit);
        m->d_fit = NULL; m->fit_cap = 0;
        TSK_CUDA(cudaMalloc(&m->d_fit, sizeof(double) * (size_t)rows));
        m->fit_cap = (size_t)rows;
    }
    const unsigned long long *rpos = m->d_acc + (positive ? 2 * m->hw : 0), *rneg = m->d_acc + m->hw;
    if (tsk_launch_fit(ctx, (const uint64_t *)rpos, (const uint64_t *)rneg, m->cycles, m->bins, min_spots, kde_bandwidth,
                       search_low, search_high, nsearch_high, m->d_fit) != 0) return -1;
    if (rows > m->cycles &&
        tsk_launch_fit(ctx, (const uint64_t *)m->d_tile_pos, (const uint64_t *)m->d_tile_neg, rows - m->cycles, m->bins,
                       min_spots, kde_bandwidth, search_low, search_high, nsearch_high, m->d_fit + m->cycles) != 0) return -1;
    TSK_CUDA(cudaMemcpyAsync(cutoffs_out, m->d_fit, sizeof(double) * (size_t)rows, cudaMemcpyDeviceToHost, ctx->stream));
    TSK_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int tsk_merge_tile_count(tsk_merge *m) { return m ? m->ntiles : -1; }
