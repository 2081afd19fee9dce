/* Internal declarations shared by the CUDA translation units of libtailseeker_b200.so. */
#ifndef TSK_INTERNAL_H
#define TSK_INTERNAL_H

#include <cuda_runtime.h>
#include <stdint.h>
#include <nvtx3/nvToolsExt.h>          /* header-only: ranges show up in Nsight Systems / ncu --nvtx */
#include "../../include/tailseeker_b200.h"

/* one NVTX range per stage of the path (host-side bracket around the stage's launches) */
struct TskRange {
    explicit TskRange(const char *name) { nvtxRangePushA(name); }
    ~TskRange() { nvtxRangePop(); }
};

#define TSK_K1_THREADS   128     /* clusters per decode_demux_seed CTA */
#define TSK_K1_SLAB      32      /* cycles per TMA copy of its base-call slab */
#define TSK_K2_THREADS   128
#define TSK_MAX_MODS     64      /* upper bound on max_terminal_modifications */
#define TSK_MAX_RANK     8       /* upper bound on balancer num-positive/negative samples */
#define TSK_NCOUNT       6       /* counters per sample */
#define TSK_PROF_RING    64      /* event sets in flight with tsk_profile_enable() */

/* cluster kinds produced by decode_demux_seed (bit field, one byte per cluster) */
#define KIND_PROCESS   0x01      /* reaches process_polya_signal() and passes the balancer check */
#define KIND_EMIT      0x02      /* seed poly(A) >= signal-analysis-trigger: owns a record slot */
#define KIND_POS       0x04      /* seed poly(A) >= seed-trigger: contrast test / positive sampling */
#define KIND_NEG       0x08      /* negative-sampling candidate (before fair sampling) */
#define KIND_CONTROL   0x10      /* barcode-unassigned with a control configured: to be classified */

/* Device-side sample descriptor (128 bytes), built by the host from tsk_sample. */
struct DevSample {
    uint32_t index_pk[4];        /* index read as 3-bit base codes, 10 per word */
    uint8_t  delim_mask[TSK_MAX_SEQ];   /* per delimiter position: IUPAC bit mask over A,C,G,T */
    uint8_t  fingerprint[TSK_MAX_SEQ];  /* base codes (0-4), 7 = matches nothing */
    int16_t  max_index_mm;
    int16_t  delimiter_pos, delimiter_length, max_delim_mm;
    int16_t  fp_pos, fp_len, max_fp_mm;
    int16_t  has_umi;
    int16_t  limit_threep, dump_len;
    int16_t  pad_[14];
};
static_assert(sizeof(DevSample) == 128, "DevSample must be 128 bytes");

/* Scalar parameters, passed to kernels by value (__grid_constant__). */
struct DevParams {
    int32_t total_cycles, index_start, index_length, threep_start, threep_length;
    int32_t keep_no_delimiter, keep_low_quality_balancer;
    int32_t balancer_start, balancer_length, balancer_min_quality, balancer_min_bases_passes;
    int32_t balancer_minimum_occurrence, npos, nneg;
    int32_t seed_trigger, neg_polya_len, cctr_left, cctr_right;
    float   required_contrast;
    int32_t boundary_pos, sampling_gap, bins;
    int32_t fair_fp_len, fair_hash_size, fair_max_count;
    int32_t w_polyA[5], w_nonA[5];
    int32_t w_small;             /* every w_polyA fits int8: the seed scan's packed-key long stretch applies */
    int32_t max_mods, min_polya_len, sigproc_trigger;
    float   dark_threshold;
    int32_t max_dark_cycles;
    float   maximum_entropy;
    int32_t num_samples;
    int32_t index_words;         /* ceil(index_length / 10) */
    int32_t has_control;
    float   inv[16];             /* inverse colour matrix of the resident tile */
};

/* Work descriptors of the resident tile (device pointers). */
struct TileView {
    const uint8_t *bcl;  size_t bcl_pitch;
    const int16_t *cif;  size_t cif_pitch;     /* elements */
    uint32_t nclusters, firstclusterno;
};

struct EncState;
struct IngestState;

struct tsk_ctx {
    int device;
    EncState *enc;               /* device-side output encoding (tsk_encode.cu), NULL until configured */
    IngestState *ingest;         /* row-wise ingestion + .bcl.gz inflate (tsk_ingest.cu), NULL until used */
    cudaStream_t stream;
    tsk_params hp;               /* host copy of the parameters */
    DevParams dp;
    DevSample *d_samples;        /* [num_samples] */
    float *d_tscore;             /* [201] */
    double *d_rcp;               /* [288] 1.0 / i for the contrast scan */
    void *d_logki;               /* double2 [2048]: (exponent, index) table of the fast kernel's logf (tsk_exact.cuh) */
    int maxdump;                 /* max signal_dump_length over samples */

    /* resident tile */
    TileView tile;
    /* two upload slots: tile t+1 is copied in on the copy stream while tile t is processed */
    uint8_t *d_bcl_own[2];  size_t bcl_cap[2];
    int16_t *d_cif_own[2];  size_t cif_cap[2];
    cudaStream_t cstream;
    cudaEvent_t ev_up[2];
    cudaEvent_t ev_done[2];      /* recorded on the main stream after the last kernel that reads slot u */
    int slot_busy[2];            /* ev_done[u] has been recorded and not yet waited for by an upload */
    int tile_slot;               /* upload slot the resident tile lives in (-1: caller's device memory) */
    int up_slot;                 /* slot of the next host upload */
    struct { TileView tv; float inv[16]; int slot; /* -1: caller's device memory */ } pending[2];
    int npending, tile_valid;
    uint32_t cap_clusters;       /* capacity of the per-cluster buffers below */

    tsk_call *d_calls;           /* [N] */
    uint8_t  *d_kind;            /* [N] */
    uint32_t *d_bucket;          /* [N] fair-sampling hash bucket of KIND_NEG clusters */
    uint32_t *d_worklist;        /* [N] clusters with KIND_PROCESS, ascending */
    uint32_t *d_blkcnt;          /* [nblk][num_samples+1] per-CTA counts, then exclusive bases */
    uint32_t *d_totals;          /* [num_samples+1] */
    uint64_t *d_recbase;         /* [num_samples+1] word offset of each sample's slots */
    uint32_t *d_records;  size_t rec_cap_words;
    float    *d_scratch;  size_t scratch_cap;   /* 2 x scratch_cap floats: per-cycle scores that are read back (tsk_score.cu) */
    uint32_t *d_pos, *d_neg;     /* [T][bins] */
    uint32_t *d_counters;        /* [num_samples][6] */
    uint32_t *d_fair;            /* [hash_size] candidate counts per bucket */
    uint32_t *d_fairslots;  size_t fairslots_cap;   /* [hash_size][fair_max_count] smallest contended clusters per bucket */
    uint32_t *d_lists;           /* undo / contended / accepted lists + their counters */
    /* control classification (tsk_control.cu); allocated only when a control is configured */
    uint32_t *d_ctl_queue;       /* [4]: [1] = next chunk of clusters to look at */
    uint8_t  *d_ctl_index;       /* PhiX target bases | exact-substring hash | q-gram indexes */
    size_t ctl_off_hash, ctl_off_qoff, ctl_off_qpos;
    cudaStream_t xstream;        /* side stream of the classifier */
    cudaEvent_t ev_ctl[2];       /* K1 done -> classifier may start; classifier done */
    uint32_t *d_ruler_pos;  size_t ruler_pos_words;   /* capacity */
    unsigned long long *d_ruler_called;               /* records measured since the last ruler reset */
    uint32_t h_cutoffs[TSK_MAX_CUTOFF_CYCLES];  int h_ncut;   /* what d_cutoffs currently holds */
    int rlen_resident;           /* d_rlen holds the lengths of every slot of the resident tile */
    /* tsk_ruler_tile_all_async() runs on its own stream, next to the decode / scan kernels of the
     * following tile; it reads a snapshot of the slot counts and is joined before that tile's
     * scoring kernel overwrites the record slots */
    cudaStream_t rstream;
    cudaEvent_t ev_rul[2];       /* tile ready for the ruler; ruler done */
    int ruler_inflight;
    uint32_t *d_rtotals;         /* [TSK_MAX_SAMPLES + 1] snapshot of d_totals */
    uint64_t *d_rrecbase;        /* [TSK_MAX_SAMPLES + 1] snapshot of d_recbase */
    int ruler_ncut, ruler_bins;                       /* current shape of the ruler histogram */
    uint32_t *d_cutoffs;         /* [TSK_MAX_CUTOFF_CYCLES] */
    int16_t  *d_rlen;  size_t rlen_cap;
    uint32_t *d_rrec;  size_t rrec_cap;       /* staging of host-provided ruler records */
    uint64_t *d_fit_hist;  size_t fit_cap;   /* pos | neg u64 histograms of fit_cutoffs */
    double   *d_fit_out;   size_t fit_out_cap;

    /* host mirrors filled by tsk_tile_process() */
    uint32_t h_totals[TSK_MAX_SAMPLES + 1];
    uint64_t h_recbase[TSK_MAX_SAMPLES + 1];
    int processed;
    uint64_t launches;
    size_t k1_smem_configured;   /* cudaFuncSetAttribute state of this context's device */
    int k2_attr_set;

    /* optional per-stage device timing (tsk_profile_enable) */
    int profile, own_stream, force_generic;
    /* a ring of event sets, harvested when their work has finished, so that timing does not
     * force the host to wait for the device between tiles */
    struct ProfSlot { cudaEvent_t ev[5]; int nev, first_stage; } prof[TSK_PROF_RING];
    int prof_head, prof_count;   /* oldest pending slot, number pending */
    double stage_ms[6];          /* decode, scan, normalise/score/pack, fix-ups, ruler, fit */
    uint64_t stage_runs[6];
};

void tsk_set_error(const char *fmt, ...);
#define TSK_CUDA(call)                                                              \
    do {                                                                            \
        cudaError_t e_ = (call);                                                    \
        if (e_ != cudaSuccess) {                                                    \
            tsk_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),   \
                          __FILE__, __LINE__);                                      \
            return -1;                                                              \
        }                                                                           \
    } while (0)

/* kernel launchers (tsk_kernels.cu / tsk_ruler.cu / tsk_fit.cu) */
int tsk_launch_import(tsk_ctx *ctx, cudaEvent_t *ev /* optional [5] */);
int tsk_launch_score(tsk_ctx *ctx);
int tsk_score_setup(tsk_ctx *ctx);
void tsk_encode_free(tsk_ctx *ctx);       /* tsk_encode.cu */
void tsk_ingest_free(tsk_ctx *ctx);       /* tsk_ingest.cu */
int tsk_upload_slot_acquire(tsk_ctx *ctx, uint32_t nclusters, uint32_t firstclusterno, const float invmatrix[16],
                            int *slot_out, size_t *bpitch_out, size_t *cpitch_out);
int tsk_upload_slot_commit(tsk_ctx *ctx, int slot);        /* per-context tables of the scoring kernels */
int tsk_launch_fixups(tsk_ctx *ctx);
int tsk_control_setup(tsk_ctx *ctx);      /* builds the PhiX index (tsk_control.cu) */
int tsk_launch_control(tsk_ctx *ctx);
int tsk_launch_ruler(tsk_ctx *ctx, const uint32_t *d_records, size_t nrecords, int dump_len,
                     int ncutoffs, int min_len, float weight, int bins, float gap,
                     int16_t *d_len_out);
int tsk_launch_ruler_resident(tsk_ctx *ctx, cudaStream_t st, int ncutoffs, int min_len, float weight, int bins,
                              float gap, int16_t *d_len_out);
int tsk_ruler_join(tsk_ctx *ctx);         /* main stream waits for a ruler still running on rstream */
int tsk_launch_ruler_segs(tsk_ctx *ctx, const uint32_t *d_records, int nseg, const uint32_t *first,
                          const uint64_t *base, const int *dump, int ncutoffs, int min_len, float weight,
                          int bins, float gap, int16_t *d_len_out);
int tsk_launch_fit(tsk_ctx *ctx, const uint64_t *d_pos, const uint64_t *d_neg, int cycles, int bins,
                   uint64_t min_spots, double bw, double low, const double *high, int nhigh,
                   double *d_out);

#endif
