/*
 * tailseeker_b200.h -- C ABI of the B200-native TAIL-seq signal-processing hot path.
 *
 * This is the drop-in boundary.  The reference (hyeshik/tailseeker 3.2.1) has no
 * plugin/FFI API: its two native programs are driven by argv + files.  The seam this
 * library replaces sits one call level below those programs' main():
 *
 *   tailseq-import:       distribute_processing(cfg, intensities, basecalls, firstclusterno)
 *                         -> run_spot_processing -> process_spots(...)
 *                         (reference src/importer/tailseq-import.c:516-565, :414-513,
 *                          src/importer/spotanalyzer.c:617-755)
 *   tailseq-polya-ruler:  process_polya_ruling(...) -> measure_polya_length(...)
 *                         (reference src/polyaruler/polyaruler.c:201-314, :104-152)
 *   cutoff estimator:     calculate_optimal_cutoffs / find_eqodds_point
 *                         (reference scripts/calculate-optimal-parameters.py:57-171)
 *
 * Conventions follow that seam: every function returns 0 on success and -1 on error
 * (message via tsk_last_error()), the caller owns every buffer it passes, parameters
 * are a read-only POD mirror of the reference's config structs
 * (src/importer/tailseq-import.h:172-215,220-274).  No torch / C++ types appear here.
 *
 * There is NO CPU fallback: every entry point that computes runs CUDA kernels built for
 * sm_100a and fails with -1 when no such device is usable.
 */
#ifndef TAILSEEKER_B200_H
#define TAILSEEKER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TSK_ABI_VERSION          3

#define TSK_NUM_CHANNELS         4      /* tailseq-import.h:40 */
#define TSK_MAX_SAMPLES          64
#define TSK_MAX_SEQ              32     /* longest index / delimiter / fingerprint */
#define TSK_T_SCORE_BINS         200    /* T_INTENSITY_SCORE_BINS, tailseq-import.h:206 */
#define TSK_SCORE_MAX            ((1u << 24) - 1u)   /* signal-packs.h:36-37 */
#define TSK_MAX_CUTOFF_CYCLES    1024   /* MAX_NUM_CYCLES, polyaruler.c:39 */
#define TSK_MAX_CONTROL_READ     64     /* longest phix-match-length */

/* PAFLAG_* bits, src/sigproc-flags.h:30-43 (same values, same meaning). */
#define TSK_PAFLAG_POLYA_DETECTED              0x0001
#define TSK_PAFLAG_DELIMITER_HAS_MISMATCH      0x0002
#define TSK_PAFLAG_HAVE_3P_MODIFICATION        0x0004
#define TSK_PAFLAG_MEASURED_FROM_FLUORESCENCE  0x0008
#define TSK_PAFLAG_BARCODE_HAS_MISMATCHES      0x0010
#define TSK_PAFLAG_DELIMITER_IS_SHIFTED        0x0020
#define TSK_PAFLAG_DARKCYCLE_EXISTS            0x0040
#define TSK_PAFLAG_DELIMITER_NOT_FOUND         0x0080
#define TSK_PAFLAG_BALANCER_CALL_QUALITY_BAD   0x0100
#define TSK_PAFLAG_BALANCER_BIASED             0x0200
#define TSK_PAFLAG_BALANCER_SIGNAL_BAD         0x0400
#define TSK_PAFLAG_DARKCYCLE_OVER_THRESHOLD    0x0800

/* One entry of the sample list, in the reference's list order after
 * compute_derived_values() (parseconfig.c:575-641): entry 0 is "Unknown", entry 1 is
 * the PhiX pseudo-sample when a control is configured, then the [sample:*] sections in
 * REVERSE file order (LIFO, parseconfig.c:407-409).  Positions are absolute 0-based
 * cycles (delimiter_pos already rebased by threep_start, parseconfig.c:638-639). */
typedef struct tsk_sample {
    char    index[TSK_MAX_SEQ];          /* NUL-terminated, "XXXXXX" for pseudo-samples */
    char    delimiter[TSK_MAX_SEQ];      /* IUPAC, "" when none */
    char    fingerprint[TSK_MAX_SEQ];    /* "" when none */
    int32_t maximum_index_mismatches;
    int32_t delimiter_pos;               /* -1 when none */
    int32_t delimiter_length;
    int32_t maximum_delimiter_mismatches;
    int32_t fingerprint_pos;
    int32_t fingerprint_length;
    int32_t maximum_fingerprint_mismatches;
    int32_t umi_ranges_count;            /* only ">0" matters on this path (spotanalyzer.c:697) */
    int32_t limit_threep_processing;     /* 0 = no limit */
    int32_t signal_dump_length;          /* parseconfig.c:630-636 */
} tsk_sample;

/* POD mirror of BalancerParameters / PolyASeederParameters / PolyAFinderParameters /
 * PolyARulerParameters and the read layout (tailseq-import.h:172-215,226-236). */
typedef struct tsk_params {
    int32_t abi_version;                 /* = TSK_ABI_VERSION */

    /* [read_format] -- 0-based starts */
    int32_t total_cycles;
    int32_t index_start, index_length;
    int32_t threep_start, threep_length;

    /* [options] */
    int32_t keep_no_delimiter;
    int32_t keep_low_quality_balancer;

    /* [control]: 1 when the PhiX pseudo-sample (list entry 1) exists.  Barcode-unassigned
     * clusters are then classified against the PhiX174 genome (controlaligner.c:107-145):
     * control_read_length bases from cycle 0 by exact substring, else the same number of
     * bases from control_first_cycle by local alignment (match 1, mismatch 2, gap 4+1,
     * score >= (int)(control_read_length * 0.65), tailseq-import.h:151-156). */
    int32_t has_control;
    int32_t control_first_cycle;         /* phix-match-start - 1 (parseconfig.c:195) */
    int32_t control_read_length;         /* phix-match-length, 8..64 */

    /* [balancer] */
    int32_t balancer_start, balancer_length;   /* end = start + length */
    int32_t balancer_min_quality;
    int32_t balancer_min_bases_passes;         /* (int)(length * min_fraction), parseconfig.c:544 */
    int32_t balancer_minimum_occurrence;
    int32_t balancer_num_positive_samples;
    int32_t balancer_num_negative_samples;

    /* [polyA_seeder] */
    int32_t seed_trigger_polya_length;
    int32_t negative_sample_polya_length;
    int32_t max_cctr_scan_left_space, max_cctr_scan_right_space;
    float   required_cdf_contrast;
    int32_t polya_boundary_pos, polya_sampling_gap;
    int32_t dist_sampling_bins;
    int32_t fair_sampling_fingerprint_length;
    int32_t fair_sampling_hash_space_size;
    int32_t fair_sampling_max_count;           /* uint8_t in the reference */

    /* [polyA_finder]: weights indexed A,C,G,T,N */
    int32_t weights_polyA[5];
    int32_t weights_nonA[5];
    int32_t max_terminal_modifications;
    int32_t min_polya_length;
    int32_t sigproc_trigger_polya_length;

    /* [polyA_ruler] */
    float   dark_cycles_threshold;
    int32_t max_dark_cycles;
    float   maximum_entropy;                       /* signalproc.c:54-64,73 (host libm) */
    float   t_intensity_score[TSK_T_SCORE_BINS + 1]; /* signalproc.c:75-81 (host libm expf) */

    int32_t num_samples;
    tsk_sample samples[TSK_MAX_SAMPLES];
} tsk_params;

/* Per-cluster call, the data the reference keeps in locals of process_spots()
 * (spotanalyzer.c:641-643) and hands to write_measurements_to_buffers() (:725-727). */
typedef struct tsk_call {
    int16_t  sample;         /* list index (numindex); always valid */
    uint16_t flags;          /* procflags, TSK_PAFLAG_* */
    int16_t  delimiter_end;  /* absolute cycle, -1 = none */
    int16_t  polya_len;      /* polya_status: seed poly(A) length, or -1 */
    int16_t  terminal_mods;  /* -1 when process_polya_signal() was not reached */
    uint8_t  written;        /* 1 = reference writes taginfo/seqqual lines for it */
    uint8_t  emitted;        /* 1 = a sigpack record exists for it */
    uint32_t record_slot;    /* slot within its sample's record array (valid if emitted) */
} tsk_call;                  /* 16 bytes */

/* Per-sample demultiplexing counters (tailseq-import.h:126-131, CSV tailseq-import.c:698-709). */
typedef struct tsk_counters {
    uint32_t clusters_mm0, clusters_mm1, clusters_mm2plus;
    uint32_t clusters_nodelim, clusters_fpmismatch, clusters_qcfailed;
} tsk_counters;

typedef struct tsk_ctx tsk_ctx;

/* Last error message of the calling thread ("" if none). */
const char *tsk_last_error(void);

/* Library / device probe: returns ABI version, or -1 if no sm_100-class device is usable. */
int tsk_probe(int device);

/* ---- context ------------------------------------------------------------------- */
/* Replaces parse_config()'s product (a read-only TailseekerConfig shared by all workers,
 * tailseq-import.c:741, parseconfig.c:683-708): validates and uploads the parameters. */
int tsk_ctx_create(tsk_ctx **out, int device, const tsk_params *params);
int tsk_ctx_destroy(tsk_ctx *ctx);

/* ---- importer path --------------------------------------------------------------- */
/* Replaces load_intensities_and_basecalls()'s buffers (tailseq-import.c:203-228) minus the
 * AoS transpose of cifreader.c:163-164: data stay in the files' native layouts,
 *   bcl[total_cycles][bcl_pitch] u8,   cif[threep_length][4][cif_pitch] i16  (A,C,G,T planes)
 * pitches in ELEMENTS, >= nclusters.  on_device != 0: the pointers are device pointers that
 * the caller keeps alive AND UNMODIFIED until that tile has been processed and the device has
 * finished with it, i.e. until a getter or tsk_synchronize() has returned after the tile's
 * tsk_tile_process[_async]() -- not merely until the next upload: two tiles may be pending and
 * tsk_tile_process_async() returns before its kernels run (zero copy) -- bcl must then be 16-byte aligned
 * with bcl_pitch a multiple of 16 (the rows are fetched by TMA tensor copies; refused otherwise),
 * and cif 16-byte aligned with cif_pitch a multiple of 8 for the fast scoring kernel; otherwise
 * they are host pointers (pinned or pageable) and are copied to HBM on the context's stream.
 * invmatrix = inverse colour matrix (load_color_matrix(), signalproc.c:345-375).
 * Up to two tiles may be uploaded ahead of tsk_tile_process(), which takes them in order: the
 * copy of tile t+1 (on its own stream) then overlaps the kernels of tile t; the copy into an
 * upload slot waits (on the device) for the kernels of the tile that used the slot before.
 * An upload larger than every tile before it re-allocates the per-cluster result buffers: the
 * call then waits for everything queued on the device, and results of the previous tile that
 * have not been read by then (records, calls, lengths) are discarded -- histograms and counters
 * are kept.  Upload the largest tile first, or read each tile's results before the next upload,
 * when tile sizes vary. */
int tsk_tile_upload(tsk_ctx *ctx, const uint8_t *bcl, size_t bcl_pitch,
                    const int16_t *cif, size_t cif_pitch,
                    const float invmatrix[16],
                    uint32_t nclusters, uint32_t firstclusterno, int on_device);

/* The same hand-over cycle file by cycle file (replaces the loops of load_intensities_and_basecalls(),
 * tailseq-import.c:203-228, and load_cif_data() / load_bcl_data(), cifreader.c:118-170,
 * bclreader.c:95-119): between tsk_tile_begin() and tsk_tile_end() every row of the tile is sent as
 * soon as the host has it -- asynchronous copies on the context's copy stream, so file reads,
 * transfers and the previous tile's kernels overlap; the put calls may come from several reader
 * threads at once.  A gzip-compressed base-call file (".bcl.gz", bclreader.c:60-72) is handed over
 * COMPRESSED and inflated on the device: `gz` holds the whole file, `inflated_size` what it inflates
 * to (4 + clusters in the file), `skip` the offset of this block's first cluster in that stream
 * (4 + first cluster of the block).  Host buffers passed to the put calls must stay unchanged until
 * tsk_tile_end() returns; it fails (-1, message names the cycle) for a corrupt or short stream.
 * The tile then waits for tsk_tile_process() like one given to tsk_tile_upload(). */
int tsk_tile_begin(tsk_ctx *ctx, uint32_t nclusters, uint32_t firstclusterno, const float invmatrix[16]);
int tsk_tile_put_bcl(tsk_ctx *ctx, int cycle, const uint8_t *calls /* [nclusters] */);
int tsk_tile_put_bcl_gz(tsk_ctx *ctx, int cycle, const uint8_t *gz, size_t gzbytes, uint64_t inflated_size, uint64_t skip);
int tsk_tile_put_cif(tsk_ctx *ctx, int threep_cycle, const int16_t *planes /* [4][plane_pitch] */, size_t plane_pitch);
/* Page-locked staging for the readers of the cycle files (the reference reads each file into its row of
 * one big pageable array, bclreader.c:95-119 / cifreader.c): a slot of at least `bytes`, to be filled, passed
 * to tsk_tile_put_bcl / tsk_tile_put_cif and released; the library hands it out again only after that
 * copy has completed, so a released slot needs no further care (unlike caller-owned buffers, above).
 * Thread-safe; blocks while all slots (24) are in use. */
int tsk_tile_stage_acquire(tsk_ctx *ctx, size_t bytes, void **ptr, int *slot);
int tsk_tile_stage_release(tsk_ctx *ctx, int slot);
int tsk_tile_end(tsk_ctx *ctx);

/* Replaces distribute_processing() (tailseq-import.c:516-565): runs the per-cluster path
 * for the uploaded block.  Like prepare_split_jobs() (:340-392) it starts from zeroed
 * pos/neg histograms and a zeroed fair-sampling table; counters accumulate across calls
 * (they live in SampleInfo in the reference) until tsk_counters_reset(). */
int tsk_tile_process(tsk_ctx *ctx);
/* The same work, only enqueued on the context's stream: returns without waiting for the device,
 * so that a caller can queue tile after tile (tsk_tile_upload / _process_async /
 * tsk_ruler_tile_all_async) and keep the GPU busy back to back.  Any getter below waits. */
int tsk_tile_process_async(tsk_ctx *ctx);

/* Results of the last tsk_tile_process().  All copies are synchronous w.r.t. the call. */
int tsk_tile_get_calls(tsk_ctx *ctx, tsk_call *calls /* [nclusters] */);
/* Number of record slots of one sample (some may be invalid, see below). */
int tsk_tile_get_record_count(tsk_ctx *ctx, int sample, uint32_t *nslots);
/* Record slots of one sample in cluster order, each (2 + signal_dump_length) u32 words:
 * SignalRecordHeader {u32 clusterno; i16 first_cycle; i16 valid_cycle_count} followed by
 * signal_dump_length packets (bit0 downhill, bits1-24 score), signal-packs.h:30-41.
 * A slot whose valid_cycle_count is negative was withdrawn (the reference returns before
 * write_polya_score() on a dark-cycle abort, spotanalyzer.c:561-565) and must be skipped. */
int tsk_tile_get_records(tsk_ctx *ctx, int sample, uint32_t *records, size_t capacity_words);
int tsk_get_counters(tsk_ctx *ctx, tsk_counters *counters /* [num_samples] */);
int tsk_counters_reset(tsk_ctx *ctx);
/* pos / neg score histograms u32[total_cycles][dist_sampling_bins]
 * (GloballyAggregatedOutput, tailseq-import.h:291-294). */
int tsk_get_hist(tsk_ctx *ctx, uint32_t *pos, uint32_t *neg);
/* Device addresses of the same histograms and counters, for an in-place all-reduce by the
 * caller (NCCL / torch.distributed) -- the cross-tile sum of
 * calculate-optimal-parameters.py:139-140 and stats-merge-demultiplexing-counts.py:33-53. */
int tsk_get_device_ptrs(tsk_ctx *ctx, void **pos_hist, void **neg_hist, void **counters,
                        void **ruler_pos_hist);

/* ---- output encoding on the device ----------------------------------------------------- */
/* Replaces what the reference's workers do after the per-cluster analysis: format_basecalls()
 * (bclreader.c:146-166), write_seqqual_entry() / write_taginfo_entry() (spotanalyzer.c:206-307),
 * the record bytes of write_polya_score() (:389-438) and the BGZF compression of all three
 * (sync_write_out_buffer() -> bgzf_write(), :355-386, :419-428) -- 71 % of the reference's
 * single-thread time.  The tile's text is formatted from the resident calls and base calls and
 * every stream is deflated on the GPU as a series of BGZF blocks (RFC 1952 members with the `BC`
 * extra field, one dynamic-Huffman block each); the host only write()s the bytes it gets back.
 * Decompressed, a stream is byte for byte what the reference writes with threads = 1. */
#define TSK_ENCODE_MAX_UMI 8
typedef struct tsk_encode_sample {
    int32_t want_seqqual, want_taginfo, want_signal;   /* which of the sample's files exist (none for
                                                          Unknown / control, tailseq-import.c:46-48) */
    int32_t umi_count;                                 /* [sample:*] umi ranges, absolute 0-based cycles */
    int32_t umi_start[TSK_ENCODE_MAX_UMI], umi_length[TSK_ENCODE_MAX_UMI];
} tsk_encode_sample;
typedef struct tsk_encode_layout {
    int32_t fivep_start, fivep_length;                 /* 0-based */
    int32_t threep_seqqual_output_length;              /* parseconfig.c:491 */
    tsk_encode_sample samples[TSK_MAX_SAMPLES];        /* list order of tsk_params.samples */
} tsk_encode_layout;
enum { TSK_STREAM_SEQQUAL = 0, TSK_STREAM_TAGINFO = 1, TSK_STREAM_SIGNAL = 2 };
int tsk_encode_configure(tsk_ctx *ctx, const tsk_encode_layout *layout);
/* Optional: page-lock the host buffer of the encoded streams for blocks of `nclusters` clusters now (on a
 * start-up thread beside the first tile's upload, say) rather than inside the first tsk_tile_encode().
 * Not to be called while tsk_encode_configure() / tsk_tile_encode() run on the same context. */
int tsk_encode_reserve(tsk_ctx *ctx, uint32_t nclusters);
/* Encodes the outputs of the tile last processed (any block of it: cluster numbers start at the
 * block's firstclusterno) and brings the compressed streams to pinned host memory. */
int tsk_tile_encode(tsk_ctx *ctx);
/* One stream of the last tsk_tile_encode(): *data (host memory owned by the context, valid until
 * the next tsk_tile_encode()) holds *nbytes bytes of whole BGZF blocks to append to the sample's
 * file; 0 bytes when the sample wrote nothing.  The signal stream holds the records only: the
 * 12-byte header {4, total clusters, dump length} (tailseq-import.c:106-110) is the caller's. */
int tsk_tile_get_encoded(tsk_ctx *ctx, int sample, int stream, const uint8_t **data, size_t *nbytes);
/* Device time of the last tsk_tile_encode() (CUDA events around its kernels). */
int tsk_encode_elapsed_ms(tsk_ctx *ctx, double *ms);
/* The same deflater for any host bytes: `n` bytes -> a run of whole BGZF blocks in pinned host memory
 * owned by the object (valid until its next call).  tailseq-writefastq's output side: replaces
 * bgzf_write() on its two FASTQ streams (src/exporter/tailseq-writefastq.c:148-240, bgzf_mt :387,399). */
typedef struct tsk_bgzf tsk_bgzf;
int tsk_bgzf_create(tsk_bgzf **out, int device);
void tsk_bgzf_destroy(tsk_bgzf *z);
int tsk_bgzf_deflate(tsk_bgzf *z, const void *data, size_t n, const uint8_t **out, size_t *outn);
uint64_t tsk_bgzf_launch_count(tsk_bgzf *z);
/* Self-test: `n` arbitrary host bytes deflated by the same kernels as one stream; the BGZF blocks
 * (<= n + n / 1000 + 64 bytes per started 32 KB) are written to out[0 .. *outn). */
int tsk_selftest_deflate(tsk_ctx *ctx, const uint8_t *data, size_t n, uint8_t *out, size_t capacity, size_t *outn);
/* The 28-byte empty BGZF block that ends a file (htslib's end-of-file marker). */
int tsk_bgzf_eof_block(const uint8_t **data, size_t *nbytes);

/* ---- ruler path -------------------------------------------------------------------- */
/* load_score_cutoffs()'s float->24-bit conversion (polyaruler.c:93-96) for one value. */
uint32_t tsk_cutoff_to_packed(double cutoff);

/* Replaces process_polya_ruling()+measure_polya_length() (polyaruler.c:201-314,104-152) for
 * `nrecords` record slots of `dump_len` packets each.  records: host pointer, or (on_device)
 * device pointer.  Outputs: polya_len_out[i] per slot (-1 when below minimum_polya_len or
 * withdrawn slot; host pointer), and the positive-sample histogram
 * u32[ncutoffs][dist_sampling_bins] ADDED into the context's ruler histogram
 * (zero it with tsk_ruler_hist_reset(); a call with another (ncutoffs, bins) shape resets it). */
int tsk_ruler_records(tsk_ctx *ctx, const uint32_t *records, size_t nrecords, int dump_len,
                      const uint32_t *cutoffs, int ncutoffs,
                      int minimum_polya_len, float downhill_ext_weight,
                      int dist_sampling_bins, float dist_sampling_gap,
                      int16_t *polya_len_out, int on_device);
/* Same, on the resident record slots of `sample` from the last tsk_tile_process(). */
int tsk_ruler_tile(tsk_ctx *ctx, int sample, const uint32_t *cutoffs, int ncutoffs,
                   int minimum_polya_len, float downhill_ext_weight,
                   int dist_sampling_bins, float dist_sampling_gap,
                   int16_t *polya_len_out /* [nslots] host */);
/* Same for EVERY sample of the resident tile in one launch (what measure-polya-lengths.py:53-70
 * loops over): polya_len_out holds the slots of sample 0, then sample 1, ...;
 * slot_offsets[num_samples + 1] receives where each sample's slots start. */
int tsk_ruler_tile_all(tsk_ctx *ctx, const uint32_t *cutoffs, int ncutoffs,
                       int minimum_polya_len, float downhill_ext_weight, int dist_sampling_bins,
                       float dist_sampling_gap, int16_t *polya_len_out, uint32_t *slot_offsets);
/* tsk_ruler_tile_all() in two halves.  The first only enqueues the measurement of every record
 * slot of the resident tile (the per-sample slot counts are read on the device, so it may
 * directly follow tsk_tile_process_async()); lengths stay in device memory and the
 * histogram / called counter accumulate.  The second waits and copies the lengths out. */
int tsk_ruler_tile_all_async(tsk_ctx *ctx, const uint32_t *cutoffs, int ncutoffs,
                             int minimum_polya_len, float downhill_ext_weight, int dist_sampling_bins,
                             float dist_sampling_gap);
int tsk_ruler_get_lengths(tsk_ctx *ctx, int16_t *polya_len_out, uint32_t *slot_offsets);
/* Number of records whose measured length reached minimum_polya_len (the lines
 * process_polya_ruling() prints with a length, polyaruler.c:289-299) since the last
 * tsk_ruler_hist_reset(). */
int tsk_ruler_get_called(tsk_ctx *ctx, uint64_t *ncalled);
int tsk_ruler_hist_reset(tsk_ctx *ctx, int ncutoffs, int dist_sampling_bins);
int tsk_ruler_get_hist(tsk_ctx *ctx, uint32_t *pos /* [ncutoffs][bins] */);

/* ---- cutoff estimation --------------------------------------------------------------- */
/* Replaces find_all_eqodds_point() (calculate-optimal-parameters.py:57-78): for every
 * cycle with >= min_spots positive and negative samples, the root of the weighted-KDE
 * log-likelihood ratio found by bisection in [low, high[k]] (first k whose bracket has a
 * sign change and whose root lies strictly inside).  pos/neg: u64[cycles][bins] host.
 * cutoffs_out[c] = root, or NaN when the cycle has too few spots or no root. */
int tsk_fit_cutoffs(tsk_ctx *ctx, const uint64_t *pos, const uint64_t *neg,
                    int cycles, int bins, uint64_t min_spots, double kde_bandwidth,
                    double search_low, const double *search_high, int nsearch_high,
                    double *cutoffs_out);

/* ---- cross-tile sums on the devices (one tile set per GPU) ----------------------------- */
/* Replaces the sums the reference forms over per-tile files: pos/neg histograms before cutoff
 * estimation (scripts/calculate-optimal-parameters.py:139-140), the ruler histogram per round
 * (scripts/measure-polya-lengths.py:64-70), the demultiplexing counters
 * (scripts/stats-merge-demultiplexing-counts.py:33-53).  Integer sums: the result equals the
 * single-device one bit for bit.
 *
 * NCCL communicators are passed as void* (an ncclComm_t).  Callers that have none can make one
 * here; NCCL is loaded at run time ("libnccl.so.2", or the file named by TSK_NCCL_LIB). */
#define TSK_NCCL_ID_BYTES 128
int tsk_nccl_unique_id(uint8_t id[TSK_NCCL_ID_BYTES]);                 /* rank 0, then hand to all */
int tsk_nccl_init_rank(void **comm, int device, int nranks, int rank, const uint8_t id[TSK_NCCL_ID_BYTES]);
int tsk_nccl_init_all(void **comms, int ndev, const int *devices);    /* one process driving ndev GPUs */
int tsk_nccl_destroy(void *comm);
int tsk_nccl_group_start(void);      /* one thread issuing the collective for several devices */
int tsk_nccl_group_end(void);
/* The context's pos / neg histograms, demultiplexing counters and ruler histogram summed over
 * the communicator's ranks, in place, as ONE grouped collective on the context's stream. */
int tsk_hist_allreduce(tsk_ctx *ctx, void *nccl_comm);

/* Run-wide accumulators (u64, like the Python's int64 sums) on one device, fed tile by tile without
 * leaving the device; with max_tiles > 0 the tiles' own histograms are kept too (per-tile fits). */
typedef struct tsk_merge tsk_merge;
int tsk_merge_create(tsk_merge **out, int device, int cycles, int bins, int nsamples, int max_tiles);
void tsk_merge_destroy(tsk_merge *m);
/* pos / neg of the tile last processed by ctx (enqueued on its stream: may follow tsk_tile_process_async) */
int tsk_merge_add_tile(tsk_merge *m, tsk_ctx *ctx);
/* the context's ruler histogram and counters (they accumulate inside the context): once, at the end */
int tsk_merge_add_totals(tsk_merge *m, tsk_ctx *ctx);
/* ONE ncclAllReduce(uint64, sum) over the packed accumulators.  ctx may be NULL here and in
 * tsk_merge_get() (the contexts of a run are gone by then): the device is synchronised first. */
int tsk_merge_allreduce(tsk_merge *m, tsk_ctx *ctx, void *nccl_comm);
/* host copies: u64 [cycles][bins] x 3, u64 [nsamples][6] (tsk_counters order); any may be NULL */
int tsk_merge_get(tsk_merge *m, tsk_ctx *ctx, uint64_t *pos, uint64_t *neg, uint64_t *ruler_pos, uint64_t *counters);
/* what: 1 pos/neg + kept tiles, 2 ruler histogram, 4 counters */
int tsk_merge_reset(tsk_merge *m, tsk_ctx *ctx, int what);
int tsk_merge_tile_count(tsk_merge *m);
/* find_all_eqodds_point() on device memory: cutoffs of the run-wide histograms (first `cycles`
 * values) and of every kept tile (then `cycles` per tile, in the order they were added).
 * positive = 0: importer positives; 1: the ruler's run-wide positives (main.py:214-215). */
int tsk_merge_fit(tsk_merge *m, tsk_ctx *ctx, int positive, uint64_t min_spots, double kde_bandwidth,
                  double search_low, const double *search_high, int nsearch_high, double *cutoffs_out);

/* ---- duplicate collapsing: tailseq-dedup-perfect ------------------------------------- */
/* Replaces the per-UMI-group loop of src/deduplicator/tailseq-dedup-perfect.c (main()
 * :368-400, tqueue_append() :168-214, process_tag_duplicates() :277-345, tag priority
 * :87-105) and, with sort_by_umi, the `sort -t "\t" -k6,6 -k1,1 -k2,2n` that feeds it
 * (tailseeker/main.py:272-278).  One record = one taginfo line as parse_line() (:216-274)
 * reads it: flags, poly(A) length, UMI.  Tile name, cluster number and modification text stay
 * with the caller (they are only printed).
 *
 * UMI key: 3 bits per base, first base in the most significant bits of key[0], 21 bases per
 * word (bit 63 unused), end-of-string 0, A 1, C 2, G 3, N 4, T 5, so comparing (key[0], key[1])
 * as integers is strcmp() on the strings (what `sort` does in the C locale).  <= 42 bases.
 * Returns -1 for a longer UMI or another character. */
#define TSK_DEDUP_MAX_UMI 42
int tsk_dedup_pack_umi(const char *umi, size_t length, uint64_t key[2]);

typedef struct tsk_dedup tsk_dedup;
/* Device buffers for up to `capacity` records (< 2^32). */
int tsk_dedup_create(tsk_dedup **out, int device, uint64_t capacity);
void tsk_dedup_destroy(tsk_dedup *d);
/* Host or device pointers, n <= capacity. */
int tsk_dedup_upload(tsk_dedup *d, uint64_t n, const uint64_t *umi_hi, const uint64_t *umi_lo,
                     const int32_t *flags, const int32_t *polya);
/* sort_by_umi == 0: the records are taken in the given order and a group is a run of equal
 * UMIs, exactly like the reference program reading its (pre-sorted) stdin.
 * sort_by_umi != 0: records are first ordered by UMI key, ties in the given order (a stable
 * sort: when the input is the tiles' taginfo files concatenated in (tile, cluster) order this
 * is the reference's `sort -k6,6 -k1,1 -k2,2n`).  "Position" below means position in that order. */
int tsk_dedup_run(tsk_dedup *d, int sort_by_umi, uint64_t *n_groups);
/* Per position p (arrays of n, any may be NULL):
 *   order[p]       index of the record in the uploaded arrays (p itself without sorting)
 *   group[p]       number of its UMI group, 0-based in order of appearance (`next_group`, :71)
 *   trace_value[p] last column of its line in the duplicate trace: -3 dropped for a
 *                  lower tag priority (:184,208), -2 kept but not the representative (:335),
 *                  else the poly(A) length reported for the group (:291,335)
 *   trace_slot[p]  0-based line number of that line in the trace output */
int tsk_dedup_get_records(tsk_dedup *d, uint32_t *order, uint32_t *group, int32_t *trace_value,
                          uint32_t *trace_slot);
/* Per group g (arrays of n_groups, any may be NULL):
 *   rep[g]    position of the representative: the first highest-priority record nearest to
 *             the mean poly(A) length of the highest-priority records (:298-315)
 *   polya[g]  (int)roundf(mean), -1 if the representative's own length is negative (:323);
 *             the record's own length when kept[g] == 1 (its line is printed as is, :286-287)
 *   ndup[g]   records in the group (`num_duplicates`)
 *   kept[g]   records of the highest priority (the queue length at :282) */
int tsk_dedup_get_groups(tsk_dedup *d, uint32_t *rep, int32_t *polya, uint32_t *ndup, uint32_t *kept);
/* Positions (ascending) of the records the reference loses: its queue has 2048 slots for the whole
 * run, grows by half when a record arrives at a full queue, and tqueue_expand() (:142-166) does
 * not carry the arriving record over, which continues as tile "", cluster 0, flags 0, length 0,
 * no modifications.  Lengths, representatives and trace values above already account for it; the
 * caller prints such a record with those empty fields.  At most 56 per run; none unless more than
 * 2047 records of one UMI share a priority.  positions may be NULL to ask for the count. */
int tsk_dedup_get_lost(tsk_dedup *d, uint32_t *positions, uint32_t capacity, uint32_t *count);
/* Device time of the last tsk_dedup_run (CUDA events on its stream): sort, grouping + collapse. */
int tsk_dedup_elapsed_ms(tsk_dedup *d, float *sort_ms, float *collapse_ms);
/* Number of kernel launches issued by tsk_dedup_run so far. */
uint64_t tsk_dedup_launch_count(tsk_dedup *d);

/* ---- stream control (for callers that overlap copies and kernels) -------------------- */
int tsk_synchronize(tsk_ctx *ctx);
/* Timed re-run of the kernels of tsk_tile_process() on the resident tile: CUDA events on the
 * context's stream (the stream the kernels are launched on), `iters` back-to-back runs,
 * mean ms per run per stage in ms_out[5] = {decode_demux_seed, slot scan + assign,
 * normalise_score_pack, fair-sampling fix-ups, total}.  Counters are left as one run
 * leaves them. */
int tsk_tile_process_timed(tsk_ctx *ctx, int iters, float ms_out[5]);
/* Use the caller's CUDA stream (a cudaStream_t, e.g. torch's current stream) for every copy
 * and kernel of this context instead of the context's own stream. */
int tsk_set_stream(tsk_ctx *ctx, void *cuda_stream);
/* Per-stage device time, accumulated with CUDA events on the context's stream while
 * enabled: stage_ms[6] / stage_runs[6] = {decode_demux_seed, slot scan + assign,
 * normalise_score_pack, fair-sampling fix-ups, polya_ruler, fit_cutoffs}.  Enabling resets. */
int tsk_profile_enable(tsk_ctx *ctx, int on);
int tsk_profile_get(tsk_ctx *ctx, double stage_ms[6], uint64_t stage_runs[6]);
/* Number of kernel launches issued by this context so far. */
uint64_t tsk_launch_count(tsk_ctx *ctx);

/* Self-test: out[i] = the device's logf(in[i]) (finite in[i] > 0), which must equal the host
 * libm's logf bit for bit because shannon_entropy() (signalproc.c:39-51) feeds 24-bit
 * truncated scores.  Host pointers. */
int tsk_selftest_logf(tsk_ctx *ctx, const float *in, float *out, size_t n);
/* Self-test sweeps on the device: (1) the FMA form of logf used by the fast scoring kernel
 * against the separately-rounded form above on EVERY normal float in (0,1]; (2) the
 * reciprocal-multiply-and-correct division against IEEE division on `div_pairs` pseudo-random
 * operand pairs.  Both mismatch counts must be 0. */
int tsk_selftest_sweeps(tsk_ctx *ctx, uint64_t div_pairs, uint64_t *logf_mismatches,
                        uint64_t *div_mismatches);
/* Debug: run normalise_score_pack with the literal operation-by-operation kernel for every
 * cluster instead of the fast kernel (tests pin one against the other and the oracle). */
int tsk_debug_force_generic(tsk_ctx *ctx, int on);

#ifdef __cplusplus
}
#endif
#endif /* TAILSEEKER_B200_H */
