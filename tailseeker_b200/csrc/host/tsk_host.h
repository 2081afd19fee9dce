/*
 * tsk_host.h -- C host side of the drop-in programs tailseq-import / tailseq-polya-ruler.
 *
 * Keeps the reference's command-line, INI and file-format surface (src/importer/tailseq-import.c,
 * parseconfig.c, cifreader.c, bclreader.c, spotanalyzer.c writers; src/polyaruler/polyaruler.c) and
 * hands the per-cluster work to libtailseeker_b200.so through include/tailseeker_b200.h.
 * There is no CPU implementation of the per-cluster path here.
 */
#ifndef TSK_HOST_H
#define TSK_HOST_H

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <zlib.h>
#include "../../../include/tailseeker_b200.h"

#define TSK_MAX_UMI 8
#define TSK_NAME_MAX 128

struct host_umi { int start, length; };

struct host_sample {
    char name[TSK_NAME_MAX];
    char *index, *delimiter, *fingerprint;       /* NULL when the key was never given */
    int delimiter_pos, delimiter_length;
    int fingerprint_pos, fingerprint_length;
    int maximum_index_mismatches, maximum_delimiter_mismatches, maximum_fingerprint_mismatches;
    struct host_umi umi[TSK_MAX_UMI];
    int umi_count;
    int limit_threep_processing, signal_dump_length;
    int numindex;
    FILE *stream_seqqual, *stream_taginfo, *stream_signal;   /* series of gzip members (writers.c) */
    struct host_sample *next;
};

struct host_altcall {
    char *filename;
    int first_cycle;                    /* 0-based */
    struct host_altcall *next;
};

struct host_config {
    /* [source] */
    char *datadir, *laneid, *threep_colormatrix_filename;
    int lane, tile;
    /* [read_format] */
    int total_cycles, fivep_start, fivep_length, index_start, index_length, threep_start, threep_length;
    int threep_seqqual_output_length;
    /* [options] */
    int keep_no_delimiter, keep_low_quality_balancer, threads;
    size_t read_buffer_size;
    int read_buffer_entry_count;
    /* [output] */
    char *seqqual_output, *taginfo_output, *signal_output, *signal_dists_output, *stats_output,
         *length_dists_output;
    /* [alternative_calls]: FASTQ files that override the base calls from `first_cycle` on; newest
     * entry first, like the reference's list (parseconfig.c:153-185) */
    struct host_altcall *altcalls;
    /* [control] */
    char control_name[TSK_NAME_MAX];
    int control_first_cycle, control_read_length;
    /* [balancer] */
    int balancer_start, balancer_length, balancer_min_quality, balancer_minimum_occurrence;
    int balancer_npos, balancer_nneg;
    float balancer_min_fraction_passes;
    /* [polyA_seeder] */
    int seed_trigger, neg_polya_len, cctr_left, cctr_right, boundary_pos, sampling_gap, bins;
    float required_cdf_contrast;
    int fair_fp_len, fair_hash_size, fair_max_count;
    /* [polyA_finder] */
    int w_polyA[5], w_nonA[5];
    int min_polya_length, max_terminal_modifications, sigproc_trigger;
    /* [polyA_ruler] */
    float dark_cycles_threshold, t_intensity_k, t_intensity_center;
    int max_dark_cycles;
    /* samples: list order of the reference (Unknown, control, then LIFO) */
    struct host_sample *samples;
    int num_samples;
};

/* ini.c */
typedef int (*tsk_ini_handler)(void *user, const char *section, const char *name, const char *value);
int tsk_ini_parse(const char *filename, tsk_ini_handler handler, void *user);

/* config.c */
struct host_config *tsk_parse_config(const char *filename);
void tsk_free_config(struct host_config *cfg);
int tsk_build_params(const struct host_config *cfg, tsk_params *out);
char *tsk_replace_placeholder(const char *format, const char *old, const char *repl);

/* tileio.c */
int tsk_invert_matrix(const float m[16], float out[16]);
int tsk_load_color_matrix(float inv[16], const char *filename);
struct tile_files;
struct tile_files *tsk_open_tile(const struct host_config *cfg, const char *msgprefix);
uint32_t tsk_tile_nclusters(const struct tile_files *tf);
/* read the next `count` clusters of every cycle into bcl[T][count], cif[L3][4][count] */
int tsk_read_block(struct tile_files *tf, uint32_t count, uint8_t *bcl, int16_t *cif);
/* the same, with every cycle file's rows sent to the device as soon as they are read and .bcl.gz files
 * inflated there (rows of `bcl` that came from .bcl.gz stay unfilled on the host) */
int tsk_read_block_to_device(struct tile_files *tf, uint32_t count, uint8_t *bcl, int16_t *cif, tsk_ctx *ctx,
                             const float invmatrix[16]);
/* after the last block: -1 when a FASTQ override still holds records */
int tsk_check_tile_end(struct tile_files *tf);
void tsk_close_tile(struct tile_files *tf);

/* writers.c */
int tsk_open_writers(struct host_config *cfg);
void tsk_close_writers(struct host_config *cfg);
int tsk_write_signal_headers(struct host_config *cfg, uint32_t total_clusters);
int tsk_write_block_outputs(struct host_config *cfg, const tsk_call *calls, uint32_t nclusters,
                            uint32_t firstclusterno, const uint8_t *bcl, uint32_t **records,
                            const uint32_t *nslots);
/* device-encoded outputs (tsk_tile_encode): the files become series of BGZF blocks; the host adds
 * the signal header block and the end-of-file block and appends what the GPU hands back */
void tsk_writers_device_encoding(int on);
int tsk_write_encoded_outputs(struct host_config *cfg, tsk_ctx *ctx);
int tsk_build_encode_layout(const struct host_config *cfg, tsk_encode_layout *out);
int tsk_write_sigdists(const char *filename_format, const char *type, const uint32_t *counts,
                       int total_cycles, int bins);
int tsk_write_stats(const char *output, const struct host_config *cfg, const tsk_counters *counters);
/* one gzip member holding src[0..n) (*out is malloc'ed; level from TSK_GZ_LEVEL), and a few worker
 * threads over an array of independent tasks */
struct tsk_task { int (*fn)(void *); void *arg; };
int tsk_gz_member(const void *src, size_t n, unsigned char **out, size_t *outn);
int tsk_run_tasks(struct tsk_task *tasks, int ntasks, int nthreads);

#endif
