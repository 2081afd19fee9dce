/* TEST INFRASTRUCTURE ONLY -- see tsk_oracle.c. */
#ifndef TSK_ORACLE_H
#define TSK_ORACLE_H

#include "../include/tailseeker_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* CPU restatement of distribute_processing()/process_spots() for one block of clusters,
 * sequential in cluster order (== the reference with threads=1).
 *   calls      [nclusters]
 *   records    flat u32 buffer, capacity_words long; on return holds, sample after sample
 *              (list order), the records the reference writes to that sample's .sigpack
 *              (header + signal_dump_length packets each), in cluster order
 *   rec_offset [num_samples] word offset of each sample's first record
 *   rec_count  [num_samples]
 *   counters   [num_samples]   (ADDED to, like SampleInfo counters)
 *   pos, neg   u32[total_cycles][dist_sampling_bins]  (zeroed here, like prepare_split_jobs)
 *   scores_out optional float[nclusters][threep_length] (NaN where not computed), the raw
 *              poly(A) scores of compute_polya_score() for tolerance tests; may be NULL
 * returns 0, or -1 (capacity too small / bad parameters). */
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
                       float *scores_out);

/* inverse_4x4_matrix() restated (contrib/misc.c:11-58): float cofactor expansion. */
int tsk_oracle_invert_matrix(const float m[16], float out[16]);

/* load_score_cutoffs()'s conversion (polyaruler.c:93-96). */
uint32_t tsk_oracle_cutoff_to_packed(double cutoff);

/* process_polya_ruling()+measure_polya_length()+add_polya_score_sample()
 * (polyaruler.c:104-172,275-300) over `nrecords` records of (2 + dump_len) words.
 * polya_len_out[i] = measured length if >= minimum_polya_len else -1.
 * pos: u32[ncutoffs][bins], ADDED to. */
int tsk_oracle_ruler(const uint32_t *records, size_t nrecords, int dump_len,
                     const uint32_t *cutoffs, int ncutoffs,
                     int minimum_polya_len, float downhill_ext_weight,
                     int bins, float gap,
                     int16_t *polya_len_out, uint32_t *pos);

/* Best local alignment score of read[0..n) (characters ACGTN) against PhiX forward + 20 N +
 * reverse, the quantity ssw_align() reports as score1 in try_alignment_to_control()
 * (controlaligner.c:127-135). */
int tsk_oracle_control_score(const char *read, int n);

/* try_alignment_to_control() (controlaligner.c:107-145) on a whole read: exact substring of
 * seq[0..n) in either strand, else the score of seq[first_cycle..+n) reaches (int)(n * 0.65). */
int tsk_oracle_control_try(const char *seq, int first_cycle, int n);

/* host libm logf, exported so tests can pin the device logf against it */
float tsk_oracle_logf(float x);

#ifdef __cplusplus
}
#endif
#endif
