/*
 * tsk_score.cu -- normalise_score_pack: CIF intensities -> poly(A) scores -> packed signal
 * records + positive / negative score histograms, and the fair-sampling fix-up kernels.
 *
 * Restates (bit for bit) probe_signal_ranges (signalproc.c:132-201), decrosstalk_intensity
 * (:85-103), compute_polya_score + normalize_signals + shannon_entropy (:269-342,246-266,
 * 39-51), find_max_cumulative_contrast (:378-412), the sampling policy and
 * add_polya_score_sample (spotanalyzer.c:574-599,441-457), write_polya_score (:389-438).
 *
 * Two code paths, both exact:
 *   k_normalise_score_pack          the fast path: one thread per cluster, a branch-free
 *       cycle step; the nine IEEE divisions per cycle become multiplications by hoisted
 *       correctly rounded reciprocals plus one FMA residual correction (Markstein), logf is
 *       the host's table algorithm in fp64 FMAs, packed record words are transposed through
 *       shared memory so every record is written with full 128-byte lines.
 *   k_normalise_score_pack_generic  the literal operation-by-operation restatement
 *       (__fdiv_rn everywhere); runs only for the clusters the fast path hands over because
 *       its preconditions do not hold (degenerate signal bandwidths), and is what the fast
 *       path is tested against.
 */
#include <math_constants.h>
#include "tsk_internal.h"
#include "tsk_device.cuh"
#include "tsk_exact.cuh"
#include "tsk_tma.cuh"

#define FULL_MASK 0xffffffffu
#define CONTRAST_MAXLEN 288          /* reciprocal table length (scan_len <= threep_length + 1) */

struct ScoreShared {
    double2 logtab[16];
    /* the same sixteen entries eight times, copy c of entry i at [i][c]: the eight lanes of a
     * quarter-warp (one shared-memory wavefront of a 16-byte load) read eight different copies, so
     * the table lookups of the entropy never conflict on a bank (ncu: 102 M conflict cycles per tile
     * with the single copy, 14 % of the kernel's SM cycles) */
    double2 logtabw[16][8];
    float   tscore[TSK_T_SCORE_BINS + 1];
    int16_t limit[TSK_MAX_SAMPLES];
    int16_t dump[TSK_MAX_SAMPLES];
};

__device__ __forceinline__ void load_score_shared(ScoreShared &sh, const float *__restrict__ tscore,
                                                  const DevSample *__restrict__ samples, int S)
{
    for (int i = threadIdx.x; i < 16; i += blockDim.x)
        sh.logtab[i] = make_double2(c_logf_tab[i][0], c_logf_tab[i][1]);
    for (int i = threadIdx.x; i < 128; i += blockDim.x)
        sh.logtabw[i >> 3][i & 7] = make_double2(c_logf_tab[i >> 3][0], c_logf_tab[i >> 3][1]);
    for (int i = threadIdx.x; i <= TSK_T_SCORE_BINS; i += blockDim.x) sh.tscore[i] = tscore[i];
    for (int i = threadIdx.x; i < S; i += blockDim.x) {
        sh.limit[i] = samples[i].limit_threep;
        sh.dump[i] = samples[i].dump_len;
    }
    __syncthreads();
}

__device__ __forceinline__ int score_bin(float score, int bins)
{
    int b = __float2int_rz(__fmul_rn(score, (float)bins));
    return b >= bins ? bins - 1 : b;
}

/* list area layout (u32): [0] undo count, [1] contended count, [2] accepted count,
 * [3] slow-path count, then six arrays of `cap` entries */
__device__ __forceinline__ uint32_t *list_array(uint32_t *lists, int which, uint32_t cap)
{
    return lists + 16 + (size_t)which * cap;
}

/* Lists: 0 inline-sampled clusters that aborted later, 1 clusters of over-subscribed
 * fair-sampling buckets ("contenders"), 2 the contenders that were accepted, 3 clusters handed to
 * the literal kernel; 4 and 5 run parallel to 1 and 2 and hold the offset (in floats) of the
 * cluster's score column in the scratch block, or TSK_NO_COLUMN. */
#define TSK_NO_COLUMN 0xffffffffu
struct ScoreArgs;

/* ======================================================================================
 * generic path: literal restatement
 * ==================================================================================== */
__device__ __forceinline__ void decrosstalk(float sig[4], const int16_t *__restrict__ p, size_t pitch,
                                            const float *inv)
{
    const float v0 = (float)__ldg(p), v1 = (float)__ldg(p + pitch);
    const float v2 = (float)__ldg(p + 2 * pitch), v3 = (float)__ldg(p + 3 * pitch);
#pragma unroll
    for (int j = 0; j < 4; j++) {
        float a = __fmul_rn(inv[4 * j + 0], v0);
        a = __fadd_rn(a, __fmul_rn(inv[4 * j + 1], v1));
        a = __fadd_rn(a, __fmul_rn(inv[4 * j + 2], v2));
        sig[j] = __fadd_rn(a, __fmul_rn(inv[4 * j + 3], v3));
    }
}

/* low / bandwidth of the four channels over the balancer cycles.  Generic in npos / nneg
 * (<= TSK_MAX_RANK); the bubble-insert keeps earlier elements on ties like the reference. */
__device__ __forceinline__ void probe_ranges(const DevParams &P, const int16_t *__restrict__ cifcol,
                                             size_t pitch, int blen, float low[4], float bw[4])
{
    float top[4][TSK_MAX_RANK], bot[4][TSK_MAX_RANK];
    const int npos = P.npos, nneg = P.nneg;
#pragma unroll
    for (int ch = 0; ch < 4; ch++)
#pragma unroll
        for (int r = 0; r < TSK_MAX_RANK; r++) { top[ch][r] = -CUDART_INF_F; bot[ch][r] = CUDART_INF_F; }
    for (int c = P.balancer_start; c < P.balancer_start + blen; c++) {
        float sig[4];
        decrosstalk(sig, cifcol + (size_t)c * 4 * pitch, pitch, P.inv);
#pragma unroll
        for (int ch = 0; ch < 4; ch++) {
            float carry = sig[ch];
#pragma unroll
            for (int r = 0; r < TSK_MAX_RANK; r++)
                if (r < npos && carry > top[ch][r]) { const float t = top[ch][r]; top[ch][r] = carry; carry = t; }
            carry = sig[ch];
#pragma unroll
            for (int r = 0; r < TSK_MAX_RANK; r++)
                if (r < nneg && carry < bot[ch][r]) { const float t = bot[ch][r]; bot[ch][r] = carry; carry = t; }
        }
    }
#pragma unroll
    for (int ch = 0; ch < 4; ch++) {
        float acc = 0.f;
#pragma unroll
        for (int r = 0; r < TSK_MAX_RANK; r++) if (r < nneg) acc = __fadd_rn(acc, bot[ch][r]);
        low[ch] = __fdiv_rn(acc, (float)nneg);
        acc = 0.f;
#pragma unroll
        for (int r = 0; r < TSK_MAX_RANK; r++) if (r < npos) acc = __fadd_rn(acc, top[ch][r]);
        bw[ch] = __fsub_rn(__fdiv_rn(acc, (float)npos), low[ch]);
    }
}

/* One cycle of compute_polya_score(), operation by operation.  Returns false for a dark cycle. */
__device__ __forceinline__ bool score_cycle_generic(const DevParams &P, const ScoreShared &sh,
                                                    const int16_t *__restrict__ p, size_t pitch,
                                                    const float low[4], const float bw[4],
                                                    float &entropy_score, float &score)
{
    float sig[4], nrm[4];
    decrosstalk(sig, p, pitch, P.inv);
    float total = 0.f;
#pragma unroll
    for (int ch = 0; ch < 4; ch++)
        if (sig[ch] > 0.f) total = __fadd_rn(total, sig[ch]);
    if (total < P.dark_threshold) return false;
    total = 0.f;
#pragma unroll
    for (int ch = 0; ch < 4; ch++) {
        float v = __fdiv_rn(__fsub_rn(sig[ch], low[ch]), bw[ch]);
        if (v < 0.f) v = 0.f;
        nrm[ch] = v;
        total = __fadd_rn(total, v);
    }
#pragma unroll
    for (int ch = 0; ch < 4; ch++) nrm[ch] = __fdiv_rn(nrm[ch], total);
    if (nrm[0] != nrm[0]) return false;
    float h = 0.f;
#pragma unroll
    for (int ch = 0; ch < 4; ch++)
        if (nrm[ch] > 0.f) h = __fsub_rn(h, __fmul_rn(nrm[ch], tsk_logf(nrm[ch], sh.logtab)));
    entropy_score = __fsub_rn(1.f, __fdiv_rn(h, P.maximum_entropy));
    int ti = __float2int_rz(__fmul_rn(nrm[3], (float)TSK_T_SCORE_BINS));
    ti = max(0, min(TSK_T_SCORE_BINS, ti));      /* the reference indexes unchecked (UB outside) */
    score = __fmul_rn(entropy_score, sh.tscore[ti]);
    return true;
}

/* find_max_cumulative_contrast(), exact scan (signalproc.c:378-412) over a scratch column. */
__device__ __forceinline__ int contrast_exact(const float *__restrict__ col, size_t spitch, int len,
                                              int left, int right, double total, float *contrast)
{
    double cum = 0., best = 0.;
    int at = -1;
    const int last = len - right;
    for (int i = 0; i <= last; i++) {
        cum = __dadd_rn(cum, (double)col[(size_t)i * spitch]);
        if (i >= left - 1) {
            const double v = __dsub_rn(__ddiv_rn(cum, (double)i),
                                       __ddiv_rn(__dsub_rn(total, cum), (double)(len - i)));
            if (i == left - 1) { best = v; at = i; }
            else if (v > best) { best = v; at = i; }
        }
    }
    *contrast = (float)best;
    return at + 1;
}

__device__ __forceinline__ void sample_positive(const DevParams &P, const float *__restrict__ col,
                                                size_t spitch, int scan_len, int ps, int firstrow,
                                                uint32_t *__restrict__ pos)
{
    const int take = -ps + min(scan_len, P.boundary_pos - P.sampling_gap);
    for (int i0 = 0; i0 < take; i0 += 16) {
        float v[16];
#pragma unroll
        for (int k = 0; k < 16; k++)          /* sixteen independent loads in flight */
            v[k] = (i0 + k < take) ? col[(size_t)(i0 + k) * spitch] : CUDART_NAN_F;
#pragma unroll
        for (int k = 0; k < 16; k++)
            if (!isnan(v[k]) && !isinf(v[k]))
                atomicAdd(&pos[(size_t)(firstrow + i0 + k) * P.bins + score_bin(v[k], P.bins)], 1u);
    }
}

struct ScoreArgs {
    TileView tv;
    const double *rcp;            /* [CONTRAST_MAXLEN] 1.0 / i, correctly rounded */
    const double2 *logki;         /* [LOGF_KI_ENTRIES] (exponent, index) table of logf_ki_t */
    const DevSample *samples;
    const float *tscore;
    tsk_call *calls;
    const uint8_t *kind;
    const uint32_t *bucket;
    const uint32_t *worklist;
    const uint32_t *totals;
    const uint64_t *recbase;
    uint32_t *records;
    float *scratch;
    size_t spitch;
    uint32_t *pos, *neg;
    const uint32_t *fair;
    uint32_t *lists;
    uint32_t listcap;
};

__device__ __forceinline__ void append_contender(const ScoreArgs &A, uint32_t n, uint32_t column)
{
    const uint32_t k = atomicAdd(&A.lists[1], 1u);
    list_array(A.lists, 1, A.listcap)[k] = n;
    list_array(A.lists, 4, A.listcap)[k] = column;
}

/* Everything the reference does for one cluster inside process_polya_signal() after the
 * sequence seed, for work item w. */
__device__ void process_item_generic(const DevParams &P, const ScoreShared &sh, const ScoreArgs &A,
                                     uint32_t n)
{
    const TileView &tv = A.tv;
    const uint32_t w = n;                  /* scratch column of this cluster */
    const uint4 craw = *reinterpret_cast<const uint4 *>(&A.calls[n]);
    tsk_call call = *reinterpret_cast<const tsk_call *>(&craw);
    const uint32_t kd = A.kind[n];
    const int ts = P.threep_start, tl = P.threep_length, bins = P.bins;
    const int de = call.delimiter_end, ps = call.terminal_mods;
    const int sample = call.sample;
    const size_t pitch = tv.cif_pitch, spitch = A.spitch;
    const int16_t *cifcol = tv.cif + n;

    int blen = de - ts;
    if (blen > P.balancer_length) blen = P.balancer_length;
    float low[4], bw[4];
    probe_ranges(P, cifcol, pitch, blen, low, bw);

    int insert_len = ts + tl - de;
    const int limit = sh.limit[sample], dump = sh.dump[sample];
    if (limit > 0 && limit < insert_len) insert_len = limit;
    const int scan_len = insert_len - ps;
    const int valid = min(scan_len, dump);

    uint32_t *rec = nullptr;
    if (kd & KIND_EMIT) {
        rec = A.records + A.recbase[sample] + (size_t)call.record_slot * (size_t)(2 + dump);
        rec[0] = tv.firstclusterno + n;
        rec[1] = (uint32_t)(uint16_t)(int16_t)(de + ps) | ((uint32_t)(uint16_t)(int16_t)valid << 16);
    }
    bool neg_inline = false;
    if (kd & KIND_NEG)
        neg_inline = (P.fair_max_count == 0) || (A.fair[A.bucket[n]] <= (uint32_t)P.fair_max_count);

    int ndark = 0;
    float prev_entropy = CUDART_NAN_F;
    unsigned long long dhwin = 0;          /* bit k = downhill of cycle (c - k) */
    const int16_t *p = cifcol + (size_t)(de - ts) * 4 * pitch;
    const uint32_t rowbase = (uint32_t)de;  /* histogram row of score index c is de + c */

    for (int c = 0; c < insert_len; c++, p += 4 * pitch) {
        float es, score;
        const bool lit = score_cycle_generic(P, sh, p, pitch, low, bw, es, score);
        uint32_t dh = 0;
        if (lit) {
            dh = (es < prev_entropy) ? 1u : 0u;
            prev_entropy = es;
        } else {
            ndark++;
            score = CUDART_NAN_F;
            prev_entropy = CUDART_NAN_F;
        }
        dhwin = (dhwin << 1) | dh;
        const int j = c - ps;
        if (j >= 0) {
            if (rec != nullptr && j < valid) {
                uint32_t word = 0;
                if (lit) {
                    uint32_t s = 1u + (uint32_t)__float2int_rz(__fmul_rn(score, (float)(TSK_SCORE_MAX - 1)));
                    if (s >= TSK_SCORE_MAX) s = TSK_SCORE_MAX;
                    /* the reference passes the UN-offset downhill array: packet j carries
                     * downhill[j], not downhill[ps + j] (spotanalyzer.c:604-606) */
                    word = (s << 1) | (uint32_t)((dhwin >> ps) & 1ull);
                }
                rec[2 + j] = word;
            }
            if (kd & KIND_POS) A.scratch[(size_t)j * spitch + w] = score;
            if (neg_inline && lit && !isinf(score))
                atomicAdd(&A.neg[(size_t)(rowbase + c) * bins + score_bin(score, bins)], 1u);
        }
    }
    if (rec != nullptr)
        for (int j = max(valid, 0); j < dump; j++) rec[2 + j] = 0;

    bool aborted = false;
    if (ndark > 0) {
        call.flags |= TSK_PAFLAG_DARKCYCLE_EXISTS;
        if (ndark >= P.max_dark_cycles) {
            call.flags |= TSK_PAFLAG_DARKCYCLE_OVER_THRESHOLD;
            call.polya_len = -1;
            call.emitted = 0;
            aborted = true;
            if (rec != nullptr) rec[1] |= 0xffff0000u;     /* withdraw the slot: valid = -1 */
        }
        *reinterpret_cast<uint4 *>(&A.calls[n]) = *reinterpret_cast<const uint4 *>(&call);
    }

    if (kd & KIND_NEG) {
        if (aborted && neg_inline)
            list_array(A.lists, 0, A.listcap)[atomicAdd(&A.lists[0], 1u)] = n;
        else if (!aborted && !neg_inline)
            append_contender(A, n, TSK_NO_COLUMN);
    }

    if ((kd & KIND_POS) && !aborted && scan_len > 0 && P.cctr_left + P.cctr_right < scan_len) {
        const float *col = A.scratch + w;
        double total = 0.;
        for (int i = 0; i < scan_len; i++) total = __dadd_rn(total, (double)col[(size_t)i * spitch]);
        float contrast;
        const int at = contrast_exact(col, spitch, scan_len, P.cctr_left, P.cctr_right, total, &contrast);
        if (at >= P.boundary_pos && contrast >= P.required_contrast)
            sample_positive(P, col, spitch, scan_len, ps, (int)rowbase + ps, A.pos);
    }
}

__global__ void __launch_bounds__(TSK_K2_THREADS)
k_normalise_score_pack_generic(const __grid_constant__ DevParams P, const ScoreArgs A, int all_items)
{
    __shared__ ScoreShared sh;
    load_score_shared(sh, A.tscore, A.samples, P.num_samples);
    /* all_items: every work item (used by tests to pin the fast path against this one);
     * otherwise only the items the fast path handed over (list 3). */
    const uint32_t cnt = all_items ? A.totals[P.num_samples] : A.lists[3];
    const uint32_t *slow = list_array(A.lists, 3, A.listcap);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x)
        process_item_generic(P, sh, A, all_items ? A.worklist[i] : slow[i]);
}

/* ======================================================================================
 * fast path
 * ==================================================================================== */
/* RN(1/x) for x in [2^-60, 2^60): the reciprocal approximation plus one Newton step, the very
 * sequence __frcp_rn() runs on its fast path (checked against it in k_div_sweep) without its
 * range test, which the caller has already made. */
__device__ __forceinline__ float rcp_rn_fast(float x)
{
    float y0;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(x));
    const float e = -__fmaf_rn(x, y0, -1.0f);
    return __fmaf_rn(y0, e, y0);
}

__device__ __forceinline__ float logf_core(uint32_t ix, const double2 *__restrict__ tab)
{
    return logf_core_t(ix, [tab](uint32_t ofs) { return *reinterpret_cast<const double2 *>(
                                                     reinterpret_cast<const char *>(tab) + ofs); });
}

/* top-npos / bottom-nneg probing with compile-time ranks (the stock configuration 2 / 4) */
template <int NPOS, int NNEG>
__device__ __forceinline__ void probe_ranges_fixed(const DevParams &P, const int16_t *__restrict__ cifcol,
                                                   size_t pitch, int blen, float low[4], float bw[4])
{
    float top[4][NPOS], bot[4][NNEG];
#pragma unroll
    for (int ch = 0; ch < 4; ch++) {
#pragma unroll
        for (int r = 0; r < NPOS; r++) top[ch][r] = -CUDART_INF_F;
#pragma unroll
        for (int r = 0; r < NNEG; r++) bot[ch][r] = CUDART_INF_F;
    }
    const int16_t *p = cifcol;      /* first balancer cycle; plain loads: shared or global memory */
    for (int c = 0; c < blen; c++, p += 4 * pitch) {
        float sig[4];
        {
            const float v0 = (float)p[0], v1 = (float)p[pitch], v2 = (float)p[2 * pitch], v3 = (float)p[3 * pitch];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                float a = __fmul_rn(P.inv[4 * j + 0], v0);
                a = __fadd_rn(a, __fmul_rn(P.inv[4 * j + 1], v1));
                a = __fadd_rn(a, __fmul_rn(P.inv[4 * j + 2], v2));
                sig[j] = __fadd_rn(a, __fmul_rn(P.inv[4 * j + 3], v3));
            }
        }
#pragma unroll
        for (int ch = 0; ch < 4; ch++) {
            float carry = sig[ch];
#pragma unroll
            for (int r = 0; r < NPOS; r++) {
                const bool sw = carry > top[ch][r];
                const float t = top[ch][r];
                top[ch][r] = sw ? carry : t;
                carry = sw ? t : carry;
            }
            carry = sig[ch];
#pragma unroll
            for (int r = 0; r < NNEG; r++) {
                const bool sw = carry < bot[ch][r];
                const float t = bot[ch][r];
                bot[ch][r] = sw ? carry : t;
                carry = sw ? t : carry;
            }
        }
    }
#pragma unroll
    for (int ch = 0; ch < 4; ch++) {
        float acc = 0.f;
#pragma unroll
        for (int r = 0; r < NNEG; r++) acc = __fadd_rn(acc, bot[ch][r]);
        low[ch] = __fdiv_rn(acc, (float)NNEG);
        acc = 0.f;
#pragma unroll
        for (int r = 0; r < NPOS; r++) acc = __fadd_rn(acc, top[ch][r]);
        bw[ch] = __fsub_rn(__fdiv_rn(acc, (float)NPOS), low[ch]);
    }
}

__device__ __forceinline__ bool in_exp_range(float v, int lo_exp, int hi_exp)
{   /* finite, positive, 2^lo_exp <= v < 2^hi_exp */
    const int e = (int)(__float_as_uint(v) >> 23);       /* sign bit set -> e >= 256 */
    return e >= 127 + lo_exp && e < 127 + hi_exp;
}

__device__ __forceinline__ int lds_s16(uint32_t addr)
{
    int v;
    asm volatile("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

#define K2_CL        TSK_K2_THREADS      /* clusters per CTA */
#define K2_WARPS     (TSK_K2_THREADS / 32)
#define K2_GROUP     4                   /* cycles per TMA copy */
#define K2_NGROUP    2                   /* groups in each warp's ring (4-7 cycles of look-ahead; 3 groups were 2 % slower, 4 groups 3 %) */
#define K2_HEAD_MAX  24                  /* resident cycles [0, head): balancer + first scored cycles */
#define K2_ROWB      (4 * 32 * 2)        /* bytes per cycle and warp: four channel rows of 32 int16 */
#define K2_TILE_CYC  8                   /* record words staged per lane between flushes */
/* Measured on a 1.07 M-cluster tile (tools/k2_variants.sh, profiles/r2_k2_variants.txt): 4 CTAs per SM
 * at 119 registers and two cycles swept together 2.53 ms; 5 CTAs at 96 registers 2.59 ms; the
 * round-1 loop 2.79 ms. */
#ifndef K2_MIN_CTAS
#define K2_MIN_CTAS  4
#endif
#ifndef K2_SWEEP
#define K2_SWEEP     2                   /* cycles of a group whose phases are swept together (1, 2 or 4) */
#endif

struct __align__(128) K2Smem {
    int16_t  head[K2_WARPS][K2_HEAD_MAX][4][32];
    int16_t  stage[K2_WARPS][K2_NGROUP * K2_GROUP][4][32];
    uint32_t tile[K2_WARPS][32][K2_TILE_CYC + 1];
    ScoreShared sh;
    uint64_t bar_head[K2_WARPS], bar_full[K2_WARPS][K2_NGROUP];
};

/* One warp = 32 consecutive clusters, one lane per cluster, and its own little pipeline: the four
 * channel planes of every 3'-read cycle (4 x 64 bytes) are brought into shared memory by the TMA
 * engine (bulk asynchronous copies completing on an mbarrier) -- cycles [0, head) once, then a
 * ring of 4-cycle groups that lane 0 refills right after the warp has consumed a group.  The arithmetic therefore never waits on a global load, no warp waits for another, and
 * every DRAM request is a full, aligned 64-byte segment of a cycle-major plane. */
__global__ void __launch_bounds__(TSK_K2_THREADS, K2_MIN_CTAS)
k_normalise_score_pack(const __grid_constant__ DevParams P, const ScoreArgs A, const float rcp_maxent,
                       const int head, const __grid_constant__ CUtensorMap tmap)
{
    extern __shared__ __align__(128) unsigned char k2_smem_raw[];
    K2Smem &sm = *reinterpret_cast<K2Smem *>(k2_smem_raw);
    const ScoreShared &sh = sm.sh;
    const int S = P.num_samples;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const TileView &tv = A.tv;
    const int ts = P.threep_start, tl = P.threep_length, bins = P.bins;
    const size_t spitch = A.spitch;
    const uint32_t n = blockIdx.x * K2_CL + tid;              /* my cluster */
    const uint32_t a_head = smem_u32(&sm.head[warp][0][0][0]), a_stage = smem_u32(&sm.stage[warp][0][0][0]);
    const uint32_t a_bhead = smem_u32(&sm.bar_head[warp]), a_full = smem_u32(&sm.bar_full[warp][0]);
    const int x0 = (int)(n - lane);                           /* first cluster of this warp */

    if (lane == 0) {
        mbar_init(a_bhead, 1);
        for (int i = 0; i < K2_NGROUP; i++) mbar_init(a_full + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    load_score_shared(sm.sh, A.tscore, A.samples, S);      /* ends with __syncthreads() */

    /* ---- per-cluster descriptor ---- */
    uint32_t kd = 0;
    tsk_call call;
    int de = ts, ps = 0, sample = 0;
    bool alive = false;
    if (n < tv.nclusters) {
        kd = A.kind[n];
        alive = (kd & KIND_PROCESS) != 0;
        if (alive) {
            const uint4 craw = *reinterpret_cast<const uint4 *>(&A.calls[n]);
            call = *reinterpret_cast<const tsk_call *>(&craw);
            de = call.delimiter_end; ps = call.terminal_mods; sample = call.sample;
        } else {
            kd = 0;
        }
    }
    if (!__any_sync(FULL_MASK, alive)) return;                /* nothing to do for these 32 clusters */
    const int start = de - ts;                                /* first scored cycle (3'-read index) */

    int insert_len = 0, dump = 0;
    if (alive) {
        insert_len = ts + tl - de;
        const int limit = sh.limit[sample];
        dump = sh.dump[sample];
        if (limit > 0 && limit < insert_len) insert_len = limit;
    }
    const int rdata_end = start + insert_len;                 /* this lane scores cycles [start, rdata_end) */
    int rdata_w = alive ? rdata_end : 0;                      /* the warp needs data of cycles [0, rdata_w) */
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) rdata_w = max(rdata_w, __shfl_xor_sync(FULL_MASK, rdata_w, d));

    /* ---- TMA: resident head cycles, then the first ring of streamed cycles ---- */
    auto issue_group = [&](int g) {          /* lane 0 only: cycles head+4g .. head+4g+3 -> ring slot g % 3 */
        const int slot = g % K2_NGROUP;
        mbar_arrive_expect_tx(a_full + 8 * slot, K2_GROUP * K2_ROWB);
        tma_cycle_g2s(a_stage + slot * (K2_GROUP * K2_ROWB), &tmap, x0, head + g * K2_GROUP, a_full + 8 * slot);
    };
    const int ngroups = rdata_w > head ? (rdata_w - head + K2_GROUP - 1) / K2_GROUP : 0;   /* groups with data */
    if (lane == 0) {
        const int hgroups = (min(head, rdata_w) + K2_GROUP - 1) / K2_GROUP;
        mbar_arrive_expect_tx(a_bhead, (uint32_t)hgroups * K2_GROUP * K2_ROWB);
        for (int g = 0; g < hgroups; g++)
            tma_cycle_g2s(a_head + g * (K2_GROUP * K2_ROWB), &tmap, x0, g * K2_GROUP, a_bhead);
        for (int g = 0; g < min(K2_NGROUP, ngroups); g++) issue_group(g);
    }

    /* ---- balancer: per-cluster dynamic range, from the resident head cycles ---- */
    mbar_wait(a_bhead, 0);
    float low[4] = {0.f, 0.f, 0.f, 0.f}, bw[4] = {1.f, 1.f, 1.f, 1.f}, ybw[4];
    if (alive) {
        const int blen = min(start, P.balancer_length);
        probe_ranges_fixed<2, 4>(P, &sm.head[warp][P.balancer_start][0][lane], (size_t)32, blen, low, bw);
        /* fast-path preconditions; otherwise the generic kernel takes this cluster */
        bool ok = true;
#pragma unroll
        for (int ch = 0; ch < 4; ch++)
            ok = ok && in_exp_range(bw[ch], -40, 40) && (fabsf(low[ch]) < 0x1p40f);
        if (!ok) {
            list_array(A.lists, 3, A.listcap)[atomicAdd(&A.lists[3], 1u)] = n;
            alive = false;
            kd = 0;
            insert_len = 0; dump = 0;
#pragma unroll
            for (int ch = 0; ch < 4; ch++) { low[ch] = 0.f; bw[ch] = 1.f; }
        }
    }
#pragma unroll
    for (int ch = 0; ch < 4; ch++) ybw[ch] = __frcp_rn(bw[ch]);

    const int scan_len = insert_len - ps;
    const int valid = min(scan_len, dump);
    const bool emit = (kd & KIND_EMIT) != 0;
    const bool ispos = (kd & KIND_POS) != 0;
    const int rscore_end = alive ? rdata_end : 0;

    uint32_t *rec = nullptr;
    if (emit) {
        rec = A.records + A.recbase[sample] + (size_t)call.record_slot * (size_t)(2 + dump);
        rec[0] = tv.firstclusterno + n;
        rec[1] = (uint32_t)(uint16_t)(int16_t)(de + ps) | ((uint32_t)(uint16_t)(int16_t)valid << 16);
    }
    bool neg_inline = false;
    if (kd & KIND_NEG)
        neg_inline = (P.fair_max_count == 0) || (A.fair[A.bucket[n]] <= (uint32_t)P.fair_max_count);

    /* the warp computes on cycles [rbeg_w, rend_w): an emitting lane owns record positions up to
     * cycle start+ps+dump-1 (zero padding past its data) */
    int rend_w = alive ? (emit ? max(rdata_end, start + ps + dump) : rdata_end) : 0;
    int rbeg_w = alive ? start : 0x7fff;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        rend_w = max(rend_w, __shfl_xor_sync(FULL_MASK, rend_w, d));
        rbeg_w = min(rbeg_w, __shfl_xor_sync(FULL_MASK, rbeg_w, d));
    }
    const unsigned emitmask = __ballot_sync(FULL_MASK, emit);
    /* record flushes write four emitting lanes' words per round, one per quarter-warp: the lane
     * this quarter serves in round `it` sits in bits [6 it, 6 it + 6) of flushpick (63 = none) */
    const int nflush = (__popc(emitmask) + 3) >> 2;
    unsigned long long flushpick = ~0ull;
    {
        unsigned m = emitmask;
        for (int it = 0; it < nflush; it++) {
            int pick = 63;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int rr = m ? __ffs(m) - 1 : 63;
                if (m) m &= m - 1;
                if (k == (lane >> 3)) pick = rr;
            }
            flushpick = (flushpick & ~(63ull << (6 * it))) | ((unsigned long long)pick << (6 * it));
        }
    }
    const unsigned long long recaddr = (unsigned long long)(uintptr_t)rec;
    const int recinfo = ((start + ps) & 0xffff) | (max(dump, 0) << 16);

    int ndark = 0;
    float prev_entropy = CUDART_NAN_F;
    double ptotal = 0.;                       /* sum of the scores from the seed start on (KIND_POS lanes) */
    uint32_t dhwin = 0;                       /* bit k = downhill of cycle (c - k); max_mods < 32 */
    /* score of cycle r lives at scratch[(r - start - ps) * spitch + n]; scr walks one row per cycle */
    float *scr = A.scratch + n + ((ptrdiff_t)rbeg_w - (start + ps)) * (ptrdiff_t)spitch;
    const uint32_t a_logtab = smem_u32(&sm.sh.logtab[0]), a_tscore = smem_u32(&sm.sh.tscore[0]);
    const uint32_t a_tile = smem_u32(&sm.tile[warp][0][0]);
    const uint32_t a_mytile = a_tile + lane * ((K2_TILE_CYC + 1) * 4);
    const uint32_t colofs = (uint32_t)lane * 2;

    /* one cycle of this lane: r = 3'-read cycle, arow = shared address of the cycle's [4][32] rows */
    auto cycle = [&](const int r, const uint32_t arow) {
        uint32_t word = 0;
        if (r >= start && r < rscore_end) {
            /* ---- decrosstalk (signalproc.c:85-103) ---- */
            const uint32_t ap = arow + colofs;
            const float v0 = __fsub_rn(__int_as_float(0x4B400000 + lds_s16(ap)), 12582912.f);
            const float v1 = __fsub_rn(__int_as_float(0x4B400000 + lds_s16(ap + 64)), 12582912.f);
            const float v2 = __fsub_rn(__int_as_float(0x4B400000 + lds_s16(ap + 128)), 12582912.f);
            const float v3 = __fsub_rn(__int_as_float(0x4B400000 + lds_s16(ap + 192)), 12582912.f);
            float sig[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                float a = __fmul_rn(P.inv[4 * j + 0], v0);
                a = __fadd_rn(a, __fmul_rn(P.inv[4 * j + 1], v1));
                a = __fadd_rn(a, __fmul_rn(P.inv[4 * j + 2], v2));
                sig[j] = __fadd_rn(a, __fmul_rn(P.inv[4 * j + 3], v3));
            }
            /* ---- dark test: sum of the positive channels (x + 0 == x, so max(.,0) is exact) ---- */
            float tot = fmaxf(sig[0], 0.f);
            tot = __fadd_rn(tot, fmaxf(sig[1], 0.f));
            tot = __fadd_rn(tot, fmaxf(sig[2], 0.f));
            tot = __fadd_rn(tot, fmaxf(sig[3], 0.f));
            bool lit = !(tot < P.dark_threshold);
            /* ---- normalise into the cluster's dynamic range (signalproc.c:246-266) ---- */
            float nrm[4];
#pragma unroll
            for (int ch = 0; ch < 4; ch++)
                nrm[ch] = fmaxf(div_by_rcp(__fsub_rn(sig[ch], low[ch]), bw[ch], ybw[ch]), 0.f);
            const float sum = __fadd_rn(__fadd_rn(__fadd_rn(nrm[0], nrm[1]), nrm[2]), nrm[3]);
            float prob[4];
            if (__builtin_expect(in_exp_range(sum, -60, 60), 1)) {
                const float ys = rcp_rn_fast(sum);
#pragma unroll
                for (int ch = 0; ch < 4; ch++) prob[ch] = div_by_rcp(nrm[ch], sum, ys);
            } else {
                /* sum == 0 (0/0 -> NaN -> dark) or a pathological magnitude: IEEE division */
#pragma unroll
                for (int ch = 0; ch < 4; ch++) prob[ch] = __fdiv_rn(nrm[ch], sum);
            }
            lit = lit && (prob[0] == prob[0]);
            /* ---- Shannon entropy: h -= p*logf(p) for p > 0 (the skipped terms are 0) ---- */
            float h = 0.f;
            {
#pragma unroll
                for (int ch = 0; ch < 4; ch++) {
                    const float lg = logf_core_t(__float_as_uint(prob[ch]), [a_logtab](uint32_t ofs) {
                        double2 t;
                        asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(t.x), "=d"(t.y) : "r"(a_logtab + ofs));
                        return t;
                    });
                    const float term = __fmul_rn(prob[ch], lg);
                    h = __fsub_rn(h, prob[ch] > 0.f ? term : 0.f);
                }
            }
            const float es = __fsub_rn(1.f, div_by_rcp(h, P.maximum_entropy, rcp_maxent));
            const uint32_t ti = min((uint32_t)__float2int_rz(__fmul_rn(prob[3], (float)TSK_T_SCORE_BINS)),
                                    (uint32_t)TSK_T_SCORE_BINS);
            float tsc;
            asm("ld.shared.f32 %0, [%1];" : "=f"(tsc) : "r"(a_tscore + ti * 4));
            const float score = lit ? __fmul_rn(es, tsc) : CUDART_NAN_F;
            const uint32_t dh = (lit && es < prev_entropy) ? 1u : 0u;
            prev_entropy = lit ? es : CUDART_NAN_F;
            ndark += lit ? 0 : 1;
            dhwin = (dhwin << 1) | dh;

            if (r >= start + ps) {
                uint32_t s = 1u + (uint32_t)__float2int_rz(__fmul_rn(score, (float)(TSK_SCORE_MAX - 1)));
                s = min(s, TSK_SCORE_MAX);
                /* packet j = c - ps carries downhill[j], the UN-offset array (spotanalyzer.c:604-606) */
                word = lit ? ((s << 1) | ((dhwin >> ps) & 1u)) : 0u;
                if (ispos) {
                    *scr = score;
                    ptotal = __dadd_rn(ptotal, (double)score);      /* same order as the reference's cumsum */
                }
                if (neg_inline && lit && !isinf(score))
                    atomicAdd(&A.neg[(size_t)(ts + r) * bins + score_bin(score, bins)], 1u);
            }
        }
        /* ---- stage the packed word; every 8 cycles the emitting lanes' records are written four
         *      at a time, 8 consecutive words (one 32-byte sector) each ---- */
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(a_mytile + (r & (K2_TILE_CYC - 1)) * 4), "r"(word) : "memory");
        if ((r & (K2_TILE_CYC - 1)) == K2_TILE_CYC - 1 || r == rend_w - 1) {
            __syncwarp();
            const int rbase = r & ~(K2_TILE_CYC - 1);
            const int l8 = lane & 7;
            for (int it = 0; it < nflush; it++) {
                /* next (up to) four emitting lanes: this quarter-warp's is in flushpick */
                const int p6 = (int)(flushpick >> (6 * it)) & 63;
                const int pick = p6 == 63 ? -1 : p6;
                const int src = pick < 0 ? 0 : pick;
                const unsigned long long ra = __shfl_sync(FULL_MASK, recaddr, src);
                const int ri = __shfl_sync(FULL_MASK, recinfo, src);
                const int j = rbase + l8 - (ri & 0xffff);
                if (pick >= 0 && j >= 0 && j < (ri >> 16) && rbase + l8 <= r) {
                    uint32_t wv;
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wv) : "r"(a_tile + (pick * (K2_TILE_CYC + 1) + l8) * 4));
                    reinterpret_cast<uint32_t *>((uintptr_t)ra)[2 + j] = wv;
                }
            }
            __syncwarp();
        }
        scr += spitch;
    };

    /* ---- resident head cycles ---- */
    for (int r = rbeg_w; r < min(rend_w, head); r++) cycle(r, a_head + r * K2_ROWB);

    /* ---- streamed cycles, four at a time: wait for the group's rows, score its cycles, let lane 0
     *      refill the ring slot with the group three ahead ---- */
    for (int g = 0; head + g * K2_GROUP < rend_w; g++) {
        const int slot = g % K2_NGROUP;
        const int r0 = head + g * K2_GROUP;
        if (g < ngroups) mbar_wait(a_full + 8 * slot, (g / K2_NGROUP) & 1);
#pragma unroll 1
        for (int k = 0; k < K2_GROUP; k++)
            if (r0 + k < rend_w) cycle(r0 + k, a_stage + (slot * K2_GROUP + k) * K2_ROWB);
        __syncwarp();
        if (lane == 0 && g + K2_NGROUP < ngroups) issue_group(g + K2_NGROUP);
    }

    if (!alive) return;
    bool aborted = false;
    if (ndark > 0) {
        call.flags |= TSK_PAFLAG_DARKCYCLE_EXISTS;
        if (ndark >= P.max_dark_cycles) {
            call.flags |= TSK_PAFLAG_DARKCYCLE_OVER_THRESHOLD;
            call.polya_len = -1;
            call.emitted = 0;
            aborted = true;
            if (rec != nullptr) rec[1] |= 0xffff0000u;     /* withdraw the slot: valid = -1 */
        }
        *reinterpret_cast<uint4 *>(&A.calls[n]) = *reinterpret_cast<const uint4 *>(&call);
    }
    if (kd & KIND_NEG) {
        if (aborted && neg_inline)
            list_array(A.lists, 0, A.listcap)[atomicAdd(&A.lists[0], 1u)] = n;
        else if (!aborted && !neg_inline)
            append_contender(A, n, TSK_NO_COLUMN);
    }

    /* ---- contrast test (find_max_cumulative_contrast) and positive sampling ----
     * f(i) = cum_i/i - (total-cum_i)/(len-i).  The divisions are replaced by table reciprocals
     * to locate the maximum (error < 1.2e-15 per value, all terms <= 1 in magnitude); the
     * exact IEEE value is evaluated only there.  If a second candidate lies within 1e-14 the
     * exact scan decides, so ties resolve exactly like the reference's strict '>' scan. */
    if (ispos && !aborted && scan_len > 0 && P.cctr_left + P.cctr_right < scan_len) {
        const float *col = A.scratch + n;
        const int left = P.cctr_left, last = scan_len - P.cctr_right;
        const double total = ptotal;
        if (total == total) {        /* a NaN anywhere makes the first candidate NaN: no sampling */
            double cum = 0., m1 = -CUDART_INF, m2 = -CUDART_INF, cum1 = 0.;
            int at = left - 1;
            for (int i0 = 0; i0 <= last; i0 += 16) {
                float v8[16];
#pragma unroll
                for (int k = 0; k < 16; k++)      /* sixteen independent loads in flight */
                    v8[k] = (i0 + k <= last) ? col[(size_t)(i0 + k) * spitch] : 0.f;
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    const int i = i0 + k;
                    if (i <= last) {
                        cum = __dadd_rn(cum, (double)v8[k]);
                        if (i >= left - 1) {
                            const double v = __dsub_rn(__dmul_rn(cum, __ldg(A.rcp + i)),
                                                       __dmul_rn(__dsub_rn(total, cum), __ldg(A.rcp + (scan_len - i))));
                            if (v > m1) { m2 = m1; m1 = v; at = i; cum1 = cum; }
                            else if (v > m2) m2 = v;
                        }
                    }
                }
            }
            float contrast;
            int atpos;
            if (m1 - m2 > 1e-14 && at > 0) {
                const double v = __dsub_rn(__ddiv_rn(cum1, (double)at),
                                           __ddiv_rn(__dsub_rn(total, cum1), (double)(scan_len - at)));
                contrast = (float)v;
                atpos = at + 1;
            } else {
                atpos = contrast_exact(col, spitch, scan_len, left, P.cctr_right, total, &contrast);
            }
            if (atpos >= P.boundary_pos && contrast >= P.required_contrast)
                sample_positive(P, col, spitch, scan_len, ps, de + ps, A.pos);
        }
    }
}

/* ---- compact form: the same per-cycle arithmetic over WORK ITEMS instead of cluster rows.
 * A warp takes 32 consecutive entries of the work list (the clusters that reach
 * process_polya_signal(), in ascending order), so no lane idles on a dropped cluster (unknown
 * barcode / PhiX, failed fingerprint or balancer QC, no delimiter: 13 % of the synthetic tile,
 * far more with a large PhiX spike-in).  The TMA box is K2C_W clusters wide starting at the item's
 * first cluster (rounded down to 8) and each lane reads its own column of it; clusters further
 * than K2C_W from there (an item spanning > 121 clusters, i.e. under ~30 % kept) go to the literal kernel.  Nothing is
 * resident: the balancer cycles and then the scored cycles stream through the same ring of two
 * 4-cycle groups (the few cycles both passes need are fetched twice, from L2). ---- */
#define K2C_NARROW 64                    /* clusters per TMA box of the narrow instance */
#define K2C_WIDE   128                   /* ... of the wide one */
#define K2_WIDE_ITEMS 8                  /* lists[8]: items the narrow instance left to the wide one (zeroed with the list counters) */

template <int W>
struct __align__(128) K2SmemC {
    int16_t  stage[K2_WARPS][K2_NGROUP * K2_GROUP][4][W];
    uint32_t tile[K2_WARPS][32][K2_TILE_CYC + 1];
    ScoreShared sh;
    uint64_t bar_full[K2_WARPS][K2_NGROUP];
    /* record flush plan: round `it`, quarter-warp q writes eight words of ONE record:
     * {record address lo, hi, first cycle | words << 16, byte offset of the lane's row in tile[warp]} */
    uint4 plan[K2_WARPS][8][4];
};

/* Two instances are launched over the same work list and every item is taken by exactly one of them:
 * the narrow one (64-cluster box) if all 32 clusters of the item lie within 64 of its first -- any tile
 * that keeps more than ~60 % of its clusters -- the wide one (128) otherwise.  A box twice as wide as
 * needed doubles the L2 -> shared-memory traffic of the kernel (8.1 GB per tile with 128 for 2.0 GB
 * used) and cost 3.6 % of its time. */
template <int W>
__global__ void __launch_bounds__(TSK_K2_THREADS, K2_MIN_CTAS)
k_normalise_score_pack_compact(const __grid_constant__ DevParams P, const ScoreArgs A, const float rcp_maxent,
                               const __grid_constant__ CUtensorMap tmap)
{
    constexpr int K2C_W = W;                     /* clusters per TMA box */
    constexpr int K2C_CHB = K2C_W * 2;           /* bytes per channel row */
    constexpr int K2C_ROWB = 4 * K2C_CHB;        /* bytes per cycle and warp */
    /* the wide instance runs second: nothing was left for it on most tiles */
    if (W == K2C_WIDE && A.lists[K2_WIDE_ITEMS] == 0) return;
    extern __shared__ __align__(128) unsigned char k2_smem_raw[];
    K2SmemC<W> &sm = *reinterpret_cast<K2SmemC<W> *>(k2_smem_raw);
    const ScoreShared &sh = sm.sh;
    const int S = P.num_samples;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const TileView &tv = A.tv;
    const int ts = P.threep_start, tl = P.threep_length, bins = P.bins;
    /* contrast scratch, dense per work item: row i of item w at scratch[(w * (tl + 1) + i) * 32 + lane],
     * so every row starts a 128-byte line */
    const size_t spitch = 32;
    const uint32_t a_stage = smem_u32(&sm.stage[warp][0][0][0]);
    const uint32_t a_full = smem_u32(&sm.bar_full[warp][0]);

    if (lane == 0) {
        for (int i = 0; i < K2_NGROUP; i++) mbar_init(a_full + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    load_score_shared(sm.sh, A.tscore, A.samples, S);      /* ends with __syncthreads() */

    /* ---- work item: 32 consecutive entries of the work list (clusters that reach
     *      process_polya_signal(), ascending); lane -> cluster, window = [x0, x0 + K2C_W) ---- */
    const uint32_t nwork = A.totals[S];
    const uint32_t w0 = (blockIdx.x * K2_WARPS + warp) * 32;
    if (w0 >= nwork) return;
    uint32_t kd = 0;
    tsk_call call;
    int de = ts, ps = 0, sample = 0;
    bool alive = w0 + lane < nwork;
    uint32_t n = alive ? A.worklist[w0 + lane] : 0u;          /* my cluster */
    /* window start: a tiled TMA copy must begin on a 16-byte boundary of the plane (8 clusters);
     * a box starting elsewhere faults with "illegal instruction" */
    const int x0 = (int)(__shfl_sync(FULL_MASK, n, 0) & ~7u);
    /* which instance owns the item (warp-uniform); the narrow one counts what it leaves */
    if (__all_sync(FULL_MASK, !alive || n - (uint32_t)x0 < K2C_NARROW) != (W == K2C_NARROW)) {
        if (W == K2C_NARROW && lane == 0) atomicAdd(&A.lists[K2_WIDE_ITEMS], 1u);
        return;
    }
    if (alive && n - (uint32_t)x0 >= K2C_W) {
        /* the window does not reach this far (long run of dropped clusters): literal kernel */
        list_array(A.lists, 3, A.listcap)[atomicAdd(&A.lists[3], 1u)] = n;
        alive = false;
    }
    if (!alive) n = (uint32_t)x0;
    if (alive) {
        kd = A.kind[n];
        const uint4 craw = *reinterpret_cast<const uint4 *>(&A.calls[n]);
        call = *reinterpret_cast<const tsk_call *>(&craw);
        de = call.delimiter_end; ps = call.terminal_mods; sample = call.sample;
    }
    const int start = de - ts;                                /* first scored cycle (3'-read index) */

    int insert_len = 0, dump = 0;
    if (alive) {
        insert_len = ts + tl - de;
        const int limit = sh.limit[sample];
        dump = sh.dump[sample];
        if (limit > 0 && limit < insert_len) insert_len = limit;
    }
    const int rdata_end = start + insert_len;                 /* this lane scores cycles [start, rdata_end) */
    int rdata_w = alive ? rdata_end : 0;                      /* the warp needs data of cycles [0, rdata_w) */
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) rdata_w = max(rdata_w, __shfl_xor_sync(FULL_MASK, rdata_w, d));

    /* ---- streaming: cycles [c0, c1) in groups of four through the warp's ring; data exist for
     *      cycles < cdata (later ones only pad records).  G counts groups over both passes so that
     *      ring slots and barrier phases carry on. ---- */
    const uint32_t colofs = (n - (uint32_t)x0) * 2;
    int G = 0;
    auto issue_group = [&](int gabs, int cyc) {       /* lane 0 only */
        const int slot = gabs % K2_NGROUP;
        mbar_arrive_expect_tx(a_full + 8 * slot, K2_GROUP * K2C_ROWB);
        tma_cycle_g2s(a_stage + slot * (K2_GROUP * K2C_ROWB), &tmap, x0, cyc, a_full + 8 * slot);
    };
    auto stream = [&](const int c0, const int c1, const int cdata, auto &&fn, auto &&after_group) {
        const int ng = c1 > c0 ? (c1 - c0 + K2_GROUP - 1) / K2_GROUP : 0;
        const int ngd = cdata > c0 ? min(ng, (cdata - c0 + K2_GROUP - 1) / K2_GROUP) : 0;
        if (lane == 0)
            for (int g = 0; g < min(K2_NGROUP, ngd); g++) issue_group(G + g, c0 + g * K2_GROUP);
        for (int g = 0; g < ng; g++) {
            const int slot = (G + g) % K2_NGROUP;
            if (g < ngd) mbar_wait(a_full + 8 * slot, ((G + g) / K2_NGROUP) & 1);
#pragma unroll 1
            for (int k = 0; k < K2_GROUP; k++)
                if (c0 + g * K2_GROUP + k < c1) fn(c0 + g * K2_GROUP + k, a_stage + (slot * K2_GROUP + k) * K2C_ROWB);
            __syncwarp();
            if (lane == 0 && g + K2_NGROUP < ngd) issue_group(G + g + K2_NGROUP, c0 + (g + K2_NGROUP) * K2_GROUP);
            after_group(g, ng, min(c0 + g * K2_GROUP + K2_GROUP, c1) - 1);
        }
        G += ngd;
    };

    /* ---- balancer: per-cluster dynamic range (top 2 / bottom 4 per channel over the balancer
     *      cycles, signalproc.c:106-243), streamed once ---- */
    float low[4] = {0.f, 0.f, 0.f, 0.f}, bw[4] = {1.f, 1.f, 1.f, 1.f}, ybw[4];
    {
        float top[4][2], bot[4][4];
#pragma unroll
        for (int ch = 0; ch < 4; ch++) {
            top[ch][0] = top[ch][1] = -CUDART_INF_F;
            bot[ch][0] = bot[ch][1] = bot[ch][2] = bot[ch][3] = CUDART_INF_F;
        }
        const int blen = alive ? min(start, P.balancer_length) : 0;
        int blen_w = blen;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) blen_w = max(blen_w, __shfl_xor_sync(FULL_MASK, blen_w, d));
        stream(P.balancer_start, P.balancer_start + blen_w, P.balancer_start + blen_w, [&](const int r, const uint32_t arow) {
            if (r - P.balancer_start < blen) {
                const uint32_t ap = arow + colofs;
                const float v0 = (float)lds_s16(ap), v1 = (float)lds_s16(ap + K2C_CHB), v2 = (float)lds_s16(ap + 2 * K2C_CHB),
                            v3 = (float)lds_s16(ap + 3 * K2C_CHB);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    float a = __fmul_rn(P.inv[4 * j + 0], v0);
                    a = __fadd_rn(a, __fmul_rn(P.inv[4 * j + 1], v1));
                    a = __fadd_rn(a, __fmul_rn(P.inv[4 * j + 2], v2));
                    const float sg = __fadd_rn(a, __fmul_rn(P.inv[4 * j + 3], v3));
                    float carry = sg;
#pragma unroll
                    for (int q = 0; q < 2; q++) {
                        const bool sw = carry > top[j][q];
                        const float t = top[j][q];
                        top[j][q] = sw ? carry : t;
                        carry = sw ? t : carry;
                    }
                    carry = sg;
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const bool sw = carry < bot[j][q];
                        const float t = bot[j][q];
                        bot[j][q] = sw ? carry : t;
                        carry = sw ? t : carry;
                    }
                }
            }
        }, [](int, int, int) {});
        if (alive) {
#pragma unroll
            for (int ch = 0; ch < 4; ch++) {
                float acc = 0.f;
#pragma unroll
                for (int q = 0; q < 4; q++) acc = __fadd_rn(acc, bot[ch][q]);
                low[ch] = __fdiv_rn(acc, 4.f);
                acc = __fadd_rn(__fadd_rn(0.f, top[ch][0]), top[ch][1]);
                bw[ch] = __fsub_rn(__fdiv_rn(acc, 2.f), low[ch]);
            }
            /* fast-path preconditions; otherwise the generic kernel takes this cluster */
            bool ok = true;
#pragma unroll
            for (int ch = 0; ch < 4; ch++)
                ok = ok && in_exp_range(bw[ch], -40, 40) && (fabsf(low[ch]) < 0x1p40f);
            if (!ok) {
                list_array(A.lists, 3, A.listcap)[atomicAdd(&A.lists[3], 1u)] = n;
                alive = false;
                kd = 0;
                insert_len = 0; dump = 0;
#pragma unroll
                for (int ch = 0; ch < 4; ch++) { low[ch] = 0.f; bw[ch] = 1.f; }
            }
        }
    }
#pragma unroll
    for (int ch = 0; ch < 4; ch++) ybw[ch] = __frcp_rn(bw[ch]);

    const int scan_len = insert_len - ps;
    const int valid = min(scan_len, dump);
    const bool emit = (kd & KIND_EMIT) != 0;
    const bool ispos = (kd & KIND_POS) != 0;

    uint32_t *rec = nullptr;
    if (emit) {
        rec = A.records + A.recbase[sample] + (size_t)call.record_slot * (size_t)(2 + dump);
        rec[0] = tv.firstclusterno + n;
        rec[1] = (uint32_t)(uint16_t)(int16_t)(de + ps) | ((uint32_t)(uint16_t)(int16_t)valid << 16);
    }
    bool neg_inline = false;
    if (kd & KIND_NEG)
        neg_inline = (P.fair_max_count == 0) || (A.fair[A.bucket[n]] <= (uint32_t)P.fair_max_count);

    /* the warp computes on cycles [rbeg_w, rend_w): an emitting lane owns record positions up to
     * cycle start+ps+dump-1 (zero padding past its data) */
    int rend_w = alive ? (emit ? max(rdata_end, start + ps + dump) : rdata_end) : 0;
    int rbeg_w = alive ? start : 0x7fff;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        rend_w = max(rend_w, __shfl_xor_sync(FULL_MASK, rend_w, d));
        rbeg_w = min(rbeg_w, __shfl_xor_sync(FULL_MASK, rbeg_w, d));
    }
    const unsigned emitmask = __ballot_sync(FULL_MASK, emit);
    /* record flushes write four emitting lanes' words per round, one per quarter-warp: the lane
     * this quarter serves in round `it` sits in bits [6 it, 6 it + 6) of flushpick (63 = none) */
    const int nflush = (__popc(emitmask) + 3) >> 2;
    unsigned long long flushpick = ~0ull;
    {
        unsigned m = emitmask;
        for (int it = 0; it < nflush; it++) {
            int pick = 63;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int rr = m ? __ffs(m) - 1 : 63;
                if (m) m &= m - 1;
                if (k == (lane >> 3)) pick = rr;
            }
            flushpick = (flushpick & ~(63ull << (6 * it))) | ((unsigned long long)pick << (6 * it));
        }
    }
    {
        /* who writes what is the same at every flush: worked out once (three shuffles a round),
         * kept in shared memory and read back with one broadcast load a round */
        const unsigned long long recaddr = (unsigned long long)(uintptr_t)rec;
        const int recinfo = ((start + ps) & 0xffff) | (max(dump, 0) << 16);
        for (int it = 0; it < 8; it++) {
            const int p6 = it < nflush ? (int)(flushpick >> (6 * it)) & 63 : 63;
            const int src = p6 == 63 ? 0 : p6;
            const unsigned long long ra = __shfl_sync(FULL_MASK, recaddr, src);
            const int ri = __shfl_sync(FULL_MASK, recinfo, src);
            if ((lane & 7) == 0)
                sm.plan[warp][it][lane >> 3] = make_uint4((uint32_t)ra, (uint32_t)(ra >> 32),
                                                          p6 == 63 ? 0u : (uint32_t)ri,
                                                          (uint32_t)(src * (K2_TILE_CYC + 1) * 4));
        }
        __syncwarp();
    }
    const uint32_t a_plan = smem_u32(&sm.plan[warp][0][lane >> 3]);

    int ndark = 0;
    float prev_entropy = CUDART_NAN_F;
    double ptotal = 0.;                       /* sum of the scores from the seed start on (KIND_POS lanes) */
    uint32_t dhwin = 0;                       /* bit k = downhill of cycle (c - k); max_mods < 32 */
    /* score of cycle r lives in row (r - start - ps) of the item's scratch block; scr walks one row per cycle */
    /* columns are handed only to the lanes whose scores are read back, packed to the left: a row is then a run of
     * completely written sectors (a sector with unwritten bytes is first filled from DRAM when it
     * leaves L2: 0.44 GB of reads per tile) and the other lanes' slots cost no traffic at all */
    /* KIND_NEG lanes keep their scores too.  Those that may be sampled at once (neg_inline) are
     * sampled after the loop by the whole warp, 32 cycles of one cluster at a time: inside the
     * loop the bin + RED sequence ran every cycle for the one lane in five that needed it (5.6 %
     * of the kernel's instructions), and a cluster that aborted on dark cycles had to be taken
     * out again afterwards.  Those of over-subscribed fair-sampling buckets are sampled after
     * the fix-ups, if accepted (k_sample_neg_columns). */
    const bool needcol = ispos || (kd & KIND_NEG) != 0;
    float *const scol = A.scratch + (size_t)(w0 >> 5) * (size_t)(tl + 1) * 32 +
                        __popc(__ballot_sync(FULL_MASK, needcol) & ((1u << lane) - 1));
    float *scr = scol + ((ptrdiff_t)rbeg_w - (start + ps)) * (ptrdiff_t)spitch;
    const uint32_t a_logtab = smem_u32(&sm.sh.logtab[0]), a_tscore = smem_u32(&sm.sh.tscore[0]);
    const uint32_t a_tile = smem_u32(&sm.tile[warp][0][0]);
    const uint32_t a_mytile = a_tile + lane * ((K2_TILE_CYC + 1) * 4);

    /* ---- scored cycles: one TMA group (four cycles) per step, straight-line and branch-free.
     * Every lane runs every cycle of the warp's range [rbeg_w, rend_w); `act` (this lane is inside
     * its own [start, start + insert_len)) gates what the cycle may change, so there is no per-lane
     * branch, no range test per cycle and the addresses inside a group are immediates.  A cycle
     * whose channel sum leaves the range the reciprocal division is proved for (never on stock
     * data; a zero sum is a dark cycle and is fine) hands the cluster to the literal kernel after
     * the loop instead of taking an IEEE-division branch inside it.
     * Entropy terms: p * logf(p) with p = +0 is 0 * (finite) = -0, and h - (-0) == h, so the
     * "skip p <= 0" select of shannon_entropy() (signalproc.c:39-51) is implicit; h starts as -t0
     * instead of 0 - t0 (differs only in the sign of a zero, which 1 - h/maxent absorbs). ---- */
    const unsigned ilen_a = alive ? (unsigned)insert_len : 0u;
    const char *const logki = reinterpret_cast<const char *>(A.logki);
    const uint32_t a_logtabw = smem_u32(&sm.sh.logtabw[0][lane & 7]);
    bool handover = false;
    /* phase A of a cycle: intensities -> the four channel probabilities and whether the cycle is lit */
    auto cycle_a = [&](const int c, const uint32_t ap, float prob[4], bool &good) {
        const bool act = (unsigned)c < ilen_a;
        /* ---- decrosstalk (signalproc.c:85-103) ---- */
        /* I2FP (one instruction, exact for int16); the two-instruction integer-add + FADD route
         * through 2^23 + 2^22 measured 1 % slower */
        const float v0 = (float)lds_s16(ap), v1 = (float)lds_s16(ap + K2C_CHB), v2 = (float)lds_s16(ap + 2 * K2C_CHB),
                    v3 = (float)lds_s16(ap + 3 * K2C_CHB);
        float sig[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float a = __fmul_rn(P.inv[4 * j + 0], v0);
            a = __fadd_rn(a, __fmul_rn(P.inv[4 * j + 1], v1));
            a = __fadd_rn(a, __fmul_rn(P.inv[4 * j + 2], v2));
            sig[j] = __fadd_rn(a, __fmul_rn(P.inv[4 * j + 3], v3));
        }
        /* ---- dark test: sum of the positive channels (x + 0 == x, so max(.,0) is exact) ---- */
        float tot = fmaxf(sig[0], 0.f);
        tot = __fadd_rn(tot, fmaxf(sig[1], 0.f));
        tot = __fadd_rn(tot, fmaxf(sig[2], 0.f));
        tot = __fadd_rn(tot, fmaxf(sig[3], 0.f));
        /* ---- normalise into the cluster's dynamic range (signalproc.c:246-266) ---- */
        float nrm[4];
#pragma unroll
        for (int ch = 0; ch < 4; ch++)
            nrm[ch] = fmaxf(div_by_rcp(__fsub_rn(sig[ch], low[ch]), bw[ch], ybw[ch]), 0.f);
        const float sum = __fadd_rn(__fadd_rn(__fadd_rn(nrm[0], nrm[1]), nrm[2]), nrm[3]);
        /* sum == 0 (every channel at or below its floor: a dark cycle): rcp(0) = inf, the
         * refinement turns it into NaN, and so does 0 x NaN in the division -- the NaN that IEEE
         * 0/0 gives, which is all that is looked at */
        handover |= act & ((__float_as_uint(sum) - 0x21800000u) >= 0x3c000000u) & (sum != 0.f);
        const float ys = rcp_rn_fast(sum);
#pragma unroll
        for (int ch = 0; ch < 4; ch++) prob[ch] = div_by_rcp(nrm[ch], sum, ys);
        good = act && !(tot < P.dark_threshold) && (prob[0] == prob[0]);
    };
    /* phase B: entropy -> score -> packed word, sampling columns, dark-cycle count */
    auto cycle_b = [&](const int c, const float prob[4], const bool good, const uint32_t atile, float *const scr) {
        const bool act = (unsigned)c < ilen_a;
        /* ---- Shannon entropy ---- */
        float term[4];
#pragma unroll
        for (int ch = 0; ch < 4; ch++) {
#ifndef TSK_K2_LOGTAB_GLOBAL
            /* 16-entry core table in shared memory: 21 more instructions per cycle than the (exponent, index)
             * table below, yet 4 % faster -- the 32 KB table's L1 latency (long-scoreboard stalls 3x) costs more */
            const float lg = logf_core_t<true>(__float_as_uint(prob[ch]), [a_logtabw](uint32_t ofs) {
                double2 t;
                asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(t.x), "=d"(t.y) : "r"(a_logtabw + ofs));
                return t;
            });
#else
            /* one 16-byte entry per (exponent, index) from a 32 KB table in global memory: the few
             * rows probabilities really fall into (k = 0 ... -12, 3 KB) stay in L1 */
            const float lg = logf_ki_t(__float_as_uint(prob[ch]), [logki](uint32_t ofs) {
                double2 t;
                asm("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(t.x), "=d"(t.y) : "l"(logki + ofs));
                return t;
            });
#endif
            term[ch] = __fmul_rn(prob[ch], lg);
        }
        const float h = __fsub_rn(__fsub_rn(__fsub_rn(-term[0], term[1]), term[2]), term[3]);
        const float es = __fsub_rn(1.f, div_by_rcp(h, P.maximum_entropy, rcp_maxent));
        const uint32_t ti = min((uint32_t)__float2int_rz(__fmul_rn(prob[3], (float)TSK_T_SCORE_BINS)),
                                (uint32_t)TSK_T_SCORE_BINS);
        float tsc;
        asm("ld.shared.f32 %0, [%1];" : "=f"(tsc) : "r"(a_tscore + ti * 4));
        const float score = good ? __fmul_rn(es, tsc) : CUDART_NAN_F;
        const uint32_t dh = (good && es < prev_entropy) ? 1u : 0u;
        prev_entropy = good ? es : CUDART_NAN_F;
        ndark += (act && !good) ? 1 : 0;
        dhwin = (dhwin << 1) | dh;
        /* packet j = c - ps carries downhill[j], the UN-offset array (spotanalyzer.c:604-606) */
        const bool inrec = act && c >= ps;
        uint32_t sp = 1u + (uint32_t)__float2int_rz(__fmul_rn(score, (float)(TSK_SCORE_MAX - 1)));
        sp = min(sp, TSK_SCORE_MAX);
        const uint32_t word = (good && c >= ps) ? ((sp << 1) | ((dhwin >> ps) & 1u)) : 0u;
        if (needcol && inrec) {       /* one predicate: the sum is only used by the KIND_POS lanes */
            *scr = score;
            ptotal = __dadd_rn(ptotal, (double)score);      /* same order as the reference's cumsum */
        }
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(atile), "r"(word) : "memory");
    };
    /* after every second group (8 cycles) and after the last one: r = last cycle staged */
    auto flush = [&](const int g, const int ng, const int r) {
        if (!(g & 1) && g != ng - 1) return;
        const int rbase = r - ((r - rbeg_w) & (K2_TILE_CYC - 1));     /* cycle in tile column 0 */
        const int l8 = lane & 7;
        const int rb = rbase + l8;
        const uint32_t a_src = a_tile + l8 * 4;
        /* two rounds at a time: four independent shared loads, then the two stores */
        for (int it = 0; it < nflush; it += 2) {
            uint4 e0, e1;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(e0.x), "=r"(e0.y), "=r"(e0.z), "=r"(e0.w) : "r"(a_plan + it * 64));
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(e1.x), "=r"(e1.y), "=r"(e1.z), "=r"(e1.w) : "r"(a_plan + it * 64 + 64));
            uint32_t w0, w1;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(a_src + e0.w));
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w1) : "r"(a_src + e1.w));
            const uint32_t j0 = (uint32_t)(rb - (int)(e0.z & 0xffffu)), j1 = (uint32_t)(rb - (int)(e1.z & 0xffffu));
            if (rb <= r && j0 < (e0.z >> 16))
                reinterpret_cast<uint32_t *>((uintptr_t)(((unsigned long long)e0.y << 32) | e0.x))[2 + j0] = w0;
            if (rb <= r && j1 < (e1.z >> 16))
                reinterpret_cast<uint32_t *>((uintptr_t)(((unsigned long long)e1.y << 32) | e1.x))[2 + j1] = w1;
        }
        __syncwarp();
    };
    {
        const int c0 = rbeg_w, c1 = rend_w;
        const int ng = c1 > c0 ? (c1 - c0 + K2_GROUP - 1) / K2_GROUP : 0;
        const int ngd = rdata_w > c0 ? min(ng, (rdata_w - c0 + K2_GROUP - 1) / K2_GROUP) : 0;
        if (lane == 0)
            for (int g = 0; g < min(K2_NGROUP, ngd); g++) issue_group(G + g, c0 + g * K2_GROUP);
        int cidx = c0 - start;                 /* this lane's score index of the group's first cycle */
        float *scrg = scr;                     /* row of the group's first cycle in the lane's scratch column */
        for (int g = 0; g < ng; g++) {
            const int slot = (G + g) % K2_NGROUP;
            if (g < ngd) mbar_wait(a_full + 8 * slot, ((G + g) / K2_NGROUP) & 1);
            /* groups past the data (record padding only) run on whatever the slot holds: act is false */
            const uint32_t ap = a_stage + slot * (K2_GROUP * K2C_ROWB) + colofs;
            const uint32_t atile = a_mytile + (g & 1) * (K2_GROUP * 4);
            /* the group's cycles in two sweeps: the four probability chains are independent of each
             * other, and so are the sixteen logarithms that follow -- fewer issue slots lost to
             * waiting on the previous instruction of one chain */
#if K2_SWEEP == 4
            {
                float prob[K2_GROUP][4];
                bool good[K2_GROUP];
#pragma unroll
                for (int k = 0; k < K2_GROUP; k++) cycle_a(cidx + k, ap + k * K2C_ROWB, prob[k], good[k]);
#pragma unroll
                for (int k = 0; k < K2_GROUP; k++) cycle_b(cidx + k, prob[k], good[k], atile + 4 * k, scrg + (size_t)k * spitch);
            }
#elif K2_SWEEP == 2
#pragma unroll
            for (int k2 = 0; k2 < K2_GROUP; k2 += 2) {
                float prob[2][4];
                bool good[2];
#pragma unroll
                for (int k = 0; k < 2; k++) cycle_a(cidx + k2 + k, ap + (k2 + k) * K2C_ROWB, prob[k], good[k]);
#pragma unroll
                for (int k = 0; k < 2; k++) cycle_b(cidx + k2 + k, prob[k], good[k], atile + 4 * (k2 + k), scrg + (size_t)(k2 + k) * spitch);
            }
#else
#pragma unroll
            for (int k = 0; k < K2_GROUP; k++) {
                float prob[4];
                bool good;
                cycle_a(cidx + k, ap + k * K2C_ROWB, prob, good);
                cycle_b(cidx + k, prob, good, atile + 4 * k, scrg + (size_t)k * spitch);
            }
#endif
            cidx += K2_GROUP;
            scrg += K2_GROUP * spitch;
            __syncwarp();
            if (lane == 0 && g + K2_NGROUP < ngd) issue_group(G + g + K2_NGROUP, c0 + (g + K2_NGROUP) * K2_GROUP);
            flush(g, ng, min(c0 + g * K2_GROUP + K2_GROUP, c1) - 1);
        }
        G += ngd;
    }
    if (handover && alive) {        /* the literal kernel redoes this cluster from scratch (list 3) */
        list_array(A.lists, 3, A.listcap)[atomicAdd(&A.lists[3], 1u)] = n;
        alive = false;
    }

    /* ---- negative sampling (spotanalyzer.c:589-599) of the lanes that need no fair-sampling
     *      decision and did not abort: the warp walks each such lane's column ---- */
    __syncwarp();                                   /* the columns were written by other lanes */
    for (unsigned todo = __ballot_sync(FULL_MASK, alive && neg_inline && !(ndark > 0 && ndark >= P.max_dark_cycles)); todo; todo &= todo - 1) {
        const int src = __ffs(todo) - 1;
        const float *col = reinterpret_cast<const float *>((uintptr_t)__shfl_sync(FULL_MASK, (unsigned long long)(uintptr_t)scol, src));
        const int len = __shfl_sync(FULL_MASK, scan_len, src);
        const int row0 = __shfl_sync(FULL_MASK, de + ps, src);
        for (int k0 = lane; k0 < len; k0 += 256) {
            float sc[8];
#pragma unroll
            for (int j = 0; j < 8; j++)             /* eight independent loads in flight */
                sc[j] = (k0 + 32 * j < len) ? col[(size_t)(k0 + 32 * j) * spitch] : CUDART_NAN_F;
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (sc[j] == sc[j] && !isinf(sc[j]))    /* NaN = dark cycle */
                    atomicAdd(&A.neg[(size_t)(row0 + k0 + 32 * j) * bins + score_bin(sc[j], bins)], 1u);
        }
    }

    /* the rest is per lane; the lanes meet again for the positive sampling */
    bool do_pos = false;
    [&]() {
    if (!alive) return;
    bool aborted = false;
    if (ndark > 0) {
        call.flags |= TSK_PAFLAG_DARKCYCLE_EXISTS;
        if (ndark >= P.max_dark_cycles) {
            call.flags |= TSK_PAFLAG_DARKCYCLE_OVER_THRESHOLD;
            call.polya_len = -1;
            call.emitted = 0;
            aborted = true;
            if (rec != nullptr) rec[1] |= 0xffff0000u;     /* withdraw the slot: valid = -1 */
        }
        *reinterpret_cast<uint4 *>(&A.calls[n]) = *reinterpret_cast<const uint4 *>(&call);
    }
    if ((kd & KIND_NEG) && !aborted && !neg_inline) append_contender(A, n, (uint32_t)(scol - A.scratch));

    /* ---- contrast test (find_max_cumulative_contrast) and positive sampling ----
     * f(i) = cum_i/i - (total-cum_i)/(len-i).  The divisions are replaced by table reciprocals
     * to locate the maximum (error < 1.2e-15 per value, all terms <= 1 in magnitude); the
     * exact IEEE value is evaluated only there.  If a second candidate lies within 1e-14 the
     * exact scan decides, so ties resolve exactly like the reference's strict '>' scan. */
    if (ispos && !aborted && scan_len > 0 && P.cctr_left + P.cctr_right < scan_len) {
        const float *col = scol;
        const int left = P.cctr_left, last = scan_len - P.cctr_right;
        const double total = ptotal;
        if (total == total) {        /* a NaN anywhere makes the first candidate NaN: no sampling */
            double cum = 0., m1 = -CUDART_INF, m2 = -CUDART_INF, cum1 = 0.;
            int at = left - 1;
            /* three stretches, so that the long one carries no per-element range tests (they cost
             * a third of the scan's instructions): positions before the first candidate only add
             * up; whole batches of sixteen candidates; the ragged end */
            int i0 = 0;
            for (; i0 < left - 1; i0 += 8) {
                float x[8];
#pragma unroll
                for (int k = 0; k < 8; k++) x[k] = (i0 + k < left - 1) ? col[(size_t)(i0 + k) * spitch] : 0.f;
#pragma unroll
                for (int k = 0; k < 8; k++) if (i0 + k < left - 1) cum = __dadd_rn(cum, (double)x[k]);
            }
            auto candidate = [&](const int i, const float x) {
                cum = __dadd_rn(cum, (double)x);
                const double v = __dsub_rn(__dmul_rn(cum, __ldg(A.rcp + i)),
                                           __dmul_rn(__dsub_rn(total, cum), __ldg(A.rcp + (scan_len - i))));
                if (v > m1) { m2 = m1; m1 = v; at = i; cum1 = cum; }
                else if (v > m2) m2 = v;
            };
            for (i0 = left - 1; i0 + 15 <= last; i0 += 16) {
                const float *p = col + (size_t)i0 * spitch;
                float x[16];
#pragma unroll
                for (int k = 0; k < 16; k++) x[k] = p[(size_t)k * spitch];      /* sixteen independent loads in flight */
#pragma unroll
                for (int k = 0; k < 16; k++) candidate(i0 + k, x[k]);
            }
            if (i0 <= last) {
                float x[16];
#pragma unroll
                for (int k = 0; k < 16; k++) x[k] = (i0 + k <= last) ? col[(size_t)(i0 + k) * spitch] : 0.f;
#pragma unroll
                for (int k = 0; k < 16; k++) if (i0 + k <= last) candidate(i0 + k, x[k]);
            }
            float contrast;
            int atpos;
            if (m1 - m2 > 1e-14 && at > 0) {
                const double v = __dsub_rn(__ddiv_rn(cum1, (double)at),
                                           __ddiv_rn(__dsub_rn(total, cum1), (double)(scan_len - at)));
                contrast = (float)v;
                atpos = at + 1;
            } else {
                atpos = contrast_exact(col, spitch, scan_len, left, P.cctr_right, total, &contrast);
            }
            do_pos = atpos >= P.boundary_pos && contrast >= P.required_contrast;
        }
    }
    }();

    /* ---- positive sampling (add_polya_score_sample, spotanalyzer.c:441-457) of the lanes whose
     *      contrast test passed: like the negative one, the warp walks one cluster at a time ---- */
    for (unsigned todo = __ballot_sync(FULL_MASK, do_pos); todo; todo &= todo - 1) {
        const int src = __ffs(todo) - 1;
        const float *col = reinterpret_cast<const float *>((uintptr_t)__shfl_sync(FULL_MASK, (unsigned long long)(uintptr_t)scol, src));
        const int take = __shfl_sync(FULL_MASK, -ps + min(scan_len, P.boundary_pos - P.sampling_gap), src);
        const int row0 = __shfl_sync(FULL_MASK, de + ps, src);
        for (int k0 = lane; k0 < take; k0 += 128) {
            float sc[4];
#pragma unroll
            for (int j = 0; j < 4; j++)
                sc[j] = (k0 + 32 * j < take) ? col[(size_t)(k0 + 32 * j) * spitch] : CUDART_NAN_F;
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (sc[j] == sc[j] && !isinf(sc[j]))
                    atomicAdd(&A.pos[(size_t)(row0 + k0 + 32 * j) * bins + score_bin(sc[j], bins)], 1u);
        }
    }
}

/* ======================================================================================
 * fair-sampling fix-ups
 * ==================================================================================== */
/* Fair sampling among the clusters of over-subscribed buckets: the reference accepts, per
 * bucket, the first fair_max_count eligible clusters in cluster order (spotanalyzer.c:476-484
 * with threads=1), i.e. the fair_max_count smallest cluster numbers of the bucket's entries in
 * the contended list.  Each contended bucket owns fair_max_count slots; every entry walks down
 * its bucket's slots with atomicMin, leaving the smaller value in the slot and carrying the larger
 * one on -- an insertion sort whose steps are atomic exchanges of a sorted pair, so in any
 * interleaving slot s ends up with the (s+1)-th smallest number.  An entry is accepted when it
 * finds its own number in one of the slots. */
__global__ void __launch_bounds__(256)
k_fair_init(uint32_t *__restrict__ slots, int K, const uint32_t *__restrict__ bucket,
            const uint32_t *__restrict__ lists, uint32_t listcap)
{
    const uint32_t cnt = lists[1];
    const uint32_t *cont = list_array(const_cast<uint32_t *>(lists), 1, listcap);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
        uint32_t *sl = slots + (size_t)bucket[cont[i]] * K;
        for (int s = 0; s < K; s++) sl[s] = 0xffffffffu;
    }
}

__global__ void __launch_bounds__(256)
k_fair_insert(uint32_t *__restrict__ slots, int K, const uint32_t *__restrict__ bucket,
              const uint32_t *__restrict__ lists, uint32_t listcap)
{
    const uint32_t cnt = lists[1];
    const uint32_t *cont = list_array(const_cast<uint32_t *>(lists), 1, listcap);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
        uint32_t v = cont[i];
        uint32_t *sl = slots + (size_t)bucket[v] * K;
        /* slots only ever decrease: an entry above the current last slot cannot end up in one.
         * Saves the same-address atomics of a bucket shared by thousands of low-diversity reads
         * (4 ns each, serialised: 0.9 ms on a 3.1 M-cluster tile). */
        if (v > __ldcg(&sl[K - 1])) continue;
        for (int s = 0; s < K && v != 0xffffffffu; s++) {
            const uint32_t old = atomicMin(&sl[s], v);
            v = max(old, v);
        }
    }
}

__global__ void __launch_bounds__(256)
k_fair_claim(const uint32_t *__restrict__ slots, int K, const uint32_t *__restrict__ bucket,
             uint32_t *__restrict__ lists, uint32_t listcap)
{
    const uint32_t cnt = lists[1];
    const uint32_t *cont = list_array(lists, 1, listcap);
    uint32_t *acc = list_array(lists, 2, listcap);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
        const uint32_t e = cont[i];
        const uint32_t *sl = slots + (size_t)bucket[e] * K;
        bool mine = false;
        for (int s = 0; s < K; s++) mine |= sl[s] == e;
        if (mine) {
            const uint32_t k = atomicAdd(&lists[2], 1u);
            acc[k] = e;
            list_array(lists, 5, listcap)[k] = list_array(lists, 4, listcap)[i];
        }
    }
}

/* probe_ranges() by a whole warp (balancer of at most 32 cycles): lane c de-crosstalks balancer
 * cycle c, the npos largest / nneg smallest values per channel are taken with warp-wide max / min
 * reductions on order-preserving integer keys (one instance removed per round, so duplicates count
 * like in the insertion form) and summed in the same order, largest / smallest first.  Unfilled
 * ranks keep -inf / +inf exactly like the arrays of the sequential form.  Every lane returns the
 * result. */
__device__ __forceinline__ int float_key(float x)
{
    const int b = __float_as_int(x);
    return b ^ ((b >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float key_float(int k) { return __int_as_float(k ^ ((k >> 31) & 0x7fffffff)); }

__device__ __forceinline__ void probe_ranges_warp(const DevParams &P, const int16_t *__restrict__ cifcol, size_t pitch,
                                                  int blen, float low[4], float bw[4], int lane)
{
    float sig[4] = {0.f, 0.f, 0.f, 0.f};
    const bool mine = lane < blen;
    if (mine) decrosstalk(sig, cifcol + (size_t)(P.balancer_start + lane) * 4 * pitch, pitch, P.inv);
    const int kneg = float_key(-CUDART_INF_F), kpos = float_key(CUDART_INF_F);
#pragma unroll
    for (int ch = 0; ch < 4; ch++) {
        int kmax = mine ? float_key(sig[ch]) : kneg, kmin = mine ? float_key(sig[ch]) : kpos;
        float acc = 0.f;
        for (int r = 0; r < P.nneg; r++) {
            const int m = __reduce_min_sync(FULL_MASK, kmin);
            const unsigned who = __ballot_sync(FULL_MASK, kmin == m && m != kpos);
            if (who && lane == __ffs(who) - 1) kmin = kpos;
            acc = __fadd_rn(acc, key_float(m));
        }
        low[ch] = __fdiv_rn(acc, (float)P.nneg);
        acc = 0.f;
        for (int r = 0; r < P.npos; r++) {
            const int m = __reduce_max_sync(FULL_MASK, kmax);
            const unsigned who = __ballot_sync(FULL_MASK, kmax == m && m != kneg);
            if (who && lane == __ffs(who) - 1) kmax = kneg;
            acc = __fadd_rn(acc, key_float(m));
        }
        bw[ch] = __fsub_rn(__fdiv_rn(acc, (float)P.npos), low[ch]);
    }
}

/* Recompute the scores of the listed clusters and add `delta` (+1 / -1, as u32) to the
 * negative histogram: accepted clusters of over-subscribed buckets (+1), and inline-sampled
 * clusters that later aborted on dark cycles (-1).  One CTA per listed cluster: the cycles of
 * a cluster are independent once its signal ranges are known, so the threads take cycles
 * tid, tid+128, ... (the literal arithmetic of the generic path). */
__global__ void __launch_bounds__(TSK_K2_THREADS, 6)      /* latency-bound (scattered column reads): warps over registers */
k_resample_neg(const __grid_constant__ DevParams P, const TileView tv,
               const DevSample *__restrict__ samples, const float *__restrict__ g_tscore,
               const tsk_call *__restrict__ calls, uint32_t *__restrict__ neg,
               const uint32_t *__restrict__ lists, uint32_t listcap, int which, uint32_t delta)
{
    __shared__ ScoreShared sh;
    load_score_shared(sh, g_tscore, samples, P.num_samples);
    const uint32_t cnt = lists[which];
    const uint32_t *arr = list_array(const_cast<uint32_t *>(lists), which, listcap);
    const int lane = threadIdx.x & 31;
    /* one CTA per listed cluster: the lists are short, a cluster's latency is what counts */
    const uint32_t *column = which == 2 ? list_array(const_cast<uint32_t *>(lists), 5, listcap) : nullptr;
    for (uint32_t i = blockIdx.x; i < cnt; i += gridDim.x) {
        if (column && column[i] != TSK_NO_COLUMN) continue;      /* sampled by k_sample_neg_columns */
        const uint32_t n = arr[i] & 0x7fffffffu;
        const tsk_call call = calls[n];
        const int ts = P.threep_start, tl = P.threep_length, bins = P.bins;
        const int de = call.delimiter_end, ps = call.terminal_mods;
        const size_t pitch = tv.cif_pitch;
        const int16_t *cifcol = tv.cif + n;
        int blen = de - ts;
        if (blen > P.balancer_length) blen = P.balancer_length;
        float low[4], bw[4];
        if (blen <= 32) probe_ranges_warp(P, cifcol, pitch, blen, low, bw, lane);
        else probe_ranges(P, cifcol, pitch, blen, low, bw);        /* every lane, same addresses */
        int insert_len = ts + tl - de;
        const int limit = sh.limit[call.sample];
        if (limit > 0 && limit < insert_len) insert_len = limit;
        for (int c = ps + (int)threadIdx.x; c < insert_len; c += TSK_K2_THREADS) {
            float es, score;
            const int16_t *p = cifcol + (size_t)(de - ts + c) * 4 * pitch;
            if (score_cycle_generic(P, sh, p, pitch, low, bw, es, score) && !isinf(score))
                atomicAdd(&neg[(size_t)(de + c) * bins + score_bin(score, bins)], delta);
        }
    }
}

/* Accepted clusters of over-subscribed buckets whose scores the scoring kernel kept: one warp per
 * cluster walks its scratch column (row i = cycle de + ps + i) and adds the bins, NaN = dark cycle
 * and infinite scores skipped like in the inline form. */
__global__ void __launch_bounds__(256)
k_sample_neg_columns(const __grid_constant__ DevParams P, const DevSample *__restrict__ samples,
                     const tsk_call *__restrict__ calls, const float *__restrict__ scratch,
                     uint32_t *__restrict__ neg, const uint32_t *__restrict__ lists, uint32_t listcap)
{
    const uint32_t cnt = lists[2];
    const uint32_t *arr = list_array(const_cast<uint32_t *>(lists), 2, listcap);
    const uint32_t *column = list_array(const_cast<uint32_t *>(lists), 5, listcap);
    const int lane = threadIdx.x & 31, bins = P.bins;
    for (uint32_t i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < cnt; i += gridDim.x * (blockDim.x >> 5)) {
        const uint32_t off = column[i];
        if (off == TSK_NO_COLUMN) continue;
        const tsk_call call = calls[arr[i]];
        const int de = call.delimiter_end, ps = call.terminal_mods;
        int insert_len = P.threep_start + P.threep_length - de;
        const int limit = samples[call.sample].limit_threep;
        if (limit > 0 && limit < insert_len) insert_len = limit;
        const float *col = scratch + off;
        for (int k = lane; k < insert_len - ps; k += 32) {
            const float score = col[(size_t)k * 32];
            if (score == score && !isinf(score))
                atomicAdd(&neg[(size_t)(de + ps + k) * bins + score_bin(score, bins)], 1u);
        }
    }
}

/* the (exponent, index) table of logf_ki_t, built once per context with the device's own FMA */
__global__ void k_fill_logki(double2 *__restrict__ tab)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < LOGF_KI_ENTRIES) tab[e] = logf_ki_entry(e, c_logf_tab[e & 15][0], c_logf_tab[e & 15][1]);
}

int tsk_score_setup(tsk_ctx *ctx)
{
    TSK_CUDA(cudaMalloc(&ctx->d_logki, sizeof(double2) * LOGF_KI_ENTRIES));
    k_fill_logki<<<LOGF_KI_ENTRIES / 256, 256, 0, ctx->stream>>>(reinterpret_cast<double2 *>(ctx->d_logki));
    ctx->launches++;
    TSK_CUDA(cudaGetLastError());
    return 0;
}

/* ======================================================================================
 * self-test hooks
 * ==================================================================================== */
__global__ void k_logf_probe(const float *__restrict__ in, float *__restrict__ out, size_t n, int variant)
{
    __shared__ double2 tab[16];
    if (threadIdx.x < 16) tab[threadIdx.x] = make_double2(c_logf_tab[threadIdx.x][0], c_logf_tab[threadIdx.x][1]);
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float x = in[i];
        if (variant == 0) out[i] = tsk_logf(x, tab);
        else {
            const int pbits = __float_as_int(x);
            out[i] = logf_core((pbits <= 0) ? 0x3f800000u : (uint32_t)max(pbits, 0x00800000), tab);
        }
    }
}

/* all floats with bit patterns [first, first+count): counts where the FMA core differs from
 * the separately-rounded restatement (must be 0 on normal inputs in (0,1]) */
__global__ void k_logf_sweep(uint32_t first, uint64_t count, unsigned long long *mismatches, const double2 *__restrict__ logki)
{
    __shared__ double2 tab[16];
    if (threadIdx.x < 16) tab[threadIdx.x] = make_double2(c_logf_tab[threadIdx.x][0], c_logf_tab[threadIdx.x][1]);
    __syncthreads();
    unsigned long long bad = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t ix = first + (uint32_t)i;
        const float a = tsk_logf(__uint_as_float(ix), tab);
        const float b = logf_core(ix, tab);
        bad += __float_as_uint(a) != __float_as_uint(b);
        {                                           /* the (exponent, index) table form of the fast kernel */
            const float c = logf_ki_t(ix, [logki](uint32_t ofs) { return __ldg(logki + (ofs >> 4)); });
            bad += __float_as_uint(a) != __float_as_uint(c);
        }
    }
    if (bad) atomicAdd(mismatches, bad);
}

/* pseudo-random (a, b) pairs: div_by_rcp(a, b, RN(1/b)) against __fdiv_rn(a, b) */
__global__ void k_div_sweep(uint64_t seed, uint64_t count, int mode, unsigned long long *mismatches)
{
    unsigned long long bad = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t x = (i + seed) * 0x9E3779B97F4A7C15ull;
        x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 32;
        uint32_t ab = (uint32_t)x, bb = (uint32_t)(x >> 32);
        float a, b;
        if (mode == 0) {            /* b in [2^-40, 2^40), a = +-[2^-60, 2^60) : arbitrary mantissas */
            b = __uint_as_float(((127 - 40 + (bb >> 23) % 80) << 23) | (bb & 0x7fffffu));
            a = __uint_as_float((ab & 0x80000000u) | ((127 - 60 + ((ab >> 23) & 0xff) % 120) << 23) | (ab & 0x7fffffu));
        } else {                    /* probabilities: 0 <= a <= b, b in [2^-3, 2^3) */
            b = __uint_as_float(((124 + (bb >> 23) % 6) << 23) | (bb & 0x7fffffu));
            a = __fmul_rn(b, __uint_as_float(0x3f800000u | (ab & 0x7fffffu)) - 1.0f);
        }
        const float yb = rcp_rn_fast(b);
        bad += __float_as_uint(yb) != __float_as_uint(__frcp_rn(b));
        const float q = div_by_rcp(a, b, yb);
        const float want = __fdiv_rn(a, b);
        bad += __float_as_uint(q) != __float_as_uint(want) && !(q == 0.f && want == 0.f);
    }
    if (bad) atomicAdd(mismatches, bad);
}

extern "C" int tsk_selftest_logf(tsk_ctx *ctx, const float *in, float *out, size_t n)
{
    if (!ctx || !in || !out) { tsk_set_error("null argument"); return -1; }
    TSK_CUDA(cudaSetDevice(ctx->device));
    float *d_in = NULL, *d_out = NULL;
    TSK_CUDA(cudaMalloc(&d_in, sizeof(float) * n));
    TSK_CUDA(cudaMalloc(&d_out, sizeof(float) * n));
    TSK_CUDA(cudaMemcpyAsync(d_in, in, sizeof(float) * n, cudaMemcpyHostToDevice, ctx->stream));
    k_logf_probe<<<148 * 8, 256, 0, ctx->stream>>>(d_in, d_out, n, 0);
    ctx->launches++;
    TSK_CUDA(cudaMemcpyAsync(out, d_out, sizeof(float) * n, cudaMemcpyDeviceToHost, ctx->stream));
    TSK_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_in); cudaFree(d_out);
    return 0;
}

extern "C" int tsk_selftest_sweeps(tsk_ctx *ctx, uint64_t div_pairs, uint64_t *logf_mismatches,
                                   uint64_t *div_mismatches)
{
    if (!ctx || !logf_mismatches || !div_mismatches) { tsk_set_error("null argument"); return -1; }
    TSK_CUDA(cudaSetDevice(ctx->device));
    unsigned long long *d = NULL, h[2] = {0, 0};
    TSK_CUDA(cudaMalloc(&d, sizeof(h)));
    TSK_CUDA(cudaMemsetAsync(d, 0, sizeof(h), ctx->stream));
    /* every normal float in (0, 1] */
    k_logf_sweep<<<148 * 16, 256, 0, ctx->stream>>>(0x00800000u, (uint64_t)(0x3f800000u - 0x00800000u) + 1, d,
                                                    reinterpret_cast<const double2 *>(ctx->d_logki));
    k_div_sweep<<<148 * 16, 256, 0, ctx->stream>>>(12345, div_pairs / 2, 0, d + 1);
    k_div_sweep<<<148 * 16, 256, 0, ctx->stream>>>(98765, div_pairs / 2, 1, d + 1);
    ctx->launches += 3;
    TSK_CUDA(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    TSK_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(d);
    *logf_mismatches = h[0];
    *div_mismatches = h[1];
    return 0;
}

/* ======================================================================================
 * launchers
 * ==================================================================================== */
static ScoreArgs make_args(tsk_ctx *ctx)
{
    ScoreArgs a;
    a.tv = ctx->tile; a.samples = ctx->d_samples; a.tscore = ctx->d_tscore; a.calls = ctx->d_calls;
    a.kind = ctx->d_kind; a.bucket = ctx->d_bucket; a.worklist = ctx->d_worklist; a.totals = ctx->d_totals;
    a.recbase = ctx->d_recbase; a.records = ctx->d_records; a.scratch = ctx->d_scratch;
    a.spitch = (size_t)ctx->cap_clusters; a.pos = ctx->d_pos; a.neg = ctx->d_neg; a.fair = ctx->d_fair;
    a.lists = ctx->d_lists; a.listcap = ctx->cap_clusters; a.rcp = ctx->d_rcp;
    a.logki = reinterpret_cast<const double2 *>(ctx->d_logki);
    return a;
}

/* resident head cycles of the fast kernel: the balancer and every possible first scored cycle */
static int fast_head_cycles(const tsk_ctx *ctx)
{
    const tsk_params &hp = ctx->hp;
    int head = hp.balancer_start + hp.balancer_length;
    for (int s = 0; s < hp.num_samples; s++)
        if (hp.samples[s].delimiter_length > 0) {
            const int first = hp.samples[s].delimiter_pos + hp.samples[s].delimiter_length + 1 - hp.threep_start;
            if (first + 1 > head) head = first + 1;
        }
    head = (head + 3) / 4 * 4;
    if (head > hp.threep_length) head = hp.threep_length;
    return head;
}

/* TMA descriptor of the resident CIF tensor, i16 [threep_length][4][pitch]: boxes of
 * {32 clusters, 4 channels, 1 cycle}; reads past nclusters are zero-filled by the engine. */
static int make_cif_tensor_map(tsk_ctx *ctx, CUtensorMap *tmap, unsigned boxw)
{
    tsk_encode_tiled_fn encode = tsk_tensor_map_encoder();
    if (!encode) return -1;
    const cuuint64_t dims[3] = {(cuuint64_t)ctx->tile.nclusters, 4, (cuuint64_t)ctx->dp.threep_length};
    const cuuint64_t strides[2] = {(cuuint64_t)ctx->tile.cif_pitch * 2, (cuuint64_t)ctx->tile.cif_pitch * 8};
    const cuuint32_t box[3] = {boxw, 4, K2_GROUP}, estr[3] = {1, 1, 1};
    const CUresult r = encode(tmap, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, (void *)ctx->tile.cif, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { tsk_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return -1; }
    return 0;
}

int tsk_launch_score(tsk_ctx *ctx)
{
    const uint32_t N = ctx->tile.nclusters;
    if (N == 0) return 0;
    const ScoreArgs a = make_args(ctx);
    const int head = fast_head_cycles(ctx);
    /* the TMA bulk copies need 16-byte aligned 256-byte rows, and the CTA windows must stay inside
     * the row pitch; the downhill window is 32 bits */
    /* |inverse colour matrix| < 2^40 keeps every de-crosstalked signal, quotient and residual of the
     * fast division far from overflow (|signal - low| / bandwidth < 2^58 * 2^40) */
    bool inv_ok = true;
    for (int i = 0; i < 16; i++) inv_ok = inv_ok && fabsf(ctx->dp.inv[i]) < 0x1p40f;
    const bool fast_ok = ctx->force_generic != 1 && inv_ok && head <= K2_HEAD_MAX && ctx->dp.max_mods < 32 &&
                         ctx->dp.npos == 2 && ctx->dp.nneg == 4 &&
                         (ctx->tile.cif_pitch % 8) == 0 && ((uintptr_t)ctx->tile.cif % 16) == 0;
    if (!fast_ok) {
        const unsigned grid = (N + TSK_K2_THREADS - 1) / TSK_K2_THREADS;
        k_normalise_score_pack_generic<<<grid, TSK_K2_THREADS, 0, ctx->stream>>>(ctx->dp, a, 1);
        ctx->launches += 1;
    } else {
        if (!ctx->k2_attr_set) {            /* per device, hence per context */
            TSK_CUDA(cudaFuncSetAttribute(k_normalise_score_pack, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)sizeof(K2Smem)));
            TSK_CUDA(cudaFuncSetAttribute(k_normalise_score_pack_compact<K2C_NARROW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)sizeof(K2SmemC<K2C_NARROW>)));
            TSK_CUDA(cudaFuncSetAttribute(k_normalise_score_pack_compact<K2C_WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)sizeof(K2SmemC<K2C_WIDE>)));
            ctx->k2_attr_set = 1;
        }
        const float rcp_maxent = 1.0f / ctx->dp.maximum_entropy;      /* IEEE RN on the host */
        /* at most N work items, whose number is only known on the device: surplus warps return */
        const unsigned grid = (N + K2_CL - 1) / K2_CL;
        CUtensorMap tmap;
        if (ctx->force_generic == 2) {       /* row form (tests): warp = 32 consecutive clusters */
            if (make_cif_tensor_map(ctx, &tmap, 32) != 0) return -1;
            k_normalise_score_pack<<<grid, TSK_K2_THREADS, sizeof(K2Smem), ctx->stream>>>(ctx->dp, a, rcp_maxent, head, tmap);
        } else {
            CUtensorMap tmapw;
            if (make_cif_tensor_map(ctx, &tmap, K2C_NARROW) != 0 || make_cif_tensor_map(ctx, &tmapw, K2C_WIDE) != 0) return -1;
            k_normalise_score_pack_compact<K2C_NARROW><<<grid, TSK_K2_THREADS, sizeof(K2SmemC<K2C_NARROW>), ctx->stream>>>(ctx->dp, a, rcp_maxent, tmap);
            k_normalise_score_pack_compact<K2C_WIDE><<<grid, TSK_K2_THREADS, sizeof(K2SmemC<K2C_WIDE>), ctx->stream>>>(ctx->dp, a, rcp_maxent, tmapw);
            ctx->launches += 1;
        }
        /* handed-over clusters: usually none, many on tiles that are mostly dropped clusters.
         * They keep their scores in the second half of the scratch buffer: the first half still
         * holds the columns of the over-subscribed fair-sampling buckets' clusters, which are
         * sampled after the fix-ups (k_sample_neg_columns) */
        ScoreArgs ah = a;
        ah.scratch = ctx->d_scratch + ctx->scratch_cap;
        k_normalise_score_pack_generic<<<148 * 2, TSK_K2_THREADS, 0, ctx->stream>>>(ctx->dp, ah, 0);
        ctx->launches += 2;
    }
    TSK_CUDA(cudaGetLastError());
    return 0;
}

int tsk_launch_fixups(tsk_ctx *ctx)
{
    const uint32_t N = ctx->tile.nclusters;
    if (N == 0) return 0;
    const DevParams &P = ctx->dp;
    const uint32_t listcap = ctx->cap_clusters;
    cudaStream_t st = ctx->stream;
    const int K = P.fair_max_count;
    if (K > 0) {
        const size_t need = (size_t)K * (size_t)P.fair_hash_size;
        if (ctx->fairslots_cap < need) {
            if (ctx->d_fairslots) cudaFree(ctx->d_fairslots);
            ctx->d_fairslots = NULL; ctx->fairslots_cap = 0;
            TSK_CUDA(cudaMalloc(&ctx->d_fairslots, sizeof(uint32_t) * need));
            ctx->fairslots_cap = need;
        }
        k_fair_init<<<148, 256, 0, st>>>(ctx->d_fairslots, K, ctx->d_bucket, ctx->d_lists, listcap);
        k_fair_insert<<<148, 256, 0, st>>>(ctx->d_fairslots, K, ctx->d_bucket, ctx->d_lists, listcap);
        k_fair_claim<<<148, 256, 0, st>>>(ctx->d_fairslots, K, ctx->d_bucket, ctx->d_lists, listcap);
        ctx->launches += 3;
    }
    k_sample_neg_columns<<<148 * 8, 256, 0, st>>>(P, ctx->d_samples, ctx->d_calls, ctx->d_scratch, ctx->d_neg,
                                                  ctx->d_lists, listcap);
    k_resample_neg<<<148 * 16, TSK_K2_THREADS, 0, st>>>(P, ctx->tile, ctx->d_samples, ctx->d_tscore, ctx->d_calls,
                                                      ctx->d_neg, ctx->d_lists, listcap, 2, 1u);
    k_resample_neg<<<148 * 16, TSK_K2_THREADS, 0, st>>>(P, ctx->tile, ctx->d_samples, ctx->d_tscore, ctx->d_calls,
                                                      ctx->d_neg, ctx->d_lists, listcap, 0, 0xffffffffu);
    ctx->launches += 3;
    TSK_CUDA(cudaGetLastError());
    return 0;
}
