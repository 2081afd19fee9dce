/*
 * tsk_ruler.cu -- polya_ruler kernel: per-record poly(A) length from packed signal records.
 *
 * One WARP per record: the 32 lanes read 32 consecutive packets (one 128-byte line), so the
 * AoS record layout of signal-packs.h is read fully coalesced; the +-1 cumulative sum, its
 * first strict maximum and the downhill extension are warp scans / ballots held in registers.
 *
 * Restates measure_polya_length() (src/polyaruler/polyaruler.c:104-152), the length gate and
 * positive resampling of process_polya_ruling() (:289-299) and add_polya_score_sample()
 * (:154-172).
 */
#include "tsk_internal.h"
#include "tsk_scorebin.cuh"

#define FULL_MASK 0xffffffffu
#define RULER_THREADS 256

/* Record slots come in up to TSK_MAX_SAMPLES segments (one per sample, each with its own
 * signal_dump_length): segment g holds slots [seg.first[g], seg.first[g+1]) of the output
 * array, stored at records + seg.base[g] with (2 + seg.dump[g]) words per slot. */
struct RulerSegs {
    int nseg;
    uint32_t first[TSK_MAX_SAMPLES + 1];
    unsigned long long base[TSK_MAX_SAMPLES];
    int dump[TSK_MAX_SAMPLES];
};

/* Records longer than 256 cycles: the same steps with the packets re-read through L1. */
__device__ __noinline__ int ruler_long_record(const uint32_t *__restrict__ rec, int first, int valid,
                                              const uint32_t *s_cut, float weight, int lane)
{
    int carry = 0, bestkey = INT_MIN;
    const unsigned le_mask = 0xffffffffu >> (31 - lane);
    for (int base = 0; base < valid; base += 32) {
        const int i = base + lane;
        const uint32_t cut = (i < valid && first + i < TSK_MAX_CUTOFF_CYCLES) ? s_cut[first + i] : 0u;
        const uint32_t sc = i < valid ? (__ldg(rec + 2 + i) >> 1) & TSK_SCORE_MAX : 0u;
        const unsigned mu = __ballot_sync(FULL_MASK, cut != 0 && sc >= cut), md = __ballot_sync(FULL_MASK, sc < cut);
        const int cum = carry + __popc(mu & le_mask) - __popc(md & le_mask);
        const int key = (i < valid) ? (((cum + 2048) << 16) | (0xffff - i)) : INT_MIN;
        bestkey = max(bestkey, __reduce_max_sync(FULL_MASK, key));
        carry += __popc(mu) - __popc(md);
    }
    const int maxcum = (bestkey >> 16) - 2048;
    const int argmax = (maxcum > 0) ? (0xffff - (bestkey & 0xffff)) : -1;
    int ext = -1;
    if (argmax >= 0) {
        for (int base = ((argmax + 1) >> 5) << 5; base < valid; base += 32) {
            const int i = base + lane;
            uint32_t x = 0;
            if (i < valid && i > argmax) x = __ldg(rec + 2 + i);
            const bool lit = ((x >> 1) & TSK_SCORE_MAX) != 0;
            const unsigned stop = __ballot_sync(FULL_MASK, lit && !(x & 1u));
            unsigned go = __ballot_sync(FULL_MASK, lit && (x & 1u));
            if (stop) go &= (1u << (__ffs(stop) - 1)) - 1u;
            if (go) ext = base + 31 - __clz(go);
            if (stop) break;
        }
    }
    int len = argmax + 1;
    if (ext >= 0) len += (int)roundf(__fmul_rn((float)(ext - argmax), weight));
    return len;
}

#define RULER_CUT_PAD 320       /* zero cutoffs past the table: first + i needs no range check */

__global__ void __launch_bounds__(RULER_THREADS)
k_polya_ruler(const uint32_t *__restrict__ records, const __grid_constant__ RulerSegs hseg,
              const uint32_t *__restrict__ d_totals, const uint64_t *__restrict__ d_recbase,
              const DevSample *__restrict__ d_samples,
              const uint32_t *__restrict__ g_cutoffs, int ncut, int min_len, float weight,
              int bins, float gap, int16_t *__restrict__ len_out, uint32_t *__restrict__ pos, int intbins,
              unsigned long long *__restrict__ called)
{
    __shared__ uint32_t s_cut[TSK_MAX_CUTOFF_CYCLES + RULER_CUT_PAD];
    __shared__ RulerSegs seg;
    __shared__ unsigned s_called;
    for (int i = threadIdx.x; i < TSK_MAX_CUTOFF_CYCLES + RULER_CUT_PAD; i += RULER_THREADS)
        s_cut[i] = i < ncut ? g_cutoffs[i] : 0u;
    if (threadIdx.x == 0) {
        s_called = 0;
        if (d_totals != nullptr) {
            /* streamed mode: the slot counts of the resident tile are still on the device
             * (hseg.nseg = number of samples) */
            uint32_t at = 0;
            seg.nseg = hseg.nseg;
            for (int s = 0; s < hseg.nseg; s++) {
                seg.first[s] = at;
                seg.base[s] = d_recbase[s];
                seg.dump[s] = d_samples[s].dump_len;
                at += d_totals[s];
            }
            seg.first[hseg.nseg] = at;
        } else {
            seg.nseg = hseg.nseg;
            for (int s = 0; s < hseg.nseg; s++) { seg.first[s] = hseg.first[s]; seg.base[s] = hseg.base[s]; seg.dump[s] = hseg.dump[s]; }
            seg.first[hseg.nseg] = hseg.first[hseg.nseg];
        }
    }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const size_t warps_total = (size_t)gridDim.x * (RULER_THREADS / 32);
    const size_t nrecords = seg.first[seg.nseg];
    const unsigned le_mask = 0xffffffffu >> (31 - lane);
    /* larger cumulative sum wins, then the smaller cycle index */
    const int keybase = (2048 << 16) | (0xffff - lane);
    int g = 0;
    unsigned ncalled = 0;

    for (size_t r = (size_t)blockIdx.x * (RULER_THREADS / 32) + (threadIdx.x >> 5); r < nrecords;
         r += warps_total) {
        while (r >= seg.first[g + 1]) g++;           /* r only grows: segments are visited in order */
        const int dump = seg.dump[g];
        const uint32_t *rec = records + seg.base[g] + (r - seg.first[g]) * (size_t)(2 + dump);
        /* the header and all packets of the slot in flight at once: nine independent loads.  The
         * packet loads do not wait for the header's valid count (a slot always holds `dump`
         * packets); what lies past the count is masked after the fact -- one memory round trip
         * per record instead of two. */
        const uint32_t hdr = __ldg(rec + 1);
        uint32_t w[8];
#pragma unroll
        for (int k = 0; k < 8; k++) w[k] = (k * 32 + lane < dump) ? __ldg(rec + 2 + k * 32 + lane) : 0u;
        int first = (int16_t)(hdr & 0xffffu);
        const int valid = (int16_t)(hdr >> 16);
        if (first < 0) first = 0;
        if (first > TSK_MAX_CUTOFF_CYCLES) first = TSK_MAX_CUTOFF_CYCLES;
        int len = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) w[k] = (k * 32 + lane < valid) ? w[k] : 0u;

        if (valid > 256) {
            len = ruler_long_record(rec, first, valid, s_cut, weight, lane);
        } else if (valid > 0) {
            /* ---- +-1 cumulative sum and its first strict maximum ----
             * the steps are +1 / -1 / 0, so a lane's inclusive prefix sum is the difference of two
             * population counts over ballots; the maximum of (cum, -index) keys is one REDUX. */
            int carry = 0, bestkey = INT_MIN;
            const uint32_t *cutp = s_cut + first + lane;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (k * 32 < valid) {
                    const bool in = k * 32 + lane < valid;
                    const uint32_t cut = in ? cutp[k * 32] : 0u;          /* 0: this cycle does not count */
                    const uint32_t sc = (w[k] >> 1) & TSK_SCORE_MAX;
                    const unsigned mu = __ballot_sync(FULL_MASK, cut != 0 && sc >= cut);
                    const unsigned md = __ballot_sync(FULL_MASK, sc < cut);
                    const int cum = carry + __popc(mu & le_mask) - __popc(md & le_mask);
                    bestkey = max(bestkey, in ? (cum << 16) + (keybase - 32 * k) : INT_MIN);
                    carry += __popc(mu) - __popc(md);
                }
            }
            bestkey = __reduce_max_sync(FULL_MASK, bestkey);
            const int maxcum = (bestkey >> 16) - 2048;
            const int argmax = (maxcum > 0) ? (0xffff - (bestkey & 0xffff)) : -1;

            /* ---- downhill extension: lit downhill packets right after the maximum, up to the
             *      first lit packet that is not downhill ---- */
            int ext = -1;
            if (argmax >= 0) {
                bool open = true;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    if (open && k * 32 + 31 > argmax && k * 32 < valid) {
                        const int i = k * 32 + lane;
                        const uint32_t x = (i > argmax) ? w[k] : 0u;          /* w[] is 0 past valid */
                        const bool lit = ((x >> 1) & TSK_SCORE_MAX) != 0;
                        const unsigned stop = __ballot_sync(FULL_MASK, lit && !(x & 1u));
                        unsigned go = __ballot_sync(FULL_MASK, lit && (x & 1u));
                        if (stop) { go &= (1u << (__ffs(stop) - 1)) - 1u; open = false; }
                        if (go) ext = k * 32 + 31 - __clz(go);
                    }
                }
            }
            len = argmax + 1;
            if (ext >= 0) len += (int)roundf(__fmul_rn((float)(ext - argmax), weight));
        }

        const bool pass = len >= min_len;
        if (lane == 0) { len_out[r] = pass ? (int16_t)len : (int16_t)-1; ncalled += pass; }
        if (pass) {
            /* positive score samples of the called tail, polyaruler.c:154-172,296-298 */
            const int take = min(len - __float2int_rz(__fmul_rn((float)len, gap)), ncut - first);
            uint32_t *row = pos + (size_t)(first + lane) * bins;
            if (valid <= 256) {
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    if (k * 32 < take) {
                        const uint32_t sc = (w[k] >> 1) & TSK_SCORE_MAX;
                        if (k * 32 + lane < take && sc > 0)
                            atomicAdd(row + (size_t)(k * 32) * bins + score_bin(sc, bins, intbins), 1u);
                    }
                }
            } else {
                for (int i = lane; i < take; i += 32) {
                    const uint32_t sc = (__ldg(rec + 2 + i) >> 1) & TSK_SCORE_MAX;
                    if (sc > 0) atomicAdd(pos + (size_t)(first + i) * bins + score_bin(sc, bins, intbins), 1u);
                }
            }
        }
    }
    if (lane == 0 && ncalled) atomicAdd(&s_called, ncalled);
    __syncthreads();
    if (threadIdx.x == 0 && s_called) atomicAdd(called, (unsigned long long)s_called);
}

int tsk_launch_ruler_segs(tsk_ctx *ctx, const uint32_t *d_records, int nseg, const uint32_t *first,
                          const uint64_t *base, const int *dump, int ncutoffs, int min_len, float weight,
                          int bins, float gap, int16_t *d_len_out)
{
    RulerSegs seg;
    seg.nseg = nseg;
    for (int i = 0; i < nseg; i++) { seg.first[i] = first[i]; seg.base[i] = base[i]; seg.dump[i] = dump[i]; }
    seg.first[nseg] = first[nseg];
    const size_t nrecords = first[nseg];
    if (nrecords == 0) return 0;
    size_t blocks = (nrecords + (RULER_THREADS / 32) - 1) / (RULER_THREADS / 32);
    const size_t maxblocks = 148 * 8 * 4;       /* persistent-ish: a few waves of 148 SMs x 8 CTAs */
    if (blocks > maxblocks) blocks = maxblocks;
    /* 2^23 - 1 = 47 * 178481 */
    const int intbins = (bins % 47 != 0) && (bins % 178481 != 0) && bins <= 65536;
    k_polya_ruler<<<(unsigned)blocks, RULER_THREADS, 0, ctx->stream>>>(
        d_records, seg, nullptr, nullptr, nullptr, ctx->d_cutoffs, ncutoffs, min_len, weight, bins, gap, d_len_out,
        ctx->d_ruler_pos, intbins, ctx->d_ruler_called);
    ctx->launches++;
    TSK_CUDA(cudaGetLastError());
    return 0;
}

/* Every record slot of the resident tile, with the per-sample slot counts read on the device
 * (nothing about the tile needs to be known on the host yet). */
int tsk_launch_ruler_resident(tsk_ctx *ctx, cudaStream_t st, int ncutoffs, int min_len, float weight, int bins,
                              float gap, int16_t *d_len_out)
{
    TskRange nvtx("tsk:polya_ruler");
    RulerSegs seg;
    seg.nseg = ctx->dp.num_samples;
    size_t blocks = ((size_t)ctx->tile.nclusters + (RULER_THREADS / 32) - 1) / (RULER_THREADS / 32);
    const size_t maxblocks = 148 * 8 * 4;
    if (blocks > maxblocks) blocks = maxblocks;
    if (blocks == 0) return 0;
    const int intbins = (bins % 47 != 0) && (bins % 178481 != 0) && bins <= 65536;
    k_polya_ruler<<<(unsigned)blocks, RULER_THREADS, 0, st>>>(
        ctx->d_records, seg, ctx->d_rtotals, ctx->d_rrecbase, ctx->d_samples, ctx->d_cutoffs, ncutoffs, min_len,
        weight, bins, gap, d_len_out, ctx->d_ruler_pos, intbins, ctx->d_ruler_called);
    ctx->launches++;
    TSK_CUDA(cudaGetLastError());
    return 0;
}

int tsk_launch_ruler(tsk_ctx *ctx, const uint32_t *d_records, size_t nrecords, int dump_len,
                     int ncutoffs, int min_len, float weight, int bins, float gap, int16_t *d_len_out)
{
    TskRange nvtx("tsk:polya_ruler");
    const uint32_t first[2] = {0, (uint32_t)nrecords};
    const uint64_t base[1] = {0};
    const int dump[1] = {dump_len};
    return tsk_launch_ruler_segs(ctx, d_records, 1, first, base, dump, ncutoffs, min_len, weight, bins, gap,
                                 d_len_out);
}
