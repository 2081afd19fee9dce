/*
 * tsk_encode.cu -- the importer's output files made on the device.
 *
 * What the reference does on the host for every written cluster -- format_basecalls()
 * (bclreader.c:146-166), write_seqqual_entry() / write_taginfo_entry() (spotanalyzer.c:206-307),
 * write_polya_score()'s record bytes (:389-438) and, above all, bgzf_write() (spotanalyzer.c:355-386,
 * 71 % of its single-thread time) -- happens here in four steps on the resident tile:
 *
 *   k_enc_lengths   per cluster: the lengths of its taginfo and seqqual lines; per-sample
 *                   offsets in cluster order (the reference's job order with threads = 1)
 *   k_enc_format    the text: the CTA's [cycles][128 clusters] base-call slab is transposed into
 *                   shared memory once, then a warp writes one line at a time, lane = character
 *   k_enc_live      record slots that were withdrawn after scoring are squeezed out
 *   k_enc_hist / k_enc_codes   byte histogram of every stream (seqqual, taginfo, signal records of
 *                   each sample) and ONE length-limited Huffman code per stream
 *   k_enc_deflate   every stream is cut into 32 KB pieces; one CTA turns a piece into one BGZF
 *                   block: a dynamic-Huffman deflate block WITHOUT matches (or a stored block when
 *                   that is smaller) and its CRC-32.  On these streams (bases, qualities, noisy
 *                   24-bit scores) zlib's LZ77 finds next to nothing: level 1 gives 0.394 / 0.497 /
 *                   0.782 of the input, Huffman only 0.417 / 0.492 / 0.781 (profiles/r2_encode.txt).
 *   k_enc_gather    the blocks of a stream are packed back to back
 *
 * The host then only write()s: every stream is a valid BGZF file body (RFC 1952 members with the
 * `BC` extra field, <= 64 KB each), which zlib / gzip / htslib read like the reference's output.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "tsk_internal.h"
#include "tsk_device.cuh"

#define FULL_MASK 0xffffffffu

#define ENC_CHUNK    32768                     /* input bytes per BGZF block */
#define ENC_THREADS  256
#define ENC_SLICE    (ENC_CHUNK / ENC_THREADS) /* 128 bytes per thread */
#define ENC_SLOT     (ENC_CHUNK + 64)          /* room for one member (stored form: 18 + 5 + 32768 + 8) */
#define ENC_NSYM     257                       /* literals + end of block */
#define ENC_HDR_BITS (3 + 5 + 5 + 4 + 19 * 3 + (ENC_NSYM + 1) * 4)     /* = 1106 */
#define ENC_FMT_CL   128                       /* clusters per CTA of the two text kernels */

struct DevEncSample {
    int32_t want[3];                           /* seqqual, taginfo, signal */
    int32_t umi_count, umi_total;
    int32_t umi_start[TSK_ENCODE_MAX_UMI], umi_length[TSK_ENCODE_MAX_UMI];
};

struct DevStream {                             /* one output stream: 3 * sample + kind */
    const uint8_t *src;
    uint64_t nbytes;
    uint32_t first_block, nblocks;
};

struct StreamCode;

struct EncState {
    tsk_encode_layout layout;
    DevEncSample *d_samples;
    uint2    *d_lineoff;                       /* [cap] in-CTA offsets of the cluster's two lines */
    uint32_t *d_ctasum;                        /* [nblk][2 S] line bytes per CTA, then exclusive bases */
    uint32_t *d_texttot;                       /* [2 S] */
    uint8_t  *d_text;      size_t text_cap;
    uint32_t *d_liverank;                      /* [slots + 1] live slots before each slot */
    uint32_t *d_livecta;   size_t livecta_cap; /* per 1024 slots */
    uint32_t *d_livetot;                       /* [1] */
    uint32_t *d_live;      size_t live_cap_words;
    DevStream *d_streams;                      /* [3 S + 1] */
    uint32_t *d_nblocks;                       /* [1] */
    uint8_t  *d_slots;     size_t nblk_cap;
    uint32_t *d_sizes;                         /* [nblk_cap] */
    uint64_t *d_offsets;                       /* [nblk_cap + 1] */
    uint64_t *d_streamoff;                     /* [3 S + 1] */
    uint8_t  *d_final;     size_t final_cap;
    uint32_t *d_crctab;                        /* [256] byte table | [256] x^(1024 m) shifts */
    uint32_t *d_hist;                          /* [3 S][256] byte histogram of every stream */
    StreamCode *d_codes;                       /* [3 S] */
    uint8_t  *h_final;     size_t h_final_cap; /* pinned */
    uint64_t h_streamoff[3 * TSK_MAX_SAMPLES + 1];
    uint32_t cap_clusters;
    int tagmax, sqmax, have;
    size_t layout_smem;
    double last_ms;
    cudaEvent_t ev[2];
};

/* ---- small helpers ---------------------------------------------------------------------- */
__device__ __forceinline__ int ndigits_u32(uint32_t v)
{
    int n = 1;
    if (v >= 10u) n = 2;
    if (v >= 100u) n = 3;
    if (v >= 1000u) n = 4;
    if (v >= 10000u) n = 5;
    if (v >= 100000u) n = 6;
    if (v >= 1000000u) n = 7;
    if (v >= 10000000u) n = 8;
    if (v >= 100000000u) n = 9;
    if (v >= 1000000000u) n = 10;
    return n;
}
__device__ __forceinline__ int nchars_i32(int v) { return v < 0 ? 1 + ndigits_u32((uint32_t)(-(int64_t)v)) : ndigits_u32((uint32_t)v); }

__constant__ uint32_t c_pow10[10] = {1u, 10u, 100u, 1000u, 10000u, 100000u, 1000000u, 10000000u, 100000000u, 1000000000u};

/* character `pos` (from the left) of printf("%d", v) */
__device__ __forceinline__ char char_of_i32(int v, int pos)
{
    uint32_t m = (uint32_t)v;
    if (v < 0) { if (pos == 0) return '-'; pos--; m = (uint32_t)(-(int64_t)v); }
    const int nd = ndigits_u32(m);
    return (char)('0' + (m / c_pow10[nd - 1 - pos]) % 10u);
}

/* lengths of the two lines of a written cluster (0 when its sample has no such file) */
__device__ __forceinline__ void line_lengths(const tsk_call &c, const DevEncSample &es, uint32_t clusterno,
                                             int ts, int tl, int fl, int out3, int &taglen, int &sqlen)
{
    const int nd = ndigits_u32(clusterno);
    taglen = sqlen = 0;
    if (es.want[1])
        taglen = nd + 1 + nchars_i32((int)c.flags) + 1 + nchars_i32((int)c.polya_len) + 1 +
                 (c.terminal_mods > 0 ? c.terminal_mods : 0) + 1 + es.umi_total + 1;
    if (es.want[0]) {
        int len3 = c.delimiter_end >= 0 ? ts + tl - c.delimiter_end : tl;
        if (len3 > out3) len3 = out3;
        sqlen = nd + 1 + fl + 1 + fl + 1 + len3 + 1 + len3 + 1;
    }
}

/* ---- k_enc_lengths ------------------------------------------------------------------------ */
__global__ void __launch_bounds__(ENC_FMT_CL)
k_enc_lengths(const __grid_constant__ DevParams P, const TileView tv, const tsk_call *__restrict__ calls,
              const DevEncSample *__restrict__ g_es, int fl, int out3,
              uint2 *__restrict__ lineoff, uint32_t *__restrict__ ctasum)
{
    __shared__ DevEncSample s_es[TSK_MAX_SAMPLES];
    __shared__ uint32_t s_w[ENC_FMT_CL / 32][2 * TSK_MAX_SAMPLES];
    const int S = P.num_samples, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < (int)(sizeof(DevEncSample) / 4) * S; i += ENC_FMT_CL)
        reinterpret_cast<uint32_t *>(s_es)[i] = reinterpret_cast<const uint32_t *>(g_es)[i];
    for (int i = tid; i < (ENC_FMT_CL / 32) * 2 * TSK_MAX_SAMPLES; i += ENC_FMT_CL) (&s_w[0][0])[i] = 0;
    __syncthreads();
    const uint32_t n = blockIdx.x * ENC_FMT_CL + tid;
    int sample = -1, taglen = 0, sqlen = 0;
    if (n < tv.nclusters) {
        const uint4 raw = *reinterpret_cast<const uint4 *>(&calls[n]);
        const tsk_call c = *reinterpret_cast<const tsk_call *>(&raw);
        if (c.written) {
            sample = c.sample;
            line_lengths(c, s_es[sample], tv.firstclusterno + n, P.threep_start, P.threep_length, fl, out3, taglen, sqlen);
        }
    }
    /* exclusive sums over the earlier clusters of the SAME sample: one warp scan per sample present */
    uint32_t tagoff = 0, sqoff = 0;
    for (unsigned todo = __ballot_sync(FULL_MASK, sample >= 0); todo;) {
        const int s = __shfl_sync(FULL_MASK, sample, __ffs(todo) - 1);
        const bool mine = sample == s;
        uint32_t a = mine ? (uint32_t)taglen : 0u, b = mine ? (uint32_t)sqlen : 0u;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t ua = __shfl_up_sync(FULL_MASK, a, d), ub = __shfl_up_sync(FULL_MASK, b, d);
            if (lane >= d) { a += ua; b += ub; }
        }
        if (mine) { tagoff = a - (uint32_t)taglen; sqoff = b - (uint32_t)sqlen; }
        if (lane == 31) { s_w[warp][2 * s] = b; s_w[warp][2 * s + 1] = a; }      /* [2 s] seqqual, [2 s + 1] taginfo */
        todo &= ~__ballot_sync(FULL_MASK, mine);
    }
    __syncthreads();
    if (sample >= 0)
        for (int w = 0; w < warp; w++) { sqoff += s_w[w][2 * sample]; tagoff += s_w[w][2 * sample + 1]; }
    if (n < tv.nclusters) lineoff[n] = make_uint2(tagoff, sqoff);
    for (int i = tid; i < 2 * S; i += ENC_FMT_CL) {
        uint32_t t = 0;
        for (int w = 0; w < ENC_FMT_CL / 32; w++) t += s_w[w][i];
        ctasum[(size_t)blockIdx.x * (2 * S) + i] = t;
    }
}

/* exclusive scan of the columns of a [rows][ncol] u32 table, one CTA per column */
__global__ void __launch_bounds__(1024)
k_enc_colscan(uint32_t *__restrict__ tab, uint32_t *__restrict__ totals, int rows, int ncol)
{
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const int col = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < rows; base += 1024 * 8) {
        const int r0 = base + tid * 8;
        uint32_t v[8], sum = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) { v[k] = (r0 + k < rows) ? tab[(size_t)(r0 + k) * ncol + col] : 0u; sum += v[k]; }
        uint32_t incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(FULL_MASK, incl, d); if (lane >= d) incl += o; }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const uint32_t w = s_warp[lane];
            uint32_t wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(FULL_MASK, wi, d); if (lane >= d) wi += o; }
            s_warp[lane] = wi - w;
        }
        __syncthreads();
        uint32_t run = s_carry + s_warp[warp] + (incl - sum);
#pragma unroll
        for (int k = 0; k < 8; k++) { if (r0 + k < rows) tab[(size_t)(r0 + k) * ncol + col] = run; run += v[k]; }
        __syncthreads();
        if (tid == 1023) s_carry = run;
        __syncthreads();
    }
    if (tid == 0) totals[col] = s_carry;
}

/* ---- k_enc_live: ranks of the record slots that were not withdrawn ------------------------- */
__device__ __forceinline__ int sample_of_slot(const uint32_t *totals, int S, uint32_t g, uint32_t &k)
{
    uint32_t at = 0;
    for (int s = 0; s < S; s++) {
        if (g < at + totals[s]) { k = g - at; return s; }
        at += totals[s];
    }
    k = 0;
    return -1;
}

__global__ void __launch_bounds__(1024)
k_enc_live_count(const uint32_t *__restrict__ records, const uint32_t *__restrict__ totals,
                 const uint64_t *__restrict__ recbase, const DevSample *__restrict__ samples, int S,
                 uint32_t *__restrict__ liverank, uint32_t *__restrict__ livecta)
{
    __shared__ uint32_t s_warp[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t nslots = 0;
    for (int s = 0; s < S; s++) nslots += totals[s];
    if ((uint64_t)blockIdx.x * 1024 >= nslots) return;
    const uint32_t g = blockIdx.x * 1024 + tid;
    uint32_t live = 0;
    if (g < nslots) {
        uint32_t k;
        const int s = sample_of_slot(totals, S, g, k);
        const uint32_t *rec = records + recbase[s] + (size_t)k * (size_t)(2 + samples[s].dump_len);
        live = (int16_t)(rec[1] >> 16) >= 0 ? 1u : 0u;
    }
    const unsigned b = __ballot_sync(FULL_MASK, live);
    const uint32_t inwarp = __popc(b & ((1u << lane) - 1u));
    if (lane == 0) s_warp[warp] = __popc(b);
    __syncthreads();
    uint32_t before = 0;
    for (int w = 0; w < warp; w++) before += s_warp[w];
    if (g < nslots) liverank[g] = before + inwarp;            /* within this CTA; the base is added later */
    if (tid == 1023) livecta[blockIdx.x] = before + inwarp + live;
}

/* ranks become global (CTA-local rank + scanned CTA base); entry [nslots] = number of live slots */
__global__ void __launch_bounds__(256)
k_enc_live_ranks(const uint32_t *__restrict__ totals, int S, uint32_t *__restrict__ liverank,
                 const uint32_t *__restrict__ livecta, const uint32_t *__restrict__ livetot)
{
    uint32_t nslots = 0;
    for (int s = 0; s < S; s++) nslots += totals[s];
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g <= nslots; g += gridDim.x * blockDim.x)
        liverank[g] = g == nslots ? *livetot : liverank[g] + livecta[g >> 10];
}

/* the squeeze itself, once the ranks are settled: nothing to do (and nothing copied) when every
 * slot is live -- the streams then read the record array in place */
__global__ void __launch_bounds__(256)
k_enc_live_place(const uint32_t *__restrict__ records, const uint32_t *__restrict__ totals,
                 const uint64_t *__restrict__ recbase, const DevSample *__restrict__ samples, int S,
                 const uint32_t *__restrict__ liverank, uint32_t *__restrict__ live)
{
    __shared__ uint32_t s_first[TSK_MAX_SAMPLES + 1];
    __shared__ unsigned long long s_wbase[TSK_MAX_SAMPLES + 1];
    if (threadIdx.x == 0) {
        uint32_t at = 0;
        unsigned long long w = 0;
        for (int s = 0; s < S; s++) {
            s_first[s] = at; s_wbase[s] = w;
            const uint32_t nl = liverank[at + totals[s]] - liverank[at];
            w += (unsigned long long)nl * (unsigned long long)(2 + samples[s].dump_len);
            at += totals[s];
        }
        s_first[S] = at; s_wbase[S] = w;
    }
    __syncthreads();
    const uint32_t nslots = s_first[S];
    if (liverank[nslots] == nslots) return;                    /* nothing withdrawn: read in place */
    const int lane = threadIdx.x & 31;
    for (uint32_t g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); g < nslots; g += gridDim.x * (blockDim.x >> 5)) {
        int s = 0;
        while (g >= s_first[s + 1]) s++;
        const int words = 2 + samples[s].dump_len;
        const uint32_t *rec = records + recbase[s] + (size_t)(g - s_first[s]) * (size_t)words;
        if ((int16_t)(rec[1] >> 16) < 0) continue;
        uint32_t *dst = live + s_wbase[s] + (size_t)(liverank[g] - liverank[s_first[s]]) * (size_t)words;
        for (int w = lane; w < words; w += 32) dst[w] = rec[w];
    }
}

/* ---- stream table --------------------------------------------------------------------------- */
__global__ void k_enc_streams(const DevEncSample *__restrict__ es, const DevSample *__restrict__ samples, int S,
                              const uint32_t *__restrict__ texttot, uint8_t *text,
                              const uint32_t *__restrict__ totals, const uint64_t *__restrict__ recbase,
                              const uint32_t *__restrict__ records, const uint32_t *__restrict__ liverank,
                              const uint32_t *__restrict__ live, DevStream *__restrict__ streams,
                              uint32_t *__restrict__ nblocks, uint32_t *__restrict__ textbase)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint32_t nslots = 0;
    for (int s = 0; s < S; s++) nslots += totals[s];
    const bool inplace = liverank[nslots] == nslots;
    uint64_t tat = 0, wat = 0;
    uint32_t blk = 0, first = 0;
    for (int s = 0; s < S; s++) {
        for (int kind = 0; kind < 3; kind++) {
            DevStream st;
            st.src = nullptr; st.nbytes = 0;
            if (kind < 2) {
                textbase[2 * s + kind] = (uint32_t)tat;
                if (es[s].want[kind]) {
                    st.src = text + tat; st.nbytes = texttot[2 * s + kind];
                    tat += (st.nbytes + 15) & ~(uint64_t)15;
                }
            } else {
                const uint32_t nl = liverank[first + totals[s]] - liverank[first];
                const uint64_t words = (uint64_t)nl * (uint64_t)(2 + samples[s].dump_len);
                if (es[s].want[2]) {
                    st.src = reinterpret_cast<const uint8_t *>(inplace ? records + recbase[s] : live + wat);
                    st.nbytes = words * 4;
                }
                wat += words;
                first += totals[s];
            }
            st.first_block = blk;
            st.nblocks = (uint32_t)((st.nbytes + ENC_CHUNK - 1) / ENC_CHUNK);
            blk += st.nblocks;
            streams[3 * s + kind] = st;
        }
    }
    DevStream end;
    end.src = nullptr; end.nbytes = 0; end.first_block = blk; end.nblocks = 0;
    streams[3 * S] = end;
    *nblocks = blk;
}

/* ---- k_enc_format --------------------------------------------------------------------------- */
__global__ void __launch_bounds__(ENC_FMT_CL)
k_enc_format(const __grid_constant__ DevParams P, const TileView tv, const tsk_call *__restrict__ calls,
             const DevEncSample *__restrict__ g_es, int f0, int fl, int out3, int tpitch,
             const uint2 *__restrict__ lineoff, const uint32_t *__restrict__ ctabase,
             const uint32_t *__restrict__ textbase, uint8_t *__restrict__ text)
{
    extern __shared__ __align__(16) unsigned char enc_smem[];
    /* [128 clusters][tpitch] raw base-call bytes (a cluster's cycles are consecutive) | samples */
    uint8_t *slab = enc_smem;
    DevEncSample *s_es = reinterpret_cast<DevEncSample *>(enc_smem + (size_t)ENC_FMT_CL * tpitch);
    const int S = P.num_samples, T = P.total_cycles, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < (int)(sizeof(DevEncSample) / 4) * S; i += ENC_FMT_CL)
        reinterpret_cast<uint32_t *>(s_es)[i] = reinterpret_cast<const uint32_t *>(g_es)[i];
    const uint32_t n0 = blockIdx.x * ENC_FMT_CL;
    {   /* transpose: thread = cluster column, one coalesced row of 128 bytes per cycle */
        const uint32_t n = n0 + tid;
        const uint8_t *col = tv.bcl + n;
        uint8_t *mine = slab + (size_t)tid * tpitch;
        if (n < tv.nclusters)
            for (int r = 0; r < T; r++) mine[r] = col[(size_t)r * tv.bcl_pitch];
    }
    __syncthreads();
    const int ts = P.threep_start, tl = P.threep_length;
    for (int i = 0; i < 32; i++) {
        const int ci = warp * 32 + i;
        const uint32_t n = n0 + ci;
        if (n >= tv.nclusters) break;
        const uint4 raw = *reinterpret_cast<const uint4 *>(&calls[n]);
        const tsk_call c = *reinterpret_cast<const tsk_call *>(&raw);
        if (!c.written) continue;
        const DevEncSample &es = s_es[c.sample];
        const uint32_t clusterno = tv.firstclusterno + n;
        int taglen, sqlen;
        line_lengths(c, es, clusterno, ts, tl, fl, out3, taglen, sqlen);
        const uint2 off = lineoff[n];
        const uint8_t *row = slab + (size_t)ci * tpitch;
        const int nd = ndigits_u32(clusterno);
        const size_t cb = (size_t)blockIdx.x * (2 * S);
        if (sqlen) {
            uint8_t *dst = text + textbase[2 * c.sample] + ctabase[cb + 2 * c.sample] + off.y;
            const int start3 = c.delimiter_end >= 0 ? c.delimiter_end : ts;
            int len3 = c.delimiter_end >= 0 ? ts + tl - c.delimiter_end : tl;
            if (len3 > out3) len3 = out3;
            /* cluster \t seq5 \t qual5 \t seq3 \t qual3 \n */
            const int b1 = nd + 1, b2 = b1 + fl + 1, b3 = b2 + fl + 1, b4 = b3 + len3 + 1;
            for (int p = lane; p < sqlen; p += 32) {
                char ch;
                if (p < nd) ch = (char)('0' + (clusterno / c_pow10[nd - 1 - p]) % 10u);
                else if (p == nd || p == b2 - 1 || p == b3 - 1 || p == b4 - 1) ch = '\t';
                else if (p == sqlen - 1) ch = '\n';
                else {
                    int cyc; bool q;
                    if (p < b2) { cyc = f0 + (p - b1); q = false; }
                    else if (p < b3) { cyc = f0 + (p - b2); q = true; }
                    else if (p < b4) { cyc = start3 + (p - b3); q = false; }
                    else { cyc = start3 + (p - b4); q = true; }
                    const uint32_t b = row[cyc];
                    /* format_basecalls(): byte 0 = no call: 'N' with quality 2 */
                    ch = q ? (char)(33 + tsk_base_qual(b)) : (b ? "ACGT"[b & 3] : 'N');
                }
                dst[p] = (uint8_t)ch;
            }
        }
        if (taglen) {
            uint8_t *dst = text + textbase[2 * c.sample + 1] + ctabase[cb + 2 * c.sample + 1] + off.x;
            /* cluster \t flags \t polyA \t mods \t umi \n */
            const int nf = nchars_i32((int)c.flags), np = nchars_i32((int)c.polya_len);
            const int tm = c.terminal_mods > 0 ? c.terminal_mods : 0;
            const int b1 = nd + 1, b2 = b1 + nf + 1, b3 = b2 + np + 1, b4 = b3 + tm + 1;
            for (int p = lane; p < taglen; p += 32) {
                char ch;
                if (p < nd) ch = (char)('0' + (clusterno / c_pow10[nd - 1 - p]) % 10u);
                else if (p == nd || p == b2 - 1 || p == b3 - 1 || p == b4 - 1) ch = '\t';
                else if (p == taglen - 1) ch = '\n';
                else if (p < b2) ch = char_of_i32((int)c.flags, p - b1);
                else if (p < b3) ch = char_of_i32((int)c.polya_len, p - b2);
                else if (p < b4) {
                    /* reverse complement of the modification bases (spotanalyzer.c:247-267) */
                    const uint32_t b = row[c.delimiter_end + tm - 1 - (p - b3)];
                    ch = b ? "TGCA"[b & 3] : 'N';
                } else {
                    int k = p - b4, cyc = 0;
                    for (int u = 0; u < es.umi_count; u++) {
                        const int ln = es.umi_length[u] > 0 ? es.umi_length[u] : 0;
                        if (k < ln) { cyc = es.umi_start[u] + k; break; }
                        k -= ln;
                    }
                    const uint32_t b = row[cyc];
                    ch = b ? "ACGT"[b & 3] : 'N';
                }
                dst[p] = (uint8_t)ch;
            }
        }
    }
}

/* ---- CRC-32 (the gzip polynomial, reflected) ------------------------------------------------- */
#define CRC_POLY 0xedb88320u
/* a(x) * b(x) mod p(x) on reflected 32-bit polynomials (bit 31 = x^0) */
__host__ __device__ static inline uint32_t crc_mulmod(uint32_t a, uint32_t b)
{
    uint32_t p = 0;
    for (int i = 31; i >= 0; i--) {
        if ((a >> i) & 1u) p ^= b;
        b = (b & 1u) ? (b >> 1) ^ CRC_POLY : b >> 1;
    }
    return p;
}

/* ---- deflate: one Huffman code per STREAM, one BGZF block per 32 KB piece ---------------------
 * A stream's pieces are statistically alike (text of one sample, records of one sample), so the
 * code is built once per stream from its whole byte histogram (every symbol counted at least once,
 * so any piece can be written with it) instead of once per piece: the only serial part of the
 * encoder -- the Huffman merge, a few tens of microseconds on one thread -- runs 3 x samples times
 * per tile, not 30 000 times.  Each piece still carries its own deflate header (138 bytes), so every
 * BGZF block stands alone. */
struct StreamCode {                       /* built by k_enc_codes, read by k_enc_deflate */
    uint32_t header[(ENC_HDR_BITS + 31) / 32];
    uint16_t code[ENC_NSYM + 3];          /* bit-reversed: Huffman codes go out MSB first */
    uint8_t  len[ENC_NSYM + 3];
};

static size_t enc_codes_bytes(int nstreams) { return sizeof(StreamCode) * (size_t)nstreams; }

__device__ __forceinline__ int stream_of_block(const DevStream *__restrict__ streams, int nstreams, uint32_t b)
{
    int s = 0;
    while (s + 1 < nstreams && b >= streams[s + 1].first_block) s++;
    while (streams[s].nblocks == 0) s++;                   /* empty streams share first_block with the next one */
    return s;
}

/* byte histogram of every stream (global atomics of per-CTA counts) */
__global__ void __launch_bounds__(ENC_THREADS)
k_enc_hist(const DevStream *__restrict__ streams, int nstreams, const uint32_t *__restrict__ nblocks,
           uint32_t *__restrict__ hist /* [nstreams][256] */)
{
    __shared__ uint32_t s_hist[ENC_THREADS / 32][256];
    __shared__ int s_stream;
    const uint32_t b = blockIdx.x;
    if (b >= *nblocks) return;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (ENC_THREADS / 32) * 256; i += ENC_THREADS) (&s_hist[0][0])[i] = 0;
    if (tid == 0) s_stream = stream_of_block(streams, nstreams, b);
    __syncthreads();
    const DevStream st = streams[s_stream];
    const uint64_t off = (uint64_t)(b - st.first_block) * ENC_CHUNK;
    const uint64_t left = st.nbytes - off;
    const uint32_t L = left < ENC_CHUNK ? (uint32_t)left : ENC_CHUNK;
    const uint8_t *src = st.src + off;
    if ((reinterpret_cast<uintptr_t>(src) & 3) == 0) {
        const uint32_t *s4 = reinterpret_cast<const uint32_t *>(src);
        for (uint32_t i = tid; i < (L >> 2); i += ENC_THREADS) {
            const uint32_t w = s4[i];
            atomicAdd(&s_hist[warp][w & 0xff], 1u); atomicAdd(&s_hist[warp][(w >> 8) & 0xff], 1u);
            atomicAdd(&s_hist[warp][(w >> 16) & 0xff], 1u); atomicAdd(&s_hist[warp][w >> 24], 1u);
        }
        for (uint32_t i = (L & ~3u) + tid; i < L; i += ENC_THREADS) atomicAdd(&s_hist[warp][src[i]], 1u);
    } else {
        for (uint32_t i = tid; i < L; i += ENC_THREADS) atomicAdd(&s_hist[warp][src[i]], 1u);
    }
    __syncthreads();
    uint32_t f = 0;
    for (int w = 0; w < ENC_THREADS / 32; w++) f += s_hist[w][tid];
    if (f) atomicAdd(&hist[(size_t)s_stream * 256 + tid], f);
}

/* one CTA per stream: length-limited Huffman code over {256 literals, end of block}, canonical
 * codes, and the deflate block header that announces them */
__global__ void __launch_bounds__(ENC_THREADS)
k_enc_codes(const DevStream *__restrict__ streams, const uint32_t *__restrict__ hist, StreamCode *__restrict__ codes)
{
    __shared__ uint32_t freq[ENC_NSYM + 3];
    __shared__ uint32_t nodefreq[ENC_NSYM];
    __shared__ uint16_t order[ENC_NSYM + 3];
    __shared__ uint16_t parent[2 * ENC_NSYM];
    __shared__ uint8_t  depth[2 * ENC_NSYM];
    __shared__ uint8_t  len[ENC_NSYM + 3];
    __shared__ uint16_t code[ENC_NSYM + 3];
    __shared__ uint32_t header[(ENC_HDR_BITS + 31) / 32];
    const int sidx = blockIdx.x, tid = threadIdx.x;
    if (streams[sidx].nblocks == 0) return;
    /* every literal counted at least once: a code exists for whatever a piece holds */
    for (int s = tid; s < ENC_NSYM; s += ENC_THREADS) {
        freq[s] = s < 256 ? hist[(size_t)sidx * 256 + s] + 1u : streams[sidx].nblocks;
        len[s] = 0;
    }
    for (int i = tid; i < (ENC_HDR_BITS + 31) / 32; i += ENC_THREADS) header[i] = 0;
    __syncthreads();
    /* symbols ordered by (frequency, symbol): rank = number of smaller keys */
    for (int s = tid; s < ENC_NSYM; s += ENC_THREADS) {
        const uint32_t f = freq[s];
        int rank = 0;
        for (int q = 0; q < ENC_NSYM; q++) {
            const uint32_t g = freq[q];
            rank += g < f || (g == f && q < s);
        }
        order[rank] = (uint16_t)s;
    }
    __syncthreads();
    if (tid == 0) {
        const int n = ENC_NSYM;
        /* two queues: sorted leaves and internal nodes in creation order (their weights ascend too) */
        int li = 0, ni = 0;
        for (int made = 0; made < n - 1; made++) {
            uint32_t w = 0;
            for (int pick = 0; pick < 2; pick++) {
                const bool leaf = li < n && (ni >= made || freq[order[li]] <= nodefreq[ni]);
                if (leaf) { w += freq[order[li]]; parent[li] = (uint16_t)(n + made); li++; }
                else { w += nodefreq[ni]; parent[n + ni] = (uint16_t)(n + made); ni++; }
            }
            nodefreq[made] = w;
        }
        int count[33];
        for (int i = 0; i < 33; i++) count[i] = 0;
        depth[2 * n - 2] = 0;
        for (int id = 2 * n - 3; id >= 0; id--) {
            int d = depth[parent[id]] + 1;
            if (d > 40) d = 40;                               /* everything above 15 is folded below anyway */
            depth[id] = (uint8_t)d;
            if (id < n) count[d > 32 ? 32 : d]++;
        }
        /* fold the lengths above 15 into 15 and repair the Kraft sum: each step gives one code of
         * the longest shorter length a sibling, at the cost of one 15-bit code */
        for (int i = 16; i <= 32; i++) count[15] += count[i];
        uint32_t total = 0;
        for (int i = 15; i > 0; i--) total += (uint32_t)count[i] << (15 - i);
        while (total != (1u << 15)) {
            count[15]--;
            for (int i = 14; i > 0; i--)
                if (count[i]) { count[i]--; count[i + 1] += 2; break; }
            total--;
        }
        /* the most frequent symbols get the shortest codes */
        int j = n;
        for (int i = 1; i <= 15; i++)
            for (int l = count[i]; l > 0; l--) len[order[--j]] = (uint8_t)i;
        /* canonical codes (RFC 1951 3.2.2) */
        int blcount[16], next[16];
        for (int i = 0; i < 16; i++) blcount[i] = 0;
        for (int s = 0; s < ENC_NSYM; s++) blcount[len[s]]++;
        blcount[0] = 0;
        int c = 0;
        for (int bits = 1; bits < 16; bits++) { c = (c + blcount[bits - 1]) << 1; next[bits] = c; }
        for (int s = 0; s < ENC_NSYM; s++) {
            const int l = len[s];
            code[s] = (uint16_t)(__brev((uint32_t)next[l]++) >> (32 - l));
        }
        /* header: BFINAL = 1, BTYPE = dynamic, 257 literal/length codes, 1 distance code of length 0
         * (no distances are used), the code-length alphabet {0..15} with 4 bits each: 19 three-bit
         * lengths in the order 16 17 18 0 8 7 9 6 10 5 11 4 12 3 13 2 14 1 15 from bit 17 */
        header[0] |= 5u | (0u << 3) | (0u << 8) | (15u << 13);
        for (int k = 3; k < 19; k++) {
            const uint32_t bit = 17 + 3 * k;
            const unsigned long long v = 4ull << (bit & 31);
            header[bit >> 5] |= (uint32_t)v;
            if (v >> 32) header[(bit >> 5) + 1] |= (uint32_t)(v >> 32);
        }
        for (int s = 0; s < ENC_NSYM + 1; s++) {
            /* the four-bit code of length value v is v itself, sent MSB first */
            const uint32_t v = s < ENC_NSYM ? len[s] : 0u;
            const uint32_t bit = 17 + 57 + 4 * s;
            const unsigned long long w = (unsigned long long)(__brev(v) >> 28) << (bit & 31);
            header[bit >> 5] |= (uint32_t)w;
            if (w >> 32) header[(bit >> 5) + 1] |= (uint32_t)(w >> 32);
        }
    }
    __syncthreads();
    StreamCode &out = codes[sidx];
    for (int s = tid; s < ENC_NSYM; s += ENC_THREADS) { out.code[s] = code[s]; out.len[s] = len[s]; }
    for (int i = tid; i < (ENC_HDR_BITS + 31) / 32; i += ENC_THREADS) out.header[i] = header[i];
}

struct __align__(16) DeflateSmem {
    uint8_t  in[ENC_CHUNK];
    uint32_t out[ENC_CHUNK / 4 + 8];
    uint32_t crctab[256];
    uint16_t code[ENC_NSYM + 3];
    uint8_t  len[ENC_NSYM + 3];
    uint32_t warpsum[ENC_THREADS / 32];
    uint32_t crcw[ENC_THREADS / 32];
    int stream;
};

__global__ void __launch_bounds__(ENC_THREADS)
k_enc_deflate(const DevStream *__restrict__ streams, int nstreams, const uint32_t *__restrict__ nblocks,
              const uint32_t *__restrict__ crctab, const StreamCode *__restrict__ codes,
              uint8_t *__restrict__ slots, uint32_t *__restrict__ sizes)
{
    extern __shared__ __align__(16) unsigned char deflate_smem[];
    DeflateSmem &sm = *reinterpret_cast<DeflateSmem *>(deflate_smem);
    const uint32_t b = blockIdx.x;
    if (b >= *nblocks) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) sm.stream = stream_of_block(streams, nstreams, b);
    sm.crctab[tid] = crctab[tid];
    __syncthreads();
    const DevStream st = streams[sm.stream];
    const StreamCode &sc = codes[sm.stream];
    const uint64_t off = (uint64_t)(b - st.first_block) * ENC_CHUNK;
    const uint64_t left = st.nbytes - off;
    const uint32_t L = left < ENC_CHUNK ? (uint32_t)left : ENC_CHUNK;
    const uint8_t *src = st.src + off;
    for (int s = tid; s < ENC_NSYM; s += ENC_THREADS) { sm.code[s] = sc.code[s]; sm.len[s] = sc.len[s]; }
    /* the output starts as the stream's block header, zeros behind it */
    for (int i = tid; i < ENC_CHUNK / 4 + 8; i += ENC_THREADS) sm.out[i] = i < (ENC_HDR_BITS + 31) / 32 ? sc.header[i] : 0u;
    /* stage the piece: 4-byte words where the source allows it */
    if ((reinterpret_cast<uintptr_t>(src) & 3) == 0) {
        const uint32_t nw = L >> 2;
        const uint32_t *s4 = reinterpret_cast<const uint32_t *>(src);
        uint32_t *d4 = reinterpret_cast<uint32_t *>(sm.in);
        for (uint32_t i = tid; i < nw; i += ENC_THREADS) d4[i] = s4[i];
        for (uint32_t i = (nw << 2) + tid; i < L; i += ENC_THREADS) sm.in[i] = src[i];
    } else {
        for (uint32_t i = tid; i < L; i += ENC_THREADS) sm.in[i] = src[i];
    }
    __syncthreads();
    /* slices are aligned to the END of the piece, so that every byte after a slice belongs to whole
     * later slices: thread t covers [L - (256 - t) 128, L - (255 - t) 128) cut to [0, L) */
    const int hi = (int)L - (ENC_THREADS - 1 - tid) * ENC_SLICE;
    int lo = hi - ENC_SLICE;
    const bool holds_first = lo <= 0 && hi > 0;
    if (lo < 0) lo = 0;

    /* ---- CRC of the slice and its size in the Huffman form ---- */
    uint32_t crc = holds_first ? 0xffffffffu : 0u, bits = 0;
    for (int i = lo; i < hi; i++) {
        const uint32_t v = sm.in[i];
        bits += sm.len[v];
        crc = sm.crctab[(crc ^ v) & 0xffu] ^ (crc >> 8);
    }
    /* move the slice's register to the end of the piece: times x^(8 * bytes after it) */
    crc = crc_mulmod(crc, crctab[256 + (ENC_THREADS - 1 - tid)]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) crc ^= __shfl_xor_sync(FULL_MASK, crc, d);
    uint32_t incl = bits;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(FULL_MASK, incl, d); if (lane >= d) incl += o; }
    if (lane == 31) sm.warpsum[warp] = incl;
    if (lane == 0) sm.crcw[warp] = crc;
    __syncthreads();
    uint32_t before = incl - bits, allbits = 0;
    for (int w = 0; w < ENC_THREADS / 32; w++) { if (w < warp) before += sm.warpsum[w]; allbits += sm.warpsum[w]; }
    const uint32_t total_bits = ENC_HDR_BITS + allbits + sm.len[256];
    const uint32_t huff_bytes = (total_bits + 7) >> 3;
    const bool stored = huff_bytes >= L + 5;
    uint32_t crc32 = 0;
    for (int w = 0; w < ENC_THREADS / 32; w++) crc32 ^= sm.crcw[w];
    crc32 = ~crc32;

    uint8_t *memb = slots + (size_t)b * ENC_SLOT + 2;        /* +2: the payload then starts on a word */
    const uint32_t payload = stored ? L + 5 : huff_bytes;
    const uint32_t msize = 18 + payload + 8;
    if (tid < 18) {
        const uint8_t hdr[18] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0,
                                 (uint8_t)((msize - 1) & 0xff), (uint8_t)((msize - 1) >> 8)};
        memb[tid] = hdr[tid];
    }
    if (tid >= 32 && tid < 40) {
        const int k = tid - 32;
        memb[18 + payload + k] = (uint8_t)((k < 4 ? crc32 >> (8 * k) : L >> (8 * (k - 4))) & 0xff);
    }
    if (tid == 0) sizes[b] = msize;

    if (stored) {          /* incompressible piece: one stored block */
        if (tid == 0) {
            memb[18] = 1; memb[19] = (uint8_t)(L & 0xff); memb[20] = (uint8_t)(L >> 8);
            memb[21] = (uint8_t)(~L & 0xff); memb[22] = (uint8_t)((~L >> 8) & 0xff);
        }
        for (uint32_t i = tid; i < L; i += ENC_THREADS) memb[23 + i] = sm.in[i];
        return;
    }
    /* ---- the literals of this thread's slice ---- */
    {
        uint32_t bit = ENC_HDR_BITS + before;
        uint32_t word = bit >> 5;
        unsigned long long acc = 0;
        int nb = (int)(bit & 31);
        auto put = [&](uint32_t sym) {
            acc |= (unsigned long long)sm.code[sym] << nb;
            nb += sm.len[sym];
            if (nb >= 32) { atomicOr(&sm.out[word++], (uint32_t)acc); acc >>= 32; nb -= 32; }
        };
        for (int i = lo; i < hi; i++) put(sm.in[i]);
        if (tid == ENC_THREADS - 1) put(256);
        if (nb) atomicOr(&sm.out[word], (uint32_t)acc);
    }
    __syncthreads();
    {
        uint32_t *d4 = reinterpret_cast<uint32_t *>(memb + 18);
        const uint32_t nw = huff_bytes >> 2;
        for (uint32_t i = tid; i < nw; i += ENC_THREADS) d4[i] = sm.out[i];
        const uint8_t *o8 = reinterpret_cast<const uint8_t *>(sm.out);
        for (uint32_t i = (nw << 2) + tid; i < huff_bytes; i += ENC_THREADS) memb[18 + i] = o8[i];
    }
}

/* exclusive scan of the block sizes (one CTA): final offsets, and where every stream starts */
__global__ void __launch_bounds__(1024)
k_enc_offsets(const uint32_t *__restrict__ sizes, const uint32_t *__restrict__ nblocks,
              const DevStream *__restrict__ streams, int nstreams, uint64_t *__restrict__ offsets,
              uint64_t *__restrict__ streamoff)
{
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_carry;
    const uint32_t nb = *nblocks;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nb; base += 1024) {
        const uint32_t i = base + tid;
        const unsigned long long v = i < nb ? sizes[i] : 0ull;
        unsigned long long incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const unsigned long long o = __shfl_up_sync(FULL_MASK, incl, d); if (lane >= d) incl += o; }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const unsigned long long w = s_warp[lane];
            unsigned long long wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const unsigned long long o = __shfl_up_sync(FULL_MASK, wi, d); if (lane >= d) wi += o; }
            s_warp[lane] = wi - w;
        }
        __syncthreads();
        const unsigned long long excl = s_carry + s_warp[warp] + incl - v;
        if (i < nb) offsets[i] = excl;
        __syncthreads();
        if (tid == 1023) s_carry = excl + v;
        __syncthreads();
    }
    if (tid == 0) offsets[nb] = s_carry;
    __syncthreads();
    for (int s = tid; s <= nstreams; s += 1024) streamoff[s] = offsets[streams[s].first_block];
}

__global__ void __launch_bounds__(256)
k_enc_gather(const uint8_t *__restrict__ slots, const uint32_t *__restrict__ sizes, const uint32_t *__restrict__ nblocks,
             const uint64_t *__restrict__ offsets, uint8_t *__restrict__ out)
{
    const uint32_t b = blockIdx.x;
    if (b >= *nblocks) return;
    const uint8_t *src = slots + (size_t)b * ENC_SLOT + 2;
    uint8_t *dst = out + offsets[b];
    const uint32_t n = sizes[b];
    /* destination words, sources read as bytes through the two words they straddle */
    const uint32_t head = (uint32_t)((4 - (reinterpret_cast<uintptr_t>(dst) & 3)) & 3);
    for (uint32_t i = threadIdx.x; i < (head < n ? head : n); i += 256) dst[i] = src[i];
    if (n > head) {
        const uint32_t nw = (n - head) >> 2;
        uint32_t *d4 = reinterpret_cast<uint32_t *>(dst + head);
        const uint8_t *s = src + head;                 /* src is 2 mod 4 */
        for (uint32_t i = threadIdx.x; i < nw; i += 256) {
            const uint8_t *q = s + 4 * i;
            d4[i] = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24);
        }
        for (uint32_t i = head + (nw << 2) + threadIdx.x; i < n; i += 256) dst[i] = src[i];
    }
}

/* ---- host side ---------------------------------------------------------------------------------- */
template <typename T>
static int enc_ensure(T **ptr, size_t *cap, size_t need)
{
    if (*cap >= need && *ptr) return 0;
    if (*ptr) cudaFree(*ptr);
    *ptr = NULL; *cap = 0;
    TSK_CUDA(cudaMalloc(ptr, need * sizeof(T)));
    *cap = need;
    return 0;
}

void tsk_encode_free(tsk_ctx *ctx)
{
    EncState *e = ctx->enc;
    if (!e) return;
    void *ptrs[] = {e->d_samples, e->d_lineoff, e->d_ctasum, e->d_texttot, e->d_text, e->d_liverank, e->d_livecta,
                    e->d_livetot, e->d_live, e->d_streams, e->d_nblocks, e->d_slots, e->d_sizes, e->d_offsets,
                    e->d_streamoff, e->d_final, e->d_crctab, e->d_hist, e->d_codes};
    for (size_t i = 0; i < sizeof(ptrs) / sizeof(ptrs[0]); i++) if (ptrs[i]) cudaFree(ptrs[i]);
    if (e->h_final) cudaFreeHost(e->h_final);
    cudaEventDestroy(e->ev[0]); cudaEventDestroy(e->ev[1]);
    free(e);
    ctx->enc = NULL;
}

extern "C" int tsk_encode_configure(tsk_ctx *ctx, const tsk_encode_layout *layout)
{
    if (!ctx || !layout) { tsk_set_error("null argument"); return -1; }
    TSK_CUDA(cudaSetDevice(ctx->device));
    const int S = ctx->dp.num_samples, T = ctx->dp.total_cycles;
    if (layout->fivep_start < 0 || layout->fivep_length < 0 || layout->fivep_start + layout->fivep_length > T ||
        layout->threep_seqqual_output_length < 0) { tsk_set_error("bad 5' read / output length"); return -1; }
    DevEncSample hs[TSK_MAX_SAMPLES];
    memset(hs, 0, sizeof(hs));
    int umimax = 0;
    for (int s = 0; s < S; s++) {
        const tsk_encode_sample *ls = &layout->samples[s];
        if (ls->umi_count < 0 || ls->umi_count > TSK_ENCODE_MAX_UMI) { tsk_set_error("sample %d: bad umi_count", s); return -1; }
        hs[s].want[0] = ls->want_seqqual != 0; hs[s].want[1] = ls->want_taginfo != 0; hs[s].want[2] = ls->want_signal != 0;
        hs[s].umi_count = ls->umi_count;
        for (int u = 0; u < ls->umi_count; u++) {
            const int ln = ls->umi_length[u] > 0 ? ls->umi_length[u] : 0;
            if (ln && (ls->umi_start[u] < 0 || ls->umi_start[u] + ln > T)) { tsk_set_error("sample %d: UMI range outside the read", s); return -1; }
            hs[s].umi_start[u] = ls->umi_start[u]; hs[s].umi_length[u] = ls->umi_length[u];
            hs[s].umi_total += ln;
        }
        if (hs[s].umi_total > umimax) umimax = hs[s].umi_total;
    }
    if (ctx->enc && memcmp(&ctx->enc->layout, layout, sizeof(*layout)) == 0) return 0;    /* already set up for it */
    if (ctx->enc) tsk_encode_free(ctx);
    EncState *e = (EncState *)calloc(1, sizeof(EncState));
    if (!e) { tsk_set_error("out of memory"); return -1; }
    ctx->enc = e;
    e->layout = *layout;
    int out3 = layout->threep_seqqual_output_length;
    if (out3 > ctx->dp.threep_length) out3 = ctx->dp.threep_length;
    e->tagmax = 10 + 1 + 5 + 1 + 6 + 1 + ctx->dp.max_mods + 1 + umimax + 1;
    e->sqmax = 10 + 1 + 2 * (layout->fivep_length + 1) + 2 * (out3 + 1);
    TSK_CUDA(cudaEventCreate(&e->ev[0]));
    TSK_CUDA(cudaEventCreate(&e->ev[1]));
    TSK_CUDA(cudaMalloc(&e->d_samples, sizeof(DevEncSample) * S));
    TSK_CUDA(cudaMemcpy(e->d_samples, hs, sizeof(DevEncSample) * S, cudaMemcpyHostToDevice));
    TSK_CUDA(cudaMalloc(&e->d_texttot, sizeof(uint32_t) * 4 * TSK_MAX_SAMPLES));    /* totals | bases */
    TSK_CUDA(cudaMalloc(&e->d_livetot, sizeof(uint32_t)));
    TSK_CUDA(cudaMalloc(&e->d_streams, sizeof(DevStream) * (3 * TSK_MAX_SAMPLES + 1)));
    TSK_CUDA(cudaMalloc(&e->d_nblocks, sizeof(uint32_t)));
    TSK_CUDA(cudaMalloc(&e->d_streamoff, sizeof(uint64_t) * (3 * TSK_MAX_SAMPLES + 1)));
    TSK_CUDA(cudaMalloc(&e->d_hist, sizeof(uint32_t) * 256 * 3 * TSK_MAX_SAMPLES));
    TSK_CUDA(cudaMalloc((void **)&e->d_codes, enc_codes_bytes(3 * TSK_MAX_SAMPLES)));
    /* CRC tables: the byte table, and x^(8 * 128 * m) mod p for the slices' distances to the end */
    uint32_t tab[512];
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c & 1u) ? (c >> 1) ^ CRC_POLY : c >> 1;
        tab[i] = c;
    }
    uint32_t x1024 = 0x40000000u;                      /* x^1 */
    for (int k = 0; k < 10; k++) x1024 = crc_mulmod(x1024, x1024);       /* x^(2^10) = x^(8 * 128) */
    tab[256] = 0x80000000u;                            /* x^0 */
    for (int m = 1; m < 256; m++) tab[256 + m] = crc_mulmod(tab[256 + m - 1], x1024);
    TSK_CUDA(cudaMalloc(&e->d_crctab, sizeof(tab)));
    TSK_CUDA(cudaMemcpy(e->d_crctab, tab, sizeof(tab), cudaMemcpyHostToDevice));
    TSK_CUDA(cudaFuncSetAttribute(k_enc_deflate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DeflateSmem)));
    return 0;
}

/* The page-locked host buffer of the encoded streams, sized for a block of `nclusters` clusters, made NOW
 * (e.g. on a start-up thread, beside the first tile's upload and kernels) instead of inside the first
 * tsk_tile_encode(): page-locking ~0.8 GB is the largest first-use cost of a process (0.25 s).  Touches
 * only the encoder's own state; must not run concurrently with tsk_encode_configure / tsk_tile_encode of
 * the same context.  The size is an estimate (0.45 of the uncompressed bound; the streams deflate to
 * ~0.36 of it); tsk_tile_encode() still grows the buffer when a tile needs more. */
extern "C" int tsk_encode_reserve(tsk_ctx *ctx, uint32_t nclusters)
{
    if (!ctx || !ctx->enc) { tsk_set_error("tsk_encode_configure() has not been called"); return -1; }
    TSK_CUDA(cudaSetDevice(ctx->device));
    EncState *e = ctx->enc;
    const double bound = (double)nclusters * ((double)(e->tagmax + e->sqmax) + 4.0 * (2 + ctx->maxdump));
    const size_t want = (size_t)(0.45 * bound) + (1u << 20);
    if (want > e->h_final_cap) {
        if (e->h_final) cudaFreeHost(e->h_final);
        e->h_final = NULL; e->h_final_cap = 0;
        TSK_CUDA(cudaHostAlloc((void **)&e->h_final, want, cudaHostAllocDefault));
        e->h_final_cap = want;
    }
    return 0;
}

static int enc_buffers(tsk_ctx *ctx)
{
    EncState *e = ctx->enc;
    const uint32_t cap = ctx->cap_clusters;
    if (e->cap_clusters == cap && e->d_lineoff) return 0;
    const int S = ctx->dp.num_samples;
    void *old[] = {e->d_lineoff, e->d_ctasum, e->d_liverank};
    for (size_t i = 0; i < 3; i++) if (old[i]) cudaFree(old[i]);
    e->d_lineoff = NULL; e->d_ctasum = NULL; e->d_liverank = NULL;
    const size_t nblk = (cap + ENC_FMT_CL - 1) / ENC_FMT_CL;
    TSK_CUDA(cudaMalloc(&e->d_lineoff, sizeof(uint2) * (size_t)cap));
    TSK_CUDA(cudaMalloc(&e->d_ctasum, sizeof(uint32_t) * nblk * 2 * S));
    TSK_CUDA(cudaMalloc(&e->d_liverank, sizeof(uint32_t) * ((size_t)cap + 1)));
    if (enc_ensure(&e->d_livecta, &e->livecta_cap, (size_t)cap / 1024 + 2) != 0) return -1;
    const size_t textneed = (size_t)cap * (size_t)(e->tagmax + e->sqmax) + 32 * (size_t)S + 64;
    if (textneed >= 0xffffffffull) { tsk_set_error("block too large for device-side text (4 GB per block)"); return -1; }
    if (enc_ensure(&e->d_text, &e->text_cap, textneed) != 0) return -1;
    if (enc_ensure(&e->d_live, &e->live_cap_words, ctx->rec_cap_words) != 0) return -1;
    const size_t maxbytes = textneed + ctx->rec_cap_words * 4;
    const size_t nb = maxbytes / ENC_CHUNK + 3 * (size_t)S + 8;
    if (nb > e->nblk_cap) {
        void *o2[] = {e->d_slots, e->d_sizes, e->d_offsets, e->d_final};
        for (size_t i = 0; i < 4; i++) if (o2[i]) cudaFree(o2[i]);
        e->d_slots = NULL; e->d_sizes = NULL; e->d_offsets = NULL; e->d_final = NULL; e->nblk_cap = 0;
        TSK_CUDA(cudaMalloc(&e->d_slots, nb * ENC_SLOT));
        TSK_CUDA(cudaMalloc(&e->d_sizes, sizeof(uint32_t) * nb));
        TSK_CUDA(cudaMalloc(&e->d_offsets, sizeof(uint64_t) * (nb + 1)));
        e->final_cap = nb * ENC_SLOT;
        TSK_CUDA(cudaMalloc(&e->d_final, e->final_cap));
        e->nblk_cap = nb;
    }
    e->cap_clusters = cap;
    return 0;
}

extern "C" int tsk_tile_encode(tsk_ctx *ctx)
{
    if (!ctx || !ctx->enc) { tsk_set_error("tsk_encode_configure() has not been called"); return -1; }
    if (!ctx->processed) { tsk_set_error("tsk_tile_process() has not run on this tile"); return -1; }
    TSK_CUDA(cudaSetDevice(ctx->device));
    TskRange nvtx("tsk:encode_outputs");
    EncState *e = ctx->enc;
    e->have = 0;
    if (enc_buffers(ctx) != 0) return -1;
    const DevParams &P = ctx->dp;
    const TileView &tv = ctx->tile;
    const int S = P.num_samples;
    const uint32_t N = tv.nclusters;
    cudaStream_t st = ctx->stream;
    const int fl = e->layout.fivep_length, f0 = e->layout.fivep_start;
    int out3 = e->layout.threep_seqqual_output_length;
    if (out3 > P.threep_length) out3 = P.threep_length;
    const unsigned nblk = (N + ENC_FMT_CL - 1) / ENC_FMT_CL;
    uint32_t *textbase = e->d_texttot + 2 * TSK_MAX_SAMPLES;
    TSK_CUDA(cudaEventRecord(e->ev[0], st));
    if (N > 0) {
        k_enc_lengths<<<nblk, ENC_FMT_CL, 0, st>>>(P, tv, ctx->d_calls, e->d_samples, fl, out3, e->d_lineoff, e->d_ctasum);
        k_enc_colscan<<<2 * S, 1024, 0, st>>>(e->d_ctasum, e->d_texttot, (int)nblk, 2 * S);
        ctx->launches += 2;
    } else {
        TSK_CUDA(cudaMemsetAsync(e->d_texttot, 0, sizeof(uint32_t) * 2 * S, st));
    }
    /* record slots: at most one per cluster */
    const unsigned nlcta = N / 1024 + 1;
    TSK_CUDA(cudaMemsetAsync(e->d_livecta, 0, sizeof(uint32_t) * (nlcta + 1), st));
    k_enc_live_count<<<nlcta, 1024, 0, st>>>(ctx->d_records, ctx->d_totals, ctx->d_recbase, ctx->d_samples, S,
                                             e->d_liverank, e->d_livecta);
    k_enc_colscan<<<1, 1024, 0, st>>>(e->d_livecta, e->d_livetot, (int)nlcta, 1);
    k_enc_live_ranks<<<148 * 4, 256, 0, st>>>(ctx->d_totals, S, e->d_liverank, e->d_livecta, e->d_livetot);
    k_enc_live_place<<<148 * 8, 256, 0, st>>>(ctx->d_records, ctx->d_totals, ctx->d_recbase, ctx->d_samples, S,
                                              e->d_liverank, e->d_live);
    k_enc_streams<<<1, 32, 0, st>>>(e->d_samples, ctx->d_samples, S, e->d_texttot, e->d_text, ctx->d_totals,
                                    ctx->d_recbase, ctx->d_records, e->d_liverank, e->d_live, e->d_streams,
                                    e->d_nblocks, textbase);
    ctx->launches += 5;
    if (N > 0) {
        const int tpitch = (P.total_cycles + 3) / 4 * 4 + 4;
        const size_t smem = (size_t)ENC_FMT_CL * tpitch + sizeof(DevEncSample) * S;
        if (smem > e->layout_smem) {
            TSK_CUDA(cudaFuncSetAttribute(k_enc_format, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            e->layout_smem = smem;
        }
        k_enc_format<<<nblk, ENC_FMT_CL, smem, st>>>(P, tv, ctx->d_calls, e->d_samples, f0, fl, out3, tpitch,
                                                      e->d_lineoff, e->d_ctasum, textbase, e->d_text);
        ctx->launches++;
    }
    /* the number of pieces is only known on the device: launch for the most there can be, the
     * surplus CTAs return at once */
    const size_t textbytes = (size_t)N * (size_t)(e->tagmax + e->sqmax) + 32 * (size_t)S;
    size_t recbytes = (size_t)N * (size_t)(2 + ctx->maxdump) * 4;
    size_t nbmax = (textbytes + recbytes) / ENC_CHUNK + 3 * (size_t)S + 4;
    if (nbmax > e->nblk_cap) nbmax = e->nblk_cap;
    TSK_CUDA(cudaMemsetAsync(e->d_hist, 0, sizeof(uint32_t) * 256 * 3 * S, st));
    k_enc_hist<<<(unsigned)nbmax, ENC_THREADS, 0, st>>>(e->d_streams, 3 * S, e->d_nblocks, e->d_hist);
    k_enc_codes<<<3 * S, ENC_THREADS, 0, st>>>(e->d_streams, e->d_hist, e->d_codes);
    k_enc_deflate<<<(unsigned)nbmax, ENC_THREADS, sizeof(DeflateSmem), st>>>(e->d_streams, 3 * S, e->d_nblocks, e->d_crctab,
                                                                            e->d_codes, e->d_slots, e->d_sizes);
    ctx->launches += 2;
    k_enc_offsets<<<1, 1024, 0, st>>>(e->d_sizes, e->d_nblocks, e->d_streams, 3 * S, e->d_offsets, e->d_streamoff);
    k_enc_gather<<<(unsigned)nbmax, 256, 0, st>>>(e->d_slots, e->d_sizes, e->d_nblocks, e->d_offsets, e->d_final);
    ctx->launches += 3;
    TSK_CUDA(cudaEventRecord(e->ev[1], st));
    TSK_CUDA(cudaGetLastError());
    TSK_CUDA(cudaMemcpyAsync(e->h_streamoff, e->d_streamoff, sizeof(uint64_t) * (3 * S + 1), cudaMemcpyDeviceToHost, st));
    TSK_CUDA(cudaStreamSynchronize(st));
    const size_t total = (size_t)e->h_streamoff[3 * S];
    if (total > e->h_final_cap) {
        if (e->h_final) cudaFreeHost(e->h_final);
        e->h_final = NULL; e->h_final_cap = 0;
        const size_t want = total + total / 4 + (1u << 20);
        TSK_CUDA(cudaHostAlloc((void **)&e->h_final, want, cudaHostAllocDefault));
        e->h_final_cap = want;
    }
    if (total) TSK_CUDA(cudaMemcpyAsync(e->h_final, e->d_final, total, cudaMemcpyDeviceToHost, st));
    TSK_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e->ev[0], e->ev[1]);
    e->last_ms = ms;
    e->have = 1;
    return 0;
}

extern "C" int tsk_tile_get_encoded(tsk_ctx *ctx, int sample, int kind, const uint8_t **data, size_t *nbytes)
{
    if (!ctx || !ctx->enc || !ctx->enc->have) { tsk_set_error("tsk_tile_encode() has not run on this tile"); return -1; }
    if (sample < 0 || sample >= ctx->dp.num_samples || kind < 0 || kind > 2 || !data || !nbytes) { tsk_set_error("bad argument"); return -1; }
    const EncState *e = ctx->enc;
    const int j = 3 * sample + kind;
    *data = e->h_final + e->h_streamoff[j];
    *nbytes = (size_t)(e->h_streamoff[j + 1] - e->h_streamoff[j]);
    return 0;
}

extern "C" int tsk_encode_elapsed_ms(tsk_ctx *ctx, double *ms)
{
    if (!ctx || !ctx->enc || !ms) { tsk_set_error("bad argument"); return -1; }
    *ms = ctx->enc->last_ms;
    return 0;
}

/* The 28-byte empty block htslib appends to a BGZF file (and looks for when it opens one). */
extern "C" int tsk_bgzf_eof_block(const uint8_t **data, size_t *nbytes)
{
    static const uint8_t eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0,
                                    0, 0, 0, 0, 0, 0, 0, 0};
    if (!data || !nbytes) { tsk_set_error("null argument"); return -1; }
    *data = eof; *nbytes = sizeof(eof);
    return 0;
}

/* Self-test hook: `n` arbitrary host bytes as ONE stream through the deflate, offset and gather
 * kernels; the BGZF blocks come back in out[0 .. *outn).  (The tile path feeds the same kernels
 * with the streams it formats.) */
__global__ void k_enc_one_stream(const uint8_t *src, uint64_t n, DevStream *streams, uint32_t *nblocks)
{
    if (threadIdx.x || blockIdx.x) return;
    DevStream st;
    st.src = src; st.nbytes = n; st.first_block = 0;
    st.nblocks = (uint32_t)((n + ENC_CHUNK - 1) / ENC_CHUNK);
    streams[0] = st;
    DevStream end;
    end.src = nullptr; end.nbytes = 0; end.first_block = st.nblocks; end.nblocks = 0;
    streams[1] = end;
    *nblocks = st.nblocks;
}

extern "C" int tsk_selftest_deflate(tsk_ctx *ctx, const uint8_t *data, size_t n, uint8_t *out, size_t capacity, size_t *outn)
{
    if (!ctx || (n && !data) || !out || !outn) { tsk_set_error("null argument"); return -1; }
    TSK_CUDA(cudaSetDevice(ctx->device));
    const size_t nb = (n + ENC_CHUNK - 1) / ENC_CHUNK;
    uint8_t *d_in = NULL, *d_slots = NULL, *d_final = NULL;
    uint32_t *d_sizes = NULL, *d_nblocks = NULL, *d_crctab = NULL, *d_hist = NULL;
    StreamCode *d_codes = NULL;
    uint64_t *d_offsets = NULL, *d_streamoff = NULL, h_off[2] = {0, 0};
    DevStream *d_streams = NULL;
    int rc = -1;
    cudaStream_t st = ctx->stream;
    uint32_t tab[512];
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c & 1u) ? (c >> 1) ^ CRC_POLY : c >> 1;
        tab[i] = c;
    }
    uint32_t x1024 = 0x40000000u;
    for (int k = 0; k < 10; k++) x1024 = crc_mulmod(x1024, x1024);
    tab[256] = 0x80000000u;
    for (int m = 1; m < 256; m++) tab[256 + m] = crc_mulmod(tab[256 + m - 1], x1024);
    do {
        if (cudaMalloc(&d_in, n + 16) != cudaSuccess || cudaMalloc(&d_slots, (nb + 1) * ENC_SLOT) != cudaSuccess ||
            cudaMalloc(&d_final, (nb + 1) * ENC_SLOT) != cudaSuccess || cudaMalloc(&d_sizes, sizeof(uint32_t) * (nb + 1)) != cudaSuccess ||
            cudaMalloc(&d_nblocks, sizeof(uint32_t)) != cudaSuccess || cudaMalloc(&d_crctab, sizeof(tab)) != cudaSuccess ||
            cudaMalloc(&d_offsets, sizeof(uint64_t) * (nb + 2)) != cudaSuccess || cudaMalloc(&d_streamoff, sizeof(uint64_t) * 2) != cudaSuccess ||
            cudaMalloc(&d_streams, sizeof(DevStream) * 2) != cudaSuccess || cudaMalloc(&d_hist, sizeof(uint32_t) * 256) != cudaSuccess ||
            cudaMalloc((void **)&d_codes, sizeof(StreamCode)) != cudaSuccess) { tsk_set_error("cudaMalloc failed"); break; }
        if (cudaFuncSetAttribute(k_enc_deflate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DeflateSmem)) != cudaSuccess) break;
        if (n) cudaMemcpyAsync(d_in, data, n, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(d_crctab, tab, sizeof(tab), cudaMemcpyHostToDevice, st);
        k_enc_one_stream<<<1, 1, 0, st>>>(d_in, (uint64_t)n, d_streams, d_nblocks);
        cudaMemsetAsync(d_hist, 0, sizeof(uint32_t) * 256, st);
        k_enc_hist<<<(unsigned)(nb + 1), ENC_THREADS, 0, st>>>(d_streams, 1, d_nblocks, d_hist);
        k_enc_codes<<<1, ENC_THREADS, 0, st>>>(d_streams, d_hist, d_codes);
        k_enc_deflate<<<(unsigned)(nb + 1), ENC_THREADS, sizeof(DeflateSmem), st>>>(d_streams, 1, d_nblocks, d_crctab, d_codes, d_slots, d_sizes);
        k_enc_offsets<<<1, 1024, 0, st>>>(d_sizes, d_nblocks, d_streams, 1, d_offsets, d_streamoff);
        k_enc_gather<<<(unsigned)(nb + 1), 256, 0, st>>>(d_slots, d_sizes, d_nblocks, d_offsets, d_final);
        ctx->launches += 6;
        cudaMemcpyAsync(h_off, d_streamoff, sizeof(h_off), cudaMemcpyDeviceToHost, st);
        if (cudaStreamSynchronize(st) != cudaSuccess) { tsk_set_error("deflate self-test: %s", cudaGetErrorString(cudaGetLastError())); break; }
        if (h_off[1] > capacity) { tsk_set_error("output buffer too small"); break; }
        if (h_off[1] && cudaMemcpy(out, d_final, (size_t)h_off[1], cudaMemcpyDeviceToHost) != cudaSuccess) break;
        *outn = (size_t)h_off[1];
        rc = 0;
    } while (0);
    void *ptrs[] = {d_in, d_slots, d_final, d_sizes, d_nblocks, d_crctab, d_offsets, d_streamoff, d_streams, d_hist, d_codes};
    for (size_t i = 0; i < sizeof(ptrs) / sizeof(ptrs[0]); i++) if (ptrs[i]) cudaFree(ptrs[i]);
    return rc;
}

/* ---- the BGZF deflater on its own: any host bytes -> BGZF blocks (tailseq-writefastq's output side) ---- */
struct tsk_bgzf {
    int device;
    cudaStream_t stream;
    size_t cap;                       /* input bytes the buffers hold */
    uint8_t *d_in, *d_slots, *d_final, *h_final;
    uint32_t *d_sizes, *d_nblocks, *d_crctab, *d_hist;
    uint64_t *d_offsets, *d_streamoff;
    DevStream *d_streams;
    StreamCode *d_codes;
    uint64_t launches;
};

static void bgzf_release(tsk_bgzf *z)
{
    void *ptrs[] = {z->d_in, z->d_slots, z->d_final, z->d_sizes, z->d_nblocks, z->d_crctab, z->d_hist, z->d_offsets,
                    z->d_streamoff, z->d_streams, z->d_codes};
    for (size_t i = 0; i < sizeof(ptrs) / sizeof(ptrs[0]); i++) if (ptrs[i]) cudaFree(ptrs[i]);
    if (z->h_final) cudaFreeHost(z->h_final);
    z->d_in = z->d_slots = z->d_final = z->h_final = NULL;
    z->d_sizes = z->d_nblocks = z->d_crctab = z->d_hist = NULL;
    z->d_offsets = z->d_streamoff = NULL; z->d_streams = NULL; z->d_codes = NULL;
}

static int bgzf_reserve(tsk_bgzf *z, size_t n)
{
    if (n <= z->cap && z->d_in) return 0;
    bgzf_release(z);
    const size_t cap = n + n / 4 + (1u << 20);
    const size_t nb = cap / ENC_CHUNK + 2;
    uint32_t tab[512];
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c & 1u) ? (c >> 1) ^ CRC_POLY : c >> 1;
        tab[i] = c;
    }
    uint32_t x1024 = 0x40000000u;
    for (int k = 0; k < 10; k++) x1024 = crc_mulmod(x1024, x1024);
    tab[256] = 0x80000000u;
    for (int m = 1; m < 256; m++) tab[256 + m] = crc_mulmod(tab[256 + m - 1], x1024);
    TSK_CUDA(cudaMalloc(&z->d_in, cap + 16));
    TSK_CUDA(cudaMalloc(&z->d_slots, nb * ENC_SLOT));
    TSK_CUDA(cudaMalloc(&z->d_final, nb * ENC_SLOT));
    TSK_CUDA(cudaHostAlloc((void **)&z->h_final, nb * ENC_SLOT, cudaHostAllocDefault));
    TSK_CUDA(cudaMalloc(&z->d_sizes, sizeof(uint32_t) * nb));
    TSK_CUDA(cudaMalloc(&z->d_nblocks, sizeof(uint32_t)));
    TSK_CUDA(cudaMalloc(&z->d_crctab, sizeof(tab)));
    TSK_CUDA(cudaMalloc(&z->d_hist, sizeof(uint32_t) * 256));
    TSK_CUDA(cudaMalloc(&z->d_offsets, sizeof(uint64_t) * (nb + 1)));
    TSK_CUDA(cudaMalloc(&z->d_streamoff, sizeof(uint64_t) * 2));
    TSK_CUDA(cudaMalloc(&z->d_streams, sizeof(DevStream) * 2));
    TSK_CUDA(cudaMalloc((void **)&z->d_codes, sizeof(StreamCode)));
    TSK_CUDA(cudaMemcpy(z->d_crctab, tab, sizeof(tab), cudaMemcpyHostToDevice));
    TSK_CUDA(cudaFuncSetAttribute(k_enc_deflate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DeflateSmem)));
    z->cap = cap;
    return 0;
}

extern "C" int tsk_bgzf_create(tsk_bgzf **out, int device)
{
    if (!out) { tsk_set_error("null argument"); return -1; }
    *out = NULL;
    if (tsk_probe(device) < 0) return -1;
    TSK_CUDA(cudaSetDevice(device));
    tsk_bgzf *z = (tsk_bgzf *)calloc(1, sizeof(tsk_bgzf));
    if (!z) { tsk_set_error("out of memory"); return -1; }
    z->device = device;
    TSK_CUDA(cudaStreamCreateWithFlags(&z->stream, cudaStreamNonBlocking));
    *out = z;
    return 0;
}

extern "C" void tsk_bgzf_destroy(tsk_bgzf *z)
{
    if (!z) return;
    cudaSetDevice(z->device);
    bgzf_release(z);
    if (z->stream) cudaStreamDestroy(z->stream);
    free(z);
}

extern "C" int tsk_bgzf_deflate(tsk_bgzf *z, const void *data, size_t n, const uint8_t **out, size_t *outn)
{
    if (!z || (n && !data) || !out || !outn) { tsk_set_error("null argument"); return -1; }
    TSK_CUDA(cudaSetDevice(z->device));
    *out = NULL; *outn = 0;
    if (n == 0) return 0;
    if (bgzf_reserve(z, n) != 0) return -1;
    const size_t nb = (n + ENC_CHUNK - 1) / ENC_CHUNK;
    cudaStream_t st = z->stream;
    uint64_t h_off[2] = {0, 0};
    TSK_CUDA(cudaMemcpyAsync(z->d_in, data, n, cudaMemcpyHostToDevice, st));
    TSK_CUDA(cudaMemsetAsync(z->d_hist, 0, sizeof(uint32_t) * 256, st));
    k_enc_one_stream<<<1, 1, 0, st>>>(z->d_in, (uint64_t)n, z->d_streams, z->d_nblocks);
    k_enc_hist<<<(unsigned)nb, ENC_THREADS, 0, st>>>(z->d_streams, 1, z->d_nblocks, z->d_hist);
    k_enc_codes<<<1, ENC_THREADS, 0, st>>>(z->d_streams, z->d_hist, z->d_codes);
    k_enc_deflate<<<(unsigned)nb, ENC_THREADS, sizeof(DeflateSmem), st>>>(z->d_streams, 1, z->d_nblocks, z->d_crctab, z->d_codes,
                                                                         z->d_slots, z->d_sizes);
    k_enc_offsets<<<1, 1024, 0, st>>>(z->d_sizes, z->d_nblocks, z->d_streams, 1, z->d_offsets, z->d_streamoff);
    k_enc_gather<<<(unsigned)nb, 256, 0, st>>>(z->d_slots, z->d_sizes, z->d_nblocks, z->d_offsets, z->d_final);
    z->launches += 6;
    TSK_CUDA(cudaGetLastError());
    TSK_CUDA(cudaMemcpyAsync(h_off, z->d_streamoff, sizeof(h_off), cudaMemcpyDeviceToHost, st));
    TSK_CUDA(cudaStreamSynchronize(st));
    if (h_off[1]) TSK_CUDA(cudaMemcpyAsync(z->h_final, z->d_final, (size_t)h_off[1], cudaMemcpyDeviceToHost, st));
    TSK_CUDA(cudaStreamSynchronize(st));
    *out = z->h_final; *outn = (size_t)h_off[1];
    return 0;
}

extern "C" uint64_t tsk_bgzf_launch_count(tsk_bgzf *z) { return z ? z->launches : 0; }
