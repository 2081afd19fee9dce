/*
 * tsk_ingest.cu -- a tile handed over cycle file by cycle file.
 *
 * The reference fills whole-tile host arrays first (load_intensities_and_basecalls(),
 * tailseq-import.c:203-228: per cycle 4 x (fseek + fread) + AoS transpose, cifreader.c:149-165;
 * gzread of every base-call file, bclreader.c:95-119) and only then starts to compute.  Here every
 * cycle file goes to its rows of the tile's HBM copy as soon as the host has read it
 * (tsk_tile_put_cif / tsk_tile_put_bcl: asynchronous copies on the copy stream, issued by the
 * reader threads themselves, so file reads and transfers overlap and the previous tile's kernels
 * run underneath), and a gzip-compressed base-call file (.bcl.gz, bclreader.c:60-72) travels
 * COMPRESSED: k_inflate_rows decodes the deflate streams on the device, one warp per file.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <condition_variable>
#include "tsk_internal.h"

#define INF_FAST_BITS 9
#define INF_OK        0
#define INF_BAD       1          /* malformed stream */
#define INF_SHORT     2          /* input ended early */
#define INF_SIZE      3          /* inflated size differs from the expected one */

struct InflateRow {
    unsigned long long src_off;          /* into the compressed arena */
    unsigned long long src_bytes;
    unsigned long long scratch_off;      /* into the inflated scratch */
    unsigned long long inflated_size;    /* expected size of the whole stream */
    unsigned long long skip;             /* first byte of the stream that belongs to the row */
    uint8_t *dst;                        /* the row: count bytes */
    uint32_t count;
    uint32_t pad_;
};

/* page-locked staging slots of the file readers (tsk_tile_stage_acquire / _release) */
#define STAGE_SLOTS 24
struct StageSlot {
    uint8_t *p;
    size_t cap;
    cudaEvent_t done;        /* the copy out of the slot, recorded on the copy stream at release */
    int busy, has_event;
};

struct IngestState {
    std::mutex lock;
    std::condition_variable slot_free;
    StageSlot stage[STAGE_SLOTS];
    int open, slot;
    uint32_t nclusters;
    size_t bpitch, cpitch;
    uint8_t *d_gz;        size_t gz_cap, gz_used;
    uint8_t *d_scratch;   size_t scratch_cap, scratch_used;
    InflateRow *h_rows;   int nrows, rows_cap;
    InflateRow *d_rows;   int d_rows_cap;
    int *d_status;        int status_cap;
    uint8_t *h_gz;        size_t h_gz_cap;       /* host staging of the compressed bytes (kept until tsk_tile_end) */
};

__constant__ uint16_t c_len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t  c_len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint16_t c_dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__constant__ uint8_t  c_dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__constant__ uint8_t  c_clen_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

/* A canonical Huffman code for decoding: symbols sorted by (length, value) with counts per length
 * (decoded bit by bit for codes longer than INF_FAST_BITS), and a table over the next INF_FAST_BITS
 * bits of the stream for the shorter ones: entry = length << 9 | symbol, 0 = not a short code. */
struct HuffDec {
    int ncoded;                           /* symbols with a code */
    uint16_t count[16];
    uint16_t symbol[288];
    uint16_t fast[1 << INF_FAST_BITS];
};

struct BitIn {
    const uint8_t *p, *end;
    unsigned long long buf;
    int n;
};

__device__ __forceinline__ void bits_fill(BitIn &b)
{
    while (b.n <= 56 && b.p < b.end) { b.buf |= (unsigned long long)(*b.p++) << b.n; b.n += 8; }
}
__device__ __forceinline__ uint32_t bits_take(BitIn &b, int k)       /* k <= 16; missing bits read as 0 */
{
    if (b.n < k) bits_fill(b);
    const uint32_t v = (uint32_t)b.buf & ((1u << k) - 1u);
    b.buf >>= k; b.n -= k;
    return v;
}

/* builds the decoder of `n` code lengths (lane 0 does the serial part, the warp fills the table);
 * returns 0 for a complete code, 1 for an over-subscribed set, 2 for an incomplete one */
__device__ int huff_build(HuffDec &h, const uint8_t *len, int n, int lane)
{
    __shared__ uint16_t s_code[288];
    __shared__ int s_ok;
    for (int i = lane; i < (1 << INF_FAST_BITS); i += 32) h.fast[i] = 0;
    if (lane == 0) {
        int offs[16], next[16];
        for (int l = 0; l < 16; l++) h.count[l] = 0;
        for (int s = 0; s < n; s++) h.count[len[s]]++;
        int left = 1, ok = 0;
        for (int l = 1; l < 16; l++) { left <<= 1; left -= h.count[l]; if (left < 0) ok = 1; }
        if (!ok && left > 0) ok = 2;
        h.ncoded = n - h.count[0];
        offs[1] = 0;
        for (int l = 1; l < 15; l++) offs[l + 1] = offs[l] + h.count[l];
        for (int s = 0; s < n; s++) if (len[s]) h.symbol[offs[len[s]]++] = (uint16_t)s;
        int code = 0;
        h.count[0] = 0;
        for (int l = 1; l < 16; l++) { code = (code + h.count[l - 1]) << 1; next[l] = code; }
        for (int s = 0; s < n; s++) if (len[s]) s_code[s] = (uint16_t)next[len[s]]++;
        s_ok = ok;
    }
    __syncwarp();
    for (int s = lane; s < n; s += 32) {
        const int l = len[s];
        if (l == 0 || l > INF_FAST_BITS) continue;
        const uint32_t r = __brev((uint32_t)s_code[s]) >> (32 - l);
        for (uint32_t i = r; i < (1u << INF_FAST_BITS); i += 1u << l) h.fast[i] = (uint16_t)((l << 9) | s);
    }
    __syncwarp();
    return s_ok;
}

__device__ __forceinline__ int huff_decode(const HuffDec &h, BitIn &b)
{
    if (b.n < 15) bits_fill(b);
    const uint32_t e = h.fast[(uint32_t)b.buf & ((1u << INF_FAST_BITS) - 1u)];
    if (e) { const int l = e >> 9; b.buf >>= l; b.n -= l; return (int)(e & 511u); }
    int code = 0, first = 0, index = 0;
    unsigned long long bits = b.buf;
    for (int l = 1; l < 16; l++) {
        code |= (int)(bits & 1ull); bits >>= 1;
        const int cnt = h.count[l];
        if (code - cnt < first) { b.buf >>= l; b.n -= l; return h.symbol[index + (code - first)]; }
        index += cnt; first += cnt; first <<= 1; code <<= 1;
    }
    return -1;
}

/* One warp per compressed file: gzip header (RFC 1952), deflate blocks (RFC 1951: stored, fixed,
 * dynamic) decoded by lane 0 into the stream's scratch area, then the warp copies the row's window
 * [skip, skip + count) to its place in the tile. */
__global__ void __launch_bounds__(32)
k_inflate_rows(const InflateRow *__restrict__ rows, const uint8_t *__restrict__ gz, uint8_t *__restrict__ scratch,
               int *__restrict__ status)
{
    __shared__ HuffDec s_lit, s_dist;
    __shared__ uint8_t s_len[320];
    __shared__ int s_state;               /* 0 decoding, 1 done, > 1 error */
    const int lane = threadIdx.x;
    const InflateRow row = rows[blockIdx.x];
    uint8_t *out = scratch + row.scratch_off;
    const unsigned long long cap = row.inflated_size;
    unsigned long long pos = 0;
    BitIn b;
    b.p = gz + row.src_off; b.end = b.p + row.src_bytes; b.buf = 0; b.n = 0;
    int err = INF_OK;
    if (lane == 0) {
        /* ---- gzip member header ---- */
        if (row.src_bytes < 18 || b.p[0] != 0x1f || b.p[1] != 0x8b || b.p[2] != 8) err = INF_BAD;
        else {
            const int flg = b.p[3];
            b.p += 10;
            if (flg & 4) { if (b.p + 2 <= b.end) { const int xlen = b.p[0] | (b.p[1] << 8); b.p += 2 + xlen; } else err = INF_SHORT; }
            if (flg & 8) { while (b.p < b.end && *b.p) b.p++; b.p++; }
            if (flg & 16) { while (b.p < b.end && *b.p) b.p++; b.p++; }
            if (flg & 2) b.p += 2;
            if (b.p >= b.end) err = INF_SHORT;
        }
        s_state = err ? 1 + err : 0;
    }
    __syncwarp();
    while (s_state == 0) {
        /* ---- block header (lane 0), code tables (warp) ---- */
        int final = 0, type = 0, nlen = 0, ndist = 0;
        if (lane == 0) {
            final = (int)bits_take(b, 1);
            type = (int)bits_take(b, 2);
            if (type == 0) {                                   /* stored */
                b.buf >>= b.n & 7; b.n -= b.n & 7;
                const uint32_t len = bits_take(b, 16), nl = bits_take(b, 16);
                if ((len ^ 0xffffu) != nl) err = INF_BAD;
                else {
                    /* the bit buffer holds whole bytes here: drain it, then copy from the input */
                    uint32_t left = len;
                    while (left && b.n >= 8 && pos < cap) { out[pos++] = (uint8_t)b.buf; b.buf >>= 8; b.n -= 8; left--; }
                    if ((unsigned long long)(b.end - b.p) < left) err = INF_SHORT;
                    else if (pos + left > cap) err = INF_SIZE;
                    else { for (uint32_t i = 0; i < left; i++) out[pos + i] = b.p[i]; pos += left; b.p += left; }
                }
            } else if (type == 1) {                            /* fixed code */
                for (int s = 0; s < 144; s++) s_len[s] = 8;
                for (int s = 144; s < 256; s++) s_len[s] = 9;
                for (int s = 256; s < 280; s++) s_len[s] = 7;
                for (int s = 280; s < 288; s++) s_len[s] = 8;
                for (int s = 0; s < 32; s++) s_len[288 + s] = 5;      /* 30 and 31 never occur in valid data */
                nlen = 288; ndist = 32;
            } else if (type == 2) {                            /* dynamic code: read the code lengths */
                nlen = (int)bits_take(b, 5) + 257; ndist = (int)bits_take(b, 5) + 1;
                const int ncode = (int)bits_take(b, 4) + 4;
                if (nlen > 286 || ndist > 30) err = INF_BAD;
                for (int i = 0; i < 19; i++) s_len[i] = 0;
                for (int i = 0; i < ncode && !err; i++) s_len[c_clen_order[i]] = (uint8_t)bits_take(b, 3);
            } else err = INF_BAD;
            s_state = err ? 1 + err : (type == 0 ? (final ? 1 : 0) : 0);
        }
        type = __shfl_sync(0xffffffffu, type, 0);
        nlen = __shfl_sync(0xffffffffu, nlen, 0);
        ndist = __shfl_sync(0xffffffffu, ndist, 0);
        final = __shfl_sync(0xffffffffu, final, 0);
        __syncwarp();
        if (s_state != 0 || type == 0) continue;
        if (type == 2) {
            /* the code-length code, then the literal/length + distance lengths it describes */
            const int bad = huff_build(s_lit, s_len, 19, lane);
            if (lane == 0) {
                if (bad == 1 || (bad == 2 && s_lit.ncoded != 1)) err = INF_BAD;
                __shared__ uint8_t s_all[320];
                int idx = 0;
                while (idx < nlen + ndist && !err) {
                    int sym = huff_decode(s_lit, b), rep = 0, val = 0;
                    if (sym < 0) { err = INF_BAD; break; }
                    if (sym < 16) { s_all[idx++] = (uint8_t)sym; continue; }
                    if (sym == 16) { if (idx == 0) { err = INF_BAD; break; } val = s_all[idx - 1]; rep = 3 + (int)bits_take(b, 2); }
                    else if (sym == 17) rep = 3 + (int)bits_take(b, 3);
                    else rep = 11 + (int)bits_take(b, 7);
                    if (idx + rep > nlen + ndist) { err = INF_BAD; break; }
                    while (rep--) s_all[idx++] = (uint8_t)val;
                }
                if (!err && s_all[256] == 0) err = INF_BAD;            /* no end-of-block code */
                for (int i = 0; i < nlen + ndist; i++) s_len[i < nlen ? i : 288 + (i - nlen)] = s_all[i];
                for (int i = nlen; i < 288; i++) s_len[i] = 0;
                if (err) s_state = 1 + err;
            }
            __syncwarp();
            if (s_state != 0) continue;
        }
        {
            const int bad1 = huff_build(s_lit, s_len, type == 1 ? 288 : nlen, lane);
            const int bad2 = huff_build(s_dist, s_len + 288, ndist, lane);
            /* incomplete codes are legal only with a single coded symbol; no distance code at all
             * (a block of literals) is legal too */
            if (lane == 0 && (bad1 == 1 || (bad1 == 2 && s_lit.ncoded != 1) || bad2 == 1 || (bad2 == 2 && s_dist.ncoded > 1))) {
                err = INF_BAD; s_state = 1 + err;
            }
            __syncwarp();
            if (s_state != 0) continue;
        }
        /* ---- the block's symbols (lane 0) ---- */
        if (lane == 0) {
            for (;;) {
                const int sym = huff_decode(s_lit, b);
                if (b.n < 0) { err = INF_SHORT; break; }
                if (sym < 0) { err = INF_BAD; break; }
                if (sym < 256) { if (pos >= cap) { err = INF_SIZE; break; } out[pos++] = (uint8_t)sym; continue; }
                if (sym == 256) break;
                if (sym > 285) { err = INF_BAD; break; }
                const int li = sym - 257;
                const uint32_t len = c_len_base[li] + bits_take(b, c_len_extra[li]);
                const int ds = huff_decode(s_dist, b);
                if (ds < 0 || ds > 29) { err = INF_BAD; break; }
                const uint32_t dist = c_dist_base[ds] + bits_take(b, c_dist_extra[ds]);
                if (dist > pos) { err = INF_BAD; break; }
                if (pos + len > cap) { err = INF_SIZE; break; }
                for (uint32_t i = 0; i < len; i++) out[pos + i] = out[pos + i - dist];
                pos += len;
                if (b.n < 0) { err = INF_SHORT; break; }
            }
            s_state = err ? 1 + err : (final ? 1 : 0);
        }
        __syncwarp();
    }
    if (lane == 0) {
        if (!err && s_state == 1 && pos != cap) err = INF_SIZE;
        status[blockIdx.x] = err;
    }
    __syncwarp();
    __threadfence_block();
    /* the row's window of the inflated stream, by the whole warp */
    if (row.skip + row.count <= cap)
        for (uint32_t i = lane; i < row.count; i += 32) row.dst[i] = out[row.skip + i];
}

/* ---- host side -------------------------------------------------------------------------------------- */
static IngestState *ingest_state(tsk_ctx *ctx)
{
    if (!ctx->ingest) ctx->ingest = new IngestState();
    return ctx->ingest;
}

void tsk_ingest_free(tsk_ctx *ctx)
{
    IngestState *g = ctx->ingest;
    if (!g) return;
    if (g->d_gz) cudaFree(g->d_gz);
    if (g->d_scratch) cudaFree(g->d_scratch);
    if (g->d_rows) cudaFree(g->d_rows);
    if (g->d_status) cudaFree(g->d_status);
    free(g->h_rows);
    if (g->h_gz) cudaFreeHost(g->h_gz);
    for (int i = 0; i < STAGE_SLOTS; i++) {
        if (g->stage[i].has_event) cudaEventDestroy(g->stage[i].done);
        if (g->stage[i].p) cudaFreeHost(g->stage[i].p);
    }
    delete g;
    ctx->ingest = NULL;
}

/* A page-locked buffer of at least `bytes` for one cycle file's rows: the reader fills it, passes it to
 * tsk_tile_put_bcl / tsk_tile_put_cif and releases it; the slot goes to the next reader once that copy
 * has completed.  From pageable memory the same copies are staged by the driver at a fifth of the
 * PCIe rate and hold the calling thread meanwhile.  Blocks while all slots are taken. */
extern "C" int tsk_tile_stage_acquire(tsk_ctx *ctx, size_t bytes, void **ptr, int *slot)
{
    if (!ctx || !ptr || !slot) { tsk_set_error("null argument"); return -1; }
    IngestState *g = ingest_state(ctx);
    int s = -1;
    {
        std::unique_lock<std::mutex> guard(g->lock);
        for (;;) {
            s = -1;
            for (int i = 0; i < STAGE_SLOTS; i++) {
                if (g->stage[i].busy) continue;
                /* a slot that is large enough already, else any free one */
                if (s < 0 || (g->stage[s].cap < bytes && g->stage[i].cap >= bytes)) s = i;
            }
            if (s >= 0) break;
            g->slot_free.wait(guard);
        }
        g->stage[s].busy = 1;
    }
    StageSlot &t = g->stage[s];
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e == cudaSuccess && t.has_event) e = cudaEventSynchronize(t.done);
    if (e == cudaSuccess && !t.has_event) {
        e = cudaEventCreateWithFlags(&t.done, cudaEventDisableTiming);
        if (e == cudaSuccess) t.has_event = 1;
    }
    if (e == cudaSuccess && t.cap < bytes) {
        if (t.p) cudaFreeHost(t.p);
        t.p = NULL; t.cap = 0;
        e = cudaHostAlloc((void **)&t.p, bytes, cudaHostAllocDefault);
        if (e == cudaSuccess) t.cap = bytes;
    }
    if (e != cudaSuccess) {
        tsk_set_error("staging slot: %s", cudaGetErrorString(e));
        std::lock_guard<std::mutex> guard(g->lock);
        t.busy = 0;
        g->slot_free.notify_one();
        return -1;
    }
    *ptr = t.p; *slot = s;
    return 0;
}

extern "C" int tsk_tile_stage_release(tsk_ctx *ctx, int slot)
{
    IngestState *g = ctx ? ctx->ingest : NULL;
    if (!g || slot < 0 || slot >= STAGE_SLOTS || !g->stage[slot].busy) { tsk_set_error("bad staging slot"); return -1; }
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e == cudaSuccess) e = cudaEventRecord(g->stage[slot].done, ctx->cstream);
    {
        std::lock_guard<std::mutex> guard(g->lock);
        g->stage[slot].busy = 0;
    }
    g->slot_free.notify_one();
    if (e != cudaSuccess) { tsk_set_error("staging slot: %s", cudaGetErrorString(e)); return -1; }
    return 0;
}

extern "C" int tsk_tile_begin(tsk_ctx *ctx, uint32_t nclusters, uint32_t firstclusterno, const float invmatrix[16])
{
    if (!ctx || !invmatrix) { tsk_set_error("null argument"); return -1; }
    IngestState *g = ingest_state(ctx);
    if (g->open) { tsk_set_error("tsk_tile_begin(): the previous tile was not ended"); return -1; }
    size_t bpitch, cpitch;
    int slot;
    if (tsk_upload_slot_acquire(ctx, nclusters, firstclusterno, invmatrix, &slot, &bpitch, &cpitch) != 0) return -1;
    g->open = 1; g->slot = slot; g->nclusters = nclusters; g->bpitch = bpitch; g->cpitch = cpitch;
    g->nrows = 0; g->gz_used = 0; g->scratch_used = 0;
    return 0;
}

extern "C" int tsk_tile_put_bcl(tsk_ctx *ctx, int cycle, const uint8_t *calls)
{
    IngestState *g = ctx ? ctx->ingest : NULL;
    if (!g || !g->open || !calls || cycle < 0 || cycle >= ctx->dp.total_cycles) { tsk_set_error("bad argument / no open tile"); return -1; }
    TSK_CUDA(cudaSetDevice(ctx->device));
    if (g->nclusters)
        TSK_CUDA(cudaMemcpyAsync(ctx->d_bcl_own[g->slot] + (size_t)cycle * g->bpitch, calls, g->nclusters,
                                 cudaMemcpyHostToDevice, ctx->cstream));
    return 0;
}

extern "C" int tsk_tile_put_cif(tsk_ctx *ctx, int cycle3, const int16_t *planes, size_t plane_pitch)
{
    IngestState *g = ctx ? ctx->ingest : NULL;
    if (!g || !g->open || !planes || cycle3 < 0 || cycle3 >= ctx->dp.threep_length || plane_pitch < g->nclusters) {
        tsk_set_error("bad argument / no open tile"); return -1;
    }
    TSK_CUDA(cudaSetDevice(ctx->device));
    if (g->nclusters)
        TSK_CUDA(cudaMemcpy2DAsync(ctx->d_cif_own[g->slot] + (size_t)cycle3 * 4 * g->cpitch, g->cpitch * sizeof(int16_t),
                                   planes, plane_pitch * sizeof(int16_t), (size_t)g->nclusters * sizeof(int16_t), 4,
                                   cudaMemcpyHostToDevice, ctx->cstream));
    return 0;
}

extern "C" int tsk_tile_put_bcl_gz(tsk_ctx *ctx, int cycle, const uint8_t *gz, size_t gzbytes, uint64_t inflated_size, uint64_t skip)
{
    IngestState *g = ctx ? ctx->ingest : NULL;
    if (!g || !g->open || !gz || cycle < 0 || cycle >= ctx->dp.total_cycles) { tsk_set_error("bad argument / no open tile"); return -1; }
    if (skip + g->nclusters > inflated_size) { tsk_set_error("cycle %d: the window lies outside the inflated stream", cycle + 1); return -1; }
    std::lock_guard<std::mutex> guard(g->lock);       /* reader threads call this side by side */
    if (g->nrows == g->rows_cap) {
        const int cap = g->rows_cap ? g->rows_cap * 2 : 512;
        InflateRow *q = (InflateRow *)realloc(g->h_rows, sizeof(InflateRow) * cap);
        if (!q) { tsk_set_error("out of memory"); return -1; }
        g->h_rows = q; g->rows_cap = cap;
    }
    /* the compressed bytes are staged on the host (pinned) until tsk_tile_end() sends them in one copy */
    const size_t at = (g->gz_used + 15) & ~(size_t)15;
    if (at + gzbytes > g->h_gz_cap) {
        const size_t cap = (at + gzbytes) * 2 + (1u << 20);
        uint8_t *q = NULL;
        if (cudaHostAlloc((void **)&q, cap, cudaHostAllocDefault) != cudaSuccess) { tsk_set_error("cudaHostAlloc failed"); return -1; }
        if (g->h_gz) { memcpy(q, g->h_gz, g->gz_used); cudaFreeHost(g->h_gz); }
        g->h_gz = q; g->h_gz_cap = cap;
    }
    memcpy(g->h_gz + at, gz, gzbytes);
    InflateRow &r = g->h_rows[g->nrows++];
    r.src_off = at; r.src_bytes = gzbytes;
    r.scratch_off = (g->scratch_used + 127) & ~(size_t)127;
    r.inflated_size = inflated_size; r.skip = skip;
    r.dst = ctx->d_bcl_own[g->slot] + (size_t)cycle * g->bpitch;
    r.count = g->nclusters; r.pad_ = (uint32_t)cycle;
    g->gz_used = at + gzbytes;
    g->scratch_used = r.scratch_off + inflated_size;
    return 0;
}

extern "C" int tsk_tile_end(tsk_ctx *ctx)
{
    IngestState *g = ctx ? ctx->ingest : NULL;
    if (!g || !g->open) { tsk_set_error("no open tile"); return -1; }
    TSK_CUDA(cudaSetDevice(ctx->device));
    g->open = 0;
    int rc = 0;
    if (g->nrows > 0) {
        TskRange nvtx("tsk:inflate_bcl_gz");
        if (g->gz_used > g->gz_cap) {
            if (g->d_gz) cudaFree(g->d_gz);
            g->d_gz = NULL; g->gz_cap = 0;
            TSK_CUDA(cudaMalloc(&g->d_gz, g->gz_used * 2));
            g->gz_cap = g->gz_used * 2;
        }
        if (g->scratch_used > g->scratch_cap) {
            if (g->d_scratch) cudaFree(g->d_scratch);
            g->d_scratch = NULL; g->scratch_cap = 0;
            TSK_CUDA(cudaMalloc(&g->d_scratch, g->scratch_used + (g->scratch_used >> 2)));
            g->scratch_cap = g->scratch_used + (g->scratch_used >> 2);
        }
        if (g->nrows > g->d_rows_cap) {
            if (g->d_rows) cudaFree(g->d_rows);
            if (g->d_status) cudaFree(g->d_status);
            g->d_rows = NULL; g->d_status = NULL; g->d_rows_cap = 0;
            TSK_CUDA(cudaMalloc(&g->d_rows, sizeof(InflateRow) * g->nrows * 2));
            TSK_CUDA(cudaMalloc(&g->d_status, sizeof(int) * g->nrows * 2));
            g->d_rows_cap = g->nrows * 2;
        }
        TSK_CUDA(cudaMemcpyAsync(g->d_gz, g->h_gz, g->gz_used, cudaMemcpyHostToDevice, ctx->cstream));
        TSK_CUDA(cudaMemcpyAsync(g->d_rows, g->h_rows, sizeof(InflateRow) * g->nrows, cudaMemcpyHostToDevice, ctx->cstream));
        k_inflate_rows<<<g->nrows, 32, 0, ctx->cstream>>>(g->d_rows, g->d_gz, g->d_scratch, g->d_status);
        ctx->launches++;
        TSK_CUDA(cudaGetLastError());
        int *st = (int *)malloc(sizeof(int) * g->nrows);
        if (!st) { tsk_set_error("out of memory"); return -1; }
        TSK_CUDA(cudaMemcpyAsync(st, g->d_status, sizeof(int) * g->nrows, cudaMemcpyDeviceToHost, ctx->cstream));
        TSK_CUDA(cudaStreamSynchronize(ctx->cstream));
        for (int i = 0; i < g->nrows; i++)
            if (st[i] != INF_OK) {
                tsk_set_error("cycle %u: %s", g->h_rows[i].pad_ + 1,
                              st[i] == INF_SIZE ? "the base-call file inflates to another size than its cluster count says"
                                                : st[i] == INF_SHORT ? "Unexpected EOF in a BCL" : "the base-call file is not a valid gzip stream");
                rc = -1;
                break;
            }
        free(st);
    }
    /* the caller's buffers are free again once the row copies have gone */
    TSK_CUDA(cudaStreamSynchronize(ctx->cstream));
    if (rc != 0) return -1;                 /* the tile is dropped: its slot is taken again by the next one */
    return tsk_upload_slot_commit(ctx, g->slot);
}
