/*
 * tsk_dedup.cu -- duplicate collapsing by UMI (reference src/deduplicator/tailseq-dedup-perfect.c)
 * and the UMI sort in front of it (tailseeker/main.py:272-278), on the device.
 *
 * The reference walks its sorted stdin once with a queue per UMI group:
 *   - a record whose tag priority (:87-105) is below the running maximum of its group is
 *     dropped at once, one whose priority exceeds it empties the queue (the queued records are
 *     dropped) and starts a new one, an equal one joins the queue (:168-214);
 *   - at the end of the group the queue holds the records of the highest priority; the mean of
 *     their poly(A) lengths is formed in float, the first record nearest to it represents the
 *     group and carries (int)roundf(mean) (:277-345).
 * Nothing in that depends on the order of evaluation, only on the order of the records, so it is
 * restated with prefix operations over a group:
 *   rex(i)   = max priority before i (exclusive running maximum), gmax = maximum of the group
 *   imm(i)   = prio(i) <  rex(i)      dropped when read            (trace -3, written at once)
 *   brk(i)   = prio(i) >  rex(i)      starts a queue ("epoch"); at most 2047 per group because
 *                                     priorities are < 2048 and strictly increase
 *   kept(i)  = prio(i) == gmax        in the final queue
 *   q(i)     = neither                queued, dropped by the next brk after i
 * and the line number of a dropped record in the trace follows from counts:
 *   imm(i):  #imm before i + #q before the brk that opened i's epoch
 *   q(i):    #imm before the next brk after i + #q before i
 *   kept(i): #dropped in the group + #kept before i.
 * One more thing is order-dependent in the reference: its queue starts with 2048 slots for the
 * whole run and grows by half when a record arrives at a full queue, and tqueue_expand()
 * (:142-166) does not carry the arriving record over, so that record continues as zeros and
 * empty strings (see dd_lost_records()).  Reproduced: the few records that arrive at a queue of
 * exactly one of the possible sizes are listed by the kernel, the host walks them in order.
 *
 * Groups of up to 32 records are finished by one lane (the classes above are 32-bit masks and the
 * counts population counts); larger groups by one CTA in two passes of block scans.
 *
 * Sort: least-significant-digit radix sort, 8 bits a pass, stable (per-warp match_any ranks inside
 * a 256-record tile, tiles of a block in order, block bases from a digit-major histogram scan).
 * Only digits in which the keys differ at all get a pass (30-base UMIs: 12 passes).
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include "tsk_internal.h"

#define FULL 0xffffffffu
#define DD_THREADS 256
#define DD_WARPS 8
#define DD_BIG_THREADS 1024                 /* CTA of the large-group kernel */
#define DD_BIG_WARPS 32
#define DD_SCAN_ITEMS 8                      /* records per thread in the group-number scan */
#define DD_SCAN_TILE (DD_THREADS * DD_SCAN_ITEMS)
#define DD_MAX_EPOCHS 2048

struct tsk_dedup {
    int device;
    uint64_t cap, n, ngroups;
    cudaStream_t stream;
    cudaEvent_t ev[3];
    uint64_t *kin_hi, *kin_lo;               /* as uploaded */
    uint64_t *khi[2], *klo[2];               /* sort ping-pong */
    uint32_t *ord[2];
    int32_t *flags, *polya;
    int32_t *prio, *pa;                      /* in sorted order */
    uint32_t *gid, *gstart;
    int32_t *tval; uint32_t *tslot;
    uint32_t *rep, *ndup, *nkept; int32_t *gpolya;
    uint32_t *biglist;                       /* [0] count, then group numbers */
    uint32_t *cands; uint32_t cand_cap;      /* [0] count, then (position, size index) pairs */
    uint32_t *lost; uint32_t nlost;          /* host: positions whose record the reference loses */
    uint32_t *hist, *rowtot, *partials;
    uint64_t *varbits;                       /* [2] */
    uint32_t nblk;
    int cur;                                 /* which of the ping-pong buffers holds the sorted keys */
    int sorted, done;
    uint64_t launches;
    float ms_sort, ms_collapse;
};

/* ---- keys ----------------------------------------------------------------------------- */
extern "C" int tsk_dedup_pack_umi(const char *umi, size_t length, uint64_t key[2])
{
    if (length > TSK_DEDUP_MAX_UMI) { tsk_set_error("UMI longer than %d bases", TSK_DEDUP_MAX_UMI); return -1; }
    uint64_t k[2] = {0, 0};
    for (size_t i = 0; i < length; i++) {
        uint64_t c;
        switch (umi[i]) {
        case 'A': c = 1; break; case 'C': c = 2; break; case 'G': c = 3; break;
        case 'N': c = 4; break; case 'T': c = 5; break;
        default: tsk_set_error("UMI character 0x%02x is not one of ACGNT", (unsigned char)umi[i]); return -1;
        }
        const size_t w = i / 21, r = i % 21;
        k[w] |= c << (60 - 3 * r);
    }
    key[0] = k[0]; key[1] = k[1];
    return 0;
}

__device__ __forceinline__ unsigned dd_digit(uint64_t hi, uint64_t lo, int shift)
{
    return (unsigned)((shift < 64 ? (lo >> shift) : (hi >> (shift - 64))) & 0xffu);
}

/* calculate_tag_prority(), tailseq-dedup-perfect.c:87-105 (flag bits: sigproc-flags.h:30-41) */
__device__ __forceinline__ int dd_priority(int f)
{
    return ((f & 0x0010) == 0) * 1024 + ((f & 0x0080) == 0) * 512 + ((f & 0x0002) == 0) * 256 +
           ((f & 0x0008) != 0) * 128 + ((f & 0x0800) == 0) * 64 + ((f & 0x0200) == 0) * 32 +
           ((f & 0x0400) == 0) * 16 + ((f & 0x0040) == 0) * 8 + ((f & 0x0020) == 0) * 4 +
           ((f & 0x0001) != 0) * 2;
}

/* ---- sort ----------------------------------------------------------------------------- */
__global__ void __launch_bounds__(DD_THREADS)
k_dedup_varbits(const uint64_t *__restrict__ hi, const uint64_t *__restrict__ lo, uint32_t n, uint64_t *out)
{
    const uint64_t h0 = hi[0], l0 = lo[0];
    uint64_t vh = 0, vl = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        vh |= hi[i] ^ h0; vl |= lo[i] ^ l0;
    }
    for (int o = 16; o; o >>= 1) {
        vh |= __shfl_xor_sync(FULL, vh, o); vl |= __shfl_xor_sync(FULL, vl, o);
    }
    if ((threadIdx.x & 31) == 0 && (vh | vl)) {
        if (vh) atomicOr((unsigned long long *)&out[0], (unsigned long long)vh);
        if (vl) atomicOr((unsigned long long *)&out[1], (unsigned long long)vl);
    }
}

__global__ void __launch_bounds__(DD_THREADS)
k_dedup_sort_hist(const uint64_t *__restrict__ hi, const uint64_t *__restrict__ lo, uint32_t n, int shift,
                  uint32_t chunk, uint32_t *__restrict__ hist, uint32_t nblk)
{
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t beg = blockIdx.x * chunk;
    const uint32_t end = (n - beg < chunk) ? n : beg + chunk;
    const uint64_t *src = shift < 64 ? lo : hi;
    const int sh = shift & 63;
    for (uint32_t i = beg + threadIdx.x; i < end; i += DD_THREADS)
        atomicAdd(&h[(unsigned)(src[i] >> sh) & 0xffu], 1u);
    __syncthreads();
    hist[threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];
}

/* digit-major histogram [256][nblk]: block d scans row d in place (exclusive) and leaves the row
 * total; the scatter kernel adds the exclusive prefix of the 256 totals itself */
__global__ void __launch_bounds__(DD_THREADS) k_dedup_hist_rowscan(uint32_t *hist, uint32_t nblk, uint32_t *rowtot)
{
    __shared__ uint32_t ws[DD_WARPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t *row = hist + (size_t)blockIdx.x * nblk;
    const uint32_t per = (nblk + DD_THREADS - 1) / DD_THREADS;
    const uint32_t beg = min(nblk, tid * per), end = min(nblk, beg + per);
    uint32_t s = 0;
    for (uint32_t i = beg; i < end; i++) s += row[i];
    uint32_t inc = s;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += v; }
    if (lane == 31) ws[warp] = inc;
    __syncthreads();
    uint32_t run = inc - s;
    for (int w = 0; w < warp; w++) run += ws[w];
    for (uint32_t i = beg; i < end; i++) { const uint32_t v = row[i]; row[i] = run; run += v; }
    if (tid == DD_THREADS - 1) rowtot[blockIdx.x] = run;
}

/* exclusive scan of `count` u32 in place, one block */
__global__ void __launch_bounds__(1024) k_dedup_scan_single(uint32_t *data, uint32_t count)
{
    __shared__ uint32_t wsum[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t per = (count + 1023) / 1024;
    const uint32_t beg = min(count, tid * per), end = min(count, beg + per);
    uint32_t s = 0;
    for (uint32_t i = beg; i < end; i++) s += data[i];
    uint32_t inc = s;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += v; }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = wsum[lane], wi = w;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(FULL, wi, o); if (lane >= o) wi += v; }
        wsum[lane] = wi - w;
    }
    __syncthreads();
    uint32_t run = wsum[warp] + inc - s;
    for (uint32_t i = beg; i < end; i++) { const uint32_t v = data[i]; data[i] = run; run += v; }
}

/* One pass of the sort.  A block walks its chunk in tiles of DD_SORT_ROWS x 256 records.  Rank of a
 * record among the tile's records of the same digit (stable: rows in order, warps in order, lanes
 * in order) from __match_any_sync per warp-row and a running per-digit count; the tile is then put
 * in digit order in shared memory and written out from there, so that the records of one digit
 * leave as one run of consecutive addresses (a scattered 8-byte store takes a whole sector slot in
 * L2: three of those per record made the first version L2-write-bound at 0.30 ms a pass). */
#define DD_SORT_ROWS 4
#define DD_SORT_TILE (DD_SORT_ROWS * DD_THREADS)
__global__ void __launch_bounds__(DD_THREADS)
k_dedup_sort_scatter(const uint64_t *__restrict__ hi, const uint64_t *__restrict__ lo, const uint32_t *__restrict__ ord,
                     uint64_t *__restrict__ hi2, uint64_t *__restrict__ lo2, uint32_t *__restrict__ ord2,
                     uint32_t n, int shift, uint32_t chunk, const uint32_t *__restrict__ hist, uint32_t nblk,
                     const uint32_t *__restrict__ rowtot)
{
    __shared__ uint32_t base[256];               /* next global slot of every digit for this block */
    __shared__ uint32_t tcnt[256];               /* records of the digit in the tile so far */
    __shared__ uint32_t dstart[256];             /* first staged slot of the digit */
    __shared__ uint32_t wsum[DD_WARPS];
    __shared__ uint32_t wcnt[DD_WARPS][256];
    __shared__ uint64_t s_hi[DD_SORT_TILE], s_lo[DD_SORT_TILE];
    __shared__ uint32_t s_ord[DD_SORT_TILE];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {   /* first slot of digit `tid` in this block = records of smaller digits + this digit's in earlier blocks */
        const uint32_t t = rowtot[tid];
        uint32_t inc = t;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += v; }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        uint32_t pre = inc - t;
        for (int w = 0; w < warp; w++) pre += wsum[w];
        base[tid] = pre + hist[tid * nblk + blockIdx.x];
        __syncthreads();
    }
    const uint32_t beg = blockIdx.x * chunk;
    const uint32_t end = (n - beg < chunk) ? n : beg + chunk;
    for (uint32_t t0 = beg; t0 < end; t0 += DD_SORT_TILE) {
        uint64_t h[DD_SORT_ROWS], l[DD_SORT_ROWS];
        uint32_t o[DD_SORT_ROWS], rk[DD_SORT_ROWS];
        unsigned d[DD_SORT_ROWS];
#pragma unroll
        for (int k = 0; k < DD_SORT_ROWS; k++) {
            const uint32_t i = t0 + k * DD_THREADS + tid;
            d[k] = 256; h[k] = 0; l[k] = 0; o[k] = 0;
            if (i < end) { h[k] = hi[i]; l[k] = lo[i]; o[k] = ord ? ord[i] : i; d[k] = dd_digit(h[k], l[k], shift); }
        }
        tcnt[tid] = 0;
#pragma unroll
        for (int k = 0; k < DD_SORT_ROWS; k++) {
#pragma unroll
            for (int w = 0; w < DD_WARPS; w++) wcnt[w][tid] = 0;
            __syncthreads();
            const unsigned peers = __match_any_sync(FULL, d[k]);
            const unsigned rank = __popc(peers & ((1u << lane) - 1));
            if (d[k] < 256 && rank == 0) wcnt[warp][d[k]] = __popc(peers);
            __syncthreads();
            {
                uint32_t acc = tcnt[tid];
#pragma unroll
                for (int w = 0; w < DD_WARPS; w++) { const uint32_t c = wcnt[w][tid]; wcnt[w][tid] = acc; acc += c; }
                tcnt[tid] = acc;
            }
            __syncthreads();
            rk[k] = d[k] < 256 ? wcnt[warp][d[k]] + rank : 0;
            __syncthreads();
        }
        {   /* staged slot of the first record of every digit: exclusive scan of the tile's counts */
            const uint32_t t = tcnt[tid];
            uint32_t inc = t;
            for (int q = 1; q < 32; q <<= 1) { const uint32_t v = __shfl_up_sync(FULL, inc, q); if (lane >= q) inc += v; }
            if (lane == 31) wsum[warp] = inc;
            __syncthreads();
            uint32_t pre = inc - t;
            for (int w = 0; w < warp; w++) pre += wsum[w];
            dstart[tid] = pre;
            __syncthreads();
        }
#pragma unroll
        for (int k = 0; k < DD_SORT_ROWS; k++)
            if (d[k] < 256) {
                const uint32_t at = dstart[d[k]] + rk[k];
                s_hi[at] = h[k]; s_lo[at] = l[k]; s_ord[at] = o[k];
            }
        __syncthreads();
        const uint32_t cnt = (end - t0 < DD_SORT_TILE) ? end - t0 : DD_SORT_TILE;
        for (uint32_t i = tid; i < cnt; i += DD_THREADS) {
            const uint64_t hh = s_hi[i], ll = s_lo[i];
            const unsigned dg = dd_digit(hh, ll, shift);
            const uint32_t pos = base[dg] + (i - dstart[dg]);
            hi2[pos] = hh; lo2[pos] = ll; ord2[pos] = s_ord[i];
        }
        __syncthreads();
        base[tid] += tcnt[tid];
        __syncthreads();
    }
}

/* ---- group numbers: scan of "UMI differs from the one before" --------------------------- */
__device__ __forceinline__ bool dd_head(const uint64_t *hi, const uint64_t *lo, uint32_t p)
{
    return p == 0 || hi[p] != hi[p - 1] || lo[p] != lo[p - 1];
}

/* a block takes DD_SCAN_TILE positions, a warp 8 rows of 32 consecutive ones (coalesced); a row's
 * heads are one ballot */
__global__ void __launch_bounds__(DD_THREADS)
k_dedup_heads_count(const uint64_t *__restrict__ hi, const uint64_t *__restrict__ lo, uint32_t n, uint32_t *partials)
{
    __shared__ uint32_t ws[DD_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t w0 = blockIdx.x * DD_SCAN_TILE + warp * (32 * DD_SCAN_ITEMS);
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < DD_SCAN_ITEMS; k++) {
        const uint32_t p = w0 + k * 32 + lane;
        c += __popc(__ballot_sync(FULL, p < n && dd_head(hi, lo, p)));
    }
    if (lane == 0) ws[warp] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t s = 0;
        for (int w = 0; w < DD_WARPS; w++) s += ws[w];
        partials[blockIdx.x] = s;
    }
}

/* second half of the scan: group number of every position, start of every group, and the
 * per-position priority / length gathered through the sort order */
__global__ void __launch_bounds__(DD_THREADS)
k_dedup_heads_apply(const uint64_t *__restrict__ hi, const uint64_t *__restrict__ lo, const uint32_t *__restrict__ ord,
                    const int32_t *__restrict__ flags, const int32_t *__restrict__ polya, uint32_t n,
                    const uint32_t *__restrict__ partials,
                    uint32_t *__restrict__ gid, uint32_t *__restrict__ gstart,
                    int32_t *__restrict__ prio, int32_t *__restrict__ pa)
{
    __shared__ uint32_t ws[DD_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t w0 = blockIdx.x * DD_SCAN_TILE + warp * (32 * DD_SCAN_ITEMS);
    unsigned heads[DD_SCAN_ITEMS];
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < DD_SCAN_ITEMS; k++) {
        const uint32_t p = w0 + k * 32 + lane;
        heads[k] = __ballot_sync(FULL, p < n && dd_head(hi, lo, p));
        c += __popc(heads[k]);
    }
    if (lane == 0) ws[warp] = c;
    __syncthreads();
    uint32_t run = partials[blockIdx.x];                      /* already scanned (exclusive) */
    for (int w = 0; w < warp; w++) run += ws[w];
    const unsigned le = 0xffffffffu >> (31 - lane);
#pragma unroll
    for (int k = 0; k < DD_SCAN_ITEMS; k++) {
        const uint32_t p = w0 + k * 32 + lane;
        if (p < n) {
            const uint32_t g = run + __popc(heads[k] & le) - 1;
            gid[p] = g;
            if ((heads[k] >> lane) & 1u) gstart[g] = p;
            const uint32_t r = ord ? ord[p] : p;
            prio[p] = dd_priority(flags[r]);
            pa[p] = polya[r];
        }
        run += __popc(heads[k]);
    }
}

__global__ void k_dedup_close_groups(uint32_t *gstart, const uint32_t *gid, uint32_t n)
{
    gstart[gid[n - 1] + 1] = n;
}

/* ---- collapse --------------------------------------------------------------------------- */
struct DdArgs {
    const int32_t *prio, *pa;
    const uint32_t *gstart;
    const uint32_t *last_gid;                /* group number of the last position */
    int32_t *tval; uint32_t *tslot;
    uint32_t *rep, *ndup, *nkept; int32_t *gpolya;
    uint32_t *biglist;
    uint32_t *cands;                         /* NULL: do not look for full-queue arrivals */
    uint32_t cand_cap;
};

/* queue sizes of the reference: 2048, then (ssize_t)(size * 1.5f) (:38,43,148) */
#define DD_MAX_SIZES 56
struct DdSizes { uint32_t n; uint32_t s[DD_MAX_SIZES]; };

__device__ __forceinline__ int dd_final_length(int rep_polya, float mean, uint32_t nkept)
{
    if (nkept == 1) return rep_polya;                         /* the line is printed as it came (:286-287) */
    return rep_polya >= 0 ? (int)roundf(mean) : -1;           /* :323 */
}

/* one group of 2..32 records on one lane: the four record classes as bit masks */
__device__ __forceinline__ void dd_lane_group(const DdArgs &A, uint32_t g, uint32_t s, uint32_t m)
{
    unsigned Bimm = 0, Bbrk = 0, Bk = 0;
    int rmax = -1, sum = 0;
    for (uint32_t j = 0; j < m; j++) {
        const int pr = A.prio[s + j];
        const unsigned bit = 1u << j;
        if (pr > rmax) { rmax = pr; Bbrk |= bit; Bk = bit; sum = A.pa[s + j]; }
        else if (pr == rmax) { Bk |= bit; sum += A.pa[s + j]; }
        else Bimm |= bit;
    }
    const unsigned all = 0xffffffffu >> (32 - m);
    const unsigned Bq = all & ~Bimm & ~Bk;
    const uint32_t nkept = __popc(Bk);
    const float mean = __fdiv_rn((float)sum, (float)nkept);                     /* :303 */
    int rep = __ffs(Bk) - 1, rep_pa = 0;
    {
        float best = 0.f;
        bool first = true;
        for (unsigned rest = Bk; rest; rest &= rest - 1) {
            const int j = __ffs(rest) - 1;
            const int pa = A.pa[s + j];
            const float dist = fabsf(__fsub_rn((float)pa, mean));               /* :307,310 */
            if (first || dist < best) { best = dist; rep = j; rep_pa = pa; first = false; }
        }
    }
    const int final_len = dd_final_length(rep_pa, mean, nkept);
    for (uint32_t j = 0; j < m; j++) {
        const unsigned bit = 1u << j, lt = bit - 1, le = lt | bit;
        uint32_t pos;
        int val = -3;
        if (Bimm & bit) {
            const int b = 31 - __clz(Bbrk & le);
            pos = __popc(Bimm & lt) + __popc(Bq & ((1u << b) - 1));
        } else if (Bq & bit) {
            const int k = __ffs(Bbrk & ~le) - 1;
            pos = __popc(Bimm & ((1u << k) - 1)) + __popc(Bq & lt);
        } else {
            pos = (m - nkept) + __popc(Bk & lt);
            val = (int)j == rep ? final_len : -2;
        }
        A.tval[s + j] = val;
        A.tslot[s + j] = s + pos;
    }
    A.rep[g] = s + rep; A.gpolya[g] = final_len; A.ndup[g] = m; A.nkept[g] = nkept;
}

/* a lane takes a group: single records directly, 2..32 with bit masks, more -> list */
__global__ void __launch_bounds__(DD_THREADS) k_dedup_collapse(const DdArgs A)
{
    const uint32_t ngroups = *A.last_gid + 1;
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += gridDim.x * blockDim.x) {
        const uint32_t s = A.gstart[g], m = A.gstart[g + 1] - s;
        if (m == 1) {
            const int pa = A.pa[s];
            A.tval[s] = pa; A.tslot[s] = s;
            A.rep[g] = s; A.gpolya[g] = pa; A.ndup[g] = 1; A.nkept[g] = 1;
        } else if (m <= 32) dd_lane_group(A, g, s, m);
        else A.biglist[1 + atomicAdd(&A.biglist[0], 1u)] = g;
    }
}

struct DdBlock {
    int wmax[DD_BIG_WARPS];
    unsigned long long wsum[DD_BIG_WARPS];
    int cmax;
    unsigned long long carry;
    uint32_t epoch_ci[DD_MAX_EPOCHS + 2], epoch_cq[DD_MAX_EPOCHS + 2], epoch_ce[DD_MAX_EPOCHS + 2];
    int rmax[DD_BIG_WARPS]; unsigned long long rcnt[DD_BIG_WARPS]; long long rsum[DD_BIG_WARPS];
    unsigned long long rkey[DD_BIG_WARPS];
};

/* exclusive running maximum over the block's threads on top of `carry` */
__device__ __forceinline__ int dd_block_exmax(int v, int carry, int *wmax, int tid)
{
    const int lane = tid & 31, warp = tid >> 5;
    int inc = v;
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc = max(inc, t); }
    if (lane == 31) wmax[warp] = inc;
    int ex = __shfl_up_sync(FULL, inc, 1);
    if (lane == 0) ex = -1;
    __syncthreads();
    int pre = carry;
    for (int w = 0; w < warp; w++) pre = max(pre, wmax[w]);
    __syncthreads();
    return max(ex, pre);
}

/* exclusive sum of four 16-bit counters packed in 64 bits (a tile has 1024 records) */
__device__ __forceinline__ unsigned long long dd_block_exsum(unsigned long long v, unsigned long long *wsum, int tid,
                                                             unsigned long long *total)
{
    const int lane = tid & 31, warp = tid >> 5;
    unsigned long long inc = v;
    for (int o = 1; o < 32; o <<= 1) { const unsigned long long t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    unsigned long long pre = 0, tot = 0;
    for (int w = 0; w < DD_BIG_WARPS; w++) { if (w < warp) pre += wsum[w]; tot += wsum[w]; }
    __syncthreads();
    *total = tot;
    return pre + inc - v;
}

/* groups of more than 32 records: one CTA each, two passes over the group */
__global__ void __launch_bounds__(DD_BIG_THREADS) k_dedup_collapse_big(const DdArgs A, const DdSizes Z)
{
    __shared__ DdBlock sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t nbig = A.biglist[0];
    for (uint32_t b = blockIdx.x; b < nbig; b += gridDim.x) {
        const uint32_t g = A.biglist[1 + b];
        const uint32_t s = A.gstart[g], m = A.gstart[g + 1] - s;
        __syncthreads();

        /* pass 1: epochs, #imm before each brk; (max priority, how many have it, their length sum) */
        int cmax = -1;
        uint32_t c_imm = 0, c_brk = 0;
        int tmax = -1; uint32_t tcnt = 0; int tsum = 0;
        for (uint32_t t0 = 0; t0 < m; t0 += DD_BIG_THREADS) {
            const uint32_t j = t0 + tid;
            const bool act = j < m;
            const int pr = act ? A.prio[s + j] : -1;
            const int rex = dd_block_exmax(pr, cmax, sm.wmax, tid);
            const bool imm = act && pr < rex, brk = act && pr > rex;
            unsigned long long tot;
            const unsigned long long ex = dd_block_exsum((unsigned long long)imm | ((unsigned long long)brk << 16), sm.wsum, tid, &tot);
            if (brk) sm.epoch_ci[c_brk + (uint32_t)((ex >> 16) & 0xffff)] = c_imm + (uint32_t)(ex & 0xffff);
            if (act) {
                const int pa = A.pa[s + j];
                if (pr > tmax) { tmax = pr; tcnt = 1; tsum = pa; }
                else if (pr == tmax) { tcnt++; tsum += pa; }
            }
            c_imm += (uint32_t)(tot & 0xffff); c_brk += (uint32_t)((tot >> 16) & 0xffff);
            /* running maximum after this tile */
            int wm = max(pr, rex);
            wm = __reduce_max_sync(FULL, wm);
            if (lane == 0) sm.wmax[warp] = wm;
            __syncthreads();
            for (int w = 0; w < DD_BIG_WARPS; w++) cmax = max(cmax, sm.wmax[w]);
            __syncthreads();
        }
        const int gmax = cmax;
        if (tmax != gmax) { tcnt = 0; tsum = 0; }
        const uint32_t wc = __reduce_add_sync(FULL, tcnt);
        const int wsu = __reduce_add_sync(FULL, tsum);
        if (lane == 0) { sm.rcnt[warp] = wc; sm.rsum[warp] = wsu; }
        __syncthreads();
        uint32_t nkept = 0; int sum = 0;
        for (int w = 0; w < DD_BIG_WARPS; w++) { nkept += (uint32_t)sm.rcnt[w]; sum += (int)sm.rsum[w]; }
        __syncthreads();
        const float mean = __fdiv_rn((float)sum, (float)nkept);

        /* pass 2: positions, nearest kept record */
        cmax = -1; c_imm = 0; c_brk = 0;
        uint32_t c_q = 0, c_k = 0;
        unsigned long long best = ~0ull;
        for (uint32_t t0 = 0; t0 < m; t0 += DD_BIG_THREADS) {
            const uint32_t j = t0 + tid;
            const bool act = j < m;
            const int pr = act ? A.prio[s + j] : -1;
            const int rex = dd_block_exmax(pr, cmax, sm.wmax, tid);
            const bool imm = act && pr < rex, brk = act && pr > rex, kept = act && pr == gmax;
            const bool q = act && !imm && !kept;
            unsigned long long tot;
            const unsigned long long ex = dd_block_exsum((unsigned long long)imm | ((unsigned long long)brk << 16) |
                                                         ((unsigned long long)q << 32) | ((unsigned long long)kept << 48),
                                                         sm.wsum, tid, &tot);
            const uint32_t ci = c_imm + (uint32_t)(ex & 0xffff), eb = c_brk + (uint32_t)((ex >> 16) & 0xffff);
            const uint32_t cq = c_q + (uint32_t)((ex >> 32) & 0xffff), ck = c_k + (uint32_t)((ex >> 48) & 0xffff);
            const uint32_t epoch = brk ? eb : eb - 1;         /* the first record of a group is a brk */
            if (brk) { sm.epoch_cq[epoch] = cq; sm.epoch_ce[epoch] = cq + ck; }
            __syncthreads();
            if (act && !brk && A.cands) {
                /* records in the queue when this one arrives (the brk path returns before the
                 * full-queue test, :178-198) */
                const uint32_t qlen = cq + ck - sm.epoch_ce[epoch];
                if (qlen + 1 >= Z.s[0])
                    for (uint32_t k = 0; k < Z.n; k++)
                        if (qlen + 1 == Z.s[k]) {
                            const uint32_t at = atomicAdd(&A.cands[0], 1u);
                            if (at < A.cand_cap) { A.cands[1 + 2 * at] = s + j; A.cands[2 + 2 * at] = k; }
                        }
            }
            if (act) {
                uint32_t pos; int val = -3;
                if (imm) pos = ci + sm.epoch_cq[epoch];
                else if (q) pos = sm.epoch_ci[epoch + 1] + cq;
                else {
                    pos = (m - nkept) + ck; val = -2;
                    const float dist = fabsf(__fsub_rn((float)A.pa[s + j], mean));
                    const unsigned long long key = ((unsigned long long)__float_as_uint(dist) << 32) | j;
                    best = key < best ? key : best;
                }
                A.tval[s + j] = val; A.tslot[s + j] = s + pos;
            }
            c_imm += (uint32_t)(tot & 0xffff); c_brk += (uint32_t)((tot >> 16) & 0xffff);
            c_q += (uint32_t)((tot >> 32) & 0xffff); c_k += (uint32_t)((tot >> 48) & 0xffff);
            int wm = max(pr, rex);
            wm = __reduce_max_sync(FULL, wm);
            if (lane == 0) sm.wmax[warp] = wm;
            __syncthreads();
            for (int w = 0; w < DD_BIG_WARPS; w++) cmax = max(cmax, sm.wmax[w]);
            __syncthreads();
        }
        for (int o = 16; o; o >>= 1) { const unsigned long long t = __shfl_xor_sync(FULL, best, o); best = t < best ? t : best; }
        if (lane == 0) sm.rkey[warp] = best;
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < DD_BIG_WARPS; w++) best = sm.rkey[w] < best ? sm.rkey[w] : best;
            const uint32_t rep = (uint32_t)(best & 0xffffffffu);
            const int final_len = dd_final_length(A.pa[s + rep], mean, nkept);
            A.tval[s + rep] = final_len;                      /* after the barrier: overwrites the -2 */
            A.rep[g] = s + rep; A.gpolya[g] = final_len; A.ndup[g] = m; A.nkept[g] = nkept;
        }
        __syncthreads();
    }
}

/* ---- C ABI ------------------------------------------------------------------------------ */
static void dd_free(tsk_dedup *d)
{
    void *bufs[] = {d->kin_hi, d->kin_lo, d->khi[0], d->khi[1], d->klo[0], d->klo[1], d->ord[0], d->ord[1], d->flags, d->polya, d->prio, d->pa,
                    d->gid, d->gstart, d->tval, d->tslot, d->rep, d->ndup, d->nkept, d->gpolya, d->biglist, d->cands, d->hist, d->rowtot,
                    d->partials, d->varbits};
    for (size_t i = 0; i < sizeof(bufs) / sizeof(bufs[0]); i++) if (bufs[i]) cudaFree(bufs[i]);
    for (int i = 0; i < 3; i++) if (d->ev[i]) cudaEventDestroy(d->ev[i]);
    if (d->stream) cudaStreamDestroy(d->stream);
    free(d->lost);
    free(d);
}

extern "C" void tsk_dedup_destroy(tsk_dedup *d)
{
    if (!d) return;
    cudaSetDevice(d->device);
    dd_free(d);
}

static int dd_alloc(tsk_dedup *d)
{
    const size_t c = (size_t)d->cap;
    TSK_CUDA(cudaMalloc(&d->kin_hi, 8 * c)); TSK_CUDA(cudaMalloc(&d->kin_lo, 8 * c));
    for (int i = 0; i < 2; i++) {
        TSK_CUDA(cudaMalloc(&d->khi[i], 8 * c)); TSK_CUDA(cudaMalloc(&d->klo[i], 8 * c));
        TSK_CUDA(cudaMalloc(&d->ord[i], 4 * c));
    }
    TSK_CUDA(cudaMalloc(&d->flags, 4 * c)); TSK_CUDA(cudaMalloc(&d->polya, 4 * c));
    TSK_CUDA(cudaMalloc(&d->prio, 4 * c)); TSK_CUDA(cudaMalloc(&d->pa, 4 * c));
    TSK_CUDA(cudaMalloc(&d->gid, 4 * c)); TSK_CUDA(cudaMalloc(&d->gstart, 4 * (c + 1)));
    TSK_CUDA(cudaMalloc(&d->tval, 4 * c)); TSK_CUDA(cudaMalloc(&d->tslot, 4 * c));
    TSK_CUDA(cudaMalloc(&d->rep, 4 * c)); TSK_CUDA(cudaMalloc(&d->ndup, 4 * c));
    TSK_CUDA(cudaMalloc(&d->nkept, 4 * c)); TSK_CUDA(cudaMalloc(&d->gpolya, 4 * c));
    TSK_CUDA(cudaMalloc(&d->biglist, 4 * (c / 33 + 2)));
    d->cand_cap = (uint32_t)(DD_MAX_SIZES * (c / 2047 + 1));
    TSK_CUDA(cudaMalloc(&d->cands, 4 * (1 + 2 * (size_t)d->cand_cap)));
    d->nblk = 148 * 8;
    TSK_CUDA(cudaMalloc(&d->hist, 4 * 256 * (size_t)d->nblk));
    TSK_CUDA(cudaMalloc(&d->rowtot, 4 * 256));
    TSK_CUDA(cudaMalloc(&d->partials, 4 * (c / DD_SCAN_TILE + 2)));
    TSK_CUDA(cudaMalloc(&d->varbits, 16));
    TSK_CUDA(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
    for (int i = 0; i < 3; i++) TSK_CUDA(cudaEventCreate(&d->ev[i]));
    return 0;
}

extern "C" int tsk_dedup_create(tsk_dedup **out, int device, uint64_t capacity)
{
    if (!out) { tsk_set_error("tsk_dedup_create: null argument"); return -1; }
    *out = NULL;
    if (capacity < 1 || capacity >= 0xffffffffull - DD_SCAN_TILE) { tsk_set_error("tsk_dedup_create: capacity must be 1 .. 2^32 - %d", DD_SCAN_TILE + 1); return -1; }
    if (tsk_probe(device) < 0) return -1;
    TSK_CUDA(cudaSetDevice(device));
    tsk_dedup *d = (tsk_dedup *)calloc(1, sizeof(*d));
    if (!d) { tsk_set_error("out of memory"); return -1; }
    d->device = device; d->cap = capacity;
    if (dd_alloc(d) != 0) { dd_free(d); return -1; }
    *out = d;
    return 0;
}

extern "C" int tsk_dedup_upload(tsk_dedup *d, uint64_t n, const uint64_t *umi_hi, const uint64_t *umi_lo,
                                const int32_t *flags, const int32_t *polya)
{
    if (!d || (n && (!umi_hi || !umi_lo || !flags || !polya))) { tsk_set_error("tsk_dedup_upload: null argument"); return -1; }
    if (n > d->cap) { tsk_set_error("tsk_dedup_upload: %llu records exceed the capacity %llu", (unsigned long long)n, (unsigned long long)d->cap); return -1; }
    TSK_CUDA(cudaSetDevice(d->device));
    d->n = n; d->ngroups = 0; d->sorted = 0; d->cur = 0; d->done = 0;
    if (n == 0) return 0;
    TSK_CUDA(cudaMemcpyAsync(d->kin_hi, umi_hi, 8 * n, cudaMemcpyDefault, d->stream));
    TSK_CUDA(cudaMemcpyAsync(d->kin_lo, umi_lo, 8 * n, cudaMemcpyDefault, d->stream));
    TSK_CUDA(cudaMemcpyAsync(d->flags, flags, 4 * n, cudaMemcpyDefault, d->stream));
    TSK_CUDA(cudaMemcpyAsync(d->polya, polya, 4 * n, cudaMemcpyDefault, d->stream));
    TSK_CUDA(cudaStreamSynchronize(d->stream));
    return 0;
}

static int dd_sort(tsk_dedup *d)
{
    const uint32_t n = (uint32_t)d->n;
    uint64_t var[2];
    TSK_CUDA(cudaMemsetAsync(d->varbits, 0, 16, d->stream));
    k_dedup_varbits<<<148 * 4, DD_THREADS, 0, d->stream>>>(d->kin_hi, d->kin_lo, n, d->varbits);
    d->launches++;
    TSK_CUDA(cudaMemcpyAsync(var, d->varbits, 16, cudaMemcpyDeviceToHost, d->stream));
    TSK_CUDA(cudaStreamSynchronize(d->stream));
    uint32_t nblk = (n + 2047) / 2048;
    if (nblk > d->nblk) nblk = d->nblk;
    uint32_t chunk = (n + nblk - 1) / nblk;
    chunk = (chunk + DD_SORT_TILE - 1) / DD_SORT_TILE * DD_SORT_TILE;
    nblk = (n + chunk - 1) / chunk;
    /* the first pass reads the uploaded keys, later ones alternate between the two work buffers */
    const uint64_t *shi = d->kin_hi, *slo = d->kin_lo;
    const uint32_t *sord = NULL;
    int dst = 0, passes = 0;
    for (int shift = 0; shift < 128; shift += 8) {
        const uint64_t v = shift < 64 ? var[1] >> shift : var[0] >> (shift - 64);
        if ((v & 0xff) == 0) continue;
        k_dedup_sort_hist<<<nblk, DD_THREADS, 0, d->stream>>>(shi, slo, n, shift, chunk, d->hist, nblk);
        k_dedup_hist_rowscan<<<256, DD_THREADS, 0, d->stream>>>(d->hist, nblk, d->rowtot);
        k_dedup_sort_scatter<<<nblk, DD_THREADS, 0, d->stream>>>(shi, slo, sord, d->khi[dst], d->klo[dst], d->ord[dst],
                                                                 n, shift, chunk, d->hist, nblk, d->rowtot);
        d->launches += 3;
        shi = d->khi[dst]; slo = d->klo[dst]; sord = d->ord[dst];
        d->cur = dst; dst ^= 1; passes++;
    }
    d->sorted = passes > 0;                    /* all keys equal: the given order stands */
    TSK_CUDA(cudaGetLastError());
    return 0;
}

__global__ void k_dedup_iota(uint32_t *out, uint32_t n)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = i;
}

/* The reference's queue has `size` slots for the whole run (2048 at first).  A record that is not
 * the first of a higher priority and finds size - 1 records queued makes tqueue_append() call
 * tqueue_expand() (:196-198), which moves the queued records to a calloc'ed array of (ssize_t)
 * (size * 1.5f) slots but not the arriving one: it goes on as tile "", cluster 0, flags 0, length 0,
 * no modifications, with the priority computed from its real flags.  The kernel listed every
 * record arriving at a queue of size_k - 1 records for some k; in order of position, the one that
 * meets the current k is lost and k advances.  Their lengths become 0 and their groups are
 * collapsed again. */
static int cmp_u32_pair(const void *a, const void *b)
{
    const uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return x < y ? -1 : x > y;
}

static int dd_lost_records(tsk_dedup *d, DdArgs a, const DdSizes &z, uint32_t ncand)
{
    if (ncand > d->cand_cap) { tsk_set_error("tsk_dedup_run: internal candidate list overflow"); return -1; }
    uint32_t *c = (uint32_t *)malloc(8 * (size_t)ncand);
    uint32_t *groups = (uint32_t *)malloc(4 * ((size_t)z.n + 1));
    if (!c || !groups) { free(c); free(groups); tsk_set_error("out of memory"); return -1; }
    if (cudaMemcpyAsync(c, d->cands + 1, 8 * (size_t)ncand, cudaMemcpyDeviceToHost, d->stream) != cudaSuccess ||
        cudaStreamSynchronize(d->stream) != cudaSuccess) { free(c); free(groups); tsk_set_error("candidate copy failed"); return -1; }
    qsort(c, ncand, 8, cmp_u32_pair);
    free(d->lost);
    d->lost = (uint32_t *)malloc(4 * ((size_t)z.n + 1));
    if (!d->lost) { free(c); free(groups); tsk_set_error("out of memory"); return -1; }
    uint32_t k = 0, nl = 0;
    for (uint32_t i = 0; i < ncand && k < z.n; i++)
        if (c[2 * i + 1] == k) { d->lost[nl++] = c[2 * i]; k++; }
    free(c);
    d->nlost = nl;
    int rc = 0;
    uint32_t ng = 0;
    const int32_t zero = 0;
    for (uint32_t i = 0; i < nl && rc == 0; i++) {
        uint32_t g;
        if (cudaMemcpyAsync(d->pa + d->lost[i], &zero, 4, cudaMemcpyHostToDevice, d->stream) != cudaSuccess ||
            cudaMemcpyAsync(&g, d->gid + d->lost[i], 4, cudaMemcpyDeviceToHost, d->stream) != cudaSuccess ||
            cudaStreamSynchronize(d->stream) != cudaSuccess) { tsk_set_error("lost-record fix-up failed"); rc = -1; break; }
        if (ng == 0 || groups[ng] != g) groups[++ng] = g;      /* positions ascend, so do groups */
    }
    if (rc == 0 && ng > 0) {
        groups[0] = ng;
        if (cudaMemcpyAsync(d->biglist, groups, 4 * ((size_t)ng + 1), cudaMemcpyHostToDevice, d->stream) != cudaSuccess) {
            tsk_set_error("lost-record fix-up failed"); rc = -1;
        } else {
            a.cands = NULL;
            k_dedup_collapse_big<<<ng < 148 ? ng : 148, DD_BIG_THREADS, 0, d->stream>>>(a, z);
            d->launches++;
            if (cudaStreamSynchronize(d->stream) != cudaSuccess) { tsk_set_error("lost-record fix-up kernel failed"); rc = -1; }
        }
    }
    free(groups);
    return rc;
}

extern "C" int tsk_dedup_run(tsk_dedup *d, int sort_by_umi, uint64_t *n_groups)
{
    if (!d) { tsk_set_error("tsk_dedup_run: null argument"); return -1; }
    TSK_CUDA(cudaSetDevice(d->device));
    if (n_groups) *n_groups = 0;
    d->ngroups = 0; d->ms_sort = d->ms_collapse = 0.f; d->sorted = 0; d->cur = 0; d->done = 0; d->nlost = 0;
    if (d->n == 0) { d->done = 1; return 0; }
    const uint32_t n = (uint32_t)d->n;
    TSK_CUDA(cudaEventRecord(d->ev[0], d->stream));
    if (sort_by_umi && dd_sort(d) != 0) return -1;
    TSK_CUDA(cudaEventRecord(d->ev[1], d->stream));
    const uint64_t *hi = d->sorted ? d->khi[d->cur] : d->kin_hi, *lo = d->sorted ? d->klo[d->cur] : d->kin_lo;
    const uint32_t *ord = d->sorted ? d->ord[d->cur] : NULL;
    const uint32_t nparts = (n + DD_SCAN_TILE - 1) / DD_SCAN_TILE;
    k_dedup_heads_count<<<nparts, DD_THREADS, 0, d->stream>>>(hi, lo, n, d->partials);
    k_dedup_scan_single<<<1, 1024, 0, d->stream>>>(d->partials, nparts);
    k_dedup_heads_apply<<<nparts, DD_THREADS, 0, d->stream>>>(hi, lo, ord, d->flags, d->polya, n,
                                                              d->partials, d->gid, d->gstart, d->prio, d->pa);
    k_dedup_close_groups<<<1, 1, 0, d->stream>>>(d->gstart, d->gid, n);
    TSK_CUDA(cudaMemsetAsync(d->biglist, 0, 4, d->stream));
    TSK_CUDA(cudaMemsetAsync(d->cands, 0, 4, d->stream));
    DdArgs a = {d->prio, d->pa, d->gstart, d->gid + (n - 1), d->tval, d->tslot, d->rep, d->ndup, d->nkept, d->gpolya,
                d->biglist, d->cands, d->cand_cap};
    DdSizes z;
    z.n = 0;
    for (uint64_t sz = 2048; sz <= 0xffffffffull && z.n < DD_MAX_SIZES; sz = (uint64_t)(long long)((float)(long long)sz * 1.5f))
        z.s[z.n++] = (uint32_t)sz;
    uint32_t grid = (n + DD_THREADS - 1) / DD_THREADS;              /* at most n groups */
    if (grid > 148 * 8) grid = 148 * 8;
    k_dedup_collapse<<<grid, DD_THREADS, 0, d->stream>>>(a);
    k_dedup_collapse_big<<<148, DD_BIG_THREADS, 0, d->stream>>>(a, z);
    d->launches += 6;
    TSK_CUDA(cudaGetLastError());
    uint32_t last = 0, ncand = 0;
    TSK_CUDA(cudaMemcpyAsync(&last, d->gid + (n - 1), 4, cudaMemcpyDeviceToHost, d->stream));
    TSK_CUDA(cudaMemcpyAsync(&ncand, d->cands, 4, cudaMemcpyDeviceToHost, d->stream));
    TSK_CUDA(cudaStreamSynchronize(d->stream));
    d->ngroups = (uint64_t)last + 1;
    d->nlost = 0;
    if (ncand > 0 && dd_lost_records(d, a, z, ncand) != 0) return -1;
    TSK_CUDA(cudaEventRecord(d->ev[2], d->stream));
    TSK_CUDA(cudaStreamSynchronize(d->stream));
    TSK_CUDA(cudaEventElapsedTime(&d->ms_sort, d->ev[0], d->ev[1]));
    TSK_CUDA(cudaEventElapsedTime(&d->ms_collapse, d->ev[1], d->ev[2]));
    if (n_groups) *n_groups = d->ngroups;
    d->done = 1;
    return 0;
}

static int dd_fetch(tsk_dedup *d, void *dst, const void *src, size_t bytes)
{
    if (!dst || bytes == 0) return 0;
    TSK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, d->stream));
    return 0;
}

extern "C" int tsk_dedup_get_records(tsk_dedup *d, uint32_t *order, uint32_t *group, int32_t *trace_value, uint32_t *trace_slot)
{
    if (!d) { tsk_set_error("tsk_dedup_get_records: null argument"); return -1; }
    TSK_CUDA(cudaSetDevice(d->device));
    if (!d->done) { tsk_set_error("tsk_dedup_get_records before tsk_dedup_run"); return -1; }
    const size_t n = (size_t)d->n;
    if (n == 0) return 0;
    if (order) {
        if (!d->sorted) {
            k_dedup_iota<<<148 * 4, DD_THREADS, 0, d->stream>>>(d->ord[0], (uint32_t)n);
            TSK_CUDA(cudaGetLastError());
        }
        if (dd_fetch(d, order, d->ord[d->sorted ? d->cur : 0], 4 * n) != 0) return -1;
    }
    if (dd_fetch(d, group, d->gid, 4 * n) != 0 || dd_fetch(d, trace_value, d->tval, 4 * n) != 0 ||
        dd_fetch(d, trace_slot, d->tslot, 4 * n) != 0) return -1;
    TSK_CUDA(cudaStreamSynchronize(d->stream));
    return 0;
}

extern "C" int tsk_dedup_get_groups(tsk_dedup *d, uint32_t *rep, int32_t *polya, uint32_t *ndup, uint32_t *kept)
{
    if (!d) { tsk_set_error("tsk_dedup_get_groups: null argument"); return -1; }
    TSK_CUDA(cudaSetDevice(d->device));
    if (!d->done) { tsk_set_error("tsk_dedup_get_groups before tsk_dedup_run"); return -1; }
    const size_t g = (size_t)d->ngroups;
    if (dd_fetch(d, rep, d->rep, 4 * g) != 0 || dd_fetch(d, polya, d->gpolya, 4 * g) != 0 ||
        dd_fetch(d, ndup, d->ndup, 4 * g) != 0 || dd_fetch(d, kept, d->nkept, 4 * g) != 0) return -1;
    TSK_CUDA(cudaStreamSynchronize(d->stream));
    return 0;
}

extern "C" int tsk_dedup_get_lost(tsk_dedup *d, uint32_t *positions, uint32_t capacity, uint32_t *count)
{
    if (!d || !count) { tsk_set_error("tsk_dedup_get_lost: null argument"); return -1; }
    if (!d->done) { tsk_set_error("tsk_dedup_get_lost before tsk_dedup_run"); return -1; }
    *count = d->nlost;
    if (positions) {
        if (capacity < d->nlost) { tsk_set_error("tsk_dedup_get_lost: %u positions, room for %u", d->nlost, capacity); return -1; }
        memcpy(positions, d->lost, 4 * (size_t)d->nlost);
    }
    return 0;
}

extern "C" int tsk_dedup_elapsed_ms(tsk_dedup *d, float *sort_ms, float *collapse_ms)
{
    if (!d) { tsk_set_error("tsk_dedup_elapsed_ms: null argument"); return -1; }
    if (sort_ms) *sort_ms = d->ms_sort;
    if (collapse_ms) *collapse_ms = d->ms_collapse;
    return 0;
}

extern "C" uint64_t tsk_dedup_launch_count(tsk_dedup *d) { return d ? d->launches : 0; }
