/*
 * tsk_control.cu -- PhiX control classification of the barcode-unassigned clusters.
 *
 * Restates try_alignment_to_control() (src/importer/controlaligner.c:107-145) for the GPU:
 *   1. exact path (:117-119): the first L = phix-match-length bases of the read, taken from
 *      cycle 0, occur verbatim in either strand;
 *   2. otherwise the best local alignment of read[first .. first+L) against
 *      forward + 20 N + reverse (:74-104) -- match +1, mismatch -2, N 0, gap 4 + 1 per extra
 *      base (tailseq-import.h:152-155; ssw.c's striped kernel returns the exact optimum at
 *      these score ranges) -- reaches S = (int)(L * 0.65) (:135,162).
 *
 * The reference runs the Smith-Waterman matrix over all 10 792 target columns for every
 * unassigned read.  Here only the decision "score >= S" is needed, so a q-gram filter picks
 * the few target windows that can hold such an alignment and the matrix is evaluated there:
 *
 *   An alignment with M matches, X mismatches, n read Ns (score 0) and G gap events of total
 *   length Lg scores M - 2X - 3G - Lg.  Reaching S needs M >= S + 2X + 4G, and
 *   M + X + n <= L bounds 3X + 4G <= L - S - n.  Its match and read-N columns fall into at most
 *   X + G + 1 gap- and mismatch-free runs of total length M + n; a run of m columns holds
 *   m - (q-1) read q-grams that equal the target on the run's diagonal when a read N is
 *   taken as a wildcard.  So the alignment shares at least
 *       T = min over n of  S + n - (q-1) - max over feasible (X, G) of [(q-3) X + (q-5) G]
 *   wildcard q-grams with the target (L = 40, q = 6: T = 9, attained at n = 0, X = 4).
 *   Its diagonals span at most Lg + 1 <= L - S - 2 <= 17 values, i.e. two adjacent
 *   16-diagonal bins.  Crossing the 20-N spacer costs 20 read bases or a gap and cannot reach
 *   S while L - 20 < S.  (control_plan() checks these side conditions and otherwise scans the
 *   whole target.)
 *
 * One warp per read: the hits of the read's q-grams (a q-gram with k <= 3 Ns is looked up under
 * all 4^k spellings; one with more lowers T by one instead) are counted per diagonal bin in
 * shared memory through an index of the target, about 90 hits for a clean read; every run of
 * bin pairs with >= T hits is evaluated as one window by an anti-diagonal wavefront, two read
 * rows per lane, H / F / target base handed to the next lane in one packed shuffle.
 */
#include "tsk_internal.h"
#include "phix174.h"

#include <stdlib.h>
#include <string.h>

#define FULL_MASK 0xffffffffu
#define CTL_THREADS   256
#define CTL_WARPS     (CTL_THREADS / 32)
#define CTL_SPACING   20                                  /* controlaligner.c:39 */
#define CTL_TLEN      (2 * TSK_PHIX174_LENGTH + CTL_SPACING)
#define CTL_TPAD      128                                 /* N padding after the target */
#define CTL_BINW      16
#define CTL_NBIN      704                                 /* >= (CTL_TLEN + 64) / 16 + 2, multiple of 64 */
#define CTL_HASH      32768
#define CTL_CHUNK     128                                 /* clusters per work unit */
#define CTL_MAXWILD   3                                   /* Ns expanded per q-gram */

struct ControlPlan {
    int first, len, minscore;
    int filter;                  /* 0: evaluate the whole target for every read */
    int q;                       /* q-gram length of the filter: 6, or 5 for short control reads */
    int threshold;               /* wildcard q-grams an accepted alignment shares with the target */
    int hash_k;                  /* bases hashed for the exact path: min(len, 16) */
};

struct ControlIndex {
    const uint8_t *target;       /* [CTL_TLEN + CTL_TPAD] codes 0-3, 4 = N */
    const uint32_t *hash;        /* [CTL_HASH] target position + 1 of a length-L substring, 0 = empty */
    const uint16_t *qoff;        /* [4^q + 1] */
    const uint16_t *qpos;        /* target positions per q-gram */
};

__host__ __device__ static inline uint32_t ctl_hash_key(uint32_t key)
{
    return (key * 2654435761u) >> 17;                     /* 15 bits */
}

/* ---- windowed Smith-Waterman: does any local alignment of the read rows against target
 *      columns [t0, t0 + W) reach minscore?  Lane l owns read rows 2l and 2l+1. ---- */
__device__ static bool window_reaches(const uint8_t *__restrict__ target, int t0, int W,
                                      uint32_t tbl0, uint32_t tbl1, int minscore, int lane)
{
    /* tblX: substitution score + 2 of the row's base against target codes 0..4, 4 bits each */
    int hl0 = 0, hl1 = 0, e0 = 0, e1 = 0;      /* H(row, c-1), E(row, c-1) */
    int hup_prev = 0;                          /* H(2l-1, c-1) */
    int hbot = 0, fbot = 0, tb = 4, best = 0;
    uint32_t chunk = 0;
    const int steps = W + 31;
    for (int s = 0; s < steps; s++) {
        if ((s & 31) == 0) chunk = target[t0 + s + lane];             /* padded past the end */
        const uint32_t fresh = __shfl_sync(FULL_MASK, chunk, s & 31);
        const uint32_t pk = (uint32_t)hbot | ((uint32_t)fbot << 8) | ((uint32_t)tb << 16);
        uint32_t up = __shfl_up_sync(FULL_MASK, pk, 1);
        if (lane == 0) up = fresh << 16;
        const int c = s - lane;
        if (c >= 0 && c < W) {
            const int hup = up & 0xff, fup = (up >> 8) & 0xff;
            tb = up >> 16;
            const int sh = tb * 4;
            /* row 2l */
            const int s0 = (int)((tbl0 >> sh) & 0xfu) - 2;
            const int ee0 = max(e0 - 1, hl0 - 4);
            const int ff0 = max(fup - 1, hup - 4);
            const int h0 = max(max(hup_prev + s0, ee0), max(ff0, 0));
            /* row 2l+1 */
            const int s1 = (int)((tbl1 >> sh) & 0xfu) - 2;
            const int ee1 = max(e1 - 1, hl1 - 4);
            const int ff1 = max(max(ff0, 0) - 1, h0 - 4);
            const int h1 = max(max(hl0 + s1, ee1), max(ff1, 0));
            hup_prev = hup;
            hl0 = h0; hl1 = h1;
            e0 = max(ee0, 0); e1 = max(ee1, 0);
            hbot = h1; fbot = max(ff1, 0);
            best = max(best, max(h0, h1));
        }
        if ((s & 7) == 7 && __any_sync(FULL_MASK, best >= minscore)) return true;
    }
    return __any_sync(FULL_MASK, best >= minscore);
}

__global__ void __launch_bounds__(CTL_THREADS)
k_control_classify(const TileView tv, const __grid_constant__ ControlPlan plan,
                   const __grid_constant__ ControlIndex ix,
                   const uint8_t *__restrict__ kind, uint32_t *__restrict__ queue, tsk_call *__restrict__ calls,
                   uint32_t *__restrict__ counters)
{
    __shared__ uint32_t s_bins[CTL_WARPS][CTL_NBIN / 2];   /* two 16-bit counts per word */
    __shared__ uint8_t s_read[CTL_WARPS][256];
    __shared__ uint32_t s_tot[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int L = plan.len, first = plan.first;
    const int ncyc = max(L, first + L);
    uint32_t *bins = s_bins[warp];
    uint8_t *rd = s_read[warp];
    if (threadIdx.x < 2) s_tot[threadIdx.x] = 0;
    __syncthreads();

    /* K1 marked the clusters to classify in kind[].  Reads differ a lot in cost (none / one /
     * many windows), so warps pull chunks of CTL_CHUNK clusters from a counter. */
    const uint32_t nchunks = (tv.nclusters + CTL_CHUNK - 1) / CTL_CHUNK;
    uint32_t n_unknown = 0, n_control = 0;
    uint32_t chunk = 0, lanemask = 0;      /* lanemask: this lane's 4 clusters of the chunk still to do */
    unsigned lanes = 0;                    /* lanes with work left in the chunk */
    for (;;) {
        if (lanes == 0) {
            if (lane == 0) chunk = atomicAdd(&queue[1], 1u);
            chunk = __shfl_sync(FULL_MASK, chunk, 0);
            if (chunk >= nchunks) break;
            const uint32_t c0 = chunk * CTL_CHUNK + lane * 4;
            const uint32_t kd4 = __ldg(reinterpret_cast<const uint32_t *>(kind + c0));   /* capacity is padded */
            lanemask = 0;
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (((kd4 >> (8 * j)) & KIND_CONTROL) && c0 + j < tv.nclusters) lanemask |= 1u << j;
            lanes = __ballot_sync(FULL_MASK, lanemask != 0);
            continue;
        }
        const int src = __ffs(lanes) - 1;
        uint32_t sm = __shfl_sync(FULL_MASK, lanemask, src);
        const int bit = __ffs(sm) - 1;
        sm &= sm - 1;
        if (lane == src) lanemask = sm;
        if (sm == 0) lanes &= lanes - 1;
        const uint32_t n = chunk * CTL_CHUNK + src * 4 + bit;
        for (int c = lane; c < ncyc; c += 32) {
            const uint32_t b = __ldg(tv.bcl + (size_t)c * tv.bcl_pitch + n);
            rd[c] = b == 0 ? 4 : (uint8_t)(b & 3);
        }
        __syncwarp();

        /* ---- 1. exact substring of read[0 .. L) ---- */
        bool aligned = false;
        {
            bool hasn = false;
            for (int i = lane; i < L; i += 32) hasn |= rd[i] == 4;
            if (!__any_sync(FULL_MASK, hasn)) {
                uint32_t key = 0;
                for (int i = 0; i < plan.hash_k; i++) key |= (uint32_t)rd[i] << (2 * i);
                uint32_t h = ctl_hash_key(key);
                for (;;) {
                    const uint32_t ent = __ldg(ix.hash + h);
                    if (ent == 0) break;
                    bool same = true;
                    for (int i = lane; i < L; i += 32) same &= __ldg(ix.target + ent - 1 + i) == rd[i];
                    if (__all_sync(FULL_MASK, same)) { aligned = true; break; }
                    h = (h + 1) & (CTL_HASH - 1);
                }
            }
        }

        /* ---- 2. local alignment of read[first .. first + L) ---- */
        if (!aligned) {
            const uint8_t *w = rd + first;
            int nn = 0;
            for (int i = lane; i < L; i += 32) nn += w[i] == 4;
            for (int o = 16; o; o >>= 1) nn += __shfl_xor_sync(FULL_MASK, nn, o);
            /* substitution tables of this lane's two rows (rows >= L behave like N) */
            uint32_t tbl[2];
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const int row = 2 * lane + k;
                const int rb = row < L ? w[row] : 4;
                uint32_t t = 0;
                for (int tbase = 0; tbase < 5; tbase++) {
                    const int sc = (rb == 4 || tbase == 4) ? 0 : (rb == tbase ? 1 : -2);
                    t |= (uint32_t)(sc + 2) << (4 * tbase);
                }
                tbl[k] = t;
            }
            if (L - nn < plan.minscore) {
                /* fewer than minscore callable bases (e.g. an all-N read): cannot align */
            } else if (!plan.filter) {
                aligned = window_reaches(ix.target, 0, CTL_TLEN, tbl[0], tbl[1], plan.minscore, lane);
            } else {
                const int q = plan.q;
                for (int i = lane; i < CTL_NBIN / 2; i += 32) bins[i] = 0;
                __syncwarp();
                int skipped = 0;
                for (int r = lane; r <= L - q; r += 32) {
                    uint32_t code = 0;
                    int wild[CTL_MAXWILD] = {0, 0, 0}, nw = 0;
                    for (int k = 0; k < q; k++) {
                        const uint32_t b = w[r + k];
                        if (b == 4) { if (nw < CTL_MAXWILD) wild[nw] = 2 * k; nw++; }
                        else code |= b << (2 * k);
                    }
                    if (nw > CTL_MAXWILD) { skipped++; continue; }
                    for (int sp = 0; sp < (1 << (2 * nw)); sp++) {
                        const uint32_t cd = code | ((uint32_t)(sp & 3) << wild[0]) | ((uint32_t)((sp >> 2) & 3) << wild[1]) |
                                            ((uint32_t)((sp >> 4) & 3) << wild[2]);
                        const int a = __ldg(ix.qoff + cd), b = __ldg(ix.qoff + cd + 1);
                        for (int j = a; j < b; j++) {
                            const int bin = ((int)__ldg(ix.qpos + j) - r + L) / CTL_BINW;
                            atomicAdd(&bins[bin >> 1], 1u << (16 * (bin & 1)));
                        }
                    }
                }
                for (int o = 16; o; o >>= 1) skipped += __shfl_xor_sync(FULL_MASK, skipped, o);
                const int T = max(plan.threshold - skipped, 1);
                __syncwarp();
                /* pair k = bins k and k+1; word i holds bins 2i (low) and 2i+1 (high).  Runs of
                 * consecutive flagged pairs [k0, k1] are evaluated as one window: diagonals
                 * [16 k0, 16 (k1 + 2)), i.e. target columns [16 k0 - L, 16 k1 + 31). */
                int run0 = -1, run1 = -2;
                auto flush = [&]() {
                    if (run0 >= 0 && !aligned) {
                        int t0 = CTL_BINW * run0 - L, t1 = CTL_BINW * run1 + 2 * CTL_BINW - 1;
                        if (t0 < 0) t0 = 0;
                        if (t1 > CTL_TLEN) t1 = CTL_TLEN;
                        if (t1 > t0)
                            aligned = window_reaches(ix.target, t0, t1 - t0, tbl[0], tbl[1], plan.minscore, lane);
                    }
                    run0 = -1;
                };
                for (int i0 = 0; i0 < CTL_NBIN / 2 && !aligned; i0 += 32) {
                    const int i = i0 + lane;
                    const uint32_t cur = bins[i];
                    const uint32_t nxt = (i + 1 < CTL_NBIN / 2) ? bins[i + 1] : 0u;
                    const int lo = cur & 0xffff, hi = cur >> 16, nlo = nxt & 0xffff;
                    unsigned even = __ballot_sync(FULL_MASK, lo + hi >= T);
                    unsigned odd = __ballot_sync(FULL_MASK, hi + nlo >= T);
                    while ((even | odd) && !aligned) {
                        int k;
                        const int je = even ? __ffs(even) - 1 : 64, jo = odd ? __ffs(odd) - 1 : 64;
                        if (je <= jo) { k = 2 * (i0 + je); even &= even - 1; }
                        else { k = 2 * (i0 + jo) + 1; odd &= odd - 1; }
                        if (run0 >= 0 && k == run1 + 1) run1 = k;
                        else { flush(); run0 = run1 = k; }
                    }
                }
                flush();
            }
        }
        __syncwarp();
        if (lane == 0) {
            if (aligned) { calls[n].sample = 1; n_control++; }
            else n_unknown++;
        }
    }
    if (lane == 0) {
        if (n_unknown) atomicAdd(&s_tot[0], n_unknown);
        if (n_control) atomicAdd(&s_tot[1], n_control);
    }
    __syncthreads();
    /* spotanalyzer.c:668-670: both pseudo-samples count under clusters_mm0 */
    if (threadIdx.x < 2 && s_tot[threadIdx.x]) atomicAdd(&counters[threadIdx.x * TSK_NCOUNT + 0], s_tot[threadIdx.x]);
}

/* ---- host: plan and index ---- */
static int wildcard_threshold(int L, int S, int q)
{
    int best = 1 << 20;
    for (int n = 0; n <= L - S; n++) {
        const int room = L - S - n;
        int worst = -(1 << 20);
        for (int X = 0; 3 * X <= room; X++)
            for (int G = 0; 3 * X + 4 * G <= room; G++) {
                const int loss = (q - 3) * X + (q - 5) * G;
                if (loss > worst) worst = loss;
            }
        const int t = S + n - (q - 1) - worst;
        if (t < best) best = t;
    }
    return best;
}

static void control_plan(const tsk_params *p, ControlPlan *pl)
{
    const int L = p->control_read_length;
    const int S = (int)(L * 0.65);                       /* controlaligner.c:162-163 */
    pl->first = p->control_first_cycle;
    pl->len = L;
    pl->minscore = S;
    pl->hash_k = L < 16 ? L : 16;
    /* 6-mers give the fewest index hits; short control reads need 5-mers for a useful threshold */
    pl->q = wildcard_threshold(L, S, 6) >= 6 ? 6 : 5;
    pl->threshold = wildcard_threshold(L, S, pl->q);
    /* side conditions of the filter (see the header comment) */
    pl->filter = (L - S - 2 <= CTL_BINW + 1) && (L - CTL_SPACING < S) && pl->threshold >= 2;
}

int tsk_control_setup(tsk_ctx *ctx)
{
    const tsk_params *p = &ctx->hp;
    const int L = p->control_read_length;
    ControlPlan pl;
    control_plan(p, &pl);
    /* one device buffer: target | hash | qoff | qpos */
    const uint32_t nq = 1u << (2 * pl.q);
    size_t off_hash, off_qoff, off_qpos, total = 0;
    total += (CTL_TLEN + CTL_TPAD + 15) & ~(size_t)15;
    off_hash = total; total += sizeof(uint32_t) * CTL_HASH;
    off_qoff = total; total += ((sizeof(uint16_t) * (nq + 1)) + 15) & ~(size_t)15;
    off_qpos = total; total += ((sizeof(uint16_t) * CTL_TLEN) + 15) & ~(size_t)15;
    uint8_t *blob = (uint8_t *)calloc(total, 1);
    uint32_t *fill = (uint32_t *)calloc(nq, sizeof(uint32_t));
    if (!blob || !fill) { tsk_set_error("out of host memory"); return -1; }
    uint8_t *target = blob;
    uint32_t *hash = (uint32_t *)(blob + off_hash);

    /* forward + 20 N + reverse complement (load_control_sequence, controlaligner.c:74-104) */
    memset(target, 4, CTL_TLEN + CTL_TPAD);
    for (int i = 0; i < TSK_PHIX174_LENGTH; i++) {
        const uint8_t b = (uint8_t)((TSK_PHIX174_2BIT[i / 16] >> (2 * (i % 16))) & 3);
        target[i] = b;
        target[CTL_TLEN - 1 - i] = (uint8_t)(3 - b);
    }
    /* exact path: every length-L substring of either strand, hashed on its first hash_k bases */
    for (int strand = 0; strand < 2; strand++) {
        const int base = strand ? TSK_PHIX174_LENGTH + CTL_SPACING : 0;
        for (int s = 0; s + L <= TSK_PHIX174_LENGTH; s++) {
            uint32_t key = 0;
            for (int i = 0; i < pl.hash_k; i++) key |= (uint32_t)target[base + s + i] << (2 * i);
            uint32_t h = ctl_hash_key(key);
            while (hash[h] != 0) h = (h + 1) & (CTL_HASH - 1);
            hash[h] = (uint32_t)(base + s + 1);
        }
    }
    /* q-gram index of the target (q-grams touching the spacer are left out) */
    {
        const int q = pl.q;
        uint16_t *qoff = (uint16_t *)(blob + off_qoff);
        uint16_t *qpos = (uint16_t *)(blob + off_qpos);
        for (int pass = 0; pass < 2; pass++) {
            for (int t = 0; t + q <= CTL_TLEN; t++) {
                uint32_t code = 0;
                int ok = 1;
                for (int j = 0; j < q; j++) { ok &= target[t + j] != 4; code |= (uint32_t)(target[t + j] & 3) << (2 * j); }
                if (!ok) continue;
                if (pass == 0) qoff[code + 1]++;
                else qpos[qoff[code] + fill[code]++] = (uint16_t)t;
            }
            if (pass == 0)
                for (uint32_t c = 0; c < nq; c++) qoff[c + 1] = (uint16_t)(qoff[c + 1] + qoff[c]);
        }
    }
    TSK_CUDA(cudaMalloc(&ctx->d_ctl_index, total));
    TSK_CUDA(cudaMemcpy(ctx->d_ctl_index, blob, total, cudaMemcpyHostToDevice));
    ctx->ctl_off_hash = off_hash; ctx->ctl_off_qoff = off_qoff; ctx->ctl_off_qpos = off_qpos;
    free(blob); free(fill);
    return 0;
}

/* Classifies the clusters K1 marked KIND_CONTROL, on the side stream ctx->xstream
 * (tsk_launch_import orders it after K1 and joins it). */
int tsk_launch_control(tsk_ctx *ctx)
{
    ControlPlan pl;
    control_plan(&ctx->hp, &pl);
    ControlIndex ix;
    ix.target = ctx->d_ctl_index;
    ix.hash = (const uint32_t *)(ctx->d_ctl_index + ctx->ctl_off_hash);
    ix.qoff = (const uint16_t *)(ctx->d_ctl_index + ctx->ctl_off_qoff);
    ix.qpos = (const uint16_t *)(ctx->d_ctl_index + ctx->ctl_off_qpos);
    /* the queue length lives on the device: a fixed grid of warps pulls reads from it */
    const unsigned want = (ctx->tile.nclusters + CTL_WARPS - 1) / CTL_WARPS;
    const unsigned blocks = want < 148u * 8u ? (want ? want : 1u) : 148u * 8u;
    k_control_classify<<<blocks, CTL_THREADS, 0, ctx->xstream>>>(ctx->tile, pl, ix, ctx->d_kind, ctx->d_ctl_queue,
                                                               ctx->d_calls, ctx->d_counters);
    ctx->launches++;
    TSK_CUDA(cudaGetLastError());
    return 0;
}
