/*
 * tsk_kernels.cu -- sm_100a kernels of the importer path (decode / demultiplex / seed,
 * slot scan, normalise / score / pack, fair-sampling fix-ups).
 *
 * One thread per cluster.  Inputs stay in the files' cycle-major layouts, so consecutive
 * lanes read consecutive bytes (BCL) or int16 (CIF planes): every request is coalesced.
 * K1 stages its CTA's [cycles][128 clusters] slab of base calls in shared memory with TMA
 * tensor copies first: the per-cluster scans then read shared memory, not a global byte per step.
 * All per-cluster state lives in registers; nothing here is a dense contraction, so no
 * tensor cores.  Compiled with -fmad=false: the reference is built without FMA
 * contraction and parity of the packed scores needs separately rounded IEEE operations.
 *
 * Reference behaviour restated here (cited per function): src/importer/spotanalyzer.c,
 * signalproc.c, findpolya.c, bclreader.c of hyeshik/tailseeker 3.2.1.
 */
#include <math_constants.h>
#include "tsk_internal.h"
#include "tsk_device.cuh"
#include "tsk_tma.cuh"

#define FULL_MASK 0xffffffffu

/* =======================================================================================
 * K1  decode_demux_seed : BCL bytes -> per-cluster call, kind, fair-sampling bucket
 *   format_basecalls (bclreader.c:146-166), assign_barcode (spotanalyzer.c:67-103),
 *   counters + fingerprint + balancer QC + delimiter (spotanalyzer.c:648-715,106-203),
 *   find_polya (findpolya.c:35-84), check_balance_minimum (signalproc.c:106-129),
 *   fair-sampling hash (spotanalyzer.c:460-474)
 * ===================================================================================== */
__device__ __forceinline__ int mismatch3(uint32_t a, uint32_t b)
{
    uint32_t x = a ^ b;
    x = (x | (x >> 1) | (x >> 2)) & 0x09249249u;
    return __popc(x);
}

template <int N> struct tsk_int { static constexpr int value = N; };

__global__ void __launch_bounds__(TSK_K1_THREADS)
k_decode_demux_seed(const __grid_constant__ DevParams P, const TileView tv,
                    const DevSample *__restrict__ g_samples,
                    tsk_call *__restrict__ calls, uint8_t *__restrict__ kind,
                    uint32_t *__restrict__ bucket, uint32_t *__restrict__ blkcnt,
                    uint32_t *__restrict__ counters, uint32_t *__restrict__ fair,
                    const __grid_constant__ CUtensorMap bclmap, const int slab_rows)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int S = P.num_samples;
    /* [slab_rows][128] base-call bytes of this CTA's clusters | mbarrier | samples | counters */
    const uint8_t *s_bcl = smem_raw;
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem_raw + (size_t)slab_rows * TSK_K1_THREADS);
    DevSample *s_samples = reinterpret_cast<DevSample *>(s_bar + 2);
    uint32_t *s_cnt = reinterpret_cast<uint32_t *>(s_samples + S);                   /* [S][6] */
    uint32_t *s_wcnt = s_cnt + S * TSK_NCOUNT;       /* [warps][S+1] */
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = TSK_K1_THREADS / 32;
    /* per base-call BYTE: poly(A) weight of its base * 512 - 1, the step of the packed
     * (running sum, position) key of the seed scan's long stretch */
    int32_t *s_wkey = reinterpret_cast<int32_t *>(s_wcnt + nwarps * (S + 1));

    if (tid == 0) {
        const uint32_t bar = smem_u32(s_bar);
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_arrive_expect_tx(bar, (uint32_t)slab_rows * TSK_K1_THREADS);
        for (int r = 0; r < slab_rows; r += TSK_K1_SLAB)
            tma_2d_g2s(smem_u32(s_bcl) + r * TSK_K1_THREADS, &bclmap, (int)(blockIdx.x * TSK_K1_THREADS), r, bar);
    }

    for (int i = tid; i < (int)(sizeof(DevSample) / 4) * S; i += TSK_K1_THREADS)
        reinterpret_cast<uint32_t *>(s_samples)[i] = reinterpret_cast<const uint32_t *>(g_samples)[i];
    for (int i = tid; i < S * TSK_NCOUNT + nwarps * (S + 1); i += TSK_K1_THREADS)
        s_cnt[i] = 0;
    if (P.w_small)
        for (int b = tid; b < 256; b += TSK_K1_THREADS) s_wkey[b] = P.w_polyA[tsk_base_code(b)] * 512 - 1;
    __syncthreads();

    const uint32_t n = blockIdx.x * TSK_K1_THREADS + tid;
    const bool active = n < tv.nclusters;
    mbar_wait(smem_u32(s_bar), 0);              /* init is visible: the __syncthreads() above */
    const uint8_t *col = s_bcl + tid;
    auto rd = [&](int c) -> uint32_t { return (uint32_t)col[c * TSK_K1_THREADS]; };

    int sample = 0, flags = 0, de = -1, status = -1, mods = -1;
    int written = 0;
    uint32_t kd = 0, bk = 0;
    bool to_control = false;     /* unassigned and a control is configured: k_control_classify decides */

    if (active) {
        const int ts = P.threep_start, tl = P.threep_length;
        bool dropped = false;

        /* ---- barcode ---- */
        uint32_t rpk[4] = {0, 0, 0, 0};
#pragma unroll
        for (int w = 0; w < 4; w++) {
            if (w < P.index_words) {
                const int k0 = w * 10;
                const int kn = min(10, P.index_length - k0);
                uint32_t acc = 0;
                for (int k = 0; k < kn; k++)
                    acc |= tsk_base_code(rd(P.index_start + k0 + k)) << (3 * k);
                rpk[w] = acc;
            }
        }
        int best = -1, bestmm = P.index_length + 1;
        bool tie = false;
        for (int s = 1; s < S; s++) {
            int mm = 0;
#pragma unroll
            for (int w = 0; w < 4; w++)
                if (w < P.index_words) mm += mismatch3(rpk[w], s_samples[s].index_pk[w]);
            if (mm < bestmm) { best = s; bestmm = mm; tie = false; }
            else if (mm == bestmm) tie = true;
        }
        int mismatches;
        if (best < 0 || tie || bestmm > s_samples[best].max_index_mm) { mismatches = -1; sample = 0; }
        else { mismatches = bestmm; sample = best; }
        const DevSample &sm = s_samples[sample];

        to_control = mismatches < 0 && P.has_control;
        if (to_control) { /* counted as Unknown or control once classified (spotanalyzer.c:652-670) */ }
        else if (mismatches <= 0) atomicAdd(&s_cnt[sample * TSK_NCOUNT + 0], 1u);
        else {
            flags |= TSK_PAFLAG_BARCODE_HAS_MISMATCHES;
            atomicAdd(&s_cnt[sample * TSK_NCOUNT + (mismatches == 1 ? 1 : 2)], 1u);
        }

        /* ---- fingerprint ---- */
        if (sm.fp_len > 0) {
            int mm = 0;
            for (int k = 0; k < sm.fp_len; k++)
                mm += tsk_base_code(rd(sm.fp_pos + k)) != sm.fingerprint[k];
            if (mm > sm.max_fp_mm) {
                atomicAdd(&s_cnt[sample * TSK_NCOUNT + 4], 1u);
                dropped = true;
            }
        }

        /* ---- balancer base-call quality (whole-read indices) ---- */
        if (!dropped && sm.has_umi > 0 && P.balancer_min_bases_passes > 0) {
            int pass = 0;
            for (int j = P.balancer_start; j < P.balancer_start + P.balancer_length; j++)
                pass += (int)tsk_base_qual(rd(j)) >= P.balancer_min_quality;
            if (pass < P.balancer_min_bases_passes) {
                flags |= TSK_PAFLAG_BALANCER_CALL_QUALITY_BAD;
                atomicAdd(&s_cnt[sample * TSK_NCOUNT + 5], 1u);
                if (!P.keep_low_quality_balancer) dropped = true;
            }
        }

        if (!dropped) {
            written = 1;
            if (sm.delimiter_length > 0) {
                /* ---- delimiter: offsets 0, -1, +1 ---- */
                int found = -1, foundshift = 0, foundmm = 0;
#pragma unroll
                for (int t = 0; t < 3; t++) {
                    const int shift = (t == 0) ? 0 : (t == 1 ? -1 : 1);
                    if (found < 0) {
                        int nd = 0;
                        for (int k = 0; k < sm.delimiter_length; k++) {
                            const uint32_t code = tsk_base_code(rd(sm.delimiter_pos + shift + k));
                            nd += !((sm.delim_mask[k] >> code) & 1u);
                        }
                        if (nd <= sm.max_delim_mm) { found = t; foundshift = shift; foundmm = nd; }
                    }
                }
                if (found < 0) {
                    flags |= TSK_PAFLAG_DELIMITER_NOT_FOUND;
                    atomicAdd(&s_cnt[sample * TSK_NCOUNT + 3], 1u);
                    if (!P.keep_no_delimiter) written = 0;
                } else {
                    if (foundmm > 0) flags |= TSK_PAFLAG_DELIMITER_HAS_MISMATCH;
                    if (foundshift != 0) flags |= TSK_PAFLAG_DELIMITER_IS_SHIFTED;
                    de = sm.delimiter_pos + sm.delimiter_length + foundshift;

                    /* ---- sequence seed: find_polya as one backward pass ----
                     * score(i,j) = V(i) + S(i) - S(j+1), S(k) = sum_{m>=k} polyA_w[seq m],
                     * V(i) = sum_{m<i} nonA_w[seq m]; the reference's (i asc, j asc, strict >)
                     * scan ends on the lexicographically first maximum, which is: per i the
                     * smallest j >= i+minlen-1 minimising S(j+1), then the first i maximising. */
                    const int L = ts + tl - de;
                    const int M = min(P.max_mods, L);
                    const int mlen = max(P.min_polya_len, 1);
                    int bi = -1, bj = -1, bscore = -1;
                    int Srun = 0, minS = 0x3fffffff, argj = -1;
                    const int captop = M + mlen - 1;        /* j below this may be a j_min(i) */
                    int j = L - 1;
                    if (P.w_small && j >= captop) {
                        /* the long stretch above the captured region only carries the running sum
                         * and its minimum.  Both live in ONE register: key = sum * 512 + position
                         * (positions < 512, |sum| < 2^15), so "smaller sum, then smaller position" --
                         * the update rule `sum <= min` of a scan that walks downwards -- is a plain
                         * integer minimum, and a step adds weight * 512 - 1, looked up by the raw
                         * base-call byte: five instructions a cycle (two shared loads, address,
                         * minimum, add) instead of ten (base code, byte-permuted weight, compare,
                         * select, two adds). */
                        int key = j, minkey = INT_MAX;         /* Srun = 0 so far */
#pragma unroll 4
                        for (; j >= captop; j--) {
                            minkey = min(minkey, key);
                            key += s_wkey[rd(de + j)];
                        }
                        minS = minkey >> 9; argj = minkey & 511;
                        Srun = (key - j) >> 9;                 /* key == Srun * 512 + j, exactly */
                    }
                    if (mlen - 1 <= 8) {
                        /* Candidate start i = j is complete at step j: S(i) has just been formed, and
                         * (min, arg min) of S(j'+1) over j' >= i + mlen - 1 was taken mlen - 1 steps
                         * earlier -- kept in an eight-deep shift register instead of arrays in local
                         * memory.  With Vsuf(i) = sum of the non-A weights of positions i..M-1,
                         * V(i) = V(M) - Vsuf(i), so the start is chosen on g(i) = S(i) - min - Vsuf(i);
                         * descending i with >= ends on the smallest maximiser, like the ascending
                         * strict > of the reference. */
                        const int DL = mlen - 1;
                        int bestg = INT_MIN, Vsuf = 0;
                        /* one instance per register depth: with the depth a compile-time constant
                         * the tap is a plain register and only DL - 1 moves are left (the
                         * run-time select over eight taps was 11 % of this kernel's instructions
                         * and 21 % of its stall samples) */
                        auto lower = [&](auto depth) {
                            constexpr int D = decltype(depth)::value;
                            int dS[D > 0 ? D : 1], dJ[D > 0 ? D : 1];
#pragma unroll
                            for (int k = 0; k < (D > 0 ? D : 1); k++) { dS[k] = 0; dJ[k] = -1; }
                            for (; j >= 0; j--) {
                                const uint32_t code = tsk_base_code(rd(de + j));
                                if (Srun <= minS) { minS = Srun; argj = j; }
                                const int cS = D > 0 ? dS[D > 0 ? D - 1 : 0] : minS, cJ = D > 0 ? dJ[D > 0 ? D - 1 : 0] : argj;
#pragma unroll
                                for (int k = (D > 0 ? D : 1) - 1; k > 0; k--) { dS[k] = dS[k - 1]; dJ[k] = dJ[k - 1]; }
                                dS[0] = minS; dJ[0] = argj;
                                Srun += P.w_polyA[code];
                                if (j < M) {
                                    Vsuf += P.w_nonA[code];
                                    if (j + D <= L - 1) {
                                        const int g = Srun - cS - Vsuf;
                                        if (g >= bestg) { bestg = g; bi = j; bj = cJ; }
                                    }
                                }
                            }
                        };
                        switch (DL) {
                        case 0: lower(tsk_int<0>{}); break;
                        case 1: lower(tsk_int<1>{}); break;
                        case 2: lower(tsk_int<2>{}); break;
                        case 3: lower(tsk_int<3>{}); break;
                        case 4: lower(tsk_int<4>{}); break;
                        case 5: lower(tsk_int<5>{}); break;
                        case 6: lower(tsk_int<6>{}); break;
                        case 7: lower(tsk_int<7>{}); break;
                        default: lower(tsk_int<8>{}); break;
                        }
                        if (bi >= 0) bscore = Vsuf + bestg;       /* Vsuf is V(M) now */
                    } else {
                        int candS[TSK_MAX_MODS], candJ[TSK_MAX_MODS], sfull[TSK_MAX_MODS];
                        uint8_t codes[TSK_MAX_MODS];
                        for (; j >= 0; j--) {
                            const uint32_t code = tsk_base_code(rd(de + j));
                            if (Srun <= minS) { minS = Srun; argj = j; }
                            if (j < captop) {
                                const int i = j - (mlen - 1);
                                if (i >= 0 && i < M) { candS[i] = minS; candJ[i] = argj; }
                            }
                            Srun += P.w_polyA[code];
                            if (j < M) { sfull[j] = Srun; codes[j] = (uint8_t)code; }
                        }
                        int V = 0;
                        for (int i = 0; i < M; i++) {
                            if (i + mlen - 1 <= L - 1) {
                                const int sc = V + sfull[i] - candS[i];
                                if (sc > bscore) { bscore = sc; bi = i; bj = candJ[i]; }
                            }
                            V += P.w_nonA[codes[i]];
                        }
                    }
                    int ps = 0, plen = 0;
                    if (bi >= 0 && bscore >= 1 && (bj - bi + 1) >= P.min_polya_len) {
                        ps = bi; plen = bj + 1 - bi;
                    }
                    mods = ps;
                    if (ps > 0) flags |= TSK_PAFLAG_HAVE_3P_MODIFICATION;
                    if (plen > 0) flags |= TSK_PAFLAG_POLYA_DETECTED;

                    /* ---- balancer composition ---- */
                    int blen = de - ts;
                    if (blen > P.balancer_length) blen = P.balancer_length;
                    uint32_t comp = 0;          /* four byte counters A,C,G,T */
                    for (int c = P.balancer_start; c < P.balancer_start + blen; c++) {
                        const uint32_t code = tsk_base_code(rd(ts + c));
                        if (code < 4) comp += 1u << (8 * code);
                    }
                    const int mo = P.balancer_minimum_occurrence;
                    if ((int)(comp & 255) < mo || (int)((comp >> 8) & 255) < mo ||
                        (int)((comp >> 16) & 255) < mo || (int)(comp >> 24) < mo) {
                        flags |= TSK_PAFLAG_BALANCER_BIASED;
                        status = -1;
                    } else {
                        status = plen;
                        kd = KIND_PROCESS;
                        int insert_len = L;
                        if (sm.limit_threep > 0 && sm.limit_threep < insert_len)
                            insert_len = sm.limit_threep;
                        const int scan_len = insert_len - ps;
                        if (scan_len > 0) {
                            if (plen >= P.seed_trigger) kd |= KIND_POS;
                            else if (plen <= P.neg_polya_len && scan_len >= P.fair_fp_len) {
                                /* 63-bit rolling hash of the first bases after the seed start */
                                unsigned long long v = 0;
                                for (int k = 0; k < P.fair_fp_len; k++) {
                                    const uint32_t code = tsk_base_code(rd(de + ps + k));
                                    /* 'A','C','G','T','N' & 7 */
                                    const unsigned long long ch = (0x64731ull >> (4 * code)) & 7ull;
                                    v = ((v << 3) | (ch ^ (v >> 60))) & 0x7fffffffffffffffull;
                                }
                                bk = (uint32_t)(v % (unsigned long long)P.fair_hash_size);
                                kd |= KIND_NEG;
                                atomicAdd(&fair[bk], 1u);
                            }
                            if (plen >= P.sigproc_trigger) kd |= KIND_EMIT;
                        }
                    }
                }
            }
        }
    }

    if (to_control) kd |= KIND_CONTROL;     /* k_control_classify picks these up from kind[] */

    /* ---- order-preserving ranks inside the CTA: record slots per sample, work items ---- */
    const int ekey = (kd & KIND_EMIT) ? sample : -1;
    const unsigned lt = (1u << lane) - 1u;
    const unsigned peers = __match_any_sync(FULL_MASK, ekey);
    int erank = __popc(peers & lt);
    if (ekey >= 0 && erank == 0) s_wcnt[warp * (S + 1) + ekey] = __popc(peers);
    const unsigned pb = __ballot_sync(FULL_MASK, (kd & KIND_PROCESS) != 0);
    int prank = __popc(pb & lt);
    if (lane == 0) s_wcnt[warp * (S + 1) + S] = __popc(pb);
    __syncthreads();
    for (int w = 0; w < warp; w++) {
        if (ekey >= 0) erank += s_wcnt[w * (S + 1) + ekey];
        prank += s_wcnt[w * (S + 1) + S];
    }
    if (tid <= S) {
        uint32_t tot = 0;
        for (int w = 0; w < nwarps; w++) tot += s_wcnt[w * (S + 1) + tid];
        blkcnt[(size_t)blockIdx.x * (S + 1) + tid] = tot;
    }
    for (int i = tid; i < S * TSK_NCOUNT; i += TSK_K1_THREADS)
        if (s_cnt[i]) atomicAdd(&counters[i], s_cnt[i]);

    if (active) {
        tsk_call c;
        c.sample = (int16_t)sample;
        c.flags = (uint16_t)flags;
        c.delimiter_end = (int16_t)de;
        c.polya_len = (int16_t)status;
        c.terminal_mods = (int16_t)mods;
        c.written = (uint8_t)written;
        c.emitted = (kd & KIND_EMIT) ? 1 : 0;
        c.record_slot = ((uint32_t)prank << 16) | (uint32_t)erank;   /* CTA-local, see k_assign */
        *reinterpret_cast<uint4 *>(&calls[n]) = *reinterpret_cast<const uint4 *>(&c);
        kind[n] = (uint8_t)kd;
        bucket[n] = bk;
    }
}

/* Exclusive scan of the per-CTA counts, one CTA per column (sample 0..S-1 = record slots,
 * column S = work items). */
#define SCAN_THREADS 1024
#define SCAN_CHUNK   16          /* entries per thread and pass, all loads in flight at once */
__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_blocks(uint32_t *__restrict__ blkcnt, uint32_t *__restrict__ totals, int nblk, int ncol)
{
    __shared__ uint32_t s_warp[SCAN_THREADS / 32];
    __shared__ uint32_t s_carry;
    const int col = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    /* the column is walked in passes of SCAN_THREADS * SCAN_CHUNK entries (one pass for a tile of up
     * to 2 M clusters); thread t owns SCAN_CHUNK consecutive entries of the pass */
    for (int base = 0; base < nblk; base += SCAN_THREADS * SCAN_CHUNK) {
        const int b0 = base + tid * SCAN_CHUNK;
        uint32_t v[SCAN_CHUNK], sum = 0;
#pragma unroll
        for (int k = 0; k < SCAN_CHUNK; k++) {
            v[k] = (b0 + k < nblk) ? blkcnt[(size_t)(b0 + k) * ncol + col] : 0u;
            sum += v[k];
        }
        uint32_t incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const uint32_t w = s_warp[lane];
            uint32_t wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= d) wi += o;
            }
            s_warp[lane] = wi - w;                       /* exclusive prefix of the warp sums */
        }
        __syncthreads();
        uint32_t run = s_carry + s_warp[warp] + (incl - sum);
#pragma unroll
        for (int k = 0; k < SCAN_CHUNK; k++) {
            if (b0 + k < nblk) blkcnt[(size_t)(b0 + k) * ncol + col] = run;
            run += v[k];
        }
        __syncthreads();
        if (tid == SCAN_THREADS - 1) s_carry = run;      /* the last thread ends on the pass total */
        __syncthreads();
    }
    if (tid == 0) totals[col] = s_carry;
}

__global__ void k_record_bases(const uint32_t *__restrict__ totals, const DevSample *__restrict__ samples,
                               uint64_t *__restrict__ recbase, int S)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        uint64_t at = 0;
        for (int s = 0; s < S; s++) {
            recbase[s] = at;
            at += (uint64_t)totals[s] * (uint64_t)(2 + samples[s].dump_len);
        }
        recbase[S] = at;
    }
}

__global__ void __launch_bounds__(256)
k_assign(const TileView tv, tsk_call *__restrict__ calls, const uint8_t *__restrict__ kind,
         const uint32_t *__restrict__ blkbase, uint32_t *__restrict__ worklist, int S)
{
    const uint32_t n = blockIdx.x * 256 + threadIdx.x;
    if (n >= tv.nclusters) return;
    const uint32_t kd = kind[n];
    const uint32_t tmp = calls[n].record_slot;
    const size_t row = (size_t)(n / TSK_K1_THREADS) * (S + 1);
    uint32_t slot = 0;
    if (kd & KIND_EMIT) slot = blkbase[row + calls[n].sample] + (tmp & 0xffffu);
    if (kd & KIND_PROCESS) worklist[blkbase[row + S] + (tmp >> 16)] = n;
    calls[n].record_slot = slot;
}

/* ===================================================================================== */
int tsk_launch_import(tsk_ctx *ctx, cudaEvent_t *ev)
{
    const DevParams &P = ctx->dp;
    const TileView &tv = ctx->tile;
    const int S = P.num_samples;
    const uint32_t N = tv.nclusters;
    const int nblk = (int)((N + TSK_K1_THREADS - 1) / TSK_K1_THREADS);
    cudaStream_t st = ctx->stream;
    const size_t histbytes = sizeof(uint32_t) * (size_t)P.total_cycles * P.bins;
    TSK_CUDA(cudaMemsetAsync(ctx->d_pos, 0, histbytes, st));
    TSK_CUDA(cudaMemsetAsync(ctx->d_neg, 0, histbytes, st));
    TSK_CUDA(cudaMemsetAsync(ctx->d_fair, 0, sizeof(uint32_t) * (size_t)P.fair_hash_size, st));
    TSK_CUDA(cudaMemsetAsync(ctx->d_lists, 0, sizeof(uint32_t) * 16, st));
    const bool control = ctx->hp.has_control != 0;
    if (control) TSK_CUDA(cudaMemsetAsync(ctx->d_ctl_queue, 0, sizeof(uint32_t) * 4, st));
    if (ev) TSK_CUDA(cudaEventRecord(ev[0], st));

    if (N > 0) {
        TskRange nvtx("tsk:decode_demux_seed");
        /* base calls u8 [T][pitch] as a 2-D tensor: boxes of {128 clusters, TSK_K1_SLAB cycles};
         * rows past T and clusters past N are zero-filled (= no-calls, never looked at) */
        const int slab_rows = (P.total_cycles + TSK_K1_SLAB - 1) / TSK_K1_SLAB * TSK_K1_SLAB;
        tsk_encode_tiled_fn encode = tsk_tensor_map_encoder();
        if (!encode) return -1;
        if ((uintptr_t)tv.bcl % 16 != 0 || tv.bcl_pitch % 16 != 0) {
            tsk_set_error("base-call rows must be 16-byte aligned"); return -1;
        }
        CUtensorMap bclmap;
        {
            const cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)P.total_cycles};
            const cuuint64_t strides[1] = {(cuuint64_t)tv.bcl_pitch};
            const cuuint32_t box[2] = {TSK_K1_THREADS, TSK_K1_SLAB}, estr[2] = {1, 1};
            const CUresult r = encode(&bclmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void *)tv.bcl, dims, strides, box, estr,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { tsk_set_error("cuTensorMapEncodeTiled (base calls) failed (%d)", (int)r); return -1; }
        }
        const size_t smem1 = (size_t)slab_rows * TSK_K1_THREADS + 16 + sizeof(DevSample) * S +
                             sizeof(uint32_t) * (S * TSK_NCOUNT + (TSK_K1_THREADS / 32) * (S + 1)) + 256 * sizeof(int32_t);
        /* the attribute is per device: remembered in the context, not in a process-wide static */
        if (smem1 > ctx->k1_smem_configured) {
            TSK_CUDA(cudaFuncSetAttribute(k_decode_demux_seed, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
            ctx->k1_smem_configured = smem1;
        }
        k_decode_demux_seed<<<nblk, TSK_K1_THREADS, smem1, st>>>(
            P, tv, ctx->d_samples, ctx->d_calls, ctx->d_kind, ctx->d_bucket, ctx->d_blkcnt,
            ctx->d_counters, ctx->d_fair, bclmap, slab_rows);
        ctx->launches++;
    }
    if (ev) TSK_CUDA(cudaEventRecord(ev[1], st));
    if (N > 0 && control) {
        /* The classifier only touches calls[].sample of unassigned clusters and the two
         * pseudo-sample counters, which nothing below reads: it runs on a side stream next to
         * the scan and normalise/score/pack kernels and is joined before the fix-ups. */
        TSK_CUDA(cudaEventRecord(ctx->ev_ctl[0], st));
        TSK_CUDA(cudaStreamWaitEvent(ctx->xstream, ctx->ev_ctl[0], 0));
        { TskRange nvtx("tsk:control_classify"); if (tsk_launch_control(ctx) != 0) return -1; }
        TSK_CUDA(cudaEventRecord(ctx->ev_ctl[1], ctx->xstream));
    }
    if (N > 0) {
        TskRange nvtx("tsk:slot_scan_assign");
        k_scan_blocks<<<S + 1, SCAN_THREADS, 0, st>>>(ctx->d_blkcnt, ctx->d_totals, nblk, S + 1);
        k_record_bases<<<1, 32, 0, st>>>(ctx->d_totals, ctx->d_samples, ctx->d_recbase, S);
        k_assign<<<(N + 255) / 256, 256, 0, st>>>(tv, ctx->d_calls, ctx->d_kind, ctx->d_blkcnt,
                                                  ctx->d_worklist, S);
        ctx->launches += 3;
    } else {
        TSK_CUDA(cudaMemsetAsync(ctx->d_totals, 0, sizeof(uint32_t) * (S + 1), st));
        TSK_CUDA(cudaMemsetAsync(ctx->d_recbase, 0, sizeof(uint64_t) * (S + 1), st));
    }
    if (ev) TSK_CUDA(cudaEventRecord(ev[2], st));
    if (tsk_ruler_join(ctx) != 0) return -1;     /* the previous tile's ruler still reads the record slots */
    { TskRange nvtx("tsk:normalise_score_pack"); if (tsk_launch_score(ctx) != 0) return -1; }
    if (N > 0 && control) TSK_CUDA(cudaStreamWaitEvent(st, ctx->ev_ctl[1], 0));
    if (ev) TSK_CUDA(cudaEventRecord(ev[3], st));
    { TskRange nvtx("tsk:fair_sampling_fixups"); if (tsk_launch_fixups(ctx) != 0) return -1; }
    if (ev) TSK_CUDA(cudaEventRecord(ev[4], st));
    /* nothing after this point reads the tile's base calls or intensities: its upload slot may be
     * refilled (tsk_tile_upload waits for this event on the copy stream) */
    if (ctx->tile_slot >= 0) {
        TSK_CUDA(cudaEventRecord(ctx->ev_done[ctx->tile_slot], st));
        ctx->slot_busy[ctx->tile_slot] = 1;
    }
    TSK_CUDA(cudaGetLastError());
    return 0;
}
