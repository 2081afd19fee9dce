// CPU check of the scoring kernels' exact-arithmetic helpers, compiled from THEIR OWN SOURCE
// (tailseeker_b200/csrc/tsk_exact.cuh, tsk_device.cuh through tsk_hostcheck.h):
//   div_by_rcp(a, b, RN(1/b))  ==  a / b                      (signalproc.c:246-266,329 divisions)
//   tsk_logf, logf_core_t, logf_ki_t  ==  the host's logf     (shannon_entropy, signalproc.c:39-51)
// g++ -O2 -mfma -ffp-contract=off -DTSK_HOST_CHECK -I tailseeker_b200/csrc
//   score_bin(sc, bins, integer form) == min(bins - 1, (int)(((sc - 1.) / 16777214) * bins)) in fp64
//                                          for every packed score (polyaruler.c:163-168)
//     usage: exact_check [stride over the floats in (0,1], default 1 = every one] [division pairs]
//                        [stride over the divisor significands of the hard-case search, default 17]
// Hard cases of the division (see tools/div_onestep_proof.c): the quotients nearest to a rounding
// midpoint, X*B + k = 2^S * A with X odd and |k| <= 33 * 2^ctz(B), for the chosen divisors B.
#include <cstdio>
#include <cstdlib>
#include <pthread.h>
#include "tsk_exact.cuh"
#include "tsk_scorebin.cuh"

#define NT 8
static double2 TAB[16], KI[LOGF_KI_ENTRIES];
struct job { int id; uint32_t stride, bstride; uint64_t pairs, n, bad_logf[3], bad_div, hard, bad_hard, nbins, bad_bins; };

static uint64_t inv_odd(uint64_t b) { uint64_t x = b; for (int i = 0; i < 6; i++) x *= 2 - b * x; return x; }

static void hard_cases(job *j)
{
    for (uint64_t B = (1u << 23) + (uint64_t)j->id * j->bstride; B < (1u << 24); B += (uint64_t)NT * j->bstride) {
        const int t = __builtin_ctzll(B);
        const uint64_t iB = inv_odd(B >> t);
        for (int S = 24; S <= 25; S++) {
            if (S - t < 2) continue;
            const uint64_t M = 1ull << (S - t);
            for (int kk = -33; kk <= 33; kk++) {
                if (kk == 0) continue;
                for (uint64_t X = ((uint64_t)(-(int64_t)kk) * iB) & (M - 1); X < (1ull << 25); X += M) {
                    if (X <= (1ull << 24) || !(X & 1)) continue;
                    const int64_t num = (int64_t)(X * B) + (int64_t)kk * (int64_t)(1ull << t);
                    if (num <= 0 || (num & ((1ll << S) - 1))) continue;
                    const uint64_t A = (uint64_t)num >> S;
                    if (A < (1u << 23) || A >= (1u << 24)) continue;
                    const float a = ldexpf((float)A, -10), b = ldexpf((float)B, -7);
                    j->hard++;
                    j->bad_hard += __float_as_uint(div_by_rcp(a, b, 1.0f / b)) != __float_as_uint(a / b);
                }
            }
        }
    }
}

static void *run(void *p)
{
    job *j = (job *)p;
    const double2 *tab = TAB, *ki = KI;
    for (uint64_t ix = 0x00800000u + (uint64_t)j->id * j->stride; ix <= 0x3f800000u; ix += (uint64_t)NT * j->stride) {
        const uint32_t u = (uint32_t)ix;
        const uint32_t want = __float_as_uint(logf(__uint_as_float(u)));
        j->n++;
        j->bad_logf[0] += __float_as_uint(tsk_logf(__uint_as_float(u), tab)) != want;
        j->bad_logf[1] += __float_as_uint(logf_core_t(u, [tab](uint32_t ofs) { return tab[ofs >> 4]; })) != want;
        j->bad_logf[2] += __float_as_uint(logf_ki_t(u, [ki](uint32_t ofs) { return ki[ofs >> 4]; })) != want;
    }
    // the operand pairs of the GPU self-test (k_div_sweep, both modes), a stretch per thread
    for (int mode = 0; mode < 2; mode++)
        for (uint64_t i = j->id; i < j->pairs; i += NT) {
            uint64_t x = (i + (mode ? 98765 : 12345)) * 0x9E3779B97F4A7C15ull;
            x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 32;
            const uint32_t ab = (uint32_t)x, bb = (uint32_t)(x >> 32);
            float a, b;
            if (mode == 0) {
                b = __uint_as_float(((127 - 40 + (bb >> 23) % 80) << 23) | (bb & 0x7fffffu));
                a = __uint_as_float((ab & 0x80000000u) | ((127 - 60 + ((ab >> 23) & 0xff) % 120) << 23) | (ab & 0x7fffffu));
            } else {
                b = __uint_as_float(((124 + (bb >> 23) % 6) << 23) | (bb & 0x7fffffu));
                a = __fmul_rn(b, __uint_as_float(0x3f800000u | (ab & 0x7fffffu)) - 1.0f);
            }
            const float q = div_by_rcp(a, b, 1.0f / b), want = a / b;
            j->bad_div += __float_as_uint(q) != __float_as_uint(want) && !(q == 0.f && want == 0.f);
        }
    hard_cases(j);
    // every packed score 1 .. 2^24 - 1, for bin counts the launcher sends down the integer form
    // (tsk_ruler.cu: bins <= 65536 sharing no factor with 2^23 - 1 = 47 * 178481)
    static const int BINS[] = {1000, 470 + 1, 64, 30000, 65536, 1, 2, 3, 999, 1024, 4096, 46, 48, 12345};
    for (unsigned b = j->id; b < sizeof(BINS) / sizeof(BINS[0]); b += NT)
        for (uint32_t sc = 1; sc <= TSK_SCORE_MAX; sc++) {
            const int bins = BINS[b];
            int want = (int)((((double)sc - 1.) / 16777214) * bins);
            if (want >= bins) want = bins - 1;
            j->nbins++;
            j->bad_bins += score_bin(sc, bins, 1) != want || score_bin(sc, bins, 0) != want;
        }
    return NULL;
}

int main(int argc, char **argv)
{
    const uint32_t stride = argc > 1 ? (uint32_t)strtoul(argv[1], NULL, 10) : 1;
    const uint64_t pairs = argc > 2 ? strtoull(argv[2], NULL, 10) : (1ull << 30);
    for (int i = 0; i < 16; i++) TAB[i] = make_double2(c_logf_tab[i][0], c_logf_tab[i][1]);
    for (int e = 0; e < LOGF_KI_ENTRIES; e++)             // as k_fill_logki fills the device table
        KI[e] = logf_ki_entry(e, c_logf_tab[e & 15][0], c_logf_tab[e & 15][1]);
    // what the kernels rely on for the bit patterns outside (0, 1]: +0 gives a finite value
    {
        const double2 *ki = KI;
        const float z = logf_ki_t(0u, [ki](uint32_t ofs) { return ki[ofs >> 4]; });
        if (!(z == z) || z - z != 0.f) { printf("logf_ki_t(+0) is not finite\n"); return 1; }
        for (uint32_t u = 0; u < 0x7f800000u; u += 9973)          // any non-NaN pattern: in bounds and finite
            if (!(logf_ki_t(u, [ki](uint32_t ofs) { return ki[ofs >> 4]; }) - 0.f == logf_ki_t(u, [ki](uint32_t ofs) { return ki[ofs >> 4]; }))) { printf("logf_ki_t(%08x) is NaN\n", u); return 1; }
    }
    const uint32_t bstride = argc > 3 ? (uint32_t)strtoul(argv[3], NULL, 10) : 17;
    pthread_t t[NT];
    job j[NT];
    uint64_t n = 0, bl[3] = {0, 0, 0}, bd = 0, hard = 0, bh = 0, nb = 0, bb = 0;
    for (int i = 0; i < NT; i++) { j[i] = job(); j[i].id = i; j[i].stride = stride ? stride : 1; j[i].pairs = pairs; j[i].bstride = bstride ? bstride : 1; pthread_create(&t[i], NULL, run, &j[i]); }
    for (int i = 0; i < NT; i++) {
        pthread_join(t[i], NULL);
        n += j[i].n; bd += j[i].bad_div; hard += j[i].hard; bh += j[i].bad_hard; nb += j[i].nbins; bb += j[i].bad_bins;
        for (int v = 0; v < 3; v++) bl[v] += j[i].bad_logf[v];
    }
    printf("{\"floats\": %llu, \"tsk_logf\": %llu, \"logf_core_t\": %llu, \"logf_ki_t\": %llu, \"division_pairs\": %llu, \"div_by_rcp\": %llu, \"hard_quotients\": %llu, \"div_by_rcp_hard\": %llu, \"score_bins\": %llu, \"score_bin\": %llu}\n",
           (unsigned long long)n, (unsigned long long)bl[0], (unsigned long long)bl[1], (unsigned long long)bl[2],
           (unsigned long long)(2 * pairs), (unsigned long long)bd, (unsigned long long)hard, (unsigned long long)bh,
           (unsigned long long)nb, (unsigned long long)bb);
    return (bl[0] || bl[1] || bl[2] || bd || bh || bb) ? 1 : 0;
}
