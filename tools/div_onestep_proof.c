/*
 * div_onestep_proof.c -- CPU evidence that K2's division helper is exact with ONE residual correction:
 *     q0 = RN(a*y),  r = a - b*q0 (one FMA, exact),  q = RN(q0 + r*y),   y = RN(1/b)
 * equals the IEEE quotient RN(a/b)  (tailseeker_b200/csrc/tsk_score.cu: div_by_rcp; it replaces the
 * nine divisions per cycle of signalproc.c:246-266,329).
 *
 *   gcc -O2 -mfma -ffp-contract=off -o div_onestep_proof tools/div_onestep_proof.c -lm -lpthread
 *   ./div_onestep_proof            # about three minutes on 8 cores
 *
 * 1. hard cases, exhaustively: q0 + r*y misses a/b by < 1.5 * 2^-24 ulp, so only quotients that close
 *    to a rounding midpoint X/2^S (X odd) can round differently.  With A, B the 24-bit significands,
 *    a/b - X/2^S = k / (2^S * B) where X*B + k = 2^S * A: for EVERY B and every k = k' * 2^ctz(B),
 *    |k'| <= 33, all solutions (X, A) are generated and tested, in both quotient binades.
 *    The same search with y off by one ulp reports failures (control, so the search is known to bite).
 * 2. the 2^31 operand pairs of the GPU self-test (k_div_sweep modes 0 and 1, same hash, same seeds,
 *    tests/test_gpu_parity.py::test_fast_arithmetic_sweeps), replayed on the CPU.
 * 3. random and structured pairs (divisor significands near 1 and near 2, short significands).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define NT 8
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float onestep(float a, float b, float y)
{
    const float q = a * y, r = fmaf(-b, q, a);
    return fmaf(r, y, q);
}
static uint64_t inv_odd(uint64_t b) { uint64_t x = b; for (int i = 0; i < 6; i++) x *= 2 - b * x; return x; }

struct job { int id, what; uint64_t n, tested, bad, bad_control; };

static void hard_cases(struct job *j)
{
    for (uint64_t B = (1u << 23) + j->id; B < (1u << 24); B += NT) {
        const int t = __builtin_ctzll(B);
        const uint64_t iB = inv_odd(B >> t);
        for (int S = 24; S <= 25; S++) {
            if (S - t < 2) continue;
            const uint64_t M = 1ull << (S - t);
            for (int kk = -33; kk <= 33; kk++) {
                if (kk == 0) continue;
                const uint64_t X0 = ((uint64_t)(-(int64_t)kk) * iB) & (M - 1);
                for (uint64_t X = X0; X < (1ull << 25); X += M) {
                    if (X <= (1ull << 24) || !(X & 1)) continue;
                    const int64_t num = (int64_t)(X * B) + (int64_t)kk * (int64_t)(1ull << t);
                    if (num <= 0 || (num & ((1ll << S) - 1))) continue;
                    const uint64_t A = (uint64_t)num >> S;
                    if (A < (1u << 23) || A >= (1u << 24)) continue;
                    const float a = ldexpf((float)A, -10), b = ldexpf((float)B, -7), y = 1.0f / b, q = a / b;
                    j->tested++;
                    j->bad += f2u(onestep(a, b, y)) != f2u(q);
                    j->bad_control += f2u(onestep(a, b, u2f(f2u(y) + 1))) != f2u(q);
                }
            }
        }
    }
}

static void gpu_pairs(struct job *j, uint64_t seed, int mode)
{
    for (uint64_t i = j->id; i < j->n; i += NT) {
        uint64_t x = (i + seed) * 0x9E3779B97F4A7C15ull;
        x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 32;
        const uint32_t ab = (uint32_t)x, bb = (uint32_t)(x >> 32);
        float a, b;
        if (mode == 0) {
            b = u2f(((127 - 40 + (bb >> 23) % 80) << 23) | (bb & 0x7fffffu));
            a = u2f((ab & 0x80000000u) | ((127 - 60 + ((ab >> 23) & 0xff) % 120) << 23) | (ab & 0x7fffffu));
        } else {
            b = u2f(((124 + (bb >> 23) % 6) << 23) | (bb & 0x7fffffu));
            a = b * (u2f(0x3f800000u | (ab & 0x7fffffu)) - 1.0f);
        }
        const float q = onestep(a, b, 1.0f / b), want = a / b;
        j->tested++;
        j->bad += f2u(q) != f2u(want) && !(q == 0.f && want == 0.f);
    }
}

static void random_pairs(struct job *j, int mode)
{
    uint64_t s = 0x9e3779b97f4a7c15ull * (j->id + 1) + mode;
    for (uint64_t i = 0; i < j->n; i++) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        const uint64_t r = s;
        uint32_t mb = (uint32_t)(r & 0x7fffff), ma = (uint32_t)((r >> 23) & 0x7fffff);
        int eb, ea;
        if (mode == 0) { eb = 127 + (int)((r >> 46) % 120) - 60; ea = eb - (int)((r >> 54) % 40); }
        else if (mode == 1) {
            const uint32_t k = (uint32_t)((r >> 46) & 0x3ff);
            mb = ((r >> 56) & 1) ? (0x7fffff - k) : k;
            eb = 127 + (int)((r >> 57) % 60) - 30; ea = eb - (int)((r >> 60) & 7);
        } else { ma &= 0x7ff000; mb &= 0x7ff800; eb = 127; ea = 127 - (int)((r >> 46) & 15); }
        const float b = u2f(((uint32_t)eb << 23) | mb), a = u2f(((uint32_t)ea << 23) | ma);
        j->tested++;
        j->bad += f2u(onestep(a, b, 1.0f / b)) != f2u(a / b);
    }
}

static void *run(void *p)
{
    struct job *j = p;
    if (j->what == 0) hard_cases(j);
    else if (j->what == 1) gpu_pairs(j, 12345, 0);
    else if (j->what == 2) gpu_pairs(j, 98765, 1);
    else random_pairs(j, j->what - 3);
    return NULL;
}

int main(int argc, char **argv)
{
    static const char *names[6] = {"hard cases (all divisor significands, |k'| <= 33)",
                                   "GPU self-test pairs, mode 0", "GPU self-test pairs, mode 1",
                                   "random pairs", "divisors next to 1 and 2", "short significands"};
    const uint64_t nrandom = argc > 1 ? strtoull(argv[1], NULL, 10) : 500000000ull;
    int failed = 0;
    for (int what = 0; what < 6; what++) {
        pthread_t t[NT];
        struct job j[NT];
        uint64_t tested = 0, bad = 0, ctl = 0;
        for (int i = 0; i < NT; i++) {
            j[i] = (struct job){i, what, what == 1 || what == 2 ? (1ull << 30) : nrandom, 0, 0, 0};
            pthread_create(&t[i], NULL, run, &j[i]);
        }
        for (int i = 0; i < NT; i++) { pthread_join(t[i], NULL); tested += j[i].tested; bad += j[i].bad; ctl += j[i].bad_control; }
        printf("%-52s %12llu quotients, %llu differ from a / b", names[what], (unsigned long long)tested,
               (unsigned long long)bad);
        if (what == 0) printf(" (control with y one ulp off: %llu differ)", (unsigned long long)ctl);
        printf("\n");
        failed |= bad != 0 || (what == 0 && ctl == 0);
    }
    return failed;
}
