/*
 * logf_variants.c -- CPU sweep over EVERY normal float in (0, 1] of cheaper forms of K2's logf
 * (tailseeker_b200/csrc/tsk_score.cu: logf_core_t, the host libm's table algorithm in fp64 FMAs,
 * shannon_entropy in signalproc.c:39-51).  Prepared for the next round: none of these is in the
 * default build yet; V2 is in the kernels behind -DTSK_LOGF_KI=1 (TSK_NVCC_EXTRA), waiting for its
 * first GPU run.
 *
 *   gcc -O2 -mfma -ffp-contract=off -o logf_variants tools/logf_variants.c -lm -lpthread && ./logf_variants
 *
 *   kernel   y0 = fma((double)k, ln2, logc[i]); r = fma(z, invc[i], -1); ...        8 fp64 operations
 *   V1       y0 = KLN2[k] + logc[i]           (k*ln2 rounded once, from a 1 KB table)   7, no int->double
 *   V2       {invc[i] * 2^-k, y0} from one (k, i) table entry; r = fma((double)x, ., -1)
 *            6 fp64 operations and no mantissa extraction (-7 of 23 instructions per logarithm);
 *            the table holds 16 entries of 16 bytes per k, so a kernel keeps k >= -KMIN in shared
 *            memory and sends the rare smaller p down the present code.
 * All three must give the float the host's logf gives.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <pthread.h>

static const double T[16][2] = {       /* {invc, logc}: glibc e_logf_data.c, SURVEY.md appendix C */
    {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2}, {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2},
    {0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2},  {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3},
    {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3}, {0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4}, {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4},
    {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5}, {0x1p+0, 0x0p+0},
    {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5},  {0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4},
    {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3},
    {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},  {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2},
};
static const double LN2 = 0x1.62e42fefa39efp-1, A1 = 0x1.5575b0be00b6ap-2, A2 = -0x1.ffffef20a4123p-2,
                    A0 = -0x1.00ea348b88334p-2;
static double KLN2[128];                /* [(-k)]            */
static double KI[128 * 16][2];          /* [(k & 127) * 16 + i] = {invc * 2^-k, fma(k, ln2, logc)} */

static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float tail(double r, double y0)
{
    const double r2 = r * r;
    double y = fma(A1, r, A2);
    y = fma(A0, r2, y);
    return (float)fma(y, r2, y0 + r);
}
static inline float kernel_form(uint32_t ix)
{
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (tmp >> 19) & 15, k = (int32_t)tmp >> 23;
    const double z = (double)u2f(ix - (tmp & 0xff800000u));
    return tail(fma(z, T[i][0], -1.0), fma((double)k, LN2, T[i][1]));
}
static inline float v1(uint32_t ix)
{
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (tmp >> 19) & 15, k = (int32_t)tmp >> 23;
    const double z = (double)u2f(ix - (tmp & 0xff800000u));
    return tail(fma(z, T[i][0], -1.0), KLN2[-k] + T[i][1]);
}
static inline float v2(uint32_t ix)
{
    const uint32_t tmp = ix - 0x3f330000u;
    const double *e = KI[(tmp >> 19) & 0x7ff];          /* the kernel's byte offset: (tmp >> 15) & 0x7ff0 */
    return tail(fma((double)u2f(ix), e[0], -1.0), e[1]);
}

/* V2 as the kernel has it behind TSK_LOGF_KI: 16 rows (k = -15 .. 0), byte offset (tmp >> 15) & 0xff0,
 * used for LOGF_KI_FIRST <= ix <= 0x3f800000; below that the kernel form runs */
#define LOGF_KI_FIRST (0x3f330000u - (15u << 23))
static double KI16[16 * 16][2];
static inline float v2_kernel(uint32_t ix)
{
    if (ix - 1u < LOGF_KI_FIRST - 1u) return kernel_form(ix);
    const uint32_t tmp = ix - 0x3f330000u;
    const double *e = KI16[((tmp >> 15) & 0xff0u) >> 4];
    return tail(fma((double)u2f(ix), e[0], -1.0), e[1]);
}

struct job { int id; uint64_t bad[4]; };
static void *run(void *p)
{
    struct job *j = p;
    for (uint64_t ix = 0x00800000u + j->id; ix <= 0x3f800000u; ix += 8) {
        const uint32_t want = f2u(logf(u2f((uint32_t)ix)));
        j->bad[0] += f2u(kernel_form((uint32_t)ix)) != want;
        j->bad[1] += f2u(v1((uint32_t)ix)) != want;
        j->bad[2] += f2u(v2((uint32_t)ix)) != want;
        j->bad[3] += f2u(v2_kernel((uint32_t)ix)) != want;
    }
    return NULL;
}

int main(void)
{
    pthread_t t[8];
    struct job j[8];
    uint64_t bad[4] = {0, 0, 0, 0};
    for (int k = 0; k < 128; k++) KLN2[k] = (double)(-k) * LN2;
    for (int k = -127; k <= 0; k++)
        for (int i = 0; i < 16; i++) {
            KI[(k & 127) * 16 + i][0] = ldexp(T[i][0], -k);
            KI[(k & 127) * 16 + i][1] = fma((double)k, LN2, T[i][1]);
        }
    for (int e = 0; e < 256; e++) {                      /* as load_score_shared() fills sh.logki */
        const int k = (e >> 4) ? (e >> 4) - 16 : 0, i = e & 15;
        KI16[e][0] = scalbn(T[i][0], -k);
        KI16[e][1] = fma((double)k, LN2, T[i][1]);
    }
    for (int i = 0; i < 8; i++) { memset(&j[i], 0, sizeof(j[i])); j[i].id = i; pthread_create(&t[i], NULL, run, &j[i]); }
    for (int i = 0; i < 8; i++) { pthread_join(t[i], NULL); for (int v = 0; v < 4; v++) bad[v] += j[i].bad[v]; }
    printf("normal floats in (0,1]: %u; differing from the host logf: kernel form %llu, V1 %llu, V2 %llu, "
           "V2 as compiled under TSK_LOGF_KI %llu\n", 0x3f800000u - 0x00800000u + 1, (unsigned long long)bad[0],
           (unsigned long long)bad[1], (unsigned long long)bad[2], (unsigned long long)bad[3]);
    return bad[0] || bad[1] || bad[2] || bad[3];
}
