/* The ruler's histogram bin (tsk_ruler.cu), in a header of its own so that tests/exact_check.cpp can
 * compile the same source for the host (-DTSK_HOST_CHECK) and compare its integer form with the
 * reference's fp64 expression for every packed score. */
#ifndef TSK_SCOREBIN_CUH
#define TSK_SCOREBIN_CUH

#ifdef TSK_HOST_CHECK
#include "tsk_hostcheck.h"
#endif
#include <stdint.h>
#include "../../include/tailseeker_b200.h"

/* floor((sc - 1) * bins / 16777214) of add_polya_score_sample() (polyaruler.c:163-168). */
__device__ __forceinline__ int score_bin(uint32_t sc, int bins, int intbins)
{
    int b;
    if (intbins) {
        /* in integers: with x = a*2^24 + b0 and 2^24 = 16777214 + 2, x = a*16777214 + (2a + b0),
         * 2a + b0 < 2*16777214.  Equals the reference's fp64 expression whenever bins shares no
         * factor with 2^23-1 (the host checks): the exact value is then an integer only for
         * sc-1 in {0, 8388607, 16777214}, which fp64 evaluates exactly, and is otherwise
         * >= 2^-24 away from one, far beyond the fp64 rounding error. */
        const unsigned long long x = (unsigned long long)(sc - 1) * (unsigned)bins;
        const uint32_t a = (uint32_t)(x >> 24), b0 = (uint32_t)x & 0xffffffu;
        b = (int)(a + ((2 * a + b0) >= 16777214u ? 1u : 0u));
    } else {
        b = __double2int_rz(__dmul_rn(__ddiv_rn(__dsub_rn((double)sc, 1.0), (double)(TSK_SCORE_MAX - 1)),
                                      (double)bins));
    }
    return b >= bins ? bins - 1 : b;
}

#endif
