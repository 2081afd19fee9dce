/* Device helpers: arithmetic that must reproduce the host (x86-64, glibc) results bit for bit.
 *
 * The reference is built without FMA contraction (src/Makefile:3: -O3, no -march), so every
 * float expression is a sequence of separately rounded IEEE operations.  The library is
 * compiled with -fmad=false; fused operations below are spelled explicitly where they are
 * provably harmless.
 */
#ifndef TSK_DEVICE_CUH
#define TSK_DEVICE_CUH

#ifdef TSK_HOST_CHECK
#include "tsk_hostcheck.h"            /* CPU checks of this very source (tests/exact_check.cpp) */
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>

/* {invc, logc} table of the logf the host libm implements (glibc >= 2.28,
 * sysdeps/ieee754/flt-32/e_logf.c + e_logf_data.c: the ARM optimized-routines design,
 * LOGF_TABLE_BITS = 4, cubic polynomial, evaluated in double).  Constants from
 * SURVEY.md appendix C, where the recipe was checked against the host logf on every normal
 * float in (0,1].  tests/test_gpu_parity.py::test_logf_exact re-checks it on the GPU box. */
__device__ __constant__ double c_logf_tab[16][2] = {
    {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2}, {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2},
    {0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2},  {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3},
    {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3}, {0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4}, {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4},
    {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5}, {0x1p+0, 0x0p+0},
    {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5},  {0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4},
    {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3},
    {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},  {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2},
};

/* logf(x) for finite x > 0, bit-identical to the host libm.  `tab` points at a copy of
 * c_logf_tab in shared memory (per-lane indices would serialise in the constant cache). */
__device__ __forceinline__ float tsk_logf(float x, const double2 *__restrict__ tab)
{
    uint32_t ix = __float_as_uint(x);
    if (ix == 0x3f800000u)
        return 0.0f;
    if (ix < 0x00800000u) {                 /* subnormal: scale by 2^23 like the host */
        ix = __float_as_uint(__fmul_rn(x, 0x1p23f));
        ix -= 23u << 23;
    }
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (tmp >> 19) & 15;
    const int k = (int32_t)tmp >> 23;
    const uint32_t iz = ix - (tmp & 0xff800000u);
    const double2 t = tab[i];
    const double z = (double)__uint_as_float(iz);
    const double r = __dadd_rn(__dmul_rn(z, t.x), -1.0);
    const double y0 = __dadd_rn(t.y, __dmul_rn((double)k, 0x1.62e42fefa39efp-1));
    const double r2 = __dmul_rn(r, r);
    double y = __dadd_rn(__dmul_rn(0x1.5575b0be00b6ap-2, r), -0x1.ffffef20a4123p-2);
    y = __dadd_rn(__dmul_rn(-0x1.00ea348b88334p-2, r2), y);
    y = __dadd_rn(__dmul_rn(y, r2), __dadd_rn(y0, r));
    return (float)y;
}

/* base code of a BCL byte: 0-3 = A,C,G,T, 4 = no-call 'N' (bclreader.c:153-162) */
__device__ __forceinline__ uint32_t tsk_base_code(uint32_t b) { return b ? (b & 3u) : 4u; }
__device__ __forceinline__ uint32_t tsk_base_qual(uint32_t b) { return b ? (b >> 2) : 2u; }

#endif
