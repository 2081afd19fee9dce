/* TEST SUPPORT: lets tsk_device.cuh / tsk_exact.cuh compile for the host (g++ -DTSK_HOST_CHECK
 * -mfma -ffp-contract=off) so that tests/exact_check.cpp can exercise the kernels' exact-arithmetic
 * helpers from their real source on a machine without a GPU.  Every intrinsic maps onto the IEEE
 * operation it denotes (separately rounded * and +, fused fma); nothing here is used by the library. */
#ifndef TSK_HOSTCHECK_H
#define TSK_HOSTCHECK_H
#ifndef TSK_HOST_CHECK
#error "tsk_hostcheck.h is for -DTSK_HOST_CHECK builds of the CPU checks only"
#endif
#include <math.h>
#include <stdint.h>
#include <string.h>

#define __device__
#define __constant__ static const
#define __forceinline__ static inline
#define __restrict__

struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 d = {x, y}; return d; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline double __ddiv_rn(double a, double b) { return a / b; }
/* scalbn: libm's */
static inline int __double2int_rz(double a) { return (int)a; }
static inline double __hiloint2double(int hi, int lo)
{
    const uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double d;
    memcpy(&d, &u, 8);
    return d;
}
#endif
