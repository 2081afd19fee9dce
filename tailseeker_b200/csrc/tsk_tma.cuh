/* TMA engine plumbing shared by the kernels that stage their input through shared memory:
 * bulk asynchronous tensor copies global -> shared that complete on an mbarrier, and the host-side
 * encoder of the tensor maps they use. */
#ifndef TSK_TMA_CUH
#define TSK_TMA_CUH

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "tsk_internal.h"

/* All helpers take 32-bit shared-window addresses computed once per thread. */
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_drop(uint32_t bar)
{   /* arrive for the current phase and leave the barrier for all later phases */
    asm volatile("mbarrier.arrive_drop.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    /* try_wait suspends the thread in hardware up to the hint (ns) instead of spinning */
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity), "r"(0x989680u) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
/* one TMA tensor copy: box {32 clusters, 4 channels, 4 cycles} of the CIF tensor -> shared */
__device__ __forceinline__ void tma_cycle_g2s(uint32_t dst, const CUtensorMap *tmap, int x, int cyc, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(tmap), "r"(x), "r"(0), "r"(cyc), "r"(bar) : "memory");
}
/* one TMA tensor copy of a 2-D box at (x, y) -> shared */
__device__ __forceinline__ void tma_2d_g2s(uint32_t dst, const CUtensorMap *tmap, int x, int y, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(tmap), "r"(x), "r"(y), "r"(bar) : "memory");
}

/* cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda at link time).
 * NOTE for callers: with CU_TENSOR_MAP_INTERLEAVE_NONE a box must START on a 16-byte boundary of
 * the innermost dimension; a copy from any other coordinate faults as "illegal instruction". */
typedef CUresult (*tsk_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                        const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);

static inline tsk_encode_tiled_fn tsk_tensor_map_encoder(void)
{
    static tsk_encode_tiled_fn encode = NULL;
    if (!encode) {
        void *fn = NULL;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess || !fn) {
            tsk_set_error("cuTensorMapEncodeTiled unavailable");
            return NULL;
        }
        encode = (tsk_encode_tiled_fn)fn;
    }
    return encode;
}

#endif
