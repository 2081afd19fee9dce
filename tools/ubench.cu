// Instruction-throughput microbenchmarks for the design choices of normalise_score_pack
// (B200, sm_100a).  Reports warp-instructions-equivalent lane-ops per clock per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o ubench tools/ubench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

template <typename Op>
__global__ void __launch_bounds__(256) k_bench(Op op, float *out, float seed)
{
    typename Op::T v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) v[i] = Op::init(seed + threadIdx.x * 0.001f + i);
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) v[i] = op(v[i]);
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; i++) acc += Op::fin(v[i]);
    if (acc == 12345.678f) out[0] = acc;
}

struct OpFmul  { using T = float;  static __device__ T init(float s) { return s; } static __device__ float fin(T v) { return v; }
                 __device__ T operator()(T v) const { return __fmul_rn(v, 1.0000001f); } };
struct OpFfma  { using T = float;  static __device__ T init(float s) { return s; } static __device__ float fin(T v) { return v; }
                 __device__ T operator()(T v) const { return __fmaf_rn(v, 1.0000001f, 0.5f); } };
struct OpFdiv  { using T = float;  static __device__ T init(float s) { return s; } static __device__ float fin(T v) { return v; }
                 __device__ T operator()(T v) const { return __fdiv_rn(1.5f, v) + 1.0f; } };
struct OpFrcp  { using T = float;  static __device__ T init(float s) { return s; } static __device__ float fin(T v) { return v; }
                 __device__ T operator()(T v) const { return __frcp_rn(v) + 1.0f; } };
struct OpMufuRcp { using T = float; static __device__ T init(float s) { return s; } static __device__ float fin(T v) { return v; }
                 __device__ T operator()(T v) const { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; } };
struct OpDfma  { using T = double; static __device__ T init(float s) { return s; } static __device__ float fin(T v) { return (float)v; }
                 __device__ T operator()(T v) const { return __fma_rn(v, 1.0000001, 0.5); } };
struct OpDmul  { using T = double; static __device__ T init(float s) { return s; } static __device__ float fin(T v) { return (float)v; }
                 __device__ T operator()(T v) const { return __dmul_rn(v, 1.0000001); } };
struct OpDadd  { using T = double; static __device__ T init(float s) { return s; } static __device__ float fin(T v) { return (float)v; }
                 __device__ T operator()(T v) const { return __dadd_rn(v, 1.0000001); } };
struct OpDdiv  { using T = double; static __device__ T init(float s) { return s; } static __device__ float fin(T v) { return (float)v; }
                 __device__ T operator()(T v) const { return __ddiv_rn(1.5, v) + 1.0; } };
struct OpDrcp  { using T = double; static __device__ T init(float s) { return s; } static __device__ float fin(T v) { return (float)v; }
                 __device__ T operator()(T v) const { return __drcp_rn(v) + 1.0; } };
struct OpF2D2F { using T = float;  static __device__ T init(float s) { return s; } static __device__ float fin(T v) { return v; }
                 __device__ T operator()(T v) const { double d = (double)v; d = __dadd_rn(d, 1e-9); return (float)d; } };
struct OpF2D   { using T = float;  static __device__ T init(float s) { return s; } static __device__ float fin(T v) { return v; }
                 __device__ T operator()(T v) const { double d = (double)v; return __double2hiint(d) ? __int_as_float(__double2loint(d) ^ __double2hiint(d)) : v; } };
struct OpI2F   { using T = float;  static __device__ T init(float s) { return s; } static __device__ float fin(T v) { return v; }
                 __device__ T operator()(T v) const { return (float)(__float_as_int(v) >> 9); } };
struct OpF2I   { using T = float;  static __device__ T init(float s) { return s; } static __device__ float fin(T v) { return v; }
                 __device__ T operator()(T v) const { return __int_as_float(__float2int_rz(v) + 0x3f800000); } };
struct OpI2D   { using T = double; static __device__ T init(float s) { return s; } static __device__ float fin(T v) { return (float)v; }
                 __device__ T operator()(T v) const { return (double)(__double2hiint(v) >> 5); } };
struct OpLogfMufu { using T = float; static __device__ T init(float s) { return s; } static __device__ float fin(T v) { return v; }
                 __device__ T operator()(T v) const { return __logf(v) + 2.0f; } };

template <typename Op>
void run(const char *name, Op op, int clock_khz)
{
    float *out;
    cudaMalloc(&out, 4);
    const int blocks = 148 * 8, threads = 256;
    k_bench<<<blocks, threads>>>(op, out, 1.5f);
    cudaDeviceSynchronize();
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k_bench<<<blocks, threads>>>(op, out, 1.5f);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double ops = (double)blocks * threads * ITERS * ILP;
    const double clk = ms * 1e-3 * clock_khz * 1e3;
    printf("%-12s %8.3f ms  %7.2f lane-ops/clk/SM (at %d MHz)\n", name, ms, ops / clk / 148.0, clock_khz / 1000);
    cudaFree(out);
}

int main()
{
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    run("FMUL", OpFmul(), khz);
    run("FFMA", OpFfma(), khz);
    run("fdiv_rn(+add)", OpFdiv(), khz);
    run("frcp_rn(+add)", OpFrcp(), khz);
    run("MUFU.RCP", OpMufuRcp(), khz);
    run("DFMA", OpDfma(), khz);
    run("DMUL", OpDmul(), khz);
    run("DADD", OpDadd(), khz);
    run("ddiv_rn(+add)", OpDdiv(), khz);
    run("drcp_rn(+add)", OpDrcp(), khz);
    run("F2D+DADD+D2F", OpF2D2F(), khz);
    run("F2D(+int)", OpF2D(), khz);
    run("I2F(+shift)", OpI2F(), khz);
    run("F2I(+iadd)", OpF2I(), khz);
    run("I2D(+shift)", OpI2D(), khz);
    run("__logf+add", OpLogfMufu(), khz);
    return 0;
}
