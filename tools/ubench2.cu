// Issue-slot microbenchmarks for packed FP32 (FFMA2 / FADD2) on B200 (sm_100a): does a packed
// instruction cost one issue slot for two lane-operations, and what does it leave for the
// integer / FP64 / conversion instructions it is mixed with in normalise_score_pack?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o ubench2.bin tools/ubench2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

typedef unsigned long long u64;
#define ITERS 2048

__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(u64 v) { float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a + b; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float fadd(float a, float b) { float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ uint32_t lop(uint32_t a, uint32_t b) { uint32_t r; asm volatile("xor.b32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ double dfma(double a, double b, double c) { double r; asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(r) : "d"(a), "d"(b), "d"(c)); return r; }
__device__ __forceinline__ float i2f(int a) { float r; asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(r) : "r"(a)); return r; }
__device__ __forceinline__ double f2d(float a) { double r; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(r) : "f"(a)); return r; }
__device__ __forceinline__ float d2f(double a) { float r; asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(r) : "d"(a)); return r; }

// NF scalar FFMA, NP packed FFMA2, NA packed FADD2, NI integer ops, ND DFMA, NC f32->f64->f32 round trips per iteration,
// eight independent chains each
template <int NF, int NP, int NA, int NI, int ND, int NC>
__global__ void __launch_bounds__(256) k_mix(float *out, float seed, u64 negzero)
{
    float f[8]; u64 p[8]; uint32_t x[8]; double d[8]; float c[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        f[i] = seed + threadIdx.x * 0.001f + i; p[i] = pk(f[i], f[i] + 1.f); x[i] = threadIdx.x * 77 + i; d[i] = f[i]; c[i] = f[i] * 0.5f;
    }
    const u64 m = pk(1.0000001f, 0.9999999f);
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (k < NF) f[k] = ffma(f[k], 1.0000001f, 0.5f);
            if (k < NP) p[k] = fma2(p[k], m, negzero);
            if (k < NA) p[(k + 4) & 7] = add2(p[(k + 4) & 7], m);
            if (k < NI) x[k] = lop(x[k], 0x9e3779b9u + it);
            if (k < ND) d[k] = dfma(d[k], 1.0000001, 0.5);
            if (k < NC) c[k] = d2f(f2d(c[k]));
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) acc += f[i] + lo(p[i]) + (float)x[i] + (float)d[i] + c[i];
    if (acc == 12345.678f) out[0] = acc;
}

template <int NF, int NP, int NA, int NI, int ND, int NC>
void run(const char *name, int khz)
{
    float *out; cudaMalloc(&out, 4);
    const int blocks = 148 * 8, threads = 256;
    k_mix<NF, NP, NA, NI, ND, NC><<<blocks, threads>>>(out, 1.5f, 0x8000000080000000ull);
    cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k_mix<NF, NP, NA, NI, ND, NC><<<blocks, threads>>>(out, 1.5f, 0x8000000080000000ull);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double warps_per_smsp = (double)blocks * threads / 32 / 148 / 4;
    const double clk = ms * 1e-3 * khz * 1e3;
    const int ninst = NF + NP + NA + NI + ND + 2 * NC;
    printf("%-44s %7.3f ms  %6.2f clk per warp-iteration (%d instr -> %.2f instr/clk/SMSP)\n", name, ms,
           clk / ITERS / warps_per_smsp, ninst, ninst / (clk / ITERS / warps_per_smsp));
    cudaFree(out);
}

int main()
{
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    printf("clock %d MHz\n", khz / 1000);
    run<8, 0, 0, 0, 0, 0>("8 FFMA", khz);
    run<0, 8, 0, 0, 0, 0>("8 FFMA2", khz);
    run<0, 0, 8, 0, 0, 0>("8 FADD2", khz);
    run<0, 4, 4, 0, 0, 0>("4 FFMA2 + 4 FADD2", khz);
    run<0, 0, 0, 8, 0, 0>("8 LOP3", khz);
    run<8, 0, 0, 8, 0, 0>("8 FFMA + 8 LOP3", khz);
    run<0, 4, 0, 8, 0, 0>("4 FFMA2 + 8 LOP3 (same flops as above)", khz);
    run<0, 8, 0, 8, 0, 0>("8 FFMA2 + 8 LOP3", khz);
    run<0, 0, 0, 0, 8, 0>("8 DFMA", khz);
    run<8, 0, 0, 0, 4, 0>("8 FFMA + 4 DFMA", khz);
    run<0, 4, 0, 0, 4, 0>("4 FFMA2 + 4 DFMA", khz);
    run<8, 0, 0, 8, 4, 0>("8 FFMA + 8 LOP3 + 4 DFMA", khz);
    run<0, 4, 0, 8, 4, 0>("4 FFMA2 + 8 LOP3 + 4 DFMA", khz);
    run<0, 0, 0, 0, 0, 8>("8 x (F2F.F64.F32 + F2F.F32.F64)", khz);
    run<8, 0, 0, 8, 4, 2>("8 FFMA + 8 LOP3 + 4 DFMA + 2 cvt pairs", khz);
    run<0, 4, 0, 8, 4, 2>("4 FFMA2 + 8 LOP3 + 4 DFMA + 2 cvt pairs", khz);
    return 0;
}
