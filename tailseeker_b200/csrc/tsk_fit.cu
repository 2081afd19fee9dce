/*
 * tsk_fit.cu -- fit_cutoffs kernel: per-cycle equal-odds point of the positive / negative
 * score distributions.
 *
 * One CTA per cycle.  Restates find_eqodds_point() / find_all_eqodds_point()
 * (scripts/calculate-optimal-parameters.py:57-78) with the weighted gaussian_kde of
 * tailseeker/stats.py:411-466,540-558 (bandwidth = scalar factor) and scipy's bisect
 * (scipy/optimize/Zeros/bisect.c: halve, keep the side where f has the sign of f(a);
 * stop when f == 0 or |dm| < xtol + rtol*|xm|; xtol = 2e-12, rtol = 4*eps, 100 iterations).
 * All arithmetic is fp64; sums are fixed-order tree reductions, so results are
 * deterministic, and agree with the NumPy pairwise sums to ~1e-15 relative (the printed
 * precision is 1e-6).
 */
#include <math_constants.h>
#include "tsk_internal.h"

#define FIT_THREADS 256
#define FIT_MAX_BINS 4096

__device__ __forceinline__ double block_sum(double v, double *s_red)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.;
    for (int w = 0; w < FIT_THREADS / 32; w++) t += s_red[w];
    return t;
}

struct Kde { double inv_cov, norm; };

__device__ Kde kde_setup(const unsigned long long *__restrict__ counts, int bins, double step,
                         double bw, double *w, double *s_red, double *total_out)
{
    double part = 0.;
    for (int j = threadIdx.x; j < bins; j += FIT_THREADS) part += (double)counts[j];
    const double total = block_sum(part, s_red);
    *total_out = total;
    double p2 = 0., pm = 0.;
    for (int j = threadIdx.x; j < bins; j += FIT_THREADS) {
        const double wj = (double)counts[j] / total;
        w[j] = wj;
        p2 += wj * wj;
        pm += wj * ((double)j * step);
    }
    const double sumw2 = block_sum(p2, s_red);
    const double mean = block_sum(pm, s_red);
    double pc = 0.;
    for (int j = threadIdx.x; j < bins; j += FIT_THREADS) {
        const double r = (double)j * step - mean;
        pc += (r * w[j]) * r;
    }
    const double cov = block_sum(pc, s_red) / (1. - sumw2);
    Kde k;
    k.inv_cov = (1. / cov) / (bw * bw);
    k.norm = sqrt(2. * CUDART_PI * (cov * (bw * bw)));
    return k;
}

__device__ double kde_eval(const Kde &k, const double *w, int bins, double step, double x, double *s_red)
{
    double part = 0.;
    for (int j = threadIdx.x; j < bins; j += FIT_THREADS) {
        const double d = x - (double)j * step;
        const double m = sqrt(d * k.inv_cov * d);
        part += exp(-.5 * (m * m)) * w[j];
    }
    return block_sum(part, s_red) / k.norm;
}

__global__ void __launch_bounds__(FIT_THREADS)
k_fit_cutoffs(const unsigned long long *__restrict__ pos, const unsigned long long *__restrict__ neg,
              int bins, unsigned long long min_spots, double bw, double low,
              const double h0, const double h1, const double h2, const double h3,
              const double h4, const double h5, const double h6, const double h7, int nhigh,
              double *__restrict__ out)
{
    extern __shared__ double s_w[];          /* [2][bins] weights */
    __shared__ double s_red[FIT_THREADS / 32];
    const int cyc = blockIdx.x;
    const double highs[8] = {h0, h1, h2, h3, h4, h5, h6, h7};
    const double step = 1. / (double)bins;
    double *wp = s_w, *wn = s_w + bins;
    double tp, tn;
    const Kde kp = kde_setup(pos + (size_t)cyc * bins, bins, step, bw, wp, s_red, &tp);
    const Kde kn = kde_setup(neg + (size_t)cyc * bins, bins, step, bw, wn, s_red, &tn);
    double result = CUDART_NAN;
    if (tp >= (double)min_spots && tn >= (double)min_spots) {
        auto f = [&](double x) -> double {
            const double a = kde_eval(kp, wp, bins, step, x, s_red);
            const double b = kde_eval(kn, wn, bins, step, x, s_red);
            return log(fmax(1e-3, a) / fmax(1e-3, b));
        };
        const double fa0 = f(low);
        for (int h = 0; h < nhigh; h++) {
            double xa = low, xb = highs[h], fa = fa0;
            const double fb = f(xb);
            double r;
            if (!(fa == fa) || !(fb == fb)) continue;     /* degenerate KDE */
            if (fa * fb > 0.) continue;                   /* scipy: ValueError -> next bracket */
            if (fa == 0.) r = xa;
            else if (fb == 0.) r = xb;
            else {
                double dm = xb - xa;
                r = xa;
                for (int it = 0; it < 100; it++) {
                    dm *= .5;
                    const double xm = xa + dm;
                    const double fm = f(xm);
                    if (fm * fa >= 0.) xa = xm;
                    r = xm;
                    if (fm == 0. || fabs(dm) < 2e-12 + 8.881784197001252e-16 * fabs(xm)) break;
                }
            }
            if (low < r && r < highs[h]) { result = r; break; }
        }
    }
    if (threadIdx.x == 0) out[cyc] = result;
}

int tsk_launch_fit(tsk_ctx *ctx, const uint64_t *d_pos, const uint64_t *d_neg, int cycles, int bins,
                   uint64_t min_spots, double bw, double low, const double *high, int nhigh, double *d_out)
{
    TskRange nvtx("tsk:fit_cutoffs");
    if (bins > FIT_MAX_BINS) { tsk_set_error("too many bins for fit_cutoffs"); return -1; }
    double h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < nhigh; i++) h[i] = high[i];
    const size_t smem = sizeof(double) * 2 * (size_t)bins;
    TSK_CUDA(cudaFuncSetAttribute(k_fit_cutoffs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_fit_cutoffs<<<cycles, FIT_THREADS, smem, ctx->stream>>>(
        (const unsigned long long *)d_pos, (const unsigned long long *)d_neg, bins,
        (unsigned long long)min_spots, bw, low, h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], nhigh, d_out);
    ctx->launches++;
    TSK_CUDA(cudaGetLastError());
    return 0;
}
