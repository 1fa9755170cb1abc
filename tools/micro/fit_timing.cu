// microbenchmark: where does fit_kernel spend its time?  One CTA, m points, instrumented copies of the
// chain primitives (clock64 on the accumulating thread).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#define FIT_PROFILE 1
__device__ long long g_t_sum, g_t_wait, g_t_total, g_n_chunks;
#include "../../orthofinder_b200/csrc/ofg_fit_core.h"

__global__ void __launch_bounds__(256) k_enorm(const double* x, long m, double* out, int reps)
{
    __shared__ __align__(16) FitShared sh;
    long long t0 = clock64();
    double r = 0;
    for (int i = 0; i < reps; i++) r += fit_enorm_big(x, m, &sh);
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = r; out[1] = (double)(t1 - t0) / reps; }
}
__global__ void __launch_bounds__(256) k_dot(const double* x, const double* y, long m, double* out, int reps)
{
    __shared__ __align__(16) FitShared sh;
    long long t0 = clock64();
    double r = 0;
    for (int i = 0; i < reps; i++) r += fit_dot_big(x, y, m, &sh);
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[2] = r; out[3] = (double)(t1 - t0) / reps; }
}
__global__ void __launch_bounds__(256) k_osum(double* out, int reps)
{
    __shared__ __align__(16) double q[512];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) q[i] = 1.0 + i * 1e-9;
    __syncthreads();
    double s = 0.0;
    long long t0 = clock64();
    if (threadIdx.x == 0) for (int i = 0; i < reps; i++) s = fit_ordered_sum(s, q, 512);
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[6] = s; out[7] = (double)(t1 - t0) / (reps * 512.0); }
    // same with a barrier per chunk
    __syncthreads();
    t0 = clock64();
    for (int i = 0; i < reps; i++) { if (threadIdx.x == 0) s = fit_ordered_sum(s, q, 512); __syncthreads(); }
    t1 = clock64();
    if (threadIdx.x == 0) { out[14] = s; out[15] = (double)(t1 - t0) / (reps * 512.0); }
}
__global__ void __launch_bounds__(256) k_elem(double* x, const double* y, long m, double* out, int reps)
{
    long long t0 = clock64();
    for (int i = 0; i < reps; i++) {
        for (long k = threadIdx.x; k < m; k += blockDim.x) x[k] = __ddiv_rn(__dsub_rn(x[k], y[k]), 1.0000001);
        __syncthreads();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[4] = (double)(t1 - t0) / reps;
}
__global__ void __launch_bounds__(256) k_fit(const double* lx, const double* ly, long m, double* scratch, double* out)
{
    __shared__ __align__(16) FitShared sh;
    long long t0 = clock64();
    fit_loglinear(lx, ly, m, scratch, scratch + m, scratch + 2 * m, scratch + 3 * m, &sh, out + 8);
    long long t1 = clock64();
    if (threadIdx.x == 0) out[5] = (double)(t1 - t0);
}
int main(int argc, char** argv)
{
    long m = argc > 1 ? atol(argv[1]) : 53000;
    int nblk = argc > 2 ? atoi(argv[2]) : 1;
    std::vector<double> lx(m), ly(m);
    srand(1);
    for (long i = 0; i < m; i++) { lx[i] = 3.0 + 3.0 * rand() / RAND_MAX; ly[i] = 0.46 * lx[i] + 0.4 + 0.3 * (rand() / (double)RAND_MAX - 0.5); }
    double *dx, *dy, *ds, *dout;
    cudaMalloc(&dx, m * 8 * nblk); cudaMalloc(&dy, m * 8 * nblk); cudaMalloc(&ds, 4 * m * 8 * nblk); cudaMalloc(&dout, 256 * nblk);
    for (int b = 0; b < nblk; b++) { cudaMemcpy(dx + b * m, lx.data(), m * 8, cudaMemcpyHostToDevice); cudaMemcpy(dy + b * m, ly.data(), m * 8, cudaMemcpyHostToDevice); }
    k_enorm<<<1, 256>>>(dx, m, dout, 10);
    { long long a, b, c; cudaMemcpyFromSymbol(&a, g_t_sum, 8); cudaMemcpyFromSymbol(&b, g_t_wait, 8); cudaMemcpyFromSymbol(&c, g_n_chunks, 8);
      printf("enorm: chunks %lld, cycles in ordered_sum %.0f per chunk, at barrier %.0f per chunk\n", c, (double)a / c, (double)b / c); }
    k_dot<<<1, 256>>>(dx, dy, m, dout, 10);
    k_elem<<<1, 256>>>(ds, dy, m, dout, 10);
    k_osum<<<1, 256>>>(dout, 100);
    k_fit<<<1, 256>>>(dx, dy, m, ds, dout);
    double h[16];
    cudaMemcpy(h, dout, 128, cudaMemcpyDeviceToHost);
    printf("ordered_sum alone: %.2f cycles/elem; with __syncthreads per 512: %.2f\n", h[7], h[15]);
    printf("m=%ld\nenorm chain: %.0f cycles (%.2f/elem)\ndot chain:   %.0f cycles (%.2f/elem)\nelementwise pass (load2, sub, div, store): %.0f cycles (%.2f/elem)\nwhole fit: %.0f cycles = %.3f ms @1.965GHz; a=%.12f b=%.12f info=%g nfev=%g\n",
           m, h[1], h[1] / m, h[3], h[3] / m, h[4], h[4] / m, h[5], h[5] / 1.965e6, h[8], h[9], h[11], h[12]);
    return 0;
}
