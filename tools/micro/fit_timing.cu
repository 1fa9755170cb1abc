// microbenchmark: cycles per element of the ordered-sum chains of fit_kernel, by term kind and by number of
// chains running side by side; 296 CTAs (2 per SM, like production), m points each.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -std=c++17 [-DFIT_SKIP_SUMS=1] -o fit_timing fit_timing.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../../orthofinder_b200/csrc/ofg_fit_core.h"

__global__ void __launch_bounds__(256, 2) k_chain(const double* lx, const double* ly, long m, int nv, int k0, int k1, int k2, double* out, int reps)
{
    extern __shared__ __align__(16) unsigned char fit_smem_raw[];  // sizeof(FitShared) > 48 KB: dynamic, opt-in
    FitShared& sh = *reinterpret_cast<FitShared*>(fit_smem_raw);
    FitState S;
    S.x0 = 0.46; S.x1 = 0.4; S.h0 = 1e-8; S.h1 = 1e-8; S.xp00 = S.x0 + S.h0; S.xp01 = S.x1; S.xp10 = S.x0; S.xp11 = S.x1 + S.h1;
    S.aj0 = 3.0; S.t = 0.1; S.aj1 = 2.0; S.temp0 = 0.01; S.xn0 = 0.461; S.xn1 = 0.401; S.hn0 = 1e-8; S.hn1 = 1e-8;
    S.xnp0 = S.xn0 + S.hn0; S.xnp1 = S.xn1 + S.hn1; S.piv = 0; S.hh1 = 1; S.hh2 = 1;
    S.rh0 = fit_rcp(S.h0); S.rh1 = fit_rcp(S.h1); S.raj0 = fit_rcp(S.aj0); S.raj1 = fit_rcp(S.aj1); S.rhn0 = fit_rcp(S.hn0); S.rhn1 = fit_rcp(S.hn1);
    double o[3], r = 0;
    long long t0 = clock64();
    for (int i = 0; i < reps; i++) { fit_chain_multi(nv, k0, k1, k2, 1, m - 1, lx, ly, S, &sh, o); r += o[0]; }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = r; out[1] = (double)(t1 - t0) / reps / (double)(m - 1); }
}

int main()
{
    const long m = 50000;
    std::vector<double> hx(m), hy(m);
    for (long i = 0; i < m; i++) { hx[i] = 4.0 + 3.0 * (double)rand() / RAND_MAX; hy[i] = 0.46 * hx[i] + 0.4 + 0.1 * ((double)rand() / RAND_MAX - 0.5); }
    double *dx, *dy, *dout;
    cudaMalloc(&dx, m * 8); cudaMalloc(&dy, m * 8); cudaMalloc(&dout, 64);
    cudaMemcpy(dx, hx.data(), m * 8, cudaMemcpyHostToDevice); cudaMemcpy(dy, hy.data(), m * 8, cudaMemcpyHostToDevice);
    struct { const char* name; int nv, k0, k1, k2; } cases[] = {
        {"FNEW", 1, FIT_K_FNEW, 0, 0}, {"F", 1, FIT_K_F, 0, 0}, {"B1", 1, FIT_K_B1, 0, 0}, {"B2W", 1, FIT_K_B2W, 0, 0},
        {"AB+AF", 2, FIT_K_AB, FIT_K_AF, 0}, {"J0+J1", 2, FIT_K_J0, FIT_K_J1, 0}, {"J0+J1+F", 3, FIT_K_J0, FIT_K_J1, FIT_K_F},
        {"FNEW+J0N+J1N", 3, FIT_K_FNEW, FIT_K_J0N, FIT_K_J1N}};
    for (auto& c : cases) {
        cudaFuncSetAttribute(k_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FitShared));
        k_chain<<<296, 256, sizeof(FitShared)>>>(dx, dy, m, c.nv, c.k0, c.k1, c.k2, dout, 3);
        cudaDeviceSynchronize();
        double h[2];
        cudaMemcpy(h, dout, 16, cudaMemcpyDeviceToHost);
        printf("%-14s %d chain(s): %6.2f cycles per element%s\n", c.name, c.nv, h[1], FIT_SKIP_SUMS ? " (staging only)" : "");
    }
    return 0;
}
