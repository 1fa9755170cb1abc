// curve_fit / MINPACK lmdif emulation kernels (gathering.py:119-121), one CTA per species pair.  Separate translation
// unit: ptxas spends most of the library's build time here.
#include "ofg_common.cuh"
#include "ofg_fit_api.h"
#include "ofg_fit_core.h"

__global__ void __launch_bounds__(FIT_THREADS) fit_kernel(FitArgs a)
{
    extern __shared__ __align__(16) unsigned char fit_smem_raw[];  // sizeof(FitShared) > 48 KB: dynamic, opt-in
    FitShared& sh = *reinterpret_cast<FitShared*>(fit_smem_raw);
    for (int pair = blockIdx.x; pair < a.P; pair += gridDim.x) {
        const u64 o0 = a.pair_sel_off[pair], o1 = a.pair_sel_off[pair + 1];
        const long m = (long)(o1 - o0);
        double* out = a.fit + (size_t)pair * 8;
        if (o1 > a.sel_cap) {
            if (threadIdx.x == 0) { atomicOr(a.err_flags, 2u); out[6] = 0.0; out[2] = (double)m; out[3] = -1.0; }
            continue;
        }
        if (m <= 1) {  // len(topScores) > 1 fails: the pair's hits are ignored (gathering.py:145,152-155)
            if (threadIdx.x == 0) { out[0] = 0.0; out[1] = 0.0; out[2] = (double)m; out[3] = 0.0; out[4] = 0.0; out[5] = 0.0; out[6] = 0.0; }
            continue;
        }
        const int info = fit_loglinear(a.lx + o0, a.ly + o0, m, &sh, out);
        if (threadIdx.x == 0) {
            const bool ok = info >= 1 && info <= 4;
            out[6] = ok ? 1.0 : 0.0;
            if (!ok) atomicOr(a.err_flags, 1u);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(FIT_THREADS) test_fit_kernel(const double* lx, const double* ly, long m, double* out)
{
    extern __shared__ __align__(16) unsigned char fit_smem_raw[];
    FitShared& sh = *reinterpret_cast<FitShared*>(fit_smem_raw);
    fit_loglinear(lx, ly, m, &sh, out);
}

cudaError_t ofg_launch_fit(const FitArgs& a, int grid, cudaStream_t stream)
{
    cudaError_t e = cudaFuncSetAttribute(fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FitShared));
    if (e != cudaSuccess) return e;
    fit_kernel<<<grid, FIT_THREADS, sizeof(FitShared), stream>>>(a);
    return cudaGetLastError();
}

cudaError_t ofg_launch_test_fit(const double* lx, const double* ly, long m, double* out, cudaStream_t stream)
{
    cudaError_t e = cudaFuncSetAttribute(test_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FitShared));
    if (e != cudaSuccess) return e;
    test_fit_kernel<<<1, FIT_THREADS, sizeof(FitShared), stream>>>(lx, ly, m, out);
    return cudaGetLastError();
}
