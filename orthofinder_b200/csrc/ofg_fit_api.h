// Launchers of the curve_fit emulation kernels (ofg_fit.cu), callable from the orchestration unit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int FIT_THREADS = 256;

struct FitArgs {
    int P;
    const unsigned long long* pair_sel_off;  // [P+1] offsets of each pair's selected points in lx / ly
    unsigned long long sel_cap;
    const double* lx;
    const double* ly;
    double* fit;      // [P][8]: a, b, m, info, nfev, iterative, ok, pad
    uint32_t* err_flags;   // bit 0: a fit failed (curve_fit would raise); bit 1: selection overflowed sel_cap
};

// one CTA per species pair; returns the launch status
cudaError_t ofg_launch_fit(const FitArgs& a, int grid, cudaStream_t stream);
// one problem (unit tests): out = a, b, m, info, nfev, iterative
cudaError_t ofg_launch_test_fit(const double* lx, const double* ly, long m, double* out, cudaStream_t stream);
