// microbenchmark: latency of a dependent fp64 add chain (registers / shared-memory fed) on one thread
#include <cstdio>
#include <cuda_runtime.h>
__global__ void chain_reg(double* out, int n, double x)
{
    double s = 0.0;
    long long t0 = clock64();
    for (int i = 0; i < n; i += 8) {
        s = __dadd_rn(s, x); s = __dadd_rn(s, x); s = __dadd_rn(s, x); s = __dadd_rn(s, x);
        s = __dadd_rn(s, x); s = __dadd_rn(s, x); s = __dadd_rn(s, x); s = __dadd_rn(s, x);
    }
    long long t1 = clock64();
    out[0] = s; out[1] = (double)(t1 - t0) / n;
}
__global__ void chain_smem(double* out, int n)
{
    __shared__ __align__(16) double q[512];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) q[i] = 1.0 + i * 1e-9;
    __syncthreads();
    if (threadIdx.x != 0) return;
    double s = 0.0;
    const double2* q2 = reinterpret_cast<const double2*>(q);
    long long t0 = clock64();
    for (int rep = 0; rep < n / 512; rep++) {
        for (int i = 0; i < 512; i += 8) {
            double2 a = q2[i / 2], b = q2[i / 2 + 1], c = q2[i / 2 + 2], d = q2[i / 2 + 3];
            s = __dadd_rn(s, a.x); s = __dadd_rn(s, a.y); s = __dadd_rn(s, b.x); s = __dadd_rn(s, b.y);
            s = __dadd_rn(s, c.x); s = __dadd_rn(s, c.y); s = __dadd_rn(s, d.x); s = __dadd_rn(s, d.y);
        }
    }
    long long t1 = clock64();
    out[2] = s; out[3] = (double)(t1 - t0) / n;
}
__global__ void chain_ffma(float* out, int n, float x)
{
    float s = 0.f;
    long long t0 = clock64();
    for (int i = 0; i < n; i += 8) {
        s = __fadd_rn(s, x); s = __fadd_rn(s, x); s = __fadd_rn(s, x); s = __fadd_rn(s, x);
        s = __fadd_rn(s, x); s = __fadd_rn(s, x); s = __fadd_rn(s, x); s = __fadd_rn(s, x);
    }
    long long t1 = clock64();
    out[0] = s; out[1] = (float)(t1 - t0) / n;
}
int main()
{
    double* d; float* f;
    cudaMalloc(&d, 64); cudaMalloc(&f, 64);
    chain_reg<<<1, 32>>>(d, 1 << 20, 1.000000001);
    chain_smem<<<1, 128>>>(d, 1 << 20);
    chain_ffma<<<1, 32>>>(f, 1 << 20, 1.0001f);
    double h[4]; float hf[2];
    cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost); cudaMemcpy(hf, f, 8, cudaMemcpyDeviceToHost);
    printf("DADD chain (regs): %.2f cycles/add\nDADD chain (LDS.128 fed): %.2f cycles/add\nFADD chain: %.2f cycles/add\n", h[1], h[3], hf[1]);
    return 0;
}
