// Throughput of __match_any_sync vs a ballot-based peer mask for 9-bit digits (B200).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o match_rate match_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned u32;
template <int BITS>
__device__ __forceinline__ u32 ballot_peers(u32 d)
{
    u32 m = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < BITS; b++) {
        const u32 bal = __ballot_sync(0xffffffffu, (d >> b) & 1u);
        m &= ((d >> b) & 1u) ? bal : ~bal;
    }
    return m;
}
template <int MODE>
__global__ void k(const u32* in, u32* out, int iters)
{
    u32 v[8];
    for (int i = 0; i < 8; i++) v[i] = in[(threadIdx.x + i * 97 + blockIdx.x) & 4095];
    u32 acc = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const u32 d = (v[i] + it * 7 + acc) & 511u;
            u32 p;
            if (MODE == 0) p = __match_any_sync(0xffffffffu, d);
            else if (MODE == 1) p = ballot_peers<9>(d);
            else p = d;
            acc += __popc(p);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
// independent matches (no serial dependency through acc)
template <int MODE>
__global__ void k_indep(const u32* in, u32* out, int iters)
{
    u32 v[8];
    for (int i = 0; i < 8; i++) v[i] = in[(threadIdx.x + i * 97 + blockIdx.x) & 4095];
    u32 acc = 0;
    for (int it = 0; it < iters; it++) {
        u32 p[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const u32 d = (v[i] + it * 7) & 511u;
            if (MODE == 0) p[i] = __match_any_sync(0xffffffffu, d);
            else p[i] = ballot_peers<9>(d);
        }
#pragma unroll
        for (int i = 0; i < 8; i++) acc += __popc(p[i]);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
// one match + two ballot-based peer masks per three items: do MATCH and VOTE overlap?
__global__ void k_mixed(const u32* in, u32* out, int iters)
{
    u32 v[9];
    for (int i = 0; i < 9; i++) v[i] = in[(threadIdx.x + i * 97 + blockIdx.x) & 4095];
    u32 acc = 0;
    for (int it = 0; it < iters; it++) {
        u32 p[9];
#pragma unroll
        for (int i = 0; i < 9; i++) {
            const u32 d = (v[i] + it * 7) & 511u;
            if (i % 3 == 0) p[i] = __match_any_sync(0xffffffffu, d);
            else p[i] = ballot_peers<9>(d);
        }
#pragma unroll
        for (int i = 0; i < 9; i++) acc += __popc(p[i]);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
// number of distinct values per warp: does MATCH.ANY cost depend on it?
template <int DISTINCT>
__global__ void k_distinct(u32* out, int iters)
{
    u32 acc = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const u32 d = ((threadIdx.x + it + i) % DISTINCT) + (acc & 0u);
            acc += __popc(__match_any_sync(0xffffffffu, d));
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// 9-bit digit matched as two narrower matches (low 5 bits, high 4 bits) whose results are ANDed
template <int MODE>
__global__ void k_split(const u32* in, u32* out, int iters)
{
    u32 v[8];
    for (int i = 0; i < 8; i++) v[i] = in[(threadIdx.x + i * 97 + blockIdx.x) & 4095];
    u32 acc = 0;
    for (int it = 0; it < iters; it++) {
        u32 p[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const u32 d = (v[i] + it * 7) & 511u;
            if (MODE == 0) p[i] = __match_any_sync(0xffffffffu, d & 31u) & __match_any_sync(0xffffffffu, d >> 5);
            else if (MODE == 1) p[i] = __match_any_sync(0xffffffffu, d & 15u) & __match_any_sync(0xffffffffu, (d >> 4) & 7u) & __match_any_sync(0xffffffffu, d >> 7);
            else if (MODE == 3) p[i] = __match_any_sync(0xffffffffu, d & 7u) & __match_any_sync(0xffffffffu, (d >> 3) & 7u) & __match_any_sync(0xffffffffu, d >> 6);
            else if (MODE == 4) p[i] = __match_any_sync(0xffffffffu, d & 3u) & __match_any_sync(0xffffffffu, (d >> 2) & 3u) & __match_any_sync(0xffffffffu, (d >> 4) & 3u)
                                     & __match_any_sync(0xffffffffu, (d >> 6) & 3u) & __match_any_sync(0xffffffffu, d >> 8);
            else p[i] = __match_any_sync(0xffffffffu, d);
        }
#pragma unroll
        for (int i = 0; i < 8; i++) acc += __popc(p[i]);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main()
{
    u32 *in, *out;
    cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, 148 * 8 * 1024 * 4);
    u32 h[4096]; for (int i = 0; i < 4096; i++) h[i] = (u32)(i * 2654435761u >> 7);
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2000;
    for (int warps = 1; warps <= 32; warps *= 2) {
        for (int mode = 0; mode < 5; mode++) {
            const int threads = warps * 32 > 1024 ? 1024 : warps * 32;
            const int blocks = 148 * (warps * 32 / threads);
            for (int rep = 0; rep < 2; rep++) {
                cudaEventRecord(e0);
                if (mode == 0) k<0><<<blocks, threads>>>(in, out, iters);
                if (mode == 1) k<1><<<blocks, threads>>>(in, out, iters);
                if (mode == 2) k<2><<<blocks, threads>>>(in, out, iters);
                if (mode == 3) k_indep<0><<<blocks, threads>>>(in, out, iters);
                if (mode == 4) k_indep<1><<<blocks, threads>>>(in, out, iters);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
            }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double cyc = ms * 1e-3 * 1.965e9;  // assumes boost clock
            const char* nm[] = {"match dep", "ballot9 dep", "none", "match indep", "ballot9 indep"};
            printf("warps/SM %2d  %-14s  %8.1f cycles per warp-item (per warp)  %7.2f SM-cycles per warp-item\n", warps, nm[mode],
                   cyc / (iters * 8.0), cyc / (iters * 8.0) / warps);
        }
    }
    {
        const int threads = 1024, blocks = 148;
        float ms;
        for (int rep = 0; rep < 2; rep++) { cudaEventRecord(e0); k_mixed<<<blocks, threads>>>(in, out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); }
        cudaEventElapsedTime(&ms, e0, e1);
        printf("mixed (1 match + 2 ballot9 per 3 items), 32 warps/SM: %7.2f SM-cycles per warp-item\n", ms * 1e-3 * 1.965e9 / (iters * 9.0) / 32);
#define DIST(N) for (int rep = 0; rep < 2; rep++) { cudaEventRecord(e0); k_distinct<N><<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); } \
        cudaEventElapsedTime(&ms, e0, e1); printf("match.any with %2d distinct values per warp: %7.2f SM-cycles per warp-instruction\n", N, ms * 1e-3 * 1.965e9 / (iters * 8.0) / 32);
        DIST(1) DIST(2) DIST(4) DIST(8) DIST(16) DIST(32)
#define SPLIT(M, NAME) for (int rep = 0; rep < 2; rep++) { cudaEventRecord(e0); k_split<M><<<blocks, threads>>>(in, out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); } \
        cudaEventElapsedTime(&ms, e0, e1); printf("random 9-bit digit, %s: %7.2f SM-cycles per warp-item\n", NAME, ms * 1e-3 * 1.965e9 / (iters * 8.0) / 32);
        SPLIT(2, "one match") SPLIT(0, "match(5 bits) & match(4 bits)") SPLIT(1, "match(4) & match(3) & match(2)") SPLIT(3, "match(3) x 3") SPLIT(4, "match(2) x 4 & match(1)")
    }
    return 0;
}
