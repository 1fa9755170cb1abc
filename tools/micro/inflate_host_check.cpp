// Host build of the device inflate (orthofinder_b200/csrc/ofg_inflate.cuh) with a one-lane "warp": checks the decoder's
// logic against zlib without a GPU.   g++ -O1 -o /tmp/inflate_check tools/micro/inflate_host_check.cpp && /tmp/inflate_check file.gz file.raw
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#define OFG_HOST_SHIM
typedef unsigned long long u64;
typedef uint32_t u32;
#define OFG_FULL 0xffffffffu
#define __device__
#define __global__
#define __forceinline__ inline
#define __constant__ static
#define __shared__ static
#define __launch_bounds__(x)
static inline void __syncwarp() {}
static inline void __syncwarp(unsigned) {}
static inline int __any_sync(unsigned, int p) { return p; }
static inline u32 lane_id() { return 0; }
static inline u32 __brev(u32 x) { u32 r = 0; for (int i = 0; i < 32; i++) if (x & (1u << i)) r |= 1u << (31 - i); return r; }
static inline u32 __funnelshift_r(u32 lo, u32 hi, int sh) { sh &= 31; return sh ? (lo >> sh) | (hi << (32 - sh)) : lo; }
static inline unsigned long long atomicMin(unsigned long long* a, unsigned long long v) { unsigned long long o = *a; if (v < o) *a = v; return o; }
struct D3 { int x; };
static D3 blockIdx = {0}, threadIdx = {0}, gridDim = {1};
#include "../../orthofinder_b200/csrc/ofg_inflate.cuh"

int main(int argc, char** argv)
{
    FILE* f = fopen(argv[1], "rb");
    std::vector<uint8_t> z(1 << 26);
    size_t n = fread(z.data(), 1, z.size() - 64, f);
    fclose(f);
    f = fopen(argv[2], "rb");
    std::vector<uint8_t> raw(1 << 27);
    size_t nr = fread(raw.data(), 1, raw.size(), f);
    fclose(f);
    // single member with a 10-byte header
    GzMember m;
    m.in_off = 10; m.in_len = n - 18; m.out_off = 0; m.file = 0; m.pad = 0;
    m.out_len = z[n - 4] | (z[n - 3] << 8) | (z[n - 2] << 16) | ((u64)z[n - 1] << 24);
    int shift = argc > 3 ? atoi(argv[3]) : 0;   // misalign the output by 0..3 bytes
    m.out_off = shift;
    std::vector<uint8_t> out(m.out_len + 64 + 16);
    unsigned long long err = ~0ULL;
    // 4-byte aligned copy of the stream (the kernel reads aligned words)
    uint8_t* zz = (uint8_t*)aligned_alloc(16, n + 64);
    memcpy(zz, z.data(), n);
    memset(zz + n, 0, 64);
    inflate_kernel(zz, &m, 1, out.data(), &err);
    printf("isize %llu raw %zu err %llx equal %d\n", m.out_len, nr, err, (int)(m.out_len == nr && !memcmp(out.data() + shift, raw.data(), nr)));
    if (err == ~0ULL && m.out_len == nr) for (size_t i = 0; i < nr; i++) if (out[i + shift] != raw[i]) { printf("first diff at %zu\n", i); break; }
    return 0;
}
