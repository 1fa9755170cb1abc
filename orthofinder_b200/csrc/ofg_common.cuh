// Shared device helpers and descriptors for libofg (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef unsigned long long u64;
typedef uint32_t u32;

#define OFG_FULL 0xffffffffu
#define OFG_ROW_SENTINEL 0xffffffffu

// One species pair (p, j) handled by this context: Blast{p}_{j} -> B_pj.
struct PairDesc {
    int32_t p, j;          // list positions of query / target species
    int32_t same;          // p == j : self hits dropped (blast_file_processor.py:83-84)
    int32_t swap;          // read query/hit from columns 1/0 (-1 mode, :41-54)
    int32_t Gi, Gj;        // nSeqs of query / target species
    int32_t present;       // hit text was provided
    int32_t owned_row;     // p is an owned query species (else: column pair of design A)
    int64_t text_off;      // byte offset of the file in the text arena (16-byte aligned)
    int64_t text_len;      // padded length (multiple of 16, padding is '\n')
    int64_t text_raw;      // real file size
    int64_t row_base;      // first global row index of this pair (rows of all pairs concatenated)
    int64_t col_base;      // first index into the per-pair column factor array
    int64_t li_off, lj_off;  // offsets of the two length vectors in lengths_concat
    int64_t gi_start, gj_start;  // seqStartingIndices of the two species (global gene ids)
};

__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ u32 lanemask_lt() { return (1u << (threadIdx.x & 31u)) - 1u; }

__device__ __forceinline__ u32 warp_incl_scan(u32 v)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 t = __shfl_up_sync(OFG_FULL, v, d);
        if ((int)lane_id() >= d) v += t;
    }
    return v;
}

// Exclusive scan of one u32 per thread over a CTA of NT threads.  smem: NT/32 + 1 words.
// Returns the exclusive prefix; *total receives the CTA sum.  Contains __syncthreads.
template <int NT>
__device__ __forceinline__ u32 block_excl_scan(u32 v, u32* smem, u32* total)
{
    u32 inc = warp_incl_scan(v);
    int w = threadIdx.x >> 5;
    if (lane_id() == 31) smem[w] = inc;
    __syncthreads();
    if (w == 0) {
        u32 s = (lane_id() < NT / 32) ? smem[lane_id()] : 0u;
        u32 si = warp_incl_scan(s);
        if (lane_id() < NT / 32) smem[lane_id()] = si - s;
        if (lane_id() == NT / 32 - 1) smem[NT / 32] = si;
    }
    __syncthreads();
    u32 r = smem[w] + inc - v;
    *total = smem[NT / 32];
    __syncthreads();
    return r;
}

// first index i in [0, n) with a[i] > x   (a ascending)
template <typename T>
__device__ __forceinline__ int upper_bound_dev(const T* a, int n, T x)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (a[mid] > x) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// positive doubles order like their bit patterns
__device__ __forceinline__ void atomic_max_pos_double(double* addr, double v)
{
    atomicMax((u64*)addr, (u64)__double_as_longlong(v));
}
__device__ __forceinline__ void atomic_min_pos_double(double* addr, double v)
{
    atomicMin((u64*)addr, (u64)__double_as_longlong(v));
}
