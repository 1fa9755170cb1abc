// DendroBLAST distance matrices on the raw slotted CSR.
//
// Reference: scripts_of/orthologues.py
//   :264-291  Worker_OGMatrices_ReadBLASTAndUpdateDistances   M_ij = 0.5 * max(B_ij, Bmin_i) * (1 / Bmax_i)
//   :226-236  lil_minmax                                      min / max over a gene's stored hits in every target species
//   :396-410  CompleteAndWriteOGMatrices                      D_ij = D_ji = -log(M_ij + M_ji)  (j < i), max_og
// B is GetBLAST6Scores with qExcludeSelfHits = False: the context is built to OFG_STAGE_RAW with option keep_self_hits.
// All arithmetic is the reference's, operation by operation (IEEE division and products, no FMA contraction); only log()
// is CUDA's (<= 1 ulp) instead of numpy's.
#pragma once
#include "ofg_common.cuh"

struct DendroArgs {
    const PairDesc* pairs;
    const int32_t* pair_of;     // [S * S]
    int S;
    const int64_t* seq_start;   // [S]
    const u32* row_start;
    const u32* rowlen;
    const u32* cols;
    const double* vals;
    double* gmin;               // [N] per gene
    double* gmaxinv;            // [N]
    int64_t N;
    // orthogroups
    int64_t n_members;
    const int32_t* mem_og;      // [n_members] orthogroup of a member
    const int64_t* og_off;      // [n_ogs + 1] first member of an orthogroup
    const int64_t* out_off;     // [n_ogs + 1] first matrix element of an orthogroup
    const int32_t* mem_sp;      // [n_members] species (list position)
    const int32_t* mem_seq;     // [n_members] iSeq
    double* M;
    double* D;
    unsigned long long* og_max; // [n_ogs] max_og as ordered bits
};

// total order of doubles as unsigned integers (for atomicMax)
__device__ __forceinline__ unsigned long long dendro_order_bits(double x)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

// one warp per gene: min / max over the gene's rows in every pair (p, j) (:275-281); 9e99 / 0 without hits
__global__ void __launch_bounds__(256) dendro_minmax_kernel(DendroArgs a)
{
    const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= a.N) return;
    const u32 lane = lane_id();
    int p = 0;
    while (p + 1 < a.S && a.seq_start[p + 1] <= g) p++;
    const int64_t q = g - a.seq_start[p];
    double mn = 9e99, mx = 0.0;
    for (int j = 0; j < a.S; j++) {
        const int pi = a.pair_of[(size_t)p * a.S + j];
        if (pi < 0) continue;
        const int64_t r = a.pairs[pi].row_base + q;
        const u32 s0 = a.row_start[r], n = a.rowlen[r];
        for (u32 k = lane; k < n; k += 32) {
            const double v = a.vals[s0 + k];
            mn = fmin(mn, v);
            mx = fmax(mx, v);
        }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        mn = fmin(mn, __shfl_xor_sync(OFG_FULL, mn, d));
        mx = fmax(mx, __shfl_xor_sync(OFG_FULL, mx, d));
    }
    if (lane == 0) {
        a.gmin[g] = mn;
        a.gmaxinv[g] = __ddiv_rn(1.0, mx);   // 1./maxes (:282): inf for a gene without hits
    }
}

// one warp per member i of an orthogroup: row i of M (:283-287); the lanes take the members j
__global__ void __launch_bounds__(256) dendro_rows_kernel(DendroArgs a)
{
    const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (m >= a.n_members) return;
    const u32 lane = lane_id();
    const int og = a.mem_og[m];
    const int64_t m0 = a.og_off[og], n = a.og_off[og + 1] - m0, i = m - m0;
    const int p = a.mem_sp[m];
    const int64_t q = a.mem_seq[m];
    const int64_t gi = a.seq_start[p] + q;
    const double mn = a.gmin[gi], inv = a.gmaxinv[gi];
    double* out = a.M + a.out_off[og] + i * n;
    for (int64_t j = lane; j < n; j += 32) {
        const int pj = a.mem_sp[m0 + j];
        const u32 cj = (u32)a.mem_seq[m0 + j];
        double b = 0.0;   // lil_matrix[...] of an absent entry
        const int pi = a.pair_of[(size_t)p * a.S + pj];
        if (pi >= 0) {
            const int64_t r = a.pairs[pi].row_base + q;
            const u32 s0 = a.row_start[r];
            u32 lo = 0, hi = a.rowlen[r];
            while (lo < hi) {   // columns of a row are ascending after the row sort
                const u32 mid = (lo + hi) >> 1;
                const u32 c = a.cols[s0 + mid] & 0x0fffffffu;
                if (c < cj) lo = mid + 1; else hi = mid;
            }
            if (lo < a.rowlen[r] && (a.cols[s0 + lo] & 0x0fffffffu) == cj) b = a.vals[s0 + lo];
        }
        const double mxv = (mn > b) ? mn : b;   // max(B, mins) of two Python floats
        out[j] = __dmul_rn(__dmul_rn(0.5, mxv), inv);
    }
}

// one warp per member i: D_ij = D_ji = -log(M_ij + M_ji) for j < i; the diagonal keeps M's value (:403-409)
__global__ void __launch_bounds__(256) dendro_complete_kernel(DendroArgs a)
{
    const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (m >= a.n_members) return;
    const u32 lane = lane_id();
    const int og = a.mem_og[m];
    const int64_t m0 = a.og_off[og], n = a.og_off[og + 1] - m0, i = m - m0;
    const double* M = a.M + a.out_off[og];
    double* D = a.D + a.out_off[og];
    double best = -9e99;
    for (int64_t j = lane; j < i; j += 32) {
        const double d = -log(__dadd_rn(M[i * n + j], M[j * n + i]));
        D[i * n + j] = d;
        D[j * n + i] = d;
        best = (d > best) ? d : best;   // max(max_og, m[i][j])
    }
    if (lane == 0) D[i * n + i] = M[i * n + i];
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        const double o = __shfl_xor_sync(OFG_FULL, best, d);
        best = (o > best) ? o : best;
    }
    if (lane == 0) atomicMax(&a.og_max[og], dendro_order_bits(best));
}
