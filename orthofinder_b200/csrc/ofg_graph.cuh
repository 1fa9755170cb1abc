// Stages NORM (part 3), THRESH and GRAPH: everything after the fit.
//
//  fit_kernel      : one CTA per species pair, ofg_fit_core.h (curve_fit / MINPACK lmdif emulation)
//  factors_kernel  : per-gene factors 10^-b * Lq^-a and Lh^-a   (gathering.py:124-131)
//  scale_bh_kernel : B' = ((10^-b * Lq^-a) * B) * Lh^-a left to right, row maxima, best hits with the
//                    1e-3 band for j != p, per-gene best hit in another species (gathering.py:213-234)
//  bh_self_kernel  : best-hit flags of the within-species matrix (gathering.py:236-247)
//  rbh_kernel      : RBH_pk = BH_pk AND BH_kp^T by binary search in the transposed pair, per-gene
//                    "most distant" threshold with last-column-wins semantics (gathering.py:294-306)
//  thresh_final    : genes without an RBH get bestHit + 1e-6 (gathering.py:308-309)
//  connect_kernel  : connect_pj = B' >= threshold (gathering.py:391-408) and the reverse flag
//                    connect_jp^T scattered into the transposed pair (gathering.py:27-31); books the edge
//                    count and text bytes of every graph row while it sets the flags (edge_account)
//  v2_ratio_kernel : --scores-v2 only: percentile-ratio conversion factors (gathering.py:344-364)
//  emit_*          : W = (connect + connect^T) .* B' as CSR with global gene ids and as MCL text with
//                    glibc-exact "%d:%.3f " formatting (gathering.py:39-48); offsets from the booked sizes
#pragma once
#include "ofg_common.cuh"

// ---------------------------------------------------------------- fit
// fit_kernel lives in its own translation unit (ofg_fit.cu): it is by far the slowest piece to compile.
#include "ofg_fit_api.h"

// ---------------------------------------------------------------- normalisation factors
struct FactorArgs {
    const PairDesc* pairs;
    int P;
    const double* fit;
    const double* lengths;
    double* rowfac;   // [total rows]
    double* colfac;   // [total cols]
    const double* given;  // --scores-v2 with ofg_set_length_factors: L ** -0.5 per gene, indexed like lengths; else null
};

__global__ void __launch_bounds__(256) factors_kernel(FactorArgs a)
{
    // grid.x = pair (S^2 pairs can exceed the 65535 limit of grid.y), grid.y strides over Gi + Gj genes
    const int pair = blockIdx.x;
    const PairDesc pd = a.pairs[pair];
    const double* f = a.fit + (size_t)pair * 8;
    if (f[6] == 0.0) return;
    const double pa = f[0], pb = f[1];
    const double ten_mb = pow(10.0, -pb);  // 10**(-params[1])
    const int64_t n = (int64_t)pd.Gi + pd.Gj;
    for (int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.y * blockDim.x) {
        if (i < pd.Gi) a.rowfac[pd.row_base + i] = __dmul_rn(ten_mb, a.given ? a.given[pd.li_off + i] : pow(a.lengths[pd.li_off + i], -pa));
        else a.colfac[pd.col_base + (i - pd.Gi)] = a.given ? a.given[pd.lj_off + (i - pd.Gi)] : pow(a.lengths[pd.lj_off + (i - pd.Gi)], -pa);
    }
}

// rows are visited in increasing order by every warp: the pair index only moves forward
__device__ __forceinline__ int advance_pair(const int64_t* pair_row_base, int P, int pair, int64_t row)
{
    if (pair < 0) pair = upper_bound_dev(pair_row_base, P + 1, row) - 1;
    while (row >= pair_row_base[pair + 1]) pair++;
    return pair;
}

// ---------------------------------------------------------------- scale + best hits
struct ScaleArgs {
    int64_t row_begin;   // rows [row_begin, n_rows) are processed (one wave of query species)
    int64_t n_rows;
    const u32* row_start;
    u32* rowlen;
    u32* cols;       // flags are OR-ed into the top bits
    double* vals;    // raw scores in, normalised scores out
    const PairDesc* pairs;
    const int64_t* pair_row_base;
    int P;
    const double* fit;
    const double* rowfac;
    const double* colfac;
    double* rowmax;  // [n_rows]
    double* best;    // [nSeqs] bit patterns, 0 = no hit in another species
    u32* row_flagcnt;  // [n_rows] best hits per row (pre-zeroed): sizes of the design-B export lists
};

// Row kernels below give a part of a warp (ROW_G = 16 or 8 lanes) to each matrix row: rows hold ~50 entries, the work per row is
// a short dependent chain of loads, and several rows in flight per warp hide that latency better than one.  All
// groups run the same trip counts (max over the warp's rows) so that warp-wide shuffles stay converged.

// largest value of n over the groups of a warp (uniform trip counts)
template <int ROW_G>
__device__ __forceinline__ int warp_max_over_groups(int n)
{
#pragma unroll
    for (int d = ROW_G; d < 32; d <<= 1) n = max(n, __shfl_xor_sync(OFG_FULL, n, d));
    return n;
}
template <int ROW_G>
__device__ __forceinline__ double group_max(double v)
{
#pragma unroll
    for (int d = ROW_G / 2; d > 0; d >>= 1) v = fmax(v, __shfl_xor_sync(OFG_FULL, v, d));
    return v;
}
template <int ROW_G>
__device__ __forceinline__ u32 group_sum(u32 v)
{
#pragma unroll
    for (int d = ROW_G / 2; d > 0; d >>= 1) v += __shfl_xor_sync(OFG_FULL, v, d);
    return v;
}

template <int ROW_G>
__global__ void __launch_bounds__(256) scale_bh_kernel(ScaleArgs a)
{
    const int lane = lane_id(), sub = lane / ROW_G, gl = lane % ROW_G;
    const int64_t n_groups = (((int64_t)gridDim.x * blockDim.x) >> 5) * (32 / ROW_G);
    int pair = -1;
    for (int64_t row0 = a.row_begin + (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * (32 / ROW_G); row0 < a.n_rows; row0 += n_groups) {
        const int64_t row = row0 + sub;
        int n = row < a.n_rows ? (int)a.rowlen[row] : 0;
        u32 s0 = 0;
        double rf = 0.0;
        const double* cf = nullptr;
        bool same = true, owned = false;
        int64_t gene = 0;
        if (n > 0) {
            pair = advance_pair(a.pair_row_base, a.P, pair, row);
            const PairDesc& pd = a.pairs[pair];
            if (a.fit[(size_t)pair * 8 + 6] == 0.0) {  // zeroed pair: sparse.lil_matrix(B.get_shape())
                if (gl == 0) a.rowlen[row] = 0;
                n = 0;
            } else {
                s0 = a.row_start[row];
                rf = a.rowfac[row];
                cf = a.colfac + pd.col_base;
                same = pd.same != 0;
                owned = pd.owned_row != 0;
                gene = pd.gi_start + (row - pd.row_base);
            }
        }
        const int nmax = warp_max_over_groups<ROW_G>(n);
        double m = 0.0;
        for (int k = gl; k < nmax; k += ROW_G) {
            if (k < n) {
                const u32 c = a.cols[s0 + k];
                const double v = __dmul_rn(__dmul_rn(rf, a.vals[s0 + k]), cf[c]);
                a.vals[s0 + k] = v;
                m = fmax(m, v);
            }
        }
        m = group_max<ROW_G>(m);
        if (gl == 0 && n > 0) a.rowmax[row] = m;
        const double thr = __dsub_rn(m, 1e-3);
        u32 nbh = 0;
        if (!same)
            for (int k = gl; k < n; k += ROW_G)
                if (a.vals[s0 + k] > thr) { a.cols[s0 + k] |= OFG_FLAG_BH; nbh++; }
        nbh = group_sum<ROW_G>(nbh);
        if (gl == 0 && n > 0 && !same) {
            a.row_flagcnt[row] = nbh;
            if (owned && m > 0.0) atomic_max_pos_double(&a.best[gene], m);
        }
    }
}

__global__ void __launch_bounds__(256) bh_self_kernel(ScaleArgs a, const int32_t* diag_pairs)
{
    // grid.y = index into the list of within-species pairs; warps stride over that pair's rows
    const int lane = lane_id();
    const PairDesc& pd = a.pairs[diag_pairs[blockIdx.y]];
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < pd.Gi; q += n_warps) {
        const int64_t row = pd.row_base + q;
        const int n = (int)a.rowlen[row];
        if (n == 0) continue;
        const u32 s0 = a.row_start[row];
        const double b = a.best[pd.gi_start + q];
        const double bestv = (__double_as_longlong(b) == 0) ? -1.0 : b;  // bestHitForSequence starts at -1
        const double thr = __dsub_rn(bestv, 1e-3);
        for (int k = lane; k < n; k += 32)
            if (a.vals[s0 + k] > thr) a.cols[s0 + k] |= OFG_FLAG_BH;
    }
}

// ---------------------------------------------------------------- RBH + thresholds
struct GraphArgs {
    int64_t n_rows;
    const u32* row_start;
    const u32* rowlen;
    u32* cols;
    const double* vals;
    const PairDesc* pairs;
    const int64_t* pair_row_base;
    int P;
    const int32_t* pair_of;  // [S*S] pair index of (p, j) or -1
    int S;
    double* most;            // [nSeqs] thresholds (bit patterns while being reduced)
    const double* best;
    u32* row_flagcnt;        // [n_rows] connect entries per row (connect_kernel; pre-zeroed)
    double* zrow;            // [n_rows] --scores-v2: the row's RBH score if it has exactly one RBH, else 0
    const double* pair_factor;  // [P] --scores-v2: conversion factor of the pair's target species
    u32* gene_cnt;           // [nSeqs] edges of every graph row, [nSeqs] text bytes of its entries (edge_account)
    u32* gene_bytes;
    u32* err_flags;
};

// index of column `want` in the row [s0, s0+n) of the slotted CSR, or -1
__device__ __forceinline__ int64_t find_col(const u32* cols, u32 s0, int n, u32 want)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const u32 c = cols[s0 + mid] & OFG_COL_MASK;
        if (c < want) lo = mid + 1; else hi = mid;
    }
    if (lo < n && (cols[s0 + lo] & OFG_COL_MASK) == want) return (int64_t)s0 + lo;
    return -1;
}

template <int ROW_G, int MODE>
__global__ void __launch_bounds__(256) rbh_kernel(GraphArgs a)
{
    // MODE 0: default thresholds (gathering.py:294-306).  --scores-v2 (:314-388) runs MODE 1 (per-row RBH score Z,
    // :340) before the percentile ratios and MODE 2 (thresholds scaled by the species factor, :381-384) after.
    const int lane = lane_id(), sub = lane / ROW_G, gl = lane % ROW_G;
    const int64_t n_groups = (((int64_t)gridDim.x * blockDim.x) >> 5) * (32 / ROW_G);
    int pair = -1;
    for (int64_t row0 = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * (32 / ROW_G); row0 < a.n_rows; row0 += n_groups) {
        const int64_t row = row0 + sub;
        int n = row < a.n_rows ? (int)a.rowlen[row] : 0;
        int tp = -1;
        u32 s0 = 0, r = 0;
        int64_t gene = 0;
        if (n > 0) {
            pair = advance_pair(a.pair_row_base, a.P, pair, row);
            const PairDesc& pd = a.pairs[pair];
            if (pd.same || !pd.owned_row) {
                n = 0;
            } else {
                tp = a.pair_of[pd.j * a.S + pd.p];
                s0 = a.row_start[row];
                r = (u32)(row - pd.row_base);
                gene = pd.gi_start + r;
            }
        }
        const int nmax = warp_max_over_groups<ROW_G>(n);
        int last = -1;  // largest k (column order) with an RBH
        int n_rbh = 0;
        double factor = 1.0;
        if (MODE == 2 && n > 0) factor = a.pair_factor[pair];
        for (int k0 = 0; k0 < nmax; k0 += ROW_G) {
            const int k = k0 + gl;
            bool rbh = false;
            if (k < n) {
                const u32 cw = a.cols[s0 + k];
                if (cw & OFG_FLAG_BH) {
                    if (tp >= 0) {
                        const PairDesc& td = a.pairs[tp];
                        const int64_t trow = td.row_base + (cw & OFG_COL_MASK);
                        const int64_t e = find_col(a.cols, a.row_start[trow], (int)a.rowlen[trow], r);
                        rbh = e >= 0 && (a.cols[e] & OFG_FLAG_BH);
                    } else {
                        rbh = (cw & OFG_FLAG_TBH) != 0;  // transposed pair lives in another context: flag was imported
                    }
                }
            }
            const u32 bal = (__ballot_sync(OFG_FULL, rbh) >> (sub * ROW_G)) & ((1u << ROW_G) - 1u);
            if (bal) last = k0 + (31 - __clz(bal));
            n_rbh += __popc(bal);
        }
        if (MODE == 1) {
            if (n > 0 && gl == 0) a.zrow[row] = n_rbh == 1 ? a.vals[s0 + last] : 0.0;
        } else if (last >= 0 && gl == 0) {
            atomic_min_pos_double(&a.most[gene], MODE == 2 ? __dmul_rn(factor, a.vals[s0 + last]) : a.vals[s0 + last]);
        }
    }
}

// --scores-v2, gathering.py:344-364: for query species p and every ordered pair (a, b) of other species, the
// (100 - 90)th percentile (numpy 'linear') of Z[a][i] * (1 / Z[b][i]) over the genes i where both RBH scores exist;
// the factor of species b is the minimum over a (with 1.0 for a == b and for "no common gene").  One CTA per
// (p, a, b): the ratios are recomputed from the two Z rows in every pass of an 8-bit radix select on their bit
// patterns (they stay L2-resident), no scratch.  pair_factor must hold 1.0 when this starts.
struct RatioArgs {
    const PairDesc* pairs;
    const int32_t* pair_of;     // [S*S]
    int S;
    const int32_t* owned;       // [n_owned] query species of this context
    int n_owned;
    const double* zrow;
    double* pair_factor;        // [P]
};

__device__ __forceinline__ bool v2_ratio(const double* za, const double* zb, int64_t i, double* out)
{
    const double x = za[i], y = zb[i];
    const double r = __dmul_rn(x, __ddiv_rn(1.0, y));  // Z[a] * Zr[b] with Zr = 1. / Z (:342,:349)
    *out = r;
    return r > 0.0 && r < INFINITY;                     // np.isfinite(ratios) & (ratios > 0); nan (0 * inf) fails both
}

__global__ void __launch_bounds__(256) v2_ratio_kernel(RatioArgs a)
{
    __shared__ u32 s_hist[256];
    __shared__ u64 s_red[8];
    __shared__ u32 s_cnt[8];
    __shared__ u64 s_prefix;
    __shared__ int s_kk;
    const int S = a.S;
    const int64_t task = blockIdx.x;
    const int p = a.owned[task / ((int64_t)S * S)];
    const int ka = (int)((task / S) % S), kb = (int)(task % S);
    if (ka == p || kb == p || ka == kb) return;
    const int pa = a.pair_of[p * S + ka], pb = a.pair_of[p * S + kb];
    if (pa < 0 || pb < 0) return;  // a missing species pair has no RBHs: no conversion, factor 1.0 (:352-353)
    const PairDesc& da = a.pairs[pa];
    const double* za = a.zrow + da.row_base;
    const double* zb = a.zrow + a.pairs[pb].row_base;
    const int64_t n_i = da.Gi;
    const int tid = threadIdx.x, lane = lane_id(), w = tid >> 5;
    // ---- how many genes have both scores
    u32 c = 0;
    for (int64_t i = tid; i < n_i; i += 256) { double r; c += v2_ratio(za, zb, i, &r) ? 1u : 0u; }
    c = warp_incl_scan(c);
    if (lane == 31) s_cnt[w] = c;
    __syncthreads();
    u32 n = 0;
    for (int i = 0; i < 8; i++) n += s_cnt[i];
    if (n == 0) return;
    // ---- numpy percentile(ratios, 10), method 'linear': virtual index (n - 1) * 0.1
    const double vi = __dmul_rn((double)(n - 1), 10.0 / 100.0);
    const bool top = vi >= (double)(n - 1);          // only n == 1: both neighbours are the last element
    const u32 k = top ? n - 1 : (u32)floor(vi);
    const double gamma = top ? 0.0 : __dsub_rn(vi, floor(vi));
    // ---- k-th smallest by radix select over the 64-bit patterns (positive doubles: unsigned order == numeric order)
    if (tid == 0) { s_prefix = 0; s_kk = (int)k; }
    for (int shift = 56; shift >= 0; shift -= 8) {
        s_hist[tid] = 0;
        __syncthreads();
        const u64 prefix = s_prefix;
        for (int64_t i = tid; i < n_i; i += 256) {
            double r;
            if (v2_ratio(za, zb, i, &r)) {
                const u64 b = (u64)__double_as_longlong(r);
                if (shift == 56 || (b >> (shift + 8)) == prefix) atomicAdd(&s_hist[(b >> shift) & 255u], 1u);
            }
        }
        __syncthreads();
        if (tid == 0) {
            int kk = s_kk;
            u32 run = 0;
            int bucket = 255;
            for (int d = 0; d < 256; d++) {
                if ((u32)kk < run + s_hist[d]) { bucket = d; break; }
                run += s_hist[d];
            }
            s_kk = kk - (int)run;
            s_prefix = (prefix << 8) | (u64)bucket;
        }
        __syncthreads();
    }
    const u64 ak = s_prefix;
    // ---- the next order statistic: ak again if it has duplicates reaching index k + 1, else the smallest value above
    u32 n_le = 0;
    u64 above = ~0ULL;
    for (int64_t i = tid; i < n_i; i += 256) {
        double r;
        if (v2_ratio(za, zb, i, &r)) {
            const u64 b = (u64)__double_as_longlong(r);
            n_le += b <= ak;
            if (b > ak && b < above) above = b;
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        n_le += __shfl_xor_sync(OFG_FULL, n_le, d);
        const u64 o = __shfl_xor_sync(OFG_FULL, above, d);
        above = o < above ? o : above;
    }
    __syncthreads();
    if (lane == 0) { s_cnt[w] = n_le; s_red[w] = above; }
    __syncthreads();
    if (tid == 0) {
        u32 tot = 0;
        u64 m = ~0ULL;
        for (int i = 0; i < 8; i++) { tot += s_cnt[i]; m = s_red[i] < m ? s_red[i] : m; }
        const u64 ak1 = (top || tot > k + 1 || m == ~0ULL) ? ak : m;
        const double va = __longlong_as_double((long long)ak), vb = __longlong_as_double((long long)ak1);
        // numpy _lerp: a + (b - a) t, and b - (b - a)(1 - t) where t >= 0.5
        const double diff = __dsub_rn(vb, va);
        double r = __dadd_rn(va, __dmul_rn(diff, gamma));
        if (gamma >= 0.5) r = __dsub_rn(vb, __dmul_rn(diff, __dsub_rn(1.0, gamma)));
        atomic_min_pos_double(&a.pair_factor[pb], r);  // column minimum over a (:364)
    }
}

__global__ void thresh_final_kernel(double* most, const double* best, const uint8_t* owned_gene, int64_t n)
{
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += (int64_t)gridDim.x * blockDim.x) {
        if (!owned_gene[g]) continue;
        const double md = most[g];
        if (md > 1e8) most[g] = __dadd_rn(best[g], 1e-6);  // best bit pattern 0 == 0.0 == "no hit"
    }
}

// ---------------------------------------------------------------- connect
// ---------------------------------------------------------------- text formatting (used by connect and emit)
// exact "%.3f" digits of a non-negative finite double: X = round-half-even(v * 1000) on the exact
// binary value (what glibc printf does); returns false if it does not fit 64 bits.
__host__ __device__ __forceinline__ bool ofg_fixed3(double v, unsigned long long* X)
{
    unsigned long long bits;
#if defined(__CUDA_ARCH__)
    bits = (unsigned long long)__double_as_longlong(v);
#else
    memcpy(&bits, &v, 8);
#endif
    const int eb = (int)((bits >> 52) & 0x7ff);
    unsigned long long mant = bits & 0x000fffffffffffffULL;
    int e;
    if (eb == 0) e = -1074; else { mant |= 0x0010000000000000ULL; e = eb - 1075; }
    const unsigned long long P = mant * 1000ULL;  // < 2^63
    if (e >= 0) {
        if (e > 0 && (P >> (64 - e)) != 0) return false;
        *X = P << e;
        return true;
    }
    const int sh = -e;
    if (sh >= 64) { *X = 0; return true; }
    unsigned long long q = P >> sh;
    const unsigned long long rem = P & ((1ULL << sh) - 1ULL), half = 1ULL << (sh - 1);
    if (rem > half || (rem == half && (q & 1ULL))) q++;
    *X = q;
    return true;
}

__host__ __device__ __forceinline__ int ofg_ndigits(unsigned long long v)
{
    int n = 1;
    while (v >= 10ULL) { v /= 10ULL; n++; }
    return n;
}

__host__ __device__ __forceinline__ int ofg_put_u64(char* out, unsigned long long v)
{
    char tmp[24];
    int n = 0;
    do { tmp[n++] = (char)('0' + (int)(v % 10ULL)); v /= 10ULL; } while (v);
    for (int k = 0; k < n; k++) out[k] = tmp[n - 1 - k];
    return n;
}

// "%d:%.3f " for (global column, weight); returns bytes (written only if out != nullptr)
__host__ __device__ __forceinline__ int ofg_format_entry(unsigned long long gcol, double w, char* out, bool* ok)
{
    unsigned long long X = 0;
    if (!ofg_fixed3(w, &X)) { *ok = false; X = 0; }
    const unsigned long long ip = X / 1000ULL;
    const int fr = (int)(X % 1000ULL);
    if (!out) return ofg_ndigits(gcol) + 1 + ofg_ndigits(ip) + 1 + 3 + 1;
    int n = ofg_put_u64(out, gcol);
    out[n++] = ':';
    n += ofg_put_u64(out + n, ip);
    out[n++] = '.';
    out[n++] = (char)('0' + fr / 100);
    out[n++] = (char)('0' + (fr / 10) % 10);
    out[n++] = (char)('0' + fr % 10);
    out[n++] = ' ';
    return n;
}

// Length of "%d:%.3f " without formatting it.  The integer part of the printed weight has more than j digits iff
// round-half-even(w * 1000) >= 10^(j+3), i.e. iff w >= T_j, the smallest double >= (2 * 10^(j+3) - 1) / 2000 (that
// rational is never a double, so there is no tie to break); T_j computed with exact rational arithmetic.
__device__ __constant__ double OFG_IP_DIGIT_THR[16] = {
    0x1.3ffbe76c8b43ap+3, 0x1.8fff7ced91688p+6, 0x1.f3ffef9db22d1p+9, 0x1.387ffef9db22ep+13, 0x1.869fffdf3b646p+16,
    0x1.e847fffbe76c9p+19, 0x1.312cffffbe76dp+23, 0x1.7d783ffff7ceep+26, 0x1.dcd64ffffef9ep+29, 0x1.2a05f1ffffefap+33,
    0x1.74876e7ffffe0p+36, 0x1.d1a94a1fffffcp+39, 0x1.2309ce5400000p+43, 0x1.6bcc41e900000p+46, 0x1.c6bf526340000p+49,
    0x1.1c37937e08000p+53};
__device__ __forceinline__ int ofg_entry_len(unsigned long long gcol, double w, bool* ok)
{
    int nd = 1;
#pragma unroll
    for (int j = 0; j < 16; j++) nd += w >= OFG_IP_DIGIT_THR[j];
    if (!(w < 18446744073709552.0)) *ok = false;  // w * 1000 >= 2^64: ofg_fixed3 does not fit (or w is not finite)
    return ofg_ndigits(gcol) + 1 + nd + 1 + 3 + 1;
}

// Edge bookkeeping of the graph rows, kept while the flags are set so that no separate counting sweep is needed.
// An entry is an edge if it or its transposed entry passes the threshold test.  The thread whose OWN entry passes
// looks the transposed entry up anyway; it evaluates that entry's test as well and books (a) its own edge with weight
// B or 2 B and (b) the transposed edge if the transposed test fails (otherwise the transposed thread books it).  No
// atomic has to return a value.  When the transposed pair lives in another context the REVERSE import books the
// second half (import_flags_kernel).
__device__ __forceinline__ void edge_account(u32* gene_cnt, u32* gene_bytes, u32* err_flags, int64_t gene, u64 gcol, double w, bool new_edge,
                                             double w_before)
{
    bool ok = true;
    if (new_edge) {
        atomicAdd(&gene_cnt[gene], 1u);
        atomicAdd(&gene_bytes[gene], (u32)ofg_entry_len(gcol, w, &ok));
    } else {  // the weight of an edge booked earlier changes from w_before to w
        const int d = ofg_entry_len(gcol, w, &ok) - ofg_entry_len(gcol, w_before, &ok);
        if (d) atomicAdd(&gene_bytes[gene], (u32)d);
    }
    if (!ok) atomicOr(err_flags, 4u);
}

template <int ROW_G>
__global__ void __launch_bounds__(256) connect_kernel(GraphArgs a)
{
    const int lane = lane_id(), sub = lane / ROW_G, gl = lane % ROW_G;
    const int64_t n_groups = (((int64_t)gridDim.x * blockDim.x) >> 5) * (32 / ROW_G);
    int pair = -1;
    for (int64_t row0 = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * (32 / ROW_G); row0 < a.n_rows; row0 += n_groups) {
        const int64_t row = row0 + sub;
        const int n = row < a.n_rows ? (int)a.rowlen[row] : 0;
        u32 nconn = 0;
        if (n > 0) {
            pair = advance_pair(a.pair_row_base, a.P, pair, row);
            const PairDesc& pd = a.pairs[pair];
            const u32 s0 = a.row_start[row];
            const u32 r = (u32)(row - pd.row_base);
            const double md = a.most[pd.gi_start + r];
            const int tp = a.pair_of[pd.j * a.S + pd.p];
            for (int k = gl; k < n; k += ROW_G) {
                const u32 cw = a.cols[s0 + k];
                const u32 c = cw & OFG_COL_MASK;
                const double v = a.vals[s0 + k];
                if (v >= md && !(pd.same && c == r)) {
                    nconn++;
                    atomicOr(&a.cols[s0 + k], OFG_FLAG_CONNECT);
                    bool t_pass = false;  // does the transposed entry pass its own test?
                    if (tp >= 0) {
                        const PairDesc& td = a.pairs[tp];
                        const int64_t trow = td.row_base + c;
                        const int64_t e = find_col(a.cols, a.row_start[trow], (int)a.rowlen[trow], r);
                        if (e >= 0) {
                            atomicOr(&a.cols[e], OFG_FLAG_REVERSE);
                            const double tv = a.vals[e];
                            t_pass = tv >= a.most[pd.gj_start + c];   // (c, r) is off the diagonal because (r, c) is
                            if (!t_pass) edge_account(a.gene_cnt, a.gene_bytes, a.err_flags, pd.gj_start + c, (u64)(pd.gi_start + r), tv, true, 0.0);
                        }
                    }
                    edge_account(a.gene_cnt, a.gene_bytes, a.err_flags, pd.gi_start + r, (u64)(pd.gj_start + c), t_pass ? __dmul_rn(2.0, v) : v, true, 0.0);
                }
            }
        }
        nconn = group_sum<ROW_G>(nconn);
        if (gl == 0 && n > 0) a.row_flagcnt[row] = nconn;
    }
}

// ---------------------------------------------------------------- sparse exchange (design B)
struct ExportArgs {
    int64_t n_rows;
    const u32* row_start;
    const u32* rowlen;
    const u32* cols;
    const PairDesc* pairs;
    const int64_t* pair_row_base;
    int P;
    const uint8_t* pair_export;  // [P] 1: pair (k, p) whose column species p is owned by another context
    u32 flag;                    // OFG_FLAG_BH or OFG_FLAG_CONNECT
    const u64* row_off;          // [n_rows + 1] exclusive scan of the per-row flag counts
    const u64* pair_off;         // [P] first list slot of the pair (pairs grouped by destination rank)
    int2* out;                   // (gene of remote species p, gene of local species k), global ids
};

// Deterministic single pass: every flagged entry of an exported pair goes to
// pair_off[pair] + (entries of earlier rows of the pair) + (rank inside the row).
__global__ void __launch_bounds__(256) export_flags_kernel(ExportArgs a)
{
    const int lane = lane_id();
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    int pair = -1;
    for (int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < a.n_rows; row += n_warps) {
        const u64 o0 = a.row_off[row];
        if (a.row_off[row + 1] == o0) continue;  // nothing flagged in this row
        pair = advance_pair(a.pair_row_base, a.P, pair, row);
        if (!a.pair_export[pair]) continue;
        const PairDesc& pd = a.pairs[pair];
        const int n = (int)a.rowlen[row];
        const u32 s0 = a.row_start[row];
        const int r = (int)(row - pd.row_base);
        u64 pos = a.pair_off[pair] + (o0 - a.row_off[pd.row_base]);
        for (int k0 = 0; k0 < n; k0 += 32) {
            const int k = k0 + lane;
            const u32 cw = k < n ? a.cols[s0 + k] : 0u;
            const bool on = (cw & a.flag) != 0;
            const u32 bal = __ballot_sync(OFG_FULL, on);
            if (on) a.out[pos + __popc(bal & lanemask_lt())] = make_int2((int)(pd.gj_start + (cw & OFG_COL_MASK)), (int)(pd.gi_start + r));
            pos += __popc(bal);
        }
    }
}

// ---- peer variant: the lists go straight into the destination's inbox (peer memory over NVLink, or another
// context of the same GPU), all offsets computed on the device - no host round trip, no collective for the payload.
// Inbox of a destination: n_ranks segments of seg_bytes; segment `src` = [u64 count][count x int2 entries].
struct PeerExportArgs {
    int64_t n_rows;
    const u32* row_start;
    const u32* rowlen;
    const u32* cols;
    const PairDesc* pairs;
    const int64_t* pair_row_base;
    int P;
    const int32_t* pair_dst;     // [P] destination rank of the pair's list, -1: not exported
    u32 flag;
    const u64* row_off;          // [n_rows + 1] exclusive scan of the per-row flag counts
    u64* pair_off;               // [P] first entry of the pair inside its destination's segment (export_layout_kernel)
    const int32_t* order;        // [n_exported] exported pairs grouped by destination
    const int32_t* dst_first;    // [n_ranks + 1] group boundaries in order
    unsigned char* const* inbox; // [n_ranks] base of every rank's inbox
    u64 seg_bytes;
    int my_rank, n_ranks;
    u32* err_flags;              // bit 5: a segment overflowed
};

// one warp per destination rank: offsets of the exported pairs inside the segment, entry count into the remote header
__global__ void __launch_bounds__(1024) export_layout_kernel(PeerExportArgs a)
{
    const int r = threadIdx.x >> 5, lane = lane_id();
    if (r >= a.n_ranks) return;
    const int g0 = a.dst_first[r], g1 = a.dst_first[r + 1];
    u64 run = 0;
    for (int b = g0; b < g1; b += 32) {
        const int i = b + lane;
        u64 cnt = 0;
        int pair = -1;
        if (i < g1) {
            pair = a.order[i];
            cnt = a.row_off[a.pair_row_base[pair + 1]] - a.row_off[a.pair_row_base[pair]];
        }
        u64 inc = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u64 t = __shfl_up_sync(OFG_FULL, inc, d);
            if (lane >= d) inc += t;
        }
        if (pair >= 0) a.pair_off[pair] = run + inc - cnt;
        run += __shfl_sync(OFG_FULL, inc, 31);
    }
    if (lane == 0 && r != a.my_rank) {
        const u64 cap = (a.seg_bytes - 8) / 8;
        if (run > cap) { atomicOr(a.err_flags, 32u); run = 0; }
        *reinterpret_cast<u64*>(a.inbox[r] + (u64)a.my_rank * a.seg_bytes) = run;
    }
}

// ROW_G lanes per matrix row (several rows in flight per warp: the work per row is a short chain of dependent loads)
template <int ROW_G>
__global__ void __launch_bounds__(256) export_peer_kernel(PeerExportArgs a)
{
    const int lane = lane_id(), sub = lane / ROW_G, gl = lane % ROW_G;
    const int64_t n_groups = (((int64_t)gridDim.x * blockDim.x) >> 5) * (32 / ROW_G);
    const u64 cap = (a.seg_bytes - 8) / 8;
    const u32 gmask = (ROW_G == 32) ? 0xffffffffu : (((1u << ROW_G) - 1u) << (sub * ROW_G));
    int pair = -1;
    for (int64_t row0 = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * (32 / ROW_G); row0 < a.n_rows; row0 += n_groups) {
        const int64_t row = row0 + sub;
        int n = 0;
        u64 pos = 0;
        u32 s0 = 0;
        int r = 0;
        int2* out = nullptr;
        int64_t gi = 0, gj = 0;
        if (row < a.n_rows) {
            const u64 o0 = a.row_off[row];
            if (a.row_off[row + 1] != o0) {   // something is flagged in this row
                pair = advance_pair(a.pair_row_base, a.P, pair, row);
                const int dst = a.pair_dst[pair];
                if (dst >= 0) {
                    const PairDesc& pd = a.pairs[pair];
                    n = (int)a.rowlen[row];
                    s0 = a.row_start[row];
                    r = (int)(row - pd.row_base);
                    gi = pd.gi_start; gj = pd.gj_start;
                    pos = a.pair_off[pair] + (o0 - a.row_off[pd.row_base]);
                    out = reinterpret_cast<int2*>(a.inbox[dst] + (u64)a.my_rank * a.seg_bytes + 8);
                }
            }
        }
        const int nmax = warp_max_over_groups<ROW_G>(n);
        for (int k0 = 0; k0 < nmax; k0 += ROW_G) {
            const int k = k0 + gl;
            const u32 cw = k < n ? a.cols[s0 + k] : 0u;
            const bool on = (cw & a.flag) != 0;
            const u32 bal = __ballot_sync(OFG_FULL, on) & gmask;
            const u64 o = pos + __popc(bal & lanemask_lt());
            if (on && o < cap) out[o] = make_int2((int)(gj + (cw & OFG_COL_MASK)), (int)(gi + r));
            pos += __popc(bal);
        }
    }
}

struct ImportArgs {
    const int2* list;
    int64_t n;
    const int64_t* seq_start;  // [S]
    int S;
    const int32_t* pair_of;
    const PairDesc* pairs;
    const u32* row_start;
    const u32* rowlen;
    u32* cols;
    u32 set_flag;              // OFG_FLAG_TBH or OFG_FLAG_REVERSE
    u32* err_flags;            // bit 3: an entry addressed a pair this context does not hold
    const double* vals;        // REVERSE imports also keep the edge bookkeeping (edge_account)
    u32* gene_cnt;
    u32* gene_bytes;
    const unsigned char* inbox;  // peer variant: this rank's inbox (n_ranks segments), else nullptr
    u64 seg_bytes;
    int my_rank;
};

__global__ void __launch_bounds__(256) import_flags_kernel(ImportArgs a)
{
    // peer variant (inbox != nullptr): blockIdx.y = source rank, its segment holds [u64 count][entries]
    const int2* list = a.list;
    int64_t n = a.n;
    if (a.inbox) {
        if ((int)blockIdx.y == a.my_rank) return;
        const unsigned char* seg = a.inbox + (u64)blockIdx.y * a.seg_bytes;
        n = (int64_t)*reinterpret_cast<const u64*>(seg);
        list = reinterpret_cast<const int2*>(seg + 8);
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int2 e = list[i];
        const int p = upper_bound_dev(a.seq_start, a.S, (int64_t)e.x) - 1;  // local (row) species
        const int k = upper_bound_dev(a.seq_start, a.S, (int64_t)e.y) - 1;  // species of the sender
        const int pair = a.pair_of[p * a.S + k];
        if (pair < 0) { atomicOr(a.err_flags, 8u); continue; }
        const PairDesc& pd = a.pairs[pair];
        const int64_t row = pd.row_base + ((int64_t)e.x - pd.gi_start);
        const int64_t slot = find_col(a.cols, a.row_start[row], (int)a.rowlen[row], (u32)((int64_t)e.y - pd.gj_start));
        if (slot >= 0) {  // no stored hit in this direction: nothing to mark
            const u32 old = atomicOr(&a.cols[slot], a.set_flag);
            if (a.set_flag == OFG_FLAG_REVERSE && !(old & OFG_FLAG_REVERSE)) {  // connect_kernel has finished: CONNECT is final
                const double v = a.vals[slot];
                if (old & OFG_FLAG_CONNECT) edge_account(a.gene_cnt, a.gene_bytes, a.err_flags, (int64_t)e.x, (u64)e.y, __dmul_rn(2.0, v), false, v);
                else edge_account(a.gene_cnt, a.gene_bytes, a.err_flags, (int64_t)e.x, (u64)e.y, v, true, 0.0);
            }
        }
    }
}

// ---------------------------------------------------------------- emission
struct EmitArgs {
    GraphArgs g;
    int64_t n_genes;          // owned genes, list order
    const int64_t* gene_of_local;     // [owned genes] global gene id
    const int32_t* species_of_local;  // [owned genes] list position p
    const int64_t* seq_start; // [S] seqStartingIndices
    u32* gene_cnt;            // [n_genes] edges of the graph row
    u32* gene_bytes;          // [n_genes] text bytes of the graph row
    const u64* cnt_off;       // [n_genes + 1]
    const u64* byte_off;      // [n_genes + 1]
    int32_t* out_indices;
    double* out_data;
    char* out_text;
    u32* err_flags;           // bit 2: a weight could not be formatted; bit 6: the output buffers are too small
    u64 cap_edges, cap_text;  // capacity of out_indices / out_data and of out_text (checked on the device: the sizes are
                              // known only there, the host reads them back once at the end of the build)
};

// Per graph row (owned gene, list order): number of edges and text bytes = "%d    " + entries + "$\n" (gathering.py:42-47),
// from the counters kept by edge_account.
__global__ void emit_sizes_kernel(EmitArgs a, const u32* gene_cnt_global, const u32* gene_bytes_global)
{
    for (int64_t lg = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; lg < a.n_genes; lg += (int64_t)gridDim.x * blockDim.x) {
        const int64_t gene = a.gene_of_local[lg];
        a.gene_cnt[lg] = gene_cnt_global[gene];
        a.gene_bytes[lg] = gene_bytes_global[gene] + (u32)ofg_ndigits((unsigned long long)gene) + 4u + 2u;
    }
}

constexpr int EMIT_WARPS = 8;
constexpr int EMIT_LIST = 64;  // edges buffered per warp before they are formatted

// One warp per graph row (gene).  The row's entries live in S matrix rows (one per target species); the
// lanes first fetch those S row descriptors in parallel, then sweep the concatenated entries 32 at a time
// (independent loads), collect the flagged ones in order into a small shared list and format that list
// densely (one edge per lane).  WRITE = false only counts edges and text bytes.
template <bool WRITE>
__global__ void __launch_bounds__(EMIT_WARPS * 32) emit_kernel(EmitArgs a)
{
    __shared__ u32 s_s0[EMIT_WARPS][32];
    __shared__ u32 s_pre[EMIT_WARPS][33];
    __shared__ u64 s_goff[EMIT_WARPS][32];
    __shared__ u64 s_gc[EMIT_WARPS][EMIT_LIST];
    __shared__ double s_w[EMIT_WARPS][EMIT_LIST];
    const int lane = lane_id(), w = threadIdx.x >> 5;
    const int S = a.g.S;
    if (WRITE && (a.cnt_off[a.n_genes] > a.cap_edges || a.byte_off[a.n_genes] > a.cap_text)) {
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(a.err_flags, 64u);  // the host grows the buffers and emits again
        return;
    }
    const int64_t n_warps = (int64_t)gridDim.x * EMIT_WARPS;
    for (int64_t lg = (int64_t)blockIdx.x * EMIT_WARPS + w; lg < a.n_genes; lg += n_warps) {
        const int p = a.species_of_local[lg];
        const int64_t gene = a.gene_of_local[lg];
        u32 cnt = 0, bytes = 0;
        u64 ocnt = 0, obyte = 0;
        if (WRITE) { ocnt = a.cnt_off[lg]; obyte = a.byte_off[lg]; }
        {   // "%d    " (gathering.py:42)
            const int nd = ofg_ndigits((unsigned long long)gene);
            if (WRITE && lane == 0) {
                char* o = a.out_text + obyte;
                const int n = ofg_put_u64(o, (unsigned long long)gene);
                o[n] = ' '; o[n + 1] = ' '; o[n + 2] = ' '; o[n + 3] = ' ';
            }
            bytes += nd + 4;
            obyte += nd + 4;
        }
        int nlist = 0;
        // flush the buffered edges: one edge per lane
        auto flush = [&]() {
            for (int b0 = 0; b0 < nlist; b0 += 32) {
                const int i = b0 + lane;
                const bool on = i < nlist;
                int len = 0;
                bool ok = true;
                u64 gc = 0;
                double wt = 0.0;
                if (on) {
                    gc = s_gc[w][i];
                    wt = s_w[w][i];
                    len = ofg_format_entry(gc, wt, nullptr, &ok);
                    if (!ok) atomicOr(a.err_flags, 4u);
                }
                const u32 binc = warp_incl_scan((u32)len);
                const u32 btot = __shfl_sync(OFG_FULL, binc, 31);
                if (WRITE && on) {
                    a.out_indices[ocnt + i - b0] = (int32_t)gc;
                    a.out_data[ocnt + i - b0] = wt;
                    ofg_format_entry(gc, wt, a.out_text + obyte + (binc - (u32)len), &ok);
                }
                const int nb = nlist - b0 < 32 ? nlist - b0 : 32;
                ocnt += nb;
                obyte += btot;
                bytes += btot;
                cnt += nb;
            }
            nlist = 0;
            __syncwarp();
        };
        for (int j0 = 0; j0 < S; j0 += 32) {
            // row descriptors of up to 32 target species
            const int j = j0 + lane;
            u32 n = 0, s0 = 0;
            if (j < S) {
                const int pair = a.g.pair_of[p * S + j];
                if (pair >= 0) {
                    const PairDesc& pd = a.g.pairs[pair];
                    const int64_t row = pd.row_base + (gene - pd.gi_start);
                    n = a.g.rowlen[row];
                    s0 = a.g.row_start[row];
                }
                s_goff[w][lane] = (u64)a.seq_start[j];
            }
            const u32 inc = warp_incl_scan(n);
            s_s0[w][lane] = s0;
            s_pre[w][lane] = inc - n;
            if (lane == 31) s_pre[w][32] = inc;
            __syncwarp();
            const u32 total = s_pre[w][32];
            for (u32 e0 = 0; e0 < total; e0 += 32) {
                const u32 e = e0 + lane;
                bool on = false;
                u64 gc = 0;
                double wt = 0.0;
                if (e < total) {
                    // species slot holding entry e: last jj with pre[jj] <= e
                    int lo = 0, hi = 32;
                    while (hi - lo > 1) {
                        const int mid = (lo + hi) >> 1;
                        if (s_pre[w][mid] <= e) lo = mid; else hi = mid;
                    }
                    const u32 slot = s_s0[w][lo] + (e - s_pre[w][lo]);
                    const u32 cw = a.g.cols[slot];
                    const int mult = ((cw & OFG_FLAG_CONNECT) ? 1 : 0) + ((cw & OFG_FLAG_REVERSE) ? 1 : 0);
                    if (mult) {
                        on = true;
                        gc = s_goff[w][lo] + (cw & OFG_COL_MASK);
                        wt = __dmul_rn((double)mult, a.g.vals[slot]);  // (connect + connect^T) .* B
                    }
                }
                const u32 bal = __ballot_sync(OFG_FULL, on);
                const int nb = __popc(bal);
                if (nlist + nb > EMIT_LIST) flush();
                if (on) {
                    const int pos = nlist + __popc(bal & lanemask_lt());
                    s_gc[w][pos] = gc;
                    s_w[w][pos] = wt;
                }
                nlist += nb;
                __syncwarp();
            }
            __syncwarp();
        }
        flush();
        // "$\n" (gathering.py:47)
        if (WRITE && lane == 0) { a.out_text[obyte] = '$'; a.out_text[obyte + 1] = '\n'; }
        bytes += 2;
        if (!WRITE && lane == 0) { a.gene_cnt[lg] = cnt; a.gene_bytes[lg] = bytes; }
    }
}

// ---------------------------------------------------------------- homology graph (`--c-homologs`)
// gathering.py:51-73 WriteGraph_perSpecies_homology (gathering_version (3, 2), dispatched at :520-521): no best hits, no
// thresholds - for query species p the graph row of gene q holds, per target species j, the structural union of row q of
// B_pj and column q of B_jp, every entry with weight 1.0.  The entries that exist only in the transposed direction
// ("extras") are collected per matrix row with a counting pass, a scan and a fill pass, then merged with the row's own
// (sorted) columns while the text is written.
struct HomArgs {
    GraphArgs g;
    u32* extra_cnt;      // [n_rows] extras of every matrix row
    const u32* extra_off;  // [n_rows + 1] exclusive scan
    u32* extra_fill;     // [n_rows]
    u32* extra_col;      // [total extras]
};

template <bool FILL>
__global__ void __launch_bounds__(256) hom_extras_kernel(HomArgs a)
{
    // ROW_G = 4 lanes per matrix row (x of pair (j, p)); every entry (x, y) looks for (y, x) in pair (p, j)
    constexpr int ROW_G = 4;
    const int lane = lane_id(), sub = lane / ROW_G, gl = lane % ROW_G;
    const int64_t n_groups = (((int64_t)gridDim.x * blockDim.x) >> 5) * (32 / ROW_G);
    int pair = -1;
    for (int64_t row0 = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * (32 / ROW_G); row0 < a.g.n_rows; row0 += n_groups) {
        const int64_t row = row0 + sub;
        const int n = row < a.g.n_rows ? (int)a.g.rowlen[row] : 0;
        if (n > 0) {
            pair = advance_pair(a.g.pair_row_base, a.g.P, pair, row);
            const PairDesc& pd = a.g.pairs[pair];
            const int tp = a.g.pair_of[pd.j * a.g.S + pd.p];
            if (tp >= 0) {
                const PairDesc& td = a.g.pairs[tp];
                const u32 s0 = a.g.row_start[row];
                const u32 x = (u32)(row - pd.row_base);
                for (int k = gl; k < n; k += ROW_G) {
                    if (!(a.g.vals[s0 + k] > 0.0)) continue;            // (w1 + w2tr > 0): stored normalised scores are positive
                    const u32 y = a.g.cols[s0 + k] & OFG_COL_MASK;
                    const int64_t trow = td.row_base + y;
                    const int64_t e = find_col(a.g.cols, a.g.row_start[trow], (int)a.g.rowlen[trow], x);
                    if (e < 0 || !(a.g.vals[e] > 0.0)) {
                        if (FILL) a.extra_col[a.extra_off[trow] + atomicAdd(&a.extra_fill[trow], 1u)] = x;
                        else atomicAdd(&a.extra_cnt[trow], 1u);
                    }
                }
            }
        }
    }
}

struct HomEmitArgs {
    GraphArgs g;
    const u32* extra_off;
    const u32* extra_col;
    int64_t n_genes;
    const int64_t* gene_of_local;
    const int32_t* species_of_local;
    const int64_t* seq_start;
    u32* gene_cnt;            // [n_genes] edges of the graph row
    u32* gene_bytes;          // [n_genes] text bytes of the graph row
    const u64* cnt_off;
    const u64* byte_off;
    int32_t* out_indices;
    double* out_data;
    char* out_text;
    u32* err_flags;           // bit 6: the output buffers are too small
    u64 cap_edges, cap_text;
};

constexpr int HOM_WARPS = 4;
constexpr int HOM_LIST = 256;   // merged entries staged per warp and pass

// One warp per graph row.  WRITE = false only counts edges and text bytes.
template <bool WRITE>
__global__ void __launch_bounds__(HOM_WARPS * 32) hom_emit_kernel(HomEmitArgs a)
{
    __shared__ u32 s_list[HOM_WARPS][HOM_LIST];
    const int lane = lane_id(), w = threadIdx.x >> 5;
    const int S = a.g.S;
    if (WRITE && (a.cnt_off[a.n_genes] > a.cap_edges || a.byte_off[a.n_genes] > a.cap_text)) {
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(a.err_flags, 64u);
        return;
    }
    const int64_t n_warps = (int64_t)gridDim.x * HOM_WARPS;
    for (int64_t lg = (int64_t)blockIdx.x * HOM_WARPS + w; lg < a.n_genes; lg += n_warps) {
        const int p = a.species_of_local[lg];
        const int64_t gene = a.gene_of_local[lg];
        u64 ocnt = 0, obyte = 0;
        u32 cnt = 0, bytes = 0;
        if (WRITE) { ocnt = a.cnt_off[lg]; obyte = a.byte_off[lg]; }
        {
            const int nd = ofg_ndigits((unsigned long long)gene);
            if (WRITE && lane == 0) {
                char* o = a.out_text + obyte;
                const int n = ofg_put_u64(o, (unsigned long long)gene);
                o[n] = ' '; o[n + 1] = ' '; o[n + 2] = ' '; o[n + 3] = ' ';
            }
            bytes += nd + 4;
            obyte += nd + 4;
        }
        for (int j = 0; j < S; j++) {
            const int pair = a.g.pair_of[p * S + j];
            if (pair < 0) continue;
            const PairDesc& pd = a.g.pairs[pair];
            const int64_t row = pd.row_base + (gene - pd.gi_start);
            const u32 s0 = a.g.row_start[row];
            const int n = (int)a.g.rowlen[row];
            const u32 x0 = a.extra_off[row];
            const int ne = (int)(a.extra_off[row + 1] - x0);
            const u64 goff = (u64)a.seq_start[j];
            const int total = n + ne;
            if (!WRITE) {
                // sizes need no merge: "%d:%.3f " of weight 1.0 is the column's digits + ":1.000 "
                u32 b = 0, c = 0;
                for (int k = lane; k < total; k += 32) {
                    u32 col;
                    bool on = true;
                    if (k < n) { col = a.g.cols[s0 + k] & OFG_COL_MASK; on = a.g.vals[s0 + k] > 0.0; }
                    else col = a.extra_col[x0 + (k - n)];
                    if (on) { b += (u32)ofg_ndigits(goff + col) + 7u; c++; }
                }
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) { b += __shfl_xor_sync(OFG_FULL, b, d); c += __shfl_xor_sync(OFG_FULL, c, d); }
                bytes += b;
                cnt += c;
                continue;
            }
            // merged order: rank of a column = its index among the kept columns + extras below it; rank of an extra = columns below it
            // + extras below it (the two sets are disjoint).  Passes of HOM_LIST ranks are staged in shared memory and formatted.
            int kept = 0;   // kept columns (positive values) - all of them in practice
            for (int k0 = 0; k0 < n; k0 += 32) {
                const int k = k0 + lane;
                kept += __popc(__ballot_sync(OFG_FULL, k < n && a.g.vals[s0 + k] > 0.0));
            }
            const int tot = kept + ne;
            for (int r0 = 0; r0 < tot; r0 += HOM_LIST) {
                for (int k = lane; k < total; k += 32) {
                    u32 col;
                    int rank;
                    if (k < n) {
                        if (!(a.g.vals[s0 + k] > 0.0)) continue;
                        col = a.g.cols[s0 + k] & OFG_COL_MASK;
                        rank = k;
                        if (kept != n) { rank = 0; for (int q = 0; q < k; q++) rank += a.g.vals[s0 + q] > 0.0; }
                        for (int q = 0; q < ne; q++) rank += a.extra_col[x0 + q] < col;
                    } else {
                        col = a.extra_col[x0 + (k - n)];
                        int lo = 0, hi = n;   // columns below col (sorted)
                        while (lo < hi) {
                            const int mid = (lo + hi) >> 1;
                            if ((a.g.cols[s0 + mid] & OFG_COL_MASK) < col) lo = mid + 1; else hi = mid;
                        }
                        rank = lo;
                        if (kept != n) { rank = 0; for (int q = 0; q < lo; q++) rank += a.g.vals[s0 + q] > 0.0; }
                        for (int q = 0; q < ne; q++) rank += a.extra_col[x0 + q] < col;
                    }
                    if (rank >= r0 && rank < r0 + HOM_LIST) s_list[w][rank - r0] = col;
                }
                __syncwarp();
                const int m = tot - r0 < HOM_LIST ? tot - r0 : HOM_LIST;
                for (int b0 = 0; b0 < m; b0 += 32) {
                    const int i = b0 + lane;
                    const bool on = i < m;
                    u64 gc = 0;
                    int len = 0;
                    if (on) { gc = goff + s_list[w][i]; len = ofg_ndigits(gc) + 7; }
                    const u32 binc = warp_incl_scan((u32)len);
                    const u32 btot = __shfl_sync(OFG_FULL, binc, 31);
                    if (on) {
                        a.out_indices[ocnt + i - b0] = (int32_t)gc;
                        a.out_data[ocnt + i - b0] = 1.0;
                        char* o = a.out_text + obyte + (binc - (u32)len);
                        const int nd = ofg_put_u64(o, gc);
                        o[nd] = ':'; o[nd + 1] = '1'; o[nd + 2] = '.'; o[nd + 3] = '0'; o[nd + 4] = '0'; o[nd + 5] = '0'; o[nd + 6] = ' ';
                    }
                    const int nb = m - b0 < 32 ? m - b0 : 32;
                    ocnt += nb;
                    obyte += btot;
                }
                __syncwarp();
            }
        }
        if (WRITE && lane == 0) { a.out_text[obyte] = '$'; a.out_text[obyte + 1] = '\n'; }
        bytes += 2;
        if (!WRITE && lane == 0) { a.gene_cnt[lg] = cnt; a.gene_bytes[lg] = bytes; }
    }
}
