// Stage RAW, part 2: COO -> per-row sorted, max-deduplicated CSR ("slotted": row r owns
// [row_start[r], row_start[r+1]) and uses the first rowlen[r] slots).  Replaces the LIL matrix
// semantics of blast_file_processor.py:55,87-88 (B[q,s] = max over duplicate HSP lines, columns
// ascending within a row).
//
//  scan      : device-wide exclusive scan of the row histogram -> row_start
//  scatter   : counting sort by row (warp-aggregated slot allocation)
//  rowsort   : one warp per row, rank sort in shared memory on (col asc, score desc), keep the head
//              of every column run; also emits the length-product key Lf = L_i[q] * L_j[s]
//              (gathering.py:85-90,142) for the normalisation sort.
#pragma once
#include "ofg_common.cuh"

// ---------------------------------------------------------------- device-wide exclusive scan
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_CHUNK = SCAN_THREADS * SCAN_ITEMS;

template <typename TO>
__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const u32* in, int64_t n, TO* partial)
{
    __shared__ TO s_w[SCAN_THREADS / 32];
    for (int64_t chunk = blockIdx.x; chunk * SCAN_CHUNK < n; chunk += gridDim.x) {
        const int64_t base = chunk * SCAN_CHUNK;
        TO s = 0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            int64_t i = base + k * SCAN_THREADS + threadIdx.x;
            if (i < n) s += (TO)in[i];
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) s += __shfl_down_sync(OFG_FULL, s, d);
        if (lane_id() == 0) s_w[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            TO t = 0;
            for (int w = 0; w < SCAN_THREADS / 32; w++) t += s_w[w];
            partial[chunk] = t;
        }
        __syncthreads();
    }
}

template <typename TO>
__global__ void __launch_bounds__(1024) scan_partials_kernel(TO* partial, int64_t n_chunks, TO* total_out)
{
    // single CTA: exclusive scan of the chunk sums, in place
    __shared__ TO s_w[32];
    __shared__ TO s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int64_t b = 0; b < n_chunks; b += 1024) {
        int64_t i = b + threadIdx.x;
        TO v = i < n_chunks ? partial[i] : 0;
        TO inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            TO t = __shfl_up_sync(OFG_FULL, inc, d);
            if ((int)lane_id() >= d) inc += t;
        }
        if (lane_id() == 31) s_w[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) {
            TO s = s_w[threadIdx.x];
            TO si = s;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                TO t = __shfl_up_sync(OFG_FULL, si, d);
                if ((int)lane_id() >= d) si += t;
            }
            s_w[threadIdx.x] = si - s;
        }
        __syncthreads();
        TO excl = s_carry + s_w[threadIdx.x >> 5] + inc - v;
        if (i < n_chunks) partial[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = s_carry;
}

// out has n + 1 entries when write_total != 0 (out[n] = grand total)
template <typename TO>
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const u32* in, int64_t n, const TO* partial, TO* out,
                                                                  int write_total)
{
    __shared__ u32 s_scan[SCAN_THREADS / 32 + 1];
    for (int64_t chunk = blockIdx.x; chunk * SCAN_CHUNK < n; chunk += gridDim.x) {
        const int64_t base = chunk * SCAN_CHUNK + (int64_t)threadIdx.x * SCAN_ITEMS;
        u32 v[SCAN_ITEMS];
        u32 s = 0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            int64_t i = base + k;
            v[k] = i < n ? in[i] : 0u;
            s += v[k];
        }
        u32 tot;
        u32 ex = block_excl_scan<SCAN_THREADS>(s, s_scan, &tot);
        TO run = partial[chunk] + (TO)ex;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            int64_t i = base + k;
            if (i < n) out[i] = run;
            run += (TO)v[k];
            if (write_total && i == n - 1) out[n] = run;
        }
    }
}

// ---------------------------------------------------------------- scatter (counting sort by row)
__global__ void __launch_bounds__(256) scatter_rows_kernel(const u32* __restrict__ coo_row, const u32* __restrict__ coo_col,
                                                           const double* __restrict__ coo_val, u64 n,
                                                           const u32* __restrict__ row_start, u32* rowfill, u32* cols,
                                                           double* vals)
{
    const u64 stride = (u64)gridDim.x * blockDim.x;
    const u64 n_round = (n + 31) & ~(u64)31;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        const u32 row = i < n ? coo_row[i] : OFG_ROW_SENTINEL;
        // records of one query are consecutive: aggregate the slot reservation over runs of equal rows
        // (MATCH.ANY retires one warp instruction per ~64 SM cycles on B200, tools/micro/match_rate.cu)
        const int lane = (int)lane_id();
        const u32 prev_row = __shfl_up_sync(OFG_FULL, row, 1);
        const u32 heads = __ballot_sync(OFG_FULL, lane == 0 || row != prev_row);
        const int leader = 31 - __clz(heads & (lanemask_lt() | (1u << lane)));   // first lane of this lane's run
        const u32 later = heads & ~(lanemask_lt() | (1u << lane));
        const int run_end = later ? __ffs(later) - 1 : 32;
        u32 basev = 0;
        if (row != OFG_ROW_SENTINEL && lane == leader) basev = atomicAdd(&rowfill[row], (u32)(run_end - leader));
        basev = __shfl_sync(OFG_FULL, basev, leader);
        if (row != OFG_ROW_SENTINEL) {
            const u32 pos = row_start[row] + basev + (u32)(lane - leader);
            cols[pos] = coo_col[i];
            vals[pos] = coo_val[i];
        }
    }
}

// ---------------------------------------------------------------- per-row sort + max-dedup
constexpr int ROWSORT_MAXN = 256;
constexpr int ROWSORT_WARPS = 4;

struct RowSortArgs {
    int64_t n_rows;
    const u32* row_start;      // [n_rows + 1]
    u32* rowlen;               // [n_rows] out: stored entries after dedup
    u32* cols;
    double* vals;
    const PairDesc* pairs;
    const int64_t* pair_row_base;  // [P + 1]
    int P;
    const double* lengths;
    u32* key;                  // out: Lf per slot (sentinel for unused slots)
    double* keyval;            // out: score per slot
    u32 key_sentinel;
    u32* long_rows;            // out: rows with more than ROWSORT_MAXN records
    u32* n_long;
    u64* pair_nnz;             // [P]
    u32* scratch_cols;         // long-row scratch (same indexing as cols)
    double* scratch_vals;
};

__device__ __forceinline__ bool rs_before(u32 cj, double vj, int j, u32 ck, double vk, int k)
{
    return (cj < ck) || (cj == ck && (vj > vk || (vj == vk && j < k)));
}

__global__ void __launch_bounds__(ROWSORT_WARPS * 32) rowsort_kernel(RowSortArgs a)
{
    __shared__ __align__(16) u32 s_c[ROWSORT_WARPS][ROWSORT_MAXN + 4];
    __shared__ double s_v[ROWSORT_WARPS][ROWSORT_MAXN];
    __shared__ u32 s_oc[ROWSORT_WARPS][ROWSORT_MAXN];
    __shared__ u64 s_ov[ROWSORT_WARPS][ROWSORT_MAXN];  // best score (bit pattern, > 0) per output slot
    const int w = threadIdx.x >> 5, lane = lane_id();
    u32* c = s_c[w];
    double* v = s_v[w];
    u32* oc = s_oc[w];
    u64* ov = s_ov[w];
    const int64_t n_warps = (int64_t)gridDim.x * ROWSORT_WARPS;
    int pair = -1;
    for (int64_t row = (int64_t)blockIdx.x * ROWSORT_WARPS + w; row < a.n_rows; row += n_warps) {
        const u32 s0 = a.row_start[row];
        const int n = (int)(a.row_start[row + 1] - s0);
        if (n == 0) {
            if (lane == 0) a.rowlen[row] = 0;
            continue;
        }
        if (n > ROWSORT_MAXN) {
            if (lane == 0) a.long_rows[atomicAdd(a.n_long, 1u)] = (u32)row;
            continue;
        }
        // rows are visited in increasing order: the pair index only moves forward
        if (pair < 0) pair = upper_bound_dev(a.pair_row_base, a.P + 1, row) - 1;
        while (row >= a.pair_row_base[pair + 1]) pair++;
        const PairDesc& pd = a.pairs[pair];
        const double Lq = a.lengths[pd.li_off + (row - pd.row_base)];
        const double* Lh = a.lengths + pd.lj_off;
        for (int k = lane; k < n; k += 32) {
            c[k] = a.cols[s0 + k];
            v[k] = a.vals[s0 + k];
            ov[k] = 0ULL;
        }
        if (lane < 4) c[n + lane] = 0xffffffffu;  // padding for the 4-wide compares (never < or == a real column)
        __syncwarp();
        const uint4* c4 = reinterpret_cast<const uint4*>(c);
        const int n4 = (n + 3) >> 2;
        for (int k = lane; k < n; k += 32) {
            const u32 ck = c[k];
            const double vk = v[k];
            int rank = 0;  // number of strictly smaller columns: duplicates of a column share the slot
            for (int j4 = 0; j4 < n4; j4++) {
                const uint4 q = c4[j4];
                rank += (q.x < ck) + (q.y < ck) + (q.z < ck) + (q.w < ck);
            }
            oc[rank] = ck;
            atomicMax(&ov[rank], (u64)__double_as_longlong(vk));  // B[q,s] = max over duplicate HSP lines
        }
        __syncwarp();
        int outn = 0;
        for (int k0 = 0; k0 < n; k0 += 32) {
            const int k = k0 + lane;
            const bool h = k < n && ov[k] != 0ULL;
            const u32 bal = __ballot_sync(OFG_FULL, h);
            if (h) {
                const u32 pos = s0 + outn + __popc(bal & lanemask_lt());
                const u32 col = oc[k];
                const double val = __longlong_as_double((long long)ov[k]);
                a.cols[pos] = col;
                a.vals[pos] = val;
                a.key[pos] = (u32)(Lq * Lh[col]);
                a.keyval[pos] = val;
            }
            outn += __popc(bal);
        }
        for (int k = outn + lane; k < n; k += 32) a.key[s0 + k] = a.key_sentinel;
        if (lane == 0) {
            a.rowlen[row] = (u32)outn;
            atomicAdd(&a.pair_nnz[pair], (u64)outn);
        }
        __syncwarp();
    }
}

// Rows longer than ROWSORT_MAXN: one CTA per row, O(n^2) rank sort through global scratch.  Rare
// (a query with hundreds of targets in one species); correctness path, not a throughput path.
__global__ void __launch_bounds__(256) rowsort_long_kernel(RowSortArgs a)
{
    __shared__ u32 s_scan[256 / 32 + 1];
    __shared__ u32 s_outn;
    const u32 n_long = *a.n_long;
    for (u32 li = blockIdx.x; li < n_long; li += gridDim.x) {
        const int64_t row = a.long_rows[li];
        const u32 s0 = a.row_start[row];
        const int n = (int)(a.row_start[row + 1] - s0);
        const int pair = upper_bound_dev(a.pair_row_base, a.P + 1, row) - 1;
        const PairDesc& pd = a.pairs[pair];
        const double Lq = a.lengths[pd.li_off + (row - pd.row_base)];
        const double* Lh = a.lengths + pd.lj_off;
        for (int k = threadIdx.x; k < n; k += blockDim.x) {
            const u32 ck = a.cols[s0 + k];
            const double vk = a.vals[s0 + k];
            int rank = 0;
            bool head = true;
            for (int j = 0; j < n; j++) {
                const u32 cj = a.cols[s0 + j];
                const bool b = rs_before(cj, a.vals[s0 + j], j, ck, vk, k);
                rank += b;
                head = head && !(b && cj == ck);
            }
            a.scratch_cols[s0 + rank] = ck | (head ? 0x80000000u : 0u);
            a.scratch_vals[s0 + rank] = vk;
        }
        if (threadIdx.x == 0) s_outn = 0;
        __syncthreads();
        for (int k0 = 0; k0 < n; k0 += 256) {
            const int k = k0 + threadIdx.x;
            const u32 oc = k < n ? a.scratch_cols[s0 + k] : 0u;
            const bool h = (oc >> 31) != 0;
            u32 tot;
            const u32 ex = block_excl_scan<256>(h ? 1u : 0u, s_scan, &tot);
            const u32 base = s_outn;
            if (h) {
                const u32 pos = s0 + base + ex;
                const u32 col = oc & 0x7fffffffu;
                const double val = a.scratch_vals[s0 + k];
                a.cols[pos] = col;
                a.vals[pos] = val;
                a.key[pos] = (u32)(Lq * Lh[col]);
                a.keyval[pos] = val;
            }
            __syncthreads();
            if (threadIdx.x == 0) s_outn = base + tot;
            __syncthreads();
        }
        const int outn = (int)s_outn;
        for (int k = outn + threadIdx.x; k < n; k += blockDim.x) a.key[s0 + k] = a.key_sentinel;
        if (threadIdx.x == 0) {
            a.rowlen[row] = (u32)outn;
            atomicAdd(&a.pair_nnz[pair], (u64)outn);
        }
        __syncthreads();
    }
}
