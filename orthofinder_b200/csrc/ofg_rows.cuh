// Stage RAW, part 2: COO -> per-row sorted, max-deduplicated CSR ("slotted": row r owns
// [row_start[r], row_start[r+1]) and uses the first rowlen[r] slots).  Replaces the LIL matrix
// semantics of blast_file_processor.py:55,87-88 (B[q,s] = max over duplicate HSP lines, columns
// ascending within a row).
//
//  scan      : device-wide exclusive scan of the row histogram -> row_start
//  scatter   : counting sort by row (slots reserved per run of equal rows)
//  rowsort   : one warp per row: rows of up to 64 records are sorted in registers by a bitonic shuffle network
//              on packed (column, input index) words, longer ones by a rank sort in shared memory; duplicates
//              of a column share a slot (max score); also emits the length-product key Lf = L_i[q] * L_j[s]
//              (gathering.py:85-90,142) for the normalisation sort.
#pragma once
#include "ofg_common.cuh"

// ---------------------------------------------------------------- device-wide exclusive scan
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_CHUNK = SCAN_THREADS * SCAN_ITEMS;

// The element count of a scan may live on the device (n_dev != nullptr: n = *n_dev << n_shift, clamped to the
// host-side bound n): the sizes of later stages follow from earlier kernels without a host round trip.
__device__ __forceinline__ int64_t scan_count(int64_t n, const int64_t* n_dev, int n_shift)
{
    if (!n_dev) return n;
    const int64_t m = *n_dev << n_shift;
    return m < n ? m : n;
}

template <typename TO>
__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const u32* in, int64_t n, TO* partial, const int64_t* n_dev, int n_shift)
{
    __shared__ TO s_w[SCAN_THREADS / 32];
    n = scan_count(n, n_dev, n_shift);
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
__global__ void __launch_bounds__(1024) scan_partials_kernel(TO* partial, int64_t n, TO* total_out, const int64_t* n_dev, int n_shift, const TO* base_dev)
{
    // single CTA: exclusive scan of the chunk sums, in place (starting at *base_dev)
    const TO base = base_dev ? *base_dev : (TO)0;
    __shared__ TO s_w[32];
    __shared__ TO s_carry;
    n = scan_count(n, n_dev, n_shift);
    const int64_t n_chunks = (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
    if (threadIdx.x == 0) s_carry = base;
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
                                                                  int write_total, const int64_t* n_dev, int n_shift, const TO* base_dev)
{
    __shared__ u32 s_scan[SCAN_THREADS / 32 + 1];
    n = scan_count(n, n_dev, n_shift);
    if (n <= 0) {
        if (write_total && blockIdx.x == 0 && threadIdx.x == 0) out[0] = base_dev ? *base_dev : (TO)0;
        return;
    }
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
                                                           const double* __restrict__ coo_val, const u64* __restrict__ n_dev, u64 cap,
                                                           const u32* __restrict__ row_start, u32* rowfill, u32* cols,
                                                           double* vals)
{
    const u64 n = *n_dev < cap ? *n_dev : cap;  // record slots the parse reserved (clamped: the host checks the overflow)
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
constexpr int ROWSORT_MAXN = 128;
constexpr int ROWSORT_WARPS = 4;

struct RowSortArgs {
    int64_t row_begin;         // rows [row_begin, n_rows) are processed (one wave of query species)
    int64_t n_rows;
    const u32* row_start;      // [n_rows + 1]
    u32* rowlen;               // [n_rows] out: stored entries after dedup
    u32* cols;
    double* vals;
    const PairDesc* pairs;
    const int64_t* pair_row_base;  // [P + 1]
    int P;
    const double* lengths;
    u32* key;                  // out: Lf per slot of the wave (index: slot - *slot0; sentinel for unused slots)
    const u32* slot0;          // first CSR slot of the wave (device)
    double* keyval;            // (unused: the first sort pass reads the scores from vals, which sits at the same slots)
    u32 key_sentinel;
    u32* long_rows;            // out: rows with more than ROWSORT_MAXN records
    u32* n_long;
    u64* pair_nnz;             // [P]
    u32* scratch_cols;         // long-row scratch (same indexing as cols)
    double* scratch_vals;
    int short_max;             // rows of at most this many records were sorted by rowsort_short_kernel (0: none)
};

__device__ __forceinline__ bool rs_before(u32 cj, double vj, int j, u32 ck, double vk, int k)
{
    return (cj < ck) || (cj == ck && (vj > vk || (vj == vk && j < k)));
}

// Shared memory of one warp of rowsort_kernel
struct RowSortWarp {
    u32 c[ROWSORT_MAXN + 4];   // staged columns (padded for the 4-wide compares of the medium path)
    double v[ROWSORT_MAXN];    // staged scores, by input index
    u32 lf[ROWSORT_MAXN];      // Lq * Lh[col] per input index
    u32 oc[ROWSORT_MAXN];      // column per output slot
    u32 okey[ROWSORT_MAXN];    // Lf per output slot
    u64 ov[ROWSORT_MAXN];      // best score (bit pattern, > 0) per output slot
};

// Medium rows (64 < n <= ROWSORT_MAXN): O(n^2) rank sort out of shared memory, duplicates of a column share a slot.
__device__ __forceinline__ int rowsort_medium(const RowSortArgs& a, RowSortWarp& sw, u32 s0, int n, double Lq, const double* Lh)
{
    const int lane = (int)lane_id();
    for (int k = lane; k < n; k += 32) {
        const u32 col = a.cols[s0 + k];
        sw.c[k] = col;
        sw.v[k] = a.vals[s0 + k];
        sw.lf[k] = (u32)(Lq * Lh[col]);
        sw.ov[k] = 0ULL;
    }
    if (lane < 4) sw.c[n + lane] = 0xffffffffu;  // padding for the 4-wide compares (never < a real column)
    __syncwarp();
    const uint4* c4 = reinterpret_cast<const uint4*>(sw.c);
    const int n4 = (n + 3) >> 2;
    for (int k = lane; k < n; k += 32) {
        const u32 ck = sw.c[k];
        int rank = 0;  // number of strictly smaller columns: duplicates of a column share the slot
        for (int j4 = 0; j4 < n4; j4++) {
            const uint4 q = c4[j4];
            rank += (q.x < ck) + (q.y < ck) + (q.z < ck) + (q.w < ck);
        }
        sw.oc[rank] = ck;
        sw.okey[rank] = sw.lf[k];
        atomicMax(reinterpret_cast<unsigned long long*>(&sw.ov[rank]), (unsigned long long)__double_as_longlong(sw.v[k]));  // max over duplicate HSP lines
    }
    __syncwarp();
    // compact the used slots (ascending column) to the front
    int outn = 0;
    for (int k0 = 0; k0 < n; k0 += 32) {
        const int k = k0 + lane;
        const bool h = k < n && sw.ov[k] != 0ULL;
        const u32 bal = __ballot_sync(OFG_FULL, h);
        u32 col = 0, key = 0;
        u64 val = 0;
        if (h) { col = sw.oc[k]; key = sw.okey[k]; val = sw.ov[k]; }
        __syncwarp();
        if (h) {
            const int o = outn + __popc(bal & lanemask_lt());
            sw.oc[o] = col; sw.okey[o] = key; sw.ov[o] = val;
        }
        outn += __popc(bal);
        __syncwarp();
    }
    return outn;
}

// compare-exchange step of the 64-element bitonic network: element i = 32 h + lane, partner i ^ j
__device__ __forceinline__ u32 bitonic_cx(u32 mine, u32 other, bool keep_min) { return keep_min ? min(mine, other) : max(mine, other); }

// Per-row sort by column + max-dedup, one warp per row.  Rows of up to 64 records (practically all: a query has
// a few dozen hits per target species) stay in registers: two records per lane, prefetched one row ahead
// (row extents two rows ahead), sorted as packed (column, input index) words by a bitonic network of shuffles;
// duplicates of a column are adjacent afterwards and share an output slot (max score).  The length product
// Lq * Lh[col] that keys the later length sort is gathered right after the columns arrive, so no load
// depends on the sort result.
__global__ void __launch_bounds__(ROWSORT_WARPS * 32, 8) rowsort_kernel(RowSortArgs a)
{
    __shared__ __align__(16) RowSortWarp s_w[ROWSORT_WARPS];
    const int w = threadIdx.x >> 5, lane = (int)lane_id();
    RowSortWarp& sw = s_w[w];
    u32* const key_rel = a.key - *a.slot0;  // the key array is wave scratch, indexed from the wave's first slot
    // every CTA owns a contiguous run of rows (the pair changes rarely), its warps interleave inside it
    const int64_t rows_per_cta = (a.n_rows - a.row_begin + gridDim.x - 1) / gridDim.x;
    const int64_t row_begin = a.row_begin + (int64_t)blockIdx.x * rows_per_cta;
    const int64_t row_end = row_begin + rows_per_cta < a.n_rows ? row_begin + rows_per_cta : a.n_rows;
    int64_t row = row_begin + w;
    if (row >= row_end) return;
    // ---- prefetch state
    int pair = upper_bound_dev(a.pair_row_base, a.P + 1, row) - 1;   // pair of the row being prefetched
    int64_t pair_end = a.pair_row_base[pair + 1];
    int64_t p_li = 0, p_lj = 0, p_rb = 0;
    {
        const PairDesc& pd = a.pairs[pair];
        p_li = pd.li_off; p_lj = pd.lj_off; p_rb = pd.row_base;
    }
    auto advance_pair = [&](int64_t r) {
        if (r >= pair_end) {
            while (r >= a.pair_row_base[pair + 1]) pair++;
            pair_end = a.pair_row_base[pair + 1];
            const PairDesc& pd = a.pairs[pair];
            p_li = pd.li_off; p_lj = pd.lj_off; p_rb = pd.row_base;
        }
    };
    // extents of the row two steps ahead, records of the row one step ahead
    u32 nx_s0 = a.row_start[row];
    int nx_n = (int)(a.row_start[row + 1] - nx_s0);
    u32 nn_s0 = 0;
    int nn_n = 0;
    if (row + ROWSORT_WARPS < row_end) { nn_s0 = a.row_start[row + ROWSORT_WARPS]; nn_n = (int)(a.row_start[row + ROWSORT_WARPS + 1] - nn_s0); }
    u32 nx_c0 = 0, nx_c1 = 0;
    double nx_v0 = 0.0, nx_v1 = 0.0, nx_Lq;
    int64_t nx_lj;
    int nx_pair;
    auto fetch_records = [&]() {  // records + per-row constants of row `row` (extents nx_s0, nx_n) into the nx_* registers
        advance_pair(row);
        nx_pair = pair;
        nx_lj = p_lj;
        nx_Lq = a.lengths[p_li + (row - p_rb)];
        if (nx_n <= 64 && nx_n > a.short_max) {
            if (lane < nx_n) { nx_c0 = a.cols[nx_s0 + lane]; nx_v0 = a.vals[nx_s0 + lane]; }
            if (lane + 32 < nx_n) { nx_c1 = a.cols[nx_s0 + lane + 32]; nx_v1 = a.vals[nx_s0 + lane + 32]; }
        }
    };
    fetch_records();
    for (; row < row_end; row += ROWSORT_WARPS) {
        // ---- rotate the pipeline: current <- next, next <- next-next, issue the loads two rows ahead
        const u32 s0 = nx_s0;
        const int n = nx_n;
        const u32 c0 = nx_c0, c1 = nx_c1;
        const double v0 = nx_v0, v1 = nx_v1, Lq = nx_Lq;
        const double* Lh = a.lengths + nx_lj;
        const int cur_pair = nx_pair;
        const int64_t cur_row = row;
        nx_s0 = nn_s0; nx_n = nn_n;
        if (row + ROWSORT_WARPS < row_end) {
            const int64_t r2 = row + 2 * ROWSORT_WARPS;
            if (r2 < row_end) { nn_s0 = a.row_start[r2]; nn_n = (int)(a.row_start[r2 + 1] - nn_s0); }
            row += ROWSORT_WARPS;   // fetch_records works on `row`
            fetch_records();
            row -= ROWSORT_WARPS;
        }
        if (a.short_max > 0 && n <= a.short_max) continue;   // done by rowsort_short_kernel
        if (n == 0) {
            if (lane == 0) a.rowlen[cur_row] = 0;
            continue;
        }
        if (n > ROWSORT_MAXN) {
            if (lane == 0) a.long_rows[atomicAdd(a.n_long, 1u)] = (u32)cur_row;
            continue;
        }
        int outn;
        const bool v0ok = lane < n, v1ok = lane + 32 < n;
        if (n <= 64 && !__any_sync(OFG_FULL, (v0ok && c0 >= (1u << 26)) || (v1ok && c1 >= (1u << 26)))) {
            // ---- register path
            const double l0 = v0ok ? Lh[c0] : 0.0, l1 = v1ok ? Lh[c1] : 0.0;   // gathers in flight during the sort
            u32 k0 = v0ok ? ((c0 << 6) | (u32)lane) : 0xffffffffu;
            u32 k1 = v1ok ? ((c1 << 6) | (u32)(lane + 32)) : 0xffffffffu;
            sw.v[lane] = v0; sw.v[lane + 32] = v1;
            if (n <= 32) {
                // 32-element network on k0 only
#pragma unroll
                for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
                    for (int j = k >> 1; j > 0; j >>= 1) {
                        const u32 o = __shfl_xor_sync(OFG_FULL, k0, j);
                        const bool lower = (lane & j) == 0, asc = (lane & k) == 0 || k == 32;
                        k0 = bitonic_cx(k0, o, lower == asc);
                    }
                }
            } else {
#pragma unroll
                for (int k = 2; k <= 64; k <<= 1) {
#pragma unroll
                    for (int j = k >> 1; j > 0; j >>= 1) {
                        if (j == 32) {  // partner is the lane's other element; k == 64: ascending
                            const u32 lo = min(k0, k1), hi = max(k0, k1);
                            k0 = lo; k1 = hi;
                        } else {
                            const u32 o0 = __shfl_xor_sync(OFG_FULL, k0, j), o1 = __shfl_xor_sync(OFG_FULL, k1, j);
                            const bool lower = (lane & j) == 0;
                            // direction of element i = 32 h + lane: ascending iff (i & k) == 0
                            const bool asc0 = k == 64 || (k == 32 ? true : (lane & k) == 0);
                            const bool asc1 = k == 64 || (k == 32 ? false : (lane & k) == 0);
                            k0 = bitonic_cx(k0, o0, lower == asc0);
                            k1 = bitonic_cx(k1, o1, lower == asc1);
                        }
                    }
                }
            }
            sw.lf[lane] = (u32)(Lq * l0); sw.lf[lane + 32] = (u32)(Lq * l1);
            __syncwarp();
            // sorted position i = 32 h + lane holds k_h; duplicates of a column are adjacent
            const u32 p0 = __shfl_up_sync(OFG_FULL, k0, 1), p1u = __shfl_up_sync(OFG_FULL, k1, 1), last0 = __shfl_sync(OFG_FULL, k0, 31);
            const u32 p1 = lane == 0 ? last0 : p1u;
            const bool in0 = k0 != 0xffffffffu, in1 = k1 != 0xffffffffu;
            const bool h0 = in0 && (lane == 0 || (p0 >> 6) != (k0 >> 6));
            const bool h1 = in1 && ((p1 >> 6) != (k1 >> 6));
            const u32 hb0 = __ballot_sync(OFG_FULL, h0), hb1 = __ballot_sync(OFG_FULL, h1);
            const u32 le = lanemask_lt() | (1u << lane);
            const int n_h0 = __popc(hb0);
            const int slot0 = __popc(hb0 & le) - 1, slot1 = n_h0 + __popc(hb1 & le) - 1;
            // the first record of every column run opens its slot ...
            if (h0) { const int idx = (int)(k0 & 63u); sw.oc[slot0] = k0 >> 6; sw.okey[slot0] = sw.lf[idx]; sw.ov[slot0] = (u64)__double_as_longlong(sw.v[idx]); }
            if (h1) { const int idx = (int)(k1 & 63u); sw.oc[slot1] = k1 >> 6; sw.okey[slot1] = sw.lf[idx]; sw.ov[slot1] = (u64)__double_as_longlong(sw.v[idx]); }
            // ... and duplicate HSP lines of a column (rare) raise it to their maximum
            if (__any_sync(OFG_FULL, (in0 && !h0) || (in1 && !h1))) {
                __syncwarp();
                if (in0 && !h0) atomicMax(reinterpret_cast<unsigned long long*>(&sw.ov[slot0]), (unsigned long long)__double_as_longlong(sw.v[k0 & 63u]));
                if (in1 && !h1) atomicMax(reinterpret_cast<unsigned long long*>(&sw.ov[slot1]), (unsigned long long)__double_as_longlong(sw.v[k1 & 63u]));
            }
            outn = n_h0 + __popc(hb1);
            __syncwarp();
        } else {
            outn = rowsort_medium(a, sw, s0, n, Lq, Lh);
        }
        for (int k = lane; k < outn; k += 32) {
            const u32 pos = s0 + (u32)k;
            const double val = __longlong_as_double((long long)sw.ov[k]);
            a.cols[pos] = sw.oc[k];
            a.vals[pos] = val;
            key_rel[pos] = sw.okey[k];
        }
        for (int k = outn + lane; k < n; k += 32) key_rel[s0 + k] = a.key_sentinel;
        if (lane == 0) {
            a.rowlen[cur_row] = (u32)outn;
            atomicAdd(reinterpret_cast<unsigned long long*>(&a.pair_nnz[cur_pair]), (unsigned long long)outn);
        }
        __syncwarp();
    }
}

// Short rows (at most ROWSORT_SHORT = 16 records: every row of a job with a handful of hits per query and target species,
// BASELINE C4) - half a warp per row, two rows per warp, everything in registers: one record per lane, a 16-element bitonic
// network on packed (column << 4 | input index) words, the scores and length products fetched from the lane that read them.
// A warp-per-row kernel leaves 3/4 of its lanes idle on such rows (C4: 8.8 records per row).  Launched before rowsort_kernel,
// which then skips these rows (RowSortArgs.short_max).
constexpr int ROWSORT_SHORT = 16;
__global__ void __launch_bounds__(256) rowsort_short_kernel(RowSortArgs a)
{
    const int lane = (int)lane_id(), hl = lane & 15, hbase = lane & 16;
    u32* const key_rel = a.key - *a.slot0;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nr = a.n_rows - a.row_begin;
    int64_t per_warp = (nr + n_warps - 1) / n_warps;
    per_warp += per_warp & 1;   // even: the two halves of a warp alternate rows of one contiguous run
    const int64_t r0 = a.row_begin + wid * per_warp;
    const int64_t r1 = r0 + per_warp < a.n_rows ? r0 + per_warp : a.n_rows;
    if (r0 >= r1) return;
    int pair = upper_bound_dev(a.pair_row_base, a.P + 1, r0) - 1;
    int64_t pair_end = a.pair_row_base[pair + 1];
    int64_t p_li = a.pairs[pair].li_off, p_lj = a.pairs[pair].lj_off, p_rb = a.pairs[pair].row_base;
    unsigned long long acc_nnz = 0;   // stored entries of pair acc_pair not yet added to pair_nnz (first lane of each half)
    int acc_pair = pair;
    // software pipeline: extents of the row two steps ahead and records of the row one step ahead are in flight while a row is sorted
    const int half = hbase >> 4;
    auto extent = [&](int64_t rb, u32* s0, int* n) {
        const int64_t r = rb + half;
        *s0 = 0; *n = ROWSORT_SHORT + 1;
        if (rb < r1 && r < r1) { *s0 = a.row_start[r]; *n = (int)(a.row_start[r + 1] - *s0); }
    };
    u32 nx_k = 0xffffffffu;
    double nx_v = 0.0, nx_Lq = 0.0;
    int64_t nx_lj = p_lj;
    int nx_pair = pair;
    auto fetch = [&](int64_t rb, u32 s0, int n) {   // records of this half's row of step rb (extents s0, n)
        const int64_t r = rb + half;
        nx_k = 0xffffffffu; nx_v = 0.0;
        if (rb < r1 && r < r1) {
            if (r >= pair_end) {
                while (r >= a.pair_row_base[pair + 1]) pair++;
                pair_end = a.pair_row_base[pair + 1];
                p_li = a.pairs[pair].li_off; p_lj = a.pairs[pair].lj_off; p_rb = a.pairs[pair].row_base;
            }
            if (n <= ROWSORT_SHORT && hl < n) {
                nx_k = (a.cols[s0 + hl] << 4) | (u32)hl;
                nx_v = a.vals[s0 + hl];
                nx_Lq = a.lengths[p_li + (r - p_rb)];
            }
        }
        nx_pair = pair;
        nx_lj = p_lj;
    };
    u32 e1_s0, e2_s0;
    int e1_n, e2_n;
    extent(r0, &e1_s0, &e1_n);
    extent(r0 + 2, &e2_s0, &e2_n);
    fetch(r0, e1_s0, e1_n);
    for (int64_t rb = r0; rb < r1; rb += 2) {
        const int64_t row = rb + half;
        const bool valid = row < r1;
        const u32 s0 = e1_s0;
        const int n = e1_n;
        u32 k = nx_k;
        const double v = nx_v;
        const int cur_pair = nx_pair;
        // the length of the hit sequence is gathered while the next row is fetched and this one is sorted
        const double lprod = k != 0xffffffffu ? nx_Lq * a.lengths[nx_lj + (k >> 4)] : 0.0;
        e1_s0 = e2_s0; e1_n = e2_n;
        extent(rb + 4, &e2_s0, &e2_n);
        fetch(rb + 2, e1_s0, e1_n);
        const bool todo = valid && n <= ROWSORT_SHORT;
        if (!__any_sync(OFG_FULL, todo)) continue;
        if (todo && cur_pair != acc_pair) {
            if (hl == 0 && acc_nnz) atomicAdd(reinterpret_cast<unsigned long long*>(&a.pair_nnz[acc_pair]), acc_nnz);
            acc_nnz = 0;
            acc_pair = cur_pair;
        }
#pragma unroll
        for (int kk = 2; kk <= 16; kk <<= 1) {
#pragma unroll
            for (int j = kk >> 1; j > 0; j >>= 1) {
                const u32 o = __shfl_xor_sync(OFG_FULL, k, j);
                const bool lower = (hl & j) == 0, asc = (hl & kk) == 0 || kk == 16;
                k = (lower == asc) ? min(k, o) : max(k, o);
            }
        }
        // sorted position hl holds k; the score and length product sit in the lane that loaded record (k & 15)
        const int src = hbase | (int)(k & 15u);
        double vs = __shfl_sync(OFG_FULL, v, src);
        const u32 lfs = __shfl_sync(OFG_FULL, (u32)lprod, src);
        const u32 prev = __shfl_up_sync(OFG_FULL, k, 1);
        const bool in = k != 0xffffffffu;
        const bool head = in && (hl == 0 || (prev >> 4) != (k >> 4));
        const u32 hb = (__ballot_sync(OFG_FULL, head) >> hbase) & 0xffffu;
        if (__any_sync(OFG_FULL, in && !head)) {
            // duplicate HSP lines of a column (rare): the first record of the run takes the maximum of the run
#pragma unroll
            for (int d = 1; d < 16; d <<= 1) {
                const double ov = __shfl_down_sync(OFG_FULL, vs, d);
                const u32 ok = __shfl_down_sync(OFG_FULL, k, d);
                if (hl + d < 16 && ok != 0xffffffffu && (ok >> 4) == (k >> 4) && ov > vs) vs = ov;
            }
        }
        if (todo) {
            const int outn = __popc(hb);
            if (head) {
                const u32 pos = s0 + (u32)(__popc(hb & ((2u << hl) - 1u)) - 1);
                a.cols[pos] = k >> 4;
                a.vals[pos] = vs;
                key_rel[pos] = lfs;
            }
            if (hl >= outn && hl < n) key_rel[s0 + hl] = a.key_sentinel;
            if (hl == 0) { a.rowlen[row] = (u32)outn; acc_nnz += (unsigned long long)outn; }
        }
    }
    if (hl == 0 && acc_nnz) atomicAdd(reinterpret_cast<unsigned long long*>(&a.pair_nnz[acc_pair]), acc_nnz);
}

// Rows longer than ROWSORT_MAXN: one CTA per row, O(n^2) rank sort through global scratch.  Rare
// (a query with hundreds of targets in one species); correctness path, not a throughput path.
__global__ void __launch_bounds__(256) rowsort_long_kernel(RowSortArgs a)
{
    __shared__ u32 s_scan[256 / 32 + 1];
    __shared__ u32 s_outn;
    const u32 n_long = *a.n_long;
    u32* const key_rel = a.key - *a.slot0;
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
                key_rel[pos] = (u32)(Lq * Lh[col]);
                }
            __syncthreads();
            if (threadIdx.x == 0) s_outn = base + tot;
            __syncthreads();
        }
        const int outn = (int)s_outn;
        for (int k = outn + threadIdx.x; k < n; k += blockDim.x) key_rel[s0 + k] = a.key_sentinel;
        if (threadIdx.x == 0) {
            a.rowlen[row] = (u32)outn;
            atomicAdd(&a.pair_nnz[pair], (u64)outn);
        }
        __syncthreads();
    }
}
