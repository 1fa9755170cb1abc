// Stage NORM, parts 1-2: the per-pair length binning and top-5% selection of
// scnorm.GetTopPercentileOfScores (gathering.py:93-116).
//
//  radix sort : STABLE segmented LSD radix sort of (Lf, score) records by Lf inside each pair.  The
//               input order is the matrix order (row, col), so stability reproduces the reference's
//               sorted(zip(L, range(n))) tie-break by original index (:96) without carrying an index.
//               Unused slots carry a key one bit above the largest Lf and sink to the end of the pair.
//  bin_cut    : one warp per length bin: 95th percentile with numpy's linear interpolation via a radix select on
//               the fp64 bit patterns (starting at the first bit where the bin's min and max differ, stopping
//               once the k-th value is alone in its bucket); the hits >= cut-off are compacted in order, their
//               two np.log10 values (gathering.py:82,120; bit-exact SVML emulation, ofg_log10.cuh) are taken
//               on full lanes and staged at the bin's own slots.
//  bin_select : concatenates the staged selections of all bins (exclusive scan of the counts).
#pragma once
#include "ofg_common.cuh"
#include "ofg_log10.cuh"

// ---------------------------------------------------------------- segmented stable radix sort
constexpr int RS_THREADS = 512;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 4096 records
constexpr int RS_MAX_BITS = 9;
constexpr int RS_MAX_DIGITS = 1 << RS_MAX_BITS;
static_assert(RS_MAX_DIGITS == RS_THREADS, "one digit per thread in the offset phase");

struct RadixArgs {
    const u32* key_in;
    const double* val_in;
    u32* key_out;
    double* val_out;
    const int64_t* n_tiles_dev;    // sort tiles of this wave (device-side: follows from the row scan)
    int64_t val_in_off_dev_valid;  // != 0: val_in is the CSR value array (absolute slots): add *slot0 to the wave-relative slot
    const u32* slot0;              // first CSR slot of the wave (device)
    int shift, bits;               // digit = (key >> shift) & ((1 << bits) - 1)
    u32* counters;                 // [n_tiles << bits], layout [pair][digit][tile in pair]
    const struct RsTileDesc* desc; // [n_tiles]
};

// Sort / bin layout of one wave of species pairs, derived on the device from the row scan (no host round trip):
// the records of pair i occupy the CSR slots [row_start[first row of i], row_start[first row of i + 1]).
struct WaveLayoutArgs {
    const u32* row_start;          // absolute
    const int64_t* pair_row_base;  // [P_all + 1] absolute
    int P0, P1;                    // pairs of the wave
    const u64* pair_nnz;           // [P_all] stored entries after dedup (rowsort)
    int64_t* pair_slot;            // out [Pw + 1] wave-relative first slot of each pair
    int64_t* tile_prefix;          // out [Pw + 1] sort tiles before each pair
    u64* task_prefix;              // out [Pw + 1] length bins before each pair
    int64_t* n_tiles;              // out
    int64_t* n_tasks;              // out
    int64_t max_tiles, max_tasks;  // capacity of the per-tile / per-bin arrays
    u32* err_flags;                // bit 4: capacity exceeded
};

// bins of a pair with H stored hits (gathering.py:100-107): all hits when H < 100, else H // binSize bins
__host__ __device__ __forceinline__ u64 ofg_bins_of(u64 H, u32* n_in)
{
    if (H == 0) { *n_in = 0; return 0; }
    if (H < 100) { *n_in = (u32)H; return 1; }
    *n_in = H > 5000 ? 1000u : (H > 1000 ? 200u : 20u);
    return H / *n_in;
}

__global__ void __launch_bounds__(1024) wave_layout_kernel(WaveLayoutArgs a, int with_sort)
{
    // single CTA: two exclusive scans over the pairs of the wave
    __shared__ u64 s_w[2][32];
    __shared__ u64 s_carry[2];
    const int Pw = a.P1 - a.P0;
    const u32 s0 = a.row_start[a.pair_row_base[a.P0]];
    if (threadIdx.x == 0) { s_carry[0] = 0; s_carry[1] = 0; }
    __syncthreads();
    for (int b = 0; b < Pw; b += 1024) {
        const int i = b + (int)threadIdx.x;
        u64 v0 = 0, v1 = 0;
        if (i < Pw) {
            const u32 sa = a.row_start[a.pair_row_base[a.P0 + i]], sb = a.row_start[a.pair_row_base[a.P0 + i + 1]];
            a.pair_slot[i] = (int64_t)(sa - s0);
            if (i == Pw - 1) a.pair_slot[Pw] = (int64_t)(sb - s0);
            v0 = with_sort ? ((u64)(sb - sa) + RS_TILE - 1) / RS_TILE : 0;
            u32 n_in;
            v1 = with_sort ? ofg_bins_of(a.pair_nnz[a.P0 + i], &n_in) : 0;
        }
        u64 i0 = v0, i1 = v1;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u64 t0 = __shfl_up_sync(OFG_FULL, i0, d), t1 = __shfl_up_sync(OFG_FULL, i1, d);
            if ((int)lane_id() >= d) { i0 += t0; i1 += t1; }
        }
        if (lane_id() == 31) { s_w[0][threadIdx.x >> 5] = i0; s_w[1][threadIdx.x >> 5] = i1; }
        __syncthreads();
        if (threadIdx.x < 32) {
            const u64 a0 = s_w[0][threadIdx.x], a1 = s_w[1][threadIdx.x];
            u64 c0 = a0, c1 = a1;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const u64 t0 = __shfl_up_sync(OFG_FULL, c0, d), t1 = __shfl_up_sync(OFG_FULL, c1, d);
                if ((int)lane_id() >= d) { c0 += t0; c1 += t1; }
            }
            s_w[0][threadIdx.x] = c0 - a0;
            s_w[1][threadIdx.x] = c1 - a1;
        }
        __syncthreads();
        const u64 e0 = s_carry[0] + s_w[0][threadIdx.x >> 5] + i0 - v0, e1 = s_carry[1] + s_w[1][threadIdx.x >> 5] + i1 - v1;
        if (i < Pw) { a.tile_prefix[i] = (int64_t)e0; a.task_prefix[i] = e1; }
        __syncthreads();
        if (threadIdx.x == 1023) { s_carry[0] = e0 + v0; s_carry[1] = e1 + v1; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        a.tile_prefix[Pw] = (int64_t)s_carry[0];
        a.task_prefix[Pw] = s_carry[1];
        int64_t nt = (int64_t)s_carry[0], nk = (int64_t)s_carry[1];
        if (nt > a.max_tiles || nk > a.max_tasks) { atomicOr(a.err_flags, 16u); nt = 0; nk = 0; }
        *a.n_tiles = nt;
        *a.n_tasks = nk;
    }
}

// Per-tile constants, computed once per sort by rs_desc_kernel so that the per-tile path of the three passes
// has no dependent global loads: every thread prefetches the next tile's descriptor while it works.
struct __align__(16) RsTileDesc {
    int64_t e0;     // first slot of the tile
    int64_t cidx0;  // counter index of (digit 0, this tile);  + digit * ntp
    int cnt;        // records in the tile
    int ntp;        // tiles in the tile's pair
    int pair;       // wave-relative pair of the tile
    int pad;
};

__global__ void rs_desc_kernel(const int64_t* tile_prefix, const int64_t* pair_slot, int Pw, const int64_t* n_tiles_dev, int bits, RsTileDesc* desc)
{
    const int64_t n_tiles = *n_tiles_dev;
    for (int64_t tile = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; tile < n_tiles; tile += (int64_t)gridDim.x * blockDim.x) {
        const int pair = upper_bound_dev(tile_prefix, Pw + 1, tile) - 1;
        const int64_t tp0 = tile_prefix[pair], tp1 = tile_prefix[pair + 1];
        const int64_t t = tile - tp0;
        RsTileDesc d;
        d.e0 = pair_slot[pair] + t * RS_TILE;
        const int64_t rem = pair_slot[pair + 1] - d.e0;
        d.cnt = (int)(rem < RS_TILE ? rem : RS_TILE);
        d.cidx0 = (tp0 << bits) + t;
        d.ntp = (int)(tp1 - tp0);
        d.pair = pair;
        d.pad = 0;
        desc[tile] = d;
    }
}

__device__ __forceinline__ RsTileDesc rs_load_desc(const RsTileDesc* desc, int64_t tile, int64_t n_tiles)
{
    RsTileDesc d;
    d.e0 = 0; d.cidx0 = 0; d.cnt = 0; d.ntp = 0; d.pair = 0; d.pad = 0;
    if (tile < n_tiles) d = desc[tile];
    return d;
}

__global__ void __launch_bounds__(RS_THREADS) radix_hist_kernel(RadixArgs a)
{
    __shared__ u32 s_hist[RS_MAX_DIGITS];
    const int nd = 1 << a.bits;
    const u32 mask = (u32)nd - 1u;
    const int64_t n_tiles = *a.n_tiles_dev;
    RsTileDesc next = rs_load_desc(a.desc, blockIdx.x, n_tiles);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const RsTileDesc td = next;
        next = rs_load_desc(a.desc, tile + gridDim.x, n_tiles);
        const int cnt = td.cnt;
        const int64_t e0 = td.e0, cidx0 = td.cidx0, ntp = td.ntp;
        s_hist[threadIdx.x] = 0;
        __syncthreads();
        for (int k = threadIdx.x; k < cnt; k += RS_THREADS) atomicAdd(&s_hist[(a.key_in[e0 + k] >> a.shift) & mask], 1u);
        __syncthreads();
        if ((int)threadIdx.x < nd) a.counters[cidx0 + (int64_t)threadIdx.x * ntp] = s_hist[threadIdx.x];
        __syncthreads();
    }
}

// Dynamic shared memory of radix_scatter_kernel
struct RadixSmem {
    double val[RS_TILE];                    // 32 KB  records in tile-sorted order
    u32 key[RS_TILE];                       // 16 KB
    unsigned short cnt[RS_WARPS][RS_MAX_DIGITS];  // 16 KB per-warp digit counts -> per-warp offsets inside a digit run
    u32 gbase[RS_MAX_DIGITS];               // global destination of the first record of each digit run
    u32 tbase[RS_MAX_DIGITS];               // tile-local start of each digit run
    u32 scan[RS_WARPS + 1];
};

// counters must hold the exclusive scan (global destination of the first record of each (pair, digit, tile)
// group) when this runs.  Stable: ranks follow (warp, iteration, lane) == input order.  Records are first
// permuted into tile-sorted order in shared memory so that every digit run leaves as one coalesced write.
__global__ void __launch_bounds__(RS_THREADS, 2) radix_scatter_kernel(RadixArgs a)
{
    extern __shared__ __align__(16) unsigned char rs_smem_raw[];
    RadixSmem& sm = *reinterpret_cast<RadixSmem*>(rs_smem_raw);
    const int nd = 1 << a.bits;
    const u32 mask = (u32)nd - 1u;
    const int w = threadIdx.x >> 5, lane = lane_id();
    const int64_t n_tiles = *a.n_tiles_dev;
    const int64_t voff = a.val_in_off_dev_valid ? (int64_t)*a.slot0 : 0;  // first pass: scores come from the CSR values
    RsTileDesc next = rs_load_desc(a.desc, blockIdx.x, n_tiles);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const RsTileDesc td = next;
        next = rs_load_desc(a.desc, tile + gridDim.x, n_tiles);
        const int cnt = td.cnt;
        const int64_t e0 = td.e0, cidx0 = td.cidx0, ntp = td.ntp;
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ww++) sm.cnt[ww][threadIdx.x] = 0;
        sm.gbase[threadIdx.x] = (int)threadIdx.x < nd ? a.counters[cidx0 + (int64_t)threadIdx.x * ntp] : 0u;
        u32 key[RS_ITEMS];
        double val[RS_ITEMS];
        u32 rk[RS_ITEMS];  // peer mask, then the record's rank among its warp's records of the same digit
        const int kbase = w * (RS_ITEMS * 32) + lane;  // warp w owns records [w*256, w*256+256) of the tile
#pragma unroll
        for (int it = 0; it < RS_ITEMS; it++) {
            const int k = kbase + it * 32;
            key[it] = k < cnt ? a.key_in[e0 + k] : 0u;
            val[it] = k < cnt ? a.val_in[voff + e0 + k] : 0.0;
        }
        // all match operations first: they are independent and their latency overlaps
#pragma unroll
        for (int it = 0; it < RS_ITEMS; it++) {
            const int k = kbase + it * 32;
            const u32 d = k < cnt ? ((key[it] >> a.shift) & mask) : 0xffffffffu;
            rk[it] = __match_any_sync(OFG_FULL, d);
        }
        __syncthreads();
#pragma unroll
        for (int it = 0; it < RS_ITEMS; it++) {
            const int k = kbase + it * 32;
            const bool valid = k < cnt;
            const u32 d = (key[it] >> a.shift) & mask;
            const u32 peers = rk[it];
            const int leader = __ffs(peers) - 1;
            u32 old = 0;
            if (valid && lane == leader) {
                old = sm.cnt[w][d];
                sm.cnt[w][d] = (unsigned short)(old + (u32)__popc(peers));
            }
            old = __shfl_sync(OFG_FULL, old, leader);
            rk[it] = old + (u32)__popc(peers & lanemask_lt());
            __syncwarp();
        }
        __syncthreads();
        // thread d: exclusive prefix over warps inside digit d, then tile-local start of the digit run
        u32 total = 0;
        {
            const int d = threadIdx.x;
#pragma unroll
            for (int ww = 0; ww < RS_WARPS; ww++) {
                const u32 c = sm.cnt[ww][d];
                sm.cnt[ww][d] = (unsigned short)total;
                total += c;
            }
        }
        u32 tile_total;
        const u32 tb = block_excl_scan<RS_THREADS>(total, sm.scan, &tile_total);
        sm.tbase[threadIdx.x] = tb;
        __syncthreads();
#pragma unroll
        for (int it = 0; it < RS_ITEMS; it++) {
            const int k = kbase + it * 32;
            if (k < cnt) {
                const u32 d = (key[it] >> a.shift) & mask;
                const u32 tpos = sm.tbase[d] + sm.cnt[w][d] + rk[it];
                sm.key[tpos] = key[it];
                sm.val[tpos] = val[it];
            }
        }
        __syncthreads();
#pragma unroll
        for (int it = 0; it < RS_ITEMS; it++) {
            const int i = it * RS_THREADS + threadIdx.x;
            if (i < cnt) {
                const u32 kk = sm.key[i];
                const u32 d = (kk >> a.shift) & mask;
                const u32 dst = sm.gbase[d] + ((u32)i - sm.tbase[d]);
                a.key_out[dst] = kk;
                a.val_out[dst] = sm.val[i];
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------- bins: percentile cut-off + selection
constexpr int BIN_THREADS = 128;
constexpr int BIN_MAXN = 1024;

struct BinTask {      // host-built, one per length bin (or one per pair with < 100 hits: "take all")
    u32 first;        // first sorted slot of the bin
    u32 n;            // records in the bin
    int32_t pair;
    int32_t take_all; // nScores < 100 (gathering.py:100-102)
};

// Expands the per-pair bin counts (task_prefix of wave_layout_kernel) into one BinTask per length bin.
__global__ void bin_tasks_kernel(const u64* task_prefix, const int64_t* pair_slot, const u64* pair_nnz, int P0, int Pw, const int64_t* n_tasks_dev,
                                 BinTask* tasks)
{
    const int64_t n_tasks = *n_tasks_dev;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_tasks; t += (int64_t)gridDim.x * blockDim.x) {
        const int pair = upper_bound_dev(task_prefix, Pw + 1, (u64)t) - 1;
        u32 n_in;
        ofg_bins_of(pair_nnz[P0 + pair], &n_in);
        BinTask bt;
        bt.first = (u32)(pair_slot[pair] + (int64_t)(t - (int64_t)task_prefix[pair]) * n_in);
        bt.n = n_in;
        bt.pair = pair;            // wave-relative
        bt.take_all = pair_nnz[P0 + pair] < 100 ? 1 : 0;
        tasks[t] = bt;
    }
}

struct BinArgs {
    const BinTask* tasks;
    const int64_t* n_tasks_dev;
    const u32* key;        // sorted Lf
    const double* val;     // scores in the same order
    // numpy percentile(…, 95) 'linear': k = floor((n-1)*0.95), g = (n-1)*0.95 - k, for n = 20, 200, 1000
    int k20, k200, k1000;
    double g20, g200, g1000;
    double* cut;           // [n_tasks]
    u32* cnt;              // [n_tasks] selected per bin
    const u64* sel_off;    // [n_tasks + 1] exclusive scan of cnt (bin_select only)
    double* lx;            // out: log10(Lf) of selected hits, reference order
    double* ly;            // out: log10(score)
    u64 sel_cap;
    double* stage_lx;      // [n_slots] per-bin staging of the selected hits (bin_cut -> bin_select)
    double* stage_ly;
    const uint4* rec;      // UNORDERED only: candidate hits of every bin (slot, Lf, score lo, score hi) at the bin's slots
    const u32* cand_cnt;   // UNORDERED only: [n_tasks] candidates per bin
    const unsigned short* below;  // UNORDERED only: [n_tasks] hits of the bin below the candidates (shifts the order statistic)
};

// One WARP per length bin (no CTA barriers).  k-th smallest (0-based) of n positive doubles held in this warp's
// shared-memory slice as bit patterns (unsigned order == numeric order), by a radix select over 8-bit digits with
// a warp-private histogram.  The select starts at the highest byte in which the bin's minimum and maximum
// differ (the bytes above are common to all values) and stops as soon as the k-th value is alone in its
// bucket, which for bit scores happens after two or three digits.  Also returns how many values are <= it.
__device__ __forceinline__ u64 warp_kth(const u64* s_bits, int n, int k, u64 vmin, u64 vmax, u32* s_hist, int* n_le)
{
    const int lane = lane_id();
    if (vmin == vmax) {
        *n_le = n;
        return vmin;
    }
    // digits of up to 8 bits, the first one starting at the highest bit in which minimum and maximum differ
    int hi = 64 - __clzll((long long)(vmin ^ vmax));   // bits [hi, 64) are common to all values
    u64 prefix = hi >= 64 ? 0ULL : (vmin >> hi);
    int kk = k;
    while (hi > 0) {
        const int width = hi < 8 ? hi : 8;
        const int shift = hi - width;
        const u32 dmask = (1u << width) - 1u;
#pragma unroll
        for (int t = 0; t < 8; t++) s_hist[lane * 8 + t] = 0;
        __syncwarp();
        for (int i = lane; i < n; i += 32) {
            const u64 b = s_bits[i];
            const bool match = hi >= 64 || (b >> hi) == prefix;
            if (match) atomicAdd(&s_hist[(u32)(b >> shift) & dmask], 1u);
        }
        __syncwarp();
        // lane l owns buckets [8l, 8l+8)
        u32 h[8], sum = 0;
#pragma unroll
        for (int t = 0; t < 8; t++) { h[t] = s_hist[lane * 8 + t]; sum += h[t]; }
        const u32 inc = warp_incl_scan(sum);
        const u32 exc = inc - sum;
        const bool mine = (u32)kk >= exc && (u32)kk < inc;
        u32 bsel = 0, below = exc, hsel = h[0];
        if (mine) {
            u32 run = exc;
#pragma unroll
            for (int t = 0; t < 7; t++) {
                if ((u32)kk >= run + h[t]) { below = run + h[t]; bsel = t + 1; hsel = h[t + 1]; }
                run += h[t];
            }
        }
        const u32 src = __ffs(__ballot_sync(OFG_FULL, mine)) - 1;
        const u32 bucket = __shfl_sync(OFG_FULL, (u32)(lane * 8) + bsel, src);
        const u32 blw = __shfl_sync(OFG_FULL, below, src);
        const u32 alone = __shfl_sync(OFG_FULL, hsel, src);
        prefix = (prefix << width) | (u64)bucket;
        kk -= (int)blw;
        hi = shift;
        __syncwarp();
        if (alone == 1u && hi > 0) {  // the k-th value is the only one with this prefix: fetch it
            u64 cand = 0;
            for (int i = lane; i < n; i += 32) {
                const u64 b = s_bits[i];
                if ((b >> hi) == prefix) cand = b;
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                const u64 o = __shfl_xor_sync(OFG_FULL, cand, d);
                cand = o > cand ? o : cand;
            }
            *n_le = k + 1;   // k values are smaller, this one is unique
            return cand;
        }
    }
    u32 c = 0;
    for (int i = lane; i < n; i += 32) c += (s_bits[i] <= prefix);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(OFG_FULL, c, d);
    *n_le = (int)c;
    return prefix;
}

constexpr int BIN_WARPS = 4;
static_assert(BIN_WARPS * 32 == BIN_THREADS, "one warp per bin");

// Per bin: the 95th-percentile cut-off (numpy 'linear'), the number of hits >= cut-off, and those hits themselves
// - log10(Lf), log10(score) in sorted order - staged at the bin's own slot range (stage_lx / stage_ly [first ..]).
// The selected hits are first compacted in shared memory so that the two log10 evaluations run on full lanes.
// UNORDERED (the bucket path of ofg_bins.cuh): the bin's hits arrive in arbitrary order; the selected ones are put into
// the reference's order by ranking their (Lf, original index) words among each other.
template <bool UNORDERED>
__global__ void __launch_bounds__(BIN_THREADS) bin_cut_kernel(BinArgs a)
{
    __shared__ u64 s_bits_all[BIN_WARPS][BIN_MAXN];
    __shared__ u32 s_hist_all[BIN_WARPS][256];
    __shared__ uint16_t s_idx_all[BIN_WARPS][BIN_MAXN];
    const int w = threadIdx.x >> 5, lane = lane_id();
    u64* s_bits = s_bits_all[w];
    u32* s_hist = s_hist_all[w];
    uint16_t* s_idx = s_idx_all[w];
    const int64_t n_warps = (int64_t)gridDim.x * BIN_WARPS;
    const int64_t n_tasks = *a.n_tasks_dev;
    for (int64_t t = (int64_t)blockIdx.x * BIN_WARPS + w; t < n_tasks; t += n_warps) {
        const BinTask bt = a.tasks[t];
        const int n_bin = (int)bt.n;                                    // hits of the bin (fixes numpy's virtual index)
        const int n = UNORDERED ? (int)a.cand_cnt[t] : n_bin;           // hits looked at: all of them, or the top candidates
        const int n_below = UNORDERED ? (int)a.below[t] : 0;
        u64 vmin = ~0ULL, vmax = 0ULL;
        for (int i0 = 0; i0 < n; i0 += 256) {  // eight independent loads in flight per lane
            u64 tb[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int i = i0 + u * 32 + lane;
                if (UNORDERED) {
                    uint4 r = make_uint4(0u, 0u, 0u, 0u);
                    if (i < n) r = a.rec[bt.first + i];
                    tb[u] = (u64)r.z | ((u64)r.w << 32);
                } else {
                    tb[u] = i < n ? (u64)__double_as_longlong(a.val[bt.first + i]) : 0ULL;
                }
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int i = i0 + u * 32 + lane;
                if (i < n) {
                    s_bits[i] = tb[u];
                    vmin = tb[u] < vmin ? tb[u] : vmin;
                    vmax = tb[u] > vmax ? tb[u] : vmax;
                }
            }
        }
        double cut = 0.0;
        if (!bt.take_all) {
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                const u64 o1 = __shfl_xor_sync(OFG_FULL, vmin, d), o2 = __shfl_xor_sync(OFG_FULL, vmax, d);
                vmin = o1 < vmin ? o1 : vmin;
                vmax = o2 > vmax ? o2 : vmax;
            }
            __syncwarp();
            const int k = (n_bin == 1000 ? a.k1000 : (n_bin == 200 ? a.k200 : a.k20)) - n_below;   // a[k] and a[k + 1] are candidates
            const double g = n_bin == 1000 ? a.g1000 : (n_bin == 200 ? a.g200 : a.g20);
            int n_le;
            const u64 ak = warp_kth(s_bits, n, k, vmin, vmax, s_hist, &n_le);
            u64 ak1 = ak;
            if (n_le <= k + 1) {  // a[k+1] is the smallest value above a[k]
                u64 m = ~0ULL;
                for (int i = lane; i < n; i += 32) {
                    const u64 b = s_bits[i];
                    if (b > ak && b < m) m = b;
                }
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    const u64 o = __shfl_xor_sync(OFG_FULL, m, d);
                    m = o < m ? o : m;
                }
                ak1 = (m == ~0ULL) ? ak : m;
            }
            const double va = __longlong_as_double((long long)ak), vb = __longlong_as_double((long long)ak1);
            // numpy _lerp for gamma < 0.5: a + (b - a) * gamma, separately rounded
            cut = __dadd_rn(va, __dmul_rn(__dsub_rn(vb, va), g));
        }
        __syncwarp();
        // order-preserving compaction of the selected hits, in place (the write position never passes the read position)
        int cnt = 0;
        for (int i0 = 0; i0 < n; i0 += 32) {
            const int i = i0 + lane;
            const u64 b = i < n ? s_bits[i] : 0ULL;
            const bool sel = i < n && (bt.take_all || __longlong_as_double((long long)b) >= cut);
            const u32 bal = __ballot_sync(OFG_FULL, sel);
            __syncwarp();
            if (sel) {
                const int o = cnt + __popc(bal & lanemask_lt());
                s_bits[o] = b;
                s_idx[o] = (uint16_t)i;
            }
            cnt += __popc(bal);
            __syncwarp();
        }
        if (lane == 0) { a.cut[t] = cut; a.cnt[t] = (u32)cnt; }
        if (UNORDERED) {
            constexpr int SC = 128;                       // words kept in shared memory (the select histogram is free now)
            u64* s_word = reinterpret_cast<u64*>(s_hist);
            const uint2* rec2 = reinterpret_cast<const uint2*>(a.rec);   // (slot, Lf) = the low half of a record
            auto word_of = [&](int r) { const uint2 q = rec2[2 * (size_t)(bt.first + s_idx[r])]; return (u64)q.x | ((u64)q.y << 32); };
            for (int r = lane; r < cnt && r < SC; r += 32) s_word[r] = word_of(r);
            __syncwarp();
            for (int r = lane; r < cnt; r += 32) {
                const u64 wr = r < SC ? s_word[r] : word_of(r);
                int rank = 0;
                for (int j = 0; j < cnt; j++) {
                    const u64 wj = j < SC ? s_word[j] : word_of(j);
                    rank += wj < wr;
                }
                a.stage_lx[bt.first + rank] = ofg_log10_svml((double)(u32)(wr >> 32));
                a.stage_ly[bt.first + rank] = ofg_log10_svml(__longlong_as_double((long long)s_bits[r]));
            }
        } else {
            for (int r = lane; r < cnt; r += 32) {
                const u32 i = s_idx[r];
                a.stage_lx[bt.first + r] = ofg_log10_svml((double)a.key[bt.first + i]);
                a.stage_ly[bt.first + r] = ofg_log10_svml(__longlong_as_double((long long)s_bits[r]));
            }
        }
        __syncwarp();
    }
}

// Concatenates the staged selections of all bins (exclusive scan of their counts in sel_off) into lx / ly.
__global__ void __launch_bounds__(BIN_THREADS) bin_select_kernel(BinArgs a)
{
    const int w = threadIdx.x >> 5, lane = lane_id();
    const int64_t n_warps = (int64_t)gridDim.x * BIN_WARPS;
    const int64_t n_tasks = *a.n_tasks_dev;
    for (int64_t t = (int64_t)blockIdx.x * BIN_WARPS + w; t < n_tasks; t += n_warps) {
        const u32 first = a.tasks[t].first;
        const u64 o0 = a.sel_off[t];
        const u32 cnt = (u32)(a.sel_off[t + 1] - o0);
        for (u32 r = lane; r < cnt; r += 32) {
            const u64 o = o0 + r;
            if (o < a.sel_cap) {
                a.lx[o] = a.stage_lx[first + r];
                a.ly[o] = a.stage_ly[first + r];
            }
        }
    }
}
