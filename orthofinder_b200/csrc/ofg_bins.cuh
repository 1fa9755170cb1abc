// Stage NORM, parts 1-2 (fast path): length bins without a full sort.
//
// scnorm.GetTopPercentileOfScores (gathering.py:93-116) sorts a pair's hits by (Lf, original index), cuts the order into
// bins of binSize consecutive hits and keeps the top 5 % of every bin by score, in sorted order.  Only two things of
// that order are ever observed: WHICH bin a hit falls into, and the relative order of the ~5 % that are kept.  So
// instead of sorting 12-byte records three times (the LSD radix path of ofg_norm.cuh, kept as the fallback), this
// path
//   bk_hist       histograms the keys of a pair over 65 536 monotone buckets (exponent + 11 mantissa bits of Lf),
//   bk_scan       turns the histogram into the rank of every bucket's first hit and finds the buckets that contain a
//                 bin boundary (a rank that is a multiple of binSize),
//   bk_gather     collects the (Lf, index) words of the hits in those boundary buckets (10-20 % of all hits),
//   bk_split      selects, per boundary, the exact (Lf, index) word at the boundary rank (radix select in shared memory),
//   bk_binhist    assigns every hit its bin (bucket rank / binSize, plus one comparison against the split word in a
//                 boundary bucket) and counts it in a 512-bucket histogram of the bin's scores,
//   bk_thr        finds, per bin, the score bucket that holds the order statistic of the 95th percentile,
//   bk_cand       appends the hits at or above that bucket (~6 %) to the bin's slot range - unordered,
//   bin_cut       (ofg_norm.cuh, unordered variant) takes the percentile among the candidates and orders just the
//                 selected hits by their (Lf, index) words.
// All comparisons are on integers, so bins and order are exactly the reference's.  Inputs that defeat the buckets
// (a boundary bucket with more than BK_CAP hits, e.g. all lengths equal; more than BK_MAXBINS bins in a pair) raise
// a device-side flag and the wave runs the radix path instead; both paths are enqueued, the one that is not taken
// sees zero work.
#pragma once
#include "ofg_common.cuh"
#include "ofg_norm.cuh"

constexpr int BK_M = 11;                 // mantissa bits of the bucket index
constexpr int BK_NB = 1 << 16;           // buckets per pair: 31 exponents x 2^11 (keys < 2^31)
constexpr int BK_CAP = 2048;             // most hits of a boundary bucket (shared memory of one warp in bk_split)
constexpr int BK_MAXBINS = 8192;         // most bins of a pair (shared-memory counters of bk_scatter; bins are kept as 16-bit ids)
constexpr u32 BK_FALLBACK = 1u;

// monotone in key: position of the highest set bit and the BK_M bits below it
__device__ __forceinline__ u32 bk_bucket(u32 key)
{
    if (key == 0u) return 0u;
    const u32 bits = __float_as_uint(__uint2float_rz(key));   // round towards zero keeps the map monotone
    return (((bits >> 23) - 127u) << BK_M) | ((bits >> (23 - BK_M)) & ((1u << BK_M) - 1u));
}

struct __align__(16) BbDesc {   // one boundary bucket
    int pair;        // wave-relative
    u32 bucket;
    u32 r0;          // rank of the bucket's first hit inside the pair
    u32 cnt;
    u64 off;         // first entry in the gathered list
    u32 t0;          // bin of the bucket's first hit
    u32 last;        // last boundary inside the bucket (<= number of bins of the pair): boundaries t0 + 1 .. last
};

struct BkArgs {
    const u32* key;                 // [slots of the wave] Lf per slot, key_sentinel for unused slots
    u32 key_sentinel;
    const double* vals;             // CSR values (absolute slots)
    const u32* slot0;               // first CSR slot of the wave
    const RsTileDesc* desc;         // tiles of RS_TILE slots, never across pairs
    const int64_t* n_tiles_all;     // tiles of the wave
    const int64_t* n_tiles_nw;      // the same, or 0 when the wave falls back to the radix path
    const int64_t* pair_slot;       // [Pw + 1]
    const u64* task_prefix;         // [Pw + 1] bins before each pair
    const u64* pair_nnz;            // [Pw] stored hits of each pair (already offset to the wave)
    int Pw;
    u32* hist;                      // [Pw][BK_NB]
    u32* binof;                     // [Pw][BK_NB] 0: beyond the last whole bin (dropped); (bin + 1) << 1: the bucket lies inside one bin;
                                    //             (index into bb) << 1 | 1: boundary bucket
    BbDesc* bb;
    u32 bb_cap;
    u32* n_bb;                      // device counters
    unsigned long long* list_used;
    u64 list_cap;
    u64* list;                      // gathered (Lf << 32 | wave-relative slot) words of the boundary buckets
    u32* bbfill;                    // [bb_cap]
    u64* split;                     // [bins of the wave] split[task_prefix[p] + t - 1]: word at rank t * binSize of pair p
    u32* gcursor;                   // [bins of the wave] hits appended to each bin so far
    uint4* rec_b;                   // out [slots] binned hits: (slot, Lf, score lo, score hi) - x | y << 32 is the (Lf, index) word
    u32* flags;                     // BK_FALLBACK
};

// Consecutive tiles of one pair, at most `max_tiles` of them, starting at tile t0: returns the number of tiles and the
// slot range [e0, e0 + n).  Tiles of a pair are consecutive and all but the last are full.
__device__ __forceinline__ int bk_segment(const RsTileDesc* desc, int64_t t0, int64_t t_end, int max_tiles, int* pair, int64_t* e0, int* n)
{
    const RsTileDesc d0 = desc[t0];
    *pair = d0.pair;
    *e0 = d0.e0;
    int nt = 1, cnt = d0.cnt;
    while (nt < max_tiles && t0 + nt < t_end) {
        const RsTileDesc d = desc[t0 + nt];
        if (d.pair != d0.pair) break;
        cnt += d.cnt;
        nt++;
    }
    *n = cnt;
    return nt;
}

// ---- 1: histogram.  A CTA aggregates up to 15 tiles (61 440 slots) of one pair in shared memory - two 16-bit counters per
// word, which cannot overflow below 65 536 slots - and adds the non-zero counters to the pair's table.
constexpr int BKH_TILES = 15;
constexpr int BKH_THREADS = 1024;
__global__ void __launch_bounds__(BKH_THREADS, 1) bk_hist_kernel(BkArgs a)
{
    extern __shared__ __align__(16) u32 s_cnt[];   // BK_NB / 2 words
    const int64_t n_tiles = *a.n_tiles_all;
    for (int64_t g0 = (int64_t)blockIdx.x * BKH_TILES; g0 < n_tiles; g0 += (int64_t)gridDim.x * BKH_TILES) {
        const int64_t g1 = g0 + BKH_TILES < n_tiles ? g0 + BKH_TILES : n_tiles;
        int64_t t0 = g0;
        while (t0 < g1) {
            int pair, n;
            int64_t e0;
            const int nt = bk_segment(a.desc, t0, g1, BKH_TILES, &pair, &e0, &n);
            t0 += nt;
            for (int i = threadIdx.x; i < BK_NB / 2; i += BKH_THREADS) s_cnt[i] = 0;
            __syncthreads();
            for (int k0 = threadIdx.x; k0 < n; k0 += 8 * BKH_THREADS) {   // eight independent loads in flight per thread
                u32 key[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int k = k0 + u * BKH_THREADS;
                    key[u] = k < n ? a.key[e0 + k] : a.key_sentinel;
                }
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    if (key[u] != a.key_sentinel) {
                        const u32 b = bk_bucket(key[u]);
                        atomicAdd(&s_cnt[b >> 1], (b & 1u) ? 0x10000u : 1u);
                    }
                }
            }
            __syncthreads();
            u32* h = a.hist + (size_t)pair * BK_NB;
            for (int i = threadIdx.x; i < BK_NB / 2; i += BKH_THREADS) {
                const u32 v = s_cnt[i];
                if (v & 0xffffu) atomicAdd(&h[2 * i], v & 0xffffu);
                if (v >> 16) atomicAdd(&h[2 * i + 1], v >> 16);
            }
            __syncthreads();
        }
    }
}

// ---- 2: bin of every bucket and the boundary buckets; one CTA per pair
__global__ void __launch_bounds__(1024) bk_scan_kernel(BkArgs a)
{
    __shared__ u32 s_scan[1024 / 32 + 1];
    const int pair = blockIdx.x;
    const u64 H = a.pair_nnz[pair];
    u32 n_in = 0;
    const u64 n_bins = ofg_bins_of(H, &n_in);
    if (n_bins > (u64)BK_MAXBINS && threadIdx.x == 0) atomicOr(a.flags, BK_FALLBACK);
    const u32* h = a.hist + (size_t)pair * BK_NB;
    u32* bo = a.binof + (size_t)pair * BK_NB;
    constexpr int PER = BK_NB / 1024;
    const int b0 = (int)threadIdx.x * PER;
    u32 sum = 0;
#pragma unroll
    for (int q = 0; q < PER / 4; q++) {
        const uint4 v = reinterpret_cast<const uint4*>(h + b0)[q];
        sum += v.x + v.y + v.z + v.w;
    }
    u32 total;
    u32 r0 = block_excl_scan<1024>(sum, s_scan, &total);
    if (sum == 0 || n_in == 0) return;   // nothing in these buckets (they are never looked up)
    for (int q = 0; q < PER; q++) {
        const u32 c = h[b0 + q];
        if (c) {
            // boundaries strictly inside the bucket: ranks t * n_in with r0 < t * n_in < r0 + c, t <= n_bins
            const u64 t0 = (u64)r0 / n_in, t_hi = ((u64)r0 + c - 1) / n_in;
            u32 e = t0 < n_bins ? (u32)((t0 + 1) << 1) : 0u;   // hits at or beyond rank n_bins * n_in are dropped (gathering.py:104-107)
            if (t_hi > t0 && t0 + 1 <= n_bins) {
                bool ok = false;
                if (c <= (u32)BK_CAP) {
                    const u32 idx = atomicAdd(a.n_bb, 1u);
                    if (idx < a.bb_cap) {
                        const u64 off = atomicAdd(a.list_used, (unsigned long long)c);
                        if (off + c <= a.list_cap) {
                            BbDesc d;
                            d.pair = pair; d.bucket = (u32)(b0 + q); d.r0 = r0; d.cnt = c; d.off = off;
                            d.t0 = (u32)t0; d.last = (u32)(t_hi < n_bins ? t_hi : n_bins);
                            a.bb[idx] = d;
                            e = (idx << 1) | 1u;
                            ok = true;
                        }
                    }
                }
                if (!ok) atomicOr(a.flags, BK_FALLBACK);
            }
            bo[b0 + q] = e;
            r0 += c;
        }
    }
}

// after bk_scan: which path runs
__global__ void bk_route_kernel(const u32* flags, const int64_t* n_tiles, const int64_t* n_tasks, u32* n_bb, u32 bb_cap,
                                int64_t* n_tiles_fb, int64_t* n_tasks_fb, int64_t* n_tiles_nw, int64_t* n_tasks_nw, u64* fallback_waves)
{
    const bool fb = (*flags & BK_FALLBACK) != 0;
    if (fb) *fallback_waves += 1;
    *n_tiles_fb = fb ? *n_tiles : 0;
    *n_tasks_fb = fb ? *n_tasks : 0;
    *n_tiles_nw = fb ? 0 : *n_tiles;
    *n_tasks_nw = fb ? 0 : *n_tasks;
    if (fb || *n_bb > bb_cap) *n_bb = 0;
}

// ---- 3: (Lf, slot) words of the hits in boundary buckets
__global__ void __launch_bounds__(512) bk_gather_kernel(BkArgs a)
{
    const int64_t n_tiles = *a.n_tiles_nw;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const RsTileDesc td = a.desc[tile];
        const u32* bo = a.binof + (size_t)td.pair * BK_NB;
        u32 key[RS_TILE / 512], e[RS_TILE / 512];
#pragma unroll
        for (int it = 0; it < RS_TILE / 512; it++) {
            const int k = it * 512 + (int)threadIdx.x;
            key[it] = k < td.cnt ? a.key[td.e0 + k] : a.key_sentinel;
        }
#pragma unroll
        for (int it = 0; it < RS_TILE / 512; it++) e[it] = key[it] != a.key_sentinel ? bo[bk_bucket(key[it])] : 0u;
#pragma unroll
        for (int it = 0; it < RS_TILE / 512; it++) {
            if (e[it] & 1u) {
                const int k = it * 512 + (int)threadIdx.x;
                const u32 i = e[it] >> 1;
                const u32 pos = atomicAdd(&a.bbfill[i], 1u);
                a.list[a.bb[i].off + pos] = ((u64)key[it] << 32) | (u64)(u32)(td.e0 + k);
            }
        }
    }
}

// ---- 4: the word at every boundary rank; one warp per boundary bucket
constexpr int BKS_WARPS = 8;
constexpr int BKS_STAGE = 512;   // words of a boundary bucket staged in shared memory (larger buckets, up to BK_CAP, are selected from global memory)
__global__ void __launch_bounds__(BKS_WARPS * 32) bk_split_kernel(BkArgs a)
{
    __shared__ u64 s_bits_all[BKS_WARPS][BKS_STAGE];
    __shared__ u32 s_hist_all[BKS_WARPS][256];
    const int w = threadIdx.x >> 5, lane = lane_id();
    u64* s_stage = s_bits_all[w];
    u32* s_hist = s_hist_all[w];
    const u32 n_bb = *a.n_bb;
    const int64_t n_warps = (int64_t)gridDim.x * BKS_WARPS;
    for (int64_t i = (int64_t)blockIdx.x * BKS_WARPS + w; i < (int64_t)n_bb; i += n_warps) {
        const BbDesc d = a.bb[i];
        const int n = (int)d.cnt;
        const bool staged = n <= BKS_STAGE;
        const u64* s_bits = staged ? s_stage : a.list + d.off;
        u64 vmin = ~0ULL, vmax = 0ULL;
        for (int k = lane; k < n; k += 32) {
            const u64 v = a.list[d.off + k];
            if (staged) s_stage[k] = v;
            vmin = v < vmin ? v : vmin;
            vmax = v > vmax ? v : vmax;
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            const u64 o1 = __shfl_xor_sync(OFG_FULL, vmin, s), o2 = __shfl_xor_sync(OFG_FULL, vmax, s);
            vmin = o1 < vmin ? o1 : vmin;
            vmax = o2 > vmax ? o2 : vmax;
        }
        __syncwarp();
        u32 n_in;
        const u64 n_bins = ofg_bins_of(a.pair_nnz[d.pair], &n_in);
        const u64 t_lo = (u64)d.r0 / n_in + 1;
        u64 t_hi = ((u64)d.r0 + d.cnt - 1) / n_in;
        if (t_hi > n_bins) t_hi = n_bins;
        for (u64 t = t_lo; t <= t_hi; t++) {
            const int k = (int)(t * n_in - d.r0);   // 1 <= k <= cnt - 1: the hit of in-bucket rank k opens bin t
            int n_le;
            const u64 word = warp_kth(s_bits, n, k, vmin, vmax, s_hist, &n_le);
            if (lane == 0) a.split[a.task_prefix[d.pair] + t - 1] = word;
            __syncwarp();
        }
        __syncwarp();
    }
}

// bin of one hit (0xffff: dropped); word = Lf << 32 | slot
__device__ __forceinline__ u32 bk_bin_of(const BkArgs& a, u32 e, u64 word, u32 n_bins, u64 tp)
{
    if (!(e & 1u)) return e ? (e >> 1) - 1u : 0xffffu;
    const uint2 tl = *reinterpret_cast<const uint2*>(&a.bb[e >> 1].t0);
    u32 tt = tl.x;
    for (u32 q = tl.x + 1; q <= tl.y; q++)   // almost always one boundary
        if (word >= a.split[tp + q - 1]) tt = q;
    return tt < n_bins ? tt : 0xffffu;
}

// ---- 5: bin and score histogram of every hit.  The 95th percentile of a bin only concerns its top ~5 %: instead of moving
// all hits into their bins, every hit adds itself to a 512-bucket histogram of its bin's scores (monotone buckets: exponent
// + 5 mantissa bits); bk_thr then finds the bucket that holds the percentile's order statistic, and only the hits at or
// above that bucket (the candidates, ~6 %) are appended to the bin's slot range by bk_cand.
constexpr int BK_SB = 512;                    // score buckets per bin, two 16-bit counters per word (a bin has <= 1000 hits)

__device__ __forceinline__ u32 bk_sbucket(double v)   // monotone in v > 0; 2^-2 .. 2^14 resolved, the rest clamped
{
    const int s = (int)((u64)__double_as_longlong(v) >> 47) - (1021 << 5);
    return (u32)(s < 0 ? 0 : (s > BK_SB - 1 ? BK_SB - 1 : s));
}

struct BkCandArgs {
    unsigned short* bin16;   // [slots of the wave] bin of every slot (0xffff: none)
    u32* shist;              // [bins of the wave][BK_SB / 2]
    unsigned short* thr;     // [bins] first candidate bucket
    unsigned short* below;   // [bins] hits of the bin in lower buckets
    const int64_t* n_tasks_nw;
    const BinTask* tasks;
    int k20, k200, k1000;    // numpy's order statistic index floor((n - 1) * 0.95)
};

__global__ void __launch_bounds__(512) bk_binhist_kernel(BkArgs a, BkCandArgs c)
{
    const int64_t n_tiles = *a.n_tiles_nw;
    const int64_t voff = (int64_t)*a.slot0;
    constexpr int ITEMS = RS_TILE / 512;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const RsTileDesc td = a.desc[tile];
        const int pair = td.pair;
        u32 n_in;
        const u32 n_bins = (u32)ofg_bins_of(a.pair_nnz[pair], &n_in);
        const u64 tp = a.task_prefix[pair];
        const u32* bo = a.binof + (size_t)pair * BK_NB;
        u32 key[ITEMS], e[ITEMS];
        double v[ITEMS];
#pragma unroll
        for (int it = 0; it < ITEMS; it++) {
            const int k = it * 512 + (int)threadIdx.x;
            key[it] = k < td.cnt ? a.key[td.e0 + k] : a.key_sentinel;
        }
#pragma unroll
        for (int it = 0; it < ITEMS; it++) {
            const int k = it * 512 + (int)threadIdx.x;
            const bool on = key[it] != a.key_sentinel && n_in;
            e[it] = on ? bo[bk_bucket(key[it])] : 0u;
            v[it] = on ? a.vals[voff + td.e0 + k] : 1.0;
        }
#pragma unroll
        for (int it = 0; it < ITEMS; it++) {
            const int k = it * 512 + (int)threadIdx.x;
            if (k < td.cnt) {
                const u32 bin = bk_bin_of(a, e[it], ((u64)key[it] << 32) | (u64)(u32)(td.e0 + k), n_bins, tp);
                c.bin16[td.e0 + k] = (unsigned short)bin;
                if (bin != 0xffffu) {
                    const u32 sb = bk_sbucket(v[it]);
                    atomicAdd(&c.shist[(tp + bin) * (BK_SB / 2) + (sb >> 1)], (sb & 1u) ? 0x10000u : 1u);
                }
            }
        }
    }
}

// ---- 6: per bin, the score bucket that holds the order statistic a[k] of numpy's percentile (k = floor((n - 1) * 0.95)) and
// the number of hits below that bucket; bins that are taken whole (pairs with < 100 hits) start at bucket 0.  One warp per bin.
__global__ void __launch_bounds__(256) bk_thr_kernel(BkCandArgs c)
{
    const int lane = lane_id();
    const int64_t n_tasks = *c.n_tasks_nw;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < n_tasks; t += n_warps) {
        const BinTask bt = c.tasks[t];
        const int n = (int)bt.n;
        const u32 k = bt.take_all ? 0u : (u32)(n == 1000 ? c.k1000 : (n == 200 ? c.k200 : c.k20));
        // lane l owns buckets [16 l, 16 l + 16)
        const uint4* w4 = reinterpret_cast<const uint4*>(c.shist + (size_t)t * (BK_SB / 2) + lane * 8);
        const uint4 q0 = w4[0], q1 = w4[1];
        const u32 w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
        u32 sum = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) sum += (w[i] & 0xffffu) + (w[i] >> 16);
        const u32 inc = warp_incl_scan(sum);
        const u32 exc = inc - sum;
        const bool mine = !bt.take_all && k >= exc && k < inc;
        if (mine) {
            u32 run = exc, thr = 0, below = exc;
            bool found = false;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const u32 cnt = (i & 1) ? (w[i >> 1] >> 16) : (w[i >> 1] & 0xffffu);
                if (!found && k < run + cnt) { thr = (u32)(lane * 16 + i); below = run; found = true; }
                run += cnt;
            }
            c.thr[t] = (unsigned short)thr;
            c.below[t] = (unsigned short)below;
        }
        if (bt.take_all && lane == 0) { c.thr[t] = 0; c.below[t] = 0; }
    }
}

// ---- 7: the candidates of every bin, appended (in arbitrary order) to the bin's slot range
__global__ void __launch_bounds__(512, 3) bk_cand_kernel(BkArgs a, BkCandArgs c)
{
    const int64_t n_tiles = *a.n_tiles_nw;
    const int64_t voff = (int64_t)*a.slot0;
    constexpr int ITEMS = RS_TILE / 512;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const RsTileDesc td = a.desc[tile];
        const int pair = td.pair;
        u32 n_in;
        ofg_bins_of(a.pair_nnz[pair], &n_in);
        const u64 tp = a.task_prefix[pair];
        const int64_t ps = a.pair_slot[pair];
        u32 bin[ITEMS], thr[ITEMS];
        double v[ITEMS];
#pragma unroll
        for (int it = 0; it < ITEMS; it++) {
            const int k = it * 512 + (int)threadIdx.x;
            bin[it] = k < td.cnt ? (u32)c.bin16[td.e0 + k] : 0xffffu;
            v[it] = bin[it] != 0xffffu ? a.vals[voff + td.e0 + k] : 0.0;
        }
#pragma unroll
        for (int it = 0; it < ITEMS; it++) thr[it] = bin[it] != 0xffffu ? (u32)c.thr[tp + bin[it]] : 0xffffu;
#pragma unroll
        for (int it = 0; it < ITEMS; it++) {
            if (bin[it] != 0xffffu && bk_sbucket(v[it]) >= thr[it]) {
                const int k = it * 512 + (int)threadIdx.x;
                const u32 p = atomicAdd(&a.gcursor[tp + bin[it]], 1u);
                const u64 vb = (u64)__double_as_longlong(v[it]);
                a.rec_b[ps + (int64_t)bin[it] * n_in + p] = make_uint4((u32)(td.e0 + k), a.key[td.e0 + k], (u32)vb, (u32)(vb >> 32));
            }
        }
    }
}
