// Stage RAW, part 1: device-side parse of decompressed Blast{p}_{j}.txt into COO records
// (global row id, hit id, bit score).  Replaces the csv.reader loop of
// blast_file_processor.py:66-88 (GetBLAST6Scores).
//
// Layout: all files of the context live back to back in one text arena, each padded with '\n' to a
// multiple of 16 bytes (blank lines are skipped by the reference, :68-69, so padding is neutral).
// A CTA takes a 16 KB tile of one file, stages it (plus 1 KB of overlap for the line that crosses
// the tile end) in shared memory with 16-byte loads, marks line starts with SIMD-in-word byte
// compares + a CTA scan, then parses one line per thread out of shared memory.  Records are written
// at (tile base + line index), the tile base coming from one atomicAdd per tile; lines that are
// dropped (self hits, score <= 0) leave a sentinel row id.  Row histogram (hits per query gene)
// is accumulated with warp-aggregated atomics (consecutive lines share the query).
#pragma once
#include "ofg_common.cuh"
#include "ofg_parse_line.h"

constexpr int PARSE_TILE = 16384;
constexpr int PARSE_OVERLAP = 1024;
constexpr int PARSE_THREADS = 256;
constexpr int PARSE_CPT = PARSE_TILE / 32 / PARSE_THREADS;  // 32-byte chunks of the tile body per thread (adjacent)
static_assert(PARSE_CPT * 32 * PARSE_THREADS == PARSE_TILE, "tile must split evenly");
constexpr int PARSE_WORDS = (PARSE_TILE + PARSE_OVERLAP) / 32;

struct ParseArgs {
    const uint8_t* text;
    const PairDesc* pairs;
    int P;
    const int64_t* tile_prefix;  // [P+1] tiles before each pair
    int64_t n_tiles;
    u32* coo_row;
    u32* coo_col;
    double* coo_val;
    u64 cap;
    u64* coo_count;   // records reserved so far (may exceed cap: host checks)
    u32* rowcount;    // [total rows] kept records per (pair, query)
    u64* pair_lines;  // [P]
    u64* pair_valid;  // [P]
    u64* pair_err;    // [P] min over offending lines of (byte offset << 4 | kind), ~0 if none
};

struct SmemSrc {
    const uint8_t* s;
    int limit;
    bool truncated;
    __device__ __forceinline__ int get(int64_t pos)
    {
        if (pos >= limit) { truncated = true; return '\n'; }
        return s[pos];
    }
};
struct GlobalSrc {
    const uint8_t* g;
    int64_t limit;
    __device__ __forceinline__ int get(int64_t pos) { return pos >= limit ? '\n' : g[pos]; }
};

// 0x80 in every byte of w that equals the pattern byte (exact, no cross-byte carries)
__device__ __forceinline__ u32 bytes_eq_hi(u32 w, u32 pattern)
{
    const u32 y = w ^ pattern;
    return ~(((y & 0x7f7f7f7fu) + 0x7f7f7f7fu) | y | 0x7f7f7f7fu);
}
// gathers the four 0x80 flags into bits 0..3 (the partial products never collide below bit 32)
__device__ __forceinline__ u32 hi_to_nibble(u32 t) { return (t * 0x00204081u) >> 28; }

// a * b + c on the FMA pipe (the kernel is bound by the integer ALU pipe)
__device__ __forceinline__ u32 umad32(u32 a, u32 b, u32 c)
{
    u32 r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

// Structural bitmaps of one 32-byte chunk (bit i = byte i): line terminators (\n, \r) and tabs.
// Fast form: for bytes below 0x80, (byte ^ c) + 0x7f has its top bit clear exactly when byte == c, with no
// carries between bytes.  Returns false if the chunk holds a byte >= 0x80 (carries possible) or a double
// quote (csv quoting is not supported); the caller then uses the exact, carry-free form.
__device__ __forceinline__ bool chunk_masks_fast(const u32 (&xs)[8], u32* term, u32* tabs)
{
    const u32 H = 0x80808080u, A = 0x7f7f7f7fu;
    u32 good = H, tm = 0, tb = 0;
#pragma unroll
    for (int b = 0; b < 8; b++) {
        const u32 x = xs[b];
        const u32 n_nl = (x ^ 0x0a0a0a0au) + A, n_cr = (x ^ 0x0d0d0d0du) + A;
        const u32 n_tb = (x ^ 0x09090909u) + A, n_qu = (x ^ 0x22222222u) + A;
        good = good & n_qu & ~x;
        tm = umad32(hi_to_nibble(~(n_nl & n_cr) & H), 1u << (4 * b), tm);   // nibbles do not overlap: multiply-add == shift-or
        tb = umad32(hi_to_nibble(~n_tb & H), 1u << (4 * b), tb);
    }
    *term = tm;
    *tabs = tb;
    return good == H;
}
__device__ __forceinline__ void chunk_masks_exact(const u32 (&xs)[8], u32* term, u32* tabs, u32* quotes)
{
    u32 tm = 0, tb = 0, qu = 0;
#pragma unroll
    for (int b = 0; b < 8; b++) {
        const u32 x = xs[b];
        tm |= hi_to_nibble(bytes_eq_hi(x, 0x0a0a0a0au) | bytes_eq_hi(x, 0x0d0d0d0du)) << (4 * b);
        tb |= hi_to_nibble(bytes_eq_hi(x, 0x09090909u)) << (4 * b);
        qu |= hi_to_nibble(bytes_eq_hi(x, 0x22222222u)) << (4 * b);
    }
    *term = tm;
    *tabs = tb;
    *quotes = qu;
}

// ---- warp-converged fast path --------------------------------------------------------------------------
// Every lane of a warp holds one record (or none).  All loops below have warp-uniform trip counts and use
// predication instead of early exits, so the 32 records are parsed in lock step; a record that does not fit
// the plain shapes BLAST / DIAMOND / MMseqs2 write ("<species>_<digits>", "<digits>[.<digits>]", <= 9 digits,
// 12+ fields within 128 bytes) is handed to the generic parser afterwards.

// 128-bit window of a bitmap starting at byte st (bit i = byte st + i)
__device__ __forceinline__ void rel_mask128(const u32* m, int st, u64* lo, u64* hi)
{
    const int w0 = st >> 5, sh = st & 31;
    const u32 a0 = m[w0], a1 = m[w0 + 1], a2 = m[w0 + 2], a3 = m[w0 + 3], a4 = m[w0 + 4];
    const u32 r0 = __funnelshift_r(a0, a1, sh), r1 = __funnelshift_r(a1, a2, sh);
    const u32 r2 = __funnelshift_r(a2, a3, sh), r3 = __funnelshift_r(a3, a4, sh);
    *lo = (u64)r0 | ((u64)r1 << 32);
    *hi = (u64)r2 | ((u64)r3 << 32);
}
__device__ __forceinline__ int ctz128(u64 lo, u64 hi) { return lo ? __ffsll((long long)lo) - 1 : 64 + __ffsll((long long)hi) - 1; }

// Field boundaries relative to the record start for a record with EXACTLY 12 fields inside a 128-byte
// window: first and second tab, the last (11th) tab and the line end.  Anything else is left to the
// generic parser.
__device__ __forceinline__ bool locate_fields_conv(const u32* s_tab, const u32* s_term, int st, bool act, int* tab1, int* tab2,
                                                   int* tab11, int* line_end)
{
    u64 ea, eb, ta, tb;
    rel_mask128(s_term, st, &ea, &eb);
    rel_mask128(s_tab, st, &ta, &tb);
    const int le = (ea | eb) ? ctz128(ea, eb) : 128;
    bool ok = act && le < 128 && st + le < PARSE_TILE + PARSE_OVERLAP;
    // tabs of this record only
    if (le < 64) { ta &= (1ULL << le) - 1ULL; tb = 0; }
    else if (le < 128) { tb &= (1ULL << (le - 64)) - 1ULL; }
    ok = ok && (__popcll(ta) + __popcll(tb) == 11);
    const u64 t2m = ta & (ta - 1ULL);
    ok = ok && t2m != 0;  // the two id fields end inside the first 64 bytes
    *tab1 = __ffsll((long long)ta) - 1;
    *tab2 = __ffsll((long long)t2m) - 1;
    *tab11 = tb ? 127 - __clzll((long long)tb) : 63 - __clzll((long long)ta);
    *line_end = le;
    return ok;
}

// 8 bytes [end - 8, end) of the staged text as a little-endian u64 (byte end-1 is the most significant).
// The buffer has a 16-byte front pad, so end >= 0 is enough.
__device__ __forceinline__ u64 load8_ending_at(const uint8_t* s_text, int end)
{
    const int a = end - 8;
    const u32* wp = reinterpret_cast<const u32*>(s_text + (a & ~3));  // (a & ~3) >= -16: inside the pad
    const u32 a0 = wp[0], a1 = wp[1], a2 = wp[2];
    const int sh = (a & 3) * 8;
    return (u64)__funnelshift_r(a0, a1, sh) | ((u64)__funnelshift_r(a1, a2, sh) << 32);
}

// 0x80 in every byte of y that is zero (exact)
__device__ __forceinline__ u64 zero_bytes64(u64 y)
{
    const u64 m = 0x7f7f7f7f7f7f7f7fULL;
    return ~(((y & m) + m) | y | m);
}

// X holds 8 ASCII characters, first character in the lowest byte.  Returns false unless all are digits;
// *out = their decimal value (SWAR, three multiplies).
__device__ __forceinline__ bool swar_digits8(u64 X, u32* out)
{
    const bool ok = (((X + 0x4646464646464646ULL) | (X - 0x3030303030303030ULL)) & 0x8080808080808080ULL) == 0;
    u64 v = X & 0x0F0F0F0F0F0F0F0FULL;
    v = (v * 2561ULL) >> 8;
    v = ((v & 0x00FF00FF00FF00FFULL) * 6553601ULL) >> 16;
    v = ((v & 0x0000FFFF0000FFFFULL) * 42949672960001ULL) >> 32;
    *out = (u32)v;
    return ok;
}

// "<anything>_<1..7 digits>" of length 2..8 with exactly one underscore, ending at byte `end` (the tab).
__device__ __forceinline__ bool swar_id(const uint8_t* s_text, int end, int len, bool act, u32* out)
{
    act = act && len >= 2 && len <= 8;
    const int l = act ? len : 8, e = act ? end : 8;
    const u64 w = load8_ending_at(s_text, e);
    const u64 vm = ~0ULL << (8 * (8 - l));
    const u64 us = zero_bytes64(w ^ 0x5f5f5f5f5f5f5f5fULL) & vm;
    const bool one = us != 0 && (us & (us - 1ULL)) == 0;
    const int u = one ? (__ffsll((long long)us) - 1) >> 3 : 0;  // byte index of the underscore
    const int nd = 7 - u;
    const bool shape = one && nd >= 1;
    const int ndc = shape ? nd : 1;
    const u64 D = w >> (8 * (8 - ndc));                       // digits, first digit in the lowest byte
    const u64 X = (D << (8 * (8 - ndc))) | (0x3030303030303030ULL >> (8 * ndc));
    u32 v;
    const bool dig = swar_digits8(X, &v);
    *out = v;
    return act && shape && dig;
}

// "<digits>[.<digits>]" of length 1..8 ending at byte `end`, at most 8 digits: exact fp64 value.
__device__ __forceinline__ bool swar_score(const uint8_t* s_text, int end, int len, bool act, double* out)
{
    act = act && len >= 1 && len <= 8;
    const int l = act ? len : 8, e = act ? end : 8;
    const u64 w = load8_ending_at(s_text, e);
    const u64 vm = ~0ULL << (8 * (8 - l));
    const u64 dots = zero_bytes64(w ^ 0x2e2e2e2e2e2e2e2eULL) & vm;
    const bool has_dot = dots != 0;
    const bool one = (dots & (dots - 1ULL)) == 0;  // zero or one dot
    const int u = has_dot ? (__ffsll((long long)dots) - 1) >> 3 : 7;
    const int frac = has_dot ? 7 - u : 0;
    // remove the dot: everything below it moves up one byte
    const u64 lower = w & ((1ULL << (8 * u)) - 1ULL);
    const u64 upper = (u == 7) ? 0ULL : ((w >> (8 * (u + 1))) << (8 * (u + 1)));
    const u64 Y = has_dot ? (upper | (lower << 8)) : w;
    const int ndig = has_dot ? l - 1 : l;
    const bool shape = one && ndig >= 1;
    const int nc = shape ? ndig : 1;
    const u64 X = (nc == 8) ? Y : ((Y & (~0ULL << (8 * (8 - nc)))) | (0x3030303030303030ULL >> (8 * nc)));
    u32 m;
    const bool dig = swar_digits8(X, &m);
    const double v = (double)m;
    *out = frac > 0 ? v / OFG_POW10_DEV[frac] : v;  // exact: m < 10^8, one correctly rounded division
    return act && shape && dig;
}

__global__ void __launch_bounds__(PARSE_THREADS, 5) parse_kernel(ParseArgs a)
{
    __shared__ __align__(16) uint8_t s_text_raw[16 + PARSE_TILE + PARSE_OVERLAP];  // 16-byte front pad for load8_ending_at
    uint8_t* const s_text = s_text_raw + 16;
    __shared__ uint16_t s_lstart[PARSE_TILE / 2];
    __shared__ u32 s_term[PARSE_WORDS + 8];  // + padding read by the 128-bit windows: all terminators
    __shared__ u32 s_tab[PARSE_WORDS + 8];
    __shared__ u32 s_scan[PARSE_THREADS / 32 + 1];
    __shared__ u64 s_base;
    __shared__ u32 s_valid;
    __shared__ int s_prev_term;
    const int tid = threadIdx.x;
    if (tid < 8) { s_term[PARSE_WORDS + tid] = 0xffffffffu; s_tab[PARSE_WORDS + tid] = 0u; }

    int pair = -1;
    // per-pair constants, reloaded only when the tile loop crosses into another file
    int64_t p_tile0 = 0, p_tile1 = 0, p_text_len = 0, p_row_base = 0;
    const uint8_t* gfile = nullptr;
    u32 p_Gi = 0, p_Gj = 0;
    bool p_same = false, p_swap = false;
    for (int64_t tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        // tiles are visited in increasing order: the pair index only moves forward
        if (pair < 0 || tile >= p_tile1) {
            if (pair < 0) pair = upper_bound_dev(a.tile_prefix, a.P + 1, tile) - 1;
            while (tile >= a.tile_prefix[pair + 1]) pair++;
            const PairDesc& d = a.pairs[pair];
            p_tile0 = a.tile_prefix[pair]; p_tile1 = a.tile_prefix[pair + 1];
            p_text_len = d.text_len; p_row_base = d.row_base; gfile = a.text + d.text_off;
            p_Gi = (u32)d.Gi; p_Gj = (u32)d.Gj; p_same = d.same != 0; p_swap = d.swap != 0;
        }
        const int64_t t0 = (tile - p_tile0) * (int64_t)PARSE_TILE;
        const int64_t remain = p_text_len - t0;
        const int nload = (int)(remain < (int64_t)(PARSE_TILE + PARSE_OVERLAP) ? remain : (int64_t)(PARSE_TILE + PARSE_OVERLAP));
        const int body = (int)(remain < (int64_t)PARSE_TILE ? remain : (int64_t)PARSE_TILE);
        {
            const uint4* g16 = reinterpret_cast<const uint4*>(gfile + t0);
            uint4* s16 = reinterpret_cast<uint4*>(s_text);
            constexpr int FULL16 = (PARSE_TILE + PARSE_OVERLAP) / 16;
            if (nload == PARSE_TILE + PARSE_OVERLAP) {  // every tile but a file's last: fixed trip count, loads in flight together
                uint4 v[FULL16 / PARSE_THREADS + 1];
#pragma unroll
                for (int k = 0; k < FULL16 / PARSE_THREADS; k++) v[k] = __ldg(g16 + (k * PARSE_THREADS + tid));
                constexpr int REST = FULL16 % PARSE_THREADS;
                if (REST && tid < REST) v[FULL16 / PARSE_THREADS] = __ldg(g16 + ((FULL16 / PARSE_THREADS) * PARSE_THREADS + tid));
#pragma unroll
                for (int k = 0; k < FULL16 / PARSE_THREADS; k++) s16[k * PARSE_THREADS + tid] = v[k];
                if (REST && tid < REST) s16[(FULL16 / PARSE_THREADS) * PARSE_THREADS + tid] = v[FULL16 / PARSE_THREADS];
            } else {
                for (int i = tid; i < nload / 16; i += PARSE_THREADS) s16[i] = __ldg(g16 + i);
                const uint4 nl = make_uint4(0x0a0a0a0au, 0x0a0a0a0au, 0x0a0a0a0au, 0x0a0a0a0au);
                for (int i = nload / 16 + tid; i < FULL16; i += PARSE_THREADS) s16[i] = nl;
            }
            if (tid == 0) {
                s_prev_term = (t0 == 0) ? 1 : (int)ofg_is_term(gfile[t0 - 1]);
                s_valid = 0;
            }
        }
        __syncthreads();
        // structural bitmaps of the whole staged window (32 bytes per word): terminators and tabs.
        // Thread t owns the adjacent body words [t*CPT, (t+1)*CPT) (their line starts come out in order);
        // the overlap words are shared out to the first threads.
        u32 starts[PARSE_CPT];
        u32 quotes = 0;
        int n_starts = 0;
#pragma unroll
        for (int k = 0; k < PARSE_CPT + 1; k++) {
            int ch;
            if (k < PARSE_CPT) ch = tid * PARSE_CPT + k;
            else { ch = PARSE_TILE / 32 + tid; if (ch >= PARSE_WORDS) break; }
            const uint4* w4 = reinterpret_cast<const uint4*>(s_text + ch * 32);
            const uint4 lo4 = w4[0], hi4 = w4[1];
            const u32 xs[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
            u32 term, tabs, qu = 0;
            if (!chunk_masks_fast(xs, &term, &tabs)) chunk_masks_exact(xs, &term, &tabs, &qu);
            s_term[ch] = term;
            s_tab[ch] = tabs;
            if (k < PARSE_CPT) {  // body word: line starts
                const int off = ch * 32;
                const u32 prev = (ch == 0) ? (u32)s_prev_term : (u32)ofg_is_term(s_text[off - 1]);
                u32 st = ~term & ((term << 1) | prev);
                const int nvalid = body - off;  // bytes of this word inside the body
                if (nvalid <= 0) { st = 0; qu = 0; }
                else if (nvalid < 32) { st &= (1u << nvalid) - 1u; qu &= (1u << nvalid) - 1u; }
                starts[k] = st;
                n_starts += __popc(st);
                if (qu && !quotes) quotes = ((u32)off + (u32)__ffs(qu) - 1u) | 0x80000000u;
            }
        }
        if (quotes)  // csv quoting is not supported: report the byte offset of the quote
            atomicMin(&a.pair_err[pair], ((u64)(t0 + (quotes & 0x7fffffffu)) << 4) | (u64)OFG_LINE_QUOTE);
        u32 nl_total;
        u32 pos = block_excl_scan<PARSE_THREADS>((u32)n_starts, s_scan, &nl_total);
#pragma unroll
        for (int k = 0; k < PARSE_CPT; k++) {
            u32 st = starts[k];
            while (st) {
                const int b = __ffs(st) - 1;
                s_lstart[pos++] = (uint16_t)((tid * PARSE_CPT + k) * 32 + b);
                st &= st - 1;
            }
        }
        if (tid == 0) {
            s_base = atomicAdd(a.coo_count, (u64)nl_total);
            atomicAdd(&a.pair_lines[pair], (u64)nl_total);
        }
        __syncthreads();
        const u64 base = s_base;
        for (u32 li0 = 0; li0 < nl_total; li0 += PARSE_THREADS) {
            if (li0 + (u32)(tid & ~31) >= nl_total) continue;  // warp without records (warp-uniform)
            const u32 li = li0 + tid;
            u32 row = OFG_ROW_SENTINEL, col = 0;
            double score = 0.0;
            const bool act = li < nl_total;
            const int st = act ? (int)s_lstart[li] : 0;
            u32 f0 = 0, f1 = 0;
            int tab1, tab2, tab11, lend;
            bool fast = locate_fields_conv(s_tab, s_term, st, act, &tab1, &tab2, &tab11, &lend);
            const bool ok0 = swar_id(s_text, st + tab1, tab1, fast, &f0);
            const bool ok1 = swar_id(s_text, st + tab2, tab2 - tab1 - 1, fast, &f1);
            const bool ok2 = swar_score(s_text, st + lend, lend - tab11 - 1, fast, &score);
            fast = fast && ok0 && ok1 && ok2;
            if (act) {
                int rc = OFG_PL_OK;
                if (!fast) {  // anything unusual: the generic record parser decides (and names the error)
                    SmemSrc ss{s_text, PARSE_TILE + PARSE_OVERLAP, false};
                    rc = ofg_parse_record(ss, (int64_t)st, &f0, &f1, &score);
                    if (ss.truncated) {  // line longer than the staged window: re-parse from global memory
                        GlobalSrc gs{gfile, p_text_len};
                        rc = ofg_parse_record(gs, t0 + st, &f0, &f1, &score);
                    }
                }
                int kind = rc;
                if (rc == OFG_PL_OK) {
                    const u32 q = p_swap ? f1 : f0, s = p_swap ? f0 : f1;
                    if (p_same && q == s) {
                        // self hit: skipped before any range check (:83-84)
                    } else if (q >= p_Gi) {
                        kind = OFG_LINE_RANGE_Q;
                    } else if (s >= p_Gj) {
                        kind = OFG_LINE_RANGE_S;
                    } else if (score > 0.0) {  // stored only if score > current (initially 0) (:87)
                        row = (u32)p_row_base + q;
                        col = s;
                    }
                }
                if (kind != OFG_LINE_OK) atomicMin(&a.pair_err[pair], ((u64)(t0 + st) << 4) | (u64)kind);
                const u64 idx = base + li;
                if (idx < a.cap) {
                    a.coo_row[idx] = row;
                    a.coo_col[idx] = col;
                    a.coo_val[idx] = score;
                }
            }
            // row histogram, aggregated over runs of equal rows (a query's hits are consecutive lines); MATCH.ANY
            // would cost ~64 SM cycles per warp instruction on B200 (tools/micro/match_rate.cu)
            const u32 prev_row = __shfl_up_sync(OFG_FULL, row, 1);
            const u32 heads = __ballot_sync(OFG_FULL, lane_id() == 0 || row != prev_row);
            if (row != OFG_ROW_SENTINEL && ((heads >> lane_id()) & 1u)) {
                const u32 later = heads & ~lanemask_lt() & ~(1u << lane_id());  // heads after this lane
                const int run_end = later ? __ffs(later) - 1 : 32;
                atomicAdd(&a.rowcount[row], (u32)(run_end - (int)lane_id()));
            }
            const u32 vb = __ballot_sync(OFG_FULL, row != OFG_ROW_SENTINEL);
            if (lane_id() == 0 && vb) atomicAdd(&s_valid, (u32)__popc(vb));
        }
        __syncthreads();
        if (tid == 0 && s_valid) atomicAdd(&a.pair_valid[pair], (u64)s_valid);
        __syncthreads();
    }
}
