// Stage RAW, part 1: device-side parse of decompressed Blast{p}_{j}.txt into COO records
// (global row id, hit id, bit score).  Replaces the csv.reader loop of
// blast_file_processor.py:66-88 (GetBLAST6Scores).
//
// Layout: all files of the context live back to back in one text arena, each padded with '\n' to a
// multiple of 16 bytes (blank lines are skipped by the reference, :68-69, so padding is neutral).
// A CTA takes an 8 KB tile of one file, stages it (plus 1 KB of overlap for the line that crosses
// the tile end) in shared memory with 16-byte loads, marks line starts with SIMD-in-word byte
// compares + a CTA scan, then parses one line per thread out of shared memory.  Records are written
// at (tile base + line index), the tile base coming from one atomicAdd per tile; lines that are
// dropped (self hits, score <= 0) leave a sentinel row id.  Row histogram (hits per query gene)
// is accumulated with warp-aggregated atomics (consecutive lines share the query).
#pragma once
#include "ofg_common.cuh"
#include "ofg_parse_line.h"

constexpr int PARSE_TILE = 8192;
constexpr int PARSE_OVERLAP = 1024;
constexpr int PARSE_THREADS = 256;
static_assert(PARSE_TILE / 32 == PARSE_THREADS, "one 32-byte chunk per thread");
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

// Field boundaries of the record starting at byte st of the staged window, from the bitmaps: first and
// second tab, the tab before field 11 and the line end.  False if the record leaves the window or has
// fewer than 12 fields (the generic parser then handles it).
__device__ __forceinline__ bool locate_fields(const u32* s_tab, const u32* s_term, int st, int* tab1, int* tab2, int* tab11,
                                              int* score_end)
{
    const int w0 = st >> 5;
    const u32 low = ~0u << (st & 31);
    int ww = w0;
    u32 m = s_term[ww] & low;
    while (!m) {
        if (++ww >= PARSE_WORDS) return false;
        m = s_term[ww];
    }
    const int le = ww * 32 + __ffs(m) - 1;
    const int wend = le >> 5;
    int cnt = 0, t1 = -1, t2 = -1;
    ww = w0;
    m = s_tab[ww] & low;
    for (;;) {
        const int c = __popc(m);
        if (t1 < 0 && cnt + c >= 1) { u32 x = m; t1 = ww * 32 + __ffs(x) - 1; }
        if (t2 < 0 && cnt + c >= 2) { u32 x = m; for (int k = cnt + 1; k < 2; k++) x &= x - 1; t2 = ww * 32 + __ffs(x) - 1; }
        if (cnt + c >= 11) {
            u32 x = m;
            for (int k = cnt + 1; k < 11; k++) x &= x - 1;
            *tab11 = ww * 32 + __ffs(x) - 1;
            // field 11 ends at the next tab of the line, if any, else at the line end
            x &= x - 1;
            int se = le;
            for (;;) {
                if (x) { const int t12 = ww * 32 + __ffs(x) - 1; if (t12 < le) se = t12; break; }
                if (++ww > wend) break;
                x = s_tab[ww];
            }
            *score_end = se;
            break;
        }
        cnt += c;
        if (++ww > wend) return false;
        m = s_tab[ww];
    }
    *tab1 = t1; *tab2 = t2;
    return *tab11 < le;
}

// Plain fields, the only shapes BLAST / DIAMOND / MMseqs2 write: "<species>_<digits>" and
// "<digits>[.<digits>]" with at most 9 digits.  Anything else returns false and takes the generic parser.
__device__ __forceinline__ bool fast_id(const uint8_t* s, int p, int end, u32* out)
{
    while (p < end && s[p] != '_') p++;
    p++;
    const int nd = end - p;
    if (nd < 1 || nd > 9) return false;
    u32 v = 0;
    for (; p < end; p++) {
        const u32 d = (u32)s[p] - (u32)'0';
        if (d > 9u) return false;
        v = v * 10u + d;
    }
    *out = v;
    return true;
}

__device__ __forceinline__ bool fast_score(const uint8_t* s, int p, int end, double* out)
{
    const int len = end - p;
    if (len < 1 || len > 10) return false;
    u32 m = 0;
    int frac = -1, nd = 0;
    for (; p < end; p++) {
        const u32 c = s[p];
        const u32 d = c - (u32)'0';
        if (d <= 9u) { m = m * 10u + d; nd++; if (frac >= 0) frac++; }
        else if (c == (u32)'.' && frac < 0) frac = 0;
        else return false;
    }
    if (nd < 1 || nd > 9) return false;
    const double v = (double)m;
    *out = frac > 0 ? v / OFG_POW10_DEV[frac] : v;  // exact: m < 2^30, one correctly rounded division
    return true;
}

__global__ void __launch_bounds__(PARSE_THREADS) parse_kernel(ParseArgs a)
{
    __shared__ __align__(16) uint8_t s_text[PARSE_TILE + PARSE_OVERLAP];
    __shared__ uint16_t s_lstart[PARSE_TILE / 2];
    __shared__ u32 s_term[PARSE_WORDS];
    __shared__ u32 s_tab[PARSE_WORDS];
    __shared__ u32 s_scan[PARSE_THREADS / 32 + 1];
    __shared__ u64 s_base;
    __shared__ u32 s_valid;
    __shared__ int s_prev_term;
    const int tid = threadIdx.x;

    int pair = -1;
    for (int64_t tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        // tiles are visited in increasing order: the pair index only moves forward
        if (pair < 0) pair = upper_bound_dev(a.tile_prefix, a.P + 1, tile) - 1;
        while (tile >= a.tile_prefix[pair + 1]) pair++;
        const PairDesc pd = a.pairs[pair];
        const int64_t t0 = (tile - a.tile_prefix[pair]) * (int64_t)PARSE_TILE;
        const uint8_t* gfile = a.text + pd.text_off;
        const int64_t remain = pd.text_len - t0;
        const int nload = (int)(remain < (int64_t)(PARSE_TILE + PARSE_OVERLAP) ? remain : (int64_t)(PARSE_TILE + PARSE_OVERLAP));
        const int body = (int)(remain < (int64_t)PARSE_TILE ? remain : (int64_t)PARSE_TILE);
        {
            const uint4* g16 = reinterpret_cast<const uint4*>(gfile + t0);
            uint4* s16 = reinterpret_cast<uint4*>(s_text);
            for (int i = tid; i < nload / 16; i += PARSE_THREADS) s16[i] = __ldg(g16 + i);
            const uint4 nl = make_uint4(0x0a0a0a0au, 0x0a0a0a0au, 0x0a0a0a0au, 0x0a0a0a0au);
            for (int i = nload / 16 + tid; i < (PARSE_TILE + PARSE_OVERLAP) / 16; i += PARSE_THREADS) s16[i] = nl;
            if (tid == 0) {
                s_prev_term = (t0 == 0) ? 1 : (int)ofg_is_term(gfile[t0 - 1]);
                s_valid = 0;
            }
        }
        __syncthreads();
        // structural bitmaps of the whole staged window (32 bytes per word): terminators and tabs
        u32 starts = 0, quotes = 0;
        for (int ch = tid; ch < PARSE_WORDS; ch += PARSE_THREADS) {
            const u32* w = reinterpret_cast<const u32*>(s_text + ch * 32);
            u32 term = 0, tabs = 0, qu = 0;
            const uint4 lo4 = reinterpret_cast<const uint4*>(w)[0], hi4 = reinterpret_cast<const uint4*>(w)[1];
            const u32 xs[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const u32 x = xs[k];
                term |= hi_to_nibble(bytes_eq_hi(x, 0x0a0a0a0au) | bytes_eq_hi(x, 0x0d0d0d0du)) << (4 * k);
                tabs |= hi_to_nibble(bytes_eq_hi(x, 0x09090909u)) << (4 * k);
                const u32 tq = bytes_eq_hi(x, 0x22222222u);
                if (tq) qu |= hi_to_nibble(tq) << (4 * k);
            }
            s_term[ch] = term;
            s_tab[ch] = tabs;
            if (ch == tid) {  // body chunk of this thread: line starts
                const int off = tid * 32;
                const u32 prev = (tid == 0) ? (u32)s_prev_term : (u32)ofg_is_term(s_text[off - 1]);
                starts = ~term & ((term << 1) | prev);
                quotes = qu;
                const int nvalid = body - off;  // bytes of this chunk inside the body
                if (nvalid <= 0) { starts = 0; quotes = 0; }
                else if (nvalid < 32) { starts &= (1u << nvalid) - 1u; quotes &= (1u << nvalid) - 1u; }
            }
        }
        if (quotes) {
            int b = __ffs(quotes) - 1;
            // report the line holding the quote: approximate by the byte offset of the quote itself
            atomicMin(&a.pair_err[pair], ((u64)(t0 + tid * 32 + b) << 4) | (u64)OFG_LINE_QUOTE);
        }
        u32 nl_total;
        u32 pos = block_excl_scan<PARSE_THREADS>(__popc(starts), s_scan, &nl_total);
        while (starts) {
            int b = __ffs(starts) - 1;
            s_lstart[pos++] = (uint16_t)(tid * 32 + b);
            starts &= starts - 1;
        }
        if (tid == 0) {
            s_base = atomicAdd(a.coo_count, (u64)nl_total);
            atomicAdd(&a.pair_lines[pair], (u64)nl_total);
        }
        __syncthreads();
        const u64 base = s_base;
        const int64_t Gi = pd.Gi, Gj = pd.Gj;
        for (u32 li0 = 0; li0 < nl_total; li0 += PARSE_THREADS) {
            const u32 li = li0 + tid;
            u32 row = OFG_ROW_SENTINEL, col = 0;
            double score = 0.0;
            if (li < nl_total) {
                const int st = s_lstart[li];
                u32 f0 = 0, f1 = 0;
                SmemSrc ss{s_text, PARSE_TILE + PARSE_OVERLAP, false};
                int rc = -1;
                int tab1, tab2, tab11, send;
                if (locate_fields(s_tab, s_term, st, &tab1, &tab2, &tab11, &send)) {
                    // fast path: the bitmaps give the field boundaries, only ids and score are read bytewise
                    if (fast_id(s_text, st, tab1, &f0) && fast_id(s_text, tab1 + 1, tab2, &f1) &&
                        fast_score(s_text, tab11 + 1, send, &score))
                        rc = OFG_PL_OK;
                }
                if (rc != OFG_PL_OK) {  // anything unusual: the generic record parser decides (and names the error)
                    ss.truncated = false;
                    rc = ofg_parse_record(ss, (int64_t)st, &f0, &f1, &score);
                }
                if (ss.truncated) {  // line longer than the staged window: re-parse from global memory
                    GlobalSrc gs{gfile, pd.text_len};
                    rc = ofg_parse_record(gs, t0 + st, &f0, &f1, &score);
                }
                int kind = rc;
                if (rc == OFG_PL_OK) {
                    const u32 q = pd.swap ? f1 : f0, s = pd.swap ? f0 : f1;
                    if (pd.same && q == s) {
                        // self hit: skipped before any range check (:83-84)
                    } else if ((int64_t)q >= Gi) {
                        kind = OFG_LINE_RANGE_Q;
                    } else if ((int64_t)s >= Gj) {
                        kind = OFG_LINE_RANGE_S;
                    } else if (score > 0.0) {  // stored only if score > current (initially 0) (:87)
                        row = (u32)(pd.row_base + q);
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
            // row histogram, warp aggregated
            const u32 peers = __match_any_sync(OFG_FULL, row);
            if (row != OFG_ROW_SENTINEL && (int)lane_id() == __ffs(peers) - 1) atomicAdd(&a.rowcount[row], (u32)__popc(peers));
            const u32 vb = __ballot_sync(OFG_FULL, row != OFG_ROW_SENTINEL);
            if (lane_id() == 0 && vb) atomicAdd(&s_valid, (u32)__popc(vb));
        }
        __syncthreads();
        if (tid == 0 && s_valid) atomicAdd(&a.pair_valid[pair], (u64)s_valid);
        __syncthreads();
    }
}
