// Stage RAW, part 1: device-side parse of decompressed Blast{p}_{j}.txt into COO records
// (global row id, hit id, bit score).  Replaces the csv.reader loop of
// blast_file_processor.py:66-88 (GetBLAST6Scores).
//
// Layout: all files of the context live back to back in one text arena, each padded with '\n' to a
// multiple of 16 bytes (blank lines are skipped by the reference, :68-69, so padding is neutral).
//
// Streaming warps: the unit of work is the set of lines that START inside one 64 KB segment of one file, and
// one WARP owns it from the first byte to the last - there is no CTA-level barrier anywhere.  The warp keeps
// the last 4 KB of text in a private shared-memory ring that it refills 1 KB at a time with cp.async, one step
// ahead of use.  Per step every lane builds the structural bitmaps (line terminators, tabs) of one 32-byte
// chunk with SIMD-in-word compares, the line starts are ranked with a warp scan and appended to a queue, and
// whenever 32 complete lines are queued they are parsed in lock step, one line per lane (full lanes except for
// the final flush of a segment).  Lines are never re-read: the ring is contiguous across steps, so a line
// that crosses a step boundary is simply parsed once its terminator has arrived.  Records go to slots reserved
// 256 at a time per warp (one atomicAdd per 8 rounds); lines that are dropped (self hits, score <= 0) and the
// unused tail of the last reservation carry a sentinel row id.  The row histogram (hits per query gene) is
// accumulated over runs of equal rows (a query's hits are consecutive lines).
#pragma once
#include "ofg_common.cuh"
#include "ofg_parse_line.h"

constexpr int PS_RING = 4096;            // bytes of text a warp keeps staged (power of two)
constexpr int PS_STEP = 1024;            // bytes per refill: 32 chunks of 32 bytes, one per lane
constexpr int PS_SEG = 65536;            // lines starting in a 64 KB segment of one file are one warp's unit of work
constexpr int PS_WARPS = 8;
constexpr int PARSE_THREADS = PS_WARPS * 32;
constexpr int PS_QCAP = 36 + PS_STEP / 2;  // pending (<= 33: 32 complete lines + the open one) + the most line starts one step can add
constexpr int PS_RESERVE = 256;          // record slots reserved per atomicAdd
constexpr u32 PS_RMASK = PS_RING - 1;
constexpr u32 PS_WMASK = PS_RING / 32 - 1;
static_assert(PS_SEG <= 65536, "queued line starts are 16-bit stream offsets");

struct ParseArgs {
    const uint8_t* text;
    const PairDesc* pairs;
    int P;
    const int64_t* seg_prefix;   // [P+1] segments before each pair
    int64_t n_segs;
    u32* coo_row;
    u32* coo_col;
    double* coo_val;
    u64 cap;
    u64* coo_count;   // record slots reserved so far (may exceed cap: host checks)
    u32* rowcount;    // [total rows] kept records per (pair, query)
    u64* pair_lines;  // [P]
    u64* pair_valid;  // [P]
    u64* pair_err;    // [P] min over offending lines of (byte offset << 4 | kind), ~0 if none
    u32 one;          // 1 (opaque to the compiler: see chunk_masks_fast)
    int keep_self;    // qExcludeSelfHits = False (DendroBLAST's raw matrices, orthologues.py:272-274)
};

// One warp's staging area.  Positions are "stream offsets": bytes from the start of the segment; the ring index
// is the offset modulo PS_RING, the bitmap word index is (offset / 32) modulo (PS_RING / 32).
struct __align__(16) PsWarp {
    uint8_t text[PS_RING];
    u32 term[PS_RING / 32];   // bit i of word w: byte 32 w + i is '\n' or '\r'
    u32 tab[PS_RING / 32];
    uint16_t q[PS_QCAP];      // stream offsets of the line starts not parsed yet, ascending
    u64 bar;                  // mbarrier of the bulk refills (byte count of one step)
};

struct RingSrc {  // generic parser over the ring
    const uint8_t* s;
    int64_t limit;            // stream offset one past the staged text
    __device__ __forceinline__ int get(int64_t pos) { return pos >= limit ? '\n' : s[(u32)pos & PS_RMASK]; }
};
struct GlobalSrc {
    const uint8_t* g;
    int64_t limit;
    __device__ __forceinline__ int get(int64_t pos) { return pos >= limit ? '\n' : g[pos]; }
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// 1-D bulk copy (TMA engine) global -> shared with an mbarrier that counts the bytes: one lane arms the barrier and issues the
// copy of a whole 1 KB step, every lane of the warp then waits on the barrier's phase.  No per-lane addresses, no LDGSTS.
__device__ __forceinline__ void mbar_init(u64* bar, u32 count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void bulk_load_step(void* smem_dst, const void* gmem_src, u32 bytes, u64* bar)
{
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(b)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "OFG_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra OFG_MBAR_WAIT;\n"
        "}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
        "r"(parity)
        : "memory");
}

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
__device__ __forceinline__ bool chunk_masks_fast(const u32 (&xs)[8], u32 one, u32* term, u32* tabs)
{
    const u32 H = 0x80808080u, A = 0x7f7f7f7fu;
    u32 good = H, tm = 0, tb = 0;
#pragma unroll
    for (int b = 0; b < 8; b++) {
        const u32 x = xs[b];
        // (y * one + A): `one` is a kernel parameter equal to 1, which keeps the add a multiply-add on the FMA pipe
        const u32 n_nl = umad32(x ^ 0x0a0a0a0au, one, A), n_cr = umad32(x ^ 0x0d0d0d0du, one, A);
        const u32 n_tb = umad32(x ^ 0x09090909u, one, A), n_qu = umad32(x ^ 0x22222222u, one, A);
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

// Offset of the first line terminator at or after stream offset st, searched in a 128-byte window (128 if none).  Only
// the last line of a segment needs it: every other line ends right before the next queued line start.
__device__ __forceinline__ u32 find_line_end(const u32* m, u32 st)
{
    const u32 w0 = st >> 5;
    const int sh = (int)(st & 31u);
    const u32 a0 = m[w0 & PS_WMASK], a1 = m[(w0 + 1) & PS_WMASK], a2 = m[(w0 + 2) & PS_WMASK], a3 = m[(w0 + 3) & PS_WMASK],
              a4 = m[(w0 + 4) & PS_WMASK];
    const u32 r0 = __funnelshift_r(a0, a1, sh), r1 = __funnelshift_r(a1, a2, sh);
    const u32 r2 = __funnelshift_r(a2, a3, sh), r3 = __funnelshift_r(a3, a4, sh);
    if (r0) return (u32)__ffs(r0) - 1u;
    if (r1) return 32u + (u32)__ffs(r1) - 1u;
    if (r2) return 64u + (u32)__ffs(r2) - 1u;
    if (r3) return 96u + (u32)__ffs(r3) - 1u;
    return 128u;
}

// Field boundaries of a record with EXACTLY 12 fields, relative to the record start `st`.  The caller knows where the
// NEXT line starts (lines are queued in order), so the line end is the byte before it - provided the two lines are
// separated by exactly one terminator (no "\r\n", no blank lines; otherwise the generic parser takes the record).
// The tabs of a line of up to 96 bytes are all inside two windows of the tab bitmap: the 64 bits from the record start
// (which also give the first two tabs) and the 32 bits before the line end (which also give the last tab).  Anything
// else is left to the generic parser.
__device__ __forceinline__ bool locate_fields_conv(const PsWarp& sw, u32 st, u32 nx, bool act, int* tab1, int* tab2, int* tab11, int* line_end)
{
    const u32 end = nx - 1u;                       // stream offset of the terminator
    const u32 le = end - st;
    // exactly one terminator between the two lines: the byte before it belongs to this line
    const u32 pb = nx - 2u;
    const bool single = ((sw.term[(pb >> 5) & PS_WMASK] >> (pb & 31u)) & 1u) == 0u;
    bool ok = act && single && le <= 96u;
    // 64-bit window starting at st, restricted to the line
    const u32 ws = (st >> 5) & PS_WMASK;
    const int sh = (int)(st & 31u);
    const u32 a0 = sw.tab[ws], a1 = sw.tab[(ws + 1) & PS_WMASK], a2 = sw.tab[(ws + 2) & PS_WMASK];
    u32 lo = __funnelshift_r(a0, a1, sh), hi = __funnelshift_r(a1, a2, sh);
    lo &= le >= 32u ? 0xffffffffu : ((1u << le) - 1u);
    hi &= le >= 64u ? 0xffffffffu : (le > 32u ? ((1u << (le - 32u)) - 1u) : 0u);
    // 32 bits before the line end: bit i = byte end - 32 + i; the part beyond the first 64 bytes of the line is counted from here
    const u32 wb = end - 32u;
    const u32 b0 = sw.tab[(wb >> 5) & PS_WMASK], b1 = sw.tab[((wb >> 5) + 1u) & PS_WMASK];
    const u32 last32 = __funnelshift_r(b0, b1, (int)(wb & 31u));
    const u32 tail = le > 64u ? (last32 & (0xffffffffu << (96u - le))) : 0u;
    ok = ok && (__popc(lo) + __popc(hi) + __popc(tail)) == 11;
    const int f1 = lo ? __ffs(lo) - 1 : 32 + __ffs(hi) - 1;
    const u32 lo2 = lo & (lo - 1u), hi2 = lo ? hi : (hi & (hi - 1u));
    const int f2 = lo2 ? __ffs(lo2) - 1 : 32 + __ffs(hi2) - 1;
    ok = ok && (lo2 | hi2) != 0u;                   // the two id fields end inside the first 64 bytes
    // last tab: highest tab bit among the 32 bytes before the line end (bytes before st may be another line's: rejected by t11 > f2)
    const int t11 = (int)le - 1 - __clz((int)last32);
    ok = ok && last32 != 0u && t11 > f2;
    *tab1 = f1;
    *tab2 = f2;
    *tab11 = t11;
    *line_end = (int)le;
    return ok;
}

// 8 bytes [end - 8, end) of the ring as a little-endian u64 (byte end-1 is the most significant); end is a
// stream offset.  Bytes before the line's first byte may be stale: callers mask them by the field length.
__device__ __forceinline__ u64 load8_ending_at(const uint8_t* s_text, u32 end)
{
    const u32 a = end - 8u;
    const u32* wp = reinterpret_cast<const u32*>(s_text);
    const u32 wi = a >> 2;
    const u32 a0 = wp[wi & (PS_RMASK >> 2)], a1 = wp[(wi + 1) & (PS_RMASK >> 2)], a2 = wp[(wi + 2) & (PS_RMASK >> 2)];
    const int sh = (int)(a & 3u) * 8;
    return (u64)__funnelshift_r(a0, a1, sh) | ((u64)__funnelshift_r(a1, a2, sh) << 32);
}

// 0x80 in every byte of y that is zero (exact)
__device__ __forceinline__ u64 zero_bytes64(u64 y)
{
    const u64 m = 0x7f7f7f7f7f7f7f7fULL;
    return ~(((y & m) + m) | y | m);
}

// X holds 8 ASCII characters, first character in the lowest byte.  Returns false unless all are digits;
// *out = their decimal value (SWAR on the two 32-bit halves, two multiplies each).
__device__ __forceinline__ u32 swar_digits4(u32 w)
{
    u32 v = w & 0x0f0f0f0fu;
    v = ((v * 2561u) >> 8) & 0x00ff00ffu;   // (10 a + b) | (10 c + d) << 16
    return (v * 6553601u) >> 16;            // 100 (10 a + b) + (10 c + d)
}
__device__ __forceinline__ bool swar_digits8(u64 X, u32* out)
{
    const u32 lo = (u32)X, hi = (u32)(X >> 32);
    const bool ok = ((((lo + 0x46464646u) | (lo - 0x30303030u)) | ((hi + 0x46464646u) | (hi - 0x30303030u))) & 0x80808080u) == 0;
    *out = swar_digits4(lo) * 10000u + swar_digits4(hi);
    return ok;
}

// "<anything>_<1..7 digits>" of length 2..8 with exactly one underscore, ending at byte `end` (the tab).
// The digits are the bytes above the underscore: everything at or below it is replaced by '0' with one mask, no shifts.
__device__ __forceinline__ bool swar_id(const uint8_t* s_text, u32 end, int len, bool act, u32* out)
{
    act = act && len >= 2 && len <= 8;
    const int l = act ? len : 8;
    const u32 e = act ? end : 8u;
    const u64 w = load8_ending_at(s_text, e);
    const u64 vm = ~0ULL << (8 * (8 - l));                       // bytes of the field
    const u64 us = zero_bytes64(w ^ 0x5f5f5f5f5f5f5f5fULL) & vm; // 0x80 in the underscore bytes
    const u64 um1 = us - 1ULL;
    const bool one = us != 0 && (us & um1) == 0;                 // exactly one underscore
    const u64 M = ~(us | um1);                                    // the bytes above it (the flag sits in the top bit of its byte)
    const bool shape = one && M != 0;                             // at least one digit
    const u64 X = (w & M) | (0x3030303030303030ULL & ~M);
    u32 v;
    const bool dig = swar_digits8(X, &v);
    *out = v;
    return act && shape && dig;
}
// The same for fields of 9..16 bytes ("123_45678": more than 99 species or more than 99 999 sequences per species): the
// underscore must lie in the last 8 bytes, the bytes in front of that window are only checked for a second underscore.
// Called from a warp-uniform branch that is not taken while every id of a round fits 8 bytes.
__device__ __forceinline__ u64 swar_id_long(const uint8_t* s_text, u32 end, int len, bool act)   // (ok << 32) | value
{
    act = act && len > 8 && len <= 16;
    const u32 e = act ? end : 16u;
    const int l = act ? len : 16;
    const u64 w = load8_ending_at(s_text, e), w2 = load8_ending_at(s_text, e - 8u);
    const u64 us = zero_bytes64(w ^ 0x5f5f5f5f5f5f5f5fULL);
    const u64 um1 = us - 1ULL;
    const bool one = us != 0 && (us & um1) == 0;
    const u64 M = ~(us | um1);
    const bool shape = one && M != 0;
    const u64 X = (w & M) | (0x3030303030303030ULL & ~M);
    u32 v;
    const bool dig = swar_digits8(X, &v);
    const u64 vm2 = ~0ULL << (8 * (16 - l));                     // the len - 8 bytes in front of the window
    const bool head_ok = (zero_bytes64(w2 ^ 0x5f5f5f5f5f5f5f5fULL) & vm2) == 0;
    return ((u64)(act && shape && dig && head_ok) << 32) | (u64)v;
}

// "<digits>[.<digits>]" of length 1..8 ending at byte `end`, at most 8 digits: exact fp64 value.
__device__ __forceinline__ bool swar_score(const uint8_t* s_text, u32 end, int len, bool act, double* out)
{
    act = act && len >= 1 && len <= 8;
    const int l = act ? len : 8;
    const u32 e = act ? end : 8u;
    const u64 w = load8_ending_at(s_text, e);
    const u64 vm = ~0ULL << (8 * (8 - l));
    const u64 dots = zero_bytes64(w ^ 0x2e2e2e2e2e2e2e2eULL) & vm;
    const bool has_dot = dots != 0;
    const u64 dm1 = dots - 1ULL;
    const bool one = (dots & dm1) == 0;                          // zero or one dot
    // remove the dot: everything below it moves up one byte
    const u64 above = ~(dots | dm1);                              // bytes above the dot
    const u64 below = (dots >> 7) - 1ULL;                         // bytes below the dot
    const u64 Y = has_dot ? ((w & above) | ((w & below) << 8)) : w;
    const u64 dmask = has_dot ? (vm << 8) : vm;                   // bytes that hold digits now
    const int frac = has_dot ? (__popcll((long long)above) >> 3) : 0;
    const int ndig = has_dot ? l - 1 : l;
    const bool shape = one && ndig >= 1;
    const u64 X = (Y & dmask) | (0x3030303030303030ULL & ~dmask);
    u32 m;
    const bool dig = swar_digits8(X, &m);
    const double v = (double)m;
    // m / 10^frac, correctly rounded, without the division routine: for m < 10^8 and frac <= 8 one Newton step
    // on the rounded product is exact (exhaustive check: tools/micro/div_pow10_check.c)
    const double p = OFG_POW10_DEV[frac], r = OFG_RCP10_DEV[frac];
    const double q0 = __dmul_rn(v, r);
    *out = __fma_rn(__fma_rn(-q0, p, v), r, q0);
    return act && shape && dig;
}


// Per-file constants a warp keeps while it works on a segment
struct PsFile {
    const uint8_t* gfile;
    int64_t text_len;   // padded file length
    int64_t seg_a;      // file offset of the segment's first byte
    u32 row_base, Gi, Gj;
    bool same, swap;
    int pair;
};

struct PsOut {  // record slots of this warp (warp-uniform)
    u64 base;
    u32 left;
};

// Parses up to 32 queued lines in lock step: lane l takes queue entry first + l (n entries are live).
// open_last: the newest entry has not been terminated inside the ring (it is being evicted): parse it from
// global memory.  loaded_end: stream offset one past the staged text.
__device__ __forceinline__ void ps_round(const ParseArgs& a, PsWarp& sw, const PsFile& f, PsOut& out, int first, int n, int q_total, bool open_last,
                                         u32 loaded_end, u32& n_valid)
{
    const int lane = (int)lane_id();
    if (out.left == 0) {  // warp-uniform
        u64 b = 0;
        if (lane == 0) b = atomicAdd(reinterpret_cast<unsigned long long*>(a.coo_count), (unsigned long long)PS_RESERVE);
        out.base = __shfl_sync(OFG_FULL, b, 0);
        out.left = PS_RESERVE;
    }
    const bool act = lane < n;
    const u32 st = act ? (u32)sw.q[first + lane] : 0u;
    const bool from_global = open_last && lane == n - 1;
    // start of the next queued line (lines are queued in order): gives this line's end without a search
    bool has_next = act && first + lane + 1 < q_total;
    u32 nx = has_next ? (u32)sw.q[first + lane + 1] : st + 2u;
    if (__any_sync(OFG_FULL, act && !has_next && !(open_last && lane == n - 1))) {
        // the last queued line (end of a segment): its terminator is staged, find it in the bitmap
        if (act && !has_next && !(open_last && lane == n - 1)) {
            const u32 le = find_line_end(sw.term, st);
            if (le < 128u) { nx = st + le + 1u; has_next = true; }
        }
    }
    u32 row = OFG_ROW_SENTINEL, col = 0;
    double score = 0.0;
    u32 f0 = 0, f1 = 0;
    int tab1, tab2, tab11, lend;
    bool fast = locate_fields_conv(sw, st, nx, has_next && !from_global, &tab1, &tab2, &tab11, &lend);
    bool ok0 = swar_id(sw.text, st + (u32)tab1, tab1, fast, &f0);
    bool ok1 = swar_id(sw.text, st + (u32)tab2, tab2 - tab1 - 1, fast, &f1);
    if (__any_sync(OFG_FULL, fast && (tab1 > 8 || tab2 - tab1 - 1 > 8))) {   // ids of 9..16 characters in this round
        if (!ok0) { const u64 r = swar_id_long(sw.text, st + (u32)tab1, tab1, fast); if (r >> 32) { ok0 = true; f0 = (u32)r; } }
        if (!ok1) { const u64 r = swar_id_long(sw.text, st + (u32)tab2, tab2 - tab1 - 1, fast); if (r >> 32) { ok1 = true; f1 = (u32)r; } }
    }
    const bool ok2 = swar_score(sw.text, st + (u32)lend, lend - tab11 - 1, fast, &score);
    fast = fast && ok0 && ok1 && ok2;
    if (act) {
        int rc = OFG_PL_OK;
        if (!fast) {  // anything unusual: the generic record parser decides (and names the error)
            if (from_global) {
                GlobalSrc gs{f.gfile, f.text_len};
                rc = ofg_parse_record(gs, f.seg_a + (int64_t)st, &f0, &f1, &score);
            } else {
                RingSrc rs{sw.text, (int64_t)loaded_end};
                rc = ofg_parse_record(rs, (int64_t)st, &f0, &f1, &score);
            }
        }
        int kind = rc;
        if (rc == OFG_PL_OK) {
            const u32 q = f.swap ? f1 : f0, s = f.swap ? f0 : f1;
            if (f.same && q == s) {
                // self hit: skipped before any range check (:83-84)
            } else if (q >= f.Gi) {
                kind = OFG_LINE_RANGE_Q;
            } else if (s >= f.Gj) {
                kind = OFG_LINE_RANGE_S;
            } else if (score > 0.0) {  // stored only if score > current (initially 0) (:87)
                row = f.row_base + q;
                col = s;
            }
        }
        if (kind != OFG_LINE_OK) atomicMin(reinterpret_cast<unsigned long long*>(&a.pair_err[f.pair]),
                                           (unsigned long long)(((u64)(f.seg_a + (int64_t)st) << 4) | (u64)kind));
    }
    {
        const u64 idx = out.base + (u64)(PS_RESERVE - out.left) + (u64)lane;
        if (idx < a.cap) {
            a.coo_row[idx] = row;
            if (act) {
                a.coo_col[idx] = col;
                a.coo_val[idx] = score;
            }
        }
        out.left -= 32;
    }
    // row histogram, aggregated over runs of equal rows (a query's hits are consecutive lines); MATCH.ANY
    // would cost ~64 SM cycles per warp instruction on B200 (tools/micro/match_rate.cu)
    const u32 prev_row = __shfl_up_sync(OFG_FULL, row, 1);
    const u32 heads = __ballot_sync(OFG_FULL, lane == 0 || row != prev_row);
    if (row != OFG_ROW_SENTINEL && ((heads >> lane) & 1u)) {
        const u32 later = heads & ~lanemask_lt() & ~(1u << lane);  // heads after this lane
        const int run_end = later ? __ffs(later) - 1 : 32;
        atomicAdd(&a.rowcount[row], (u32)(run_end - lane));
    }
    n_valid += (u32)__popc(__ballot_sync(OFG_FULL, row != OFG_ROW_SENTINEL));
}

__global__ void __launch_bounds__(PARSE_THREADS, 4) parse_kernel(ParseArgs a)
{
    extern __shared__ __align__(16) unsigned char ps_smem_raw[];
    PsWarp& sw = reinterpret_cast<PsWarp*>(ps_smem_raw)[threadIdx.x >> 5];
    const int lane = (int)lane_id();
    const int64_t n_warps = (int64_t)gridDim.x * PS_WARPS;
    PsOut out{0, 0};
    u32 bar_phase = 0;   // parity of the next bulk refill of this warp
    if (lane == 0) { mbar_init(&sw.bar, 1u); mbar_fence_init(); }
    __syncwarp();
    for (int64_t seg = (int64_t)blockIdx.x * PS_WARPS + (threadIdx.x >> 5); seg < a.n_segs; seg += n_warps) {
        PsFile f;
        {
            const int pair = upper_bound_dev(a.seg_prefix, a.P + 1, seg) - 1;
            const PairDesc& d = a.pairs[pair];
            f.pair = pair;
            f.gfile = a.text + d.text_off;
            f.text_len = d.text_len;
            f.seg_a = (seg - a.seg_prefix[pair]) * (int64_t)PS_SEG;
            f.row_base = (u32)d.row_base; f.Gi = (u32)d.Gi; f.Gj = (u32)d.Gj;
            f.same = d.same != 0 && !a.keep_self; f.swap = d.swap != 0;
        }
        const int64_t len_rel64 = f.text_len - f.seg_a;  // bytes of the file from the segment start on (multiple of 16)
        const u32 b_rel = (u32)(len_rel64 < (int64_t)PS_SEG ? len_rel64 : (int64_t)PS_SEG);  // starts below this offset are ours
        // steps of PS_STEP bytes the file still has from here: the first full_steps are whole, then at most one partial
        const int64_t cl = len_rel64 < (int64_t)0x40000000 ? len_rel64 : (int64_t)0x40000000;  // far beyond any step this segment reads
        const u32 full_steps = (u32)cl / PS_STEP, tail_bytes = (u32)cl % PS_STEP;
        const u32 file_steps = full_steps + (tail_bytes ? 1u : 0u);
        u32 prev_term = (f.seg_a == 0) ? 1u : (u32)ofg_is_term(f.gfile[f.seg_a - 1]);
        int qn = 0;                         // queued line starts
        bool tail_open = false;             // the newest queued line has not seen its terminator yet
        bool foreign_open = prev_term == 0; // the line running into the staged text is not in the queue
        u32 n_lines = 0, n_valid = 0;
        bool bulk_pending = false;          // a bulk refill is in flight (its barrier phase has not been consumed)
        const uint8_t* glane = f.gfile + f.seg_a + lane * 16;  // this lane's first 16-byte piece of step 0
        auto issue_step = [&](u32 k) {  // whole steps: one bulk copy per warp; the last, partial step: 16-byte pieces per lane, '\n' past the end of the file
            uint8_t* dst = sw.text + ((k * PS_STEP) & PS_RMASK) + lane * 16;
            const uint8_t* src = glane + (size_t)k * PS_STEP;
            if (k < full_steps) {
                // the ring slot was read by the whole warp before (generic proxy): every lane is past those reads here
                __syncwarp();
                if (lane == 0) bulk_load_step(sw.text + ((k * PS_STEP) & PS_RMASK), f.gfile + f.seg_a + (size_t)k * PS_STEP, PS_STEP, &sw.bar);
                bulk_pending = true;
            } else {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    if ((u32)(lane * 16 + 512 * h) < tail_bytes) cp_async16(dst + 512 * h, src + 512 * h);
                    else *reinterpret_cast<uint4*>(dst + 512 * h) = make_uint4(0x0a0a0a0au, 0x0a0a0a0au, 0x0a0a0a0au, 0x0a0a0a0au);
                }
            }
            cp_async_commit();
        };
        issue_step(0);
        for (u32 k = 0;; k++) {
            if (k < full_steps) { mbar_wait(&sw.bar, bar_phase); bar_phase ^= 1u; bulk_pending = false; }   // the bulk copy of step k has landed
            else cp_async_wait_all();
            __syncwarp();
            const u32 step_so = k * PS_STEP;
            const u32 loaded_end = step_so + PS_STEP;   // bytes past the file end are staged as '\n'
            const bool more = k + 1 < file_steps;
            if (more) {
                // the next refill overwrites stream bytes [step_so - 3 KB, step_so - 2 KB): lines starting there go now
                if (qn > 0 && (int)sw.q[0] < (int)step_so - 2 * PS_STEP) {
                    for (int r0 = 0; r0 < qn; r0 += 32) {
                        const int n = qn - r0 < 32 ? qn - r0 : 32;
                        ps_round(a, sw, f, out, r0, n, qn, tail_open && r0 + n == qn, step_so, n_valid);
                        n_lines += (u32)n;
                    }
                    qn = 0;
                    foreign_open = foreign_open || tail_open;
                    tail_open = false;
                    __syncwarp();
                }
                issue_step(k + 1);
            }
            // ---- structural bitmaps of this step: lane l owns the chunk at stream offset step_so + 32 l
            const u32 so = step_so + (u32)lane * 32u;
            u32 term, tabs;
            {
                const uint4* w4 = reinterpret_cast<const uint4*>(sw.text + (so & PS_RMASK));
                const uint4 lo4 = w4[0], hi4 = w4[1];
                const u32 xs[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
                if (!chunk_masks_fast(xs, a.one, &term, &tabs)) {
                    u32 qu;
                    chunk_masks_exact(xs, &term, &tabs, &qu);
                    const u32 nv = k < full_steps ? (u32)PS_STEP : (k == full_steps ? tail_bytes : 0u);
                    const u32 inside = so < step_so + nv ? (step_so + nv - so >= 32u ? 0xffffffffu : (1u << (step_so + nv - so)) - 1u) : 0u;
                    qu &= inside;
                    if (qu)  // csv quoting is not supported: report the byte offset of the quote
                        atomicMin(reinterpret_cast<unsigned long long*>(&a.pair_err[f.pair]),
                                  (unsigned long long)(((u64)(f.seg_a + (int64_t)so + (int64_t)(__ffs(qu) - 1)) << 4) | (u64)OFG_LINE_QUOTE));
                }
                sw.term[(so >> 5) & PS_WMASK] = term;
                sw.tab[(so >> 5) & PS_WMASK] = tabs;
            }
            // ---- line starts: first byte after a terminator that is not itself a terminator
            const u32 up = __shfl_up_sync(OFG_FULL, term, 1);
            const u32 prev = lane == 0 ? prev_term : (up >> 31);
            u32 stb = ~term & ((term << 1) | prev);
            const u32 any_term = __ballot_sync(OFG_FULL, term != 0u);
            prev_term = (any_term >> 31) ? (__shfl_sync(OFG_FULL, term, 31) >> 31) : 0u;
            if (loaded_end <= b_rel) {
                // common case: every start of this step is ours
                const u32 has_start = __ballot_sync(OFG_FULL, stb != 0u);
                const u32 multi = __ballot_sync(OFG_FULL, (stb & (stb - 1u)) != 0u);
                if (multi == 0u) {  // at most one start per 32-byte chunk (lines longer than 32 bytes)
                    if (stb) sw.q[qn + __popc(has_start & lanemask_lt())] = (uint16_t)(so + (u32)__ffs(stb) - 1u);
                    qn += __popc(has_start);
                } else {
                    const u32 cnt = (u32)__popc(stb);
                    const u32 inc = warp_incl_scan(cnt);
                    int pos = qn + (int)(inc - cnt);
                    u32 t = stb;
                    while (t) {
                        sw.q[pos++] = (uint16_t)(so + (u32)__ffs(t) - 1u);
                        t &= t - 1;
                    }
                    qn += (int)__shfl_sync(OFG_FULL, inc, 31);
                }
                if (any_term) foreign_open = false;
                tail_open = prev_term == 0u && !foreign_open;  // the line open at the end of the step is the newest queued one
            } else {
                // last steps of the segment: starts at or beyond b_rel belong to the next segment
                if (so >= b_rel) stb = 0;
                else if (b_rel - so < 32u) stb &= (1u << (b_rel - so)) - 1u;
                const u32 has_start = __ballot_sync(OFG_FULL, stb != 0u);
                if (has_start) {
                    const u32 cnt = (u32)__popc(stb);
                    const u32 inc = warp_incl_scan(cnt);
                    int pos = qn + (int)(inc - cnt);
                    u32 t = stb;
                    while (t) {
                        sw.q[pos++] = (uint16_t)(so + (u32)__ffs(t) - 1u);
                        t &= t - 1;
                    }
                    qn += (int)__shfl_sync(OFG_FULL, inc, 31);
                    // is the newest queued start followed by a terminator?
                    const int last_lane = 31 - __clz(has_start);
                    const int hb = 31 - __clz(stb | 1u);  // highest start bit of this lane (used on last_lane)
                    const bool closed_here = (hb < 31 && (term >> (hb + 1)) != 0u) || (last_lane < 31 && (any_term >> (last_lane + 1)) != 0u);
                    tail_open = !__shfl_sync(OFG_FULL, (int)closed_here, last_lane);
                } else if (any_term) {
                    tail_open = false;
                }
                if (tail_open && !more && tail_bytes == 0u) {
                    // the file ends exactly at this step without a final newline: plant a terminator after it
                    if (lane == 0) { sw.term[(loaded_end >> 5) & PS_WMASK] = 0xffffffffu; sw.tab[(loaded_end >> 5) & PS_WMASK] = 0u; }
                    tail_open = false;
                }
            }
            __syncwarp();
            // ---- parse complete lines, 32 at a time.  A round needs the start of the line after its last one (that is where
            // the last line ends), so 33 queued starts make a round; the final flush of a segment hands its very last line to
            // the generic parser.
            const bool seg_done = loaded_end >= b_rel && !tail_open;
            const int ready = tail_open ? qn - 1 : qn;
            int done = 0;
            while (qn - done >= 33 && ready - done >= 32 || (seg_done && done < qn)) {
                const int n = qn - done < 32 ? qn - done : 32;
                ps_round(a, sw, f, out, done, n, qn, false, loaded_end, n_valid);
                done += n;
                n_lines += (u32)n;
            }
            if (seg_done) {
                // a refill that was issued ahead and is not needed any more must still land before the barrier is armed again
                if (bulk_pending) { mbar_wait(&sw.bar, bar_phase); bar_phase ^= 1u; bulk_pending = false; }
                break;
            }
            if (done > 0) {  // keep the leftovers (at most 32 complete lines and the open one) at the front of the queue
                const int left = qn - done;
                uint16_t v = 0, v2 = 0;
                if (lane < left) v = sw.q[done + lane];
                if (lane == 0 && left > 32) v2 = sw.q[done + 32];
                __syncwarp();
                if (lane < left) sw.q[lane] = v;
                if (lane == 0 && left > 32) sw.q[32] = v2;
                qn = left;
            }
            __syncwarp();
        }
        if (lane == 0) {
            if (n_lines) atomicAdd(reinterpret_cast<unsigned long long*>(&a.pair_lines[f.pair]), (unsigned long long)n_lines);
            if (n_valid) atomicAdd(reinterpret_cast<unsigned long long*>(&a.pair_valid[f.pair]), (unsigned long long)n_valid);
        }
        __syncwarp();
    }
    // unused tail of the last reservation: sentinel rows
    for (u32 i = (u32)lane; i < out.left; i += 32) {
        const u64 idx = out.base + (u64)(PS_RESERVE - out.left) + (u64)i;
        if (idx < a.cap) a.coo_row[idx] = OFG_ROW_SENTINEL;
    }
}
